"""On-disk formats of the path: UniLM / TNLRv3 ``.bin`` conversion, the
``epoch-{n}.pt`` training checkpoint and the pickled teacher embedding tables.  Pure host code.

Reference: Tiny-NewsRec/tnlrv3/convert_state_dict.py:39-76 (``load_model``), tnlrv3/modeling.py:88-118
(position-embedding resize), :120-128 (``replace_prefix``), run.py:205-214 (checkpoint dict),
run.py:458-459 (teacher ``.pkl``), run.py:61-88 (loading teachers / first-stage students).
"""
import pickle

import numpy as np
import torch


def convert_unilm_state_dict(state_dict):
    """UniLM checkpoint keys -> the module's keys (convert_state_dict.py:39-71): fused ``qkv_linear.weight``
    split into query / key / value, ``q_bias`` / ``v_bias`` flattened, the key bias created as zeros,
    ``bert.encoder.rel_pos_bias.weight`` moved to ``bert.rel_pos_bias.weight``."""
    out = {}
    for key, value in state_dict.items():
        if key.endswith("attention.self.q_bias"):
            out[key.replace("attention.self.q_bias", "attention.self.query.bias")] = value.view(-1)
        elif key.endswith("attention.self.v_bias"):
            out[key.replace("attention.self.v_bias", "attention.self.value.bias")] = value.view(-1)
            out[key.replace("attention.self.v_bias", "attention.self.key.bias")] = torch.zeros_like(value.view(-1))
        elif key.endswith("attention.self.qkv_linear.weight"):
            rows = value.shape[0]
            if rows % 3:
                raise ValueError(f"{key}: fused QKV weight with {rows} rows is not divisible by 3")
            q, k, v = torch.split(value, rows // 3, dim=0)
            base = key[:-len("qkv_linear.weight")]
            out[base + "query.weight"], out[base + "key.weight"], out[base + "value.weight"] = q, k, v
        elif key == "bert.encoder.rel_pos_bias.weight":
            out["bert.rel_pos_bias.weight"] = value
        else:
            out[key] = value
    return out


def resize_position_embeddings(state_dict, max_position_embeddings, initializer_range=0.02,
                               reuse_position_embedding=None, generator=None):
    """tnlrv3/modeling.py:88-118: grow (new rows ~ N(0, initializer_range); with ``reuse_position_embedding``
    the old table is tiled over the new one, otherwise copied once) or truncate the position table."""
    k = "bert.embeddings.position_embeddings.weight"
    if k not in state_dict:
        return state_dict
    old = state_dict[k]
    n_old = old.shape[0]
    if max_position_embeddings > n_old:
        new = torch.empty(max_position_embeddings, old.shape[1], dtype=torch.float32)
        new.normal_(mean=0.0, std=initializer_range, generator=generator)
        max_range = max_position_embeddings if reuse_position_embedding else n_old
        shift = 0
        while shift < max_range:
            delta = min(n_old, max_range - shift)
            new[shift:shift + delta] = old[:delta]
            shift += delta
        state_dict[k] = new
    elif max_position_embeddings < n_old:
        state_dict[k] = old[:max_position_embeddings].clone().float()
    return state_dict


def strip_prefix(state_dict, replace_prefix):
    """tnlrv3/modeling.py:120-128."""
    if replace_prefix is None:
        return state_dict
    return {(k[len(replace_prefix):] if k.startswith(replace_prefix) else k): v for k, v in state_dict.items()}


def load_unilm_bin(bert_model, path_or_state_dict, max_position_embeddings=None, initializer_range=0.02,
                   reuse_position_embedding=None, replace_prefix=None):
    """What ``TuringNLRv3ForSequenceClassification.from_pretrained(model_name, config=...)`` does for a local
    ``.bin`` (model_bert.py:107-113): convert, resize positions, then a NON-strict load -- layers beyond the
    module's depth are dropped (12-layer file into a 4-layer student), the pooler / classifier the file lacks
    keep their initialisation.  Returns (missing_keys, unexpected_keys)."""
    sd = path_or_state_dict
    if not isinstance(sd, dict):
        sd = torch.load(path_or_state_dict, map_location="cpu")
    sd = convert_unilm_state_dict(sd)
    if max_position_embeddings is None:
        max_position_embeddings = bert_model.bert.embeddings.position_embeddings.weight.shape[0]
    sd = resize_position_embeddings(sd, max_position_embeddings, initializer_range, reuse_position_embedding)
    sd = strip_prefix(sd, replace_prefix)
    res = bert_model.load_state_dict(sd, strict=False)
    return list(res.missing_keys), list(res.unexpected_keys)


def save_checkpoint(path, model, category_dict=None, word_dict=None, subcategory_dict=None):
    """run.py:205-214."""
    torch.save({"model_state_dict": {k: v.detach().cpu().clone() for k, v in model.state_dict().items()},
                "category_dict": category_dict, "word_dict": word_dict, "subcategory_dict": subcategory_dict}, path)


def load_checkpoint(path, model=None, strict=True):
    """run.py:232-240 (test) / :395-399 (get_teacher_emb): returns the checkpoint dict, loading
    ``model_state_dict`` into ``model`` when given."""
    ckpt = torch.load(path, map_location="cpu")
    if model is not None:
        model.load_state_dict(ckpt["model_state_dict"], strict=strict)
    return ckpt


def save_teacher_table(path, table):
    """run.py:458-459: float32 ndarray [N+1, D], pickled."""
    arr = np.asarray(table.detach().cpu().numpy() if isinstance(table, torch.Tensor) else table, dtype=np.float32)
    with open(path, "wb") as f:
        pickle.dump(arr, f)


def load_teacher_table(path):
    """run.py:48-51."""
    with open(path, "rb") as f:
        return np.asarray(pickle.load(f), dtype=np.float32)
