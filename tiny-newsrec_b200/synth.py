"""Deterministic random-init weights and synthetic MIND-shaped data.

There is no network for datasets or checkpoints, so benchmarks and parity tests
use (a) weights drawn from the same distributions the reference constructors use
and (b) impressions / news shaped like MIND.  Every tensor is
generated from ``(seed, key)`` with a CPU ``torch.Generator`` so the build
container, the tests and the GPU box all see identical values.

Reference init rules restated here:
  * encoder Linear / Embedding weights ~ N(0, 0.02), biases 0, LayerNorm 1 / 0
    (Tiny-NewsRec/tnlrv3/modeling.py:41-51; applies to pooler, classifier and
    rel_pos_bias too)
  * attention-pooling / dense heads: torch ``nn.Linear`` default
    U(-1/sqrt(fan_in), 1/sqrt(fan_in)) for weight and bias (model_bert.py:12-13,117)
  * ``pad_doc`` ~ U(-1, 1) (model_bert.py:151-153)
  * ``transform_matrix`` xavier-uniform, zero bias (model_bert.py:258-260)
"""
import math
import zlib
from types import SimpleNamespace

import numpy as np
import torch

BERT_BASE = dict(vocab_size=30522, hidden_size=768, num_attention_heads=12,
                 intermediate_size=3072, max_position_embeddings=512, type_vocab_size=2,
                 rel_pos_bins=32, max_rel_pos=128, layer_norm_eps=1e-12, num_labels=2,
                 hidden_dropout_prob=0.1, attention_probs_dropout_prob=0.1)


def _gen(seed, key):
    g = torch.Generator(device="cpu")
    g.manual_seed((seed * 1000003 + zlib.crc32(key.encode())) % (2 ** 63 - 1))
    return g


def _normal(seed, key, shape, std):
    return torch.randn(shape, generator=_gen(seed, key), dtype=torch.float32) * std


def _uniform(seed, key, shape, bound):
    return (torch.rand(shape, generator=_gen(seed, key), dtype=torch.float32) * 2 - 1) * bound


def bert_model_state(pfx, num_layers, seed, cfg=None, noisy=False):
    """State of ``TuringNLRv3ForSequenceClassification`` under prefix ``pfx``
    (keys as enumerated in SURVEY.md section 8b).  ``noisy=True`` also randomises
    biases and LayerNorm parameters so parity tests exercise them."""
    c = dict(BERT_BASE)
    c.update(cfg or {})
    E, F, A = c["hidden_size"], c["intermediate_size"], c["num_attention_heads"]
    sd = {}

    def lin(name, out_f, in_f):
        sd[name + ".weight"] = _normal(seed, name + ".weight", (out_f, in_f), 0.02)
        sd[name + ".bias"] = (_normal(seed, name + ".bias", (out_f,), 0.02) if noisy
                              else torch.zeros(out_f))

    def ln(name):
        sd[name + ".weight"] = (1.0 + _normal(seed, name + ".weight", (E,), 0.1) if noisy
                                else torch.ones(E))
        sd[name + ".bias"] = (_normal(seed, name + ".bias", (E,), 0.1) if noisy
                              else torch.zeros(E))

    e = pfx + "bert.embeddings."
    sd[e + "word_embeddings.weight"] = _normal(seed, e + "word", (c["vocab_size"], E), 0.02)
    sd[e + "position_embeddings.weight"] = _normal(seed, e + "pos", (c["max_position_embeddings"], E), 0.02)
    sd[e + "token_type_embeddings.weight"] = _normal(seed, e + "type", (c["type_vocab_size"], E), 0.02)
    ln(e + "LayerNorm")
    for l in range(num_layers):
        p = f"{pfx}bert.encoder.layer.{l}."
        for n in ("query", "key", "value"):
            lin(p + "attention.self." + n, E, E)
        lin(p + "attention.output.dense", E, E)
        ln(p + "attention.output.LayerNorm")
        lin(p + "intermediate.dense", F, E)
        lin(p + "output.dense", E, F)
        ln(p + "output.LayerNorm")
    lin(pfx + "bert.pooler.dense", E, E)
    sd[pfx + "bert.rel_pos_bias.weight"] = _normal(seed, pfx + "relpos", (A, c["rel_pos_bins"]),
                                                    0.5 if noisy else 0.02)
    lin(pfx + "classifier", c["num_labels"], E)
    return sd


def _default_linear(sd, seed, name, out_f, in_f):
    b = 1.0 / math.sqrt(in_f)
    sd[name + ".weight"] = _uniform(seed, name + ".weight", (out_f, in_f), b)
    sd[name + ".bias"] = _uniform(seed, name + ".bias", (out_f,), b)


def attention_pooling_state(pfx, emb, hidden, seed):
    sd = {}
    _default_linear(sd, seed, pfx + "att_fc1", hidden, emb)
    _default_linear(sd, seed, pfx + "att_fc2", 1, hidden)
    return sd


def user_encoder_state(pfx, news_dim, qdim, seed, model="NAML", n_heads=16):
    """``UserEncoder`` (model_bert.py:140-153).  model='NRMS' adds the multi-head self-attention
    (W_Q / W_K / W_V: xavier-uniform weights :73-76, default nn.Linear biases) and pools over
    n_heads * 16 features; key order = the reference's state_dict order."""
    sd = {pfx + "pad_doc": _uniform(seed, pfx + "pad_doc", (1, news_dim), 1.0)}
    pool_dim = news_dim
    if model == "NRMS":
        pool_dim = n_heads * 16
        for nm in ("W_Q", "W_K", "W_V"):
            k = pfx + "multi_head_self_attn." + nm
            sd[k + ".weight"] = _uniform(seed, k + ".weight", (pool_dim, news_dim), math.sqrt(6.0 / (news_dim + pool_dim)))
            sd[k + ".bias"] = _uniform(seed, k + ".bias", (pool_dim,), 1.0 / math.sqrt(news_dim))
    sd.update(attention_pooling_state(pfx + "attn.", pool_dim, qdim, seed))
    return sd


def model_bert_state(pfx, num_layers, seed, news_dim=256, news_q=200, user_q=200, cfg=None,
                     noisy=False, model="NAML", n_heads=16):
    """``ModelBert`` = news_encoder (bert_model + attn + dense) + user_encoder."""
    E = (cfg or {}).get("hidden_size", BERT_BASE["hidden_size"])
    sd = bert_model_state(pfx + "news_encoder.bert_model.", num_layers, seed, cfg, noisy)
    sd.update(attention_pooling_state(pfx + "news_encoder.attn.", E, news_q, seed))
    _default_linear(sd, seed, pfx + "news_encoder.dense", news_dim, E)
    sd.update(user_encoder_state(pfx + "user_encoder.", news_dim, user_q, seed, model, n_heads))
    return sd


def kd_model_state(num_layers, num_teachers, seed, news_dim=256, news_q=200, user_q=200,
                   cfg=None, noisy=False, model="NAML", n_heads=16):
    """``Model`` (KD wrapper): teachers.{i} user encoders, student ModelBert,
    transform_matrix.{i}.  Key order follows the reference module order."""
    sd = {}
    for i in range(num_teachers):
        sd.update(user_encoder_state(f"teachers.{i}.", news_dim, user_q, seed, model, n_heads))
    sd.update(model_bert_state("student.", num_layers, seed, news_dim, news_q, user_q, cfg, noisy, model, n_heads))
    for i in range(num_teachers):
        b = math.sqrt(6.0 / (news_dim + news_dim))
        sd[f"transform_matrix.{i}.weight"] = _uniform(seed, f"tm{i}.w", (news_dim, news_dim), b)
        sd[f"transform_matrix.{i}.bias"] = (_normal(seed, f"tm{i}.b", (news_dim,), 0.02) if noisy
                                            else torch.zeros(news_dim))
    return sd


# --------------------------------------------------------------------------
# synthetic MIND-shaped data
# --------------------------------------------------------------------------
def news_table(n_news, L=30, seed=1234, vocab=30522, mean_len=14.0, std_len=4.0, min_len=4):
    """int32 [n_news+1, 2L] = token ids | attention mask; row 0 is the all-zero pad
    news (preprocess.py:49-53).  [CLS]=101 ... [SEP]=102, ids uniform in [1000, vocab)."""
    rng = np.random.default_rng(seed)
    N = n_news + 1
    lens = np.clip(np.rint(rng.normal(mean_len, std_len, N)), min_len, L).astype(np.int64)
    ids = rng.integers(1000, vocab, size=(N, L), dtype=np.int64)
    pos = np.arange(L)[None, :]
    ids[:, 0] = 101
    ids[np.arange(N), lens - 1] = 102
    mask = (pos < lens[:, None])
    ids = ids * mask
    out = np.concatenate([ids, mask.astype(np.int64)], axis=1).astype(np.int32)
    out[0] = 0
    return out


def train_impressions(n, n_news, H=50, K=5, seed=1234):
    """Index-mapped training impressions: hist_idx int32 [n,H] (front-padded with 0),
    hist_mask f32 [n,H], cand_idx int32 [n,K], label int64 [n]
    (dataloader.py:73-83,129-137 semantics)."""
    rng = np.random.default_rng(seed + 17)
    hl = np.minimum(H, rng.geometric(1.0 / 25.0, n)).astype(np.int64)
    hist = rng.integers(1, n_news + 1, size=(n, H), dtype=np.int64)
    valid = np.arange(H)[None, :] >= (H - hl[:, None])
    hist = hist * valid
    cand = rng.integers(1, n_news + 1, size=(n, K), dtype=np.int64)
    label = rng.integers(0, K, size=n, dtype=np.int64)
    return (hist.astype(np.int32), valid.astype(np.float32), cand.astype(np.int32), label)


def eval_impressions(n, n_news, H=50, seed=1234, max_c=300):
    """Eval impressions with ragged candidate lists (CSR): hist_idx int32 [n,H],
    hist_mask f32 [n,H], cand_ptr int64 [n+1], cand_idx int32 [nnz], labels int8 [nnz].
    C ~ clip(round(LogNormal(3.2, 0.8)), 2, 300); labels Bernoulli(0.1), non-constant."""
    rng = np.random.default_rng(seed + 29)
    hl = np.minimum(H, rng.geometric(1.0 / 25.0, n)).astype(np.int64)
    hist = rng.integers(1, n_news + 1, size=(n, H), dtype=np.int64)
    valid = np.arange(H)[None, :] >= (H - hl[:, None])
    hist = hist * valid
    C = np.clip(np.rint(rng.lognormal(3.2, 0.8, n)), 2, max_c).astype(np.int64)
    ptr = np.zeros(n + 1, dtype=np.int64)
    np.cumsum(C, out=ptr[1:])
    nnz = int(ptr[-1])
    cand = rng.integers(1, n_news + 1, size=nnz, dtype=np.int64)
    lab = (rng.random(nnz) < 0.1).astype(np.int8)
    # force non-constant labels: first candidate positive if none, second negative if all
    s = np.add.reduceat(lab.astype(np.int64), ptr[:-1])
    lab[ptr[:-1][s == 0]] = 1
    lab[ptr[:-1][s == C] + 1] = 0
    return (hist.astype(np.int32), valid.astype(np.float32), ptr, cand.astype(np.int32), lab)


def teacher_tables(n_news, M=4, D=256, seed=1234):
    rng = np.random.default_rng(seed + 41)
    return [(rng.standard_normal((n_news + 1, D), dtype=np.float32) * 0.1) for _ in range(M)]


def demo_args(**over):
    """Namespace with the demo.sh values (Tiny-NewsRec/demo.sh:8-33); the reference
    modules read these fields from ``args``."""
    d = dict(pooling="att", model_type="tnlrv3", config_name=None, model_name=None,
             num_teacher_layers=12, num_student_layers=4, num_hidden_layers=12,
             news_query_vector_dim=200, user_query_vector_dim=200, news_dim=256, model="NAML",
             num_attention_heads=16, user_log_mask=False, user_log_length=50, num_teachers=4,
             temperature=1.0, coef=0.2, npratio=4, batch_size=32, num_words_title=30,
             bert_trainable_layer=[2, 3], lr=1e-4, enable_hvd=False, enable_gpu=True)
    d.update(over)
    return SimpleNamespace(**d)
