"""Import shim that makes the read-only reference at /root/reference importable
with the installed torch 2.x / transformers 5.x (it pins transformers 3.0.2).

Used ONLY by tests/golden/make_golden.py in the build container to generate the
committed fixtures.  Nothing under tests/ (at run time), bench.py or the
product imports this file: /root/reference does not exist on the GPU box.

What it patches:
  * old flat module paths `transformers.modeling_bert`, `transformers.tokenization_bert`
    (reference: Tiny-NewsRec/tnlrv3/modeling.py:12, tokenization_tnlrv3.py:12)
  * `cached_path`, `TF2_WEIGHTS_NAME`, `TF_WEIGHTS_NAME`, `WEIGHTS_NAME`
    (tnlrv3/convert_state_dict.py:4, modeling.py:15)
  * `TuringNLRv3PreTrainedModel.from_pretrained` -> construct from config with
    the reference's own `_init_weights` (no UniLM .bin exists offline).
"""
import os
import sys
import types

REF_ROOT = os.environ.get("TNR_REFERENCE_ROOT", "/root/reference")


def install(app="Tiny-NewsRec"):
    import torch  # noqa: F401
    import transformers
    import transformers.models.bert.modeling_bert as hf_bert
    import transformers.models.bert.tokenization_bert as hf_tok
    import transformers.modeling_utils as hf_mu
    import transformers.file_utils as hf_fu

    sys.modules.setdefault("transformers.modeling_bert", hf_bert)
    sys.modules.setdefault("transformers.tokenization_bert", hf_tok)
    if not hasattr(hf_bert, "BertPreTrainedModel"):
        raise RuntimeError("unexpected transformers layout")
    for name, val in (("cached_path", lambda *a, **k: None),
                      ("TF2_WEIGHTS_NAME", "tf_model.h5"),
                      ("TF_WEIGHTS_NAME", "model.ckpt"),
                      ("WEIGHTS_NAME", "pytorch_model.bin")):
        for mod in (hf_mu, hf_fu):
            if not hasattr(mod, name):
                setattr(mod, name, val)
    if not hasattr(hf_tok, "whitespace_tokenize"):
        hf_tok.whitespace_tokenize = lambda s: s.strip().split()

    app_dir = os.path.join(REF_ROOT, app)
    if app_dir not in sys.path:
        sys.path.insert(0, app_dir)

    import tnlrv3.modeling as ref_modeling

    base = ref_modeling.TuringNLRv3PreTrainedModel

    def _init_weights_all(self):
        self.apply(self._init_weights)

    base.init_weights = _init_weights_all

    @classmethod
    def _from_pretrained(cls, name_or_path, *a, config=None, **kw):
        return cls(config)

    base.from_pretrained = _from_pretrained
    return ref_modeling


def make_args(**over):
    """argparse-like namespace with the demo.sh values (Tiny-NewsRec/demo.sh:8-33)."""
    d = dict(pooling="att", model_type="tnlrv3",
             config_name=os.path.join(REF_ROOT, "Tiny-NewsRec/tnlrv3/config/tnlrv3-base-uncased-config.json"),
             model_name="unused", num_teacher_layers=12, num_student_layers=4,
             num_hidden_layers=12, news_query_vector_dim=200, user_query_vector_dim=200,
             news_dim=256, model="NAML", num_attention_heads=16, user_log_mask=False,
             user_log_length=50, num_teachers=4, temperature=1.0, coef=0.2)
    d.update(over)
    return types.SimpleNamespace(**d)
