"""Drop-in for ``Tiny-NewsRec/model_bert_2.py`` (teacher model used by get_teacher_emb) and
``PLM-NR/model_bert.py`` (PLM-NR fine-tuning): ``ModelBert(args).forward(history,
history_mask, candidate, label) -> (loss, score)``.

Reference: Tiny-NewsRec/model_bert_2.py:103-138 (NewsEncoder with args.num_hidden_layers),
:141-183 (UserEncoder), :186-213 (ModelBert); PLM-NR/model_bert.py:187-207.
"""
import torch
from torch import nn

from ._lib import TinyRecError
from .model_bert import AttentionPooling, NewsEncoder as _NewsEncoder, TrainState, UserEncoder, _StepFn, _get_state  # noqa: F401


class NewsEncoder(_NewsEncoder):
    """model_bert_2.py:103-138: depth comes from ``args.num_hidden_layers``."""

    def __init__(self, args):
        super().__init__(args, num_layers=args.num_hidden_layers)


class ModelBert(nn.Module):
    def __init__(self, args):
        super().__init__()
        self.args = args
        self.news_encoder = NewsEncoder(args)
        self.user_encoder = UserEncoder(args)
        self.loss_fn = nn.CrossEntropyLoss()

    def train_state(self):
        dev = self.user_encoder.pad_doc.device
        if dev.type != "cuda":
            raise TinyRecError("tinyrec ModelBert needs its parameters on a CUDA device (no CPU path)")
        return _get_state(self, lambda: TrainState(self.news_encoder, self.user_encoder, [], [], dev))

    def forward(self, history, history_mask, candidate, label):
        st = self.train_state()
        want_grad = torch.is_grad_enabled() and st.flat is not None
        # CE only: M = 0 teachers, coef = 1 -> total == target loss (model_bert_2.py:212)
        args = (history, history_mask, candidate, label, [], [], 1.0, 1.0, bool(self.args.user_log_mask), want_grad,
                bool(self.training))
        if want_grad:
            total, _, _, _, score = _StepFn.apply(st.anchor, st, args)
            return total, score
        w = st.step_forward(*args)
        return w["losses"][3].clone(), w["score"].clone()
