"""First-stage (post-training) knowledge distillation: the model wrappers of the reference's
``Post-train_KD.ipynb`` on the B200 path.

    TitleBodySimModel(args).forward(title [bz, 1+K, 2*Lt], body [bz, 2*Lb]) -> (scores, title_emb, body_emb)   (cell 12)
    DistillModel(args).forward(title, body, labels, teacher_titles, teacher_bodies)
        -> (loss, target_loss, distill_loss, emb_loss, student_score)                                          (cell 14)

The news encoder is the same ``NewsEncoder`` (cell 11 == Tiny-NewsRec/model_bert.py:103-137), run twice per step:
over the 1+K titles (24 tokens) and over the body (up to 512 tokens -> the streamed-KV attention kernels, forward
and backward).  The loss is the finetuning-stage loss of model_bert.py:262-306 with no click history: the titles
play the candidates, the body vector plays the user vector, so ``tnr_kd_loss_fwdbwd`` computes it (H = 0, coef = 1,
temperature 1) together with every gradient; state_dict keys are the notebook's
(``student.news_encoder.*``, ``transform_matrix.{i}.*``).

Two notes on the published notebook: cell 14 multiplies the Python list ``teacher_MSEs`` by a tensor (a TypeError as
written) -- the evident intent, ``torch.stack(teacher_MSEs, dim=-1)``, is what is implemented (and what
``oracle.model.post_train_distill_forward`` restates: parity for this wrapper is pinned on that restatement, not on
a run of the notebook); cell 18 trains ``bert_model`` at lr 1e-6 and everything else at 1e-5 --
``DistillModel.lr_ranges(lr_bert, lr_rest)`` gives the two flat-buffer ranges for ``tinyrec.optim.Adam.set_lr_ranges``.
"""
from types import SimpleNamespace

import torch
from torch import nn

from . import ops
from ._lib import TinyRecError
from .engine import F32, FlatParams
from .model_bert import NewsEncoder as _NewsEncoder, _const_stride, _get_state, _trainable_signature


def _encoder_args(args):
    return SimpleNamespace(pooling="att", model_type="tnlrv3", config_name=getattr(args, "config_name", None),
                           model_name=getattr(args, "model_name", None), num_student_layers=args.num_hidden_layers,
                           num_teacher_layers=args.num_hidden_layers, news_query_vector_dim=args.news_query_vector_dim,
                           news_dim=args.news_dim)


TITLE_SITE_BASE = 1 << 12          # dropout tensor-id offset of the title pass (encoder ids stay below 8 * 26)


class TitleBodySimModel(nn.Module):
    """Post-train_KD.ipynb cell 12 (inference semantics here; training goes through ``DistillModel``)."""

    def __init__(self, args):
        super().__init__()
        self.news_encoder = _NewsEncoder(_encoder_args(args), False)

    def forward(self, title, body):
        bz, candi_num, input_num = title.shape
        body_emb = self.news_encoder(body)
        title_emb = self.news_encoder(title.reshape(-1, input_num)).reshape(bz, candi_num, -1)
        scores = torch.bmm(title_emb, body_emb.unsqueeze(-1)).squeeze(-1)          # [bz, 1+K] fp32: plumbing-sized
        return scores, title_emb, body_emb


class _DistillState:
    """Flat parameter / gradient buffers and workspaces of one DistillModel (cf. model_bert.TrainState)."""

    def __init__(self, news_encoder, transform, device):
        self.ne, self.transform = news_encoder, transform
        self.enc = news_encoder.engine()
        self.enc.check_supported()
        enc_params = self.enc.trainable_order()
        tparams = []
        for lin in transform:
            if lin.weight.requires_grad != lin.bias.requires_grad:
                raise TinyRecError("transform_matrix weight/bias must share requires_grad")
            if lin.weight.requires_grad:
                tparams += [lin.weight, lin.bias]
        if tparams and len(tparams) != 2 * len(transform):
            raise TinyRecError("transform_matrix modules must be trainable or frozen together")
        self.tm_trainable = bool(tparams)
        ordered = enc_params + tparams
        self.sig = tuple(id(p) for p in ordered)
        self.flat = FlatParams(ordered, device) if ordered else None
        self.head_begin = self.flat.off(tparams[0]) if tparams else (self.flat.numel if self.flat else 0)
        self.head_end = self.flat.numel if self.flat else 0
        self.stage = torch.zeros(max(self.head_end - self.head_begin, 1), device=device, dtype=F32)
        news_encoder._flat = self.flat
        self.hw = {}
        self.anchor = torch.zeros(1, device=device, requires_grad=True)
        self.comm_hook = None          # DistributedOptimizer falls back to one all-reduce of the flat gradient in step()

    def valid(self, module):
        want = tuple(id(p) for p in _trainable_signature(module))
        return set(want) == set(self.sig) and (self.flat is None or self.flat.attached())

    def _stage_view(self, p):
        o = self.flat.off(p) - self.head_begin
        return self.stage[o:o + p.numel()]

    def step_forward(self, title, body, labels, t_titles, t_bodies, want_grad, training):
        B, K1, Wt = title.shape
        M = len(t_titles)
        dev = title.device
        flat = self.flat
        if flat is not None:
            if flat.stale():
                flat.refresh_shadow()
            if want_grad:
                flat.reattach_grads()
        D = self.enc.D
        R = B * K1
        key = (B, K1, M)
        w = self.hw.get(key)
        if w is None:
            mk = lambda *s: torch.empty(*s, device=dev, dtype=F32)  # noqa: E731
            w = dict(news=mk(R, D), user=mk(B, D), score=mk(B, K1), losses=torch.zeros(4 + 4 * B + 1, device=dev, dtype=F32),
                     d_news=mk(R, D), d_user=mk(B, D), T=mk(max(M, 1), R + B, D), TP=mk(max(M, 1), R + B, D),
                     G=mk(max(M, 1), R + B, D))
            self.hw[key] = w
        drop = None
        if training:
            drop = self.ne.drop_state(dev)
            drop.advance()
        body_emb = self.enc.forward(body, flat, save=want_grad, out=w["user"], drop=drop, tag="body")
        w["ws_body"] = self.enc.last_ws
        # the title pass draws masks from its own Philox streams (tensor ids shifted): nn.Dropout in the reference is
        # independent between the two news_encoder calls of a step
        title_emb = self.enc.forward(title.reshape(R, Wt), flat, save=want_grad, out=w["news"],
                                     drop=drop.offset(TITLE_SITE_BASE) if drop is not None else None, tag="title")
        w["ws_title"] = self.enc.last_ws
        T, TP, G = w["T"], w["TP"], w["G"]
        for i in range(M):
            T[i, :R].copy_(t_titles[i].reshape(R, D))
            T[i, R:].copy_(t_bodies[i].reshape(B, D))
        if M:
            ws_, bs_ = [lin.weight for lin in self.transform], [lin.bias for lin in self.transform]
            sw, sb = _const_stride(ws_), _const_stride(bs_)
            if sw is not None and sb is not None:
                ops.sgemm_nt(T, ws_[0], bs_[0], TP, R + B, D, D, M, (R + B) * D, sw, sb, (R + B) * D)
            else:
                for i in range(M):
                    ops.sgemm_nt(T[i], ws_[i], bs_[i], TP[i], R + B, D, D, 1, 0, 0, 0, 0)
        # H = 0: the titles are the candidates, the body vector stands where the user vector does; coef = 1, tau = 1
        ops.kd_loss(title_emb, body_emb, labels.contiguous(), T if M else None, TP if M else None, M, B, 0, K1, D, 1.0, 1.0,
                    want_grad, w["score"], w["losses"], w["d_news"], w["d_user"], G if M else None)
        if want_grad:
            self.stage.zero_()
            if self.tm_trainable and M:
                gw = [self._stage_view(lin.weight) for lin in self.transform]
                gb = [self._stage_view(lin.bias) for lin in self.transform]
                sw, sb = _const_stride(gw), _const_stride(gb)
                if sw is not None and sb is not None:
                    ops.sgemm_tn_acc(G, T, gw[0], gb[0], R + B, D, D, M, (R + B) * D, (R + B) * D, sw, sb)
                else:
                    for i in range(M):
                        ops.sgemm_tn_acc(G[i], T[i], gw[i], gb[i], R + B, D, D, 1, 0, 0, 0, 0)
        self.last = w
        return w

    def step_backward(self, g_total):
        w, flat = self.last, self.flat
        w["d_news"].mul_(g_total)
        w["d_user"].mul_(g_total)
        if self.head_end > self.head_begin:
            self.stage.mul_(g_total)
            flat.grad[self.head_begin:self.head_end].add_(self.stage[:self.head_end - self.head_begin])
        self.enc.backward(w["d_news"], flat, ws=w["ws_title"])
        self.enc.backward(w["d_user"], flat, ws=w["ws_body"])


class _DistillFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, anchor, state, args):
        w = state.step_forward(*args)
        ctx.state = state
        ls = w["losses"]
        return ls[3].clone(), ls[2].clone(), ls[0].clone(), ls[1].clone(), w["score"].clone()

    @staticmethod
    def backward(ctx, g_total, g_target, g_distill, g_emb, g_score):
        ctx.state.step_backward(g_total)
        return None, None, None


class DistillModel(nn.Module):
    """Post-train_KD.ipynb cell 14.  Only ``loss`` carries gradient (that is all cell 19 back-propagates)."""

    def __init__(self, args):
        super().__init__()
        self.args = args
        self.student = TitleBodySimModel(args)
        self.target_loss = nn.CrossEntropyLoss()
        self.transform_matrix = nn.ModuleList([nn.Linear(args.news_dim, args.news_dim) for _ in range(args.num_teachers)])
        for module in self.transform_matrix:
            nn.init.xavier_uniform_(module.weight, gain=1.0)
            nn.init.constant_(module.bias, 0.0)

    def train_state(self):
        dev = self.student.news_encoder.dense.weight.device
        if dev.type != "cuda":
            raise TinyRecError("tinyrec DistillModel needs its parameters on a CUDA device (no CPU path)")
        return _get_state(self, lambda: _DistillState(self.student.news_encoder, list(self.transform_matrix), dev))

    def lr_ranges(self, lr_bert=1e-6, lr_rest=1e-5):
        """The two parameter groups of cell 18 as flat-buffer ranges for ``Adam.set_lr_ranges``: the trainable
        ``bert_model`` layers, then the pooling head, ``dense`` and ``transform_matrix``."""
        st = self.train_state()
        ne = self.student.news_encoder
        rest = [p for p in (ne.attn.att_fc1.weight, ne.dense.weight) if st.flat.has(p)]
        rest += [lin.weight for lin in self.transform_matrix if st.flat.has(lin.weight)]
        cut = min(st.flat.off(p) for p in rest) if rest else st.flat.numel
        return [(0, cut, lr_bert), (cut, st.flat.numel, lr_rest)]

    def forward(self, title, body, labels, teacher_titles, teacher_bodies):
        st = self.train_state()
        want_grad = torch.is_grad_enabled() and st.flat is not None
        args = (title, body, labels, list(teacher_titles), list(teacher_bodies), want_grad, bool(self.training))
        if want_grad:
            return _DistillFn.apply(st.anchor, st, args)
        w = st.step_forward(*args)
        ls = w["losses"]
        return ls[3].clone(), ls[2].clone(), ls[0].clone(), ls[1].clone(), w["score"].clone()


class DomainTitleBodySimModel(nn.Module):
    """``TitleBodySimModel`` of Domian-specific_Post-train.ipynb (cells 10-11): the same two-pass title / body encoder
    with a plain cross-entropy over the 1+K title scores, ``forward(title, body, labels) -> (scores, loss)``;
    state_dict keys ``news_encoder.*``.  This one runs as published, so it IS pinned on the reference
    (tests/golden/post_train.npz, generated by executing the notebook's own cells)."""

    def __init__(self, args):
        super().__init__()
        self.news_encoder = _NewsEncoder(_encoder_args(args), False)
        self.loss = nn.CrossEntropyLoss()

    def train_state(self):
        dev = self.news_encoder.dense.weight.device
        if dev.type != "cuda":
            raise TinyRecError("tinyrec DomainTitleBodySimModel needs its parameters on a CUDA device (no CPU path)")
        return _get_state(self, lambda: _DistillState(self.news_encoder, [], dev))

    def forward(self, title, body, labels):
        st = self.train_state()
        want_grad = torch.is_grad_enabled() and st.flat is not None
        args = (title, body, labels, [], [], want_grad, bool(self.training))        # no teachers: total == CE
        if want_grad:
            total, _, _, _, score = _DistillFn.apply(st.anchor, st, args)
            return score, total
        w = st.step_forward(*args)
        return w["score"].clone(), w["losses"][3].clone()
