"""Fused Adam(amsgrad) over the model's flat parameter buffer and the NCCL data-parallel
wrapper that replaces Horovod.

Reference: ``optim.Adam(model.parameters(), lr=args.lr, amsgrad=True)`` (Tiny-NewsRec/run.py:134),
``hvd.broadcast_parameters`` / ``hvd.DistributedOptimizer(op=hvd.Average)`` (run.py:142-149).
"""
import torch
import torch.distributed as dist

from . import ops
from ._lib import TinyRecError


class Adam:
    """``Adam(model, lr, amsgrad=True)``: one kernel launch per step over all trainable parameters
    (fp32 master weights, m, v, vmax) that also refreshes the bf16 shadow weights the GEMMs read."""

    def __init__(self, model, lr=1e-4, betas=(0.9, 0.999), eps=1e-8, amsgrad=True, capturable=False):
        self.amsgrad = bool(amsgrad)          # False: torch's default Adam (Post-train_KD.ipynb cell 18); True: run.py:134
        self.model, self.lr, self.betas, self.eps = model, lr, betas, eps
        self.step_count = 0
        self.m = self.v = self.vmax = None
        self.grad_scale = 1.0
        # capturable: the step counter lives on the device so that a captured CUDA graph of the train step
        # (tinyrec.run.GraphedTrainStep) replays with the right bias corrections
        self.capturable = bool(capturable)
        self.step_dev = self.bc_ws = None

    def _flat(self):
        st = self.model.train_state()
        if st.flat is None:
            raise TinyRecError("model has no trainable parameters")
        return st.flat

    def zero_grad(self, set_to_none=False):
        flat = self._flat()
        flat.reattach_grads()
        flat.grad.zero_()

    def make_capturable(self):
        if not self.capturable:
            self.capturable = True
            self.step_dev = None

    def ensure_state(self):
        """Allocate m / v / vmax (and the device step counter when capturable) without stepping."""
        flat = self._flat()
        if self.m is None or self.m.numel() != flat.numel or self.m.device != flat.data.device:
            self.m = torch.zeros_like(flat.data)
            self.v = torch.zeros_like(flat.data)
            self.vmax = torch.zeros_like(flat.data) if self.amsgrad else None
        if self.capturable and self.step_dev is None:
            self.step_dev = torch.tensor([self.step_count], device=flat.data.device, dtype=torch.int32)
            self.bc_ws = torch.zeros(2, device=flat.data.device, dtype=torch.float32)
        return flat

    def set_lr_ranges(self, ranges):
        """Per-range learning rates over the flat buffer: ``[(lo, hi, lr), ...]`` in elements, contiguous and
        covering it (the post-training notebook's two groups: 1e-6 for ``bert_model``, 1e-5 for the rest,
        Post-train_KD.ipynb cell 18).  One kernel launch per range."""
        self.lr_ranges = [(int(lo), int(hi), float(lr)) for lo, hi, lr in ranges]

    def step(self):
        flat = self.ensure_state()
        ops.bump_param_generation()
        ranges = getattr(self, "lr_ranges", None)
        if ranges:
            if self.capturable:
                raise TinyRecError("lr ranges are not supported together with the device step counter")
            self.step_count += 1
            for lo, hi, lr in ranges:
                if hi > lo:
                    ops.adam_amsgrad(flat.data[lo:hi], flat.grad[lo:hi], self.m[lo:hi], self.v[lo:hi],
                                     self.vmax[lo:hi] if self.vmax is not None else None,
                                     flat.shadow[lo:hi], lr, self.betas[0], self.betas[1], self.eps, self.step_count,
                                     self.grad_scale)
            return
        if self.capturable:
            ops.adam_amsgrad_devstep(flat.data, flat.grad, self.m, self.v, self.vmax, flat.shadow, self.lr, self.betas[0],
                                     self.betas[1], self.eps, self.step_dev, self.bc_ws, self.grad_scale)
            return
        self.step_count += 1
        ops.adam_amsgrad(flat.data, flat.grad, self.m, self.v, self.vmax, flat.shadow, self.lr, self.betas[0],
                         self.betas[1], self.eps, self.step_count, self.grad_scale)

    def steps_done(self):
        return int(self.step_dev) if self.capturable and self.step_dev is not None else self.step_count

    def state_dict(self):
        return dict(step=self.steps_done(), m=self.m, v=self.v, vmax=self.vmax, lr=self.lr)

    def load_state_dict(self, sd):
        self.step_count, self.m, self.v, self.vmax, self.lr = sd["step"], sd["m"], sd["v"], sd["vmax"], sd["lr"]
        self.step_dev = None


def broadcast_parameters(model, root_rank=0):
    """hvd.broadcast_parameters(model.state_dict(), root_rank=0), run.py:142,247."""
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size() == 1:
        return
    st = getattr(model, "train_state", None)
    flat = st().flat if st is not None else None
    if flat is not None:
        dist.broadcast(flat.data, src=root_rank)
        flat.refresh_shadow()
    seen = set(id(p) for p in flat.params) if flat is not None else set()
    for t in list(model.parameters()) + list(model.buffers()):
        if id(t) not in seen:
            dist.broadcast(t.data, src=root_rank)


class DistributedOptimizer:
    """Average the flat gradient buffer over ranks (NCCL all-reduce over NVLink), then step.

    ``overlap=True`` launches one all-reduce per gradient bucket on a side stream as soon as the
    backward has finished that bucket (TrainState.comm_hook: top encoder layer + heads first, lower
    layers as they complete), so the exchange of layer i overlaps the backward of layer i-1 like
    Horovod's per-tensor hooks did; ``step()`` waits for all of them."""

    def __init__(self, optimizer, overlap=True):
        self.opt = optimizer
        self.world = dist.get_world_size() if dist.is_initialized() else 1
        self.opt.grad_scale = 1.0 / self.world          # SUM all-reduce, scale folded into Adam
        self.stream = torch.cuda.Stream() if (self.world > 1 and overlap) else None
        self.pending = []
        if self.stream is not None:
            self.opt.model.train_state().comm_hook = self._launch

    def _launch(self, flat, lo=0, hi=None):
        hi = flat.numel if hi is None else hi
        ev = torch.cuda.Event()
        ev.record()
        self.stream.wait_event(ev)
        with torch.cuda.stream(self.stream):
            dist.all_reduce(flat.grad[lo:hi], op=dist.ReduceOp.SUM)
            done = torch.cuda.Event()
            done.record()
        self.pending.append(done)

    def zero_grad(self, set_to_none=False):
        self.opt.zero_grad()

    def step(self):
        if self.world > 1:
            if self.stream is not None and self.pending:
                for ev in self.pending:
                    torch.cuda.current_stream().wait_event(ev)
                self.pending = []
            else:
                dist.all_reduce(self.opt._flat().grad, op=dist.ReduceOp.SUM)
        self.opt.step()
