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
        ranges = getattr(self, "lr_ranges", None)
        if ranges:
            if self.capturable:
                raise TinyRecError("lr ranges are not supported together with the device step counter")
            ops.bump_param_generation()
            self.step_count += 1
            for lo, hi, lr in ranges:
                if hi > lo:
                    ops.adam_amsgrad(flat.data[lo:hi], flat.grad[lo:hi], self.m[lo:hi], self.v[lo:hi],
                                     self.vmax[lo:hi] if self.vmax is not None else None,
                                     flat.shadow[lo:hi], lr, self.betas[0], self.betas[1], self.eps, self.step_count,
                                     self.grad_scale)
            return
        self.step_ranges([(0, flat.numel)], advance=True)

    def step_ranges(self, ranges, advance=True):
        """Update the element ranges ``[(lo, hi), ...]`` of the flat buffer (8-element aligned bounds) with the bias
        corrections of ONE optimizer step: ``advance`` starts that step (counter + 1), further calls of the same step
        pass ``advance=False``.  Adam is element-wise, so a step may be applied range by range."""
        flat = self.ensure_state()
        ops.bump_param_generation()
        if advance and not self.capturable:
            self.step_count += 1
        for lo, hi in ranges:
            if hi <= lo and not (advance and self.capturable):
                continue
            sl = slice(lo, hi)
            vmax = self.vmax[sl] if self.vmax is not None else None
            if self.capturable:
                ops.adam_amsgrad_devstep(flat.data[sl], flat.grad[sl], self.m[sl], self.v[sl], vmax, flat.shadow[sl], self.lr,
                                         self.betas[0], self.betas[1], self.eps, self.step_dev, self.bc_ws, self.grad_scale,
                                         advance=advance)
            else:
                ops.adam_amsgrad(flat.data[sl], flat.grad[sl], self.m[sl], self.v[sl], vmax, flat.shadow[sl], self.lr,
                                 self.betas[0], self.betas[1], self.eps, self.step_count, self.grad_scale)
            advance = False

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

    ``overlap=True`` launches one all-reduce per gradient bucket on a side stream as soon as the backward has
    finished that bucket (TrainState.comm_hook: top encoder layer + heads first; the lowest trainable layer range by
    range), so the exchange overlaps the remaining backward like Horovod's per-tensor hooks did.  ``step()`` waits
    for all buckets but the LAST, updates everything outside it, then waits for the last one and updates that range:
    the trailing exchange (2.4 MB) hides behind the Adam pass over the other 57 MB.

    While exchanges are in flight the persistent GEMM grids leave ``sm_reserve`` SMs free (``tnr_set_sm_reserve``) and
    NCCL is capped at that many CTAs (``NCCL_MAX_CTAS``, set by ``tinyrec.parallel.init_distributed`` before the
    communicator exists): otherwise a collective's CTAs wait for a whole one-CTA-per-SM GEMM, and the GEMM CTAs they
    displace afterwards run as a second wave (+0.44 ms per step at 8 GPUs in round 1)."""

    def __init__(self, optimizer, overlap=True, sm_reserve=None):
        import os
        self.opt = optimizer
        self.world = dist.get_world_size() if dist.is_initialized() else 1
        self.opt.grad_scale = 1.0 / self.world          # SUM all-reduce, scale folded into Adam
        self.stream = torch.cuda.Stream() if (self.world > 1 and overlap) else None
        self.pending = []
        self.use_p2p = os.environ.get("TNR_P2P_ALLREDUCE", "1") != "0"
        p2p = False
        if self.stream is not None:
            flat = self.opt.model.train_state().flat
            p2p = self.use_p2p and flat is not None and getattr(flat, "symm", None) is not None
        if sm_reserve is None:
            # the peer-memory kernel needs few CTAs (it is latency-, not bandwidth-bound at these bucket sizes and runs
            # hidden behind the backward); NCCL gets what NCCL_MAX_CTAS allows it
            default = "4" if p2p else os.environ.get("NCCL_MAX_CTAS", "0")
            sm_reserve = int(os.environ.get("TNR_COMM_SM_RESERVE", default) or 0)
        self.sm_reserve = (sm_reserve + 1) // 2 * 2 if self.stream is not None else 0
        self.reserved = False
        self.trace = None           # list -> per step a dict of CUDA events (bench.py --gpus N: comm timeline)
        self.comm_ctas = max(1, self.sm_reserve or 4)
        if self.stream is not None:
            self.opt.model.train_state().comm_hook = self._launch

    def _launch(self, flat, lo=0, hi=None):
        hi = flat.numel if hi is None else hi
        if self.sm_reserve and not self.reserved:
            ops.set_sm_reserve(self.sm_reserve)         # GEMMs enqueued from here to step() leave room for the collective
            self.reserved = True
        timed = self.trace is not None
        ev = torch.cuda.Event(enable_timing=timed)
        ev.record()
        self.stream.wait_event(ev)
        with torch.cuda.stream(self.stream):
            start = None
            if timed:
                start = torch.cuda.Event(enable_timing=True)
                start.record()
            if flat.symm is not None and self.use_p2p:
                flat.symm.all_reduce(lo, hi, self.comm_ctas)        # one kernel over NVLink peer memory
            else:
                dist.all_reduce(flat.grad[lo:hi], op=dist.ReduceOp.SUM)
            done = torch.cuda.Event(enable_timing=timed)
            done.record()
        self.pending.append((lo, hi, done))
        if timed:
            self._cur.setdefault("buckets", []).append(dict(lo=lo, hi=hi, ready=ev, start=start, done=done))

    def zero_grad(self, set_to_none=False):
        if self.trace is not None:          # a step starts here (run.py:193)
            self._cur = dict(t0=torch.cuda.Event(enable_timing=True))
            self._cur["t0"].record()
        self.opt.zero_grad()

    def _mark(self, name):
        if self.trace is not None:
            ev = torch.cuda.Event(enable_timing=True)
            ev.record()
            self._cur[name] = ev

    def step(self):
        self._mark("backward_end")
        self._step()
        self._mark("step_end")
        if self.trace is not None:
            self.trace.append(self._cur)

    def _step(self):
        if self.world > 1:
            if self.stream is not None and self.pending:
                if self.reserved:
                    ops.set_sm_reserve(0)
                    self.reserved = False
                cur = torch.cuda.current_stream()
                pend, self.pending = self.pending, []
                flat = self.opt._flat()
                lo, hi, last = pend[-1]
                splittable = len(pend) > 1 and not getattr(self.opt, "lr_ranges", None) and lo % 8 == 0 and hi % 8 == 0
                for _, _, ev in pend[:-1]:
                    cur.wait_event(ev)
                if splittable:
                    self.opt.step_ranges([(0, lo), (hi, flat.numel)], advance=True)
                    cur.wait_event(last)
                    self.opt.step_ranges([(lo, hi)], advance=False)
                    return
                cur.wait_event(last)
            else:
                dist.all_reduce(self.opt._flat().grad, op=dist.ReduceOp.SUM)
        self.opt.step()
