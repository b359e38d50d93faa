"""Oracle (test infrastructure): Adam(amsgrad=True) as used at
Tiny-NewsRec/run.py:134 (torch.optim.Adam defaults: betas (0.9, 0.999),
eps 1e-8, weight_decay 0; third-party arithmetic = torch's documented
algorithm), and the Horovod average all-reduce it is wrapped in
(run.py:145-149) restated as a mean over rank gradients."""
import math

import torch


def adam_amsgrad_step(p, g, m, v, vmax, step, lr=1e-4, b1=0.9, b2=0.999, eps=1e-8):
    """In-place single-tensor update; `step` is the 1-based step count."""
    m.mul_(b1).add_(g, alpha=1 - b1)
    v.mul_(b2).addcmul_(g, g, value=1 - b2)
    torch.maximum(vmax, v, out=vmax)
    bc1 = 1 - b1 ** step
    bc2 = 1 - b2 ** step
    denom = (vmax.sqrt() / math.sqrt(bc2)).add_(eps)
    p.addcdiv_(m, denom, value=-lr / bc1)


def allreduce_average(grads_per_rank):
    """hvd.DistributedOptimizer(op=Average), run.py:145-149."""
    return sum(grads_per_rank) / float(len(grads_per_rank))
