#!/usr/bin/env python
"""Device time of the small fp32 head kernels of one KD step (kd_loss, user encoders, pooling) from the per-op events of
tools/step_profile.py's machinery, 20 steps.  Usage: python tools/head_timing.py"""
import collections
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402


def main():
    import tinyrec.ops as ops
    import tinyrec.optim as topt
    dev = torch.device("cuda", 0)
    wl = bench.WORKLOADS["kd4"]
    model, _ = bench.make_model(wl["layers"], wl["trainable"], dev)
    opt = topt.Adam(model, lr=1e-4)
    batcher, dev_batches, _ = bench.make_inputs(0, dev)

    def step(idx_batch):
        b = bench.assembled(batcher, idx_batch)
        opt.zero_grad()
        model(*b)[0].backward()
        opt.step()

    for i in range(3):
        step(dev_batches[i % len(dev_batches)])
    torch.cuda.synchronize()
    ops.stats.op_events = []
    n = 20
    for i in range(n):
        step(dev_batches[i % len(dev_batches)])
    torch.cuda.synchronize()
    ev, ops.stats.op_events = ops.stats.op_events, None
    agg = collections.OrderedDict()
    for name, tag, s, e in ev:
        if name == "gemm" and not tag.startswith("1760") and "x200" not in tag and not tag.startswith("200x") and not tag.startswith("256x"):
            continue
        c = agg.setdefault(f"{name} {tag}".strip(), [0, 0.0])
        c[0] += 1
        c[1] += s.elapsed_time(e)
    for k, (c, ms) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        print(f"{1e3 * ms / n:8.1f} us/step {c / n:5.1f}x  {k}")


if __name__ == "__main__":
    main()
