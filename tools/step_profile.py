#!/usr/bin/env python
"""Per-op device-time breakdown of one KD train step (warm caches, CUDA events around every op
wrapper of tinyrec.ops, same workload as bench.py).  Not a bench number: bracketing every launch
with events serialises nothing but adds host overhead; use it to rank kernels.

    python tools/step_profile.py [--workload kd4|kd2] [--steps 5] [--dropout 0.1]
"""
import argparse
import collections
import json
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--workload", default="kd4")
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--out", default="gpurun_out/step_profile.json")
    a = ap.parse_args()
    import tinyrec.ops as ops
    import tinyrec.optim as topt
    dev = torch.device("cuda", 0)
    torch.cuda.set_device(0)
    wl = bench.WORKLOADS[a.workload]
    model, _ = bench.make_model(wl["layers"], wl["trainable"], dev)
    opt = topt.Adam(model, lr=1e-4)
    batcher, dev_batches, _ = bench.make_inputs(0, dev)

    def step(idx_batch):
        b = bench.assembled(batcher, idx_batch)
        opt.zero_grad()
        out = model(*b)
        out[0].backward()
        opt.step()

    for i in range(3):
        step(dev_batches[i % len(dev_batches)])
    torch.cuda.synchronize()
    ops.stats.op_events = []
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(a.steps):
        step(dev_batches[i % len(dev_batches)])
    e1.record()
    torch.cuda.synchronize()
    ev, ops.stats.op_events = ops.stats.op_events, None
    total = e0.elapsed_time(e1) / a.steps
    agg = collections.OrderedDict()
    for name, tag, s, e in ev:
        k = f"{name} {tag}".strip()
        c = agg.setdefault(k, [0, 0.0])
        c[0] += 1
        c[1] += s.elapsed_time(e)
    rows = sorted(agg.items(), key=lambda kv: -kv[1][1])
    covered = sum(v[1] for _, v in rows) / a.steps
    print(f"step {total:.3f} ms (with event overhead); ops cover {covered:.3f} ms")
    for k, (cnt, ms) in rows:
        print(f"{ms / a.steps:9.3f} ms {cnt / a.steps:6.1f}x {100 * ms / a.steps / total:5.1f}%  {k}")
    os.makedirs(os.path.dirname(a.out), exist_ok=True)
    with open(a.out, "w") as f:
        json.dump({"step_ms": total, "ops": {k: {"count": v[0] / a.steps, "ms": v[1] / a.steps} for k, v in rows}}, f, indent=1)


if __name__ == "__main__":
    main()
