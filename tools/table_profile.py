#!/usr/bin/env python
"""Per-op device-time breakdown of one table-build batch (tools/step_profile.py for the forward-only
workloads).  python tools/table_profile.py --workload table_long"""
import argparse
import collections
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--workload", default="table_long")
    ap.add_argument("--steps", type=int, default=3)
    a = ap.parse_args()
    import tinyrec.model_bert_2 as mb2
    import tinyrec.ops as ops
    import tinyrec.synth as synth
    wl = bench.TABLE_WORKLOADS[a.workload]
    dev = torch.device("cuda", 0)
    torch.cuda.set_device(0)
    model = mb2.ModelBert(synth.demo_args(num_hidden_layers=wl["layers"]))
    model.load_state_dict(synth.model_bert_state("", wl["layers"], 0), strict=True)
    model.to(dev).eval()
    news = synth.news_table(4 * wl["rows"], L=wl["L"], seed=1234, mean_len=wl["mean_len"], std_len=wl["std_len"],
                            min_len=wl["min_len"])
    tab = torch.from_numpy(news[:4 * wl["rows"]]).to(dev)
    with torch.no_grad():
        for i in range(2):
            model.news_encoder(tab[i * wl["rows"]:(i + 1) * wl["rows"]].to(torch.int64))
        torch.cuda.synchronize()
        ops.stats.op_events = []
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for i in range(a.steps):
            model.news_encoder(tab[(i % 4) * wl["rows"]:(i % 4 + 1) * wl["rows"]].to(torch.int64))
        e1.record()
        torch.cuda.synchronize()
    ev, ops.stats.op_events = ops.stats.op_events, None
    total = e0.elapsed_time(e1) / a.steps
    agg = collections.OrderedDict()
    for name, tag, s, e in ev:
        c = agg.setdefault(f"{name} {tag}".strip(), [0, 0.0])
        c[0] += 1
        c[1] += s.elapsed_time(e)
    print(f"{a.workload}: step {total:.3f} ms; ops cover {sum(v[1] for v in agg.values()) / a.steps:.3f} ms")
    for k, (cnt, ms) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        print(f"{ms / a.steps:9.3f} ms {cnt / a.steps:6.1f}x {100 * ms / a.steps / total:5.1f}%  {k}")


if __name__ == "__main__":
    main()
