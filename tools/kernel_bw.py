#!/usr/bin/env python
"""Achieved HBM bandwidth of the memory-bound kernels at the kd4 step's shapes (T = 52 800 tokens,
n = 1 760 news, E = 768): CUDA events around `reps` back-to-back launches over rotating buffers whose
total footprint exceeds the 126 MB L2, algorithmic bytes per launch over the
average launch time, against MEASURED_PEAKS.json:hbm_gbs.

    python tools/kernel_bw.py [--out gpurun_out/kernel_bw.json]
"""
import argparse
import json
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402


def timed(fn, n_variants, reps=40):
    for i in range(min(n_variants, 4)):
        fn(i % n_variants)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(reps):
        fn(i % n_variants)
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps * 1e-3


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--out", default="gpurun_out/kernel_bw.json")
    a = ap.parse_args()
    import tinyrec.engine as eng
    import tinyrec.ops as ops
    import tinyrec.synth as synth
    dev = torch.device("cuda", 0)
    torch.cuda.set_device(0)
    _, peak_bw, how = bench.peaks()
    n, L, E, Q, F, D, NV = 1760, 30, 768, 200, 3072, 256, 4       # NV rotating buffer sets
    T = n * L
    BF, F32 = torch.bfloat16, torch.float32
    g = torch.Generator(device=dev).manual_seed(0)
    rn = lambda *s, dt=BF: (torch.randn(*s, device=dev, generator=g) * 0.5).to(dt)  # noqa: E731
    rows = []

    def rec(name, byts, sec, note=""):
        rows.append({"kernel": name, "algorithmic_MB": byts / 1e6, "us": sec * 1e6, "GBps": byts / sec / 1e9,
                     "frac_of_hbm_peak": byts / sec / 1e9 / peak_bw, "note": note})

    # ---- embedding gather + LN (tnr_embed_ln_fwd)
    news = synth.news_table(bench.N_NEWS, L=L, seed=1234)
    idx = torch.randint(1, bench.N_NEWS, (NV, n), generator=torch.Generator().manual_seed(1))
    xs = [torch.from_numpy(news[idx[i].numpy()].astype("int64")).to(dev) for i in range(NV)]
    word = rn(30522, E)
    pos, type0 = rn(512, E, dt=F32), rn(E, dt=F32)
    gamma, beta = rn(E, dt=F32) + 1, rn(E, dt=F32)
    outs = [torch.empty(T, E, device=dev, dtype=BF) for _ in range(NV)]
    sec = timed(lambda i: ops.embed_ln(xs[i], L, word, pos, type0, gamma, beta, 1e-12, outs[i]), NV)
    rec("tnr_embed_ln_fwd", T * E * 4, sec, "bf16 word rows gathered + bf16 rows written; ids MIND-shaped")
    seed = torch.tensor([5], device=dev, dtype=torch.int64)
    sec = timed(lambda i: ops.embed_ln(xs[i], L, word, pos, type0, gamma, beta, 1e-12, outs[i],
                                       drop=ops.make_drop(seed, 0, 0.1)), NV)
    rec("tnr_embed_ln_fwd+dropout", T * E * 4, sec)

    # ---- LayerNorm fwd / bwd
    pre = [rn(T, E) for _ in range(NV)]
    sec = timed(lambda i: ops.layernorm_fwd(pre[i], gamma, beta, 1e-12, outs[i]), NV)
    rec("tnr_layernorm_fwd", T * E * 4, sec)
    dy = [rn(T, E) for _ in range(NV)]
    dxs = [torch.empty(T, E, device=dev, dtype=BF) for _ in range(NV)]
    dg, db, ds = torch.zeros(E, device=dev), torch.zeros(E, device=dev), torch.zeros(E, device=dev)
    sec = timed(lambda i: ops.layernorm_bwd(dy[i], pre[i], gamma, 1e-12, dxs[i], dg, db, dsum=ds), NV)
    rec("tnr_layernorm_bwd", T * E * 6, sec, "reads dy, x; writes dx")
    dxd = torch.empty(T, E, device=dev, dtype=BF)
    sec = timed(lambda i: ops.layernorm_bwd(dy[i], pre[i], gamma, 1e-12, dxs[i], dg, db, dx_drop=dxd,
                                            drop=ops.make_drop(seed, 9, 0.1), dsum=ds), NV)
    rec("tnr_layernorm_bwd+dropout", T * E * 8, sec, "also writes the masked dx")

    # ---- word-level additive pooling
    e = [torch.tanh(rn(T, Q)) for _ in range(NV)]
    w2, b2 = rn(Q, dt=F32) * 0.1, torch.zeros(1, device=dev)
    pooled = torch.empty(n, E, device=dev, dtype=BF)
    aw = torch.empty(n, L, device=dev, dtype=F32)
    sec = timed(lambda i: ops.attnpool_fwd(pre[i], e[i], Q, w2, b2, None, pooled, aw, n, L), NV)
    rec("tnr_attnpool_fwd", T * (E + Q) * 2, sec)
    dout = rn(n, E, dt=F32)
    du = torch.empty(T, Q, device=dev, dtype=BF)
    dw2, db2 = torch.zeros(Q, device=dev), torch.zeros(1, device=dev)
    sec = timed(lambda i: ops.attnpool_bwd(pre[i], e[i], Q, w2, aw, dout, dxs[i], du, dw2, db2, n, L), NV)
    rec("tnr_attnpool_bwd", T * (E + Q) * 2 * 2, sec, "reads x, e; writes dx, du")

    # ---- attention
    qkv = [rn(T, 3 * E) for _ in range(NV)]
    relpos = rn(12, 2 * L - 1, dt=F32)
    sec = timed(lambda i: ops.attn_fwd(qkv[i], xs[i], L, relpos, outs[i], 12), NV)
    rec("tnr_attn_relpos_fwd", T * 4 * E * 2, sec)
    sec = timed(lambda i: ops.attn_fwd(qkv[i], xs[i], L, relpos, outs[i], 12, drop=ops.make_drop(seed, 8, 0.1)), NV)
    rec("tnr_attn_relpos_fwd+dropout", T * 4 * E * 2, sec)
    dqkv = torch.empty(T, 3 * E, device=dev, dtype=BF)
    sec = timed(lambda i: ops.attn_bwd(qkv[i], xs[i], L, relpos, dy[i], dqkv, 12, drop=ops.make_drop(seed, 8, 0.1)), NV)
    rec("tnr_attn_relpos_bwd+dropout", T * 7 * E * 2, sec)
    dbq = torch.zeros(3 * E, device=dev)
    sec = timed(lambda i: ops.attn_bwd(qkv[i], xs[i], L, relpos, dy[i], dqkv, 12, drop=ops.make_drop(seed, 8, 0.1), dbias=dbq), NV)
    rec("tnr_attn_relpos_bwd+dropout+dbias", T * 7 * E * 2, sec, "as the train step calls it: + column sums of dqkv (the QKV bias gradient)")

    # ---- column sums (bias gradients)
    dz = [rn(T, F) for _ in range(2)]
    cs = torch.zeros(F, device=dev)
    sec = timed(lambda i: ops.colsum(dz[i], cs), 2)
    rec("tnr_colsum_bf16 [T,3072]", T * F * 2, sec)

    # ---- batch-assembly gathers (teacher embedding rows, news token rows)
    table = rn(bench.N_NEWS + 1, D, dt=F32)
    ncat = torch.from_numpy(news).to(dev)
    gi = [torch.randint(0, bench.N_NEWS, (32 * 55,), device=dev, dtype=torch.int32) for _ in range(NV)]
    go = torch.empty(32 * 55, D, device=dev)
    sec = timed(lambda i: ops.gather_rows_f32(table, gi[i], go), NV)
    rec("tnr_gather_rows_f32 (1 760 rows x 1 KB)", 32 * 55 * D * 4 * 2, sec, "launch-latency bound at this size")
    big = torch.randint(0, bench.N_NEWS, (1 << 20,), device=dev, dtype=torch.int32)
    gob = torch.empty(1 << 20, D, device=dev)
    sec = timed(lambda i: ops.gather_rows_f32(table, big, gob), 1, reps=10)
    rec("tnr_gather_rows_f32 (1 Mi rows x 1 KB)", (1 << 20) * D * 4 * 2, sec, "table (165 MB) read + 1 GiB written")
    go64 = torch.empty(1 << 20, ncat.shape[1], device=dev, dtype=torch.int64)
    sec = timed(lambda i: ops.gather_rows_i32_i64(ncat, big, go64), 1, reps=10)
    rec("tnr_gather_rows_i32_i64 (1 Mi rows)", (1 << 20) * ncat.shape[1] * 12, sec, "int32 row read, int64 row written")

    # ---- Adam
    nP = 14841634 // 8 * 8
    p, gr, m, v, vm = (torch.zeros(nP, device=dev) for _ in range(5))
    sh = torch.zeros(nP, device=dev, dtype=BF)
    sec = timed(lambda i: ops.adam_amsgrad(p, gr, m, v, vm, sh, 1e-4, 0.9, 0.999, 1e-8, i + 1), 1)
    rec("tnr_adam_amsgrad", nP * (9 * 4 + 2), sec, "59 MB per array: partly L2-resident between launches")

    os.makedirs(os.path.dirname(a.out) or ".", exist_ok=True)
    with open(a.out, "w") as f:
        json.dump({"hbm_peak_GBps": peak_bw, "peak_source": how, "rows": rows}, f, indent=1)
    print(f"HBM peak {peak_bw:.0f} GB/s ({how})")
    for r in rows:
        print(f"{r['kernel']:<44s} {r['algorithmic_MB']:9.1f} MB {r['us']:8.1f} us {r['GBps']:8.0f} GB/s "
              f"{100 * r['frac_of_hbm_peak']:5.1f}%  {r['note']}")


if __name__ == "__main__":
    main()
