#!/usr/bin/env python
"""Bring-up diagnostics + timing for tnr_gemm_bf16 on a B200 (run under gpurun).

Prints, per operand-layout variant, the relative error against torch matmul; on a
mismatch it probes with identity operands to show which K / MN permutation the tensor
core actually saw (wrong swizzle / descriptor stride shows up as a permutation).
Writes a JSON summary to gpurun_out/gemm_bringup.json.
"""
import json
import os
import sys
import time

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import tinyrec.ops as ops  # noqa: E402

BF = torch.bfloat16
dev = "cuda"
results = {}


def rel(got, ref):
    return float((got.float() - ref.float()).norm() / (ref.float().norm() + 1e-12))


def probe_perm(a_t, b_t):
    """A = identity-ish, B random: C[m, n] should equal B[n, m] for m < K."""
    M, N, K = 128, 256, 128
    eye = torch.eye(K, device=dev, dtype=BF)
    a = eye.t().contiguous() if a_t else eye            # logical A[M=K, K]
    g = torch.Generator(device=dev).manual_seed(5)
    blog = torch.randn(N, K, generator=g, device=dev).to(BF)     # logical B[N, K]
    b = blog.t().contiguous() if b_t else blog
    out = torch.empty(M, N, device=dev, dtype=torch.float32)
    ops.gemm(a, b, out, a_t=a_t, b_t=b_t)
    torch.cuda.synchronize()
    ref = blog.float().t()                               # [K(=M), N]
    bad_rows = (out - ref).abs().amax(1) > 1e-3
    msg = [f"  identity probe a_t={a_t} b_t={b_t}: {int(bad_rows.sum())}/{M} rows wrong"]
    # for a few wrong rows find which reference row they equal
    for m in torch.nonzero(bad_rows).flatten()[:8].tolist():
        d = (ref - out[m][None, :]).abs().amax(1)
        j = int(d.argmin())
        msg.append(f"    row {m}: matches ref row {j} (err {float(d[j]):.3g}); col-wise match of row m: "
                   f"{int(((out[m] - ref[m]).abs() < 1e-3).sum())}/{N}")
    return "\n".join(msg)


def check(name, M, N, K, a_t=False, b_t=False, **kw):
    g = torch.Generator(device=dev).manual_seed(1)
    alog = torch.randn(M, K, generator=g, device=dev).to(BF)
    blog = torch.randn(N, K, generator=g, device=dev).to(BF)
    a = alog.t().contiguous() if a_t else alog
    b = blog.t().contiguous() if b_t else blog
    split = kw.pop("split_k", 1)
    out = torch.zeros(M, N, device=dev, dtype=torch.float32 if (split > 1 or a_t) else BF)
    try:
        ops.gemm(a, b, out, a_t=a_t, b_t=b_t, split_k=split, accumulate=split > 1, **kw)
        torch.cuda.synchronize()
        r = rel(out, alog.float() @ blog.float().t())
    except Exception as ex:  # noqa: BLE001
        print(f"[{name}] EXCEPTION {ex}")
        results[name] = {"error": str(ex)}
        return False
    ok = r < 5e-3
    print(f"[{name}] M={M} N={N} K={K} a_t={a_t} b_t={b_t} split={split}: rel_err={r:.3e} {'OK' if ok else 'FAIL'}")
    results[name] = {"rel_err": r, "ok": ok}
    if not ok:
        try:
            print(probe_perm(a_t, b_t))
        except Exception as ex:  # noqa: BLE001
            print("  probe failed:", ex)
    return ok


def bench(name, M, N, K, a_t=False, b_t=False, iters=20, **kw):
    g = torch.Generator(device=dev).manual_seed(1)
    a = torch.randn((K, M) if a_t else (M, K), generator=g, device=dev).to(BF)
    b = torch.randn((K, N) if b_t else (N, K), generator=g, device=dev).to(BF)
    split = kw.get("split_k", 1)
    out = torch.zeros(M, N, device=dev, dtype=torch.float32 if split > 1 else BF)
    if split > 1:
        kw["accumulate"] = True
    for _ in range(3):
        ops.gemm(a, b, out, a_t=a_t, b_t=b_t, **kw)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters):
        ops.gemm(a, b, out, a_t=a_t, b_t=b_t, **kw)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / iters
    tf = 2.0 * M * N * K / ms / 1e9
    # cuBLAS for context
    al = a.t() if a_t else a
    bl = b.t() if b_t else b
    for _ in range(3):
        torch.matmul(al, bl.t())
    torch.cuda.synchronize()
    e0.record()
    for _ in range(iters):
        torch.matmul(al, bl.t())
    e1.record()
    torch.cuda.synchronize()
    ms_cb = e0.elapsed_time(e1) / iters
    print(f"[bench {name}] M={M} N={N} K={K}: {ms:.3f} ms  {tf:.0f} TFLOP/s   (cuBLAS {ms_cb:.3f} ms "
          f"{2.0 * M * N * K / ms_cb / 1e9:.0f} TFLOP/s)")
    results["bench_" + name] = {"ms": ms, "tflops": tf, "cublas_ms": ms_cb}


if __name__ == "__main__":
    print(torch.cuda.get_device_name(0))
    ok = True
    ok &= check("nt_small", 128, 256, 64)
    ok &= check("nt_k256", 128, 256, 256)
    ok &= check("nt_tails", 300, 200, 200)
    ok &= check("nt_bn128", 256, 128, 128)
    ok &= check("nt_big", 4096, 2304, 768)
    ok &= check("pair_small", 4096, 256, 64)             # CTA-pair kernels (M >= 4096, N > 128)
    ok &= check("pair_odd", 4200, 768, 768)
    ok &= check("pair_bmn", 4200, 768, 2304, b_t=True)
    ok &= check("bmn_small", 128, 256, 64, b_t=True)
    ok &= check("bmn", 1000, 768, 3072, b_t=True)
    ok &= check("wgrad_small", 128, 256, 64, a_t=True, b_t=True)
    ok &= check("wgrad", 768, 3072, 4096, a_t=True, b_t=True, split_k=4)
    if ok or os.environ.get("TNR_BENCH_ANYWAY"):
        T = 52800
        bench("qkv", T, 2304, 768)
        bench("oproj", T, 768, 768)
        bench("ffn1_gelu", T, 3072, 768, act=ops.ACT_GELU)
        bench("ffn2", T, 768, 3072)
        bench("dgrad_ffn2", T, 3072, 768, b_t=True)
        bench("dgrad_ffn1", T, 768, 3072, b_t=True)
        bench("wgrad_ffn1", 3072, 768, T, a_t=True, b_t=True, split_k=8)
        bench("wgrad_qkv", 2304, 768, T, a_t=True, b_t=True, split_k=8)
        bench("wgrad_o", 768, 768, T, a_t=True, b_t=True, split_k=16)
    os.makedirs("gpurun_out", exist_ok=True)
    with open(os.environ.get("TNR_BRINGUP_OUT", "gpurun_out/gemm_bringup.json"), "w") as f:
        json.dump(results, f, indent=1)
    print("ALL OK" if ok else "SOME FAILED")
