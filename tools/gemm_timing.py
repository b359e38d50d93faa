#!/usr/bin/env python
"""Debug: clock64 stamps of CTA 0 of the tcgen05 GEMM (library built with TNR_EXTRA_NVCC_FLAGS=-DGEMM_TIMING): when the MMA
thread waited for / got the accumulator stage and each k-block's operands, when it had issued them, and when epilogue
warp 0 saw the accumulator, released it and issued its store.  Usage: python tools/gemm_timing.py [qkv|oproj|ffn2]"""
import ctypes
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import tinyrec._lib as L  # noqa: E402
import tinyrec.ops as ops  # noqa: E402


def main():
    if len(sys.argv) > 2:                        # several shapes: launch times only, one process
        for w in sys.argv[1:]:
            one(w, stamps=False)
        return
    one(sys.argv[1] if len(sys.argv) > 1 else "qkv")


def one(which, stamps=True):
    M = 52800
    N, K, res, drop = {"qkv": (2304, 768, False, False), "oproj": (768, 768, True, True), "ffn2": (768, 3072, True, True),
                       "ffn1": (3072, 768, False, False), "ffn1_frozen": (3072, 768, False, False), "dffn2": (3072, 768, False, False)}[which]
    g = torch.Generator(device="cuda").manual_seed(0)
    a = (torch.randn(M, K, device="cuda", generator=g)).to(torch.bfloat16)
    b = (torch.randn(N, K, device="cuda", generator=g) * 0.05).to(torch.bfloat16)
    bias = torch.randn(N, device="cuda", generator=g)
    out = torch.empty(M, N, device="cuda", dtype=torch.bfloat16)
    r = torch.randn(M, N, device="cuda", generator=g).to(torch.bfloat16) if res else None
    seed = torch.tensor([1], device="cuda", dtype=torch.int64)
    d = ops.make_drop(seed, 3, 0.1) if drop else None
    act, aux = ops.ACT_NONE, None
    if which == "ffn1":
        act, aux = ops.ACT_GELU_DAUX, torch.empty_like(out)
    elif which == "ffn1_frozen":
        act = ops.ACT_GELU
    elif which == "dffn2":                       # dX = (dY @ W2) * gelu'(z): B given [K, N], the derivative read as aux
        act, aux, bias = ops.ACT_MULAUX, torch.randn(M, N, device="cuda", generator=g).to(torch.bfloat16), None
        b = b.t().contiguous()
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(2)]
    n_rep = 3 if stamps else 20
    for i in range(3 + n_rep):
        if i == 3:
            ev[0].record()
        ops.gemm(a, b, out, bias=bias, residual=r, drop=d, act=act, aux=aux, b_t=(which == "dffn2"))
    ev[1].record()
    torch.cuda.synchronize()
    us = ev[0].elapsed_time(ev[1]) / n_rep * 1e3
    print(f"{which}: {us:.1f} us per launch, {2.0 * M * N * K / us * 1e-6:.0f} TFLOP/s")
    lib = L.load()
    if not stamps or not hasattr(lib, "tnr_debug_gemm_stamps"):
        return                                   # production build: only the launch time
    buf = (ctypes.c_longlong * 8192)()
    lib.tnr_debug_gemm_stamps.argtypes = [ctypes.c_void_p, ctypes.c_int]
    assert lib.tnr_debug_gemm_stamps(buf, 8192) == 0
    t0 = buf[0]
    print(f"{which}: M={M} N={N} K={K}; cycles relative to the MMA thread's first stamp")
    print("tile  tempty_wait  tempty_ok | per k-block: operands_ok -> issued ... | epi0: tfull_wait tfull_ok released store_issued")
    for tl in range(9):
        m = [buf[tl * 32 + i] - t0 for i in range(32)]
        e = [buf[4096 + tl * 8 + i] - t0 for i in range(4)]
        kbs = " ".join(f"{m[2 + 2 * k]}>{m[3 + 2 * k]}" for k in range(min(12, K // 64)))
        print(f"{tl:3d}  {m[0]:9d} {m[1]:9d} | {kbs} | {e[0]} {e[1]} {e[2]} {e[3]}")


if __name__ == "__main__":
    main()
