#!/usr/bin/env python
"""Debug: clock64 stamps of CTA 0 of ue_logits_kernel (library built with TNR_EXTRA_NVCC_FLAGS=-DUL_TIMING).
Prints, per (tile, k-block) item, when the MMA thread saw its own / the peer's stage and issued, when loader thread 0
waited for / got the free stage, and the epilogue's per-tile stamps, in cycles relative to the first stamp."""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import tinyrec.ops as ops  # noqa: E402


def main():
    B, H, D, Q, N = 4096, 50, 256, 200, 161014
    g = torch.Generator(device="cuda").manual_seed(0)
    table = torch.randn(N, D, device="cuda", generator=g) * 0.1
    idx = torch.randint(0, N, (B, H), device="cuda", generator=g).int()
    mask = (torch.rand(B, H, device="cuda", generator=g) > 0.5).float()
    pad = torch.randn(D, device="cuda", generator=g)
    W1 = torch.randn(Q, D, device="cuda", generator=g) * 0.06
    b1, w2, b2 = torch.randn(Q, device="cuda", generator=g) * 0.1, torch.randn(Q, device="cuda", generator=g) * 0.1, torch.zeros(1, device="cuda")
    user, a = torch.empty(B, D, device="cuda"), torch.empty(B, H, device="cuda")
    wp = ops.user_encoder_pack_w1(W1, pad, b1, w2)
    for _ in range(3):
        ops.user_encoder_score(table, idx, mask, pad, wp, Q, b1, w2, b2, True, user, a, B, H)
    torch.cuda.synchronize()
    ws = list(ops._score_ws.values())[0]
    dbg = ws[(B * H + 4) * 4:].view(torch.int64).cpu()
    mma = dbg[:256].view(64, 4)
    ld = dbg[256:512].view(64, 4)
    ep = dbg[512:544].view(8, 4)
    t0 = int(min(mma[0, 0], ld[0, 0]))
    print("item  mma:landed  peer  issued | loader0: before_empty  after_empty  issued")
    for i in range(48):
        print(f"{i:3d}  {int(mma[i,0])-t0:9d} {int(mma[i,1])-t0:9d} {int(mma[i,2])-t0:9d} | {int(ld[i,0])-t0:9d} {int(ld[i,1])-t0:9d} {int(ld[i,2])-t0:9d}")
    print("tile  epi: wait_start  tfull  released")
    for t in range(6):
        print(f"{t:3d}  {int(ep[t,0])-t0:9d} {int(ep[t,1])-t0:9d} {int(ep[t,2])-t0:9d}")


if __name__ == "__main__":
    main()
