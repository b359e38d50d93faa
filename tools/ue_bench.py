#!/usr/bin/env python
"""Time the scoring-path user encoder variants at the eval-step shape (4 096 impressions x 50 clicks x 256)."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import tinyrec.ops as ops  # noqa: E402


def timed(fn, n=20):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n * 1e3


def main():
    B, H, D, Q, N = 4096, 50, 256, 200, 161014
    g = torch.Generator(device="cuda").manual_seed(0)
    table = torch.randn(N, D, device="cuda", generator=g) * 0.1
    idx = torch.randint(0, N, (B, H), device="cuda", generator=g).int()
    seq = torch.arange(B * H, device="cuda").view(B, H).int() % N
    mask = (torch.rand(B, H, device="cuda", generator=g) > 0.3).float()
    pad = torch.randn(D, device="cuda", generator=g)
    W1 = torch.randn(Q, D, device="cuda", generator=g) * 0.06
    b1, w2, b2 = torch.randn(Q, device="cuda", generator=g) * 0.1, torch.randn(Q, device="cuda", generator=g) * 0.1, torch.zeros(1, device="cuda")
    user, a = torch.empty(B, D, device="cuda"), torch.empty(B, H, device="cuda")
    vecs = torch.empty(B * H, D, device="cuda")
    wp = ops.user_encoder_pack_w1(W1, pad, b1, w2)
    for um in (False, True):
        print(f"use_mask={um}")
        print("  score, random gather      %8.1f us" % timed(lambda: ops.user_encoder_score(table, idx, mask, pad, wp, Q, b1, w2, b2, um, user, a, B, H)))
        print("  score, sequential idx     %8.1f us" % timed(lambda: ops.user_encoder_score(table, seq, mask, pad, wp, Q, b1, w2, b2, um, user, a, B, H)))
        ops.gather_rows_f32(table, idx.reshape(-1), vecs)
        print("  score, contiguous vecs    %8.1f us" % timed(lambda: ops.user_encoder_score(vecs, None, mask, pad, wp, Q, b1, w2, b2, um, user, a, B, H)))
        print("  per-impression kernel     %8.1f us" % timed(lambda: ops.user_encoder_fwd(vecs, mask, pad, W1, b1, w2, b2, um, user, a, None, B, H)))
    print("  gather_rows_f32           %8.1f us" % timed(lambda: ops.gather_rows_f32(table, idx.reshape(-1), vecs)))


if __name__ == "__main__":
    main()
