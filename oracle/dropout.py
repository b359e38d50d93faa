"""TEST INFRASTRUCTURE ONLY (see oracle/__init__.py): numpy restatement of the counter-based
dropout generator of the CUDA path (tiny-newsrec_b200/csrc/common.cuh: philox4x32_7,
dropout_keep8), used to give the CPU oracle the *same* masks as the kernels so training-mode
forward / backward can be compared exactly.

The reference draws its masks from torch's global Philox stream (nn.Dropout at
tnlrv3/modeling.py:177,223 and transformers BertSelfOutput / BertOutput); that stream cannot be
reproduced outside torch's kernels, so parity for the dropout path is defined as: same
algorithm (keep with probability 1-p, scale kept values by 1/(1-p), masks independent per
element / site / step) with masks injected into the oracle.  parity unpinned for the RNG itself.
"""
import numpy as np

M0, M1 = np.uint64(0xD2511F53), np.uint64(0xCD9E8D57)
W0, W1 = np.uint32(0x9E3779B9), np.uint32(0xBB67AE85)
MASK32 = np.uint64(0xFFFFFFFF)


def philox4x32_7(c0, c1, c2, c3, k0, k1):
    """Vectorised Philox4x32 with 7 rounds; all arguments uint32 arrays (or scalars)."""
    c0, c1, c2, c3 = (np.asarray(c, dtype=np.uint32).copy() for c in (c0, c1, c2, c3))
    k0 = np.uint32(k0)
    k1 = np.uint32(k1)
    with np.errstate(over="ignore"):
        for _ in range(7):
            p0 = M0 * c0.astype(np.uint64)
            p1 = M1 * c2.astype(np.uint64)
            hi0, lo0 = (p0 >> np.uint64(32)).astype(np.uint32), (p0 & MASK32).astype(np.uint32)
            hi1, lo1 = (p1 >> np.uint64(32)).astype(np.uint32), (p1 & MASK32).astype(np.uint32)
            c0, c1, c2, c3 = hi1 ^ c1 ^ k0, lo1, hi0 ^ c3 ^ k1, lo0
            k0 = np.uint32((int(k0) + int(W0)) & 0xFFFFFFFF)
            k1 = np.uint32((int(k1) + int(W1)) & 0xFFFFFFFF)
    return c0, c1, c2, c3


def keep_mask(seed, site, n, p):
    """bool [n]: keep flag of linear element index 0..n-1 of dropout tensor ``site``."""
    if not p > 0.0:
        return np.ones(n, dtype=bool)
    thr = np.uint32(int(np.float32(p) * np.float32(65536.0) + np.float32(0.5)))
    groups = (n + 7) // 8
    g = np.arange(groups, dtype=np.uint64)
    r = philox4x32_7((g & MASK32).astype(np.uint32), (g >> np.uint64(32)).astype(np.uint32),
                     np.full(groups, site, dtype=np.uint32), np.zeros(groups, dtype=np.uint32),
                     seed & 0xFFFFFFFF, (seed >> 32) & 0xFFFFFFFF)
    lanes = np.empty((groups, 8), dtype=np.uint32)
    for i in range(4):
        lanes[:, 2 * i] = r[i] & np.uint32(0xFFFF)
        lanes[:, 2 * i + 1] = r[i] >> np.uint32(16)
    return (lanes >= thr).reshape(-1)[:n]


def attention_keep(seed, site, n_items, L, p):
    """bool [n_items, L, L] keep flags of the attention probabilities (item = news * A + head).
    Element (item, i, j) has linear index ((item*32 + i)*4 + ((j >> 1) & 3))*8 + (j >> 3)*2 + (j & 1)
    (tiny-newsrec_b200/csrc/attention.cu: attn_keep8)."""
    flat = keep_mask(seed, site, n_items * 32 * 32, p).reshape(n_items, 32, 4, 8)
    j = np.arange(L)
    return flat[:, :L][:, :, (j >> 1) & 3, (j >> 3) * 2 + (j & 1)]


# dropout tensor ids of the CUDA path (tiny-newsrec_b200/engine.py: SITE_EMB, drop_site)
SITE_EMB = 0
KIND_ATTN, KIND_ATT_OUT, KIND_FFN_OUT = 0, 1, 2


def drop_site(layer, kind):
    return 8 * (layer + 1) + kind


class Plan:
    """Multipliers keep / (1 - p) for every dropout tensor of one encoder forward over the
    concatenated news batch (rows = news; the CUDA path encodes [history rows | candidate rows] in one
    pass, so callers pass the row offset of the slice they encode)."""

    def __init__(self, seed, p_hidden=0.1, p_attn=0.1):
        self.seed, self.p_hidden, self.p_attn = int(seed), float(p_hidden), float(p_attn)

    def _rows(self, site, row0, n, L, E, p):
        import torch
        full = keep_mask(self.seed, site, (row0 + n) * L * E, p)[row0 * L * E:]
        return torch.from_numpy(full.reshape(n, L, E).astype(np.float32)) / (1.0 - p)

    def emb(self, row0, n, L, E):
        return self._rows(SITE_EMB, row0, n, L, E, self.p_hidden)

    def dense(self, layer, kind, row0, n, L, E):
        return self._rows(drop_site(layer, kind), row0, n, L, E, self.p_hidden)

    def attn(self, layer, row0, n, A, L):
        import torch
        k = attention_keep(self.seed, drop_site(layer, KIND_ATTN), (row0 + n) * A, L, self.p_attn)[row0 * A:]
        return torch.from_numpy(k.reshape(n, A, L, L).astype(np.float32)) / (1.0 - self.p_attn)
