"""Oracle (test infrastructure): batch assembly (index gathers, padding, masks).

Reference: Tiny-NewsRec/dataloader.py:73-83 (id->index, front padding),
:118-172 (train batch), :285-314 (eval batch); sharding of impressions over
ranks: streaming.py:40-58 (files r, r+world, ...), split_file.py:36-43.
Pure numpy / Python ints -- results must be matched bit-exactly.
"""
import numpy as np


def to_index(nids, news_index):
    """dataloader.py:73-74: unknown ids map to row 0 (the all-pad news)."""
    return [news_index[i] if i in news_index else 0 for i in nids]


def pad_history(idx, fix_length):
    """dataloader.py:76-83 with padding_front=True: keep the LAST `fix_length`
    clicks, left-pad with 0; mask = [0..0, 1..1]."""
    idx = list(idx)
    n = len(idx)
    pad = [0] * (fix_length - n) + idx[-fix_length:]
    mask = [0] * (fix_length - n) + [1] * min(fix_length, n)
    return pad, mask


def insert_positive(pos, neg, label):
    """dataloader.py:135-137: candidate order = neg[:label] + pos + neg[label:]."""
    return list(neg[:label]) + list(pos) + list(neg[label:])


def train_batch(hist_idx, cand_idx, news_combined, teacher_tables):
    """dataloader.py:129-160 for already index-mapped impressions.

    hist_idx int [B,H], cand_idx int [B,K] -> (history int64 [B,H,2L],
    candidate int64 [B,K,2L], [M x f32 [B,H,D]], [M x f32 [B,K,D]])."""
    history = news_combined[hist_idx].astype(np.int64)
    candidate = news_combined[cand_idx].astype(np.int64)
    th = [t[hist_idx].astype(np.float32) for t in teacher_tables]
    tc = [t[cand_idx].astype(np.float32) for t in teacher_tables]
    return history, candidate, th, tc


def eval_batch(hist_idx, cand_idx_list, news_scoring):
    """dataloader.py:292-301: log vecs f32 [B,H,D]; ragged candidate vecs."""
    log_vecs = news_scoring[hist_idx].astype(np.float32)
    cands = [news_scoring[np.asarray(c, dtype=np.int64)] for c in cand_idx_list]
    return log_vecs, cands


def shard_round_robin(n_items, rank, world):
    """streaming.py:53-54: worker r takes items r, r+world, ..."""
    return list(range(rank, n_items, world))
