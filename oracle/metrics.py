"""Oracle (test infrastructure): ranking metrics of the eval loop.

Reference: Tiny-NewsRec/metrics.py:5-23 (dcg/ndcg/mrr over
``np.argsort(score)[::-1]``), ``sklearn.metrics.roc_auc_score`` (metrics.py:1,
third-party, unpinned; restated as the tie-aware Mann-Whitney statistic, which is
what its trapezoid ROC integration equals for binary labels), and the per
impression loop / final reduction of Tiny-NewsRec/run.py:346-379.
"""
import numpy as np


def auc(y_true, y_score):
    """P(score_pos > score_neg) + 0.5 P(==)  == sklearn roc_auc_score (binary)."""
    y_true = np.asarray(y_true)
    s = np.asarray(y_score, dtype=np.float64)
    pos, neg = s[y_true == 1], s[y_true == 0]
    if len(pos) == 0 or len(neg) == 0:
        raise ValueError("AUC undefined for constant labels")
    gt = (pos[:, None] > neg[None, :]).sum()
    eq = (pos[:, None] == neg[None, :]).sum()
    return (gt + 0.5 * eq) / (len(pos) * len(neg))


def _order(y_score):
    return np.argsort(y_score)[::-1]          # metrics.py:6,19


def dcg(y_true, y_score, k=10):
    """metrics.py:5-10."""
    top = np.take(y_true, _order(y_score)[:k])
    return np.sum((2.0 ** top - 1) / np.log2(np.arange(len(top)) + 2))


def ndcg(y_true, y_score, k=10):
    """metrics.py:13-16."""
    return dcg(y_true, y_score, k) / dcg(y_true, y_true, k)


def mrr(y_true, y_score):
    """metrics.py:19-23."""
    y = np.take(y_true, _order(y_score))
    return np.sum(y / (np.arange(len(y)) + 1)) / np.sum(y)


def impression_metrics(label, score):
    """One impression of run.py:346-361 -> (auc, mrr, ndcg5, ndcg10) or None if
    the labels are constant (skipped, run.py:348)."""
    label = np.asarray(label)
    if label.mean() == 0 or label.mean() == 1:
        return None
    return (auc(label, score), mrr(label, score), ndcg(label, score, 5), ndcg(label, score, 10))


def eval_reduce(per_impression, total_count):
    """run.py:372-379: metric sums over *valid* impressions divided by the count
    of *all* impressions (skipped ones included)."""
    sums = np.zeros(4, dtype=np.float64)
    for m in per_impression:
        if m is not None:
            sums += np.asarray(m, dtype=np.float64)
    return sums / float(total_count), sums


def doc_sim(news_scoring, n_pairs, rng):
    """Tiny-NewsRec/run.py:292-299, verbatim semantics: ``rng`` is a ``random.Random`` (the reference uses the
    module-level generator)."""
    total = 0
    for _ in range(n_pairs):
        i = rng.randrange(1, len(news_scoring))
        j = rng.randrange(1, len(news_scoring))
        if i != j:
            total += np.dot(news_scoring[i], news_scoring[j]) / (
                np.linalg.norm(news_scoring[i]) * np.linalg.norm(news_scoring[j]))
    return total / n_pairs
