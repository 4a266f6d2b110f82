"""AUROC / AUPR on the device: what `sklearn.metrics.roc_curve` + `precision_recall_curve` + `auc`
compute in the reference's `evaluate` (evaluation.py:58-65), restated as a stable radix sort of the logits
(dg_sort_pairs_u64, our own kernel), a running count of positives and two trapezoid sums in float64.

sklearn's algorithm (`_binary_clf_curve`): sort by score descending (stable), keep one operating point per DISTINCT
score (ties collapse into a single point: the last index of each run), tps = cumulative positives, fps = rank - tps;
ROC prepends (0, 0); the PR curve starts at (recall 0, precision 1). `roc_curve`'s drop_intermediate only removes
collinear points, so the areas are those of the full curves. ~520 k scores on lrssl: the CPU round trip (D2H copy +
two sklearn curves, 1.2 s [probe]) becomes a ~30-launch device computation with one scalar read-back.
"""
import torch as th

from . import ops


def _descending_keys(score):
    """Monotone map float32 -> uint32 (stored in int64) under which ascending key order = descending score."""
    s = score.to(th.float32) + 0.0                               # -0.0 -> +0.0: sklearn compares values, not bit patterns
    bits = s.contiguous().view(th.int32).to(th.int64) & 0xffffffff
    neg = bits >= 0x80000000
    asc = th.where(neg, 0xffffffff - bits, bits + 0x80000000)   # ascending in score
    return 0xffffffff - asc


def binary_curve_areas(y_true, y_score):
    """(auroc, aupr) as Python floats, equal to sklearn's `auc(*roc_curve(y, s)[:2])` and
    `auc(recall, precision)` of `precision_recall_curve(y, s)`; nan when one class is absent (sklearn warns and
    returns nan there too). y_true: {0,1} labels (any numeric dtype), y_score: logits; both 1-D CUDA tensors."""
    if not y_score.is_cuda:
        raise RuntimeError('dreamgnn_b200.metrics needs CUDA tensors (no CPU fallback)')
    n = y_score.numel()
    if n == 0 or y_true.numel() != n:
        raise ValueError('binary_curve_areas: empty input or length mismatch')
    keys = _descending_keys(y_score.reshape(-1))
    labels = (y_true.reshape(-1) > 0).to(th.int32)
    keys, labels = ops.sort_pairs_u64(keys.clone(), labels.contiguous(), 32)      # stable: ties keep input order
    tps_all = th.cumsum(labels, 0, dtype=th.int64)
    last_of_run = th.ones(n, dtype=th.bool, device=keys.device)
    last_of_run[:-1] = keys[1:] != keys[:-1]
    idx = th.nonzero(last_of_run).reshape(-1)                    # operating points (one host sync for the count)
    tps = tps_all[idx].to(th.float64)
    fps = (idx + 1).to(th.float64) - tps
    p_tot, n_tot = tps[-1], fps[-1]
    zero = th.zeros(1, dtype=th.float64, device=keys.device)
    one = th.ones(1, dtype=th.float64, device=keys.device)
    tpr, fpr = th.cat([zero, tps / p_tot]), th.cat([zero, fps / n_tot])
    auroc = th.trapezoid(tpr, fpr)
    precision, recall = th.cat([one, tps / (tps + fps)]), th.cat([zero, tps / p_tot])
    aupr = th.trapezoid(precision, recall)
    auroc, aupr, p_tot, n_tot = th.stack([auroc, aupr, p_tot, n_tot]).tolist()      # the one D2H read
    nan = float('nan')
    return (auroc if p_tot > 0 and n_tot > 0 else nan), (aupr if p_tot > 0 else nan)
