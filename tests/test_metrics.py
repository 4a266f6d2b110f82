"""Device AUROC / AUPR (dreamgnn_b200/metrics.py) against sklearn, the reference's metric code (evaluation.py:58-65)."""
import numpy as np
import pytest
import torch as th

from dreamgnn_b200 import metrics as M


def test_descending_keys_order_scores():
    """CPU: the uint32 sort key is strictly decreasing in the score, and +0.0 / -0.0 share a key (sklearn compares
    values)."""
    s = th.tensor([float('-inf'), -3.5, -1e-30, -0.0, 0.0, 1e-30, 2.0, 2.0000002, float('inf')], dtype=th.float32)
    k = M._descending_keys(s)
    assert k[3] == k[4]
    d = k[1:] - k[:-1]
    assert (d[[0, 1, 2, 4, 5, 6, 7]] < 0).all() and d[3] == 0
    assert int(k.min()) >= 0 and int(k.max()) <= 0xffffffff


def _sk(y, s):
    from sklearn import metrics
    fpr, tpr, _ = metrics.roc_curve(y, s)
    p, r, _ = metrics.precision_recall_curve(y, s)
    return metrics.auc(fpr, tpr), metrics.auc(r, p)


@pytest.mark.gpu
@pytest.mark.parametrize('n,decimals,pos', [(2, None, 0.5), (17, None, 0.3), (1000, 1, 0.1), (4096, 0, 0.5), (52000, None, 0.006),
                                           (520000, 2, 0.006)])
def test_binary_curve_areas_match_sklearn(n, decimals, pos):
    rng = np.random.default_rng(n)
    y = (rng.random(n) < pos).astype(np.int64)
    y[0], y[1] = 1, 0
    s = (rng.normal(size=n) + 1.5 * y).astype(np.float32)
    if decimals is not None:
        s = np.round(s, decimals).astype(np.float32)          # heavy ties, including +0.0 / -0.0
    dev = th.device('cuda:0')
    a, p = M.binary_curve_areas(th.tensor(y, device=dev), th.tensor(s, device=dev))
    ra, rp = _sk(y, s)
    assert abs(a - ra) <= 1e-12 and abs(p - rp) <= 1e-12
    # float labels, as the training loop holds them
    a2, p2 = M.binary_curve_areas(th.tensor(y, device=dev, dtype=th.float32), th.tensor(s, device=dev))
    assert (a2, p2) == (a, p)


@pytest.mark.gpu
def test_binary_curve_areas_single_class_is_nan():
    dev = th.device('cuda:0')
    s = th.randn(100, device=dev)
    a, p = M.binary_curve_areas(th.zeros(100, device=dev), s)
    assert np.isnan(a) and np.isnan(p)
    a, p = M.binary_curve_areas(th.ones(100, device=dev), s)
    assert np.isnan(a) and p == 1.0                           # sklearn: roc undefined, PR curve is constant 1
    with pytest.raises(ValueError):
        M.binary_curve_areas(th.ones(0, device=dev), th.ones(0, device=dev))
    with pytest.raises(RuntimeError):
        M.binary_curve_areas(th.ones(3), th.ones(3))          # no CPU fallback
