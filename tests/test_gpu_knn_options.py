"""GPU: the kNN builder's option set beyond the training default -- `symm=False` (data_loader.py:302 skipped),
`utils.knn_graph` (utils.py:106-140) and `augmented_knn_graph` (augmentation.py:341-399; both dead code in the reference) --
against oracle/restate.py; index sets bit-exact, fp32 values bit-exact, float64 values to 1e-12."""
import numpy as np
import pytest
import torch as th

from oracle import restate as R

pytestmark = pytest.mark.gpu


@pytest.fixture(scope='module')
def dev():
    from dreamgnn_b200 import _lib
    _lib.load()
    return th.device('cuda:0')


def _coo(t):
    t = t.coalesce()
    return t.indices()[0].cpu().numpy(), t.indices()[1].cpu().numpy(), t.values().cpu().numpy()


@pytest.mark.parametrize('n,k', [(40, 3), (257, 15), (9, 20)])
def test_directed_knn_graph(dev, n, k):
    from dreamgnn_b200 import graph_build as GB
    rng = np.random.default_rng(n)
    x = rng.standard_normal((n, 8))
    sim = x @ x.T
    t = GB.create_similarity_graph(sim, k, dev, symm=False)
    row, col, val = R.similarity_knn_graph(sim, k, symm=False)
    gr, gc, gv = _coo(t)
    np.testing.assert_array_equal(gr, row)
    np.testing.assert_array_equal(gc, col)
    np.testing.assert_array_equal(gv, val)
    f = GB.create_feature_similarity_graph(x, k, dev, symm=False)
    row, col, val = R.similarity_knn_graph(R.feature_cosine_similarity(x), k, symm=False)
    gr, gc, gv = _coo(f)
    np.testing.assert_array_equal(gr, row)
    np.testing.assert_array_equal(gc, col)
    np.testing.assert_array_equal(gv, val)


@pytest.mark.parametrize('n,k', [(40, 0), (40, 4), (300, 15)])
def test_binary_and_augmented_knn_graph(dev, n, k):
    from dreamgnn_b200 import graph_build as GB
    rng = np.random.default_rng(n + k)
    x = rng.standard_normal((n, 6))
    sim = x @ x.T
    row, col, val = R.knn_graph_binary(sim, k)
    gr, gc, gv = _coo(GB.knn_graph(sim, k, dev))
    np.testing.assert_array_equal(gr, row)
    np.testing.assert_array_equal(gc, col)
    np.testing.assert_array_equal(gv, val)
    if k == 0:
        return
    noise = rng.standard_normal(len(val))
    keep = rng.permutation(len(val))[:max(1, int(len(val) * 0.8))]
    row, col, want = R.augmented_knn_graph(sim, k, keep=keep, noise=noise, noise_scale=0.1)
    gr, gc, gv = _coo(GB.augmented_knn_graph(sim, k, dev, noise=noise, keep=keep, noise_scale=0.1))
    np.testing.assert_array_equal(gr, row)
    np.testing.assert_array_equal(gc, col)
    np.testing.assert_allclose(gv, want, rtol=1e-12, atol=0)
    own = GB.augmented_knn_graph(sim, k, dev, dropout_rate=0.2, add_noise=True).coalesce()        # own draws: structure checks
    d = own.to_dense()
    assert th.equal(d, d.t()) and float(d.diagonal().min()) >= 1.0
