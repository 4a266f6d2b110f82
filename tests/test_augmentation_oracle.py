"""a13: the perturbation methods of augmentation.py:127-337 (`add_random_edges`, `graph_noise`, `feature_masking`,
`mix_up`) against oracle/restate.py with the random draws injected on both sides, and -- where /root/reference or the
staged archive exists -- the restatements against the reference's own functions under a fixed seed.

`add_random_edges`, `feature_masking` and `mix_up` are device-agnostic tensor code and run here on CPU tensors as well;
`graph_noise` needs the CSR sidecar kernels (gpu)."""
import random

import numpy as np
import pytest
import torch as th

from oracle import ref_runner as rr
from oracle import restate as R


def _graph(rng, n_d, n_s, n_edges, device='cpu'):
    from dreamgnn_b200.graph import heterograph
    cells = rng.choice(n_d * n_s, size=n_edges, replace=False)
    lab = rng.random(n_edges) < 0.3
    d, s = cells // n_s, cells % n_s
    data = {('drug', '0', 'disease'): (d[~lab], s[~lab]), ('disease', 'rev-0', 'drug'): (s[~lab], d[~lab]),
            ('drug', '1', 'disease'): (d[lab], s[lab]), ('disease', 'rev-1', 'drug'): (s[lab], d[lab])}
    g = heterograph({k: (th.tensor(a), th.tensor(b)) for k, (a, b) in data.items()}, {'drug': n_d, 'disease': n_s})
    return g.to(device) if device != 'cpu' else g, data


@pytest.mark.parametrize('n_d,n_s,n_edges,rate', [(30, 20, 200, 0.05), (12, 9, 100, 0.3), (50, 40, 37, 0.03), (8, 8, 60, 0.5)])
def test_add_random_edges_matches_sequential_walk(n_d, n_s, n_edges, rate):
    """Dense relations (most draws hit existing edges, the attempt cap binds), repeated candidates, tiny relations."""
    from dreamgnn_b200.augmentation import GraphAugmentation as GA
    rng = np.random.default_rng(n_edges)
    g, data = _graph(rng, n_d, n_s, n_edges)
    cand = {}
    for c, (a, b) in data.items():
        m = max(1, int(len(a) * rate)) * 10
        ns, nd = (n_d, n_s) if c[0] == 'drug' else (n_s, n_d)
        cand[c] = (rng.integers(0, ns, size=m), rng.integers(0, nd, size=m))
    out = GA.add_random_edges(g, rate, candidates=cand)
    for c, (a, b) in data.items():
        ns, nd = (n_d, n_s) if c[0] == 'drug' else (n_s, n_d)
        add_s, add_d = R.add_random_edges(a, b, ns, nd, rate, cand[c])
        es, ed = out.edges(etype=c)
        np.testing.assert_array_equal(es.numpy(), np.concatenate([a, add_s]))      # appended after the existing edges, draw order
        np.testing.assert_array_equal(ed.numpy(), np.concatenate([b, add_d]))
        assert g.number_of_edges(c) == len(a)                                      # the input graph is untouched


def test_add_random_edges_own_draws_are_new_and_distinct():
    from dreamgnn_b200.augmentation import GraphAugmentation as GA
    g, data = _graph(np.random.default_rng(3), 40, 30, 300)
    out = GA.add_random_edges(g, 0.1)
    for c, (a, b) in data.items():
        es, ed = out.edges(etype=c)
        n_dst = 30 if c[0] == 'drug' else 40
        keys = es.numpy() * n_dst + ed.numpy()
        assert len(np.unique(keys)) == len(keys) and len(keys) == len(a) + max(1, int(len(a) * 0.1))


def test_feature_masking_and_mix_up_match_oracle():
    from dreamgnn_b200.augmentation import GraphAugmentation as GA
    gen = th.Generator().manual_seed(4)
    x = th.randn(37, 24, generator=gen)
    u = th.rand(37, 24, generator=gen)
    assert th.equal(GA.feature_masking(x, 0.1, u=u), R.feature_masking(x, u, 0.1))
    idx = th.randperm(37, generator=gen)
    assert th.equal(GA.mix_up_features(x, 0.2, indices=idx, lam=0.37), R.mix_up_features(x, idx, 0.37))
    noise = th.randn(37, 24, generator=gen)
    np.testing.assert_allclose(GA.feature_noise(x, 0.05, noise=noise).numpy(), (x + noise * 0.05).numpy(), rtol=3e-7, atol=1e-7)      # fused multiply-add: <= 1 ulp


@pytest.mark.skipif(not rr.reference_available(), reason='no reference tree and no staged archive (oracle/_ref)')
def test_restatements_match_reference_under_seed():
    """oracle/restate.py's perturbation restatements vs the reference's own statics: same generator state -> same output
    (add_random_edges: Python's `random` stream replayed as the injected candidates)."""
    mods = rr.import_reference()
    GA = mods['augmentation'].GraphAugmentation
    import dgl                                           # the stand-in (oracle/dgl), put on sys.path by import_reference
    rng = np.random.default_rng(9)
    cells = rng.choice(25 * 18, size=150, replace=False)
    d, s = th.tensor(cells // 18), th.tensor(cells % 18)
    g = dgl.heterograph({('drug', '0', 'disease'): (d, s), ('disease', 'rev-0', 'drug'): (s, d)},
                        num_nodes_dict={'drug': 25, 'disease': 18})
    random.seed(5)
    out = GA.add_random_edges(g, 0.2)
    random.seed(5)
    for c in g.canonical_etypes:                         # the reference draws relation by relation in canonical order
        ns, nd = g.number_of_nodes(c[0]), g.number_of_nodes(c[2])
        a, b = (x.numpy() for x in g.edges(etype=c))
        num_add = max(1, int(len(a) * 0.2))
        cs, cd, accepted, have = [], [], set(), set(zip(a.tolist(), b.tolist()))
        while len(accepted) < num_add and len(cs) < num_add * 10:      # replay exactly as many draws as the walk consumed
            x, y = random.randint(0, ns - 1), random.randint(0, nd - 1)
            cs.append(x)
            cd.append(y)
            if (x, y) not in have:
                accepted.add((x, y))
        add_s, add_d = R.add_random_edges(a, b, ns, nd, 0.2, (np.array(cs), np.array(cd)))
        es, ed = out.edges(etype=c)
        np.testing.assert_array_equal(es.numpy(), np.concatenate([a, add_s]))
        np.testing.assert_array_equal(ed.numpy(), np.concatenate([b, add_d]))
    x = th.randn(20, 8)
    th.manual_seed(11)
    want = GA.feature_masking(x, 0.25)
    th.manual_seed(11)
    assert th.equal(R.feature_masking(x, th.rand_like(x), 0.25), want)
    adj = th.sparse_coo_tensor(th.tensor([[0, 1, 2, 2], [1, 0, 2, 0]]), th.tensor([0.5, 0.25, 1.0, 0.01]), (3, 3))
    th.manual_seed(12)
    want = GA.sparse_graph_noise(adj, 0.05)
    th.manual_seed(12)
    assert th.equal(R.sparse_graph_noise(adj._values(), th.randn_like(adj._values()), 0.05), want._values())
    th.manual_seed(13)
    np.random.seed(13)
    want = GA.mix_up_features(x, 0.2)
    th.manual_seed(13)
    np.random.seed(13)
    idx = th.randperm(20)
    assert th.equal(R.mix_up_features(x, idx, np.random.beta(0.2, 0.2)), want)


@pytest.mark.gpu
def test_graph_noise_on_device_keeps_both_orientations_consistent():
    """sparse_graph_noise on a base kNN graph and on an edge-dropped one (whose edge ids still name the parent's entries):
    values equal the oracle's for the injected draw, the CSR / transposed CSR carry them to the right slots, and a later
    edge dropout needs no device read (the advisor's capture hazard)."""
    from dreamgnn_b200 import _lib, graph_build as GB
    from dreamgnn_b200.augmentation import GraphAugmentation as GA
    from dreamgnn_b200.layers import adjacency_csr
    _lib.load()
    dev = th.device('cuda:0')
    gen = th.Generator(dev).manual_seed(2)
    sim = th.rand(40, 40, generator=gen, device=dev, dtype=th.float64)
    base = GB.create_similarity_graph(sim + sim.t(), 4, dev)
    for adj in (base, GA.random_edge_dropout_sparse(base, 0.2)):
        noise = th.randn(adj._values().numel(), generator=gen, device=dev)
        out = GA.sparse_graph_noise(adj, 0.05, noise=noise)
        want = R.sparse_graph_noise(adj._values().cpu(), noise.cpu(), 0.05)
        assert th.equal(out._values().cpu(), want)
        assert th.equal(out._indices(), adj._indices())
        dense = out.to_dense()
        csr = adjacency_csr(out)
        x = th.randn(40, 8, generator=gen, device=dev)
        from dreamgnn_b200 import ops
        y = ops.spmm(csr, x)
        np.testing.assert_allclose(y.cpu().numpy(), (dense.double() @ x.double()).cpu().numpy(), rtol=1e-5, atol=1e-6)
        yt = ops.spmm(csr.transpose(), x)
        np.testing.assert_allclose(yt.cpu().numpy(), (dense.double().t() @ x.double()).cpu().numpy(), rtol=1e-5, atol=1e-6)
        again = GA.random_edge_dropout_sparse(out, 0.1)                 # noise -> dropout chain
        assert again._values().numel() == max(1, int(out._values().numel() * 0.9))
        d2 = again.to_dense()
        assert float((d2 - dense * (d2 != 0)).abs().max()) == 0.0
