"""CPU (no GPU needed): the C-ABI library loads and exports every symbol include/dreamgnn.h declares,
the host-side mirrors keep the reference's parameter names / initialisation, and the product refuses
to run without CUDA instead of falling back."""
import argparse
import os
import re
import subprocess

import numpy as np
import pytest
import torch as th

import dreamgnn_b200
from dreamgnn_b200 import _lib, graph as G, layers, ops
from dreamgnn_b200.model import Net
from tests import helpers as H

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope='module')
def built():
    from dreamgnn_b200 import build
    return build.build()


def test_header_symbols_exported(built):
    hdr = open(os.path.join(REPO, 'include', 'dreamgnn.h')).read()
    declared = set(re.findall(r'DG_API[^;(]*?\b(dg_[a-z0-9_]+)\s*\(', hdr))
    assert len(declared) >= 20
    out = subprocess.run(['nm', '-D', '--defined-only', built], capture_output=True, text=True, check=True).stdout
    exported = set(re.findall(r' T (dg_[a-z0-9_]+)', out))
    assert declared == exported, (declared ^ exported)
    assert declared == set(_lib.EXPORTED_SYMBOLS)


def test_library_loads_and_reports_version(built):
    lib = _lib.load()
    assert lib.dg_abi_version() == _lib.ABI_VERSION
    assert lib.dg_scan_workspace_bytes(1000) > 0
    assert lib.dg_csr_build_workspace_bytes(1000, 10) > lib.dg_csr_build_workspace_bytes(10, 10)
    assert isinstance(_lib.launch_count(), int)


def test_no_cpu_fallback():
    x = th.randn(4, 8)
    csr = ops.CSR(th.tensor([0, 1, 2, 3, 4], dtype=th.int32), th.tensor([0, 1, 2, 3], dtype=th.int32),
                  th.arange(4, dtype=th.int32), None, 4, 4)
    with pytest.raises(RuntimeError, match='CUDA'):
        ops.spmm(csr, x)
    with pytest.raises(RuntimeError, match='CUDA'):
        ops.decoder_mlp(th.randn(2, 128), th.randn(2, 128), th.randn(64, 128), th.randn(64), th.randn(1, 64),
                        th.randn(1), None)
    with pytest.raises(RuntimeError):
        layers.adjacency_csr(th.eye(3).to_sparse())
    with pytest.raises(RuntimeError, match='CUDA'):
        ops.bce_with_logits(th.zeros(3), th.zeros(3))
    from dreamgnn_b200.optim import FusedAdam
    p = th.zeros(4, requires_grad=True)
    p.grad = th.ones(4)
    with pytest.raises(RuntimeError, match='CUDA'):
        FusedAdam([p]).step()
    assert th.equal(p.detach(), th.zeros(4))


def test_fused_adam_keeps_torch_adams_contract():
    """Same param_groups / defaults / state_dict layout as torch.optim.Adam (checkpoints interchange); options the
    kernel does not implement are refused at construction; the seed pool is a CUDA-only shortcut."""
    from dreamgnn_b200.optim import FusedAdam
    p = th.zeros(4, requires_grad=True)
    opt = FusedAdam([p], lr=0.002, weight_decay=1e-5)
    ref = th.optim.Adam([p], lr=0.002, weight_decay=1e-5)
    g, r = opt.param_groups[0], ref.param_groups[0]
    assert {k: g[k] for k in ('lr', 'betas', 'eps', 'weight_decay')} == {k: r[k] for k in ('lr', 'betas', 'eps', 'weight_decay')}
    assert g['capturable'] is True and isinstance(opt, th.optim.Adam)
    assert set(opt.state_dict()) == set(ref.state_dict())
    with pytest.raises(ValueError, match='amsgrad'):
        FusedAdam([p], amsgrad=True)
    ops.begin_seed_pool('cpu')                                        # no-op off CUDA
    assert ops.fresh_seed(th.device('cpu')).shape == (1,)


def _args(g, name):
    cfg = {'tinyA': dict(layers=3, gcn_agg_units=105, gcn_out_units=16, nhid1=40, nhid2=16),
           'tinyB': dict(layers=2, gcn_agg_units=96, gcn_out_units=8, nhid1=20, nhid2=8)}[name]
    return argparse.Namespace(model_activation='leaky', gcn_agg_accum='sum', share_param=True, device='cpu',
                              dropout=0.0, attention_dropout=0.0, rating_vals=[0, 1],
                              src_in_units=g['feat.drug'].shape[1], dst_in_units=g['feat.disease'].shape[1],
                              fdim_drug=g['feat.drug'].shape[0], fdim_disease=g['feat.disease'].shape[0], **cfg)


@pytest.mark.parametrize('name', H.CASES)
def test_net_state_dict_and_init_match_reference(name):
    """Same keys, shapes AND values as the reference's Net built from the same seed (make_golden.py
    used th.manual_seed(2024)): construction order and initialisers are mirrored exactly."""
    g = H.load_golden(name)
    th.manual_seed(2024)
    net = Net(_args(g, name))
    sd = net.state_dict()
    want = {k[3:]: v for k, v in g.items() if k.startswith('sd.')}
    assert set(sd) == set(want)
    for k, v in want.items():
        np.testing.assert_array_equal(sd[k].numpy(), v, err_msg=k)
    net.load_state_dict({k: th.tensor(v) for k, v in want.items()})   # checkpoints interchange


def test_graph_handle_structure_on_cpu():
    g = H.load_golden('tinyA')
    data = {}
    for et in ('0', '1'):
        s, d = g[f'train.enc.{et}']
        data[('drug', et, 'disease')] = (s, d)
        data[('disease', 'rev-' + et, 'drug')] = (d, s)
    hg = G.heterograph(data, num_nodes_dict={'drug': 60, 'disease': 45})
    assert hg.canonical_etypes == [('disease', 'rev-0', 'drug'), ('disease', 'rev-1', 'drug'),
                                   ('drug', '0', 'disease'), ('drug', '1', 'disease')]
    assert hg.etypes == ['rev-0', 'rev-1', '0', '1'] and hg.ntypes == ['disease', 'drug']
    assert hg.number_of_edges('0') == g['train.enc.0'].shape[1]
    rel = hg['0']
    assert rel.number_of_src_nodes() == 60 and rel.number_of_dst_nodes() == 45
    np.testing.assert_array_equal(rel.in_degrees().numpy(), np.bincount(g['train.enc.0'][1], minlength=45))
    hg.nodes['drug'].data['ci'] = th.ones(60, 1)
    assert rel.srcdata['ci'] is hg.nodes['drug'].data['ci']           # slices share node data
    with rel.local_scope():
        rel.srcdata['h'] = th.zeros(60, 3)
    assert 'h' not in hg.nodes['drug'].data
    i32 = hg.int()
    assert i32.idtype == th.int32 and hg.int() is i32                  # conversions are cached
    with pytest.raises(RuntimeError, match='CUDA'):
        hg.block('drug')
    c = hg.clone()
    c.add_edges([0], [0], etype='0')
    assert c.number_of_edges('0') == hg.number_of_edges('0') + 1


def test_package_has_no_oracle_dependency():
    pkg = os.path.dirname(dreamgnn_b200.__file__)
    for root, _, files in os.walk(pkg):
        for f in files:
            if f.endswith(('.py', '.cu', '.cuh')):
                src = open(os.path.join(root, f)).read()
                assert 'oracle' not in src.replace('oracle/', '').lower() or f == 'none', f
                assert 'import dgl' not in src, f


def test_new_ops_refuse_cpu_and_workspaces_are_host_functions(built):
    lib = _lib.load()
    assert lib.dg_colsum_workspace_bytes(100000, 128) >= 128 * 4
    assert lib.dg_colsum_workspace_bytes(0, 128) > 0
    assert lib.dg_random_subset_workspace_bytes() >= 256 * 4 + 16
    with pytest.raises(RuntimeError, match='CUDA'):
        ops.colsum(th.randn(8, 4))
    with pytest.raises(RuntimeError, match='CUDA'):
        ops.gram_common_loss(th.randn(8, 4), th.randn(8, 4))
    with pytest.raises(RuntimeError, match='CUDA'):
        ops.linear(th.randn(8, 4), th.randn(4, 4), th.randn(4))
    with pytest.raises(RuntimeError):
        ops.random_subset_flags(4, 2, th.zeros(4, dtype=th.uint8), rnd=th.zeros(4, dtype=th.int64))


def test_branches_are_plain_calls_unless_enabled():
    """ops.branches: off by default (results in order, same stream); the context manager scopes the switch."""
    calls = []
    out = ops.branches([lambda: calls.append('a') or 1, lambda: calls.append('b') or 2, lambda: calls.append('c') or 3])
    assert out == [1, 2, 3] and calls == ['a', 'b', 'c'] and ops.PARALLEL_BRANCHES is False
    with ops.parallel_branches(True):
        assert ops.PARALLEL_BRANCHES is True
        with ops.parallel_branches(False):                 # False leaves the current setting alone
            assert ops.PARALLEL_BRANCHES is True
    assert ops.PARALLEL_BRANCHES is False
    assert list(ops._each_tensor((th.zeros(1), [th.ones(1), {'k': th.ones(2)}], 3))).__len__() == 3


def test_common_loss_cpu_path_is_the_reference_expression():
    """On CPU tensors utils.common_loss is utils.py:87-95 literally (the oracle compares against it); the Gram identity
    it is replaced by on CUDA gives the same value."""
    from dreamgnn_b200.utils import common_loss, common_loss_dense, common_loss_gram_torch
    gen = th.Generator().manual_seed(0)
    a, b = th.randn(50, 8, generator=gen, dtype=th.float64), th.randn(50, 8, generator=gen, dtype=th.float64)
    assert th.equal(common_loss(a, b), common_loss_dense(a, b))
    assert abs(float(common_loss_gram_torch(a.float(), b.float())) - float(common_loss_dense(a, b))) <= 1e-6 * float(common_loss_dense(a, b))


def test_edge_sampler_choice_on_cpu(monkeypatch):
    from dreamgnn_b200.augmentation import _randperm, num_keep_edges
    p = _randperm(10, 'cpu')
    assert isinstance(p, th.Tensor) and sorted(p.tolist()) == list(range(10))
    monkeypatch.setenv('DG_EDGE_SAMPLER', 'select')
    assert isinstance(_randperm(10, 'cpu'), th.Tensor)        # the select sampler is a CUDA kernel: CPU graphs keep randperm
    assert num_keep_edges(467641, 0.1) == int(467641 * 0.9) and num_keep_edges(1, 0.99) == 1


def test_bench_side_files_parse():
    """profiles/roofline_traffic.json and profiles/l2_peak.json feed bench.py's roofline block."""
    import json
    tr = json.load(open(os.path.join(REPO, 'profiles', 'roofline_traffic.json')))
    assert 'kernel_digest' in tr and all(isinstance(tr[k], int) for k in tr if k.endswith(('d128', 'd344', 'd768')))
    l2 = json.load(open(os.path.join(REPO, 'profiles', 'l2_peak.json')))
    assert l2['l2_gather_d344']['GBps'] > 1000 and l2['hbm_seq']['GBps'] > 1000


def test_arena_gives_twin_layouts_and_one_copy_clone():
    """graphed._Arena: two arenas filled by the same sequence of requests have identical layouts (256-byte aligned), so a
    tree of tensors is cloned into its twin by ONE buffer copy; dense entries of a staged augmentation go through it."""
    from dreamgnn_b200.graphed import StagedAugmentation, _Arena
    ts = [th.arange(7, dtype=th.int32), th.randn(3, 5), th.zeros(0, 4), th.arange(6, dtype=th.int64).view(2, 3)]
    a = _Arena(sum(_Arena.span(t) for t in ts), 'cpu')
    held = [a.take_like(t) for t in ts]
    assert all(th.equal(h, t) and h.dtype == t.dtype and h.shape == t.shape for h, t in zip(held, ts))
    assert all(h.data_ptr() % 4 == 0 and (h.data_ptr() - a.buf.data_ptr()) % 256 == 0 for h in held if h.numel())
    b = _Arena(a.buf.numel(), 'cpu')
    views = [b.view_like(t) for t in ts]
    b.buf.copy_(a.buf)
    assert all(th.equal(v, t) for v, t in zip(views, ts))
    with pytest.raises(ValueError):
        b.view_like(th.zeros(1))
    base = {'drug_feat': th.randn(4, 3), 'disease_feat': th.randn(5, 3), 'enc_graph': None}
    aug = {'drug_feat': base['drug_feat'] + 1.0, 'disease_feat': base['disease_feat'], 'enc_graph': None}
    st = StagedAugmentation(aug, base)
    assert st.keys == ['drug_feat'] and st.passthrough['disease_feat'] is base['disease_feat']
    live = st.clone()
    assert th.equal(live['drug_feat'], aug['drug_feat']) and live['drug_feat'].data_ptr() != st.tree['drug_feat'].data_ptr()
    st.refresh({'drug_feat': aug['drug_feat'] * 2.0})
    assert th.equal(st.tree['drug_feat'], aug['drug_feat'] * 2.0) and th.equal(live['drug_feat'], aug['drug_feat'])
    with pytest.raises(ValueError, match='changed shape'):
        st.refresh({'drug_feat': th.zeros(2, 2)})


def test_weighted_loss_sum_on_cpu():
    xs = [th.tensor(v, requires_grad=True) for v in (0.7, 2.5, -1.25)]
    total = ops.WeightedLossSum.apply(xs[0], xs[1], xs[2], 0.001)
    assert abs(float(total) - (0.7 + 0.001 * (2.5 - 1.25))) < 1e-7
    (total * 2.0).backward()
    assert [round(float(x.grad), 7) for x in xs] == [2.0, 0.002, 0.002]


@pytest.mark.parametrize('T', [1, 4, 7, 64])
def test_split_plan_tiles_every_row_in_order(T):
    """ops.split_plan (chunked aggregation of skewed CSRs): the virtual rows tile each real row exactly, in edge order,
    in chunks of <= T; the unused tail of the static bound is empty; chunk sums then row sums equal the direct sums."""
    rng = np.random.default_rng(T)
    deg = np.concatenate([[0, 0, 1, T, T + 1, 5 * T, 37 * T + 3], rng.integers(0, 3 * T + 2, 40), [0]])
    rng.shuffle(deg)
    indptr = th.tensor(np.concatenate([[0], np.cumsum(deg)]), dtype=th.int32)
    n_rows, nnz = len(deg), int(deg.sum())
    v_indptr, comb, n_v = ops.split_plan(indptr, n_rows, nnz, T)
    assert n_v == n_rows + nnz // T and v_indptr.dtype == comb.dtype == th.int32
    assert v_indptr.numel() == n_v + 1 and comb.numel() == n_rows + 1
    v, c = v_indptr.numpy().astype(np.int64), comb.numpy().astype(np.int64)
    assert v[0] == 0 and v[-1] == nnz and (np.diff(v) >= 0).all() and (np.diff(v) <= T).all()
    assert c[0] == 0 and c[-1] <= n_v and (np.diff(c) == np.maximum(1, -(-deg // T))).all()
    for r in range(n_rows):
        assert v[c[r]] == indptr[r] and v[c[r + 1]] == indptr[r + 1]           # row r = its chunks, contiguous
        assert (np.diff(v[c[r]:c[r + 1] + 1])[:-1] == T).all()                   # all but the last chunk are full
    assert (v[c[-1]:] == nnz).all()                                              # tail: empty rows
    # two-pass sums == direct sums (integers: exact)
    w = th.tensor(rng.integers(-5, 6, nnz), dtype=th.float64)
    seg = lambda ptr, x: th.stack([x[int(ptr[i]):int(ptr[i + 1])].sum() for i in range(len(ptr) - 1)]) if len(ptr) > 1 else x[:0]
    assert th.equal(seg(c, seg(v, w)), seg(indptr.numpy(), w))


def test_split_decision_is_inherited_by_derived_structures():
    base = ops.CSR(th.tensor([0, 2]), th.tensor([0, 1]), th.tensor([0, 1]), None, 1, 2)
    assert base.split_T == 0 and base._plan is None
    base.split_T = 512
    assert ops.inherit_layout(ops.CSR(base.indptr, base.indices, base.eid, None, 1, 2), base).split_T == 512
    from dreamgnn_b200 import graphed
    assert graphed._clone_csr(base, lambda t: t.clone()).split_T == 512


def test_chunked_spmm_composition_with_a_stand_in_kernel(monkeypatch):
    """Host logic of the skewed-graph SpMM path (ops._spmm_raw with csr.split_T): with the C entry point replaced by a
    torch stand-in of its contract (out = relu?(ds * (A (ss * x)) + bias)), the two launches -- chunk partials, then per-row
    sums with the epilogue -- compose to exactly what one launch gives. (The kernel itself is covered by the GPU tests.)"""
    calls = []

    class FakeLib:
        @staticmethod
        def dg_spmm_csr_f32(indptr, indices, vals, ss, ds, bias, x, x_stride, out, out_stride, n_rows, d, flags, stream):
            calls.append((int(n_rows), int(flags)))
            assert indptr.numel() == n_rows + 1 and out.shape == (n_rows, d)
            rows = th.repeat_interleave(th.arange(n_rows), (indptr[1:] - indptr[:-1]).long())
            nnz = int(indptr[-1])
            g = x.double()[indices[:nnz].long()] * (1.0 if ss is None else ss.double()[indices[:nnz].long()][:, None])
            if vals is not None:
                g = g * vals.double()[:nnz, None]
            acc = th.zeros(n_rows, d, dtype=th.float64).index_add_(0, rows, g)
            if ds is not None:
                acc = acc * ds.double()[:, None]
            if bias is not None:
                acc = acc + bias.double()
            out.copy_(acc.relu() if flags & ops.SPMM_RELU else acc)
            return 0

    class FakeL:
        load = staticmethod(lambda: FakeLib)
        ptr = staticmethod(lambda t, *a: t)
        check = staticmethod(lambda rc, what: None)
        stream = staticmethod(lambda: 0)
    monkeypatch.setattr(ops, 'L', FakeL)
    rng = np.random.default_rng(0)
    deg = np.concatenate([[0, 300, 1, 0, 77], rng.integers(0, 9, 30)])
    indptr = th.tensor(np.concatenate([[0], np.cumsum(deg)]), dtype=th.int32)
    nnz, n_rows, n_cols, d = int(deg.sum()), len(deg), 50, 8
    g = th.Generator().manual_seed(0)
    indices = th.randint(0, n_cols, (nnz,), generator=g, dtype=th.int32)
    vals = th.randint(1, 4, (nnz,), generator=g).float()
    x = th.randint(-3, 4, (n_cols, d), generator=g).float()                 # small integers: every sum is exact
    ss, ds = th.randint(1, 3, (n_cols,), generator=g).float(), th.randint(1, 3, (n_rows,), generator=g).float()
    bias = th.randint(-2, 3, (d,), generator=g).float()
    csr = ops.CSR(indptr, indices, th.arange(nnz, dtype=th.int32), vals, n_rows, n_cols)
    one = ops._spmm_raw(csr, x, ss, ds, bias, ops.SPMM_RELU)
    assert len(calls) == 1
    csr.split_T = 16
    two = ops._spmm_raw(csr, x, ss, ds, bias, ops.SPMM_RELU)
    assert [c[0] for c in calls[1:]] == [n_rows + nnz // 16, n_rows] and calls[1][1] & ops.SPMM_RELU == 0
    assert th.equal(one, two) and csr._plan is not None
    assert th.equal(ops._spmm_raw(csr, x, ss, ds, bias, ops.SPMM_RELU), one) and len(calls) == 5       # cached plan


def test_zipf_cells_stress_set_generator():
    """synthetic.zipf_cells (secondary stress set): distinct in-range cells, seeded, with the popular
    drugs saturated at every disease and a long tail of short rows."""
    from dreamgnn_b200 import synthetic
    n_d, n_s = 2000, 300
    draw = lambda seed: synthetic.zipf_cells(n_d, n_s, 400_000, th.Generator().manual_seed(seed), 'cpu')
    cells = draw(5)
    assert th.equal(cells, draw(5)) and not th.equal(cells, draw(6))
    assert th.equal(cells, th.unique(cells)) and 0 <= int(cells.min()) and int(cells.max()) < n_d * n_s
    deg = th.bincount(cells // n_s, minlength=n_d)
    assert int(deg.max()) == n_s and int((deg == n_s).sum()) >= 5          # the head saturates
    assert float(deg.float().median()) < 0.2 * n_s                         # the tail is short
    spec = dict(synthetic.SHAPES['syn20m'], pair_dist='zipf')
    assert synthetic.scaled(spec, 0.5)['pair_dist'] == 'zipf'              # the option survives scaling
