"""CPU: pin oracle/restate.py (and the test-side data preparation of tests/shapes.py) against digests of the UNMODIFIED
reference at the full BASELINE.json dataset shapes (lrssl 763 x 681, Gdataset 593 x 313, Cdataset 663 x 409; CLI-default
model). Digests come from tests/golden/make_golden_shapes.py. Runs everywhere, including the GPU box."""
import numpy as np
import pytest
import torch as th

from oracle import restate as R
from tests import helpers as H
from tests import shapes as S


@pytest.fixture(scope='module', params=list(S.DATASETS))
def shape(request):
    name = request.param
    ds = S.dataset(name)
    enc, knn = S.oracle_graphs(ds)
    return name, ds, enc, knn, S.load_shape_golden(name)


def test_inputs_bit_exact(shape):
    """Fold split, encoder edge lists, ci / cj, the four kNN graphs, the normalised features: SHA-256 equal."""
    _, ds, enc, knn, g = shape
    for split in ('train', 'test'):
        (rows, cols), vals = ds['split'][split]
        assert S.sha(np.stack([rows, cols]).astype(np.int64)) == str(g[f'hash.{split}.pairs'])
        assert S.sha(vals.astype(np.float32)) == str(g[f'hash.{split}.labels'])
        assert rows.size == int(g[f'meta.{split}.n_pairs'])
        for et in ('0', '1', 'rev-0', 'rev-1'):
            assert S.sha(np.stack(enc[split]['edges'][et]).astype(np.int64)) == str(g[f'hash.{split}.enc.{et}']), et
        for nt in ('drug', 'disease'):
            assert S.sha(enc[split]['ci'][nt]) == str(g[f'hash.{split}.ci.{nt}'])
            assert S.sha(enc[split]['cj'][nt]) == str(g[f'hash.{split}.cj.{nt}'])
    for gk, (row, col, val, _) in zip(('drug_graph', 'disease_graph', 'drug_feature_graph', 'disease_feature_graph'), knn):
        assert val.size == int(g[f'meta.knn.{gk}.nnz']), gk
        assert S.sha(np.stack([row, col]).astype(np.int64)) == str(g[f'hash.knn.{gk}.indices']), gk
        assert S.sha(val.astype(np.float32)) == str(g[f'hash.knn.{gk}.values']), gk
    assert S.sha(ds['drug_feat'].numpy()) == str(g['hash.feat.drug'])
    assert S.sha(ds['dis_feat'].numpy()) == str(g['hash.feat.disease'])


def test_seeded_init_equals_reference(shape):
    _, ds, _, _, g = shape
    assert S.state_dict_hash(S.init_state_dict(ds)) == str(g['hash.sd'])


def test_forward_and_gradients(shape):
    """fp32 restatement vs the reference's fp32 run: eval-mode outputs <= 2e-6, loss, every gradient within the
    per-tensor budget of tests/shapes.py:budget (1e-5, or 1.5 x the reference's own distance from the float64 value)."""
    name, ds, enc, knn, g = shape
    sd = S.init_state_dict(ds)
    with th.no_grad():
        _, out = S.oracle_forward(ds, enc, knn, sd)
    for nm, t in zip(('pred', 'drug_out', 'drug_sim_out', 'dis_out', 'dis_sim_out'), out):
        e_s, e_n = S.digest_errors('fwd.' + nm, t, float(g[f'fwd.{nm}.norm']), g[f'fwd.{nm}.samples'])
        assert e_s <= 2e-6 and e_n <= 2e-6, (nm, e_s, e_n)
    _, loss32, g32 = S.oracle_loss_and_grads(ds, enc, knn, sd, th.float32)
    _, loss64, g64 = S.oracle_loss_and_grads(ds, enc, knn, sd, th.float64)
    assert abs(loss32 - float(g['loss'])) <= 2e-6 and abs(loss64 - float(g['loss'])) <= 2e-6
    for k in (key[8:] for key in g if key.startswith('hasgrad.')):      # named_parameters(): shared ifc listed once (as ufc)
        if not bool(g['hasgrad.' + k]):
            assert k not in g32 or float(g32[k].abs().max()) == 0.0, k
            continue
        gold = (float(g[f'grad.{k}.norm']), g[f'grad.{k}.samples'])
        ref_vs_exact = S.digest_errors('grad.' + k, g64[k].float(), *gold)[0]
        e_s, e_n = S.digest_errors('grad.' + k, g32[k], *gold)
        # two fp32 evaluations that are each within e of the exact value can sit 2e apart
        assert e_s <= S.budget(ref_vs_exact, slack=3.0) and e_n <= S.budget(ref_vs_exact, slack=3.0), (k, e_s, e_n, ref_vs_exact)
        assert H.rel_err(g32[k], g64[k]) <= S.budget(ref_vs_exact), (k, ref_vs_exact)


def test_evaluate_auc(shape):
    """evaluation.py:4-74 at the reference's weights: AUROC / AUPR within 1e-3 (north star) on both splits."""
    _, ds, enc, knn, g = shape
    sd = S.init_state_dict(ds)
    feats = (ds['drug_feat'], ds['dis_feat'], ds['drug_sim'], ds['dis_sim'])
    for split in ('train', 'test'):
        pairs, labels = ds['split'][split]
        auroc, aupr = R.evaluate_auc(S.oracle_params(sd), enc[split], pairs, labels, knn, feats, dict(layers=3))
        assert abs(auroc - float(g[f'eval.{split}.auroc'])) <= 1e-3, (split, auroc)
        assert abs(aupr - float(g[f'eval.{split}.aupr'])) <= 1e-3, (split, aupr)
