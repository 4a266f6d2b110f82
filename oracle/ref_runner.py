"""TEST INFRASTRUCTURE ONLY -- run the UNMODIFIED reference (`/root/reference/*.py`) on CPU through
the DGL stand-in in `oracle/dgl`, and write same-schema synthetic `.mat` datasets for it.

Only usable where `/root/reference` exists (the build container). It is used by
`tests/golden/make_golden.py` to freeze golden vectors and by the `not gpu` tests that validate
`oracle/restate.py` against the live reference; nothing that runs on the GPU box imports it.

`.mat` schema follows README.md:76-83 and data_loader.py:110-129: `didr` [N_dis x N_drug],
`drug` [N_drug x N_drug], `disease` [N_dis x N_dis], `drug_embed` [N_drug x F],
`disease_embed` [N_dis x F], `Wrname` [N_drug x 1] object array.
"""
import contextlib
import importlib
import io
import os
import sys

import numpy as np

REFERENCE_DIR = os.environ.get('DREAMGNN_REFERENCE_DIR', '/root/reference')
ORACLE_DIR = os.path.dirname(os.path.abspath(__file__))
STAGED_ARCHIVE = os.path.join(ORACLE_DIR, '_ref', 'dreamgnn_reference.zip')     # oracle/stage_ref.py
_REF_MODULES = ('utils', 'augmentation', 'data_loader', 'layers', 'model', 'evaluation', 'train')


def reference_root():
    """Where the unmodified reference modules import from: the source tree in the build container, else the archive
    staged by oracle/stage_ref.py (what travels to the GPU box), else None."""
    if os.path.isfile(os.path.join(REFERENCE_DIR, 'layers.py')):
        return REFERENCE_DIR
    if os.path.isfile(STAGED_ARCHIVE):
        return STAGED_ARCHIVE
    return None


def reference_available():
    return reference_root() is not None


def import_reference(quiet=True, replace=None):
    """Return a dict {name: module} of the reference's modules, imported unmodified with `import dgl` resolving to the
    stand-in in oracle/dgl. `replace` maps module names to already-imported modules that take their place BEFORE the
    others are imported (the drop-in test puts this repo's `layers` and graph handle under the reference's own
    `model.py` / `train.py` that way: {'dgl': ..., 'layers': ...})."""
    root = reference_root()
    if root is None:
        raise RuntimeError('reference sources not present at %s and no staged archive at %s' % (REFERENCE_DIR, STAGED_ARCHIVE))
    for p in (REFERENCE_DIR, STAGED_ARCHIVE, ORACLE_DIR):
        if p in sys.path:
            sys.path.remove(p)
    sys.path.insert(0, root)
    sys.path.insert(0, ORACLE_DIR)        # `import dgl` -> oracle/dgl
    replace = dict(replace or {})
    for name in list(sys.modules):        # a different wiring than last time: re-import everything
        if name in _REF_MODULES or name == 'dgl' or name.startswith('dgl.'):
            del sys.modules[name]
    sys.modules.update(replace)
    mods = {}
    sink = io.StringIO()
    with contextlib.redirect_stdout(sink if quiet else sys.stdout):
        for name in _REF_MODULES:
            mods[name] = replace[name] if name in replace else importlib.import_module(name)
    return mods


def synthetic_mat_arrays(n_drug, n_dis, n_pos, embed_dim=768, sim_rank=32, seed=0):
    """Seeded synthetic dataset (SURVEY.md 8d config 1 generator)."""
    rng = np.random.default_rng(seed)
    ed_drug, ed_dis = embed_dim if isinstance(embed_dim, (tuple, list)) else (embed_dim, embed_dim)
    cells = rng.choice(n_drug * n_dis, size=n_pos, replace=False)
    assoc = np.zeros((n_drug, n_dis), dtype=np.float64)
    assoc[cells // n_dis, cells % n_dis] = 1.0

    def sim(n):
        x = rng.standard_normal((n, sim_rank))
        x /= np.linalg.norm(x, axis=1, keepdims=True)
        s = (x @ x.T + 1.0) / 2.0
        np.fill_diagonal(s, 1.0)
        return s

    names = np.empty((n_drug, 1), dtype=object)
    for i in range(n_drug):
        names[i, 0] = np.array(['DB%05d' % i])
    return {
        'didr': assoc.T.copy(),
        'drug': sim(n_drug),
        'disease': sim(n_dis),
        'drug_embed': rng.standard_normal((n_drug, ed_drug)),
        'disease_embed': rng.standard_normal((n_dis, ed_dis)),
        'Wrname': names,
    }


def write_synthetic_mat(root, name, **kw):
    """Write `<root>/raw_data/drug_data/<name>/<name>.mat` (data_loader.py:33-38 relative paths)."""
    import scipy.io as sio
    d = os.path.join(root, 'raw_data', 'drug_data', name)
    os.makedirs(d, exist_ok=True)
    arrays = synthetic_mat_arrays(**kw)
    sio.savemat(os.path.join(d, name + '.mat'), arrays)
    return arrays


@contextlib.contextmanager
def chdir(path):
    old = os.getcwd()
    os.chdir(path)
    try:
        yield
    finally:
        os.chdir(old)


def load_reference_dataset(root, name, k=4, quiet=True):
    """Instantiate the reference's DrugDataLoader on a synthetic `.mat` under `root`."""
    import torch as th
    mods = import_reference(quiet)
    sink = io.StringIO()
    with chdir(root), contextlib.redirect_stdout(sink if quiet else sys.stdout):
        ds = mods['data_loader'].DrugDataLoader(name, th.device('cpu'), symm=True, k=k)
    return mods, ds
