"""The `dgl` module the reference imports, served by this package.

    import dreamgnn_b200.dgl_compat as dgl_compat
    dgl_compat.install()            # before the reference's modules are imported
    import train                    # the reference's own train.py / model.py / layers.py / augmentation.py / data_loader.py

registers module objects named `dgl`, `dgl.function`, `dgl.fn`, `dgl.nn` and `dgl.nn.pytorch` in `sys.modules` that expose
exactly the DGL surface the reference touches: `heterograph`, `bipartite_from_scipy`, `DGLGraph`,
`DGLHeteroGraph`, `DGLError`, `function.copy_u / copy_src / sum`, `nn.pytorch.HeteroGraphConv`. Everything resolves to
`dreamgnn_b200.graph.HeteroGraph` (relation-block CSRs on device, `update_all(copy_u, sum)` = the sm_100a SpMM kernel
with a deterministic transposed-CSR backward, `apply_edges` = row gathers) and `dreamgnn_b200.layers.HeteroGraphConv`.
With only this shim installed the reference's `layers.py` runs unmodified on the CUDA kernels (layers.py:174-233 ->
`dg_spmm_csr_f32`); replacing `layers` by `dreamgnn_b200.layers` as well switches to the fused relation-block path.
"""
import sys
import types

from . import graph as _graph


def build_modules():
    from .layers import HeteroGraphConv
    dgl = types.ModuleType('dgl')
    dgl.__doc__ = 'DGL surface of DREAM-GNN served by dreamgnn_b200.graph (see dreamgnn_b200/dgl_compat.py)'
    dgl.__path__ = []                                   # a package: its `function` / `fn` / `nn.pytorch` submodules must resolve
    for name in ('heterograph', 'bipartite_from_scipy', 'DGLGraph', 'DGLHeteroGraph', 'DGLError', 'HeteroGraph'):
        setattr(dgl, name, getattr(_graph, name))
    fn = types.ModuleType('dgl.function')
    fn.copy_u, fn.copy_src, fn.sum = _graph.function.copy_u, _graph.function.copy_src, _graph.function.sum
    nn_ = types.ModuleType('dgl.nn')
    nn_.__path__ = []
    nn_pt = types.ModuleType('dgl.nn.pytorch')
    nn_pt.HeteroGraphConv = HeteroGraphConv
    nn_.pytorch = nn_pt
    dgl.function, dgl.fn, dgl.nn = fn, fn, nn_
    return {'dgl': dgl, 'dgl.function': fn, 'dgl.fn': fn, 'dgl.nn': nn_, 'dgl.nn.pytorch': nn_pt}


def install():
    """Put the shim into sys.modules (replacing a real or stand-in `dgl` if one was imported). Returns the modules."""
    mods = build_modules()
    for name in [m for m in sys.modules if m == 'dgl' or m.startswith('dgl.')]:
        del sys.modules[name]
    sys.modules.update(mods)
    return mods
