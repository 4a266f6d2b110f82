"""TEST INFRASTRUCTURE ONLY -- `dgl.nn.pytorch` stand-in: HeteroGraphConv as used at
layers.py:98 (ctor) and layers.py:129 (forward with `mod_args`)."""
import torch as th
import torch.nn as nn


class HeteroGraphConv(nn.Module):
    def __init__(self, mods, aggregate='sum'):
        super().__init__()
        self.mods = nn.ModuleDict(mods)
        self.aggregate = aggregate

    def forward(self, g, inputs, mod_args=None, mod_kwargs=None):
        mod_args = mod_args or {}
        mod_kwargs = mod_kwargs or {}
        outputs = {nt: [] for nt in g.dsttypes}
        for stype, etype, dtype in g.canonical_etypes:
            if stype not in inputs:
                continue
            rel = g[stype, etype, dtype]
            out = self.mods[etype](rel, (inputs[stype], inputs[dtype]),
                                   *mod_args.get(etype, ()), **mod_kwargs.get(etype, {}))
            outputs[dtype].append(out)
        rsts = {}
        for nt, alist in outputs.items():
            if not alist:
                continue
            if self.aggregate == 'sum':
                rsts[nt] = th.stack(alist, dim=0).sum(0)
            elif self.aggregate == 'stack':
                rsts[nt] = th.stack(alist, dim=1)
            elif self.aggregate == 'mean':
                rsts[nt] = th.stack(alist, dim=0).mean(0)
            elif self.aggregate == 'max':
                rsts[nt] = th.stack(alist, dim=0).max(0)[0]
            else:
                raise ValueError(self.aggregate)
        return rsts
