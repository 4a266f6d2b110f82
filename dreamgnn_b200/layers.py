"""Drop-in mirrors of the reference's `layers.py` modules (same constructor arguments, forward
signatures, parameter names / state_dict keys, initialisation order) running on the sm_100a kernels
behind `dreamgnn_b200.ops`. Citations are into /root/reference/layers.py.

What changes underneath:
  * GCMCLayer (layers.py:18-143): one projection GEMM per node type produces the messages of all R
    relations in one [R, N_src, D] buffer; one SpMM launch per destination type walks the combined
    relation-block CSR with `dropout(cj)[src]` and `ci[dst]` fused, i.e. HeteroGraphConv's
    per-etype update_all + stack/sum (K2-K6 in SURVEY.md) become GEMM + SpMM. Backward is the same
    SpMM on the transposed block -- deterministic, no atomics.
  * GCN / FGCN (layers.py:238-285): `x @ W` is computed once and reused for the similarity and the
    feature graph; spmm + bias (+ReLU) run in one kernel on a cached CSR of the COO adjacency.
  * MLPDecoder (layers.py:341-379): exact split of lin1 over the concat + fused gather/MLP kernel.
The per-relation `GCMCGraphConv.forward` / `HeteroGraphConv.forward` entry points stay available.
"""
import math

import torch as th
import torch.nn as nn
import torch.nn.functional as F
from torch.nn import init
from torch.nn.parameter import Parameter

from . import ops
from .graph import DGLError
from .utils import get_activation, to_etype_name

# feature storage for the gathered GCMC messages: th.float32 (1e-5 parity path) or th.bfloat16
# (2e-2 path: bf16 storage, fp32 accumulate)
MESSAGE_DTYPE = th.float32
# message width is padded (zero weight columns) to a multiple of this many elements; 0 = the vector width (4 fp32 / 8 bf16).
# 32 fp32 = whole 128-byte lines per gathered row (341 -> 352 instead of 344): see DESIGN.md on the d = 344 SpMM
import os as _os
MESSAGE_PAD = int(_os.environ.get('DG_MSG_PAD', '0'))


def _pad_cols(t, mult):
    pad = (-t.shape[-1]) % mult
    return F.pad(t, (0, pad)) if pad else t


def _flat_f32(t):
    return t.reshape(-1).to(th.float32).contiguous()


def _act_code(act):
    """(code, slope) of an activation the fused act + dropout kernel implements, else None."""
    if isinstance(act, nn.LeakyReLU):
        return 'leaky', float(act.negative_slope)
    if isinstance(act, nn.ReLU):
        return 'relu', 0.0
    if not isinstance(act, nn.Module) and act is not None and act(th.tensor(-2.0)).item() == -2.0:
        return None, 0.0                                     # get_activation(None): the identity lambda
    return False


def adjacency_csr(adj):
    """CSR sidecar of a torch sparse-COO adjacency (cached on the tensor object). The COO may be
    uncoalesced (augmentation.py:124): duplicates simply stay separate entries of the row."""
    csr = getattr(adj, '_dg_csr', None)
    if csr is None:
        if not (isinstance(adj, th.Tensor) and adj.is_sparse):
            raise TypeError('adjacency must be a torch sparse COO tensor')
        if not adj.is_cuda:
            raise RuntimeError('GraphConvolution needs CUDA tensors (no CPU fallback)')
        idx = adj._indices()
        csr = ops.CSR.from_coo(idx[0], idx[1], adj.shape[0], adj.shape[1], adj._values().to(th.float32))
        adj._dg_csr = csr
    return csr


class HeteroGraphConv(nn.Module):
    """Mirror of dgl.nn.pytorch.HeteroGraphConv as used at layers.py:98 / :129 (generic per-relation
    loop; GCMCLayer.forward takes the fused relation-block path instead)."""

    def __init__(self, mods, aggregate='sum'):
        super().__init__()
        self.mods = nn.ModuleDict(mods)
        self.aggregate = aggregate

    def forward(self, g, inputs, mod_args=None, mod_kwargs=None):
        mod_args, mod_kwargs = mod_args or {}, mod_kwargs or {}
        outputs = {nt: [] for nt in g.dsttypes}
        for stype, etype, dtype in g.canonical_etypes:
            if stype not in inputs:
                continue
            out = self.mods[etype](g[stype, etype, dtype], (inputs[stype], inputs[dtype]),
                                   *mod_args.get(etype, ()), **mod_kwargs.get(etype, {}))
            outputs[dtype].append(out)
        rsts = {}
        for nt, alist in outputs.items():
            if alist:
                rsts[nt] = th.stack(alist, dim=0).sum(0) if self.aggregate == 'sum' else th.stack(alist, dim=1)
        return rsts


class GCMCGraphConv(nn.Module):
    """layers.py:146-236."""

    def __init__(self, in_feats, out_feats, weight=True, device=None, dropout_rate=0.1):
        super().__init__()
        self._in_feats, self._out_feats, self.device = in_feats, out_feats, device
        self.dropout = nn.Dropout(dropout_rate)
        if weight:
            self.weight = nn.Parameter(th.Tensor(in_feats, out_feats))
        else:
            self.register_parameter('weight', None)
        self.reset_parameters()

    def reset_parameters(self):
        if self.weight is not None:
            init.xavier_uniform_(self.weight)

    def forward(self, graph, feat, weight=None, Two_Stage=False):
        with graph.local_scope():
            if isinstance(feat, tuple):
                feat, _ = feat
            cj, ci = graph.srcdata['cj'], graph.dstdata['ci']
            if feat.size(0) != graph.number_of_src_nodes():
                raise ValueError('feat has %d rows for %d source nodes' % (feat.size(0), graph.number_of_src_nodes()))
            if weight is not None:
                if self.weight is not None:
                    raise DGLError('External weight provided but module also has its own weight parameter, '
                                   'please set weight=False.')
            else:
                weight = self.weight
            if weight is not None:
                feat = dot_or_identity(feat, weight, self.device)
            d = feat.shape[1]
            rst = ops.spmm(graph.etype_csr(), _pad_cols(feat, 4), src_scale=_flat_f32(self.dropout(cj)),
                           dst_scale=_flat_f32(ci), tag='gcmc')
        return rst[:, :d]


class GCMCLayer(nn.Module):
    """layers.py:18-143."""

    def __init__(self, rating_vals, user_in_units, movie_in_units, msg_units, out_units, dropout_rate=0.1,
                 agg='stack', agg_act=None, ini=True, share_user_item_param=False, basis_units=2, device=None):
        super().__init__()
        self.rating_vals = rating_vals
        self.agg = agg
        self.share_user_item_param = share_user_item_param
        self.user_in_units = user_in_units
        effective = msg_units
        if agg == 'stack':
            assert effective % len(rating_vals) == 0
            effective = effective // len(rating_vals)
        if ini:
            effective = effective // 3
        self.msg_units = effective
        self.ufc = nn.Linear(effective, out_units)
        self.ifc = self.ufc if share_user_item_param else nn.Linear(effective, out_units)
        self.dropout = nn.Dropout(dropout_rate)
        self.W_r = {}
        self.basis_units = basis_units
        self.att = nn.Parameter(th.randn(len(self.rating_vals), basis_units))
        self.basis = nn.Parameter(th.randn(basis_units, user_in_units, effective))
        sub = {}
        shared_dims = share_user_item_param and user_in_units == movie_in_units
        for rating in rating_vals:
            rating = to_etype_name(rating)
            if shared_dims:
                sub[rating] = GCMCGraphConv(user_in_units, effective, False, device, dropout_rate)
                sub['rev-%s' % rating] = GCMCGraphConv(user_in_units, effective, False, device, dropout_rate)
            else:
                self.W_r = None
                sub[rating] = GCMCGraphConv(user_in_units, effective, True, device, dropout_rate)
                sub['rev-%s' % rating] = GCMCGraphConv(movie_in_units, effective, True, device, dropout_rate)
        self.conv = HeteroGraphConv(sub, aggregate=agg)
        self.agg_act = get_activation(agg_act)
        self._act_code = _act_code(self.agg_act)
        self.device = device
        self.reset_parameters()

    def partial_to(self, device):
        assert device == self.device
        if device is not None:
            self.ufc.cuda(device)
            if not self.share_user_item_param:
                self.ifc.cuda(device)
            self.dropout.cuda(device)

    def reset_parameters(self):
        for p in self.parameters():
            if p.dim() > 1:
                nn.init.xavier_uniform_(p)

    def _relation_weights(self):
        """etype -> [in, msg] weight: W = att @ basis for the shared-dims branch (layers.py:120-127),
        the per-relation module's own parameter otherwise (layers.py:86-97)."""
        if self.W_r is None:
            return {et: m.weight for et, m in self.conv.mods.items()}
        if self.att.is_cuda:
            # one elementwise kernel each way, written at the padded message width the aggregation wants (K2)
            mult = MESSAGE_PAD or (8 if MESSAGE_DTYPE == th.bfloat16 else 4)
            self.W_padded = ops.basis_combine(self.att, self.basis, mult)
            W = self.W_padded[:, :, :self.msg_units]
        else:
            W = th.matmul(self.att, self.basis.view(self.basis_units, -1)).view(-1, self.user_in_units, self.msg_units)
            self.W_padded = None
        self.W = W
        out = {}
        for i, rating in enumerate(self.rating_vals):
            rating = to_etype_name(rating)
            out[rating] = W[i]
            out['rev-%s' % rating] = W[i]
        return out

    def _rating_index(self, etype):
        """Position in `rating_vals` of the rating behind an etype name ('3' or 'rev-3')."""
        name = etype[4:] if etype.startswith('rev-') else etype
        for i, rating in enumerate(self.rating_vals):
            if to_etype_name(rating) == name:
                return i
        raise KeyError(etype)

    def forward(self, graph, drug_feat=None, dis_feat=None, Two_Stage=False):
        if self.agg != 'sum':
            raise NotImplementedError("only agg='sum' is usable downstream (as in the reference)")
        feats = {'drug': drug_feat, 'disease': dis_feat}
        weights = self._relation_weights()
        D = self.msg_units
        mult = MESSAGE_PAD or (8 if MESSAGE_DTYPE == th.bfloat16 else 4)
        seen = []
        for c in graph.canonical_etypes:                 # blocks in the reference's etype order
            if c[2] not in seen:
                seen.append(c[2])

        part = getattr(graph, 'partition', None)                     # row-partitioned graph: features hold the owned rows
        memo = {}

        def weight_stack(blk):
            """[R, in, Dp] weights of a block's relations, message width zero-padded to the vector width. Shared-dims
            branch: both node types' blocks use W = att @ basis in rating order, so it is padded once per layer and
            shared (autograd then adds two gradients instead of un-stacking / un-padding / scattering each slice)."""
            names = [c[1] for c in blk.etypes]
            if self.W_r is None:
                return _pad_cols(th.stack([weights[n] for n in names], dim=0), mult)
            order = tuple(self._rating_index(n) for n in names)
            if 'padded' not in memo:
                memo['padded'] = self.W_padded if self.W_padded is not None else _pad_cols(self.W, mult)
            if order not in memo:
                memo[order] = memo['padded'] if order == tuple(range(self.W.shape[0])) else memo['padded'][list(order)]
            return memo[order]

        def padded(w):
            """Zero-padded copy of an output-layer weight, made once per forward (ifc is ufc with share_param)."""
            if id(w) not in memo:
                memo[id(w)] = _pad_cols(w, mult)
            return memo[id(w)]

        def messages(dst_type):
            """Projection operands of one destination type: (block, x, [R, in, Dp] weights, [R * N_src] dropout(cj) scales)."""
            blk = graph.block(dst_type)
            x = feats[blk.src_type]
            if part is None and x.size(0) != blk.n_src:
                raise ValueError('%s features have %d rows for %d nodes' % (blk.src_type, x.size(0), blk.n_src))
            wstack = weight_stack(blk)                                   # [R, in, Dp]
            cj = _flat_f32(graph.nodes[blk.src_type].data['cj'])
            drops = [self.conv.mods[c[1]].dropout for c in blk.etypes]
            if all(d.p == drops[0].p and d.training == drops[0].training for d in drops):
                # the R per-relation dropout(cj) draws (layers.py:222) as ONE draw over the [R, N_src] expansion
                scale = F.dropout(cj.unsqueeze(0).expand(len(drops), -1), drops[0].p, drops[0].training).reshape(-1)
            else:
                scale = th.stack([d(cj) for d in drops], dim=0).reshape(-1)
            return blk, x, wstack, scale

        def aggregate(dst_type):
            """All relations into one node type: one batched projection, one SpMM."""
            blk, x, wstack, scale = messages(dst_type)
            h = ops.project(x, wstack)                               # [R, N_src, Dp]: all relations' messages
            dp = wstack.shape[2]
            ci = _flat_f32(graph.nodes[dst_type].data['ci'])
            if MESSAGE_DTYPE != th.float32:
                h = h.to(MESSAGE_DTYPE)
            return ops.spmm(blk.csr, h.reshape(blk.num_rel * blk.n_src, dp), src_scale=scale, dst_scale=ci, tag='gcmc')

        def aggregate_partitioned():
            """Row-partitioned graph: both node types' exchanges are issued before either aggregation, so the all-gather
            of one overlaps the projection / SpMM of the other (NVLink transfer under compute)."""
            from . import dist as _dist
            pend = {}
            for t in seen:
                blk, x, wstack, scale = messages(t)
                pend[t] = (blk,) + _dist.gcmc_exchange(x, wstack, scale)
            out = {}
            for t in seen:
                blk, buf, src_scale, finish = pend[t]
                ci = _flat_f32(graph.nodes[t].data['ci'])
                out[t] = ops.spmm(blk.csr, finish(buf), src_scale=src_scale, dst_scale=ci, tag='gcmc')
            return out

        def tail(dst_type, agg):
            """Activation, dropout, output layer of one node type."""
            # The padded message columns (341 -> 344) are exactly zero through aggregation, activation and dropout, so
            # the tail runs on the padded width with zero weight columns appended to ifc / ufc: no slice copy forward,
            # no re-padding of the gradient backward, and the GEMM operands stay 16-byte aligned for TMA.
            code = self._act_code
            if code is False:                                      # tanh / gelu / ...: no fused instance
                y = self.dropout(self.agg_act(agg))
            else:                                                  # activation + dropout in one launch (K7)
                y = ops.act_dropout(agg, code[0], code[1], p=self.dropout.p, training=self.training)
            fc = self.ifc if dst_type == 'drug' else self.ufc
            w = padded(fc.weight) if y.shape[1] != D else fc.weight
            return ops.linear(y, w, fc.bias)

        missing = [t for t in ('drug', 'disease') if t not in seen]
        if missing:
            raise KeyError('no relation into node type %r' % missing[0])
        def aggregate_one_partitioned(dst_type):
            """Row-partitioned graph, one node type on its own stream branch: exchange, then aggregate. The collectives of
            this branch (forward all-gather, backward reduce-scatter: blocking on ITS stream) overlap the other branch's
            projection / SpMM in both directions."""
            from . import dist as _dist
            blk, x, wstack, scale = messages(dst_type)
            buf, src_scale, finish = _dist.gcmc_exchange(x, wstack, scale)
            ci = _flat_f32(graph.nodes[dst_type].data['ci'])
            return ops.spmm(blk.csr, finish(buf), src_scale=src_scale, dst_scale=ci, tag='gcmc')

        if ops.PARALLEL_BRANCHES:                                  # the two node types as parallel stream branches
            one = aggregate if part is None else aggregate_one_partitioned
            drug, dis = ops.branches([lambda: tail('drug', one('drug')), lambda: tail('disease', one('disease'))])
            return drug, dis
        # serial: the reference's dropout draw order (cj per etype in canonical order, then drug, then disease)
        aggs = aggregate_partitioned() if part is not None else {t: aggregate(t) for t in seen}
        return tail('drug', aggs['drug']), tail('disease', aggs['disease'])


class GraphConvolution(nn.Module):
    """layers.py:287-321."""

    def __init__(self, in_features, out_features, bias=True):
        super().__init__()
        self.in_features, self.out_features = in_features, out_features
        self.weight = Parameter(th.FloatTensor(in_features, out_features))
        if bias:
            self.bias = Parameter(th.FloatTensor(out_features))
        else:
            self.register_parameter('bias', None)
        self.reset_parameters()

    def reset_parameters(self):
        stdv = 1. / math.sqrt(self.weight.size(1))
        self.weight.data.uniform_(-stdv, stdv)
        if self.bias is not None:
            self.bias.data.uniform_(-stdv, stdv)

    def support(self, input):
        return ops.project(input, self.weight.unsqueeze(0))[0]

    def aggregate(self, support, adj, relu=False, gathered=None):
        d = support.shape[1]
        bias = self.bias
        if d % 4:
            support = _pad_cols(support, 4)
            bias = _pad_cols(bias, 4) if bias is not None else None
        if hasattr(adj, 'partition'):                            # owned rows of a row-partitioned kNN graph
            from . import dist as _dist
            full = gathered if gathered is not None else _dist.all_gather_rows(support)
            out = ops.spmm(adj.csr, full, bias=bias, relu=relu, tag='fgcn')
        else:
            out = ops.spmm(adjacency_csr(adj), support, bias=bias, relu=relu, tag='fgcn')
        return out[:, :d] if out.shape[1] != d else out

    def forward(self, input, adj):
        if not hasattr(adj, 'partition') and adj.device != input.device:
            adj = adj.to(input.device)
        return self.aggregate(self.support(input), adj)

    def __repr__(self):
        return '%s (%d -> %d)' % (self.__class__.__name__, self.in_features, self.out_features)


class GCN(nn.Module):
    """layers.py:238-249."""

    def __init__(self, features, nhid, nhid2, dropout):
        super().__init__()
        self.gc1 = GraphConvolution(features, nhid)
        self.gc2 = GraphConvolution(nhid, nhid2)
        self.dropout = dropout

    def _tail(self, support1, adj, gathered=None):
        x = self.gc1.aggregate(support1, adj, relu=True, gathered=gathered)      # spmm + bias + ReLU in one kernel
        x = ops.act_dropout(x, None, p=self.dropout, training=self.training)
        return self.gc2(x, adj)

    def forward(self, x, adj):
        return self._tail(self.gc1.support(x), adj)

    def forward_shared(self, x, adjs):
        """Same layer on several graphs over the same input: x @ W1 is computed once (the reference
        recomputes it per graph, layers.py:263-270)."""
        s1 = self.gc1.support(x)
        gathered = None
        if adjs and hasattr(adjs[0], 'partition') and s1.shape[1] % 4 == 0:
            # row-partitioned graphs: the [n_loc, nhid1] support is all-gathered ONCE for both graphs (its gradient comes
            # back through one reduce-scatter of the two consumers' summed gradients)
            from . import dist as _dist
            gathered = _dist.all_gather_rows(s1)
        return [self._tail(s1, adj, gathered) for adj in adjs]


class FGCN(nn.Module):
    """layers.py:251-285."""

    def __init__(self, fdim_drug, fdim_disease, nhid1, nhid2, dropout):
        super().__init__()
        self.FGCN_drug = GCN(fdim_drug, nhid1, nhid2, dropout)
        self.FGCN_disease = GCN(fdim_disease, nhid1, nhid2, dropout)
        self.dropout = dropout
        self.drug_fusion = nn.Linear(nhid2 * 2, nhid2)
        self.disease_fusion = nn.Linear(nhid2 * 2, nhid2)

    def forward(self, drug_graph, drug_sim_feat, dis_graph, disease_sim_feat,
                drug_feature_graph=None, disease_feature_graph=None):
        if drug_feature_graph is None or disease_feature_graph is None:
            emb1_sim = self.FGCN_drug(drug_sim_feat, drug_graph)
            emb2_sim = self.FGCN_disease(disease_sim_feat, dis_graph)
            return emb1_sim, emb2_sim, emb1_sim, None, emb2_sim, None
        if ops.PARALLEL_BRANCHES:
            # the two node types as parallel stream branches (dropout draws then go drug, drug, disease, disease)
            def side(gcn, x, g_sim, g_feat, fusion):
                e_sim, e_feat = gcn.forward_shared(x, [g_sim, g_feat])
                fused = ops.linear(th.cat([e_sim, e_feat], dim=1), fusion.weight, fusion.bias)
                return ops.act_dropout(fused, 'relu', p=self.dropout, training=self.training), e_sim, e_feat
            (emb1, emb1_sim, emb1_feat), (emb2, emb2_sim, emb2_feat) = ops.branches([
                lambda: side(self.FGCN_drug, drug_sim_feat, drug_graph, drug_feature_graph, self.drug_fusion),
                lambda: side(self.FGCN_disease, disease_sim_feat, dis_graph, disease_feature_graph, self.disease_fusion)])
            return emb1, emb2, emb1_sim, emb1_feat, emb2_sim, emb2_feat
        if self.training and self.dropout > 0:
            # keep the reference's dropout draw order: drug(sim), disease(sim), drug(feat), disease(feat)
            s_d, s_s = self.FGCN_drug.gc1.support(drug_sim_feat), self.FGCN_disease.gc1.support(disease_sim_feat)
            g_d = g_s = None
            if hasattr(drug_graph, 'partition') and s_d.shape[1] % 4 == 0:      # row-partitioned: one all-gather per node type
                from . import dist as _dist
                g_d, g_s = _dist.all_gather_rows(s_d), _dist.all_gather_rows(s_s)
            emb1_sim = self.FGCN_drug._tail(s_d, drug_graph, g_d)
            emb2_sim = self.FGCN_disease._tail(s_s, dis_graph, g_s)
            emb1_feat = self.FGCN_drug._tail(s_d, drug_feature_graph, g_d)
            emb2_feat = self.FGCN_disease._tail(s_s, disease_feature_graph, g_s)
        else:
            emb1_sim, emb1_feat = self.FGCN_drug.forward_shared(drug_sim_feat, [drug_graph, drug_feature_graph])
            emb2_sim, emb2_feat = self.FGCN_disease.forward_shared(disease_sim_feat, [dis_graph, disease_feature_graph])
        fused_drug = ops.linear(th.cat([emb1_sim, emb1_feat], dim=1), self.drug_fusion.weight, self.drug_fusion.bias)
        fused_disease = ops.linear(th.cat([emb2_sim, emb2_feat], dim=1), self.disease_fusion.weight, self.disease_fusion.bias)
        emb1 = ops.act_dropout(fused_drug, 'relu', p=self.dropout, training=self.training)        # ReLU + dropout, one launch
        emb2 = ops.act_dropout(fused_disease, 'relu', p=self.dropout, training=self.training)
        return emb1, emb2, emb1_sim, emb1_feat, emb2_sim, emb2_feat


class Attention(nn.Module):
    """layers.py:324-338."""

    def __init__(self, in_size, hidden_size=16, dropout_rate=0.1):
        super().__init__()
        self.project = nn.Sequential(nn.Linear(in_size, hidden_size), nn.Tanh(), nn.Linear(hidden_size, 1, bias=False))
        self.dropout = nn.Dropout(dropout_rate)

    def fuse(self, za, zb):
        """`forward(th.stack((za, zb), dim=1))` without materialising the stack: one fused row kernel each way (K12)."""
        lin1, lin2 = self.project[0], self.project[2]
        if za.is_cuda and ops.attention_eligible(za, zb, lin1.weight):
            out, beta = ops.attention_fuse(za, zb, lin1.weight, lin1.bias, lin2.weight, p=self.dropout.p, training=self.training)
            return out, beta.unsqueeze(-1)
        return self._forward_torch(th.stack((za, zb), dim=1))

    def _forward_torch(self, z):
        beta = self.dropout(th.softmax(self.project(z), dim=1))
        return (beta * z).sum(1), beta

    def forward(self, z):
        if z.dim() == 3 and z.shape[1] == 2 and z.is_cuda and z.stride(2) == 1:
            return self.fuse(z[:, 0], z[:, 1])
        if not z.is_cuda:
            raise RuntimeError('Attention needs CUDA tensors (no CPU fallback)')
        return self._forward_torch(z)                              # K != 2 views: the reference's expression


class MLPDecoder(nn.Module):
    """layers.py:341-379."""

    def __init__(self, in_units, dropout_rate=0.1):
        super().__init__()
        self.dropout = nn.Dropout(dropout_rate)
        self.sigmoid = nn.Sigmoid()
        self.lin1 = nn.Linear(2 * in_units, 128)
        self.lin2 = nn.Linear(128, 64)
        self.lin3 = nn.Linear(64, 1)
        self.reset_parameters()

    def reset_parameters(self):
        self.lin1.reset_parameters()
        self.lin2.reset_parameters()
        self.lin3.reset_parameters()

    def forward(self, graph, drug_feat, dis_feat):
        n_in = drug_feat.shape[1]
        w1 = self.lin1.weight
        pd = ops.linear(drug_feat, w1[:, :n_in], self.lin1.bias)   # drug half of lin1 (+ bias)
        ps = ops.linear(dis_feat, w1[:, n_in:])                    # disease half
        if hasattr(graph, 'partition'):                            # this rank's slice of the pairs; node rows gathered
            from . import dist as _dist
            pairs, pd, ps = graph.pairs, _dist.all_gather_rows(pd), _dist.all_gather_rows(ps)
        else:
            pairs = graph.pair_graph()
        seed = 0
        if self.training and self.dropout.p > 0:
            # drawn on device (no host sync, and a captured CUDA graph gets a fresh seed on every replay)
            seed = ops.fresh_seed(pd.device)
        return ops.decoder_mlp(pd, ps, self.lin2.weight, self.lin2.bias, self.lin3.weight, self.lin3.bias, pairs,
                               p=self.dropout.p, seed=seed, training=self.training)


def udf_u_mul_e(edges):
    """layers.py:378-379 (kept for API compatibility with graph.apply_edges)."""
    return {'m': th.cat([edges.src['h'], edges.dst['h']], 1)}


def dot_or_identity(A, B, device=None):
    """layers.py:382-392."""
    if A is None:
        return B
    if A.shape[1] == 3:
        out = th.cat([B[A[:, 0].long()], B[A[:, 1].long()], B[A[:, 2].long()]], 1)
        return out if device is None else out.to(device)
    return th.matmul(A, B)
