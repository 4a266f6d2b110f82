"""`Net` -- drop-in for the reference's top-level module (model.py:4-103).

Contract kept: `Net(args)` consumes the same `args` attributes, registers its children under the same
names (`TGCN.<l>`, `FGCN`, `attention`, `decoder` -> identical state_dict keys), builds them in the same
order (so a given torch seed yields identical initial weights), and `forward` takes the reference's eleven
arguments and returns `(pred, drug_out, drug_sim_out, dis_out, dis_sim_out)`.
"""
import torch as th
import torch.nn as nn

from . import layers as L
from . import ops
from .utils import get_activation


def _topology_layer(args, act, first):
    """GCMC layer l: layer 0 maps the raw embeddings (src/dst widths, msg = gcn_agg_units // 3);
    deeper layers are out_units -> out_units with `ini=False` (model.py:9-42)."""
    n_out = args.gcn_out_units
    if first:
        dims, msg, extra = (args.src_in_units, args.dst_in_units), args.gcn_agg_units, {}
    else:
        stacked = args.gcn_agg_accum == 'stack'
        dims, msg, extra = (n_out, n_out), n_out * (len(args.rating_vals) if stacked else 1), {'ini': False}
    return L.GCMCLayer(args.rating_vals, dims[0], dims[1], msg, n_out, args.dropout, args.gcn_agg_accum,
                       agg_act=act, share_user_item_param=args.share_param, device=args.device, **extra)


class Net(nn.Module):
    def __init__(self, args):
        super().__init__()
        for name in ('layers', 'gcn_agg_accum', 'rating_vals', 'device', 'gcn_agg_units', 'src_in_units'):
            setattr(self, name, getattr(args, name))
        self._act = get_activation(args.model_activation)
        self.TGCN = nn.ModuleList(_topology_layer(args, self._act, first=(l == 0)) for l in range(args.layers))
        self.FGCN = L.FGCN(args.fdim_drug, args.fdim_disease, args.nhid1, args.nhid2, args.dropout)
        self.attention = L.Attention(args.gcn_out_units, dropout_rate=args.attention_dropout)
        self.decoder = L.MLPDecoder(in_units=args.gcn_out_units, dropout_rate=args.dropout)
        self.parallel_routes = False        # not part of the reference API: see `embed`
        self.routes_ready = None

    def _topology_route(self, enc_graph, drug, dis, two_stage):
        """Stacked GCMC layers with the 1/(l+1)-weighted sum of their outputs (model.py:67-76)."""
        acc = None
        for depth, layer in enumerate(self.TGCN, start=1):
            drug, dis = layer(enc_graph, drug, dis, two_stage)
            # out += o / (l + 1) (model.py:70-76) as one scaled add
            acc = (drug, dis) if acc is None else (th.add(acc[0], drug, alpha=1.0 / depth), th.add(acc[1], dis, alpha=1.0 / depth))
        return acc

    def _fuse(self, topo, feat):
        """Shared two-view attention (model.py:93-97)."""
        return self.attention.fuse(topo, feat)[0]

    def embed(self, enc_graph, drug_graph, drug_sim_feat, drug_feat, dis_graph, disease_sim_feat, dis_feat,
              drug_feature_graph=None, disease_feature_graph=None, Two_Stage=False):
        """Everything of `forward` up to the decoder: the four route outputs and the two fused node embeddings the
        decoder scores pairs from (model.py:65-97). Not part of the reference API: `forward` is this + the decoder,
        and the all-pairs scoring of `predict.get_top_novel_predictions` calls it once instead of once per batch."""
        # The topology route (GCMC layers on the association graph) and the feature route (FGCN on the kNN graphs)
        # share no intermediate: with `parallel_routes` (or ops.PARALLEL_BRANCHES) they -- and the per-node-type halves
        # inside them -- are recorded as parallel stream branches. Pays on the launch-bound real-dataset shapes, where
        # no single branch fills the GPU; `graphed.GraphedIteration` switches it on for those.
        with ops.parallel_branches(self.parallel_routes and drug_feat.is_cuda):
            (drug_out, dis_out), (drug_sim_out, dis_sim_out) = ops.branches([
                lambda: self._topology_route(enc_graph, drug_feat, dis_feat, Two_Stage),
                lambda: self.FGCN(drug_graph, drug_sim_feat, dis_graph, disease_sim_feat, drug_feature_graph,
                                  disease_feature_graph)[:2]])
        # with parallel branches on, mark the point where the four route outputs exist: the common losses (train.py:292-293)
        # depend on nothing later and can run beside the attention + decoder (train.train_iteration forks there)
        self.routes_ready = None
        if self.parallel_routes and drug_feat.is_cuda:
            self.routes_ready = th.cuda.Event()
            self.routes_ready.record()
        return (drug_out, drug_sim_out, dis_out, dis_sim_out,
                self._fuse(drug_out, drug_sim_out), self._fuse(dis_out, dis_sim_out))

    def forward(self, enc_graph, dec_graph, drug_graph, drug_sim_feat, drug_feat, dis_graph, disease_sim_feat,
                dis_feat, drug_feature_graph=None, disease_feature_graph=None, Two_Stage=False):
        drug_out, drug_sim_out, dis_out, dis_sim_out, drug_emb, dis_emb = self.embed(
            enc_graph, drug_graph, drug_sim_feat, drug_feat, dis_graph, disease_sim_feat, dis_feat, drug_feature_graph,
            disease_feature_graph, Two_Stage)
        scores = self.decoder(dec_graph, drug_emb, dis_emb)
        return scores, drug_out, drug_sim_out, dis_out, dis_sim_out
