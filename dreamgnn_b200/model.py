"""Drop-in mirror of the reference's `model.py` `Net` (model.py:4-103): same `args` fields, same
submodule / parameter names (checkpoints interchange), same forward signature and 5-tuple return."""
import torch as th
import torch.nn as nn

from .layers import FGCN, Attention, GCMCLayer, MLPDecoder
from .utils import get_activation


class Net(nn.Module):
    def __init__(self, args):
        super().__init__()
        self.layers = args.layers
        self._act = get_activation(args.model_activation)
        self.TGCN = nn.ModuleList()
        self.TGCN.append(GCMCLayer(args.rating_vals, args.src_in_units, args.dst_in_units, args.gcn_agg_units,
                                   args.gcn_out_units, args.dropout, args.gcn_agg_accum, agg_act=self._act,
                                   share_user_item_param=args.share_param, device=args.device))
        self.gcn_agg_accum = args.gcn_agg_accum
        self.rating_vals = args.rating_vals
        self.device = args.device
        self.gcn_agg_units = args.gcn_agg_units
        self.src_in_units = args.src_in_units
        for _ in range(1, args.layers):
            if args.gcn_agg_accum == 'stack':
                gcn_out_units = args.gcn_out_units * len(args.rating_vals)
            else:
                gcn_out_units = args.gcn_out_units
            self.TGCN.append(GCMCLayer(args.rating_vals, args.gcn_out_units, args.gcn_out_units, gcn_out_units,
                                       args.gcn_out_units, args.dropout, args.gcn_agg_accum, agg_act=self._act,
                                       share_user_item_param=args.share_param, ini=False, device=args.device))
        self.FGCN = FGCN(args.fdim_drug, args.fdim_disease, args.nhid1, args.nhid2, args.dropout)
        self.attention = Attention(args.gcn_out_units, dropout_rate=args.attention_dropout)
        self.decoder = MLPDecoder(in_units=args.gcn_out_units, dropout_rate=args.dropout)

    def forward(self, enc_graph, dec_graph, drug_graph, drug_sim_feat, drug_feat, dis_graph, disease_sim_feat,
                dis_feat, drug_feature_graph=None, disease_feature_graph=None, Two_Stage=False):
        # topology route: layer-weighted sum o0 + o1/2 + o2/3 (model.py:67-76)
        for i in range(self.layers):
            drug_o, dis_o = self.TGCN[i](enc_graph, drug_feat, dis_feat, Two_Stage)
            if i == 0:
                drug_out, dis_out = drug_o, dis_o
            else:
                drug_out = drug_out + drug_o / float(i + 1)
                dis_out = dis_out + dis_o / float(i + 1)
            drug_feat, dis_feat = drug_o, dis_o
        # feature route over the kNN similarity graphs (model.py:79-83)
        drug_sim_out, dis_sim_out = self.FGCN(drug_graph, drug_sim_feat, dis_graph, disease_sim_feat,
                                              drug_feature_graph, disease_feature_graph)[:2]
        # shared attention over the two views, drug first (model.py:93-97)
        drug_feats, _ = self.attention(th.stack([drug_out, drug_sim_out], dim=1))
        dis_feats, _ = self.attention(th.stack([dis_out, dis_sim_out], dim=1))
        pred_ratings = self.decoder(dec_graph, drug_feats, dis_feats)
        return pred_ratings, drug_out, drug_sim_out, dis_out, dis_sim_out
