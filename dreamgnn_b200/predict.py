"""`get_top_novel_predictions` (train.py:26-151): score every (drug, disease) pair that is NOT a known
association and keep the top K.

The reference walks the N_d x N_s grid in a Python double loop, then runs the WHOLE model once per 5 000-pair
batch (104 batches on lrssl), each time rebuilding a decoder graph, and sorts a list of dicts with pandas. In
eval mode the encoder output does not depend on the scored pairs, so here the node embeddings are computed
once (`Net.embed`), all novel pairs go through the fused decoder kernel in a single launch, and the top K
are selected on the device. Same signature, same DataFrame columns, same CSV file name.
"""
import os

import numpy as np
import torch as th

from . import ops


def novel_pair_scores(args, model, dataset, cv_idx):
    """(drug_id, disease_id, probability) device tensors over all pairs with association_matrix == 0, in the
    reference's row-major (drug-major) order."""
    dev = th.device(args.device)
    model.eval()
    gd = dataset.get_graph_data_for_training(cv_idx)
    known = th.as_tensor(np.asarray(dataset.association_matrix) != 0, device=dev)
    drug_id, dis_id = th.nonzero(~known, as_tuple=True)                    # row-major = the reference's nested loops
    with th.no_grad():
        drug_emb, dis_emb = model.embed(gd['train_enc_graph'].int().to(dev), gd['drug_graph'], gd['drug_sim_features'],
                                        gd['drug_features'], gd['disease_graph'], gd['disease_sim_features'],
                                        gd['disease_features'], gd['drug_feature_graph'], gd['disease_feature_graph'])[4:]
        dec = model.decoder
        n_in = drug_emb.shape[1]
        w1 = dec.lin1.weight
        pd = ops.linear(drug_emb, w1[:, :n_in], dec.lin1.bias)
        ps = ops.linear(dis_emb, w1[:, n_in:])
        pairs = ops.PairGraph(drug_id, dis_id, dataset.num_drug, dataset.num_disease)
        logits = ops.decoder_mlp(pd, ps, dec.lin2.weight, dec.lin2.bias, dec.lin3.weight, dec.lin3.bias, pairs,
                                 p=0.0, seed=0, training=False)
    return drug_id, dis_id, th.sigmoid(logits.squeeze(-1))


def get_top_novel_predictions(args, model, dataset, cv_idx, top_k=200):
    import pandas as pd_
    print('Generating top %d novel predictions for fold %d...' % (top_k, cv_idx + 1))
    drug_id, dis_id, score = novel_pair_scores(args, model, dataset, cv_idx)
    print('Found %d potential novel drug-disease pairs.' % score.numel())
    if score.numel() == 0:
        return pd_.DataFrame(columns=['drug_id', 'disease_id', 'score'])
    k = min(int(top_k), score.numel())
    top = th.topk(score, k, sorted=True)
    idx = top.indices
    df = pd_.DataFrame({'drug_id': drug_id[idx].cpu().numpy(), 'disease_id': dis_id[idx].cpu().numpy(),
                        'score': top.values.cpu().numpy().astype(np.float64)})
    names = getattr(dataset, 'drug_ids', None)
    if names is not None:
        df['drug_name'] = df['drug_id'].map({i: n for i, n in enumerate(names)})
    try:
        path = os.path.join(args.save_dir, 'top%d_novel_predictions_fold%d.csv' % (top_k, cv_idx + 1))
        df.to_csv(path, index=False)
        print('Top %d novel predictions saved to %s' % (top_k, path))
    except Exception as e:                                                   # noqa: BLE001 -- the reference reports and goes on
        print('Error saving CSV file: %s' % e)
    return df
