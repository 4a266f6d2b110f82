"""`evaluate` with the reference's signature and protocol (evaluation.py:4-74): the model is scored in eval
mode on the evaluated split's OWN encoder / decoder graphs with un-augmented inputs. AUROC / AUPR are computed on
the device (`metrics.binary_curve_areas`: sklearn's curve construction restated as a radix sort + scan, equal to
sklearn to ~1e-15); `DG_EVAL=sklearn` switches back to the reference's CPU metric code, which is what the parity
tests compare against."""
import os

import torch as th

from .metrics import binary_curve_areas


def _areas_sklearn(y_true, y_score):
    from sklearn import metrics
    fpr, tpr, _ = metrics.roc_curve(y_true, y_score)
    precision, recall, _ = metrics.precision_recall_curve(y_true, y_score)
    return metrics.auc(fpr, tpr), metrics.auc(recall, precision)


def evaluate(args, model, graph_data, drug_graph, drug_feat, drug_sim_feat, dis_graph, dis_feat, dis_sim_feat,
             drug_feature_graph=None, disease_feature_graph=None, return_predictions=False):
    enc, dec, labels = graph_data['test'][:3]
    dev = args.device
    on_dev = lambda g: None if g is None else g.to(dev)
    enc, dec = enc.int().to(dev), dec.int().to(dev)            # cached on the graph handle after the first call
    was_training = model.training
    model.eval()
    with th.no_grad():
        logits = model(enc, dec, on_dev(drug_graph), drug_sim_feat, drug_feat, on_dev(dis_graph), dis_sim_feat,
                       dis_feat, on_dev(drug_feature_graph), on_dev(disease_feature_graph))[0]
    del was_training                                             # the reference leaves the model in eval mode too
    if os.environ.get('DG_EVAL') == 'sklearn':
        y_score, y_true = logits.view(-1).cpu().numpy(), labels.cpu().numpy()
        auroc, aupr = _areas_sklearn(y_true, y_score)
        return (auroc, aupr, (y_score, y_true)) if return_predictions else (auroc, aupr)
    auroc, aupr = binary_curve_areas(labels.to(dev), logits.view(-1))
    if return_predictions:                                       # only this path copies the scores to the host
        return auroc, aupr, (logits.view(-1).cpu().numpy(), labels.cpu().numpy())
    return auroc, aupr
