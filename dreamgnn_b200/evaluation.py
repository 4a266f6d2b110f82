"""`evaluate` with the reference's signature and protocol (evaluation.py:4-74): the model is scored in eval
mode on the evaluated split's OWN encoder / decoder graphs with un-augmented inputs, and AUROC / AUPR come
from sklearn on the raw logits -- deliberately the reference's CPU metric code, it is the parity instrument."""
import torch as th
from sklearn import metrics


def _areas(y_true, y_score):
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
    y_score, y_true = logits.view(-1).cpu().numpy(), labels.cpu().numpy()
    auroc, aupr = _areas(y_true, y_score)
    return (auroc, aupr, (y_score, y_true)) if return_predictions else (auroc, aupr)
