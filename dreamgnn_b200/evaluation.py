"""Mirror of the reference's `evaluate` (evaluation.py:4-74): eval-mode forward on the evaluated split's
own encoder / decoder graphs, un-augmented inputs, sklearn ROC / PR areas on the raw logits. The metric
code is the reference's CPU sklearn path on purpose -- it is the AUROC/AUPR parity instrument."""
import torch as th
from sklearn import metrics


def evaluate(args, model, graph_data, drug_graph, drug_feat, drug_sim_feat, dis_graph, dis_feat, dis_sim_feat,
             drug_feature_graph=None, disease_feature_graph=None, return_predictions=False):
    rating_values = graph_data['test'][2]
    enc_graph = graph_data['test'][0].int().to(args.device)      # conversions are cached on the graph handle
    dec_graph = graph_data['test'][1].int().to(args.device)
    drug_graph, dis_graph = drug_graph.to(args.device), dis_graph.to(args.device)
    if drug_feature_graph is not None:
        drug_feature_graph = drug_feature_graph.to(args.device)
    if disease_feature_graph is not None:
        disease_feature_graph = disease_feature_graph.to(args.device)
    model.eval()
    with th.no_grad():
        pred_ratings = model(enc_graph, dec_graph, drug_graph, drug_sim_feat, drug_feat, dis_graph, dis_sim_feat,
                             dis_feat, drug_feature_graph, disease_feature_graph)[0]
    y_score = pred_ratings.view(-1).cpu().numpy()
    y_true = rating_values.cpu().numpy()
    fpr, tpr, _ = metrics.roc_curve(y_true, y_score)
    auc = metrics.auc(fpr, tpr)
    precision, recall, _ = metrics.precision_recall_curve(y_true, y_score)
    aupr = metrics.auc(recall, precision)
    if return_predictions:
        return auc, aupr, (y_score, y_true)
    return auc, aupr
