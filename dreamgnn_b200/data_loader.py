"""`DrugDataLoader` with the reference's constructor, attributes and fold layout (data_loader.py:41-597),
building every graph with the device kernels of `graph_build`.

Host-side preprocessing that is not on the hot path (loadmat, KFold over positives and over ALL negatives
with random_state=1024, feature normalisation) follows the reference step by step so that folds, pair order
and labels are identical; the four kNN graphs do not depend on the fold and are built ONCE (the reference
rebuilds them ten times, data_loader.py:230-276) and shared by every fold's entry.
"""
import os

import numpy as np
import torch as th
import torch.nn.functional as F

from . import graph_build as GB

_paths = {
    'Gdataset': './raw_data/drug_data/Gdataset/Gdataset.mat',
    'Cdataset': './raw_data/drug_data/Cdataset/Cdataset.mat',
    'Ldataset': './raw_data/drug_data/Ldataset/lagcn',
    'lrssl': './raw_data/drug_data/lrssl/lrssl.mat',
}


class DrugDataLoader(object):
    def __init__(self, name, device, symm=True, k=5, use_augmentation=False, aug_params=None, n_folds=10):
        self._name, self._device, self._symm = name, th.device(device), symm
        self.num_neighbor = k
        self.use_augmentation = use_augmentation
        self.aug_params = aug_params or {}
        self._n_folds = n_folds
        self._dir = os.path.join(_paths[name])
        self._load_raw_data(self._dir, name)
        self.cv_data_dict = self._create_cv_splits()
        self.embedding_mode = 'pretrained'
        self._generate_feat()
        self.cv_specific_graphs = {}
        self._generate_cv_specific_graphs()
        self.data_cv = self._build_all_cv_data()

    # data_loader.py:99-134
    def _load_raw_data(self, file_path, data_name):
        import scipy.io as sio
        if data_name not in ('Gdataset', 'Cdataset', 'lrssl'):
            raise ValueError('no loader branch for %r (the reference has none either)' % data_name)
        data = sio.loadmat(file_path)
        self.association_matrix = data['didr'].T
        self.disease_sim_features = data['disease']
        self.drug_sim_features = data['drug']
        self.drug_ids = [str(x[0][0]).strip() for x in data['Wrname']] if 'Wrname' in data else None
        n_drug, n_dis = self.association_matrix.shape
        self.drug_embed = data['drug_embed'] if 'drug_embed' in data else np.random.normal(0, 0.1, (n_drug, 768))
        self.disease_embed = data['disease_embed'] if 'disease_embed' in data else np.random.normal(0, 0.1, (n_dis, 768))
        self._num_drug, self._num_disease = n_drug, n_dis

    # data_loader.py:136-203 (KFold on positives and on all negatives, positives listed first)
    def _create_cv_splits(self):
        from sklearn.model_selection import KFold
        inter = self.association_matrix
        pos_row, pos_col = np.nonzero(inter)
        neg_row, neg_col = np.nonzero(1 - inter)
        kfold = KFold(n_splits=self._n_folds, shuffle=True, random_state=1024)
        cv = {}
        for i, ((tr_p, te_p), (tr_n, te_n)) in enumerate(zip(kfold.split(pos_row), kfold.split(neg_row))):
            def pack(pi, ni):
                rows = np.concatenate([pos_row[pi], neg_row[ni]]).astype(np.int64)
                cols = np.concatenate([pos_col[pi], neg_col[ni]]).astype(np.int64)
                vals = np.zeros(rows.size, dtype=np.float32)
                vals[:len(pi)] = 1
                return {'drug_id': rows, 'disease_id': cols, 'values': vals}
            cv[i] = [pack(tr_p, tr_n), pack(te_p, te_n), np.array([0, 1])]
        return cv

    # data_loader.py:205-228
    def _generate_feat(self):
        self.drug_feature = F.normalize(th.FloatTensor(self.drug_embed).to(self._device), p=2, dim=1)
        self.disease_feature = F.normalize(th.FloatTensor(self.disease_embed).to(self._device), p=2, dim=1)
        self.drug_feature_shape = self.drug_feature.shape
        self.disease_feature_shape = self.disease_feature.shape

    # data_loader.py:230-344 -- fold-independent, built once
    def _generate_cv_specific_graphs(self):
        k, dev = self.num_neighbor, self._device
        shared = {
            'drug_graph': GB.create_similarity_graph(self.drug_sim_features, k, dev, self._symm),
            'disease_graph': GB.create_similarity_graph(self.disease_sim_features, k, dev, self._symm),
            'drug_feature_graph': GB.create_feature_similarity_graph(self.drug_embed, k, dev, symm=self._symm),
            'disease_feature_graph': GB.create_feature_similarity_graph(self.disease_embed, k, dev, symm=self._symm),
        }
        for cv_idx in range(self._n_folds):
            train = self.cv_data_dict[cv_idx][0]
            assoc = np.zeros_like(self.association_matrix)
            pos = train['values'] == 1
            assoc[train['drug_id'][pos], train['disease_id'][pos]] = 1
            self.cv_specific_graphs[cv_idx] = dict(shared, train_association_matrix=assoc)

    # data_loader.py:346-398
    def _build_all_cv_data(self):
        out = {}
        for cv_idx in range(self._n_folds):
            entry = {}
            for split, info in zip(('train', 'test'), self.cv_data_dict[cv_idx][:2]):
                pairs, values = self._generate_pair_value(info)
                entry[split] = [self._generate_enc_graph(pairs, values, add_support=True),
                                self._generate_dec_graph(pairs), th.FloatTensor(values)]
            out[cv_idx] = entry
        return out

    @staticmethod
    def _generate_pair_value(rel_info):
        return ((np.asarray(rel_info['drug_id'], dtype=np.int64), np.asarray(rel_info['disease_id'], dtype=np.int64)),
                np.asarray(rel_info['values'], dtype=np.float32))

    def _generate_enc_graph(self, rating_pairs, rating_values, add_support=False):
        return GB.generate_enc_graph(rating_pairs, rating_values, self._num_drug, self._num_disease, self._device,
                                     symm=self._symm, add_support=add_support)

    def _generate_dec_graph(self, rating_pairs):
        return GB.generate_dec_graph(rating_pairs, self._num_drug, self._num_disease, self._device)

    # data_loader.py:511-582
    def augment_features(self):
        """Feature-level augmentation gated by --use_augmentation (noise, masking, optional mix-up)."""
        if not self.use_augmentation:
            return self.drug_feature, self.disease_feature
        from .augmentation import GraphAugmentation as GA
        p = self.aug_params
        out = []
        for feat in (self.drug_feature, self.disease_feature):
            f = GA.feature_masking(GA.feature_noise(feat.clone(), p.get('feature_noise_scale', 0.05)),
                                   p.get('feature_mask_rate', 0.1))
            out.append(GA.mix_up_features(f, p.get('mixup_alpha', 0.2)) if p.get('use_mixup', False) else f)
        return tuple(out)

    def get_graph_data_for_training(self, cv_idx):
        cv_data, graphs = self.data_cv[cv_idx], self.cv_specific_graphs[cv_idx]
        drug_feat, dis_feat = self.augment_features()
        dev = self._device
        return {'train_enc_graph': cv_data['train'][0].to(dev), 'train_dec_graph': cv_data['train'][1].to(dev),
                'train_labels': cv_data['train'][2].to(dev), 'test_enc_graph': cv_data['test'][0].to(dev),
                'test_dec_graph': cv_data['test'][1].to(dev), 'test_labels': cv_data['test'][2].to(dev),
                'drug_graph': graphs['drug_graph'].to(dev), 'disease_graph': graphs['disease_graph'].to(dev),
                'drug_feature_graph': graphs['drug_feature_graph'].to(dev),
                'disease_feature_graph': graphs['disease_feature_graph'].to(dev),
                'drug_features': drug_feat.to(dev), 'disease_features': dis_feat.to(dev),
                'drug_sim_features': th.as_tensor(self.drug_sim_features, dtype=th.float32).to(dev),
                'disease_sim_features': th.as_tensor(self.disease_sim_features, dtype=th.float32).to(dev)}

    @property
    def num_links(self):
        return len(np.unique(self.association_matrix))

    @property
    def num_disease(self):
        return self._num_disease

    @property
    def num_drug(self):
        return self._num_drug
