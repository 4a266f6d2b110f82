"""Host-side glue with the reference's names and semantics (utils.py in the reference)."""
import random

import numpy as np
import torch as th
import torch.nn as nn


def get_activation(act):
    """String -> activation module, same table as the reference (utils.py:47-80)."""
    if act is None:
        return lambda x: x
    if not isinstance(act, str):
        return act
    table = {'leaky': lambda: nn.LeakyReLU(0.1), 'relu': nn.ReLU, 'tanh': nn.Tanh, 'sigmoid': nn.Sigmoid,
             'softsign': nn.Softsign, 'gelu': nn.GELU, 'elu': nn.ELU, 'selu': nn.SELU}
    if act not in table:
        raise NotImplementedError(act)
    return table[act]()


def to_etype_name(rating):
    """utils.py:83-84."""
    return str(rating).replace('.', '_')


def _gram_eligible(emb1, emb2):
    import os
    from . import ops
    return (os.environ.get('DG_COMMON_LOSS', 'fused') != 'torch' and emb1.is_cuda and emb2.is_cuda
            and emb1.dtype == th.float32 and emb2.dtype == th.float32 and emb1.dim() == 2 and emb1.shape == emb2.shape
            and emb1.shape[0] > 1 and ops._rows_ok(emb1) and ops._rows_ok(emb2))


def common_loss(emb1, emb2):
    """Covariance-difference loss (utils.py:87-95): mean((Z1 Z1^T - Z2 Z2^T)^2) over centred, row-normalised
    embeddings. On CUDA the value is computed through the Frobenius identity
    ||Z1 Z1^T - Z2 Z2^T||_F^2 = ||Z1^T Z1||_F^2 + ||Z2^T Z2||_F^2 - 2 ||Z1^T Z2||_F^2 (float64 accumulation, the N x N
    products never materialised) by the explicit-kernel function `ops.GramCommonLoss`; `common_loss_dense` is the
    reference's literal expression (CPU tensors, DG_COMMON_LOSS=torch, and what the tests compare against)."""
    if _gram_eligible(emb1, emb2):
        from . import ops
        return ops.gram_common_loss(emb1, emb2)
    return common_loss_dense(emb1, emb2)


def common_loss_dense(emb1, emb2):
    """utils.py:87-95 as written."""
    emb1 = th.nn.functional.normalize(emb1 - th.mean(emb1, dim=0, keepdim=True), p=2, dim=1)
    emb2 = th.nn.functional.normalize(emb2 - th.mean(emb2, dim=0, keepdim=True), p=2, dim=1)
    return th.mean((emb1 @ emb1.t() - emb2 @ emb2.t()) ** 2)


def common_loss_gram(emb1, emb2):
    """Same value as `common_loss` without the N x N matrices:
    ||Z1 Z1^T - Z2 Z2^T||_F^2 = ||Z1^T Z1||_F^2 + ||Z2^T Z2||_F^2 - 2 ||Z1^T Z2||_F^2.
    Used only where N x N does not fit (the synthetic 100k-node shapes); accumulates in float64. Runs as the
    explicit-kernel autograd function `ops.GramCommonLoss` where the layout allows (DG_COMMON_LOSS=torch keeps the
    traced torch expression below, which is also what the tests compare it with)."""
    if _gram_eligible(emb1, emb2):
        from . import ops
        return ops.gram_common_loss(emb1, emb2)
    return common_loss_gram_torch(emb1, emb2)


def common_loss_gram_torch(emb1, emb2):
    """The Gram form written with torch ops (autograd-traced)."""
    n = emb1.shape[0]
    z1 = th.nn.functional.normalize(emb1 - th.mean(emb1, dim=0, keepdim=True), p=2, dim=1).double()
    z2 = th.nn.functional.normalize(emb2 - th.mean(emb2, dim=0, keepdim=True), p=2, dim=1).double()
    g11, g22, g12 = z1.t() @ z1, z2.t() @ z2, z1.t() @ z2
    return (((g11 ** 2).sum() + (g22 ** 2).sum() - 2.0 * (g12 ** 2).sum()) / float(n) / float(n)).float()


def setup_seed(seed):
    """utils.py:98-103."""
    th.manual_seed(seed)
    th.cuda.manual_seed_all(seed)
    np.random.seed(seed)
    random.seed(seed)
    th.backends.cudnn.deterministic = True
