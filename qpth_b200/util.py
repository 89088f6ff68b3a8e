"""Shape / broadcast helpers with the reference's names and behaviour (`qpth/util.py`)."""
import torch


def bger(x, y):
    """Batched outer product (util.py:18-19)."""
    return x.unsqueeze(2) * y.unsqueeze(1)


def get_sizes(G, A=None):
    """(nineq, nz, neq, nBatch) (util.py:22-33)."""
    if G.dim() == 2:
        nineq, nz = G.size()
        nBatch = 1
    elif G.dim() == 3:
        nBatch, nineq, nz = G.size()
    else:
        raise RuntimeError("Unexpected number of dimensions.")
    neq = None
    if A is not None:
        neq = A.size(1) if A.nelement() > 0 else 0
    return nineq, nz, neq, nBatch


def expandParam(X, nBatch, nDim):
    """Un-batched -> stride-0 batch view and a flag saying so (util.py:44-50)."""
    if X.ndimension() in (0, nDim) or X.nelement() == 0:
        return X, False
    elif X.ndimension() == nDim - 1:
        return X.unsqueeze(0).expand(*([nBatch] + list(X.size()))), True
    else:
        raise RuntimeError("Unexpected number of dimensions.")


def extract_nBatch(Q, p, G, h, A, b):
    """Batch size = size(0) of the first fully-batched argument, else 1 (util.py:53-59)."""
    dims = [3, 2, 3, 2, 3, 2]
    params = [Q, p, G, h, A, b]
    for param, dim in zip(params, dims):
        if param.ndimension() == dim:
            return param.size(0)
    return 1
