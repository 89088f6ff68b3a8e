"""Shape / broadcast helpers with the names and behaviour of the reference's `qpth/util.py`.

Only `expandParam` / `extract_nBatch` are on the hot path (they define how un-batched parameters broadcast, rows a3 of
SURVEY.md section 8); the rest is kept so that `from qpth.util import ...` keeps working after
`qpth_b200.install_as_qpth()`.
"""
import numpy as np
import torch

_PARAM_RANKS = (3, 2, 3, 2, 3, 2)      # batched rank of Q, p, G, h, A, b


def extract_nBatch(Q, p, G, h, A, b):
    """Batch size: size(0) of the first argument given with its batched rank, 1 if none is (util.py:53-59)."""
    return next((int(t.size(0)) for t, rank in zip((Q, p, G, h, A, b), _PARAM_RANKS) if t.ndimension() == rank), 1)


def expandParam(X, nBatch, nDim):
    """(tensor seen as a batch, was_unbatched) (util.py:44-50).

    0-dim and empty tensors and tensors that already have `nDim` dimensions pass through; one dimension short means
    "shared by the batch" and becomes a stride-0 view; anything else is the reference's RuntimeError.
    """
    have = X.ndimension()
    if X.nelement() == 0 or have == 0 or have == nDim:
        return X, False
    if have + 1 != nDim:
        raise RuntimeError("Unexpected number of dimensions.")
    return X.unsqueeze(0).expand(nBatch, *X.shape), True


def check_shapes(Q, p, G, h, A, b):
    """Validate every trailing dimension and batch size BEFORE raw pointers and strides go to the kernels.

    The reference gets these errors for free from `bmm` / `expand` (a torch shape RuntimeError somewhere inside
    `pre_factor_kkt` or `forward`); here nothing downstream would notice - the kernels would read out of bounds.
    Returns (nBatch, nz, nineq, neq). Rank errors are `expandParam`'s ("Unexpected number of dimensions.").
    """
    nBatch = extract_nBatch(Q, p, G, h, A, b)
    for X, nd in zip((Q, p, G, h, A, b), _PARAM_RANKS):
        expandParam(X, nBatch, nd)

    def fail(msg):
        raise RuntimeError("qpth_b200: inconsistent shapes: " + msg)

    if Q.nelement() == 0 or Q.dim() < 2:
        fail("Q must be (nBatch, nz, nz) or (nz, nz), got %s" % (tuple(Q.shape),))
    nz = int(Q.size(-1))
    if Q.size(-2) != nz:
        fail("Q is not square: %s" % (tuple(Q.shape),))
    nineq = int(G.size(-2)) if (G.nelement() > 0 and G.dim() >= 2) else 0
    neq = int(A.size(-2)) if (A.nelement() > 0 and A.dim() >= 2) else 0
    want = (("p", p, (nz,), True), ("G", G, (nineq, nz), nineq > 0), ("h", h, (nineq,), nineq > 0),
            ("A", A, (neq, nz), neq > 0), ("b", b, (neq,), neq > 0))
    for name, X, trail, needed in want:
        if not needed:
            if X.nelement() != 0:
                fail("%s given %s but its constraint block is empty" % (name, tuple(X.shape)))
            continue
        if X.nelement() == 0 or tuple(X.shape[-len(trail):]) != trail:
            fail("%s has shape %s, expected trailing dimensions %s" % (name, tuple(X.shape), trail))
    for name, X, rank in zip("QpGhAb", (Q, p, G, h, A, b), _PARAM_RANKS):
        if X.nelement() > 0 and X.dim() == rank and X.size(0) != nBatch:
            fail("%s has batch size %d, the batch is %d" % (name, X.size(0), nBatch))
    return nBatch, nz, nineq, neq


def get_sizes(G, A=None):
    """(nineq, nz, neq, nBatch) from G (2-D or 3-D) and optionally A; neq is None when A is not given (util.py:22-33)."""
    if G.dim() not in (2, 3):
        raise RuntimeError("Unexpected number of dimensions.")
    nBatch = G.size(0) if G.dim() == 3 else 1
    nineq, nz = G.shape[-2], G.shape[-1]
    neq = None if A is None else (A.size(1) if A.nelement() > 0 else 0)
    return nineq, nz, neq, nBatch


def bger(x, y):
    """Batched outer product: (B,k),(B,l) -> (B,k,l) (util.py:18-19)."""
    return x[:, :, None] * y[:, None, :]


def bdiag(d):
    """(B,n) -> (B,n,n) with d on the diagonals (util.py:36-41)."""
    return torch.diag_embed(d)


def to_np(t):
    """Tensor -> numpy (None stays None, empty becomes an empty array) (util.py:9-15)."""
    if t is None:
        return None
    return np.array([]) if t.nelement() == 0 else t.detach().cpu().numpy()


def print_header(msg):
    print('===>', msg)


def copy_lower_(dst, src, band=20):
    """dst[..., lower] = src[..., lower] for a batch of SYMMETRIC matrices (B, n, n), between a (pinned) host tensor
    and a device tensor, on the current CUDA stream: only the lower triangle crosses PCIe, as `band`-row strips
    (C ABI `qpb200_copy_lower`). What lies above the strips in `dst` is left as it was. Used for Q on its way to the
    device and for dQ on its way back (both symmetric: qp.py:81-85 requires an SPD Q, qp.py:157-158 symmetrises dQ)."""
    import ctypes
    from . import _lib
    assert dst.shape == src.shape and dst.dim() == 3 and dst.size(1) == dst.size(2)
    assert dst.dtype == torch.float64 and src.dtype == torch.float64 and dst.is_contiguous() and src.is_contiguous()
    assert dst.is_cuda != src.is_cuda, "one side on the host, one on the device"
    dev = dst.device if dst.is_cuda else src.device
    with torch.cuda.device(dev):
        st = ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)
        _lib.check(_lib.load().qpb200_copy_lower(ctypes.c_void_p(src.data_ptr()), ctypes.c_void_p(dst.data_ptr()),
                                                 int(dst.size(0)), int(dst.size(1)), int(band), 1 if src.is_cuda else 0, st))
    return dst
