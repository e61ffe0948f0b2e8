"""Context-free wrappers around the K1 kernels (CSR <-> dense on the device)."""
import torch

from . import _lib
from ._lib import check, ptr, stream_ptr


def expand_rows(csr, rows, width=None):
    """Dense float32 [B x width] CUDA tensor of the given rows of a DeviceCSR (width defaults to the matrix's
    column count; a larger one leaves zero columns on the right, e.g. for condition flags)."""
    B = int(rows.numel())
    width = csr.shape[1] if width is None else int(width)
    out = torch.empty((B, width), dtype=torch.float32, device=csr.device)
    with torch.cuda.device(csr.device):
        check(_lib.lib().b200vae_expand_rows_raw(ptr(csr.indptr), ptr(csr.indices), ptr(csr.values), ptr(rows),
                                                 B, width, ptr(out), stream_ptr()))
    return out


def dense_to_csr(dense):
    """(indptr int64 [B+1], indices int32 [nnz], values float32 [nnz]) device tensors of a dense
    float32 CUDA matrix.  One host sync to size the outputs."""
    B, n_items = dense.shape
    dev = dense.device
    with torch.cuda.device(dev):
        lens = torch.empty(B + 1, dtype=torch.int64, device=dev)
        indptr = torch.empty(B + 1, dtype=torch.int64, device=dev)
        check(_lib.lib().b200vae_dense_to_csr_raw(ptr(dense), B, n_items, ptr(lens), ptr(indptr), None, None, 0,
                                                  stream_ptr()))
        nnz = int(indptr[-1].item())
        indices = torch.empty(max(nnz, 1), dtype=torch.int32, device=dev)
        values = torch.empty(max(nnz, 1), dtype=torch.float32, device=dev)
        check(_lib.lib().b200vae_dense_to_csr_raw(ptr(dense), B, n_items, ptr(lens), ptr(indptr), ptr(indices),
                                                  ptr(values), max(nnz, 1), stream_ptr()))
    return indptr, indices[:nnz], values[:nnz]
