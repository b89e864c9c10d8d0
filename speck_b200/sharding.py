"""Multi-GPU host logic (the reference is single-GPU, so this has no reference file):
rows of C depend only on the matching rows of A and on all of B, so A is cut into contiguous row
blocks balanced by intermediate products, B is replicated once at setup (a broadcast over
NCCL/NVLink on GPUs, gloo in the CPU tests) and the slabs of C are concatenated; there is no
collective in the per-multiply path."""
import numpy as np
import torch
import torch.distributed as dist

from .matrices import HostCSR, product_balanced_cuts


def shard_rows(A: HostCSR, b_row_offsets, world, rank):
    """-> (cuts int64[world+1], this rank's slab of A with re-based row_offsets)."""
    cuts = product_balanced_cuts(A, np.asarray(b_row_offsets), world)
    return cuts, A.row_slice(int(cuts[rank]), int(cuts[rank + 1]))


def broadcast_csr(B, src=0, device="cpu"):
    """Broadcast a HostCSR from `src` to every rank (setup only).  On `device` = a CUDA device the
    three arrays travel GPU-to-GPU over NCCL; returns a HostCSR on CPU ranks, or a dict of device
    tensors (rows, cols, nnz, rp, ci, v) for CUDA devices."""
    rank = dist.get_rank()
    dev = torch.device(device)
    meta = torch.zeros(4, dtype=torch.int64, device=dev)
    if rank == src:   # the value type travels with the sizes: fp32 and fp64 are both instantiated
        meta = torch.tensor([B.rows, B.cols, B.nnz, np.dtype(B.data.dtype).itemsize], dtype=torch.int64, device=dev)
    dist.broadcast(meta, src)
    rows, cols, nnz, itemsize = (int(x) for x in meta.tolist())
    vdtype = torch.float32 if itemsize == 4 else torch.float64
    if rank == src:
        rp = torch.from_numpy(np.ascontiguousarray(B.row_offsets).view(np.int32)).to(dev)
        ci = torch.from_numpy(np.ascontiguousarray(B.col_ids).view(np.int32)).to(dev)
        v = torch.from_numpy(np.ascontiguousarray(B.data)).to(dev)
    else:
        rp = torch.empty(rows + 1, dtype=torch.int32, device=dev)
        ci = torch.empty(nnz, dtype=torch.int32, device=dev)
        v = torch.empty(nnz, dtype=vdtype, device=dev)
    for t in (rp, ci, v):
        dist.broadcast(t, src)
    if dev.type == "cpu":
        return HostCSR(rows, cols, rp.numpy().view(np.uint32), ci.numpy().view(np.uint32), v.numpy())
    return {"rows": rows, "cols": cols, "nnz": nnz, "rp": rp, "ci": ci, "v": v}


def concat_slabs(slabs):
    """Concatenate per-rank (row_offsets, col_ids, data) slabs into one CSR.  The single-matrix
    form keeps the API's u32 row_offsets, so the total nnz must stay below 2^32."""
    total = sum(int(rp[-1]) for rp, _, _ in slabs)
    if total >= 2 ** 32:
        raise OverflowError("concatenated nnz(C) does not fit u32 row_offsets; keep C distributed")
    out_rp = [np.zeros(1, np.uint32)]
    base = 0
    for rp, _, _ in slabs:
        out_rp.append((rp[1:].astype(np.int64) + base).astype(np.uint32))
        base += int(rp[-1])
    return (np.concatenate(out_rp), np.concatenate([ci for _, ci, _ in slabs]),
            np.concatenate([v for _, _, v in slabs]))
