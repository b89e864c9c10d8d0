"""world_size-2 gloo test (CPU) of the multi-GPU host logic: contiguous row shards of A
balanced by products, B replicated by a broadcast at setup, every rank multiplies its slab with no
data-path collective, slabs concatenate to the single-device result.  The slab products are done
by the oracle here (no GPU in this suite); on the GPU box the same sharding code drives the CUDA
path (tests/test_gpu_sharded.py, bench.py --gpus N)."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

import oracle
from speck_b200 import matrices as M
from speck_b200.sharding import broadcast_csr, concat_slabs, shard_rows


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, out_dir):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    oracle.set_threads(1)
    A = M.rmat(10, 8, seed=31)                      # every rank can generate A; only rank 0 owns B
    B = broadcast_csr(A if rank == 0 else None, src=0, device="cpu")
    np.testing.assert_array_equal(B.col_ids, A.col_ids)
    cuts, slab = shard_rows(A, B.row_offsets, world, rank)
    rp, ci, v = oracle.spgemm(slab.row_offsets, slab.col_ids, slab.data, B.row_offsets, B.col_ids, B.data, B.cols)
    np.savez(os.path.join(out_dir, f"slab{rank}.npz"), rp=rp, ci=ci, v=v, cuts=cuts)
    # the only collective after setup: the per-slab nnz (8 bytes per rank)
    counts = [torch.zeros(1, dtype=torch.int64) for _ in range(world)]
    dist.all_gather(counts, torch.tensor([ci.size], dtype=torch.int64))
    assert int(counts[rank]) == ci.size
    dist.barrier()
    dist.destroy_process_group()


def test_two_rank_row_sharding_concatenates_to_full_product(tmp_path):
    world = 2
    mp.spawn(_worker, args=(world, _free_port(), str(tmp_path)), nprocs=world, join=True)
    A = M.rmat(10, 8, seed=31)
    slabs = [np.load(tmp_path / f"slab{r}.npz") for r in range(world)]
    rp, ci, v = concat_slabs([(s["rp"], s["ci"], s["v"]) for s in slabs])
    frp, fci, fv = oracle.spgemm(A.row_offsets, A.col_ids, A.data, A.row_offsets, A.col_ids, A.data, A.cols)
    np.testing.assert_array_equal(rp, frp)
    np.testing.assert_array_equal(ci, fci)
    np.testing.assert_array_equal(v, fv)
    cuts = slabs[0]["cuts"]
    assert cuts[0] == 0 and cuts[-1] == A.rows


def test_concat_rejects_u32_overflow():
    big = (np.array([0, 2 ** 31], np.uint32), np.zeros(0, np.uint32), np.zeros(0))
    with pytest.raises(OverflowError):
        concat_slabs([big, big, big])
