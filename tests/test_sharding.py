"""N > 1 path.  CPU: world_size-2 gloo processes exercise the host-side plumbing of candidate sharding (ranges, padding,
wire format, gather) with the filter injected from the oracle.  GPU: the same sharder with the real filter and the
library's own NCCL communicator on however many GPUs the box has (torchrun-free: spawned processes)."""
import os
import socket
from types import SimpleNamespace

import numpy as np
import pytest
import torch
import torch.multiprocessing as mp

from oracle import oracle


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def test_shard_range_covers_everything():
    from face_detection_and_recognition_b200.sharding import shard_range
    for n in (0, 1, 7, 16, 1000, 1001, 12345):
        for world in (1, 2, 3, 8):
            covered = []
            m_locals = set()
            for r in range(world):
                a, b, m = shard_range(n, world, r)
                assert 0 <= a <= b <= n and b - a <= m
                covered += list(range(a, b))
                m_locals.add(m)
            assert covered == list(range(n)) and len(m_locals) == 1
    with pytest.raises(ValueError):
        shard_range(10, 2, 2)


def test_pack_unpack_roundtrip():
    from face_detection_and_recognition_b200.sharding import pack_results, shard_range, unpack_results
    rng = np.random.default_rng(0)
    n, world = 1003, 3
    keep = torch.from_numpy(rng.integers(0, 2, n).astype(np.uint8))
    idx = torch.from_numpy(rng.integers(0, 1 << 30, n).astype(np.int32))
    bufs = []
    for r in range(world):
        a, b, m = shard_range(n, world, r)
        bufs.append(pack_results(keep[a:b], idx[a:b], m))
    k2, i2 = unpack_results(torch.cat(bufs), world, m, n)
    assert torch.equal(k2, keep) and torch.equal(i2, idx)


def _oracle_filter(ref, cand, thr, metric="cosine"):
    k, i, v = oracle.filter_cosine(ref.numpy(), cand.numpy(), thr)
    return SimpleNamespace(keep=torch.from_numpy(k), best_idx=torch.from_numpy(i), best_val=torch.from_numpy(v))


def _gloo_worker(rank, world, port, n_cand, out_dir):
    import torch.distributed as dist
    from face_detection_and_recognition_b200.sharding import CandidateSharder
    dist.init_process_group("gloo", init_method=f"tcp://127.0.0.1:{port}", rank=rank, world_size=world)
    ref, cand = oracle.make_synthetic(64, n_cand, 32, seed=11)
    sh = CandidateSharder(rank, world, gather="dist", filter_fn=_oracle_filter)
    a, b, _ = sh.local_range(n_cand)
    keep, idx, _ = sh.filter(torch.from_numpy(ref), torch.from_numpy(cand[a:b]), 0.5, n_cand)
    np.savez(os.path.join(out_dir, f"r{rank}.npz"), keep=keep.numpy(), idx=idx.numpy())
    dist.destroy_process_group()


@pytest.mark.parametrize("n_cand", [1001, 64, 3])
def test_candidate_sharding_gloo_world2(tmp_path, n_cand):
    world = 2
    mp.spawn(_gloo_worker, args=(world, _free_port(), n_cand, str(tmp_path)), nprocs=world, join=True)
    ref, cand = oracle.make_synthetic(64, n_cand, 32, seed=11)
    ko, io, _ = oracle.filter_cosine(ref, cand, 0.5)
    for r in range(world):
        g = np.load(tmp_path / f"r{r}.npz")
        assert np.array_equal(g["keep"], ko) and np.array_equal(g["idx"], io), f"rank {r}"


def _nccl_worker(rank, world, port, n_ref, n_cand, dim, out_dir):
    import torch.distributed as dist
    from face_detection_and_recognition_b200.sharding import CandidateSharder
    torch.cuda.set_device(rank)
    dist.init_process_group("gloo", init_method=f"tcp://127.0.0.1:{port}", rank=rank, world_size=world)
    ref, cand = oracle.make_synthetic(n_ref, n_cand, dim, seed=13, n_adversarial=50, n_dup_refs=8)
    sh = CandidateSharder(rank, world, device=rank, gather="nccl")
    a, b, _ = sh.local_range(n_cand)
    keep, idx, _ = sh.filter(torch.from_numpy(ref).cuda(), torch.from_numpy(cand[a:b]).cuda(), 0.5, n_cand)
    torch.cuda.synchronize()
    np.savez(os.path.join(out_dir, f"r{rank}.npz"), keep=keep.cpu().numpy(), idx=idx.cpu().numpy())
    sh.close()
    dist.destroy_process_group()


@pytest.mark.gpu
def test_candidate_sharding_nccl(tmp_path, ffr_lib, cuda_dev):
    """Real filter + the library's NCCL allgather (K4), one process per visible GPU (a single-GPU box runs world 1...
    plus a 2-rank-on-1-GPU run is deliberately NOT attempted: NCCL refuses duplicate devices)."""
    world = min(torch.cuda.device_count(), 8)
    n_ref, n_cand, dim = 300, 20_003, 128
    mp.spawn(_nccl_worker, args=(world, _free_port(), n_ref, n_cand, dim, str(tmp_path)), nprocs=world, join=True)
    ref, cand = oracle.make_synthetic(n_ref, n_cand, dim, seed=13, n_adversarial=50, n_dup_refs=8)
    ko, io, so = oracle.filter_cosine(ref, cand, 0.5)
    _, _, s64 = oracle.filter_cosine(ref, cand, 0.5, dtype=np.float64)
    far = np.abs(s64 - 0.5) > 1e-6
    for r in range(world):
        g = np.load(tmp_path / f"r{r}.npz")
        assert np.array_equal(g["keep"][far], ko[far]), f"rank {r}"
        assert np.mean(g["idx"] == io) > 0.999, f"rank {r}"


@pytest.mark.gpu
def test_allgather_pack_unpack_single_rank(ffr_lib, cuda_dev):
    """K4 pack -> ncclAllGather -> unpack on a 1-rank communicator reproduces the inputs (odd length: padding path)."""
    from face_detection_and_recognition_b200 import ops
    rg = ops.ResultGather(0, 1, 0)
    keep = torch.randint(0, 2, (1237,), dtype=torch.uint8, device="cuda")
    idx = torch.randint(0, 1 << 30, (1237,), dtype=torch.int32, device="cuda")
    k2, i2 = rg.all_gather(keep, idx)
    torch.cuda.synchronize()
    assert torch.equal(k2, keep) and torch.equal(i2, idx)
    rg.close()


@pytest.mark.gpu
def test_allgather_inplace_single_rank(ffr_lib, cuda_dev):
    """K4, in-place form: the filter's outputs ARE this rank's slice of the gathered arrays; one grouped NCCL launch, no
    pack / unpack kernels of our own (launch counter unchanged by the gather)."""
    from face_detection_and_recognition_b200 import ops
    rg = ops.ResultGather(0, 1, 0)
    keep_all, idx_all, keep_mine, idx_mine = rg.buffers(1237)
    ref, cand = oracle.make_synthetic(300, 1237, 128, seed=2)
    res = ops.face_filter(torch.from_numpy(ref).cuda(), torch.from_numpy(cand).cuda(), 0.5, out=(keep_mine, idx_mine, torch.empty(1237, device="cuda")))
    l0 = ops.launch_count()
    k2, i2 = rg.all_gather_inplace(keep_all, idx_all)
    torch.cuda.synchronize()
    assert ops.launch_count() == l0
    ko, io, _ = oracle.filter_cosine(ref, cand, 0.5)
    assert np.mean(i2.cpu().numpy() == io) > 0.999 and k2.data_ptr() == res.keep.data_ptr()
    rg.close()
