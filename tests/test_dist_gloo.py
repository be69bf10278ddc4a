"""CPU, world_size=2 over gloo: the only collectives on the path - naiveSyncBN statistics exchange -
and the frame sharding used by bench.py (no data-path collective)."""
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp


def _port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, out):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from sst_b200.norm import NaiveSyncBatchNorm1d
    torch.manual_seed(0)
    bn = NaiveSyncBatchNorm1d(8, eps=1e-3, momentum=0.01).train()
    g = torch.Generator().manual_seed(100)
    full = torch.randn(64, 8, generator=g)          # same on every rank; each rank takes an equal slice
    x = full[rank * 32:(rank + 1) * 32].clone().requires_grad_(True)
    y = bn(x)
    y.square().sum().backward()
    out[rank] = (y.detach(), x.grad.detach(), bn.running_mean.clone(), bn.running_var.clone())
    # frame sharding (bench.py): rank r owns frames r, r+world, ... ; totals are all-reduced for reporting only
    frames = torch.tensor([float(len(range(rank, 10, world)))])
    dist.all_reduce(frames)
    out[f"frames{rank}"] = frames.item()
    dist.destroy_process_group()


@pytest.mark.timeout(120)
def test_naive_sync_bn_matches_global_batchnorm():
    world = 2
    mgr = mp.Manager()
    out = mgr.dict()
    mp.spawn(_worker, args=(world, _port(), out), nprocs=world, join=True)
    g = torch.Generator().manual_seed(100)
    full = torch.randn(64, 8, generator=g).requires_grad_(True)
    torch.manual_seed(0)
    ref = torch.nn.BatchNorm1d(8, eps=1e-3, momentum=0.01).train()
    y = ref(full)
    y.square().sum().backward()
    got_y = torch.cat([out[0][0], out[1][0]])
    got_g = torch.cat([out[0][1], out[1][1]])
    torch.testing.assert_close(got_y, y.detach(), rtol=1e-4, atol=1e-5)
    torch.testing.assert_close(got_g, full.grad, rtol=1e-3, atol=1e-4)
    # the reference updates running_var with the biased variance (ops/norm.py:76-77); mean must agree exactly
    torch.testing.assert_close(out[0][2], ref.running_mean, rtol=1e-5, atol=1e-6)
    torch.testing.assert_close(out[0][2], out[1][2])
    assert out["frames0"] == 10.0 and out["frames1"] == 10.0


def _ddp_worker(rank, world, port, out):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from sst_b200.dist_utils import allreduce_grads_flat, max_over_ranks, shard_range
    torch.manual_seed(0)
    model = torch.nn.Sequential(torch.nn.Linear(6, 5), torch.nn.GELU(), torch.nn.Linear(5, 3))
    g = torch.Generator().manual_seed(7)
    frames = torch.randn(7, 4, 6, generator=g)            # 7 "frames" of 4 rows: uneven shard (4 + 3)
    lo, hi = shard_range(7, rank, world)
    loss = sum(model(frames[i]).square().mean() for i in range(lo, hi)) / 7 * world   # mean over ALL frames after averaging
    loss.backward()
    n = allreduce_grads_flat(model.parameters())
    out[rank] = ([p.grad.clone() for p in model.parameters()], (lo, hi), n, max_over_ranks(1.0 + rank))
    dist.destroy_process_group()


@pytest.mark.timeout(120)
def test_frame_sharding_and_flat_gradient_allreduce():
    """Training-side collective of SURVEY 8e(1) on 2 gloo ranks: contiguous frame shards + ONE flat gradient all-reduce give the
    single-process gradient of the whole batch; timings reduce with MAX."""
    world = 2
    mgr = mp.Manager()
    out = mgr.dict()
    mp.spawn(_ddp_worker, args=(world, _port(), out), nprocs=world, join=True)
    assert out[0][1] == (0, 4) and out[1][1] == (4, 7)
    torch.manual_seed(0)
    model = torch.nn.Sequential(torch.nn.Linear(6, 5), torch.nn.GELU(), torch.nn.Linear(5, 3))
    g = torch.Generator().manual_seed(7)
    frames = torch.randn(7, 4, 6, generator=g)
    (sum(model(frames[i]).square().mean() for i in range(7)) / 7).backward()
    for r in range(world):
        for got, p in zip(out[r][0], model.parameters()):
            torch.testing.assert_close(got, p.grad, rtol=1e-5, atol=1e-7)
        assert out[r][2] == sum(p.numel() for p in model.parameters())
        assert out[r][3] == 2.0
