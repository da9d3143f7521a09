"""N>1 host logic on CPU: two gloo ranks each fill only the tiles they own (tile k -> rank k % N) and the
frame is gathered to rank 0 with a SUM reduce - exact because tile supports are disjoint.  Pixel values come from the
oracle so no GPU is needed; the same partition + reduce is what bench.py runs over NCCL."""
import os
import socket
import sys

import numpy as np
import pytest

import oracle_api as orc

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, xml, out_path):
    import torch
    import torch.distributed as dist
    sys.path.insert(0, ROOT)
    import __graft_entry__ as ge
    rt = ge.load_package()
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    hs = rt.HostScene(xml)
    cam = hs.camera(0)
    full = orc.oracle_render(hs.desc, cam, threads=2)
    mask = rt.tiling.tile_mask(cam.nx, cam.ny, rank, world)
    assert mask.sum() > 0
    part = np.where(mask[..., None], full, 0.0).astype(np.float32)
    # the same helper bench.py uses at N>1, on its reduce path (gloo has no peer memory): frame buffer + finish()
    gather = rt.multigpu.FrameGather(part.size, torch.device("cpu"), mode="nccl")
    assert gather.mode == "nccl" and gather.resolve_owned_only == 0 and gather.out_ptr == gather.frame.data_ptr()
    gather.frame.copy_(torch.from_numpy(part).reshape(-1))
    gather.finish()
    t = gather.frame.reshape(part.shape)
    cover = torch.from_numpy(mask.astype(np.int32))
    dist.reduce(cover, dst=0, op=dist.ReduceOp.SUM)
    if rank == 0:
        assert int(cover.min()) == 1 and int(cover.max()) == 1          # every pixel owned by exactly one rank
        assert np.array_equal(t.numpy(), full)                           # exact gather
        np.save(out_path, t.numpy())
    dist.destroy_process_group()


def test_two_rank_tile_gather_gloo(rt, scene_dir, tmp_path):
    import torch.multiprocessing as mp
    xml = rt.scenes.scene_c1(os.path.join(scene_dir, "c1gloo"), res=72)
    out = str(tmp_path / "frame.npy")
    mp.start_processes(_worker, args=(2, _free_port(), xml, out), nprocs=2, join=True, start_method="spawn")
    assert os.path.exists(out)


def test_tile_owner_properties(rt):
    t = rt.tiling
    for (nx, ny, n) in [(1920, 1080, 8), (800, 800, 4), (33, 17, 2), (16, 16, 3)]:
        own = t.tile_owner(nx, ny, n)
        assert own.shape == (ny, nx) and own.min() == 0 and own.max() <= n - 1
        counts = [t.local_tile_count(nx, ny, r, n) for r in range(n)]
        assert sum(counts) == ((nx + 15) // 16) * ((ny + 15) // 16)
        assert max(counts) - min(counts) <= 1                              # interleaving balances tile counts
