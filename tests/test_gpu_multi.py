"""Two real GPUs (skipped on a one-GPU box): the p2p frame gather - every rank's resolve kernel storing its own tiles into
rank 0's framebuffer over NVLink peer memory (multigpu.FrameGather, RtRenderOpts.resolve_owned_only) - against the
single-GPU frame, and against the NCCL reduce fallback."""
import os
import socket
import sys

import numpy as np
import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, xml, out_dir):
    import torch
    import torch.distributed as dist
    sys.path.insert(0, ROOT)
    import __graft_entry__ as ge
    rt = ge.load_package()
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    os.environ.setdefault("NCCL_DEBUG", "WARN")
    torch.cuda.set_device(rank)
    dev = torch.device("cuda", rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=dev)
    hs = rt.HostScene(xml)
    cam = hs.camera(0)
    sc = rt.Scene(hs.desc, rank)
    stream = torch.cuda.current_stream()
    for mode in ("p2p", "nccl"):
        g = rt.multigpu.FrameGather(cam.nx * cam.ny * 3, dev, mode=mode)
        if rank == 0:
            g.frame.fill_(float("nan"))                                   # every pixel must be written by exactly one rank
        dist.barrier(); torch.cuda.synchronize()
        if g.mode == "nccl":
            g.frame.zero_()
        o = rt.default_opts(tile_rank=rank, tile_count=world, resolve_owned_only=g.resolve_owned_only)
        sc.render_device(cam, o, g.out_ptr, stream.cuda_stream)
        g.finish()
        torch.cuda.synchronize(); dist.barrier()
        if rank == 0:
            np.save(os.path.join(out_dir, f"{mode}_{g.mode}.npy"), g.frame.cpu().numpy().reshape(cam.ny, cam.nx, 3))
    # two DIFFERENT frames back to back, rank 0 reading each with a non-blocking copy: FrameGather.release() must keep frame n+1's peer
    # stores out of the buffer until frame n has been read (write-after-read across ranks)
    import copy
    cams = [cam, copy.copy(cam)]
    cams[1] = rt.RtCameraDesc.from_buffer_copy(cam)
    cams[1].position[0] += 1.5; cams[1].position[1] += 0.7
    g = rt.multigpu.FrameGather(cam.nx * cam.ny * 3, dev, mode="p2p")
    host = [torch.empty(cam.nx * cam.ny * 3, dtype=torch.float32).pin_memory() for _ in cams]
    o = rt.default_opts(tile_rank=rank, tile_count=world, resolve_owned_only=g.resolve_owned_only)
    for rep in range(3):
        for k, c in enumerate(cams):
            if g.mode == "nccl":
                g.frame.zero_()
            sc.render_device(c, o, g.out_ptr, stream.cuda_stream)
            g.finish()
            if rank == 0:
                host[k].copy_(g.frame, non_blocking=True)
            g.release()
    torch.cuda.synchronize(); dist.barrier()
    if rank == 0:
        for k, c in enumerate(cams):
            np.save(os.path.join(out_dir, f"seq{k}.npy"), host[k].numpy().reshape(cam.ny, cam.nx, 3))
            alone_k = torch.empty(cam.nx * cam.ny * 3, dtype=torch.float32, device=dev)
            sc.render_device(c, rt.default_opts(), alone_k.data_ptr(), stream.cuda_stream)
            torch.cuda.synchronize()
            np.save(os.path.join(out_dir, f"seq{k}_alone.npy"), alone_k.cpu().numpy().reshape(cam.ny, cam.nx, 3))
    if rank == 0:
        alone = torch.empty(cam.nx * cam.ny * 3, dtype=torch.float32, device=dev)
        sc.render_device(cam, rt.default_opts(), alone.data_ptr(), stream.cuda_stream)
        torch.cuda.synchronize()
        np.save(os.path.join(out_dir, "alone.npy"), alone.cpu().numpy().reshape(cam.ny, cam.nx, 3))
    dist.barrier()
    sc.close(); hs.close()
    dist.destroy_process_group()


def test_two_gpu_p2p_gather_matches_single_gpu(rt, scene_dir, tmp_path):
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs two GPUs")
    import torch.multiprocessing as mp
    xml = rt.scenes.scene_c3(os.path.join(scene_dir, "c3p2p"), res=(203, 117), spp=4, n=48)
    mp.start_processes(_worker, args=(2, _free_port(), xml, str(tmp_path)), nprocs=2, join=True, start_method="spawn")
    alone = np.load(tmp_path / "alone.npy")
    for k in range(2):
        assert np.allclose(np.load(tmp_path / f"seq{k}.npy"), np.load(tmp_path / f"seq{k}_alone.npy"), rtol=1e-5, atol=1e-5), k
    assert not np.allclose(np.load(tmp_path / "seq0.npy"), np.load(tmp_path / "seq1.npy"), rtol=1e-2, atol=1e-2)
    files = sorted(f for f in os.listdir(tmp_path) if f.endswith(".npy") and f != "alone.npy" and not f.startswith("seq"))
    assert any(f.startswith("p2p_") for f in files) and "nccl_nccl.npy" in files
    for f in files:
        got = np.load(tmp_path / f)
        assert np.isfinite(got).all(), f
        assert np.allclose(got, alone, rtol=1e-5, atol=1e-5), f


# ------------------------------------------------------------------------------------------------ multi-GPU behind the C ABI (rt_group_*)
def _group_test_exe():
    return os.path.join(ROOT, "raytracer-795_b200", "bin", "group_render_test")


@pytest.mark.parametrize("no_peer", [False, True])
def test_group_render_c_program(rt, scene_dir, no_peer):
    """tests/cpp/group_render_test.cpp (plain C ABI, no Python / torch): rt_group_render over every device of the box (on a one-GPU
    box: two replicas on device 0, which still exercises the worker threads, the tile partition and the owned-tile stores into one
    frame) equals rt_render on one device to 1e-5; per-rank tile options are rejected.  no_peer: every replica but the root takes
    the copy-and-merge path a device without peer access gets."""
    import subprocess
    exe = _group_test_exe()
    assert os.path.exists(exe), "build it: make -C raytracer-795_b200"
    import torch
    ndev = torch.cuda.device_count()
    devs = [str(i) for i in range(ndev)] if ndev >= 2 else ["0", "0"]
    xml = rt.scenes.scene_c2(os.path.join(scene_dir, "c2grp"), res=203, spp=4)      # ragged tiles + branching rays
    env = dict(os.environ)
    if no_peer:
        env["RT795_GROUP_NO_PEER"] = "1"
    r = subprocess.run([exe, xml] + devs, capture_output=True, text=True, env=env)
    print(r.stdout.strip())
    assert r.returncode == 0, (r.stdout, r.stderr[-2000:])
    assert f"devices={len(devs)}" in r.stdout and "rejected_tile_opts=1" in r.stdout
    if no_peer:
        assert "peer=1 " in r.stdout              # only the root writes directly


def test_group_render_binding_matches_single_device(rt, scene_dir):
    """The same through the ctypes binding, three replicas (or every GPU of the box), C4-style instanced scene with motion blur."""
    import torch
    ndev = torch.cuda.device_count()
    devs = list(range(ndev)) if ndev >= 2 else [0, 0, 0]
    xml = rt.scenes.scene_c4(os.path.join(scene_dir, "c4grp"), res=(256, 144), spp=4, n=40)
    hs = rt.HostScene(xml)
    cam = hs.camera(0)
    sc = rt.Scene(hs.desc, 0)
    alone = sc.render(cam)
    sc.close()
    g = rt.SceneGroup(hs.desc, devs)
    a = g.render(cam)
    b = g.render(cam)                                                          # a second frame reuses workers, frame and queues
    assert np.allclose(a, alone, rtol=1e-5, atol=1e-5) and np.allclose(b, alone, rtol=1e-5, atol=1e-5)
    st = [g.stats(i) for i in range(len(devs))]
    assert sum(s.rays_primary for s in st) == cam.nx * cam.ny * cam.num_samples     # the replicas partition the frame exactly
    assert all(s.rays_primary > 0 for s in st)
    with pytest.raises(rt.RtError):
        g.render(cam, rt.default_opts(tile_rank=0, tile_count=2))
    g.close(); hs.close()
