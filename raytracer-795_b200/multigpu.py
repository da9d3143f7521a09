"""Multi-GPU frame gather for the tile-interleaved partition (replaces the reference's 8 row-band threads,
src/Scene.cpp:285-312, at process granularity: one process per GPU, tile k -> rank k % N).

Two ways to land every rank's tiles in the root GPU's framebuffer:

  p2p   the root's framebuffer is a symmetric-memory allocation mapped into every rank (NVLink / NVSwitch peer memory).
        Each rank calls rt_render_device with `out` = the root's buffer and RtRenderOpts.resolve_owned_only = 1, so its
        resolve kernel (K8) stores its own tiles straight into the root's HBM: the resolve IS the gather, 1/N of the frame
        per rank crosses NVLink once, no reduction and no zero traffic.  A device-side barrier on the stream orders the
        root's later reads after everybody's stores.
  nccl  every rank resolves a full-size frame with zeros outside its tiles and the frames are SUM-reduced to the root
        (exact: x + 0 = x).  Moves N full frames; kept as the fallback when peer mapping is unavailable.

torch is plumbing here (device memory, the process group, the symmetric-memory rendezvous); the kernels are the library's.
"""
from __future__ import annotations

import os
import sys


class FrameGather:
    def __init__(self, npix3: int, device, mode: str | None = None):
        import torch
        import torch.distributed as dist

        self.torch, self.dist = torch, dist
        self.world = dist.get_world_size() if dist.is_initialized() else 1
        self.rank = dist.get_rank() if dist.is_initialized() else 0
        self.device = device
        self.npix3 = npix3
        self.hdl = None
        mode = (mode or os.environ.get("RT795_GATHER", "p2p")).lower()
        if mode not in ("p2p", "nccl"):
            raise ValueError(f"RT795_GATHER must be p2p or nccl, got {mode!r}")
        if self.world == 1:
            mode = "local"
        self.why_not_p2p = None
        if mode == "p2p":
            try:
                import torch.distributed._symmetric_memory as symm_mem

                self.local = symm_mem.empty(npix3, dtype=torch.float32, device=device)
                self.local.zero_()
                self.hdl = symm_mem.rendezvous(self.local, dist.group.WORLD)
                self.root = self.hdl.get_buffer(0, (npix3,), torch.float32)
            except Exception as e:  # peer mapping refused (no P2P, container limits): fall back to the NCCL reduce
                self.why_not_p2p = f"{type(e).__name__}: {e}".splitlines()[0][:200]
                print(f"[rt795] rank {self.rank}: p2p frame gather unavailable ({self.why_not_p2p}); using NCCL reduce", file=sys.stderr)
                mode = "nccl"
            # every rank must take the same path
            flag = torch.tensor([1 if mode == "p2p" else 0], dtype=torch.int32, device=device)
            dist.all_reduce(flag, op=dist.ReduceOp.MIN)
            if int(flag.item()) == 0:
                mode = "nccl"
        if mode != "p2p":
            self.hdl = None
            self.local = torch.zeros(npix3, dtype=torch.float32, device=device)
            self.root = self.local
        self.mode = mode

    @property
    def frame(self):
        """The gathered frame (valid on rank 0 after finish())."""
        return self.local

    @property
    def out_ptr(self) -> int:
        return (self.root if self.mode == "p2p" else self.local).data_ptr()

    @property
    def resolve_owned_only(self) -> int:
        return 1 if self.mode == "p2p" else 0

    def finish(self):
        """Enqueue, on the current stream, whatever completes the gather after this rank's rt_render_device returned.
        Contract (p2p): after finish() rank 0 may READ the frame on this stream; before ANY rank renders the next frame into it, every
        rank must call release() after rank 0 has enqueued its reads (copy to the host, tone map ...) on the stream - otherwise a fast
        peer's stores of frame n+1 land in a frame rank 0 is still reading (write after read; invisible only while consecutive frames
        are identical)."""
        if self.mode == "p2p":
            self.hdl.barrier()                        # device-side signal-pad barrier: all ranks' peer stores have landed
        elif self.mode == "nccl":
            self.dist.reduce(self.local, dst=0, op=self.dist.ReduceOp.SUM)

    def release(self):
        """Second half of the frame hand-over: enqueued by every rank after rank 0's reads of the gathered frame were enqueued on
        its stream.  p2p: a second device barrier, so no rank's next resolve starts storing into the root's frame before rank 0's reads
        are done.  nccl / local: nothing to do (every rank writes only its own buffer, stream order covers it)."""
        if self.mode == "p2p":
            self.hdl.barrier()
