"""Row-band sharding of one frame across the GPUs of one box (one process per GPU).

The reference is single-GPU (no collectives anywhere in its tree).  Pixels are independent
(`kernel.cu:1682,1688` write only `pixels[y*width+x]`), so a frame shards by rows with no
data-path exchange; the ONE exchange step is gathering the bands to the presenting GPU
(rank 0), which owns the framebuffer handed to `setPixelBuff` (`window.cpp:130-132`).

Partition: interleaved rows - rank r renders rows r, r+P, r+2P, ... (row cost varies ~8x
between sky rows and sphere rows; interleaving balances it
without a cost model).  Each rank needs the GLOBAL row index (dy depends on it,
`kernel.cu:1625`), which the C ABI takes as (y0, y1, y_step).

Gather (two implementations):
  * "nccl"  : `torch.distributed.gather` of equal-size (padded) bands over NVLink, then a
              strided de-interleave copy on the presenter.
  * "peer"  : the presenter's framebuffer is CUDA-IPC-mapped into every rank; each rank's
              render kernels store their pixels straight into their rows of the presenter's
              frame through NVLink (compute fused with its "collective": no separate
              gather kernel, no staging copy).  Ranks render into a strided row view.

The band producer is pluggable so the host logic can be tested on CPU with the `gloo`
backend (tests/test_multigpu_cpu.py) - there the producer is the CPU checker.
"""
from __future__ import annotations

from dataclasses import dataclass

import numpy as np


@dataclass
class BandPlan:
    """Rows of one rank: y0 = rank, y_step = world (interleaved)."""

    rank: int
    world: int
    width: int
    height: int

    @property
    def y0(self) -> int:
        return self.rank

    @property
    def y_step(self) -> int:
        return self.world

    @property
    def rows(self) -> int:
        return rows_of(self.rank, self.world, self.height)

    @property
    def max_rows(self) -> int:
        return rows_of(0, self.world, self.height)


def rows_of(rank: int, world: int, height: int) -> int:
    return max(0, (height - rank + world - 1) // world)


def deinterleave(bands, height: int, width: int, world: int, out=None):
    """bands[r] holds rows r, r+world, ... (possibly padded to max_rows); returns [height, width]."""
    import torch

    if out is None:
        out = torch.empty((height, width), dtype=bands[0].dtype, device=bands[0].device)
    for r in range(world):
        n = rows_of(r, world, height)
        if n:
            out[r::world] = bands[r][:n]
    return out


class BandGatherer:
    """Gathers interleaved row bands to rank 0 with torch.distributed (nccl on GPUs, gloo on CPU)."""

    def __init__(self, width: int, height: int, device, dtype=None, group=None):
        import torch
        import torch.distributed as dist

        self.dist = dist
        self.group = group
        self.rank = dist.get_rank(group)
        self.world = dist.get_world_size(group)
        self.plan = BandPlan(self.rank, self.world, width, height)
        self.device = device
        dtype = dtype or torch.int32
        # equal-size bands for the collective: pad to the largest band (differs by <= 1 row)
        self.band = torch.zeros((self.plan.max_rows, width), dtype=dtype, device=device)
        self.gather_list = (
            [torch.empty_like(self.band) for _ in range(self.world)] if self.rank == 0 else None
        )
        self.frame = torch.empty((height, width), dtype=dtype, device=device) if self.rank == 0 else None

    def gather(self):
        """band (this rank) -> frame on rank 0; returns the frame on rank 0, None elsewhere."""
        if self.world == 1:
            self.frame[:] = self.band[: self.plan.rows]
            return self.frame
        self.dist.gather(self.band, self.gather_list, dst=0, group=self.group)
        if self.rank == 0:
            return deinterleave(self.gather_list, self.plan.height, self.plan.width, self.world, out=self.frame)
        return None


def render_frame_sharded(gatherer: BandGatherer, produce_band):
    """produce_band(plan, band_tensor) fills this rank's rows; returns the frame on rank 0."""
    produce_band(gatherer.plan, gatherer.band)
    return gatherer.gather()


# ---- CUDA IPC peer mapping (presenter framebuffer visible to every rank) -----------------------

class PeerFrame:
    """Rank 0 owns `n_buffers` ping-pong framebuffers (plain cudaMalloc through the C ABI); every
    rank maps them with CUDA IPC and its render kernels store rows straight into them."""

    def __init__(self, renderer, width: int, height: int, n_buffers: int = 2, group=None):
        import torch.distributed as dist

        self.r = renderer
        multi = dist.is_available() and dist.is_initialized()
        self.rank = dist.get_rank(group) if multi else 0
        self.world = dist.get_world_size(group) if multi else 1
        self.width, self.height = width, height
        self.nbytes = 4 * width * height
        self.ptrs = []
        self._owned, self._opened = [], []
        if self.rank == 0:
            for _ in range(n_buffers):
                p = renderer.dev_alloc(self.nbytes)
                self._owned.append(p)
                self.ptrs.append(p)
            payload = [[renderer.ipc_export(p) for p in self.ptrs]]
        else:
            payload = [None]
        if self.world > 1:
            dist.broadcast_object_list(payload, src=0, group=group)
        if self.rank != 0:
            for handle in payload[0]:
                p = renderer.ipc_import(handle)
                self._opened.append(p)
                self.ptrs.append(p)

    ROW_BLOCK = 8  # rows are dealt to the ranks in blocks of 8 (= the primary kernel's tile height)

    def band_args(self, buf: int) -> dict:
        """kwargs for Renderer.render_device: this rank's block-interleaved rows of buffer `buf`, stored at
        their image position in the presenter's frame."""
        b = self.ROW_BLOCK if self.world > 1 else 1
        y0 = b * self.rank
        return dict(out_ptr=self.ptrs[buf] + 4 * self.width * y0, y0=y0, y1=self.height,
                    y_step=b * self.world, y_block=b, out_pitch=self.width)

    def close(self):
        for p in self._opened:
            self.r.ipc_close(p)
        for p in self._owned:
            self.r.dev_free(p)
        self._opened, self._owned = [], []


def numpy_band_from_frame(frame: np.ndarray, rank: int, world: int) -> np.ndarray:
    return frame[rank::world]
