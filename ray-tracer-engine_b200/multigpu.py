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

Completion (round 2): there is no collective per frame.  Rank r announces "my rows of frame g are in
place" by storing g+1 into ITS slot of a small flag array owned by the presenter, in stream order behind
its render kernels (`ore_flag_write`: cuStreamWriteValue32, or a one-thread st.release.sys kernel over
NVLink for a peer address); the presenter's present stream waits on the slots (`ore_flag_wait_geq`:
cuStreamWaitValue32) and acknowledges a consumed frame by storing into every rank's `ack` flag, which is
what lets a rank reuse a ring buffer.  Nobody blocks on anybody else's kernels.

Host-resident frames (the drop-in contract: `setPixelBuff` reads a HOST pointer, window.cpp:130-132):
`SharedHostFrame` is ONE POSIX shared-memory frame ring mapped and pinned by every rank; each rank copies
its own row blocks there over its own PCIe link (`ore_render_async_signal`, rows in place) and the
completion / consumed counters live in the same shared memory.

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

ROW_BLOCK = 8  # rows are dealt to the ranks in blocks of 8 (= the primary kernel's tile height)


def block_band(rank: int, world: int, height: int, block: int = ROW_BLOCK) -> dict:
    """Block-interleaved rows of one rank as ore_frame fields: blocks of `block` rows, rank r owns blocks r, r+P, ...
    A rank whose first block starts past the image gets an EMPTY band (y0 == y1), not an error."""
    if world <= 1:
        return dict(y0=0, y1=height, y_step=1, y_block=1)
    y0 = min(block * rank, height)
    return dict(y0=y0, y1=height, y_step=block * world, y_block=block)


def block_rows(rank: int, world: int, height: int, block: int = ROW_BLOCK):
    """image rows of `rank` under block_band (host-side mirror of the C ABI's row rule)"""
    if world <= 1:
        return list(range(height))
    return [y for y in range(height) if (y // block) % world == rank]


class PeerFrame:
    """Rank 0 owns `n_buffers` ring framebuffers (plain cudaMalloc through the C ABI); every rank maps them with
    CUDA IPC and its render kernels store rows straight into them over NVLink.  Completion and buffer reuse are
    signalled with stream-ordered flags (module docstring), not with a collective:
      done[r]  (presenter memory, one slot per rank): frames rank r has finished storing
      ack      (one flag in EVERY rank's memory, written by the presenter): frames the presenter has consumed"""

    ROW_BLOCK = ROW_BLOCK
    FLAG_BYTES = 4096

    def __init__(self, renderer, width: int, height: int, n_buffers: int = 2, group=None, band_of=None, local_frames=False):
        """band_of = (rank, world): render the rows THAT rank of THAT many ranks would own while the flags stay those
        of this process (single-process tools that look at one rank's share of a frame).

        Where the flags live: with one process, in device memory.  With several, in a page of POSIX shared memory that
        every rank pins (ore_host_register): a stream writes / waits on pinned host memory with cuStreamWriteValue32 /
        cuStreamWaitValue32, which needs NO SM - a one-thread kernel storing to a peer GPU's memory would queue behind
        the persistent render kernels that fill every SM (measured: acknowledgements milliseconds late)."""
        import torch.distributed as dist

        self.r = renderer
        multi = dist.is_available() and dist.is_initialized()
        self.rank = dist.get_rank(group) if multi else 0
        self.world = dist.get_world_size(group) if multi else 1
        self.band_rank, self.band_world = band_of if band_of else (self.rank, self.world)
        self.width, self.height = width, height
        self.nbytes = 4 * width * height
        self.ptrs = []
        self._owned, self._opened = [], []
        self.n_buffers = n_buffers
        self.submitted = 0    # frames this rank has submitted
        self.presented = 0    # frames the presenter has enqueued for presentation (rank 0 only)
        self._shm = None
        # local_frames (diagnostic only): every rank stores into ring buffers of its OWN memory - the frame is never
        # assembled; isolates the cost of the NVLink stores from everything else in a multi-GPU measurement
        self.local_frames = bool(local_frames)
        if self.rank == 0 or self.local_frames:
            for _ in range(n_buffers):
                p = renderer.dev_alloc(self.nbytes)
                self._owned.append(p)
                self.ptrs.append(p)
            payload = [[renderer.ipc_export(p) for p in self.ptrs]]
        else:
            payload = [None]
        if self.world > 1:
            from multiprocessing import shared_memory
            if self.rank == 0:
                self._shm = shared_memory.SharedMemory(create=True, size=self.FLAG_BYTES)
                self._shm.buf[:] = bytes(self.FLAG_BYTES)
                payload[0].append(self._shm.name)
            dist.broadcast_object_list(payload, src=0, group=group)
            if self.rank != 0:
                self._shm = shared_memory.SharedMemory(name=payload[0][-1])
                if not self.local_frames:
                    for handle in payload[0][:-1]:
                        self._opened.append(renderer.ipc_import(handle))
                    self.ptrs = list(self._opened)
            self._flag_view = np.ndarray((self.FLAG_BYTES // 4,), dtype=np.uint32, buffer=self._shm.buf)
            base = self._flag_view.ctypes.data
            renderer.host_register(base, self.FLAG_BYTES)
            self._flag_base = base
            self.done = base                                  # done[r] at base + 256 r
            self.ack = base + 256 * self.world                # ONE acknowledgement counter, read by every rank
            self.acks = [self.ack]
            dist.barrier(group=group)                         # everybody has mapped and pinned the page
        else:
            self.done = renderer.dev_alloc(256)               # single process: device memory
            self.ack = renderer.dev_alloc(256)
            self._owned += [self.done, self.ack]
            self.acks = [self.ack]

    def band_args(self, buf: int) -> dict:
        """kwargs for Renderer.render_device: this rank's block-interleaved rows of buffer `buf`, stored at
        their image position in the presenter's frame."""
        b = block_band(self.band_rank, self.band_world, self.height, self.ROW_BLOCK)
        return dict(out_ptr=self.ptrs[buf] + 4 * self.width * b["y0"], out_pitch=self.width, **b)

    # ---- one frame, flag protocol ----
    def submit(self, camera, stream: int = 0, flags: int = 0, renderer=None):
        """Render this rank's rows of the next frame into its ring buffer on `stream` and announce them.  Waits (on
        the stream, not on the host) until the presenter has consumed the frame that used the buffer before."""
        rr = renderer or self.r
        g = self.submitted
        if g >= self.n_buffers:
            rr.flag_wait_geq(self.ack, g - self.n_buffers + 1, stream)
        rr.render_device(camera, self.width, self.height, stream=stream, flags=flags, **self.band_args(g % self.n_buffers))
        # frames in flight on several streams may finish out of order: the flag goes out on the primary context's
        # in-order signal stream, behind this frame's kernels
        self.r.flag_write_after(self.done + 256 * self.rank, g + 1, stream)
        self.submitted = g + 1
        return g

    def submit_batch(self, cameras, stream: int = 0, flags: int = 0, renderer=None):
        """Render this rank's rows of the next len(cameras) frames in ONE launch set (ore_render_batch_device) into their
        ring buffers and announce them together.  len(cameras) <= n_buffers."""
        rr = renderer or self.r
        g, k = self.submitted, len(cameras)
        assert 1 <= k <= self.n_buffers
        if g + k > self.n_buffers:
            rr.flag_wait_geq(self.ack, g + k - self.n_buffers, stream)   # the oldest buffer this batch reuses is free
        args = [self.band_args((g + i) % self.n_buffers) for i in range(k)]
        band = {key: v for key, v in args[0].items() if key != "out_ptr"}
        rr.render_batch_device(cameras, self.width, self.height, [a["out_ptr"] for a in args], stream=stream, flags=flags, **band)
        self.r.flag_write_after(self.done + 256 * self.rank, g + k, stream)
        self.submitted = g + k
        return g

    def present(self, stream: int, consume=None):
        """Presenter only: on `stream`, wait for every rank's rows of the next frame, run `consume(buffer_ptr, g)`
        (enqueue-only work such as a device->host copy), then acknowledge the frame to every rank."""
        assert self.rank == 0
        g = self.presented
        for r in range(self.world):
            self.r.flag_wait_geq(self.done + 256 * r, g + 1, stream)
        if consume is not None:
            consume(self.ptrs[g % self.n_buffers], g)
        for a in self.acks:
            self.r.flag_write(a, g + 1, stream)
        self.presented = g + 1
        return g

    def present_batch(self, stream: int, k: int, consume=None):
        """Presenter only: the same for the next k frames at once (one wait per rank, one acknowledgement)"""
        assert self.rank == 0
        g = self.presented
        for r in range(self.world):
            self.r.flag_wait_geq(self.done + 256 * r, g + k, stream)
        if consume is not None:
            for i in range(k):
                consume(self.ptrs[(g + i) % self.n_buffers], g + i)
        for a in self.acks:
            self.r.flag_write(a, g + k, stream)
        self.presented = g + k
        return g

    def close(self):
        for p in self._opened:
            self.r.ipc_close(p)
        for p in self._owned:
            self.r.dev_free(p)
        self._opened, self._owned = [], []
        if self._shm is not None:
            try:
                self.r.host_unregister(self._flag_base)
            except Exception:
                pass
            self._flag_view = None
            try:
                self._shm.close()
            except BufferError:
                pass
            if self.rank == 0:
                try:
                    self._shm.unlink()
                except FileNotFoundError:
                    pass
            self._shm = None


# ---- NUMA placement of host pages next to the GPU that writes them ---------------------------------------------

def gpu_local_cpus(pci_bus_id: str | None):
    """CPUs NVML reports as local to the GPU (its NUMA node), or None when that cannot be determined"""
    try:
        import pynvml
        pynvml.nvmlInit()
        h = pynvml.nvmlDeviceGetHandleByPciBusId(pci_bus_id.encode() if isinstance(pci_bus_id, str) else pci_bus_id)
        import os
        n_cpu = os.cpu_count() or 1
        words = pynvml.nvmlDeviceGetCpuAffinity(h, (n_cpu + 63) // 64)
        cpus = {64 * w + b for w, m in enumerate(words) for b in range(64) if (int(m) >> b) & 1}
        return cpus or None
    except Exception:
        return None


class run_near_gpu:
    """Context manager: pin the calling thread to the CPUs local to a GPU (so that pages it first touches land on that
    GPU's NUMA node), restoring the affinity afterwards.  Does nothing when the topology is unknown."""

    def __init__(self, pci_bus_id):
        self.cpus = gpu_local_cpus(pci_bus_id) if pci_bus_id else None
        self.saved = None

    def __enter__(self):
        import os
        if self.cpus:
            try:
                self.saved = os.sched_getaffinity(0)
                want = self.cpus & self.saved
                if want:
                    os.sched_setaffinity(0, want)
                else:
                    self.saved = None
            except Exception:
                self.saved = None
        return self

    def __exit__(self, *exc):
        import os
        if self.saved is not None:
            try:
                os.sched_setaffinity(0, self.saved)
            except Exception:
                pass
        return False


# ---- one shared, pinned HOST frame ring written by every rank over its own PCIe link ------------------------

class SharedHostFrame:
    """POSIX shared memory: [header: done[world] and consumed, one 64-byte line each][n_buffers frames].
    Every rank maps it, pins it (`register(ptr, nbytes)`, e.g. Renderer.host_register) and copies its own row blocks
    into it; rank 0 is the presenter: the host code that would hand the frame to setPixelBuff.
      done[r]   = frames whose rows rank r has landed in host memory (written by the GPU's copy stream)
      consumed  = frames the presenter has consumed (written by the presenter's host thread)"""

    LINE = 64
    HEADER = 4096

    def __init__(self, width: int, height: int, rank: int, world: int, n_buffers: int = 3, name: str | None = None,
                 register=None, unregister=None, block: int = ROW_BLOCK):
        from multiprocessing import shared_memory

        self.width, self.height, self.rank, self.world = width, height, rank, world
        self.n_buffers, self.block = n_buffers, block
        self.frame_bytes = 4 * width * height
        total = self.HEADER + n_buffers * self.frame_bytes
        assert (world + 1) * self.LINE <= self.HEADER
        self.owner = name is None
        if self.owner:
            self.shm = shared_memory.SharedMemory(create=True, size=total)
            self.shm.buf[: self.HEADER] = bytes(self.HEADER)
        else:
            self.shm = shared_memory.SharedMemory(name=name)
        self.name = self.shm.name
        self._flags = np.ndarray((self.HEADER // 4,), dtype=np.uint32, buffer=self.shm.buf)
        self.frames = [np.ndarray((height, width), dtype=np.uint32, buffer=self.shm.buf,
                                  offset=self.HEADER + i * self.frame_bytes) for i in range(n_buffers)]
        self.base = self._flags.ctypes.data
        self.total_bytes = total
        self._register, self._unregister = register, unregister
        self._registered = False
        if register is not None and world == 1:
            self.pin()      # (with several ranks: touch_own_rows() on every rank, a barrier, then pin() - see there)
        self.submitted = 0
        self.presented = 0

    def touch_own_rows(self, pci_bus_id=None):
        """First touch of this rank's row blocks in every ring buffer, from a CPU next to this rank's GPU: the pages a
        GPU will write over PCIe are then allocated on ITS NUMA node instead of all on the creator's (shared memory is
        sparse until touched; pinning faults the rest in wherever the pinning process runs, so touch first, pin after)."""
        rows = block_rows(self.rank, self.world, self.height, self.block)
        if not rows:
            return
        with run_near_gpu(pci_bus_id):
            for fr in self.frames:
                for y0 in rows[:: self.block]:
                    fr[y0:min(y0 + self.block, self.height)] = 0

    def pin(self):
        if self._register is not None and not self._registered:
            self._register(self.base, self.total_bytes)
            self._registered = True

    # addresses / views
    def done_addr(self, r: int) -> int:
        return self.base + r * self.LINE

    def done(self, r: int) -> int:
        return int(self._flags[r * self.LINE // 4])

    def set_done(self, r: int, value: int):   # CPU producers (tests); GPU producers write it from the copy stream
        self._flags[r * self.LINE // 4] = value

    @property
    def consumed(self) -> int:
        return int(self._flags[self.world * self.LINE // 4])

    def _set_consumed(self, v: int):
        self._flags[self.world * self.LINE // 4] = v

    def band(self) -> dict:
        return block_band(self.rank, self.world, self.height, self.block)

    def row_addr(self, buf: int, y: int) -> int:
        return self.base + self.HEADER + buf * self.frame_bytes + 4 * self.width * y

    # protocol
    def can_submit(self) -> bool:
        """the ring buffer of the next frame is free once the presenter has consumed the frame that used it before"""
        return self.consumed + self.n_buffers > self.submitted

    def can_submit_batch(self, k: int) -> bool:
        return self.consumed + self.n_buffers >= self.submitted + k

    def next_slot(self):
        """(frame number g, buffer index) of this rank's next frame; call can_submit() first"""
        g = self.submitted
        self.submitted = g + 1
        return g, g % self.n_buffers

    def ready(self) -> bool:
        """presenter: every rank's rows of the next frame to present have landed"""
        f = self.presented
        return all(self.done(r) >= f + 1 for r in range(self.world))

    def present(self, consume=None) -> int:
        """presenter: consume the next frame (call ready() first) and release its buffer"""
        f = self.presented
        if consume is not None:
            consume(self.frames[f % self.n_buffers], f)
        self.presented = f + 1
        self._set_consumed(f + 1)
        return f

    def close(self):
        if self._registered and self._unregister is not None:
            self._unregister(self.base)
        self._registered = False
        self._flags = None
        self.frames = []
        try:
            self.shm.close()
        except BufferError:
            pass
        if self.owner:
            try:
                self.shm.unlink()
            except FileNotFoundError:
                pass


def numpy_band_from_frame(frame: np.ndarray, rank: int, world: int) -> np.ndarray:
    return frame[rank::world]
