"""Host-side multi-GPU logic on CPU: world_size-2 `gloo`, band producer = the CPU checker.

Covers the interleaved row plan, the equal-size padded gather and the de-interleave on the
presenter; the CUDA producer and the peer-mapped path are covered by the -m gpu tests/bench."""
import os
import socket
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, W, H, q):
    sys.path.insert(0, ROOT)
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    import oraclelib
    import rte_b200

    pkg = rte_b200.pkg
    orc = oraclelib.load("port")
    sc = pkg.scene.reference_scene(16, 3)
    cam = pkg.scene.reference_camera()
    g = pkg.multigpu.BandGatherer(W, H, torch.device("cpu"))

    def produce(plan, band):
        out = orc.render(sc, cam, W, H, y0=plan.y0, y1=H, y_step=plan.y_step, n_threads=1, want_ids=False, want_t=False)
        band[: plan.rows] = torch.from_numpy(out["pixels"].view(np.int32))

    frame = pkg.multigpu.render_frame_sharded(g, produce)
    if rank == 0:
        full = orc.render(sc, cam, W, H, n_threads=1, want_ids=False, want_t=False)["pixels"]
        q.put(bool(np.array_equal(frame.numpy().view(np.uint32), full)))
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("world,H", [(2, 37), (3, 20)])
def test_band_gather_reassembles_the_frame(world, H):
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, 48, H, q)) for r in range(world)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(120)
        assert p.exitcode == 0
    assert q.get(timeout=5) is True


def test_row_plan_covers_every_row_once(pkg):
    mg = pkg.multigpu
    for world in (1, 2, 3, 4, 8):
        for H in (1, 7, 8, 9, 2160):
            rows = sorted(y for r in range(world) for y in range(r, H, world))
            assert rows == list(range(H))
            assert sum(mg.rows_of(r, world, H) for r in range(world)) == H
            assert mg.rows_of(0, world, H) == max(mg.rows_of(r, world, H) for r in range(world))


# ---- round 2: one shared host frame ring, completion / consumed counters instead of a collective -------------

def _shared_worker(rank, world, port, W, H, n_frames, n_buffers, q):
    sys.path.insert(0, ROOT)
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    import time

    import oraclelib
    import rte_b200

    pkg = rte_b200.pkg
    mg = pkg.multigpu
    orc = oraclelib.load("port")
    sc = pkg.scene.scaled_scene(16, 3)
    name = [None]
    shared = None
    if rank == 0:
        shared = mg.SharedHostFrame(W, H, 0, world, n_buffers=n_buffers)
        name = [shared.name]
    dist.broadcast_object_list(name, src=0)
    if rank != 0:
        shared = mg.SharedHostFrame(W, H, rank, world, n_buffers=n_buffers, name=name[0])
    rows = mg.block_rows(rank, world, H)
    band = shared.band()
    seen = []

    def consume(frame, f):
        full = orc.render(sc, pkg.scene.orbit_camera(sc, f), W, H, n_threads=1, want_ids=False, want_t=False)["pixels"]
        seen.append(bool(np.array_equal(frame, full)))

    def drain(block):
        while rank == 0 and shared.presented < n_frames and (shared.ready() or block):
            if shared.ready():
                shared.present(consume)
            else:
                time.sleep(0.0005)

    max_ahead = 0
    for _ in range(n_frames):
        while not shared.can_submit():      # back-pressure: the ring buffer is still being presented
            drain(False)
            time.sleep(0.0005)
        g, buf = shared.next_slot()
        max_ahead = max(max_ahead, g - shared.consumed)
        assert pkg.Renderer.rows(H, **band) == len(rows)
        if rows:   # this rank's rows only (the CPU checker has no block-interleaved mode: rows are picked afterwards)
            full = orc.render(sc, pkg.scene.orbit_camera(sc, g), W, H, n_threads=1, want_ids=False, want_t=False)["pixels"]
            shared.frames[buf][rows] = full[rows]
        shared.set_done(rank, g + 1)
        drain(False)
    drain(True)
    dist.barrier()
    if rank == 0:
        q.put((len(seen), all(seen)))
    q.put(("ahead", rank, max_ahead))
    shared.close()
    dist.destroy_process_group()


@pytest.mark.parametrize("world,H", [(2, 37), (3, 20)])
def test_shared_host_frame_ring_presents_every_frame(world, H):
    """every rank lands its own block-interleaved rows in ONE shared host frame ring; the presenter consumes frames
    in order once all ranks' counters say so, and nobody runs more than n_buffers frames ahead of it"""
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    n_frames, n_buffers = 7, 2
    procs = [ctx.Process(target=_shared_worker, args=(r, world, port, 40, H, n_frames, n_buffers, q)) for r in range(world)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(180)
        assert p.exitcode == 0
    got = [q.get(timeout=5) for _ in range(world + 1)]
    frames = [g for g in got if g[0] != "ahead"]
    assert frames == [(n_frames, True)]
    assert all(g[2] < n_buffers for g in got if g[0] == "ahead"), got


def test_block_band_matches_block_rows(pkg):
    mg = pkg.multigpu
    for world in (1, 2, 3, 8):
        for H in (1, 8, 9, 20, 63, 64, 65, 4320):
            seen = []
            for r in range(world):
                b = mg.block_band(r, world, H)
                n = pkg.Renderer.rows(H, b["y0"], b["y1"], b["y_step"], b["y_block"])
                rows = mg.block_rows(r, world, H)
                assert n == len(rows), (world, H, r, b)
                assert b["y0"] <= b["y1"] <= H      # a rank past the end of a short frame gets an EMPTY band
                seen += rows
            assert sorted(seen) == list(range(H))
