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
