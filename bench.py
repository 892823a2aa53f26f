#!/usr/bin/env python3
"""bench.py - Mrays/s of the render hot path on N B200s (one process per GPU).

    python bench.py --gpus 1 --steps 20 --warmup 3
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 \
        --master-port P bench.py --gpus N --steps K --warmup W
    python bench.py --impl reference ...      # the reference's own CPU code on the host cores

A "step" is one frame: one pass of the hot path (frame prep -> primary nearest hit -> shadow +
shade -> packed framebuffer) over one camera of the orbit.  Default workload = BASELINE.json
configs[3] (7680x4320, 1024 spheres): the configuration the 1/2/4/8-GPU metric is quoted on; it fits
one GPU, so the SAME workload runs at every N (at N = 1 the line also carries the 4K / 1024-sphere
figures of configs[2], the scene of the single-GPU roofline target).  With N > 1 the frame is split into interleaved row bands
(strong scaling) that the ranks store straight into the presenting GPU's framebuffer over
NVLink (peer-mapped), plus one tiny NCCL all-reduce per frame as the completion signal.

Prints ONE JSON line on rank 0 (contract in the task statement): value = whole-job Mrays/s with
inputs resident in HBM; e2e = the same through the C ABI with HOST output buffers; roofline =
FP32 accounting of the dominant kernel; cpu_baseline = the CPU checker timed on a bounded sample.
"""
from __future__ import annotations

import argparse
import json
import os
import statistics
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

WORKLOADS = {
    # name: (W, H, n_spheres, seed, kind, description)
    "vga64": (640, 480, 64, 1, "reference", "configs[0]: 640x480, reference-formula scene R(64,1), reference camera"),
    "1080p64": (1920, 1080, 64, 2, "scaled", "configs[1]: 1920x1080, 64 random spheres S(64,2), orbit"),
    "4k1024": (3840, 2160, 1024, 3, "scaled", "configs[2]: 3840x2160, 1024 random spheres S(1024,3), 240-frame camera orbit"),
    "8k1024": (7680, 4320, 1024, 3, "scaled", "configs[3]: 7680x4320, 1024 spheres S(1024,3), row bands"),
    "4k16384": (3840, 2160, 16384, 5, "scaled", "configs[4]: 3840x2160, 16384 spheres S(16384,5)"),
}
FLOP_PER_TEST = 18
NOMINAL_FP32_TFLOPS = 148 * 128 * 2 * 1.965e9 / 1e12


def make_workload(pkg, name):
    W, H, n, seed, kind, desc = WORKLOADS[name]
    sc = pkg.scene.reference_scene(n, seed) if kind == "reference" else pkg.scene.scaled_scene(n, seed)

    def camera(frame):
        return pkg.scene.reference_camera() if kind == "reference" else pkg.scene.orbit_camera(sc, frame % 240)

    return W, H, sc, camera, desc


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled DURING the timed region (profiling recipe)."""

    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index):
        self.gpu = gpu_index
        self.rows = []
        self.proc = None

    def start(self):
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", f"--id={self.gpu}", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "100"],
                stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.th = threading.Thread(target=self._read, daemon=True)
            self.th.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append(line.strip())

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.05)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm, mx, reasons, power = [], [], set(), []
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for r in self.rows:
            f = [x.strip() for x in r.split(",")]
            if len(f) < 9:
                continue
            try:
                sm.append(float(f[1]))
                mx.append(float(f[2]))
                power.append(float(f[3]))
            except ValueError:
                continue
            for i, n in enumerate(names):
                if f[5 + i].lower().startswith("active"):
                    reasons.add(n)
        return {"sm_mhz": statistics.median(sm) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "power_w_max": max(power) if power else None, "samples": len(sm), "reasons": sorted(reasons)}


def issue_block(workload):
    """Executed-instruction side of the default path's kernels, from the committed ncu summaries (profiles/): how busy
    the issue slots and the FMA pipe are.  Static context for the line (ncu never runs inside bench.py); None when the
    summaries are for another workload or unreadable."""
    try:
        if workload != "8k1024":
            return None
        out = {"source": "profiles/r01_ncu_{shade_setup,shadow_beam_staged,primary_tile}_8k1024.json (ncu --set full, "
                         "one launch each, --clock-control none)", "kernels": {}}
        for key, fn in (("shade_setup_kernel", "r01_ncu_shade_setup_8k1024.json"),
                        ("shadow_beam_kernel (staged)", "r01_ncu_shadow_beam_staged_8k1024.json"),
                        ("primary_tile_kernel", "r01_ncu_primary_tile_8k1024.json")):
            m = json.load(open(os.path.join(ROOT, "profiles", fn)))["metrics"]

            def num(name):
                return float(str(m[name]).split()[0])

            out["kernels"][key] = {
                "issue_slot_utilisation": num("smsp__issue_active.avg.pct_of_peak_sustained_active") / 100.0,
                "fma_pipe_utilisation": num("sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active") / 100.0,
                "ncu_ms": num("gpu__time_duration.sum"),
                "registers": int(num("launch__registers_per_thread")),
            }
        return out
    except Exception:
        return None


def hbm_block(traffic_bytes, kernel_ms):
    """DRAM side of the dominant kernel against the measured HBM peak (MEASURED_PEAKS.json, else the recipe's fallback)"""
    peak, src = 6650.0, "fallback (B200_PROFILING.md)"
    try:
        peak = float(json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"])
        src = "MEASURED_PEAKS.json"
    except Exception:
        pass
    if not traffic_bytes:
        return {"peak_gbs": peak, "peak_source": src, "achieved_gbs": None, "frac": None}
    ach = traffic_bytes / (kernel_ms * 1e-3) / 1e9
    return {"peak_gbs": peak, "peak_source": src, "achieved_gbs": ach, "frac": ach / peak,
            "note": "ncu dram bytes of the shadow pass of one frame / its CUDA-event time (the staging buffer between the "
                    "two shadow kernels is most of it); the pass is instruction-issue bound, not HBM-bound"}


def dist_env():
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    return rank, local, world


# ---------------------------------------------------------------------------------------------
# reference arm / cpu baseline: the reference's own CPU code (oracle/_ref) or its C restatement
# ---------------------------------------------------------------------------------------------
def cpu_sample_plan(orc, sc, cam, W, H, budget_s):
    """pick a row stride so that one sample costs about `budget_s` seconds on the host cores"""
    probe_step = max(1, H // 8)
    t0 = time.perf_counter()
    orc.render(sc, cam, W, H, y0=probe_step // 2, y_step=probe_step, want_ids=False, want_t=False)
    dt = time.perf_counter() - t0
    rows_probe = (H - probe_step // 2 + probe_step - 1) // probe_step
    per_row = dt / max(1, rows_probe)
    rows = int(max(1, min(H, budget_s / max(per_row, 1e-9))))
    step = max(1, H // rows)
    return step


def run_cpu(args, pkg, what):
    """times the CPU checker on a bounded sample; returns (mrays, info)"""
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import oraclelib  # the one place bench.py executes oracle/ (cpu_baseline + --impl reference)

    orc = oraclelib.load("best")
    W, H, sc, camera, desc = make_workload(pkg, args.workload)
    cores = os.cpu_count() or 1
    n_steps = args.steps if what == "reference" else 1
    n_warm = args.warmup if what == "reference" else 0
    budget = min(10.0, 150.0 / max(1, n_steps + n_warm)) if what == "reference" else 12.0
    step = cpu_sample_plan(orc, sc, camera(0), W, H, budget)
    y0 = step // 2
    rows = (H - y0 + step - 1) // step
    times = []
    for i in range(n_warm + n_steps):
        cam = camera(i if what == "reference" else 0)
        t0 = time.perf_counter()
        orc.render(sc, cam, W, H, y0=y0, y_step=step, want_ids=False, want_t=False)
        dt = time.perf_counter() - t0
        if i >= n_warm:
            times.append(dt)
    total = sum(times)
    mrays = rows * W * len(times) / total / 1e6
    info = {"value": mrays, "unit": "Mrays/s", "cores": cores, "kind": orc.kind,
            "sample": f"rows {y0}::{step} of each {W}x{H} frame ({rows} rows, {rows * W} primary rays per step, "
                      f"{len(times)} step(s), {total:.1f} s), all {cores} host threads (OpenMP)"}
    return mrays, info, total / len(times)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="8k1024", choices=sorted(WORKLOADS))
    ap.add_argument("--gather", default="peer", choices=["peer", "nccl"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--in-flight", type=int, default=4, help="frames in flight per GPU (contexts/streams used round-robin)")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "ours" else args.warmup
    rank, local, world = dist_env()

    import rte_b200
    pkg = rte_b200.pkg
    W, H, sc, camera, desc = make_workload(pkg, args.workload)
    config = {"workload": desc, "width": W, "height": H, "n_spheres": sc.n_spheres, "n_lights": int(sc.lights.shape[0]),
              "rays_per_pixel": 1, "scene": sc.name}

    if args.impl == "reference":
        if rank != 0:
            return 0
        mrays, info, sec = run_cpu(args, pkg, "reference")
        line = {"impl": "reference", "metric": "Mrays/s (primary rays)", "value": mrays, "unit": "Mrays/s",
                "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": sec * 1e3,
                "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
                "config": config, "cpu_baseline": info,
                "e2e": {"value": mrays, "unit": "Mrays/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
                "gpu_launches": 0}
        print(json.dumps(line))
        return 0

    # ---------------- our arm ----------------
    import torch
    import torch.distributed as dist

    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device - the render path has no CPU fallback")
    torch.cuda.set_device(local)
    if world > 1:
        if os.environ.get("NCCL_DEBUG", "VERSION").upper() == "VERSION":
            os.environ["NCCL_DEBUG"] = "WARN"  # NCCL prints its version banner on stdout; rank 0's stdout is ONE JSON line
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    pkg.build.build_library()
    r = pkg.Renderer(local)
    r.set_scene(sc)
    stream = torch.cuda.Stream(device=local)
    # a second context + stream: consecutive frames alternate between the two, so frame f+1's kernels fill the
    # SMs that frame f's persistent CTAs vacate at the end of a kernel (double-buffered rendering; every frame
    # still completes inside the timed region)
    ctxs = [(r, stream)]
    for _ in range(max(1, args.in_flight) - 1):
        rx = pkg.Renderer(local)
        rx.set_scene(sc)
        ctxs.append((rx, torch.cuda.Stream(device=local)))
    NF = len(ctxs)
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=f"cuda:{local}")  # > 126 MB L2
    sync_t = torch.zeros(1, dtype=torch.int32, device=f"cuda:{local}")

    peer = None
    gatherer = None
    gather_mode = args.gather
    if not (world > 1 and gather_mode == "nccl"):
        err = None
        try:
            peer = pkg.multigpu.PeerFrame(r, W, H, n_buffers=max(2, args.in_flight))
        except Exception as e:  # CUDA IPC refused (e.g. separate IPC namespaces): decided collectively below
            err = e
        if world > 1:
            bad = torch.tensor([1 if err is not None else 0], dtype=torch.int32, device=f"cuda:{local}")
            dist.all_reduce(bad, op=dist.ReduceOp.MAX)
            if int(bad.item()):
                if rank == 0:
                    print(f"bench.py: peer-mapped framebuffer unavailable ({err}); using the NCCL band gather", file=sys.stderr)
                if peer is not None:
                    peer.close()
                peer, gather_mode = None, "nccl"
        elif err is not None:
            raise err
    if peer is None:
        gatherer = pkg.multigpu.BandGatherer(W, H, torch.device("cuda", local))

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def render_step(f, which=0):
        """device-resident step: this rank's rows of frame f into the presenter's framebuffer"""
        cam = camera(f)
        rr, st = ctxs[which]
        with torch.cuda.stream(st):
            if peer is not None:
                rr.render_device(cam, W, H, stream=st.cuda_stream, **peer.band_args(f % len(peer.ptrs)))
                if world > 1:
                    dist.all_reduce(sync_t)  # completion signal: every rank's rows of frame f have landed
            else:
                plan = gatherer.plan
                rr.render_device(cam, W, H, out_ptr=gatherer.band.data_ptr(), stream=st.cuda_stream,
                                 y0=plan.y0, y1=H, y_step=plan.y_step)
                gatherer.gather()

    def flush_l2():
        with torch.cuda.stream(stream):
            flush.fill_(1)

    overlap = (gatherer is None)   # the NCCL-gather variant reuses one band buffer: no frame overlap there
    frame_bytes = W * H * 4
    l2_note = ("not flushed: every step writes a fresh framebuffer plus per-pixel hit records "
               f"({3 * frame_bytes / 1e6:.0f} MB per step, L2 is 126 MB) and reads 64 KB of scene records that the kernels "
               "keep in shared memory by design; nothing a step touches can be served from the previous step's L2")
    small_frame = 3 * frame_bytes / max(1, world) < (160 << 20)
    if small_frame:
        l2_note = "flushed before every step (256 MiB fill on the step's stream, inside the timed region)"

    # warm-up (both contexts)
    for f in range(max(args.warmup, 2) * (NF if overlap else 1)):
        render_step(f, f % NF if overlap else 0)
    barrier()

    sampler = ClockSampler(local)
    sampler.start()
    kernel_ms = []
    ev_start = torch.cuda.Event(enable_timing=True)
    ev_end = torch.cuda.Event(enable_timing=True)
    barrier()
    ev_start.record(stream)
    for _, st in ctxs[1:]:
        st.wait_event(ev_start)
    for i in range(args.steps):
        which = i % NF if overlap else 0
        if small_frame:
            with torch.cuda.stream(ctxs[which][1]):
                flush.fill_(1)
        render_step(args.warmup + i, which)   # no host sync inside the timed region
    for _, st in ctxs[1:]:
        ev_other = torch.cuda.Event()
        ev_other.record(st)
        stream.wait_event(ev_other)
    ev_end.record(stream)
    barrier()
    clocks = sampler.stop()
    # per-kernel device times (CUDA events inside the library, on the launching stream): separate untimed pass,
    # one frame at a time
    for i in range(min(args.steps, 8)):
        render_step(args.warmup + i, 0)
        stream.synchronize()
        kernel_ms.append(r.kernel_ms())
    launches_per_frame = int(r.counters()["kernel_launches"])   # counted by the library per render call
    barrier()
    total_ms = torch.tensor([ev_start.elapsed_time(ev_end)], dtype=torch.float64, device=f"cuda:{local}")
    if world > 1:
        dist.all_reduce(total_ms, op=dist.ReduceOp.MAX)
    total_ms = float(total_ms.item())
    value = W * H * args.steps / (total_ms * 1e-3) / 1e6

    # ---- e2e: through the C ABI with HOST output, device->host copy inside the timed region ----
    # N = 1: ore_render_async into two pinned host framebuffers (frame f copies out while f+1 renders; all K
    #        frames are on the host before the clock stops) and, for reference, the synchronous ore_render.
    # N > 1: every rank renders its rows into the presenter's frame (peer stores), completion all-reduce, then the
    #        presenter copies the frame to pinned host memory.
    host = [r.host_alloc((H, W)) for _ in range(2)] if rank == 0 else None
    frame_check = None

    def e2e_sync_step(f):
        cam = camera(f)
        if world == 1:
            r.render(cam, W, H, out=host[f % 2])     # ore_render: kernels + D2H of the frame, synchronous
        else:
            render_step(f)
            stream.synchronize()
            if rank == 0:
                if peer is not None:
                    r.copy_to_host(host[f % 2], peer.ptrs[f % len(peer.ptrs)])
                else:
                    host[f % 2][:] = gatherer.frame.cpu().numpy().view(np.uint32)

    def timed_e2e(step_fn, finish_fn=None):
        for f in range(2):
            step_fn(f)
        if finish_fn:
            finish_fn()
        barrier()
        t0 = time.perf_counter()
        for i in range(args.steps):
            step_fn(args.warmup + i)
        if finish_fn:
            finish_fn()
        if world > 1:
            dist.barrier()
        dt = time.perf_counter() - t0
        tt = torch.tensor([dt], dtype=torch.float64, device=f"cuda:{local}")
        if world > 1:
            dist.all_reduce(tt, op=dist.ReduceOp.MAX)
        return W * H * args.steps / float(tt.item()) / 1e6

    e2e_sync_value = timed_e2e(e2e_sync_step)
    if world == 1:
        e2e_value = timed_e2e(lambda f: r.render_async(camera(f), W, H, out=host[f % 2]), r.wait)
        e2e_mode = "pipelined: ore_render_async x K + ore_wait, two pinned host framebuffers"
    else:
        e2e_value = e2e_sync_value
        e2e_mode = "peer-written frame, completion all-reduce, presenter D2H to pinned host memory, per frame"
        # correctness of the sharded frame: the presenter re-renders the last frame alone and compares
        if rank == 0:
            last = args.warmup + args.steps - 1
            alone = r.render(camera(last), W, H)
            frame_check = "identical" if np.array_equal(alone, host[last % 2]) else "MISMATCH"
    barrier()

    # ---- roofline accounting (untimed): reference-order test counts of the SAME frames ----
    own = {"primary": 0.0, "shadow": 0.0, "sky": 0.0, "hits": 0.0}
    if peer is not None:
        band_kw = peer.band_args(0)
    else:
        band_kw = dict(out_ptr=gatherer.band.data_ptr(), y0=gatherer.plan.y0, y1=H, y_step=gatherer.plan.y_step)
    count_out = peer.ptrs[0] if peer is not None else gatherer.band.data_ptr()
    for i in range(args.steps):
        r.render_device(camera(args.warmup + i), W, H, flags=pkg.capi.ORE_FLAG_COUNT_REFERENCE_TESTS, **band_kw)
        c = r.counters()
        own["primary"] += c["primary_tests"]
        own["shadow"] += c["shadow_tests_ref"]
        own["sky"] += c["sky_tests"]
        own["hits"] += c["hit_pixels"]
    counts = torch.tensor([own["primary"], own["shadow"], own["sky"], own["hits"]], dtype=torch.float64,
                          device=f"cuda:{local}")
    if world > 1:
        dist.all_reduce(counts)
    tests_primary, tests_shadow, tests_sky, hits = [float(v) for v in counts.tolist()]
    # the same accounting for the per-ray shadow kernel (every sample ray x every sphere; the kernel the
    # FP32-pipe utilisation in profiles/ is quoted on), 3 untimed frames
    per_ray = None
    if world == 1:
        pr_ms, pr_tests = [], 0.0
        for i in range(3):
            r.render_device(camera(args.warmup + i), W, H, out_ptr=count_out,
                            flags=pkg.capi.ORE_FLAG_PER_RAY_SHADOW | pkg.capi.ORE_FLAG_COUNT_REFERENCE_TESTS)
            pr_ms.append(r.kernel_ms()[2])
            pr_tests += r.counters()["shadow_tests_ref"]
        per_ray = {"kernel": "shadow_kernel<3> (ORE_FLAG_PER_RAY_SHADOW)", "kernel_ms": statistics.mean(pr_ms),
                   "achieved": FLOP_PER_TEST * pr_tests / 3 / (statistics.mean(pr_ms) * 1e-3) / 1e12, "unit": "TFLOP/s"}
    peak_tf, nominal_mhz = r.measure_fp32_peak()
    # dominant kernel = shadow kernel; roofline of rank 0's own launches against rank 0's own tests
    k_shadow_ms = statistics.mean(m[2] for m in kernel_ms)
    k_primary_ms = statistics.mean(m[1] for m in kernel_ms)
    k_prep_ms = statistics.mean(m[0] for m in kernel_ms)
    own_shadow = own["shadow"]
    shadow_flop_per_launch = FLOP_PER_TEST * own_shadow / args.steps
    achieved_tf = shadow_flop_per_launch / (k_shadow_ms * 1e-3) / 1e12
    step_flop = FLOP_PER_TEST * (tests_primary + tests_shadow + tests_sky) / args.steps
    step_tf = step_flop / (total_ms / args.steps * 1e-3) / 1e12

    traffic = None
    tpath = os.path.join(ROOT, "profiles", "ncu_traffic.json")
    if os.path.isfile(tpath):
        try:
            traffic = json.load(open(tpath)).get(args.workload, {}).get("shadow_kernel_dram_bytes_per_launch")
        except Exception:
            traffic = None

    if traffic and world > 1:
        traffic = traffic / world   # the ncu capture is a whole frame on one GPU; a rank sweeps 1/world of the hit pixels
    also_4k = None
    if world == 1 and args.workload == "8k1024":
        # configs[2] (4K / 1024 spheres, the single-GPU roofline scene): same scene, quarter of the pixels
        W4, H4 = 3840, 2160
        ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        ms4, kms4 = [], []
        for i in range(args.warmup + 10):
            flush_l2()
            ev0.record(stream)
            with torch.cuda.stream(stream):
                r.render_device(camera(i), W4, H4, out_ptr=count_out, stream=stream.cuda_stream)
            ev1.record(stream)
            ev1.synchronize()
            if i >= args.warmup:
                ms4.append(ev0.elapsed_time(ev1))
                kms4.append(r.kernel_ms())
        also_4k = {"workload": WORKLOADS["4k1024"][5], "value": W4 * H4 / statistics.mean(ms4) / 1e3, "unit": "Mrays/s",
                   "ms_per_step": statistics.mean(ms4), "steps": 10,
                   "kernel_ms": {"prep": statistics.mean(m[0] for m in kms4), "primary": statistics.mean(m[1] for m in kms4),
                                 "shadow": statistics.mean(m[2] for m in kms4)}}
    primary_only = None
    if world == 1:
        # the primary-only fraction (light_size = 0 is a legal reference configuration: the light
        # loop runs zero times, kernel.cu:1665): tests = W*H*N exactly
        r.set_lights(sc.lights[:0])
        ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        for i in range(3):
            r.render_device(camera(i), W, H, out_ptr=count_out, stream=stream.cuda_stream)
        stream.synchronize()
        ev0.record(stream)
        for i in range(10):
            r.render_device(camera(args.warmup + i), W, H, out_ptr=count_out, stream=stream.cuda_stream)
        ev1.record(stream)
        ev1.synchronize()
        po_ms = ev0.elapsed_time(ev1) / 10
        r.set_lights(sc.lights)
        po_tf = FLOP_PER_TEST * W * H * sc.n_spheres / (po_ms * 1e-3) / 1e12
        primary_only = {"what": "n_lights = 0: primary nearest hit + sky only; tests = W*H*N", "ms_per_step": po_ms,
                        "value": W * H / po_ms / 1e3, "unit": "Mrays/s", "achieved": po_tf, "achieved_unit": "TFLOP/s"}
    fast_libm = None
    if world == 1:
        # the same frames with ORE_FLAG_FAST_LIBM (CUDA's libm: within 1 LSB instead of bit-identical)
        ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        for i in range(3):
            r.render_device(camera(i), W, H, out_ptr=count_out, stream=stream.cuda_stream, flags=pkg.capi.ORE_FLAG_FAST_LIBM)
        stream.synchronize()
        ev0.record(stream)
        for i in range(10):
            r.render_device(camera(args.warmup + i), W, H, out_ptr=count_out, stream=stream.cuda_stream,
                            flags=pkg.capi.ORE_FLAG_FAST_LIBM)
        ev1.record(stream)
        ev1.synchronize()
        fl_ms = ev0.elapsed_time(ev1) / 10
        fast_libm = {"flag": "ORE_FLAG_FAST_LIBM", "value": W * H / fl_ms / 1e3, "unit": "Mrays/s", "ms_per_step": fl_ms, "steps": 10,
                     "note": "CUDA's cosf/sinf/acosf/atan2f instead of the glibc-bit-compatible device functions; ids/t unchanged, "
                             "pixels within 1 LSB on >= 99.9 % instead of bit-identical"}
    ref_gpu = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        # the reference's OWN CUDA kernel built for sm_100 (oracle/_ref/libref_sm100*.so), same scene, 3 frames:
        # reported alongside, never on our path
        try:
            sys.path.insert(0, os.path.join(ROOT, "tests"))
            import oraclelib
            ref_gpu = {}
            cams = [camera(args.warmup + i) for i in range(3)]
            for fast in (False, True):
                if oraclelib.have_ref_gpu(fast):
                    _, ms_up, ms_k = oraclelib.RefGpu(fast=fast).render(sc, cams, W, H)
                    ref_gpu["use_fast_math" if fast else "default_flags"] = {
                        "ms_per_update": ms_up, "ms_per_kernel": ms_k,
                        "mrays_per_s_update": W * H / ms_up / 1e3, "mrays_per_s_kernel": W * H / ms_k / 1e3}
            ref_gpu["what"] = ("the reference's rayTrace kernel / update() compiled unmodified in arithmetic for sm_100 "
                               "(oracle/ref_build/make_ref_gpu.py), run headless on this GPU in this run")
        except Exception as e:  # measurement extra only
            ref_gpu = {"unavailable": str(e)[:200]}
    if rank == 0:
        cpu = None
        if world == 1 and not args.no_cpu_baseline:
            _, cpu, _ = run_cpu(args, pkg, "cpu_baseline")
        line = {
            "metric": "Mrays/s (primary rays)", "value": value, "unit": "Mrays/s", "n_gpus": world,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": total_ms / args.steps,
            "higher_is_better": True, "scaling": "strong", "vs_baseline": None,
            "dtype": "f32", "data": "synthetic",
            "config": dict(config, l2=l2_note,
                           frame_overlap=(f"{NF} frames in flight per GPU (contexts/streams used round-robin)"
                                          if overlap and NF > 1 else "none"),
                           parallelism=(f"row-bands x{world} (8-row blocks dealt round-robin), presenter = rank 0, gather = {gather_mode}"
                                        if world > 1 else "1 GPU"),
                           hit_pixel_fraction=hits / (W * H * args.steps),
                           tests_per_pixel={"primary": tests_primary / (W * H * args.steps),
                                            "shadow_reference_order": tests_shadow / (W * H * args.steps)}),
            "clocks": clocks,
            "e2e": {"value": e2e_value, "unit": "Mrays/s", "h2d_bytes_per_step": 36 + 32,
                    "d2h_bytes_per_step": W * H * 4, "mode": e2e_mode,
                    "synchronous_value": e2e_sync_value,
                    "note": "inputs per step are the 36-byte camera and the 32-byte frame descriptor (kernel arguments); "
                            "output per step is the whole framebuffer read back to pinned host memory; "
                            "synchronous_value = ore_render (launch + sync + copy per call, the reference update() semantics)",
                    "sharded_frame_check": frame_check},
            "gpu_launches": launches_per_frame * args.steps,
            "gpu_launches_per_frame": {"count": launches_per_frame,
                                       "kernels": "prep_frame, primary_tile, one catch-all fused shadow_beam (normally empty), then per "
                                                  "hit-list chunk shade_setup + shadow_beam (as many chunks as the previous "
                                                  "frame's hit count suggests); one fused shadow_beam only when the sphere "
                                                  "records exceed shared memory"},
            "roofline": {
                "bound": "fp32", "kernel": "shadow pass = shade_setup_kernel + shadow_beam_kernel (soft-shadow any-hit + shading; default path)",
                "achieved": achieved_tf, "peak": peak_tf, "unit": "TFLOP/s", "frac": achieved_tf / peak_tf if peak_tf else None,
                "traffic": traffic,
                "peak_source": "measured: FFMA burn on this GPU in this run (MEASURED_PEAKS.json carries HBM and bf16 only); "
                               f"nominal 148 SM x 128 lanes x 2 x 1.965 GHz = {NOMINAL_FP32_TFLOPS:.2f}",
                "flop_per_test": FLOP_PER_TEST,
                "tests_per_launch": own_shadow / args.steps,
                "kernel_ms": k_shadow_ms,
                "definition": "achieved = 18 FLOP x reference-order sphere tests of the launch / CUDA-event kernel time "
                              "(SURVEY.md 8d). The kernel reaches the reference's results with far fewer executed FLOP "
                              "(shared L/C per pixel, 3-FMA filter, per-light cone and per-warp beam tests ahead of the "
                              "per-ray tests), so frac is an ALGORITHMIC-work rate and exceeds 1; executed-instruction "
                              "utilisation of the kernels (ncu, profiles/) is under `executed`, the FP32-pipe-bound "
                              "kernel generation under per_ray_kernel.",
                "per_ray_kernel": (dict(per_ray, frac=per_ray["achieved"] / peak_tf) if per_ray and peak_tf else per_ray),
                "whole_step": {"achieved": step_tf, "frac": step_tf / peak_tf if peak_tf else None,
                               "frac_of_nominal": step_tf / NOMINAL_FP32_TFLOPS},
                "kernel_ms_all": {"prep": k_prep_ms, "primary": k_primary_ms, "shadow": k_shadow_ms},
                "dram_write_gbs_framebuffer": W * H * 4 / world / ((k_primary_ms + k_shadow_ms) * 1e-3) / 1e9,
                "hbm": hbm_block(traffic, k_shadow_ms),
                "executed": issue_block(args.workload),
            },
            "cpu_baseline": cpu,
            "reference_kernel_on_b200": ref_gpu,
            "also_configs2_4k1024": also_4k,
            "fast_libm": fast_libm,
            "primary_only": (dict(primary_only, frac=primary_only["achieved"] / peak_tf) if primary_only and peak_tf else primary_only),
        }
        print(json.dumps(line))
    if host:
        for h_ in host:
            r.host_free(h_)
    if peer is not None:
        peer.close()
    for rx, _ in ctxs[1:]:
        rx.close()
    r.close()
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()
    return 0


if __name__ == "__main__":
    sys.exit(main())
