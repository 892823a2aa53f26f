"""Executed FP32 work per unit, from an ncu launch list with the thread-level op counters.

    ncu --metrics smsp__sass_thread_inst_executed_op_fadd_pred_on.sum,smsp__sass_thread_inst_executed_op_fmul_pred_on.sum,\
smsp__sass_thread_inst_executed_op_ffma_pred_on.sum,smsp__inst_executed.sum,smsp__issue_active.avg.pct_of_peak_sustained_active,\
dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum --clock-control none --csv --log-file ops.csv \
        python tools/profile_target.py --workload 8k1024 --frames 1 --json gpurun_out/ops_units.json

    python tools/executed_flops.py ops.csv ops_units.json 8k1024 profiles/r02_executed_flops.json

FLOP = fadd + fmul + 2 * ffma (predicated-on thread instructions).  Units: hit pixels for the shadow-pass kernels,
pixels for the primary kernel - both reported by the profiled program itself (device counters of that frame)."""
import csv
import json
import sys


def main():
    ops_csv, units_json, workload, out = sys.argv[1:5]
    rows = list(csv.reader(open(ops_csv, errors="replace")))
    hi = [i for i, r in enumerate(rows) if r and r[0] == "ID"][0]
    hdr = rows[hi]
    ix = {h: i for i, h in enumerate(hdr)}
    per = {}
    for r in rows[hi + 1:]:
        if len(r) < len(hdr):
            continue
        key = (r[ix["ID"]], r[ix["Kernel Name"]])
        v = r[ix["Metric Value"]].replace(",", "")
        try:
            per.setdefault(key, {})[r[ix["Metric Name"]]] = float(v)
        except ValueError:
            pass
        per[key]["_unit_" + r[ix["Metric Name"]]] = r[ix["Metric Unit"]]
    units = json.load(open(units_json))
    frame = units["frames"][-1]
    hits, pixels = frame["hit_pixels"], frame["pixels"]

    def scale_bytes(m, name):
        u = m.get("_unit_" + name, "byte")
        return m.get(name, 0.0) * {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}.get(u, 1)

    def scale_ms(m):
        u = m.get("_unit_gpu__time_duration.sum", "ns")
        return m.get("gpu__time_duration.sum", 0.0) * {"ns": 1e-6, "us": 1e-3, "usecond": 1e-3, "ms": 1, "msecond": 1, "nsecond": 1e-6}.get(u, 1e-6)

    kernels = {}
    # the LAST frame's launches: walk backwards until a prep_frame_kernel is met
    keys = sorted(per, key=lambda k: int(k[0]))
    last = []
    for k in reversed(keys):
        last.append(k)
        if "prep_frame" in k[1]:
            break
    for k in reversed(last):
        m = per[k]
        name = k[1].split("(")[0].replace("void ", "").replace("ore::", "")
        fl = (m.get("smsp__sass_thread_inst_executed_op_fadd_pred_on.sum", 0) + m.get("smsp__sass_thread_inst_executed_op_fmul_pred_on.sum", 0)
              + 2 * m.get("smsp__sass_thread_inst_executed_op_ffma_pred_on.sum", 0))
        e = kernels.setdefault(name, {"flop": 0.0, "warp_inst": 0.0, "dram_bytes": 0.0, "ncu_ms": 0.0, "launches": 0, "issue_w": 0.0})
        e["flop"] += fl
        e["warp_inst"] += m.get("smsp__inst_executed.sum", 0)
        e["dram_bytes"] += scale_bytes(m, "dram__bytes_read.sum") + scale_bytes(m, "dram__bytes_write.sum")
        ms = scale_ms(m)
        e["ncu_ms"] += ms
        e["issue_w"] += ms * m.get("smsp__issue_active.avg.pct_of_peak_sustained_active", 0.0)
        e["launches"] += 1
    for e in kernels.values():
        e["issue_slot_utilisation"] = e.pop("issue_w") / e["ncu_ms"] / 100 if e["ncu_ms"] else None
    shadow = [n for n in kernels if "shade_setup" in n or "shadow" in n]
    prim = [n for n in kernels if "primary" in n]
    sh_flop = sum(kernels[n]["flop"] for n in shadow)
    sh_dram = sum(kernels[n]["dram_bytes"] for n in shadow)
    pr_flop = sum(kernels[n]["flop"] for n in prim)
    entry = {
        "source": f"{ops_csv} (ncu thread-level op counters, --clock-control none, last frame of tools/profile_target.py "
                  f"--workload {workload}); FLOP = fadd + fmul + 2*ffma, predicated on",
        "profiled_frame": frame,
        "kernels": kernels,
        "shadow_pass": {"kernels": shadow, "flop": sh_flop, "flop_per_hit_pixel": sh_flop / max(1, hits),
                        "dram_bytes_per_hit_pixel": sh_dram / max(1, hits)},
        "primary": {"kernels": prim, "flop": pr_flop, "flop_per_pixel": pr_flop / max(1, pixels)},
        "issue_slot_utilisation": {n: kernels[n]["issue_slot_utilisation"] for n in kernels},
        "exact_shadow_per_hit_pixel": frame.get("exact_shadow", 0) / max(1, hits),
        "cone_tests_per_hit_pixel": frame.get("beam_l2", 0) / max(1, hits),
    }
    try:
        table = json.load(open(out))
    except Exception:
        table = {}
    table[workload] = entry
    json.dump(table, open(out, "w"), indent=1)
    print(json.dumps({k: entry[k] for k in ("shadow_pass", "primary", "issue_slot_utilisation")}, indent=1))


if __name__ == "__main__":
    main()
