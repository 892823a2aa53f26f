"""Condense an .ncu-rep (read on the CPU box) into the JSON summary kept under profiles/."""
import csv
import json
import subprocess
import sys

KEEP = [
    "gpu__time_duration.sum", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
    "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active", "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_elapsed",
    "sm__pipe_fmaheavy_cycles_active.avg.pct_of_peak_sustained_elapsed", "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active",
    "smsp__issue_active.avg.pct_of_peak_sustained_active", "sm__warps_active.avg.pct_of_peak_sustained_active",
    "launch__registers_per_thread", "launch__grid_size", "launch__block_size", "launch__occupancy_limit_registers",
    "dram__bytes_read.sum", "dram__bytes_write.sum", "dram__bytes_write.sum.per_second", "dram__bytes_read.sum.per_second",
    "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
    "smsp__inst_executed.sum", "sm__cycles_elapsed.avg.per_second", "sm__inst_executed.avg.per_cycle_elapsed",
    "smsp__thread_inst_executed_per_inst_executed.ratio", "lts__t_bytes.sum", "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum",
    # instruction caches: SM-level hit rate and the GPC-level cache's request rate against its peak
    "sm__icc_request_hit_rate.pct", "sm__icc_requests.sum.pct_of_peak_sustained_elapsed",
    "gcc__cache_requests_type_instruction.sum.pct_of_peak_sustained_elapsed", "gcc__average_cache_request_hit_rate.pct",
]


def to_bytes(v, unit):
    m = {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}
    return float(v.replace(",", "")) * m.get(unit, 1)


def main():
    rep, out, note = sys.argv[1], sys.argv[2], (sys.argv[3] if len(sys.argv) > 3 else "")
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(raw.splitlines()))
    hdr, units = rows[0], rows[1]
    # several launches in the report: summarise the longest one
    ti = hdr.index("gpu__time_duration.sum")
    def dur(r):
        try:
            return float(r[ti].replace(",", ""))
        except Exception:
            return 0.0
    vals = max((r for r in rows[2:] if len(r) == len(hdr)), key=dur)
    d = {"report": rep, "note": note, "kernel": vals[hdr.index("Kernel Name")], "metrics": {}}
    col = {}
    for h, u, v in zip(hdr, units, vals):
        col[h] = (v, u)
        if h in KEEP or ("issue_stalled" in h and "per_issue_active" in h):
            d["metrics"][h] = f"{v} {u}".strip()
    rd = to_bytes(*col["dram__bytes_read.sum"])
    wr = to_bytes(*col["dram__bytes_write.sum"])
    d["dram_bytes_per_launch"] = rd + wr
    json.dump(d, open(out, "w"), indent=1)
    print(out, d["kernel"], d["metrics"].get("gpu__time_duration.sum"), "dram", rd + wr)


if __name__ == "__main__":
    main()
