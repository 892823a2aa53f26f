"""Per-source-line samples / executed instructions from `ncu -i rep --page source --csv --print-source cuda,sass`.
Usage: python tools/ncu_lines.py <rep> [min_pct] -> prints file:line rows (source lines only) sorted by position."""
import csv, subprocess, sys
rep = sys.argv[1]; min_pct = float(sys.argv[2]) if len(sys.argv) > 2 else 0.4
raw = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass"], capture_output=True, text=True).stdout
rows = list(csv.reader(raw.splitlines()))
cur = None; hdr = None; out = []
for r in rows:
    if not r: continue
    if r[0] == "File Path": cur = r[1].split("/")[-1]; continue
    if r[0] == "Function Name": continue
    if r[0] == "Line No": hdr = r; ix = {}; [ix.setdefault(h, i) for i, h in enumerate(hdr)]; continue
    if hdr and r[0].isdigit():
        def I(k):
            try: return int(r[ix[k]])
            except Exception: return 0
        out.append((cur, int(r[0]), r[1].strip()[:90], I("# Samples"), I("Instructions Executed"), I("Thread Instructions Executed")))
tot = sum(o[3] for o in out) or 1; toti = sum(o[4] for o in out) or 1
print(f"total samples {tot} instr {toti/1e6:.1f}M thread-instr/instr {sum(o[5] for o in out)/toti:.1f}")
agg = {}
for f, l, s, sm, ie, te in out:
    a = agg.setdefault((f, l), [s, 0, 0, 0]); a[1] += sm; a[2] += ie; a[3] += te
for (f, l), (s, sm, ie, te) in sorted(agg.items()):
    if 100 * sm / tot >= min_pct or 100 * ie / toti >= min_pct:
        print(f"{f}:{l:5d} smp {100*sm/tot:5.1f}% ins {100*ie/toti:5.1f}% act {te/max(ie,1):4.1f} | {s}")
