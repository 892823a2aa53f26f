"""Stall-reason breakdown per block of SASS instructions from `ncu -i rep --page source --csv --print-source sass` output.
Usage: python tools/ncu_stalls.py <csv> <work items (for per-item instruction counts)> [block size]"""
import csv,sys
rows=list(csv.reader(open(sys.argv[1])))
items=float(sys.argv[2]); blk=int(sys.argv[3]) if len(sys.argv)>3 else 100
hi=[i for i,r in enumerate(rows) if r and r[0]=="Address"][0]
hdr=rows[hi]; data=rows[hi+1:]
ix={h:i for i,h in enumerate(hdr)}
def I(r,k):
    try: return int(r[ix[k]])
    except: return 0
tot=sum(I(r,'# Samples') for r in data)
keys=['stall_no_inst','stall_wait','stall_long_sb','stall_branch_resolving','stall_short_sb','stall_not_selected','stall_selected','stall_dispatch','stall_math','stall_mio','stall_lg','stall_barrier']
print("total",tot,{k:round(100*sum(I(r,k) for r in data)/tot,1) for k in keys})
ex=[I(r,'Instructions Executed') for r in data]
print(len(data), "hot>=0.5/item:", sum(1 for e in ex if e>=0.5*items), "toti", sum(ex)/1e6, "per item", sum(ex)/items)
for b in range(0,len(data),blk):
    seg=data[b:b+blk]
    sm=sum(I(r,'# Samples') for r in seg)
    ie=sum(I(r,'Instructions Executed') for r in seg)
    if sm>0.01*tot:
        print(f"{b:5d}: smp {100*sm/tot:5.1f}% instr/item {ie/items:7.1f} ", {k.replace('stall_',''):round(100*sum(I(r,k) for r in seg)/tot,1) for k in keys[:6]})
