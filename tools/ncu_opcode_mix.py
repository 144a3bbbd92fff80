#!/usr/bin/env python
"""Dynamic opcode mix, shared-memory wavefronts and stall split of one kernel from the SASS source page of an
.ncu-rep.  usage: ncu -i x.ncu-rep --page source --csv --print-source sass --kernel-name regex:NAME > k.csv;
python tools/ncu_opcode_mix.py k.csv"""
import csv, sys, collections
rows = list(csv.reader(open(sys.argv[1])))
hdr = rows[1]
iS = hdr.index("Source"); iI = hdr.index("Instructions Executed"); iN = hdr.index("# Samples")
iW = hdr.index("L1 Wavefronts Shared"); iWi = hdr.index("L1 Wavefronts Shared Ideal")
stall_cols = [(i,h) for i,h in enumerate(hdr) if h.startswith("stall_") and "Not Issued" not in h]
opc = collections.defaultdict(lambda: [0,0,0,0]); tot=0; tots=0
st = collections.Counter()
end = next((i for i in range(2,len(rows)) if rows[i] and rows[i][0]=="Kernel Name"), len(rows))
for r in rows[2:end]:
    if len(r) < len(hdr): continue
    s = r[iS].split()
    if not s: continue
    op = s[1] if s[0].startswith('@') else s[0]
    full = op
    op = op.split('.')[0]
    if op in ("LDS","STS"): op = full.rstrip(';')
    ie = int(r[iI]); sm = int(r[iN])
    opc[op][0] += ie; opc[op][1] += sm; opc[op][2] += int(r[iW]); opc[op][3] += int(r[iWi])
    tot += ie; tots += sm
    for i,h in stall_cols: st[h] += int(r[i])
print("total warp inst", tot, "samples", tots)
for k, v in sorted(opc.items(), key=lambda kv: -kv[1][0])[:22]:
    print(f"{k:16s} inst={v[0]:>12d} {100*v[0]/tot:5.1f}% samp={100*v[1]/tots:5.1f}%")
print({k: round(100*v/tots,1) for k,v in st.most_common(12)})
