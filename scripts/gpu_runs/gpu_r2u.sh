#!/bin/bash
mkdir -p gpurun_out
R=/tmp/ncu; mkdir -p $R
timeout -k 10 900 ncu --set full --clock-control none --import-source on -k regex:dit_attention_v5 -s 20 -c 1 -o $R/dattn -f python scripts/prof_flow.py 1 > gpurun_out/r2u_ncu.log 2>&1
tail -2 gpurun_out/r2u_ncu.log
ncu -i $R/dattn.ncu-rep --page source --csv --print-source sass 2>/dev/null > $R/src.csv
python - <<'P'
import csv
rows=list(csv.reader(open('/tmp/ncu/src.csv')))
hdr=rows[1]
si=hdr.index("# Samples"); ie=hdr.index("Instructions Executed")
cols=[i for i,h in enumerate(hdr) if h.startswith("stall_") or "Stall" in h]
data=[]
for k,r in enumerate(rows[2:]):
    if len(r)<=ie: continue
    try: data.append((int(r[si]), int(r[ie]), k, r[1].strip(), r))
    except: pass
tot=sum(d[0] for d in data)
out=open('gpurun_out/r2u_attn_source_top.txt','w')
print("total samples", tot, "sass rows", len(data), file=out)
print("header:", [h for h in hdr[:12]], file=out)
stall_cols=[i for i,h in enumerate(hdr) if i>ie+5]
# aggregate stall reason columns over all rows
agg={}
for d in data:
    for i in stall_cols:
        try: v=int(d[4][i])
        except: continue
        if v: agg[hdr[i]]=agg.get(hdr[i],0)+v
print("aggregate of per-row counters (top 25):", file=out)
for k,v in sorted(agg.items(), key=lambda kv:-kv[1])[:25]: print(f"   {v:10d}  {k}", file=out)
print("top 60 SASS rows by samples:", file=out)
for d in sorted(data, key=lambda x:-x[0])[:60]:
    print(f"{100*d[0]/max(tot,1):5.1f}%  exec {d[1]:8d}  row {d[2]:5d}  {d[3][:100]}", file=out)
print("samples by 50-row region:", file=out)
for i in range(0,len(data),50):
    seg=data[i:i+50]; s=sum(x[0] for x in seg)
    if s: print(f"rows {i:5d}-{i+49:5d}: {100*s/max(tot,1):5.1f}%  first: {seg[0][3][:60]}", file=out)
out.close()
P
head -40 gpurun_out/r2u_attn_source_top.txt
