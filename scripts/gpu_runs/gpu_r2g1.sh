#!/bin/bash
mkdir -p gpurun_out
R=/tmp/ncu; mkdir -p $R
HVX_FLOW_PRECISE=1 timeout -k 10 900 ncu --set full --clock-control none --import-source on -k regex:gemm_pair3_kernel -s 9 -c 5 -o $R/ff1 -f python scripts/prof_flow.py 1 > gpurun_out/r2g1_ncu.log 2>&1
tail -2 gpurun_out/r2g1_ncu.log
ncu -i $R/ff1.ncu-rep --page source --csv --print-source sass 2>/dev/null > $R/src.csv
python - <<'P'
import csv
allrows=list(csv.reader(open('/tmp/ncu/src.csv')))
starts=[i for i,r in enumerate(allrows) if r and r[0]=="Kernel Name"]
print("kernels in report:", [allrows[i][1][:60] for i in starts])
pick=[i for i in starts if "(int)0, (int)1" in allrows[i][1]]
i0=pick[0]; i1=min([j for j in starts if j>i0]+[len(allrows)])
rows=allrows[i0:i1]
hdr=rows[1]
si=hdr.index("# Samples"); ie=hdr.index("Instructions Executed")
data=[]
for k,r in enumerate(rows[2:]):
    if len(r)<=ie: continue
    try: data.append((int(r[si]), int(r[ie]), k, r[1].strip(), r))
    except: pass
tot=sum(d[0] for d in data)
out=open('gpurun_out/r2g1_ff1_source_top.txt','w')
print("kernel:", rows[0][1][:120], file=out)
print("total samples", tot, "sass rows", len(data), file=out)
agg={}
for d in data:
    for i,h in enumerate(hdr):
        if h.startswith("stall_") and "Not Issued" not in h:
            try: v=int(d[4][i])
            except: continue
            if v: agg[h]=agg.get(h,0)+v
for k,v in sorted(agg.items(), key=lambda kv:-kv[1])[:14]: print(f"   {v:8d}  {k}", file=out)
print("top 50 SASS rows by samples:", file=out)
for d in sorted(data, key=lambda x:-x[0])[:50]:
    print(f"{100*d[0]/max(tot,1):5.1f}%  exec {d[1]:8d}  row {d[2]:5d}  {d[3][:100]}", file=out)
print("samples by 40-row region:", file=out)
for i in range(0,len(data),40):
    seg=data[i:i+40]; s=sum(x[0] for x in seg)
    if s*200>tot: print(f"rows {i:5d}-{i+39:5d}: {100*s/max(tot,1):5.1f}%  inst {sum(x[1] for x in seg):9d}  first: {seg[0][3][:70]}", file=out)
out.close()
P
head -75 gpurun_out/r2g1_ff1_source_top.txt
