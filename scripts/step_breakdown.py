import csv, re, sys, collections
with open(sys.argv[1]) as f:
    lines = [l for l in f if l.startswith('"')]
r = csv.reader(lines); hdr = next(r)
ki, vi = hdr.index("Kernel Name"), hdr.index("Metric Value")
gi = hdr.index("Grid Size"); bi = hdr.index("Block Size")
rows = [(re.sub(r"\(.*", "", x[ki]).replace("void hvx::","").replace("hvx::",""), float(x[vi].replace(",",""))/1000.0, x[gi], x[bi]) for x in r if len(x) > vi]
idx = [i for i,x in enumerate(rows) if x[0].startswith("llm_sampler")]
a, b = idx[-2]+1, idx[-1]+1
step = rows[a:b]
print("kernels in last step:", len(step), "sum us %.1f" % sum(x[1] for x in step))
agg = collections.OrderedDict()
for x in step:
    k=(x[0],x[2],x[3]); agg.setdefault(k,[0,0.0]); agg[k][0]+=1; agg[k][1]+=x[1]
for k,v in sorted(agg.items(), key=lambda kv:-kv[1][1]): print(f"{v[1]:8.1f} us n={v[0]:3d} avg {v[1]/v[0]:7.2f}  {k}")
