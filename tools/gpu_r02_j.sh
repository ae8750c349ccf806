#!/bin/bash
# round-2 GPU call J (1 GPU): per-launch time + DRAM bytes of the CG and BiCGStab chains in context
mkdir -p gpurun_out
for w in cg bicgstab; do
  timeout 300 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none --csv --log-file gpurun_out/r02j_$w.csv python tools/cg_probe.py $w 24 > gpurun_out/r02j_$w.log 2>&1
  tail -2 gpurun_out/r02j_$w.log
  python - <<PY
import csv, collections, re, statistics
rows=[r for r in csv.reader(open('gpurun_out/r02j_$w.csv', errors='replace')) if len(r)>10]
h=rows[0]; ki,mi,vi,ui=h.index('Kernel Name'),h.index('Metric Name'),h.index('Metric Value'),h.index('Metric Unit')
per=collections.defaultdict(lambda: collections.defaultdict(list))
for r in rows[1:]:
    try: v=float(r[vi].replace(',',''))
    except ValueError: continue
    name=re.sub(r'\(.*','',re.sub(r'\(anonymous namespace\)::|<unnamed>::','',r[ki]))[:40]
    u=r[ui]
    if 'time' in r[mi]: v*= {'ns':1e-3,'us':1,'ms':1e3,'nsecond':1e-3,'usecond':1,'msecond':1e3}.get(u,1e-3)
    else: v*= {'byte':1,'Kbyte':1e3,'Mbyte':1e6,'Gbyte':1e9}.get(u,1)
    per[name][r[mi]].append(v)
for k,m in per.items():
    t=m['gpu__time_duration.sum']; rd=m['dram__bytes_read.sum']; wr=m['dram__bytes_write.sum']
    print(f"{k:40s} n={len(t):4d} med {statistics.median(t):8.1f} us  rd {statistics.median(rd)/1e6:8.1f} MB wr {statistics.median(wr)/1e6:8.1f} MB  -> {(statistics.median(rd)+statistics.median(wr))/statistics.median(t)/1e3:7.1f} GB/s")
PY
done
