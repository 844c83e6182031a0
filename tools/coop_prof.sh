#!/bin/bash
# usage: tools/coop_prof.sh <tag>   (runs on the GPU box through gpurun, prints the consumer loop)
tag=$1
/usr/local/graft/bin/gpurun --timeout 600 -- "ncu --set full --clock-control none --import-source on -k regex:rans_decode_coop -s 1 -c 1 -o gpurun_out/prof_$tag python tools/prof_coop.py > gpurun_out/ncu_$tag.log 2>&1; tail -1 gpurun_out/ncu_$tag.log" 2>&1 | tail -2
python tools/ncu_sass_hot.py gpurun_out/prof_$tag.ncu-rep rans_decode_coop 400000 0 > /tmp/coop_hot_$tag.txt 2>&1
python - <<PY
rows=[]
for l in open('/tmp/coop_hot_$tag.txt').read().splitlines()[1:]:
    p=l.split()
    try:
        a=int(p[0],16); w=float(p[1]); smp=int(p[2])
    except: continue
    rows.append((a,w,smp,l[:120]))
print("total", sum(r[2] for r in rows))
cons=[r for r in rows if 30<=r[1]<=33 or 60<=r[1]<=64]
print(len(cons), sum(r[2] for r in cons))
for r in cons: print(r[3])
PY
