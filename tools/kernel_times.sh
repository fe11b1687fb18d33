#!/bin/sh
# pure durations of the 4 frame kernels on warm caches (ncu launch list, later frames of the trajectory)
ncu --metrics gpu__time_duration.sum --clock-control none --cache-control none -k regex:"k_project|k_scatter|k_column|k_fuse" -s 12 -c 12 --csv --log-file gpurun_out/launches_tmp.csv python tools/profile_frame.py 8 > /dev/null 2>&1
python - <<PY
import csv
rows=list(csv.reader(open("gpurun_out/launches_tmp.csv")))
i=[k for k,r in enumerate(rows) if r and r[0]=="ID"][0]
h=rows[i]
acc={}
for r in rows[i+1:]:
    if len(r)>5:
        acc.setdefault(r[h.index("Kernel Name")].split("(")[0],[]).append(float(r[h.index("Metric Value")].replace(",",""))/1000)
for k,v in acc.items(): print(f"{k:24s} us: "+" ".join(f"{x:.1f}" for x in v))
print("sum of means: %.1f us" % sum(sum(v)/len(v) for v in acc.values()))
PY
