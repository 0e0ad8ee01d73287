#!/bin/bash
# On the GPU box: a few ncu counters of the tile kernel for every build_variants/*.so
for so in build_variants/*.so; do
  cp $so square_wheel_b200/libsqw_cuda.so
  n=$(basename $so .so)
  ncu --metrics gpu__time_duration.sum,smsp__inst_executed.sum,smsp__issue_active.avg.pct_of_peak_sustained_active,sm__warps_active.avg.pct_of_peak_sustained_active,smsp__thread_inst_executed_per_inst_executed.ratio,sm__cycles_active.avg,sm__cycles_elapsed.max \
    --clock-control none -k regex:raster_tile --launch-skip 40 -c 4 --csv --log-file gpurun_out/nv_$n.csv python bench.py --steps 1 --warmup 3 --no-cpu-baseline > /dev/null 2>&1
  python - <<PY
import csv
rows=[r for r in csv.reader(open("gpurun_out/nv_$n.csv")) if len(r)>10]
h=rows[0]; i_m=h.index("Metric Name"); i_v=h.index("Metric Value"); i_id=h.index("ID")
d={}
for r in rows[1:]:
    d.setdefault(r[i_m],[]).append(float(r[i_v].replace(",","")))
print("$n", {k.split(".")[0].replace("smsp__","").replace("sm__",""): round(sum(v)/len(v),1) for k,v in d.items()})
PY
done
