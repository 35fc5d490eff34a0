#!/bin/bash
# per-launch durations of one quick_bench run for a variant library: scripts/launch_list.sh <variant> <tag> "<args>"
lib=path-tracing__ray-tracer_b200/build/variants/libb200rt_$1.so
B200RT_LIB=$lib ncu --metrics gpu__time_duration.sum --clock-control none -c 60 --csv --log-file gpurun_out/launches_$2.csv python scripts/quick_bench.py $3 > gpurun_out/launches_$2.log 2>&1
python - <<PY
import csv,collections
rows=[r for r in csv.reader(open("gpurun_out/launches_$2.csv")) if len(r)>5]
hdr=rows[0]; ik=hdr.index("Kernel Name"); iv=hdr.index("Metric Value")
for r in rows[1:40]:
    print("%-60s %10.1f us" % (r[ik][:60], float(r[iv].replace(",",""))/1000))
PY
