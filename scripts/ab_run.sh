#!/bin/bash
# A/B on the GPU box: for every variant library run quick_bench (one line per variant into gpurun_out/ab_<tag>.log)
# usage: scripts/ab_run.sh <tag> "<quick_bench args>" name1 name2 ...
tag=$1; shift; args=$1; shift
for v in "$@"; do
  lib=path-tracing__ray-tracer_b200/build/variants/libb200rt_$v.so
  echo -n "$v: " >> gpurun_out/ab_$tag.log
  B200RT_LIB=$lib python scripts/quick_bench.py $args 2>&1 | tail -1 >> gpurun_out/ab_$tag.log
done
cat gpurun_out/ab_$tag.log
