#!/bin/bash
# Session D: ncu launch list of one bench step (per-kernel device times).
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.build()" > gpurun_out/build.log 2>&1 || tail -20 gpurun_out/build.log
timeout 1200 ncu --metrics gpu__time_duration.sum --clock-control none -s ${1:-1100} -c ${2:-420} --csv \
  --log-file gpurun_out/launches.csv python bench.py --steps 1 --warmup 3 --no-e2e --no-cpu-baseline > gpurun_out/ncu_bench.log 2>&1
tail -2 gpurun_out/ncu_bench.log | cut -c1-300
wc -l gpurun_out/launches.csv
