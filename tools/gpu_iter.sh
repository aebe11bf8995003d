#!/bin/bash
# Iteration pass: selected GPU tests + bench line.  usage: gpu_iter.sh <tag> [pytest args...]
tag=${1:-cur}; shift
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.build()" > gpurun_out/build.log 2>&1 || tail -20 gpurun_out/build.log
timeout 900 python -m pytest "${@:-tests}" -q -m gpu -x 2>&1 | tail -15
timeout 600 python bench.py --steps 10 --warmup 3 > gpurun_out/bench_${tag}.json 2> gpurun_out/bench_${tag}.err; tail -c 3500 gpurun_out/bench_${tag}.json; tail -5 gpurun_out/bench_${tag}.err
