#!/bin/bash
# Full verification pass: GPU parity tests, smoke, bench line, ncu launch list of one step.
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.build()" > gpurun_out/build.log 2>&1 || tail -20 gpurun_out/build.log
timeout 900 python -m pytest tests -q -m gpu -x 2>&1 | tail -8
timeout 300 python __graft_entry__.py smoke 2>&1 | tail -3
timeout 600 python bench.py --steps 10 --warmup 3 > gpurun_out/bench_${1:-cur}.json 2> gpurun_out/bench_${1:-cur}.err; tail -c 3000 gpurun_out/bench_${1:-cur}.json; tail -3 gpurun_out/bench_${1:-cur}.err
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 1500 --csv --log-file gpurun_out/launches_${1:-cur}.csv python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-e2e > gpurun_out/ncu_bench.log 2>&1; tail -2 gpurun_out/ncu_bench.log
