#!/bin/bash
# tests given as args, then the bench configs named in $BENCHES (space separated)
tools/gpu_t1.sh "$@"
for c in $BENCHES; do
  timeout 900 python bench.py --config $c --steps ${STEPS:-10} --warmup 3 > gpurun_out/bench_${c}_${TAG:-cur}.json 2> gpurun_out/bench_${c}.err
  echo "== bench $c rc=$?"; tail -c 2500 gpurun_out/bench_${c}_${TAG:-cur}.json; tail -5 gpurun_out/bench_${c}.err
done
