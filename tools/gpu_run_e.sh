#!/bin/bash
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.build()" > gpurun_out/build.log 2>&1 || tail -20 gpurun_out/build.log
timeout 900 python -m pytest tests/test_gpu_targets.py tests/test_gpu_model.py -q -m gpu -x 2>&1 | tail -${1:-40} | tee gpurun_out/pytest_e.log
echo "== bench"
timeout 900 python bench.py --steps 10 --warmup 3 > gpurun_out/bench.json 2> gpurun_out/bench.err; tail -5 gpurun_out/bench.err; cat gpurun_out/bench.json
