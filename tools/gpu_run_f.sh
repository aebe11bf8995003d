#!/bin/bash
# conv kernel iteration: parity tests, model tests, per-layer timing.
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.build()" > gpurun_out/build.log 2>&1 || tail -20 gpurun_out/build.log
timeout 600 python -m pytest tests/test_gpu_conv_tc.py tests/test_gpu_wgrad_tc.py -q -m gpu -x 2>&1 | tail -15
timeout 600 python -m pytest tests/test_gpu_model.py -q -m gpu -x 2>&1 | tail -8
timeout 300 python tools/layer_bench.py --out gpurun_out/layer_bench_${1:-v2}.json 2>&1 | head -${2:-45}
