#!/bin/bash
# Session B: model-level parity tests.
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.build()" > gpurun_out/build.log 2>&1 || tail -20 gpurun_out/build.log
timeout 900 python -m pytest tests/test_gpu_model.py -q -m gpu -x 2>&1 | tail -${1:-60} | tee gpurun_out/pytest_b.log
