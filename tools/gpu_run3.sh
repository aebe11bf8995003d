#!/bin/bash
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.build()" > gpurun_out/build.log 2>&1
echo "== pytest wgrad"
timeout 300 python -m pytest tests/test_gpu_wgrad_tc.py -q -m gpu -x 2>&1 | tail -30
