#!/bin/bash
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.build()" > gpurun_out/build.log 2>&1
echo "== pytest roi"
timeout 900 python -m pytest tests/test_gpu_roi_align.py -q -m gpu 2>&1 | tail -15
echo "== pytest wgrad"
timeout 300 python -m pytest tests/test_gpu_wgrad_tc.py -q -m gpu 2>&1 | tail -40
echo "== microbench roi"
timeout 600 python tools/microbench.py roi --out gpurun_out/micro_roi2.jsonl > gpurun_out/micro_roi2.log 2>&1; grep -E '"R": 1000' gpurun_out/micro_roi2.log | cut -c1-200
