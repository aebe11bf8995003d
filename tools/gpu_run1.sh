#!/bin/bash
# First GPU session: parity tests (safe kernels first), microbench, tcgen05 test, ncu.
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/smi.txt 2>&1
python -c "import __graft_entry__ as g; g.build()" > gpurun_out/build.log 2>&1
echo "== pytest roi/nms" 
timeout 900 python -m pytest tests/test_gpu_roi_align.py tests/test_gpu_nms.py -q -m gpu 2>&1 | tail -40 | tee gpurun_out/pytest_roi_nms.log
echo "== smoke"
timeout 300 python __graft_entry__.py smoke 2>&1 | tail -5 | tee gpurun_out/smoke.log
echo "== microbench roi/nms"
timeout 600 python tools/microbench.py roi --out gpurun_out/micro_roi.jsonl > gpurun_out/micro_roi.log 2>&1; tail -3 gpurun_out/micro_roi.log
timeout 300 python tools/microbench.py nms --out gpurun_out/micro_nms.jsonl > gpurun_out/micro_nms.log 2>&1; tail -3 gpurun_out/micro_nms.log
echo "== pytest conv_tc"
timeout 300 python -m pytest tests/test_gpu_conv_tc.py -q -m gpu 2>&1 | tail -60 | tee gpurun_out/pytest_conv.log
echo "== microbench conv"
timeout 300 python tools/microbench.py conv --out gpurun_out/micro_conv.jsonl > gpurun_out/micro_conv.log 2>&1; tail -30 gpurun_out/micro_conv.log
echo "== ncu roi"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:roi_align_nchw_fwd -s 3 -c 1 -o gpurun_out/roi_fwd python tools/microbench.py roi > gpurun_out/ncu_roi.log 2>&1
ls -la gpurun_out
