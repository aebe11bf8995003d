#!/bin/bash
# Session A: existing parity tests, microbenchmarks, cuBLAS TF32 peak.
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/smi.txt 2>&1
python -c "import os; print('cpus', os.cpu_count())" >> gpurun_out/smi.txt
python -c "import __graft_entry__ as g; g.build()" > gpurun_out/build.log 2>&1
echo "== pytest gpu"
timeout 900 python -m pytest tests -q -m gpu -x 2>&1 | tail -15 | tee gpurun_out/pytest_a.log
echo "== peaks"
timeout 300 python tools/microbench.py peaks --out gpurun_out/micro_peaks.jsonl 2>&1 | tail -5
echo "== conv"
timeout 300 python tools/microbench.py conv --out gpurun_out/micro_conv.jsonl > gpurun_out/micro_conv.log 2>&1; cut -c1-230 gpurun_out/micro_conv.log | tail -40
echo "== roi"
timeout 600 python tools/microbench.py roi --out gpurun_out/micro_roi.jsonl > gpurun_out/micro_roi.log 2>&1; grep -E '"R": 1000' gpurun_out/micro_roi.log | cut -c1-220
echo "== nms"
timeout 300 python tools/microbench.py nms --out gpurun_out/micro_nms.jsonl > gpurun_out/micro_nms.log 2>&1; cut -c1-220 gpurun_out/micro_nms.log | tail -20
