#!/bin/bash
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.build()" > gpurun_out/build.log 2>&1 || tail -20 gpurun_out/build.log
timeout 1300 python -m pytest "$@" -q -m gpu > gpurun_out/t1_full.log 2>&1; grep -n "^FAILED\|^ERROR\|passed\|failed" gpurun_out/t1_full.log | tail -40
