#!/bin/bash
# ncu --set full captures of the two tensor-core kernels.
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.build()" > gpurun_out/build.log 2>&1 || tail -20 gpurun_out/build.log
timeout 600 ncu --set full --clock-control none --import-source on -k regex:conv_gemm_tc -s 2 -c 1 -o gpurun_out/ncu_conv_gemm_${1:-v2} -f python tools/ncu_targets.py conv > gpurun_out/ncu_g1.log 2>&1; tail -2 gpurun_out/ncu_g1.log
timeout 600 ncu --set full --clock-control none --import-source on -k regex:conv_wgrad_tc -s 2 -c 1 -o gpurun_out/ncu_wgrad_${1:-v2} -f python tools/ncu_targets.py wgrad > gpurun_out/ncu_g2.log 2>&1; tail -2 gpurun_out/ncu_g2.log
ls -la gpurun_out/*.ncu-rep
