#!/bin/bash
# ncu --set full captures.  usage: gpu_ncu.sh <tag> <target>:<kernel-regex> ...
tag=$1; shift; export CMR_ROI_CL_CTAS=${CMR_ROI_CL_CTAS:-2}
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.build()" > gpurun_out/build.log 2>&1 || tail -20 gpurun_out/build.log
for spec in "$@"; do
  tgt=${spec%%:*}; rx=${spec#*:}
  timeout 600 ncu --set full --clock-control none --import-source on -k regex:$rx -s 2 -c 1 -o gpurun_out/ncu_${tgt}_${tag} -f python tools/ncu_targets.py $tgt > gpurun_out/ncu_${tgt}.log 2>&1; tail -1 gpurun_out/ncu_${tgt}.log
done
ls -la gpurun_out/*.ncu-rep
