#!/bin/bash
# End-of-round evidence pass on one B200: every GPU test, smoke, the four bench configs, the
# launch list of one bench run and ncu --set full captures of the named kernels.
tag=${1:-r2}
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.build()" > gpurun_out/build.log 2>&1 || tail -20 gpurun_out/build.log
timeout 1200 python -m pytest tests -q -m gpu > gpurun_out/${tag}_gputests.log 2>&1; tail -3 gpurun_out/${tag}_gputests.log
timeout 300 python __graft_entry__.py smoke 2>&1 | tail -2 | tee gpurun_out/${tag}_smoke.log
for c in train r101 infer roi_nms; do
  timeout 900 python bench.py --config $c --steps 20 --warmup 3 > gpurun_out/${tag}_bench_${c}.json 2> gpurun_out/${tag}_bench_${c}.err
  echo "== bench $c rc=$?"; head -c 400 gpurun_out/${tag}_bench_${c}.json; echo; tail -2 gpurun_out/${tag}_bench_${c}.err
done
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 1500 --csv --log-file gpurun_out/${tag}_launches_step.csv python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-e2e --sustain 0 > gpurun_out/ncu_bench.log 2>&1; tail -1 gpurun_out/ncu_bench.log
for spec in ${NCU_SPECS:-conv:conv_gemm_tc conv1x1:conv_gemm_tc wgrad:conv_wgrad roi_cl:roi_align_cl2_fwd roi_cl_2:roi_align_cl_bwd roi:roi_align_nhwc_fwd roi_2:roi_align_nhwc_bwd}; do
  tgt=${spec%%:*}; rx=${spec#*:}
  timeout 600 ncu --set full --clock-control none --import-source on -k regex:$rx -s 2 -c 1 -o gpurun_out/ncu_${tgt}_${tag} -f python tools/ncu_targets.py $tgt > gpurun_out/ncu_${tgt}.log 2>&1; tail -1 gpurun_out/ncu_${tgt}.log
done
timeout 300 ncu --set full --clock-control none --import-source on -k regex:nms_sweep -s 2 -c 1 -o gpurun_out/ncu_nms_sweep_${tag} -f python tools/ncu_nms.py 2000 > gpurun_out/ncu_nms.log 2>&1; tail -1 gpurun_out/ncu_nms.log
ls -la gpurun_out/*_${tag}.ncu-rep
