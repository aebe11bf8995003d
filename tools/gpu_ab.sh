#!/bin/bash
# usage: tools/gpu_ab.sh [bench args]   -- same-box A/B of bench.py: an older conv_tc.cu (put it at
# tools/ab/conv_tc_old.cu.txt first: git show REF:chainer_mask_rcnn_b200/csrc/conv_tc.cu, plus stubs for
# entry points REF lacks) vs the tree's conv_tc.cu, twice each
mkdir -p gpurun_out
run() {
  python -c "import __graft_entry__ as g; g.build()" > gpurun_out/build.log 2>&1 || tail -5 gpurun_out/build.log
  python bench.py --no-cpu-baseline --no-e2e --sustain 0 "$@" > gpurun_out/ab_$tag.json 2> gpurun_out/ab_$tag.err
  python - <<PY
import json
d=json.loads(open('gpurun_out/ab_$tag.json').read().strip().splitlines()[-1]); r=d['roofline']
print('$tag: %.2f ms  conv %.0f TF  tensor %.0f  hbm %.0f GB/s (%.2f ms)  wgrad %.0f  clk %s' % (d['ms_per_step'], r['achieved'], r['split']['tensor_bound_launches']['achieved'], r['split']['hbm_bound_launches']['achieved'], r['split']['hbm_bound_launches']['ms_per_step'], r['wgrad']['achieved'], d['clocks']['sm_mhz']))
PY
}
src=chainer_mask_rcnn_b200/csrc/conv_tc.cu
cp $src /tmp/conv_tc_new.cu
tag=new1; run "$@"
cp tools/ab/conv_tc_old.cu.txt $src; tag=old1; run "$@"
cp /tmp/conv_tc_new.cu $src; tag=new2; run "$@"
cp tools/ab/conv_tc_old.cu.txt $src; tag=old2; run "$@"
cp /tmp/conv_tc_new.cu $src
