#!/bin/bash
# usage: tools/roi_ab.sh "<nvcc -D flags A>" "<flags B>" ...   -- same-box A/B of bench.py --config roi_nms per build
mkdir -p gpurun_out
i=0
for rep in 1 2; do for flags in "$@"; do
  i=$((i+1))
  CMR_EXTRA_NVCC_FLAGS="$flags" python -c "import __graft_entry__ as g; g.build()" > gpurun_out/build.log 2>&1 || tail -5 gpurun_out/build.log
  CMR_EXTRA_NVCC_FLAGS="$flags" python bench.py --config roi_nms --steps 20 --no-cpu-baseline > gpurun_out/roi_ab_$i.json 2> gpurun_out/roi_ab_$i.err
  CMR_EXTRA_NVCC_FLAGS="$flags" python -c "
import json
d=json.loads(open('gpurun_out/roi_ab_$i.json').read().strip().splitlines()[-1])
print('[$flags]', ' '.join('R%d fwd %.0f bwd %.0f |' % (k['R'],k['fwd_kernel_gbs'],k['bwd_kernel_gbs']) for k in d['cases'] if k['op']=='roi_align_2d' and k['map']=='nchw'))"
done; done
