#!/bin/bash
# usage: tools/gpu_env_ab.sh VAR [bench args]   -- same-box A/B of bench.py with VAR=0 and VAR=1, twice each
var=$1; shift
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.build()" > gpurun_out/build.log 2>&1 || tail -5 gpurun_out/build.log
for rep in 1 2; do for v in 0 1; do
  env $var=$v python bench.py --no-cpu-baseline --no-e2e --sustain 0 "$@" > gpurun_out/envab_${v}_$rep.json 2> gpurun_out/envab_${v}_$rep.err
  python - <<PY
import json
try:
    d=json.loads(open('gpurun_out/envab_${v}_$rep.json').read().strip().splitlines()[-1]); r=d['roofline']
    print('$var=$v rep $rep: %.2f ms  conv %.0f TF  tensor %.0f  hbm %.0f GB/s (%.2f ms)  wgrad %.0f  clk %s loss %.4f' % (d['ms_per_step'], r['achieved'], r['split']['tensor_bound_launches']['achieved'], r['split']['hbm_bound_launches']['achieved'], r['split']['hbm_bound_launches']['ms_per_step'], r['wgrad']['achieved'], d['clocks']['sm_mhz'], d['loss_last']))
except Exception as e:
    print('no line', e); print(open('gpurun_out/envab_${v}_$rep.err').read()[-800:])
PY
done; done
