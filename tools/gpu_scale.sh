#!/bin/bash
# usage: tools/gpu_scale.sh <tag> <N> [bench args...]   -- one torchrun bench line on N GPUs of this box
tag=$1; n=$2; shift 2
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.build()" > gpurun_out/build.log 2>&1 || tail -20 gpurun_out/build.log
if [ "$n" = "1" ]; then
  timeout 420 python bench.py --gpus 1 "$@" > gpurun_out/scale_${tag}_n1.json 2> gpurun_out/scale_${tag}_n1.err
else
  timeout 420 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus $n "$@" > gpurun_out/scale_${tag}_n$n.json 2> gpurun_out/scale_${tag}_n$n.err
fi
echo "== $tag n=$n rc=$?"
python - <<PY
import json
try:
    d=json.loads(open('gpurun_out/scale_${tag}_n$n.json').read().strip().splitlines()[-1])
    print('value %.1f images/s  ms/step %.2f  e2e %s  sustained %s' % (d['value'], d['ms_per_step'], d.get('e2e',{}).get('value'), d.get('sustained',{}).get('value')))
except Exception as e:
    print('no line:', e)
PY
tail -4 gpurun_out/scale_${tag}_n$n.err
