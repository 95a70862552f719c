#!/bin/bash
# usage: tools/nccl_sweep.sh N "cfg1" "cfg2" ...   where cfg is a list of ENV=VALUE words ("" = defaults)
n=$1; shift
i=0
for cfg in "$@"; do
  i=$((i+1))
  env $cfg timeout 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 2970$i bench.py --gpus $n --steps 20 --warmup 5 --no-extras --no-cpu-baseline > gpurun_out/nccl_${n}_$i.json 2> gpurun_out/nccl_${n}_$i.err
  python - <<PY
import json
try:
    l=json.loads([x for x in open('gpurun_out/nccl_${n}_$i.json') if x.startswith('{')][0])
    sp=l.get('step_profile',{})
    print('N=$n [$cfg] ms_per_step', l['ms_per_step'], 'nccl active/exposed', sp.get('nccl_active_ms'), sp.get('nccl_exposed_ms'), 'span', sp.get('span_ms'), 'dp ok', l.get('dp_check',{}).get('ok'))
except Exception as e:
    print('N=$n [$cfg] failed', e)
PY
done
