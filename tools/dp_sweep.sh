#!/bin/bash
# N-GPU sweep of the GEMM grid cap used while the all-reduce is in flight.  usage: tools/dp_sweep.sh N "caps..."
n=${1:-2}; shift
for cap in ${@:-0 140 132 124 116}; do
  SHG_DP_GEMM_CTAS=$cap timeout 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 2960$n bench.py --gpus $n --steps 20 --warmup 5 --no-extras --no-cpu-baseline > gpurun_out/dp_cap_${n}_$cap.json 2> gpurun_out/dp_cap_${n}_$cap.err
  python - <<PY
import json
try:
    l=json.loads([x for x in open('gpurun_out/dp_cap_${n}_$cap.json') if x.startswith('{')][0])
    sp=l.get('step_profile',{})
    print('N=$n cap=$cap ms_per_step', l['ms_per_step'], 'nccl active/exposed', sp.get('nccl_active_ms'), sp.get('nccl_exposed_ms'), 'span', sp.get('span_ms'), 'dp ok', l.get('dp_check',{}).get('ok'))
except Exception as e:
    print('N=$n cap=$cap failed', e)
PY
done
