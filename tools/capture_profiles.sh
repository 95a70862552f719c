#!/bin/bash
# Round-1 evidence: launch list of one eager bs32 train step + ncu --set full of the hot kernels (run under gpurun).
set -x
O=gpurun_out
python bench.py --steps 1 --warmup 3 --eager --no-cpu-baseline > $O/plain_eager.log 2>&1 || exit 1
ncu --metrics gpu__time_duration.sum --clock-control none -s 2600 -c 1100 --csv --log-file $O/launches_r01_v3.csv \
    python bench.py --steps 1 --warmup 3 --eager --no-cpu-baseline > $O/ncu_launches.log 2>&1
python tools/prof_case.py gemm 12576 3072 768 gelu > /dev/null 2>&1 || exit 1
ncu --set full --clock-control none --import-source on -k regex:gemm -s 2 -c 1 -o $O/prof_gemm_ffn1_r01 \
    python tools/prof_case.py gemm 12576 3072 768 gelu > $O/ncu_g1.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:gemm -s 2 -c 1 -o $O/prof_gemm_plain_r01 \
    python tools/prof_case.py gemm 12576 3072 768 > $O/ncu_g2.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:attn_fwd -s 2 -c 1 -o $O/prof_attn_fwd_r01b \
    python tools/prof_case.py attn > $O/ncu_a1.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:attn_bwd -s 4 -c 2 -o $O/prof_attn_bwd_r01b \
    python tools/prof_case.py attn_bwd > $O/ncu_a2.log 2>&1
ls -la $O/*.ncu-rep
