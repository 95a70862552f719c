#!/bin/bash
# ncu --set full of the hot kernels in their end-of-round state (run under gpurun).
set -x
O=gpurun_out
python tools/prof_case.py gemm 12576 3072 768 gelu > /dev/null 2>&1 || exit 1
ncu --set full --clock-control none --import-source on -k regex:gemm -s 2 -c 1 -o $O/prof_gemm_ffn1_r01c \
    python tools/prof_case.py gemm 12576 3072 768 gelu > $O/ncu_g1.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:gemm -s 2 -c 1 -o $O/prof_gemm_plain_r01c \
    python tools/prof_case.py gemm 12576 3072 768 > $O/ncu_g2.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:attn_fwd -s 2 -c 1 -o $O/prof_attn_fwd_r01c \
    python tools/prof_case.py attn > $O/ncu_a1.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:attn_bwd -s 4 -c 2 -o $O/prof_attn_bwd_r01c \
    python tools/prof_case.py attn_bwd > $O/ncu_a2.log 2>&1
ls -la $O/*r01c.ncu-rep
