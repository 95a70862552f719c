#!/bin/bash
# ncu --set full of the Conv3d implicit-GEMM kernels at bs32 (run under gpurun), plus a fresh graph-replay timeline.
set -x
O=gpurun_out
python tools/prof_case.py conv_fwd 1 > /dev/null 2>&1 || exit 1
ncu --set full --clock-control none --import-source on -k regex:conv_gemm -s 2 -c 1 -o $O/prof_conv_fwd_r01 \
    python tools/prof_case.py conv_fwd 1 > $O/ncu_c1.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:conv_wgrad -s 2 -c 1 -o $O/prof_conv_wgrad_r01 \
    python tools/prof_case.py conv_wgrad 1 > $O/ncu_c2.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:conv_gemm -s 2 -c 1 -o $O/prof_conv_dgrad_r01 \
    python tools/prof_case.py conv_dgrad 2 > $O/ncu_c3.log 2>&1
python tools/timeline.py > $O/timeline3.txt 2>&1
ls -la $O/*.ncu-rep
