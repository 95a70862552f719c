#!/bin/bash
# End-of-round evidence: bench line, launch list of one eager step (ncu), CUPTI timeline of the graph replay.
set -x
O=gpurun_out
python bench.py > $O/bench_final.log 2> $O/bench_final.err || exit 1
python bench.py --steps 1 --warmup 3 --eager --no-cpu-baseline --no-frontend > $O/plain_eager.log 2>&1 || exit 1
ncu --metrics gpu__time_duration.sum --clock-control none -s 2300 -c 1100 --csv --log-file $O/launches_final.csv \
    python bench.py --steps 1 --warmup 3 --eager --no-cpu-baseline --no-frontend > $O/ncu_launches.log 2>&1
python tools/timeline.py > $O/timeline_final.txt 2>&1
ncu --set full --clock-control none --import-source on -k regex:attn_bwd_dq -s 2 -c 1 -o $O/prof_attn_bwd_merged_r01 \
    python tools/prof_case.py attn_bwd 128 393 > $O/ncu_m1.log 2>&1
