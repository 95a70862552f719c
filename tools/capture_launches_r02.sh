#!/bin/bash
# Launch list of the bench command (B200_PROFILING.md's "--metrics gpu__time_duration.sum --clock-control none" pass):
# every kernel of one replayed bs32 train step with its device time (cold-cache, serialised: compare SHARES).
O=gpurun_out
python bench.py --steps 2 --warmup 3 --no-extras --no-cpu-baseline > $O/plain_launches.log 2>&1 && \
ncu --metrics gpu__time_duration.sum --clock-control none --graph-profiling node -c 6000 --csv \
    --log-file $O/launches_r02.csv python bench.py --steps 2 --warmup 3 --no-extras --no-cpu-baseline > $O/ncu_launches.log 2>&1
tail -2 $O/ncu_launches.log; wc -l $O/launches_r02.csv
