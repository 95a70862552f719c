#!/bin/bash
# ncu captures of the hot kernels in their end-of-round-2 state (run under gpurun; one ncu "use" per call).
# Each command first exits 0 without ncu, then runs under ncu (profiles/README.md lists what each file shows).
O=gpurun_out
cap() {   # name, kernel regex, skip, count, command...
  local name=$1 rx=$2 skip=$3 cnt=$4; shift 4
  "$@" > $O/plain_$name.log 2>&1 && \
  ncu --set full --clock-control none --import-source on -k regex:$rx -s $skip -c $cnt -f -o $O/prof_${name}_r02 "$@" > $O/ncu_$name.log 2>&1
  tail -1 $O/ncu_$name.log
}
cap gemm_plain gemm 2 1 python tools/prof_case.py gemm 12576 3072 768
cap gemm_ffn1 gemm 2 1 python tools/prof_case.py gemm 12576 3072 768 gelu
cap gemm_mid gemm 2 1 python tools/prof_case.py gemm 5632 768 768
cap attn_fwd attn_fwd 2 1 python tools/prof_case.py attn
cap attn_bwd attn_bwd 4 2 python tools/prof_case.py attn_bwd
cap ln layernorm 4 2 python tools/prof_case.py ln
cap colsum colsum 2 1 python tools/prof_case.py colsum
cap match "match_from_logits|set_loss" 4 2 python tools/prof_case.py match
cap adam "bert_adam|grad_sumsq" 4 2 python tools/prof_case.py adam
ls -la $O/*_r02.ncu-rep
