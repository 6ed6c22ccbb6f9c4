#!/usr/bin/env bash
set -u
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
export CUDA_LAUNCH_BLOCKING=1
for c in interior right_oob chan_oob neg_aligned neg_aligned_big chan_neg pos_unaligned neg_unaligned; do
  T1_VARIANT=r4_sw128_param T1_CASE=$c timeout 120 python tools/umma_probe.py t1 > gpurun_out/probe_t1_$c.log 2>&1
  echo "t1 $c rc=$? : $(grep -E 'PASS|fail|Error|timed out' gpurun_out/probe_t1_$c.log | tr '\n' ';' | cut -c1-300)"
done
for v in as_designed a_swapped b_swapped both_swapped; do
  U1_VARIANT=$v timeout 120 python tools/umma_probe.py u1 > gpurun_out/probe_u1_$v.log 2>&1
  echo "u1 $v rc=$? : $(grep -E 'PASS|fail|Error|timed out' gpurun_out/probe_u1_$v.log | tr '\n' ';' | cut -c1-400)"
done
