#!/usr/bin/env bash
# bench lines of every BASELINE config (C2..C5 + the real-valued baseline), the fp32 parity mode, batch 4 and the
# reference's unmodified model.py on the drop-in: bash tools/trip_configs.sh <tag>
set -u
cd "$(dirname "$0")/.."
TAG="${1:-cfg}"
O=gpurun_out; mkdir -p $O
run() { # name, args...
  local name=$1; shift
  timeout 600 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-frontend "$@" > $O/${TAG}_bench_$name.log 2>&1
  echo "$name rc=$? $(tail -1 $O/${TAG}_bench_$name.log | cut -c1-170)"
  tail -1 $O/${TAG}_bench_$name.log > $O/${TAG}_bench_$name.json
}
run c2_b4 --batch 4
run c2_refmodel --model reference
run c2_fp32 --precision fp32
run c3_q_parallel --config QSELD-TCN-S1-PHI_parallel_8ch
run c4_16ch --config DQSELD-TCN-S1-PHI_16chMagPhase
run c5_two_branch --config DQSELD-TCN-S1-PHI_micAMagPhaseParallelmicBMagPhase
run real_8ch --config SELD-TCN-S1-PHI_8ch
