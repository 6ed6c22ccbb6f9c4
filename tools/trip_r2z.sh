#!/usr/bin/env bash
# quick trip: single-launch glue steps (parity, bench A/B)
set -u
cd "$(dirname "$0")/.."
O=gpurun_out; mkdir -p $O
timeout 400 python -m pytest tests -m gpu -q -p no:cacheprovider -x > $O/r2z_pytest.log 2>&1; echo "pytest rc=$? $(tail -1 $O/r2z_pytest.log)"; grep -E "^E  |FAILED|timed out" $O/r2z_pytest.log | head -10
timeout 300 python bench.py --steps 20 --warmup 5 > $O/r2z_bench.log 2>&1; echo "bench rc=$?"; tail -1 $O/r2z_bench.log | cut -c1-200
SELDQ_TCN_FUSED_GLUE=0 timeout 300 python bench.py --steps 20 --warmup 5 > $O/r2z_bench_twolaunch.log 2>&1; echo "bench (two-launch glue) rc=$?"; tail -1 $O/r2z_bench_twolaunch.log | cut -c1-200
SELDQ_SIDE_WGRAD=0 timeout 300 python bench.py --steps 20 --warmup 5 > $O/r2z_bench_noside.log 2>&1; echo "bench (no side stream) rc=$?"; tail -1 $O/r2z_bench_noside.log | cut -c1-200
timeout 300 python tools/step_profile.py --top 40 2>&1 | grep -v "Warn\|_warn\|_ACCUMULATE" > $O/r2z_step_profile_b1.txt; echo "step profile rc=$?"; head -30 $O/r2z_step_profile_b1.txt | cut -c1-150
