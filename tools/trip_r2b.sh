#!/usr/bin/env bash
# sibling launches + fused epilogues: parity, then A/B of the step time
set -u
cd "$(dirname "$0")/.."
O=gpurun_out; T=r2c
mkdir -p $O
timeout 900 python -m pytest tests -m gpu -q -x -p no:cacheprovider > $O/${T}_pytest.log 2>&1; echo "pytest rc=$? $(tail -1 $O/${T}_pytest.log)"

timeout 600 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-frontend > $O/${T}_bench.log 2>&1; echo "bench rc=$?"; tail -1 $O/${T}_bench.log | cut -c1-330
SELDQ_TCN_PAIR=0 timeout 600 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-frontend > $O/${T}_bench_nopair.log 2>&1; echo "bench nopair rc=$?"; tail -1 $O/${T}_bench_nopair.log | cut -c1-330
timeout 300 python tools/step_profile.py --top 40 2>&1 | grep -v "Warn\|_warn\|_ACCUMULATE" > $O/${T}_step_profile_b1.txt; echo "profile b1 rc=$?"; head -30 $O/${T}_step_profile_b1.txt | cut -c1-150
