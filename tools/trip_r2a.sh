#!/usr/bin/env bash
# round-2 first visit: state of the tree on the GPU + where batch > 1 loses its time
set -u
cd "$(dirname "$0")/.."
O=gpurun_out; T=r2a
mkdir -p $O
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > $O/${T}_host.txt 2>&1
lscpu | grep -E "Model name|^CPU\(s\)" >> $O/${T}_host.txt 2>&1
timeout 600 python -m pytest tests -m gpu -q -x -p no:cacheprovider > $O/${T}_pytest.log 2>&1; echo "pytest rc=$? $(tail -1 $O/${T}_pytest.log)"
timeout 600 python bench.py --steps 20 --warmup 5 > $O/${T}_bench.log 2>&1; echo "bench rc=$?"; tail -1 $O/${T}_bench.log | cut -c1-400
timeout 600 python bench.py --steps 20 --warmup 5 --model reference --no-cpu-baseline --no-frontend > $O/${T}_bench_refmodel.log 2>&1; echo "bench refmodel rc=$?"; tail -1 $O/${T}_bench_refmodel.log | cut -c1-300
timeout 600 python bench.py --steps 10 --warmup 3 --batch 4 --no-cpu-baseline --no-frontend > $O/${T}_bench_b4.log 2>&1; echo "bench b4 rc=$?"; tail -1 $O/${T}_bench_b4.log | cut -c1-300
timeout 300 python tools/step_profile.py --top 70 2>&1 | grep -v "Warn\|_warn\|_ACCUMULATE" > $O/${T}_step_profile_b1.txt; echo "profile b1 rc=$?"
timeout 300 python tools/step_profile.py --top 70 --batch 4 2>&1 | grep -v "Warn\|_warn\|_ACCUMULATE" > $O/${T}_step_profile_b4.txt; echo "profile b4 rc=$?"
timeout 300 python bench.py --steps 10 --warmup 3 --config DQSELD-TCN-S1-PHI_16chMagPhase --no-cpu-baseline --no-frontend > $O/${T}_bench_c4.log 2>&1; echo "bench c4 rc=$?"; tail -1 $O/${T}_bench_c4.log | cut -c1-300
ls -la $O/${T}_* | cut -c30-
