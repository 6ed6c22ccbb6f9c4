#!/usr/bin/env bash
# Final-state evidence in one short GPU-box visit: parity tests, smoke, the bench line of both arms, the per-kernel
# table of one step and the ncu launch list of one step (shares): bash tools/trip_final.sh <tag>
set -u
cd "$(dirname "$0")/.."
TAG="${1:-fin}"
O=gpurun_out; mkdir -p $O
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw,memory.total --format=csv > $O/${TAG}_host.txt 2>&1
timeout 400 python -m pytest tests -m gpu -q -p no:cacheprovider > $O/${TAG}_pytest_gpu.log 2>&1; echo "pytest rc=$? $(tail -1 $O/${TAG}_pytest_gpu.log)"
timeout 200 python -c "import __graft_entry__ as g; g.smoke()" > $O/${TAG}_smoke.log 2>&1; echo "smoke rc=$?"; tail -3 $O/${TAG}_smoke.log
timeout 600 python bench.py --steps 20 --warmup 5 > $O/${TAG}_bench.log 2>&1; echo "bench rc=$?"; tail -1 $O/${TAG}_bench.log > $O/${TAG}_bench.json; cut -c1-240 $O/${TAG}_bench.json
timeout 300 python bench.py --impl reference --steps 2 --warmup 1 > $O/${TAG}_bench_reference.log 2>&1; echo "ref rc=$?"; tail -1 $O/${TAG}_bench_reference.log > $O/${TAG}_bench_reference.json; cut -c1-200 $O/${TAG}_bench_reference.json
timeout 300 python tools/step_profile.py --top 60 2>&1 | grep -v "Warn\|_warn\|_ACCUMULATE" > $O/${TAG}_step_profile_b1.txt; echo "step profile rc=$?"
timeout 400 ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv --log-file $O/${TAG}_ncu_launches.csv python tools/ncu_step.py --stft > $O/${TAG}_ncu_launches.log 2>&1; echo "ncu launches rc=$?"
