#!/usr/bin/env bash
set -u
cd "$(dirname "$0")/.."
O=gpurun_out; T=r2i
mkdir -p $O
timeout 900 python -m pytest tests -m gpu -q -p no:cacheprovider > $O/${T}_pytest.log 2>&1; echo "pytest rc=$? $(tail -1 $O/${T}_pytest.log)"; grep "^FAILED\|^E  " $O/${T}_pytest.log | cut -c1-300 | head -20
timeout 600 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-frontend > $O/${T}_bench.log 2>&1; echo "bench rc=$?"; tail -1 $O/${T}_bench.log | cut -c1-230
SELDQ_ATTN=fp32 timeout 600 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-frontend > $O/${T}_bench_noattn.log 2>&1; echo "bench torch-attn rc=$?"; tail -1 $O/${T}_bench_noattn.log | cut -c1-230
timeout 300 python tools/step_profile.py --top 45 2>&1 | grep -v "Warn\|_warn\|_ACCUMULATE" > $O/${T}_step_profile_b1.txt; head -48 $O/${T}_step_profile_b1.txt | cut -c1-150
