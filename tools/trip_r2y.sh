#!/usr/bin/env bash
# quick trip: wgrad fold rewrite (parity, timeline, bench), aten op list of the step
set -u
cd "$(dirname "$0")/.."
O=gpurun_out; mkdir -p $O
timeout 300 python -m pytest tests -m gpu -q -p no:cacheprovider -x > $O/r2y_pytest.log 2>&1; echo "pytest rc=$? $(tail -1 $O/r2y_pytest.log)"
timeout 200 python tools/wgrad_trace.py > $O/r2y_wgrad_trace.txt 2>&1; echo "wgrad trace rc=$?"; grep -A1 "splits auto" $O/r2y_wgrad_trace.txt | cut -c1-400
timeout 300 python bench.py --steps 20 --warmup 5 > $O/r2y_bench.log 2>&1; echo "bench rc=$?"; tail -1 $O/r2y_bench.log | cut -c1-200
timeout 300 python tools/step_profile.py --top 70 --ops 2>&1 | grep -v "Warn\|_warn\|_ACCUMULATE" > $O/r2y_step_profile_ops.txt; echo "step profile rc=$?"
