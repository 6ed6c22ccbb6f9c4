#!/usr/bin/env bash
set -u
cd "$(dirname "$0")/.."
O=gpurun_out
timeout 400 python -m pytest tests -m gpu -q -p no:cacheprovider > $O/r1m_pytest_gpu.log 2>&1; echo "pytest rc=$? $(tail -1 $O/r1m_pytest_gpu.log)"
timeout 300 python bench.py --steps 20 --warmup 5 --no-cpu-baseline > $O/r1m_bench.log 2>&1; echo "bench rc=$?"
tail -1 $O/r1m_bench.log | python -c "import json,sys; d=json.loads(sys.stdin.read()); print(d['value'], d['ms_per_step'], d['e2e']['value'], d['roofline']['frac'], d['roofline']['other_kernels_ms_per_step'])"
timeout 300 python tools/step_profile.py --top 70 --ops 2>&1 | grep -v "Warn\|_warn\|_ACCUMULATE" > $O/r1m_step_profile_b1.txt; echo "step profile rc=$?"
