cd /root/repo
mkdir -p gpurun_out
timeout 500 python -m pytest tests -m gpu -q -p no:cacheprovider > gpurun_out/pytest_all.log 2>&1; echo "pytest all rc=$? $(tail -1 gpurun_out/pytest_all.log)"; grep -E "^FAILED|^E  |Error" gpurun_out/pytest_all.log | cut -c1-300 | head -20
timeout 600 python bench.py --steps 20 --warmup 3 --no-cpu-baseline > gpurun_out/bench_r1p.log 2>&1; echo "bench rc=$? $(tail -1 gpurun_out/bench_r1p.log | grep -o '"ms_per_step": [0-9.]*' | head -1)"
SELDQ_PDL=0 SELDQ_SIDE_WGRAD=0 timeout 300 python tools/step_profile.py --top 50 2>&1 | grep -v "Warn\|_warn\|_ACCUM" > gpurun_out/step_profile_b1.log; echo "prof rc=$?"; grep -E "total|tcn::|epi::|first::" gpurun_out/step_profile_b1.log | cut -c1-150
