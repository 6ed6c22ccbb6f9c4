cd /root/repo
mkdir -p gpurun_out
timeout 500 python -m pytest tests -m gpu -q -p no:cacheprovider > gpurun_out/pytest_all.log 2>&1; echo "pytest all rc=$? $(tail -1 gpurun_out/pytest_all.log)"; grep -E "^FAILED|^E  |Error" gpurun_out/pytest_all.log | cut -c1-300 | head -20
timeout 300 python tools/step_profile.py 2>&1 | grep -v "Warn\|_warn" > gpurun_out/step_profile_b1.log; echo "prof rc=$?"; head -16 gpurun_out/step_profile_b1.log | cut -c1-150
timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/bench_r1l.log 2>&1; echo "bench rc=$?"; tail -1 gpurun_out/bench_r1l.log | cut -c1-600
