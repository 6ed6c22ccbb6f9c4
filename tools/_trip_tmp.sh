cd /root/repo
mkdir -p gpurun_out
timeout 500 python -m pytest tests -m gpu -q -p no:cacheprovider > gpurun_out/pytest_all.log 2>&1; echo "pytest all rc=$? $(tail -1 gpurun_out/pytest_all.log)"; grep -E "^FAILED|^E  |Error" gpurun_out/pytest_all.log | cut -c1-300 | head -20
timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/bench_r1m.log 2>&1; echo "bench rc=$?"; tail -1 gpurun_out/bench_r1m.log | cut -c1-200; tail -1 gpurun_out/bench_r1m.log | grep -o '"e2e": {[^}]*}'
