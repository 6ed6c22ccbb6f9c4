cd /root/repo
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q -p no:cacheprovider -x > gpurun_out/pytest_all.log 2>&1; echo "pytest all rc=$? $(tail -1 gpurun_out/pytest_all.log)"; grep -E "^FAILED|^E  " gpurun_out/pytest_all.log | cut -c1-300 | head -20
timeout 900 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/bench_r1c.log 2>&1; echo "bench rc=$?"; tail -1 gpurun_out/bench_r1c.log | cut -c1-600
