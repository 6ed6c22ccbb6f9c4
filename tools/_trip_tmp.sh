cd /root/repo
mkdir -p gpurun_out
timeout 500 python -m pytest tests -m gpu -q -p no:cacheprovider > gpurun_out/pytest_all.log 2>&1; echo "pytest all rc=$? $(tail -1 gpurun_out/pytest_all.log)"; grep -E "^FAILED|^E  |Error" gpurun_out/pytest_all.log | cut -c1-300 | head -20
for i in 1 2; do
timeout 600 python bench.py --steps 20 --warmup 3 --no-cpu-baseline > gpurun_out/bench_w1.log 2>&1; echo "side wgrad on:  rc=$? $(tail -1 gpurun_out/bench_w1.log | grep -o '"ms_per_step": [0-9.]*' | head -1)"
SELDQ_SIDE_WGRAD=0 timeout 600 python bench.py --steps 20 --warmup 3 --no-cpu-baseline > gpurun_out/bench_w0.log 2>&1; echo "side wgrad off: rc=$? $(tail -1 gpurun_out/bench_w0.log | grep -o '"ms_per_step": [0-9.]*' | head -1)"
done
tail -3 gpurun_out/bench_w1.log | cut -c1-300
