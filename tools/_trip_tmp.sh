cd /root/repo
mkdir -p gpurun_out
timeout 500 python -m pytest tests -m gpu -q -p no:cacheprovider > gpurun_out/pytest_all.log 2>&1; echo "pytest all rc=$? $(tail -1 gpurun_out/pytest_all.log)"; grep -E "^FAILED|^E  |Error" gpurun_out/pytest_all.log | cut -c1-300 | head -20
timeout 600 python bench.py --steps 20 --warmup 3 --no-cpu-baseline > gpurun_out/bench_r1q.log 2>&1; echo "bench rc=$? $(tail -1 gpurun_out/bench_r1q.log | grep -o '"ms_per_step": [0-9.]*' | head -1)"; tail -1 gpurun_out/bench_r1q.log | grep -o '"roofline": {[^}]*' | cut -c1-160
SELDQ_PDL=0 timeout 300 python tools/kprof.py --layers cnn0,cnn1,cnn2,tcn3,tcn1 > gpurun_out/kprof_b1.log 2>&1; grep -E "^==|fprop|wgrad" gpurun_out/kprof_b1.log | cut -c1-140
