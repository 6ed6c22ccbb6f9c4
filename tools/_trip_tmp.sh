cd /root/repo
mkdir -p gpurun_out
timeout 500 python -m pytest tests -m gpu -q -p no:cacheprovider > gpurun_out/pytest_all.log 2>&1; echo "pytest all rc=$? $(tail -1 gpurun_out/pytest_all.log)"; grep -E "^FAILED|^E  |Error" gpurun_out/pytest_all.log | cut -c1-300 | head -20
timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/bench_pdl1.log 2>&1; echo "bench PDL on rc=$?"; tail -1 gpurun_out/bench_pdl1.log | cut -c1-200; tail -1 gpurun_out/bench_pdl1.log | grep -o '"e2e": {[^}]*}'
SELDQ_PDL=0 timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/bench_pdl0.log 2>&1; echo "bench PDL off rc=$?"; tail -1 gpurun_out/bench_pdl0.log | cut -c1-200; tail -1 gpurun_out/bench_pdl0.log | grep -o '"e2e": {[^}]*}'
timeout 300 python tools/kprof.py --layers tcn3,tcn1 > gpurun_out/kprof_b1.log 2>&1; grep -E "^==|fprop|wgrad" gpurun_out/kprof_b1.log | cut -c1-140
