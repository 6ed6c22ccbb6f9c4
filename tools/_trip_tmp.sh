cd /root/repo
mkdir -p gpurun_out
timeout 500 python -m pytest tests -m gpu -q -p no:cacheprovider > gpurun_out/pytest_all.log 2>&1; echo "pytest all rc=$? $(tail -1 gpurun_out/pytest_all.log)"; grep -E "^FAILED|^E  |Error" gpurun_out/pytest_all.log | cut -c1-300 | head -20
timeout 600 python bench.py --steps 20 --warmup 3 --no-cpu-baseline > gpurun_out/bench_r1n.log 2>&1; echo "bench rc=$?"; tail -1 gpurun_out/bench_r1n.log | cut -c1-200; tail -1 gpurun_out/bench_r1n.log | grep -o '"e2e": {[^}]*}'
timeout 300 python tools/step_profile.py 2>&1 | grep -v "Warn\|_warn\|_ACCUM" > gpurun_out/step_profile_b1.log; echo "prof rc=$?"; head -12 gpurun_out/step_profile_b1.log | cut -c1-150; grep -E "add<float>|elementwise_kernel<128" gpurun_out/step_profile_b1.log | cut -c1-120
for c in QSELD-TCN-S1-PHI_parallel_8ch DQSELD-TCN-S1-PHI_16chMagPhase; do timeout 600 python bench.py --config $c --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/bench_$c.log 2>&1; echo "$c rc=$? $(tail -1 gpurun_out/bench_$c.log | cut -c1-160)"; done
