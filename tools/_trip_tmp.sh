cd /root/repo
mkdir -p gpurun_out
timeout 600 python tools/step_profile.py 2>&1 | grep -v "Warn\|_warn" > gpurun_out/step_profile_b1.log; echo "prof rc=$?"; head -32 gpurun_out/step_profile_b1.log | cut -c1-150
timeout 900 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/bench_r1b.log 2>&1; echo "bench rc=$?"; tail -1 gpurun_out/bench_r1b.log | cut -c1-1800
