cd /root/repo
mkdir -p gpurun_out
timeout 500 python -m pytest tests -m gpu -q -p no:cacheprovider > gpurun_out/pytest_all.log 2>&1; echo "pytest all rc=$? $(tail -1 gpurun_out/pytest_all.log)"; grep -E "^FAILED|^E  " gpurun_out/pytest_all.log | cut -c1-300 | head -20
timeout 300 python tools/kernel_bench.py --skip-conv --out gpurun_out/kbench_stft.json > gpurun_out/kbench_stft.log 2>&1; echo "stft bench rc=$?"; grep stft gpurun_out/kbench_stft.log | cut -c1-200
timeout 300 python tools/kprof.py > gpurun_out/kprof_b1.log 2>&1; grep -E "^==|fprop|wgrad" gpurun_out/kprof_b1.log | cut -c1-140
for s in 4 8; do echo "== wgrad splits $s"; SELDQ_WGRAD_SPLITS=$s timeout 200 python tools/kprof.py --layers tcn3,tcn1 2>&1 | grep -E "wgrad" | cut -c1-140; done
timeout 300 python tools/step_profile.py 2>&1 | grep -v "Warn\|_warn" > gpurun_out/step_profile_b1.log; echo "prof rc=$?"; head -32 gpurun_out/step_profile_b1.log | cut -c1-150
timeout 600 python bench.py --steps 10 --warmup 3 > gpurun_out/bench_r1i.log 2>&1; echo "bench rc=$?"; tail -1 gpurun_out/bench_r1i.log | cut -c1-1500
