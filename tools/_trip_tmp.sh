cd /root/repo
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw,memory.total --format=csv > gpurun_out/smi.txt 2>&1
nproc > gpurun_out/host.txt; lscpu | grep -E "Model name|^CPU\(s\)|Thread|Socket" >> gpurun_out/host.txt
timeout 900 python -m pytest tests -m gpu -x -q -p no:cacheprovider > gpurun_out/pytest_all.log 2>&1; echo "pytest rc=$? $(tail -1 gpurun_out/pytest_all.log)"
timeout 600 python tools/step_profile.py > gpurun_out/step_profile_b1.log 2>&1; echo "prof rc=$?"; head -30 gpurun_out/step_profile_b1.log
timeout 900 python bench.py --steps 10 --warmup 3 > gpurun_out/bench_r1a.log 2>&1; echo "bench rc=$?"; tail -1 gpurun_out/bench_r1a.log | cut -c1-2500
timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_ref_r1a.log 2>&1; echo "ref rc=$?"; tail -1 gpurun_out/bench_ref_r1a.log | cut -c1-1200
