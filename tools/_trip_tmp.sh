cd /root/repo
GROUPS_TO_RUN="model_fp32 model_bf16" KBENCH=0 bash tools/gpu_trip.sh
timeout 600 python bench.py --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/bench_first.log 2>&1; echo "bench rc=$?"; tail -5 gpurun_out/bench_first.log | cut -c1-1500
