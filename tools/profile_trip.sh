#!/usr/bin/env bash
# One GPU-box visit that produces the evidence kept under profiles/: parity tests, the bench line of both arms,
# the per-kernel table of one step, the ncu launch list of one step and full ncu captures of the top kernels.
# Usage (from the repo root, on the GPU box): bash tools/profile_trip.sh <tag>
set -u
cd "$(dirname "$0")/.."
TAG="${1:-rX}"
O=gpurun_out
mkdir -p $O
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw,memory.total --format=csv > $O/${TAG}_host.txt 2>&1
lscpu | grep -E "Model name|^CPU\(s\)" >> $O/${TAG}_host.txt 2>&1
timeout 500 python -m pytest tests -m gpu -q -p no:cacheprovider > $O/${TAG}_pytest_gpu.log 2>&1; echo "pytest rc=$? $(tail -1 $O/${TAG}_pytest_gpu.log)"
timeout 900 python bench.py --steps 20 --warmup 5 > $O/${TAG}_bench.log 2>&1; echo "bench rc=$?"; tail -1 $O/${TAG}_bench.log > $O/${TAG}_bench.json; cut -c1-300 $O/${TAG}_bench.json
timeout 600 python bench.py --impl reference --steps 2 --warmup 1 > $O/${TAG}_bench_reference.log 2>&1; echo "ref rc=$?"; tail -1 $O/${TAG}_bench_reference.log > $O/${TAG}_bench_reference.json; cut -c1-300 $O/${TAG}_bench_reference.json
timeout 300 python tools/step_profile.py --top 60 2>&1 | grep -v "Warn\|_warn\|_ACCUMULATE" > $O/${TAG}_step_profile_b1.txt; echo "step profile rc=$?"
timeout 300 python tools/kernel_bench.py --skip-conv --out $O/${TAG}_stft.json > $O/${TAG}_stft.log 2>&1; echo "stft rc=$?"; grep stft $O/${TAG}_stft.log | cut -c1-160
timeout 600 ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv --log-file $O/${TAG}_ncu_launches.csv python tools/ncu_step.py --stft > $O/${TAG}_ncu_launches.log 2>&1; echo "ncu launches rc=$?"
# DRAM traffic of every convolution launch of one step (the roofline's `traffic`: tools/ncu_summary.py traffic)
timeout 600 ncu --profile-from-start off --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum --clock-control none \
  -k regex:qconv_cl_fprop_kernel --csv --log-file $O/${TAG}_ncu_traffic.csv python tools/ncu_step.py > $O/${TAG}_ncu_traffic.log 2>&1; echo "ncu traffic rc=$?"
# full captures of the top kernels, a few launches each (ncu replays every kernel ~40 times)
i=0
for spec in "qconv_cl_fprop_kernel:14" "qconv_cl_wgrad_kernel:6" "first_layer_bwd_kernel:1" "cnn_tail_fwd_vec_kernel:2" "cnn_tail_bwd_apply_vec_kernel:1" "stft_:2" "gate_fwd_kernel:1" "gate_bwd_apply_kernel:1"; do
  k="${spec%%:*}"; c="${spec##*:}"; i=$((i+1))
  SELDQ_PDL=0 SELDQ_SIDE_WGRAD=0 timeout 600 ncu --profile-from-start off --set full --clock-control none --import-source on \
    -k regex:"$k" -c $c -o $O/${TAG}_ncu_full_$i -f python tools/ncu_step.py --stft > $O/${TAG}_ncu_full_$i.log 2>&1; echo "ncu full $k rc=$?"
done
ls -la $O/${TAG}_* | cut -c30-
