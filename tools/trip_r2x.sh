#!/usr/bin/env bash
# quick trip: Adam parity, ring fix under ncu --set full, wgrad / fprop CTA timelines
set -u
cd "$(dirname "$0")/.."
O=gpurun_out; mkdir -p $O
timeout 300 python -m pytest tests -m gpu -q -p no:cacheprovider -x > $O/r2x_pytest.log 2>&1; echo "pytest rc=$? $(tail -1 $O/r2x_pytest.log)"
timeout 200 python tools/wgrad_trace.py > $O/r2x_wgrad_trace.txt 2>&1; echo "wgrad trace rc=$?"; cat $O/r2x_wgrad_trace.txt | cut -c1-400
timeout 200 python tools/fprop_trace.py > $O/r2x_fprop_trace.txt 2>&1; echo "fprop trace rc=$?"; cat $O/r2x_fprop_trace.txt | cut -c1-300
timeout 300 python bench.py --steps 20 --warmup 5 > $O/r2x_bench.log 2>&1; echo "bench rc=$?"; tail -1 $O/r2x_bench.log | cut -c1-200
SELDQ_PDL=0 SELDQ_SIDE_WGRAD=0 timeout 600 ncu --profile-from-start off --set full --clock-control none --import-source on \
    -k regex:qconv_cl_fprop_kernel -c 14 -o $O/r2x_ncu_full_1 -f python tools/ncu_step.py --stft > $O/r2x_ncu_full_1.log 2>&1; echo "ncu full fprop rc=$?"
