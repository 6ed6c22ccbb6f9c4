#!/usr/bin/env bash
# One GPU-box visit: parity tests (isolated per group so a faulting kernel cannot poison the
# others' CUDA context), smoke, per-kernel timings.  Everything lands in gpurun_out/.
set -u
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw,memory.total --format=csv > gpurun_out/smi.txt 2>&1
run() { # name, -k expression
  timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -q -k "$2" -p no:cacheprovider > "gpurun_out/pytest_$1.log" 2>&1
  echo "pytest $1 rc=$? : $(tail -1 gpurun_out/pytest_$1.log)"; grep -E "^FAILED|AssertionError: \(" "gpurun_out/pytest_$1.log" | cut -c1-400 | head -12
}
for g in ${GROUPS_TO_RUN:-conv_bf16 model_fp32 model_bf16 conv_fp32 linear stft}; do
  case $g in
    conv_bf16) run conv_bf16 "conv_bf16 or rejects";;
    *) run $g "$g";;
  esac
done
echo "== smoke"; timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1; echo "smoke rc=$?"; tail -4 gpurun_out/smoke.log
if [ "${KBENCH:-1}" = 1 ]; then
  echo "== kbench"; timeout 900 python tools/kernel_bench.py --batch ${KBENCH_BATCH:-1 4} > gpurun_out/kbench.log 2>&1; echo "kbench rc=$?"; grep -vE "^Traceback|^  File|^    " gpurun_out/kbench.log | cut -c1-260 | tail -50
fi
