#!/usr/bin/env bash
# Builds the tools-only probe library (tcgen05 / TMA bring-up and issue-rate probes) for sm_100a.
set -euo pipefail
HERE="$(cd "$(dirname "${BASH_SOURCE[0]}")" && pwd)"
NVCC="${NVCC:-/usr/local/cuda/bin/nvcc}"
"$NVCC" -gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -std=c++17 --expt-relaxed-constexpr --extended-lambda \
  -Xcompiler -fPIC -shared "$HERE/umma_probe.cu" -o "$HERE/libseldq_probe.so"
echo "built $HERE/libseldq_probe.so"
