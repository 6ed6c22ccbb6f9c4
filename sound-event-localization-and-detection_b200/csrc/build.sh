#!/usr/bin/env bash
# Builds libseldq.so for sm_100a in-tree (next to this script's parent package).
# nvcc cross-compiles without a GPU; the .so travels to the GPU box with the repo snapshot.
set -euo pipefail
HERE="$(cd "$(dirname "${BASH_SOURCE[0]}")" && pwd)"
OUT="${1:-$HERE/../libseldq.so}"
NVCC="${NVCC:-/usr/local/cuda/bin/nvcc}"
FLAGS=(-gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -std=c++17 --expt-relaxed-constexpr
       --extended-lambda -Xcompiler -fPIC)
SRCS=(seldq_api simt_kernels stft conv_cl wgrad_cl epilogue tcn_glue conv_umma wgrad_umma wgrad_first attention tail eval rotation)
OBJDIR="$HERE/build"
mkdir -p "$OBJDIR"
pids=()
for s in "${SRCS[@]}"; do
  "$NVCC" "${FLAGS[@]}" -c "$HERE/$s.cu" -o "$OBJDIR/$s.o" &
  pids+=($!)
done
for p in "${pids[@]}"; do wait "$p"; done
OBJS=()
for s in "${SRCS[@]}"; do OBJS+=("$OBJDIR/$s.o"); done
"$NVCC" -shared -o "$OUT" "${OBJS[@]}"
echo "built $OUT"
