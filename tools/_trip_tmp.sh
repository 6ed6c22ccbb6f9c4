#!/usr/bin/env bash
set -u
cd "$(dirname "$0")/.."
O=gpurun_out
for cfg in "B:SELDQ_QUAD_FUSE=2 SELDQ_ACC_DOUBLE=1" "C:SELDQ_QUAD_FUSE=2"; do
  tag="${cfg%%:*}"; envs="${cfg#*:}"
  echo "== $tag [$envs]"
  env $envs timeout 300 python -m pytest tests -m gpu -q -p no:cacheprovider -k "conv_bf16 or model_bf16 or full_size_conv" 2>&1 | tail -1
  env $envs timeout 300 python bench.py --steps 30 --warmup 5 --no-cpu-baseline > $O/r1o_bench_$tag.log 2>&1; echo "bench rc=$?"
  tail -1 $O/r1o_bench_$tag.log | python -c "import json,sys; d=json.loads(sys.stdin.read()); print(d['value'], d['ms_per_step'], d['e2e']['value'], d['roofline']['frac'], d['roofline']['other_kernels_ms_per_step']['qconv_cl_fprop_kernel'])"
  env $envs timeout 300 python tools/step_profile.py --top 8 2>&1 | grep -A1 "fprop_kernel" | cut -c1-120
done
