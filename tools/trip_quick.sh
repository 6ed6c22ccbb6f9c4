#!/usr/bin/env bash
# quick trip: parity + bench (+ optional extra env A/B lines): bash tools/trip_quick.sh <tag> ["ENV=1 ENV2=0" ...]
set -u
cd "$(dirname "$0")/.."
TAG="${1:-q}"; shift || true
O=gpurun_out; mkdir -p $O
timeout 400 python -m pytest tests -m gpu -q -p no:cacheprovider -x > $O/${TAG}_pytest.log 2>&1; echo "pytest rc=$? $(tail -1 $O/${TAG}_pytest.log)"; grep -E "^E  |FAILED|timed out" $O/${TAG}_pytest.log | head -10
timeout 300 python bench.py --steps 30 --warmup 5 > $O/${TAG}_bench.log 2>&1; echo "bench rc=$?"; tail -1 $O/${TAG}_bench.log > $O/${TAG}_bench.json; cut -c1-200 $O/${TAG}_bench.json
i=0
for envs in "$@"; do
  i=$((i+1))
  env $envs timeout 300 python bench.py --steps 30 --warmup 5 > $O/${TAG}_bench_ab$i.log 2>&1; echo "bench [$envs] rc=$?"; tail -1 $O/${TAG}_bench_ab$i.log | cut -c1-200
done
