#!/usr/bin/env bash
set -u
cd "$(dirname "$0")/.."
O=gpurun_out; T=r2g
mkdir -p $O
python tools/pair_bench.py > $O/${T}_pair_bench.txt 2>&1; head -24 $O/${T}_pair_bench.txt
timeout 900 python -m pytest tests -m gpu -q -x -p no:cacheprovider > $O/${T}_pytest.log 2>&1; echo "pytest rc=$? $(tail -1 $O/${T}_pytest.log)"
SELDQ_TCN_EPI=1 timeout 600 python -m pytest tests/test_gpu_parity.py -q -x -p no:cacheprovider -k "fused or full_size_training or bf16_matches" > $O/${T}_pytest_epi.log 2>&1; echo "pytest epi rc=$? $(tail -1 $O/${T}_pytest_epi.log)"
timeout 600 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-frontend > $O/${T}_bench.log 2>&1; echo "bench rc=$?"; tail -1 $O/${T}_bench.log | cut -c1-230
SELDQ_TCN_EPI=1 timeout 600 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-frontend > $O/${T}_bench_epi.log 2>&1; echo "bench epi rc=$?"; tail -1 $O/${T}_bench_epi.log | cut -c1-230
