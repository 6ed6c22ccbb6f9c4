#!/usr/bin/env bash
# two-GPU visit: NCCL correctness of the two-part exchange + the 2-GPU bench line (with and without the overlap)
set -u
cd "$(dirname "$0")/.."
O=gpurun_out; T=${1:-r2n2}
mkdir -p $O
timeout 600 python -m pytest tests/test_gpu_nccl.py -q -x -s -p no:cacheprovider > $O/${T}_pytest_nccl.log 2>&1; echo "nccl test rc=$? $(tail -1 $O/${T}_pytest_nccl.log)"; grep "^E  \|^[01] {" $O/${T}_pytest_nccl.log | cut -c1-400 | head
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 20 --warmup 5 --no-frontend > $O/${T}_bench_n2.log 2>&1; echo "bench n2 rc=$?"; grep '^{' $O/${T}_bench_n2.log | tail -1 | cut -c1-260
SELDQ_AR_OVERLAP=0 timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 2 --steps 20 --warmup 5 --no-frontend > $O/${T}_bench_n2_noverlap.log 2>&1; echo "bench n2 no-overlap rc=$?"; grep '^{' $O/${T}_bench_n2_noverlap.log | tail -1 | cut -c1-260
timeout 600 python bench.py --steps 20 --warmup 5 --no-frontend --no-cpu-baseline > $O/${T}_bench_n1.log 2>&1; grep '^{' $O/${T}_bench_n1.log | tail -1 | cut -c1-260
