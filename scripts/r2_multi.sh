#!/bin/bash
# Multi-GPU evidence (run under gpurun --gpus N): the 2-GPU parity tests and the N-rank bench line.
# Usage: scripts/r2_multi.sh <tag> <N>
set -u
T=${1:-r2a}; N=${2:-2}; O=gpurun_out; mkdir -p $O
nvidia-smi -L | head -n 8
timeout 600 python -m pytest tests/test_gpu_multi.py -m gpu -q -rs > $O/${T}_pytest_multi_${N}gpu.log 2>&1; echo "pytest rc $?"; tail -n 6 $O/${T}_pytest_multi_${N}gpu.log
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29517 \
  bench.py --gpus $N > $O/${T}_bench_all_${N}gpu.json 2> $O/${T}_bench_all_${N}gpu.err; echo "bench rc $?"
cut -c1-400 $O/${T}_bench_all_${N}gpu.json; tail -n 5 $O/${T}_bench_all_${N}gpu.err
