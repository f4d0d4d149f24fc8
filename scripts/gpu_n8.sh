#!/bin/bash
# N-GPU confirmation: NCCL/peer exchange parity test (world 4 when >= 4 GPUs) + the driver's default line at N ranks
TAG=${1:-R2s}; N=${2:-8}; OUT=gpurun_out/$TAG; mkdir -p $OUT
nvidia-smi topo -m > $OUT/topo.txt 2>&1; nproc > $OUT/nproc.txt; lscpu | grep -i "numa\|socket\|model name" >> $OUT/nproc.txt
timeout 600 python -m pytest tests/test_dist_gpu.py -x -q > $OUT/pytest_dist.txt 2>&1; tail -3 $OUT/pytest_dist.txt
bash scripts/gpu_default_n.sh $TAG $N
