#!/bin/bash
# builds the standalone tcgen05 probes (run them under gpurun)
set -e
cd "$(dirname "$0")/.."
F="-gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -lineinfo -diag-suppress 177"
nvcc $F -o tools/umma_probe tools/umma_probe.cu
nvcc $F -DWUN_UMMA_TIMING -o tools/umma_probe_timing tools/umma_probe.cu
nvcc $F -o tools/umma_layout_bench tools/umma_layout_bench.cu
nvcc $F -o tools/presplit_probe tools/presplit_probe.cu
nvcc $F -o tools/first_layer_probe tools/first_layer_probe.cu

nvcc $F -o tools/cluster_probe tools/cluster_probe.cu
nvcc $F -o tools/umma_rate_bench tools/umma_rate_bench.cu
echo probes built
