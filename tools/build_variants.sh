#!/bin/bash
# Experimental builds of libwun.so for same-box A/B runs (bench.py / tests pick one with WUN_LIB=<path>).
set -e
cd "$(dirname "$0")/../wave-u-net_b200"
NVCC=${NVCC:-/usr/local/cuda/bin/nvcc}
SRC="csrc/plan.cpp csrc/crc32c.cpp csrc/kernels_simt.cu csrc/kernels_first.cu csrc/kernels_feed.cu csrc/kernels_umma.cu csrc/engine.cu"
F="-gencode arch=compute_100a,code=sm_100a -lineinfo -O3 -std=c++17 -diag-suppress 177 -Xcompiler -fPIC -shared"
for v in "$@"; do
  case $v in
    eb8) D="-DWUN_EPI_BATCH=8";;
    krb2) D="-DWUN_KRB=2";;
    krb4) D="-DWUN_KRB=4";;
    eb1) D="-DWUN_EPI_BATCH=1";;
    eb2) D="-DWUN_EPI_BATCH=2";;
    eb4) D="-DWUN_EPI_BATCH=4";;
    eb3) D="-DWUN_EPI_BATCH=3";;
    eb6) D="-DWUN_EPI_BATCH=6";;
    noslope) D="-DWUN_EXP_NOSLOPE";;
    *) echo "unknown variant $v"; exit 1;;
  esac
  $NVCC $F $D -o libwun_$v.so $SRC -lcuda &
done
wait
ls -la libwun_*.so
