#!/usr/bin/env bash
# Builds vsiquantization_b200/libvsiq.so for sm_100a (nvcc cross-compiles without a GPU).
# The CUDA runtime is linked statically so the library loads with no libcudart on the path.
set -euo pipefail
HERE="$(cd "$(dirname "${BASH_SOURCE[0]}")" && pwd)"
OUT="${1:-$HERE/../libvsiq.so}"
NVCC="${NVCC:-/usr/local/cuda/bin/nvcc}"
HOSTCXX="${VSIQ_HOSTCXX:-/usr/bin/g++}"
[ -x "$HOSTCXX" ] || HOSTCXX=g++
"$NVCC" -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -lineinfo \
    -ccbin "$HOSTCXX" -Xcompiler -fPIC,-O2,-Wall -shared -cudart static \
    ${VSIQ_NVCC_EXTRA:-} \
    -o "$OUT" "$HERE/abi.cu" "$HERE/fake_quant.cu" "$HERE/observer.cu" "$HERE/bn_fold.cu" "$HERE/channels_inner.cu" "$HERE/host_pipeline.cu"
echo "built $OUT"
