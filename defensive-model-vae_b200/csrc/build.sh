#!/usr/bin/env bash
# Build libdmvae.so in-tree for sm_100a (B200).  nvcc cross-compiles without a GPU.
set -euo pipefail
cd "$(dirname "$0")"
NVCC=${NVCC:-/usr/local/cuda/bin/nvcc}
FLAGS=(-gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -std=c++17 -Xcompiler -fPIC -Xptxas -v)
SRCS=(dmvae_api.cu dmvae_pack.cu dmvae_decode.cu dmvae_decode_tc.cu)
for extra in dmvae_train.cu dmvae_train_tc.cu dmvae_adam.cu dmvae_loss.cu dmvae_prof.cu dmvae_probe_tc.cu dmvae_metrics.cu dmvae_mpc.cu dmvae_dense.cu; do [ -f "$extra" ] && SRCS+=("$extra"); done
mkdir -p build
objs=()
for s in "${SRCS[@]}"; do
  o="build/${s%.cu}.o"
  # stale when the source or any header of this directory / include/ is newer (every .cu sees most of them)
  stale=0
  [ -f "$o" ] || stale=1
  for dep in "$s" build.sh *.cuh *.h ../../include/*.h; do [ "$dep" -nt "$o" ] && stale=1; done
  if [ "$stale" = 1 ]; then
    echo "nvcc $s"
    "$NVCC" "${FLAGS[@]}" -c "$s" -o "$o" 2> "build/${s%.cu}.ptxas.log" || { cat "build/${s%.cu}.ptxas.log"; exit 1; }
  fi
  objs+=("$o")
done
"$NVCC" -gencode arch=compute_100a,code=sm_100a -shared -o libdmvae.so "${objs[@]}"
echo "built $(pwd)/libdmvae.so"
