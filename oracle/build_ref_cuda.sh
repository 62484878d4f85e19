#!/usr/bin/env bash
# TEST / MEASUREMENT INFRASTRUCTURE -- never on the product path.
#
# Builds the reference's own CUDA backend (the kernel to beat, SURVEY.md
# Appendix B) for sm_100 from the UNMODIFIED sources where they lie under
# /root/reference, with plain nvcc (the reference's cmake build is not run).
# Output: oracle/_ref/libtfhe_cuda_backend_ref.so (git-ignored, travels to the
# GPU box with gpurun).  It exports the same extern "C" symbols as our
# library, so the B200_LIB_PATH loader of tfhe-rs_b200/_lib.py can drive it
# through the same harness:  same-box A/B for throughput, latency, keyswitch
# and multi-bit, plus cross-implementation parity on the oracle's keys.
#
# Only the hot-path translation units are compiled (PBS classic + multi-bit,
# key conversion, FFT twiddle tables, keyswitch, ciphertext helpers, multi-GPU
# helper, device wrappers); flags follow the reference's CMakeLists.txt
# (cuda/CMakeLists.txt:104-107: -O3 -std=c++17 --no-exceptions
# --expt-relaxed-constexpr -rdc=true --use_fast_math, CUDA_ARCH=<cc>0).
set -euo pipefail
REF=${REF_ROOT:-/root/reference}
CUDA_DIR=$REF/backends/tfhe-cuda-backend/cuda
COMMON=$REF/backends/tfhe-cuda-common/cuda
HERE=$(cd "$(dirname "$0")" && pwd)
OUT=$HERE/_ref
OBJ=$OUT/obj
JOBS=${JOBS:-$(nproc)}
if [ ! -d "$CUDA_DIR" ]; then
  echo "build_ref_cuda: $CUDA_DIR absent (GPU box) -- using the prebuilt .so" >&2
  exit 0
fi
mkdir -p "$OBJ"
SRCS=(
  "$COMMON/src/device.cu"
  "$CUDA_DIR/src/pbs/programmable_bootstrap_classic.cu"
  "$CUDA_DIR/src/pbs/programmable_bootstrap_multibit.cu"
  "$CUDA_DIR/src/pbs/bootstrapping_key.cu"
  "$CUDA_DIR/src/pbs/programmable_bootstrap.cu"
  "$CUDA_DIR/src/fft/twiddles.cu"
  "$CUDA_DIR/src/fft/fft16x4x16_twiddles.cu"
  "$CUDA_DIR/src/fft128/twiddles.cu"
  "$CUDA_DIR/src/crypto/keyswitch.cu"
  "$CUDA_DIR/src/crypto/ciphertext.cu"
  "$CUDA_DIR/src/utils/helper_multi_gpu.cu"
)
FLAGS=(-gencode arch=compute_100,code=sm_100 -O3 -std=c++17 --no-exceptions
  --expt-relaxed-constexpr -rdc=true --use_fast_math -lineinfo
  -Xcompiler -fPIC -Xcompiler -fopenmp -DCUDA_ARCH=1000
  -I"$CUDA_DIR/include" -I"$CUDA_DIR/src" -I"$COMMON/include")
pids=()
objs=()
for s in "${SRCS[@]}"; do
  o=$OBJ/$(basename "$(dirname "$s")")_$(basename "${s%.cu}").o
  objs+=("$o")
  if [ -f "$o" ] && [ "$o" -nt "$s" ]; then continue; fi
  ( nvcc "${FLAGS[@]}" -c "$s" -o "$o" > "$o.log" 2>&1 || { echo "FAILED: $s"; tail -30 "$o.log"; exit 1; } ) &
  pids+=($!)
  while [ "$(jobs -rp | wc -l)" -ge "$JOBS" ]; do sleep 1; done
done
for p in "${pids[@]:-}"; do [ -n "$p" ] && wait "$p"; done
SO=$OUT/libtfhe_cuda_backend_ref.so
fresh=1
[ -f "$SO" ] || fresh=0
for o in "${objs[@]}"; do [ "$o" -nt "$SO" ] && fresh=0; done
if [ "$fresh" = 1 ]; then echo "up to date: $SO"; exit 0; fi
nvcc -gencode arch=compute_100,code=sm_100 -shared -Xcompiler -fPIC -Xcompiler -fopenmp \
  "${objs[@]}" -o "$OUT/libtfhe_cuda_backend_ref.so" -lcudart -lgomp
echo "built $OUT/libtfhe_cuda_backend_ref.so"
