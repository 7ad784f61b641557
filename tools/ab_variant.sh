#!/bin/bash
# Build an A/B variant of libmh_b200.so next to the default one and (on a GPU box) compare them.
#   tools/ab_variant.sh build fwd4cta -DMH_FWD_KV_STAGES=1     # here (no GPU needed): -> speech_ssl_compression_b200/variants/libmh_b200_fwd4cta.so
#   tools/ab_variant.sh run   fwd4cta                          # under gpurun: attention tests + in-graph timeline with both libraries
set -e
ROOT="$(cd "$(dirname "$0")/.." && pwd)"
CSRC="$ROOT/speech_ssl_compression_b200/csrc"
VAR="$ROOT/speech_ssl_compression_b200/variants"
cmd="$1"; name="$2"; shift 2 || true
case "$cmd" in
  build)
    mkdir -p "$VAR" "/tmp/mh_variant_$name"
    for f in "$CSRC"/*.cu; do
      nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -lineinfo -Xcompiler -fPIC -I"$ROOT/include" "$@" \
           -c "$f" -o "/tmp/mh_variant_$name/$(basename "$f" .cu).o" &
    done
    wait
    nvcc -gencode arch=compute_100a,code=sm_100a -shared -o "$VAR/libmh_b200_$name.so" /tmp/mh_variant_$name/*.o -lcudart_static -lpthread -ldl -lrt
    ls -la "$VAR/libmh_b200_$name.so"
    ;;
  run)
    mkdir -p "$ROOT/gpurun_out"
    for lib in default "$name"; do
      if [ "$lib" = default ]; then unset MH_B200_LIB; else export MH_B200_LIB="$VAR/libmh_b200_$lib.so"; fi
      echo "== $lib"
      timeout 120 python -m pytest "$ROOT/tests/test_gpu_kernels.py" -q -x -k attention 2>&1 | tail -2
      timeout 100 python "$ROOT/tools/profile_step.py" --graph 2>&1 | grep -A4 "^wall" | tee "$ROOT/gpurun_out/ab_${lib}_timeline.txt"
    done
    ;;
  *) echo "usage: $0 build|run NAME [nvcc flags]"; exit 1;;
esac
