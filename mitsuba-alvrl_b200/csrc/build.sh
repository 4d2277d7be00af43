#!/bin/bash
# Builds libalvrl.so in-tree for sm_100a.  The strict flavour and the exact-arithmetic kernels are compiled with
# -fmad=false; host code never uses FMA contraction either (bit-exact slice / cluster decisions).
set -e
cd "$(dirname "$0")"
OUT=../libalvrl.so
NV="nvcc $ALVRL_EXTRA_NVCC -gencode arch=compute_100a,code=sm_100a -lineinfo -O3 -std=c++17 -Xcompiler -fPIC,-ffp-contract=off,-fno-fast-math,-Wall,-Wno-unused-function"
mkdir -p obj
build_one() { # src flags...
  local src=$1; shift
  local obj=obj/${src%.cu}.o
  if [ ! -f "$obj" ] || [ -n "$(find . -maxdepth 1 \( -name '*.cu' -o -name '*.cuh' -o -name '*.h' -o -name '*.inl' \) -newer "$obj" -print -quit)" ] \
     || [ ../../include/alvrl.h -nt "$obj" ] || [ ../../include/alvrl_rng.h -nt "$obj" ]; then
    rm -f "$obj"            # a failed compile must never leave a stale object for the link step
    $NV "$@" -c "$src" -o "$obj" &
    pids+=($!)
  fi
}
pids=()
build_one primary.cu -fmad=false
build_one transport_strict.cu -fmad=false
build_one transport_fast.cu
build_one clustering.cu -fmad=false
build_one capi.cu -fmad=false
build_one group.cu -fmad=false
build_one slices_dev.cu -fmad=false
build_one film.cu -fmad=false
build_one chain.cu -fmad=false
build_one tracer.cu -fmad=false
build_one volpath.cu -fmad=false
fail=0
for p in "${pids[@]}"; do wait "$p" || fail=1; done
if [ $fail -ne 0 ]; then echo "build.sh: compilation failed" >&2; exit 1; fi
nvcc -gencode arch=compute_100a,code=sm_100a -shared -o $OUT obj/primary.o obj/transport_strict.o obj/transport_fast.o obj/clustering.o obj/capi.o obj/group.o obj/slices_dev.o obj/film.o obj/chain.o obj/tracer.o obj/volpath.o -ldl
echo "built $OUT"
