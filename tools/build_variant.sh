#!/bin/bash
# Kernel-variant experiments: rebuild only the fast transport flavour with extra nvcc flags and link it with the objects
# of the regular build into build/libalvrl_<name>.so (select it with ALVRL_LIB=build/libalvrl_<name>.so).
#   tools/build_variant.sh t128c6 -DALVRL_TILE_VRLS=128 -DALVRL_MIN_CTAS=6
set -e
cd "$(dirname "$0")/../mitsuba-alvrl_b200/csrc"
name=$1; shift
mkdir -p ../../build/obj_$name
nvcc -gencode arch=compute_100a,code=sm_100a -lineinfo -O3 -std=c++17 -Xcompiler -fPIC,-ffp-contract=off,-fno-fast-math "$@" -c transport_fast.cu -o ../../build/obj_$name/transport_fast.o
nvcc -gencode arch=compute_100a,code=sm_100a -shared -o ../../build/libalvrl_$name.so obj/primary.o obj/transport_strict.o ../../build/obj_$name/transport_fast.o obj/clustering.o obj/capi.o obj/group.o obj/slices_dev.o obj/film.o obj/chain.o obj/tracer.o obj/volpath.o -ldl
echo "built build/libalvrl_$name.so"
