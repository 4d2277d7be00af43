#!/bin/bash
# AddressSanitizer + UndefinedBehaviorSanitizer over the host-side code that the CPU suite reaches: libalvrl_host.so (the product's
# host headers: slices.h, occluders.h, occ_query.h, sharding.h, hostio.h, shapes.h, heap_order.h, host_sampler.h) and the oracle
# (with the plugin shim on top of it, tests/test_plugin_oracle_cpu.py).  Builds the instrumented libraries into a scratch
# directory, swaps them in for the run and puts the regular builds back.  Last run (round 2, final state): the whole CPU suite
# (150 tests, fuzz tests included) clean, after two fixes of the same kind in test code (memcpy from an empty vector's null
# data() in host_test_api.cpp and in the oracle's getters).
set -e
cd "$(dirname "$0")/.."
T=$(mktemp -d)
SAN="-O1 -g -fsanitize=address,undefined -fno-sanitize-recover=undefined"
PRE="$(gcc -print-file-name=libasan.so) $(gcc -print-file-name=libubsan.so)"
cp mitsuba-alvrl_b200/libalvrl_host.so oracle/liborc.so "$T"/
restore() { cp "$T"/libalvrl_host.so mitsuba-alvrl_b200/libalvrl_host.so; cp "$T"/liborc.so oracle/liborc.so; touch oracle/liborc.so oracle/liborc_fast.so; }
trap restore EXIT
g++ -std=c++17 $SAN -fPIC -shared -ffp-contract=off -fno-fast-math -I/usr/local/cuda/include mitsuba-alvrl_b200/csrc/host_test_api.cpp -o mitsuba-alvrl_b200/libalvrl_host.so
(cd oracle && g++ -std=c++17 $SAN -fPIC -shared -pthread -ffp-contract=off -fno-fast-math -Wno-sign-compare oracle_capi.cpp -o liborc.so)
touch oracle/liborc.so oracle/liborc_fast.so
LD_PRELOAD="$PRE" ASAN_OPTIONS=detect_leaks=0 python -m pytest tests -m "not gpu" -q -x -p no:cacheprovider \
    --deselect tests/test_sharding_cpu.py "$@"
