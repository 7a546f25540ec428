#!/bin/bash
# tests/emul/build.sh -- TEST INFRASTRUCTURE: compiles the product's own .cu sources with g++ against the CUDA-on-CPU
# shim (tests/emul/cuda_runtime.h) into tests/emul/libsb_emul_TESTONLY.so.  Used only by tests/test_emul_cpu.py.
set -e
HERE="$(cd "$(dirname "$0")" && pwd)"
SRC="$HERE/../../r1cs-spartan_b200/csrc"
OUT="$HERE/libsb_emul_TESTONLY.so"
CXX_BIN=/usr/bin/g++; [ -x "$CXX_BIN" ] || CXX_BIN=g++
newest=$(ls -t "$SRC"/*.cu "$SRC"/*.cuh "$SRC"/*.h "$HERE"/*.h "$HERE"/*.cpp "$HERE"/build.sh "$HERE/../../include/spartan_b200.h" | head -1)
if [ -f "$OUT" ] && [ "$OUT" -nt "$newest" ]; then exit 0; fi
mkdir -p "$HERE/obj"
FLAGS="-std=c++20 -O2 -fPIC -pthread -I$HERE -include $HERE/cuda_runtime.h -Wno-unknown-pragmas -Wno-attributes"
pids=()
for f in kernels_fr msm prover comm_shm indexer; do
  $CXX_BIN $FLAGS -x c++ -c "$SRC/$f.cu" -o "$HERE/obj/$f.o" & pids+=($!)
done
$CXX_BIN $FLAGS -c "$HERE/emul_runtime.cpp" -o "$HERE/obj/emul_runtime.o" & pids+=($!)
for p in "${pids[@]}"; do wait $p; done
$CXX_BIN -shared -pthread -o "$OUT" "$HERE"/obj/*.o -lrt
