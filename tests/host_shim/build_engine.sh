#!/bin/sh
# TEST INFRASTRUCTURE ONLY: the engine's host logic (api.cu + engine.cu, compiled by g++) linked against a stand-in CUDA runtime
# and host versions of the train-mode forward launchers (host_engine_stubs.cpp) and the host-shim build of train_backward.cu.
# -Bsymbolic: the stand-in cuda* symbols must win over a real libcudart that torch may already have loaded into the process.
set -e
here="$(cd "$(dirname "$0")" && pwd)"
root="$(cd "$here/../.." && pwd)"
src="$root/monocon_pytorch_b200/csrc"
out="$here/_build"
mkdir -p "$out"
CUDA_INC="${CUDA_HOME:-/usr/local/cuda}/include"
FLAGS="-O2 -std=c++17 -fPIC -I$CUDA_INC -I$src"
g++ $FLAGS -x c++ -c "$src/api.cu" -o "$out/api.o"
g++ $FLAGS -x c++ -c "$src/engine.cu" -o "$out/engine.o"
g++ $FLAGS -c "$here/host_engine_stubs.cpp" -o "$out/stubs.o"
g++ -O2 -std=c++17 -fPIC -DMC_HOST_SHIM -I"$here" -I"$src" -x c++ -c "$src/train_backward.cu" -o "$out/train_backward_host.o"
g++ -shared -Wl,-Bsymbolic -o "$out/libmonocon_host_engine.so" "$out/api.o" "$out/engine.o" "$out/stubs.o" "$out/train_backward_host.o"
