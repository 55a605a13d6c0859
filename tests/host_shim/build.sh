#!/bin/sh
# TEST INFRASTRUCTURE ONLY: g++ build of monocon_pytorch_b200/csrc/train_backward.cu as host code (see host_shim.h).
set -e
here="$(cd "$(dirname "$0")" && pwd)"
root="$(cd "$here/../.." && pwd)"
mkdir -p "$here/_build"
g++ -O2 -std=c++17 -fPIC -shared -DMC_HOST_SHIM -I"$here" -I"$root/monocon_pytorch_b200/csrc" \
    -x c++ "$root/monocon_pytorch_b200/csrc/train_backward.cu" -o "$here/_build/libtrain_backward_host.so"
