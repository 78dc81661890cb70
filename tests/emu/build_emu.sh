#!/bin/sh
# TEST INFRASTRUCTURE: compile the product sources against the host SIMT emulator.
set -e
cd "$(dirname "$0")/../.."
g++ -O1 -g -std=c++20 -DEQ_HOST_EMU -ffp-contract=off -fno-fast-math -mno-fma -fPIC -shared -pthread \
    -Itests/emu -x c++ equilibrium_b200/csrc/eq_api.cu tests/emu/cuda_emu.cpp \
    -o tests/emu/libequilibrium_emu.so -lrt
