# Native build without Python.  `make` = the product library; the other targets are test infrastructure.
NVCC      ?= nvcc
NVCCFLAGS ?= -O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a -lineinfo -fmad=false -Xcompiler -fPIC -shared
CSRC      := $(wildcard equilibrium_b200/csrc/*)
LIB       := equilibrium_b200/libequilibrium_cuda.so
EMU       := tests/emu/libequilibrium_emu.so
ORACLE    := oracle/libfluid_ref.so

all: $(LIB)

$(LIB): $(CSRC) include/equilibrium_cuda.h
	$(NVCC) $(NVCCFLAGS) -o $@ equilibrium_b200/csrc/eq_api.cu

# the same sources against the host SIMT emulator (CPU tests)
$(EMU): $(CSRC) include/equilibrium_cuda.h tests/emu/cuda_emu.h tests/emu/cuda_emu.cpp
	tests/emu/build_emu.sh

$(ORACLE): oracle/fluid_ref.c oracle/fluid_ref.h
	$(MAKE) -C oracle

# the C++ host mirror (include/equilibrium.hpp) against the oracle: on the GPU, or emulated on the CPU
build/host_mirror_test: tests/cpp/host_mirror_test.cpp include/equilibrium.hpp $(LIB) $(ORACLE)
	mkdir -p build
	g++ -std=c++17 -O1 -Wall -Wextra -Iinclude -Ioracle $< -o $@ -Lequilibrium_b200 -l:libequilibrium_cuda.so \
	    -Loracle -l:libfluid_ref.so -Wl,-rpath,$(CURDIR)/equilibrium_b200 -Wl,-rpath,$(CURDIR)/oracle
build/host_mirror_test_emu: tests/cpp/host_mirror_test.cpp include/equilibrium.hpp $(EMU) $(ORACLE)
	mkdir -p build
	g++ -std=c++17 -O1 -Wall -Wextra -Iinclude -Ioracle $< -o $@ -Ltests/emu -l:libequilibrium_emu.so \
	    -Loracle -l:libfluid_ref.so -Wl,-rpath,$(CURDIR)/tests/emu -Wl,-rpath,$(CURDIR)/oracle

test-cpp: build/host_mirror_test
	build/host_mirror_test
test-cpp-emu: build/host_mirror_test_emu
	EQ_EMU_SMS=4 build/host_mirror_test_emu

.PHONY: all test-cpp test-cpp-emu
