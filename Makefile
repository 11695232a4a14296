# Build everything in-tree (the .so files travel to the GPU box with the snapshot).
#   make            -> gpu library, host mirror, oracle, C++ test binaries
#   make gpu|host|oracle
NVCC      ?= /usr/local/cuda/bin/nvcc
CXX       ?= g++
CC        ?= gcc
ARCH      := -gencode arch=compute_100a,code=sm_100a
# -fmad=false: the reference's f64 arithmetic has no fused multiply-add; the
# device must round exactly like the oracle (BASELINE.json north_star).
NVFLAGS   := -O3 -std=c++17 $(ARCH) -lineinfo -fmad=false -Xcompiler -fPIC -Xptxas -v -Iinclude
CXXFLAGS  := -std=c++20 -O2 -ffp-contract=off -fPIC -Wall -Wextra -Iinclude
CFLAGS    := -std=c11 -O2 -ffp-contract=off -fPIC -Wall -Wextra -Iinclude

PKG       := portrayer_b200
GPU_LIB   := $(PKG)/lib/libportrayer_gpu.so
HOST_LIB  := $(PKG)/lib/libportrayer_host.so
# scene_blob.c + tiles.c alone: the pure-host part of the C ABI (pack / unpack / tile ownership), for processes that
# must not map the CUDA library (the host mirror's scene half, bench.py --impl reference)
BLOB_LIB  := $(PKG)/lib/libportrayer_blob.so
# Image::render of the host mirror: the one part of it that calls libportrayer_gpu.so
RENDER_LIB := $(PKG)/lib/libportrayer_render.so
ORACLE_LIB:= oracle/liboracle.so
HOST_TEST := tests/cpp/test_host
GROUP_TEST := tests/cpp/test_group
EXAMPLE_BIN := $(PKG)/lib/portrayer_example

GPU_SRC   := $(wildcard $(PKG)/csrc/*.cu)
GPU_HDR   := $(wildcard $(PKG)/csrc/*.cuh) $(wildcard $(PKG)/csrc/*.h) include/portrayer_gpu.h
RENDER_SRC := $(PKG)/host/render.cpp $(PKG)/host/capi_render.cpp
HOST_SRC  := $(filter-out $(RENDER_SRC),$(wildcard $(PKG)/host/*.cpp)) $(wildcard $(PKG)/host/examples/*.cpp)
HOST_HDR  := $(wildcard $(PKG)/host/*.hpp) $(wildcard $(PKG)/host/*.h) $(wildcard $(PKG)/host/examples/*.hpp) include/portrayer_gpu.h

all: gpu host oracle hosttest example
gpu: $(GPU_LIB)
host: $(HOST_LIB) $(RENDER_LIB)
oracle: $(ORACLE_LIB)
hosttest: $(HOST_TEST) $(GROUP_TEST)
example: $(EXAMPLE_BIN)

$(PKG)/lib:
	mkdir -p $@

build/%.o: $(PKG)/csrc/%.c include/portrayer_gpu.h
	mkdir -p build
	$(CC) $(CFLAGS) -c $< -o $@

$(GPU_LIB): $(GPU_SRC) $(GPU_HDR) build/scene_blob.o build/tiles.o | $(PKG)/lib
	$(NVCC) $(NVFLAGS) -shared $(GPU_SRC) build/scene_blob.o build/tiles.o -o $@ 2> build/ptxas_gpu.log || (cat build/ptxas_gpu.log; false)
	@grep -E "error|warning" build/ptxas_gpu.log || true

$(BLOB_LIB): build/scene_blob.o build/tiles.o | $(PKG)/lib
	$(CC) -shared build/scene_blob.o build/tiles.o -o $@

# this image's g++ links libstdc++ statically: keep that copy private to the library (not exported, bound at link
# time), or it gets interposed by whatever libstdc++.so another module of the process (numpy's OpenBLAS, the CUDA
# library) has already loaded, and two C++ runtimes end up sharing state
CXX_SO_FLAGS := -Wl,-Bsymbolic -Wl,--exclude-libs,ALL
$(HOST_LIB): $(HOST_SRC) $(HOST_HDR) $(BLOB_LIB) | $(PKG)/lib
	$(CXX) $(CXXFLAGS) -shared $(HOST_SRC) -o $@ $(CXX_SO_FLAGS) -L$(PKG)/lib -lportrayer_blob -Wl,-rpath,'$$ORIGIN'

$(RENDER_LIB): $(RENDER_SRC) $(HOST_HDR) $(HOST_LIB) $(GPU_LIB) | $(PKG)/lib
	$(CXX) $(CXXFLAGS) -shared $(RENDER_SRC) -o $@ $(CXX_SO_FLAGS) -L$(PKG)/lib -lportrayer_host -lportrayer_gpu -Wl,-rpath,'$$ORIGIN'

# the CPU baseline: -O3 as BASELINE.md section 2 says (still -ffp-contract=off: no FMA, like the reference)
$(ORACLE_LIB): oracle/oracle.c oracle/oracle.h include/portrayer_gpu.h
	$(CC) $(CFLAGS) -O3 -shared oracle/oracle.c -o $@ -lm -lpthread

$(HOST_TEST): tests/cpp/test_host.cpp $(HOST_LIB) $(RENDER_LIB)
	$(CXX) $(CXXFLAGS) -I$(PKG)/host $< -o $@ -L$(PKG)/lib -lportrayer_render -lportrayer_host -lportrayer_gpu -lportrayer_blob -Wl,-rpath,'$$ORIGIN/../../$(PKG)/lib'

$(GROUP_TEST): tests/cpp/test_group.cpp $(HOST_LIB) $(RENDER_LIB)
	$(CXX) $(CXXFLAGS) -I$(PKG)/host $< -o $@ -L$(PKG)/lib -lportrayer_render -lportrayer_host -lportrayer_gpu -lportrayer_blob -Wl,-rpath,'$$ORIGIN/../../$(PKG)/lib'

$(EXAMPLE_BIN): $(PKG)/host/tools/example_main.cpp $(HOST_LIB) $(RENDER_LIB)
	$(CXX) $(CXXFLAGS) -I$(PKG)/host $< -o $@ -L$(PKG)/lib -lportrayer_render -lportrayer_host -lportrayer_gpu -lportrayer_blob -Wl,-rpath,'$$ORIGIN'

clean:
	rm -rf build $(PKG)/lib $(ORACLE_LIB) $(HOST_TEST) $(GROUP_TEST)

.PHONY: all gpu host oracle hosttest example clean
