# Builds the C-ABI library in-tree (the .so travels to the GPU box with the snapshot).
NVCC ?= /usr/local/cuda/bin/nvcc
ARCH := -gencode arch=compute_100a,code=sm_100a
NVFLAGS := $(ARCH) -O3 -std=c++17 -lineinfo -Xcompiler -fPIC -Xptxas -v --expt-relaxed-constexpr
SRC := fac_fake_b200/csrc/ff_engine.cu fac_fake_b200/csrc/ff_blaze.cu
HDR := $(wildcard fac_fake_b200/csrc/*.cuh) include/facfake.h
LIB := fac_fake_b200/libfacfake.so

all: $(LIB)

$(LIB): $(SRC) $(HDR)
	$(NVCC) $(NVFLAGS) -shared -o $@ $(SRC) 2> build_ptxas.log || (cat build_ptxas.log; exit 1)

clean:
	rm -f $(LIB) build_ptxas.log
.PHONY: all clean
