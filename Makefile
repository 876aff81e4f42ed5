# Builds the C-ABI library in-tree (the .so travels to the GPU box with the snapshot).  One object per engine
# translation unit, so `make -j` compiles them in parallel.
NVCC ?= /usr/local/cuda/bin/nvcc
ARCH := -gencode arch=compute_100a,code=sm_100a
NVFLAGS := $(ARCH) -O3 -std=c++17 -lineinfo -Xcompiler -fPIC -Xptxas -v --expt-relaxed-constexpr
CSRC := fac_fake_b200/csrc
UNITS := ff_host ff_cvit ff_resvitkan ff_ggca ff_s3d ff_blaze
OBJ := $(patsubst %,build/%.o,$(UNITS))
HDR := $(wildcard $(CSRC)/*.cuh) $(wildcard $(CSRC)/*.h) include/facfake.h
LIB := fac_fake_b200/libfacfake.so

all: $(LIB)

build/%.o: $(CSRC)/%.cu $(HDR)
	@mkdir -p build
	$(NVCC) $(NVFLAGS) -c -o $@ $< 2> build/$*.ptxas.log || (cat build/$*.ptxas.log; exit 1)

$(LIB): $(OBJ)
	$(NVCC) $(ARCH) -shared -o $@ $(OBJ)
	@cat $(patsubst %,build/%.ptxas.log,$(UNITS)) > build_ptxas.log

# Developer build with the clock64 pipeline traces of the encoder and the fused layer-1/2 kernel compiled in
# (tools/xf_trace.py loads it; profiles/r02_xf_trace.txt, profiles/r02_c12_trace.txt).  Never loaded by the package.
trace: $(LIB)
	$(NVCC) $(NVFLAGS) -DFF_XF_TRACE -DFF_C12_TRACE -DFF_PTC_TRACE -c -o build/ff_cvit_trace.o $(CSRC)/ff_cvit.cu 2> build/ff_cvit_trace.ptxas.log
	$(NVCC) $(ARCH) -shared -o build/libfacfake_trace.so build/ff_cvit_trace.o $(filter-out build/ff_cvit.o,$(OBJ))

clean:
	rm -rf build $(LIB) build_ptxas.log
.PHONY: all clean trace
