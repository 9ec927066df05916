# Builds the product library (sm_100a only) and the CPU oracle (test infrastructure).
NVCC ?= nvcc
NVCCFLAGS ?= -O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a -lineinfo -Xcompiler -fPIC -Xptxas -v --expt-relaxed-constexpr
CSRC := modppl_b200/csrc
SRCS := $(CSRC)/pf.cu $(CSRC)/is_mh.cu $(CSRC)/multi_gpu.cu $(CSRC)/jit.cu
# the kernel headers a run-time model (csrc/jit.cu) is compiled against by NVRTC: embedded in the library as string literals
JIT_HDRS := common.cuh models.cuh nested_quant.cuh pf_kernels.cuh scan2.cuh nested.cuh
EMBED := modppl_b200/lib/embedded_headers.inc
HDRS := $(wildcard $(CSRC)/*.cuh) $(CSRC)/engine.h include/modppl_b200.h
LIB := modppl_b200/lib/libmodppl_b200.so

all: $(LIB) oracle

OBJS := $(patsubst $(CSRC)/%.cu,modppl_b200/lib/%.o,$(SRCS))

# one object per translation unit so that `make -j` compiles them side by side; ptxas -v output is kept per unit
$(EMBED): $(addprefix $(CSRC)/,$(JIT_HDRS))
	@mkdir -p modppl_b200/lib
	python3 -c "import sys; [sys.stdout.write('{\"%s\", R\"MPLHDR(%s)MPLHDR\"},\n' % (n, open('$(CSRC)/' + n).read())) for n in '$(JIT_HDRS)'.split()]" > $@

modppl_b200/lib/jit.o: $(EMBED)

modppl_b200/lib/%.o: $(CSRC)/%.cu $(HDRS)
	@mkdir -p modppl_b200/lib
	$(NVCC) $(NVCCFLAGS) -c -o $@ $< 2> modppl_b200/lib/$*.ptxas.log || (cat modppl_b200/lib/$*.ptxas.log; false)

$(LIB): $(OBJS)
	$(NVCC) -gencode arch=compute_100a,code=sm_100a -shared -o $@ $(OBJS) -ldl
	@cat modppl_b200/lib/*.ptxas.log > modppl_b200/lib/ptxas.log

oracle:
	$(MAKE) -C oracle

clean:
	rm -rf modppl_b200/lib oracle/libmodppl_oracle.so

.PHONY: all oracle clean
