# Builds the product library (sm_100a only) and the CPU oracle (test infrastructure).
NVCC ?= nvcc
NVCCFLAGS ?= -O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a -lineinfo -Xcompiler -fPIC -Xptxas -v --expt-relaxed-constexpr
CSRC := modppl_b200/csrc
SRCS := $(CSRC)/pf.cu $(CSRC)/is_mh.cu $(CSRC)/multi_gpu.cu
HDRS := $(wildcard $(CSRC)/*.cuh) $(CSRC)/engine.h include/modppl_b200.h
LIB := modppl_b200/lib/libmodppl_b200.so

all: $(LIB) oracle

OBJS := $(patsubst $(CSRC)/%.cu,modppl_b200/lib/%.o,$(SRCS))

# one object per translation unit so that `make -j` compiles them side by side; ptxas -v output is kept per unit
modppl_b200/lib/%.o: $(CSRC)/%.cu $(HDRS)
	@mkdir -p modppl_b200/lib
	$(NVCC) $(NVCCFLAGS) -c -o $@ $< 2> modppl_b200/lib/$*.ptxas.log || (cat modppl_b200/lib/$*.ptxas.log; false)

$(LIB): $(OBJS)
	$(NVCC) -shared -o $@ $(OBJS)
	@cat modppl_b200/lib/*.ptxas.log > modppl_b200/lib/ptxas.log

oracle:
	$(MAKE) -C oracle

clean:
	rm -rf modppl_b200/lib oracle/libmodppl_oracle.so

.PHONY: all oracle clean
