# Top-level build: liblsdsort.so (the product) + the oracle checkers.
# `python -c "import __graft_entry__ as g; g.build()"` runs the same commands.
NVCC    ?= nvcc
CXX     ?= g++
CUDA_HOME ?= /usr/local/cuda
GENCODE := -gencode arch=compute_100a,code=sm_100a
NVFLAGS := -O3 -std=c++17 -lineinfo $(GENCODE) -Xcompiler -fPIC -Xcompiler -fvisibility=hidden --expt-relaxed-constexpr
CSRC    := lsdradixsort_b200/csrc
OBJDIR  := build/obj
# TUNING=1 appends the measured-and-rejected kernel families (onesweep_r8_{b,c,d}.cu and the #ifdef'd table entries) as
# lsd_sort_options.variant values: a separate library, lsdradixsort_b200/liblsdsort_tuning.so (use it with
# LSDSORT_LIB=...; bench_tools/ and the variant tests need it).  The product library holds the shipped shapes only.
TUNING  ?= 0
SRCS    := api.cu multi.cu sort.cu keys64.cu histogram.cu scan.cu onesweep_r1.cu onesweep_r2.cu onesweep_r4.cu onesweep_r8.cu onesweep_r8_a.cu
ifeq ($(TUNING),1)
SRCS    += onesweep_r8_b.cu onesweep_r8_c.cu onesweep_r8_d.cu
NVFLAGS += -DLSD_TUNING_VARIANTS
OBJDIR  := build/obj_tuning
LIB     := lsdradixsort_b200/liblsdsort_tuning.so
else
LIB     := lsdradixsort_b200/liblsdsort.so
endif
OBJS    := $(addprefix $(OBJDIR)/,$(SRCS:.cu=.o))

.PHONY: all lib oracle tools clean
all: lib oracle tools

lib: $(LIB)

$(OBJDIR)/%.o: $(CSRC)/%.cu $(wildcard $(CSRC)/*.cuh) $(wildcard $(CSRC)/*.h) include/lsdsort.h
	@mkdir -p $(OBJDIR)
	$(NVCC) $(NVFLAGS) -Xptxas -v -c $< -o $@ 2> $(OBJDIR)/$*.ptxas.log || (cat $(OBJDIR)/$*.ptxas.log; false)

$(LIB): $(OBJS)
	$(NVCC) $(GENCODE) -shared -o $@ $(OBJS)

# lsd_bench: the reference's Test*/Benchmark* drivers over the C ABI (tools/lsd_bench.cpp)
tools: build/lsd_bench build/lsd_multi_check
# lsd_multi_check: lsd_sort_multi called from C++ threads with NCCL communicators (tools/lsd_multi_check.cpp); skipped
# without nccl.h
build/lsd_multi_check: tools/lsd_multi_check.cpp include/lsdsort.h include/lsdsort_nccl.h $(LIB)
	@mkdir -p build
	@if [ -f /usr/include/nccl.h ]; then \
	  $(CXX) -O2 -std=c++17 -pthread -Iinclude -I$(CUDA_HOME)/include -o $@ $< -Llsdradixsort_b200 -llsdsort \
	    -L$(CUDA_HOME)/lib64 -lcudart -lnccl -Wl,-rpath,'$$ORIGIN/../lsdradixsort_b200' ; \
	else echo "nccl.h not found: skipping lsd_multi_check"; fi
build/lsd_bench: tools/lsd_bench.cpp include/lsdsort.h $(LIB)
	@mkdir -p build
	$(CXX) -O2 -std=c++17 -Iinclude -I$(CUDA_HOME)/include -o $@ $< -Llsdradixsort_b200 -llsdsort \
	    -L$(CUDA_HOME)/lib64 -lcudart -Wl,-rpath,'$$ORIGIN/../lsdradixsort_b200'

oracle:
	$(MAKE) -C oracle oracle ref

clean:
	rm -rf build $(LIB)
	$(MAKE) -C oracle clean
