# Builds the B200 tensor-verification engine in-tree.
# (-l:libstdc++.so.6 is spelled out on every g++ link line: the compiler wrapper of the build image, CXX=/opt/gcc/bin/g++,
#  finds only libstdc++.a and links it STATICALLY into shared objects; with GCC 13 the stream / locale initialiser lives in
#  the library, the static copy never runs it, and the first `istream >> double` inside libqsxhost.so crashed)
#   make            -> qsyn_b200/lib/libqsx.so   (C ABI of include/qsx.h; CUDA sm_100a)
#   make oracle     -> oracle/_build/*           (CPU checker, test infrastructure only)
NVCC      ?= /usr/local/cuda/bin/nvcc
CXX       ?= g++
ARCH      := -gencode arch=compute_100a,code=sm_100a
NVFLAGS   := -O3 -std=c++20 -lineinfo $(ARCH) -Xcompiler -fPIC,-Wall,-Wextra -Xptxas -v
CXXFLAGS  := -O2 -std=c++20 -fPIC -Wall -Wextra -Werror
# gcc 13 reports a false -Wfree-nonheap-object inside std::vector<T*> growth at -O2
HOSTFLAGS := $(CXXFLAGS) -Wno-free-nonheap-object

CSRC      := qsyn_b200/csrc
LIBDIR    := qsyn_b200/lib
OBJDIR    := build/obj

LIBQSX    := $(LIBDIR)/libqsx.so
OBJS      := $(OBJDIR)/kernels.o $(OBJDIR)/runtime.o $(OBJDIR)/planner.o $(OBJDIR)/planner_lower.o \
             $(OBJDIR)/planner_schedule.o $(OBJDIR)/planner_emit.o $(OBJDIR)/comm.o \
             $(OBJDIR)/jit.o $(OBJDIR)/jit_codegen.o
CUDAINC   := -I/usr/local/cuda/include

HOST      := qsyn_b200/host
SHIM      := bin/qsyn-b200
HOSTHDRS  := $(wildcard $(HOST)/*/*.hpp) include/qsx.h

FP64PEAK  := bin/fp64_peak
BCAST     := bin/bcast_probe

LIBHOST   := $(LIBDIR)/libqsxhost.so

all: $(LIBQSX) $(LIBHOST) $(SHIM) $(FP64PEAK) $(BCAST)

$(OBJDIR)/host_%.o: $(HOST)/capi/%.cpp $(HOSTHDRS) include/qsx_host.h
	@mkdir -p $(OBJDIR)
	$(CXX) $(HOSTFLAGS) -c $< -o $@

# C ABI of the host layer (include/qsx_host.h) for non-C++ callers (tests, bench.py)
HOSTOBJS  := $(OBJDIR)/host_qcir_io.o $(OBJDIR)/host_qcir_to_tensor.o $(OBJDIR)/host_zx_io.o $(OBJDIR)/host_zxgraph_to_tensor.o

$(LIBHOST): $(HOSTOBJS) $(OBJDIR)/host_qsx_host.o $(LIBQSX)
	$(CXX) -shared -o $@ $(HOSTOBJS) $(OBJDIR)/host_qsx_host.o \
	    -L$(LIBDIR) -lqsx -l:libstdc++.so.6 -lm -Wl,-rpath,'$$ORIGIN'

$(BCAST): $(CSRC)/tools/bcast_probe.cu
	@mkdir -p bin
	$(NVCC) -O3 -std=c++20 -lineinfo $(ARCH) -o $@ $<

$(FP64PEAK): $(CSRC)/tools/fp64_peak.cu
	@mkdir -p bin
	$(NVCC) -O3 -std=c++20 -lineinfo $(ARCH) -o $@ $<

# host layer (reference-shaped QTensor / to_tensor / QCir API) + the dofile-runner shim
$(OBJDIR)/host_%.o: $(HOST)/qcir/%.cpp $(HOSTHDRS)
	@mkdir -p $(OBJDIR)
	$(CXX) $(HOSTFLAGS) -c $< -o $@
$(OBJDIR)/host_%.o: $(HOST)/convert/%.cpp $(HOSTHDRS)
	@mkdir -p $(OBJDIR)
	$(CXX) $(HOSTFLAGS) -c $< -o $@
$(OBJDIR)/host_%.o: $(HOST)/zx/%.cpp $(HOSTHDRS)
	@mkdir -p $(OBJDIR)
	$(CXX) $(HOSTFLAGS) -c $< -o $@

$(OBJDIR)/host_%.o: $(HOST)/cli/%.cpp $(HOSTHDRS)
	@mkdir -p $(OBJDIR)
	$(CXX) $(HOSTFLAGS) -c $< -o $@

$(SHIM): $(HOSTOBJS) $(OBJDIR)/host_shim_main.o $(LIBQSX)
	@mkdir -p bin
	$(CXX) -o $@ $(HOSTOBJS) $(OBJDIR)/host_shim_main.o \
	    -L$(LIBDIR) -lqsx -l:libstdc++.so.6 -lm -Wl,-rpath,'$$ORIGIN/../$(LIBDIR)'

$(OBJDIR)/%.o: $(CSRC)/%.cu $(wildcard $(CSRC)/*.hpp) include/qsx.h $(wildcard qsyn_b200/host/util/*.hpp)
	@mkdir -p $(OBJDIR)
	$(NVCC) $(NVFLAGS) -c $< -o $@ 2> $(OBJDIR)/$*.ptxas.log || (cat $(OBJDIR)/$*.ptxas.log; false)

$(OBJDIR)/%.o: $(CSRC)/%.cpp $(wildcard $(CSRC)/*.hpp) include/qsx.h
	@mkdir -p $(OBJDIR)
	$(CXX) $(CXXFLAGS) $(CUDAINC) -c $< -o $@

$(LIBQSX): $(OBJS)
	@mkdir -p $(LIBDIR)
	$(NVCC) $(ARCH) -shared -o $@ $(OBJS) -lcudart -ldl

# A/B builds of the kernels: make variant NAME=wide DEFS="-DQSX_EARLY_PERMS=16 -DQSX_EARLY_ORUNS=8"
#   -> build/variants/$(NAME)/libqsx.so (+ libqsxhost.so beside it); time them with scripts/ab_probe.py
variant: $(LIBQSX) $(LIBHOST)
	@test -n "$(NAME)" || (echo "usage: make variant NAME=<name> DEFS=<-D flags>"; exit 1)
	@mkdir -p build/variants/$(NAME)
	$(NVCC) -O3 -std=c++20 -lineinfo $(ARCH) -Xcompiler -fPIC $(DEFS) -c $(CSRC)/kernels.cu -o build/variants/$(NAME)/kernels.o
	$(NVCC) $(ARCH) -shared -o build/variants/$(NAME)/libqsx.so build/variants/$(NAME)/kernels.o $(filter-out $(OBJDIR)/kernels.o,$(OBJS)) -lcudart -ldl
	cp $(LIBHOST) build/variants/$(NAME)/
	rm build/variants/$(NAME)/kernels.o

oracle: oracle/_build/liboracle_c.so

oracle/_build/liboracle_c.so: oracle/oracle_c.c
	@mkdir -p oracle/_build
	@# no -march=native: the built library travels to the GPU box, whose host CPU may differ
	gcc -O3 -std=c11 -fPIC -shared -Wall -Wextra -o $@ $< -lm

clean:
	rm -rf build $(LIBDIR) bin oracle/_build

.PHONY: variant all clean oracle
