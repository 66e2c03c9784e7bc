# Builds libsvdd_b200.so (sm_100a only) in-tree so it travels with the repo
# snapshot to the GPU box.  `python -c "import __graft_entry__ as g; g.build()"`
# runs this.
NVCC      ?= nvcc
ARCH      := -gencode arch=compute_100a,code=sm_100a
NVCCFLAGS := $(ARCH) -O3 -std=c++17 -lineinfo -Xcompiler -fPIC -Xcompiler -Wall \
             --expt-relaxed-constexpr -Xptxas -v $(EXTRA_NVCCFLAGS)
CSRC      := svdd_b200/csrc
SRCS      := $(wildcard $(CSRC)/*.cu)
OBJS      := $(patsubst $(CSRC)/%.cu,build/%.o,$(SRCS))
HDRS      := $(wildcard $(CSRC)/*.cuh) include/svdd_b200.h
LIB       := svdd_b200/libsvdd_b200.so

all: $(LIB)

build/%.o: $(CSRC)/%.cu $(HDRS)
	@mkdir -p build
	$(NVCC) $(NVCCFLAGS) -c $< -o $@ 2> build/$*.ptxas.log || (cat build/$*.ptxas.log; exit 1)

$(LIB): $(OBJS)
	$(NVCC) $(ARCH) -shared -o $@ $(OBJS)

clean:
	rm -rf build $(LIB)

.PHONY: all clean
