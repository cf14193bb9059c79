# Builds the C-ABI CUDA library in-tree for sm_100a (B200).  nvcc cross-compiles without a GPU.
NVCC ?= /usr/local/cuda/bin/nvcc
ARCH := -gencode arch=compute_100a,code=sm_100a
NVCCFLAGS := -O3 -std=c++17 -lineinfo $(ARCH) -Xcompiler -fPIC -Xcompiler -Wall -Xcompiler -Wno-unused-function
PKG := tensornetworksfork_b200
SRC := $(wildcard $(PKG)/csrc/*.cu)
OBJ := $(patsubst $(PKG)/csrc/%.cu,build/%.o,$(SRC))
LIB := $(PKG)/libtn_b200.so

all: $(LIB)

build/%.o: $(PKG)/csrc/%.cu $(PKG)/csrc/common.cuh include/tn_b200.h $(wildcard $(PKG)/csrc/*.cuh)
	@mkdir -p build
	$(NVCC) $(NVCCFLAGS) $(EXTRA) -c $< -o $@

$(LIB): $(OBJ)
	$(NVCC) $(ARCH) -shared -o $@ $(OBJ) -lcuda

clean:
	rm -rf build $(LIB)

.PHONY: all clean
