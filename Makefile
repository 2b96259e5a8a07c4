# Builds libuglad_b200.so (sm_100a only) and the oracle helpers.
NVCC ?= /usr/local/cuda/bin/nvcc
ARCH := -gencode arch=compute_100a,code=sm_100a
NVFLAGS := -O3 -std=c++17 -lineinfo $(ARCH) -Xcompiler -fPIC -Xptxas -v
SRC := $(wildcard uglad_b200/csrc/*.cu)
OBJ := $(SRC:.cu=.o)
LIB := uglad_b200/lib/libuglad_b200.so

all: $(LIB)

%.o: %.cu uglad_b200/csrc/common.cuh uglad_b200/csrc/kernels.cuh include/uglad_b200.h
	$(NVCC) $(NVFLAGS) -c $< -o $@

$(LIB): $(OBJ)
	mkdir -p uglad_b200/lib
	$(NVCC) $(ARCH) -shared -o $@ $(OBJ) -lcudart

clean:
	rm -f $(OBJ) $(LIB)
