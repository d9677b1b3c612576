# Builds the product library (CUDA, sm_100a) and the checkers.
NVCC ?= nvcc
ARCH := -gencode arch=compute_100a,code=sm_100a
NVFLAGS := -O3 -std=c++17 -lineinfo $(ARCH) -Xcompiler -fPIC -Xcompiler -Wall
CSRC := rabbitsalign_b200/csrc
LIB  := rabbitsalign_b200/librsa_ext.so

all: $(LIB) tools/dpx_microbench oracle

tools/dpx_microbench: tools/dpx_microbench.cu $(CSRC)/fast_cell.cuh $(CSRC)/common.cuh
	$(NVCC) -O3 -std=c++17 -lineinfo $(ARCH) -o $@ $<

# three translation units (extension engine, seeding, SAM formatter), one product library
$(CSRC)/engine.o: $(CSRC)/engine.cu $(wildcard $(CSRC)/*.cuh) include/rsa_ext.h
	$(NVCC) $(NVFLAGS) -c -o $@ $<
$(CSRC)/seed.o: $(CSRC)/seed.cu $(CSRC)/kernels_seed.cuh include/rsa_seed.h
	$(NVCC) $(NVFLAGS) -c -o $@ $<
$(CSRC)/sam.o: $(CSRC)/sam.cu include/rsa_sam.h include/rsa_ext.h
	$(NVCC) $(NVFLAGS) -c -o $@ $<
$(LIB): $(CSRC)/engine.o $(CSRC)/seed.o $(CSRC)/sam.o
	$(NVCC) $(ARCH) -shared -o $@ $^

ptxas-info:
	$(NVCC) $(NVFLAGS) -Xptxas -v -c -o /dev/null $(CSRC)/engine.cu

oracle:
	$(MAKE) -C oracle all

clean:
	rm -f $(LIB) $(CSRC)/*.o
	$(MAKE) -C oracle clean
.PHONY: all oracle clean ptxas-info
