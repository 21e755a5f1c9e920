# Non-meson build of goldrush-b200 (meson is not installed in the build image; meson.build,
# goldrush_b200/meson.build and goldrush_path/meson.build describe the same targets for a reference
# checkout, and tests/test_host_logic.py keeps the two descriptions in step).
#
#   make            libgoldrush_b200.so (CUDA, sm_100a) + goldrush-path + goldpolish-index + grb-synth
#   make oracle     CPU checkers under oracle/ (test infrastructure)
#   make host-tools grb-synth only (no CUDA needed)
ROOT    := $(dir $(abspath $(lastword $(MAKEFILE_LIST))))
# The image exports CXX=/opt/gcc/bin/g++ (no libgomp.spec there), so do not inherit $$CXX.
GRB_CXX ?= $(firstword $(wildcard /usr/bin/g++) g++)
NVCC    ?= $(firstword $(wildcard /usr/local/cuda/bin/nvcc) nvcc)
CXXFLAGS_HOST := -std=c++17 -O2 -fopenmp -Wall -I$(ROOT)include
NVCCFLAGS := -std=c++17 -O3 -lineinfo -gencode arch=compute_100a,code=sm_100a \
             -Xcompiler -fPIC,-fopenmp,-Wall -I$(ROOT)include -I$(ROOT)goldrush_b200/csrc

LIBDIR := $(ROOT)goldrush_b200/_lib
LIB    := $(LIBDIR)/libgoldrush_b200.so
CU_SRC := $(wildcard $(ROOT)goldrush_b200/csrc/*.cu)
CU_HDR := $(wildcard $(ROOT)goldrush_b200/csrc/*.cuh) $(wildcard $(ROOT)goldrush_b200/csrc/*.h) $(ROOT)include/goldrush_b200.h
HOST_LIB_SRC := $(ROOT)goldrush_b200/host/synth.cpp $(ROOT)goldrush_b200/host/host_util.cpp $(ROOT)goldrush_b200/host/path_driver.cpp $(ROOT)goldrush_b200/host/decide_host.cpp $(ROOT)goldrush_b200/host/polish_driver.cpp $(ROOT)goldrush_b200/host/polish_inputs.cpp

all: lib goldrush-path goldpolish-index host-tools

lib: $(LIB)

$(LIB): $(CU_SRC) $(CU_HDR) $(HOST_LIB_SRC)
	@mkdir -p $(LIBDIR)
	$(NVCC) $(NVCCFLAGS) -ccbin $(GRB_CXX) -shared -o $@ $(CU_SRC) $(HOST_LIB_SRC) -lcudart -lgomp

goldrush-path: $(ROOT)build/goldrush-path

$(ROOT)build/goldrush-path: $(ROOT)goldrush_b200/host/goldrush_path_main.cpp $(ROOT)goldrush_b200/host/opt.cpp $(LIB)
	@mkdir -p $(ROOT)build
	$(GRB_CXX) $(CXXFLAGS_HOST) -o $@ $(ROOT)goldrush_b200/host/goldrush_path_main.cpp $(ROOT)goldrush_b200/host/opt.cpp \
	  -L$(LIBDIR) -lgoldrush_b200 -Wl,-rpath,'$$ORIGIN/../goldrush_b200/_lib'

goldpolish-index: $(ROOT)build/goldpolish-index

$(ROOT)build/goldpolish-index: $(ROOT)goldrush_b200/host/goldpolish_index_main.cpp $(LIB)
	@mkdir -p $(ROOT)build
	$(GRB_CXX) $(CXXFLAGS_HOST) -o $@ $(ROOT)goldrush_b200/host/goldpolish_index_main.cpp \
	  -L$(LIBDIR) -lgoldrush_b200 -Wl,-rpath,'$$ORIGIN/../goldrush_b200/_lib'

host-tools: $(ROOT)build/grb-synth

$(ROOT)build/grb-synth: $(ROOT)goldrush_b200/host/synth.cpp $(ROOT)goldrush_b200/host/grb_synth_main.cpp $(ROOT)include/goldrush_b200.h
	@mkdir -p $(ROOT)build
	$(GRB_CXX) $(CXXFLAGS_HOST) -o $@ $(ROOT)goldrush_b200/host/synth.cpp $(ROOT)goldrush_b200/host/grb_synth_main.cpp

tools: $(ROOT)build/sector-roofline

$(ROOT)build/sector-roofline: $(ROOT)tools/sector_roofline.cu
	@mkdir -p $(ROOT)build
	$(NVCC) -O3 -gencode arch=compute_100a,code=sm_100a -o $@ $<

oracle:
	$(MAKE) -C $(ROOT)oracle all

clean:
	rm -rf $(ROOT)build $(LIBDIR)

.PHONY: all lib goldrush-path goldpolish-index host-tools tools oracle clean
