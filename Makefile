# Builds the product library (libav1r.so: host parser + sm_100a CUDA kernels + C ABI) and the
# test-only oracle library (oracle/_build/liboracle.so).  nvcc cross-compiles without a GPU.
NVCC ?= /usr/local/cuda/bin/nvcc
CXX ?= g++
CC ?= gcc
ARCH := -gencode arch=compute_100a,code=sm_100a
CSRC := av1-go_b200/csrc
LIBDIR := av1-go_b200/lib
OBJDIR := build/obj
CXXFLAGS := -O2 -g -std=c++17 -fPIC -Wall -Wno-unused-function -Iinclude
NVFLAGS := $(ARCH) -O3 -lineinfo -std=c++17 -Xcompiler -fPIC -Iinclude --expt-relaxed-constexpr $(EXTRA_NVFLAGS)

HOST_SRCS := $(wildcard $(CSRC)/*.cpp)
CU_SRCS := $(wildcard $(CSRC)/kernels/*.cu) $(wildcard $(CSRC)/*.cu)
HOST_OBJS := $(patsubst $(CSRC)/%.cpp,$(OBJDIR)/%.o,$(HOST_SRCS))
CU_OBJS := $(patsubst $(CSRC)/%.cu,$(OBJDIR)/%.cu.o,$(CU_SRCS))

ORACLE_C := $(wildcard oracle/*.c)
ORACLE_CPP := $(wildcard oracle/*.cpp)

all: $(LIBDIR)/libav1r.so oracle/_build/liboracle.so

$(OBJDIR)/%.o: $(CSRC)/%.cpp $(wildcard $(CSRC)/*.h) $(wildcard include/*.h)
	@mkdir -p $(dir $@)
	$(CXX) $(CXXFLAGS) -I/usr/local/cuda/include -c $< -o $@

$(OBJDIR)/%.cu.o: $(CSRC)/%.cu $(wildcard $(CSRC)/*.h) $(wildcard $(CSRC)/kernels/*.cuh) $(wildcard include/*.h)
	@mkdir -p $(dir $@)
	$(NVCC) $(NVFLAGS) -Xptxas -v -c $< -o $@ 2> $@.ptxas.log || (cat $@.ptxas.log; false)

$(LIBDIR)/libav1r.so: $(HOST_OBJS) $(CU_OBJS)
	@mkdir -p $(LIBDIR)
	$(NVCC) $(ARCH) -shared -o $@ $^ -lcudart -lpthread

# the oracle links the product's *host parser* objects (never the other way round)
PARSER_OBJS := $(OBJDIR)/obu.o $(OBJDIR)/tile.o $(OBJDIR)/tile_inter.o $(OBJDIR)/stream_parser.o
oracle/_build/liboracle.so: $(ORACLE_C) $(ORACLE_CPP) $(PARSER_OBJS) $(wildcard oracle/*.h) $(wildcard $(CSRC)/kernels/*.h) $(wildcard $(CSRC)/tables/*.inc)
	@mkdir -p oracle/_build
	$(CC) -O2 -g -fPIC -c oracle/filmgrain.c -o oracle/_build/filmgrain.o
	$(CXX) -O2 -g -std=c++17 -fPIC -shared -Iinclude $(ORACLE_CPP) oracle/_build/filmgrain.o $(PARSER_OBJS) -o $@ -ldl

# sanitizer build of the host half (demux + parser + C entry points that need no GPU), driven by tests/test_robustness.py
ASAN_CXX ?= /usr/bin/g++
ASAN_SRCS := $(filter-out $(CSRC)/batch.cpp,$(HOST_SRCS)) tests/native/parse_fuzz_main.cpp
build/asan/parse_fuzz: $(ASAN_SRCS) $(wildcard $(CSRC)/*.h) $(wildcard include/*.h)
	@mkdir -p build/asan
	$(ASAN_CXX) -O1 -g -std=c++17 -fsanitize=address,undefined -fno-sanitize=shift-base -fno-sanitize-recover=undefined -fno-omit-frame-pointer -Iinclude -I/usr/local/cuda/include $(ASAN_SRCS) -o $@ -lpthread

clean:
	rm -rf build $(LIBDIR) oracle/_build

.PHONY: all clean asan
asan: build/asan/parse_fuzz
