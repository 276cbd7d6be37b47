# Build of the B200 path-tracing core, the C++ host library, the offlinerender driver and the CPU oracle.
#   make            -> product libraries + driver
#   make oracle     -> oracle/_build/liboracle.so (test infrastructure)
# All artefacts are git-ignored but travel to the GPU box with the gpurun snapshot.

NVCC      ?= /usr/local/cuda/bin/nvcc
CXX       := $(shell test -x /usr/bin/g++ && echo /usr/bin/g++ || echo g++)
LIBDIR    := vviewer_b200/_lib
ORCDIR    := oracle/_build

NVCCFLAGS := -gencode arch=compute_100a,code=sm_100a -lineinfo -O3 -std=c++17 \
             -Xcompiler -fPIC,-fvisibility=hidden,-O3 -Xptxas -v -Iinclude
CXXFLAGS  := -O2 -std=c++17 -fPIC -fvisibility=hidden -Wall -Wno-unused-variable -Wno-unused-function -Iinclude

CUDA_SRCS := $(wildcard vviewer_b200/csrc/*.cu)
CUDA_HDRS := $(wildcard vviewer_b200/csrc/*.cuh) include/ptc.h
HOST_SRCS := vviewer_b200/host/vengine.cpp vviewer_b200/host/io_image.cpp vviewer_b200/host/io_obj.cpp \
             vviewer_b200/host/scenes.cpp vviewer_b200/host/capi.cpp \
             vviewer_b200/host/io_jpeg.cpp vviewer_b200/host/io_gltf.cpp vviewer_b200/host/io_scene.cpp
HOST_HDRS := $(wildcard vviewer_b200/host/*.hpp) include/ptc.h include/vengine_host.h

all: $(LIBDIR)/libptc_cuda.so $(LIBDIR)/libvengine_host.so $(LIBDIR)/offlinerender

$(LIBDIR)/libptc_cuda.so: $(CUDA_SRCS) $(CUDA_HDRS)
	@mkdir -p $(LIBDIR)
	$(NVCC) $(NVCCFLAGS) -shared -o $@ vviewer_b200/csrc/ptc_cuda.cu -lcudart -ldl

$(LIBDIR)/libvengine_host.so: $(HOST_SRCS) $(HOST_HDRS)
	@mkdir -p $(LIBDIR)
	$(CXX) $(CXXFLAGS) -shared -o $@ $(HOST_SRCS) -lz -ldl

$(LIBDIR)/offlinerender: vviewer_b200/bin/offlinerender/main.cpp $(LIBDIR)/libvengine_host.so
	$(CXX) $(CXXFLAGS) -o $@ vviewer_b200/bin/offlinerender/main.cpp $(HOST_SRCS) -lz -ldl

oracle: $(ORCDIR)/liboracle.so

# diagnosis build with traversal counters (tools/trav_stats.py); never loaded by the product, tests or bench
stats: $(LIBDIR)/libptc_cuda_stats.so
$(LIBDIR)/libptc_cuda_stats.so: $(CUDA_SRCS) $(CUDA_HDRS)
	@mkdir -p $(LIBDIR)
	$(NVCC) $(NVCCFLAGS) -DPTC_TRAV_STATS -shared -o $@ vviewer_b200/csrc/ptc_cuda.cu -lcudart -ldl

# -ffp-contract=off: the world-space flatten and the LBVH reference build must round exactly like the
# device kernels, which use explicit __fmul_rn/__fadd_rn
$(ORCDIR)/liboracle.so: oracle/oracle.cpp oracle/omath.hpp oracle/bsdf.hpp oracle/accel.hpp oracle/envdist.hpp oracle/srgb_table.h include/ptc.h
	@mkdir -p $(ORCDIR)
	$(CXX) -O3 -std=c++17 -fPIC -fvisibility=hidden -ffp-contract=off -fopenmp -Wall -Iinclude -shared -o $@ oracle/oracle.cpp

# oracle/_ref: the only part of the reference that compiles here is its vendored image decoder (stb_image.h); it is built
# from the sources where they lie under /root/reference (nothing is copied) and used by tests/golden/make_image_fixtures.py
REFSTB := /root/reference/src/lib/external/stb
oracle-ref:
	@if [ -f $(REFSTB)/stb_image.h ]; then mkdir -p oracle/_ref && \
	  gcc -O2 -w -I$(REFSTB) -o oracle/_ref/stb_decode oracle/ref_tools/stb_decode.c -lm && echo built oracle/_ref/stb_decode; \
	else echo "/root/reference absent: oracle/_ref not rebuilt"; fi

clean:
	rm -rf $(LIBDIR) $(ORCDIR) oracle/_ref

.PHONY: all oracle oracle-ref stats clean
