# Top-level build.  Everything is built IN-TREE so the .so files travel to the GPU box.
#
#   make cuda     convolutionalencdec_b200/libced_cuda.so      sm_100a kernels + extern "C" ABI
#   make host     convolutionalencdec_b200/libconvencdec_{k7,k3}.so   the C drop-in library
#   make drivers  drivers/_bin/*   the reference's UNCHANGED driver sources ($(CED_REF)/*/X.c)
#                 compiled against include/ and linked to the drop-in library
#   make oracle   oracle/libced_oracle.so, oracle/_ref/*   (test infrastructure)
#   make hostsim  tests/hostsim/libswar_sim.so             (test infrastructure)
CED_REF ?= /root/reference
NVCC ?= nvcc
CC ?= gcc
PKG := convolutionalencdec_b200
CSRC := $(PKG)/csrc
NVFLAGS := -gencode arch=compute_100a,code=sm_100a -lineinfo -O3 -std=c++17 -Xcompiler -fPIC
CFLAGS := -O2 -g -std=gnu11 -fPIC -Wall -Iinclude
HOST_SRCS := $(CSRC)/host/convEncode.c $(CSRC)/host/convHelpers.c $(CSRC)/host/viterbiDecoder.c $(CSRC)/host/ced_introspect.c \
             $(CSRC)/host/viterbiDecoderQueue.c
CUDA_HDRS := $(wildcard $(CSRC)/*.cuh) include/ced_abi.h

all: cuda host oracle hostsim drivers examples

# one object per translation unit (build/ is git-ignored), so `make -j` compiles them side by side
CUDA_SRCS := $(wildcard $(CSRC)/*.cu)
CUDA_OBJS := $(patsubst $(CSRC)/%.cu,build/%.o,$(CUDA_SRCS)) build/host_pack.o

cuda: $(PKG)/libced_cuda.so
build/%.o: $(CSRC)/%.cu $(CUDA_HDRS)
	@mkdir -p build
	$(NVCC) $(NVFLAGS) -Xcompiler -pthread -c -o $@ $<
build/host_pack.o: $(CSRC)/host_pack.cpp
	@mkdir -p build
	$(NVCC) $(NVFLAGS) -Xcompiler -pthread -c -o $@ $<
$(PKG)/libced_cuda.so: $(CUDA_OBJS)
	$(NVCC) $(NVFLAGS) -Xcompiler -pthread -shared -o $@ $(CUDA_OBJS) -ldl
	python tools/sass_loop_stats.py   # instruction counts of the hot loops of THIS binary (bench.py reads them)

host: $(PKG)/libconvencdec_k7.so $(PKG)/libconvencdec_k3.so
$(PKG)/libconvencdec_k7.so: $(HOST_SRCS) $(CSRC)/host/params/default/convCodeParams.c $(PKG)/libced_cuda.so $(wildcard include/*.h)
	$(CC) $(CFLAGS) -Iinclude/params/default -shared -o $@ $(HOST_SRCS) $(CSRC)/host/params/default/convCodeParams.c \
	    -L$(PKG) -lced_cuda -Wl,-rpath,'$$ORIGIN' -Wl,-Bsymbolic
$(PKG)/libconvencdec_k3.so: $(HOST_SRCS) $(CSRC)/host/params/handTraced/convCodeParams.c $(PKG)/libced_cuda.so $(wildcard include/*.h)
	$(CC) $(CFLAGS) -Iinclude/params/handTraced -shared -o $@ $(HOST_SRCS) $(CSRC)/host/params/handTraced/convCodeParams.c \
	    -L$(PKG) -lced_cuda -Wl,-rpath,'$$ORIGIN' -Wl,-Bsymbolic

examples: examples/_bin/batch_roundtrip examples/_bin/speed_queued examples/_bin/multi_gpu_roundtrip examples/_bin/soft_decisions
examples/_bin/soft_decisions: examples/soft_decisions.c include/ced_abi.h $(PKG)/libced_cuda.so
	mkdir -p examples/_bin
	$(CC) -O2 -g -std=gnu11 -Wall -Iinclude -o $@ $< -L$(PKG) -lced_cuda -lm -Wl,-rpath,'$$ORIGIN/../../$(PKG)'
examples/_bin/multi_gpu_roundtrip: examples/multi_gpu_roundtrip.c include/ced_abi.h $(PKG)/libced_cuda.so
	mkdir -p examples/_bin
	$(CC) -O2 -g -std=gnu11 -Wall -Iinclude -o $@ $< -L$(PKG) -lced_cuda -Wl,-rpath,'$$ORIGIN/../../$(PKG)'
examples/_bin/speed_queued: examples/speed_queued.c $(wildcard include/*.h) $(PKG)/libconvencdec_k7.so
	mkdir -p examples/_bin
	$(CC) -O2 -g -std=gnu11 -Wall -Iinclude/params/default -Iinclude -o $@ $< -L$(PKG) -lconvencdec_k7 -lced_cuda -pthread \
	    -Wl,-rpath,'$$ORIGIN/../../$(PKG)'
examples/_bin/batch_roundtrip: examples/batch_roundtrip.c include/ced_abi.h $(PKG)/libced_cuda.so
	mkdir -p examples/_bin
	$(CC) -O2 -g -std=gnu11 -Wall -Iinclude -o $@ $< -L$(PKG) -lced_cuda -Wl,-rpath,'$$ORIGIN/../../$(PKG)'

oracle:
	$(MAKE) -C oracle CED_REF=$(CED_REF) all

hostsim: tests/hostsim/libswar_sim.so
tests/hostsim/libswar_sim.so: tests/hostsim/swar_sim.cpp $(CSRC)/trellis_swar.cuh $(CSRC)/trellis_swar16.cuh $(CSRC)/trellis_fused.cuh $(CSRC)/swar_generic.cuh $(CSRC)/swar_radix4.cuh
	g++ -O2 -std=c++17 -Wno-unknown-pragmas -fPIC -shared -x c++ -I$(CSRC) -o $@ $<

# Reference drivers, sources untouched.  speedDecode/speedEncode pin their worker
# to CPU 16 (speedDecode.c:23,147); drivers/affinity_wrap.c makes that call
# succeed on boxes with fewer CPUs (link-time --wrap, no source change).
DRV := drivers/_bin
DRV_FLAGS := -O2 -g -std=gnu11 -Iinclude -w
ifneq ($(wildcard $(CED_REF)/speedDecode/speedDecode.c),)
drivers: $(DRV)/handTraced $(DRV)/berTestK7 $(DRV)/speedDecode $(DRV)/speedEncode
$(DRV)/handTraced: $(CED_REF)/handTracedTest/handTraced.c $(PKG)/libconvencdec_k3.so | $(DRV)
	$(CC) $(DRV_FLAGS) -Iinclude/params/handTraced -o $@ $< -L$(PKG) -lconvencdec_k3 -lced_cuda -Wl,-rpath,'$$ORIGIN/../../$(PKG)'
$(DRV)/berTestK7: $(CED_REF)/berTestK7/berTestK7.c $(PKG)/libconvencdec_k7.so | $(DRV)
	$(CC) $(DRV_FLAGS) -Iinclude/params/default -o $@ $< -L$(PKG) -lconvencdec_k7 -lced_cuda -lm -Wl,-rpath,'$$ORIGIN/../../$(PKG)'
$(DRV)/speedDecode: $(CED_REF)/speedDecode/speedDecode.c drivers/affinity_wrap.c $(PKG)/libconvencdec_k7.so | $(DRV)
	$(CC) $(DRV_FLAGS) -Iinclude/params/default -o $@ $< drivers/affinity_wrap.c -Wl,--wrap=pthread_attr_setaffinity_np \
	    -L$(PKG) -lconvencdec_k7 -lced_cuda -pthread -lm -Wl,-rpath,'$$ORIGIN/../../$(PKG)'
$(DRV)/speedEncode: $(CED_REF)/speedEncode/speedEncode.c drivers/affinity_wrap.c $(PKG)/libconvencdec_k7.so | $(DRV)
	$(CC) $(DRV_FLAGS) -Iinclude/params/default -o $@ $< drivers/affinity_wrap.c -Wl,--wrap=pthread_attr_setaffinity_np \
	    -L$(PKG) -lconvencdec_k7 -lced_cuda -pthread -lm -Wl,-rpath,'$$ORIGIN/../../$(PKG)'
$(DRV):
	mkdir -p $@
else
drivers:
	@echo "drivers: $(CED_REF) not present; using prebuilt drivers/_bin if any"
endif

# measurement probes (not part of the library): tools/_bin/{write_probe,latency_probe}
tools: tools/_bin/write_probe tools/_bin/latency_probe
tools/_bin/%: tools/%.cu
	mkdir -p tools/_bin
	$(NVCC) -gencode arch=compute_100a,code=sm_100a -lineinfo -O3 -std=c++17 -o $@ $<

clean:
	rm -f $(PKG)/*.so tests/hostsim/*.so build/*.o
	rm -rf $(DRV) examples/_bin
	$(MAKE) -C oracle clean

.PHONY: all cuda host oracle hostsim drivers examples tools clean
