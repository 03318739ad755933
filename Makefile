# Top-level build: the sm_100a library, the five benchmark drivers (each into
# benchmarkNN/build/ like the reference's README recipe) and the test oracle.
# The CMake files next to each driver build the same targets.
HOSTCXX := $(shell [ -x /usr/bin/g++ ] && echo /usr/bin/g++ || echo g++)
CUDA    ?= /usr/local/cuda
CXXFLAGS := -O3 -std=c++17 -fopenmp -march=x86-64-v3 -Wall -I$(CUDA)/include
LIBDIR  := gpu-benchmarking_b200
LDFLAGS := -L$(LIBDIR) -lb200fe -L$(CUDA)/lib64 -lcudart -lcublas -lnccl -lpthread -Wl,-rpath,'$$ORIGIN/../../$(LIBDIR)' -Wl,-rpath,$(CUDA)/lib64
BENCH   := 01 02 03 04 05
DRIVERS := $(foreach b,$(BENCH),benchmark$(b)/build/benchmark$(b))
HDRS    := $(wildcard utils/*.h) include/b200fe.h

all: lib drivers oracle

lib:
	$(MAKE) -C $(LIBDIR)/csrc

drivers: lib $(DRIVERS)

define DRIVER_RULE
benchmark$(1)/build/benchmark$(1): benchmark$(1)/benchmark$(1).cc $$(HDRS)
	mkdir -p benchmark$(1)/build
	$$(HOSTCXX) $$(CXXFLAGS) $$< -o $$@ $$(LDFLAGS)
endef
$(foreach b,$(BENCH),$(eval $(call DRIVER_RULE,$(b))))

oracle:
	$(MAKE) -C oracle all

clean:
	$(MAKE) -C $(LIBDIR)/csrc clean
	rm -rf $(foreach b,$(BENCH),benchmark$(b)/build)

.PHONY: all lib drivers oracle clean
