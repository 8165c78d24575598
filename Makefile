# Top-level convenience Makefile (the Python entry point `__graft_entry__.build()` / `python -m ompmc_b200.build` does the same).
#   make            libompmc_b200.so (sm_100a, nvcc) + the plain-C host programs
#   make dropin     the reference's own user codes on the library (needs the reference checkout: make dropin REF=/path/to/ompMC)
#   make test       CPU test-suite;   make gputest   GPU test-suite (needs a B200)
NVCC   ?= nvcc
CC     ?= gcc
ARCH   := -gencode arch=compute_100a,code=sm_100a
NVFLAGS := -O3 -lineinfo -std=c++17 -Xcompiler -fPIC -Xcompiler -O2
PKG    := ompmc_b200
CSRC   := $(PKG)/csrc
HOST   := $(PKG)/host
OBJ    := $(PKG)/build
LIB    := $(PKG)/libompmc_b200.so
HDRS   := $(wildcard $(CSRC)/*.cuh) $(wildcard $(CSRC)/*.h) include/ompmc_b200.h
REF    ?= /root/reference

all: $(LIB) $(HOST)/omc_dosxyz_b200 $(HOST)/omc_matrad_b200

$(OBJ)/omc_lockstep.o: $(CSRC)/omc_lockstep.cu $(HDRS)
	@mkdir -p $(OBJ)
	$(NVCC) $(ARCH) $(NVFLAGS) -fmad=false -c $< -o $@    # parity kernel: no contraction of a*b+c (the reference's x86-64 build has none)
$(OBJ)/%.o: $(CSRC)/%.cu $(HDRS)
	@mkdir -p $(OBJ)
	$(NVCC) $(ARCH) $(NVFLAGS) -c $< -o $@
$(LIB): $(OBJ)/omc_lockstep.o $(OBJ)/omc_wavefront.o $(OBJ)/omc_capi.o
	$(NVCC) $(ARCH) -shared -o $@ $^

$(HOST)/omc_dosxyz_b200: $(HOST)/omc_dosxyz_b200.c $(HOST)/omc_tables.c $(HOST)/omc_tables.h $(HOST)/omc_host_common.h $(HOST)/omc_host_input.h $(LIB)
	$(CC) -O2 -Wall -ffp-contract=off -o $@ $(HOST)/omc_dosxyz_b200.c $(HOST)/omc_tables.c -Iinclude -I$(HOST) -L$(PKG) -lompmc_b200 -lm '-Wl,-rpath,$$ORIGIN/..'
$(HOST)/omc_matrad_b200: $(HOST)/omc_matrad_b200.c $(HOST)/omc_host_common.h $(LIB)
	$(CC) -O2 -Wall -o $@ $< -Iinclude -I$(HOST) -L$(PKG) -lompmc_b200 -lm '-Wl,-rpath,$$ORIGIN/..'

dropin: all
	$(MAKE) -C oracle dropin refdata REF=$(REF)

test:
	python -m pytest tests -q -m "not gpu"
gputest:
	python -m pytest tests -q -m gpu

clean:
	rm -rf $(OBJ) $(LIB) $(HOST)/omc_dosxyz_b200 $(HOST)/omc_matrad_b200 $(PKG)/libomc_format_host.so $(HOST)/libomc_tables.so
.PHONY: all dropin test gputest clean
