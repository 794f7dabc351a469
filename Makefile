# Builds libscanet_b200.so (the C-ABI engine) for sm_100a.  `make -j` ; nvcc cross-compiles without a GPU.
NVCC      ?= /usr/local/cuda/bin/nvcc
ARCH      := -gencode arch=compute_100a,code=sm_100a
NVFLAGS   := -O3 -std=c++17 -lineinfo $(ARCH) -Xcompiler -fPIC,-Wall,-Wno-unused-function -Xptxas -v --expt-relaxed-constexpr
CSRC      := scanet3_b200/csrc
BUILD     := build
OBJS      := $(BUILD)/engine.o $(BUILD)/kernels_simt.o $(BUILD)/kernels_fast.o $(BUILD)/bn_pool.o $(BUILD)/tc.o $(BUILD)/tc_gemm.o $(BUILD)/lstm.o $(BUILD)/lstm_persist.o $(BUILD)/rnn.o $(BUILD)/group.o $(BUILD)/host_init.o $(BUILD)/wire.o
LIB       := scanet3_b200/libscanet_b200.so

all: $(LIB)

$(BUILD)/%.o: $(CSRC)/%.cu $(wildcard $(CSRC)/*.h $(CSRC)/*.cuh) include/scanet_b200.h
	@mkdir -p $(BUILD)
	$(NVCC) $(NVFLAGS) -c $< -o $@ 2> $(BUILD)/$*.ptxas.log || (cat $(BUILD)/$*.ptxas.log; exit 1)

$(BUILD)/%.o: $(CSRC)/%.cpp $(wildcard $(CSRC)/*.h) include/scanet_b200.h
	@mkdir -p $(BUILD)
	g++ -O2 -std=c++17 -fPIC -Wall -c $< -o $@

$(LIB): $(OBJS)
	$(NVCC) $(ARCH) -shared -o $@ $(OBJS) -cudart static -ldl -lpthread

# developer variant with the in-situ timeline stamps compiled in (common.cuh SB_TRACE_BUILD): tools/trace_step.py loads it through SB_LIB
TBUILD    := build/trace
TOBJS     := $(patsubst $(BUILD)/%,$(TBUILD)/%,$(OBJS))
TLIB      := scanet3_b200/libscanet_b200_trace.so
trace: $(TLIB)
$(TBUILD)/%.o: $(CSRC)/%.cu $(wildcard $(CSRC)/*.h $(CSRC)/*.cuh) include/scanet_b200.h
	@mkdir -p $(TBUILD)
	$(NVCC) $(NVFLAGS) -DSB_TRACE_BUILD -c $< -o $@ 2> $(TBUILD)/$*.ptxas.log || (cat $(TBUILD)/$*.ptxas.log; exit 1)
$(TBUILD)/%.o: $(CSRC)/%.cpp $(wildcard $(CSRC)/*.h) include/scanet_b200.h
	@mkdir -p $(TBUILD)
	g++ -O2 -std=c++17 -fPIC -Wall -c $< -o $@
$(TLIB): $(TOBJS)
	$(NVCC) $(ARCH) -shared -o $@ $(TOBJS) -cudart static -ldl -lpthread

clean:
	rm -rf $(BUILD) $(LIB) $(TLIB)
.PHONY: all clean trace
