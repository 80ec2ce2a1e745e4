# Builds the product library stt_b200/libstt_b200.so (CUDA sm_100a + C++ host, no CPU path) and, via
# `make oracle`, the test-only oracle libraries.  __graft_entry__.build() drives both.
NVCC      ?= nvcc
CXX       ?= g++
ARCH      := -gencode arch=compute_100a,code=sm_100a
NVFLAGS   := $(ARCH) -lineinfo -O3 -std=c++17 -Xcompiler -fPIC,-fvisibility=hidden -Xptxas -v
CXXFLAGS  := -O2 -std=c++17 -fPIC -fvisibility=hidden -mfma -ffp-contract=off
CUDA_HOME ?= /usr/local/cuda
SRC       := stt_b200/csrc
OUT       := stt_b200/libstt_b200.so
DEVOUT    := stt_b200/libstt_b200_dev.so   # the same sources + the unit-test hooks (-DSTT_B200_DEV_HOOKS); never the product
BUILD     := build

HDRS := $(wildcard $(SRC)/*.h $(SRC)/*.cuh) include/stt_capi.h

.PHONY: all oracle clean variant
CLI       := stt_b200/stt
all: $(OUT) $(CLI) $(DEVOUT)

$(CLI): $(SRC)/client.cc include/stt_capi.h $(OUT)
	$(CXX) -O2 -std=c++17 -o $@ $(SRC)/client.cc -Lstt_b200 -lstt_b200 -Wl,-rpath,'$$ORIGIN' -L$(CUDA_HOME)/lib64 -Wl,-rpath,$(CUDA_HOME)/lib64

$(BUILD)/engine.o: $(SRC)/engine.cu $(HDRS)
	@mkdir -p $(BUILD)
	$(NVCC) $(NVFLAGS) -c $< -o $@ 2> $(BUILD)/ptxas_engine.log || (cat $(BUILD)/ptxas_engine.log; false)

$(BUILD)/%.o: $(SRC)/%.cc $(HDRS)
	@mkdir -p $(BUILD)
	$(CXX) $(CXXFLAGS) -I$(CUDA_HOME)/include -c $< -o $@

$(OUT): $(BUILD)/engine.o $(BUILD)/capi.o $(BUILD)/model_file.o $(BUILD)/tflite_reader.o $(BUILD)/scorer_image.o
	$(NVCC) $(ARCH) -shared -o $@ $^ -lcudart -Xlinker --exclude-libs,ALL

# A/B builds for measurements (tools/gpu_*.sh): `make variant V=name DEFS="-DSTT_X=1"` -> build/libstt_b200_name.so,
# selected at load time with STT_B200_LIB=<path> (stt_b200/api.py).  Never shipped: the product is $(OUT).
variant: $(BUILD)/capi.o $(BUILD)/model_file.o $(BUILD)/tflite_reader.o $(BUILD)/scorer_image.o
	$(NVCC) $(NVFLAGS) $(DEFS) -c $(SRC)/engine.cu -o $(BUILD)/engine_$(V).o 2> $(BUILD)/ptxas_engine_$(V).log || (cat $(BUILD)/ptxas_engine_$(V).log; false)
	$(NVCC) $(ARCH) -shared -o $(BUILD)/libstt_b200_$(V).so $(BUILD)/engine_$(V).o $^ -lcudart -Xlinker --exclude-libs,ALL

$(BUILD)/engine_dev.o: $(SRC)/engine.cu $(HDRS)
	@mkdir -p $(BUILD)
	$(NVCC) $(NVFLAGS) -DSTT_B200_DEV_HOOKS -c $< -o $@ 2> $(BUILD)/ptxas_engine_dev.log || (cat $(BUILD)/ptxas_engine_dev.log; false)

$(BUILD)/capi_dev.o: $(SRC)/capi.cc $(HDRS)
	@mkdir -p $(BUILD)
	$(CXX) $(CXXFLAGS) -DSTT_B200_DEV_HOOKS -I$(CUDA_HOME)/include -c $< -o $@

$(DEVOUT): $(BUILD)/engine_dev.o $(BUILD)/capi_dev.o $(BUILD)/model_file.o $(BUILD)/tflite_reader.o $(BUILD)/scorer_image.o
	$(NVCC) $(ARCH) -shared -o $@ $^ -lcudart -Xlinker --exclude-libs,ALL

oracle:
	$(MAKE) -C oracle
	if [ -d /root/reference ]; then $(MAKE) -C oracle -j8 ref; fi

clean:
	rm -rf $(BUILD) $(OUT) $(CLI) $(DEVOUT)
