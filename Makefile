# Build of the B200-native batched tensegrity step (sm_100a only) and of the CPU oracle used by the tests.
NVCC      ?= nvcc
CXX       ?= g++
ARCH      := -gencode arch=compute_100a,code=sm_100a
NVFLAGS   := $(ARCH) -lineinfo -O3 -std=c++17 -Xcompiler -fPIC
CSRC      := ntrtsim_b200/csrc
HOST      := ntrtsim_b200/host
LIBDIR    := ntrtsim_b200/lib
BUILD     := build

DEV_HDRS  := $(wildcard $(CSRC)/*.cuh) $(wildcard $(CSRC)/*.h) include/ntrtsim_b200.h
HOST_HDRS := $(wildcard $(HOST)/*.h) $(wildcard $(HOST)/LinearMath/*.h) include/ntrtsim_b200.h

all: $(LIBDIR)/libntrtsim_b200.so $(LIBDIR)/libtgb_host.so oracle/liboracle.so

$(BUILD) $(LIBDIR):
	mkdir -p $@

STEP_OBJS := $(foreach p,f32 f64,$(foreach c,simple sine cpg,$(BUILD)/tgb_step_$(p)_$(c).o))
TGB_EXTRA ?=

$(BUILD)/tgb_step_f32.o $(BUILD)/tgb_step_f32_simple.o $(BUILD)/tgb_step_f32_sine.o $(BUILD)/tgb_step_f32_cpg.o: $(BUILD)/%.o: $(CSRC)/%.cu $(DEV_HDRS) | $(BUILD)
	$(NVCC) $(NVFLAGS) $(TGB_EXTRA) -c $< -o $@
# the double instantiation evaluates the oracle's expression tree: no FMA contraction
$(BUILD)/tgb_step_f64.o $(BUILD)/tgb_step_f64_simple.o $(BUILD)/tgb_step_f64_sine.o $(BUILD)/tgb_step_f64_cpg.o: $(BUILD)/%.o: $(CSRC)/%.cu $(DEV_HDRS) | $(BUILD)
	$(NVCC) $(NVFLAGS) $(TGB_EXTRA) --fmad=false -c $< -o $@
$(BUILD)/tgb_capi.o: $(CSRC)/tgb_capi.cu $(DEV_HDRS) | $(BUILD)
	$(NVCC) $(NVFLAGS) -c $< -o $@

$(LIBDIR)/libntrtsim_b200.so: $(BUILD)/tgb_step_f32.o $(BUILD)/tgb_step_f64.o $(STEP_OBJS) $(BUILD)/tgb_capi.o | $(LIBDIR)
	$(NVCC) $(ARCH) -shared -o $@ $^

HOST_SRCS := $(HOST)/tgb_host.cpp $(HOST)/tgb_creator.cpp $(HOST)/tgb_sensors.cpp $(HOST)/tgb_yaml.cpp $(HOST)/tgb_examples.cpp $(HOST)/tgb_host_capi.cpp
$(LIBDIR)/libtgb_host.so: $(HOST_SRCS) $(HOST_HDRS) $(LIBDIR)/libntrtsim_b200.so | $(LIBDIR)
	$(CXX) -std=c++17 -O2 -fPIC -shared -o $@ $(HOST_SRCS) -L$(LIBDIR) -lntrtsim_b200 -ldl -Wl,-rpath,'$$ORIGIN'

# test infrastructure, never linked into the libraries above
oracle/liboracle.so: oracle/tg_oracle.cpp oracle/tg_oracle.h include/ntrtsim_b200.h
	$(CXX) -std=c++17 -O2 -ffp-contract=off -fPIC -shared -o $@ oracle/tg_oracle.cpp -lpthread

clean:
	rm -rf $(BUILD) $(LIBDIR)/*.so oracle/liboracle.so

.PHONY: all clean
