# Builds libgpb200.so (sm_100a) in-tree.  `python -c "import __graft_entry__ as g; g.build()"` calls this.
NVCC ?= nvcc
ARCH := -gencode arch=compute_100a,code=sm_100a
NCCL_INC := $(shell python -c "import os, nvidia; print(os.path.join(list(nvidia.__path__)[0], 'nccl', 'include'))" 2>/dev/null)
NVFLAGS := $(ARCH) $(if $(NCCL_INC),-I$(NCCL_INC),) -O3 -lineinfo -std=c++17 -Xcompiler -fPIC -Xcompiler -Wall -Wno-deprecated-gpu-targets
SRC_DIR := inference_tools_b200/csrc
BUILD := build/obj
# api_test.cu (host-buffer hooks onto internal primitives for tests/ and tools/) is NOT part of the product library:
# it goes into libgpb200_test.so, which links against libgpb200.so.
SRCS := $(filter-out $(SRC_DIR)/api_test.cu,$(wildcard $(SRC_DIR)/*.cu))
OBJS := $(patsubst $(SRC_DIR)/%.cu,$(BUILD)/%.o,$(SRCS))
LIB := inference_tools_b200/libgpb200.so
TEST_LIB := inference_tools_b200/libgpb200_test.so

all: $(LIB) $(TEST_LIB)

$(BUILD)/%.o: $(SRC_DIR)/%.cu $(wildcard $(SRC_DIR)/*.cuh) include/gpb200.h
	@mkdir -p $(BUILD)
	$(NVCC) $(NVFLAGS) -c $< -o $@

$(LIB): $(OBJS)
	$(NVCC) $(ARCH) -shared -o $@ $(OBJS) -lcudart -ldl

$(TEST_LIB): $(BUILD)/api_test.o $(LIB)
	$(NVCC) $(ARCH) -shared -o $@ $(BUILD)/api_test.o -Linference_tools_b200 -lgpb200 -lcudart -Xlinker -rpath -Xlinker '$$ORIGIN'

clean:
	rm -rf build $(LIB) $(TEST_LIB)
.PHONY: all clean
