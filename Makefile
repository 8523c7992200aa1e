# Builds libgpb200.so (sm_100a) in-tree.  `python -c "import __graft_entry__ as g; g.build()"` calls this.
NVCC ?= nvcc
ARCH := -gencode arch=compute_100a,code=sm_100a
NCCL_INC := $(shell python -c "import os, nvidia; print(os.path.join(list(nvidia.__path__)[0], 'nccl', 'include'))" 2>/dev/null)
NVFLAGS := $(ARCH) $(if $(NCCL_INC),-I$(NCCL_INC),) -O3 -lineinfo -std=c++17 -Xcompiler -fPIC -Xcompiler -Wall -Wno-deprecated-gpu-targets
SRC_DIR := inference_tools_b200/csrc
BUILD := build/obj
SRCS := $(wildcard $(SRC_DIR)/*.cu)
OBJS := $(patsubst $(SRC_DIR)/%.cu,$(BUILD)/%.o,$(SRCS))
LIB := inference_tools_b200/libgpb200.so

all: $(LIB)

$(BUILD)/%.o: $(SRC_DIR)/%.cu $(wildcard $(SRC_DIR)/*.cuh) include/gpb200.h
	@mkdir -p $(BUILD)
	$(NVCC) $(NVFLAGS) -c $< -o $@

$(LIB): $(OBJS)
	$(NVCC) $(ARCH) -shared -o $@ $(OBJS) -lcudart -ldl

clean:
	rm -rf build $(LIB)
.PHONY: all clean
