# Builds the C-ABI shared library of the sm_100a MSCKF kernels in-tree (it travels to the GPU box with the snapshot).
NVCC ?= nvcc
CSRC := monocular_visual_inertial_msckf_b200/csrc
SRCS := $(CSRC)/front.cu $(CSRC)/front_split.cu $(CSRC)/front_chol_smem.cu $(CSRC)/tsqr.cu $(CSRC)/gram.cu $(CSRC)/update_info.cu $(CSRC)/update_smem.cu $(CSRC)/apply.cu $(CSRC)/propagate.cu $(CSRC)/window.cu $(CSRC)/tracks.cu $(CSRC)/match.cu $(CSRC)/synth.cu $(CSRC)/descriptors.cu $(CSRC)/probe.cu $(CSRC)/api.cu
OBJS := $(SRCS:.cu=.o)
LIB := $(CSRC)/libmsckf_b200.so
NVFLAGS := -gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -std=c++17 -Xcompiler -fPIC -Xptxas -v

all: $(LIB)

%.o: %.cu $(CSRC)/msckf_common.cuh $(CSRC)/dense.cuh $(CSRC)/front_geom.cuh $(CSRC)/front_tiles.cuh $(CSRC)/front_gate.cuh include/msckf_b200.h
	$(NVCC) $(NVFLAGS) -c $< -o $@ 2> $@.ptxas.log || (cat $@.ptxas.log; exit 1)

$(LIB): $(OBJS)
	$(NVCC) -shared -gencode arch=compute_100a,code=sm_100a -o $@ $(OBJS)

clean:
	rm -f $(OBJS) $(LIB) $(CSRC)/*.ptxas.log
