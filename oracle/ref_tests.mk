# Builds the REFERENCE's own legacy test executables (tests/cc/transpose_test.cc, tests/cc/halo_test.cc), unmodified,
# from the sources where they lie under /root/reference, against THIS repo's cudecomp.h, MPI shim and libcudecomp.so.
# The binaries contain the reference's own known-answer generator and comparator (transpose_test.cc:103-155), so a
# pass is the reference's verdict on this library. Outputs go to oracle/_ref/ only (git-ignored; travels to the GPU box).
#   make -f oracle/ref_tests.mk            (in the build container; /root/reference does not exist on the GPU box)
REF ?= /root/reference
ROOT := $(abspath $(dir $(lastword $(MAKEFILE_LIST)))/..)
OUT := $(ROOT)/oracle/_ref
NVCC ?= nvcc
FLAGS := -gencode arch=compute_100a,code=sm_100a -O2 -std=c++17 -x cu \
         -I$(ROOT)/include -I$(ROOT)/include/mpi_shim -I$(ROOT)/oracle/stubs -I$(REF)/include \
         -L$(ROOT)/cudecomp_b200/lib -lcudecomp -Xlinker -rpath -Xlinker '$$ORIGIN/../../cudecomp_b200/lib'

TYPES := R32 R64 C32 C64
BINS := $(foreach t,$(TYPES),$(OUT)/transpose_test_$(t) $(OUT)/halo_test_$(t))
# the reference's FFT benchmark (benchmark/benchmark.cu: cuFFT per pencil + the four transposes), also unmodified
BENCH := $(OUT)/benchmark_c2c $(OUT)/benchmark_r2c $(OUT)/benchmark_c2c_f $(OUT)/benchmark_r2c_f

all: $(BINS) $(BENCH)

$(OUT)/benchmark_c2c: $(REF)/benchmark/benchmark.cu $(ROOT)/cudecomp_b200/lib/libcudecomp.so
	@mkdir -p $(OUT)
	$(NVCC) $(FLAGS) -DC2C $< -lcufft -o $@
$(OUT)/benchmark_r2c: $(REF)/benchmark/benchmark.cu $(ROOT)/cudecomp_b200/lib/libcudecomp.so
	@mkdir -p $(OUT)
	$(NVCC) $(FLAGS) -DR2C $< -lcufft -o $@
$(OUT)/benchmark_c2c_f: $(REF)/benchmark/benchmark.cu $(ROOT)/cudecomp_b200/lib/libcudecomp.so
	@mkdir -p $(OUT)
	$(NVCC) $(FLAGS) -DC2C -DUSE_FLOAT $< -lcufft -o $@
$(OUT)/benchmark_r2c_f: $(REF)/benchmark/benchmark.cu $(ROOT)/cudecomp_b200/lib/libcudecomp.so
	@mkdir -p $(OUT)
	$(NVCC) $(FLAGS) -DR2C -DUSE_FLOAT $< -lcufft -o $@

$(OUT)/transpose_test_%: $(REF)/tests/cc/transpose_test.cc $(ROOT)/cudecomp_b200/lib/libcudecomp.so
	@mkdir -p $(OUT)
	$(NVCC) $(FLAGS) -D$* $< -o $@

$(OUT)/halo_test_%: $(REF)/tests/cc/halo_test.cc $(ROOT)/cudecomp_b200/lib/libcudecomp.so
	@mkdir -p $(OUT)
	$(NVCC) $(FLAGS) -D$* $< -o $@

clean:
	rm -rf $(OUT)
