# libb200rl.so -- sm_100a only.  `python -c "import __graft_entry__ as g; g.build()"` runs the same recipe.
NVCC      ?= nvcc
NVCCFLAGS ?= -gencode arch=compute_100a,code=sm_100a -lineinfo -O3 -std=c++17 -Xcompiler -fPIC -Iinclude
SRC       := $(wildcard rl_algo_impls_b200/csrc/*.cu)
OBJ       := $(patsubst rl_algo_impls_b200/csrc/%.cu,build/%.o,$(SRC))
LIB       := rl_algo_impls_b200/libb200rl.so

all: $(LIB)

build/%.o: rl_algo_impls_b200/csrc/%.cu $(wildcard rl_algo_impls_b200/csrc/*.cuh) include/b200rl.h
	@mkdir -p build
	$(NVCC) $(NVCCFLAGS) -c $< -o $@

$(LIB): $(OBJ)
	$(NVCC) -gencode arch=compute_100a,code=sm_100a -shared -o $@ $(OBJ) -lcudart

clean:
	rm -rf build $(LIB)
.PHONY: all clean
