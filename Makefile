# Builds everything in-tree (the .so files travel to the GPU box with the
# snapshot; they are git-ignored).
#   liblighthugger_b200.so  the product: sm_100a CUDA kernels + C ABI
#   liblh_host.so           host-side scene preparation (C++20)
#   oracle/liblh_oracle.so  CPU restatement of the reference (test infrastructure)
#   oracle/_ref/...         the reference's own scene-prep code compiled in place (only where /root/reference exists)
NVCC ?= /usr/local/cuda/bin/nvcc
CXX ?= g++
REF ?= /root/reference

LIBDIR := lighthugger_b200/lib
CUDA_SRCS := $(wildcard lighthugger_b200/csrc/*.cu)
CUDA_HDRS := $(wildcard lighthugger_b200/csrc/*.cuh lighthugger_b200/csrc/*.h include/*.h)
HOST_SRCS := $(wildcard lighthugger_b200/host/*.cpp)
HOST_HDRS := $(wildcard lighthugger_b200/host/*.hpp include/*.h)
ORACLE_SRCS := $(wildcard oracle/*.cpp)
ORACLE_HDRS := $(wildcard oracle/*.h oracle/*.hpp include/*.h)

# the pinned arithmetic (no FMA contraction, IEEE div/sqrt) is written with explicit _rn intrinsics (lh_internal.cuh)
NVCCFLAGS := -gencode arch=compute_100a,code=sm_100a -O3 -std=c++20 -lineinfo \
             -Xcompiler -fPIC,-fvisibility=hidden -Iinclude -Ilighthugger_b200/csrc
LDFLAGS_SO := -Wl,--exclude-libs,ALL
CXXFLAGS := -O2 -std=c++20 -fPIC -fvisibility=hidden -ffp-contract=off -Iinclude -Wall -Wextra -Wno-unused-parameter

all: cuda host oracle ref examples probes

cuda: $(LIBDIR)/liblighthugger_b200.so
host: $(LIBDIR)/liblh_host.so
oracle: oracle/liblh_oracle.so
examples: lighthugger_b200/bin/lh_frame_loop
# hardware probes (texture-unit filter arithmetic, L2 atomic ceilings): run on the GPU box, fitted offline (tools/*.py)
PROBES := build/texfloat_probe build/texwrap_probe build/texvalue_probe build/texsrgb_probe build/texfilter_probe build/atomic_probe
probes: $(PROBES)
build/%_probe: tools/%_probe.cu
	@mkdir -p build
	$(NVCC) -gencode arch=compute_100a,code=sm_100a -O3 -o $@ $<

CUDA_OBJS := $(patsubst lighthugger_b200/csrc/%.cu,build/%.o,$(CUDA_SRCS))
# the lighting resolve keeps the reference shader's operation order with IEEE operations only (see resolve.cu)
build/resolve.o: NVCCFLAGS += -fmad=false
build/%.o: lighthugger_b200/csrc/%.cu $(CUDA_HDRS)
	@mkdir -p build
	$(NVCC) $(NVCCFLAGS) -c -o $@ $<

$(LIBDIR)/liblighthugger_b200.so: $(CUDA_OBJS)
	@mkdir -p $(LIBDIR)
	$(NVCC) -shared -o $@ $(CUDA_OBJS) -lcudart_static -ldl -lrt -lpthread

# Scene preparation meshletises with the reference's vendored meshoptimizer (north star; src/resources/meshlets.cpp),
# compiled from where the reference keeps it; without /root/reference the built-in greedy builder is the fallback.
REF_MESHOPT_SRCS := $(wildcard $(REF)/external/meshoptimizer/src/clusterizer.cpp $(REF)/external/meshoptimizer/src/allocator.cpp)
ifneq ($(words $(REF_MESHOPT_SRCS)),2)
REF_MESHOPT_SRCS :=
HOST_MESHOPT_FLAGS :=
else
HOST_MESHOPT_FLAGS := -DLH_HAVE_MESHOPTIMIZER -I$(REF)/external/meshoptimizer/src
endif
# Where the reference is absent (the GPU box) an existing library is KEPT whatever the timestamps say: it was built here
# with the reference's meshoptimizer, and rebuilding it there would silently swap the meshlet builder (and with it the
# benchmark scene) for the greedy fallback.
ifeq ($(REF_MESHOPT_SRCS)$(wildcard $(LIBDIR)/liblh_host.so),$(LIBDIR)/liblh_host.so)
$(LIBDIR)/liblh_host.so:
	@true
else
$(LIBDIR)/liblh_host.so: $(HOST_SRCS) $(HOST_HDRS)
	@mkdir -p $(LIBDIR)
	$(CXX) $(CXXFLAGS) $(HOST_MESHOPT_FLAGS) -shared -o $@ $(HOST_SRCS) $(REF_MESHOPT_SRCS) $(LDFLAGS_SO) -ldl
endif

# a C++20 host on top of the two C ABIs, the counterpart of the reference's main() frame loop
lighthugger_b200/bin/lh_frame_loop: examples/frame_loop.cpp $(LIBDIR)/liblighthugger_b200.so $(LIBDIR)/liblh_host.so $(HOST_HDRS)
	@mkdir -p lighthugger_b200/bin
	$(CXX) -O2 -std=c++20 -Iinclude -Wall -Wextra -o $@ examples/frame_loop.cpp -L$(LIBDIR) -llighthugger_b200 -llh_host -Wl,-rpath,'$$ORIGIN/../lib'

oracle/liblh_oracle.so: $(ORACLE_SRCS) $(ORACLE_HDRS)
	$(CXX) $(CXXFLAGS) -fvisibility=default -pthread -shared -o $@ $(ORACLE_SRCS) $(LDFLAGS_SO)

# The reference's scene-prep path compiles from its own few sources; its full
# build (Vulkan, fastgltf download) does not -- see DESIGN.md.
REF_GLM := $(REF)/external/VulkanMemoryAllocator-Hpp/Vulkan-Hpp/glm
REF_MESHOPT := $(REF)/external/meshoptimizer/src
ref:
	@if [ -d $(REF)/src ]; then $(MAKE) --no-print-directory oracle/_ref/libref_sceneprep.so oracle/_ref/libref_shaders.so oracle/_ref/tony_mc_mapface.dds; \
	 else echo "no $(REF): keeping prebuilt oracle/_ref"; fi

oracle/_ref/libref_sceneprep.so: oracle/ref_shim/ref_exports.cpp oracle/ref_shim/ref_pch.h
	@mkdir -p oracle/_ref
	$(CXX) -O2 -std=c++20 -fPIC -fvisibility=hidden -w -shared -o $@ \
	    -include oracle/ref_shim/ref_pch.h -Ioracle/ref_shim -I$(REF)/src -I$(REF_GLM) -I$(REF_MESHOPT) \
	    oracle/ref_shim/ref_exports.cpp $(REF_MESHOPT)/clusterizer.cpp $(REF_MESHOPT)/allocator.cpp $(LDFLAGS_SO)

# The reference's GLSL (every compute entry point of the path, the vertex shaders and the function files they include)
# compiled as C++ against its own glm and its own shared_cpu_gpu.h: glsl_filter.py drops what is not C++ (layout
# declarations, #extension) into oracle/_ref/gen/, ref_shaders.cpp includes the result.  This pins the oracle.
REF_SHADERS := common/prefix_sum.glsl common/util.glsl common/culling.glsl common/matrices.glsl common/vbuffer.glsl \
               common/pbr.glsl common/debug.glsl cull_instances.comp write_draw_calls.comp rasterization.glsl \
               render_geometry.comp display_transform.comp compute/reset_buffers.glsl compute/generate_shadow_matrices.glsl \
               compute/read_depth.glsl
REF_SHADERS_GEN := $(addprefix oracle/_ref/gen/,$(REF_SHADERS)) oracle/_ref/gen/shared_cpu_gpu.h
oracle/_ref/gen/shared_cpu_gpu.h: $(REF)/src/shared_cpu_gpu.h oracle/ref_shim/glsl_filter.py
	@mkdir -p $(dir $@)
	python oracle/ref_shim/glsl_filter.py $< $@
oracle/_ref/gen/%: $(REF)/src/shaders/% oracle/ref_shim/glsl_filter.py
	@mkdir -p $(dir $@)
	python oracle/ref_shim/glsl_filter.py $< $@

oracle/_ref/libref_shaders.so: oracle/ref_shim/ref_shaders.cpp oracle/ref_shim/glsl_shim.h $(REF_SHADERS_GEN) oracle/liblh_oracle.so
	$(CXX) -O2 -std=gnu++20 -fPIC -fvisibility=hidden -ffp-contract=off -w -shared -o $@ \
	    -Ioracle/ref_shim -Ioracle/_ref -Ioracle -Iinclude -I$(REF)/src -I$(REF_GLM) oracle/ref_shim/ref_shaders.cpp \
	    -Loracle -llh_oracle -Wl,-rpath,'$$ORIGIN/..' $(LDFLAGS_SO)

# the reference's display LUT as a build output (git-ignored; travels to the GPU box with the prebuilt libraries): the
# GPU test of the display transform runs the REAL table through lh_set_display_lut
oracle/_ref/tony_mc_mapface.dds: $(REF)/external/tony-mc-mapface/shader/tony_mc_mapface.dds
	@mkdir -p oracle/_ref
	cp $< $@

clean:
	rm -f $(LIBDIR)/*.so oracle/*.so oracle/_ref/*.so lighthugger_b200/bin/lh_frame_loop

.PHONY: all cuda host oracle ref examples clean
