// Ceiling of the rasteriser's sample writes ("Raster: L2 atomic throughput ... from a u64-atomicMin
// microbenchmark over a 66 MB (4K, L2-resident) and 265 MB (8K, HBM-resident) footprint on the same box").
// Every thread issues fire-and-forget 64-bit min reductions to pseudo-random pixels of a key buffer, as
// Target::emit does.  build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/atomic_probe tools/atomic_probe.cu
// run:   tools/atomic_probe            (prints G reductions/s for both footprints, scattered and 8x4-block-local)
#include <cstdio>
#include <cuda_runtime.h>

__global__ void probe(unsigned long long* keys, unsigned long long pixels, int per_thread, int local) {
    unsigned long long x = (blockIdx.x * (unsigned long long)blockDim.x + threadIdx.x) * 0x9E3779B97F4A7C15ull + 1;
    const unsigned lane = threadIdx.x & 31;
    for (int i = 0; i < per_thread; i++) {
        x ^= x << 13; x ^= x >> 7; x ^= x << 17;
        unsigned long long at = x % pixels;
        if (local) {   // the 32 lanes of a warp hit one 8x4 pixel block of a 3840-wide image (what a meshlet's samples look like)
            const unsigned long long base = __shfl_sync(0xFFFFFFFFu, at, 0) % (pixels - 4 * 3840);
            at = base + (lane & 7) + (lane >> 3) * 3840ull;
        }
        atomicMin(keys + at, (x >> 8) | ((unsigned long long)i << 56));
    }
}

int main() {
    for (unsigned long long pixels : {3840ull * 2160, 7680ull * 4320}) {
        unsigned long long* keys = nullptr;
        cudaMalloc(&keys, pixels * 8);
        cudaMemset(keys, 0xFF, pixels * 8);
        for (int local = 0; local < 2; local++) {
            const int blocks = 148 * 16, threads = 256, per_thread = 256;
            cudaEvent_t a, b;
            cudaEventCreate(&a); cudaEventCreate(&b);
            probe<<<blocks, threads>>>(keys, pixels, 16, local);   // warm-up
            cudaEventRecord(a);
            probe<<<blocks, threads>>>(keys, pixels, per_thread, local);
            cudaEventRecord(b);
            cudaEventSynchronize(b);
            float ms = 0;
            cudaEventElapsedTime(&ms, a, b);
            printf("%llu MB footprint, %s: %.1f G reductions/s (%.3f ms)\n", pixels * 8 >> 20, local ? "8x4-block-local" : "scattered",
                   (double)blocks * threads * per_thread / ms / 1e6, ms);
        }
        cudaFree(keys);
    }
    return cudaGetLastError() != cudaSuccess;
}
