// How does the texture unit quantise bilinear weights?  (The oracle's sampler pins the same sub-texel precision: Vulkan
// leaves it implementation-defined, VkPhysicalDeviceLimits::subTexelPrecisionBits.)  A 4x1 RGBA8 texture with texels
// 0, 255, 0, 255 (unorm, and sRGB) is sampled at 8192 positions per texel interval; prints the returned weight for
// every position as "pos value" so the staircase (step width, step phase, rounding) can be read off.
// build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/texfilter_probe tools/texfilter_probe.cu
#include <cstdio>
#include <vector>
#include <cuda_runtime.h>
__global__ void sweep(cudaTextureObject_t t, float* out, int n, float x0, float x1, float lod) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) out[i] = tex2DLod<float4>(t, x0 + (x1 - x0) * (float)i / (float)n, 0.5f, lod).x;
}
int main() {
  for (int W : {4, 1024}) {
    for (int srgb = 0; srgb < 2; srgb++) {
        cudaChannelFormatDesc d = cudaCreateChannelDesc<uchar4>();
        cudaMipmappedArray_t arr; cudaMallocMipmappedArray(&arr, &d, make_cudaExtent(W, 1, 0), 1);
        cudaArray_t lv; cudaGetMipmappedArrayLevel(&lv, arr, 0);
        std::vector<unsigned char> px(W * 4); for (int i = 0; i < W; i++) for (int c = 0; c < 4; c++) px[4 * i + c] = (i & 1) ? 255 : 0;
        cudaMemcpy2DToArray(lv, 0, 0, px.data(), W * 4, W * 4, 1, cudaMemcpyHostToDevice);
        cudaResourceDesc res{}; res.resType = cudaResourceTypeMipmappedArray; res.res.mipmap.mipmap = arr;
        cudaTextureDesc td{}; td.addressMode[0] = td.addressMode[1] = cudaAddressModeWrap; td.filterMode = cudaFilterModeLinear;
        td.mipmapFilterMode = cudaFilterModePoint; td.readMode = cudaReadModeNormalizedFloat; td.normalizedCoords = 1; td.sRGB = srgb;
        cudaTextureObject_t t; if (cudaCreateTextureObject(&t, &res, &td, nullptr)) { printf("create failed\n"); return 1; }
        const int n = 8192; float* out; cudaMalloc(&out, n * 4);
        // u from texel centre 0 (0.5/W) to texel centre 1 (1.5/W): weight of texel 1 goes 0 -> 1
        const int k = W == 4 ? 0 : 700;   // between the centres of texels k and k + 1 (k even: the weight of texel k + 1 goes 0 -> 1)
        sweep<<<(n + 255) / 256, 256>>>(t, out, n, (k + 0.5f) / W, (k + 1.5f) / W, 0.0f);
        std::vector<float> h(n); cudaMemcpy(h.data(), out, n * 4, cudaMemcpyDeviceToHost);
        printf("# srgb %d W %d\n", srgb, W);
        for (int i = 0; i < n; i++) printf("%d %.9g\n", i, h[i]);
    }
  }
    return cudaGetLastError() != cudaSuccess;
}
