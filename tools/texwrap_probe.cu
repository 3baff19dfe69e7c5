// Third texture-unit probe: does the model of tools/texvalue_probe.cu hold for a texture whose sides are not powers of
// two (48 x 80) and for coordinates outside [0, 1) under repeat addressing (u, v in [-4, 4))?  Same dump format, files
// texwrap_{unorm,srgb}.bin.
// build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o build/texwrap_probe tools/texwrap_probe.cu
#include <cstdio>
#include <cstdint>
#include <vector>
#include <cuda_runtime.h>
__global__ void sample(cudaTextureObject_t t, const float2* uv, float4* out, int n, float lod) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) out[i] = tex2DLod<float4>(t, uv[i].x, uv[i].y, lod);
}
static uint32_t lcg(uint32_t& s) { s = s * 1664525u + 1013904223u; return s >> 8; }
int main(int argc, char** argv) {
    const char* dir = argc > 1 ? argv[1] : "gpurun_out";
    const int W = 48, H = 80, n = 1 << 15;
    for (int srgb = 0; srgb < 2; srgb++) {
        cudaChannelFormatDesc d = cudaCreateChannelDesc<uchar4>();
        cudaMipmappedArray_t arr;
        cudaMallocMipmappedArray(&arr, &d, make_cudaExtent(W, H, 0), 1);
        cudaArray_t lv;
        cudaGetMipmappedArrayLevel(&lv, arr, 0);
        uint32_t seed = 12345u + srgb;
        std::vector<unsigned char> px(W * H * 4);
        for (auto& b : px) b = (unsigned char)(lcg(seed) & 255u);
        cudaMemcpy2DToArray(lv, 0, 0, px.data(), W * 4, W * 4, H, cudaMemcpyHostToDevice);
        cudaResourceDesc res{};
        res.resType = cudaResourceTypeMipmappedArray;
        res.res.mipmap.mipmap = arr;
        cudaTextureDesc td{};
        td.addressMode[0] = td.addressMode[1] = cudaAddressModeWrap;
        td.filterMode = cudaFilterModeLinear;
        td.mipmapFilterMode = cudaFilterModePoint;
        td.readMode = cudaReadModeNormalizedFloat;
        td.normalizedCoords = 1;
        td.sRGB = srgb;
        cudaTextureObject_t t;
        if (cudaCreateTextureObject(&t, &res, &td, nullptr)) { printf("create failed\n"); return 1; }
        std::vector<float2> uv(n);
        for (int i = 0; i < n; i++) {
            uv[i].x = (float)lcg(seed) / 16777216.0f * 8.0f - 4.0f;
            uv[i].y = (float)lcg(seed) / 16777216.0f * 8.0f - 4.0f;
        }
        float2* duv; float4* dout;
        cudaMalloc(&duv, n * sizeof(float2)); cudaMalloc(&dout, n * sizeof(float4));
        cudaMemcpy(duv, uv.data(), n * sizeof(float2), cudaMemcpyHostToDevice);
        sample<<<(n + 255) / 256, 256>>>(t, duv, dout, n, 0.0f);
        std::vector<float4> out(n);
        cudaMemcpy(out.data(), dout, n * sizeof(float4), cudaMemcpyDeviceToHost);
        char name[256];
        snprintf(name, sizeof name, "%s/texwrap_%s.bin", dir, srgb ? "srgb" : "unorm");
        FILE* f = fopen(name, "wb");
        fwrite(px.data(), 1, px.size(), f);
        for (int i = 0; i < n; i++) {
            float rec[6] = {uv[i].x, uv[i].y, out[i].x, out[i].y, out[i].z, out[i].w};
            fwrite(rec, 4, 6, f);
        }
        fclose(f);
    }
    cudaError_t e = cudaDeviceSynchronize();
    printf("texwrap_probe: %s\n", cudaGetErrorString(e));
    return e != cudaSuccess;
}
