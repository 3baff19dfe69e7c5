// Fourth texture-unit probe: linear filtering of 32-bit float textures, as the skybox (cube map) and the display LUT (3D)
// use it.  Dumps, for random float texels and random coordinates:
//   texfloat_2d.bin    64x64 float4, clamp addressing:   [64*64*4 floats][n * (u v  r g b a)]
//   texfloat_cube.bin  16x16x6 float4 cube map:          [6*16*16*4 floats][n * (x y z  r g b a)]
//   texfloat_3d.bin    12x12x12 float4, clamp addressing:[12^3*4 floats][n * (u v w  r g b a)]
// build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o build/texfloat_probe tools/texfloat_probe.cu
#include <cstdio>
#include <cstdint>
#include <vector>
#include <cuda_runtime.h>
__global__ void s2d(cudaTextureObject_t t, const float3* c, float4* out, int n) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) out[i] = tex2DLod<float4>(t, c[i].x, c[i].y, 0.0f);
}
__global__ void scube(cudaTextureObject_t t, const float3* c, float4* out, int n) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) out[i] = texCubemapLod<float4>(t, c[i].x, c[i].y, c[i].z, 0.0f);
}
__global__ void s3d(cudaTextureObject_t t, const float3* c, float4* out, int n) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) out[i] = tex3D<float4>(t, c[i].x, c[i].y, c[i].z);
}
static uint32_t lcg(uint32_t& s) { s = s * 1664525u + 1013904223u; return s >> 8; }
static float frand(uint32_t& s) { return (float)lcg(s) / 16777216.0f; }

template <typename K>
static void run(const char* path, cudaArray_t arr, const std::vector<float>& texels, const std::vector<float3>& coords, K kernel) {
    cudaResourceDesc res{};
    res.resType = cudaResourceTypeArray;
    res.res.array.array = arr;
    cudaTextureDesc td{};
    td.addressMode[0] = td.addressMode[1] = td.addressMode[2] = cudaAddressModeClamp;
    td.filterMode = cudaFilterModeLinear;
    td.readMode = cudaReadModeElementType;
    td.normalizedCoords = 1;
    cudaTextureObject_t t = 0;
    if (cudaCreateTextureObject(&t, &res, &td, nullptr)) { printf("create failed for %s\n", path); return; }
    const int n = (int)coords.size();
    float3* dc; float4* dout;
    cudaMalloc(&dc, n * sizeof(float3)); cudaMalloc(&dout, n * sizeof(float4));
    cudaMemcpy(dc, coords.data(), n * sizeof(float3), cudaMemcpyHostToDevice);
    kernel<<<(n + 255) / 256, 256>>>(t, dc, dout, n);
    std::vector<float4> out(n);
    cudaMemcpy(out.data(), dout, n * sizeof(float4), cudaMemcpyDeviceToHost);
    FILE* f = fopen(path, "wb");
    fwrite(texels.data(), 4, texels.size(), f);
    for (int i = 0; i < n; i++) {
        float rec[7] = {coords[i].x, coords[i].y, coords[i].z, out[i].x, out[i].y, out[i].z, out[i].w};
        fwrite(rec, 4, 7, f);
    }
    fclose(f);
    cudaFree(dc); cudaFree(dout);
    cudaDestroyTextureObject(t);
}

int main(int argc, char** argv) {
    const char* dir = argc > 1 ? argv[1] : "gpurun_out";
    char name[256];
    uint32_t seed = 777u;
    const int n = 1 << 15;
    cudaChannelFormatDesc desc = cudaCreateChannelDesc<float4>();
    {   // 2D: channel r in [0, 1), g in [0, 64) (HDR-like), b = small integers / 8 (exact products), a in [-1, 1)
        const int W = 64;
        std::vector<float> tx(W * W * 4);
        for (int i = 0; i < W * W; i++) { tx[4 * i] = frand(seed); tx[4 * i + 1] = frand(seed) * 64.0f; tx[4 * i + 2] = (float)(lcg(seed) & 255u) / 8.0f; tx[4 * i + 3] = frand(seed) * 2.0f - 1.0f; }
        cudaArray_t arr; cudaMallocArray(&arr, &desc, W, W);
        cudaMemcpy2DToArray(arr, 0, 0, tx.data(), W * 16, W * 16, W, cudaMemcpyHostToDevice);
        std::vector<float3> c(n);
        for (int i = 0; i < n; i++) c[i] = make_float3(frand(seed) * 1.1f - 0.05f, frand(seed) * 1.1f - 0.05f, 0.0f);   // a little outside: clamp
        snprintf(name, sizeof name, "%s/texfloat_2d.bin", dir);
        run(name, arr, tx, c, s2d);
    }
    {   // cube map
        const int S = 16;
        std::vector<float> tx(6 * S * S * 4);
        for (auto& v : tx) v = frand(seed) * 4.0f;
        cudaArray_t arr; cudaMalloc3DArray(&arr, &desc, make_cudaExtent(S, S, 6), cudaArrayCubemap);
        cudaMemcpy3DParms cp{};
        cp.srcPtr = make_cudaPitchedPtr(tx.data(), (size_t)S * 16, S, S);
        cp.dstArray = arr; cp.extent = make_cudaExtent(S, S, 6); cp.kind = cudaMemcpyHostToDevice;
        cudaMemcpy3D(&cp);
        std::vector<float3> c(n);
        for (int i = 0; i < n; i++) c[i] = make_float3(frand(seed) * 2.0f - 1.0f, frand(seed) * 2.0f - 1.0f, frand(seed) * 2.0f - 1.0f);
        snprintf(name, sizeof name, "%s/texfloat_cube.bin", dir);
        run(name, arr, tx, c, scube);
    }
    {   // 3D
        const int S = 12;
        std::vector<float> tx(S * S * S * 4);
        for (auto& v : tx) v = frand(seed);
        cudaArray_t arr; cudaMalloc3DArray(&arr, &desc, make_cudaExtent(S, S, S), 0);
        cudaMemcpy3DParms cp{};
        cp.srcPtr = make_cudaPitchedPtr(tx.data(), (size_t)S * 16, S, S);
        cp.dstArray = arr; cp.extent = make_cudaExtent(S, S, S); cp.kind = cudaMemcpyHostToDevice;
        cudaMemcpy3D(&cp);
        std::vector<float3> c(n);
        for (int i = 0; i < n; i++) c[i] = make_float3(frand(seed), frand(seed), frand(seed));
        snprintf(name, sizeof name, "%s/texfloat_3d.bin", dir);
        run(name, arr, tx, c, s3d);
    }
    cudaError_t e = cudaDeviceSynchronize();
    printf("texfloat_probe: %s\n", cudaGetErrorString(e));
    return e != cudaSuccess;
}
