// Diagnostic: which cudaTextureDesc combinations does the driver accept for block-compressed mipmapped arrays?
#include <cuda_runtime.h>
#include <cstdio>
#include <vector>
int main() {
    struct Fmt { const char* name; cudaChannelFormatDesc d; int block; };
    Fmt fmts[] = {
        {"rgba8", cudaCreateChannelDesc<uchar4>(), 0},
        {"bc7", cudaCreateChannelDesc<cudaChannelFormatKindUnsignedBlockCompressed7>(), 16},
        {"bc7srgb", cudaCreateChannelDesc<cudaChannelFormatKindUnsignedBlockCompressed7SRGB>(), 16},
        {"bc1srgb", cudaCreateChannelDesc<cudaChannelFormatKindUnsignedBlockCompressed1SRGB>(), 8},
    };
    for (auto& f : fmts) {
        for (int size : {32, 256}) {
            for (int levels : {1, 3, 0}) {
                int full = 0; for (int s = size; s; s >>= 1) full++;
                int L = levels ? levels : full;
                cudaMipmappedArray_t arr = nullptr;
                cudaError_t e = cudaMallocMipmappedArray(&arr, &f.d, make_cudaExtent(size, size, 0), L);
                printf("%s size %d levels %d: malloc %s\n", f.name, size, L, cudaGetErrorString(e));
                if (e) { cudaGetLastError(); continue; }
                for (int l = 0; l < L; l++) {
                    cudaArray_t lv; e = cudaGetMipmappedArrayLevel(&lv, arr, l);
                    cudaExtent ext; cudaChannelFormatDesc dd; unsigned fl;
                    cudaArrayGetInfo(&dd, &ext, &fl, lv);
                    int w = size >> l ? size >> l : 1;
                    size_t row = f.block ? (size_t)((w + 3) / 4) * f.block : (size_t)w * 4, rows = f.block ? (w + 3) / 4 : w;
                    std::vector<unsigned char> data(row * rows, 0x40);
                    cudaError_t e2 = cudaMemcpy2DToArray(lv, 0, 0, data.data(), row, row, rows, cudaMemcpyHostToDevice);
                    printf("   level %d ext %zux%zu get %s copy %s\n", l, ext.width, ext.height, cudaGetErrorString(e), cudaGetErrorString(e2));
                    cudaGetLastError();
                }
                for (int rm = 0; rm < 2; rm++) for (int srgb = 0; srgb < 2; srgb++) for (int aniso : {0, 1, 16}) for (int mipf = 0; mipf < 2; mipf++) {
                    cudaResourceDesc res{}; res.resType = cudaResourceTypeMipmappedArray; res.res.mipmap.mipmap = arr;
                    cudaTextureDesc td{};
                    td.addressMode[0] = td.addressMode[1] = td.addressMode[2] = cudaAddressModeWrap;
                    td.filterMode = cudaFilterModeLinear;
                    td.mipmapFilterMode = mipf ? cudaFilterModeLinear : cudaFilterModePoint;
                    td.readMode = rm ? cudaReadModeNormalizedFloat : cudaReadModeElementType;
                    td.normalizedCoords = 1; td.maxAnisotropy = aniso; td.maxMipmapLevelClamp = (float)(L - 1); td.sRGB = srgb;
                    cudaTextureObject_t t = 0;
                    e = cudaCreateTextureObject(&t, &res, &td, nullptr);
                    printf("   readMode %s srgb %d aniso %d mipLinear %d -> %s\n", rm ? "norm" : "elem", srgb, aniso, mipf, cudaGetErrorString(e));
                    cudaGetLastError();
                    if (!e) cudaDestroyTextureObject(t);
                }
                cudaFreeMipmappedArray(arr);
            }
        }
    }
    return 0;
}
