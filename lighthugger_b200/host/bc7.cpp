// BC7 encoder (mode 6 only: one subset, 7777.1 endpoints, 4-bit indices) used
// by the scene generator to turn the procedural RGBA8 textures into the
// block-compressed format the reference loads from DDS
// (src/resources/image_loading.cpp:14-39 registers every BC texture as sRGB).
// Mode 6 represents smooth RGBA content well and keeps the encoder small; the
// bit layout follows the published BC7 specification.
#include <algorithm>
#include <cmath>
#include <cstdint>
#include <cstring>

#include "lh_scene.h"

namespace {

const int kWeights4[16] = {0, 4, 9, 13, 17, 21, 26, 30, 34, 38, 43, 47, 51, 55, 60, 64};

struct BitWriter {
    uint8_t* dst;
    int pos = 0;
    void put(uint32_t value, int bits) {
        for (int i = 0; i < bits; i++, pos++)
            if ((value >> i) & 1u) dst[pos >> 3] |= (uint8_t)(1u << (pos & 7));
    }
};

inline int interp(int e0, int e1, int w) { return ((64 - w) * e0 + w * e1 + 32) >> 6; }

// error of a candidate endpoint pair (8-bit per channel) + best indices
long block_error(const uint8_t px[16][4], const int e0[4], const int e1[4], uint8_t idx[16]) {
    int pal[16][4];
    for (int k = 0; k < 16; k++)
        for (int c = 0; c < 4; c++) pal[k][c] = interp(e0[c], e1[c], kWeights4[k]);
    long total = 0;
    for (int t = 0; t < 16; t++) {
        long best = -1;
        int bi = 0;
        for (int k = 0; k < 16; k++) {
            long e = 0;
            for (int c = 0; c < 4; c++) {
                long d = (long)px[t][c] - pal[k][c];
                e += d * d;
            }
            if (best < 0 || e < best) {
                best = e;
                bi = k;
            }
        }
        idx[t] = (uint8_t)bi;
        total += best;
    }
    return total;
}

void encode_block(const uint8_t px[16][4], uint8_t out[16]) {
    // principal axis by power iteration on the 4-D covariance
    float mean[4] = {0, 0, 0, 0};
    for (int t = 0; t < 16; t++)
        for (int c = 0; c < 4; c++) mean[c] += px[t][c];
    for (int c = 0; c < 4; c++) mean[c] /= 16.0f;
    float cov[4][4] = {};
    for (int t = 0; t < 16; t++)
        for (int a = 0; a < 4; a++)
            for (int b = 0; b < 4; b++) cov[a][b] += (px[t][a] - mean[a]) * (px[t][b] - mean[b]);
    float axis[4] = {1, 1, 1, 1};
    for (int it = 0; it < 8; it++) {
        float n[4] = {0, 0, 0, 0};
        for (int a = 0; a < 4; a++)
            for (int b = 0; b < 4; b++) n[a] += cov[a][b] * axis[b];
        float len = std::sqrt(n[0] * n[0] + n[1] * n[1] + n[2] * n[2] + n[3] * n[3]);
        if (len < 1e-6f) break;
        for (int a = 0; a < 4; a++) axis[a] = n[a] / len;
    }
    float alen = std::sqrt(axis[0] * axis[0] + axis[1] * axis[1] + axis[2] * axis[2] + axis[3] * axis[3]);
    for (int a = 0; a < 4; a++) axis[a] /= alen;
    float tmin = 1e30f, tmax = -1e30f;
    for (int t = 0; t < 16; t++) {
        float d = 0;
        for (int c = 0; c < 4; c++) d += (px[t][c] - mean[c]) * axis[c];
        tmin = std::min(tmin, d);
        tmax = std::max(tmax, d);
    }
    float lo[4], hi[4];
    for (int c = 0; c < 4; c++) {
        lo[c] = std::min(255.0f, std::max(0.0f, mean[c] + axis[c] * tmin));
        hi[c] = std::min(255.0f, std::max(0.0f, mean[c] + axis[c] * tmax));
    }
    // quantise to 7 bits + shared p-bit per endpoint; try the four p-bit pairs
    long best_err = -1;
    int best_e0[4] = {0, 0, 0, 0}, best_e1[4] = {0, 0, 0, 0}, best_p0 = 0, best_p1 = 0;
    uint8_t best_idx[16] = {};
    for (int p0 = 0; p0 < 2; p0++)
        for (int p1 = 0; p1 < 2; p1++) {
            int e0[4], e1[4];
            for (int c = 0; c < 4; c++) {
                int q0 = (int)std::lround((lo[c] - p0) / 2.0f), q1 = (int)std::lround((hi[c] - p1) / 2.0f);
                q0 = std::min(127, std::max(0, q0));
                q1 = std::min(127, std::max(0, q1));
                e0[c] = (q0 << 1) | p0;
                e1[c] = (q1 << 1) | p1;
            }
            uint8_t idx[16];
            long err = block_error(px, e0, e1, idx);
            if (best_err < 0 || err < best_err) {
                best_err = err;
                memcpy(best_e0, e0, sizeof(e0));
                memcpy(best_e1, e1, sizeof(e1));
                memcpy(best_idx, idx, 16);
                best_p0 = p0;
                best_p1 = p1;
            }
        }
    // the anchor index (texel 0) stores 3 bits: its MSB must be 0
    if (best_idx[0] & 8) {
        std::swap_ranges(best_e0, best_e0 + 4, best_e1);
        std::swap(best_p0, best_p1);
        for (int t = 0; t < 16; t++) best_idx[t] = (uint8_t)(15 - best_idx[t]);
    }
    memset(out, 0, 16);
    BitWriter w{out};
    w.put(1u << 6, 7);  // mode 6
    for (int c = 0; c < 4; c++) {
        w.put((uint32_t)best_e0[c] >> 1, 7);
        w.put((uint32_t)best_e1[c] >> 1, 7);
    }
    w.put((uint32_t)best_p0, 1);
    w.put((uint32_t)best_p1, 1);
    w.put(best_idx[0], 3);
    for (int t = 1; t < 16; t++) w.put(best_idx[t], 4);
}

}  // namespace

extern "C" uint64_t lh_bc7_encode(const uint8_t* rgba8, uint32_t width, uint32_t height, uint8_t* dst) {
    uint32_t bw = (width + 3) / 4, bh = (height + 3) / 4;
    uint64_t total = (uint64_t)bw * bh * 16;
    if (!dst || !rgba8) return total;
    for (uint32_t by = 0; by < bh; by++)
        for (uint32_t bx = 0; bx < bw; bx++) {
            uint8_t px[16][4];
            for (int t = 0; t < 16; t++) {
                uint32_t x = std::min(width - 1, bx * 4 + (uint32_t)(t & 3));
                uint32_t y = std::min(height - 1, by * 4 + (uint32_t)(t >> 2));
                memcpy(px[t], rgba8 + ((uint64_t)y * width + x) * 4, 4);
            }
            encode_block(px, dst + ((uint64_t)by * bw + bx) * 16);
        }
    return total;
}
