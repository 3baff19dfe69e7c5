// Mutation fuzzer for the DDS / KTX2 parser, built by tests/test_image_files.py with -fsanitize=address,undefined:
// every seed file is parsed and decoded unmodified (must succeed) and then N times with a few bytes / the length
// changed (must return an error code or a consistent layout -- never read or write out of bounds).
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <vector>

#include "lh_image_files.h"

static uint64_t rng_state = 0x9E3779B97F4A7C15ull;
static uint64_t rng() {
    rng_state ^= rng_state << 13;
    rng_state ^= rng_state >> 7;
    rng_state ^= rng_state << 17;
    return rng_state;
}

static int run(const std::vector<uint8_t>& file, bool must_succeed) {
    LhImageInfo info;
    int rc = lh_image_parse(file.data(), file.size(), &info);
    if (rc != LH_IMAGE_OK) return must_succeed ? 1 : 0;
    if (info.levels == 0 || info.levels > LH_IMAGE_MAX_LEVELS) return 2;
    uint64_t sum = 0;
    for (uint32_t l = 0; l < info.levels; l++) sum += info.levels_info[l].bytes;
    if (sum != info.decoded_bytes) return 3;
    if (info.decoded_bytes > (1ull << 28)) return 0;   // a mutated size field: nothing to allocate for
    std::vector<uint8_t> out((size_t)info.decoded_bytes + 1, 0xA5);
    rc = lh_image_decode(file.data(), file.size(), &info, out.data(), info.decoded_bytes);
    if (out.back() != 0xA5) return 4;                  // wrote past the stated size
    if (rc != LH_IMAGE_OK && must_succeed) return 5;
    return 0;
}

int main(int argc, char** argv) {
    const int iterations = argc > 1 ? atoi(argv[1]) : 1000;
    for (int a = 2; a < argc; a++) {
        FILE* f = fopen(argv[a], "rb");
        if (!f) return 10;
        std::vector<uint8_t> seed;
        uint8_t buf[4096];
        size_t n;
        while ((n = fread(buf, 1, sizeof buf, f)) > 0) seed.insert(seed.end(), buf, buf + n);
        fclose(f);
        if (int rc = run(seed, true)) {
            fprintf(stderr, "seed %s failed: %d\n", argv[a], rc);
            return 20 + rc;
        }
        for (int i = 0; i < iterations; i++) {
            std::vector<uint8_t> m = seed;
            const int edits = 1 + (int)(rng() % 4);
            for (int e = 0; e < edits; e++) {
                // headers are where the structure lives: most edits land in the first 160 bytes
                const size_t at = (rng() & 3) ? (size_t)(rng() % (m.size() < 160 ? m.size() : 160)) : (size_t)(rng() % m.size());
                switch (rng() % 4) {
                case 0: m[at] = (uint8_t)rng(); break;
                case 1: m[at] ^= (uint8_t)(1u << (rng() % 8)); break;
                case 2: m[at] = 0xFF; break;
                default: m[at] = 0; break;
                }
            }
            if (rng() % 8 == 0) m.resize((size_t)(rng() % (m.size() + 1)));   // truncation
            if (int rc = run(m, false)) {
                fprintf(stderr, "mutation %d of %s: inconsistent result %d\n", i, argv[a], rc);
                return 40 + rc;
            }
        }
    }
    printf("fuzz ok\n");
    return 0;
}
