// common.hpp -- helpers of the C++ front's sample programs (tests/cpp/*.cpp): synthetic images and the
// comparison conventions of the reference's samples (samples-public/common/hipacc_helper.hpp:180-207:
// integer images |diff| <= tolerance, float images relative error), written from scratch.
#pragma once
#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <vector>

namespace tc {

inline uint64_t splitmix64(uint64_t z) {
    z += 0x9E3779B97F4A7C15ull;
    z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
    z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
    return z ^ (z >> 31);
}
// counter-based generator of SURVEY.md section 8d: any pixel of any strip is a pure function of (x, y, seed)
inline uint64_t pixel_bits(int x, int y, uint64_t seed) { return splitmix64(seed ^ (((uint64_t)(uint32_t)y << 32) + (uint32_t)x)); }
inline std::vector<unsigned char> image_u8(int w, int h, uint64_t seed) {
    std::vector<unsigned char> v((size_t)w * h);
    for (int y = 0; y < h; ++y)
        for (int x = 0; x < w; ++x) v[(size_t)y * w + x] = (unsigned char)(pixel_bits(x, y, seed) & 0xFF);
    return v;
}
inline std::vector<float> image_f32(int w, int h, uint64_t seed, float scale = 1.0f) {
    std::vector<float> v((size_t)w * h);
    for (int y = 0; y < h; ++y)
        for (int x = 0; x < w; ++x) v[(size_t)y * w + x] = (float)(pixel_bits(x, y, seed) >> 40) * (1.0f / 16777216.0f) * scale;
    return v;
}
// smooth blocks + noise: gives the Harris detector real corners
inline std::vector<unsigned char> image_blocks(int w, int h, uint64_t seed) {
    std::vector<unsigned char> v((size_t)w * h);
    for (int y = 0; y < h; ++y)
        for (int x = 0; x < w; ++x) {
            const int base = (((x / 16) + (y / 16)) & 1) ? 200 : 40;
            v[(size_t)y * w + x] = (unsigned char)(base + (int)(pixel_bits(x, y, seed) % 15));
        }
    return v;
}

inline int clampi(int v, int lo, int hi) { return v < lo ? lo : v > hi ? hi : v; }
inline int mirrori(int v, int n) { return v < 0 ? -v - 1 : v >= n ? n - (v + 1 - n) : v; }

template <typename T> long count_diff(const T *a, const T *b, size_t n, int tol, long *first = nullptr) {
    long bad = 0;
    for (size_t i = 0; i < n; ++i) {
        const long d = (long)a[i] - (long)b[i];
        if (d > tol || d < -tol) {
            if (!bad && first) *first = (long)i;
            ++bad;
        }
    }
    return bad;
}
inline long count_diff_rel(const float *a, const float *b, size_t n, double rel, double abs_floor, long *first = nullptr) {
    long bad = 0;
    for (size_t i = 0; i < n; ++i) {
        const double d = std::fabs((double)a[i] - (double)b[i]);
        if (d > rel * std::fabs((double)b[i]) + abs_floor) {
            if (!bad && first) *first = (long)i;
            ++bad;
        }
    }
    return bad;
}
inline int verdict(const char *name, long bad, size_t n, long first) {
    if (bad == 0) {
        std::printf("%s: Test PASSED (%zu pixels)\n", name, n);
        return 0;
    }
    std::printf("%s: Test FAILED, %ld of %zu pixels differ (first at index %ld)\n", name, bad, n, first);
    return 1;
}

// command line of the compiled-body programs: [width height] [--io in.raw out.raw] (raw = dense pixels, row-major); the
// GPU tests feed the inputs of tests/golden/*.npz through --io and compare the output file with the golden array
struct Args {
    int w, h;
    const char *in = nullptr, *out = nullptr;
    Args(int argc, char **argv, int w0, int h0) : w(w0), h(h0) {
        int pos = 0;
        for (int i = 1; i < argc; ++i) {
            if (!std::strcmp(argv[i], "--io") && i + 2 < argc) { in = argv[i + 1]; out = argv[i + 2]; i += 2; }
            else if (pos == 0) { w = std::atoi(argv[i]); ++pos; }
            else if (pos == 1) { h = std::atoi(argv[i]); ++pos; }
        }
    }
};
template <typename T> inline void read_raw(const char *path, std::vector<T> &v) {
    std::FILE *f = std::fopen(path, "rb");
    if (!f || std::fread(v.data(), sizeof(T), v.size(), f) != v.size()) { std::fprintf(stderr, "cannot read %s\n", path); std::exit(2); }
    std::fclose(f);
}
template <typename T> inline void write_raw(const char *path, const T *p, size_t n) {
    std::FILE *f = std::fopen(path, "wb");
    if (!f || std::fwrite(p, sizeof(T), n, f) != n) { std::fprintf(stderr, "cannot write %s\n", path); std::exit(2); }
    std::fclose(f);
}

}  // namespace tc
