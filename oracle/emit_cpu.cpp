// oracle/emit_cpu.cpp -- TEST INFRASTRUCTURE ONLY: the CPU oracle / CPU baseline.
//
// A hand restatement of what Hipacc's `-emit-cpu` backend generates for the hot path,
// driven by the SAME descriptors as the product's C ABI (include/hipacc_b200.h) but with
// HOST pointers in hb_view::data.  Only tests/, __graft_entry__.smoke() and bench.py's
// cpu_baseline / --impl reference legs may load this library; the product never does.
//
// Parity status: PINNED.  tests/test_oracle.py checks every function here
// bit-for-bit (uchar/int) or to 1 ulp-level tolerance (float) against oracle/_ref
// (the reference's own DSL headers + sample kernels executed as C++), against the samples'
// embedded plain-C checkers, against SURVEY.md appendix A known-answer vectors and against
// the committed fixtures in tests/golden/.
//
// What is restated (paths relative to the Hipacc tree):
//   loop nest + OpenMP row loop          lib/Backend/CPU_x86.cpp:3583-3691
//   IS / accessor offset arithmetic      lib/AST/MemoryAccess.cpp:98-177, dsl/image.hpp:412
//   boundary index remap and its order   lib/AST/BorderHandling.cpp:41-120,339-366
//   tap order / fold                     lib/AST/Convolution.cpp:90-99,397-435, dsl/kernel.hpp:241-315
//   interpolation (NN, LF, B5, CF, L3)    dsl/image.hpp:321-528, lib/AST/Interpolate.cpp:85-113
//   global reduction                     dsl/kernel.hpp:121-151, runtime/hipacc_cpu_red.hpp:19-68
// Build: g++ -O3 -fopenmp -ffp-contract=off -mno-fma  (no FMA contraction: float results
// equal the DSL's unfused multiply + add).

#include <algorithm>
#include <cmath>
#include <cstdint>
#include <cstring>
#include <limits>
#include <vector>
#include <omp.h>

#include "../include/hipacc_b200.h"

namespace {

typedef unsigned char uchar;

struct Region {  // boundary window of an accessor, vertical ghost extension included
    int lo_x, hi_x, lo_y, hi_y;
};

inline void norm_view(hb_view &v) {
    if (v.width <= 0 || v.height <= 0) { v.width = v.img_width; v.height = v.img_height; v.offset_x = 0; v.offset_y = 0; }
}
inline Region region_of(const hb_view &v) {
    return Region{v.offset_x, v.offset_x + v.width, v.offset_y - v.ghost_top, v.offset_y + v.height + v.ghost_bottom};
}

// lib/AST/BorderHandling.cpp:41-120: upper test first, then lower (order of :339-366)
inline int remap(int idx, int lo, int hi, int mode) {
    switch (mode) {
    case HB_BOUNDARY_CLAMP:
        if (idx >= hi) idx = hi - 1;
        if (idx < lo) idx = lo;
        break;
    case HB_BOUNDARY_REPEAT:
        while (idx >= hi) idx -= (hi - lo);
        while (idx < lo) idx += (hi - lo);
        break;
    case HB_BOUNDARY_MIRROR:
        if (idx >= hi) idx = hi - (idx + 1 - hi);
        if (idx < lo) idx = lo + (lo - idx - 1);
        break;
    default: break;
    }
    return idx;
}

template <typename T>
struct Img {
    const T *p; int stride, iw, ih;
    Region r; int mode; T cval;
    // neighbour fetch through the boundary mode (dsl/image.hpp:574-612)
    inline T at(int x, int y) const {
        if (mode == HB_BOUNDARY_CONSTANT) {
            if (x < r.lo_x || x >= r.hi_x || y < r.lo_y || y >= r.hi_y) return cval;
            return p[(size_t)y * stride + x];
        }
        x = remap(x, r.lo_x, r.hi_x, mode);
        y = remap(y, r.lo_y, r.hi_y, mode);
        // UNDEFINED (and degenerate halo > size cases): stay inside the allocation
        x = std::min(std::max(x, 0), iw - 1);
        y = std::min(std::max(y, 0), ih - 1);
        return p[(size_t)y * stride + x];
    }
    inline T raw(int x, int y) const { return p[(size_t)y * stride + x]; }
};

template <typename T>
Img<T> make_img(const hb_view &v, int mode, double cval) {
    Img<T> im;
    im.p = static_cast<const T *>(v.data); im.stride = v.stride; im.iw = v.img_width; im.ih = v.img_height;
    im.r = region_of(v); im.mode = mode; im.cval = (T)cval;
    return im;
}

// C conversions as g++/x86 performs them: float -> integer goes through int (cvttss2si)
template <typename TO, typename TA> inline TO cast_out(TA v) { return (TO)v; }
template <> inline uchar cast_out<uchar, float>(float v) { return (uchar)(int)v; }
template <> inline signed char cast_out<signed char, float>(float v) { return (signed char)(int)v; }
template <> inline short cast_out<short, float>(float v) { return (short)(int)v; }
template <> inline unsigned short cast_out<unsigned short, float>(float v) { return (unsigned short)(int)v; }

template <typename TA> inline TA fold(TA acc, TA v, int mode) {
    switch (mode) {
    case HB_REDUCE_SUM: return acc + v;
    case HB_REDUCE_MIN: return (v < acc ? v : acc);   // hipacc::math::min(fun(), result)
    case HB_REDUCE_MAX: return (v > acc ? v : acc);
    default: return acc * v;
    }
}

template <typename TO, typename TA>
inline TO epilogue(TA acc, int epi, const double *p) {
    switch (epi) {
    case HB_EPI_ADD_CAST: return cast_out<TO, TA>(acc + (TA)p[0]);
    case HB_EPI_ADD_CLAMP_CAST: {
        TA v = acc + (TA)p[0];
        v = (v < (TA)p[2] ? v : (TA)p[2]);
        v = (v > (TA)p[1] ? v : (TA)p[1]);
        return cast_out<TO, TA>(v);
    }
    case HB_EPI_DIVI_CAST: return cast_out<TO, int>((int)acc / (int)p[0]);
    case HB_EPI_DIVF_CAST: return cast_out<TO, float>((float)acc / (float)p[0]);
    default: return cast_out<TO, TA>(acc);
    }
}

struct Taps {  // visited taps in row-major order
    std::vector<int> dx, dy; std::vector<float> cf; std::vector<int> ci;
};

Taps build_taps(const hb_local_desc &d) {
    Taps t;
    int n = d.size_x * d.size_y;
    for (int i = 0; i < n; ++i) {
        int ty = i / d.size_x, tx = i % d.size_x;
        bool on = true;
        if (d.kind == HB_LOCAL_REDUCE_DOMAIN) {
            if (d.domain) on = d.domain[i] != 0;
            else if (d.tap == HB_TAP_MUL) on = d.coef_f32 ? (d.coef_f32[i] != 0.0f) : (d.coef_s32[i] != 0);
        }
        if (!on) continue;
        t.dx.push_back(tx - d.size_x / 2); t.dy.push_back(ty - d.size_y / 2);
        t.cf.push_back(d.coef_f32 ? d.coef_f32[i] : (d.coef_s32 ? (float)d.coef_s32[i] : 1.0f));
        t.ci.push_back(d.coef_s32 ? d.coef_s32[i] : 0);
    }
    return t;
}

template <typename TI, typename TA, typename TO>
int local_typed(const hb_local_desc &d) {
    hb_view in = d.in, out = d.out;
    norm_view(in); norm_view(out);
    Img<TI> im = make_img<TI>(in, d.boundary, d.boundary_const);
    TO *op = static_cast<TO *>(out.data);
    Taps t = build_taps(d);
    const int nt = (int)t.dx.size();
    if (nt == 0) return HB_ERR_INVALID;
    const bool fcoef = d.coef_f32 != nullptr;
    const int hx = d.size_x / 2, hy = d.size_y / 2;
    const int mode = d.reduce_mode, tapk = d.tap, epi = d.epilogue, accdt = d.acc_dtype;
    const double *ep = d.epi_p;

#pragma omp parallel for schedule(static)
    for (int gy = 0; gy < out.height; ++gy) {
        const int iy = in.offset_y + gy;  // centre in input coordinates (dsl/image.hpp:412)
        const bool row_in = (iy - hy >= im.r.lo_y) && (iy + hy < im.r.hi_y);
        for (int gx = 0; gx < out.width; ++gx) {
            const int ix = in.offset_x + gx;
            const bool interior = row_in && (ix - hx >= im.r.lo_x) && (ix + hx < im.r.hi_x);
            TA acc = 0;
            for (int k = 0; k < nt; ++k) {
                TI pix = interior ? im.raw(ix + t.dx[k], iy + t.dy[k]) : im.at(ix + t.dx[k], iy + t.dy[k]);
                TA v;
                if (tapk == HB_TAP_IN) v = (TA)pix;
                else if (fcoef || std::is_same<TI, float>::value)
                    v = (TA)(t.cf[k] * (float)pix);               // float operand: C promotes the product to float
                else v = (TA)(t.ci[k] * (int)pix);                // int mask on integer pixels: int product
                acc = (k == 0) ? v : fold<TA>(acc, v, mode);      // first tap initialises (dsl/kernel.hpp:250)
            }
            if (accdt == HB_S16) acc = (TA)(short)acc;
            op[(size_t)(out.offset_y + gy) * out.stride + out.offset_x + gx] = epilogue<TO, TA>(acc, epi, ep);
        }
    }
    return HB_OK;
}

template <typename TI, typename TA>
int local_out(const hb_local_desc &d) {
    switch (d.out.dtype) {
    case HB_U8: return local_typed<TI, TA, uchar>(d);
    case HB_S8: return local_typed<TI, TA, signed char>(d);
    case HB_U16: return local_typed<TI, TA, unsigned short>(d);
    case HB_S16: return local_typed<TI, TA, short>(d);
    case HB_S32: return local_typed<TI, TA, int>(d);
    case HB_F32: return local_typed<TI, TA, float>(d);
    default: return HB_ERR_UNSUPPORTED;
    }
}
template <typename TI>
int local_acc(const hb_local_desc &d) {
    if (d.acc_dtype == HB_F32) return local_out<TI, float>(d);
    if (d.acc_dtype == HB_S32 || d.acc_dtype == HB_S16) return local_out<TI, int>(d);
    return HB_ERR_UNSUPPORTED;
}

// ------------------------------------------------------------------ bilateral
template <typename T>
int bilateral_typed(const hb_bilateral_desc &d) {
    hb_view in = d.in, out = d.out;
    norm_view(in); norm_view(out);
    Img<T> im = make_img<T>(in, d.boundary, d.boundary_const);
    T *op = static_cast<T *>(out.data);
    const int s = d.size, h = s / 2;
    const float c_r = 0.5f / (d.sigma_r * d.sigma_r);
#pragma omp parallel for schedule(dynamic, 4)
    for (int gy = 0; gy < out.height; ++gy) {
        const int iy = in.offset_y + gy;
        for (int gx = 0; gx < out.width; ++gx) {
            const int ix = in.offset_x + gx;
            float dsum = 0.0f, p = 0.0f;
            const float center = (float)im.at(ix, iy);
            for (int ty = 0; ty < s; ++ty)
                for (int tx = 0; tx < s; ++tx) {
                    const float m = d.coef_f32[ty * s + tx];
                    if (m == 0.0f) continue;  // Domain(mask) holes (dsl/mask.hpp:238-250)
                    const float v = (float)im.at(ix + tx - h, iy + ty - h);
                    const float diff = v - center;
                    const float w = expf(-c_r * diff * diff) * m;
                    dsum += w;
                    p += w * v;
                }
            T o;
            if (std::is_same<T, float>::value) o = (T)(p / dsum);
            else o = cast_out<T, float>(p / dsum + 0.5f);
            op[(size_t)(out.offset_y + gy) * out.stride + out.offset_x + gx] = o;
        }
    }
    return HB_OK;
}

// ------------------------------------------------------------------ point operators
// interpolation weights (dsl/image.hpp:321-383)
inline float interp_binomial5(float diff) {
    diff = std::fabs(diff);
    return diff < 0.5f ? 6.0f / 8.0f : diff < 1.0f ? 4.0f / 8.0f : diff < 1.5f ? 1.0f / 8.0f : 0.0f;
}
inline float interp_bicubic(float diff) {   // Keys' cubic convolution, a = -0.5
    diff = std::fabs(diff);
    const float a = -0.5f;
    if (diff < 1.0f) return (a + 2.0f) * diff * diff * diff - (a + 3.0f) * diff * diff + 1.0f;
    if (diff < 2.0f) return a * diff * diff * diff - 5.0f * a * diff * diff + 8.0f * a * diff - 4.0f * a;
    return 0.0f;
}
inline float interp_lanczos3(float diff) {   // evaluated in double, rounded once (image.hpp:366-382)
    diff = std::fabs(diff);
    const float l = 3.0f;
    const double pi = std::atan(1.0) * 4;
    if (diff == 0.0f) return 1.0f;
    if (diff < l) return static_cast<float>(l * (std::sin(pi * diff / l) * std::sin(pi * diff)) / (pi * pi * diff * diff));
    return 0.0f;
}
// value of input `v` for output pixel (gx,gy) of an IS of size (isw,ish), through the
// accessor's interpolation mode (dsl/image.hpp:390-422), default boundary CLAMP (:616-620)
template <typename T>
inline float fetch_f(const Img<T> &im, const hb_view &v, int interp, int gx, int gy, int isw, int ish) {
    if (interp == HB_INTERP_NO) return (float)im.at(v.offset_x + gx, v.offset_y + gy);
    const float stride_x = v.width / (float)isw;
    const float stride_y = v.height / (float)ish;
    const float x_mapped = v.offset_x + stride_x / 2 + stride_x * (gx);
    const float y_mapped = v.offset_y + stride_y / 2 + stride_y * (gy);
    if (interp == HB_INTERP_NN) return (float)im.at((int)x_mapped, (int)y_mapped);
    float xb = x_mapped - 0.5f, yb = y_mapped - 0.5f;
    if (xb < 0.0f) xb = 0.0f;
    if (yb < 0.0f) yb = 0.0f;
    const int x_int = (int)xb, y_int = (int)yb;
    const float x_frac = xb - x_int, y_frac = yb - y_int;
    if (interp == HB_INTERP_LF) {
        const float r = (1.0f - x_frac) * (1.0f - y_frac) * (float)im.at(x_int, y_int) +
                        x_frac * (1.0f - y_frac) * (float)im.at(x_int + 1, y_int) +
                        (1.0f - x_frac) * y_frac * (float)im.at(x_int, y_int + 1) +
                        x_frac * y_frac * (float)im.at(x_int + 1, y_int + 1);
        return r;
    }
    // B5 / CF / L3 (dsl/image.hpp:424-528): a TAPS x TAPS neighbourhood starting at (x_int + X0, y_int + Y0); every row is
    // the left-to-right sum of pixel * wx[i], the result the top-to-bottom sum of row * wy[j]
    const int taps = interp == HB_INTERP_L3 ? 6 : 4;
    const int x0 = interp == HB_INTERP_B5 ? 0 : interp == HB_INTERP_CF ? -1 : -2;
    const int y0 = interp == HB_INTERP_B5 ? 0 : -1;   // L3 starts its rows at y_int - 1 like CF (image.hpp:478,484,...)
    float wx[6][6], wy[6];
    float fx = x_frac, fy = y_frac;
    if (interp == HB_INTERP_B5) { fx += 0.5; fy += 0.5; }
    for (int j = 0; j < taps; ++j) {
        for (int i = 0; i < taps; ++i) {
            if (interp == HB_INTERP_B5) wx[j][i] = interp_binomial5(fx - i);
            else if (interp == HB_INTERP_CF) wx[j][i] = interp_bicubic(fx - 1 + i);
            else wx[j][i] = interp_lanczos3(fx - 2 + ((j == 1 && i == 4) ? 5 : i));   // row 1 repeats the weight of tap 5 (image.hpp:489)
        }
        wy[j] = interp == HB_INTERP_B5 ? interp_binomial5(fy - j) : interp == HB_INTERP_CF ? interp_bicubic(fy - 1 + j) : interp_lanczos3(fy - 2 + j);
    }
    float rows[6];
    for (int j = 0; j < taps; ++j) {
        float acc = (float)im.at(x_int + x0, y_int + y0 + j) * wx[j][0];
        for (int i = 1; i < taps; ++i) acc = acc + (float)im.at(x_int + x0 + i, y_int + y0 + j) * wx[j][i];
        rows[j] = acc;
    }
    float r = rows[0] * wy[0];
    for (int j = 1; j < taps; ++j) r = r + rows[j] * wy[j];
    return r;
}
// the accessor returns data_t: interpolated value converted back (convert<T>, dsl/types.hpp:115-117)
template <typename T>
inline T fetch(const Img<T> &im, const hb_view &v, int interp, int gx, int gy, int isw, int ish) {
    if (interp == HB_INTERP_NO) return im.at(v.offset_x + gx, v.offset_y + gy);
    if (interp == HB_INTERP_NN) {
        const float stride_x = v.width / (float)isw, stride_y = v.height / (float)ish;
        const float x_mapped = v.offset_x + stride_x / 2 + stride_x * (gx);
        const float y_mapped = v.offset_y + stride_y / 2 + stride_y * (gy);
        return im.at((int)x_mapped, (int)y_mapped);
    }
    return cast_out<T, float>(fetch_f(im, v, interp, gx, gy, isw, ish));
}

template <typename TI, typename TO>
int point_typed(const hb_point_desc &d) {
    hb_view out = d.out; norm_view(out);
    hb_view in[3]; Img<TI> im[3];
    for (int i = 0; i < d.n_in; ++i) {
        in[i] = d.in[i]; norm_view(in[i]);
        im[i] = make_img<TI>(in[i], HB_BOUNDARY_CLAMP, 0.0);
    }
    TO *op = static_cast<TO *>(out.data);
    const bool isf = std::is_same<TI, float>::value;
    const int W = out.width, H = out.height;
#pragma omp parallel for schedule(static)
    for (int gy = 0; gy < H; ++gy) {
        for (int gx = 0; gx < W; ++gx) {
            TI a = fetch(im[0], in[0], d.interp[0], gx, gy, W, H);
            TI b = d.n_in > 1 ? fetch(im[1], in[1], d.interp[1], gx, gy, W, H) : TI(0);
            TI c = d.n_in > 2 ? fetch(im[2], in[2], d.interp[2], gx, gy, W, H) : TI(0);
            TO o;
            switch (d.op) {
            case HB_POINT_COPY: o = (TO)a; break;
            case HB_POINT_SQUARE: o = isf ? (TO)((float)a * (float)a) : (TO)((int)a * (int)a); break;
            case HB_POINT_MUL: o = isf ? (TO)((float)a * (float)b) : (TO)((int)a * (int)b); break;
            case HB_POINT_SUB: o = isf ? (TO)((float)a - (float)b) : (TO)((int)a - (int)b); break;
            case HB_POINT_ADD: o = isf ? (TO)((float)a + (float)b) : (TO)((int)a + (int)b); break;
            case HB_POINT_BLEND: o = isf ? (TO)((float)a + (float)b / 2) : (TO)((int)a + (int)b / 2); break;
            case HB_POINT_SOBEL_COMBINE: {
                int norm = (int)d.p[0];
                TI in1 = (TI)((int)a / norm), in2 = (TI)((int)b / norm);
                float r = sqrtf((float)((int)in1 * (int)in1 + (int)in2 * (int)in2));
                r = (r < 255.0f ? r : 255.0f);
                r = (r > 0.0f ? r : 0.0f);
                o = cast_out<TO, float>(r);
                break;
            }
            case HB_POINT_HARRIS: {
                int x = (int)a, y = (int)b, xy = (int)c;
                float k = (float)d.p[0], thr = (float)d.p[1];
                float R = ((x * y) - (xy * xy)) - (k * (x + y) * (x + y));
                o = (TO)(R > thr ? 1 : 0);
                break;
            }
            default: o = 0;
            }
            op[(size_t)(out.offset_y + gy) * out.stride + out.offset_x + gx] = o;
        }
    }
    return HB_OK;
}

template <typename TI>
int point_out(const hb_point_desc &d) {
    switch (d.out.dtype) {
    case HB_U8: return point_typed<TI, uchar>(d);
    case HB_S8: return point_typed<TI, signed char>(d);
    case HB_S16: return point_typed<TI, short>(d);
    case HB_S32: return point_typed<TI, int>(d);
    case HB_F32: return point_typed<TI, float>(d);
    default: return HB_ERR_UNSUPPORTED;
    }
}

}  // namespace

extern "C" {

int oc_num_threads(void) { return omp_get_max_threads(); }
void oc_set_num_threads(int n) { omp_set_num_threads(n); }

int oc_local_op(const hb_local_desc *d) {
    if (!d || d->size_x <= 0 || d->size_y <= 0) return HB_ERR_INVALID;
    if (d->tap == HB_TAP_MUL && !d->coef_f32 && !d->coef_s32) return HB_ERR_INVALID;
    switch (d->in.dtype) {
    case HB_U8: return local_acc<uchar>(*d);
    case HB_S8: return local_acc<signed char>(*d);
    case HB_S16: return local_acc<short>(*d);
    case HB_S32: return local_acc<int>(*d);
    case HB_F32: return local_acc<float>(*d);
    default: return HB_ERR_UNSUPPORTED;
    }
}

int oc_bilateral(const hb_bilateral_desc *d) {
    if (!d || d->size <= 0 || !d->coef_f32 || d->in.dtype != d->out.dtype) return HB_ERR_INVALID;
    if (d->in.dtype == HB_U8) return bilateral_typed<uchar>(*d);
    if (d->in.dtype == HB_F32) return bilateral_typed<float>(*d);
    return HB_ERR_UNSUPPORTED;
}

int oc_point_op(const hb_point_desc *d) {
    if (!d || d->n_in < 1 || d->n_in > 3) return HB_ERR_INVALID;
    for (int i = 1; i < d->n_in; ++i) if (d->in[i].dtype != d->in[0].dtype) return HB_ERR_UNSUPPORTED;
    switch (d->in[0].dtype) {
    case HB_U8: return point_out<uchar>(*d);
    case HB_S8: return point_out<signed char>(*d);
    case HB_S16: return point_out<short>(*d);
    case HB_S32: return point_out<int>(*d);
    case HB_F32: return point_out<float>(*d);
    default: return HB_ERR_UNSUPPORTED;
    }
}

// DSL order: strict serial row-major left fold (dsl/kernel.hpp:134-140)
int oc_reduce_serial_f32(const hb_view *v_, int mode, float *result) {
    hb_view v = *v_; norm_view(v);
    const float *p = static_cast<const float *>(v.data);
    float r = p[(size_t)v.offset_y * v.stride + v.offset_x];
    bool first = true;
    for (int y = 0; y < v.height; ++y)
        for (int x = 0; x < v.width; ++x) {
            if (first) { first = false; continue; }
            float e = p[(size_t)(v.offset_y + y) * v.stride + v.offset_x + x];
            if (mode == HB_REDUCE_SUM) r = r + e;
            else if (mode == HB_REDUCE_MIN) r = (r < e ? r : e);
            else if (mode == HB_REDUCE_MAX) r = (r > e ? r : e);
            else r = r * e;
        }
    *result = r;
    return HB_OK;
}

// -emit-cpu order: per-thread partials over row chunks, serial combine (runtime/hipacc_cpu_red.hpp:19-68);
// also returns the float64 sum (the stable target for the float SUM check)
int oc_reduce_minmaxsum_f32(const hb_view *v_, float out[3], double *sum_f64) {
    hb_view v = *v_; norm_view(v);
    const float *p = static_cast<const float *>(v.data);
    const int nthr = omp_get_max_threads();
    std::vector<float> pmin(nthr, std::numeric_limits<float>::infinity()), pmax(nthr, -std::numeric_limits<float>::infinity()), psum(nthr, 0.0f);
    std::vector<double> pd(nthr, 0.0);
    std::vector<char> used(nthr, 0);
#pragma omp parallel for schedule(static)
    for (int y = 0; y < v.height; ++y) {
        const int t = omp_get_thread_num();
        const float *row = p + (size_t)(v.offset_y + y) * v.stride + v.offset_x;
        float mn = pmin[t], mx = pmax[t], sm = psum[t]; double sd = pd[t];
        for (int x = 0; x < v.width; ++x) {
            float e = row[x];
            mn = (mn < e ? mn : e); mx = (mx > e ? mx : e); sm += e; sd += e;
        }
        pmin[t] = mn; pmax[t] = mx; psum[t] = sm; pd[t] = sd; used[t] = 1;
    }
    float mn = std::numeric_limits<float>::infinity(), mx = -mn, sm = 0.0f; double sd = 0.0;
    for (int t = 0; t < nthr; ++t) if (used[t]) { mn = std::min(mn, pmin[t]); mx = std::max(mx, pmax[t]); sm += psum[t]; sd += pd[t]; }
    out[0] = mn; out[1] = mx; out[2] = sm;
    if (sum_f64) *sum_f64 = sd;
    return HB_OK;
}

// -emit-cpu binning: BINNING_CPU_2D (runtime/hipacc_cpu_red.hpp:70-128): per-thread local bins over row
// chunks, then a per-bin combine; the Put helper drops indices >= num_bins (:71-76).  bin(idx) = val with
// reduce = + (Histogram/src/main.cpp:61-67).  C conversion float -> uint like g++/x86-64 (via a 64-bit
// truncation, so negative indices wrap above num_bins and are dropped).
int oc_binning(const hb_binning_desc *d, unsigned *bins) {
    hb_view v = d->in; norm_view(v);
    const int nb = d->num_bins, nthr = omp_get_max_threads();
    std::vector<unsigned> lb((size_t)nthr * nb, 0u);
    const float p0 = (float)d->p0;
#pragma omp parallel for schedule(static)
    for (int y = 0; y < v.height; ++y) {
        unsigned *my = lb.data() + (size_t)omp_get_thread_num() * nb;
        for (int x = 0; x < v.width; ++x) {
            unsigned idx, val;
            if (v.dtype == HB_F32) {
                const float e = static_cast<const float *>(v.data)[(size_t)(v.offset_y + y) * v.stride + v.offset_x + x];
                idx = d->index_kind == HB_BIN_INDEX_SCALE ? (unsigned)(long long)(e / p0 * (float)(unsigned)nb) : (unsigned)(long long)e;
                val = d->value_kind == HB_BIN_VALUE_ONE ? 1u : (unsigned)(long long)e;
            } else if (v.dtype == HB_U8) {
                const unsigned char e = static_cast<const unsigned char *>(v.data)[(size_t)(v.offset_y + y) * v.stride + v.offset_x + x];
                idx = d->index_kind == HB_BIN_INDEX_SCALE ? (unsigned)(long long)((float)e / p0 * (float)(unsigned)nb) : (unsigned)e;
                val = d->value_kind == HB_BIN_VALUE_ONE ? 1u : (unsigned)e;
            } else {
                continue;
            }
            if (idx < (unsigned)nb) my[idx] = my[idx] + val;
        }
    }
    for (int i = 0; i < nb; ++i) {
        unsigned a = 0;
        for (int t = 0; t < nthr; ++t) a = a + lb[(size_t)t * nb + i];
        bins[i] = a;
    }
    return (v.dtype == HB_F32 || v.dtype == HB_U8) ? HB_OK : HB_ERR_UNSUPPORTED;
}

}  // extern "C"
