// hipacc_types.hpp -- the DSL's vector pixel types and their operator set (dsl/types.hpp:56-516, device side
// runtime/hipacc_types.hpp), usable from the host compiler and from nvcc (host + device).
//
//   char4 uchar4 short4 ushort4 int4 uint4 float4      4 channels x, y, z, w (what the samples' RGBA images use)
//   make_<T>4(x, y, z, w), make_<T>4(s)                 construct / broadcast
//   a op b, a op s, s op a                              element-wise + - * / for all types; % & | ^ << >> for integers;
//                                                       results are converted back to the element type (uchar4 wraps)
//   a op= b                                             element-wise;   a op= s takes `a` BY VALUE in the reference
//                                                       (dsl/types.hpp:146-148) and therefore has no effect -- kept,
//                                                       because it decides results (Laplace_RGBA's `sum += 128`)
//   -a, ~a, comparisons (element-wise, -1 / 0 in the signed integer vector of the same width, OpenCL style)
//   convert_<T>4(v)                                     element-wise C conversion
//
// Written with templates over a traits class instead of the reference's macro expansion; under nvcc the types are
// CUDA's own built-in vector structs (same layout), under a host compiler they are defined here.
#ifndef HIPACC_B200_TYPES_HPP
#define HIPACC_B200_TYPES_HPP

#include <type_traits>

#ifdef __CUDACC__
#include <vector_types.h>
#include <vector_functions.h>
#define HB_HD __host__ __device__ inline
#else
#define HB_HD inline
#endif

#ifndef HIPACC_B200_NO_TYPEDEFS
typedef unsigned char uchar;
typedef unsigned short ushort;
typedef unsigned int uint;
#endif

#ifndef __CUDACC__
// layout of CUDA's vector_types.h: size = 4 elements, alignment = min(size, 16)
struct alignas(4) char4 { signed char x, y, z, w; };
struct alignas(4) uchar4 { unsigned char x, y, z, w; };
struct alignas(8) short4 { short x, y, z, w; };
struct alignas(8) ushort4 { unsigned short x, y, z, w; };
struct alignas(16) int4 { int x, y, z, w; };
struct alignas(16) uint4 { unsigned int x, y, z, w; };
struct alignas(16) float4 { float x, y, z, w; };
#endif

namespace hipacc_b200 {
template <typename V> struct vec4 { static constexpr bool is = false; };
#define HB_DECL_VEC4(V, E, C)                      \
    template <> struct vec4<V> {                   \
        static constexpr bool is = true;           \
        typedef E elem;                            \
        typedef C cmp; /* result of comparisons */ \
    };
HB_DECL_VEC4(char4, signed char, char4)
HB_DECL_VEC4(uchar4, unsigned char, char4)
HB_DECL_VEC4(short4, short, short4)
HB_DECL_VEC4(ushort4, unsigned short, short4)
HB_DECL_VEC4(int4, int, int4)
HB_DECL_VEC4(uint4, unsigned int, int4)
HB_DECL_VEC4(float4, float, int4)
#undef HB_DECL_VEC4
template <typename V> using elem_t = typename vec4<V>::elem;
template <typename V> using cmp_t = typename vec4<V>::cmp;
template <typename V> using if_vec = typename std::enable_if<vec4<V>::is, int>::type;
template <typename V> using if_ivec = typename std::enable_if<vec4<V>::is && std::is_integral<typename vec4<V>::elem>::value, int>::type;

template <typename V, typename A, typename B, typename C, typename D> HB_HD V mk(A x, B y, C z, D w) {
    V v;
    v.x = (elem_t<V>)x; v.y = (elem_t<V>)y; v.z = (elem_t<V>)z; v.w = (elem_t<V>)w;
    return v;
}
}  // namespace hipacc_b200

// ---- constructors ---------------------------------------------------------------------------------------------------
#define HB_MAKE_FUNCS(V, E)                                                           \
    HB_HD V make_##V(E s) { return hipacc_b200::mk<V>(s, s, s, s); }
#ifndef __CUDACC__
#define HB_MAKE_FUNCS4(V, E) HB_HD V make_##V(E x, E y, E z, E w) { return hipacc_b200::mk<V>(x, y, z, w); }
#else
#define HB_MAKE_FUNCS4(V, E)   /* vector_functions.h */
#endif
HB_MAKE_FUNCS(char4, signed char) HB_MAKE_FUNCS4(char4, signed char)
HB_MAKE_FUNCS(uchar4, unsigned char) HB_MAKE_FUNCS4(uchar4, unsigned char)
HB_MAKE_FUNCS(short4, short) HB_MAKE_FUNCS4(short4, short)
HB_MAKE_FUNCS(ushort4, unsigned short) HB_MAKE_FUNCS4(ushort4, unsigned short)
HB_MAKE_FUNCS(int4, int) HB_MAKE_FUNCS4(int4, int)
HB_MAKE_FUNCS(uint4, unsigned int) HB_MAKE_FUNCS4(uint4, unsigned int)
HB_MAKE_FUNCS(float4, float) HB_MAKE_FUNCS4(float4, float)
#undef HB_MAKE_FUNCS
#undef HB_MAKE_FUNCS4

// ---- element-wise binary operators: V op V, V op s, s op V ------------------------------------------------------------
#define HB_VEC_BINOP(OP, GUARD)                                                                                          \
    template <typename V, hipacc_b200::GUARD<V> = 0> HB_HD V operator OP(V a, V b) {                                     \
        return hipacc_b200::mk<V>(a.x OP b.x, a.y OP b.y, a.z OP b.z, a.w OP b.w);                                       \
    }                                                                                                                    \
    template <typename V, hipacc_b200::GUARD<V> = 0> HB_HD V operator OP(V a, hipacc_b200::elem_t<V> b) {                \
        return hipacc_b200::mk<V>(a.x OP b, a.y OP b, a.z OP b, a.w OP b);                                               \
    }                                                                                                                    \
    template <typename V, hipacc_b200::GUARD<V> = 0> HB_HD V operator OP(hipacc_b200::elem_t<V> a, V b) {                \
        return hipacc_b200::mk<V>(a OP b.x, a OP b.y, a OP b.z, a OP b.w);                                               \
    }                                                                                                                    \
    template <typename V, hipacc_b200::GUARD<V> = 0> HB_HD void operator OP##=(V &a, V b) {                              \
        a.x OP## = b.x; a.y OP## = b.y; a.z OP## = b.z; a.w OP## = b.w;                                                  \
    }                                                                                                                    \
    /* the reference takes `a` by value here (dsl/types.hpp:146-148): no effect */                                       \
    template <typename V, hipacc_b200::GUARD<V> = 0> HB_HD void operator OP##=(V, hipacc_b200::elem_t<V>) {}
HB_VEC_BINOP(+, if_vec)
HB_VEC_BINOP(-, if_vec)
HB_VEC_BINOP(*, if_vec)
HB_VEC_BINOP(/, if_vec)
HB_VEC_BINOP(%, if_ivec)
HB_VEC_BINOP(&, if_ivec)
HB_VEC_BINOP(|, if_ivec)
HB_VEC_BINOP(^, if_ivec)
HB_VEC_BINOP(<<, if_ivec)
HB_VEC_BINOP(>>, if_ivec)
#undef HB_VEC_BINOP

template <typename V, hipacc_b200::if_vec<V> = 0> HB_HD V operator-(V a) { return hipacc_b200::mk<V>(-a.x, -a.y, -a.z, -a.w); }
template <typename V, hipacc_b200::if_vec<V> = 0> HB_HD V operator+(V a) { return a; }
template <typename V, hipacc_b200::if_ivec<V> = 0> HB_HD V operator~(V a) { return hipacc_b200::mk<V>(~a.x, ~a.y, ~a.z, ~a.w); }

// ---- comparisons: -1 (all bits) where true, 0 where false, in the signed integer vector of the same width -------------
#define HB_VEC_CMP(OP)                                                                                                            \
    template <typename V, hipacc_b200::if_vec<V> = 0> HB_HD hipacc_b200::cmp_t<V> operator OP(V a, V b) {                         \
        return hipacc_b200::mk<hipacc_b200::cmp_t<V>>(-(a.x OP b.x), -(a.y OP b.y), -(a.z OP b.z), -(a.w OP b.w));                \
    }                                                                                                                             \
    template <typename V, hipacc_b200::if_vec<V> = 0> HB_HD hipacc_b200::cmp_t<V> operator OP(V a, hipacc_b200::elem_t<V> b) {    \
        return hipacc_b200::mk<hipacc_b200::cmp_t<V>>(-(a.x OP b), -(a.y OP b), -(a.z OP b), -(a.w OP b));                        \
    }                                                                                                                             \
    template <typename V, hipacc_b200::if_vec<V> = 0> HB_HD hipacc_b200::cmp_t<V> operator OP(hipacc_b200::elem_t<V> a, V b) {    \
        return hipacc_b200::mk<hipacc_b200::cmp_t<V>>(-(a OP b.x), -(a OP b.y), -(a OP b.z), -(a OP b.w));                        \
    }
HB_VEC_CMP(==)
HB_VEC_CMP(!=)
HB_VEC_CMP(<)
HB_VEC_CMP(<=)
HB_VEC_CMP(>)
HB_VEC_CMP(>=)
#undef HB_VEC_CMP

// ---- conversions ------------------------------------------------------------------------------------------------------
#define HB_VEC_CONVERT(V)                                                                     \
    template <typename S, hipacc_b200::if_vec<S> = 0> HB_HD V convert_##V(S v) {              \
        return hipacc_b200::mk<V>(v.x, v.y, v.z, v.w);                                        \
    }
HB_VEC_CONVERT(char4)
HB_VEC_CONVERT(uchar4)
HB_VEC_CONVERT(short4)
HB_VEC_CONVERT(ushort4)
HB_VEC_CONVERT(int4)
HB_VEC_CONVERT(uint4)
HB_VEC_CONVERT(float4)
#undef HB_VEC_CONVERT

// ---- as_float / convert<T> (dsl/types.hpp:103-125): what the interpolating accessors compute in -------------------------
namespace hipacc_b200 {
template <typename T, typename = void> struct float_of { typedef float type; };
template <typename T> struct float_of<T, typename std::enable_if<vec4<T>::is>::type> { typedef float4 type; };
template <typename T, typename std::enable_if<!vec4<T>::is, int>::type = 0> HB_HD float to_float(T v) { return (float)v; }
template <typename V, if_vec<V> = 0> HB_HD float4 to_float(V v) { return mk<float4>(v.x, v.y, v.z, v.w); }
template <typename T, typename std::enable_if<!vec4<T>::is, int>::type = 0> HB_HD T from_float(float v) { return (T)v; }
template <typename V, if_vec<V> = 0> HB_HD V from_float(float4 v) { return mk<V>(v.x, v.y, v.z, v.w); }
}  // namespace hipacc_b200

// ---- element-wise math of hipacc::math on vectors (dsl/math_functions.hpp): min, max with vector or scalar bound, the
// float functions on float4.  Declared at global scope like the reference's; hipacc::math pulls them in. ---------------
#define HB_VEC_MINMAX(NAME, CMP)                                                                                        \
    template <typename V, hipacc_b200::if_vec<V> = 0> HB_HD V NAME(V a, V b) {                                          \
        return hipacc_b200::mk<V>(b.x CMP a.x ? b.x : a.x, b.y CMP a.y ? b.y : a.y, b.z CMP a.z ? b.z : a.z, b.w CMP a.w ? b.w : a.w); \
    }                                                                                                                   \
    template <typename V, hipacc_b200::if_vec<V> = 0> HB_HD V NAME(V a, hipacc_b200::elem_t<V> b) {                     \
        return hipacc_b200::mk<V>(b CMP a.x ? b : a.x, b CMP a.y ? b : a.y, b CMP a.z ? b : a.z, b CMP a.w ? b : a.w);  \
    }
namespace hipacc_b200 { namespace vmath {
HB_VEC_MINMAX(min, <)
HB_VEC_MINMAX(max, >)
}}  // namespace hipacc_b200::vmath
#undef HB_VEC_MINMAX

#endif  // HIPACC_B200_TYPES_HPP
