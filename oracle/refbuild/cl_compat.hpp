// OpenCL C 1.2 -> C++17 compatibility layer, just wide enough to compile the reference's own
// kernel source (src/Graphics/Gudni/OpenCL/Kernels.cl, read where it lies under /root/reference by
// build_ref.py) for the host CPU.  TEST INFRASTRUCTURE, NOT PRODUCT: the resulting
// oracle/_ref/libgudni_ref.so is what pins the restated oracle (oracle/kernels_oracle.cpp) to the
// reference, and the `--impl reference` arm of bench.py.
//
// What is modelled, and how far:
//   * address-space and kernel qualifiers vanish (one flat host address space);
//   * the vector types the source uses (float2/4/8, int2/4, uint4, uchar2/4) with exactly the
//     component names it uses (.x .y .z .w .xy .zw .s0-.s7 .s0123 .s4567), component-wise
//     arithmetic with the scalar widening rule of OpenCL C 6.2.6 (the scalar is converted to the
//     element type, then splatted);
//   * vector literals `(float4)(a,b,c,d)` cannot be spelt in C++ (a cast of a comma expression):
//     build_ref.py rewrites them textually to mk_float4(a,b,c,d); nothing else in an expression
//     is touched;
//   * convert_T (default rounding: rtz for float->int, OpenCL C 6.2.3.3), as_T (bit casts),
//     clz (64 for 0, 6.12.3), mul24, min/max/fabs/fmin/fmax/isinf;
//   * get_global_id reads a thread-local pair the harness sets before each work-item;
//   * floating point is IEEE binary32 with no contraction (the harness is built with
//     -ffp-contract=off): what -cl-fast-relaxed-math (OpenCL/Setup.hs:129) would have changed on a
//     real device is by definition not reproducible.
#pragma once
#include <cfloat>
#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cstring>

typedef unsigned char uchar;
typedef unsigned short ushort;
typedef unsigned int uint;
typedef unsigned long ulong;
static_assert(sizeof(ulong) == 8 && sizeof(uint) == 4, "LP64 expected");

#define __kernel
#define __global
#define __local
#define __private
#define __constant
#define CLK_LOCAL_MEM_FENCE 1
#define CLK_GLOBAL_MEM_FENCE 2

// ---- vector types -------------------------------------------------------------------------------
struct float2 { union { struct { float x, y; }; struct { float s0, s1; }; }; };
struct int2   { union { struct { int x, y; };   struct { int s0, s1; }; }; };
struct uchar2 { union { struct { uchar x, y; }; struct { uchar s0, s1; }; }; };
struct float4 { union { struct { float x, y, z, w; }; struct { float s0, s1, s2, s3; }; struct { float2 xy, zw; }; }; };
struct int4   { union { struct { int x, y, z, w; };   struct { int s0, s1, s2, s3; };   struct { int2 xy, zw; }; }; };
struct uint4  { union { struct { uint x, y, z, w; };  struct { uint s0, s1, s2, s3; }; }; };
struct uchar4 { union { struct { uchar x, y, z, w; }; struct { uchar s0, s1, s2, s3; }; struct { uchar2 xy, zw; }; }; };
struct float8 { union { struct { float s0, s1, s2, s3, s4, s5, s6, s7; }; struct { float4 s0123, s4567; }; }; };
static_assert(sizeof(float2) == 8 && sizeof(float4) == 16 && sizeof(float8) == 32, "vector layout");
static_assert(sizeof(int2) == 8 && sizeof(int4) == 16 && sizeof(uchar4) == 4 && sizeof(uint4) == 16, "vector layout");

// ---- vector literals (targets of build_ref.py's rewrite) ------------------------------------------
template <class A, class B> inline float2 mk_float2(A a, B b) { float2 r; r.x = (float)a; r.y = (float)b; return r; }
template <class A, class B> inline int2 mk_int2(A a, B b) { int2 r; r.x = (int)a; r.y = (int)b; return r; }
template <class A, class B, class C, class D> inline float4 mk_float4(A a, B b, C c, D d) {
    float4 r; r.x = (float)a; r.y = (float)b; r.z = (float)c; r.w = (float)d; return r;
}
template <class A, class B, class C, class D> inline uchar4 mk_uchar4(A a, B b, C c, D d) {
    uchar4 r; r.x = (uchar)a; r.y = (uchar)b; r.z = (uchar)c; r.w = (uchar)d; return r;
}
inline float4 mk_float4(float2 lo, float2 hi) { float4 r; r.xy = lo; r.zw = hi; return r; }
template <class A> inline float4 mk_float4(A a) { return mk_float4(a, a, a, a); }   // (float4)(s): splat, OpenCL C 6.2.2
inline float8 mk_float8(float4 lo, float4 hi) { float8 r; r.s0123 = lo; r.s4567 = hi; return r; }

// ---- component-wise arithmetic ------------------------------------------------------------------
#define CL_BINOP2(V, E, op)                                                                    \
    inline V operator op(V a, V b) { V r; r.x = a.x op b.x; r.y = a.y op b.y; return r; }       \
    inline V operator op(V a, E s) { V r; r.x = a.x op s; r.y = a.y op s; return r; }           \
    inline V operator op(E s, V b) { V r; r.x = s op b.x; r.y = s op b.y; return r; }           \
    inline V& operator op##=(V& a, V b) { a = a op b; return a; }                               \
    inline V& operator op##=(V& a, E s) { a = a op s; return a; }
#define CL_BINOP4(V, E, op)                                                                                        \
    inline V operator op(V a, V b) { V r; r.x = a.x op b.x; r.y = a.y op b.y; r.z = a.z op b.z; r.w = a.w op b.w; return r; } \
    inline V operator op(V a, E s) { V r; r.x = a.x op s; r.y = a.y op s; r.z = a.z op s; r.w = a.w op s; return r; }         \
    inline V operator op(E s, V b) { V r; r.x = s op b.x; r.y = s op b.y; r.z = s op b.z; r.w = s op b.w; return r; }         \
    inline V& operator op##=(V& a, V b) { a = a op b; return a; }                                                  \
    inline V& operator op##=(V& a, E s) { a = a op s; return a; }
#define CL_ARITH(M, V, E) M(V, E, +) M(V, E, -) M(V, E, *) M(V, E, /)
CL_ARITH(CL_BINOP2, float2, float)
CL_ARITH(CL_BINOP2, int2, int)
CL_ARITH(CL_BINOP4, float4, float)
CL_ARITH(CL_BINOP4, int4, int)
inline float2 operator-(float2 a) { return mk_float2(-a.x, -a.y); }
inline float4 operator-(float4 a) { return mk_float4(-a.x, -a.y, -a.z, -a.w); }
#define CL_BINOP8(op)                                                                                   \
    inline float8 operator op(float8 a, float8 b) { return mk_float8(a.s0123 op b.s0123, a.s4567 op b.s4567); } \
    inline float8 operator op(float8 a, float s) { return mk_float8(a.s0123 op s, a.s4567 op s); }      \
    inline float8 operator op(float s, float8 b) { return mk_float8(s op b.s0123, s op b.s4567); }      \
    inline float8& operator op##=(float8& a, float8 b) { a = a op b; return a; }
CL_BINOP8(+) CL_BINOP8(-) CL_BINOP8(*) CL_BINOP8(/)

// ---- conversions (OpenCL C 6.2.3: float -> integer rounds toward zero by default) -------------------
inline float convert_float(int v) { return (float)v; }
inline float2 convert_float2(int2 v) { return mk_float2((float)v.x, (float)v.y); }
inline float4 convert_float4(uchar4 v) { return mk_float4((float)v.x, (float)v.y, (float)v.z, (float)v.w); }
inline int2 convert_int2(float2 v) { return mk_int2((int)v.x, (int)v.y); }
// out-of-range input is undefined in OpenCL C (no _sat); this layer does what the one real OpenCL
// device the kernels were run on does (NVIDIA OpenCL 3.0, B200: cvt.rzi.u8.f32 clamps, NaN -> 0)
inline uchar cl_f2uchar_rtz(float v) { return v >= 255.0f ? (uchar)255 : (v > 0.0f ? (uchar)(int)v : (uchar)0); }
inline uchar4 convert_uchar4(float4 v) {
    return mk_uchar4(cl_f2uchar_rtz(v.x), cl_f2uchar_rtz(v.y), cl_f2uchar_rtz(v.z), cl_f2uchar_rtz(v.w));
}

// ---- reinterpretation ---------------------------------------------------------------------------
template <class To, class From> inline To cl_bitcast(From f) {
    static_assert(sizeof(To) == sizeof(From), "as_T needs equal sizes");
    To t; std::memcpy(&t, &f, sizeof t); return t;
}
template <class From> inline uint as_uint(From f) { return cl_bitcast<uint>(f); }
template <class From> inline ushort as_ushort(From f) { return cl_bitcast<ushort>(f); }
template <class From> inline uint4 as_uint4(From f) { return cl_bitcast<uint4>(f); }
template <class From> inline uchar4 as_uchar4(From f) { return cl_bitcast<uchar4>(f); }
template <class From> inline float as_float(From f) { return cl_bitcast<float>(f); }

// ---- integer / common built-ins ---------------------------------------------------------------------
inline int clz(ulong v) { return v ? __builtin_clzl(v) : 64; }
inline int clz(uint v) { return v ? __builtin_clz(v) : 32; }
inline int mul24(int a, int b) { return a * b; }
inline int min(int a, int b) { return b < a ? b : a; }
inline int max(int a, int b) { return a < b ? b : a; }
inline uint min(uint a, uint b) { return b < a ? b : a; }
inline uint max(uint a, uint b) { return a < b ? b : a; }
// OpenCL C 6.12.4: fmin-like semantics are NOT required of min/max on floats; "y < x ? y : x"
inline float min(float a, float b) { return b < a ? b : a; }
inline float max(float a, float b) { return a < b ? b : a; }
inline void barrier(int) {}
using std::isinf;

// ---- work-item functions ------------------------------------------------------------------------
extern thread_local int cl_global_id[3];
inline int get_global_id(int dim) { return cl_global_id[dim]; }
