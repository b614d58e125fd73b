// cl_shim.hpp -- just enough of OpenCL C, on plain C++, to EXECUTE the reference's own kernel source
// (/root/reference/super_resolution/raisr.cl) on the CPU.  TEST INFRASTRUCTURE (see oracle/raisr_oracle.c for the
// rules): nothing under oclcomputervision_b200/ may use it.
//
// The reference cannot run as shipped (no OpenCL platform), so its kernel text had never produced a
// single pixel anywhere near this repository and the oracle was "unpinned".  oracle/build_ref.py takes the kernel
// text where it lies, applies four mechanical spelling rewrites that C++ needs (OpenCL vector literals `(half4)(a,b)`
// -> `half4(a,b)`, half literals `0.5h` -> `half(0.5f)`, the swizzle `.s210` -> `.s210()`, `__local` parameters), puts
// the one `#if 1` of the early return under a macro, and compiles the result against this header into
// oracle/_ref/libraisr_ref_*.so.  A work-group is 16x16 host threads that really run the kernel function concurrently,
// `__local` variables are function statics shared by them, barrier() is a pthread barrier.
//
// Arithmetic model of `half` (cl_khr_fp16): every operation is computed in binary32 and rounded to binary16
// (round-to-nearest-even), i.e. what a device with native fp16 ALUs produces for + - * / (exactly: binary32 carries
// more than 2p+2 bits) and a faithful stand-in for sqrt / atan2.  dot() accumulates left to right in half.  Mixed
// operands follow the usual arithmetic conversions of OpenCL C: half with an integer -> half, with float -> float,
// with double -> double.
#pragma once
#include <algorithm>
#include <cmath>
#include <cstdint>
#include <pthread.h>

#define __kernel
#define __global
#define __constant static const
#define __local static
#define __inline inline
#define read_only
#define write_only
#define CLK_LOCAL_MEM_FENCE 1
#define CLK_NORMALIZED_COORDS_FALSE 0
#define CLK_NORMALIZED_COORDS_TRUE 1
#define CLK_ADDRESS_CLAMP_TO_EDGE 2
#define CLK_FILTER_NEAREST 0
#define CLK_FILTER_LINEAR 0x10
#define CLK_R 0x10B0
#define CLK_RGBA 0x10B5
#define CLK_BGRA 0x10B6
#define CLK_ARGB 0x10B7

// -DCL_SHIM_HALF_IS_FLOAT: `half` keeps binary32 -- the kernel TEXT evaluated in single precision, which is the
// arithmetic the oracle restates; without it, true binary16 as on a cl_khr_fp16 device.
#ifdef CL_SHIM_HALF_IS_FLOAT
typedef float cl_half_store;
#else
typedef _Float16 cl_half_store;
#endif
struct half {
    cl_half_store v;
    half() : v((cl_half_store)0.0f) {}
    half(float f) : v((cl_half_store)f) {}
    half(double f) : v((cl_half_store)f) {}
    half(int i) : v((cl_half_store)(float)i) {}
    explicit operator float() const { return (float)v; }
    float f() const { return (float)v; }
};
inline half operator+(half a, half b) { return half(a.f() + b.f()); }
inline half operator-(half a, half b) { return half(a.f() - b.f()); }
inline half operator*(half a, half b) { return half(a.f() * b.f()); }
inline half operator/(half a, half b) { return half(a.f() / b.f()); }
inline half operator-(half a) { return half(-a.f()); }
// half with an integer -> half
inline half operator+(half a, int b) { return a + half(b); }
inline half operator+(int a, half b) { return half(a) + b; }
inline half operator-(half a, int b) { return a - half(b); }
inline half operator-(int a, half b) { return half(a) - b; }
inline half operator*(half a, int b) { return a * half(b); }
inline half operator*(int a, half b) { return half(a) * b; }
inline half operator/(half a, int b) { return a / half(b); }
// half with float -> float, with double -> double
inline float operator+(half a, float b) { return a.f() + b; }
inline float operator+(float a, half b) { return a + b.f(); }
inline float operator-(half a, float b) { return a.f() - b; }
inline float operator-(float a, half b) { return a - b.f(); }
inline float operator*(half a, float b) { return a.f() * b; }
inline float operator*(float a, half b) { return a * b.f(); }
inline float operator/(half a, float b) { return a.f() / b; }
inline double operator+(half a, double b) { return (double)a.f() + b; }
inline double operator-(half a, double b) { return (double)a.f() - b; }
inline double operator*(half a, double b) { return (double)a.f() * b; }
inline double operator/(half a, double b) { return (double)a.f() / b; }
inline half& operator+=(half& a, half b) { a = a + b; return a; }
inline half& operator+=(half& a, double b) { a = half(a + b); return a; }   // theta += PI
inline bool operator<(half a, half b) { return a.f() < b.f(); }
inline bool operator<(half a, int b) { return a.f() < (float)b; }
inline bool operator<(half a, float b) { return a.f() < b; }
inline bool operator!=(half a, int b) { return a.f() != (float)b; }
inline half sqrt(half a) { return half(std::sqrt(a.f())); }
inline half atan2(half y, half x) { return half(std::atan2(y.f(), x.f())); }
inline half clamp(half v, half lo, half hi) { return v.f() < lo.f() ? lo : (v.f() > hi.f() ? hi : v); }
inline int clamp(int v, int lo, int hi) { return v < lo ? lo : (v > hi ? hi : v); }
using std::ceil;
using std::floor;
using std::min;

typedef unsigned char uchar;
typedef unsigned short ushort;
typedef unsigned int uint;
inline float clamp(float v, float lo, float hi) { return v < lo ? lo : (v > hi ? hi : v); }
struct int2 { int x, y; int2() : x(0), y(0) {} int2(int a, int b) : x(a), y(b) {} };
struct uint4 { uint x, y, z, w; uint4() : x(0), y(0), z(0), w(0) {} uint4(uint a, uint b, uint c, uint d) : x(a), y(b), z(c), w(d) {} };
// ushortN of hist.cl:3-38 (only N = 8 is built: HIST_BINS / HIST_THREAD_NUM = 256 / 32, eq_opencl.py:26); arithmetic wraps at 16 bits
struct ushort8 {
    ushort s0, s1, s2, s3, s4, s5, s6, s7;
    ushort8(int v = 0) : s0(v), s1(v), s2(v), s3(v), s4(v), s5(v), s6(v), s7(v) {}
    ushort8& operator+=(const ushort8& o)
    {
        s0 += o.s0; s1 += o.s1; s2 += o.s2; s3 += o.s3; s4 += o.s4; s5 += o.s5; s6 += o.s6; s7 += o.s7;
        return *this;
    }
};
inline ushort8 vload8(int n, const ushort* p)
{
    ushort8 v;
    p += 8 * n;
    v.s0 = p[0]; v.s1 = p[1]; v.s2 = p[2]; v.s3 = p[3]; v.s4 = p[4]; v.s5 = p[5]; v.s6 = p[6]; v.s7 = p[7];
    return v;
}
struct float2 {
    float x, y;
    float2() : x(0), y(0) {}
    float2(float a, float b) : x(a), y(b) {}
};
inline float2 operator/(float2 a, float2 b) { return float2(a.x / b.x, a.y / b.y); }
inline float2 operator*(float2 a, float2 b) { return float2(a.x * b.x, a.y * b.y); }
inline float2 convert_float2(int2 a) { return float2((float)a.x, (float)a.y); }
struct float3 { float x, y, z; };
struct float4 {
    float x, y, z, w;
    float4() : x(0), y(0), z(0), w(0) {}
    float4(float a, float b, float c, float d) : x(a), y(b), z(c), w(d) {}
};
inline float4 vload4(int n, const float* p) { return float4(p[4 * n], p[4 * n + 1], p[4 * n + 2], p[4 * n + 3]); }
inline float3 vload3(int n, const float* p) { return float3{p[3 * n], p[3 * n + 1], p[3 * n + 2]}; }

struct half2 {
    half x, y;
    half2() {}
    half2(half a, half b) : x(a), y(b) {}
};
inline half2 sqrt(half2 a) { return half2(sqrt(a.x), sqrt(a.y)); }
struct half3 {
    half x, y, z;
    half3() {}
    half3(half a, half b, half c) : x(a), y(b), z(c) {}
    half3 s210() const { return half3(z, y, x); }
};
inline half3 convert_half3(float3 a) { return half3(half(a.x), half(a.y), half(a.z)); }
inline half dot(half3 a, half3 b) { return a.x * b.x + a.y * b.y + a.z * b.z; }
struct half4 {
    half x, y, z, w;
    half4() {}
    half4(half a, half b, half c, half d) : x(a), y(b), z(c), w(d) {}
};
inline half4 operator+(half4 a, half4 b) { return half4(a.x + b.x, a.y + b.y, a.z + b.z, a.w + b.w); }
inline half4 operator*(half4 a, half b) { return half4(a.x * b, a.y * b, a.z * b, a.w * b); }
inline half4 operator*(half a, half4 b) { return half4(a * b.x, a * b.y, a * b.z, a * b.w); }
inline half4& operator+=(half4& a, half4 b) { a = a + b; return a; }
inline half4 clamp(half4 v, half lo, half hi) { return half4(clamp(v.x, lo, hi), clamp(v.y, lo, hi), clamp(v.z, lo, hi), clamp(v.w, lo, hi)); }
inline half4 convert_half4(float4 a) { return half4(half(a.x), half(a.y), half(a.z), half(a.w)); }
inline float4 convert_float4(half4 a) { return float4(a.x.f(), a.y.f(), a.z.f(), a.w.f()); }
inline half dot(half4 a, half4 b) { return a.x * b.x + a.y * b.y + a.z * b.z + a.w * b.w; }

// ---- images (CL_R / CL_BGRA, CL_UNORM_INT8) and samplers
typedef int sampler_t;
struct image2d {
    uint8_t* data;
    int w, h, pitch, order;   // pitch in bytes
};
typedef image2d* image2d_t;
inline int get_image_width(image2d_t im) { return im->w; }
inline int get_image_height(image2d_t im) { return im->h; }
inline int get_image_channel_order(image2d_t im) { return im->order; }
// CLK_NORMALIZED_COORDS_FALSE | CLK_ADDRESS_CLAMP_TO_EDGE | CLK_FILTER_NEAREST with integer coordinates; UNORM_INT8
// decodes to v / 255; a CL_R image returns (r, 0, 0, 1), a CL_BGRA image its components in (r, g, b, a) order
inline float4 read_imagef(image2d_t im, sampler_t, int2 c)
{
    const int x = std::min(std::max(c.x, 0), im->w - 1), y = std::min(std::max(c.y, 0), im->h - 1);
    const uint8_t* p = im->data + (size_t)y * im->pitch;
    if (im->order == CLK_R) return float4(p[x] / 255.0f, 0.0f, 0.0f, 1.0f);
    p += 4 * x;
    return float4(p[2] / 255.0f, p[1] / 255.0f, p[0] / 255.0f, p[3] / 255.0f);
}
// float coordinates: only what interpolation.cl:12-13 uses -- CLK_NORMALIZED_COORDS_TRUE | CLK_ADDRESS_CLAMP_TO_EDGE |
// CLK_FILTER_LINEAR, evaluated as the OpenCL 1.2 specification writes it (section 8.2): u = s * w, i0 = floor(u - 0.5),
// a = frac(u - 0.5), T = (1-a)(1-b) T00 + a (1-b) T10 + (1-a) b T01 + a b T11 in single precision.  (Hardware samplers
// use fixed-point weights of unspecified width, so this variant is a specification reference, not a device's bits.)
inline float4 read_imagef(image2d_t im, sampler_t smp, float2 c)
{
    float u = c.x, v = c.y;
    if (smp & CLK_NORMALIZED_COORDS_TRUE) { u *= (float)im->w; v *= (float)im->h; }
    if (!(smp & CLK_FILTER_LINEAR)) return read_imagef(im, smp, int2((int)std::floor(u), (int)std::floor(v)));
    const float fu = u - 0.5f, fv = v - 0.5f;
    const int i0 = (int)std::floor(fu), j0 = (int)std::floor(fv);
    const float a = fu - std::floor(fu), b = fv - std::floor(fv);
    const float4 t00 = read_imagef(im, smp, int2(i0, j0)), t10 = read_imagef(im, smp, int2(i0 + 1, j0));
    const float4 t01 = read_imagef(im, smp, int2(i0, j0 + 1)), t11 = read_imagef(im, smp, int2(i0 + 1, j0 + 1));
    const float w00 = (1.0f - a) * (1.0f - b), w10 = a * (1.0f - b), w01 = (1.0f - a) * b, w11 = a * b;
    auto mix = [&](float p00, float p10, float p01, float p11) { return ((w00 * p00 + w10 * p10) + w01 * p01) + w11 * p11; };
    return float4(mix(t00.x, t10.x, t01.x, t11.x), mix(t00.y, t10.y, t01.y, t11.y), mix(t00.z, t10.z, t01.z, t11.z),
                  mix(t00.w, t10.w, t01.w, t11.w));
}
// CL_UNSIGNED_INT8 images (hist.cl): raw bytes in, saturated bytes out; a CL_R image returns (r, 0, 0, 1)
inline uint4 read_imageui(image2d_t im, sampler_t, int2 c)
{
    const int x = std::min(std::max(c.x, 0), im->w - 1), y = std::min(std::max(c.y, 0), im->h - 1);
    return uint4(im->data[(size_t)y * im->pitch + x], 0, 0, 1);
}
inline void write_imageui(image2d_t im, int2 c, uint4 v)
{
    if (c.x < 0 || c.y < 0 || c.x >= im->w || c.y >= im->h) return;
    im->data[(size_t)c.y * im->pitch + c.x] = (uint8_t)std::min(v.x, 255u);
}
inline uint8_t cl_unorm8(float v)
{
    if (!(v == v)) return 0;                       // NaN converts to 0
    v = std::min(std::max(v, 0.0f), 1.0f);
    return (uint8_t)std::nearbyint(v * 255.0f);    // round to nearest even (default rounding mode)
}
inline void write_imagef(image2d_t im, int2 c, float4 v)
{
    if (c.x < 0 || c.y < 0 || c.x >= im->w || c.y >= im->h) return;
    uint8_t* p = im->data + (size_t)c.y * im->pitch;
    if (im->order == CLK_R) { p[c.x] = cl_unorm8(v.x); return; }
    p += 4 * c.x;
    p[2] = cl_unorm8(v.x); p[1] = cl_unorm8(v.y); p[0] = cl_unorm8(v.z); p[3] = cl_unorm8(v.w);
}

// ---- work-item functions: one host thread per work-item of the current work-group
struct cl_item { int gid[2], lid[2], grp[2], lsz[2], gsz[2], ngrp[2]; };
extern thread_local cl_item cl_self;
extern pthread_barrier_t* cl_group_barrier;
inline int get_global_id(int d) { return cl_self.gid[d]; }
inline int get_local_id(int d) { return cl_self.lid[d]; }
inline int get_group_id(int d) { return cl_self.grp[d]; }
inline int get_local_size(int d) { return cl_self.lsz[d]; }
inline int get_global_size(int d) { return cl_self.gsz[d]; }
inline int get_num_groups(int d) { return cl_self.ngrp[d]; }
inline void barrier(int) { pthread_barrier_wait(cl_group_barrier); }
