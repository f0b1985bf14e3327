// strolle_b200 — device maths.
//
// Strict IEEE-754 binary32 arithmetic: this translation unit is compiled with
// -fmad=false (no FMA contraction), default -prec-div/-prec-sqrt, no fast-math,
// so every + - * / sqrt below is one correctly-rounded operation, evaluated in
// the order written.  Vector helpers follow glam 0.24's scalar formulas
// (dot = x*x + y*y + z*z left to right, normalize = v * (1/len), ...), which is
// what strolle-gpu's arithmetic is built from.  Elementary functions (sin, cos,
// acos, atan2, exp, pow) are explicit Cephes-style polynomial kernels instead of
// libdevice calls so that a frame is reproducible bit-for-bit against a CPU run
// of the same formulas.
//
// This header (and st_device.cuh, kernels.cu) is compiled TWICE:
//   ST_FAST undefined  -> namespace st : the strict flavour described above (nvcc -fmad=false).
//   ST_FAST = 1        -> namespace stf: the "fast shading" flavour of the ReSTIR kernels (nvcc -fmad=true -prec-div=false
//                         -prec-sqrt=false): FMA contraction, div/sqrt/rcp approximations and SFU sin/cos/ex2/lg2 for radiance,
//                         BRDF, pdf and MIS evaluation — what a GPU shader compiler emits for the reference's rust-gpu SPIR-V.
// Everything that DECIDES something discrete is written with the x*() primitives below (explicit round-to-nearest
// intrinsics that no compiler flag contracts or approximates): BVH traversal, ray/box and ray/triangle tests, the alpha test's
// texel address, RNG draws.  Both flavours therefore return the same hit for the same ray, bit for bit.
#pragma once
#include <cstdint>
#include <cuda_runtime.h>

#if defined(ST_FAST) && ST_FAST
#define ST_NS stf
#else
#define ST_NS st
#endif

namespace ST_NS {

typedef uint32_t u32;
typedef int32_t i32;

#define ST_DEV __device__ __forceinline__

static const float kPi = 3.14159265358979323846f;
static const float kHalfPi = 1.5707963267948966f;
static const float kF32Max = 3.40282347e+38f;
static const float kF32Eps = 1.1920929e-7f;

// flag-independent IEEE-754 primitives: one correctly rounded operation each, never contracted into an FMA
ST_DEV float xadd(float a, float b) { return __fadd_rn(a, b); }
ST_DEV float xsub(float a, float b) { return __fsub_rn(a, b); }
ST_DEV float xmul(float a, float b) { return __fmul_rn(a, b); }
ST_DEV float xdiv(float a, float b) { return __fdiv_rn(a, b); }
ST_DEV float xsqrt(float a) { return __fsqrt_rn(a); }

ST_DEV u32 fbits(float f) { return __float_as_uint(f); }
ST_DEV float bitsf(u32 u) { return __uint_as_float(u); }
ST_DEV float finf() { return __uint_as_float(0x7f800000u); }
ST_DEV float fnan() { return __uint_as_float(0x7fc00000u); }

// Rust f32::min/max (NaN-ignoring), spelled with comparisons so that signed
// zeros behave the same on every platform.
ST_DEV float rmin(float a, float b) { return (a != a) ? b : ((b != b) ? a : (a < b ? a : b)); }
ST_DEV float rmax(float a, float b) { return (a != a) ? b : ((b != b) ? a : (a > b ? a : b)); }
ST_DEV float rclamp(float x, float lo, float hi) { if (x < lo) x = lo; if (x > hi) x = hi; return x; }
ST_DEV float sat(float x) { return rclamp(x, 0.0f, 1.0f); }
ST_DEV float sq(float x) { return x * x; }
ST_DEV float cpsign(float mag, float sgn) { return bitsf((fbits(mag) & 0x7fffffffu) | (fbits(sgn) & 0x80000000u)); }
ST_DEV float fabs_(float x) { return bitsf(fbits(x) & 0x7fffffffu); }
ST_DEV u32 to_u32_sat(float f) { return __float2uint_rz(f); }   // truncating, saturating, NaN -> 0
ST_DEV i32 to_i32_sat(float f) { return __float2int_rz(f); }

// ---- elementary functions ---------------------------------------------------
#if defined(ST_FAST) && ST_FAST
// SFU flavour: sin.approx / cos.approx (arguments on this path are within a few multiples of pi), ex2/lg2.approx.
ST_DEV void sincos_det(float x, float* s_out, float* c_out) { __sincosf(x, s_out, c_out); }
#else
ST_DEV void sincos_det(float xx, float* s_out, float* c_out) {
    float x = fabs_(xx);
    u32 j = (u32)(1.27323954473516f * x);
    float y = (float)j;
    if (j & 1u) { j += 1u; y += 1.0f; }
    j &= 7u;
    x = ((x - y * 0.78515625f) - y * 2.4187564849853515625e-4f) - y * 3.77489497744594108e-8f;
    float z = x * x;
    float ps = ((-1.9515295891e-4f * z + 8.3321608736e-3f) * z - 1.6666654611e-1f) * z * x + x;
    float pc = ((2.443315711809948e-5f * z - 1.388731625493765e-3f) * z + 4.166664568298827e-2f) * z * z - 0.5f * z + 1.0f;
    float s = (j == 0u) ? ps : (j == 2u) ? pc : (j == 4u) ? -ps : -pc;
    float c = (j == 0u) ? pc : (j == 2u) ? -ps : (j == 4u) ? -pc : ps;
    if (fbits(xx) & 0x80000000u) s = -s;
    *s_out = s; *c_out = c;
}
#endif
ST_DEV float sin_det(float x) { float s, c; sincos_det(x, &s, &c); return s; }
ST_DEV float cos_det(float x) { float s, c; sincos_det(x, &s, &c); return c; }

ST_DEV float asin_core(float x) {
    float z = x * x;
    return ((((4.2163199048e-2f * z + 2.4181311049e-2f) * z + 4.5470025998e-2f) * z + 7.4953002686e-2f) * z + 1.6666752422e-1f) * z * x + x;
}
ST_DEV float acos_det(float x) {
    if (!(x == x)) return x;
    if (x < -1.0f || x > 1.0f) return fnan();
    if (x > 0.5f) return 2.0f * asin_core(sqrtf(0.5f * (1.0f - x)));
    if (x < -0.5f) return kPi - 2.0f * asin_core(sqrtf(0.5f * (1.0f + x)));
    if (x >= 0.0f) return kHalfPi - asin_core(x);
    return kHalfPi + asin_core(-x);
}
// glam 0.24.2 math::acos_approx (third-party, restated from its published definition = DirectXMath XMScalarACos): what
// Vec3::angle_between evaluates; on the path only the spot-light cone uses it (strolle-gpu/src/light.rs:149-152).
ST_DEV float acos_approx_glam(float v) {
    bool nonnegative = v >= 0.0f;
    float x = fabs_(v);
    float omx = 1.0f - x;
    if (omx < 0.0f) omx = 0.0f;
    float root = sqrtf(omx);
    float result = ((((((-0.0012624911f * x + 0.0066700901f) * x - 0.0170881256f) * x + 0.0308918810f) * x - 0.0501743046f) * x + 0.0889789874f) * x - 0.2145988016f) * x + 1.5707963050f;
    result *= root;
    return nonnegative ? result : kPi - result;
}
ST_DEV float atan_core(float x) {   // x >= 0
    float y;
    if (x > 2.414213562373095f) { y = kHalfPi; x = -(1.0f / x); }
    else if (x > 0.4142135623730950f) { y = 0.7853981633974483f; x = (x - 1.0f) / (x + 1.0f); }
    else y = 0.0f;
    float z = x * x;
    y += (((8.05374449538e-2f * z - 1.38776856032e-1f) * z + 1.99777106478e-1f) * z - 3.33329491539e-1f) * z * x + x;
    return y;
}
ST_DEV float atan2_det(float y, float x) {
    if (!(x == x) || !(y == y)) return fnan();
    if (y == 0.0f) {
        if (x > 0.0f || (x == 0.0f && !(fbits(x) >> 31))) return y;
        return cpsign(kPi, y);
    }
    if (x == 0.0f) return cpsign(kHalfPi, y);
    float a = atan_core(fabs_(y) / fabs_(x));
    if (x < 0.0f) a = kPi - a;
    return cpsign(a, y);
}
ST_DEV float ldexp_det(float m, int n) {
    if (n > 127) { m = m * bitsf(0x7f000000u); n -= 127; if (n > 127) n = 127; }
    else if (n < -126) { m = m * bitsf(0x00800000u); n += 126; if (n < -126) n = -126; }
    return m * bitsf((u32)(n + 127) << 23);
}
#if defined(ST_FAST) && ST_FAST
ST_DEV float exp_det(float x) { return __expf(x); }
ST_DEV float log_det(float x) { return __logf(x); }
#else
ST_DEV float exp_det(float x) {
    if (!(x == x)) return x;
    if (x > 88.72283905206835f) return finf();
    if (x < -103.278929903431851103f) return 0.0f;
    float z = floorf(1.44269504088896341f * x + 0.5f);
    float r = (x - z * 0.693359375f) - z * -2.12194440e-4f;
    int n = (int)z;
    float zz = r * r;
    float p = (((((1.9875691500e-4f * r + 1.3981999507e-3f) * r + 8.3334519073e-3f) * r + 4.1665795894e-2f) * r + 1.6666665459e-1f) * r + 5.0000001201e-1f) * zz + r + 1.0f;
    return ldexp_det(p, n);
}
ST_DEV float log_det(float x) {   // x > 0 finite
    u32 bits = fbits(x);
    int e;
    if ((bits & 0x7f800000u) == 0u) { x = x * 8388608.0f; bits = fbits(x); e = (int)((bits >> 23) & 0xffu) - 126 - 23; }
    else e = (int)((bits >> 23) & 0xffu) - 126;
    float m = bitsf((bits & 0x007fffffu) | 0x3f000000u);
    if (m < 0.707106781186547524f) { e -= 1; m = (m + m) - 1.0f; } else m = m - 1.0f;
    float z = m * m;
    float y = ((((((((7.0376836292e-2f * m - 1.1514610310e-1f) * m + 1.1676998740e-1f) * m - 1.2420140846e-1f) * m + 1.4249322787e-1f) * m - 1.6668057665e-1f) * m + 2.0000714765e-1f) * m - 2.4999993993e-1f) * m + 3.3333331174e-1f) * m * z;
    float fe = (float)e;
    y += -2.12194440e-4f * fe;
    y += -0.5f * z;
    float r = m + y;
    r += 0.693359375f * fe;
    return r;
}
#endif
ST_DEV float pow_det(float x, float y) {
    if (y == 0.0f) return 1.0f;
    if (!(x == x) || !(y == y)) return fnan();
    if (x == 1.0f) return 1.0f;
    if (x == 0.0f) return (y > 0.0f) ? 0.0f : finf();
    if (x < 0.0f) return fnan();
    if (x == finf()) return (y > 0.0f) ? finf() : 0.0f;
    if (y == 1.0f) return x;
    if (y == 2.0f) return x * x;
    if (y == 3.0f) return (x * x) * x;
    if (y == 5.0f) { float x2 = x * x; return (x2 * x2) * x; }
    if (y == 8.0f) { float x2 = x * x; float x4 = x2 * x2; return x4 * x4; }
    if (y == 64.0f) { float x2 = x * x; float x4 = x2 * x2; float x8 = x4 * x4; float x16 = x8 * x8; float x32 = x16 * x16; return x32 * x32; }
    if (y == 1.5f) return x * sqrtf(x);
#if defined(ST_FAST) && ST_FAST
    return __powf(x, y);   // ex2.approx(y * lg2.approx(x))
#else
    return exp_det(y * log_det(x));
#endif
}

// ---- vectors ----------------------------------------------------------------
ST_DEV float3 f3(float x, float y, float z) { return make_float3(x, y, z); }
ST_DEV float3 f3s(float s) { return make_float3(s, s, s); }
ST_DEV float2 f2(float x, float y) { return make_float2(x, y); }
ST_DEV float4 f4(float x, float y, float z, float w) { return make_float4(x, y, z, w); }
ST_DEV float4 f4(float3 a, float w) { return make_float4(a.x, a.y, a.z, w); }
ST_DEV float4 f4zero() { return make_float4(0.f, 0.f, 0.f, 0.f); }
ST_DEV float3 xyz(float4 a) { return make_float3(a.x, a.y, a.z); }

ST_DEV float3 operator+(float3 a, float3 b) { return f3(a.x + b.x, a.y + b.y, a.z + b.z); }
ST_DEV float3 operator-(float3 a, float3 b) { return f3(a.x - b.x, a.y - b.y, a.z - b.z); }
ST_DEV float3 operator-(float3 a) { return f3(-a.x, -a.y, -a.z); }
ST_DEV float3 operator*(float3 a, float3 b) { return f3(a.x * b.x, a.y * b.y, a.z * b.z); }
ST_DEV float3 operator*(float3 a, float s) { return f3(a.x * s, a.y * s, a.z * s); }
ST_DEV float3 operator*(float s, float3 a) { return f3(s * a.x, s * a.y, s * a.z); }
ST_DEV float3 operator/(float3 a, float3 b) { return f3(a.x / b.x, a.y / b.y, a.z / b.z); }
ST_DEV float3 operator/(float3 a, float s) { return f3(a.x / s, a.y / s, a.z / s); }
ST_DEV float3 operator/(float s, float3 a) { return f3(s / a.x, s / a.y, s / a.z); }
ST_DEV float2 operator+(float2 a, float2 b) { return f2(a.x + b.x, a.y + b.y); }
ST_DEV float2 operator-(float2 a, float2 b) { return f2(a.x - b.x, a.y - b.y); }
ST_DEV float2 operator*(float2 a, float2 b) { return f2(a.x * b.x, a.y * b.y); }
ST_DEV float2 operator*(float2 a, float s) { return f2(a.x * s, a.y * s); }
ST_DEV float2 operator*(float s, float2 a) { return f2(s * a.x, s * a.y); }
ST_DEV float2 operator/(float2 a, float2 b) { return f2(a.x / b.x, a.y / b.y); }
ST_DEV float2 operator/(float2 a, float s) { return f2(a.x / s, a.y / s); }
ST_DEV float4 operator+(float4 a, float4 b) { return f4(a.x + b.x, a.y + b.y, a.z + b.z, a.w + b.w); }
ST_DEV float4 operator-(float4 a, float4 b) { return f4(a.x - b.x, a.y - b.y, a.z - b.z, a.w - b.w); }
ST_DEV float4 operator*(float4 a, float4 b) { return f4(a.x * b.x, a.y * b.y, a.z * b.z, a.w * b.w); }
ST_DEV float4 operator*(float4 a, float s) { return f4(a.x * s, a.y * s, a.z * s, a.w * s); }
ST_DEV float4 operator/(float4 a, float s) { return f4(a.x / s, a.y / s, a.z / s, a.w / s); }
ST_DEV bool all_zero(float4 a) { return a.x == 0.0f && a.y == 0.0f && a.z == 0.0f && a.w == 0.0f; }
ST_DEV bool eq3(float3 a, float3 b) { return a.x == b.x && a.y == b.y && a.z == b.z; }

ST_DEV float dot(float2 a, float2 b) { return (a.x * b.x) + (a.y * b.y); }
ST_DEV float dot(float3 a, float3 b) { return (a.x * b.x) + (a.y * b.y) + (a.z * b.z); }
ST_DEV float dot(float4 a, float4 b) { return (a.x * b.x) + (a.y * b.y) + (a.z * b.z) + (a.w * b.w); }
ST_DEV float3 cross(float3 a, float3 b) { return f3(a.y * b.z - b.y * a.z, a.z * b.x - b.z * a.x, a.x * b.y - b.x * a.y); }
// the same formulas through the flag-independent primitives (traversal, intersection tests)
ST_DEV float3 xsub3(float3 a, float3 b) { return f3(xsub(a.x, b.x), xsub(a.y, b.y), xsub(a.z, b.z)); }
ST_DEV float xdot(float3 a, float3 b) { return xadd(xadd(xmul(a.x, b.x), xmul(a.y, b.y)), xmul(a.z, b.z)); }
ST_DEV float3 xcross(float3 a, float3 b) { return f3(xsub(xmul(a.y, b.z), xmul(b.y, a.z)), xsub(xmul(a.z, b.x), xmul(b.z, a.x)), xsub(xmul(a.x, b.y), xmul(b.x, a.y))); }
ST_DEV float3 xscale(float3 a, float s) { return f3(xmul(a.x, s), xmul(a.y, s), xmul(a.z, s)); }
ST_DEV float3 xadd3(float3 a, float3 b) { return f3(xadd(a.x, b.x), xadd(a.y, b.y), xadd(a.z, b.z)); }
ST_DEV float3 xnorm(float3 a) { return xscale(a, xdiv(1.0f, xsqrt(xdot(a, a)))); }
ST_DEV float len2(float3 a) { return dot(a, a); }
ST_DEV float len2(float2 a) { return dot(a, a); }
ST_DEV float len(float3 a) { return sqrtf(dot(a, a)); }
ST_DEV float3 norm(float3 a) { return a * (1.0f / len(a)); }
ST_DEV float dist(float3 a, float3 b) { return len(a - b); }
ST_DEV float3 min3(float3 a, float3 b) { return f3(rmin(a.x, b.x), rmin(a.y, b.y), rmin(a.z, b.z)); }
ST_DEV float3 max3(float3 a, float3 b) { return f3(rmax(a.x, b.x), rmax(a.y, b.y), rmax(a.z, b.z)); }
ST_DEV float3 clamp3(float3 a, float3 lo, float3 hi) { return min3(max3(a, lo), hi); }
ST_DEV float lerpc(float a, float b, float t) { return a + (b - a) * rclamp(t, 0.0f, 1.0f); }
ST_DEV float3 lerpc(float3 a, float3 b, float t) { return a + (b - a) * rclamp(t, 0.0f, 1.0f); }
ST_DEV float3 reflect3(float3 self, float3 other) { return self - 2.0f * dot(other, self) * other; }
ST_DEV float luma(float3 c) { return dot(c, f3(0.2126f, 0.7152f, 0.0722f)); }

ST_DEV void ortho_pair(float3 n, float3* a_out, float3* b_out) {   // Duff et al. 2017 (glam any_orthonormal_pair)
    float sign = cpsign(1.0f, n.z);
    float a = -1.0f / (sign + n.z);
    float b = n.x * n.y * a;
    *a_out = f3(1.0f + sign * n.x * n.x * a, sign * b, -sign * n.x);
    *b_out = f3(b, sign + n.y * n.y * a, -n.y);
}

struct Mat4 { float4 c[4]; };   // column-major
ST_DEV float4 mat_mul(const Mat4& m, float4 v) {
    float4 r = m.c[0] * v.x;
    r = r + m.c[1] * v.y;
    r = r + m.c[2] * v.z;
    r = r + m.c[3] * v.w;
    return r;
}
ST_DEV float3 project_point(const Mat4& m, float3 p) {
    float4 r = m.c[0] * p.x;
    r = r + m.c[1] * p.y;
    r = r + m.c[2] * p.z;
    r = r + m.c[3];
    float rw = 1.0f / r.w;
    return f3(r.x * rw, r.y * rw, r.z * rw);
}
ST_DEV u32 pack_bytes(u32 a, u32 b, u32 c, u32 d) { return a | (b << 8) | (c << 16) | (d << 24); }

}  // namespace ST_NS
