// ORACLE — TEST INFRASTRUCTURE ONLY.
//
// CPU restatement of Patryk27/strolle's hot path (strolle-gpu + strolle-shaders
// + the host slice of the `strolle` crate).  Only tests/, __graft_entry__.smoke()
// and bench.py's cpu_baseline / --impl reference legs may use anything under
// oracle/.  The product (strolle_b200/) never includes, links or calls it.
//
// Parity status: the reference ships six unit tests (layouts + Camera::contain,
// SURVEY.md §4); all six are restated in tests/test_oracle_reference_tests.py
// and pin the serialisation layouts.  For traversal, ReSTIR and SVGF numerics
// the reference holds NO golden vectors and cannot be built here (no Rust, no
// Vulkan) => "parity unpinned" for those; this oracle is the pin.
//
// This header: scalar f32 vector maths with glam 0.24.2's evaluation order
// (third-party dependency, Cargo.lock `glam 0.24.2`, not vendored under
// /root/reference — semantics restated from its published scalar
// implementation) and a deterministic libm subset.
//
// Build with -ffp-contract=off and no fast-math: every operation is a single
// IEEE-754 binary32 operation, in the order written.
#pragma once
#include <cmath>
#include <cstdint>
#include <cstring>

namespace orc {

typedef uint32_t u32;
typedef int32_t i32;

static inline u32 f2u(float f) { u32 u; std::memcpy(&u, &f, 4); return u; }
static inline float u2f(u32 u) { float f; std::memcpy(&f, &u, 4); return f; }

static const float PI = 3.14159265358979323846f;
static const float F32_MAX = 3.40282347e+38f;
static const float F32_EPSILON = 1.1920929e-7f;
static const float F32_INF = INFINITY;

// Rust f32::min / f32::max: return the non-NaN operand (== C fminf/fmaxf).
static inline float fmin_(float a, float b) { return (a != a) ? b : ((b != b) ? a : (a < b ? a : b)); }
static inline float fmax_(float a, float b) { return (a != a) ? b : ((b != b) ? a : (a > b ? a : b)); }
// Rust f32::clamp: NaN stays NaN; plain comparisons.
static inline float clampf(float x, float lo, float hi) { if (x < lo) x = lo; if (x > hi) x = hi; return x; }
static inline float saturate(float x) { return clampf(x, 0.0f, 1.0f); }
static inline float sqr(float x) { return x * x; }
static inline float copysign_(float mag, float sgn) { return u2f((f2u(mag) & 0x7fffffffu) | (f2u(sgn) & 0x80000000u)); }
static inline float abs_(float x) { return u2f(f2u(x) & 0x7fffffffu); }
static inline float sqrt_(float x) { return sqrtf(x); }  // IEEE correctly rounded

// Rust `as u32` / `as i32` from f32: truncating, saturating, NaN -> 0.
static inline u32 f2u_sat(float f) {
    if (!(f == f)) return 0u;
    if (f <= 0.0f) return 0u;
    if (f >= 4294967296.0f) return 0xffffffffu;
    return (u32)f;
}
static inline i32 f2i_sat(float f) {
    if (!(f == f)) return 0;
    if (f <= -2147483648.0f) return INT32_MIN;
    if (f >= 2147483648.0f) return INT32_MAX;
    return (i32)f;
}
static inline float floor_(float x) { return floorf(x); }   // exact ops
static inline float ceil_(float x) { return ceilf(x); }
static inline float trunc_(float x) { return truncf(x); }
static inline float round_(float x) { return roundf(x); }   // half away from zero == Rust f32::round
static inline float fmod_(float x, float y) { return fmodf(x, y); }  // exact

// ---------------------------------------------------------------------------
// Deterministic elementary functions.
//
// The reference evaluates sin/cos/acos/atan2/exp/powf through GLSL.std.450 on
// the device (driver-defined precision) and through libm on the host.  Neither
// is reproducible bit-for-bit on another platform, so the oracle carries its
// own implementations (Cephes single-precision algorithms, public domain,
// S. Moshier) built from + - * / only.  The CUDA product implements the same
// published algorithms, which is what makes whole-frame bit-exact parity
// testable.  Build with -DORC_LIBM to swap in the host libm instead (used by
// tests to show the substitution is immaterial at the 1e-3 tolerance).
// ---------------------------------------------------------------------------
#ifdef ORC_LIBM
static inline float sin_(float x) { return sinf(x); }
static inline float cos_(float x) { return cosf(x); }
static inline float acos_(float x) { return acosf(x); }
static inline float atan2_(float y, float x) { return atan2f(y, x); }
static inline float exp_(float x) { return expf(x); }
static inline float pow_(float x, float y) { return powf(x, y); }
#else
// sin/cos: Cephes sinf/cosf (octant reduction with 3-part pi/4).
static inline void sincos_core(float xx, float* s_out, float* c_out) {
    const float FOPI = 1.27323954473516f;
    const float DP1 = 0.78515625f, DP2 = 2.4187564849853515625e-4f, DP3 = 3.77489497744594108e-8f;
    float x = abs_(xx);
    u32 j = (u32)(FOPI * x);          // |x| assumed < 2^23 (angles here are < 100)
    float y = (float)j;
    if (j & 1u) { j += 1u; y += 1.0f; }
    j &= 7u;
    x = ((x - y * DP1) - y * DP2) - y * DP3;
    float z = x * x;
    float ps = ((-1.9515295891e-4f * z + 8.3321608736e-3f) * z - 1.6666654611e-1f) * z * x + x;
    float pc = ((2.443315711809948e-5f * z - 1.388731625493765e-3f) * z + 4.166664568298827e-2f) * z * z - 0.5f * z + 1.0f;
    float s, c;
    switch (j) {
        case 0: s = ps;  c = pc;  break;
        case 2: s = pc;  c = -ps; break;
        case 4: s = -ps; c = -pc; break;
        default: /*6*/ s = -pc; c = ps; break;
    }
    if (f2u(xx) & 0x80000000u) s = -s;
    *s_out = s; *c_out = c;
}
static inline float sin_(float x) { float s, c; sincos_core(x, &s, &c); return s; }
static inline float cos_(float x) { float s, c; sincos_core(x, &s, &c); return c; }

// asin on [0, 0.5] (Cephes asinf polynomial)
static inline float asin_poly(float x) {
    float z = x * x;
    return ((((4.2163199048e-2f * z + 2.4181311049e-2f) * z + 4.5470025998e-2f) * z + 7.4953002686e-2f) * z + 1.6666752422e-1f) * z * x + x;
}
static inline float acos_(float x) {
    if (!(x == x)) return x;
    if (x < -1.0f || x > 1.0f) return u2f(0x7fc00000u);
    if (x > 0.5f) return 2.0f * asin_poly(sqrt_(0.5f * (1.0f - x)));
    if (x < -0.5f) return PI - 2.0f * asin_poly(sqrt_(0.5f * (1.0f + x)));
    if (x >= 0.0f) return 1.5707963267948966f - asin_poly(x);
    return 1.5707963267948966f + asin_poly(-x);
}
// atan for x >= 0 (Cephes atanf)
static inline float atan_pos(float x) {
    float y;
    if (x > 2.414213562373095f) { y = 1.5707963267948966f; x = -(1.0f / x); }
    else if (x > 0.4142135623730950f) { y = 0.7853981633974483f; x = (x - 1.0f) / (x + 1.0f); }
    else y = 0.0f;
    float z = x * x;
    y += (((8.05374449538e-2f * z - 1.38776856032e-1f) * z + 1.99777106478e-1f) * z - 3.33329491539e-1f) * z * x + x;
    return y;
}
static inline float atan2_(float y, float x) {
    if (!(x == x) || !(y == y)) return u2f(0x7fc00000u);
    if (y == 0.0f) {
        if (x > 0.0f || (x == 0.0f && !(f2u(x) >> 31))) return y;        // +-0
        return copysign_(PI, y);
    }
    if (x == 0.0f) return copysign_(1.5707963267948966f, y);
    float a = atan_pos(abs_(y) / abs_(x));   // inf/inf not reachable here
    if (x < 0.0f) a = PI - a;
    return copysign_(a, y);
}
// 2^n scaling by exponent construction, n clamped so the result is a normal or 0/inf.
static inline float ldexp_(float m, int n) {
    if (n > 127) { m = m * u2f(0x7f000000u); n -= 127; if (n > 127) n = 127; }
    else if (n < -126) { m = m * u2f(0x00800000u); n += 126; if (n < -126) n = -126; }
    return m * u2f((u32)(n + 127) << 23);
}
// exp: Cephes expf (Cody–Waite ln2 split + degree-5 polynomial)
static inline float exp_(float x) {
    if (!(x == x)) return x;
    if (x > 88.72283905206835f) return F32_INF;
    if (x < -103.278929903431851103f) return 0.0f;
    const float LOG2EF = 1.44269504088896341f;
    const float C1 = 0.693359375f, C2 = -2.12194440e-4f;
    float z = floor_(LOG2EF * x + 0.5f);
    float r = (x - z * C1) - z * C2;
    int n = (int)z;
    float zz = r * r;
    float p = (((((1.9875691500e-4f * r + 1.3981999507e-3f) * r + 8.3334519073e-3f) * r + 4.1665795894e-2f) * r + 1.6666665459e-1f) * r + 5.0000001201e-1f) * zz + r + 1.0f;
    return ldexp_(p, n);
}
// log for x > 0, normal or subnormal (Cephes logf), returned as (hi part e*ln2 handled by caller)
static inline float log_(float x) {
    // frexp
    u32 bits = f2u(x);
    int e;
    if ((bits & 0x7f800000u) == 0) {  // subnormal
        x = x * 8388608.0f; bits = f2u(x); e = (int)((bits >> 23) & 0xff) - 126 - 23;
    } else e = (int)((bits >> 23) & 0xff) - 126;
    float m = u2f((bits & 0x007fffffu) | 0x3f000000u);   // [0.5, 1)
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
// powf for the argument ranges the path uses (x >= 0, finite y): exp(y*log x) with
// the IEEE special cases that are reachable.
static inline float pow_(float x, float y) {
    if (y == 0.0f) return 1.0f;
    if (!(x == x) || !(y == y)) return u2f(0x7fc00000u);
    if (x == 1.0f) return 1.0f;
    if (x == 0.0f) return (y > 0.0f) ? 0.0f : F32_INF;
    if (x < 0.0f) return u2f(0x7fc00000u);
    if (x == F32_INF) return (y > 0.0f) ? F32_INF : 0.0f;
    if (y == 1.0f) return x;
    if (y == 2.0f) return x * x;
    // small integer / half-integer exponents used by the path (3, 5, 8, 64, 1.5): by multiplication
    if (y == 3.0f) return (x * x) * x;
    if (y == 5.0f) { float x2 = x * x; return (x2 * x2) * x; }
    if (y == 8.0f) { float x2 = x * x; float x4 = x2 * x2; return x4 * x4; }
    if (y == 64.0f) { float x2 = x * x; float x4 = x2 * x2; float x8 = x4 * x4; float x16 = x8 * x8; float x32 = x16 * x16; return x32 * x32; }
    if (y == 1.5f) return x * sqrt_(x);
    return exp_(y * log_(x));
}
#endif

// ---------------------------------------------------------------------------
// Vectors (glam scalar semantics)
// ---------------------------------------------------------------------------
struct V2 { float x, y; };
struct V3 { float x, y, z; };
struct V4 { float x, y, z, w; };
struct IV2 { i32 x, y; };
struct UV2 { u32 x, y; };

static inline V2 v2(float x, float y) { V2 r = {x, y}; return r; }
static inline V3 v3(float x, float y, float z) { V3 r = {x, y, z}; return r; }
static inline V3 v3s(float s) { return v3(s, s, s); }
static inline V4 v4(float x, float y, float z, float w) { V4 r = {x, y, z, w}; return r; }
static inline V4 v4(V3 a, float w) { return v4(a.x, a.y, a.z, w); }
static inline V4 v4z() { return v4(0, 0, 0, 0); }
static inline V3 xyz(V4 a) { return v3(a.x, a.y, a.z); }
static inline IV2 iv2(i32 x, i32 y) { IV2 r = {x, y}; return r; }
static inline UV2 uv2(u32 x, u32 y) { UV2 r = {x, y}; return r; }

static inline V2 operator+(V2 a, V2 b) { return v2(a.x + b.x, a.y + b.y); }
static inline V2 operator-(V2 a, V2 b) { return v2(a.x - b.x, a.y - b.y); }
static inline V2 operator*(V2 a, V2 b) { return v2(a.x * b.x, a.y * b.y); }
static inline V2 operator*(V2 a, float s) { return v2(a.x * s, a.y * s); }
static inline V2 operator*(float s, V2 a) { return v2(s * a.x, s * a.y); }
static inline V2 operator/(V2 a, V2 b) { return v2(a.x / b.x, a.y / b.y); }
static inline V2 operator/(V2 a, float s) { return v2(a.x / s, a.y / s); }

static inline V3 operator+(V3 a, V3 b) { return v3(a.x + b.x, a.y + b.y, a.z + b.z); }
static inline V3 operator-(V3 a, V3 b) { return v3(a.x - b.x, a.y - b.y, a.z - b.z); }
static inline V3 operator-(V3 a) { return v3(-a.x, -a.y, -a.z); }
static inline V3 operator*(V3 a, V3 b) { return v3(a.x * b.x, a.y * b.y, a.z * b.z); }
static inline V3 operator*(V3 a, float s) { return v3(a.x * s, a.y * s, a.z * s); }
static inline V3 operator*(float s, V3 a) { return v3(s * a.x, s * a.y, s * a.z); }
static inline V3 operator/(V3 a, V3 b) { return v3(a.x / b.x, a.y / b.y, a.z / b.z); }
static inline V3 operator/(V3 a, float s) { return v3(a.x / s, a.y / s, a.z / s); }
static inline V3 operator/(float s, V3 a) { return v3(s / a.x, s / a.y, s / a.z); }
static inline V3& operator+=(V3& a, V3 b) { a = a + b; return a; }
static inline V3& operator*=(V3& a, V3 b) { a = a * b; return a; }
static inline V3& operator*=(V3& a, float s) { a = a * s; return a; }
static inline bool operator==(V3 a, V3 b) { return a.x == b.x && a.y == b.y && a.z == b.z; }
static inline bool operator!=(V3 a, V3 b) { return !(a == b); }

static inline V4 operator+(V4 a, V4 b) { return v4(a.x + b.x, a.y + b.y, a.z + b.z, a.w + b.w); }
static inline V4 operator-(V4 a, V4 b) { return v4(a.x - b.x, a.y - b.y, a.z - b.z, a.w - b.w); }
static inline V4 operator*(V4 a, V4 b) { return v4(a.x * b.x, a.y * b.y, a.z * b.z, a.w * b.w); }
static inline V4 operator*(V4 a, float s) { return v4(a.x * s, a.y * s, a.z * s, a.w * s); }
static inline V4 operator/(V4 a, float s) { return v4(a.x / s, a.y / s, a.z / s, a.w / s); }
static inline bool operator==(V4 a, V4 b) { return a.x == b.x && a.y == b.y && a.z == b.z && a.w == b.w; }
static inline bool is_zero(V4 a) { return a.x == 0.0f && a.y == 0.0f && a.z == 0.0f && a.w == 0.0f; }

static inline float dot(V2 a, V2 b) { return (a.x * b.x) + (a.y * b.y); }
static inline float dot(V3 a, V3 b) { return (a.x * b.x) + (a.y * b.y) + (a.z * b.z); }
static inline float dot(V4 a, V4 b) { return (a.x * b.x) + (a.y * b.y) + (a.z * b.z) + (a.w * b.w); }
static inline V3 cross(V3 a, V3 b) {
    return v3(a.y * b.z - b.y * a.z, a.z * b.x - b.z * a.x, a.x * b.y - b.x * a.y);
}
static inline float length_squared(V2 a) { return dot(a, a); }
static inline float length_squared(V3 a) { return dot(a, a); }
static inline float length(V3 a) { return sqrt_(dot(a, a)); }
static inline V3 normalize(V3 a) { return a * (1.0f / length(a)); }
static inline float distance(V3 a, V3 b) { return length(a - b); }
static inline V3 vmin(V3 a, V3 b) { return v3(fmin_(a.x, b.x), fmin_(a.y, b.y), fmin_(a.z, b.z)); }
static inline V3 vmax(V3 a, V3 b) { return v3(fmax_(a.x, b.x), fmax_(a.y, b.y), fmax_(a.z, b.z)); }
static inline V3 vabs(V3 a) { return v3(abs_(a.x), abs_(a.y), abs_(a.z)); }
static inline V3 vclamp(V3 a, V3 lo, V3 hi) { return vmin(vmax(a, lo), hi); }
static inline V3 vlerp(V3 a, V3 b, float s) { return a + (b - a) * s; }   // glam Vec3::lerp (unclamped)

// strolle-gpu/src/utils.rs:23-31 — strolle's own lerp clamps t
static inline float lerp_c(float a, float b, float t) { return a + (b - a) * clampf(t, 0.0f, 1.0f); }
static inline V3 lerp_c(V3 a, V3 b, float t) { return a + (b - a) * clampf(t, 0.0f, 1.0f); }

// strolle-gpu/src/utils/vec3_ext.rs
static inline V3 reflect(V3 self, V3 other) { return self - 2.0f * dot(other, self) * other; }
static inline float luma(V3 c) { return dot(c, v3(0.2126f, 0.7152f, 0.0722f)); }
static inline float perc_luma(V3 c) { return sqrt_(luma(c)); }

// glam Vec3::any_orthonormal_pair (Duff et al. 2017)
static inline void any_orthonormal_pair(V3 n, V3* a_out, V3* b_out) {
    float sign = copysign_(1.0f, n.z);
    float a = -1.0f / (sign + n.z);
    float b = n.x * n.y * a;
    *a_out = v3(1.0f + sign * n.x * n.x * a, sign * b, -sign * n.x);
    *b_out = v3(b, sign + n.y * n.y * a, -n.y);
}

// Column-major 4x4 (glam Mat4)
struct M4 { V4 c[4]; };
static inline V4 mul(const M4& m, V4 v) {
    V4 r = m.c[0] * v.x;
    r = r + m.c[1] * v.y;
    r = r + m.c[2] * v.z;
    r = r + m.c[3] * v.w;
    return r;
}
static inline M4 mul(const M4& a, const M4& b) {
    M4 r;
    for (int i = 0; i < 4; i++) r.c[i] = mul(a, b.c[i]);
    return r;
}
// glam Mat4::project_point3
static inline V3 project_point3(const M4& m, V3 p) {
    V4 r = m.c[0] * p.x;
    r = r + m.c[1] * p.y;
    r = r + m.c[2] * p.z;
    r = r + m.c[3];
    float rw = 1.0f / r.w;   // glam: res.xyz() * res.w.recip()  (Vec4 / wwww in SIMD builds; same value up to 1 ulp)
    return v3(r.x * rw, r.y * rw, r.z * rw);
}
// General 4x4 inverse (cofactor expansion, GLM/glam ordering).
static inline M4 inverse(const M4& m) {
    float m00 = m.c[0].x, m01 = m.c[0].y, m02 = m.c[0].z, m03 = m.c[0].w;
    float m10 = m.c[1].x, m11 = m.c[1].y, m12 = m.c[1].z, m13 = m.c[1].w;
    float m20 = m.c[2].x, m21 = m.c[2].y, m22 = m.c[2].z, m23 = m.c[2].w;
    float m30 = m.c[3].x, m31 = m.c[3].y, m32 = m.c[3].z, m33 = m.c[3].w;
    float coef00 = m22 * m33 - m32 * m23;
    float coef02 = m12 * m33 - m32 * m13;
    float coef03 = m12 * m23 - m22 * m13;
    float coef04 = m21 * m33 - m31 * m23;
    float coef06 = m11 * m33 - m31 * m13;
    float coef07 = m11 * m23 - m21 * m13;
    float coef08 = m21 * m32 - m31 * m22;
    float coef10 = m11 * m32 - m31 * m12;
    float coef11 = m11 * m22 - m21 * m12;
    float coef12 = m20 * m33 - m30 * m23;
    float coef14 = m10 * m33 - m30 * m13;
    float coef15 = m10 * m23 - m20 * m13;
    float coef16 = m20 * m32 - m30 * m22;
    float coef18 = m10 * m32 - m30 * m12;
    float coef19 = m10 * m22 - m20 * m12;
    float coef20 = m20 * m31 - m30 * m21;
    float coef22 = m10 * m31 - m30 * m11;
    float coef23 = m10 * m21 - m20 * m11;
    V4 fac0 = v4(coef00, coef00, coef02, coef03);
    V4 fac1 = v4(coef04, coef04, coef06, coef07);
    V4 fac2 = v4(coef08, coef08, coef10, coef11);
    V4 fac3 = v4(coef12, coef12, coef14, coef15);
    V4 fac4 = v4(coef16, coef16, coef18, coef19);
    V4 fac5 = v4(coef20, coef20, coef22, coef23);
    V4 vec0 = v4(m10, m00, m00, m00);
    V4 vec1 = v4(m11, m01, m01, m01);
    V4 vec2 = v4(m12, m02, m02, m02);
    V4 vec3_ = v4(m13, m03, m03, m03);
    V4 inv0 = (vec1 * fac0 - vec2 * fac1) + vec3_ * fac2;
    V4 inv1 = (vec0 * fac0 - vec2 * fac3) + vec3_ * fac4;
    V4 inv2 = (vec0 * fac1 - vec1 * fac3) + vec3_ * fac5;
    V4 inv3 = (vec0 * fac2 - vec1 * fac4) + vec2 * fac5;
    V4 sign_a = v4(1.0f, -1.0f, 1.0f, -1.0f);
    V4 sign_b = v4(-1.0f, 1.0f, -1.0f, 1.0f);
    M4 inv;
    inv.c[0] = inv0 * sign_a;
    inv.c[1] = inv1 * sign_b;
    inv.c[2] = inv2 * sign_a;
    inv.c[3] = inv3 * sign_b;
    V4 col0 = v4(inv.c[0].x, inv.c[1].x, inv.c[2].x, inv.c[3].x);
    V4 dot0 = m.c[0] * col0;
    float dot1 = dot0.x + dot0.y + dot0.z + dot0.w;
    float rcp_det = 1.0f / dot1;
    for (int i = 0; i < 4; i++) inv.c[i] = inv.c[i] * rcp_det;
    return inv;
}

// u32 <-> 4 bytes (strolle-gpu/src/utils/u32_ext.rs)
static inline u32 from_bytes(u32 a, u32 b, u32 c, u32 d) { return a | (b << 8) | (c << 16) | (d << 24); }

}  // namespace orc
