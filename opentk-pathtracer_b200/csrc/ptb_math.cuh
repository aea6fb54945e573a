// ptb_math.cuh — fp32 evaluation of the GLSL built-ins the reference shaders use, for sm_100a.
//
// The reference's integrator (res/shaders/PathTracing/compute.glsl) is written against GLSL built-ins whose
// precision is implementation-defined.  This header fixes one evaluation (DESIGN.md "evaluation model"):
// IEEE binary32 everywhere, no implicit contraction (the TU is compiled with -fmad=false), explicit fma only
// inside dot / mat*vec / mix / the polynomial kernels, a/b = a * rcp.rn(b), min/max = FMNMX.
// Every Monte-Carlo path is chaotic, so a 1-ulp difference flips branches; fixing the evaluation is what makes
// matched-seed parity testable at all.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace ptb {

struct V3 { float x, y, z; };

__device__ __forceinline__ V3 mk(float x, float y, float z) { V3 r; r.x = x; r.y = y; r.z = z; return r; }
__device__ __forceinline__ V3 operator+(V3 a, V3 b) { return mk(a.x + b.x, a.y + b.y, a.z + b.z); }
__device__ __forceinline__ V3 operator-(V3 a, V3 b) { return mk(a.x - b.x, a.y - b.y, a.z - b.z); }
__device__ __forceinline__ V3 operator*(V3 a, V3 b) { return mk(a.x * b.x, a.y * b.y, a.z * b.z); }
__device__ __forceinline__ V3 operator*(V3 a, float s) { return mk(a.x * s, a.y * s, a.z * s); }
__device__ __forceinline__ V3 operator-(V3 a) { return mk(-a.x, -a.y, -a.z); }

#ifdef PTB_FAST
// The fast build (ptb_fast.cu, ptb_set_precision(PTB_PRECISION_FAST)): the hardware's special-function unit for 1/x, sqrt,
// 1/sqrt, sin, cos, 2^x (MUFU, ~1-2 ulp, denormals flushed) and, through -fmad=true, fused multiply-adds wherever the
// compiler finds them — roughly what a GL driver's compiler does with the same GLSL.  Not bit-comparable with anything;
// held to the exact build by the north star's tolerance (per-channel MSE < 1e-6 at matched seeds, tests/).
__device__ __forceinline__ float rcp(float b) { float r; asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(b)); return r; }
__device__ __forceinline__ float fdiv(float a, float b) { return a * rcp(b); }
__device__ __forceinline__ float fsqrt(float a) { float r; asm("sqrt.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(a)); return r; }
__device__ __forceinline__ float frsqrt(float a) { float r; asm("rsqrt.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(a)); return r; }
#else
__device__ __forceinline__ float rcp(float b) { return __frcp_rn(b); }
__device__ __forceinline__ float fdiv(float a, float b) { return a * __frcp_rn(b); }
__device__ __forceinline__ float fsqrt(float a) { return __fsqrt_rn(a); }
#endif
__device__ __forceinline__ float fma_(float a, float b, float c) { return __fmaf_rn(a, b, c); }
__device__ __forceinline__ float fmin_(float a, float b) { return fminf(a, b); }   // FMNMX: NaN loses, -0 < +0
__device__ __forceinline__ float fmax_(float a, float b) { return fmaxf(a, b); }
__device__ __forceinline__ float stepf(float edge, float x) { return x < edge ? 0.0f : 1.0f; }
__device__ __forceinline__ float signf(float x) { return x > 0.0f ? 1.0f : (x < 0.0f ? -1.0f : 0.0f); }
__device__ __forceinline__ float mixf(float x, float y, float a) { return __fmaf_rn(y, a, x * (1.0f - a)); }
__device__ __forceinline__ float pow5(float x) { float x2 = x * x; float x4 = x2 * x2; return x4 * x; }
__device__ __forceinline__ float pow15(float x) { return x * fsqrt(x); }

__device__ __forceinline__ float dot(V3 a, V3 b) { return __fmaf_rn(a.z, b.z, __fmaf_rn(a.y, b.y, a.x * b.x)); }
__device__ __forceinline__ float length(V3 a) { return fsqrt(dot(a, a)); }
#ifdef PTB_FAST
__device__ __forceinline__ V3 normalize(V3 a) { return a * frsqrt(dot(a, a)); }
#else
__device__ __forceinline__ V3 normalize(V3 a) { return a * __frcp_rn(__fsqrt_rn(dot(a, a))); }
#endif
__device__ __forceinline__ V3 mix(V3 a, V3 b, float t) { return mk(mixf(a.x, b.x, t), mixf(a.y, b.y, t), mixf(a.z, b.z, t)); }

// GLSL reflect: I - 2*dot(N,I)*N
__device__ __forceinline__ V3 reflect(V3 I, V3 N)
{
    const float k = 2.0f * dot(N, I);
    return mk(I.x - k * N.x, I.y - k * N.y, I.z - k * N.z);
}
// GLSL refract: k = 1 - eta^2 (1 - dot(N,I)^2); k < 0 -> 0 vector
__device__ __forceinline__ V3 refract(V3 I, V3 N, float eta)
{
    const float d = dot(N, I);
    const float k = 1.0f - (eta * eta) * (1.0f - d * d);
    if (k < 0.0f) return mk(0.0f, 0.0f, 0.0f);
    const float s = eta * d + fsqrt(k);
    return mk(eta * I.x - s * N.x, eta * I.y - s * N.y, eta * I.z - s * N.z);
}

// Row r of (M * v) for a column-major mat4 held as 16 floats: fma chain over columns 0..3.
__device__ __forceinline__ float mat_row(const float* __restrict__ M, int r, float x, float y, float z, float w)
{
    float acc = M[r] * x;
    acc = __fmaf_rn(M[4 + r], y, acc);
    acc = __fmaf_rn(M[8 + r], z, acc);
    acc = __fmaf_rn(M[12 + r], w, acc);
    return acc;
}

// sin & cos by 3-term Cody-Waite reduction modulo pi/2 and degree-7 / degree-8 minimax kernels.
__device__ __forceinline__ void sincos_(float x, float& s, float& c)
{
#ifdef PTB_FAST
    asm("sin.approx.ftz.f32 %0, %1;" : "=f"(s) : "f"(x));
    asm("cos.approx.ftz.f32 %0, %1;" : "=f"(c) : "f"(x));
    return;
#endif
    const float magic = 12582912.0f;                     // 1.5 * 2^23
    const float t = __fmaf_rn(x, 0.636619747f, magic);   // round(x * 2/pi) in the low mantissa bits
    const float q = t - magic;
    const uint32_t qi = __float_as_uint(t);
    float r = __fmaf_rn(q, -1.57079601e+00f, x);
    r = __fmaf_rn(q, -3.13916473e-07f, r);
    r = __fmaf_rn(q, -5.39030253e-15f, r);
    const float r2 = r * r;
    float ps = __fmaf_rn(r2, -1.9515295891e-4f, 8.3321608736e-3f);
    ps = __fmaf_rn(ps, r2, -1.6666654611e-1f);
    const float sn = __fmaf_rn(r * r2, ps, r);
    float pc = __fmaf_rn(r2, 2.443315711809948e-5f, -1.388731625493765e-3f);
    pc = __fmaf_rn(pc, r2, 4.166664568298827e-2f);
    const float cs = __fmaf_rn(r2 * r2, pc, __fmaf_rn(r2, -0.5f, 1.0f));
    const bool swap = qi & 1u;
    float so = swap ? cs : sn;
    float co = swap ? sn : cs;
    if (qi & 2u) so = -so;
    if ((qi + 1u) & 2u) co = -co;
    s = so;
    c = co;
}

// exp(x) = 2^k * e^r, k = round(x log2 e), r = x - k ln2 (two-term), degree-5 kernel on r^2, two-step scaling.
__device__ __forceinline__ float exp_(float x)
{
#ifdef PTB_FAST
    float e2;
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(e2) : "f"(x * 1.44269504f));
    return e2;
#endif
    if (x != x) return x + x;
    if (x > 88.7228394f) return __uint_as_float(0x7f800000u);
    if (x < -103.972084f) return 0.0f;
    const float magic = 12582912.0f;
    const float t = __fmaf_rn(x, 1.44269502f, magic);
    const float kf = t - magic;
    float r = __fmaf_rn(kf, -6.93145752e-1f, x);
    r = __fmaf_rn(kf, -1.42860677e-6f, r);
    float p = 1.9875691500e-4f;
    p = __fmaf_rn(p, r, 1.3981999507e-3f);
    p = __fmaf_rn(p, r, 8.3334519073e-3f);
    p = __fmaf_rn(p, r, 4.1665795894e-2f);
    p = __fmaf_rn(p, r, 1.6666665459e-1f);
    p = __fmaf_rn(p, r, 5.0000001201e-1f);
    const float e = __fmaf_rn(p, r * r, r) + 1.0f;
    const int k = __float2int_rz(kf);
    const int k1 = k >> 1;
    const int k2 = k - k1;
    return (e * __uint_as_float((uint32_t)(k1 + 127) << 23)) * __uint_as_float((uint32_t)(k2 + 127) << 23);
}

// log(x), x > 0 (frexp to [sqrt(1/2), sqrt(2)), degree-8 kernel) and pow(x, y) = exp(y log x) for x >= 0: only the
// post-process pass and the sRGB decode use them (PostProcessing/fragment.glsl:31, GL sRGB texture decode).
__device__ __forceinline__ float log_(float x)
{
    if (x != x || x < 0.0f) return __uint_as_float(0x7fc00000u);
    if (x == 0.0f) return __uint_as_float(0xff800000u);
    if (x == __uint_as_float(0x7f800000u)) return x;
    int e = 0;
    if (x < 1.17549435e-38f) { x = x * 8388608.0f; e = -23; }
    const uint32_t b = __float_as_uint(x);
    e += (int)(b >> 23) - 126;
    float m = __uint_as_float((b & 0x007fffffu) | 0x3f000000u);
    if (m < 0.707106781186547524f) { e -= 1; m = m + m - 1.0f; } else { m = m - 1.0f; }
    const float z = m * m;
    float p = 7.0376836292e-2f;
    p = __fmaf_rn(p, m, -1.1514610310e-1f);
    p = __fmaf_rn(p, m, 1.1676998740e-1f);
    p = __fmaf_rn(p, m, -1.2420140846e-1f);
    p = __fmaf_rn(p, m, 1.4249322787e-1f);
    p = __fmaf_rn(p, m, -1.6668057665e-1f);
    p = __fmaf_rn(p, m, 2.0000714765e-1f);
    p = __fmaf_rn(p, m, -2.4999993993e-1f);
    p = __fmaf_rn(p, m, 3.3333331174e-1f);
    float y = (p * z) * m;
    const float fe = (float)e;
    y = __fmaf_rn(fe, -2.12194440e-4f, y);
    y = __fmaf_rn(z, -0.5f, y);
    const float r = m + y;
    return __fmaf_rn(fe, 0.693359375f, r);
}
__device__ __forceinline__ float pow_(float x, float y) { return exp_(y * log_(x)); }

// compute.glsl:334-344 — PCG-RXS-M-XS hash stream; float(h) / 2^32 (can return exactly 1.0).
__device__ __forceinline__ uint32_t pcg_hash(uint32_t& seed)
{
    seed = seed * 747796405u + 2891336453u;
    const uint32_t word = ((seed >> ((seed >> 28u) + 4u)) ^ seed) * 277803737u;
    return (word >> 22u) ^ word;
}
__device__ __forceinline__ float rand01(uint32_t& seed)
{
    return __uint2float_rn(pcg_hash(seed)) * 2.3283064365386963e-10f;   // * 2^-32, exact
}

// ---- packed fp32x2 (Blackwell FADD2 / FMUL2 / FFMA2): two IEEE binary32 lanes per instruction, same rounding as the
// scalar ops, half the issue slots.  NOTE: ptxas contracts mul.rn.f32x2 + add/sub.rn.f32x2 into FFMA2 even with
// --fmad=false, so a product that must be rounded before an add (e.g. b*b - c) is kept scalar by the callers.
typedef unsigned long long f32x2;
__device__ __forceinline__ f32x2 pack2(float lo, float hi) { f32x2 r; asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(lo), "f"(hi)); return r; }
__device__ __forceinline__ void unpack2(f32x2 v, float& lo, float& hi) { asm("mov.b64 {%0, %1}, %2;" : "=f"(lo), "=f"(hi) : "l"(v)); }
__device__ __forceinline__ f32x2 sub2(f32x2 a, f32x2 b) { f32x2 r; asm("sub.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b)); return r; }
__device__ __forceinline__ f32x2 mul2(f32x2 a, f32x2 b) { f32x2 r; asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b)); return r; }
__device__ __forceinline__ f32x2 fma2(f32x2 a, f32x2 b, f32x2 c) { f32x2 r; asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(r) : "l"(a), "l"(b), "l"(c)); return r; }

} // namespace ptb
