/*
 * oracle/glsl_model.h — TEST INFRASTRUCTURE ONLY (never linked into the product).
 *
 * The "GLSL evaluation model": how this oracle evaluates the GLSL 4.50 built-ins
 * and operators the reference shaders use
 *   res/shaders/PathTracing/compute.glsl            (the integrator)
 *   res/shaders/AtmosphericScattering/compute.glsl  (the environment producer)
 * GLSL leaves the precision of these implementation-defined (GLSL 4.50 §4.7.1),
 * and nothing of the reference can execute in the build container, so the oracle
 * fixes ONE admissible evaluation, in IEEE-754 binary32, that a CPU (gcc,
 * -ffp-contract=off -mfma) and a GPU (nvcc, -fmad=false) reproduce bit for bit.
 * The product (csrc/ptb_math.cuh) is a separate, hand-written CUDA implementation
 * of the same model; tests/ compare the two bitwise.
 *
 *   a+b a-b a*b        IEEE RNE, denormals kept, never contracted
 *   a/b                a * rcp(b), rcp = correctly rounded 1/b   (GLSL: 2.5 ULP)
 *   sqrt               correctly rounded
 *   dot(a,b)           fma(a.z,b.z, fma(a.y,b.y, a.x*b.x))      (FMUL,FFMA,FFMA)
 *   mat4*vec4          per row: fma chain over columns 0,1,2,3
 *   length / normalize sqrt(dot(v,v)) ; v * rcp(sqrt(dot(v,v)))  (0 -> NaN)
 *   mix(x,y,a)         fma(y, a, x*(1-a))
 *   reflect / refract  GLSL 4.50 §8.5 formulas, evaluated literally
 *   pow(x,5.0)         ((x*x)*(x*x))*x   total, sign preserving   
 *   pow(x,1.5)         x*sqrt(x)
 *   min/max            IEEE minNum/maxNum (NaN loses), -0 < +0      (FMNMX)
 *   step, sign, abs    GLSL definitions
 *   sin, cos, exp      the fixed polynomial algorithms below
 *   float(uint)        RNE;  f2i = truncation, NaN -> 0, saturating
 * PARITY PIN: the reference has no tests, fixtures or golden vectors for this
 * path and its C# + OpenGL host cannot run here, but its
 * SHADERS can: oracle/build_ref.py compiles the three GLSL files from
 * /root/reference with g++ against oracle/glsl_shim.hpp (which evaluates GLSL
 * operators and built-ins with the scalar primitives of THIS header) into
 * oracle/_ref/libglsl_ref.so.  tests/test_reference_pin.py holds the hand-written
 * restatement (pt_oracle.c, atmosphere_oracle.c) to that library bit for bit, and
 * tests/golden/ref_*.npz are its outputs.  So the algorithm is pinned to the
 * reference's source text; what remains a choice is this header — how a GL
 * driver would round `/`, sin, cos, exp, pow and filter a cubemap.
 */
#ifndef PTO_GLSL_MODEL_H
#define PTO_GLSL_MODEL_H

#include <stdint.h>
#include <string.h>

/* G_FN: `static inline` for gcc / g++; under nvcc also __host__ __device__, so the SAME model text can be compiled for the
 * GPU (oracle/build_ref.py --cuda compiles the reference's shaders with nvcc as the GL-compute proxy).  With nvcc's
 * defaults (-prec-div=true -prec-sqrt=true) and -fmad=false, 1.0f / b, sqrtf and fmaf round exactly as on the host. */
#ifdef __CUDACC__
#define G_FN static inline __host__ __device__
#else
#define G_FN static inline
#endif

typedef struct { float x, y, z; } vec3;
typedef struct { float x, y, z, w; } vec4;

G_FN uint32_t g_bits(float f) { uint32_t u; memcpy(&u, &f, 4); return u; }
G_FN float g_float(uint32_t u) { float f; memcpy(&f, &u, 4); return f; }

#if defined(__CUDA_ARCH__)
G_FN float g_fma(float a, float b, float c) { return __fmaf_rn(a, b, c); }
G_FN float g_rcp(float b) { return __frcp_rn(b); }
#else
G_FN float g_fma(float a, float b, float c) { return __builtin_fmaf(a, b, c); }
G_FN float g_rcp(float b) { return 1.0f / b; }
#endif
G_FN float g_div(float a, float b) { return a * g_rcp(b); }
#if defined(__CUDA_ARCH__)
G_FN float g_sqrt(float a) { return __fsqrt_rn(a); }
#else
G_FN float g_sqrt(float a) { return __builtin_sqrtf(a); }
#endif
G_FN float g_abs(float a) { return g_float(g_bits(a) & 0x7fffffffu); }
G_FN int g_isnan(float a) { return a != a; }

/* minNum / maxNum with -0 < +0 (what FMNMX does). */
G_FN float g_min(float a, float b)
{
    if (g_isnan(a)) return b;
    if (g_isnan(b)) return a;
    if (a == b) return (g_bits(a) >> 31) ? a : b;
    return b < a ? b : a;
}
G_FN float g_max(float a, float b)
{
    if (g_isnan(a)) return b;
    if (g_isnan(b)) return a;
    if (a == b) return (g_bits(a) >> 31) ? b : a;
    return a < b ? b : a;
}
G_FN float g_step(float edge, float x) { return x < edge ? 0.0f : 1.0f; }
G_FN float g_sign(float x) { return (float)((x > 0.0f) - (x < 0.0f)); }
G_FN float g_mix(float x, float y, float a) { return g_fma(y, a, x * (1.0f - a)); }
G_FN float g_pow5(float x) { float x2 = x * x; float x4 = x2 * x2; return x4 * x; }
G_FN float g_pow15(float x) { return x * g_sqrt(x); }

/* float -> int: truncate; NaN -> 0; saturate to +-2^30 (indices are clamped by callers anyway). */
G_FN int g_f2i(float x)
{
    if (g_isnan(x)) return 0;
    if (x >= 1073741824.0f) return 1073741824;
    if (x <= -1073741824.0f) return -1073741824;
    return (int)x;
}
/* floor as float, via truncation fix-up (|x| < 2^30 assumed by callers; otherwise x itself). */
G_FN float g_floor(float x)
{
    if (g_isnan(x)) return x;
    if (!(g_abs(x) < 1073741824.0f)) return x;
    float t = (float)(int)x;
    return t > x ? t - 1.0f : t;
}

/* ---- vec3 helpers ---------------------------------------------------------------- */
G_FN vec3 v3(float x, float y, float z) { vec3 r = { x, y, z }; return r; }
G_FN vec3 v_add(vec3 a, vec3 b) { return v3(a.x + b.x, a.y + b.y, a.z + b.z); }
G_FN vec3 v_sub(vec3 a, vec3 b) { return v3(a.x - b.x, a.y - b.y, a.z - b.z); }
G_FN vec3 v_mul(vec3 a, vec3 b) { return v3(a.x * b.x, a.y * b.y, a.z * b.z); }
G_FN vec3 v_scale(vec3 a, float s) { return v3(a.x * s, a.y * s, a.z * s); }
G_FN vec3 v_neg(vec3 a) { return v3(-a.x, -a.y, -a.z); }
G_FN float v_dot(vec3 a, vec3 b) { return g_fma(a.z, b.z, g_fma(a.y, b.y, a.x * b.x)); }
G_FN float v_length(vec3 a) { return g_sqrt(v_dot(a, a)); }
G_FN vec3 v_normalize(vec3 a) { return v_scale(a, g_rcp(g_sqrt(v_dot(a, a)))); }
G_FN vec3 v_mix(vec3 a, vec3 b, float t)
{
    return v3(g_mix(a.x, b.x, t), g_mix(a.y, b.y, t), g_mix(a.z, b.z, t));
}
/* reflect(I,N) = I - 2*dot(N,I)*N */
G_FN vec3 v_reflect(vec3 I, vec3 N)
{
    float k = 2.0f * v_dot(N, I);
    return v3(I.x - k * N.x, I.y - k * N.y, I.z - k * N.z);
}
/* refract(I,N,eta): k = 1 - eta*eta*(1 - dot(N,I)^2); k<0 -> 0; else eta*I - (eta*dot(N,I)+sqrt(k))*N */
G_FN vec3 v_refract(vec3 I, vec3 N, float eta)
{
    float d = v_dot(N, I);
    float k = 1.0f - (eta * eta) * (1.0f - d * d);
    if (k < 0.0f) return v3(0.0f, 0.0f, 0.0f);
    float s = eta * d + g_sqrt(k);
    return v3(eta * I.x - s * N.x, eta * I.y - s * N.y, eta * I.z - s * N.z);
}
/* (M * v).row r for a column-major mat4 stored as 16 floats (GLSL std140 / OpenTK row-major bytes). */
G_FN float m4_row(const float *M, int r, float x, float y, float z, float w)
{
    float acc = M[0 + r] * x;
    acc = g_fma(M[4 + r], y, acc);
    acc = g_fma(M[8 + r], z, acc);
    acc = g_fma(M[12 + r], w, acc);
    return acc;
}

/* ---- transcendental algorithms ---------------------------------------------------- */
#define G_MAGIC 12582912.0f /* 1.5 * 2^23: adding it rounds to an integer (RNE) */

/* sin and cos of x by Cody-Waite reduction to [-pi/4, pi/4] (three-term pi/2) and the
 * classic single-precision minimax kernels.  Exact same operation sequence on both sides. */
G_FN void g_sincos(float x, float *s_out, float *c_out)
{
    float t = g_fma(x, 0.636619747f, G_MAGIC);       /* x * 2/pi, rounded to integer */
    float q = t - G_MAGIC;
    uint32_t qi = g_bits(t);                          /* low mantissa bits hold q mod 4 */
    float r = g_fma(q, -1.57079601e+00f, x);          /* pi/2 split: hi */
    r = g_fma(q, -3.13916473e-07f, r);                /* mid */
    r = g_fma(q, -5.39030253e-15f, r);                /* lo */
    float r2 = r * r;
    /* sin kernel: r + r*r2*(S1 + r2*(S2 + r2*S3)) */
    float ps = g_fma(r2, -1.9515295891e-4f, 8.3321608736e-3f);
    ps = g_fma(ps, r2, -1.6666654611e-1f);
    float sn = g_fma(r * r2, ps, r);
    /* cos kernel: 1 - r2/2 + r2*r2*(C1 + r2*(C2 + r2*C3)) */
    float pc = g_fma(r2, 2.443315711809948e-5f, -1.388731625493765e-3f);
    pc = g_fma(pc, r2, 4.166664568298827e-2f);
    float cs = g_fma(r2 * r2, pc, g_fma(r2, -0.5f, 1.0f));
    float s = (qi & 1u) ? cs : sn;
    float c = (qi & 1u) ? sn : cs;
    if (qi & 2u) s = -s;
    if ((qi + 1u) & 2u) c = -c;
    *s_out = s;
    *c_out = c;
}
G_FN float g_sin(float x) { float s, c; g_sincos(x, &s, &c); return s; }
G_FN float g_cos(float x) { float s, c; g_sincos(x, &s, &c); return c; }

/* exp(x): k = rint(x*log2(e)); r = x - k*ln2 (two-term); e^r = 1 + r + r^2*P(r); scale by 2^k in two steps. */
G_FN float g_exp(float x)
{
    if (g_isnan(x)) return x + x;
    if (x > 88.7228394f) return g_float(0x7f800000u);
    if (x < -103.972084f) return 0.0f;
    float t = g_fma(x, 1.44269502f, G_MAGIC);
    float kf = t - G_MAGIC;
    float r = g_fma(kf, -6.93145752e-1f, x);
    r = g_fma(kf, -1.42860677e-6f, r);
    float p = 1.9875691500e-4f;
    p = g_fma(p, r, 1.3981999507e-3f);
    p = g_fma(p, r, 8.3334519073e-3f);
    p = g_fma(p, r, 4.1665795894e-2f);
    p = g_fma(p, r, 1.6666665459e-1f);
    p = g_fma(p, r, 5.0000001201e-1f);
    float e = g_fma(p, r * r, r) + 1.0f;
    int k = (int)kf;
    int k1 = k >> 1;              /* arithmetic shift: floor(k/2) */
    int k2 = k - k1;
    float s1 = g_float((uint32_t)(k1 + 127) << 23);
    float s2 = g_float((uint32_t)(k2 + 127) << 23);
    return (e * s1) * s2;
}

/* log(x), x > 0: frexp to [sqrt(1/2), sqrt(2)), degree-8 kernel (the classic single-precision coefficients), explicit fma.
 * log(0) = -inf, log(x<0) = NaN.  Used only by pow() in the post-process / sRGB paths. */
G_FN float g_log(float x)
{
    if (g_isnan(x) || x < 0.0f) return g_float(0x7fc00000u);
    if (x == 0.0f) return g_float(0xff800000u);
    if (x == g_float(0x7f800000u)) return x;
    int e = 0;
    if (x < 1.17549435e-38f) { x = x * 8388608.0f; e = -23; }
    uint32_t b = g_bits(x);
    e += (int)(b >> 23) - 126;
    float m = g_float((b & 0x007fffffu) | 0x3f000000u);          /* mantissa in [0.5, 1) */
    if (m < 0.707106781186547524f) { e -= 1; m = m + m - 1.0f; } else { m = m - 1.0f; }
    float z = m * m;
    float p = 7.0376836292e-2f;
    p = g_fma(p, m, -1.1514610310e-1f);
    p = g_fma(p, m, 1.1676998740e-1f);
    p = g_fma(p, m, -1.2420140846e-1f);
    p = g_fma(p, m, 1.4249322787e-1f);
    p = g_fma(p, m, -1.6668057665e-1f);
    p = g_fma(p, m, 2.0000714765e-1f);
    p = g_fma(p, m, -2.4999993993e-1f);
    p = g_fma(p, m, 3.3333331174e-1f);
    float y = (p * z) * m;
    float fe = (float)e;
    y = g_fma(fe, -2.12194440e-4f, y);
    y = g_fma(z, -0.5f, y);
    float r = m + y;
    return g_fma(fe, 0.693359375f, r);
}
/* pow(x, y) for x >= 0 (GLSL: undefined for x < 0): exp(y * log(x)); pow(0, y > 0) = 0. */
G_FN float g_pow(float x, float y) { return g_exp(y * g_log(x)); }

#endif
