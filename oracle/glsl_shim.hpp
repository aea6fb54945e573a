/*
 * oracle/glsl_shim.hpp — TEST INFRASTRUCTURE ONLY (never linked into the product).
 *
 * A small GLSL 4.50 run-time for g++: the vector / matrix types, swizzles, built-in functions and
 * image / sampler objects that the reference's shaders use, so that the shaders THEMSELVES
 *   /root/reference/OpenTK-PathTracer/res/shaders/PathTracing/compute.glsl
 *   /root/reference/OpenTK-PathTracer/res/shaders/AtmosphericScattering/compute.glsl
 *   /root/reference/OpenTK-PathTracer/res/shaders/PostProcessing/fragment.glsl
 * can be compiled from where they lie (oracle/build_ref.py applies a handful of purely lexical rewrites, listed
 * there, and g++ compiles the result into oracle/_ref/libglsl_ref.so).  That library is "the reference itself,
 * run here": the hand-written restatement in pt_oracle.c / atmosphere_oracle.c is pinned against it bit for bit
 * (tests/test_reference_pin.py) and the golden images under tests/golden/ref_*.npz come from it.
 *
 * GLSL leaves operator / built-in precision implementation-defined, so the shim has to choose one; it uses the
 * scalar primitives of glsl_model.h (the documented evaluation model: correctly rounded + - * rcp sqrt, a/b = a*rcp(b),
 * dot and mat*vec as fma chains, fixed-polynomial sin / cos / exp / log).  Everything *structural* — which operations
 * run, in which order, on which operands, the draw order of the RNG, the closest-hit fold, the std140 field a value
 * comes from — is decided by the reference's source text, not by this repo.
 *
 * Scalars: the translated shader text says `Float` wherever the GLSL says `float` (build_ref.py rule R2/R3), a
 * one-member wrapper whose operators are the model's (so a scalar `a / b` is a * rcp(b) exactly like a vector one,
 * and nothing is ever evaluated in double or folded by the C++ compiler with different rounding).
 */
#ifndef PTO_GLSL_SHIM_HPP
#define PTO_GLSL_SHIM_HPP

#include <stdint.h>
#include <string.h>
#include <stddef.h>
#include <vector>

namespace gm {
extern "C++" {
#include "glsl_model.h"
}
}

/* GLSL_UNIFORM marks what the GLSL declared `uniform` (shared, written by the harness before a dispatch); GLSL_FN marks the
 * shader's functions; GLSL_HD the shim's.  Empty for g++; for nvcc (oracle/build_ref.py --cuda: the same translated text
 * compiled as the GL-compute proxy) uniforms live in __constant__ memory and every function is __host__ __device__. */
#ifdef __CUDACC__
#ifdef GLSL_UNIFORM_DEVICE              /* uniform blocks too large for the 64 KB constant bank (--capacity builds) */
#define GLSL_UNIFORM __device__
#else
#define GLSL_UNIFORM __constant__
#endif
#define GLSL_FN __device__
#define GLSL_HD __host__ __device__
#else
#define GLSL_UNIFORM
#define GLSL_FN
#define GLSL_HD
#endif

namespace glsl {

typedef unsigned int uint;

/* GLSL `float` */
struct Float {
    float v;
    Float() = default;
    GLSL_HD Float(float f) : v(f) {}
    GLSL_HD Float(int i) : v((float)i) {}          /* implicit int -> float (§4.1.10), exact below 2^24 */
    GLSL_HD Float(uint u) : v((float)u) {}         /* float(uint): round to nearest even */
    GLSL_HD explicit operator float() const { return v; }
};
GLSL_HD inline Float operator+(Float a, Float b) { return Float(a.v + b.v); }
GLSL_HD inline Float operator-(Float a, Float b) { return Float(a.v - b.v); }
GLSL_HD inline Float operator*(Float a, Float b) { return Float(a.v * b.v); }
#if defined(GLSL_SHIM_FAST_GPU) && defined(__CUDA_ARCH__)
/* Timing-only model for the nvcc build with -use_fast_math: approximate division and MUFU transcendentals, contraction left
 * to the compiler — roughly what a GL driver's compiler emits for this shader.  Never a parity reference. */
GLSL_HD inline Float operator/(Float a, Float b) { return Float(__fdividef(a.v, b.v)); }
#elif defined(GLSL_SHIM_ALT_MODEL)
/* A second, equally admissible evaluation model, used ONLY by tools/model_sensitivity.py to measure how much of the image
 * depends on the choice: IEEE-correct division, libm sinf / cosf / expf / powf, unfused dot / mix / mat*vec. */
extern "C" { float sinf(float); float cosf(float); float expf(float); float powf(float, float); }
GLSL_HD inline Float operator/(Float a, Float b) { return Float(a.v / b.v); }
#else
GLSL_HD inline Float operator/(Float a, Float b) { return Float(gm::g_div(a.v, b.v)); }
#endif
GLSL_HD inline Float operator-(Float a) { return Float(-a.v); }
GLSL_HD inline Float &operator+=(Float &a, Float b) { a = a + b; return a; }
GLSL_HD inline Float &operator-=(Float &a, Float b) { a = a - b; return a; }
GLSL_HD inline Float &operator*=(Float &a, Float b) { a = a * b; return a; }
GLSL_HD inline Float &operator/=(Float &a, Float b) { a = a / b; return a; }
GLSL_HD inline bool operator<(Float a, Float b) { return a.v < b.v; }
GLSL_HD inline bool operator>(Float a, Float b) { return a.v > b.v; }
GLSL_HD inline bool operator<=(Float a, Float b) { return a.v <= b.v; }
GLSL_HD inline bool operator>=(Float a, Float b) { return a.v >= b.v; }
GLSL_HD inline bool operator==(Float a, Float b) { return a.v == b.v; }
GLSL_HD inline bool operator!=(Float a, Float b) { return a.v != b.v; }

struct vec2; struct vec3; struct vec4; struct ivec2; struct ivec3; struct uvec2; struct uvec3; struct bvec3;

struct ivec2 { int x, y; ivec2() = default; GLSL_HD ivec2(int a, int b) : x(a), y(b) {} GLSL_HD explicit ivec2(const uvec2 &v); };
struct uvec2 { uint x, y; uvec2() = default; GLSL_HD uvec2(uint a, uint b) : x(a), y(b) {} };
struct ivec3 {
    int x, y, z; ivec3() = default; GLSL_HD ivec3(int a, int b, int c) : x(a), y(b), z(c) {} GLSL_HD explicit ivec3(const uvec3 &v);
    GLSL_HD ivec2 xy() const { return ivec2(x, y); }
};
struct uvec3 {
    uint x, y, z; uvec3() = default; GLSL_HD uvec3(uint a, uint b, uint c) : x(a), y(b), z(c) {}
    GLSL_HD uvec2 xy() const { return uvec2(x, y); }
};
GLSL_HD inline ivec2::ivec2(const uvec2 &v) : x((int)v.x), y((int)v.y) {}
GLSL_HD inline ivec3::ivec3(const uvec3 &v) : x((int)v.x), y((int)v.y), z((int)v.z) {}
struct bvec3 { bool x, y, z; };

/* The one-argument "splat" constructors are explicit, as in GLSL (a scalar never converts to a vector implicitly). */
struct vec2 {
    Float x, y;
    vec2() = default;
    GLSL_HD explicit vec2(Float s) : x(s), y(s) {}
    GLSL_HD vec2(Float a, Float b) : x(a), y(b) {}
    GLSL_HD vec2(const ivec2 &v) : x(v.x), y(v.y) {}                  /* GLSL implicit conversion ivec2 -> vec2 (§4.1.10) */
    GLSL_HD vec2 xy() const { return *this; }
};
struct vec3 {
    Float x, y, z;
    vec3() = default;
    GLSL_HD explicit vec3(Float s) : x(s), y(s), z(s) {}
    GLSL_HD vec3(Float a, Float b, Float c) : x(a), y(b), z(c) {}
    GLSL_HD explicit vec3(const bvec3 &b) : x(b.x ? 1.0f : 0.0f), y(b.y ? 1.0f : 0.0f), z(b.z ? 1.0f : 0.0f) {}
    GLSL_HD vec2 xy() const { return vec2(x, y); }
    GLSL_HD vec3 xyz() const { return *this; }
    GLSL_HD vec3 rgb() const { return *this; }
};
/* assignable two-component swizzle (`rayEye.zw = vec2(...)`) */
struct swz2_ref {
    Float &a, &b;
    GLSL_HD swz2_ref &operator=(const vec2 &v) { a = v.x; b = v.y; return *this; }
    GLSL_HD operator vec2() const { return vec2(a, b); }
};
struct vec4 {
    Float x, y, z, w;
    vec4() = default;
    GLSL_HD explicit vec4(Float s) : x(s), y(s), z(s), w(s) {}
    GLSL_HD vec4(Float a, Float b, Float c, Float d) : x(a), y(b), z(c), w(d) {}
    GLSL_HD vec4(const vec2 &v, Float c, Float d) : x(v.x), y(v.y), z(c), w(d) {}
    GLSL_HD vec4(const vec3 &v, Float d) : x(v.x), y(v.y), z(v.z), w(d) {}
    GLSL_HD vec2 xy() const { return vec2(x, y); }
    GLSL_HD vec3 xyz() const { return vec3(x, y, z); }
    GLSL_HD vec3 rgb() const { return vec3(x, y, z); }
    GLSL_HD swz2_ref zw() { return swz2_ref{ z, w }; }
};

/* ---- arithmetic: component-wise on Float (so `/` is a * rcp(b) everywhere) ----------------------------------------- */
#define GLSL_SHIM_OP(OP)                                                                                                           \
    GLSL_HD inline vec2 operator OP(const vec2 &a, const vec2 &b) { return vec2(a.x OP b.x, a.y OP b.y); }                         \
    GLSL_HD inline vec2 operator OP(const vec2 &a, Float b) { return vec2(a.x OP b, a.y OP b); }                                   \
    GLSL_HD inline vec2 operator OP(Float a, const vec2 &b) { return vec2(a OP b.x, a OP b.y); }                                   \
    GLSL_HD inline vec3 operator OP(const vec3 &a, const vec3 &b) { return vec3(a.x OP b.x, a.y OP b.y, a.z OP b.z); }             \
    GLSL_HD inline vec3 operator OP(const vec3 &a, Float b) { return vec3(a.x OP b, a.y OP b, a.z OP b); }                         \
    GLSL_HD inline vec3 operator OP(Float a, const vec3 &b) { return vec3(a OP b.x, a OP b.y, a OP b.z); }                         \
    GLSL_HD inline vec4 operator OP(const vec4 &a, const vec4 &b) { return vec4(a.x OP b.x, a.y OP b.y, a.z OP b.z, a.w OP b.w); } \
    GLSL_HD inline vec4 operator OP(const vec4 &a, Float b) { return vec4(a.x OP b, a.y OP b, a.z OP b, a.w OP b); }               \
    GLSL_HD inline vec4 operator OP(Float a, const vec4 &b) { return vec4(a OP b.x, a OP b.y, a OP b.z, a OP b.w); }
GLSL_SHIM_OP(+) GLSL_SHIM_OP(-) GLSL_SHIM_OP(*) GLSL_SHIM_OP(/)

GLSL_HD inline vec2 operator-(const vec2 &a) { return vec2(-a.x, -a.y); }
GLSL_HD inline vec3 operator-(const vec3 &a) { return vec3(-a.x, -a.y, -a.z); }
GLSL_HD inline vec4 operator-(const vec4 &a) { return vec4(-a.x, -a.y, -a.z, -a.w); }

/* compound assignment: `a op= b` is `a = a op b` (GLSL §5.8) */
#define GLSL_SHIM_COMPOUND(T)                                               \
    GLSL_HD inline T &operator+=(T &a, const T &b) { a = a + b; return a; } \
    GLSL_HD inline T &operator-=(T &a, const T &b) { a = a - b; return a; } \
    GLSL_HD inline T &operator*=(T &a, const T &b) { a = a * b; return a; } \
    GLSL_HD inline T &operator/=(T &a, const T &b) { a = a / b; return a; } \
    GLSL_HD inline T &operator+=(T &a, Float b) { a = a + b; return a; }    \
    GLSL_HD inline T &operator-=(T &a, Float b) { a = a - b; return a; }    \
    GLSL_HD inline T &operator*=(T &a, Float b) { a = a * b; return a; }    \
    GLSL_HD inline T &operator/=(T &a, Float b) { a = a / b; return a; }
GLSL_SHIM_COMPOUND(vec2) GLSL_SHIM_COMPOUND(vec3) GLSL_SHIM_COMPOUND(vec4)

/* ---- built-in functions (GLSL 4.50 §8) ------------------------------------------------------------------------------ */
GLSL_HD inline Float abs(Float a) { return gm::g_abs(a.v); }
GLSL_HD inline Float sign(Float a) { return gm::g_sign(a.v); }
GLSL_HD inline Float sqrt(Float a) { return gm::g_sqrt(a.v); }
#if defined(GLSL_SHIM_FAST_GPU) && defined(__CUDA_ARCH__)
GLSL_HD inline Float sin(Float a) { return __sinf(a.v); }
GLSL_HD inline Float cos(Float a) { return __cosf(a.v); }
GLSL_HD inline Float exp(Float a) { return __expf(a.v); }
#elif defined(GLSL_SHIM_ALT_MODEL)
GLSL_HD inline Float sin(Float a) { return ::glsl::sinf(a.v); }
GLSL_HD inline Float cos(Float a) { return ::glsl::cosf(a.v); }
GLSL_HD inline Float exp(Float a) { return ::glsl::expf(a.v); }
#else
GLSL_HD inline Float sin(Float a) { return gm::g_sin(a.v); }
GLSL_HD inline Float cos(Float a) { return gm::g_cos(a.v); }
GLSL_HD inline Float exp(Float a) { return gm::g_exp(a.v); }
#endif
GLSL_HD inline Float min(Float a, Float b) { return gm::g_min(a.v, b.v); }
GLSL_HD inline Float max(Float a, Float b) { return gm::g_max(a.v, b.v); }
GLSL_HD inline Float step(Float edge, Float x) { return gm::g_step(edge.v, x.v); }
#if defined(GLSL_SHIM_ALT_MODEL) || (defined(GLSL_SHIM_FAST_GPU) && defined(__CUDA_ARCH__))
GLSL_HD inline Float fma_(Float a, Float b, Float c) { return a * b + c; }
#else
GLSL_HD inline Float fma_(Float a, Float b, Float c) { return gm::g_fma(a.v, b.v, c.v); }
#endif
/* mix(x, y, a) = x*(1-a) + y*a, the y*a product fused into the sum */
GLSL_HD inline Float mix(Float x, Float y, Float a) { return fma_(y, a, x * (Float(1.0f) - a)); }
GLSL_HD inline Float clamp(Float x, Float lo, Float hi) { return min(max(x, lo), hi); }
/* pow with the constant exponents the shaders use is strength-reduced as the model states; otherwise exp(y*log(x)). */
GLSL_HD inline Float pow(Float x, Float y)
{
#if defined(GLSL_SHIM_FAST_GPU) && defined(__CUDA_ARCH__)
    return __powf(x.v, y.v);
#elif defined(GLSL_SHIM_ALT_MODEL)
    return ::glsl::powf(x.v, y.v);
#else
    if (y.v == 5.0f) { Float x2 = x * x; Float x4 = x2 * x2; return x4 * x; }
    if (y.v == 1.5f) return x * sqrt(x);
    return gm::g_exp((y * Float(gm::g_log(x.v))).v);
#endif
}

GLSL_HD inline vec3 exp(const vec3 &a) { return vec3(exp(a.x), exp(a.y), exp(a.z)); }
GLSL_HD inline vec3 min(const vec3 &a, const vec3 &b) { return vec3(min(a.x, b.x), min(a.y, b.y), min(a.z, b.z)); }
GLSL_HD inline vec3 max(const vec3 &a, const vec3 &b) { return vec3(max(a.x, b.x), max(a.y, b.y), max(a.z, b.z)); }
GLSL_HD inline vec3 pow(const vec3 &a, const vec3 &b) { return vec3(pow(a.x, b.x), pow(a.y, b.y), pow(a.z, b.z)); }
GLSL_HD inline vec3 clamp(const vec3 &v, Float lo, Float hi) { return vec3(clamp(v.x, lo, hi), clamp(v.y, lo, hi), clamp(v.z, lo, hi)); }
GLSL_HD inline vec3 mix(const vec3 &x, const vec3 &y, Float a) { return vec3(mix(x.x, y.x, a), mix(x.y, y.y, a), mix(x.z, y.z, a)); }
GLSL_HD inline vec3 mix(const vec3 &x, const vec3 &y, const vec3 &a) { return vec3(mix(x.x, y.x, a.x), mix(x.y, y.y, a.y), mix(x.z, y.z, a.z)); }
GLSL_HD inline bvec3 lessThan(const vec3 &a, const vec3 &b) { return bvec3{ a.x < b.x, a.y < b.y, a.z < b.z }; }

/* dot = FMUL, FFMA, FFMA */
GLSL_HD inline Float dot(const vec3 &a, const vec3 &b) { return fma_(a.z, b.z, fma_(a.y, b.y, a.x * b.x)); }
GLSL_HD inline Float length(const vec3 &a) { return sqrt(dot(a, a)); }
#if defined(GLSL_SHIM_FAST_GPU) && defined(__CUDA_ARCH__)
GLSL_HD inline vec3 normalize(const vec3 &a) { return a * Float(rsqrtf(dot(a, a).v)); }
#else
GLSL_HD inline vec3 normalize(const vec3 &a) { return a * (Float(1.0f) / sqrt(dot(a, a))); }
#endif
/* §8.5: reflect = I - 2.0 * dot(N, I) * N */
GLSL_HD inline vec3 reflect(const vec3 &I, const vec3 &N) { return I - Float(2.0f) * dot(N, I) * N; }
/* §8.5: k = 1.0 - eta * eta * (1.0 - dot(N, I) * dot(N, I)); k < 0 ? 0 : eta * I - (eta * dot(N, I) + sqrt(k)) * N */
GLSL_HD inline vec3 refract(const vec3 &I, const vec3 &N, Float eta)
{
    Float d = dot(N, I);
    Float k = Float(1.0f) - eta * eta * (Float(1.0f) - d * d);
    if (k < Float(0.0f)) return vec3(Float(0.0f));
    return eta * I - (eta * d + sqrt(k)) * N;
}

/* column-major 4x4 (std140 memory order); M * v per row as an fma chain over the columns */
struct mat4 { float c[4][4]; };
GLSL_HD inline vec4 operator*(const mat4 &M, const vec4 &v)
{
    Float r[4];
    for (int i = 0; i < 4; i++) {
        Float acc = Float(M.c[0][i]) * v.x;
        acc = fma_(M.c[1][i], v.y, acc);
        acc = fma_(M.c[2][i], v.z, acc);
        acc = fma_(M.c[3][i], v.w, acc);
        r[i] = acc;
    }
    return vec4(r[0], r[1], r[2], r[3]);
}

/* ---- images ------------------------------------------------------------------------------------------------------------ */
struct image2D { float *texels; int width, height; };       /* rgba32f, row 0 = y 0 */
GLSL_HD inline ivec2 imageSize(const image2D &im) { return ivec2(im.width, im.height); }
GLSL_HD inline vec4 imageLoad(const image2D &im, const ivec2 &p)
{
    if (p.x < 0 || p.y < 0 || p.x >= im.width || p.y >= im.height) return vec4(Float(0.0f));      /* OOB load -> 0 */
    const float *t = im.texels + ((size_t)p.y * im.width + p.x) * 4;
    return vec4(t[0], t[1], t[2], t[3]);
}
GLSL_HD inline void imageStore(const image2D &im, const ivec2 &p, const vec4 &v)
{
    if (p.x < 0 || p.y < 0 || p.x >= im.width || p.y >= im.height) return;                 /* OOB store discarded */
    float *t = im.texels + ((size_t)p.y * im.width + p.x) * 4;
    t[0] = v.x.v; t[1] = v.y.v; t[2] = v.z.v; t[3] = v.w.v;
}
struct imageCube { float *texels; int size; };               /* 6 layers (+X,-X,+Y,-Y,+Z,-Z) of size x size rgba32f */
GLSL_HD inline ivec2 imageSize(const imageCube &im) { return ivec2(im.size, im.size); }
GLSL_HD inline void imageStore(const imageCube &im, const ivec3 &p, const vec4 &v)
{
    if (p.x < 0 || p.y < 0 || p.z < 0 || p.x >= im.size || p.y >= im.size || p.z >= 6) return;
    float *t = im.texels + (((size_t)p.z * im.size + p.y) * im.size + p.x) * 4;
    t[0] = v.x.v; t[1] = v.y.v; t[2] = v.z.v; t[3] = v.w.v;
}

/* ---- sampler2D: only ever sampled 1:1 at texel centres by the full-screen post-process pass (NEAREST == LINEAR there);
 * an unbound unit (texels == nullptr) returns (0,0,0,1) ------------------------------------------------------------------- */
struct sampler2D { const float *texels; int width, height; };
GLSL_HD inline vec4 texture(const sampler2D &s, const vec2 &uv)
{
    if (!s.texels) return vec4(0.0f, 0.0f, 0.0f, 1.0f);
    int i = gm::g_f2i(gm::g_floor(uv.x.v * (float)s.width)), j = gm::g_f2i(gm::g_floor(uv.y.v * (float)s.height));
    i = i < 0 ? 0 : (i >= s.width ? s.width - 1 : i);
    j = j < 0 ? 0 : (j >= s.height ? s.height - 1 : j);
    const float *t = s.texels + ((size_t)j * s.width + i) * 4;
    return vec4(t[0], t[1], t[2], t[3]);
}

/* ---- samplerCube: the GL texture unit, not shader code.  OpenGL 4.5 §8.13 Table 8.19 face selection, LOD 0 => LINEAR
 * (AtmosphericScatterer.cs:67-69), GL_TEXTURE_CUBE_MAP_SEAMLESS (MainWindow.cs:168).  Written independently of
 * pt_oracle.c's lattice walk: every face is padded once with a one-texel border fetched through the 3-D position of the
 * border texel's centre (CubePadder, host side), then a lookup is four taps + three lerps.  Model choices shared with
 * s = 0.5 * (sc / |ma| + 1), u = s * N - 0.5, weights = fract, lerp = mix(); a corner border texel is the mean of
 * the three texels meeting at that cube corner (sum in face order, times fl(1/3)); a non-finite or zero direction fetches 0. */
struct samplerCube {
    int size;
    const float *padded;         /* [6][size+2][size+2][4], host or device memory */
    GLSL_HD const float *at(int f, int i, int j) const { return padded + ((((size_t)f * (size + 2)) + (j + 1)) * (size + 2) + (i + 1)) * 4; }
};
/* builds the padded faces on the host */
struct CubePadder {
    int size = 0;
    std::vector<float> padded;
    float *at(int f, int i, int j) { return &padded[((((size_t)f * (size + 2)) + (j + 1)) * (size + 2) + (i + 1)) * 4]; }

    /* Table 8.19 as integer frames: a point of face f is major*N + U*a + V*b with a = 2i+1-N, b = 2j+1-N */
    static void frame(int f, int major[3], int U[3], int V[3])
    {
        static const int M[6][3] = { { 1, 0, 0 }, { -1, 0, 0 }, { 0, 1, 0 }, { 0, -1, 0 }, { 0, 0, 1 }, { 0, 0, -1 } };
        static const int UU[6][3] = { { 0, 0, -1 }, { 0, 0, 1 }, { 1, 0, 0 }, { 1, 0, 0 }, { 1, 0, 0 }, { -1, 0, 0 } };
        static const int VV[6][3] = { { 0, -1, 0 }, { 0, -1, 0 }, { 0, 0, 1 }, { 0, 0, -1 }, { 0, -1, 0 }, { 0, -1, 0 } };
        for (int k = 0; k < 3; k++) { major[k] = M[f][k]; U[k] = UU[f][k]; V[k] = VV[f][k]; }
    }
    /* texel of the cube whose centre is the doubled-lattice point P (exactly one |coordinate| == N) */
    static void locate(const int P[3], int N, int *f, int *i, int *j)
    {
        for (int g = 0; g < 6; g++) {
            int m[3], U[3], V[3];
            frame(g, m, U, V);
            if (P[0] * m[0] + P[1] * m[1] + P[2] * m[2] == N) {
                int a = P[0] * U[0] + P[1] * U[1] + P[2] * U[2], b = P[0] * V[0] + P[1] * V[1] + P[2] * V[2];
                *f = g; *i = (a + N - 1) / 2; *j = (b + N - 1) / 2;
                return;
            }
        }
        *f = 0; *i = 0; *j = 0;
    }
    samplerCube upload(const float *faces, int N)
    {
        size = N;
        padded.assign((size_t)6 * (N + 2) * (N + 2) * 4, 0.0f);
        auto src = [&](int f, int i, int j) { return faces + (((size_t)f * N + j) * N + i) * 4; };
        for (int f = 0; f < 6; f++) {
            int m[3], U[3], V[3];
            frame(f, m, U, V);
            for (int j = -1; j <= N; j++)
                for (int i = -1; i <= N; i++) {
                    float *dst = at(f, i, j);
                    bool oi = i < 0 || i >= N, oj = j < 0 || j >= N;
                    if (!oi && !oj) { memcpy(dst, src(f, i, j), 16); continue; }
                    int ci = i < 0 ? 0 : (i >= N ? N - 1 : i), cj = j < 0 ? 0 : (j >= N ? N - 1 : j);
                    int a = 2 * ci + 1 - N, b = 2 * cj + 1 - N;      /* nearest in-face texel */
                    if (oi && oj) {
                        /* the three texels at this cube corner, in ascending face order */
                        int sgn[3];
                        for (int k = 0; k < 3; k++) { int p = m[k] * N + U[k] * a + V[k] * b; sgn[k] = p < 0 ? -1 : 1; }
                        float acc[4] = { 0, 0, 0, 0 };
                        bool first = true;
                        for (int g = 0; g < 6; g++) {
                            int axis = g / 2, want = (g % 2) ? -1 : 1;
                            if (sgn[axis] != want) continue;
                            int P[3];
                            for (int k = 0; k < 3; k++) P[k] = (k == axis) ? sgn[k] * N : sgn[k] * (N - 1);
                            int ff, ii, jj;
                            locate(P, N, &ff, &ii, &jj);
                            const float *t = src(ff, ii, jj);
                            for (int c = 0; c < 4; c++) acc[c] = first ? t[c] : acc[c] + t[c];
                            first = false;
                        }
                        for (int c = 0; c < 4; c++) dst[c] = acc[c] * 0.333333343f;
                    } else {
                        /* one step over an edge: the neighbour across it sits on the adjacent face, one half-texel in */
                        int P[3];
                        const int *E = oi ? U : V;                   /* direction that left the face */
                        int es = oi ? (i < 0 ? -1 : 1) : (j < 0 ? -1 : 1);
                        int along = oi ? b : a;
                        const int *A = oi ? V : U;
                        for (int k = 0; k < 3; k++) P[k] = m[k] * (N - 1) + E[k] * es * N + A[k] * along;
                        int ff, ii, jj;
                        locate(P, N, &ff, &ii, &jj);
                        memcpy(dst, src(ff, ii, jj), 16);
                    }
                }
        }
        return samplerCube{ N, padded.data() };
    }
};
GLSL_HD inline vec4 texture(const samplerCube &s, const vec3 &dir)
{
    const float FMAX = 3.4028235e+38f;
    const float rx = dir.x.v, ry = dir.y.v, rz = dir.z.v;
    float ax = gm::g_abs(rx), ay = gm::g_abs(ry), az = gm::g_abs(rz);
    if (!(ax <= FMAX && ay <= FMAX && az <= FMAX) || (ax == 0.0f && ay == 0.0f && az == 0.0f)) return vec4(Float(0.0f));
    int f; float sc, tc, ma;
    if (ax >= ay && ax >= az) { ma = ax; if (rx >= 0.0f) { f = 0; sc = -rz; tc = -ry; } else { f = 1; sc = rz; tc = -ry; } }
    else if (ay >= ax && ay >= az) { ma = ay; if (ry >= 0.0f) { f = 2; sc = rx; tc = rz; } else { f = 3; sc = rx; tc = -rz; } }
    else { ma = az; if (rz >= 0.0f) { f = 4; sc = rx; tc = -ry; } else { f = 5; sc = -rx; tc = -ry; } }
    const int N = s.size;
    float ima = gm::g_rcp(ma);
    float u = (0.5f * (sc * ima + 1.0f)) * (float)N - 0.5f;
    float v = (0.5f * (tc * ima + 1.0f)) * (float)N - 0.5f;
    float fu = gm::g_floor(u), fv = gm::g_floor(v);
    float wa = u - fu, wb = v - fv;
    int i0 = gm::g_f2i(fu), j0 = gm::g_f2i(fv);
    i0 = i0 < -1 ? -1 : (i0 > N - 1 ? N - 1 : i0);
    j0 = j0 < -1 ? -1 : (j0 > N - 1 ? N - 1 : j0);
    const float *t00 = s.at(f, i0, j0), *t10 = s.at(f, i0 + 1, j0), *t01 = s.at(f, i0, j0 + 1), *t11 = s.at(f, i0 + 1, j0 + 1);
    float out[4];
    for (int c = 0; c < 4; c++)
        out[c] = gm::g_mix(gm::g_mix(t00[c], t10[c], wa), gm::g_mix(t01[c], t11[c], wa), wb);
    return vec4(out[0], out[1], out[2], out[3]);
}

} /* namespace glsl */
#endif
