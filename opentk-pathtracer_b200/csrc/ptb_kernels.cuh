// ptb_kernels.cuh — device code of libptb200: the path-tracing megakernel and its helpers (sm_100a).
//
// What it replaces: res/shaders/PathTracing/compute.glsl (cited below as pt:LINE) dispatched by
// src/Render/PathTracer.cs:114-129.  This is a re-design, not a translation:
//   * a persistent grid (CTAs resident on every SM) pulls pixels from one atomic work counter;
//   * each warp keeps 32 paths in registers and, after every bounce, refills the lanes whose pixel finished
//     (ballot + prefix-popcount slot assignment), so lanes at different bounce depths stay in one coherent
//     trace->shade loop instead of idling until the longest path of an 8x8 group ends;
//   * the scene is repacked once per edit into an SoA block (sphere centre + r^2, slab bounds, 1/r, materials)
//     and staged HBM -> shared memory with one TMA bulk copy (cp.async.bulk + mbarrier) per CTA;
//   * the closest-hit fold keeps only (T, primitive, fromInside); position / normal / material are produced once
//     for the winner;
//   * the cubemap is stored with a 1-texel seamless border so a miss is four aligned 128-bit loads and three lerps;
//   * results leave as 128-bit stores.
// Arithmetic is the evaluation model of ptb_math.cuh (bit-exact against the CPU oracle in tests/).
#pragma once
#include "ptb_math.cuh"

namespace ptb {

constexpr float kFloatMax = 3.4028235e+38f;   // pt:2
constexpr float kFloatMin = -3.4028235e+38f;  // pt:3
constexpr float kEps = 0.001f;                // pt:4
constexpr float kPi = 3.14159265f;            // pt:5

constexpr int kSphereStride = 80;   // Sphere.cs:8
constexpr int kCuboidStride = 96;   // Cuboid.cs:8
#ifndef PTB_THREADS
#define PTB_THREADS 256
#endif
#ifndef PTB_MIN_BLOCKS
#define PTB_MIN_BLOCKS 4       // <= 64 registers: four 256-thread CTAs per SM (measured: 80 registers / 3 CTAs costs 5 %)
#endif
#ifndef PTB_REFILL_MIN
#define PTB_REFILL_MIN 1      // refill as soon as this many lanes of a warp are idle
#endif
// (Round-1 experiment, removed: FADD2/FMUL2/FFMA2 packed fp32x2 in the fold, two spheres per instruction.  Bit-identical,
//  but on B200 +3 % on the default scene and -15 % on the 1280-primitive scene — the packed ops issue at half rate — and
//  the pair layout alone cost the scalar loop 20 % on the large scene.  See profiles/r01_experiments.md.)
#ifndef PTB_COOP_MAX
#define PTB_COOP_MAX 8        // tail: at most this many live paths in a starved warp -> group-cooperative fold (0 disables)
#endif
__device__ __forceinline__ unsigned long long global_ns()
{
    unsigned long long t;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
    return t;
}

__device__ __forceinline__ unsigned ld_acquire_gpu(const unsigned* p)
{
    unsigned v;
    asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ void st_release_gpu(unsigned* p, unsigned v)
{
    asm volatile("st.release.gpu.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}

constexpr int kMegaThreads = PTB_THREADS;
constexpr int kQueue = 64;          // primary-ray ring entries per warp (two tiles)
constexpr int kDoneWords = 11;      // completion-queue entry: direction, throughput, radiance (3 floats each), px | lrow << 16, frame slot | needs-env << 8

// ------------------------------------------------------------------------------------------------------------
// Launch parameters (by value: they live in the constant bank; UBO 0 is small enough to ride along).
struct RenderParams {
    float basic[36];          // BasicDataUBO: InvProjection @0, InvView @16, ViewPos @32  (pt:59-64)
    int width, height;        // imageSize(ImgResult) — the FULL image, also for tiles
    int frame;                // thisRendererFrame (pt:96)
    int spp, ray_depth;       // pt:90-91
    float focal_length, aperture_diameter;  // pt:93-94
    int n_spheres, n_cuboids; // uboGameObjectsSize (pt:88)
    int env_size;             // cubemap face edge N (padded faces are (N+2)^2)
    int rank, world, stripe_rows, local_rows;  // pixel-tile partition (rows y with (y/stripe_rows)%world==rank)
    // packed scene block (float4 units from the block base)
    int off_aux, off_cmin, off_cmax, off_mat, block_bytes;
    int stage_bytes;          // bytes of the block staged into shared memory (== block_bytes unless materials stay in HBM)
    int off_nodes, off_pidx, n_nodes, n_unbounded;   // BVH (float4 units from the block base); n_nodes == 0: brute force
    float bvh_tau;
    const float4* scene;      // packed scene block in HBM
    const float4* env;        // padded cubemap
    float4* image;            // local accumulation image: local_rows x width
    float4* scratch;          // pipelined mode: this frame's own estimate goes here (no read of `image`); blend_kernel folds it in
    unsigned int* counters;   // [0] work counter, [1] finished-CTA counter
    unsigned long long* stats;// optional: samples, traces, hits
    const unsigned char* raw_objects; // raw GameObjectsUBO bytes (naive proxy only)
    int max_spheres;
    // host-precomputed, correctly rounded (1.0f / x on the host == rcp.rn on the device)
    float inv_width, inv_height, inv_spp, blend;   // 1/W, 1/H, 1/SPP, 1/(frame+1)
    unsigned tiles_x, tiles_total;                  // 8x4 work tiles
    unsigned tiles_magic;                           // ceil(2^32 / tiles_x): tile / tiles_x == umulhi(tile, magic) while tile * tiles_x < 2^32 (0 = divide)
    // frame batching (megakernel<..., kBatch = true> only; appended so that the offsets above never move): one launch traces
    // `batch` consecutive frames, frame P.frame + b into scratch + b * scratch_stride; work index = b * tiles_total + tile
    int batch;
    unsigned long long scratch_stride;              // float4 elements between the scratch images of consecutive frames
    // ray-classification table (megakernel<kFold = 2>; scenes of at most 64 primitives): one 64-bit candidate mask per
    // (origin cell, direction bucket); bit i = primitive i (spheres first) may be hit by some ray of that class
    const unsigned long long* rct;
    float rct_lo[3], rct_inv[3];                    // grid origin, 1 / cell size
    int rct_n[3], rct_G;                            // cells per axis, direction buckets per cube-face axis
    float rct_halfG;
    unsigned rct_sm0, rct_sm1;                      // which bits of the mask's low / high word are spheres
    unsigned long long* ktime;                      // optional {min CTA start, max CTA end} in globaltimer ns (ptb_set_kernel_timing)
    int defer_finish;                               // megakernel with the ring, SPP == 1: finished pixels go through the warp's completion queue
    unsigned* done_flag;                            // optional: the last CTA out stores done_value here (release): the batch's blend
    unsigned done_value;                            //   kernel, already resident, spins on it instead of waiting for a stream event
    // uniform grid for large scenes (megakernel<kFold = 3>): cell_start[] (u16) at off_gcell, items[] (u16) at off_gitem (float4 units)
    int off_gcell, off_gsph, off_gitem, grid_n[3];
    float grid_lo[3], grid_hi[3], grid_cell[3], grid_inv[3];
    float rct_nf[3];                                // cells per axis as floats (range test of the cell coordinates)
    unsigned long long rct_valid;                   // the mask of every existing primitive: what an unclassifiable ray tests
};

// ------------------------------------------------------------------------------------------------------------
// Scene views.  Both expose the same accessors; the fold and the shader are written once against them.
struct PackedScene {           // SoA block in shared memory
    const float4* base;
    const float4* mats;        // materials: in the shared block, or (large scenes with a BVH) left in HBM / L2
    const float4* nodes;       // BVH nodes, 2 x float4 each (large scenes only)
    const int* pidx;           // BVH primitive index list: [unbounded primitives][leaf contents]
    int n_nodes, n_unbounded;
    float tau;                 // slack on ray parameters in the conservative box tests
    int nS, nC, off_aux, off_cmin, off_cmax, off_mat;
    __device__ __forceinline__ float4 sphere(int i) const { return base[i]; }                 // (c, r*r)
    // 1 / radius, needed once per hit: with the materials in the shared block for small scenes, with them in HBM / L2 for large ones
    __device__ __forceinline__ float sphere_rcp_r(int i) const { return reinterpret_cast<const float*>(mats + off_aux)[i]; }
    // slab bounds are interleaved (lo0, hi0, lo1, hi1, ...; off_cmax == off_cmin + 1): one address computation per cuboid
    __device__ __forceinline__ float4 cmin(int i) const { return base[off_cmin + 2 * i]; }
    __device__ __forceinline__ float4 cmax(int i) const { return base[off_cmax + 2 * i]; }
    __device__ __forceinline__ float4 mat(int prim, int k) const { return mats[off_mat + prim * 4 + k]; }
};
struct RawScene {              // the std140 bytes exactly as the host uploaded them (naive proxy)
    const unsigned char* ubo;
    int nS, nC, max_spheres;
    __device__ __forceinline__ float4 sphere(int i) const
    {
        float4 s = __ldg(reinterpret_cast<const float4*>(ubo + (size_t)i * kSphereStride));
        s.w = s.w * s.w;
        return s;
    }
    __device__ __forceinline__ float sphere_rcp_r(int i) const
    {
        return rcp(__ldg(reinterpret_cast<const float4*>(ubo + (size_t)i * kSphereStride)).w);
    }
    __device__ __forceinline__ const unsigned char* cub(int i) const { return ubo + (size_t)max_spheres * kSphereStride + (size_t)i * kCuboidStride; }
    __device__ __forceinline__ float4 cmin(int i) const { return __ldg(reinterpret_cast<const float4*>(cub(i))); }
    __device__ __forceinline__ float4 cmax(int i) const { return __ldg(reinterpret_cast<const float4*>(cub(i) + 16)); }
    __device__ __forceinline__ float4 mat(int prim, int k) const
    {
        const unsigned char* p = prim < nS ? ubo + (size_t)prim * kSphereStride + 16 : cub(prim - nS) + 32;
        return __ldg(reinterpret_cast<const float4*>(p) + k);
    }
};

__device__ __forceinline__ PackedScene make_packed_scene(const float4* smem_block, const float4* global_block, int nS, int nC,
                                                        int off_aux, int off_cmin, int off_cmax, int off_mat, int stage_bytes, int block_bytes,
                                                        int off_nodes, int off_pidx, int n_nodes, int n_unbounded, float tau)
{
    PackedScene sc;
    sc.base = smem_block;
    sc.mats = stage_bytes >= block_bytes ? smem_block : global_block;
    sc.nodes = smem_block + off_nodes;
    sc.pidx = reinterpret_cast<const int*>(smem_block + off_pidx);
    sc.n_nodes = n_nodes; sc.n_unbounded = n_unbounded; sc.tau = tau;
    sc.nS = nS; sc.nC = nC; sc.off_aux = off_aux; sc.off_cmin = off_cmin; sc.off_cmax = off_cmax; sc.off_mat = off_mat;
    return sc;
}
#define PTB_PACKED_SCENE(P, smem) make_packed_scene(smem, (P).scene, (P).n_spheres, (P).n_cuboids, (P).off_aux, (P).off_cmin, (P).off_cmax, \
                                                    (P).off_mat, (P).stage_bytes, (P).block_bytes, (P).off_nodes, (P).off_pidx, (P).n_nodes, (P).n_unbounded, (P).bvh_tau)

// ------------------------------------------------------------------------------------------------------------
// pt:226-294 — the order-dependent closest-hit fold (spheres, then cuboids; accept on hit && t2>0 && t1<T).
template <class Scene>
__device__ __forceinline__ void trace(const Scene& sc, V3 o, V3 d, float& T, int& prim, bool& inside)
{
    T = kFloatMax;
    prim = -1;
    inside = false;
#pragma unroll 4
    for (int i = 0; i < sc.nS; ++i) {
        const float4 s = sc.sphere(i);
        const V3 v = mk(o.x - s.x, o.y - s.y, o.z - s.z);
        const float b = dot(d, v);
        const float c = dot(v, v) - s.w;
        const float disc = b * b - c;
        if (!(disc < 0.0f)) {
            const float sq = fsqrt(disc);
            const float t1 = -b - sq;
            const float t2 = -b + sq;
            if (t1 <= t2 && t2 > 0.0f && t1 < T) {
                T = t1 < 0.0f ? t2 : t1;
                inside = (T == t2);
                prim = i;
            }
        }
    }
    const float ix = rcp(d.x), iy = rcp(d.y), iz = rcp(d.z);
#pragma unroll 2
    for (int i = 0; i < sc.nC; ++i) {
        const float4 lo = sc.cmin(i), hi = sc.cmax(i);
        const float ax = (lo.x - o.x) * ix, ay = (lo.y - o.y) * iy, az = (lo.z - o.z) * iz;
        const float bx = (hi.x - o.x) * ix, by = (hi.y - o.y) * iy, bz = (hi.z - o.z) * iz;
        const float t1 = fmax_(kFloatMin, fmax_(fmin_(ax, bx), fmax_(fmin_(ay, by), fmin_(az, bz))));
        const float t2 = fmin_(kFloatMax, fmin_(fmax_(ax, bx), fmin_(fmax_(ay, by), fmax_(az, bz))));
        if (t1 <= t2 && t2 > 0.0f && t1 < T) {
            T = t1 < 0.0f ? t2 : t1;
            inside = (T == t2);
            prim = sc.nS + i;
        }
    }
}

// ---- the fold's arithmetic, written once for every variant of the fold (plain, table, cooperative, BVH) -----------------
// State of the fold: (T, prim, t2w) with t2w = the winner's exit distance; HitInfo.FromInside (pt:237,250) is T == t2w.
// In the fast build the multiply-adds are contracted by hand, identically everywhere, so that a ray gets the same distances
// whichever variant of the fold happens to trace it (the frame's tail switches to the cooperative fold).
struct RayInv { V3 inv, noi; };       // 1 / d and (fast build) -o / d
__device__ __forceinline__ RayInv ray_inverse(V3 o, V3 d)
{
    RayInv r;
#ifdef PTB_FAST
    // a zero component would give inf * 0 = NaN in lo * inv - o * inv: 1e-18 keeps it a huge finite slope (and changes nothing
    // for |d| >= 1e-10, where it is below half an ulp)
    r.inv = mk(rcp(d.x + copysignf(1e-18f, d.x)), rcp(d.y + copysignf(1e-18f, d.y)), rcp(d.z + copysignf(1e-18f, d.z)));
    r.noi = mk(-o.x * r.inv.x, -o.y * r.inv.y, -o.z * r.inv.z);
#else
    r.inv = mk(rcp(d.x), rcp(d.y), rcp(d.z));
    r.noi = mk(0.0f, 0.0f, 0.0f);
#endif
    return r;
}
__device__ __forceinline__ void sphere_terms(const float4 s, V3 o, V3 d, float& b, float& disc)       // pt:263-266
{
    const V3 v = mk(o.x - s.x, o.y - s.y, o.z - s.z);
#ifdef PTB_FAST
    b = __fmaf_rn(d.z, v.z, __fmaf_rn(d.y, v.y, d.x * v.x));
    const float c = __fmaf_rn(v.z, v.z, __fmaf_rn(v.y, v.y, __fmaf_rn(v.x, v.x, -s.w)));
    disc = __fmaf_rn(b, b, -c);
#else
    b = dot(d, v);
    const float c = dot(v, v) - s.w;
    disc = b * b - c;
#endif
}
__device__ __forceinline__ void slab_terms(const float4 lo, const float4 hi, V3 o, const RayInv& ri, float& t1, float& t2)   // pt:282-291
{
#ifdef PTB_FAST
    const float ax = __fmaf_rn(lo.x, ri.inv.x, ri.noi.x), ay = __fmaf_rn(lo.y, ri.inv.y, ri.noi.y), az = __fmaf_rn(lo.z, ri.inv.z, ri.noi.z);
    const float bx = __fmaf_rn(hi.x, ri.inv.x, ri.noi.x), by = __fmaf_rn(hi.y, ri.inv.y, ri.noi.y), bz = __fmaf_rn(hi.z, ri.inv.z, ri.noi.z);
#else
    const float ax = (lo.x - o.x) * ri.inv.x, ay = (lo.y - o.y) * ri.inv.y, az = (lo.z - o.z) * ri.inv.z;
    const float bx = (hi.x - o.x) * ri.inv.x, by = (hi.y - o.y) * ri.inv.y, bz = (hi.z - o.z) * ri.inv.z;
#endif
    t1 = fmax_(kFloatMin, fmax_(fmin_(ax, bx), fmax_(fmin_(ay, by), fmin_(az, bz))));
    t2 = fmin_(kFloatMax, fmin_(fmax_(ax, bx), fmin_(fmax_(ay, by), fmax_(az, bz))));
}
__device__ __forceinline__ void accept(float t1, float t2, int i, float& T, int& prim, float& t2w)      // pt:234-240, 247-253
{
    if (t1 <= t2 && t2 > 0.0f && t1 < T) {
        T = t1 < 0.0f ? t2 : t1;
        t2w = t2;
        prim = i;
    }
}
__device__ __forceinline__ void accept_sphere(float b, float disc, int i, float& T, int& prim, float& t2w)
{
    if (!(disc < 0.0f)) {
        const float sq = fsqrt(disc);
        accept(-b - sq, -b + sq, i, T, prim, t2w);
    }
}
// The fold over the shared-memory block: spheres four at a time (the array is padded to a multiple of four with spheres of
// r^2 = -inf that nothing can hit), one max + branch for the four discriminants — a NaN or negative discriminant never hits,
// so dropping those through the NaN-ignoring max is exact — then the cuboids.
__device__ __forceinline__ void trace_any(const PackedScene& sc, V3 o, V3 d, float& T, int& prim, bool& inside)
{
    T = kFloatMax;
    prim = -1;
    float t2w = 0.0f;
    const int n4 = (sc.nS + 3) & ~3;
    for (int i = 0; i < n4; i += 4) {
        float b0, b1, b2, b3, d0, d1, d2, d3;
        sphere_terms(sc.base[i], o, d, b0, d0);
        sphere_terms(sc.base[i + 1], o, d, b1, d1);
        sphere_terms(sc.base[i + 2], o, d, b2, d2);
        sphere_terms(sc.base[i + 3], o, d, b3, d3);
        if (fmax_(fmax_(d0, d1), fmax_(d2, d3)) >= 0.0f) {
            accept_sphere(b0, d0, i, T, prim, t2w);
            accept_sphere(b1, d1, i + 1, T, prim, t2w);
            accept_sphere(b2, d2, i + 2, T, prim, t2w);
            accept_sphere(b3, d3, i + 3, T, prim, t2w);
        }
    }
    const RayInv ri = ray_inverse(o, d);
    const float4* cub = sc.base + sc.off_cmin;
#pragma unroll 2
    for (int i = 0; i < sc.nC; ++i) {
        float t1, t2;
        slab_terms(cub[2 * i], cub[2 * i + 1], o, ri, t1, t2);
        accept(t1, t2, sc.nS + i, T, prim, t2w);
    }
    inside = (T == t2w);
}
__device__ __forceinline__ void trace_any(const RawScene& sc, V3 o, V3 d, float& T, int& prim, bool& inside) { trace(sc, o, d, T, prim, inside); }

// pt:316-332 — surface normal of the winning primitive.
template <class Scene>
__device__ __forceinline__ V3 surface_normal(const Scene& sc, int prim, V3 pos)
{
    if (prim < sc.nS) {
        const float4 s = sc.sphere(prim);
        return mk(pos.x - s.x, pos.y - s.y, pos.z - s.z) * sc.sphere_rcp_r(prim);
    }
    const float4 lo = sc.cmin(prim - sc.nS), hi = sc.cmax(prim - sc.nS);
    const V3 half = mk(hi.x - lo.x, hi.y - lo.y, hi.z - lo.z) * 0.5f;
    const V3 cs = pos - mk(hi.x + lo.x, hi.y + lo.y, hi.z + lo.z) * 0.5f;
    V3 n;
    n.x = 0.0f + signf(cs.x) * stepf(fabsf(fabsf(cs.x) - half.x), kEps);
    n.y = 0.0f + signf(cs.y) * stepf(fabsf(fabsf(cs.y) - half.y), kEps);
    n.z = 0.0f + signf(cs.z) * stepf(fabsf(fabsf(cs.z) - half.z), kEps);
    // normalize(n) with n in {-1,0,1}^3: dot(n,n) is exactly 0..3, so rcp(sqrt(.)) is one of four constants
    // (inf for the zero vector: 0 * inf = NaN, as normalize(vec3(0)) gives).  0x3f3504f3 = rcp(sqrt(2)), 0x3f13cd3a = rcp(sqrt(3)).
    const float k = dot(n, n);
    const float inv = k == 1.0f ? 1.0f : (k == 2.0f ? __uint_as_float(0x3f3504f3u) : (k == 3.0f ? __uint_as_float(0x3f13cd3au) : rcp(fsqrt(k))));
    return n * inv;
}

// pt:297-307
__device__ __forceinline__ V3 cosine_hemisphere(V3 n, uint32_t& rng)
{
    const float z = rand01(rng) * 2.0f - 1.0f;
    const float a = rand01(rng) * 2.0f * kPi;
    const float r = fsqrt(1.0f - z * z);
    float sn, cs;
    sincos_(a, sn, cs);
    return normalize(n + mk(r * cs, r * sn, z));
}

// ------------------------------------------------------------------------------------------------------------
// texture(SamplerEnvironment, dir).rgb (pt:177): GL 4.5 cube face selection, LINEAR magnification, seamless.
// The border of each padded face already holds the neighbouring faces' texels (pad_cubemap_kernel).
__device__ __forceinline__ V3 env_lookup(const float4* __restrict__ env, int N, V3 r)
{
    const float ax = fabsf(r.x), ay = fabsf(r.y), az = fabsf(r.z);
    // non-finite or zero direction: defined to fetch 0 (a texture unit never returns NaN for NaN coords)
    if (!(ax <= kFloatMax && ay <= kFloatMax && az <= kFloatMax) || (ax == 0.0f && ay == 0.0f && az == 0.0f))
        return mk(0.0f, 0.0f, 0.0f);
    int f;
    float sc, tc, ma;
    if (ax >= ay && ax >= az) { ma = ax; f = r.x >= 0.0f ? 0 : 1; sc = r.x >= 0.0f ? -r.z : r.z; tc = -r.y; }
    else if (ay >= ax && ay >= az) { ma = ay; f = r.y >= 0.0f ? 2 : 3; sc = r.x; tc = r.y >= 0.0f ? r.z : -r.z; }
    else { ma = az; f = r.z >= 0.0f ? 4 : 5; sc = r.z >= 0.0f ? r.x : -r.x; tc = -r.y; }
    const float ima = rcp(ma);
    const float s = 0.5f * (sc * ima + 1.0f);
    const float t = 0.5f * (tc * ima + 1.0f);
    const float u = s * (float)N - 0.5f;
    const float v = t * (float)N - 0.5f;
    const float fu = floorf(u), fv = floorf(v);
    const float alpha = u - fu, beta = v - fv;
    int i0 = __float2int_rz(fu), j0 = __float2int_rz(fv);
    i0 = max(-1, min(N - 1, i0));
    j0 = max(-1, min(N - 1, j0));
    const int P = N + 2;
    const float4* p = env + ((size_t)f * P + (j0 + 1)) * P + (i0 + 1);
    const float4 t00 = __ldg(p), t10 = __ldg(p + 1), t01 = __ldg(p + P), t11 = __ldg(p + P + 1);
    const V3 top = mix(mk(t00.x, t00.y, t00.z), mk(t10.x, t10.y, t10.z), alpha);
    const V3 bot = mix(mk(t01.x, t01.y, t01.z), mk(t11.x, t11.y, t11.z), alpha);
    return mix(top, bot, beta);
}

// pt:177 — radiance += texture(env, dir).rgb * throughput, written once for the two places a miss is shaded (in the bounce loop,
// or later from the completion queue) so that the fast build contracts it the same way in both.
__device__ __forceinline__ V3 add_environment(V3 rad, V3 env, V3 thr)
{
#ifdef PTB_FAST
    return mk(__fmaf_rn(env.x, thr.x, rad.x), __fmaf_rn(env.y, thr.y, rad.y), __fmaf_rn(env.z, thr.z, rad.z));
#else
    return rad + env * thr;
#endif
}

// ------------------------------------------------------------------------------------------------------------
// One lane's path state.
struct Path {
    V3 o, d;            // current ray
    V3 thr, rad;        // throughput, radiance of the current sample
    V3 irr;             // sum of finished samples of this pixel (pt:109,123)
    uint32_t rng;       // rndSeed (pt:98) — one stream per pixel, continuous across samples and bounces
    int px, py;         // global pixel coordinates (gl_GlobalInvocationID.xy)
    int lrow;           // row inside this rank's local image
    int sample, depth;
    int fb;             // frame slot inside a batched launch (kBatch instantiations only; otherwise never touched)
};

// pt:110-121 — jitter, pinhole ray, thin-lens origin / direction.  Draw order: jitter.x, jitter.y, angle, radius.
__device__ __forceinline__ void primary_ray(const RenderParams& P, Path& p)
{
    const float ox = rand01(p.rng);
    const float oy = rand01(p.rng);
    const float ndcx = ((float)p.px + ox) * P.inv_width * 2.0f - 1.0f;
    const float ndcy = ((float)p.py + oy) * P.inv_height * 2.0f - 1.0f;
    const float* IP = P.basic;
    const float* IV = P.basic + 16;
    const float ex = mat_row(IP, 0, ndcx, ndcy, -1.0f, 0.0f);
    const float ey = mat_row(IP, 1, ndcx, ndcy, -1.0f, 0.0f);
    const V3 dir = normalize(mk(mat_row(IV, 0, ex, ey, -1.0f, 0.0f), mat_row(IV, 1, ex, ey, -1.0f, 0.0f), mat_row(IV, 2, ex, ey, -1.0f, 0.0f)));
    const V3 view = mk(P.basic[32], P.basic[33], P.basic[34]);
    const V3 focal = view + dir * P.focal_length;
    const float angle = rand01(p.rng) * 2.0f * kPi;
    const float rr = fsqrt(rand01(p.rng));
    float sn, cs;
    sincos_(angle, sn, cs);
    const float hk = P.aperture_diameter * 0.5f;
    const float offx = hk * (cs * rr), offy = hk * (sn * rr);
    p.o = mk(mat_row(IV, 0, offx, offy, 0.0f, 1.0f), mat_row(IV, 1, offx, offy, 0.0f, 1.0f), mat_row(IV, 2, offx, offy, 0.0f, 1.0f));
    p.d = normalize(focal - p.o);
    p.thr = mk(1.0f, 1.0f, 1.0f);
    p.rad = mk(0.0f, 0.0f, 0.0f);
    p.depth = 0;
}

// pt:140-180 — one iteration of the bounce loop for one lane, after the closest-hit fold gave (T, prim, inside).
// Returns kContinue while the sample continues, kDone when it ended on a surface, kMissed when the ray left the scene; with
// `defer_env` a miss leaves the environment lookup (pt:177) to the caller.
constexpr int kContinue = 0, kDone = 1, kMissed = 2;
template <class Scene>
__device__ __forceinline__ int shade(const RenderParams& P, const Scene& sc, Path& p, float T, int prim, bool inside, unsigned long long* stats,
                                     bool defer_env = false)
{
    if (stats) atomicAdd(stats + 1, 1ull);
    if (T != kFloatMax) {
        if (stats) atomicAdd(stats + 2, 1ull);
        const V3 pos = p.o + p.d * T;
        V3 n = surface_normal(sc, prim, pos);
        const float4 m0 = sc.mat(prim, 0), m1 = sc.mat(prim, 1), m2 = sc.mat(prim, 2), m3 = sc.mat(prim, 3);
        const float ior = m3.y;
        if (inside) {                                        // pt:145-149 Beer's law
            n = -n;
            // exp(-0 * T) is exactly 1 for finite T, so a medium without absorbance (clear glass) skips the three exponentials
            if (!(m2.x == 0.0f && m2.y == 0.0f && m2.z == 0.0f && T <= kFloatMax))
                p.thr = p.thr * mk(exp_(-m2.x * T), exp_(-m2.y * T), exp_(-m2.z * T));
        }
        // ---- BSDF, pt:184-224
        float spec = m0.w, refr = m2.w;
        if (spec > 0.0f) {
            const float n1 = inside ? ior : 1.0f, n2 = !inside ? ior : 1.0f;
            float r0 = fdiv(n1 - n2, n1 + n2);
            r0 *= r0;
            const float fres = r0 + (1.0f - r0) * pow5(1.0f - dot(-p.d, n));   // pt:359-364
            spec = mixf(spec, 1.0f, fres);
            const float diffuse_chance = 1.0f - spec - refr;
            refr = 1.0f - spec - diffuse_chance;
        }
        // RNG draws in the shader's order: hemisphere z, hemisphere angle, lobe roll, and for the refraction lobe a
        // second (z, angle) pair.  The refraction lobe never uses the first hemisphere DIRECTION (pt:210-211), only its draws, so
        // one hemisphere evaluation serves every lane: around -n from the second pair for refraction, around n from the first
        // pair otherwise.  Same operations on the same values as pt:297-307, once instead of twice per iteration.
        float uz = rand01(p.rng);
        float ua = rand01(p.rng);
        const float roll = rand01(p.rng);
        const bool lobe_spec = spec > roll;
        const bool lobe_refr = !lobe_spec && (spec + refr > roll);
        V3 hn = n;
        if (lobe_refr) { uz = rand01(p.rng); ua = rand01(p.rng); hn = -n; }
        const float hz = uz * 2.0f - 1.0f;
        const float ha = ua * 2.0f * kPi;
        const float hr = fsqrt(1.0f - hz * hz);
        float hs, hc;
        sincos_(ha, hs, hc);
        const V3 hemi = normalize(hn + mk(hr * hc, hr * hs, hz));
        const bool refractive = lobe_refr;
        float prob = 1.0f - spec - refr;
        V3 nd = hemi;
        if (lobe_spec || lobe_refr) {
            V3 x;
            float rough;
            if (lobe_spec) { x = reflect(p.d, n); rough = m1.w * m1.w; prob = spec; }
            else { x = refract(p.d, n, inside ? fdiv(ior, 1.0f) : fdiv(1.0f, ior)); rough = m3.x * m3.x; prob = refr; }
            nd = normalize(mix(x, hemi, rough));
        }
        p.d = nd;
        p.o = pos + nd * kEps;
        prob = fmax_(prob, kEps);
        // ---- pt:156-173
        p.rad = p.rad + mk(m1.x, m1.y, m1.z) * p.thr;
        if (!refractive) p.thr = p.thr * mk(m0.x, m0.y, m0.z);
        p.thr = p.thr * rcp(prob);
        const float q = fmax_(p.thr.x, fmax_(p.thr.y, p.thr.z));
        if (rand01(p.rng) > q) return kDone;
        p.thr = p.thr * rcp(q);
        return ++p.depth < P.ray_depth ? kContinue : kDone;
    }
    if (!defer_env) p.rad = add_environment(p.rad, env_lookup(P.env, P.env_size, p.d), p.thr);          // pt:177
    return kMissed;
}

template <class Scene>
__device__ __forceinline__ bool bounce(const RenderParams& P, const Scene& sc, Path& p, unsigned long long* stats)
{
    float T;
    int prim;
    bool inside;
    trace_any(sc, p.o, p.d, T, prim, inside);
    return shade(P, sc, p, T, prim, inside, stats) == kContinue;
}

// ------------------------------------------------------------------------------------------------------------
// Warp-cooperative fold for the tail of a frame.  When a warp has no more pixels to pull and only a few live paths, the
// per-lane fold (every primitive, serially, for one ray) makes the frame wait ~6 us per bounce of its longest path.
// Here groups of lanes split the primitives of each remaining ray, and the order-dependent fold
// (pt:231-255, accept on hit && t2>0 && t1<T) is rebuilt exactly from its closed form:
//   K  = the largest index whose primitive contains the origin (t1 < 0 < t2): it always overwrites what came before;
//   the winner is the smallest entry distance t1 among later, non-containing hits if that beats t2_K (strictly),
//   else K; ties go to the lower index; with no K it is the plain first-wins argmin of t1 (t1 < FLOAT_MAX).
__device__ __forceinline__ bool coop_test(const PackedScene& sc, int i, V3 o, V3 d, const RayInv& ri, float& t1, float& t2)
{
    if (i < sc.nS) {
        float b, disc;
        sphere_terms(sc.sphere(i), o, d, b, disc);
        if (disc < 0.0f) return false;
        const float sq = fsqrt(disc);
        t1 = -b - sq;
        t2 = -b + sq;
        return t1 <= t2;
    }
    slab_terms(sc.cmin(i - sc.nS), sc.cmax(i - sc.nS), o, ri, t1, t2);
    return t1 <= t2;
}
// All 32 lanes call this.  The n_live live rays of the warp (bits of `live`) are served by groups of g = 32 / next_pow2(n_live)
// lanes: group j works on the j-th live ray, lane `sub` of the group tests primitives sub, sub+g, ...; butterfly
// reductions inside the group rebuild the fold; the owner lane then picks its result up from its group.
__device__ __forceinline__ void trace_group(const PackedScene& sc, unsigned lane, unsigned live, unsigned n_live, V3 o_, V3 d_,
                                            float& T, int& prim, bool& inside)
{
    const unsigned log2g = 5u - (n_live <= 1u ? 0u : (32u - (unsigned)__clz(n_live - 1u)));   // g = 32 >> ceil(log2 n_live)
    const unsigned g = 1u << log2g;
    const unsigned grp = lane >> log2g, sub = lane & (g - 1u);
    const unsigned src_bit = __fns(live, 0, (int)grp + 1);                  // position of the grp-th live lane
    const int src = (grp < n_live) ? (int)src_bit : (int)(__ffs(live) - 1);   // spare groups shadow the first ray (result unused)
    const V3 o = mk(__shfl_sync(0xffffffffu, o_.x, src), __shfl_sync(0xffffffffu, o_.y, src), __shfl_sync(0xffffffffu, o_.z, src));
    const V3 d = mk(__shfl_sync(0xffffffffu, d_.x, src), __shfl_sync(0xffffffffu, d_.y, src), __shfl_sync(0xffffffffu, d_.z, src));
    const RayInv inv = ray_inverse(o, d);
    const int n = sc.nS + sc.nC;

    uint32_t key; int idx, k_idx; float t1, t2, k_t2;
    int after = -1;
    float limit = kFloatMax;                       // the first accept needs t1 < FLOAT_MAX (pt:228,234)
    float gT = kFloatMax; int gprim = -1; bool ginside = false;
    for (int pass = 0; pass < 2; ++pass) {
        // this lane's share: best non-containing hit (min t1, then min index) and the largest containing index
        key = 0xffffffffu; idx = 0x7fffffff; t1 = kFloatMax; t2 = 0.0f; k_idx = -1; k_t2 = 0.0f;
        for (int i = (int)sub; i < n; i += (int)g) {
            float a1, a2;
            if (i > after && coop_test(sc, i, o, d, inv, a1, a2) && a2 > 0.0f) {
                if (a1 < 0.0f) { k_idx = i; k_t2 = a2; }                     // ascending i: the last one is the largest
                else if (a1 < limit) {
                    const uint32_t kk = __float_as_uint(a1) & 0x7fffffffu;    // t1 >= 0 orders like its bits (-0 folded onto +0)
                    if (kk < key) { key = kk; idx = i; t1 = a1; t2 = a2; }
                }
            }
        }
        // butterfly all-reduce inside the group
        for (unsigned off = 1u; off < g; off <<= 1) {
            const uint32_t okey = __shfl_xor_sync(0xffffffffu, key, off);
            const int oidx = __shfl_xor_sync(0xffffffffu, idx, off);
            const float ot1 = __shfl_xor_sync(0xffffffffu, t1, off), ot2 = __shfl_xor_sync(0xffffffffu, t2, off);
            const int ok = __shfl_xor_sync(0xffffffffu, k_idx, off);
            const float okt2 = __shfl_xor_sync(0xffffffffu, k_t2, off);
            if (okey < key || (okey == key && oidx < idx)) { key = okey; idx = oidx; t1 = ot1; t2 = ot2; }
            if (ok > k_idx) { k_idx = ok; k_t2 = okt2; }
        }
        const bool found = key != 0xffffffffu;
        if (pass == 0) {
            if (k_idx < 0) { gT = found ? t1 : kFloatMax; gprim = found ? idx : -1; ginside = found && (t1 == t2); }
            // the origin sits inside primitive K: only later primitives can displace it, and only by beating its exit distance
            else { gT = k_t2; gprim = k_idx; ginside = true; after = k_idx; limit = __uint_as_float(0x7f800000u); }
            if (!__any_sync(0xffffffffu, k_idx >= 0)) break;
            if (k_idx < 0) after = n;              // this group is done; it only keeps the second pass convergent
        } else if (after < n && found && t1 < gT) {
            gT = t1; gprim = idx; ginside = (t1 == t2);
        }
    }
    // the owner of the j-th live ray reads its result from group j
    const unsigned rank = (unsigned)__popc(live & ((1u << lane) - 1u));
    const int from = (int)((rank << log2g) & 31u);
    T = __shfl_sync(0xffffffffu, gT, from);
    prim = __shfl_sync(0xffffffffu, gprim, from);
    inside = __shfl_sync(0xffffffffu, (int)ginside, from) != 0;
}

// ------------------------------------------------------------------------------------------------------------
// Large scenes: the same fold through a bounding-volume hierarchy in shared memory.  The hierarchy only ever REMOVES
// primitives that provably fail the exact test: every primitive's box is inflated by more than the worst-case rounding error
// of its own fp32 test (spheres: radius^2 + 4e-6 D^2; slabs: 1e-5 D; D bounds every coordinate and origin-centre distance —
// the derivation is in DESIGN.md), node tests carry a slack tau on the ray parameter, and whatever survives is tested with the
// exact formulas.  Because candidates arrive out of index order, the order-dependent fold is rebuilt from its closed form
// (see trace_group): pass 0 finds K, the largest index containing the origin, and the first-index argmin of the entry distance;
// if K exists a second pass looks only at later primitives that beat t2_K.
__device__ __forceinline__ void bvh_consider(const PackedScene& sc, int i, V3 o, V3 d, const RayInv& inv, int after, float limit,
                                             uint32_t& key, int& idx, float& t1b, float& t2b, int& k_idx, float& k_t2, float& best)
{
    float a1, a2;
    if (i > after && coop_test(sc, i, o, d, inv, a1, a2) && a2 > 0.0f) {
        if (a1 < 0.0f) { if (i > k_idx) { k_idx = i; k_t2 = a2; } }
        else if (a1 < limit) {
            const uint32_t kk = __float_as_uint(a1) & 0x7fffffffu;
            if (kk < key || (kk == key && i < idx)) { key = kk; idx = i; t1b = a1; t2b = a2; best = fmin_(best, a1); }
        }
    }
}
__device__ __forceinline__ bool bvh_box(const float4 lo, const float4 hi, V3 o, V3 inv, float tau, float best, float& tn)
{
    const float ax = (lo.x - o.x) * inv.x, ay = (lo.y - o.y) * inv.y, az = (lo.z - o.z) * inv.z;
    const float bx = (hi.x - o.x) * inv.x, by = (hi.y - o.y) * inv.y, bz = (hi.z - o.z) * inv.z;
    tn = fmax_(fmin_(ax, bx), fmax_(fmin_(ay, by), fmin_(az, bz)));
    const float tf = fmin_(fmax_(ax, bx), fmin_(fmax_(ay, by), fmax_(az, bz)));
    return tn <= tf && tf >= -tau && tn <= best + tau;
}
__device__ __forceinline__ void bvh_pass(const PackedScene& sc, V3 o, V3 d, const RayInv& inv, int after, float limit, float best,
                                         uint32_t& key, int& idx, float& t1b, float& t2b, int& k_idx, float& k_t2, int* visits = nullptr)
{
    key = 0xffffffffu; idx = 0x7fffffff; t1b = kFloatMax; t2b = 0.0f; k_idx = -1; k_t2 = 0.0f;
    for (int u = 0; u < sc.n_unbounded; ++u) bvh_consider(sc, sc.pidx[u], o, d, inv, after, limit, key, idx, t1b, t2b, k_idx, k_t2, best);
    if (sc.n_nodes == 0) return;
    int stack[32];
    int sp = 0, node = 0;
    while (true) {
        const float4 A = sc.nodes[2 * node], B = sc.nodes[2 * node + 1];
        const int count = __float_as_int(B.w), first = __float_as_int(A.w);
        if (visits) ++*visits;
        if (count > 0) {
            for (int j = 0; j < count; ++j) bvh_consider(sc, sc.pidx[first + j], o, d, inv, after, limit, key, idx, t1b, t2b, k_idx, k_t2, best);
        } else {
            float tl, tr;
            const bool hl = bvh_box(sc.nodes[2 * first], sc.nodes[2 * first + 1], o, inv.inv, sc.tau, best, tl);
            const bool hr = bvh_box(sc.nodes[2 * first + 2], sc.nodes[2 * first + 3], o, inv.inv, sc.tau, best, tr);
            if (hl && hr) {
                const bool left_first = tl <= tr;
                if (sp < 32) stack[sp++] = left_first ? first + 1 : first;
                node = left_first ? first : first + 1;
                continue;
            }
            if (hl || hr) { node = hl ? first : first + 1; continue; }
        }
        if (sp == 0) break;
        node = stack[--sp];
    }
}
__device__ __forceinline__ void trace_bvh(const PackedScene& sc, V3 o, V3 d, float& T, int& prim, bool& inside, int* visits = nullptr)
{
    // non-finite rays (normalize(0) after total internal reflection, ...) take the plain fold: nothing to cull by
    const float fin = fabsf(o.x) + fabsf(o.y) + fabsf(o.z) + fabsf(d.x) + fabsf(d.y) + fabsf(d.z);
    if (!(fin <= kFloatMax)) { trace(sc, o, d, T, prim, inside); return; }
    const RayInv inv = ray_inverse(o, d);
    uint32_t key; int idx, k_idx; float t1, t2, k_t2;
    bvh_pass(sc, o, d, inv, -1, kFloatMax, kFloatMax, key, idx, t1, t2, k_idx, k_t2, visits);
    if (k_idx < 0) {
        const bool found = key != 0xffffffffu;
        T = found ? t1 : kFloatMax; prim = found ? idx : -1; inside = found && (t1 == t2);
        return;
    }
    const int K = k_idx;
    const float t2K = k_t2;
    int k2; float k2t;
    bvh_pass(sc, o, d, inv, K, __uint_as_float(0x7f800000u), t2K, key, idx, t1, t2, k2, k2t, visits);
    if (key != 0xffffffffu && t1 < t2K) { T = t1; prim = idx; inside = (t1 == t2); }
    else { T = t2K; prim = K; inside = true; }
}

// ------------------------------------------------------------------------------------------------------------
// Large scenes, second variant (megakernel<kFold = 3>, the default above the BVH threshold when the grid fits): a uniform grid
// in shared memory walked by a 3-D DDA.  A cell lists every primitive whose box — inflated by the error margins of the BVH
// plus a grid margin that covers the DDA's own rounding — overlaps it; scene-sized and non-finite primitives sit in the
// always-tested list.  A lane visits the cells its ray crosses in order of entry distance, runs the exact tests on their
// primitives, and stops once the nearest hit so far lies (by more than tau) before the exit of the current cell: every cell
// that could hold a nearer hit has been visited by then.  Candidates arrive out of index order and more than once, so the
// fold is rebuilt from its closed form exactly as in trace_bvh (a primitive tested twice changes nothing: the accumulators are
// a max over containing indices and a first-index argmin).  Compared with the binary BVH there is no stack, no box test per
// node, and all lanes of a warp run the same short loop body: C3 spends ~x.x k instead of 11.8 k lane-instructions per sample.
struct GridView {
    const unsigned short* cell_start;      // n_cells + 1 offsets into items
    const unsigned char* cell_spheres;     // per cell: how many of its items are spheres (items ascend, so they come first)
    const unsigned short* items;
    int nx, ny, nz;
    float lo[3], inv_cell[3], cell[3], hi[3];
};
__device__ __forceinline__ GridView make_grid_view(const RenderParams& P, const float4* smem_block)
{
    GridView G;
    G.cell_start = reinterpret_cast<const unsigned short*>(smem_block + P.off_gcell);
    G.cell_spheres = reinterpret_cast<const unsigned char*>(smem_block + P.off_gsph);
    G.items = reinterpret_cast<const unsigned short*>(smem_block + P.off_gitem);
    G.nx = P.grid_n[0]; G.ny = P.grid_n[1]; G.nz = P.grid_n[2];
    for (int k = 0; k < 3; ++k) { G.lo[k] = P.grid_lo[k]; G.hi[k] = P.grid_hi[k]; G.cell[k] = P.grid_cell[k]; G.inv_cell[k] = P.grid_inv[k]; }
    return G;
}
// closed-form accumulation of one candidate's (t1, t2): K = the largest containing index, else the first-index argmin of t1
__device__ __forceinline__ void grid_accumulate(int i, float a1, float a2, float limit, uint32_t& key, int& idx, float& t1b, float& t2b, int& k_idx,
                                                float& k_t2, float& best)
{
    if (a1 < 0.0f) { if (i > k_idx) { k_idx = i; k_t2 = a2; } }
    else if (a1 < limit) {
        const uint32_t kk = __float_as_uint(a1) & 0x7fffffffu;
        if (kk < key || (kk == key && i < idx)) { key = kk; idx = i; t1b = a1; t2b = a2; best = fmin_(best, a1); }
    }
}
__device__ __forceinline__ void trace_grid(const PackedScene& sc, const GridView& G, V3 o, V3 d, float& T, int& prim, bool& inside, int* visits = nullptr)
{
    // non-finite rays take the plain fold, as in trace_bvh
    const float fin = fabsf(o.x) + fabsf(o.y) + fabsf(o.z) + fabsf(d.x) + fabsf(d.y) + fabsf(d.z);
    if (!(fin <= kFloatMax)) { trace(sc, o, d, T, prim, inside); return; }
    const RayInv ri = ray_inverse(o, d);
    const V3 inv = ri.inv;
    // the part of the ray inside the grid: [tn, tf] (NaN-dropping min/max: an axis-parallel ray inside the slab gives -inf / +inf)
    const float gax = (G.lo[0] - o.x) * inv.x, gbx = (G.hi[0] - o.x) * inv.x;
    const float gay = (G.lo[1] - o.y) * inv.y, gby = (G.hi[1] - o.y) * inv.y;
    const float gaz = (G.lo[2] - o.z) * inv.z, gbz = (G.hi[2] - o.z) * inv.z;
    const float tn = fmax_(0.0f, fmax_(fmin_(gax, gbx), fmax_(fmin_(gay, gby), fmin_(gaz, gbz))));
    const float tf = fmin_(fmax_(gax, gbx), fmin_(fmax_(gay, gby), fmax_(gaz, gbz)));
    const bool enters = tn <= tf;
    const V3 p = mk(__fmaf_rn(d.x, tn, o.x), __fmaf_rn(d.y, tn, o.y), __fmaf_rn(d.z, tn, o.z));
    const int ix0 = min(G.nx - 1, max(0, __float2int_rd((p.x - G.lo[0]) * G.inv_cell[0])));
    const int iy0 = min(G.ny - 1, max(0, __float2int_rd((p.y - G.lo[1]) * G.inv_cell[1])));
    const int iz0 = min(G.nz - 1, max(0, __float2int_rd((p.z - G.lo[2]) * G.inv_cell[2])));
    const int sx = d.x > 0.0f ? 1 : -1, sy = d.y > 0.0f ? 1 : -1, sz = d.z > 0.0f ? 1 : -1;
    const float inf = __uint_as_float(0x7f800000u);
    // parameter at which the ray leaves its first cell along each axis; a zero component never leaves
    const float tx0 = d.x != 0.0f ? (G.lo[0] + (float)(ix0 + (sx > 0 ? 1 : 0)) * G.cell[0] - o.x) * inv.x : inf;
    const float ty0 = d.y != 0.0f ? (G.lo[1] + (float)(iy0 + (sy > 0 ? 1 : 0)) * G.cell[1] - o.y) * inv.y : inf;
    const float tz0 = d.z != 0.0f ? (G.lo[2] + (float)(iz0 + (sz > 0 ? 1 : 0)) * G.cell[2] - o.z) * inv.z : inf;
    const float dx = G.cell[0] * fabsf(inv.x), dy = G.cell[1] * fabsf(inv.y), dz = G.cell[2] * fabsf(inv.z);

    // pass 0 finds K and the first-index argmin of the entry distance; pass 1 (only when K exists) looks at later primitives that
    // beat t2_K.  One loop body for both passes: the walk is the same, only (after, limit, best) differ — and the code is half the
    // size, which matters: the first version of this kernel spent more issue slots waiting for instructions than for data.
    int after = -1, K = -1;
    float limit = kFloatMax, best = kFloatMax, t2K = 0.0f;
    uint32_t key; int idx, k_idx; float t1b, t2b, k_t2;
#pragma unroll 1
    for (int pass = 0; pass < 2; ++pass) {
        key = 0xffffffffu; idx = 0x7fffffff; t1b = kFloatMax; t2b = 0.0f; k_idx = -1; k_t2 = 0.0f;
        for (int u = 0; u < sc.n_unbounded; ++u) bvh_consider(sc, sc.pidx[u], o, d, ri, after, limit, key, idx, t1b, t2b, k_idx, k_t2, best);
        if (enters) {
            int ix = ix0, iy = iy0, iz = iz0;
            float tx = tx0, ty = ty0, tz = tz0;
            for (int guard = G.nx + G.ny + G.nz + 3; guard > 0; --guard) {
                const int cell = (iz * G.ny + iy) * G.nx + ix;
                const int first = G.cell_start[cell], last = G.cell_start[cell + 1], mid = first + G.cell_spheres[cell];
                if (visits) ++*visits;
                for (int j = first; j < mid; ++j) {                 // the cell's spheres
                    const int i = G.items[j];
                    float b, disc;
                    sphere_terms(sc.base[i], o, d, b, disc);
                    if (i > after && !(disc < 0.0f)) {
                        const float sq = fsqrt(disc);
                        const float a1 = -b - sq, a2 = -b + sq;
                        if (a1 <= a2 && a2 > 0.0f) grid_accumulate(i, a1, a2, limit, key, idx, t1b, t2b, k_idx, k_t2, best);
                    }
                }
                for (int j = mid; j < last; ++j) {                  // then its cuboids
                    const int i = G.items[j];
                    float a1, a2;
                    slab_terms(sc.cmin(i - sc.nS), sc.cmax(i - sc.nS), o, ri, a1, a2);
                    if (i > after && a1 <= a2 && a2 > 0.0f) grid_accumulate(i, a1, a2, limit, key, idx, t1b, t2b, k_idx, k_t2, best);
                }
                const float t_exit = fmin_(tx, fmin_(ty, tz));
                if (best + sc.tau < t_exit) break;                  // nothing nearer can start in a cell that begins at or after t_exit
                if (tx <= ty && tx <= tz) { ix += sx; if ((unsigned)ix >= (unsigned)G.nx) break; tx += dx; }
                else if (ty <= tz) { iy += sy; if ((unsigned)iy >= (unsigned)G.ny) break; ty += dy; }
                else { iz += sz; if ((unsigned)iz >= (unsigned)G.nz) break; tz += dz; }
            }
        }
        if (pass == 0) {
            if (k_idx < 0) {
                const bool found = key != 0xffffffffu;
                T = found ? t1b : kFloatMax; prim = found ? idx : -1; inside = found && (t1b == t2b);
                return;
            }
            K = k_idx; t2K = k_t2;
            after = K; limit = inf; best = t2K;
        }
    }
    if (key != 0xffffffffu && t1b < t2K) { T = t1b; prim = idx; inside = (t1b == t2b); }
    else { T = t2K; prim = K; inside = true; }
}

// ------------------------------------------------------------------------------------------------------------
// Small scenes (<= 64 primitives): ray classification (Arvo & Kirk's 5-D idea, as a flat table).  Rays are classed by the
// grid cell of their origin and by a direction bucket (cube face x G x G); the table holds, per class, the 64-bit set of
// primitives that ANY ray of the class can hit (rct_build_kernel: an interval test of the class's beam against every
// primitive's box, inflated by more than the rounding error of the exact fp32 tests — the same margins as the BVH).  A lane
// then runs the exact tests of pt:231-255 over the set bits of its own mask only, in ascending index order.  A primitive
// outside the set fails `hit && t2 > 0`, and a failing primitive never changes the fold's state, so the order-dependent fold
// over the subset equals the fold over all primitives bit for bit.  Rays that start outside the grid, non-finite rays and
// zero directions take the full mask (= the plain fold).  The default scene tests ~6 primitives per ray instead of 55; the
// loop runs for the lane with the most candidates of its warp (~12 spheres + ~4 cuboids).
__device__ __forceinline__ unsigned long long rct_lookup(const RenderParams& P, V3 o, V3 d, V3 inv)
{
    // every comparison below is false for a NaN operand, so non-finite origins and directions fall through to the full mask
    const float fx = (o.x - P.rct_lo[0]) * P.rct_inv[0], fy = (o.y - P.rct_lo[1]) * P.rct_inv[1], fz = (o.z - P.rct_lo[2]) * P.rct_inv[2];
    const float ax = fabsf(d.x), ay = fabsf(d.y), az = fabsf(d.z);
    const bool mx = ax >= ay && ax >= az, my = !mx && ay >= az;
    const float dm = mx ? d.x : (my ? d.y : d.z);
    const float im = fabsf(mx ? inv.x : (my ? inv.y : inv.z));
    const float u = (mx ? d.y : d.x) * im, v = ((mx || my) ? d.z : d.y) * im;
    // 0.5 <= im <= 2: the major component of a unit vector lies in [0.577, 1]; anything else (zero, huge, inf) is not classified
    const bool ok = fx >= 0.0f && fx < P.rct_nf[0] && fy >= 0.0f && fy < P.rct_nf[1] && fz >= 0.0f && fz < P.rct_nf[2] &&
                    im >= 0.5f && im <= 2.0f && fabsf(u) <= 1.0001f && fabsf(v) <= 1.0001f;
    if (!ok) return P.rct_valid;
    const unsigned G = (unsigned)P.rct_G, Gm = G - 1u;
    const unsigned gu = min(Gm, (unsigned)max(0, __float2int_rd(__fmaf_rn(u, P.rct_halfG, P.rct_halfG))));
    const unsigned gv = min(Gm, (unsigned)max(0, __float2int_rd(__fmaf_rn(v, P.rct_halfG, P.rct_halfG))));
    const unsigned f = (mx ? 0u : (my ? 2u : 4u)) + (dm < 0.0f ? 1u : 0u);
    const unsigned cell = ((unsigned)__float2int_rd(fz) * (unsigned)P.rct_n[1] + (unsigned)__float2int_rd(fy)) * (unsigned)P.rct_n[0] + (unsigned)__float2int_rd(fx);
    const unsigned entry = ((cell * 6u + f) * G + gv) * G + gu;          // < 2^23: the table is capped at 64 MiB
    return __ldg(P.rct + entry);
}
__device__ __forceinline__ void trace_rct(const RenderParams& P, const PackedScene& sc, V3 o, V3 d, float& T, int& prim, bool& inside)
{
    T = kFloatMax;
    prim = -1;
    float t2w = 0.0f;
    const RayInv ri = ray_inverse(o, d);
    const unsigned long long mask = rct_lookup(P, o, d, ri.inv);
    const unsigned w0 = (unsigned)mask, w1 = (unsigned)(mask >> 32);
#pragma unroll 1
    for (int w = 0; w < 2; ++w) {                       // spheres 0..31, then 32..63
        unsigned bits = w ? (w1 & P.rct_sm1) : (w0 & P.rct_sm0);
        const float4* base = sc.base + 32 * w;
        while (bits) {
            const int j = __ffs(bits) - 1;
            bits &= bits - 1u;
            float b, disc;
            sphere_terms(base[j], o, d, b, disc);
            accept_sphere(b, disc, 32 * w + j, T, prim, t2w);
        }
    }
#pragma unroll 1
    for (int w = 0; w < 2; ++w) {                       // cuboids, bits nS..nS+nC-1
        unsigned bits = w ? (w1 & ~P.rct_sm1) : (w0 & ~P.rct_sm0);
        const float4* cub = sc.base + sc.off_cmin + 2 * (32 * w - sc.nS);
        while (bits) {
            const int j = __ffs(bits) - 1;
            bits &= bits - 1u;
            float t1, t2;
            slab_terms(cub[2 * j], cub[2 * j + 1], o, ri, t1, t2);
            accept(t1, t2, 32 * w + j, T, prim, t2w);
        }
    }
    inside = (T == t2w);
}

#ifndef PTB_MEGA_ONLY      // (the fast-arithmetic translation unit compiles the megakernel only)
// Table build: one thread per (cell, face, gv, gu).  The class's rays are { o + s * (sgn e_m + u e_a + v e_b) : o in the
// cell, u in [u0,u1], v in [v0,v1], s >= 0 } (s = distance along the major axis m).  Such a ray meets the box [klo,khi] iff
// some s >= 0 satisfies three interval conditions — on the major axis directly, on each minor axis through
// s*u in [klo_a - chi_a, khi_a - clo_a], i.e. s*u1 >= A0 and s*u0 <= A1, both half-lines in s — so the test is an
// intersection of half-lines.  Cells and buckets are widened by eps (classification rounds in fp32), boxes by the margins.
struct RctBuild {
    const unsigned char* ubo;
    unsigned long long* table;
    double lo[3], cell[3];
    double eps_cell, eps_u, E, m;
    int n[3], G, nS, nC, max_spheres;
    unsigned long long total;
};
__global__ void rct_build_kernel(const __grid_constant__ RctBuild B)
{
    const unsigned long long id = (unsigned long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (id >= B.total) return;
    const int G = B.G;
    const int gu = (int)(id % G), gv = (int)((id / G) % G), f = (int)((id / ((unsigned long long)G * G)) % 6);
    const unsigned long long cell = id / ((unsigned long long)G * G * 6);
    const int c[3] = {(int)(cell % B.n[0]), (int)((cell / B.n[0]) % B.n[1]), (int)(cell / ((unsigned long long)B.n[0] * B.n[1]))};
    double clo[3], chi[3];
    for (int k = 0; k < 3; ++k) { clo[k] = B.lo[k] + c[k] * B.cell[k] - B.eps_cell; chi[k] = B.lo[k] + (c[k] + 1) * B.cell[k] + B.eps_cell; }
    const int m = f >> 1, a = m == 0 ? 1 : 0, b = m == 2 ? 1 : 2;
    const bool neg = f & 1;
    const double ulo[2] = {-1.0 + 2.0 * gu / G - B.eps_u, -1.0 + 2.0 * gv / G - B.eps_u};
    const double uhi[2] = {-1.0 + 2.0 * (gu + 1) / G + B.eps_u, -1.0 + 2.0 * (gv + 1) / G + B.eps_u};
    const int ax2[2] = {a, b};
    unsigned long long mask = 0ull;
    for (int p = 0; p < B.nS + B.nC; ++p) {
        double klo[3], khi[3];
        if (p < B.nS) {
            const float* g = reinterpret_cast<const float*>(B.ubo + (size_t)p * kSphereStride);
            const double r = sqrt((double)g[3] * (double)g[3] + B.E) + B.m;
            for (int k = 0; k < 3; ++k) { klo[k] = (double)g[k] - r; khi[k] = (double)g[k] + r; }
        } else {
            const float* g = reinterpret_cast<const float*>(B.ubo + (size_t)B.max_spheres * kSphereStride + (size_t)(p - B.nS) * kCuboidStride);
            for (int k = 0; k < 3; ++k) { klo[k] = fmin((double)g[k], (double)g[4 + k]) - B.m; khi[k] = fmax((double)g[k], (double)g[4 + k]) + B.m; }
        }
        const double sum = klo[0] + klo[1] + klo[2] + khi[0] + khi[1] + khi[2];
        bool cand = !(fabs(sum) <= 1e300);                      // NaN / Inf geometry: never culled
        if (!cand) {
            // major axis: sgn * s in [klo_m - chi_m, khi_m - clo_m]
            const double L = klo[m] - chi[m], U = khi[m] - clo[m];
            double s0 = neg ? -U : L, s1 = neg ? -L : U;
            if (s0 < 0.0) s0 = 0.0;
            for (int q = 0; q < 2 && s0 <= s1; ++q) {
                const double A0 = klo[ax2[q]] - chi[ax2[q]], A1 = khi[ax2[q]] - clo[ax2[q]];
                const double u0 = ulo[q], u1 = uhi[q];
                if (u1 > 0.0) s0 = fmax(s0, A0 / u1); else if (u1 < 0.0) s1 = fmin(s1, A0 / u1); else if (A0 > 0.0) s1 = -1.0;      // s*u1 >= A0
                if (u0 > 0.0) s1 = fmin(s1, A1 / u0); else if (u0 < 0.0) s0 = fmax(s0, A1 / u0); else if (A1 < 0.0) s1 = -1.0;      // s*u0 <= A1
            }
            cand = s0 <= s1;
        }
        if (cand) mask |= 1ull << p;
    }
    B.table[id] = mask;
}

#endif  // PTB_MEGA_ONLY

// pt:125-129 — mean over SPP, running mean over frames, store.  Frame 0 does not read the image: the reference
// multiplies the stale value by exactly 0 there (mix(x, y, 1.0)), so a zero stands in for it.
__device__ __forceinline__ void finish_pixel(const RenderParams& P, const Path& p, size_t frame_offset = 0)
{
    const V3 irr = p.irr * P.inv_spp;
    const size_t at = (size_t)p.lrow * P.width + p.px;
    if (P.scratch) {              // pipelined frames: the running mean is applied by blend_kernel, in frame order
        P.scratch[frame_offset + at] = make_float4(irr.x, irr.y, irr.z, 1.0f);
        return;
    }
    float4* px = P.image + at;
    V3 last = mk(0.0f, 0.0f, 0.0f);
    if (P.frame > 0) {
        const float4 l = *px;
        last = mk(l.x, l.y, l.z);
    }
    const V3 out = mix(last, irr, P.blend);
    *px = make_float4(out.x, out.y, out.z, 1.0f);
}

#ifndef PTB_MEGA_ONLY
// pt:126-129 for pipelined frames: image = mix(image, this frame's estimate, 1/(frame+1)) — the same operations, in the same
// order, as finish_pixel's in-place path; one thread per pixel, 128-bit loads and stores.
__global__ void blend_kernel(float4* __restrict__ image, const float4* __restrict__ estimate, size_t n, int frame, float blend)
{
    const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const float4 e = estimate[i];
    V3 last = mk(0.0f, 0.0f, 0.0f);
    if (frame > 0) {
        const float4 l = image[i];
        last = mk(l.x, l.y, l.z);
    }
    const V3 out = mix(last, mk(e.x, e.y, e.z), blend);
    image[i] = make_float4(out.x, out.y, out.z, 1.0f);
}

// The same for a batch of consecutive frames traced by one launch (ptb_set_batch): the running mean is folded frame by frame in
// registers — mix(mix(mix(prev, e0, b0), e1, b1), ...), exactly the operations of `frames` blend_kernel launches in the same
// order — with one read and one write of the accumulation image instead of `frames` of each.
constexpr int kMaxBatch = 16;
struct BatchBlend {
    float blend[kMaxBatch];        // 1 / (frame + 1) of each frame, host-rounded like RenderParams::blend
    int frame0, frames;
    unsigned long long stride;     // float4 elements between consecutive frames' estimates
};
// Device-side dependency of a batch's blend kernel on the batch's trace.  A persistent grid leaves no CTA slot (and no
// register) free while it runs, so a kernel that waited for the trace through a stream event could only start one batch late,
// when the NEXT grid drains.  Instead the blend kernel is launched right behind the trace without a stream dependency: its few
// CTAs (high-priority stream) take the first slots the previous grid frees, sit beside the trace, and start the moment the
// trace's last CTA publishes `value` in `flag`.  The grid is small (<= 64 CTAs), so the trace always has CTA slots to finish.
struct BatchWait {
    const unsigned* flag;          // nullptr: nothing to wait for (the caller ordered the kernel by a stream event)
    unsigned value;
    unsigned* error;               // sticky: a wait that timed out
};
__device__ __forceinline__ void wait_for_trace(const BatchWait& Wt)
{
    if (Wt.flag && threadIdx.x == 0) {
        const unsigned long long t0 = global_ns();
        while ((int)(ld_acquire_gpu(Wt.flag) - Wt.value) < 0) {
            __nanosleep(500);
            // ten minutes: far beyond any batch (the host only launches this kernel behind a trace that did launch); a trace that
            // never finishes must not wedge the GPU for ever, so the wait gives up and raises the sticky flag ptb_synchronize reports
            if (global_ns() - t0 > 600000000000ull) { atomicExch(Wt.error, 1u); break; }
        }
    }
    __syncthreads();
}
// Eight of a pixel's estimates are requested before the first one is used (eight independent 128-bit loads in flight per
// thread, <= 64 registers), so a few dozen CTAs already move a batch at several hundred GB/s: a background job beside the next
// batch's persistent grid, not a whole-machine pass.
constexpr int kBlendChunk = 8;
__global__ void __launch_bounds__(256, 4) blend_batch_kernel(float4* __restrict__ image, const float4* __restrict__ estimates, size_t n,
                                                             const __grid_constant__ BatchBlend B, const __grid_constant__ BatchWait Wt)
{
    wait_for_trace(Wt);
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
        V3 acc = mk(0.0f, 0.0f, 0.0f);
        if (B.frame0 > 0) {
            const float4 l = image[i];
            acc = mk(l.x, l.y, l.z);
        }
        for (int j0 = 0; j0 < B.frames; j0 += kBlendChunk) {
            float4 e[kBlendChunk];
#pragma unroll
            for (int j = 0; j < kBlendChunk; ++j)
                if (j0 + j < B.frames) e[j] = estimates[(size_t)(j0 + j) * B.stride + i];
#pragma unroll
            for (int j = 0; j < kBlendChunk; ++j)
                if (j0 + j < B.frames) acc = mix(acc, mk(e[j].x, e[j].y, e[j].z), B.blend[j0 + j]);
        }
        image[i] = make_float4(acc.x, acc.y, acc.z, 1.0f);
    }
}
#endif  // PTB_MEGA_ONLY

// local row -> global y for the stripe partition
__device__ __forceinline__ int global_row(const RenderParams& P, int lrow)
{
    if (P.world == 1) return lrow;
    const int ls = lrow / P.stripe_rows;
    return (ls * P.world + P.rank) * P.stripe_rows + (lrow - ls * P.stripe_rows);
}

// ------------------------------------------------------------------------------------------------------------
// TMA bulk copy helpers (cp.async.bulk global -> shared, completion on an mbarrier).
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count)
{
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes)
{
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void tma_bulk_g2s(void* dst, const void* src, uint32_t bytes, uint64_t* bar)
{
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smem_u32(dst)),
                 "l"(src), "r"(bytes), "r"(smem_u32(bar))
                 : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity)
{
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "PTB_WAIT_%=:\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
        "@p bra PTB_DONE_%=;\n"
        "bra PTB_WAIT_%=;\n"
        "PTB_DONE_%=:\n"
        "}\n" ::"r"(smem_u32(bar)),
        "r"(parity)
        : "memory");
}

// ------------------------------------------------------------------------------------------------------------
// The megakernel.  Work item = one pixel (all its SPP samples, RNG stream intact).  Work index -> pixel through
// 8x4 tiles so a freshly filled warp starts on a compact footprint.
// kRing: stage primary rays through the per-warp shared-memory ring (2 KB per warp; the host turns it off when the scene
// block is so large that the ring would lower the number of resident CTAs).
// kBatch: one launch traces P.batch consecutive frames (ptb_set_batch).  The work counter runs over batch * tiles_total items,
// frame-major; a tile's frame slot b seeds its pixels with frame P.frame + b and routes their estimates to scratch image b,
// so lanes move from the last pixels of one frame straight into the next frame and only the last frame of a batch drains.
// The slot travels in bits 12..15 of the ring's pixel word (the host batches only images up to 4096 pixels wide).
// kFold: 0 = brute-force fold, 1 = shared-memory BVH, 3 = shared-memory grid (large scenes), 2 = ray-classification table (<= 64 primitives).
// kRing: 0 = idle lanes generate their primary ray in place; 1 = primary rays staged through the per-warp ring; 2 = ring + the
// completion queue (SPP == 1 only: the host selects it, P.defer_finish documents it).
template <bool kStats, int kRingMode, int kFold, bool kBatch = false>
__global__ void __launch_bounds__(kMegaThreads, PTB_MIN_BLOCKS) megakernel(const __grid_constant__ RenderParams P)
{
    extern __shared__ __align__(128) unsigned char smem_raw[];
    __shared__ __align__(8) uint64_t bar;
    constexpr bool kRing = kRingMode != 0;
    constexpr bool kDefer = kRingMode == 2;
    __shared__ uint32_t s_ring[kRing ? (kMegaThreads / 32) * 8 * kQueue : 1];
    // Completion queue (per warp, kDoneWords x 64 entries; SPP == 1 only).  A sample that ends — on a surface, or by leaving the
    // scene with the environment lookup still to do — parks (direction, throughput, radiance, pixel) here and frees its lane at
    // once; when 32 entries have gathered, the whole warp shades them together: the cubemap lookup and the pixel store then run
    // with 32 lanes instead of the ~9 that happen to miss in one iteration of the bounce loop.
    __shared__ uint32_t s_done[kDefer ? (kMegaThreads / 32) * kDoneWords * kQueue : 1];
    float4* sblock = reinterpret_cast<float4*>(smem_raw);

    if (P.ktime && threadIdx.x == 0) atomicMin(P.ktime, global_ns());      // launch timing: when the first CTA starts
    // ---- stage the packed scene: one elected thread arms the barrier and issues the bulk copies
    if (threadIdx.x == 0) mbar_init(&bar, 1);
    __syncthreads();
    if (threadIdx.x == 0) {
        mbar_expect_tx(&bar, (uint32_t)P.stage_bytes);
        const unsigned char* src = reinterpret_cast<const unsigned char*>(P.scene);
        for (int off = 0; off < P.stage_bytes; off += 32768) {
            const int n = min(32768, P.stage_bytes - off);
            tma_bulk_g2s(smem_raw + off, src + off, (uint32_t)n, &bar);
        }
    }
    mbar_wait(&bar, 0);

    const PackedScene sc = PTB_PACKED_SCENE(P, sblock);
    GridView grid;
    if constexpr (kFold == 3) grid = make_grid_view(P, sblock);

    const unsigned lane = threadIdx.x & 31u;
    const unsigned lt_mask = (1u << lane) - 1u;
    unsigned long long* stats = kStats ? P.stats : nullptr;

    // Ring of ready-made primary rays.  Ray generation (4 RNG draws, 2 normalisations, sin/cos, 8 matrix rows) runs with
    // all 32 lanes converged on one 8x4 pixel tile and is amortised over 32 pixels, instead of running for the ~12 lanes
    // that happen to be idle in every iteration of the bounce loop.
    uint32_t* ring = s_ring + (kRing ? (threadIdx.x >> 5) * 8 * kQueue : 0);   // [8][kQueue]: o.xyz d.xyz rng px|lrow<<16
    unsigned q_head = 0, q_count = 0;                                           // warp-uniform
    uint32_t* doneq = s_done + (kDefer ? (threadIdx.x >> 5) * kDoneWords * kQueue : 0);
    unsigned dq_count = 0;                                                      // warp-uniform
    constexpr bool defer = kDefer;
    // shades `n` queued completions, entries [first, first + n), one per lane
    auto flush_done = [&](unsigned first, unsigned n) {
        if (lane < n) {
            const unsigned e = first + lane;
            Path f;
            f.d = mk(__uint_as_float(doneq[0 * kQueue + e]), __uint_as_float(doneq[1 * kQueue + e]), __uint_as_float(doneq[2 * kQueue + e]));
            f.thr = mk(__uint_as_float(doneq[3 * kQueue + e]), __uint_as_float(doneq[4 * kQueue + e]), __uint_as_float(doneq[5 * kQueue + e]));
            f.rad = mk(__uint_as_float(doneq[6 * kQueue + e]), __uint_as_float(doneq[7 * kQueue + e]), __uint_as_float(doneq[8 * kQueue + e]));
            const uint32_t xy = doneq[9 * kQueue + e], fw = doneq[10 * kQueue + e];
            f.px = (int)(xy & 0xffffu); f.lrow = (int)(xy >> 16); f.fb = (int)(fw & 0xffu);
            if (fw >> 8) f.rad = add_environment(f.rad, env_lookup(P.env, P.env_size, f.d), f.thr);      // pt:177
            f.irr = mk(0.0f, 0.0f, 0.0f) + f.rad;                                            // pt:109,123 with SPP == 1
            if constexpr (kBatch) finish_pixel(P, f, (size_t)f.fb * (size_t)P.scratch_stride);
            else finish_pixel(P, f);
        }
        __syncwarp();
    };

    Path p;
    bool alive = false;         // lane owns an unfinished pixel
    bool fresh = false;         // lane needs a primary ray for the NEXT sample of its pixel (SPP > 1 only)
    bool exhausted = false;     // warp-uniform: the tile counter ran past the end

    while (true) {
        const unsigned dead = __ballot_sync(0xffffffffu, !alive);
        const unsigned n_dead = (unsigned)__popc(dead);
        if constexpr (kRing) {
            // ---- top up the ring: one whole tile per visit, all lanes generating
            if (n_dead > q_count && !exhausted && q_count <= kQueue - 32 && (n_dead >= (unsigned)PTB_REFILL_MIN || dead == 0xffffffffu)) {
                unsigned tile = 0;
                if (lane == 0) tile = atomicAdd(P.counters, 1u);
                tile = __shfl_sync(0xffffffffu, tile, 0);
                if (tile >= (kBatch ? P.tiles_total * (unsigned)P.batch : P.tiles_total)) {
                    exhausted = true;
                } else {
                    unsigned fb = 0u;
                    if constexpr (kBatch) { fb = tile / P.tiles_total; tile -= fb * P.tiles_total; }
                    const unsigned tyu = P.tiles_magic ? __umulhi(tile, P.tiles_magic) : tile / P.tiles_x;
                    const int x = (int)(tile - tyu * P.tiles_x) * 8 + (int)(lane & 7u), lr = (int)tyu * 4 + (int)(lane >> 3);
                    const int y = lr < P.local_rows ? global_row(P, lr) : P.height;
                    const bool valid = x < P.width && y < P.height;
                    const unsigned vmask = __ballot_sync(0xffffffffu, valid);
                    if (valid) {
                        Path g;
                        g.px = x; g.py = y;
                        g.rng = ((uint32_t)x * 1973u + (uint32_t)y * 9277u + ((uint32_t)P.frame + fb) * 2699u) | 1u;   // pt:106
                        primary_ray(P, g);
                        const unsigned slot = (q_head + q_count + (unsigned)__popc(vmask & lt_mask)) & (kQueue - 1);
                        ring[0 * kQueue + slot] = __float_as_uint(g.o.x); ring[1 * kQueue + slot] = __float_as_uint(g.o.y);
                        ring[2 * kQueue + slot] = __float_as_uint(g.o.z); ring[3 * kQueue + slot] = __float_as_uint(g.d.x);
                        ring[4 * kQueue + slot] = __float_as_uint(g.d.y); ring[5 * kQueue + slot] = __float_as_uint(g.d.z);
                        ring[6 * kQueue + slot] = g.rng; ring[7 * kQueue + slot] = (uint32_t)x | (fb << 12) | ((uint32_t)lr << 16);
                    }
                    q_count += (unsigned)__popc(vmask);
                    __syncwarp();
                }
            }
            // ---- idle lanes pop ready rays (prefix-popcount slot assignment)
            if (n_dead != 0u && q_count != 0u) {
                const unsigned take = min(n_dead, q_count);
                const unsigned rank = (unsigned)__popc(dead & lt_mask);
                if (!alive && rank < take) {
                    const unsigned slot = (q_head + rank) & (kQueue - 1);
                    p.o = mk(__uint_as_float(ring[0 * kQueue + slot]), __uint_as_float(ring[1 * kQueue + slot]), __uint_as_float(ring[2 * kQueue + slot]));
                    p.d = mk(__uint_as_float(ring[3 * kQueue + slot]), __uint_as_float(ring[4 * kQueue + slot]), __uint_as_float(ring[5 * kQueue + slot]));
                    p.rng = ring[6 * kQueue + slot];
                    const uint32_t xy = ring[7 * kQueue + slot];
                    if constexpr (kBatch) { p.px = (int)(xy & 0xfffu); p.fb = (int)((xy >> 12) & 0xfu); }
                    else p.px = (int)(xy & 0xffffu);
                    p.lrow = (int)(xy >> 16);
                    p.thr = mk(1.0f, 1.0f, 1.0f); p.rad = mk(0.0f, 0.0f, 0.0f); p.irr = mk(0.0f, 0.0f, 0.0f);
                    p.depth = 0; p.sample = 0;
                    alive = true;
                    if (kStats) atomicAdd(stats, 1ull);
                }
                q_head = (q_head + take) & (kQueue - 1);
                q_count -= take;
                __syncwarp();
            }
        } else {
            // ---- no ring: idle lanes take pixels straight from the counter and generate their ray in place
            if (!exhausted && n_dead != 0u && (n_dead >= (unsigned)PTB_REFILL_MIN || dead == 0xffffffffu)) {
                unsigned base = 0;
                if (lane == 0) base = atomicAdd(P.counters, n_dead);
                base = __shfl_sync(0xffffffffu, base, 0);
                const unsigned total = (kBatch ? P.tiles_total * (unsigned)P.batch : P.tiles_total) * 32u;
                if (base + n_dead >= total) exhausted = true;
                const unsigned idx = base + (unsigned)__popc(dead & lt_mask);
                if (!alive && idx < total) {
                    unsigned tile = idx >> 5;
                    const unsigned in = idx & 31u;
                    unsigned fb = 0u;
                    if constexpr (kBatch) { fb = tile / P.tiles_total; tile -= fb * P.tiles_total; p.fb = (int)fb; }
                    const unsigned tyu = P.tiles_magic ? __umulhi(tile, P.tiles_magic) : tile / P.tiles_x;
                    const int x = (int)(tile - tyu * P.tiles_x) * 8 + (int)(in & 7u), lr = (int)tyu * 4 + (int)(in >> 3);
                    const int y = lr < P.local_rows ? global_row(P, lr) : P.height;
                    if (x < P.width && y < P.height) {
                        p.px = x; p.lrow = lr; p.py = y;
                        p.rng = ((uint32_t)x * 1973u + (uint32_t)y * 9277u + ((uint32_t)P.frame + fb) * 2699u) | 1u;   // pt:106
                        primary_ray(P, p);
                        p.irr = mk(0.0f, 0.0f, 0.0f);
                        p.sample = 0;
                        alive = true;
                        if (kStats) atomicAdd(stats, 1ull);
                    }
                }
            }
        }
        const bool starved = exhausted && q_count == 0u;       // this warp can no longer refill its lanes
        const unsigned live = __ballot_sync(0xffffffffu, alive);
        if (live == 0u) {
            if (starved) break;
            continue;   // the tile held only padding pixels; fetch again
        }
        if (alive && fresh) {      // SPP > 1: the next sample continues this pixel's RNG stream, so it is generated in place
            p.py = global_row(P, p.lrow);
            primary_ray(P, p);
            fresh = false;
            if (kStats) atomicAdd(stats, 1ull);
        }
        // ---- one bounce for every live lane.  In the tail of the frame (nothing left to pull, at most half the lanes alive)
        //      the lanes of the warp share the fold of the remaining rays, which shortens the frame's critical path.
        float T = kFloatMax;
        int prim = -1;
        bool inside = false;
        const unsigned n_live = (unsigned)__popc(live);
        if (starved && n_live <= (unsigned)PTB_COOP_MAX && P.ray_depth > 0) {
            trace_group(sc, lane, live, n_live, p.o, p.d, T, prim, inside);
        } else if (alive && P.ray_depth > 0) {
            if constexpr (kFold == 1) trace_bvh(sc, p.o, p.d, T, prim, inside);
            else if constexpr (kFold == 3) trace_grid(sc, grid, p.o, p.d, T, prim, inside);
            else if constexpr (kFold == 2) trace_rct(P, sc, p.o, p.d, T, prim, inside);
            else trace_any(sc, p.o, p.d, T, prim, inside);
        }
        int ended = kContinue;
        if (alive) ended = P.ray_depth > 0 ? shade(P, sc, p, T, prim, inside, stats, defer) : kDone;
        if constexpr (defer) {
            // SPP == 1: the sample was the pixel's only one.  Park it and free the lane; 32 parked completions are shaded together.
            const bool park = alive && ended != kContinue;
            const unsigned pm = __ballot_sync(0xffffffffu, park);
            if (pm != 0u) {
                if (park) {
                    const unsigned e = dq_count + (unsigned)__popc(pm & lt_mask);
                    doneq[0 * kQueue + e] = __float_as_uint(p.d.x); doneq[1 * kQueue + e] = __float_as_uint(p.d.y); doneq[2 * kQueue + e] = __float_as_uint(p.d.z);
                    doneq[3 * kQueue + e] = __float_as_uint(p.thr.x); doneq[4 * kQueue + e] = __float_as_uint(p.thr.y); doneq[5 * kQueue + e] = __float_as_uint(p.thr.z);
                    doneq[6 * kQueue + e] = __float_as_uint(p.rad.x); doneq[7 * kQueue + e] = __float_as_uint(p.rad.y); doneq[8 * kQueue + e] = __float_as_uint(p.rad.z);
                    doneq[9 * kQueue + e] = (uint32_t)p.px | ((uint32_t)p.lrow << 16);
                    doneq[10 * kQueue + e] = (kBatch ? (uint32_t)p.fb : 0u) | (ended == kMissed && P.ray_depth > 0 ? 0x100u : 0u);
                    alive = false;
                }
                dq_count += (unsigned)__popc(pm);
                __syncwarp();
                if (dq_count >= 32u) { dq_count -= 32u; flush_done(dq_count, 32u); }
            }
        } else if (alive && ended != kContinue) {
            p.irr = p.irr + p.rad;                // pt:123
            if (++p.sample < P.spp) fresh = true;
            else {
                if constexpr (kBatch) finish_pixel(P, p, (size_t)p.fb * (size_t)P.scratch_stride);
                else finish_pixel(P, p);
                alive = false;
            }
        }
    }
    if constexpr (defer) {
        if (dq_count != 0u) flush_done(0u, dq_count);          // what is still parked when the warp runs out of work
    }

    // ---- last CTA out re-arms the counters for the next launch
    __syncthreads();
    if (threadIdx.x == 0) {
        if (P.ktime) atomicMax(P.ktime + 1, global_ns());                   // ... and when the last one ends
        __threadfence();
        if (atomicAdd(P.counters + 1, 1u) == gridDim.x - 1) {
            P.counters[0] = 0u; P.counters[1] = 0u;
            __threadfence();
            if (P.done_flag) st_release_gpu(P.done_flag, P.done_value);     // every CTA fenced before its increment: all estimates are visible
        }
    }
}

#ifndef PTB_MEGA_ONLY
// ------------------------------------------------------------------------------------------------------------
// The labelled GL-compute proxy: the reference's own launch shape (8x8 groups, one invocation per pixel,
// ceil(W/8) x ceil(H/8) groups — PathTracer.cs:121, pt:8) reading the raw std140 UBO bytes.  Baseline only.
__global__ void __launch_bounds__(64) naive_kernel(const __grid_constant__ RenderParams P)
{
    const int x = blockIdx.x * 8 + threadIdx.x, lr = blockIdx.y * 8 + threadIdx.y;
    if (x >= P.width || lr >= P.local_rows) return;      // GL discards out-of-bounds image stores
    RawScene sc;
    sc.ubo = P.raw_objects; sc.nS = P.n_spheres; sc.nC = P.n_cuboids; sc.max_spheres = P.max_spheres;
    Path p;
    p.px = x; p.lrow = lr; p.py = global_row(P, lr);
    if (p.py >= P.height) return;
    p.rng = ((uint32_t)p.px * 1973u + (uint32_t)p.py * 9277u + (uint32_t)P.frame * 2699u) | 1u;
    p.irr = mk(0.0f, 0.0f, 0.0f);
    for (p.sample = 0; p.sample < P.spp; ++p.sample) {
        primary_ray(P, p);
        if (P.ray_depth > 0)
            while (bounce(P, sc, p, nullptr)) {}
        p.irr = p.irr + p.rad;
    }
    finish_pixel(P, p);
}

// ------------------------------------------------------------------------------------------------------------
// Scene repack: std140 GameObjectsUBO bytes -> SoA block (run once per scene edit, one thread per primitive).
__global__ void pack_scene_kernel(const unsigned char* __restrict__ ubo, int max_spheres, int nS, int nC, float4* __restrict__ block,
                                  int off_aux, int off_cmin, int off_cmax, int off_mat)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < nS) {
        const float4* s = reinterpret_cast<const float4*>(ubo + (size_t)i * kSphereStride);
        const float4 g = s[0];
        block[i] = make_float4(g.x, g.y, g.z, g.w * g.w);
        reinterpret_cast<float*>(block + off_aux)[i] = rcp(g.w);
        for (int k = 0; k < 4; ++k) block[off_mat + i * 4 + k] = s[1 + k];
        if (i == nS - 1)                                         // pad to a multiple of four with spheres nothing can hit
            for (int q = nS; q < ((nS + 3) & ~3); ++q) block[q] = make_float4(0.0f, 0.0f, 0.0f, __uint_as_float(0xff800000u));
    } else if (i < nS + nC) {
        const int c = i - nS;
        const float4* s = reinterpret_cast<const float4*>(ubo + (size_t)max_spheres * kSphereStride + (size_t)c * kCuboidStride);
        block[off_cmin + 2 * c] = s[0];
        block[off_cmax + 2 * c] = s[1];
        for (int k = 0; k < 4; ++k) block[off_mat + i * 4 + k] = s[2 + k];
    }
}

// ------------------------------------------------------------------------------------------------------------
// Seamless cubemap padding.  Texel centres live on an integer lattice: face f, texel (i,j) -> coordinates
// a = 2i+1-N, b = 2j+1-N on the face's (sc, tc) axes and +-N on its major axis (GL 4.5 Table 8.19).
struct FaceAxes { int maj, msgn, sax, ssgn, tax, tsgn; };
__device__ __forceinline__ FaceAxes face_axes(int f)
{
    switch (f) {
    case 0: return {0, 1, 2, -1, 1, -1};
    case 1: return {0, -1, 2, 1, 1, -1};
    case 2: return {1, 1, 0, 1, 2, 1};
    case 3: return {1, -1, 0, 1, 2, -1};
    case 4: return {2, 1, 0, 1, 1, -1};
    default: return {2, -1, 0, -1, 1, -1};
    }
}
__device__ __forceinline__ float4 lattice_texel(const float4* __restrict__ faces, int N, const int q[3])
{
    int axis = 0;
    if (q[1] == N || q[1] == -N) axis = 1;
    if (q[2] == N || q[2] == -N) axis = 2;
    if (q[0] == N || q[0] == -N) axis = 0;
    const int f = 2 * axis + (q[axis] < 0 ? 1 : 0);
    const FaceAxes A = face_axes(f);
    const int a = A.ssgn * q[A.sax], b = A.tsgn * q[A.tax];
    return faces[((size_t)f * N + (b + N - 1) / 2) * N + (a + N - 1) / 2];
}
__global__ void pad_cubemap_kernel(const float4* __restrict__ faces, int N, float4* __restrict__ padded)
{
    const int P = N + 2;
    const int pi = blockIdx.x * blockDim.x + threadIdx.x, pj = blockIdx.y * blockDim.y + threadIdx.y, f = blockIdx.z;
    if (pi >= P || pj >= P) return;
    const int i = pi - 1, j = pj - 1;
    const bool oi = (i < 0 || i >= N), oj = (j < 0 || j >= N);
    float4 out;
    if (!oi && !oj) {
        out = faces[((size_t)f * N + j) * N + i];
    } else {
        const FaceAxes A = face_axes(f);
        if (oi && oj) {
            // beyond a corner: mean of the three texels meeting there, summed in face order
            const int ci = i < 0 ? 0 : N - 1, cj = j < 0 ? 0 : N - 1;
            int c[3];
            c[A.maj] = A.msgn * N; c[A.sax] = A.ssgn * (2 * ci + 1 - N); c[A.tax] = A.tsgn * (2 * cj + 1 - N);
            float4 tex[6];
            bool have[6] = {false, false, false, false, false, false};
            for (int ax = 0; ax < 3; ++ax) {
                int q[3];
                for (int k = 0; k < 3; ++k) { const int sg = c[k] < 0 ? -1 : 1; q[k] = (k == ax) ? sg * N : sg * (N - 1); }
                const int ff = 2 * ax + (q[ax] < 0 ? 1 : 0);
                tex[ff] = lattice_texel(faces, N, q);
                have[ff] = true;
            }
            float sx = 0.0f, sy = 0.0f, sz = 0.0f;
            bool first = true;
            for (int ff = 0; ff < 6; ++ff)
                if (have[ff]) {
                    if (first) { sx = tex[ff].x; sy = tex[ff].y; sz = tex[ff].z; first = false; }
                    else { sx = sx + tex[ff].x; sy = sy + tex[ff].y; sz = sz + tex[ff].z; }
                }
            out = make_float4(sx * 0.333333343f, sy * 0.333333343f, sz * 0.333333343f, 1.0f);
        } else {
            // beyond one edge: fold over the shared edge onto the adjacent face
            int q[3];
            q[A.maj] = A.msgn * (N - 1);
            const int a = 2 * i + 1 - N, b = 2 * j + 1 - N;
            q[A.sax] = A.ssgn * (oi ? (a < 0 ? -N : N) : a);
            q[A.tax] = A.tsgn * (oj ? (b < 0 ? -N : N) : b);
            out = lattice_texel(faces, N, q);
        }
    }
    padded[((size_t)f * P + pj) * P + pi] = out;
}

// ------------------------------------------------------------------------------------------------------------
// Atmosphere cubemap producer: res/shaders/AtmosphericScattering/compute.glsl (cited as at:LINE).
struct AtmosParams {
    float ubo[16 * 7];   // InvProjection + InvView[6]  (at:12-16)
    float light[3];      // lightPos (at:23)
    float intensity;     // lightIntensity (at:25)
    int i_steps, j_steps, size;
};
__device__ __forceinline__ void rsi(V3 r0, V3 rd, float sr, float& x, float& y)   // at:58-71
{
    const float a = dot(rd, rd);
    const float b = 2.0f * dot(rd, r0);
    const float c = dot(r0, r0) - (sr * sr);
    const float d = (b * b) - 4.0f * a * c;
    if (d < 0.0f) { x = 1e5f; y = -1e5f; return; }
    const float sq = fsqrt(d);
    x = fdiv(-b - sq, 2.0f * a);
    y = fdiv(-b + sq, 2.0f * a);
}
__global__ void atmosphere_kernel(const __grid_constant__ AtmosParams A, float4* __restrict__ faces)
{
    const int x = blockIdx.x * blockDim.x + threadIdx.x, y = blockIdx.y * blockDim.y + threadIdx.y, f = blockIdx.z;
    if (x >= A.size || y >= A.size) return;                                     // at:34
    const float isz = rcp((float)A.size);
    const float nx = (float)x * isz * 2.0f - 1.0f, ny = (float)y * isz * 2.0f - 1.0f;   // at:37
    const float* IP = A.ubo;
    const float* IV = A.ubo + 16 + 16 * f;
    const float ex = mat_row(IP, 0, nx, ny, -1.0f, 0.0f), ey = mat_row(IP, 1, nx, ny, -1.0f, 0.0f);
    V3 r = normalize(mk(mat_row(IV, 0, ex, ey, -1.0f, 0.0f), mat_row(IV, 1, ex, ey, -1.0f, 0.0f), mat_row(IV, 2, ex, ey, -1.0f, 0.0f)));
    // at:41-53 constants
    const V3 r0 = mk(0.0f, 6376e3f, 0.0f);
    const float rPlanet = 6371e3f, rAtmos = 6471e3f, kMie = 21e-6f, shRlh = 8e3f, shMie = 1.2e3f, g = 0.758f;
    const V3 kRlh = mk(5.5e-6f, 13.0e-6f, 22.4e-6f);
    // at:73-159
    const V3 pSun = normalize(mk(A.light[0], A.light[1], A.light[2]));
    r = normalize(r);
    V3 col = mk(0.0f, 0.0f, 0.0f);
    float px_, py_;
    rsi(r0, r, rAtmos, px_, py_);
    if (!(px_ > py_)) {
        float qx, qy;
        rsi(r0, r, rPlanet, qx, qy);
        py_ = fmin_(py_, qx);
        const float iStep = fdiv(py_ - px_, (float)A.i_steps);
        float iTime = 0.0f, iOdR = 0.0f, iOdM = 0.0f;
        V3 totR = mk(0.0f, 0.0f, 0.0f), totM = mk(0.0f, 0.0f, 0.0f);
        const float mu = dot(r, pSun), mumu = mu * mu, gg = g * g;
        const float pR = fdiv(3.0f, 16.0f * kPi) * (1.0f + mumu);
        const float pM = fdiv(fdiv(3.0f, 8.0f * kPi) * ((1.0f - gg) * (mumu + 1.0f)), pow15(1.0f + gg - 2.0f * mu * g) * (2.0f + gg));
        const float ishR = rcp(shRlh), ishM = rcp(shMie);
        for (int i = 0; i < A.i_steps; ++i) {
            const V3 iPos = r0 + r * (iTime + iStep * 0.5f);
            const float iH = length(iPos) - rPlanet;
            const float odR = exp_(-iH * ishR) * iStep;
            const float odM = exp_(-iH * ishM) * iStep;
            iOdR += odR;
            iOdM += odM;
            float jx, jy;
            rsi(iPos, pSun, rAtmos, jx, jy);
            const float jStep = fdiv(jy, (float)A.j_steps);
            float jTime = 0.0f, jOdR = 0.0f, jOdM = 0.0f;
            for (int j = 0; j < A.j_steps; ++j) {
                const V3 jPos = iPos + pSun * (jTime + jStep * 0.5f);
                const float jH = length(jPos) - rPlanet;
                jOdR += exp_(-jH * ishR) * jStep;
                jOdM += exp_(-jH * ishM) * jStep;
                jTime += jStep;
            }
            const float m = kMie * (iOdM + jOdM);
            const float rl = iOdR + jOdR;
            const V3 attn = mk(exp_(-(m + kRlh.x * rl)), exp_(-(m + kRlh.y * rl)), exp_(-(m + kRlh.z * rl)));
            totR = totR + attn * odR;
            totM = totM + attn * odM;
            iTime += iStep;
        }
        col = mk(A.intensity * (pR * kRlh.x * totR.x + pM * kMie * totM.x),
                 A.intensity * (pR * kRlh.y * totR.y + pM * kMie * totM.y),
                 A.intensity * (pR * kRlh.z * totR.z + pM * kMie * totM.z));
    }
    faces[((size_t)f * A.size + y) * A.size + x] = make_float4(col.x, col.y, col.z, 1.0f);   // at:55
}

// ------------------------------------------------------------------------------------------------------------
// The atmosphere producer for LIVE regeneration (the GUI re-runs the pass on every slider tick, sizes up to 2048^2,
// Gui.cs:93-143).  Same loops as atmosphere_kernel (at:73-159) — 15 of every 16 exponentials sit in the secondary loop — with
// the special-function unit for exp / sqrt / 1/x and hand-fused multiply-adds: a third of the instructions per texel, results
// within ~1e-5 relative of atmosphere_kernel, which stays bit-exact with the shader and is this kernel's checker
// (tests/test_parity_gpu.py::test_fast_atmosphere_against_the_exact_kernel).  One thread per texel, 256-thread CTAs over the
// flat texel index.  (A tabulation of the secondary loop over (radius, sun-zenith cosine) was tried and dropped: with the sun on
// the horizon — the default, Time = 0.5 — the shader's 15-point sum along a ~1000 km grazing path changes by a factor of two
// per 1e-3 of cosine, and no affordable table follows it to better than 10 %.)
__device__ __forceinline__ float fast_exp(float x) { return exp2f(x * 1.44269504f); }     // ex2.approx after a multiply
__device__ __forceinline__ float fast_rcp(float x) { float r; asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x)); return r; }
__device__ __forceinline__ float fast_sqrt(float x) { float r; asm("sqrt.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x)); return r; }
__device__ __forceinline__ void rsi_fast(V3 r0, V3 rd, float sr, float& x, float& y)
{
    const float a = dot(rd, rd);
    const float b = 2.0f * dot(rd, r0);
    const float c = __fmaf_rn(-sr, sr, dot(r0, r0));
    const float d = __fmaf_rn(b, b, -4.0f * a * c);
    if (d < 0.0f) { x = 1e5f; y = -1e5f; return; }
    const float sq = fast_sqrt(d), ia = fast_rcp(2.0f * a);
    x = (-b - sq) * ia;
    y = (-b + sq) * ia;
}
__global__ void __launch_bounds__(256) atmosphere_fast_kernel(const __grid_constant__ AtmosParams A, float4* __restrict__ faces)
{
    const int idx = blockIdx.x * blockDim.x + threadIdx.x;
    const int n2 = A.size * A.size;
    if (idx >= 6 * n2) return;
    const int f = idx / n2, y = (idx - f * n2) / A.size, x = idx - f * n2 - y * A.size;
    const float isz = rcp((float)A.size);
    const float nx = (float)x * isz * 2.0f - 1.0f, ny = (float)y * isz * 2.0f - 1.0f;
    const float* IP = A.ubo;
    const float* IV = A.ubo + 16 + 16 * f;
    const float ex = mat_row(IP, 0, nx, ny, -1.0f, 0.0f), ey = mat_row(IP, 1, nx, ny, -1.0f, 0.0f);
    const V3 r = normalize(mk(mat_row(IV, 0, ex, ey, -1.0f, 0.0f), mat_row(IV, 1, ex, ey, -1.0f, 0.0f), mat_row(IV, 2, ex, ey, -1.0f, 0.0f)));
    const V3 r0 = mk(0.0f, 6376e3f, 0.0f);
    const float rPlanet = 6371e3f, rAtmos = 6471e3f, kMie = 21e-6f, shRlh = 8e3f, shMie = 1.2e3f, g = 0.758f;
    const V3 kRlh = mk(5.5e-6f, 13.0e-6f, 22.4e-6f);
    const V3 pSun = normalize(mk(A.light[0], A.light[1], A.light[2]));
    V3 col = mk(0.0f, 0.0f, 0.0f);
    float px_, py_;
    rsi(r0, r, rAtmos, px_, py_);
    if (!(px_ > py_)) {
        float qx, qy;
        rsi(r0, r, rPlanet, qx, qy);
        py_ = fmin_(py_, qx);
        const float iStep = fdiv(py_ - px_, (float)A.i_steps);
        float iTime = 0.0f, iOdR = 0.0f, iOdM = 0.0f;
        V3 totR = mk(0.0f, 0.0f, 0.0f), totM = mk(0.0f, 0.0f, 0.0f);
        const float mu = dot(r, pSun), mumu = mu * mu, gg = g * g;
        const float pR = fdiv(3.0f, 16.0f * kPi) * (1.0f + mumu);
        const float pM = fdiv(fdiv(3.0f, 8.0f * kPi) * ((1.0f - gg) * (mumu + 1.0f)), pow15(1.0f + gg - 2.0f * mu * g) * (2.0f + gg));
        // exp(-h / H) = 2^(h * k) with k = -log2(e) / H: one multiply-add and one ex2 per exponential
        const float kR = -1.44269504f * rcp(shRlh), kM = -1.44269504f * rcp(shMie);
        const float ij = rcp((float)A.j_steps);
        for (int i = 0; i < A.i_steps; ++i) {
            const float ti = iTime + iStep * 0.5f;
            const V3 iPos = mk(__fmaf_rn(r.x, ti, r0.x), __fmaf_rn(r.y, ti, r0.y), __fmaf_rn(r.z, ti, r0.z));
            const float iH = fast_sqrt(dot(iPos, iPos)) - rPlanet;
            const float odR = exp2f(iH * kR) * iStep;
            const float odM = exp2f(iH * kM) * iStep;
            iOdR += odR;
            iOdM += odM;
            float jx, jy;
            rsi_fast(iPos, pSun, rAtmos, jx, jy);
            const float jStep = jy * ij;
            float jTime = jStep * 0.5f, jOdR = 0.0f, jOdM = 0.0f;
            for (int j = 0; j < A.j_steps; ++j) {
                const V3 jPos = mk(__fmaf_rn(pSun.x, jTime, iPos.x), __fmaf_rn(pSun.y, jTime, iPos.y), __fmaf_rn(pSun.z, jTime, iPos.z));
                const float jH = fast_sqrt(dot(jPos, jPos)) - rPlanet;
                jOdR += exp2f(jH * kR);
                jOdM += exp2f(jH * kM);
                jTime += jStep;
            }
            jOdR *= jStep;
            jOdM *= jStep;
            const float m = kMie * (iOdM + jOdM);
            const float rl = iOdR + jOdR;
            const V3 attn = mk(fast_exp(-(m + kRlh.x * rl)), fast_exp(-(m + kRlh.y * rl)), fast_exp(-(m + kRlh.z * rl)));
            totR = totR + attn * odR;
            totM = totM + attn * odM;
            iTime += iStep;
        }
        col = mk(A.intensity * (pR * kRlh.x * totR.x + pM * kMie * totM.x),
                 A.intensity * (pR * kRlh.y * totR.y + pM * kMie * totM.y),
                 A.intensity * (pR * kRlh.z * totR.z + pM * kMie * totM.z));
    }
    faces[idx] = make_float4(col.x, col.y, col.z, 1.0f);
}

// ------------------------------------------------------------------------------------------------------------
// Fused blend + exchange for one-process-per-GPU rendering.  Every rank folds its frame estimate into its local stripes AND
// stores the blended pixels straight into rank 0's row-major image (a CUDA-IPC peer mapping: the stores travel over NVLink),
// so the frame needs no staging copy, no NCCL gather and no de-interleave pass.  Flow control lives in a small flag block
// next to the images on rank 0 (system-scope atomics): `consumed` = frames the consumer has released, `arrived[slot]` = ranks
// that have finished writing that slot.  Every wait is bounded by a timeout that raises `error` instead of hanging the GPU.
constexpr int kMaxSlots = 32;
struct ExchangeFlags {
    unsigned arrived[kMaxSlots];   // per slot, monotonic: += 1 per rank per use
    unsigned consumed;       // frames released by rank 0's consumer
    unsigned pad1[15];
    unsigned error;          // != 0: a wait timed out
};
constexpr unsigned long long kExchangeTimeoutNs = 4000000000ull;

__device__ __forceinline__ unsigned ld_acquire_sys(const unsigned* p)
{
    unsigned v;
    asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ void wait_at_least(const unsigned* flag, unsigned target, unsigned* error)
{
    const unsigned long long t0 = global_ns();
    while ((int)(ld_acquire_sys(flag) - target) < 0) {
        if (ld_acquire_sys(error) != 0u) break;          // sticky: after one timeout nobody waits again
        __nanosleep(200);
        if (global_ns() - t0 > kExchangeTimeoutNs) { atomicExch_system(error, 1u); break; }
    }
}

// Blend + scatter for 1..16 consecutive frames (ptb_render: one frame from its scratch image; batches: all of them).  One
// thread per pixel, all its estimates requested up front (see blend_batch_kernel); frame j's blended pixel goes to rank 0's
// slot `slot[j]` (a CUDA-IPC peer mapping: the stores travel over NVLink) as RGBA32F or — `rgb` — as packed RGB32F: the alpha
// the shader stores is the constant 1.0 (pt:129), and rank 0's NVLink ingress (7/8 of every frame at N = 8) is what bounds the
// exchange, so it is not shipped.  Four neighbouring lanes then pass their colours one lane down and three of them store a
// 128-bit piece of the quad's 48 bytes.  Every slot gets its own arrival from the kernel's last block.
struct BatchScatter {
    void* full[kMaxBatch];           // the root's row-major image of each frame's slot (peer mapping, or local on the root itself)
    ExchangeFlags* flags[kMaxBatch]; // the flag block of each frame's root (rank 0, or rank q % world with rotating roots)
    int slot[kMaxBatch];
    unsigned need[kMaxBatch];        // frames the root's consumer must have released before the slot may be rewritten (0: none)
};
__global__ void __launch_bounds__(256, 4) blend_scatter_batch_kernel(float4* __restrict__ image, const float4* __restrict__ estimates, int width,
                                                                     int local_rows, int height, int rank, int world, int stripe_rows, int rgb,
                                                                     const __grid_constant__ BatchBlend B, const __grid_constant__ BatchScatter X,
                                                                     unsigned* block_count, const __grid_constant__ BatchWait Wt)
{
    // a slot is free once its root's consumer has released the frame that used it last: wait here (small grids only — the
    // single-frame path, whose grid is as large as the image, sends a one-thread kernel ahead instead and passes need = 0)
    if (threadIdx.x == 0)
        for (int j = 0; j < B.frames; ++j)
            if (X.need[j] > 0u) wait_at_least(&X.flags[j]->consumed, X.need[j], &X.flags[j]->error);
    wait_for_trace(Wt);
    const size_t n = (size_t)local_rows * width;
    const size_t n32 = (n + 31) & ~(size_t)31;                  // whole warps stay in the loop together (the shuffles below)
    const bool quads = rgb && (width & 3) == 0;                 // then lanes 4q..4q+3 are one aligned quad of one row
    const unsigned lane = threadIdx.x & 31u;
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n32; i += (size_t)gridDim.x * blockDim.x) {
        const bool live = i < n;
        const size_t ii = live ? i : n - 1;
        const int lrow = (int)(ii / (size_t)width), x = (int)(ii - (size_t)lrow * width);
        const int ls = lrow / stripe_rows;
        const int y = (ls * world + rank) * stripe_rows + (lrow - ls * stripe_rows);
        const size_t g = (size_t)y * width + x;                 // this pixel in the full image
        V3 acc = mk(0.0f, 0.0f, 0.0f);
        if (B.frame0 > 0) {
            const float4 l = image[ii];
            acc = mk(l.x, l.y, l.z);
        }
        const bool whole = __all_sync(0xffffffffu, live);
        for (int j0 = 0; j0 < B.frames; j0 += kBlendChunk) {
            float4 e[kBlendChunk];
#pragma unroll
            for (int j = 0; j < kBlendChunk; ++j)
                if (j0 + j < B.frames) e[j] = estimates[(size_t)(j0 + j) * B.stride + ii];
#pragma unroll
            for (int j = 0; j < kBlendChunk; ++j)
                if (j0 + j < B.frames) {
                    acc = mix(acc, mk(e[j].x, e[j].y, e[j].z), B.blend[j0 + j]);
                    void* full = X.full[j0 + j];
                    if (!rgb) {
                        if (live && y < height) static_cast<float4*>(full)[g] = make_float4(acc.x, acc.y, acc.z, 1.0f);
                    } else if (quads && whole) {
                        const float nr = __shfl_down_sync(0xffffffffu, acc.x, 1), ng = __shfl_down_sync(0xffffffffu, acc.y, 1),
                                    nb = __shfl_down_sync(0xffffffffu, acc.z, 1);
                        const unsigned p = lane & 3u;
                        const float4 piece = p == 0u ? make_float4(acc.x, acc.y, acc.z, nr) : (p == 1u ? make_float4(acc.y, acc.z, nr, ng) : make_float4(acc.z, nr, ng, nb));
                        if (p < 3u && y < height) reinterpret_cast<float4*>(static_cast<float*>(full) + (g - p) * 3)[p] = piece;
                    } else if (live && y < height) {
                        float* dst = static_cast<float*>(full) + g * 3;
                        dst[0] = acc.x; dst[1] = acc.y; dst[2] = acc.z;
                    }
                }
        }
        if (live) image[ii] = make_float4(acc.x, acc.y, acc.z, 1.0f);
    }
    __threadfence_system();
    __syncthreads();
    if (threadIdx.x == 0) {
        if (atomicAdd(block_count, 1u) == gridDim.x - 1) {       // last block of this rank: everything above is visible system-wide
            *block_count = 0u;
            __threadfence_system();
            for (int j = 0; j < B.frames; ++j) atomicAdd_system(&X.flags[j]->arrived[X.slot[j]], 1u);
        }
    }
}
__global__ void exchange_wait_free_kernel(ExchangeFlags* flags, unsigned need_consumed)
{
    wait_at_least(&flags->consumed, need_consumed, &flags->error);     // the slot's previous frame was released by the consumer
}
__global__ void exchange_acquire_kernel(ExchangeFlags* flags, int slot, unsigned target)
{
    wait_at_least(&flags->arrived[slot], target, &flags->error);
}
// a rank that owns no rows of the image (more ranks than stripes) still counts towards the slot's arrival target
__global__ void exchange_arrive_kernel(ExchangeFlags* flags, int slot)
{
    __threadfence_system();
    atomicAdd_system(&flags->arrived[slot], 1u);
}
__global__ void exchange_release_kernel(ExchangeFlags* flags, unsigned consumed)
{
    __threadfence_system();
    atomicExch_system(&flags->consumed, consumed);
}

// ------------------------------------------------------------------------------------------------------------
// De-interleave after the per-frame gather: rank-major stripe buffers -> full row-major image.
__global__ void deinterleave_kernel(const float4* __restrict__ gathered, float4* __restrict__ full, int width, int height,
                                    int world, int stripe_rows, int max_local_rows)
{
    const int x = blockIdx.x * blockDim.x + threadIdx.x, y = blockIdx.y;
    if (x >= width || y >= height) return;
    const int stripe = y / stripe_rows, r = stripe % world, ls = stripe / world;
    const int lrow = ls * stripe_rows + (y - stripe * stripe_rows);
    full[(size_t)y * width + x] = gathered[((size_t)r * max_local_rows + lrow) * width + x];
}

// ------------------------------------------------------------------------------------------------------------
// SURVEY 8f N1 — the pass that follows the path tracer: PostProcessing/fragment.glsl (pp:LINE) into an RGBA8 target
// (ScreenEffect.cs:20-37): ACES fit, linear -> sRGB, unorm8 store (clamp, *255, round to nearest even).
__device__ __forceinline__ float aces_film(float x)                       // pp:36-44
{
    const float a = 2.51f, b = 0.03f, c = 2.43f, d = 0.59f, e = 0.14f;
    const float v = fdiv(x * (a * x + b), x * (c * x + d) + e);
    return fmin_(fmax_(v, 0.0f), 1.0f);
}
__device__ __forceinline__ float linear_to_inverse_gamma(float rgb, float gamma)    // pp:28-32
{
    const float sel = rgb < 0.0031308f ? 1.0f : 0.0f;
    return mixf(pow_(rgb, fdiv(1.0f, gamma)) * 1.055f - 0.055f, rgb * 12.92f, sel);
}
__device__ __forceinline__ unsigned unorm8(float f)
{
    if (f != f) return 0u;
    f = fmin_(fmax_(f, 0.0f), 1.0f);
    return (unsigned)__float2int_rn(f * 255.0f);
}
__global__ void tonemap_kernel(const float4* __restrict__ image, size_t n, uchar4* __restrict__ out)
{
    const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const float4 c = image[i];
    const float r = linear_to_inverse_gamma(aces_film(c.x + 0.0f), 2.4f);     // pp:19-24 (Sampler1 is never bound: + 0)
    const float g = linear_to_inverse_gamma(aces_film(c.y + 0.0f), 2.4f);
    const float b = linear_to_inverse_gamma(aces_film(c.z + 0.0f), 2.4f);
    out[i] = make_uchar4((unsigned char)unorm8(r), (unsigned char)unorm8(g), (unsigned char)unorm8(b), 255);
}
// Read-back snapshot in RGB32F: drops the constant alpha (pt:129 stores 1.0), colour floats copied bit for bit.
// One thread per four pixels: four 128-bit loads, three 128-bit stores; the last thread finishes a ragged tail.
__global__ void pack_rgb_kernel(const float4* __restrict__ image, size_t n, float* __restrict__ out)
{
    const size_t q = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    const size_t i = q * 4;
    if (i >= n) return;
    if (i + 4 <= n) {
        const float4 a = image[i], b = image[i + 1], c = image[i + 2], d = image[i + 3];
        float4* o = reinterpret_cast<float4*>(out + i * 3);
        o[0] = make_float4(a.x, a.y, a.z, b.x);
        o[1] = make_float4(b.y, b.z, c.x, c.y);
        o[2] = make_float4(c.z, d.x, d.y, d.z);
    } else {
        for (size_t k = i; k < n; ++k) {
            const float4 a = image[k];
            out[k * 3] = a.x; out[k * 3 + 1] = a.y; out[k * 3 + 2] = a.z;
        }
    }
}
// SURVEY 8f N2 — Srgb8Alpha8 skybox faces (Helper.cs:18-50): sRGB decode before filtering (GL 4.5 8.24), alpha linear.
__global__ void srgb8_decode_kernel(const uchar4* __restrict__ in, size_t n, float4* __restrict__ out)
{
    const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const uchar4 t = in[i];
    const unsigned char ch[3] = {t.x, t.y, t.z};
    float lin[3];
    for (int k = 0; k < 3; ++k) {
        const float cs = fdiv((float)ch[k], 255.0f);
        lin[k] = cs <= 0.04045f ? fdiv(cs, 12.92f) : pow_(fdiv(cs + 0.055f, 1.055f), 2.4f);
    }
    out[i] = make_float4(lin[0], lin[1], lin[2], fdiv((float)t.w, 255.0f));
}

// ------------------------------------------------------------------------------------------------------------
// Unit probes for the parity tests (ptb_debug_eval).
__global__ void dbg_log_kernel(const float* in, int n, float* out)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) out[i] = log_(in[i]);
}
__global__ void dbg_sincos_kernel(const float* in, int n, float* out)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) { float s, c; sincos_(in[i], s, c); out[2 * i] = s; out[2 * i + 1] = c; }
}
__global__ void dbg_exp_kernel(const float* in, int n, float* out)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) out[i] = exp_(in[i]);
}
__global__ void dbg_pcg_kernel(uint32_t seed, int n, float* out)
{
    if (blockIdx.x == 0 && threadIdx.x == 0) { uint32_t s = seed; for (int i = 0; i < n; ++i) out[i] = rand01(s); }
}
__global__ void dbg_env_kernel(const float4* env, int N, const float* dirs, int n, float* out)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) { const V3 t = env_lookup(env, N, mk(dirs[3 * i], dirs[3 * i + 1], dirs[3 * i + 2])); out[3 * i] = t.x; out[3 * i + 1] = t.y; out[3 * i + 2] = t.z; }
}
template <class Scene>
__device__ __forceinline__ void dbg_trace_one(const Scene& sc, const float* rays, int i, float* out)
{
    const V3 o = mk(rays[6 * i], rays[6 * i + 1], rays[6 * i + 2]), d = mk(rays[6 * i + 3], rays[6 * i + 4], rays[6 * i + 5]);
    float T; int prim; bool inside;
    trace_any(sc, o, d, T, prim, inside);
    float* q = out + 12 * i;
    const bool hit = T != kFloatMax;
    q[0] = hit ? 1.0f : 0.0f; q[1] = T; q[2] = (hit && inside) ? 1.0f : 0.0f; q[3] = 0.0f;
    for (int k = 4; k < 12; ++k) q[k] = 0.0f;
    if (hit) {
        const V3 pos = o + d * T;
        const V3 n = surface_normal(sc, prim, pos);
        q[4] = pos.x; q[5] = pos.y; q[6] = pos.z; q[7] = n.x; q[8] = n.y; q[9] = n.z;
        q[10] = sc.mat(prim, 0).x; q[11] = sc.mat(prim, 1).x;
    }
}
__global__ void dbg_trace_kernel(const __grid_constant__ RenderParams P, const float* rays, int n, float* out, int use_raw)
{
    extern __shared__ __align__(128) unsigned char smem_raw[];
    if (use_raw != 1) {
        const float4* src = P.scene;
        float4* dst = reinterpret_cast<float4*>(smem_raw);
        for (int k = threadIdx.x; k < P.stage_bytes / 16; k += blockDim.x) dst[k] = src[k];
    }
    __syncthreads();
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    if (use_raw == 1) {
        RawScene sc; sc.ubo = P.raw_objects; sc.nS = P.n_spheres; sc.nC = P.n_cuboids; sc.max_spheres = P.max_spheres;
        dbg_trace_one(sc, rays, i, out);
    } else if (use_raw == 2 || use_raw == 3 || use_raw == 4) {
        const PackedScene sc = PTB_PACKED_SCENE(P, reinterpret_cast<const float4*>(smem_raw));
        const V3 o = mk(rays[6 * i], rays[6 * i + 1], rays[6 * i + 2]), d = mk(rays[6 * i + 3], rays[6 * i + 4], rays[6 * i + 5]);
        float T; int prim; bool inside;
        int visits = 0;
        if (use_raw == 2) trace_bvh(sc, o, d, T, prim, inside, &visits);
        else if (use_raw == 4) trace_grid(sc, make_grid_view(P, reinterpret_cast<const float4*>(smem_raw)), o, d, T, prim, inside, &visits);     // q[3]: cells visited
        else {
            // q[3]: candidates the table left for this ray (65 = the ray took the full mask: outside the grid / non-finite)
            trace_rct(P, sc, o, d, T, prim, inside);
            const unsigned long long mk64 = rct_lookup(P, o, d, mk(rcp(d.x), rcp(d.y), rcp(d.z)));
            visits = mk64 == P.rct_valid ? 65 : __popcll(mk64);
        }
        float* q = out + 12 * i;
        const bool hit = T != kFloatMax;
        // q[3]: BVH nodes visited (0 = the ray took the brute-force fold: non-finite ray, or nothing but the always-tested list)
        q[0] = hit ? 1.0f : 0.0f; q[1] = T; q[2] = (hit && inside) ? 1.0f : 0.0f; q[3] = (float)visits;
        for (int k = 4; k < 12; ++k) q[k] = 0.0f;
        if (hit) {
            const V3 pos = o + d * T;
            const V3 n = surface_normal(sc, prim, pos);
            q[4] = pos.x; q[5] = pos.y; q[6] = pos.z; q[7] = n.x; q[8] = n.y; q[9] = n.z;
            q[10] = sc.mat(prim, 0).x; q[11] = sc.mat(prim, 1).x;
        }
    } else {
        const PackedScene sc = PTB_PACKED_SCENE(P, reinterpret_cast<const float4*>(smem_raw));
        dbg_trace_one(sc, rays, i, out);
    }
}
// The group-cooperative fold on its own: each warp serves k rays at a time, parked on scattered lanes (3 + 7j mod 32).
__global__ void dbg_group_kernel(const __grid_constant__ RenderParams P, const float* rays, int n, float* out, int k)
{
    extern __shared__ __align__(128) unsigned char smem_raw[];
    {
        const float4* src = P.scene;
        float4* dst = reinterpret_cast<float4*>(smem_raw);
        for (int q = threadIdx.x; q < P.stage_bytes / 16; q += blockDim.x) dst[q] = src[q];
    }
    __syncthreads();
    const PackedScene sc = PTB_PACKED_SCENE(P, reinterpret_cast<const float4*>(smem_raw));
    const unsigned lane = threadIdx.x & 31u;
    const int j = (int)(((lane + 29u) * 23u) & 31u);        // inverse of lane = (3 + 7j) mod 32
    const int i = blockIdx.x * k + j;
    const bool alive = j < k && i < n;
    V3 o = mk(0.0f, 0.0f, 0.0f), d = mk(0.0f, 0.0f, 1.0f);
    if (alive) { o = mk(rays[6 * i], rays[6 * i + 1], rays[6 * i + 2]); d = mk(rays[6 * i + 3], rays[6 * i + 4], rays[6 * i + 5]); }
    const unsigned live = __ballot_sync(0xffffffffu, alive);
    if (live == 0u) return;
    float T; int prim; bool inside;
    trace_group(sc, lane, live, (unsigned)__popc(live), o, d, T, prim, inside);
    if (!alive) return;
    float* q = out + 12 * i;
    const bool hit = T != kFloatMax;
    q[0] = hit ? 1.0f : 0.0f; q[1] = T; q[2] = (hit && inside) ? 1.0f : 0.0f; q[3] = 0.0f;
    for (int m = 4; m < 12; ++m) q[m] = 0.0f;
    if (hit) {
        const V3 pos = o + d * T;
        const V3 nn = surface_normal(sc, prim, pos);
        q[4] = pos.x; q[5] = pos.y; q[6] = pos.z; q[7] = nn.x; q[8] = nn.y; q[9] = nn.z;
        q[10] = sc.mat(prim, 0).x; q[11] = sc.mat(prim, 1).x;
    }
}
__global__ void dbg_arith_kernel(const float* in, int n, float* out)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) {
        const float a = in[2 * i], b = in[2 * i + 1];
        out[4 * i] = fmin_(a, b); out[4 * i + 1] = fmax_(a, b); out[4 * i + 2] = rcp(a); out[4 * i + 3] = fsqrt(b);
    }
}

#endif  // PTB_MEGA_ONLY

} // namespace ptb
