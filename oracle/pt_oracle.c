/*
 * oracle/pt_oracle.c — TEST INFRASTRUCTURE ONLY.  CPU restatement of the reference's
 * path-tracing compute shader, function by function, in the order the GLSL has them.
 * All citations are /root/reference/OpenTK-PathTracer/res/shaders/PathTracing/compute.glsl:LINE
 * unless another file is named.  Arithmetic follows oracle/glsl_model.h.
 *
 * PARITY PIN: the reference ships no tests / golden vectors for this path, so this file is pinned
 * against the reference's own shader instead: compute.glsl compiled for the CPU from /root/reference
 * (oracle/build_ref.py -> oracle/_ref/libglsl_ref.so) must agree with it bit for bit on whole
 * dispatches, RayTrace() probes, the RNG stream and the post-process pass
 * (tests/test_reference_pin.py; committed outputs tests/golden/ref_*.npz).  See DESIGN.md §2.
 *
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs
 * may load this library.  The product never does.
 */
#include "glsl_model.h"
#include <stdlib.h>
#ifdef _OPENMP
#include <omp.h>
#endif

#define FLOAT_MAX 3.4028235e+38f  /* compute.glsl:2 */
#define FLOAT_MIN -3.4028235e+38f /* compute.glsl:3 */
#define EPSILON 0.001f            /* compute.glsl:4 */
#define PI 3.14159265f            /* compute.glsl:5 */

/* compute.glsl:13-26 — std140: 4 x vec4 = 64 B (Material.cs:36-51) */
typedef struct {
    vec3 Albedo;     float SpecularChance;
    vec3 Emissiv;    float SpecularRoughness;
    vec3 Absorbance; float RefractionChance;
    float RefractionRoughness; float IOR; float pad0; float pad1;
} Material;
/* compute.glsl:28-34 — 96 B: Min @0, Max @16, Material @32 (Cuboid.cs:27-35) */
typedef struct { vec3 Min; float pad0; vec3 Max; float pad1; Material Mat; } Cuboid;
/* compute.glsl:36-42 — 80 B: vec4(pos, radius), Material @16 (Sphere.cs:23-31) */
typedef struct { vec3 Position; float Radius; Material Mat; } Sphere;
/* compute.glsl:44-51 */
typedef struct { float T; int FromInside; vec3 NearHitPos; vec3 Normal; Material Mat; } HitInfo;
/* compute.glsl:53-57 */
typedef struct { vec3 Origin; vec3 Direction; } Ray;

/* Everything one dispatch reads (UBO 0, UBO 1, plain uniforms, samplerCube, image). */
typedef struct {
    int width, height;           /* imageSize(ImgResult), compute.glsl:103 */
    int frame;                   /* thisRendererFrame, compute.glsl:96 */
    int spp, ray_depth;          /* compute.glsl:90-91 */
    float focal_length;          /* compute.glsl:93 */
    float aperture_diameter;     /* compute.glsl:94 */
    float n_spheres, n_cuboids;  /* uboGameObjectsSize (a float vec2!), compute.glsl:88 */
    int max_spheres;             /* capacity of Spheres[] => byte offset of Cuboids[], Cuboid.cs:21 */
    int env_size;                /* cubemap face edge */
    int y0, y1;                  /* rows [y0,y1) to render (crop); others untouched */
    int x0, x1;                  /* columns [x0,x1) */
    int n_threads;               /* OpenMP threads; <=0 = all */
    int y_step;                  /* render rows y0, y0+y_step, ... (<=1 = every row): an evenly spread sample */
} pto_params;

typedef struct {
    uint64_t samples;        /* primary rays traced */
    uint64_t bounces;        /* RayTrace calls */
    uint64_t hits;           /* RayTrace calls that returned true */
    uint64_t rng_draws;
    uint64_t nonfinite_pixels;
    uint64_t depth_hist[64]; /* RayTrace calls per sample, clamped to 63 */
} pto_stats;

typedef struct {
    const pto_params *p;
    const float *basic;    /* UBO 0 as 36 floats: InvProjection @0, InvView @16, ViewPos @32 */
    const uint8_t *objects;/* UBO 1 bytes */
    const float *env;      /* 6 faces x N x N x RGBA32F, faces +X,-X,+Y,-Y,+Z,-Z, row = t */
    uint32_t rndSeed;      /* compute.glsl:98 */
    pto_stats *st;
} Ctx;

/* ------------------------------------------------------------------------------------------ */
/* compute.glsl:334-339 */
static uint32_t GetPCGHash(uint32_t *seed)
{
    *seed = *seed * 747796405u + 2891336453u;
    uint32_t word = ((*seed >> ((*seed >> 28u) + 4u)) ^ *seed) * 277803737u;
    return (word >> 22u) ^ word;
}
/* compute.glsl:341-344.  /4294967296.0 == * 2^-32 exactly; float(h) rounds (RNE) so 1.0 is reachable. */
static float GetRandomFloat01(Ctx *c)
{
    if (c->st) c->st->rng_draws++;
    return g_div((float)GetPCGHash(&c->rndSeed), 4294967296.0f);
}
/* compute.glsl:347-350 */
static float GetSmallestPositive(float t1, float t2) { return t1 < 0.0f ? t2 : t1; }

/* compute.glsl:261-277 */
static int RaySphereIntersect(Ray ray, const Sphere *sphere, float *t1, float *t2)
{
    *t1 = *t2 = FLOAT_MAX;
    vec3 sphereToRay = v_sub(ray.Origin, sphere->Position);
    float b = v_dot(ray.Direction, sphereToRay);
    float c = v_dot(sphereToRay, sphereToRay) - sphere->Radius * sphere->Radius;
    float discriminant = b * b - c;
    if (discriminant < 0.0f)
        return 0;
    float squareRoot = g_sqrt(discriminant);
    *t1 = -b - squareRoot;
    *t2 = -b + squareRoot;
    return *t1 <= *t2;
}

/* compute.glsl:280-294.  vec3 / vec3 is component-wise a * rcp(b). */
static int RayCuboidIntersect(Ray ray, const Cuboid *cuboid, float *t1, float *t2)
{
    *t1 = FLOAT_MIN;
    *t2 = FLOAT_MAX;
    float ix = g_rcp(ray.Direction.x), iy = g_rcp(ray.Direction.y), iz = g_rcp(ray.Direction.z);
    vec3 t0s = v3((cuboid->Min.x - ray.Origin.x) * ix, (cuboid->Min.y - ray.Origin.y) * iy, (cuboid->Min.z - ray.Origin.z) * iz);
    vec3 t1s = v3((cuboid->Max.x - ray.Origin.x) * ix, (cuboid->Max.y - ray.Origin.y) * iy, (cuboid->Max.z - ray.Origin.z) * iz);
    vec3 tsmaller = v3(g_min(t0s.x, t1s.x), g_min(t0s.y, t1s.y), g_min(t0s.z, t1s.z));
    vec3 tbigger = v3(g_max(t0s.x, t1s.x), g_max(t0s.y, t1s.y), g_max(t0s.z, t1s.z));
    *t1 = g_max(*t1, g_max(tsmaller.x, g_max(tsmaller.y, tsmaller.z)));
    *t2 = g_min(*t2, g_min(tbigger.x, g_min(tbigger.y, tbigger.z)));
    return *t1 <= *t2;
}

/* compute.glsl:316-319 */
static vec3 GetNormalSphere(const Sphere *sphere, vec3 surfacePosition)
{
    return v_scale(v_sub(surfacePosition, sphere->Position), g_rcp(sphere->Radius));
}
/* compute.glsl:322-332 */
static vec3 GetNormalCuboid(const Cuboid *cuboid, vec3 surfacePosition)
{
    vec3 halfSize = v_scale(v_sub(cuboid->Max, cuboid->Min), 0.5f);
    vec3 centerSurface = v_sub(surfacePosition, v_scale(v_add(cuboid->Max, cuboid->Min), 0.5f));
    vec3 normal = v3(0.0f, 0.0f, 0.0f);
    float s;
    s = g_step(g_abs(g_abs(centerSurface.x) - halfSize.x), EPSILON);
    normal = v_add(normal, v_scale(v3(g_sign(centerSurface.x), 0.0f, 0.0f), s));
    s = g_step(g_abs(g_abs(centerSurface.y) - halfSize.y), EPSILON);
    normal = v_add(normal, v_scale(v3(0.0f, g_sign(centerSurface.y), 0.0f), s));
    s = g_step(g_abs(g_abs(centerSurface.z) - halfSize.z), EPSILON);
    normal = v_add(normal, v_scale(v3(0.0f, 0.0f, g_sign(centerSurface.z)), s));
    return v_normalize(normal);
}

/* compute.glsl:226-258 — the order-dependent closest-hit fold. */
static int RayTrace(Ctx *c, Ray ray, HitInfo *hitInfo)
{
    const Sphere *spheres = (const Sphere *)c->objects;
    const Cuboid *cuboids = (const Cuboid *)(c->objects + (size_t)c->p->max_spheres * sizeof(Sphere));
    hitInfo->T = FLOAT_MAX;
    float t1, t2;
    for (int i = 0; (float)i < c->p->n_spheres; i++) {
        const Sphere *sphere = &spheres[i];
        if (RaySphereIntersect(ray, sphere, &t1, &t2) && t2 > 0.0f && t1 < hitInfo->T) {
            hitInfo->T = GetSmallestPositive(t1, t2);
            hitInfo->FromInside = hitInfo->T == t2;
            hitInfo->Mat = sphere->Mat;
            hitInfo->NearHitPos = v_add(ray.Origin, v_scale(ray.Direction, hitInfo->T));
            hitInfo->Normal = GetNormalSphere(sphere, hitInfo->NearHitPos);
        }
    }
    for (int i = 0; (float)i < c->p->n_cuboids; i++) {
        const Cuboid *cuboid = &cuboids[i];
        if (RayCuboidIntersect(ray, cuboid, &t1, &t2) && t2 > 0.0f && t1 < hitInfo->T) {
            hitInfo->T = GetSmallestPositive(t1, t2);
            hitInfo->FromInside = hitInfo->T == t2;
            hitInfo->Mat = cuboid->Mat;
            hitInfo->NearHitPos = v_add(ray.Origin, v_scale(ray.Direction, hitInfo->T));
            hitInfo->Normal = GetNormalCuboid(cuboid, hitInfo->NearHitPos);
        }
    }
    return hitInfo->T != FLOAT_MAX;
}

/* compute.glsl:297-307 — draws z then angle. */
static vec3 CosineSampleHemisphere(Ctx *c, vec3 normal)
{
    float z = GetRandomFloat01(c) * 2.0f - 1.0f;
    float a = GetRandomFloat01(c) * 2.0f * PI;
    float r = g_sqrt(1.0f - z * z);
    float sn, cs;
    g_sincos(a, &sn, &cs);
    float x = r * cs;
    float y = r * sn;
    return v_normalize(v_add(normal, v3(x, y, z)));
}
/* compute.glsl:309-314 — draws angle then r. */
static void UniformSampleUnitCircle(Ctx *c, float *ox, float *oy)
{
    float angle = GetRandomFloat01(c) * 2.0f * PI;
    float r = g_sqrt(GetRandomFloat01(c));
    float sn, cs;
    g_sincos(angle, &sn, &cs);
    *ox = cs * r;
    *oy = sn * r;
}
/* compute.glsl:359-364 */
static float FresnelSchlick(float cosTheta, float n1, float n2)
{
    float r0 = g_div(n1 - n2, n1 + n2);
    r0 *= r0;
    return r0 + (1.0f - r0) * g_pow5(1.0f - cosTheta);
}

/* compute.glsl:184-224 */
static float BSDF(Ctx *c, Ray *ray, const HitInfo *hitInfo, int *isRefractive)
{
    *isRefractive = 0;
    float specularChance = hitInfo->Mat.SpecularChance;
    float refractionChance = hitInfo->Mat.RefractionChance;
    if (specularChance > 0.0f) {
        specularChance = g_mix(specularChance, 1.0f,
            FresnelSchlick(v_dot(v_neg(ray->Direction), hitInfo->Normal),
                           hitInfo->FromInside ? hitInfo->Mat.IOR : 1.0f,
                           !hitInfo->FromInside ? hitInfo->Mat.IOR : 1.0f));
        float diffuseChance = 1.0f - specularChance - refractionChance;
        refractionChance = 1.0f - specularChance - diffuseChance;
    }
    vec3 diffuseRay = CosineSampleHemisphere(c, hitInfo->Normal);
    float rayProbability = 1.0f;
    float raySelectRoll = GetRandomFloat01(c);
    if (specularChance > raySelectRoll) {
        vec3 reflectionRayDir = v_reflect(ray->Direction, hitInfo->Normal);
        reflectionRayDir = v_normalize(v_mix(reflectionRayDir, diffuseRay,
            hitInfo->Mat.SpecularRoughness * hitInfo->Mat.SpecularRoughness));
        ray->Direction = reflectionRayDir;
        rayProbability = specularChance;
    } else if (specularChance + refractionChance > raySelectRoll) {
        vec3 refractionRayDir = v_refract(ray->Direction, hitInfo->Normal,
            hitInfo->FromInside ? g_div(hitInfo->Mat.IOR, 1.0f) : g_div(1.0f, hitInfo->Mat.IOR));
        refractionRayDir = v_normalize(v_mix(refractionRayDir, CosineSampleHemisphere(c, v_neg(hitInfo->Normal)),
            hitInfo->Mat.RefractionRoughness * hitInfo->Mat.RefractionRoughness));
        ray->Direction = refractionRayDir;
        rayProbability = refractionChance;
        *isRefractive = 1;
    } else {
        ray->Direction = diffuseRay;
        rayProbability = 1.0f - specularChance - refractionChance;
    }
    ray->Origin = v_add(hitInfo->NearHitPos, v_scale(ray->Direction, EPSILON));
    return g_max(rayProbability, EPSILON);
}

/* ------------------------------------------------------------------------------------------
 * texture(samplerCube, dir) — compute.glsl:177.  Restates OpenGL 4.5 §8.13 (Table 8.19 face
 * selection) and §8.14 (LINEAR magnification; a compute shader has no derivatives so LOD = 0),
 * with GL_TEXTURE_CUBE_MAP_SEAMLESS on (MainWindow.cs:168): taps beyond a face edge come from
 * the adjacent face; a tap beyond a corner is the average of the three texels at that corner.
 * Cited from memory of the spec (no spec text offline).
 * ------------------------------------------------------------------------------------------ */
/* Integer cube coordinates: a texel centre on face f at (i,j) is the 3D lattice point built from
 * a = 2i+1-N, b = 2j+1-N (odd, in (-N,N)) and major = +-N, per Table 8.19. */
static void cube_point(int f, int N, int a, int b, int P[3])
{
    switch (f) {
    case 0: P[0] = N;  P[1] = -b; P[2] = -a; break; /* +X: sc=-rz tc=-ry */
    case 1: P[0] = -N; P[1] = -b; P[2] = a;  break; /* -X: sc=+rz tc=-ry */
    case 2: P[0] = a;  P[1] = N;  P[2] = b;  break; /* +Y: sc=+rx tc=+rz */
    case 3: P[0] = a;  P[1] = -N; P[2] = -b; break; /* -Y: sc=+rx tc=-rz */
    case 4: P[0] = a;  P[1] = -b; P[2] = N;  break; /* +Z: sc=+rx tc=-ry */
    default: P[0] = -a; P[1] = -b; P[2] = -N; break; /* -Z: sc=-rx tc=-ry */
    }
}
/* Inverse: lattice point with exactly one coordinate == +-N -> (face, a, b). */
static void cube_unpoint(const int P[3], int N, int *f, int *a, int *b)
{
    if (P[0] == N)       { *f = 0; *a = -P[2]; *b = -P[1]; }
    else if (P[0] == -N) { *f = 1; *a = P[2];  *b = -P[1]; }
    else if (P[1] == N)  { *f = 2; *a = P[0];  *b = P[2]; }
    else if (P[1] == -N) { *f = 3; *a = P[0];  *b = -P[2]; }
    else if (P[2] == N)  { *f = 4; *a = P[0];  *b = -P[1]; }
    else                 { *f = 5; *a = -P[0]; *b = -P[1]; }
}
static const float *env_texel(const Ctx *c, int f, int i, int j)
{
    int N = c->p->env_size;
    return c->env + (((size_t)f * N + j) * N + i) * 4;
}
/* One tap of the 2x2 footprint, with i,j in [-1,N]. */
static vec3 cube_tap(const Ctx *c, int f, int i, int j)
{
    int N = c->p->env_size;
    int oi = (i < 0 || i >= N), oj = (j < 0 || j >= N);
    if (!oi && !oj) {
        const float *t = env_texel(c, f, i, j);
        return v3(t[0], t[1], t[2]);
    }
    if (oi && oj) {
        /* beyond a corner: average the three texels that meet there, summed in face order */
        int ci = i < 0 ? 0 : N - 1, cj = j < 0 ? 0 : N - 1;
        int P[3], Q[3], ff[3], aa[3], bb[3];
        cube_point(f, N, 2 * ci + 1 - N, 2 * cj + 1 - N, P);
        /* the corner texel on each of the three faces: put +-N on each axis in turn, +-(N-1) elsewhere */
        for (int ax = 0; ax < 3; ax++) {
            for (int k = 0; k < 3; k++) {
                int sgn = P[k] < 0 ? -1 : 1;
                Q[k] = (k == ax) ? sgn * N : sgn * (N - 1);
            }
            cube_unpoint(Q, N, &ff[ax], &aa[ax], &bb[ax]);
        }
        /* sort the three by face index (ascending) */
        for (int u = 0; u < 3; u++)
            for (int v = u + 1; v < 3; v++)
                if (ff[v] < ff[u]) {
                    int tf = ff[u]; ff[u] = ff[v]; ff[v] = tf;
                    int ta = aa[u]; aa[u] = aa[v]; aa[v] = ta;
                    int tb = bb[u]; bb[u] = bb[v]; bb[v] = tb;
                }
        vec3 sum = v3(0.0f, 0.0f, 0.0f);
        for (int u = 0; u < 3; u++) {
            const float *t = env_texel(c, ff[u], (aa[u] + N - 1) / 2, (bb[u] + N - 1) / 2);
            vec3 tv = v3(t[0], t[1], t[2]);
            sum = (u == 0) ? tv : v_add(sum, tv);
        }
        return v_scale(sum, 0.333333343f);
    }
    /* beyond exactly one edge: fold the lattice point over the shared edge onto the adjacent face */
    int P[3];
    cube_point(f, N, 2 * i + 1 - N, 2 * j + 1 - N, P);
    for (int k = 0; k < 3; k++) {
        if (P[k] == N || P[k] == -N) P[k] = P[k] < 0 ? -(N - 1) : (N - 1);   /* old major: one half-texel inside */
        else if (P[k] > N) P[k] = N;                                          /* overflow axis becomes the major */
        else if (P[k] < -N) P[k] = -N;
    }
    int nf, na, nb;
    cube_unpoint(P, N, &nf, &na, &nb);
    const float *t = env_texel(c, nf, (na + N - 1) / 2, (nb + N - 1) / 2);
    return v3(t[0], t[1], t[2]);
}
static vec3 TextureCube(const Ctx *c, vec3 r)
{
    int N = c->p->env_size;
    float ax = g_abs(r.x), ay = g_abs(r.y), az = g_abs(r.z);
    int f;
    float sc, tc, ma;
    /* A texture unit never returns NaN for a NaN coordinate; what it does return is hardware-specific.
     * The model defines it: a direction with a non-finite component, or the zero vector, fetches 0.
     * This is reached from valid scenes: total internal reflection with RefractionRoughness == 0 makes
     * refract() return 0 and normalize(0) is NaN (compute.glsl:210-211).  The reference's own
     * 68 510-spp screenshot shows no NaN pixels, which literal propagation through the lerp would give. */
    if (!(ax <= FLOAT_MAX && ay <= FLOAT_MAX && az <= FLOAT_MAX) || (ax == 0.0f && ay == 0.0f && az == 0.0f))
        return v3(0.0f, 0.0f, 0.0f);
    if (ax >= ay && ax >= az) { ma = ax; if (r.x >= 0.0f) { f = 0; sc = -r.z; tc = -r.y; } else { f = 1; sc = r.z; tc = -r.y; } }
    else if (ay >= ax && ay >= az) { ma = ay; if (r.y >= 0.0f) { f = 2; sc = r.x; tc = r.z; } else { f = 3; sc = r.x; tc = -r.z; } }
    else { ma = az; if (r.z >= 0.0f) { f = 4; sc = r.x; tc = -r.y; } else { f = 5; sc = -r.x; tc = -r.y; } }
    float ima = g_rcp(ma);
    float s = 0.5f * (sc * ima + 1.0f);
    float t = 0.5f * (tc * ima + 1.0f);
    float u = s * (float)N - 0.5f;
    float v = t * (float)N - 0.5f;
    float fu = g_floor(u), fv = g_floor(v);
    float alpha = u - fu, beta = v - fv;
    int i0 = g_f2i(fu), j0 = g_f2i(fv);
    /* s,t in [0,1] for finite input; clamp defends the NaN / Inf cases (result is NaN anyway) */
    if (i0 < -1) i0 = -1; if (i0 > N - 1) i0 = N - 1;
    if (j0 < -1) j0 = -1; if (j0 > N - 1) j0 = N - 1;
    vec3 t00 = cube_tap(c, f, i0, j0), t10 = cube_tap(c, f, i0 + 1, j0);
    vec3 t01 = cube_tap(c, f, i0, j0 + 1), t11 = cube_tap(c, f, i0 + 1, j0 + 1);
    vec3 top = v_mix(t00, t10, alpha);
    vec3 bot = v_mix(t01, t11, alpha);
    return v_mix(top, bot, beta);
}

/* compute.glsl:132-182 */
static vec3 Radiance(Ctx *c, Ray ray)
{
    vec3 throughput = v3(1.0f, 1.0f, 1.0f);
    vec3 radiance = v3(0.0f, 0.0f, 0.0f);
    HitInfo hitInfo;
    int isRefractive;
    float rayProbability;
    int traces = 0;
    for (int i = 0; i < c->p->ray_depth; i++) {
        traces++;
        if (RayTrace(c, ray, &hitInfo)) {
            if (c->st) c->st->hits++;
            if (hitInfo.FromInside) {
                hitInfo.Normal = v_scale(hitInfo.Normal, -1.0f);
                vec3 a = hitInfo.Mat.Absorbance;
                throughput = v_mul(throughput, v3(g_exp(-a.x * hitInfo.T), g_exp(-a.y * hitInfo.T), g_exp(-a.z * hitInfo.T)));
            }
            rayProbability = BSDF(c, &ray, &hitInfo, &isRefractive);
            radiance = v_add(radiance, v_mul(hitInfo.Mat.Emissiv, throughput));
            if (!isRefractive)
                throughput = v_mul(throughput, hitInfo.Mat.Albedo);
            throughput = v_scale(throughput, g_rcp(rayProbability));
            {
                float p = g_max(throughput.x, g_max(throughput.y, throughput.z));
                if (GetRandomFloat01(c) > p)
                    break;
                throughput = v_scale(throughput, g_rcp(p));
            }
        } else {
            radiance = v_add(radiance, v_mul(TextureCube(c, ray.Direction), throughput));
            break;
        }
    }
    if (c->st) {
        c->st->bounces += (uint64_t)traces;
        c->st->depth_hist[traces > 63 ? 63 : traces]++;
    }
    return radiance;
}

/* compute.glsl:352-357.  Only .xy of rayEye survive (.zw overwritten at :355). */
static Ray GetWorldSpaceRay(const float *inverseProj, const float *inverseView, vec3 viewPos, float ndcx, float ndcy)
{
    float ex = m4_row(inverseProj, 0, ndcx, ndcy, -1.0f, 0.0f);
    float ey = m4_row(inverseProj, 1, ndcx, ndcy, -1.0f, 0.0f);
    vec3 d = v3(m4_row(inverseView, 0, ex, ey, -1.0f, 0.0f),
                m4_row(inverseView, 1, ex, ey, -1.0f, 0.0f),
                m4_row(inverseView, 2, ex, ey, -1.0f, 0.0f));
    Ray r = { viewPos, v_normalize(d) };
    return r;
}

/* compute.glsl:101-130 for one invocation. img = the RGBA32F image (read then written, :126-129). */
static void shade_pixel(Ctx *c, int px, int py, float *img)
{
    const pto_params *p = c->p;
    const float *InvProjection = c->basic, *InvView = c->basic + 16;
    vec3 ViewPos = v3(c->basic[32], c->basic[33], c->basic[34]);

    c->rndSeed = ((uint32_t)px * 1973u + (uint32_t)py * 9277u + (uint32_t)p->frame * 2699u) | 1u; /* :106 */
    vec3 irradiance = v3(0.0f, 0.0f, 0.0f);
    float isx = g_rcp((float)p->width), isy = g_rcp((float)p->height);
    for (int i = 0; i < p->spp; i++) {
        if (c->st) c->st->samples++;
        float ox = GetRandomFloat01(c);  /* :113 — arguments evaluate left to right */
        float oy = GetRandomFloat01(c);
        float ndcx = ((float)px + ox) * isx * 2.0f - 1.0f;  /* :114 */
        float ndcy = ((float)py + oy) * isy * 2.0f - 1.0f;
        Ray ray = GetWorldSpaceRay(InvProjection, InvView, ViewPos, ndcx, ndcy); /* :115 */
        vec3 focalPoint = v_add(ray.Origin, v_scale(ray.Direction, p->focal_length)); /* :117 */
        float cx, cy;
        UniformSampleUnitCircle(c, &cx, &cy);
        float hk = p->aperture_diameter * 0.5f;  /* :118 — (apertureDiameter * 0.5) * vec2 */
        float offx = hk * cx, offy = hk * cy;
        ray.Origin = v3(m4_row(InvView, 0, offx, offy, 0.0f, 1.0f),  /* :120 */
                        m4_row(InvView, 1, offx, offy, 0.0f, 1.0f),
                        m4_row(InvView, 2, offx, offy, 0.0f, 1.0f));
        ray.Direction = v_normalize(v_sub(focalPoint, ray.Origin));   /* :121 */
        irradiance = v_add(irradiance, Radiance(c, ray));             /* :123 */
    }
    irradiance = v_scale(irradiance, g_rcp((float)p->spp));           /* :125 */
    float *px4 = img + ((size_t)py * p->width + px) * 4;
    vec3 last = v3(px4[0], px4[1], px4[2]);                           /* :126 */
    float a = g_div(1.0f, (float)(p->frame + 1));                     /* :128 */
    irradiance = v_mix(last, irradiance, a);
    px4[0] = irradiance.x; px4[1] = irradiance.y; px4[2] = irradiance.z; px4[3] = 1.0f; /* :129 */
    if (c->st) {
        uint32_t e = (g_bits(irradiance.x) & g_bits(irradiance.y) & g_bits(irradiance.z)) ;
        (void)e;
        if (((g_bits(irradiance.x) >> 23) & 255) == 255 || ((g_bits(irradiance.y) >> 23) & 255) == 255 ||
            ((g_bits(irradiance.z) >> 23) & 255) == 255)
            c->st->nonfinite_pixels++;
    }
}

/* ============================ exported entry points (ctypes) ============================== */

/* One dispatch (PathTracer.cs:114-129 + compute.glsl main) over the crop [x0,x1)x[y0,y1).
 * Pixels outside the image do not exist here (GL discards OOB image stores). */
int pto_render(const pto_params *p, const void *basic_ubo, const void *objects_ubo, const float *env,
               float *image, pto_stats *stats)
{
    if (!p || !basic_ubo || !objects_ubo || !env || !image) return -1;
    if (p->width <= 0 || p->height <= 0 || p->spp <= 0 || p->env_size <= 0) return -2;
    int y0 = p->y0 < 0 ? 0 : p->y0, y1 = p->y1 > p->height ? p->height : p->y1;
    int x0 = p->x0 < 0 ? 0 : p->x0, x1 = p->x1 > p->width ? p->width : p->x1;
    int nt = p->n_threads;
#ifdef _OPENMP
    if (nt <= 0) nt = omp_get_max_threads();
#else
    nt = 1;
#endif
    pto_stats *per = stats ? (pto_stats *)calloc((size_t)nt, sizeof(pto_stats)) : NULL;
#pragma omp parallel num_threads(nt)
    {
#ifdef _OPENMP
        int tid = omp_get_thread_num();
#else
        int tid = 0;
#endif
        Ctx c;
        c.p = p; c.basic = (const float *)basic_ubo; c.objects = (const uint8_t *)objects_ubo;
        c.env = env; c.rndSeed = 0; c.st = per ? &per[tid] : NULL;
        const int ys = p->y_step > 1 ? p->y_step : 1;
#pragma omp for schedule(dynamic, 4)
        for (int y = y0; y < y1; y += ys)
            for (int x = x0; x < x1; x++)
                shade_pixel(&c, x, y, image);
    }
    if (stats) {
        memset(stats, 0, sizeof(*stats));
        for (int t = 0; t < nt; t++) {
            stats->samples += per[t].samples; stats->bounces += per[t].bounces; stats->hits += per[t].hits;
            stats->rng_draws += per[t].rng_draws; stats->nonfinite_pixels += per[t].nonfinite_pixels;
            for (int k = 0; k < 64; k++) stats->depth_hist[k] += per[t].depth_hist[k];
        }
        free(per);
    }
    return 0;
}

int pto_max_threads(void)
{
#ifdef _OPENMP
    return omp_get_max_threads();
#else
    return 1;
#endif
}

/* ---- unit-level probes for known-answer tests ---------------------------------------------- */
/* seed (compute.glsl:106) and the first n hashes / floats of its stream (compute.glsl:334-344) */
uint32_t pto_seed(uint32_t x, uint32_t y, int frame) { return (x * 1973u + y * 9277u + (uint32_t)frame * 2699u) | 1u; }
void pto_pcg_stream(uint32_t seed, int n, uint32_t *hashes, float *floats)
{
    Ctx c; memset(&c, 0, sizeof c); c.rndSeed = seed;
    for (int i = 0; i < n; i++) {
        uint32_t s = c.rndSeed;
        uint32_t h = GetPCGHash(&s);
        if (hashes) hashes[i] = h;
        float f = GetRandomFloat01(&c);
        if (floats) floats[i] = f;
    }
}
void pto_sincos(const float *x, int n, float *s, float *c) { for (int i = 0; i < n; i++) g_sincos(x[i], &s[i], &c[i]); }
void pto_exp(const float *x, int n, float *y) { for (int i = 0; i < n; i++) y[i] = g_exp(x[i]); }

/* rays: n x 6 floats (origin, direction); out: n x 4 floats (hit, t1, t2, 0) */
void pto_ray_sphere(const float *rays, int n, const void *sphere80, float *out)
{
    for (int i = 0; i < n; i++) {
        Ray r = { v3(rays[6 * i], rays[6 * i + 1], rays[6 * i + 2]), v3(rays[6 * i + 3], rays[6 * i + 4], rays[6 * i + 5]) };
        float t1, t2;
        int h = RaySphereIntersect(r, (const Sphere *)sphere80, &t1, &t2);
        out[4 * i] = (float)h; out[4 * i + 1] = t1; out[4 * i + 2] = t2; out[4 * i + 3] = 0.0f;
    }
}
void pto_ray_cuboid(const float *rays, int n, const void *cuboid96, float *out)
{
    for (int i = 0; i < n; i++) {
        Ray r = { v3(rays[6 * i], rays[6 * i + 1], rays[6 * i + 2]), v3(rays[6 * i + 3], rays[6 * i + 4], rays[6 * i + 5]) };
        float t1, t2;
        int h = RayCuboidIntersect(r, (const Cuboid *)cuboid96, &t1, &t2);
        out[4 * i] = (float)h; out[4 * i + 1] = t1; out[4 * i + 2] = t2; out[4 * i + 3] = 0.0f;
    }
}
/* Closest-hit fold over a scene. out: n x 12 floats: hit, T, fromInside, primitive-agnostic pad, pos(3), normal(3), albedo.x, emissiv.x */
void pto_ray_trace(const float *rays, int n, const void *objects_ubo, int max_spheres, float n_spheres, float n_cuboids, float *out)
{
    pto_params p; memset(&p, 0, sizeof p);
    p.max_spheres = max_spheres; p.n_spheres = n_spheres; p.n_cuboids = n_cuboids;
    Ctx c; memset(&c, 0, sizeof c); c.p = &p; c.objects = (const uint8_t *)objects_ubo;
    for (int i = 0; i < n; i++) {
        Ray r = { v3(rays[6 * i], rays[6 * i + 1], rays[6 * i + 2]), v3(rays[6 * i + 3], rays[6 * i + 4], rays[6 * i + 5]) };
        HitInfo h; memset(&h, 0, sizeof h);
        int hit = RayTrace(&c, r, &h);
        float *o = out + 12 * i;
        o[0] = (float)hit; o[1] = h.T; o[2] = hit ? (float)h.FromInside : 0.0f; o[3] = 0.0f;
        if (hit) {
            o[4] = h.NearHitPos.x; o[5] = h.NearHitPos.y; o[6] = h.NearHitPos.z;
            o[7] = h.Normal.x; o[8] = h.Normal.y; o[9] = h.Normal.z;
            o[10] = h.Mat.Albedo.x; o[11] = h.Mat.Emissiv.x;
        } else {
            for (int k = 4; k < 12; k++) o[k] = 0.0f;
        }
    }
}
/* texture(samplerCube, dir).rgb for n directions (n x 3 in, n x 3 out) */
void pto_texture_cube(const float *env, int env_size, const float *dirs, int n, float *out)
{
    pto_params p; memset(&p, 0, sizeof p); p.env_size = env_size;
    Ctx c; memset(&c, 0, sizeof c); c.p = &p; c.env = env;
    for (int i = 0; i < n; i++) {
        vec3 t = TextureCube(&c, v3(dirs[3 * i], dirs[3 * i + 1], dirs[3 * i + 2]));
        out[3 * i] = t.x; out[3 * i + 1] = t.y; out[3 * i + 2] = t.z;
    }
}

/* ------------------------------------------------------------------------------------------
 * SURVEY §8f N1 — the post-process pass that follows the path tracer (MainWindow.cs:51):
 * /root/reference/OpenTK-PathTracer/res/shaders/PostProcessing/fragment.glsl (cited pp:LINE), rendered into an RGBA8 target
 * (ScreenEffect.cs:20-22).  Sampler1 is never bound by the host (ScreenEffect.Render gets one texture), so it adds 0.
 * RGBA8 store: clamp to [0,1], scale by 255, round to nearest even.
 * ------------------------------------------------------------------------------------------ */
static float ACESFilm1(float x) /* pp:36-44 */
{
    const float a = 2.51f, b = 0.03f, c = 2.43f, d = 0.59f, e = 0.14f;
    float v = g_div(x * (a * x + b), x * (c * x + d) + e);
    return g_min(g_max(v, 0.0f), 1.0f);
}
static float LinearToInverseGamma1(float rgb, float gamma) /* pp:28-32 */
{
    float sel = rgb < 0.0031308f ? 1.0f : 0.0f;
    return g_mix(g_pow(rgb, g_div(1.0f, gamma)) * 1.055f - 0.055f, rgb * 12.92f, sel);
}
static uint8_t unorm8(float f)
{
    if (g_isnan(f)) return 0;
    f = g_min(g_max(f, 0.0f), 1.0f);
    return (uint8_t)__builtin_rintf(f * 255.0f);
}
void pto_tonemap(const float *rgba32f, int n_pixels, uint8_t *rgba8)
{
    for (int i = 0; i < n_pixels; i++) {
        for (int c = 0; c < 3; c++) {
            float color = rgba32f[4 * i + c] + 0.0f;          /* pp:19-20 */
            color = ACESFilm1(color);                          /* pp:23 */
            color = LinearToInverseGamma1(color, 2.4f);        /* pp:24 */
            rgba8[4 * i + c] = unorm8(color);
        }
        rgba8[4 * i + 3] = 255;                                /* pp:25 */
    }
}

/* SURVEY §8f N2 — the alternate environment: six sRGB8 PNG faces uploaded as Srgb8Alpha8 (Helper.cs:18-50,
 * MainWindow.cs:177-187).  The texture unit decodes sRGB to linear before filtering (GL 4.5 §8.24): c/12.92 below 0.04045,
 * ((c+0.055)/1.055)^2.4 above; alpha stays linear. */
void pto_srgb8_to_linear(const uint8_t *rgba8, int n_texels, float *rgba32f)
{
    for (int i = 0; i < n_texels; i++) {
        for (int c = 0; c < 3; c++) {
            float cs = g_div((float)rgba8[4 * i + c], 255.0f);
            rgba32f[4 * i + c] = cs <= 0.04045f ? g_div(cs, 12.92f) : g_pow(g_div(cs + 0.055f, 1.055f), 2.4f);
        }
        rgba32f[4 * i + 3] = g_div((float)rgba8[4 * i + 3], 255.0f);
    }
}
void pto_log(const float *x, int n, float *y) { for (int i = 0; i < n; i++) y[i] = g_log(x[i]); }
