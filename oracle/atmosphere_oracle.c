/*
 * oracle/atmosphere_oracle.c — TEST INFRASTRUCTURE ONLY.  CPU restatement of the reference's
 * atmosphere cubemap producer, /root/reference/OpenTK-PathTracer/res/shaders/AtmosphericScattering/compute.glsl
 * (cited as atmos:LINE) driven as AtmosphericScatterer.cs:63-113 drives it.  Arithmetic per glsl_model.h.
 * PARITY PIN: bit-exact against the same shader compiled for the CPU from /root/reference (oracle/build_ref.py,
 * tests/test_reference_pin.py::test_atmosphere_shader, tests/golden/ref_atmosphere.npz).
 */
#include "glsl_model.h"
#ifdef _OPENMP
#include <omp.h>
#endif

#define PI 3.14159265f /* atmos:6 */

typedef struct { float x, y; } vec2;

/* atmos:58-71 */
static vec2 Rsi(vec3 r0, vec3 rd, float sr)
{
    float a = v_dot(rd, rd);
    float b = 2.0f * v_dot(rd, r0);
    float c = v_dot(r0, r0) - (sr * sr);
    float d = (b * b) - 4.0f * a * c;
    vec2 r;
    if (d < 0.0f) { r.x = 1e5f; r.y = -1e5f; return r; }
    r.x = g_div(-b - g_sqrt(d), 2.0f * a);
    r.y = g_div(-b + g_sqrt(d), 2.0f * a);
    return r;
}

/* atmos:73-159 */
static vec3 Atmosphere(vec3 r, vec3 r0, vec3 pSun, float iSun, float rPlanet, float rAtmos, vec3 kRlh, float kMie,
                       float shRlh, float shMie, float g, int iSteps, int jSteps)
{
    pSun = v_normalize(pSun);
    r = v_normalize(r);

    vec2 p = Rsi(r0, r, rAtmos);
    if (p.x > p.y) return v3(0.0f, 0.0f, 0.0f);
    p.y = g_min(p.y, Rsi(r0, r, rPlanet).x);
    float iStepSize = g_div(p.y - p.x, (float)iSteps);

    float iTime = 0.0f;
    vec3 totalRlh = v3(0.0f, 0.0f, 0.0f);
    vec3 totalMie = v3(0.0f, 0.0f, 0.0f);
    float iOdRlh = 0.0f;
    float iOdMie = 0.0f;

    float mu = v_dot(r, pSun);
    float mumu = mu * mu;
    float gg = g * g;
    float pRlh = g_div(3.0f, 16.0f * PI) * (1.0f + mumu);
    float pMie = g_div(g_div(3.0f, 8.0f * PI) * ((1.0f - gg) * (mumu + 1.0f)),
                       g_pow15(1.0f + gg - 2.0f * mu * g) * (2.0f + gg));
    float ishRlh = g_rcp(shRlh), ishMie = g_rcp(shMie);

    for (int i = 0; i < iSteps; i++) {
        vec3 iPos = v_add(r0, v_scale(r, iTime + iStepSize * 0.5f));
        float iHeight = v_length(iPos) - rPlanet;
        float odStepRlh = g_exp(-iHeight * ishRlh) * iStepSize;
        float odStepMie = g_exp(-iHeight * ishMie) * iStepSize;
        iOdRlh += odStepRlh;
        iOdMie += odStepMie;

        float jStepSize = g_div(Rsi(iPos, pSun, rAtmos).y, (float)jSteps);
        float jTime = 0.0f;
        float jOdRlh = 0.0f;
        float jOdMie = 0.0f;
        for (int j = 0; j < jSteps; j++) {
            vec3 jPos = v_add(iPos, v_scale(pSun, jTime + jStepSize * 0.5f));
            float jHeight = v_length(jPos) - rPlanet;
            jOdRlh += g_exp(-jHeight * ishRlh) * jStepSize;
            jOdMie += g_exp(-jHeight * ishMie) * jStepSize;
            jTime += jStepSize;
        }
        float m = kMie * (iOdMie + jOdMie);
        float rl = iOdRlh + jOdRlh;
        vec3 attn = v3(g_exp(-(m + kRlh.x * rl)), g_exp(-(m + kRlh.y * rl)), g_exp(-(m + kRlh.z * rl)));
        totalRlh = v_add(totalRlh, v_scale(attn, odStepRlh));
        totalMie = v_add(totalMie, v_scale(attn, odStepMie));
        iTime += iStepSize;
    }
    vec3 res = v3(iSun * (pRlh * kRlh.x * totalRlh.x + pMie * kMie * totalMie.x),
                  iSun * (pRlh * kRlh.y * totalRlh.y + pMie * kMie * totalMie.y),
                  iSun * (pRlh * kRlh.z * totalRlh.z + pMie * kMie * totalMie.z));
    return res;
}

/* atmos:166-171 */
static vec3 GetWorldSpaceRay(const float *inverseProj, const float *inverseView, float nx, float ny)
{
    float ex = m4_row(inverseProj, 0, nx, ny, -1.0f, 0.0f);
    float ey = m4_row(inverseProj, 1, nx, ny, -1.0f, 0.0f);
    return v_normalize(v3(m4_row(inverseView, 0, ex, ey, -1.0f, 0.0f),
                          m4_row(inverseView, 1, ex, ey, -1.0f, 0.0f),
                          m4_row(inverseView, 2, ex, ey, -1.0f, 0.0f)));
}

/* atmos:30-56 over the whole cubemap (AtmosphericScatterer.cs:102-113 dispatch (size/8, size/8, 6)).
 * ubo = AtmosphericDataUBO bytes: InvProjection @0, InvView[6] @64 (atmos:12-16); out = 6*size*size*4 floats. */
int pto_atmosphere(int size, const void *ubo, const float *lightPos, float lightIntensity, int iSteps, int jSteps,
                   float *out, int n_threads)
{
    if (size <= 0 || !ubo || !lightPos || !out) return -1;
    const float *U = (const float *)ubo;
#ifdef _OPENMP
    if (n_threads <= 0) n_threads = omp_get_max_threads();
#else
    n_threads = 1;
#endif
    float isz = g_rcp((float)size);
#pragma omp parallel for num_threads(n_threads) schedule(dynamic, 8) collapse(2)
    for (int f = 0; f < 6; f++)
        for (int y = 0; y < size; y++)
            for (int x = 0; x < size; x++) {
                float nx = (float)x * isz * 2.0f - 1.0f; /* atmos:37 — corner sampled, no +0.5 */
                float ny = (float)y * isz * 2.0f - 1.0f;
                vec3 dir = GetWorldSpaceRay(U, U + 16 + 16 * f, nx, ny);
                vec3 col = Atmosphere(dir, v3(0.0f, 6376e3f, 0.0f), v3(lightPos[0], lightPos[1], lightPos[2]),
                                      lightIntensity, 6371e3f, 6471e3f, v3(5.5e-6f, 13.0e-6f, 22.4e-6f), 21e-6f,
                                      8e3f, 1.2e3f, 0.758f, iSteps, jSteps);
                float *o = out + (((size_t)f * size + y) * size + x) * 4;
                o[0] = col.x; o[1] = col.y; o[2] = col.z; o[3] = 1.0f;
            }
    return 0;
}
