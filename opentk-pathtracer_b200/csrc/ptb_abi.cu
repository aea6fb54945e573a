// ptb_abi.cu — the C ABI of libptb200.so (see include/ptb200.h for the reference member each entry replaces).
// Host side: owns the device copies of the two UBOs, the padded environment cubemap, the accumulation image,
// the packed scene block, and launches the kernels of ptb_kernels.cuh on the context's stream.
#include "../../include/ptb200.h"
#include "ptb_kernels.cuh"
#include "ptb_fast.h"

#include <algorithm>
#include <cmath>
#include <cstdarg>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <new>
#include <vector>

using namespace ptb;

// CUDA-GL interop of the runtime (cuda_gl_interop.h:204).  Declared here because that header includes <GL/gl.h>, which a
// headless build box does not have; GLuint / GLenum are unsigned int, GL_TEXTURE_2D is 0x0DE1.
extern "C" cudaError_t CUDARTAPI cudaGraphicsGLRegisterImage(struct cudaGraphicsResource** resource, unsigned int image, unsigned int target, unsigned int flags);

namespace {

thread_local char g_err[512] = "";

int fail(int code, const char* fmt, ...)
{
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof g_err, fmt, ap);
    va_end(ap);
    return code;
}

#define CU(call)                                                                                                   \
    do {                                                                                                           \
        cudaError_t e_ = (call);                                                                                   \
        if (e_ != cudaSuccess) return fail(PTB_E_CUDA, "%s failed: %s (%s:%d)", #call, cudaGetErrorString(e_), __FILE__, __LINE__); \
    } while (0)

constexpr int kBasicBytes = 144;   // MainWindow.cs:196
constexpr int kMaxSmem = 227 * 1024;
constexpr int kMaxOverlap = 4;
// Scratch sets of the batched path.  Three, not two: with two, the trace of batch k+2 would wait for the blend of batch k, which
// itself can only start once batch k+1's persistent grid lets go of its CTA slots — trace and blend would take turns instead of
// overlapping.  With three, batch k+2 only needs the blend of batch k-1.
constexpr int kBatchSets = 3;

} // namespace

struct SceneExtent { double lo[3], hi[3], maxabs; };     // bounds of the finite primitive boxes (BVH error-margin sizing)

struct ptb_ctx {
    int device = 0;
    int sm_count = 148;
    int width = 0, height = 0;
    int max_spheres = 0, max_cuboids = 0;
    int n_spheres = 0, n_cuboids = 0;
    int ray_depth = 13, spp = 1;
    float focal_length = 20.0f, aperture_diameter = 0.14f;
    int frame = 0;
    int kernel = PTB_KERNEL_MEGA;
    int rank = 0, world = 1, stripe_rows = 8, local_rows = 0;
    // host mirrors
    unsigned char basic[kBasicBytes] = {};
    std::vector<unsigned char> objects;    // GameObjectsUBO bytes
    bool scene_dirty = true;
    // device
    unsigned char* d_objects = nullptr;
    float4* d_block = nullptr;
    size_t block_capacity = 0;
    int off_aux = 0, off_cmin = 0, off_cmax = 0, off_mat = 0, block_bytes = 0;
    // BVH for large scenes (conservative culling in front of the exact tests)
    int bvh_threshold = 96;          // primitives; below this the brute-force fold wins
    int stage_bytes = 0, off_nodes = 0, off_pidx = 0, n_nodes = 0, n_unbounded = 0;
    float bvh_tau = 0.0f, bvh_D = 0.0f;
    std::vector<float4> bvh_nodes;
    std::vector<int> bvh_pidx;
    SceneExtent bvh_extent = {};
    // ray-classification table for small scenes (<= 64 primitives): see trace_rct / rct_build_kernel
    int rct_mode = 1;                // 0 off, 1 on when the scene qualifies
    int rct_cells = 18, rct_G = 16;  // cells along the longest scene axis, direction buckets per cube-face axis (C2 batched, fast: 13,12 0.1925 / 16,16 0.1833 / 18,16 0.1804 / 20,16 0.1782 ms per frame)
    unsigned long long* d_rct = nullptr;
    size_t rct_capacity = 0;
    bool rct_on = false;
    float rct_lo[3] = {}, rct_inv[3] = {};
    int rct_n[3] = {};
    unsigned rct_sm0 = 0, rct_sm1 = 0;
    std::vector<unsigned char> rct_geometry;   // the geometry bytes the table was built from (material edits do not rebuild it)
    // uniform grid for large scenes (trace_grid): the default above bvh_threshold; the BVH stays as the fallback / alternative
    int large_mode = 1;              // 0 = BVH, 1 = grid
    float grid_density = 3.0f;       // target cells per binned primitive (ptb_set_grid_density; C3: 2 / 3 / 4 / 6 -> 1247 / 1271 / 1254 / 1115 Msamples/s)
    bool grid_on = false;
    int grid_n[3] = {}, off_gcell = 0, off_gsph = 0, off_gitem = 0;
    std::vector<unsigned char> grid_cell_spheres;
    float grid_lo[3] = {}, grid_hi[3] = {}, grid_cell[3] = {}, grid_inv[3] = {};
    std::vector<unsigned short> grid_cell_start, grid_items;
    size_t env_faces_bytes = 0, env_padded_bytes = 0;
    float4* d_env_faces = nullptr;   // unpadded 6*N*N
    float4* d_env = nullptr;         // padded 6*(N+2)^2
    int env_size = 0;
    float4* d_image = nullptr;
    size_t image_bytes = 0;
    unsigned int* d_counters = nullptr;
    unsigned long long* d_stats = nullptr;
    bool stats_on = false;
    cudaStream_t stream = nullptr;
    bool own_stream = false;
    // pipelined read-back: snapshot on the render stream, D2H on the copy stream, two staging buffers in flight
    cudaStream_t copy_stream = nullptr;
    float4* d_stage[2] = {nullptr, nullptr};
    size_t readback_bytes = 0;
    cudaEvent_t ev_snap[2] = {nullptr, nullptr}, ev_copied[2] = {nullptr, nullptr};
    int stage_next = 0;
    // pipelined frames: `overlap` trace streams, each with its own scratch image and work counters, and one blend stream.
    // Frame f is traced on stream f % overlap into scratch[f % overlap]; the blend stream folds the estimates into the image
    // in frame order; the user-visible stream waits for each blend.  overlap <= 1: classic in-place accumulation.
    int overlap = 2;
    cudaStream_t trace_stream[kMaxOverlap] = {};
    cudaStream_t blend_stream = nullptr;
    float4* d_scratch[kMaxOverlap] = {};
    unsigned int* d_slot_counters[kMaxOverlap] = {};
    cudaEvent_t ev_trace_done[kMaxOverlap] = {}, ev_blend_done[kMaxOverlap] = {};
    bool blend_recorded[kMaxOverlap] = {};
    cudaEvent_t ev_inputs = nullptr;            // scene / environment / image edits enqueued on `stream`
    unsigned inputs_version = 1, seen_version[kMaxOverlap + 1] = {};
    size_t scratch_bytes = 0;
    unsigned long long launch_seq = 0;
    // fused multi-GPU exchange (ptb_exchange_*): rank 0 owns [flags | slots x full image]; other ranks map it through CUDA IPC
    // Roots: 1 = every frame is assembled on rank 0 (xch_block = rank 0's block, mapped by the others); world = rotating roots,
    // frame q is assembled on rank q % world (every rank owns a block and maps all the others': xch_peer[]).
    bool xch_on = false, xch_mapped = false, xch_rgb = false;
    int xch_slots = 0, xch_roots = 1;
    unsigned char* xch_block = nullptr;          // this rank's own block (roots == world, or rank 0), or rank 0's mapping (roots == 1)
    unsigned char* xch_peer[16] = {};            // rotating roots: the block of every rank (own block for self)
    bool xch_peer_mapped[16] = {};
    size_t xch_image_bytes = 0;
    unsigned long long xch_seq = 0, xch_acquired = 0, xch_released = 0;
    unsigned int* d_xch_blocks = nullptr;
    cudaEvent_t ev0 = nullptr, ev1 = nullptr;
    bool timed = false;
    int launches = 0;
    int mega_smem_set = -1;
    int mega_grid = 0;
    int grid_divisor = 1;            // ptb_set_grid_divisor: launch 1/d of the resident CTA slots per frame (experiments)
    // frame batching (ptb_set_batch): up to `batch` consecutive frames of one ptb_render_frames call are traced by ONE
    // megakernel launch into a set of per-frame scratch images; kBatchSets sets rotate so that the blend of batch k runs
    // beside the trace of batch k+1
    cudaGraphicsResource* gl_result = nullptr;   // PathTracer.Result registered through CUDA-GL interop (ptb_register_gl_texture)
    unsigned* d_done = nullptr;      // [kBatchSets] "batch traced" flags (written by the trace's last CTA) + [kBatchSets] a sticky error word
    int defer_finish = 1;            // SPP == 1: finished pixels go through the warp's completion queue (PTB_DEFER=0 switches it off for A/B runs)
    int blend_ctas = 64;             // grid of the batch blend kernels (a background kernel beside the next batch's trace)
    int batch = 16;
    float4* d_batch_scratch[kBatchSets] = {};
    size_t batch_scratch_frames = 0, batch_scratch_stride = 0;     // frames per set / float4 elements per frame
    unsigned int* d_batch_counters[kBatchSets] = {};
    cudaEvent_t ev_batch_blend[kBatchSets] = {};
    bool batch_blend_recorded[kBatchSets] = {};
    unsigned long long batch_seq = 0;
    bool mega_ring = true, mega_defer = true;
    int mega_fold_set = -1;
    // ptb_set_kernel_timing: every megakernel launch brackets itself on the device (first CTA start, last CTA end, %globaltimer)
    bool kt_on = false;
    unsigned long long* d_kt = nullptr;        // 8 x {min CTA start, max CTA end}
    unsigned long long* h_kt = nullptr;        // pinned: [0..15] results, [16..17] the initial pair {~0, 0}
    cudaEvent_t kt_b[8] = {};             // recorded behind each timed launch's read-back: its results are on the host
    int kt_frames[8] = {};
    bool kt_used[8] = {};
    unsigned kt_next = 0;
    double kt_ms = 0.0;
    long long kt_frames_total = 0, kt_launches = 0;
    int precision = PTB_PRECISION_EXACT, mega_precision_set = -1;   // ptb_set_precision: which translation unit's megakernel runs
};

namespace {

int sync_all(ptb_ctx* c)
{
    CU(cudaSetDevice(c->device));
    CU(cudaStreamSynchronize(c->stream));      // handle 0 (the legacy default stream) is a valid stream too
    for (int i = 0; i < kMaxOverlap; ++i) if (c->trace_stream[i]) CU(cudaStreamSynchronize(c->trace_stream[i]));
    if (c->blend_stream) CU(cudaStreamSynchronize(c->blend_stream));
    if (c->copy_stream) CU(cudaStreamSynchronize(c->copy_stream));
    return PTB_OK;
}

// Anything enqueued on the user-visible stream that the trace / blend streams must see (scene repack, environment, image writes).
int mark_inputs(ptb_ctx* c)
{
    if (!c->ev_inputs) CU(cudaEventCreateWithFlags(&c->ev_inputs, cudaEventDisableTiming));
    CU(cudaEventRecord(c->ev_inputs, c->stream));
    c->inputs_version++;
    return PTB_OK;
}

int ensure_pipeline(ptb_ctx* c)
{
    const size_t bytes = c->image_bytes;
    for (int i = 0; i < c->overlap; ++i) {
        if (!c->trace_stream[i]) {
            CU(cudaStreamCreateWithFlags(&c->trace_stream[i], cudaStreamNonBlocking));
            CU(cudaEventCreateWithFlags(&c->ev_trace_done[i], cudaEventDisableTiming));
            CU(cudaEventCreateWithFlags(&c->ev_blend_done[i], cudaEventDisableTiming));
            CU(cudaMalloc(&c->d_slot_counters[i], 2 * sizeof(unsigned int)));
            CU(cudaMemsetAsync(c->d_slot_counters[i], 0, 2 * sizeof(unsigned int), c->stream));
            { const int rc = mark_inputs(c); if (rc != PTB_OK) return rc; }
        }
    }
    if (!c->blend_stream) {
        int lo = 0, hi = 0;
        CU(cudaDeviceGetStreamPriorityRange(&lo, &hi));
        CU(cudaStreamCreateWithPriority(&c->blend_stream, cudaStreamNonBlocking, hi));   // small kernel, must not queue behind a persistent grid
    }
    if (c->scratch_bytes != bytes) {
        { const int rc = sync_all(c); if (rc != PTB_OK) return rc; }
        for (int i = 0; i < kMaxOverlap; ++i) { if (c->d_scratch[i]) CU(cudaFree(c->d_scratch[i])); c->d_scratch[i] = nullptr; }
        for (int i = 0; i < c->overlap; ++i) CU(cudaMalloc(&c->d_scratch[i], bytes));
        c->scratch_bytes = bytes;
        for (int i = 0; i < kMaxOverlap; ++i) c->blend_recorded[i] = false;
    } else {
        for (int i = 0; i < c->overlap; ++i) if (!c->d_scratch[i]) CU(cudaMalloc(&c->d_scratch[i], bytes));
    }
    return PTB_OK;
}

int compute_local_rows(int height, int rank, int world, int stripe_rows)
{
    int rows = 0;
    const int stripes = (height + stripe_rows - 1) / stripe_rows;
    for (int s = rank; s < stripes; s += world) rows += (s * stripe_rows + stripe_rows <= height) ? stripe_rows : height - s * stripe_rows;
    return rows;
}

int alloc_image(ptb_ctx* c)
{
    { const int rc = sync_all(c); if (rc != PTB_OK) return rc; }
    c->local_rows = compute_local_rows(c->height, c->rank, c->world, c->stripe_rows);
    // capacity = the largest local image any rank holds, so gathered buffers are uniform
    const int max_rows = compute_local_rows(c->height, 0, c->world, c->stripe_rows);
    const size_t bytes = (size_t)(max_rows > 0 ? max_rows : 1) * c->width * sizeof(float4);
    if (c->d_image) { CU(cudaFree(c->d_image)); c->d_image = nullptr; }
    CU(cudaMalloc(&c->d_image, bytes));
    CU(cudaMemsetAsync(c->d_image, 0, bytes, c->stream));
    c->image_bytes = bytes;
    return mark_inputs(c);
}

// ---- conservative BVH over the primitives' inflated boxes (host build; see trace_bvh and DESIGN.md for the error bounds)
struct Box { double lo[3], hi[3]; };
#ifndef PTB_BVH_BIG
#define PTB_BVH_BIG 0.25     // a primitive wider than this fraction of the scene on any axis is tested for every ray instead
#endif
#ifndef PTB_BVH_LEAF
#define PTB_BVH_LEAF 4
#endif

SceneExtent scene_extent(const std::vector<Box>& boxes)
{
    SceneExtent e = {{1e300, 1e300, 1e300}, {-1e300, -1e300, -1e300}, 0.0};
    for (const Box& b : boxes)
        for (int k = 0; k < 3; ++k) {
            e.lo[k] = std::min(e.lo[k], b.lo[k]); e.hi[k] = std::max(e.hi[k], b.hi[k]);
            e.maxabs = std::max(e.maxabs, std::max(std::fabs(b.lo[k]), std::fabs(b.hi[k])));
        }
    return e;
}

// D bounds every coordinate that enters a test and every origin-to-primitive distance: scene boxes + camera + lens,
// rounded up to a power of two so that small camera moves do not change it.
float required_extent(ptb_ctx* c, SceneExtent e)
{
    float cam[3];
    memcpy(cam, c->basic + 128, 12);
    const double lens = std::fabs((double)c->aperture_diameter) * 0.5 + 1.0;
    for (int k = 0; k < 3; ++k) {
        if (!std::isfinite(cam[k])) continue;
        e.lo[k] = std::min(e.lo[k], cam[k] - lens); e.hi[k] = std::max(e.hi[k], cam[k] + lens);
        e.maxabs = std::max(e.maxabs, std::fabs((double)cam[k]) + lens);
    }
    double diag2 = 0.0;
    for (int k = 0; k < 3; ++k) diag2 += (e.hi[k] - e.lo[k]) * (e.hi[k] - e.lo[k]);
    const double D = 1.01 * std::max(std::sqrt(diag2), e.maxabs) + 1.0;
    float Dq = 1.0f;
    while (Dq < D && Dq < 1e30f) Dq *= 2.0f;
    return Dq;
}

void raw_boxes(ptb_ctx* c, double E, double m, std::vector<Box>& boxes, std::vector<char>& bounded)
{
    const int nS = c->n_spheres, nC = c->n_cuboids;
    boxes.assign((size_t)nS + nC, Box());
    bounded.assign((size_t)nS + nC, 1);
    for (int i = 0; i < nS + nC; ++i) {
        Box& b = boxes[i];
        if (i < nS) {
            float g[4];
            memcpy(g, c->objects.data() + (size_t)i * kSphereStride, 16);
            const double r = std::sqrt((double)g[3] * g[3] + E) + m;
            for (int k = 0; k < 3; ++k) { b.lo[k] = g[k] - r; b.hi[k] = g[k] + r; }
        } else {
            float lo[4], hi[4];
            const unsigned char* p = c->objects.data() + (size_t)c->max_spheres * kSphereStride + (size_t)(i - nS) * kCuboidStride;
            memcpy(lo, p, 16); memcpy(hi, p + 16, 16);
            for (int k = 0; k < 3; ++k) { b.lo[k] = std::min(lo[k], hi[k]) - m; b.hi[k] = std::max(lo[k], hi[k]) + m; }
        }
        double sum = 0.0;
        for (int k = 0; k < 3; ++k) sum += b.lo[k] + b.hi[k];
        if (!std::isfinite(sum)) bounded[i] = 0;      // NaN / Inf geometry: always tested, never culled
    }
}

void build_bvh(ptb_ctx* c)
{
    c->bvh_nodes.clear(); c->bvh_pidx.clear(); c->n_nodes = 0; c->n_unbounded = 0; c->bvh_tau = 0.0f; c->bvh_D = 0.0f;
    const int n = c->n_spheres + c->n_cuboids;
    if (n < c->bvh_threshold) return;
    std::vector<Box> boxes; std::vector<char> bounded;
    raw_boxes(c, 0.0, 0.0, boxes, bounded);
    std::vector<Box> finite;
    for (int i = 0; i < n; ++i) if (bounded[i]) finite.push_back(boxes[i]);
    c->bvh_extent = scene_extent(finite);
    const float D = required_extent(c, c->bvh_extent);
    const double E = 4e-6 * (double)D * D, m = 1e-5 * (double)D + 1e-6;
    raw_boxes(c, E, m, boxes, bounded);
    // primitives that span a large part of the scene (the room's walls and floor) would drag every ancestor box up to scene
    // size: they go to the always-tested list together with the non-finite ones
    double slo[3] = {1e300, 1e300, 1e300}, shi[3] = {-1e300, -1e300, -1e300};
    for (int i = 0; i < n; ++i) if (bounded[i]) for (int k = 0; k < 3; ++k) { slo[k] = std::min(slo[k], boxes[i].lo[k]); shi[k] = std::max(shi[k], boxes[i].hi[k]); }
    std::vector<int> order;
    for (int i = 0; i < n; ++i) {
        bool big = !bounded[i];
        if (!big) for (int k = 0; k < 3; ++k) if (boxes[i].hi[k] - boxes[i].lo[k] > PTB_BVH_BIG * (shi[k] - slo[k])) big = true;
        if (big && (int)c->bvh_pidx.size() < 64) c->bvh_pidx.push_back(i); else if (bounded[i]) order.push_back(i); else c->bvh_pidx.push_back(i);
    }
    c->n_unbounded = (int)c->bvh_pidx.size();
    c->bvh_D = D;
    c->bvh_tau = 1e-4f * D;
    if (order.empty()) return;
    struct Task { int node, begin, end, depth; };      // the device-side traversal stack holds 32 entries: depth is bounded below
    std::vector<Task> todo;
    auto set_box = [&](int node, int begin, int end) {
        double lo[3] = {1e300, 1e300, 1e300}, hi[3] = {-1e300, -1e300, -1e300};
        for (int q = begin; q < end; ++q) for (int k = 0; k < 3; ++k) { lo[k] = std::min(lo[k], boxes[order[q]].lo[k]); hi[k] = std::max(hi[k], boxes[order[q]].hi[k]); }
        float4 A, B;   // round outwards
        A.x = std::nextafter((float)lo[0], -INFINITY); A.y = std::nextafter((float)lo[1], -INFINITY); A.z = std::nextafter((float)lo[2], -INFINITY);
        B.x = std::nextafter((float)hi[0], INFINITY); B.y = std::nextafter((float)hi[1], INFINITY); B.z = std::nextafter((float)hi[2], INFINITY);
        A.w = 0.0f; B.w = 0.0f;
        c->bvh_nodes[2 * node] = A; c->bvh_nodes[2 * node + 1] = B;
    };
    c->bvh_nodes.resize(2);
    todo.push_back({0, 0, (int)order.size(), 0});
    while (!todo.empty()) {
        const Task t = todo.back(); todo.pop_back();
        set_box(t.node, t.begin, t.end);
        const int cnt = t.end - t.begin;
        if (cnt <= PTB_BVH_LEAF) {
            std::sort(order.begin() + t.begin, order.begin() + t.end);
            const int first = (int)c->bvh_pidx.size();
            for (int q = t.begin; q < t.end; ++q) c->bvh_pidx.push_back(order[q]);
            memcpy(&c->bvh_nodes[2 * t.node].w, &first, 4);
            memcpy(&c->bvh_nodes[2 * t.node + 1].w, &cnt, 4);
            continue;
        }
        // binned surface-area heuristic: 16 bins per axis over the centroid range, cost = area(L) * n(L) + area(R) * n(R)
        double clo[3] = {1e300, 1e300, 1e300}, chi[3] = {-1e300, -1e300, -1e300};
        for (int q = t.begin; q < t.end; ++q) for (int k = 0; k < 3; ++k) { const double ce = boxes[order[q]].lo[k] + boxes[order[q]].hi[k]; clo[k] = std::min(clo[k], ce); chi[k] = std::max(chi[k], ce); }
        auto half_area = [](const double* lo, const double* hi) { const double dx = hi[0] - lo[0], dy = hi[1] - lo[1], dz = hi[2] - lo[2]; return dx * dy + dy * dz + dz * dx; };
        constexpr int kBins = 16;
        int best_axis = -1, best_bin = -1;
        double best_cost = 1e300;
        // beyond depth 14 only balanced median splits are used, so the total depth stays under 14 + log2(n) < 32
        for (int axis = 0; axis < 3 && t.depth < 14; ++axis) {
            const double ext = chi[axis] - clo[axis];
            if (!(ext > 0.0)) continue;
            int cntb[kBins] = {};
            double blo[kBins][3], bhi[kBins][3];
            for (int b = 0; b < kBins; ++b) for (int k = 0; k < 3; ++k) { blo[b][k] = 1e300; bhi[b][k] = -1e300; }
            for (int q = t.begin; q < t.end; ++q) {
                const Box& bx = boxes[order[q]];
                int b = (int)(((bx.lo[axis] + bx.hi[axis]) - clo[axis]) / ext * kBins);
                b = std::min(std::max(b, 0), kBins - 1);
                cntb[b]++;
                for (int k = 0; k < 3; ++k) { blo[b][k] = std::min(blo[b][k], bx.lo[k]); bhi[b][k] = std::max(bhi[b][k], bx.hi[k]); }
            }
            double rlo[kBins][3], rhi[kBins][3]; int rcnt[kBins];
            double alo[3] = {1e300, 1e300, 1e300}, ahi[3] = {-1e300, -1e300, -1e300}; int acc = 0;
            for (int b = kBins - 1; b >= 0; --b) {
                for (int k = 0; k < 3; ++k) { alo[k] = std::min(alo[k], blo[b][k]); ahi[k] = std::max(ahi[k], bhi[b][k]); }
                acc += cntb[b];
                for (int k = 0; k < 3; ++k) { rlo[b][k] = alo[k]; rhi[b][k] = ahi[k]; }
                rcnt[b] = acc;
            }
            double llo[3] = {1e300, 1e300, 1e300}, lhi[3] = {-1e300, -1e300, -1e300}; int lc = 0;
            for (int b = 0; b < kBins - 1; ++b) {
                for (int k = 0; k < 3; ++k) { llo[k] = std::min(llo[k], blo[b][k]); lhi[k] = std::max(lhi[k], bhi[b][k]); }
                lc += cntb[b];
                if (lc == 0 || rcnt[b + 1] == 0) continue;
                const double cost = half_area(llo, lhi) * lc + half_area(rlo[b + 1], rhi[b + 1]) * rcnt[b + 1];
                if (cost < best_cost) { best_cost = cost; best_axis = axis; best_bin = b; }
            }
        }
        int mid;
        if (best_axis >= 0) {
            const int axis = best_axis;
            const double ext = chi[axis] - clo[axis];
            auto it = std::partition(order.begin() + t.begin, order.begin() + t.end, [&](int a) {
                int b = (int)(((boxes[a].lo[axis] + boxes[a].hi[axis]) - clo[axis]) / ext * kBins);
                b = std::min(std::max(b, 0), kBins - 1);
                return b <= best_bin;
            });
            mid = (int)(it - order.begin());
        } else {
            mid = t.begin;      // all centroids coincide
        }
        if (mid == t.begin || mid == t.end) {      // degenerate: fall back to an object-median split on the widest axis
            int axis = 0;
            if (chi[1] - clo[1] > chi[axis] - clo[axis]) axis = 1;
            if (chi[2] - clo[2] > chi[axis] - clo[axis]) axis = 2;
            mid = t.begin + cnt / 2;
            std::nth_element(order.begin() + t.begin, order.begin() + mid, order.begin() + t.end,
                             [&](int a, int b) { return boxes[a].lo[axis] + boxes[a].hi[axis] < boxes[b].lo[axis] + boxes[b].hi[axis]; });
        }
        const int left = (int)(c->bvh_nodes.size() / 2);
        c->bvh_nodes.resize(c->bvh_nodes.size() + 4);
        const int zero = 0;
        memcpy(&c->bvh_nodes[2 * t.node].w, &left, 4);
        memcpy(&c->bvh_nodes[2 * t.node + 1].w, &zero, 4);
        todo.push_back({left, t.begin, mid, t.depth + 1});
        todo.push_back({left + 1, mid, t.end, t.depth + 1});
    }
    c->n_nodes = (int)(c->bvh_nodes.size() / 2);
}

// Geometry bytes only (sphere centre + radius, cuboid min + max): what the classification table depends on.
void geometry_bytes(const ptb_ctx* c, std::vector<unsigned char>& out)
{
    out.clear();
    const int hdr[6] = {c->n_spheres, c->n_cuboids, c->rct_cells, c->rct_G, 0, 0};
    out.insert(out.end(), (const unsigned char*)hdr, (const unsigned char*)hdr + sizeof hdr);
    for (int i = 0; i < c->n_spheres; ++i) { const unsigned char* p = c->objects.data() + (size_t)i * kSphereStride; out.insert(out.end(), p, p + 16); }
    for (int i = 0; i < c->n_cuboids; ++i) { const unsigned char* p = c->objects.data() + (size_t)c->max_spheres * kSphereStride + (size_t)i * kCuboidStride; out.insert(out.end(), p, p + 32); }
}

bool rct_eligible(const ptb_ctx* c)
{
    const int n = c->n_spheres + c->n_cuboids;
    return c->rct_mode != 0 && n >= 4 && n <= 64 && n < c->bvh_threshold;
}

// Plans and (re)builds the table on c->stream.  Returns PTB_OK with c->rct_on = false when the scene has no finite extent.
int build_rct(ptb_ctx* c)
{
    c->rct_on = false;
    if (!rct_eligible(c)) { c->rct_geometry.clear(); return PTB_OK; }
    std::vector<Box> boxes; std::vector<char> bounded;
    raw_boxes(c, 0.0, 0.0, boxes, bounded);
    std::vector<Box> finite;
    for (size_t i = 0; i < boxes.size(); ++i) if (bounded[i]) finite.push_back(boxes[i]);
    if (finite.empty()) { c->rct_geometry.clear(); return PTB_OK; }
    const SceneExtent e = scene_extent(finite);
    // D bounds every coordinate and every origin-to-primitive distance of a ray that starts inside the grid (rays from
    // outside take the full mask): the same error margins as the BVH, without the camera term
    double diag2 = 0.0, ext[3], longest = 0.0;
    for (int k = 0; k < 3; ++k) { ext[k] = e.hi[k] - e.lo[k]; diag2 += ext[k] * ext[k]; longest = std::max(longest, ext[k]); }
    if (!(longest > 0.0) || !std::isfinite(longest)) { c->rct_geometry.clear(); return PTB_OK; }
    const double D = 1.01 * (std::sqrt(diag2) + e.maxabs) + 1.0;
    const double E = 4e-6 * D * D, m = 1e-5 * D + 1e-6;
    const double pad = 2.0 * m + 1e-3 * longest;
    RctBuild B = {};
    B.G = c->rct_G;
    // the table is capped at 64 MiB (its index stays below 2^23): a scene whose shape asks for more cells gets a coarser grid
    for (int want = std::max(1, c->rct_cells);; want = want * 9 / 10) {
        const double target = longest / std::max(1, want);
        unsigned long long cells = 1;
        for (int k = 0; k < 3; ++k) {
            const double lo = e.lo[k] - pad, hi = e.hi[k] + pad;
            int n = (int)std::lround((hi - lo) / target);
            n = std::min(std::max(n, 1), 64);
            B.n[k] = n; B.lo[k] = lo; B.cell[k] = (hi - lo) / n;
            c->rct_n[k] = n; c->rct_lo[k] = (float)lo; c->rct_inv[k] = (float)(1.0 / B.cell[k]);
            cells *= (unsigned long long)n;
        }
        B.total = cells * 6ull * (unsigned long long)(B.G * B.G);
        if (B.total * sizeof(unsigned long long) <= ((size_t)64 << 20) || want <= 1) break;
    }
    // classification rounds in fp32: (o - lo) * inv is off by < 4 ulp of a cell index <= 64, the grid origin / cell size by one
    // rounding each; u = d_a * rcp(|d_m|) by < 2 ulp.  1e-4 of a cell and 1e-5 in u are far above both.
    B.eps_cell = 1e-4 * std::max(B.cell[0], std::max(B.cell[1], B.cell[2])) + 4e-7 * e.maxabs;
    B.eps_u = 1e-5;
    B.E = E; B.m = m;
    B.nS = c->n_spheres; B.nC = c->n_cuboids; B.max_spheres = c->max_spheres;
    B.ubo = c->d_objects;
    const size_t bytes = (size_t)B.total * sizeof(unsigned long long);
    if (bytes > ((size_t)64 << 20)) return fail(PTB_E_INVALID, "ray-classification table of %zu bytes (cells %d, buckets %d) exceeds 64 MiB", bytes, c->rct_cells, c->rct_G);
    if (bytes > c->rct_capacity) {
        if (c->d_rct) CU(cudaFree(c->d_rct));
        c->d_rct = nullptr;
        CU(cudaMalloc(&c->d_rct, bytes));
        c->rct_capacity = bytes;
    }
    B.table = c->d_rct;
    rct_build_kernel<<<(unsigned)((B.total + 127) / 128), 128, 0, c->stream>>>(B);      // after the H2D of the UBO on the same stream
    c->launches++;
    CU(cudaGetLastError());
    const int nS = c->n_spheres;
    c->rct_sm0 = nS >= 32 ? 0xffffffffu : ((1u << nS) - 1u);
    c->rct_sm1 = nS >= 64 ? 0xffffffffu : (nS > 32 ? ((1u << (nS - 32)) - 1u) : 0u);
    c->rct_on = true;
    return PTB_OK;
}

// Uniform grid over the primitives' inflated boxes (see trace_grid).  Returns false when the scene does not fit the 16-bit
// lists (the caller then builds the BVH instead).
bool build_grid(ptb_ctx* c)
{
    c->grid_on = false; c->grid_cell_start.clear(); c->grid_items.clear(); c->grid_cell_spheres.clear();
    c->bvh_nodes.clear(); c->bvh_pidx.clear(); c->n_nodes = 0; c->n_unbounded = 0; c->bvh_tau = 0.0f; c->bvh_D = 0.0f;
    const int n = c->n_spheres + c->n_cuboids;
    if (n < c->bvh_threshold || n >= 65535) return false;
    std::vector<Box> boxes; std::vector<char> bounded;
    raw_boxes(c, 0.0, 0.0, boxes, bounded);
    std::vector<Box> finite;
    for (int i = 0; i < n; ++i) if (bounded[i]) finite.push_back(boxes[i]);
    if (finite.empty()) return false;
    c->bvh_extent = scene_extent(finite);
    const float D = required_extent(c, c->bvh_extent);
    const double E = 4e-6 * (double)D * D, m = 1e-5 * (double)D + 1e-6;
    raw_boxes(c, E, m, boxes, bounded);
    double slo[3] = {1e300, 1e300, 1e300}, shi[3] = {-1e300, -1e300, -1e300};
    for (int i = 0; i < n; ++i) if (bounded[i]) for (int k = 0; k < 3; ++k) { slo[k] = std::min(slo[k], boxes[i].lo[k]); shi[k] = std::max(shi[k], boxes[i].hi[k]); }
    std::vector<int> binned;
    for (int i = 0; i < n; ++i) {
        bool big = !bounded[i];
        if (!big) for (int k = 0; k < 3; ++k) if (boxes[i].hi[k] - boxes[i].lo[k] > PTB_BVH_BIG * (shi[k] - slo[k])) big = true;
        if (big && (int)c->bvh_pidx.size() < 64) c->bvh_pidx.push_back(i); else if (bounded[i]) binned.push_back(i); else c->bvh_pidx.push_back(i);
    }
    c->n_unbounded = (int)c->bvh_pidx.size();
    c->bvh_D = D;
    c->bvh_tau = 1e-4f * D;
    // bounds of what is binned
    double glo[3] = {1e300, 1e300, 1e300}, ghi[3] = {-1e300, -1e300, -1e300};
    for (int i : binned) for (int k = 0; k < 3; ++k) { glo[k] = std::min(glo[k], boxes[i].lo[k]); ghi[k] = std::max(ghi[k], boxes[i].hi[k]); }
    if (binned.empty()) for (int k = 0; k < 3; ++k) { glo[k] = 0.0; ghi[k] = 1.0; }
    double ext[3], vol = 1.0;
    for (int k = 0; k < 3; ++k) { ext[k] = std::max(ghi[k] - glo[k], 1e-3); vol *= ext[k]; }
    const double target = std::min(8192.0, std::max(64.0, (double)c->grid_density * (double)binned.size()));
    const double side = std::cbrt(vol / target);
    // the grid margin: a primitive is listed in every cell its box comes within g of — g covers the rounding of the DDA
    // (cell boundaries and accumulated exit parameters are off by far less than tau = 1e-4 D) and of the cell lookup
    double cellmin = 1e300;
    for (int k = 0; k < 3; ++k) {
        int nk = (int)std::ceil(ext[k] / side);
        nk = std::min(std::max(nk, 1), 32);
        c->grid_n[k] = nk;
        cellmin = std::min(cellmin, ext[k] / nk);
    }
    const double g = (double)c->bvh_tau + 1e-3 * cellmin;
    for (int k = 0; k < 3; ++k) {
        const double lo = glo[k] - 2.0 * g, hi = ghi[k] + 2.0 * g;
        c->grid_lo[k] = (float)lo; c->grid_hi[k] = (float)hi;
        c->grid_cell[k] = (float)((hi - lo) / c->grid_n[k]);
        c->grid_inv[k] = (float)(c->grid_n[k] / (hi - lo));
    }
    const int ncell = c->grid_n[0] * c->grid_n[1] * c->grid_n[2];
    std::vector<std::vector<unsigned short>> lists((size_t)ncell);
    size_t total = 0;
    for (int i : binned) {
        int a[3], b[3];
        for (int k = 0; k < 3; ++k) {
            const double cs = (double)c->grid_cell[k];
            a[k] = std::min(std::max((int)std::floor((boxes[i].lo[k] - g - (double)c->grid_lo[k]) / cs), 0), c->grid_n[k] - 1);
            b[k] = std::min(std::max((int)std::floor((boxes[i].hi[k] + g - (double)c->grid_lo[k]) / cs), 0), c->grid_n[k] - 1);
        }
        for (int z = a[2]; z <= b[2]; ++z)
            for (int y = a[1]; y <= b[1]; ++y)
                for (int x = a[0]; x <= b[0]; ++x) { lists[((size_t)z * c->grid_n[1] + y) * c->grid_n[0] + x].push_back((unsigned short)i); ++total; }
    }
    if (total >= 65535) return false;
    // per cell an offset into the item list and the number of its items that are spheres (items ascend, so the spheres come
    // first): the kernel runs two uniform loops per cell instead of one loop that branches on the primitive type
    c->grid_cell_start.resize((size_t)ncell + 1);
    c->grid_cell_spheres.assign((size_t)ncell, 0);
    for (int q = 0; q < ncell; ++q) {
        c->grid_cell_start[q] = (unsigned short)c->grid_items.size();
        size_t n_sph = 0;
        for (unsigned short v : lists[q]) if ((int)v < c->n_spheres) ++n_sph;
        if (n_sph > 255) return false;
        c->grid_cell_spheres[q] = (unsigned char)n_sph;
        c->grid_items.insert(c->grid_items.end(), lists[q].begin(), lists[q].end());      // ascending index inside a cell (binned is ascending)
    }
    c->grid_cell_start[ncell] = (unsigned short)c->grid_items.size();
    c->grid_on = true;
    return true;
}

void layout_block(ptb_ctx* c)
{
    // float4 units: [spheres][slabs][BVH nodes][always-tested list + BVH leaves][grid offsets][grid sphere counts][grid items] | [1/r][materials]
    // (what follows the bar is needed once per hit: small scenes stage it too, large scenes leave it in HBM / L2)
    const int nS = c->n_spheres, nC = c->n_cuboids;
    const int n4 = (nS + 3) & ~3;                 // the sphere array is padded to a multiple of four (never-hit dummies)
    c->off_cmin = n4;
    c->off_cmax = c->off_cmin + 1;                // slab bounds interleaved: lo0, hi0, lo1, hi1, ...
    c->off_nodes = c->off_cmin + 2 * nC;
    c->off_pidx = c->off_nodes + 2 * c->n_nodes;
    c->off_gcell = c->off_pidx + ((int)c->bvh_pidx.size() + 3) / 4;
    c->off_gsph = c->off_gcell + (c->grid_on ? ((int)c->grid_cell_start.size() + 7) / 8 : 0);
    c->off_gitem = c->off_gsph + (c->grid_on ? ((int)c->grid_cell_spheres.size() + 15) / 16 : 0);
    const int off_tail = c->off_gitem + (c->grid_on ? ((int)c->grid_items.size() + 7) / 8 : 0);
    c->off_aux = off_tail;
    c->off_mat = c->off_aux + (nS + 3) / 4;
    c->block_bytes = (c->off_mat + (nS + nC) * 4) * 16;
    if (c->block_bytes < 16) c->block_bytes = 16;
    // with a BVH the materials stay in HBM / L2 (only the winner's 64 B are read per bounce) so more CTAs fit an SM
    c->stage_bytes = c->n_nodes > 0 || c->n_unbounded > 0 || c->grid_on ? std::max(16, c->off_aux * 16) : c->block_bytes;
}

int sync_scene(ptb_ctx* c)
{
    if (!c->scene_dirty) return PTB_OK;
    if (!(c->large_mode == 1 && build_grid(c))) { c->grid_on = false; build_bvh(c); }
    layout_block(c);
    if (c->stage_bytes > kMaxSmem - 1024)
        return fail(PTB_E_INVALID, "scene block of %d bytes does not fit shared memory (%d spheres, %d cuboids)", c->block_bytes, c->n_spheres, c->n_cuboids);
    { const int rc = sync_all(c); if (rc != PTB_OK) return rc; }     // kernels in flight still read the old block / UBO copy
    if ((size_t)c->block_bytes > c->block_capacity) {
        if (c->d_block) CU(cudaFree(c->d_block));
        c->d_block = nullptr;
        CU(cudaMalloc(&c->d_block, (size_t)c->block_bytes));
        c->block_capacity = (size_t)c->block_bytes;
    }
    CU(cudaMemcpyAsync(c->d_objects, c->objects.data(), c->objects.size(), cudaMemcpyHostToDevice, c->stream));
    CU(cudaMemsetAsync(c->d_block, 0, (size_t)c->block_bytes, c->stream));
    const int n = c->n_spheres + c->n_cuboids;
    if (n > 0) {
        pack_scene_kernel<<<(n + 127) / 128, 128, 0, c->stream>>>(c->d_objects, c->max_spheres, c->n_spheres, c->n_cuboids, c->d_block, c->off_aux,
                                                                  c->off_cmin, c->off_cmax, c->off_mat);
        c->launches++;
        CU(cudaGetLastError());
    }
    {
        std::vector<unsigned char> geo;
        geometry_bytes(c, geo);
        if (!rct_eligible(c)) { c->rct_on = false; c->rct_geometry.clear(); }
        else if (geo != c->rct_geometry || !c->d_rct) {
            const int rc = build_rct(c);
            if (rc != PTB_OK) return rc;
            c->rct_geometry.swap(geo);
        }
    }
    if (c->n_nodes > 0) CU(cudaMemcpyAsync(c->d_block + c->off_nodes, c->bvh_nodes.data(), c->bvh_nodes.size() * sizeof(float4), cudaMemcpyHostToDevice, c->stream));
    if (!c->bvh_pidx.empty()) CU(cudaMemcpyAsync(c->d_block + c->off_pidx, c->bvh_pidx.data(), c->bvh_pidx.size() * sizeof(int), cudaMemcpyHostToDevice, c->stream));
    if (c->grid_on) {
        CU(cudaMemcpyAsync(c->d_block + c->off_gcell, c->grid_cell_start.data(), c->grid_cell_start.size() * sizeof(unsigned short), cudaMemcpyHostToDevice, c->stream));
        CU(cudaMemcpyAsync(c->d_block + c->off_gsph, c->grid_cell_spheres.data(), c->grid_cell_spheres.size(), cudaMemcpyHostToDevice, c->stream));
        if (!c->grid_items.empty()) CU(cudaMemcpyAsync(c->d_block + c->off_gitem, c->grid_items.data(), c->grid_items.size() * sizeof(unsigned short), cudaMemcpyHostToDevice, c->stream));
    }
    if (c->n_nodes > 0 || !c->bvh_pidx.empty() || c->grid_on) CU(cudaStreamSynchronize(c->stream));     // the host vectors may be rebuilt before the copy ran
    c->scene_dirty = false;
    return mark_inputs(c);
}

void fill_params(ptb_ctx* c, RenderParams& P)
{
    memcpy(P.basic, c->basic, kBasicBytes);
    P.width = c->width; P.height = c->height; P.frame = c->frame;
    P.spp = c->spp; P.ray_depth = c->ray_depth;
    P.focal_length = c->focal_length; P.aperture_diameter = c->aperture_diameter;
    P.n_spheres = c->n_spheres; P.n_cuboids = c->n_cuboids;
    P.env_size = c->env_size;
    P.rank = c->rank; P.world = c->world; P.stripe_rows = c->stripe_rows; P.local_rows = c->local_rows;
    P.off_aux = c->off_aux; P.off_cmin = c->off_cmin; P.off_cmax = c->off_cmax; P.off_mat = c->off_mat; P.block_bytes = c->block_bytes;
    P.stage_bytes = c->stage_bytes; P.off_nodes = c->off_nodes; P.off_pidx = c->off_pidx; P.n_nodes = c->n_nodes; P.n_unbounded = c->n_unbounded;
    P.bvh_tau = c->bvh_tau;
    P.scene = c->d_block; P.env = c->d_env; P.image = c->d_image;
    P.counters = c->d_counters; P.stats = c->d_stats;
    P.raw_objects = c->d_objects; P.max_spheres = c->max_spheres;
    P.inv_width = 1.0f / (float)c->width;
    P.inv_height = 1.0f / (float)c->height;
    P.inv_spp = 1.0f / (float)c->spp;
    P.blend = 1.0f * (1.0f / (float)(c->frame + 1));   // 1.0 / (thisRendererFrame + 1), compute.glsl:128, as a * rcp(b)
    P.tiles_x = (unsigned)((c->width + 7) / 8);
    P.tiles_total = P.tiles_x * (unsigned)((c->local_rows + 3) / 4);
    // umulhi(n, ceil(2^32/d)) == n / d exactly while n * d < 2^32 (error term n*e/(d*2^32) < 1/d); d == 1 needs no division
    P.tiles_magic = (P.tiles_x > 1 && (unsigned long long)P.tiles_total * P.tiles_x < (1ull << 32)) ? (unsigned)(((1ull << 32) + P.tiles_x - 1) / P.tiles_x) : 0u;
    P.batch = 1;
    P.scratch_stride = 0ull;
    P.ktime = nullptr;
    P.defer_finish = (c->spp == 1 && c->defer_finish) ? 1 : 0;
    P.done_flag = nullptr; P.done_value = 0u;
    P.off_gcell = c->off_gcell; P.off_gsph = c->off_gsph; P.off_gitem = c->off_gitem;
    for (int k = 0; k < 3; ++k) { P.grid_n[k] = c->grid_n[k]; P.grid_lo[k] = c->grid_lo[k]; P.grid_hi[k] = c->grid_hi[k]; P.grid_cell[k] = c->grid_cell[k]; P.grid_inv[k] = c->grid_inv[k]; }
    P.rct = c->rct_on ? c->d_rct : nullptr;
    for (int k = 0; k < 3; ++k) { P.rct_lo[k] = c->rct_lo[k]; P.rct_inv[k] = c->rct_inv[k]; P.rct_n[k] = c->rct_n[k]; }
    P.rct_G = c->rct_G; P.rct_halfG = 0.5f * (float)c->rct_G;
    P.rct_sm0 = c->rct_sm0; P.rct_sm1 = c->rct_sm1;
    for (int k = 0; k < 3; ++k) P.rct_nf[k] = (float)c->rct_n[k];
    { const int n = c->n_spheres + c->n_cuboids; P.rct_valid = n >= 64 ? ~0ull : ((1ull << n) - 1ull); }
}

// Where frame number q of the exchange goes: the root's block, the slot in it and how many frames that root's consumer must
// have released before the slot may be written again.
struct XchTarget { unsigned char* block; int slot; unsigned need; };
XchTarget xch_target(const ptb_ctx* c, unsigned long long q)
{
    const bool rot = c->xch_roots > 1;
    const unsigned long long li = rot ? q / (unsigned long long)c->world : q;          // index among the root's own frames
    XchTarget t;
    t.block = rot ? c->xch_peer[q % (unsigned long long)c->world] : c->xch_block;
    t.slot = (int)(li % (unsigned long long)c->xch_slots);
    t.need = li >= (unsigned long long)c->xch_slots ? (unsigned)(li - c->xch_slots + 1) : 0u;
    return t;
}
// frames with number < seq that this rank is the root of
unsigned long long xch_own_frames(const ptb_ctx* c, unsigned long long seq)
{
    if (c->xch_roots <= 1) return c->rank == 0 ? seq : 0ull;
    return (seq + (unsigned long long)(c->world - 1 - c->rank)) / (unsigned long long)c->world;
}

int fold_of(const ptb_ctx* c) { return c->grid_on ? 3 : ((c->n_nodes > 0 || c->n_unbounded > 0) ? 1 : (c->rct_on ? 2 : 0)); }

// One megakernel launch of the current configuration (exact or fast translation unit) on `stream`.
int launch_mega(ptb_ctx* c, const RenderParams& P, bool batch, int smem, cudaStream_t stream);

// ring mode of the next launch: 0 no ring, 1 ring, 2 ring + completion queue (SPP 1, no statistics)
int ring_mode(const ptb_ctx* c, bool stats) { return !c->mega_ring ? 0 : ((c->mega_defer && c->spp == 1 && c->defer_finish && !stats) ? 2 : 1); }

template <int kFold, class F>
int with_mega_fold(ptb_ctx* c, bool stats, F&& launch)
{
    switch (ring_mode(c, stats)) {
    case 2: return launch(megakernel<false, 2, kFold>);
    case 1: return stats ? launch(megakernel<true, 1, kFold>) : launch(megakernel<false, 1, kFold>);
    default: return stats ? launch(megakernel<true, 0, kFold>) : launch(megakernel<false, 0, kFold>);
    }
}
template <class F>
int with_mega(ptb_ctx* c, bool stats, F&& launch)
{
    switch (fold_of(c)) {
    case 1: return with_mega_fold<1>(c, stats, launch);
    case 2: return with_mega_fold<2>(c, stats, launch);
    case 3: return with_mega_fold<3>(c, stats, launch);
    default: return with_mega_fold<0>(c, stats, launch);
    }
}

template <int kFold>
int prepare_mega(ptb_ctx* c, int smem, int& with_ring, int& with_queue, int& without)
{
    CU(cudaFuncSetAttribute(megakernel<false, 2, kFold>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
    CU(cudaFuncSetAttribute(megakernel<false, 1, kFold>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
    CU(cudaFuncSetAttribute(megakernel<true, 1, kFold>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
    CU(cudaFuncSetAttribute(megakernel<false, 0, kFold>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
    CU(cudaFuncSetAttribute(megakernel<true, 0, kFold>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
    CU(cudaFuncSetAttribute(megakernel<false, 2, kFold, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
    CU(cudaFuncSetAttribute(megakernel<false, 1, kFold, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
    CU(cudaFuncSetAttribute(megakernel<false, 0, kFold, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
    CU(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&with_queue, megakernel<false, 2, kFold>, kMegaThreads, smem));
    CU(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&with_ring, megakernel<false, 1, kFold>, kMegaThreads, smem));
    CU(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&without, megakernel<false, 0, kFold>, kMegaThreads, smem));
    return PTB_OK;
}

int launch_frame(ptb_ctx* c)
{
    RenderParams P;
    fill_params(c, P);
    P.scratch = nullptr;
    if (c->kernel == PTB_KERNEL_NAIVE) {
        dim3 grid((c->width + 7) / 8, (c->local_rows + 7) / 8, 1), block(8, 8, 1);   // PathTracer.cs:121
        if (grid.y > 0) naive_kernel<<<grid, block, 0, c->stream>>>(P);
        c->launches++;
        CU(cudaGetLastError());
        c->frame++;
        return mark_inputs(c);      // the image changed on the user-visible stream
    }
    const int smem = c->stage_bytes;
    const int fold = fold_of(c);
    if (c->mega_smem_set != smem || c->mega_fold_set != fold || c->mega_precision_set != c->precision) {
        int with_ring = 0, with_queue = 0, without = 0;
        if (c->precision == PTB_PRECISION_FAST) {
            CU(ptb_fast_api::prepare(fold, smem, &with_ring, &with_queue, &without));
        } else {
            const int rc = fold == 1 ? prepare_mega<1>(c, smem, with_ring, with_queue, without)
                         : (fold == 2 ? prepare_mega<2>(c, smem, with_ring, with_queue, without)
                         : (fold == 3 ? prepare_mega<3>(c, smem, with_ring, with_queue, without) : prepare_mega<0>(c, smem, with_ring, with_queue, without)));
            if (rc != PTB_OK) return rc;
        }
        if (without < 1) return fail(PTB_E_CUDA, "megakernel does not fit an SM with %d bytes of shared memory", smem);
        c->mega_ring = with_ring >= without;          // the ring must not cost a resident CTA
        c->mega_defer = c->mega_ring && with_queue >= with_ring;      // nor may the completion queue
        c->mega_grid = std::max(c->sm_count, c->sm_count * (c->mega_ring ? with_ring : without) / c->grid_divisor);
        c->mega_smem_set = smem;
        c->mega_fold_set = fold;
        c->mega_precision_set = c->precision;
    }
    if (c->local_rows > 0) {
        if (c->overlap <= 1) {
            // classic: accumulate in place on the user-visible stream
            const int rc = launch_mega(c, P, false, smem, c->stream);
            if (rc != PTB_OK) return rc;
            c->launches++;
            CU(cudaGetLastError());
            { const int r2 = mark_inputs(c); if (r2 != PTB_OK) return r2; }
        } else {
            { const int rc = ensure_pipeline(c); if (rc != PTB_OK) return rc; }
            const int s = (int)(c->launch_seq % (unsigned long long)c->overlap);
            cudaStream_t ts = c->trace_stream[s];
            if (c->seen_version[s] != c->inputs_version) { CU(cudaStreamWaitEvent(ts, c->ev_inputs, 0)); c->seen_version[s] = c->inputs_version; }
            if (c->blend_recorded[s]) CU(cudaStreamWaitEvent(ts, c->ev_blend_done[s], 0));      // scratch[s] has been consumed
            P.scratch = c->d_scratch[s];
            P.counters = c->d_slot_counters[s];
            const int rc = launch_mega(c, P, false, smem, ts);
            if (rc != PTB_OK) return rc;
            CU(cudaEventRecord(c->ev_trace_done[s], ts));
            cudaStream_t bs = c->blend_stream;
            if (c->seen_version[kMaxOverlap] != c->inputs_version) { CU(cudaStreamWaitEvent(bs, c->ev_inputs, 0)); c->seen_version[kMaxOverlap] = c->inputs_version; }
            CU(cudaStreamWaitEvent(bs, c->ev_trace_done[s], 0));
            const size_t n = (size_t)c->local_rows * c->width;
            if (c->xch_on) {
                const XchTarget tg = xch_target(c, c->xch_seq);
                ExchangeFlags* fl = reinterpret_cast<ExchangeFlags*>(tg.block);
                // a one-thread kernel waits for the slot: this blend grid is as large as the image, and a grid that large spinning
                // on `consumed` would hold every CTA slot the consumer's own kernels (read-back snapshot, ...) need before it can release
                if (tg.need > 0u) { exchange_wait_free_kernel<<<1, 1, 0, bs>>>(fl, tg.need); c->launches++; }
                BatchBlend B = {};
                B.frame0 = c->frame; B.frames = 1; B.stride = 0ull; B.blend[0] = P.blend;
                BatchScatter X = {};
                X.slot[0] = tg.slot; X.flags[0] = fl; X.need[0] = 0u;
                X.full[0] = tg.block + 4096 + (size_t)tg.slot * c->xch_image_bytes;
                const BatchWait none = {nullptr, 0u, nullptr};      // ordered by the stream event above
                blend_scatter_batch_kernel<<<(unsigned)std::min<size_t>((n + 255) / 256, (size_t)c->sm_count * 4), 256, 0, bs>>>(
                    c->d_image, c->d_scratch[s], c->width, c->local_rows, c->height, c->rank, c->world, c->stripe_rows, c->xch_rgb ? 1 : 0, B, X,
                    c->d_xch_blocks, none);
                c->xch_seq++;
            } else {
                blend_kernel<<<(unsigned)((n + 255) / 256), 256, 0, bs>>>(c->d_image, c->d_scratch[s], n, c->frame, P.blend);
            }
            CU(cudaGetLastError());
            CU(cudaEventRecord(c->ev_blend_done[s], bs));
            c->blend_recorded[s] = true;
            CU(cudaStreamWaitEvent(c->stream, c->ev_blend_done[s], 0));       // whatever the host enqueues next sees this frame
            c->launches += 2;
            c->launch_seq++;
        }
    } else if (c->xch_on) {
        // no rows on this rank (more ranks than stripes): it still has to arrive at the frame's slot, or rank 0 would wait for ever
        { const int rc = ensure_pipeline(c); if (rc != PTB_OK) return rc; }
        cudaStream_t bs = c->blend_stream;
        const XchTarget tg = xch_target(c, c->xch_seq);
        ExchangeFlags* fl = reinterpret_cast<ExchangeFlags*>(tg.block);
        if (tg.need > 0u) { exchange_wait_free_kernel<<<1, 1, 0, bs>>>(fl, tg.need); c->launches++; }
        exchange_arrive_kernel<<<1, 1, 0, bs>>>(fl, tg.slot);
        CU(cudaGetLastError());
        c->launches++;
        c->xch_seq++;
    }
    c->frame++;   // PathTracer.cs:117 thisRenderNumFrame++
    return PTB_OK;
}

// ---- frame batching (ptb_set_batch) -----------------------------------------------------------------------------------
template <int kFold, class F>
int with_mega_batch_fold(ptb_ctx* c, F&& launch)
{
    switch (ring_mode(c, false)) {
    case 2: return launch(megakernel<false, 2, kFold, true>);
    case 1: return launch(megakernel<false, 1, kFold, true>);
    default: return launch(megakernel<false, 0, kFold, true>);
    }
}
template <class F>
int with_mega_batch(ptb_ctx* c, F&& launch)
{
    switch (fold_of(c)) {
    case 1: return with_mega_batch_fold<1>(c, launch);
    case 2: return with_mega_batch_fold<2>(c, launch);
    case 3: return with_mega_batch_fold<3>(c, launch);
    default: return with_mega_batch_fold<0>(c, launch);
    }
}

int kt_collect(ptb_ctx* c, int slot)
{
    if (!c->kt_used[slot]) return PTB_OK;
    CU(cudaEventSynchronize(c->kt_b[slot]));
    // device-side bracket: first CTA start -> last CTA end (an event pair on the stream would also count the time a launch
    // waits for SM slots behind the previous, still resident, persistent grid)
    const unsigned long long t0 = c->h_kt[2 * slot], t1 = c->h_kt[2 * slot + 1];
    if (t1 > t0) c->kt_ms += (double)(t1 - t0) * 1e-6;
    c->kt_frames_total += c->kt_frames[slot]; c->kt_launches++;
    c->kt_used[slot] = false;
    return PTB_OK;
}

int launch_mega(ptb_ctx* c, const RenderParams& P, bool batch, int smem, cudaStream_t stream)
{
    int slot = -1;
    if (c->kt_on) {
        slot = (int)(c->kt_next++ % 8u);
        { const int rc = kt_collect(c, slot); if (rc != PTB_OK) return rc; }
        CU(cudaMemcpyAsync(c->d_kt + 2 * slot, c->h_kt + 16, 16, cudaMemcpyHostToDevice, stream));
    }
    RenderParams Pt = P;
    Pt.ktime = slot >= 0 ? c->d_kt + 2 * slot : nullptr;
    if (c->precision == PTB_PRECISION_FAST) {
        CU(ptb_fast_api::launch(&Pt, fold_of(c), ring_mode(c, false), batch, c->mega_grid, smem, stream));
    } else {
        const int rc = batch ? with_mega_batch(c, [&](auto k) { k<<<c->mega_grid, kMegaThreads, smem, stream>>>(Pt); return PTB_OK; })
                             : with_mega(c, c->stats_on, [&](auto k) { k<<<c->mega_grid, kMegaThreads, smem, stream>>>(Pt); return PTB_OK; });
        if (rc != PTB_OK) return rc;
        CU(cudaGetLastError());
    }
    if (slot >= 0) {
        CU(cudaMemcpyAsync(c->h_kt + 2 * slot, c->d_kt + 2 * slot, 16, cudaMemcpyDeviceToHost, stream));
        CU(cudaEventRecord(c->kt_b[slot], stream));
        c->kt_frames[slot] = batch ? P.batch : 1;
        c->kt_used[slot] = true;
    }
    return PTB_OK;
}

bool batch_eligible(const ptb_ctx* c)
{
    // the frame slot rides in bits 12..15 of the ring's pixel word; statistics and the proxy kernel stay per frame
    return c->batch > 1 && c->kernel == PTB_KERNEL_MEGA && c->overlap >= 2 && c->local_rows > 0 && c->width <= 4096 && !c->stats_on;
}

int ensure_batch(ptb_ctx* c, int frames)
{
    { const int rc = ensure_pipeline(c); if (rc != PTB_OK) return rc; }      // trace streams 0/1, blend stream
    const size_t stride = c->image_bytes / sizeof(float4);
    if (c->batch_scratch_frames < (size_t)frames || c->batch_scratch_stride != stride) {
        { const int rc = sync_all(c); if (rc != PTB_OK) return rc; }
        for (int s = 0; s < kBatchSets; ++s) {
            if (c->d_batch_scratch[s]) { CU(cudaFree(c->d_batch_scratch[s])); c->d_batch_scratch[s] = nullptr; }
            CU(cudaMalloc(&c->d_batch_scratch[s], (size_t)frames * c->image_bytes));
            c->batch_blend_recorded[s] = false;
        }
        c->batch_scratch_frames = (size_t)frames;
        c->batch_scratch_stride = stride;
    }
    if (!c->d_done) {
        CU(cudaMalloc(&c->d_done, (kBatchSets + 1) * sizeof(unsigned)));
        CU(cudaMemsetAsync(c->d_done, 0, (kBatchSets + 1) * sizeof(unsigned), c->stream));
        { const int rc = mark_inputs(c); if (rc != PTB_OK) return rc; }
    }
    for (int s = 0; s < kBatchSets; ++s) {
        if (!c->d_batch_counters[s]) {
            CU(cudaMalloc(&c->d_batch_counters[s], 2 * sizeof(unsigned int)));
            CU(cudaMemsetAsync(c->d_batch_counters[s], 0, 2 * sizeof(unsigned int), c->stream));
            CU(cudaEventCreateWithFlags(&c->ev_batch_blend[s], cudaEventDisableTiming));
            { const int rc = mark_inputs(c); if (rc != PTB_OK) return rc; }
        }
    }
    return PTB_OK;
}

// `frames` (2..16) consecutive frames: one trace launch, then the per-frame blends in frame order on the blend stream.
int launch_batch(ptb_ctx* c, int frames)
{
    { const int rc = ensure_batch(c, std::max(frames, c->batch)); if (rc != PTB_OK) return rc; }     // sized once for the configured batch
    RenderParams P;
    fill_params(c, P);
    const int smem = c->stage_bytes;
    if (c->mega_smem_set != smem || c->mega_fold_set != fold_of(c) || c->mega_precision_set != c->precision) {
        // the grid / ring decision (and the shared-memory attributes of every instantiation) belong to the single-frame path
        const int rc = launch_frame(c);
        if (rc != PTB_OK) return rc;
        return frames > 1 ? (frames - 1 >= 2 ? launch_batch(c, frames - 1) : launch_frame(c)) : PTB_OK;
    }
    const int s = (int)(c->batch_seq % (unsigned long long)kBatchSets);      // scratch set
    const int t = (int)(c->batch_seq & 1ull);                                // trace stream
    cudaStream_t ts = c->trace_stream[t];
    if (c->seen_version[t] != c->inputs_version) { CU(cudaStreamWaitEvent(ts, c->ev_inputs, 0)); c->seen_version[t] = c->inputs_version; }
    if (c->batch_blend_recorded[s]) CU(cudaStreamWaitEvent(ts, c->ev_batch_blend[s], 0));      // this scratch set has been consumed
    P.scratch = c->d_batch_scratch[s];
    P.counters = c->d_batch_counters[s];
    P.batch = frames;
    P.scratch_stride = (unsigned long long)c->batch_scratch_stride;
    // the blend kernel of this batch does not wait for a stream event: it is launched right behind the trace, takes CTA slots
    // as soon as the previous grid frees some, and spins on this flag (see wait_for_trace)
    P.done_flag = c->d_done + s;
    P.done_value = (unsigned)(c->batch_seq + 1ull);
    const int rc = launch_mega(c, P, true, smem, ts);
    if (rc != PTB_OK) return rc;
    cudaStream_t bs = c->blend_stream;
    if (c->seen_version[kMaxOverlap] != c->inputs_version) { CU(cudaStreamWaitEvent(bs, c->ev_inputs, 0)); c->seen_version[kMaxOverlap] = c->inputs_version; }
    const size_t n = (size_t)c->local_rows * c->width;
    c->launches++;
    BatchWait Wt = {c->d_done + s, (unsigned)(c->batch_seq + 1ull), c->d_done + kBatchSets};
    // one blend kernel for the whole batch: the running mean is folded frame by frame in registers (same operations, same order)
    BatchBlend B = {};
    B.frame0 = c->frame; B.frames = frames; B.stride = (unsigned long long)c->batch_scratch_stride;
    for (int j = 0; j < frames; ++j) B.blend[j] = 1.0f * (1.0f / (float)(c->frame + j + 1));     // as fill_params: 1.0 / (thisRendererFrame + 1)
    // a batch blend is a background kernel: it has a whole batch time to move a few hundred MB, so it gets a small grid and
    // leaves the CTA slots to the next batch's persistent grid (PTB_BLEND_CTAS overrides the default for experiments)
    static const int blend_ctas_env = getenv("PTB_BLEND_CTAS") ? atoi(getenv("PTB_BLEND_CTAS")) : 0;
    const size_t blend_cap = blend_ctas_env > 0 ? (size_t)blend_ctas_env : (size_t)c->blend_ctas;
    const unsigned blend_grid = (unsigned)std::min<size_t>((n + 255) / 256, blend_cap);
    if (c->xch_on) {
        BatchScatter X = {};
        for (int j = 0; j < frames; ++j) {
            const XchTarget tg = xch_target(c, c->xch_seq + j);
            X.slot[j] = tg.slot; X.need[j] = tg.need;
            X.flags[j] = reinterpret_cast<ExchangeFlags*>(tg.block);
            X.full[j] = tg.block + 4096 + (size_t)tg.slot * c->xch_image_bytes;
        }
        blend_scatter_batch_kernel<<<blend_grid, 256, 0, bs>>>(
            c->d_image, c->d_batch_scratch[s], c->width, c->local_rows, c->height, c->rank, c->world, c->stripe_rows, c->xch_rgb ? 1 : 0, B, X,
            c->d_xch_blocks, Wt);
        c->xch_seq += frames;
    } else {
        blend_batch_kernel<<<blend_grid, 256, 0, bs>>>(c->d_image, c->d_batch_scratch[s], n, B, Wt);
    }
    CU(cudaGetLastError());
    c->launches++;
    CU(cudaEventRecord(c->ev_batch_blend[s], bs));
    c->batch_blend_recorded[s] = true;
    CU(cudaStreamWaitEvent(c->stream, c->ev_batch_blend[s], 0));       // whatever the host enqueues next sees these frames
    c->batch_seq++;
    c->frame += frames;
    return PTB_OK;
}

} // namespace

extern "C" {

static int exchange_close(ptb_ctx* c);
static int install_environment(ptb_ctx* c, int N);

const char* ptb_last_error(void) { return g_err; }
int ptb_version(void) { return 100; }

int ptb_create(ptb_ctx** out, int width, int height, int max_spheres, int max_cuboids, int device)
{
    if (!out) return fail(PTB_E_INVALID, "out is null");
    *out = nullptr;
    if (width <= 0 || height <= 0 || width > 65535 || height > 65535 || max_spheres < 0 || max_cuboids < 0) return fail(PTB_E_INVALID, "bad size %dx%d (1..65535) / capacities %d,%d", width, height, max_spheres, max_cuboids);
    int count = 0;
    CU(cudaGetDeviceCount(&count));
    if (device < 0 || device >= count) return fail(PTB_E_INVALID, "device %d out of range (%d CUDA devices)", device, count);
    CU(cudaSetDevice(device));
    cudaDeviceProp prop;
    CU(cudaGetDeviceProperties(&prop, device));
    if (prop.major != 10) return fail(PTB_E_CUDA, "libptb200 is built for sm_100a only; device %d is sm_%d%d", device, prop.major, prop.minor);
    ptb_ctx* c = new (std::nothrow) ptb_ctx();
    if (!c) return fail(PTB_E_NOMEM, "out of host memory");
    c->device = device;
    c->sm_count = prop.multiProcessorCount;
    if (getenv("PTB_DEFER")) c->defer_finish = atoi(getenv("PTB_DEFER")) != 0;
    if (getenv("PTB_RCT")) {       // experiments: "cells,buckets" of the ray-classification table
        int a = 0, b = 0;
        if (sscanf(getenv("PTB_RCT"), "%d,%d", &a, &b) == 2 && a >= 1 && a <= 64 && b >= 1 && b <= 64) { c->rct_cells = a; c->rct_G = b; }
    }
    c->width = width; c->height = height;
    c->max_spheres = max_spheres; c->max_cuboids = max_cuboids;
    c->objects.assign((size_t)max_spheres * kSphereStride + (size_t)max_cuboids * kCuboidStride + 16, 0);
    int rc = PTB_OK;
    do {
        if (cudaStreamCreateWithFlags(&c->stream, cudaStreamNonBlocking) != cudaSuccess) { rc = fail(PTB_E_CUDA, "cudaStreamCreate failed"); break; }
        c->own_stream = true;
        if (cudaEventCreate(&c->ev0) != cudaSuccess || cudaEventCreate(&c->ev1) != cudaSuccess) { rc = fail(PTB_E_CUDA, "cudaEventCreate failed"); break; }
        if (cudaMalloc(&c->d_objects, c->objects.size()) != cudaSuccess) { rc = fail(PTB_E_NOMEM, "cudaMalloc(objects) failed"); break; }
        if (cudaMalloc(&c->d_counters, 2 * sizeof(unsigned int)) != cudaSuccess) { rc = fail(PTB_E_NOMEM, "cudaMalloc(counters) failed"); break; }
        if (cudaMalloc(&c->d_stats, 4 * sizeof(unsigned long long)) != cudaSuccess) { rc = fail(PTB_E_NOMEM, "cudaMalloc(stats) failed"); break; }
        cudaMemsetAsync(c->d_counters, 0, 2 * sizeof(unsigned int), c->stream);
        cudaMemsetAsync(c->d_stats, 0, 4 * sizeof(unsigned long long), c->stream);
        rc = alloc_image(c);
    } while (0);
    if (rc != PTB_OK) { ptb_destroy(c); return rc; }
    *out = c;
    return PTB_OK;
}

void ptb_destroy(ptb_ctx* c)
{
    if (!c) return;
    cudaSetDevice(c->device);
    sync_all(c);
    for (int i = 0; i < kMaxOverlap; ++i) {
        cudaFree(c->d_scratch[i]); cudaFree(c->d_slot_counters[i]);
        if (c->ev_trace_done[i]) cudaEventDestroy(c->ev_trace_done[i]);
        if (c->ev_blend_done[i]) cudaEventDestroy(c->ev_blend_done[i]);
        if (c->trace_stream[i]) cudaStreamDestroy(c->trace_stream[i]);
    }
    for (int i = 0; i < kBatchSets; ++i) {
        cudaFree(c->d_batch_scratch[i]); cudaFree(c->d_batch_counters[i]);
        if (c->ev_batch_blend[i]) cudaEventDestroy(c->ev_batch_blend[i]);
    }
    exchange_close(c);
    if (c->gl_result) cudaGraphicsUnregisterResource(c->gl_result);
    cudaFree(c->d_xch_blocks);
    if (c->blend_stream) cudaStreamDestroy(c->blend_stream);
    if (c->ev_inputs) cudaEventDestroy(c->ev_inputs);
    cudaFree(c->d_rct);
    cudaFree(c->d_objects); cudaFree(c->d_block); cudaFree(c->d_env_faces); cudaFree(c->d_env);
    cudaFree(c->d_image); cudaFree(c->d_counters); cudaFree(c->d_stats);
    if (c->copy_stream) { cudaStreamSynchronize(c->copy_stream); cudaStreamDestroy(c->copy_stream); }
    for (int i = 0; i < 2; ++i) { cudaFree(c->d_stage[i]); if (c->ev_snap[i]) cudaEventDestroy(c->ev_snap[i]); if (c->ev_copied[i]) cudaEventDestroy(c->ev_copied[i]); }
    for (int i = 0; i < 8; ++i) if (c->kt_b[i]) cudaEventDestroy(c->kt_b[i]);
    cudaFree(c->d_kt); if (c->h_kt) cudaFreeHost(c->h_kt);
    cudaFree(c->d_done);
    if (c->ev0) cudaEventDestroy(c->ev0);
    if (c->ev1) cudaEventDestroy(c->ev1);
    if (c->own_stream) cudaStreamDestroy(c->stream);
    delete c;
}

int ptb_set_size(ptb_ctx* c, int width, int height)
{
    if (!c) return fail(PTB_E_INVALID, "ctx is null");
    if (width <= 0 || height <= 0 || width > 65535 || height > 65535) return fail(PTB_E_INVALID, "bad size %dx%d (1..65535)", width, height);
    { const int rc = sync_all(c); if (rc != PTB_OK) return rc; }
    exchange_close(c);                               // the slots were sized for the old image: a fresh ptb_exchange_init is required
    if (c->gl_result) { cudaGraphicsUnregisterResource(c->gl_result); c->gl_result = nullptr; }     // the host re-allocates Result (PathTracer.cs:134) and registers it again
    c->width = width; c->height = height;
    c->frame = 0;                                    // PathTracer.cs:133
    return alloc_image(c);
}

int ptb_reset(ptb_ctx* c)
{
    if (!c) return fail(PTB_E_INVALID, "ctx is null");
    c->frame = 0;
    return PTB_OK;
}

int ptb_set_ray_depth(ptb_ctx* c, int v) { if (!c) return fail(PTB_E_INVALID, "ctx is null"); if (v < 0 || v > 4095) return fail(PTB_E_INVALID, "rayDepth %d outside [0,4095]", v); c->ray_depth = v; return PTB_OK; }
int ptb_set_spp(ptb_ctx* c, int v) { if (!c) return fail(PTB_E_INVALID, "ctx is null"); if (v < 1 || v > 4095) return fail(PTB_E_INVALID, "SPP %d outside [1,4095]", v); c->spp = v; return PTB_OK; }
int ptb_set_focal_length(ptb_ctx* c, float v) { if (!c) return fail(PTB_E_INVALID, "ctx is null"); c->focal_length = v; return PTB_OK; }
int ptb_set_aperture_diameter(ptb_ctx* c, float v) { if (!c) return fail(PTB_E_INVALID, "ctx is null"); c->aperture_diameter = v; return PTB_OK; }
int ptb_set_num_spheres(ptb_ctx* c, int n)
{
    if (!c) return fail(PTB_E_INVALID, "ctx is null");
    if (n < 0 || n > c->max_spheres) return fail(PTB_E_INVALID, "NumSpheres %d outside [0,%d]", n, c->max_spheres);
    if (n != c->n_spheres) { c->n_spheres = n; c->scene_dirty = true; }
    return PTB_OK;
}
int ptb_set_num_cuboids(ptb_ctx* c, int n)
{
    if (!c) return fail(PTB_E_INVALID, "ctx is null");
    if (n < 0 || n > c->max_cuboids) return fail(PTB_E_INVALID, "NumCuboids %d outside [0,%d]", n, c->max_cuboids);
    if (n != c->n_cuboids) { c->n_cuboids = n; c->scene_dirty = true; }
    return PTB_OK;
}

int ptb_basic_data_subdata(ptb_ctx* c, int offset, int size, const void* data)
{
    if (!c || !data) return fail(PTB_E_INVALID, "null argument");
    if (offset < 0 || size < 0 || offset + size > kBasicBytes) return fail(PTB_E_INVALID, "BasicDataUBO SubData [%d,%d) outside 144 bytes", offset, offset + size);
    memcpy(c->basic + offset, data, (size_t)size);
    return PTB_OK;
}

int ptb_game_objects_subdata(ptb_ctx* c, int offset, int size, const void* data)
{
    if (!c || !data) return fail(PTB_E_INVALID, "null argument");
    const size_t cap = (size_t)c->max_spheres * kSphereStride + (size_t)c->max_cuboids * kCuboidStride;
    if (offset < 0 || size < 0 || (size_t)offset + (size_t)size > cap) return fail(PTB_E_INVALID, "GameObjectsUBO SubData [%d,%d) outside %zu bytes", offset, offset + size, cap);
    memcpy(c->objects.data() + offset, data, (size_t)size);
    c->scene_dirty = true;
    return PTB_OK;
}

static int install_environment(ptb_ctx* c, int N)
{
    // d_env_faces already holds 6*N*N texels on the stream; (re)build the padded copy
    const int P = N + 2;
    const size_t padded = (size_t)6 * P * P * sizeof(float4);
    if (c->env_padded_bytes != padded) {
        if (c->d_env) { CU(cudaFree(c->d_env)); c->d_env = nullptr; c->env_padded_bytes = 0; }
        CU(cudaMalloc(&c->d_env, padded));
        c->env_padded_bytes = padded;
    }
    dim3 block(16, 16, 1), grid((P + 15) / 16, (P + 15) / 16, 6);
    pad_cubemap_kernel<<<grid, block, 0, c->stream>>>(c->d_env_faces, N, c->d_env);
    c->launches++;
    CU(cudaGetLastError());
    c->env_size = N;
    return mark_inputs(c);
}

int ptb_set_environment_rgba32f(ptb_ctx* c, int face_size, const float* six_faces)
{
    if (!c || !six_faces) return fail(PTB_E_INVALID, "null argument");
    if (face_size < 1 || face_size > 8192) return fail(PTB_E_INVALID, "face size %d outside [1,8192]", face_size);
    CU(cudaSetDevice(c->device));
    { const int rc = sync_all(c); if (rc != PTB_OK) return rc; }      // frames in flight on the trace streams still sample the old map
    const size_t bytes = (size_t)6 * face_size * face_size * sizeof(float4);
    if (c->d_env_faces) { CU(cudaFree(c->d_env_faces)); c->d_env_faces = nullptr; c->env_faces_bytes = 0; }
    CU(cudaMalloc(&c->d_env_faces, bytes));
    c->env_faces_bytes = bytes;
    CU(cudaMemcpyAsync(c->d_env_faces, six_faces, bytes, cudaMemcpyHostToDevice, c->stream));
    return install_environment(c, face_size);
}

int ptb_set_environment_srgb8(ptb_ctx* c, int face_size, const unsigned char* six_faces)
{
    if (!c || !six_faces) return fail(PTB_E_INVALID, "null argument");
    if (face_size < 1 || face_size > 8192) return fail(PTB_E_INVALID, "face size %d outside [1,8192]", face_size);
    CU(cudaSetDevice(c->device));
    { const int rc = sync_all(c); if (rc != PTB_OK) return rc; }
    const size_t n = (size_t)6 * face_size * face_size;
    if (c->d_env_faces) { CU(cudaFree(c->d_env_faces)); c->d_env_faces = nullptr; c->env_faces_bytes = 0; }
    CU(cudaMalloc(&c->d_env_faces, n * sizeof(float4)));
    c->env_faces_bytes = n * sizeof(float4);
    uchar4* d_raw = nullptr;
    CU(cudaMalloc(&d_raw, n * sizeof(uchar4)));
    int rc = PTB_OK;
    if (cudaMemcpyAsync(d_raw, six_faces, n * sizeof(uchar4), cudaMemcpyHostToDevice, c->stream) != cudaSuccess) rc = fail(PTB_E_CUDA, "H2D failed");
    if (rc == PTB_OK) {
        srgb8_decode_kernel<<<(unsigned)((n + 255) / 256), 256, 0, c->stream>>>(d_raw, n, c->d_env_faces);
        c->launches++;
        if (cudaGetLastError() != cudaSuccess || cudaStreamSynchronize(c->stream) != cudaSuccess) rc = fail(PTB_E_CUDA, "sRGB decode failed");
    }
    cudaFree(d_raw);
    if (rc != PTB_OK) return rc;
    return install_environment(c, face_size);
}

static int generate_atmosphere(ptb_ctx* c, int face_size, const void* ubo, int ubo_size, const float* light_pos, float light_intensity, int i_steps, int j_steps, bool fast)
{
    if (!c || !ubo || !light_pos) return fail(PTB_E_INVALID, "null argument");
    if (face_size < 1 || face_size > 8192) return fail(PTB_E_INVALID, "face size %d outside [1,8192]", face_size);
    if (ubo_size < 448) return fail(PTB_E_INVALID, "AtmosphericDataUBO needs 448 bytes, got %d", ubo_size);
    if (i_steps < 0 || j_steps < 0) return fail(PTB_E_INVALID, "negative step count");
    CU(cudaSetDevice(c->device));
    { const int rc = sync_all(c); if (rc != PTB_OK) return rc; }      // frames in flight still sample the old map
    const size_t bytes = (size_t)6 * face_size * face_size * sizeof(float4);
    if (c->env_faces_bytes != bytes) {      // a slider tick regenerates at the same size: keep the buffers (cudaMalloc of a 2048^2 map costs more than the kernel)
        if (c->d_env_faces) { CU(cudaFree(c->d_env_faces)); c->d_env_faces = nullptr; c->env_faces_bytes = 0; }
        CU(cudaMalloc(&c->d_env_faces, bytes));
        c->env_faces_bytes = bytes;
    }
    AtmosParams A;
    memcpy(A.ubo, ubo, 448);
    A.light[0] = light_pos[0]; A.light[1] = light_pos[1]; A.light[2] = light_pos[2];
    A.intensity = light_intensity < 0.0f ? 0.0f : light_intensity;   // AtmosphericScatterer.cs:52
    A.i_steps = i_steps; A.j_steps = j_steps; A.size = face_size;
    if (fast) {
        const unsigned n = 6u * (unsigned)face_size * (unsigned)face_size;
        atmosphere_fast_kernel<<<(n + 255u) / 256u, 256, 0, c->stream>>>(A, c->d_env_faces);
        c->launches++;
    } else {
        dim3 block(8, 8, 1), grid((face_size + 7) / 8, (face_size + 7) / 8, 6);   // AtmosphericScatterer.cs:109
        atmosphere_kernel<<<grid, block, 0, c->stream>>>(A, c->d_env_faces);
        c->launches++;
    }
    CU(cudaGetLastError());
    return install_environment(c, face_size);
}
int ptb_generate_atmosphere(ptb_ctx* c, int face_size, const void* ubo, int ubo_size, const float* light_pos, float light_intensity, int i_steps, int j_steps)
{
    return generate_atmosphere(c, face_size, ubo, ubo_size, light_pos, light_intensity, i_steps, j_steps, false);
}
int ptb_generate_atmosphere_fast(ptb_ctx* c, int face_size, const void* ubo, int ubo_size, const float* light_pos, float light_intensity, int i_steps, int j_steps)
{
    return generate_atmosphere(c, face_size, ubo, ubo_size, light_pos, light_intensity, i_steps, j_steps, true);
}

int ptb_read_environment(ptb_ctx* c, float* six_faces)
{
    if (!c || !six_faces) return fail(PTB_E_INVALID, "null argument");
    if (!c->d_env_faces) return fail(PTB_E_STATE, "no environment map set");
    CU(cudaMemcpyAsync(six_faces, c->d_env_faces, (size_t)6 * c->env_size * c->env_size * sizeof(float4), cudaMemcpyDeviceToHost, c->stream));
    CU(cudaStreamSynchronize(c->stream));
    return PTB_OK;
}
int ptb_environment_size(ptb_ctx* c) { return c ? c->env_size : fail(PTB_E_INVALID, "ctx is null"); }

int ptb_render_frames(ptb_ctx* c, int n)
{
    if (!c) return fail(PTB_E_INVALID, "ctx is null");
    if (n < 0) return fail(PTB_E_INVALID, "n < 0");
    if (!c->d_env) return fail(PTB_E_STATE, "Render() before an EnvironmentMap was set");
    CU(cudaSetDevice(c->device));
    if (!c->scene_dirty && (c->n_nodes > 0 || c->n_unbounded > 0 || c->grid_on)) {
        // the BVH margins were sized for an extent D that includes the camera: a camera that left it forces a rebuild
        if (required_extent(c, c->bvh_extent) > c->bvh_D) c->scene_dirty = true;
    }
    int rc = sync_scene(c);
    if (rc != PTB_OK) return rc;
    // fused exchange: the blend of frame k waits until frame k - slots was released, and rank 0's releases can only be enqueued
    // after this call has returned (its stream waits for every blend): more un-released frames than slots would wait on itself
    if (c->xch_on && c->local_rows > 0) {
        const unsigned long long mine = xch_own_frames(c, c->xch_seq + (unsigned long long)n) - c->xch_released;     // frames this rank would be holding
        if (mine > (unsigned long long)c->xch_slots)
            return fail(PTB_E_STATE, "fused exchange: %d frame(s) requested would leave %llu un-released frames on this root with %d slot(s); acquire/release "
                                     "between calls (TiledPathTracer.render does) or raise the slot count", n, mine, c->xch_slots);
    }
    CU(cudaEventRecord(c->ev0, c->stream));
    for (int i = 0; i < n;) {
        // fused exchange: blend j of a batch waits for the release of frame (seq_j - slots), and rank 0 enqueues this batch's
        // acquires / releases only after the whole batch (its stream waits for the batch's last blend) — so a batch may not be
        // longer than the slot ring, or its own later blends would wait for its own earlier releases
        const int bmax = c->xch_on ? std::min(c->batch, c->xch_slots * c->xch_roots) : c->batch;
        const int b = batch_eligible(c) ? std::min(n - i, bmax) : 1;
        rc = b >= 2 ? launch_batch(c, b) : launch_frame(c);
        if (rc != PTB_OK) return rc;
        i += b >= 2 ? b : 1;
    }
    CU(cudaEventRecord(c->ev1, c->stream));
    c->timed = true;
    return PTB_OK;
}
int ptb_render(ptb_ctx* c) { return ptb_render_frames(c, 1); }

int ptb_samples(ptb_ctx* c) { return c ? c->frame * c->spp : fail(PTB_E_INVALID, "ctx is null"); }
int ptb_frame(ptb_ctx* c) { return c ? c->frame : fail(PTB_E_INVALID, "ctx is null"); }
int ptb_set_frame(ptb_ctx* c, int frame)
{
    if (!c) return fail(PTB_E_INVALID, "ctx is null");
    if (frame < 0) return fail(PTB_E_INVALID, "frame < 0");
    c->frame = frame;
    return PTB_OK;
}

int ptb_read_result_async(ptb_ctx* c, float* dst) { return ptb_read_result_format_async(c, PTB_FORMAT_RGBA32F, dst); }

// Pipelined read-back, shared by the compact and the scattered form.  The image is accumulated in place, so the next
// Render() would race with a slow PCIe copy: snapshot it on the render stream (HBM -> HBM; for RGB32F / RGBA8 the snapshot
// is the packing / tone-map kernel), then let the copy stream move the snapshot to the host while the next frame renders.
static int read_result_impl(ptb_ctx* c, int format, void* dst, bool scatter)
{
    if (!c || !dst) return fail(PTB_E_INVALID, "null argument");
    if (format != PTB_FORMAT_RGBA32F && format != PTB_FORMAT_RGB32F && format != PTB_FORMAT_RGBA8) return fail(PTB_E_INVALID, "unknown read-back format %d", format);
    CU(cudaSetDevice(c->device));
    const size_t n_pixels = (size_t)c->local_rows * c->width;
    const size_t bpp = format == PTB_FORMAT_RGBA32F ? 16 : (format == PTB_FORMAT_RGB32F ? 12 : 4);
    const size_t bytes = n_pixels * sizeof(float4);                  // staging buffers hold the largest format
    if (bytes == 0) return PTB_OK;
    if (!c->copy_stream) {
        CU(cudaStreamCreateWithFlags(&c->copy_stream, cudaStreamNonBlocking));
        for (int i = 0; i < 2; ++i) {
            CU(cudaEventCreateWithFlags(&c->ev_snap[i], cudaEventDisableTiming));
            CU(cudaEventCreateWithFlags(&c->ev_copied[i], cudaEventDisableTiming));
        }
    }
    if (c->readback_bytes < bytes) {
        CU(cudaStreamSynchronize(c->copy_stream));
        for (int i = 0; i < 2; ++i) { if (c->d_stage[i]) CU(cudaFree(c->d_stage[i])); c->d_stage[i] = nullptr; CU(cudaMalloc(&c->d_stage[i], bytes)); }
        c->readback_bytes = bytes;
    }
    const int k = c->stage_next;
    c->stage_next ^= 1;
    CU(cudaStreamWaitEvent(c->stream, c->ev_copied[k], 0));       // the D2H that last used this staging buffer is done
    if (format == PTB_FORMAT_RGBA32F) {
        CU(cudaMemcpyAsync(c->d_stage[k], c->d_image, bytes, cudaMemcpyDeviceToDevice, c->stream));
    } else if (format == PTB_FORMAT_RGB32F) {
        const size_t quads = (n_pixels + 3) / 4;
        pack_rgb_kernel<<<(unsigned)((quads + 255) / 256), 256, 0, c->stream>>>(c->d_image, n_pixels, reinterpret_cast<float*>(c->d_stage[k]));
        c->launches++;
        CU(cudaGetLastError());
    } else {
        tonemap_kernel<<<(unsigned)((n_pixels + 255) / 256), 256, 0, c->stream>>>(c->d_image, n_pixels, reinterpret_cast<uchar4*>(c->d_stage[k]));
        c->launches++;
        CU(cudaGetLastError());
    }
    CU(cudaEventRecord(c->ev_snap[k], c->stream));
    CU(cudaStreamWaitEvent(c->copy_stream, c->ev_snap[k], 0));
    const unsigned char* src = reinterpret_cast<const unsigned char*>(c->d_stage[k]);
    unsigned char* out = static_cast<unsigned char*>(dst);
    if (!scatter || c->world == 1) {
        CU(cudaMemcpyAsync(out, src, n_pixels * bpp, cudaMemcpyDeviceToHost, c->copy_stream));
    } else {
        // local stripe j (compact, stripe_rows rows each) is global stripe rank + j * world; only the globally last stripe can be ragged
        const size_t stripe_bytes = (size_t)c->stripe_rows * c->width * bpp;
        const size_t n_full = (size_t)c->local_rows / c->stripe_rows;
        const size_t tail_rows = (size_t)c->local_rows % c->stripe_rows;
        if (n_full)
            CU(cudaMemcpy2DAsync(out + (size_t)c->rank * stripe_bytes, (size_t)c->world * stripe_bytes, src, stripe_bytes, stripe_bytes, n_full,
                                 cudaMemcpyDeviceToHost, c->copy_stream));
        if (tail_rows)
            CU(cudaMemcpyAsync(out + ((size_t)c->rank + n_full * c->world) * stripe_bytes, src + n_full * stripe_bytes, tail_rows * c->width * bpp,
                               cudaMemcpyDeviceToHost, c->copy_stream));
    }
    CU(cudaEventRecord(c->ev_copied[k], c->copy_stream));
    return PTB_OK;
}
int ptb_read_result_format_async(ptb_ctx* c, int format, void* dst) { return read_result_impl(c, format, dst, false); }
int ptb_read_result_scatter_async(ptb_ctx* c, int format, void* full_frame) { return read_result_impl(c, format, full_frame, true); }
int ptb_read_result(ptb_ctx* c, float* dst)
{
    if (!c || !dst) return fail(PTB_E_INVALID, "null argument");
    CU(cudaMemcpyAsync(dst, c->d_image, (size_t)c->local_rows * c->width * sizeof(float4), cudaMemcpyDeviceToHost, c->stream));
    CU(cudaStreamSynchronize(c->stream));
    return PTB_OK;
}
int ptb_write_result(ptb_ctx* c, const float* src)
{
    if (!c || !src) return fail(PTB_E_INVALID, "null argument");
    CU(cudaMemcpyAsync(c->d_image, src, (size_t)c->local_rows * c->width * sizeof(float4), cudaMemcpyHostToDevice, c->stream));
    CU(cudaStreamSynchronize(c->stream));
    return PTB_OK;
}
static int run_tonemap(ptb_ctx* c, uchar4* d_out)
{
    const size_t n = (size_t)c->local_rows * c->width;
    if (n == 0) return PTB_OK;
    tonemap_kernel<<<(unsigned)((n + 255) / 256), 256, 0, c->stream>>>(c->d_image, n, d_out);
    c->launches++;
    CU(cudaGetLastError());
    return PTB_OK;
}
int ptb_tonemap_device(ptb_ctx* c, void* rgba8_device)
{
    if (!c || !rgba8_device) return fail(PTB_E_INVALID, "null argument");
    return run_tonemap(c, (uchar4*)rgba8_device);
}
int ptb_tonemap_rgba8(ptb_ctx* c, unsigned char* rgba8)
{
    if (!c || !rgba8) return fail(PTB_E_INVALID, "null argument");
    const size_t n = (size_t)c->local_rows * c->width;
    uchar4* d = nullptr;
    CU(cudaMalloc(&d, (n ? n : 1) * sizeof(uchar4)));
    int rc = run_tonemap(c, d);
    if (rc == PTB_OK && cudaMemcpyAsync(rgba8, d, n * sizeof(uchar4), cudaMemcpyDeviceToHost, c->stream) != cudaSuccess) rc = fail(PTB_E_CUDA, "D2H failed");
    if (rc == PTB_OK && cudaStreamSynchronize(c->stream) != cudaSuccess) rc = fail(PTB_E_CUDA, "tone-map kernel failed");
    cudaFree(d);
    return rc;
}

// ---- PathTracer.Result as a GL texture (PathTracer.cs:86,97-99: an Rgba32f TEXTURE_2D that ScreenEffect.Render samples,
//      MainWindow.cs:51).  The host registers that texture once; ptb_present_gl copies the accumulation image into it on the
//      device (HBM -> the texture's array, no PCIe traffic), stream-ordered after the frames rendered so far.
int ptb_register_gl_texture(ptb_ctx* c, unsigned int texture)
{
    if (!c) return fail(PTB_E_INVALID, "ctx is null");
    if (c->world != 1) return fail(PTB_E_STATE, "a tile context (ptb_set_tile) holds only its stripes; register the texture on a full-frame context");
    CU(cudaSetDevice(c->device));
    if (c->gl_result) { cudaGraphicsUnregisterResource(c->gl_result); c->gl_result = nullptr; }
    const cudaError_t e = cudaGraphicsGLRegisterImage(&c->gl_result, texture, 0x0DE1u /* GL_TEXTURE_2D */, cudaGraphicsRegisterFlagsWriteDiscard);
    if (e != cudaSuccess) {
        c->gl_result = nullptr;
        cudaGetLastError();
        return fail(PTB_E_STATE, "cudaGraphicsGLRegisterImage(texture %u) failed: %s — the calling thread needs a current OpenGL context on this GPU "
                                 "and an Rgba32f TEXTURE_2D of the render size", texture, cudaGetErrorString(e));
    }
    return PTB_OK;
}
int ptb_present_gl(ptb_ctx* c)
{
    if (!c) return fail(PTB_E_INVALID, "ctx is null");
    if (!c->gl_result) return fail(PTB_E_STATE, "no GL texture registered (ptb_register_gl_texture)");
    CU(cudaSetDevice(c->device));
    CU(cudaGraphicsMapResources(1, &c->gl_result, c->stream));
    cudaArray_t arr = nullptr;
    cudaError_t e = cudaGraphicsSubResourceGetMappedArray(&arr, c->gl_result, 0, 0);
    if (e == cudaSuccess) {
        const size_t row = (size_t)c->width * sizeof(float4);
        e = cudaMemcpy2DToArrayAsync(arr, 0, 0, c->d_image, row, row, (size_t)c->height, cudaMemcpyDeviceToDevice, c->stream);
    }
    const cudaError_t u = cudaGraphicsUnmapResources(1, &c->gl_result, c->stream);      // GL may sample the texture after this, in stream order
    if (e != cudaSuccess) return fail(PTB_E_CUDA, "copy into the GL texture failed: %s (is it %dx%d Rgba32f?)", cudaGetErrorString(e), c->width, c->height);
    if (u != cudaSuccess) return fail(PTB_E_CUDA, "cudaGraphicsUnmapResources failed: %s", cudaGetErrorString(u));
    return PTB_OK;
}
int ptb_unregister_gl_texture(ptb_ctx* c)
{
    if (!c) return fail(PTB_E_INVALID, "ctx is null");
    if (c->gl_result) { CU(cudaGraphicsUnregisterResource(c->gl_result)); c->gl_result = nullptr; }
    return PTB_OK;
}

int ptb_synchronize(ptb_ctx* c)
{
    if (!c) return fail(PTB_E_INVALID, "ctx is null");
    CU(cudaStreamSynchronize(c->stream));
    if (c->copy_stream) CU(cudaStreamSynchronize(c->copy_stream));
    if (c->d_done) {       // the batch blends' sticky flag: a blend that gave up waiting for its trace
        unsigned gave_up = 0;
        CU(cudaMemcpy(&gave_up, c->d_done + kBatchSets, sizeof gave_up, cudaMemcpyDeviceToHost));
        if (gave_up) return fail(PTB_E_STATE, "a batch blend kernel gave up waiting for its trace (ten-minute time-out): the image is incomplete");
    }
    return PTB_OK;
}

int ptb_result_device_ptr(ptb_ctx* c, void** p, size_t* bytes)
{
    if (!c || !p) return fail(PTB_E_INVALID, "null argument");
    *p = c->d_image;
    if (bytes) *bytes = c->image_bytes;
    return PTB_OK;
}
int ptb_set_stream(ptb_ctx* c, void* s)
{
    if (!c) return fail(PTB_E_INVALID, "ctx is null");
    { const int rc = sync_all(c); if (rc != PTB_OK) return rc; }
    if (c->own_stream) cudaStreamDestroy(c->stream);
    c->stream = (cudaStream_t)s;
    c->own_stream = false;
    return mark_inputs(c);
}
int ptb_width(ptb_ctx* c) { return c ? c->width : fail(PTB_E_INVALID, "ctx is null"); }
int ptb_height(ptb_ctx* c) { return c ? c->height : fail(PTB_E_INVALID, "ctx is null"); }

int ptb_set_tile(ptb_ctx* c, int rank, int world, int stripe_rows)
{
    if (!c) return fail(PTB_E_INVALID, "ctx is null");
    if (world < 1 || rank < 0 || rank >= world || stripe_rows < 1) return fail(PTB_E_INVALID, "bad tile rank=%d world=%d stripe_rows=%d", rank, world, stripe_rows);
    { const int rc = sync_all(c); if (rc != PTB_OK) return rc; }
    exchange_close(c);                               // arrival targets and row mapping depend on the partition
    c->rank = rank; c->world = world; c->stripe_rows = stripe_rows;
    c->frame = 0;
    return alloc_image(c);
}
int ptb_local_rows(ptb_ctx* c) { return c ? c->local_rows : fail(PTB_E_INVALID, "ctx is null"); }
int ptb_max_local_rows(ptb_ctx* c) { return c ? compute_local_rows(c->height, 0, c->world, c->stripe_rows) : fail(PTB_E_INVALID, "ctx is null"); }

int ptb_deinterleave_device(ptb_ctx* c, const void* gathered, void* full)
{
    if (!c || !gathered || !full) return fail(PTB_E_INVALID, "null argument");
    dim3 block(256, 1, 1), grid((c->width + 255) / 256, c->height, 1);
    deinterleave_kernel<<<grid, block, 0, c->stream>>>((const float4*)gathered, (float4*)full, c->width, c->height, c->world, c->stripe_rows,
                                                       compute_local_rows(c->height, 0, c->world, c->stripe_rows));
    c->launches++;
    CU(cudaGetLastError());
    return PTB_OK;
}

// Stream memory operations of the driver API (no libcuda link dependency: resolved through the runtime).  With them rank 0's
// acquire / release are commands of the stream's front end, not kernels — a one-thread kernel would have to wait for a CTA
// slot, and a persistent grid frees none until it drains.
typedef int (*StreamWaitValue32)(cudaStream_t, unsigned long long, unsigned, unsigned);
typedef int (*StreamWriteValue32)(cudaStream_t, unsigned long long, unsigned, unsigned);
static StreamWaitValue32 g_wait32 = nullptr;
static StreamWriteValue32 g_write32 = nullptr;
static bool g_memops_probed = false;
static bool stream_memops()
{
    if (!g_memops_probed) {
        g_memops_probed = true;
        if (getenv("PTB_NO_MEMOPS")) return false;
        void *w = nullptr, *r = nullptr;
        cudaDriverEntryPointQueryResult q1, q2;
        if (cudaGetDriverEntryPoint("cuStreamWaitValue32", &w, cudaEnableDefault, &q1) == cudaSuccess && q1 == cudaDriverEntryPointSuccess &&
            cudaGetDriverEntryPoint("cuStreamWriteValue32", &r, cudaEnableDefault, &q2) == cudaSuccess && q2 == cudaDriverEntryPointSuccess) {
            g_wait32 = (StreamWaitValue32)w;
            g_write32 = (StreamWriteValue32)r;
        }
        cudaGetLastError();
    }
    return g_wait32 && g_write32;
}

static int exchange_close(ptb_ctx* c)
{
    for (int r = 0; r < 16; ++r) {
        if (c->xch_peer[r] && c->xch_peer_mapped[r]) cudaIpcCloseMemHandle(c->xch_peer[r]);
        c->xch_peer[r] = nullptr; c->xch_peer_mapped[r] = false;
    }
    if (c->xch_block) {
        if (c->xch_mapped) cudaIpcCloseMemHandle(c->xch_block);
        else cudaFree(c->xch_block);
    }
    c->xch_block = nullptr;
    c->xch_on = false;
    c->xch_mapped = false;
    c->xch_roots = 1;
    return PTB_OK;
}

int ptb_exchange_init(ptb_ctx* c, int slots) { return ptb_exchange_init_roots(c, slots, PTB_FORMAT_RGBA32F, 0); }
int ptb_exchange_init_format(ptb_ctx* c, int slots, int format) { return ptb_exchange_init_roots(c, slots, format, 0); }
int ptb_exchange_init_roots(ptb_ctx* c, int slots, int format, int rotate)
{
    if (!c) return fail(PTB_E_INVALID, "ctx is null");
    if (format != PTB_FORMAT_RGBA32F && format != PTB_FORMAT_RGB32F) return fail(PTB_E_INVALID, "the exchange ships RGBA32F or RGB32F, not format %d", format);
    if (slots < 1 || slots > kMaxSlots) return fail(PTB_E_INVALID, "slots %d outside [1,%d]", slots, kMaxSlots);
    if (rotate && c->world > 16) return fail(PTB_E_INVALID, "rotating roots support up to 16 ranks");
    if (c->overlap < 2) return fail(PTB_E_STATE, "the fused exchange needs the pipelined mode (ptb_set_overlap >= 2)");
    { const int rc = sync_all(c); if (rc != PTB_OK) return rc; }
    exchange_close(c);
    c->xch_slots = slots;
    c->xch_roots = (rotate && c->world > 1) ? c->world : 1;
    c->xch_rgb = format == PTB_FORMAT_RGB32F;
    c->xch_image_bytes = (((size_t)c->width * c->height * (c->xch_rgb ? 12 : 16)) + 255) & ~(size_t)255;
    c->xch_seq = c->xch_acquired = c->xch_released = 0;
    if (!c->d_xch_blocks) { CU(cudaMalloc(&c->d_xch_blocks, sizeof(unsigned int))); }
    CU(cudaMemset(c->d_xch_blocks, 0, sizeof(unsigned int)));
    CU(cudaDeviceSynchronize());
    if (c->rank == 0 || c->xch_roots > 1) {
        const size_t bytes = 4096 + (size_t)slots * c->xch_image_bytes;
        CU(cudaMalloc(&c->xch_block, bytes));
        CU(cudaMemset(c->xch_block, 0, bytes));
        CU(cudaDeviceSynchronize());      // cudaMemset on device memory may return before it ran; peers must never see it land late
        if (c->xch_roots > 1) c->xch_peer[c->rank] = c->xch_block;
        c->xch_on = c->xch_roots == 1;    // rotating roots: on once every peer's block is attached
    }
    return PTB_OK;
}
int ptb_exchange_handle(ptb_ctx* c, void* handle64)
{
    if (!c || !handle64) return fail(PTB_E_INVALID, "null argument");
    if (!c->xch_block || c->xch_mapped) return fail(PTB_E_STATE, "only a root exports its exchange block (rank 0, or every rank with rotating roots), after ptb_exchange_init");
    static_assert(sizeof(cudaIpcMemHandle_t) == 64, "IPC handle size");
    cudaIpcMemHandle_t h;
    CU(cudaIpcGetMemHandle(&h, c->xch_block));
    memcpy(handle64, &h, 64);
    return PTB_OK;
}
int ptb_exchange_attach(ptb_ctx* c, const void* handle64)
{
    if (!c || !handle64) return fail(PTB_E_INVALID, "null argument");
    if (c->xch_roots > 1) return fail(PTB_E_STATE, "rotating roots: attach every peer with ptb_exchange_attach_peer");
    if (c->rank == 0) return fail(PTB_E_STATE, "rank 0 owns the exchange block");
    if (c->xch_slots < 1) return fail(PTB_E_STATE, "call ptb_exchange_init first");
    cudaIpcMemHandle_t h;
    memcpy(&h, handle64, 64);
    void* p = nullptr;
    CU(cudaIpcOpenMemHandle(&p, h, cudaIpcMemLazyEnablePeerAccess));
    c->xch_block = (unsigned char*)p;
    c->xch_mapped = true;
    c->xch_on = true;
    return PTB_OK;
}
int ptb_exchange_attach_peer(ptb_ctx* c, int peer_rank, const void* handle64)
{
    if (!c || !handle64) return fail(PTB_E_INVALID, "null argument");
    if (c->xch_roots <= 1) return fail(PTB_E_STATE, "ptb_exchange_attach_peer belongs to rotating roots (ptb_exchange_init_roots(..., 1))");
    if (peer_rank < 0 || peer_rank >= c->world) return fail(PTB_E_INVALID, "peer rank %d outside [0,%d)", peer_rank, c->world);
    if (peer_rank != c->rank && !c->xch_peer[peer_rank]) {
        cudaIpcMemHandle_t h;
        memcpy(&h, handle64, 64);
        void* p = nullptr;
        CU(cudaIpcOpenMemHandle(&p, h, cudaIpcMemLazyEnablePeerAccess));
        c->xch_peer[peer_rank] = (unsigned char*)p;
        c->xch_peer_mapped[peer_rank] = true;
    }
    bool all = true;
    for (int r = 0; r < c->world; ++r) all = all && c->xch_peer[r] != nullptr;
    c->xch_on = all;
    return PTB_OK;
}
int ptb_exchange_root(ptb_ctx* c, long long frame_seq)
{
    if (!c) return fail(PTB_E_INVALID, "ctx is null");
    const unsigned long long q = frame_seq < 0 ? (c->xch_seq ? c->xch_seq - 1 : 0ull) : (unsigned long long)frame_seq;
    return c->xch_roots > 1 ? (int)(q % (unsigned long long)c->world) : 0;
}
int ptb_exchange_pending(ptb_ctx* c)
{
    if (!c) return fail(PTB_E_INVALID, "ctx is null");
    if (!c->xch_on || !c->xch_block || c->xch_mapped) return 0;
    return (int)(xch_own_frames(c, c->xch_seq) - c->xch_acquired);
}
int ptb_exchange_acquire(ptb_ctx* c, void** full_device)
{
    if (!c || !full_device) return fail(PTB_E_INVALID, "null argument");
    if (!c->xch_on || !c->xch_block || c->xch_mapped) return fail(PTB_E_STATE, "acquire is a root's side of an initialised exchange");
    if (c->xch_acquired >= xch_own_frames(c, c->xch_seq)) return fail(PTB_E_STATE, "no rendered frame left to acquire on this root");
    const int slot = (int)(c->xch_acquired % (unsigned long long)c->xch_slots);
    const unsigned target = (unsigned)((c->xch_acquired / (unsigned long long)c->xch_slots + 1) * (unsigned long long)c->world);
    ExchangeFlags* fl = reinterpret_cast<ExchangeFlags*>(c->xch_block);
    if (stream_memops() && g_wait32(c->stream, (unsigned long long)(uintptr_t)&fl->arrived[slot], target, 0x1 /* CU_STREAM_WAIT_VALUE_GEQ */) == 0) {
        // the stream's front end polls the arrival counter: no kernel, no CTA slot needed
    } else {
        exchange_acquire_kernel<<<1, 1, 0, c->stream>>>(fl, slot, target);
        CU(cudaGetLastError());
        c->launches++;
    }
    c->xch_acquired++;
    *full_device = c->xch_block + 4096 + (size_t)slot * c->xch_image_bytes;
    return PTB_OK;
}
int ptb_exchange_release(ptb_ctx* c)
{
    if (!c) return fail(PTB_E_INVALID, "ctx is null");
    if (!c->xch_on || !c->xch_block || c->xch_mapped) return fail(PTB_E_STATE, "release is a root's side of an initialised exchange");
    if (c->xch_released >= c->xch_acquired) return fail(PTB_E_STATE, "nothing acquired");
    c->xch_released++;
    ExchangeFlags* fl = reinterpret_cast<ExchangeFlags*>(c->xch_block);
    if (stream_memops() && g_write32(c->stream, (unsigned long long)(uintptr_t)&fl->consumed, (unsigned)c->xch_released, 0x0 /* default: ordered after prior work */) == 0) {
    } else {
        exchange_release_kernel<<<1, 1, 0, c->stream>>>(fl, (unsigned)c->xch_released);
        CU(cudaGetLastError());
        c->launches++;
    }
    return PTB_OK;
}
int ptb_exchange_status(ptb_ctx* c)
{
    if (!c) return fail(PTB_E_INVALID, "ctx is null");
    if (!c->xch_on) return PTB_OK;
    unsigned err = 0;
    if (c->xch_block) CU(cudaMemcpy(&err, c->xch_block + offsetof(ExchangeFlags, error), sizeof err, cudaMemcpyDeviceToHost));
    return err ? fail(PTB_E_STATE, "a fused-exchange wait timed out (a rank stalled or the consumer never released a slot)") : PTB_OK;
}

int ptb_set_kernel(ptb_ctx* c, int k)
{
    if (!c) return fail(PTB_E_INVALID, "ctx is null");
    if (k != PTB_KERNEL_MEGA && k != PTB_KERNEL_NAIVE) return fail(PTB_E_INVALID, "unknown kernel %d", k);
    { const int rc = sync_all(c); if (rc != PTB_OK) return rc; }
    c->kernel = k;
    return PTB_OK;
}
int ptb_set_overlap(ptb_ctx* c, int n)
{
    if (!c) return fail(PTB_E_INVALID, "ctx is null");
    if (n < 0 || n > kMaxOverlap) return fail(PTB_E_INVALID, "overlap %d outside [0,%d]", n, kMaxOverlap);
    { const int rc = sync_all(c); if (rc != PTB_OK) return rc; }
    c->overlap = n;
    c->launch_seq = 0;
    return PTB_OK;
}
int ptb_set_batch(ptb_ctx* c, int frames)
{
    if (!c) return fail(PTB_E_INVALID, "ctx is null");
    if (frames < 1 || frames > kMaxBatch) return fail(PTB_E_INVALID, "batch %d outside [1,%d]", frames, kMaxBatch);
    { const int rc = sync_all(c); if (rc != PTB_OK) return rc; }
    c->batch = frames;
    return PTB_OK;
}
int ptb_set_grid_divisor(ptb_ctx* c, int d)
{
    if (!c) return fail(PTB_E_INVALID, "ctx is null");
    if (d < 1 || d > 8) return fail(PTB_E_INVALID, "grid divisor %d outside [1,8]", d);
    { const int rc = sync_all(c); if (rc != PTB_OK) return rc; }
    c->grid_divisor = d;
    c->mega_smem_set = -1;           // recompute the grid at the next launch
    return PTB_OK;
}
int ptb_set_kernel_timing(ptb_ctx* c, int enabled)
{
    if (!c) return fail(PTB_E_INVALID, "ctx is null");
    { const int rc = sync_all(c); if (rc != PTB_OK) return rc; }
    for (int i = 0; i < 8; ++i) {
        if (enabled && !c->kt_b[i]) CU(cudaEventCreateWithFlags(&c->kt_b[i], cudaEventDisableTiming));
        if (enabled && !c->d_kt) {
            CU(cudaMalloc(&c->d_kt, 16 * sizeof(unsigned long long)));
            CU(cudaMallocHost(&c->h_kt, 18 * sizeof(unsigned long long)));
            c->h_kt[16] = ~0ull; c->h_kt[17] = 0ull;
        }
        c->kt_used[i] = false;
    }
    c->kt_on = enabled != 0;
    c->kt_ms = 0.0; c->kt_frames_total = 0; c->kt_launches = 0; c->kt_next = 0;
    return PTB_OK;
}
int ptb_kernel_time(ptb_ctx* c, double* ms_total, long long* frames, long long* launches)
{
    if (!c || !ms_total || !frames || !launches) return fail(PTB_E_INVALID, "null argument");
    for (int i = 0; i < 8; ++i) { const int rc = kt_collect(c, i); if (rc != PTB_OK) return rc; }
    *ms_total = c->kt_ms; *frames = c->kt_frames_total; *launches = c->kt_launches;
    c->kt_ms = 0.0; c->kt_frames_total = 0; c->kt_launches = 0;
    return PTB_OK;
}
int ptb_set_precision(ptb_ctx* c, int precision)
{
    if (!c) return fail(PTB_E_INVALID, "ctx is null");
    if (precision != PTB_PRECISION_EXACT && precision != PTB_PRECISION_FAST) return fail(PTB_E_INVALID, "unknown precision %d", precision);
    if (precision == PTB_PRECISION_FAST) {
        if (ptb_fast_api::params_size() != sizeof(RenderParams) || ptb_fast_api::threads() != kMegaThreads)
            return fail(PTB_E_STATE, "the fast translation unit was built with a different RenderParams layout");
        if (c->stats_on) return fail(PTB_E_STATE, "path statistics are collected by the exact build only (ptb_set_stats(0) first)");
    }
    { const int rc = sync_all(c); if (rc != PTB_OK) return rc; }
    c->precision = precision;
    return PTB_OK;
}
int ptb_precision(ptb_ctx* c) { return c ? c->precision : fail(PTB_E_INVALID, "ctx is null"); }
int ptb_set_ray_classification(ptb_ctx* c, int mode, int cells, int buckets)
{
    if (!c) return fail(PTB_E_INVALID, "ctx is null");
    if (mode < 0 || mode > 1) return fail(PTB_E_INVALID, "mode %d outside [0,1]", mode);
    if (cells < 1 || cells > 64 || buckets < 1 || buckets > 64) return fail(PTB_E_INVALID, "cells %d / buckets %d outside [1,64]", cells, buckets);
    c->rct_mode = mode; c->rct_cells = cells; c->rct_G = buckets;
    c->rct_geometry.clear();
    c->scene_dirty = true;
    return PTB_OK;
}
int ptb_set_grid_density(ptb_ctx* c, float cells_per_primitive)
{
    if (!c) return fail(PTB_E_INVALID, "ctx is null");
    if (!(cells_per_primitive >= 0.05f && cells_per_primitive <= 64.0f)) return fail(PTB_E_INVALID, "density %g outside [0.05,64]", (double)cells_per_primitive);
    c->grid_density = cells_per_primitive;
    c->scene_dirty = true;
    return PTB_OK;
}
int ptb_set_large_scene_mode(ptb_ctx* c, int mode)
{
    if (!c) return fail(PTB_E_INVALID, "ctx is null");
    if (mode != 0 && mode != 1) return fail(PTB_E_INVALID, "mode %d outside [0,1]", mode);
    c->large_mode = mode;
    c->scene_dirty = true;
    return PTB_OK;
}
int ptb_set_bvh_threshold(ptb_ctx* c, int primitives)
{
    if (!c) return fail(PTB_E_INVALID, "ctx is null");
    if (primitives < 1) return fail(PTB_E_INVALID, "threshold %d < 1", primitives);
    c->bvh_threshold = primitives;
    c->scene_dirty = true;
    return PTB_OK;
}
int ptb_kernel_launches(ptb_ctx* c) { return c ? c->launches : fail(PTB_E_INVALID, "ctx is null"); }
int ptb_scene_info(ptb_ctx* c, int what)
{
    if (!c) return fail(PTB_E_INVALID, "ctx is null");
    { const int rc = sync_scene(c); if (rc != PTB_OK) return rc; }
    switch (what) {
    case PTB_INFO_BVH_NODES: return c->n_nodes;
    case PTB_INFO_ALWAYS_TESTED: return c->n_unbounded;
    case PTB_INFO_STAGED_BYTES: return c->stage_bytes;
    case PTB_INFO_GRID_CTAS: return c->mega_grid;
    case PTB_INFO_FOLD: return fold_of(c);
    case PTB_INFO_GRID_CELLS: return c->grid_on ? c->grid_n[0] * c->grid_n[1] * c->grid_n[2] : 0;
    case PTB_INFO_GRID_ITEMS: return c->grid_on ? (int)c->grid_items.size() : 0;
    case PTB_INFO_RCT_KBYTES: return c->rct_on ? (int)(((size_t)c->rct_n[0] * c->rct_n[1] * c->rct_n[2] * 6 * c->rct_G * c->rct_G * 8) >> 10) : 0;
    default: return fail(PTB_E_INVALID, "unknown info %d", what);
    }
}
float ptb_last_render_ms(ptb_ctx* c)
{
    if (!c || !c->timed) return -1.0f;
    if (cudaEventSynchronize(c->ev1) != cudaSuccess) return -1.0f;
    float ms = -1.0f;
    if (cudaEventElapsedTime(&ms, c->ev0, c->ev1) != cudaSuccess) return -1.0f;
    return ms;
}
int ptb_set_stats(ptb_ctx* c, int enabled)
{
    if (!c) return fail(PTB_E_INVALID, "ctx is null");
    if (enabled && c->precision == PTB_PRECISION_FAST) return fail(PTB_E_STATE, "path statistics are collected by the exact build only (ptb_set_precision(PTB_PRECISION_EXACT) first)");
    c->stats_on = enabled != 0;
    CU(cudaMemsetAsync(c->d_stats, 0, 4 * sizeof(unsigned long long), c->stream));
    return mark_inputs(c);
}
int ptb_read_stats(ptb_ctx* c, unsigned long long* out3)
{
    if (!c || !out3) return fail(PTB_E_INVALID, "null argument");
    CU(cudaMemcpyAsync(out3, c->d_stats, 3 * sizeof(unsigned long long), cudaMemcpyDeviceToHost, c->stream));
    CU(cudaStreamSynchronize(c->stream));
    return PTB_OK;
}

int ptb_debug_eval(ptb_ctx* c, int op, const float* in, int n, float* out)
{
    if (!c || !in || !out || n < 0) return fail(PTB_E_INVALID, "bad argument");
    if (n == 0) return PTB_OK;
    CU(cudaSetDevice(c->device));
    size_t in_f = 0, out_f = 0;
    switch (op) {
    case 0: in_f = n; out_f = 2 * (size_t)n; break;
    case 1: in_f = n; out_f = n; break;
    case 2: in_f = 1; out_f = n; break;
    case 3: in_f = 3 * (size_t)n; out_f = 3 * (size_t)n; break;
    case 4: case 6: case 9: case 10: case 11: in_f = 6 * (size_t)n; out_f = 12 * (size_t)n; break;
    case 5: in_f = 2 * (size_t)n; out_f = 4 * (size_t)n; break;
    case 7: in_f = 6 * (size_t)n + 1; out_f = 12 * (size_t)n; break;
    case 8: in_f = n; out_f = n; break;
    default: return fail(PTB_E_INVALID, "unknown debug op %d", op);
    }
    float *d_in = nullptr, *d_out = nullptr;
    CU(cudaMalloc(&d_in, in_f * sizeof(float)));
    CU(cudaMalloc(&d_out, out_f * sizeof(float)));
    int rc = PTB_OK;
    do {
        if (cudaMemcpyAsync(d_in, in, in_f * sizeof(float), cudaMemcpyHostToDevice, c->stream) != cudaSuccess) { rc = fail(PTB_E_CUDA, "H2D failed"); break; }
        const int tb = 128, gb = (n + tb - 1) / tb;
        if (op == 0) dbg_sincos_kernel<<<gb, tb, 0, c->stream>>>(d_in, n, d_out);
        else if (op == 1) dbg_exp_kernel<<<gb, tb, 0, c->stream>>>(d_in, n, d_out);
        else if (op == 2) { uint32_t seed; memcpy(&seed, in, 4); dbg_pcg_kernel<<<1, 32, 0, c->stream>>>(seed, n, d_out); }
        else if (op == 3) {
            if (!c->d_env) { rc = fail(PTB_E_STATE, "no environment map set"); break; }
            dbg_env_kernel<<<gb, tb, 0, c->stream>>>(c->d_env, c->env_size, d_in, n, d_out);
        } else if (op == 4 || op == 6 || op == 9 || op == 10 || op == 11) {
            rc = sync_scene(c);
            if (rc != PTB_OK) break;
            RenderParams P;
            fill_params(c, P);
            const int smem = op == 6 ? 0 : c->stage_bytes;
            if (op == 9 && c->n_nodes == 0 && c->n_unbounded == 0) { rc = fail(PTB_E_STATE, "the current scene has no BVH (fewer than %d primitives)", c->bvh_threshold); break; }
            if (op == 9 && c->grid_on) { rc = fail(PTB_E_STATE, "the current scene is traced through the grid, not the BVH (ptb_set_large_scene_mode(0) selects the BVH)"); break; }
            if (op == 11 && !c->grid_on) { rc = fail(PTB_E_STATE, "the current scene has no grid (fewer than %d primitives, or the BVH was selected)", c->bvh_threshold); break; }
            if (op == 10 && !c->rct_on) { rc = fail(PTB_E_STATE, "the current scene has no ray-classification table (more than 64 primitives, or switched off)"); break; }
            if (cudaFuncSetAttribute(dbg_trace_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, c->stage_bytes) != cudaSuccess) { rc = fail(PTB_E_CUDA, "smem attr failed"); break; }
            dbg_trace_kernel<<<gb, tb, smem, c->stream>>>(P, d_in, n, d_out, op == 6 ? 1 : (op == 9 ? 2 : (op == 10 ? 3 : (op == 11 ? 4 : 0))));
        } else if (op == 5) dbg_arith_kernel<<<gb, tb, 0, c->stream>>>(d_in, n, d_out);
        else if (op == 8) dbg_log_kernel<<<gb, tb, 0, c->stream>>>(d_in, n, d_out);
        else if (op == 7) {
            rc = sync_scene(c);
            if (rc != PTB_OK) break;
            const int k = (int)in[6 * (size_t)n];
            if (k < 1 || k > 32) { rc = fail(PTB_E_INVALID, "group size %d outside [1,32]", k); break; }
            RenderParams P;
            fill_params(c, P);
            if (cudaFuncSetAttribute(dbg_group_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, c->stage_bytes) != cudaSuccess) { rc = fail(PTB_E_CUDA, "smem attr failed"); break; }
            dbg_group_kernel<<<(n + k - 1) / k, 32, c->stage_bytes, c->stream>>>(P, d_in, n, d_out, k);
        }
        c->launches++;
        if (cudaGetLastError() != cudaSuccess) { rc = fail(PTB_E_CUDA, "debug kernel launch failed"); break; }
        if (cudaMemcpyAsync(out, d_out, out_f * sizeof(float), cudaMemcpyDeviceToHost, c->stream) != cudaSuccess) { rc = fail(PTB_E_CUDA, "D2H failed"); break; }
        if (cudaStreamSynchronize(c->stream) != cudaSuccess) { rc = fail(PTB_E_CUDA, "debug kernel failed: %s", cudaGetErrorString(cudaGetLastError())); break; }
    } while (0);
    cudaFree(d_in);
    cudaFree(d_out);
    return rc;
}

} // extern "C"
