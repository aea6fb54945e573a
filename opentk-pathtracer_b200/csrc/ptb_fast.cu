// ptb_fast.cu — the megakernel of ptb_kernels.cuh compiled a second time with fast arithmetic (SURVEY.md §8c protocol P2):
// MUFU rcp / sqrt / rsqrt / sin / cos / ex2 (ptb_math.cuh under PTB_FAST) and -fmad=true, denormals flushed.  Same source,
// same RNG stream, same control flow as the exact build in ptb_abi.cu; only the rounding of the floating-point operations
// differs, which is what a GL driver's compiler is free to do with compute.glsl as well.  Selected at run time with
// ptb_set_precision(ctx, PTB_PRECISION_FAST); tests hold it to the exact build by the north star's tolerance.
// Everything lives in namespace ptb_fast so that the two translation units can be linked into one library.
#define PTB_FAST 1
#define PTB_MEGA_ONLY 1
#define ptb ptb_fast
#include "ptb_kernels.cuh"
#undef ptb
#include "ptb_fast.h"

#include <cstring>

namespace ptb_fast_api {

namespace {

template <int kFold, class F>
cudaError_t pick(int ring, bool batch, F&& f)
{
    using namespace ptb_fast;
    if (batch) return ring == 2 ? f(megakernel<false, 2, kFold, true>) : (ring == 1 ? f(megakernel<false, 1, kFold, true>) : f(megakernel<false, 0, kFold, true>));
    return ring == 2 ? f(megakernel<false, 2, kFold, false>) : (ring == 1 ? f(megakernel<false, 1, kFold, false>) : f(megakernel<false, 0, kFold, false>));
}
template <class F>
cudaError_t with(int fold, int ring, bool batch, F&& f)
{
    switch (fold) {
    case 1: return pick<1>(ring, batch, f);
    case 2: return pick<2>(ring, batch, f);
    case 3: return pick<3>(ring, batch, f);
    default: return pick<0>(ring, batch, f);
    }
}

} // namespace

size_t params_size() { return sizeof(ptb_fast::RenderParams); }
int threads() { return ptb_fast::kMegaThreads; }

cudaError_t prepare(int fold, int smem, int* with_ring, int* with_queue, int* without)
{
    cudaError_t e = cudaSuccess;
    for (int ring = 0; ring < 3 && e == cudaSuccess; ++ring)
        for (int batch = 0; batch < 2 && e == cudaSuccess; ++batch)
            e = with(fold, ring, batch != 0, [&](auto k) { return cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, smem); });
    if (e != cudaSuccess) return e;
    e = with(fold, 2, false, [&](auto k) { return cudaOccupancyMaxActiveBlocksPerMultiprocessor(with_queue, k, ptb_fast::kMegaThreads, smem); });
    if (e != cudaSuccess) return e;
    e = with(fold, 1, false, [&](auto k) { return cudaOccupancyMaxActiveBlocksPerMultiprocessor(with_ring, k, ptb_fast::kMegaThreads, smem); });
    if (e != cudaSuccess) return e;
    return with(fold, 0, false, [&](auto k) { return cudaOccupancyMaxActiveBlocksPerMultiprocessor(without, k, ptb_fast::kMegaThreads, smem); });
}

cudaError_t launch(const void* render_params, int fold, int ring, bool batch, int grid, int smem, cudaStream_t stream)
{
    ptb_fast::RenderParams P;
    memcpy(&P, render_params, sizeof P);          // ptb::RenderParams and ptb_fast::RenderParams are the same declaration
    return with(fold, ring, batch, [&](auto k) {
        k<<<grid, ptb_fast::kMegaThreads, smem, stream>>>(P);
        return cudaGetLastError();
    });
}

} // namespace ptb_fast_api
