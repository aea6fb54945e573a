// ptb_fast.h — what ptb_abi.cu needs from the fast-arithmetic translation unit (ptb_fast.cu).  Internal to the library.
#pragma once
#include <cuda_runtime.h>
#include <stddef.h>

namespace ptb_fast_api {
size_t params_size();                                                     // sizeof(RenderParams) as ptb_fast.cu sees it
int threads();
// ring: 0 no primary-ray ring, 1 ring, 2 ring + completion queue (SPP 1)
cudaError_t prepare(int fold, int smem, int* with_ring, int* with_queue, int* without);     // shared-memory attributes + occupancy
cudaError_t launch(const void* render_params, int fold, int ring, bool batch, int grid, int smem, cudaStream_t stream);
} // namespace ptb_fast_api
