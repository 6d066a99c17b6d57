// kb_evict_launch.h — host entry of the reclaim / preempt kernels (kb_evict_kernels.cu, its own translation unit: built
// with L2-only global loads).
#pragma once
#include <cuda_runtime.h>
#include "kb_evict.h"
namespace kb {
cudaError_t launch_evict(bool preempt, const DevSession& S, const EvictDev& E, int sm_count, cudaStream_t stream);
}
