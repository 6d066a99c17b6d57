// kb_bind.h — host entry of the bind-list kernels (kb_bind.cu, its own translation unit: the only user of CUB)
#pragma once
#include <cuda_runtime.h>
#include <stddef.h>
#include "../../include/kbgpu.h"
namespace kb {
size_t bind_scratch_bytes(uint32_t T);
cudaError_t bind_list(const kb_decision* d_dec, uint32_t T, unsigned char* scratch, size_t scratch_bytes, uint32_t* h_task, int32_t* h_node,
                      uint32_t* n_out, cudaStream_t st);
}
