// adamw.cu -- CUDA launcher of the multi-tensor AdamW + bf16-shadow step (body in adamw.cuh).
// Arithmetic verified on the CPU against torch.optim.AdamW (tests/test_adamw_emul.py) and on the GPU
// (tests/test_gpu_adamw.py).
#include <cuda_runtime.h>

#include "../../include/coocc_b200.h"
#include "adamw.cuh"

namespace coocc {
namespace opt {
__global__ void __launch_bounds__(256) adamw_kernel(const AdamWP a, long long n) {
  for (long long i = blockIdx.x * 256LL + threadIdx.x; i < n; i += gridDim.x * 256LL) adamw_item(a, i);
}
__global__ void step_inc_kernel(float* step) { *step += 1.f; }
}  // namespace opt
}  // namespace coocc

using namespace coocc::opt;

static_assert(sizeof(TensorEntry) == 56, "coocc_adamw tensor table entry layout (5 pointers + int64 + 2 floats)");

// tensors: device array of ntensors 56-byte entries {float* p, float* g, float* m, float* v, uint16* bf16_shadow|NULL,
// int64 n, float wd_mult, float lr_mult};  dyn: device float[2] {lr multiplier, gradient scale} or NULL;
// chunk_tensor / chunk_index: device int[nchunks] (tensor id and chunk number of every chunk_elems-sized chunk);
// step: device float, incremented by this call before the update (t = 1 for the first step).
extern "C" int coocc_adamw_step(const void* tensors, int ntensors, const int* chunk_tensor, const int* chunk_index,
                                int nchunks, int chunk_elems, float lr, float beta1, float beta2, float eps,
                                float weight_decay, float* step, int zero_grad, const float* dyn, void* stream) {
  if (!tensors || !chunk_tensor || !chunk_index || !step || ntensors < 1 || nchunks < 1 || chunk_elems < 4 ||
      (chunk_elems & 3))
    return COOCC_ERR_ARG;
  cudaStream_t st = (cudaStream_t)stream;
  step_inc_kernel<<<1, 1, 0, st>>>(step);
  AdamWP a{};
  a.tensors = reinterpret_cast<const TensorEntry*>(tensors);
  a.chunk_tensor = chunk_tensor; a.chunk_index = chunk_index; a.chunk_elems = chunk_elems;
  a.lr = lr; a.beta1 = beta1; a.beta2 = beta2; a.eps = eps; a.weight_decay = weight_decay; a.step = step;
  a.zero_grad = zero_grad;
  a.dyn = dyn;
  const long long n = (long long)nchunks * (chunk_elems >> 2);
  long long b = (n + 255) / 256;
  if (b > 148LL * 32) b = 148LL * 32;
  adamw_kernel<<<(unsigned)b, 256, 0, st>>>(a, n);
  return cudaGetLastError() == cudaSuccess ? 0 : COOCC_ERR_CUDA;
}
