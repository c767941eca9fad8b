// adamw_emul.cpp -- TEST INFRASTRUCTURE ONLY: runs the work-item body of co-occ_b200/csrc/adamw.cuh on the CPU so that
// tests/test_adamw_emul.py can compare it with torch.optim.AdamW without a GPU.  Never linked into the product.
#include "../../co-occ_b200/csrc/adamw.cuh"

using namespace coocc::opt;

extern "C" int emul_adamw_step(const void* tensors, int ntensors, const int* chunk_tensor, const int* chunk_index,
                               int nchunks, int chunk_elems, float lr, float beta1, float beta2, float eps,
                               float weight_decay, float* step, int zero_grad, const float* dyn) {
  *step += 1.f;
  AdamWP a{};
  a.tensors = reinterpret_cast<const TensorEntry*>(tensors);
  a.chunk_tensor = chunk_tensor; a.chunk_index = chunk_index; a.chunk_elems = chunk_elems;
  a.lr = lr; a.beta1 = beta1; a.beta2 = beta2; a.eps = eps; a.weight_decay = weight_decay; a.step = step;
  a.zero_grad = zero_grad;
  a.dyn = dyn;
  const long long n = (long long)nchunks * (chunk_elems >> 2);
  for (long long i = 0; i < n; ++i) adamw_item(a, i);
  (void)ntensors;
  return 0;
}
