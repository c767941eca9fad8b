// adamw.cuh -- work-item body of the multi-tensor AdamW step that also emits the bf16 operand copy of every
// parameter (the copy the tensor-core convolutions consume), replacing three passes over the 148 M hot-path
// parameters per step: torch's fused AdamW, the fp32 -> bf16 weight conversions, and (optionally) the zero fill of
// the gradient buffers.  Same __host__ __device__ scheme as fine_stage.cuh: the CUDA launcher (adamw.cu) and the
// CPU test harness (tests/emul/adamw_emul.cpp) compile this one source.
// Update rule = torch.optim.AdamW (decoupled weight decay, bias correction), per element:
//   p *= 1 - lr*wd;  m += (g - m)(1 - b1);  v = b2 v + (1 - b2) g^2;
//   p -= (lr / (1 - b1^t)) * m / (sqrt(v) / sqrt(1 - b2^t) + eps)
#pragma once
#include <math.h>
#include <stdint.h>

#if defined(__CUDACC__)
#define COOCC_HD __host__ __device__ __forceinline__
#else
#define COOCC_HD inline
#endif

namespace coocc {
namespace opt {

struct TensorEntry {
  float* p;
  float* g;          // gradient (read; zeroed afterwards when zero_grad != 0)
  float* m;
  float* v;
  uint16_t* shadow;  // bf16 copy of p in the same memory order, or nullptr
  long long n;
  float wd_mult;     // parameter-group multipliers (mmcv paramwise_cfg: norm_decay_mult = 0 for the BatchNorm / GroupNorm
  float lr_mult;     // weights and biases, coocc_multi_r50_256x704.py:263-276)
};

struct AdamWP {
  const TensorEntry* tensors;
  const int* chunk_tensor;       // [nchunks] tensor index of every chunk
  const int* chunk_index;        // [nchunks] chunk number inside its tensor
  int chunk_elems;               // elements per chunk (multiple of 4)
  float lr, beta1, beta2, eps, weight_decay;
  const float* step;             // device scalar t (already incremented for this update)
  int zero_grad;
  // device float[2] or nullptr: {learning-rate multiplier (schedule), gradient scale (clip_grad_norm coefficient)} --
  // read at run time, so a step replayed from a CUDA graph follows the schedule and the clip
  const float* dyn;
};

COOCC_HD uint16_t f32_to_bf16_rn(float f) {
  uint32_t u;
#if defined(__CUDA_ARCH__)
  u = __float_as_uint(f);
#else
  union { float f; uint32_t u; } c; c.f = f; u = c.u;
#endif
  if ((u & 0x7fffffffu) > 0x7f800000u) return (uint16_t)((u >> 16) | 0x0040u);     // NaN stays NaN
  const uint32_t r = 0x7fffu + ((u >> 16) & 1u);                                     // round to nearest even
  return (uint16_t)((u + r) >> 16);
}

// item = chunk * (chunk_elems / 4) + quad
COOCC_HD void adamw_item(const AdamWP& a, long long id) {
  const int qpc = a.chunk_elems >> 2;
  const int chunk = (int)(id / qpc), quad = (int)(id % qpc);
  const TensorEntry t = a.tensors[a.chunk_tensor[chunk]];
  const long long base = (long long)a.chunk_index[chunk] * a.chunk_elems + (long long)quad * 4;
  if (base >= t.n) return;
  const float tstep = *a.step;
  const float bc1 = 1.f - powf(a.beta1, tstep);
  const float bc2 = 1.f - powf(a.beta2, tstep);
  const float lr = a.lr * t.lr_mult * (a.dyn ? a.dyn[0] : 1.f);
  const float gscale = a.dyn ? a.dyn[1] : 1.f;
  const float step_size = lr / bc1;
  const float inv_sqrt_bc2 = 1.f / sqrtf(bc2);
  const float decay = 1.f - lr * a.weight_decay * t.wd_mult;
  const int cnt = (t.n - base) < 4 ? (int)(t.n - base) : 4;
  const uintptr_t al = (uintptr_t)(t.p + base) | (uintptr_t)(t.g + base) | (uintptr_t)(t.m + base) | (uintptr_t)(t.v + base);
  if (cnt == 4 && (al & 15) == 0 && (!t.shadow || ((uintptr_t)(t.shadow + base) & 7) == 0)) {
    // 16-byte accesses: one quad per work item, consecutive items touch consecutive quads (coalesced)
    struct alignas(16) F4 { float x[4]; };
    struct alignas(8) H4 { uint16_t x[4]; };
    const F4 g4 = *reinterpret_cast<const F4*>(t.g + base);
    F4 p4 = *reinterpret_cast<const F4*>(t.p + base);
    F4 m4 = *reinterpret_cast<const F4*>(t.m + base);
    F4 v4 = *reinterpret_cast<const F4*>(t.v + base);
    H4 h4;
    for (int k = 0; k < 4; ++k) {
      const float g = g4.x[k] * gscale;
      float p = p4.x[k] * decay;
      const float m = m4.x[k] + (g - m4.x[k]) * (1.f - a.beta1);
      const float v = a.beta2 * v4.x[k] + (1.f - a.beta2) * g * g;
      p -= step_size * m / (sqrtf(v) * inv_sqrt_bc2 + a.eps);
      p4.x[k] = p; m4.x[k] = m; v4.x[k] = v;
      h4.x[k] = f32_to_bf16_rn(p);
    }
    *reinterpret_cast<F4*>(t.p + base) = p4;
    *reinterpret_cast<F4*>(t.m + base) = m4;
    *reinterpret_cast<F4*>(t.v + base) = v4;
    if (t.shadow) *reinterpret_cast<H4*>(t.shadow + base) = h4;
    if (a.zero_grad) { F4 z; z.x[0] = z.x[1] = z.x[2] = z.x[3] = 0.f; *reinterpret_cast<F4*>(t.g + base) = z; }
    return;
  }
  for (int k = 0; k < cnt; ++k) {
    const long long i = base + k;
    const float g = t.g[i] * gscale;
    float p = t.p[i] * decay;
    const float m = t.m[i] + (g - t.m[i]) * (1.f - a.beta1);
    const float v = a.beta2 * t.v[i] + (1.f - a.beta2) * g * g;
    p -= step_size * m / (sqrtf(v) * inv_sqrt_bc2 + a.eps);
    t.p[i] = p; t.m[i] = m; t.v[i] = v;
    if (t.shadow) t.shadow[i] = f32_to_bf16_rn(p);
    if (a.zero_grad) t.g[i] = 0.f;
  }
}

}  // namespace opt
}  // namespace coocc
