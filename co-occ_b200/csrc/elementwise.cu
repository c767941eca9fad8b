// elementwise.cu -- HBM-bound companions of the tensor-core convolutions:
//   BatchNorm3d (training mode) fused with ReLU and the residual add, forward and backward,
//   fed by the per-channel sums the conv epilogue already produced (no extra statistics pass);
//   x2 zero-insertion for the strided-conv data gradient.
//
// Reference semantics: torch.nn.BatchNorm3d / SyncBatchNorm (world size 1) as used at
// P/coocc/fuser/bifuser_n.py:25,28, P/coocc/backbones/resnet3d.py:41,45,54-60,
// P/coocc/necks/fpn3d.py:48-67 (ConvModule), P/coocc/dense_heads/occ_head.py:102-132:
//   y = (x - mean_batch) / sqrt(var_biased + eps) * gamma + beta ; running stats with momentum,
//   running_var from the unbiased variance.
// All tensors are [V, C] row-major (NDHWC); one thread owns 4 consecutive channels of a row, a
// warp covers 128 contiguous channels (512 B) -> fully coalesced float4 traffic.
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <stdint.h>

#include <initializer_list>

#include "../../include/coocc_b200.h"
#include "act_types.cuh"

namespace coocc {

// mean / invstd from the conv epilogue sums; updates the running statistics in place
__global__ void bn_finalize_kernel(const float* __restrict__ stats, int C, double count, float eps,
                                   float momentum, float* __restrict__ running_mean,
                                   float* __restrict__ running_var, float* __restrict__ mean_invstd) {
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= C) return;
  const double m = (double)stats[c] / count;
  double var = (double)stats[C + c] / count - m * m;
  if (var < 0.0) var = 0.0;
  mean_invstd[c] = (float)m;
  mean_invstd[C + c] = (float)(1.0 / sqrt(var + (double)eps));
  if (running_mean != nullptr) {
    const double unbiased = count > 1.0 ? var * count / (count - 1.0) : var;
    running_mean[c] = (1.f - momentum) * running_mean[c] + momentum * (float)m;
    running_var[c] = (1.f - momentum) * running_var[c] + momentum * (float)unbiased;
  }
}

// the BatchNorm affine map, one expression shared by forward and backward so that the backward can
// recompute the ReLU mask (y > 0) from x bit-for-bit instead of re-reading the stored output
// (written as x * scale + shift with explicitly rounded steps: the wide kernels keep scale / shift in registers and
// must produce the same bits)
__device__ __forceinline__ float bn_scale(float is, float g) { return __fmul_rn(is, g); }
__device__ __forceinline__ float bn_shift(float m, float sc, float b) { return __fmaf_rn(-m, sc, b); }
__device__ __forceinline__ float bn_affine(float x, float m, float is, float g, float b) {
  const float sc = bn_scale(is, g);
  return __fmaf_rn(x, sc, bn_shift(m, sc, b));
}
// dz = dout * [y > 0]: mask from the stored output `o` (use_out) or recomputed from x
__device__ __forceinline__ float4 bn_mask(float4 d, bool use_out, float4 o, float4 xv, float4 m, float4 is, float4 g,
                                          float4 b) {
  if (use_out) {
    d.x = o.x > 0.f ? d.x : 0.f; d.y = o.y > 0.f ? d.y : 0.f;
    d.z = o.z > 0.f ? d.z : 0.f; d.w = o.w > 0.f ? d.w : 0.f;
  } else {
    d.x = bn_affine(xv.x, m.x, is.x, g.x, b.x) > 0.f ? d.x : 0.f;
    d.y = bn_affine(xv.y, m.y, is.y, g.y, b.y) > 0.f ? d.y : 0.f;
    d.z = bn_affine(xv.z, m.z, is.z, g.z, b.z) > 0.f ? d.z : 0.f;
    d.w = bn_affine(xv.w, m.w, is.w, g.w, b.w) > 0.f ? d.w : 0.f;
  }
  return d;
}

// out = relu?( (x - mean) * invstd * gamma + beta (+ residual) )
template <typename T>
__global__ void __launch_bounds__(256) bn_act_fwd_kernel(const T* __restrict__ x, long long ldx,
                                                         long long V, int C,
                                                         const float* __restrict__ mean_invstd,
                                                         const float* __restrict__ gamma,
                                                         const float* __restrict__ beta,
                                                         const T* __restrict__ residual, long long ldr,
                                                         int relu, T* __restrict__ out, long long ldo) {
  const int c4 = C >> 2;
  const long long total = V * c4;
  for (long long i = blockIdx.x * 256LL + threadIdx.x; i < total; i += gridDim.x * 256LL) {
    const long long r = i / c4;
    const int c = (int)(i % c4) * 4;
    const float4 xv = load4(x + r * ldx + c);
    const float4 m = *reinterpret_cast<const float4*>(mean_invstd + c);
    const float4 is = *reinterpret_cast<const float4*>(mean_invstd + C + c);
    const float4 g = *reinterpret_cast<const float4*>(gamma + c);
    const float4 b = *reinterpret_cast<const float4*>(beta + c);
    float4 y;
    y.x = bn_affine(xv.x, m.x, is.x, g.x, b.x);
    y.y = bn_affine(xv.y, m.y, is.y, g.y, b.y);
    y.z = bn_affine(xv.z, m.z, is.z, g.z, b.z);
    y.w = bn_affine(xv.w, m.w, is.w, g.w, b.w);
    if (residual != nullptr) {
      const float4 rv = load4(residual + r * ldr + c);
      y.x += rv.x; y.y += rv.y; y.z += rv.z; y.w += rv.w;
    }
    if (relu) {
      y.x = fmaxf(y.x, 0.f); y.y = fmaxf(y.y, 0.f); y.z = fmaxf(y.z, 0.f); y.w = fmaxf(y.w, 0.f);
    }
    store4(out + r * ldo + c, y);
  }
}

// sums[c] = sum_r dz, sums[C + c] = sum_r dz * xhat, dz = dout * [y > 0] (relu) or dout.  The mask is read
// from the stored output when `out` is given (needed when a residual was added before the ReLU) and recomputed
// from x otherwise -- one tensor less to read.
// block = 32 x 8: threadIdx.x -> channel group (4 channels), threadIdx.y -> row lane.
template <typename T>
__global__ void __launch_bounds__(256) bn_act_bwd_reduce_kernel(
    const T* __restrict__ dout, long long ldd, const T* __restrict__ out, long long ldo,
    const T* __restrict__ x, long long ldx, long long V, int C, const float* __restrict__ mean_invstd,
    const float* __restrict__ gamma, const float* __restrict__ beta, int relu, int rows_per_block,
    float* __restrict__ sums) {
  const int c = (blockIdx.x * 32 + threadIdx.x) * 4;
  float4 s1 = make_float4(0.f, 0.f, 0.f, 0.f), s2 = s1;
  if (c < C) {
    const float4 m = *reinterpret_cast<const float4*>(mean_invstd + c);
    const float4 is = *reinterpret_cast<const float4*>(mean_invstd + C + c);
    const bool use_out = out != nullptr;
    float4 g = m, b = m;
    if (relu && !use_out) {
      g = *reinterpret_cast<const float4*>(gamma + c);
      b = *reinterpret_cast<const float4*>(beta + c);
    }
    const long long r0 = (long long)blockIdx.y * rows_per_block;
    const long long r1 = min(V, r0 + rows_per_block);
    for (long long r = r0 + threadIdx.y; r < r1; r += 8) {
      float4 d = load4(dout + r * ldd + c);
      const float4 xv = load4(x + r * ldx + c);
      if (relu) d = bn_mask(d, use_out, use_out ? load4(out + r * ldo + c) : xv, xv, m, is, g, b);
      s1.x += d.x; s1.y += d.y; s1.z += d.z; s1.w += d.w;
      s2.x += d.x * (xv.x - m.x) * is.x; s2.y += d.y * (xv.y - m.y) * is.y;
      s2.z += d.z * (xv.z - m.z) * is.z; s2.w += d.w * (xv.w - m.w) * is.w;
    }
  }
  __shared__ float4 sh1[8][32], sh2[8][32];
  sh1[threadIdx.y][threadIdx.x] = s1;
  sh2[threadIdx.y][threadIdx.x] = s2;
  __syncthreads();
  if (threadIdx.y == 0 && c < C) {
    for (int j = 1; j < 8; ++j) {
      const float4 a = sh1[j][threadIdx.x], b = sh2[j][threadIdx.x];
      s1.x += a.x; s1.y += a.y; s1.z += a.z; s1.w += a.w;
      s2.x += b.x; s2.y += b.y; s2.z += b.z; s2.w += b.w;
    }
    atomicAdd(&sums[c + 0], s1.x); atomicAdd(&sums[c + 1], s1.y);
    atomicAdd(&sums[c + 2], s1.z); atomicAdd(&sums[c + 3], s1.w);
    atomicAdd(&sums[C + c + 0], s2.x); atomicAdd(&sums[C + c + 1], s2.y);
    atomicAdd(&sums[C + c + 2], s2.z); atomicAdd(&sums[C + c + 3], s2.w);
  }
}

// dx = gamma * invstd * (dz - mean(dz) - xhat * mean(dz * xhat));  dres = dz (optional)
template <typename T>
__global__ void __launch_bounds__(256) bn_act_bwd_apply_kernel(
    const T* __restrict__ dout, long long ldd, const T* __restrict__ out, long long ldo,
    const T* __restrict__ x, long long ldx, long long V, int C, const float* __restrict__ mean_invstd,
    const float* __restrict__ gamma, const float* __restrict__ beta, const float* __restrict__ sums, int relu,
    T* __restrict__ dx, long long lddx, T* __restrict__ dres, long long lddr, float inv_n) {
  const int c4 = C >> 2;
  const long long total = V * c4;
  for (long long i = blockIdx.x * 256LL + threadIdx.x; i < total; i += gridDim.x * 256LL) {
    const long long r = i / c4;
    const int c = (int)(i % c4) * 4;
    float4 d = load4(dout + r * ldd + c);
    const float4 xv = load4(x + r * ldx + c);
    const float4 m = *reinterpret_cast<const float4*>(mean_invstd + c);
    const float4 is = *reinterpret_cast<const float4*>(mean_invstd + C + c);
    const float4 g = *reinterpret_cast<const float4*>(gamma + c);
    if (relu) {
      const bool use_out = out != nullptr;
      const float4 bt = use_out ? m : *reinterpret_cast<const float4*>(beta + c);
      d = bn_mask(d, use_out, use_out ? load4(out + r * ldo + c) : xv, xv, m, is, g, bt);
    }
    if (dres != nullptr) store4(dres + r * lddr + c, d);
    const float4 a = *reinterpret_cast<const float4*>(sums + c);
    const float4 b = *reinterpret_cast<const float4*>(sums + C + c);
    float4 y;
    y.x = g.x * is.x * (d.x - a.x * inv_n - (xv.x - m.x) * is.x * b.x * inv_n);
    y.y = g.y * is.y * (d.y - a.y * inv_n - (xv.y - m.y) * is.y * b.y * inv_n);
    y.z = g.z * is.z * (d.z - a.z * inv_n - (xv.z - m.z) * is.z * b.z * inv_n);
    y.w = g.w * is.w * (d.w - a.w * inv_n - (xv.w - m.w) * is.w * b.w * inv_n);
    store4(dx + r * lddx + c, y);
  }
}

// up[(2x,2y,2z), :] = src[(x,y,z), :], zeros elsewhere: the data gradient of a stride-2 conv is the
// stride-1 data gradient of the zero-inserted output gradient.
template <typename T>
__global__ void __launch_bounds__(256) dilate2_kernel(const T* __restrict__ src, long long lds, int oX,
                                                      int oY, int oZ, int C, T* __restrict__ dst,
                                                      long long ldd, int X, int Y, int Z) {
  const int cv = C / (16 / (int)sizeof(T));           // 16-byte vectors per row
  const long long total = (long long)X * Y * Z * cv;
  for (long long i = blockIdx.x * 256LL + threadIdx.x; i < total; i += gridDim.x * 256LL) {
    const long long v = i / cv;
    const int c = (int)(i % cv);
    const int z = v % Z, y = (v / Z) % Y, x = v / ((long long)Z * Y);
    uint4 val = make_uint4(0, 0, 0, 0);
    if (!(x & 1) && !(y & 1) && !(z & 1) && (x >> 1) < oX && (y >> 1) < oY && (z >> 1) < oZ) {
      const long long sv = ((long long)(x >> 1) * oY + (y >> 1)) * oZ + (z >> 1);
      val = *reinterpret_cast<const uint4*>(src + sv * lds + (long long)c * (16 / sizeof(T)));
    }
    *reinterpret_cast<uint4*>(dst + v * ldd + (long long)c * (16 / sizeof(T))) = val;
  }
}


// ---------------------------------------------------------------------------------------------
// "Wide" variants (the ones normally launched): every access is 16 bytes (8 bf16 / 4 fp32 channels), the per-channel
// constants live in registers for the whole kernel, a thread walks rows with UNROLL independent rows in flight, and
// there is no per-element 64-bit division.  The narrow kernels above ran at 43-56 % of the measured copy bandwidth
// (profiles/r02_bench*.json hbm_kernels); they remain as the fallback for C not a multiple of the vector width.
// Thread layout: 256 threads = (C / VEC channel groups) x (256 / groups row lanes).
// ---------------------------------------------------------------------------------------------
template <int VEC>
struct VecF {
  float v[VEC];
};
// raw 16-byte vectors stay packed in registers until they are used (4 registers per row in flight)
template <typename T> struct Raw;
template <> struct Raw<float> { typedef float4 type; };
template <> struct Raw<__nv_bfloat16> { typedef uint4 type; };
__device__ __forceinline__ float4 loadraw(const float* p) { return *reinterpret_cast<const float4*>(p); }
__device__ __forceinline__ uint4 loadraw(const __nv_bfloat16* p) { return *reinterpret_cast<const uint4*>(p); }
__device__ __forceinline__ VecF<4> unpack(const float4& t) {
  VecF<4> r;
  r.v[0] = t.x; r.v[1] = t.y; r.v[2] = t.z; r.v[3] = t.w;
  return r;
}
__device__ __forceinline__ VecF<8> unpack(const uint4& u) {
  VecF<8> r;
  r.v[0] = __uint_as_float(u.x << 16); r.v[1] = __uint_as_float(u.x & 0xFFFF0000u);
  r.v[2] = __uint_as_float(u.y << 16); r.v[3] = __uint_as_float(u.y & 0xFFFF0000u);
  r.v[4] = __uint_as_float(u.z << 16); r.v[5] = __uint_as_float(u.z & 0xFFFF0000u);
  r.v[6] = __uint_as_float(u.w << 16); r.v[7] = __uint_as_float(u.w & 0xFFFF0000u);
  return r;
}
__device__ __forceinline__ void storev(float* p, const VecF<4>& r) {
  *reinterpret_cast<float4*>(p) = make_float4(r.v[0], r.v[1], r.v[2], r.v[3]);
}
__device__ __forceinline__ void storev(__nv_bfloat16* p, const VecF<8>& r) {
  uint4 u;
  __nv_bfloat162 a = __floats2bfloat162_rn(r.v[0], r.v[1]), b = __floats2bfloat162_rn(r.v[2], r.v[3]);
  __nv_bfloat162 c = __floats2bfloat162_rn(r.v[4], r.v[5]), d = __floats2bfloat162_rn(r.v[6], r.v[7]);
  u.x = *reinterpret_cast<uint32_t*>(&a); u.y = *reinterpret_cast<uint32_t*>(&b);
  u.z = *reinterpret_cast<uint32_t*>(&c); u.w = *reinterpret_cast<uint32_t*>(&d);
  *reinterpret_cast<uint4*>(p) = u;
}
template <int VEC>
__device__ __forceinline__ VecF<VEC> loadc(const float* p) {        // VEC consecutive fp32 constants
  VecF<VEC> r;
#pragma unroll
  for (int k = 0; k < VEC; k += 4) {
    const float4 t = *reinterpret_cast<const float4*>(p + k);
    r.v[k] = t.x; r.v[k + 1] = t.y; r.v[k + 2] = t.z; r.v[k + 3] = t.w;
  }
  return r;
}
// scale = invstd * gamma, shift = beta - mean * scale of the thread's VEC channels
template <int VEC>
__device__ __forceinline__ void load_affine(const float* mean_invstd, const float* gamma, const float* beta, int C, int c,
                                            VecF<VEC>& sc, VecF<VEC>& sh) {
  const VecF<VEC> m = loadc<VEC>(mean_invstd + c), is = loadc<VEC>(mean_invstd + C + c);
  const VecF<VEC> g = loadc<VEC>(gamma + c), b = loadc<VEC>(beta + c);
#pragma unroll
  for (int k = 0; k < VEC; ++k) {
    sc.v[k] = bn_scale(is.v[k], g.v[k]);
    sh.v[k] = bn_shift(m.v[k], sc.v[k], b.v[k]);
  }
}

template <typename T, int VEC, int UNROLL>
__global__ void __launch_bounds__(256) bn_act_fwd_wide_kernel(const T* __restrict__ x, long long ldx, long long V,
                                                              int C, const float* __restrict__ mean_invstd,
                                                              const float* __restrict__ gamma,
                                                              const float* __restrict__ beta,
                                                              const T* __restrict__ residual, long long ldr, int relu,
                                                              T* __restrict__ out, long long ldo, int cg, int rl) {
  typedef typename Raw<T>::type R;
  const int tc = threadIdx.x % cg, tr = threadIdx.x / cg;
  if (tr >= rl) return;
  const int c = tc * VEC;
  VecF<VEC> sc, sh;
  load_affine<VEC>(mean_invstd, gamma, beta, C, c, sc, sh);
  const long long stride = (long long)gridDim.x * rl;
  for (long long r0 = (long long)blockIdx.x * rl + tr; r0 < V; r0 += stride * UNROLL) {
    R xr[UNROLL], rr[UNROLL];
#pragma unroll
    for (int u = 0; u < UNROLL; ++u) {
      const long long r = r0 + u * stride;
      if (r < V) {
        xr[u] = loadraw(x + r * ldx + c);
        if (residual != nullptr) rr[u] = loadraw(residual + r * ldr + c);
      }
    }
#pragma unroll
    for (int u = 0; u < UNROLL; ++u) {
      const long long r = r0 + u * stride;
      if (r < V) {
        const VecF<VEC> xv = unpack(xr[u]);
        VecF<VEC> y;
#pragma unroll
        for (int k = 0; k < VEC; ++k) y.v[k] = __fmaf_rn(xv.v[k], sc.v[k], sh.v[k]);
        if (residual != nullptr) {
          const VecF<VEC> rv = unpack(rr[u]);
#pragma unroll
          for (int k = 0; k < VEC; ++k) y.v[k] += rv.v[k];
        }
        if (relu) {
#pragma unroll
          for (int k = 0; k < VEC; ++k) y.v[k] = fmaxf(y.v[k], 0.f);
        }
        storev(out + r * ldo + c, y);
      }
    }
  }
}

// accumulates sum dz and sum dz * x per channel; the block's reducing threads turn the latter into
// sum dz * xhat = invstd * (sum dz * x - mean * sum dz) before the atomics
template <typename T, int VEC, int UNROLL>
__global__ void __launch_bounds__(256) bn_act_bwd_reduce_wide_kernel(
    const T* __restrict__ dout, long long ldd, const T* __restrict__ out, long long ldo, const T* __restrict__ x,
    long long ldx, long long V, int C, const float* __restrict__ mean_invstd, const float* __restrict__ gamma,
    const float* __restrict__ beta, int relu, float* __restrict__ sums, int cg, int rl) {
  typedef typename Raw<T>::type R;
  extern __shared__ float sh[];                    // [2][rl][C]
  const int tc = threadIdx.x % cg, tr = threadIdx.x / cg;
  const int c = tc * VEC;
  const bool use_out = out != nullptr;
  VecF<VEC> s1, s2;
#pragma unroll
  for (int k = 0; k < VEC; ++k) { s1.v[k] = 0.f; s2.v[k] = 0.f; }
  if (tr < rl) {
    VecF<VEC> sc, sf;
    if (relu && !use_out) load_affine<VEC>(mean_invstd, gamma, beta, C, c, sc, sf);
    // shifted accumulation (x - mean) keeps sum dz * x well conditioned
    const VecF<VEC> m = loadc<VEC>(mean_invstd + c);
    const long long stride = (long long)gridDim.x * rl;
    for (long long r0 = (long long)blockIdx.x * rl + tr; r0 < V; r0 += stride * UNROLL) {
      R dr[UNROLL], xr[UNROLL], orw[UNROLL];
#pragma unroll
      for (int u = 0; u < UNROLL; ++u) {
        const long long r = r0 + u * stride;
        if (r < V) {
          dr[u] = loadraw(dout + r * ldd + c);
          xr[u] = loadraw(x + r * ldx + c);
          if (relu && use_out) orw[u] = loadraw(out + r * ldo + c);
        }
      }
#pragma unroll
      for (int u = 0; u < UNROLL; ++u) {
        const long long r = r0 + u * stride;
        if (r < V) {
          const VecF<VEC> dv = unpack(dr[u]), xv = unpack(xr[u]);
          VecF<VEC> ov;
          if (relu && use_out) ov = unpack(orw[u]);
#pragma unroll
          for (int k = 0; k < VEC; ++k) {
            float d = dv.v[k];
            if (relu) {
              const float y = use_out ? ov.v[k] : __fmaf_rn(xv.v[k], sc.v[k], sf.v[k]);
              d = y > 0.f ? d : 0.f;
            }
            s1.v[k] += d;
            s2.v[k] += d * (xv.v[k] - m.v[k]);
          }
        }
      }
    }
#pragma unroll
    for (int k = 0; k < VEC; ++k) {
      sh[(size_t)tr * C + c + k] = s1.v[k];
      sh[(size_t)(rl + tr) * C + c + k] = s2.v[k];
    }
  }
  __syncthreads();
  for (int i = threadIdx.x; i < 2 * C; i += 256) {
    const int which = i / C, ch = i % C;
    float t = 0.f;
    for (int j = 0; j < rl; ++j) t += sh[(size_t)(which * rl + j) * C + ch];
    if (which) t *= mean_invstd[C + ch];
    atomicAdd(&sums[i], t);
  }
}

template <typename T, int VEC, int UNROLL>
__global__ void __launch_bounds__(256) bn_act_bwd_apply_wide_kernel(
    const T* __restrict__ dout, long long ldd, const T* __restrict__ out, long long ldo, const T* __restrict__ x,
    long long ldx, long long V, int C, const float* __restrict__ mean_invstd, const float* __restrict__ gamma,
    const float* __restrict__ beta, const float* __restrict__ sums, int relu, T* __restrict__ dx, long long lddx,
    T* __restrict__ dres, long long lddr, float inv_n, int cg, int rl) {
  typedef typename Raw<T>::type R;
  const int tc = threadIdx.x % cg, tr = threadIdx.x / cg;
  if (tr >= rl) return;
  const int c = tc * VEC;
  const bool use_out = out != nullptr;
  VecF<VEC> sc, sf;
  if (relu && !use_out) load_affine<VEC>(mean_invstd, gamma, beta, C, c, sc, sf);
  // dx = gamma*invstd * (dz - a/n - xhat * b/n) = ka * dz + kb * x + kc
  VecF<VEC> ka, kb, kc;
  {
    const VecF<VEC> m = loadc<VEC>(mean_invstd + c), is = loadc<VEC>(mean_invstd + C + c);
    const VecF<VEC> g = loadc<VEC>(gamma + c);
    const VecF<VEC> a = loadc<VEC>(sums + c), b = loadc<VEC>(sums + C + c);
#pragma unroll
    for (int k = 0; k < VEC; ++k) {
      ka.v[k] = g.v[k] * is.v[k];
      kb.v[k] = -ka.v[k] * is.v[k] * (b.v[k] * inv_n);
      kc.v[k] = -ka.v[k] * (a.v[k] * inv_n) - kb.v[k] * m.v[k];
    }
  }
  const long long stride = (long long)gridDim.x * rl;
  for (long long r0 = (long long)blockIdx.x * rl + tr; r0 < V; r0 += stride * UNROLL) {
    R dr[UNROLL], xr[UNROLL], orw[UNROLL];
#pragma unroll
    for (int u = 0; u < UNROLL; ++u) {
      const long long r = r0 + u * stride;
      if (r < V) {
        dr[u] = loadraw(dout + r * ldd + c);
        xr[u] = loadraw(x + r * ldx + c);
        if (relu && use_out) orw[u] = loadraw(out + r * ldo + c);
      }
    }
#pragma unroll
    for (int u = 0; u < UNROLL; ++u) {
      const long long r = r0 + u * stride;
      if (r < V) {
        const VecF<VEC> dv = unpack(dr[u]), xv = unpack(xr[u]);
        VecF<VEC> ov, dz, y;
        if (relu && use_out) ov = unpack(orw[u]);
#pragma unroll
        for (int k = 0; k < VEC; ++k) {
          float d = dv.v[k];
          if (relu) {
            const float yy = use_out ? ov.v[k] : __fmaf_rn(xv.v[k], sc.v[k], sf.v[k]);
            d = yy > 0.f ? d : 0.f;
          }
          dz.v[k] = d;
          y.v[k] = ka.v[k] * d + (kb.v[k] * xv.v[k] + kc.v[k]);
        }
        if (dres != nullptr) storev(dres + r * lddr + c, dz);
        storev(dx + r * lddx + c, y);
      }
    }
  }
}

// can the wide kernels serve this call?  (vector width divides C and every row stride, 16-byte aligned bases,
// at most 256 channel groups)
static bool wide_ok(int C, int es, std::initializer_list<long long> lds, std::initializer_list<const void*> ptrs) {
  const int vec = 16 / es;
  if (C % vec || C / vec > 256) return false;
  for (long long ld : lds)
    if (ld % vec) return false;
  for (const void* p : ptrs)
    if (p && (reinterpret_cast<uintptr_t>(p) & 15)) return false;
  return true;
}
static int wide_grid(long long V, int rl, int per_sm) {
  long long b = (V + rl - 1) / rl;
  if (b > 148LL * per_sm) b = 148LL * per_sm;
  return (int)(b < 1 ? 1 : b);
}

static int grid_for(long long total) {
  long long b = (total + 255) / 256;
  if (b > 148LL * 32) b = 148LL * 32;
  if (b < 1) b = 1;
  return (int)b;
}

}  // namespace coocc

using namespace coocc;
#define CK_LAUNCH() (cudaGetLastError() == cudaSuccess ? 0 : COOCC_ERR_CUDA)

extern "C" int coocc_bn_finalize(const float* stats, int C, long long count, float eps, float momentum,
                                 float* running_mean, float* running_var, float* mean_invstd, void* stream) {
  if (!stats || !mean_invstd || C < 1 || count < 1) return COOCC_ERR_ARG;
  bn_finalize_kernel<<<(C + 127) / 128, 128, 0, (cudaStream_t)stream>>>(stats, C, (double)count, eps, momentum,
                                                                       running_mean, running_var, mean_invstd);
  return CK_LAUNCH();
}

// act_bf16 = 1: every activation tensor of the call (x, residual, out / dout, out, x, dx, dres) is
// stored as bf16; 0: fp32.  Statistics, gamma/beta and the arithmetic are fp32 either way.
extern "C" int coocc_bn_act_fwd(const void* x, long long ldx, long long V, int C, const float* mean_invstd,
                                const float* gamma, const float* beta, const void* residual, long long ldr,
                                int relu, void* out, long long ldo, int act_bf16, void* stream) {
  if (!x || !mean_invstd || !gamma || !beta || !out || (C & 3) || (ldx & 3) || (ldo & 3)) return COOCC_ERR_ARG;
  if (residual && (ldr & 3)) return COOCC_ERR_ARG;
  if (V < 1) return 0;
  const int es = act_bf16 ? 2 : 4;
  if (wide_ok(C, es, {ldx, ldo, residual ? ldr : 0}, {x, out, residual})) {
    const int cg = C / (16 / es), rl = 256 / cg;
    cudaStream_t st = (cudaStream_t)stream;
    if (act_bf16)
      bn_act_fwd_wide_kernel<__nv_bfloat16, 8, 8><<<wide_grid(V, rl * 8, 8), 256, 0, st>>>(
          (const __nv_bfloat16*)x, ldx, V, C, mean_invstd, gamma, beta, (const __nv_bfloat16*)residual, ldr, relu,
          (__nv_bfloat16*)out, ldo, cg, rl);
    else
      bn_act_fwd_wide_kernel<float, 4, 8><<<wide_grid(V, rl * 8, 8), 256, 0, st>>>(
          (const float*)x, ldx, V, C, mean_invstd, gamma, beta, (const float*)residual, ldr, relu, (float*)out, ldo, cg,
          rl);
    return CK_LAUNCH();
  }
  const int g = grid_for(V * (C >> 2));
  if (act_bf16)
    bn_act_fwd_kernel<__nv_bfloat16><<<g, 256, 0, (cudaStream_t)stream>>>(
        (const __nv_bfloat16*)x, ldx, V, C, mean_invstd, gamma, beta, (const __nv_bfloat16*)residual, ldr, relu,
        (__nv_bfloat16*)out, ldo);
  else
    bn_act_fwd_kernel<float><<<g, 256, 0, (cudaStream_t)stream>>>((const float*)x, ldx, V, C, mean_invstd, gamma, beta,
                                                                (const float*)residual, ldr, relu, (float*)out, ldo);
  return CK_LAUNCH();
}

// relu != 0: dz = dout * [y > 0]; the mask is read from `out` when it is given (required if a residual was added
// before the ReLU) and recomputed from x, gamma, beta when out == NULL.
// Backward in two stream-ordered halves so that SyncBatchNorm can all-reduce `sums` in between:
//   reduce: sums (float[2*C], zeroed by the caller) += (sum dz, sum dz*xhat) over this rank's rows
//   apply : dx = gamma*invstd*(dz - sums[0:C]/count - xhat*sums[C:2C]/count); dres = dz (optional)
// count = number of rows the statistics were taken over (V, or the sum of V over all ranks).
// After the reduce, sums[0:C] = dbeta and sums[C:2C] = dgamma (of this rank / of all ranks).
extern "C" int coocc_bn_act_bwd_reduce(const void* dout, long long ldd, const void* out, long long ldo,
                                       const void* x, long long ldx, long long V, int C, const float* mean_invstd,
                                       const float* gamma, const float* beta, int relu, float* sums, int act_bf16,
                                       void* stream) {
  if (!dout || !x || !mean_invstd || !sums || (C & 3) || (ldd & 3) || (ldx & 3)) return COOCC_ERR_ARG;
  if (relu && out && (ldo & 3)) return COOCC_ERR_ARG;
  if (relu && !out && (!gamma || !beta)) return COOCC_ERR_ARG;
  if (V < 1) return 0;
  {
    const int es = act_bf16 ? 2 : 4;
    const int cg = C / (16 / es) > 0 ? C / (16 / es) : 1, rl = 256 / cg;
    const size_t smem = (size_t)2 * (rl > 0 ? rl : 1) * C * sizeof(float);
    if (wide_ok(C, es, {ldd, ldx, (relu && out) ? ldo : 0}, {dout, x, (relu && out) ? out : nullptr}) && smem <= 48 * 1024) {
      cudaStream_t st = (cudaStream_t)stream;
      const int grid = wide_grid(V, rl * 4, 4);
      if (act_bf16)
        bn_act_bwd_reduce_wide_kernel<__nv_bfloat16, 8, 4><<<grid, 256, smem, st>>>(
            (const __nv_bfloat16*)dout, ldd, (const __nv_bfloat16*)out, ldo, (const __nv_bfloat16*)x, ldx, V, C,
            mean_invstd, gamma, beta, relu, sums, cg, rl);
      else
        bn_act_bwd_reduce_wide_kernel<float, 4, 4><<<grid, 256, smem, st>>>(
            (const float*)dout, ldd, (const float*)out, ldo, (const float*)x, ldx, V, C, mean_invstd, gamma, beta, relu,
            sums, cg, rl);
      return CK_LAUNCH();
    }
  }
  const int cgroups = (C / 4 + 31) / 32;
  int rows_per_block = 256;
  long long nby = (V + rows_per_block - 1) / rows_per_block;
  while (nby * cgroups > 148LL * 16) {
    rows_per_block *= 2;
    nby = (V + rows_per_block - 1) / rows_per_block;
  }
  const dim3 grid(cgroups, (unsigned)nby), block(32, 8);
  if (act_bf16)
    bn_act_bwd_reduce_kernel<__nv_bfloat16><<<grid, block, 0, (cudaStream_t)stream>>>(
        (const __nv_bfloat16*)dout, ldd, (const __nv_bfloat16*)out, ldo, (const __nv_bfloat16*)x, ldx, V, C,
        mean_invstd, gamma, beta, relu, rows_per_block, sums);
  else
    bn_act_bwd_reduce_kernel<float><<<grid, block, 0, (cudaStream_t)stream>>>(
        (const float*)dout, ldd, (const float*)out, ldo, (const float*)x, ldx, V, C, mean_invstd, gamma, beta, relu,
        rows_per_block, sums);
  return CK_LAUNCH();
}

extern "C" int coocc_bn_act_bwd_apply(const void* dout, long long ldd, const void* out, long long ldo,
                                      const void* x, long long ldx, long long V, int C, const float* mean_invstd,
                                      const float* gamma, const float* beta, int relu, const float* sums,
                                      long long count, void* dx, long long lddx, int act_bf16, void* dres,
                                      long long lddr, void* stream) {
  if (!dout || !x || !mean_invstd || !gamma || !sums || !dx || (C & 3) || (ldd & 3) || (ldx & 3) || (lddx & 3) ||
      count < 1)
    return COOCC_ERR_ARG;
  if (relu && out && (ldo & 3)) return COOCC_ERR_ARG;
  if (relu && !out && !beta) return COOCC_ERR_ARG;
  cudaStream_t st = (cudaStream_t)stream;
  if (V < 1) return 0;
  const float inv_n = (float)(1.0 / (double)count);
  {
    const int es = act_bf16 ? 2 : 4;
    if (wide_ok(C, es, {ldd, ldx, lddx, (relu && out) ? ldo : 0, dres ? lddr : 0},
                {dout, x, dx, (relu && out) ? out : nullptr, dres})) {
      const int cg = C / (16 / es), rl = 256 / cg;
      const int grid = wide_grid(V, rl * 4, 8);
      if (act_bf16)
        bn_act_bwd_apply_wide_kernel<__nv_bfloat16, 8, 4><<<grid, 256, 0, st>>>(
            (const __nv_bfloat16*)dout, ldd, (const __nv_bfloat16*)out, ldo, (const __nv_bfloat16*)x, ldx, V, C,
            mean_invstd, gamma, beta, sums, relu, (__nv_bfloat16*)dx, lddx, (__nv_bfloat16*)dres, lddr, inv_n, cg, rl);
      else
        bn_act_bwd_apply_wide_kernel<float, 4, 4><<<grid, 256, 0, st>>>(
            (const float*)dout, ldd, (const float*)out, ldo, (const float*)x, ldx, V, C, mean_invstd, gamma, beta, sums,
            relu, (float*)dx, lddx, (float*)dres, lddr, inv_n, cg, rl);
      return CK_LAUNCH();
    }
  }
  const int g = grid_for(V * (C >> 2));
  if (act_bf16)
    bn_act_bwd_apply_kernel<__nv_bfloat16><<<g, 256, 0, st>>>(
        (const __nv_bfloat16*)dout, ldd, (const __nv_bfloat16*)out, ldo, (const __nv_bfloat16*)x, ldx, V, C,
        mean_invstd, gamma, beta, sums, relu, (__nv_bfloat16*)dx, lddx, (__nv_bfloat16*)dres, lddr, inv_n);
  else
    bn_act_bwd_apply_kernel<float><<<g, 256, 0, st>>>((const float*)dout, ldd, (const float*)out, ldo, (const float*)x,
                                                     ldx, V, C, mean_invstd, gamma, beta, sums, relu, (float*)dx, lddx,
                                                     (float*)dres, lddr, inv_n);
  return CK_LAUNCH();
}

namespace coocc {
// ------------------------------------------------------------------------------------------
// Backward glue of a Linear / conv with bias and ReLU epilogue (the NeRF MLP heads, P/utils/nerf_mlp.py:92-105, the
// fine-stage MLPs): dz = dy * [y > 0] written in the operand type of the following dgrad / wgrad, and the bias gradient
// db[c] = sum_rows dz -- one pass instead of torch's compare + multiply + reduce + convert (4 passes, 0.65 ms per step).
// ------------------------------------------------------------------------------------------
__device__ __forceinline__ float4 ld4_any(const void* p, int bf16, long long idx) {
  return bf16 ? load4(reinterpret_cast<const __nv_bfloat16*>(p) + idx) : load4(reinterpret_cast<const float*>(p) + idx);
}
__device__ __forceinline__ float ld1_any(const void* p, int bf16, long long idx) {
  return bf16 ? load1(reinterpret_cast<const __nv_bfloat16*>(p) + idx) : load1(reinterpret_cast<const float*>(p) + idx);
}

// C % 4 == 0, (C / 4) divides 256: a thread owns 4 channels, a block walks `rows_per_block` rows
__global__ void __launch_bounds__(256) relu_bias_bwd_vec_kernel(const void* __restrict__ dy, long long ld_dy, int dy_bf16,
                                                                const void* __restrict__ y, long long ld_y, int y_bf16,
                                                                int M, int C, void* __restrict__ out, long long ldo,
                                                                int out_bf16, float* __restrict__ db, int rows_per_block) {
  const int tpr = C >> 2;                       // threads per row
  const int rpi = 256 / tpr;                    // rows per iteration
  const int c4 = (threadIdx.x % tpr) * 4;
  const int r0 = blockIdx.x * rows_per_block + threadIdx.x / tpr;
  const int r1 = min(M, (blockIdx.x + 1) * rows_per_block);
  float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
  for (int r = r0; r < r1; r += rpi) {
    float4 d = ld4_any(dy, dy_bf16, (long long)r * ld_dy + c4);
    if (y != nullptr) {
      const float4 o = ld4_any(y, y_bf16, (long long)r * ld_y + c4);
      d.x = o.x > 0.f ? d.x : 0.f; d.y = o.y > 0.f ? d.y : 0.f;
      d.z = o.z > 0.f ? d.z : 0.f; d.w = o.w > 0.f ? d.w : 0.f;
    }
    if (out != nullptr) {
      if (out_bf16) store4(reinterpret_cast<__nv_bfloat16*>(out) + (long long)r * ldo + c4, d);
      else store4(reinterpret_cast<float*>(out) + (long long)r * ldo + c4, d);
    }
    acc.x += d.x; acc.y += d.y; acc.z += d.z; acc.w += d.w;
  }
  if (db == nullptr) return;
  __shared__ float4 sh[256];
  sh[threadIdx.x] = acc;
  __syncthreads();
  if (threadIdx.x < tpr) {
    for (int k = 1; k < rpi; ++k) {
      const float4 o = sh[threadIdx.x + k * tpr];
      acc.x += o.x; acc.y += o.y; acc.z += o.z; acc.w += o.w;
    }
    atomicAdd(&db[c4], acc.x); atomicAdd(&db[c4 + 1], acc.y);
    atomicAdd(&db[c4 + 2], acc.z); atomicAdd(&db[c4 + 3], acc.w);
  }
}

// any C <= 1024: element-wise, bias sums through shared-memory atomics
__global__ void __launch_bounds__(256) relu_bias_bwd_gen_kernel(const void* __restrict__ dy, long long ld_dy, int dy_bf16,
                                                                const void* __restrict__ y, long long ld_y, int y_bf16,
                                                                int M, int C, void* __restrict__ out, long long ldo,
                                                                int out_bf16, float* __restrict__ db) {
  __shared__ float sh[1024];
  for (int c = threadIdx.x; c < C; c += 256) sh[c] = 0.f;
  __syncthreads();
  const long long total = (long long)M * C;
  for (long long i = blockIdx.x * 256LL + threadIdx.x; i < total; i += (long long)gridDim.x * 256LL) {
    const int r = (int)(i / C), c = (int)(i % C);
    float d = ld1_any(dy, dy_bf16, (long long)r * ld_dy + c);
    if (y != nullptr && !(ld1_any(y, y_bf16, (long long)r * ld_y + c) > 0.f)) d = 0.f;
    if (out != nullptr) {
      if (out_bf16) reinterpret_cast<__nv_bfloat16*>(out)[(long long)r * ldo + c] = __float2bfloat16_rn(d);
      else reinterpret_cast<float*>(out)[(long long)r * ldo + c] = d;
    }
    if (db != nullptr) atomicAdd(&sh[c], d);
  }
  if (db == nullptr) return;
  __syncthreads();
  for (int c = threadIdx.x; c < C; c += 256) atomicAdd(&db[c], sh[c]);
}

}  // namespace coocc

extern "C" int coocc_relu_bias_bwd(const void* dy, long long ld_dy, int dy_bf16, const void* y, long long ld_y,
                                   int y_bf16, int M, int C, void* out, long long ldo, int out_bf16, float* db,
                                   void* stream) {
  using namespace coocc;
  if (!dy || M < 0 || C < 1 || C > 1024 || (!out && !db)) return COOCC_ERR_ARG;
  if (M == 0) return 0;
  const int tpr = C / 4;
  const bool vec = (C % 4) == 0 && tpr <= 256 && (256 % tpr) == 0 && (ld_dy % 4) == 0 && (!y || (ld_y % 4) == 0) &&
                   (!out || (ldo % 4) == 0);
  if (vec) {
    // ~4 blocks per SM; every block at least 8 row iterations
    int rpb = (M + 148 * 4 - 1) / (148 * 4);
    const int rpi = 256 / tpr;
    if (rpb < 8 * rpi) rpb = 8 * rpi;
    rpb = (rpb + rpi - 1) / rpi * rpi;
    const int blocks = (M + rpb - 1) / rpb;
    relu_bias_bwd_vec_kernel<<<blocks, 256, 0, (cudaStream_t)stream>>>(dy, ld_dy, dy_bf16, y, ld_y, y_bf16, M, C, out, ldo,
                                                                     out_bf16, db, rpb);
  } else {
    long long blocks = ((long long)M * C + 255) / 256;
    if (blocks > 148 * 8) blocks = 148 * 8;
    relu_bias_bwd_gen_kernel<<<(int)blocks, 256, 0, (cudaStream_t)stream>>>(dy, ld_dy, dy_bf16, y, ld_y, y_bf16, M, C, out,
                                                                          ldo, out_bf16, db);
  }
  return cudaGetLastError() == cudaSuccess ? 0 : COOCC_ERR_CUDA;
}

extern "C" int coocc_dilate2(const void* src, long long lds, int oX, int oY, int oZ, int C, void* dst, long long ldd,
                             int X, int Y, int Z, int is_bf16, void* stream) {
  if (!src || !dst) return COOCC_ERR_ARG;
  const int es = is_bf16 ? 2 : 4;
  if ((C * es) % 16 || (lds * es) % 16 || (ldd * es) % 16) return COOCC_ERR_ALIGN;
  const long long total = (long long)X * Y * Z * (C * es / 16);
  if (is_bf16)
    dilate2_kernel<__nv_bfloat16><<<grid_for(total), 256, 0, (cudaStream_t)stream>>>(
        (const __nv_bfloat16*)src, lds, oX, oY, oZ, C, (__nv_bfloat16*)dst, ldd, X, Y, Z);
  else
    dilate2_kernel<float><<<grid_for(total), 256, 0, (cudaStream_t)stream>>>((const float*)src, lds, oX, oY, oZ, C,
                                                                            (float*)dst, ldd, X, Y, Z);
  return CK_LAUNCH();
}
