// eval_hist.cu -- test-time metric of the occupancy head:
//   COOCC_Ray.evaluation_semantic   P/coocc/detectors/coocc_ray.py:659-684
//   fast_hist                       coocc_ray.py:726-730
// The reference up-samples the [1,C,X,Y,Z] logits to the label grid with
// F.interpolate(mode='trilinear', align_corners=False), takes the argmax, copies prediction and labels
// to the host and builds confusion matrices with np.bincount.  Here one kernel does all of it per label
// voxel -- interpolate the C logits (same source-index / weight arithmetic and summation nesting as
// ATen's upsample_trilinear3d, no FMA contraction), first-maximum argmax, and shared-memory histograms
// flushed with 64-bit atomics -- so only C*C + C*C + 4 counters ever leave the GPU.
#include <cuda_runtime.h>
#include <stdint.h>

#include "../../include/coocc_b200.h"

namespace coocc {

constexpr int kEvalMaxCls = 32;

struct Axis { int i0, i1; float w0, w1; };

// area_pixel_compute_source_index(scale, dst, align_corners=false, cubic=false) and the lambda pair
__device__ __forceinline__ Axis axis_of(int dst, int in, int out) {
  Axis a;
  if (in == out) { a.i0 = dst; a.i1 = dst; a.w0 = 1.f; a.w1 = 0.f; return a; }
  const float scale = (float)in / (float)out;
  float src = __fsub_rn(__fmul_rn(scale, __fadd_rn((float)dst, 0.5f)), 0.5f);
  if (src < 0.f) src = 0.f;
  const int i0 = (int)src;
  a.i0 = i0;
  a.i1 = i0 + (i0 < in - 1 ? 1 : 0);
  a.w1 = __fsub_rn(src, (float)i0);
  a.w0 = __fsub_rn(1.f, a.w1);
  return a;
}

template <typename T>
__global__ void __launch_bounds__(256) eval_confusion_kernel(const float* __restrict__ logits, long long ld, int X,
                                                             int Y, int Z, int C, const T* __restrict__ gt, int GX,
                                                             int GY, int GZ, const unsigned char* __restrict__ vis,
                                                             int empty_idx, int ignore,
                                                             unsigned long long* __restrict__ h_ssc,
                                                             unsigned long long* __restrict__ h_vis,
                                                             unsigned long long* __restrict__ h_sc) {
  __shared__ unsigned int s_ssc[kEvalMaxCls * kEvalMaxCls], s_vis[kEvalMaxCls * kEvalMaxCls], s_sc[4];
  for (int i = threadIdx.x; i < C * C; i += blockDim.x) { s_ssc[i] = 0; s_vis[i] = 0; }
  if (threadIdx.x < 4) s_sc[threadIdx.x] = 0;
  __syncthreads();
  const long long total = (long long)GX * GY * GZ;
  for (long long g = blockIdx.x * (long long)blockDim.x + threadIdx.x; g < total;
       g += (long long)gridDim.x * blockDim.x) {
    const int label = (int)gt[g];
    if (label == ignore) continue;                              // noise_mask (coocc_ray.py:666)
    const int gz = (int)(g % GZ), gy = (int)((g / GZ) % GY), gx = (int)(g / ((long long)GZ * GY));
    const Axis ax = axis_of(gx, X, GX), ay = axis_of(gy, Y, GY), az = axis_of(gz, Z, GZ);
    const float* r000 = logits + (((long long)ax.i0 * Y + ay.i0) * Z + az.i0) * ld;
    const float* r001 = logits + (((long long)ax.i0 * Y + ay.i0) * Z + az.i1) * ld;
    const float* r010 = logits + (((long long)ax.i0 * Y + ay.i1) * Z + az.i0) * ld;
    const float* r011 = logits + (((long long)ax.i0 * Y + ay.i1) * Z + az.i1) * ld;
    const float* r100 = logits + (((long long)ax.i1 * Y + ay.i0) * Z + az.i0) * ld;
    const float* r101 = logits + (((long long)ax.i1 * Y + ay.i0) * Z + az.i1) * ld;
    const float* r110 = logits + (((long long)ax.i1 * Y + ay.i1) * Z + az.i0) * ld;
    const float* r111 = logits + (((long long)ax.i1 * Y + ay.i1) * Z + az.i1) * ld;
    int best = 0;
    float bestv = 0.f;
    for (int c = 0; c < C; ++c) {
      // value = wx0*(wy0*(wz0*v000 + wz1*v001) + wy1*(...)) + wx1*(...), innermost axis first
      const float a00 = __fadd_rn(__fmul_rn(r000[c], az.w0), __fmul_rn(r001[c], az.w1));
      const float a01 = __fadd_rn(__fmul_rn(r010[c], az.w0), __fmul_rn(r011[c], az.w1));
      const float a10 = __fadd_rn(__fmul_rn(r100[c], az.w0), __fmul_rn(r101[c], az.w1));
      const float a11 = __fadd_rn(__fmul_rn(r110[c], az.w0), __fmul_rn(r111[c], az.w1));
      const float b0 = __fadd_rn(__fmul_rn(a00, ay.w0), __fmul_rn(a01, ay.w1));
      const float b1 = __fadd_rn(__fmul_rn(a10, ay.w0), __fmul_rn(a11, ay.w1));
      const float v = __fadd_rn(__fmul_rn(b0, ax.w0), __fmul_rn(b1, ax.w1));
      if (c == 0 || v > bestv) { best = c; bestv = v; }       // first maximum, like torch.argmax
    }
    if (label >= 0 && label < C) {
      atomicAdd(&s_ssc[label * C + best], 1u);
      if (vis && vis[g] != 0) atomicAdd(&s_vis[label * C + best], 1u);
    }
    atomicAdd(&s_sc[(label != empty_idx ? 2 : 0) + (best != empty_idx ? 1 : 0)], 1u);
  }
  __syncthreads();
  for (int i = threadIdx.x; i < C * C; i += blockDim.x) {
    if (s_ssc[i]) atomicAdd(&h_ssc[i], (unsigned long long)s_ssc[i]);
    if (h_vis && s_vis[i]) atomicAdd(&h_vis[i], (unsigned long long)s_vis[i]);
  }
  if (threadIdx.x < 4 && s_sc[threadIdx.x]) atomicAdd(&h_sc[threadIdx.x], (unsigned long long)s_sc[threadIdx.x]);
}

}  // namespace coocc

using namespace coocc;

extern "C" int coocc_eval_confusion(const float* logits, long long ld, int X, int Y, int Z, int C, const void* gt,
                                    int gt_bytes, int GX, int GY, int GZ, const unsigned char* visible, int empty_idx,
                                    int ignore, long long* hist_ssc, long long* hist_ssc_visible, long long* hist_sc,
                                    void* stream) {
  if (!logits || !gt || !hist_ssc || !hist_sc || X < 1 || Y < 1 || Z < 1 || GX < 1 || GY < 1 || GZ < 1 || C < 2 ||
      C > kEvalMaxCls || ld < C)
    return COOCC_ERR_ARG;
  if (visible && !hist_ssc_visible) return COOCC_ERR_ARG;
  cudaStream_t st = (cudaStream_t)stream;
  const long long total = (long long)GX * GY * GZ;
  int blocks = (int)((total + 255) / 256);
  if (blocks > 148 * 8) blocks = 148 * 8;
  auto hs = reinterpret_cast<unsigned long long*>(hist_ssc);
  auto hv = reinterpret_cast<unsigned long long*>(hist_ssc_visible);
  auto hc = reinterpret_cast<unsigned long long*>(hist_sc);
  if (cudaMemsetAsync(hs, 0, sizeof(long long) * C * C, st) != cudaSuccess) return COOCC_ERR_CUDA;
  if (hv && cudaMemsetAsync(hv, 0, sizeof(long long) * C * C, st) != cudaSuccess) return COOCC_ERR_CUDA;
  if (cudaMemsetAsync(hc, 0, sizeof(long long) * 4, st) != cudaSuccess) return COOCC_ERR_CUDA;
  if (gt_bytes == 8)
    eval_confusion_kernel<long long><<<blocks, 256, 0, st>>>(logits, ld, X, Y, Z, C, (const long long*)gt, GX, GY, GZ,
                                                            visible, empty_idx, ignore, hs, hv, hc);
  else if (gt_bytes == 4)
    eval_confusion_kernel<int><<<blocks, 256, 0, st>>>(logits, ld, X, Y, Z, C, (const int*)gt, GX, GY, GZ, visible,
                                                      empty_idx, ignore, hs, hv, hc);
  else if (gt_bytes == 1)
    eval_confusion_kernel<unsigned char><<<blocks, 256, 0, st>>>(logits, ld, X, Y, Z, C, (const unsigned char*)gt, GX,
                                                                GY, GZ, visible, empty_idx, ignore, hs, hv, hc);
  else
    return COOCC_ERR_ARG;
  return cudaGetLastError() == cudaSuccess ? 0 : COOCC_ERR_CUDA;
}
