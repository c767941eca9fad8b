// trilinear.cu -- F.interpolate(mode='trilinear', align_corners=False, size=...) on NDHWC rows,
// fused with the additions / per-voxel weights that follow it in the reference:
//   FPN3D top-down path   laterals[i-1] = laterals[i-1] + interpolate(laterals[i])   (P/coocc/necks/fpn3d.py:91-94)
//   OccHead level fusion  out += interpolate(feats) * softmax_weight[:, level]         (P/coocc/dense_heads/occ_head.py:161-165)
// out[v, :] = base[v, :] + wts[v] * sum_{8 corners} lambda * src[corner, :]
// Backward w.r.t. src is a *gather* over the output voxels that reference an input voxel (no atomics,
// deterministic); w.r.t. wts a per-voxel dot product.
#include <cuda_runtime.h>
#include <stdint.h>

#include "../../include/coocc_b200.h"
#include "act_types.cuh"

namespace coocc {

// ATen's area_pixel_compute_source_index(scale = in/out, align_corners = false)
__device__ __forceinline__ void tl_src(int o, float scale, int in_size, int& i0, int& i1, float& l1) {
  float s = scale * ((float)o + 0.5f) - 0.5f;
  if (s < 0.f) s = 0.f;
  i0 = (int)s;
  if (i0 > in_size - 1) i0 = in_size - 1;
  i1 = i0 + (i0 < in_size - 1 ? 1 : 0);
  l1 = s - (float)i0;
}

struct TlDims {
  int sX, sY, sZ, oX, oY, oZ;
  float fx, fy, fz;   // in / out
};

template <typename T>
__global__ void __launch_bounds__(256) trilinear_fwd_kernel(const T* __restrict__ src, long long lds,
                                                            TlDims d, int C, const T* __restrict__ base,
                                                            long long ldb, const float* __restrict__ wts,
                                                            long long ldw, T* __restrict__ out, long long ldo) {
  const int c4 = C >> 2;
  const long long total = (long long)d.oX * d.oY * d.oZ * c4;
  for (long long i = blockIdx.x * 256LL + threadIdx.x; i < total; i += gridDim.x * 256LL) {
    const long long v = i / c4;
    const int c = (int)(i % c4) * 4;
    const int z = v % d.oZ, y = (v / d.oZ) % d.oY, x = v / ((long long)d.oZ * d.oY);
    int x0, x1, y0, y1, z0, z1;
    float lx, ly, lz;
    tl_src(x, d.fx, d.sX, x0, x1, lx);
    tl_src(y, d.fy, d.sY, y0, y1, ly);
    tl_src(z, d.fz, d.sZ, z0, z1, lz);
    float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
    for (int k = 0; k < 8; ++k) {
      const int xi = (k & 4) ? x1 : x0, yi = (k & 2) ? y1 : y0, zi = (k & 1) ? z1 : z0;
      const float w = ((k & 4) ? lx : 1.f - lx) * ((k & 2) ? ly : 1.f - ly) * ((k & 1) ? lz : 1.f - lz);
      const float4 s = load4(src + (((long long)xi * d.sY + yi) * d.sZ + zi) * lds + c);
      acc.x += w * s.x; acc.y += w * s.y; acc.z += w * s.z; acc.w += w * s.w;
    }
    if (wts != nullptr) {
      const float w = wts[v * ldw];
      acc.x *= w; acc.y *= w; acc.z *= w; acc.w *= w;
    }
    if (base != nullptr) {
      const float4 b = load4(base + v * ldb + c);
      acc.x += b.x; acc.y += b.y; acc.z += b.z; acc.w += b.w;
    }
    store4(out + v * ldo + c, acc);
  }
}

// range of output indices whose interpolation may touch input index i (widened by one on both sides;
// the exact membership is re-derived per candidate)
__device__ __forceinline__ void tl_range(int i, float scale, int out_size, int& lo, int& hi) {
  const float inv = 1.f / scale;
  lo = (int)floorf(((float)i - 0.5f) * inv - 0.5f) - 1;
  hi = (int)ceilf(((float)i + 1.5f) * inv - 0.5f) + 1;
  if (lo < 0) lo = 0;
  if (hi > out_size - 1) hi = out_size - 1;
}

// per-axis gather lists: for input index i, the output indices o in [lo, lo+n) and their weights
constexpr int kTlMax = 24;     // enough for x8 upsampling (<= 2*8 + 3 candidates per axis)

// one warp per input voxel (lanes = channel groups of 4): the three per-axis weight lists are built
// once per voxel (lane j computes candidate j) in shared memory, the inner loops are pure
// multiply-adds over coalesced 16-byte loads
template <typename T>
__global__ void __launch_bounds__(256) trilinear_bwd_kernel(const T* __restrict__ dout, long long ldd,
                                                            TlDims d, int C, const float* __restrict__ wts,
                                                            long long ldw, T* __restrict__ dsrc,
                                                            long long lds) {
  __shared__ float sw[8][3][kTlMax];
  const int wid = threadIdx.x >> 5;
  const long long v = blockIdx.x * 8LL + wid;
  const int lane = threadIdx.x & 31;
  if (v >= (long long)d.sX * d.sY * d.sZ) return;
  const int z = v % d.sZ, y = (v / d.sZ) % d.sY, x = v / ((long long)d.sZ * d.sY);
  int lo[3], n[3];
  {
    const int idx[3] = {x, y, z};
    const float sc[3] = {d.fx, d.fy, d.fz};
    const int in_sz[3] = {d.sX, d.sY, d.sZ};
    const int out_sz[3] = {d.oX, d.oY, d.oZ};
#pragma unroll
    for (int a = 0; a < 3; ++a) {
      int hi;
      tl_range(idx[a], sc[a], out_sz[a], lo[a], hi);
      n[a] = min(hi - lo[a] + 1, kTlMax);
      if (lane < n[a]) {
        int a0, a1;
        float la;
        tl_src(lo[a] + lane, sc[a], in_sz[a], a0, a1, la);
        sw[wid][a][lane] = (a0 == idx[a] ? 1.f - la : 0.f) + (a1 == idx[a] ? la : 0.f);
      }
    }
  }
  __syncwarp();
  const float* wx = sw[wid][0];
  const float* wy = sw[wid][1];
  const float* wz = sw[wid][2];
  const int c4 = C >> 2;
  for (int cg = lane; cg < c4; cg += 32) {
    const int c = cg * 4;
    float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
    for (int ix = 0; ix < n[0]; ++ix) {
      const float wxv = wx[ix];
      if (wxv == 0.f) continue;
      for (int iy = 0; iy < n[1]; ++iy) {
        const float wxy = wxv * wy[iy];
        if (wxy == 0.f) continue;
        const long long row0 = ((long long)(lo[0] + ix) * d.oY + (lo[1] + iy)) * d.oZ + lo[2];
        for (int iz = 0; iz < n[2]; ++iz) {
          float w = wxy * wz[iz];
          if (w == 0.f) continue;
          const long long ov = row0 + iz;
          if (wts != nullptr) w *= wts[ov * ldw];
          const float4 g = load4(dout + ov * ldd + c);
          acc.x += w * g.x; acc.y += w * g.y; acc.z += w * g.z; acc.w += w * g.w;
        }
      }
    }
    store4(dsrc + v * lds + c, acc);
  }
}

// dw[v] = sum_c dout[v, c] * interpolate(src)[v, c]     (one warp per output voxel)
template <typename T>
__global__ void __launch_bounds__(256) trilinear_wgrad_kernel(const T* __restrict__ dout, long long ldd,
                                                              const T* __restrict__ src, long long lds,
                                                              TlDims d, int C, float* __restrict__ dw,
                                                              long long lddw) {
  const long long v = blockIdx.x * 8LL + (threadIdx.x >> 5);
  const int lane = threadIdx.x & 31;
  if (v >= (long long)d.oX * d.oY * d.oZ) return;
  const int z = v % d.oZ, y = (v / d.oZ) % d.oY, x = v / ((long long)d.oZ * d.oY);
  int x0, x1, y0, y1, z0, z1;
  float lx, ly, lz;
  tl_src(x, d.fx, d.sX, x0, x1, lx);
  tl_src(y, d.fy, d.sY, y0, y1, ly);
  tl_src(z, d.fz, d.sZ, z0, z1, lz);
  float acc = 0.f;
  for (int c = lane; c < C; c += 32) {
    float u = 0.f;
#pragma unroll
    for (int k = 0; k < 8; ++k) {
      const int xi = (k & 4) ? x1 : x0, yi = (k & 2) ? y1 : y0, zi = (k & 1) ? z1 : z0;
      const float w = ((k & 4) ? lx : 1.f - lx) * ((k & 2) ? ly : 1.f - ly) * ((k & 1) ? lz : 1.f - lz);
      u += w * load1(src + (((long long)xi * d.sY + yi) * d.sZ + zi) * lds + c);
    }
    acc += u * load1(dout + v * ldd + c);
  }
#pragma unroll
  for (int o = 16; o >= 1; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
  if (lane == 0) dw[v * lddw] = acc;
}

static TlDims make_dims(int sX, int sY, int sZ, int oX, int oY, int oZ) {
  TlDims d;
  d.sX = sX; d.sY = sY; d.sZ = sZ; d.oX = oX; d.oY = oY; d.oZ = oZ;
  d.fx = (float)sX / (float)oX;
  d.fy = (float)sY / (float)oY;
  d.fz = (float)sZ / (float)oZ;
  return d;
}
static int tl_grid(long long total) {
  long long b = (total + 255) / 256;
  if (b > 148LL * 32) b = 148LL * 32;
  return b < 1 ? 1 : (int)b;
}

}  // namespace coocc

using namespace coocc;
#define CK_LAUNCH() (cudaGetLastError() == cudaSuccess ? 0 : COOCC_ERR_CUDA)

typedef __nv_bfloat16 bf16_t;

extern "C" int coocc_trilinear_fwd(const void* src, long long lds, int sX, int sY, int sZ, int C,
                                   const void* base, long long ldb, const float* wts, long long ldw, void* out,
                                   long long ldo, int oX, int oY, int oZ, int act_bf16, void* stream) {
  if (!src || !out || (C & 3) || (lds & 3) || (ldo & 3) || (base && (ldb & 3))) return COOCC_ERR_ARG;
  const TlDims d = make_dims(sX, sY, sZ, oX, oY, oZ);
  const int g = tl_grid((long long)oX * oY * oZ * (C >> 2));
  if (act_bf16)
    trilinear_fwd_kernel<bf16_t><<<g, 256, 0, (cudaStream_t)stream>>>((const bf16_t*)src, lds, d, C, (const bf16_t*)base,
                                                                    ldb, wts, ldw, (bf16_t*)out, ldo);
  else
    trilinear_fwd_kernel<float><<<g, 256, 0, (cudaStream_t)stream>>>((const float*)src, lds, d, C, (const float*)base,
                                                                   ldb, wts, ldw, (float*)out, ldo);
  return CK_LAUNCH();
}

extern "C" int coocc_trilinear_bwd(const void* dout, long long ldd, int oX, int oY, int oZ, int C,
                                   const float* wts, long long ldw, void* dsrc, long long lds, int sX, int sY,
                                   int sZ, int act_bf16, void* stream) {
  if (!dout || !dsrc || (C & 3) || (lds & 3) || (ldd & 3)) return COOCC_ERR_ARG;
  const TlDims d = make_dims(sX, sY, sZ, oX, oY, oZ);
  const long long Vs = (long long)sX * sY * sZ;
  const unsigned g = (unsigned)((Vs + 7) / 8);
  if (act_bf16)
    trilinear_bwd_kernel<bf16_t><<<g, 256, 0, (cudaStream_t)stream>>>((const bf16_t*)dout, ldd, d, C, wts, ldw,
                                                                    (bf16_t*)dsrc, lds);
  else
    trilinear_bwd_kernel<float><<<g, 256, 0, (cudaStream_t)stream>>>((const float*)dout, ldd, d, C, wts, ldw,
                                                                   (float*)dsrc, lds);
  return CK_LAUNCH();
}

extern "C" int coocc_trilinear_wgrad(const void* dout, long long ldd, const void* src, long long lds, int sX,
                                     int sY, int sZ, int oX, int oY, int oZ, int C, float* dw, long long lddw,
                                     int act_bf16, void* stream) {
  if (!dout || !src || !dw) return COOCC_ERR_ARG;
  const TlDims d = make_dims(sX, sY, sZ, oX, oY, oZ);
  const long long V = (long long)oX * oY * oZ;
  const unsigned g = (unsigned)((V + 7) / 8);
  if (act_bf16)
    trilinear_wgrad_kernel<bf16_t><<<g, 256, 0, (cudaStream_t)stream>>>((const bf16_t*)dout, ldd, (const bf16_t*)src, lds,
                                                                      d, C, dw, lddw);
  else
    trilinear_wgrad_kernel<float><<<g, 256, 0, (cudaStream_t)stream>>>((const float*)dout, ldd, (const float*)src, lds,
                                                                     d, C, dw, lddw);
  return CK_LAUNCH();
}
