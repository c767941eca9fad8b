// lss_pool.cu -- Lift-Splat voxel pooling and frustum geometry, the step that
// produces the path's img_voxel_feats and the geom tensor the renderer re-reads:
//   ViewTransformerLiftSplatShootVoxel.voxel_pooling   P/coocc/image2bev/ViewTransformerLSSVoxel.py:100-123
//   bev_pool (argsort by voxel rank + interval sum)    M/ops/bev_pool/bev_pool.py:80-97,
//                                                      M/ops/bev_pool/src/bev_pool_cuda.cu:20-98
//   get_geometry                                       P/coocc/image2bev/ViewTransformerLSSBEVDepth.py:117-150
// Design: frustum point -> voxel key (same fp32 arithmetic and truncation as the reference), a stable
// radix sort of (key, point id), then one warp per voxel sums its run of points with float4 row loads and
// writes the NDHWC output row -- empty voxels get their zeros from the same kernel, no memset, no atomics.
// The fused mode never materialises the reference's [N,D,H,W,C] "volume" (1.9 GB for the r101 config):
// a point's feature row is depth_prob[n,d,h,w] * img_feat[n,h,w,:] formed on the fly.
#include <cuda_runtime.h>
#include <stdint.h>

#include "../../include/coocc_b200.h"
#include "radix_sort.cuh"

namespace coocc {

// points = frustum - post_trans; inv(post_rots) @ points; (x*z, y*z, z); (rots @ inv(intrins)) @ points + trans;
// bda @ points   (ViewTransformerLSSBEVDepth.py:126-148).  mats: per camera 3+9+9+3 floats.
__global__ void lss_geometry_kernel(const float* __restrict__ frustum, int ncam, long long per_cam,
                                    const float* __restrict__ mats, const float* __restrict__ bda,
                                    float* __restrict__ geom) {
  const long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x;
  if (i >= per_cam * ncam) return;
  const int cam = (int)(i / per_cam);
  const long long j = i % per_cam;
  const float* m = mats + cam * 24;
  float p0 = frustum[j * 3 + 0] - m[0], p1 = frustum[j * 3 + 1] - m[1], p2 = frustum[j * 3 + 2] - m[2];
  const float* R = m + 3;
  float q0 = R[0] * p0 + R[1] * p1 + R[2] * p2;
  float q1 = R[3] * p0 + R[4] * p1 + R[5] * p2;
  float q2 = R[6] * p0 + R[7] * p1 + R[8] * p2;
  q0 *= q2;
  q1 *= q2;
  const float* Cm = m + 12;
  const float* t = m + 21;
  p0 = Cm[0] * q0 + Cm[1] * q1 + Cm[2] * q2 + t[0];
  p1 = Cm[3] * q0 + Cm[4] * q1 + Cm[5] * q2 + t[1];
  p2 = Cm[6] * q0 + Cm[7] * q1 + Cm[8] * q2 + t[2];
  geom[i * 3 + 0] = bda[0] * p0 + bda[1] * p1 + bda[2] * p2;
  geom[i * 3 + 1] = bda[3] * p0 + bda[4] * p1 + bda[5] * p2;
  geom[i * 3 + 2] = bda[6] * p0 + bda[7] * p1 + bda[8] * p2;
}

// key = voxel id (x*Y + y)*Z + z of ((geom - lo) / dx).long(), or V when the point is dropped
// (ViewTransformerLSSVoxel.py:107, 113-117: truncation toward zero, then 0 <= idx < nx)
__global__ void lss_keys_kernel(const float* __restrict__ geom, long long npts, float lo0, float lo1, float lo2,
                                float dx0, float dx1, float dx2, int X, int Y, int Z, uint32_t* __restrict__ keys0,
                                uint32_t* __restrict__ keys, uint32_t* __restrict__ vals) {
  const long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x;
  if (i >= npts) return;
  const float t0 = __fdiv_rn(__fsub_rn(geom[i * 3 + 0], lo0), dx0);
  const float t1 = __fdiv_rn(__fsub_rn(geom[i * 3 + 1], lo1), dx1);
  const float t2 = __fdiv_rn(__fsub_rn(geom[i * 3 + 2], lo2), dx2);
  uint32_t key = (uint32_t)(X * Y * Z);
  const bool finite = fabsf(t0) < 1e9f && fabsf(t1) < 1e9f && fabsf(t2) < 1e9f;
  if (finite) {
    const int x = (int)t0, y = (int)t1, z = (int)t2;        // trunc toward zero: (-1, 0) lands in cell 0
    if (x >= 0 && x < X && y >= 0 && y < Y && z >= 0 && z < Z) key = (uint32_t)((x * Y + y) * Z + z);
  }
  keys0[i] = key;
  keys[i] = key;
  vals[i] = (uint32_t)i;
}

// seg[v] = first sorted position whose key is >= v, for v in [0, V+1]: position j writes the entries of the
// keys in (key[j-1], key[j]], the last position also those above key[n-1] -- every entry exactly once.
__global__ void lss_segments_kernel(const uint32_t* __restrict__ skeys, long long npts, int V, int* __restrict__ seg) {
  const long long j = blockIdx.x * (long long)blockDim.x + threadIdx.x;
  if (j >= npts) return;
  const uint32_t k = skeys[j];
  const long long prev = j == 0 ? -1 : (long long)skeys[j - 1];
  for (long long v = prev + 1; v <= (long long)k && v <= V; ++v) seg[v] = (int)j;
  if (j == npts - 1)
    for (long long v = (long long)k + 1; v <= V; ++v) seg[v] = (int)npts;
}

// one warp per voxel: out[v,:] = sum over its run of sorted points of w_i * F[row_i,:]
//   plain mode : row_i = i, w_i = 1                     (F = the reference's flattened volume [npts, C])
//   fused mode : i = ((n*D + d)*HW + hw) -> row_i = n*HW + hw, w_i = depth[i]   (F = img_feat as [ncam*HW, C])
// The run is walked 32 points at a time: the lanes fetch the 32 point ids / weights together, then the row
// loads of the batch are independent of each other (addresses come from shuffles), so several 512-byte rows
// are in flight per warp.  C <= 512 (up to 4 float4 accumulators per lane).
__global__ void __launch_bounds__(256) lss_pool_fwd_kernel(const uint32_t* __restrict__ skeys,
                                                           const uint32_t* __restrict__ svals, long long npts, int V,
                                                           int C, const float* __restrict__ F, long long ldf,
                                                           const float* __restrict__ depth, int D, int HW,
                                                           const int* __restrict__ seg, float* __restrict__ out,
                                                           long long ldo) {
  const int v = blockIdx.x * 8 + (threadIdx.x >> 5);
  const int lane = threadIdx.x & 31;
  if (v >= V) return;
  const int beg = seg[v], end = seg[v + 1];
  float4 acc[4];
#pragma unroll
  for (int u = 0; u < 4; ++u) acc[u] = make_float4(0.f, 0.f, 0.f, 0.f);
  const long long DHW = (long long)D * HW;
  for (int j0 = beg; j0 < end; j0 += 32) {
    const int j = j0 + lane;
    int row = 0;
    float w = 0.f;
    if (j < end) {
      const uint32_t i = svals[j];
      if (depth) {
        row = (int)((i / DHW) * HW + (i % HW));
        w = depth[i];
      } else {
        row = (int)i;
        w = 1.f;
      }
    }
    const int cnt = min(32, end - j0);
#pragma unroll 4
    for (int t = 0; t < cnt; ++t) {
      const int r = __shfl_sync(0xffffffffu, row, t);
      const float ww = __shfl_sync(0xffffffffu, w, t);
      const float* src = F + (long long)r * ldf;
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        const int c0 = lane * 4 + u * 128;
        if (c0 < C) {
          const float4 f = *reinterpret_cast<const float4*>(src + c0);
          acc[u].x += ww * f.x; acc[u].y += ww * f.y; acc[u].z += ww * f.z; acc[u].w += ww * f.w;
        }
      }
    }
  }
#pragma unroll
  for (int u = 0; u < 4; ++u) {
    const int c0 = lane * 4 + u * 128;
    if (c0 < C) *reinterpret_cast<float4*>(out + (long long)v * ldo + c0) = acc[u];
  }
}

// plain mode backward: d_volume[i,:] = gout[key_i,:] (0 for dropped points)   (bev_pool_cuda.cu:62-86)
__global__ void __launch_bounds__(256) lss_pool_bwd_plain_kernel(const uint32_t* __restrict__ keys0, long long npts,
                                                                 int V, int C, const float* __restrict__ gout,
                                                                 long long ldg, float* __restrict__ dvol,
                                                                 long long ldv) {
  const long long i = blockIdx.x * 8LL + (threadIdx.x >> 5);
  const int lane = threadIdx.x & 31;
  if (i >= npts) return;
  const uint32_t k = keys0[i];
  for (int c0 = lane * 4; c0 < C; c0 += 128) {
    float4 g = make_float4(0.f, 0.f, 0.f, 0.f);
    if (k < (uint32_t)V) g = *reinterpret_cast<const float4*>(gout + (long long)k * ldg + c0);
    *reinterpret_cast<float4*>(dvol + i * ldv + c0) = g;
  }
}

// fused mode backward, one warp per pixel (n, hw): walks its D depth samples,
//   d_depth[i] = <gout[key_i,:], feat[row,:]>,   d_feat[row,:] = sum_d depth[i] * gout[key_i,:]
__global__ void __launch_bounds__(256) lss_pool_bwd_fused_kernel(const uint32_t* __restrict__ keys0, int ncam, int D,
                                                                 int HW, int V, int C,
                                                                 const float* __restrict__ feat, long long ldf,
                                                                 const float* __restrict__ depth,
                                                                 const float* __restrict__ gout, long long ldg,
                                                                 float* __restrict__ dfeat, long long lddf,
                                                                 float* __restrict__ ddepth) {
  const long long row = blockIdx.x * 8LL + (threadIdx.x >> 5);
  const int lane = threadIdx.x & 31;
  if (row >= (long long)ncam * HW) return;
  const long long n = row / HW, hw = row % HW;
  // C <= 512: up to 4 float4 per lane
  float4 f[4], acc[4];
#pragma unroll
  for (int u = 0; u < 4; ++u) {
    const int c0 = lane * 4 + u * 128;
    f[u] = c0 < C ? *reinterpret_cast<const float4*>(feat + row * ldf + c0) : make_float4(0.f, 0.f, 0.f, 0.f);
    acc[u] = make_float4(0.f, 0.f, 0.f, 0.f);
  }
  for (int d0 = 0; d0 < D; d0 += 32) {
    // the lanes fetch key / weight of 32 depth bins together; the gout rows of the batch are then independent
    const int dl = d0 + lane;
    uint32_t kl = (uint32_t)V;
    float wl = 0.f;
    if (dl < D) {
      const long long il = (n * D + dl) * HW + hw;
      kl = keys0[il];
      wl = depth[il];
    }
    const int cnt = min(32, D - d0);
    float mydot = 0.f;                       // lane t keeps the dot product of depth bin d0 + t
#pragma unroll 4
    for (int t = 0; t < cnt; ++t) {
      const uint32_t k = __shfl_sync(0xffffffffu, kl, t);
      const float w = __shfl_sync(0xffffffffu, wl, t);
      float dot = 0.f;
      if (k < (uint32_t)V) {
#pragma unroll
        for (int u = 0; u < 4; ++u) {
          const int c0 = lane * 4 + u * 128;
          if (c0 < C) {
            const float4 g = *reinterpret_cast<const float4*>(gout + (long long)k * ldg + c0);
            dot += g.x * f[u].x + g.y * f[u].y + g.z * f[u].z + g.w * f[u].w;
            acc[u].x += w * g.x; acc[u].y += w * g.y; acc[u].z += w * g.z; acc[u].w += w * g.w;
          }
        }
      }
#pragma unroll
      for (int o = 16; o >= 1; o >>= 1) dot += __shfl_xor_sync(0xffffffffu, dot, o);
      if (lane == t) mydot = dot;
    }
    if (dl < D) ddepth[(n * D + dl) * HW + hw] = mydot;
  }
#pragma unroll
  for (int u = 0; u < 4; ++u) {
    const int c0 = lane * 4 + u * 128;
    if (c0 < C) *reinterpret_cast<float4*>(dfeat + row * lddf + c0) = acc[u];
  }
}

struct LssWs {
  uint32_t *keys0, *kA, *kB, *vA, *vB;
  int* counters;
  int* seg;
};
static size_t lss_align(size_t x) { return (x + 255) / 256 * 256; }
static size_t lss_layout(char* base, long long npts, long long V, LssWs* w) {
  size_t off = 0;
  auto take = [&](size_t bytes) { char* p = base ? base + off : nullptr; off += lss_align(bytes); return p; };
  uint32_t* k0 = reinterpret_cast<uint32_t*>(take((size_t)npts * 4));
  uint32_t* kA = reinterpret_cast<uint32_t*>(take((size_t)npts * 4));
  uint32_t* kB = reinterpret_cast<uint32_t*>(take((size_t)npts * 4));
  uint32_t* vA = reinterpret_cast<uint32_t*>(take((size_t)npts * 4));
  uint32_t* vB = reinterpret_cast<uint32_t*>(take((size_t)npts * 4));
  int* cnt = reinterpret_cast<int*>(take(radix_counters_bytes((int)npts, 1)));
  int* seg = reinterpret_cast<int*>(take((size_t)(V + 2) * 4));
  if (w) *w = LssWs{k0, kA, kB, vA, vB, cnt, seg};
  return off;
}

}  // namespace coocc

using namespace coocc;

extern "C" int coocc_lss_geometry(const float* frustum, int ncam, int D, int H, int W, const float* cam_mats,
                                  const float* bda, float* geom, void* stream) {
  if (!frustum || !cam_mats || !bda || !geom || ncam < 1 || D < 1 || H < 1 || W < 1) return COOCC_ERR_ARG;
  const long long per_cam = (long long)D * H * W;
  const long long n = per_cam * ncam;
  lss_geometry_kernel<<<(unsigned)((n + 255) / 256), 256, 0, (cudaStream_t)stream>>>(frustum, ncam, per_cam, cam_mats,
                                                                                   bda, geom);
  return cudaGetLastError() == cudaSuccess ? 0 : COOCC_ERR_CUDA;
}

extern "C" long long coocc_lss_workspace(long long npts, long long V) {
  if (npts < 1 || npts > 0x7fffffffLL || V < 1 || V >= (1LL << 30)) return -1;
  return (long long)lss_layout(nullptr, npts, V, nullptr);
}

// Sorts the points by voxel.  After the call the workspace holds the per-point keys (for the backward) and
// the sorted (key, point id) arrays; sorted_keys / sorted_vals receive pointers into the workspace.
extern "C" int coocc_lss_sort(const float* geom, long long npts, const float* lo3, const float* dx3, int X, int Y,
                              int Z, void* workspace, const unsigned int** sorted_keys,
                              const unsigned int** sorted_vals, const unsigned int** point_keys,
                              const int** segments, void* stream) {
  if (!geom || !lo3 || !dx3 || !workspace || !sorted_keys || !sorted_vals || npts < 1 || npts > 0x7fffffffLL || X < 1 ||
      Y < 1 || Z < 1)
    return COOCC_ERR_ARG;
  const long long V = (long long)X * Y * Z;
  if (V >= (1LL << 30)) return COOCC_ERR_CAPACITY;
  cudaStream_t st = (cudaStream_t)stream;
  LssWs w;
  lss_layout(reinterpret_cast<char*>(workspace), npts, V, &w);
  lss_keys_kernel<<<(unsigned)((npts + 255) / 256), 256, 0, st>>>(geom, npts, lo3[0], lo3[1], lo3[2], dx3[0], dx3[1],
                                                                 dx3[2], X, Y, Z, w.keys0, w.kA, w.vA);
  int bits = 1;
  while ((1LL << bits) <= V) ++bits;            // keys in [0, V]
  uint32_t *ko = nullptr, *vo = nullptr;
  if (radix_sort_pairs(w.kA, w.vA, w.kB, w.vB, (int)npts, 1, bits, w.counters, st, &ko, &vo) != 0) return COOCC_ERR_CUDA;
  lss_segments_kernel<<<(unsigned)((npts + 255) / 256), 256, 0, st>>>(ko, npts, (int)V, w.seg);
  *sorted_keys = ko;
  *sorted_vals = vo;
  if (point_keys) *point_keys = w.keys0;
  if (segments) *segments = w.seg;
  return cudaGetLastError() == cudaSuccess ? 0 : COOCC_ERR_CUDA;
}

extern "C" int coocc_lss_pool_fwd(const unsigned int* sorted_keys, const unsigned int* sorted_vals,
                                  const int* segments, long long npts, int V, int C, const float* feat, long long ldf,
                                  const float* depth, int D, int HW, float* out, long long ldo, void* stream) {
  if (!sorted_keys || !sorted_vals || !segments || !feat || !out || npts < 1 || V < 1 || C < 4 || C % 4 || C > 512 ||
      ldf % 4 || ldo % 4)
    return COOCC_ERR_ARG;
  if (depth && (D < 1 || HW < 1)) return COOCC_ERR_ARG;
  lss_pool_fwd_kernel<<<(V + 7) / 8, 256, 0, (cudaStream_t)stream>>>(sorted_keys, sorted_vals, npts, V, C, feat, ldf,
                                                                    depth, D, HW, segments, out, ldo);
  return cudaGetLastError() == cudaSuccess ? 0 : COOCC_ERR_CUDA;
}

extern "C" int coocc_lss_pool_bwd(const unsigned int* point_keys, long long npts, int V, int C, const float* gout,
                                  long long ldg, const float* feat, long long ldf, const float* depth, int ncam,
                                  int D, int HW, float* dfeat, long long lddf, float* ddepth, void* stream) {
  if (!point_keys || !gout || !dfeat || npts < 1 || V < 1 || C < 4 || C % 4 || ldg % 4 || lddf % 4) return COOCC_ERR_ARG;
  cudaStream_t st = (cudaStream_t)stream;
  if (!depth) {
    lss_pool_bwd_plain_kernel<<<(unsigned)((npts + 7) / 8), 256, 0, st>>>(point_keys, npts, V, C, gout, ldg, dfeat, lddf);
  } else {
    if (!feat || !ddepth || C > 512 || ldf % 4 || (long long)ncam * D * HW != npts) return COOCC_ERR_ARG;
    const long long rows = (long long)ncam * HW;
    lss_pool_bwd_fused_kernel<<<(unsigned)((rows + 7) / 8), 256, 0, st>>>(point_keys, ncam, D, HW, V, C, feat, ldf,
                                                                         depth, gout, ldg, dfeat, lddf, ddepth);
  }
  return cudaGetLastError() == cudaSuccess ? 0 : COOCC_ERR_CUDA;
}
