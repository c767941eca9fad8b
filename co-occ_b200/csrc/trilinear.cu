// trilinear.cu -- F.interpolate(mode='trilinear', align_corners=False, size=...) on NDHWC rows,
// fused with the additions / per-voxel weights that follow it in the reference:
//   FPN3D top-down path   laterals[i-1] = laterals[i-1] + interpolate(laterals[i])   (P/coocc/necks/fpn3d.py:91-94)
//   OccHead level fusion  out += interpolate(feats) * softmax_weight[:, level]         (P/coocc/dense_heads/occ_head.py:161-165)
// out[v, :] = base[v, :] + wts[v] * sum_{8 corners} lambda * src[corner, :]
// Backward w.r.t. src is a *gather* over the output voxels that reference an input voxel (no atomics,
// deterministic); w.r.t. wts a per-voxel dot product.
#include <cuda_runtime.h>
#include <stdint.h>
#include <string.h>

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

// Separable form of the same transpose: interp = Tx (x) Ty (x) Tz, so interp^T is three 1-D passes,
// each reading its input once with fully coalesced rows and writing a tensor `scale` times smaller.
// The direct gather above pulls every dout row 8 times through L2 (2x2x2 footprint); the passes read
// dout once.  View of a pass: in [outer][n_in][inner][C] -> out [outer][n_out][inner][C],
//   out[a, i, r, :] = sum_{o : i in {i0(o), i1(o)}} lambda(o, i) * vw(a, o, r) * in[a, o, r, :]
// (vw = optional per-voxel weight of the first pass).  One warp per output row.
template <typename Tin, typename Tout>
__global__ void __launch_bounds__(256) trilinear_bwd_axis_kernel(const Tin* __restrict__ in, long long ldi,
                                                                 int outer, int n_in, int n_out, int inner, int C,
                                                                 float scale, const float* __restrict__ vw,
                                                                 long long ldw, Tout* __restrict__ out,
                                                                 long long ldo) {
  const int lane = threadIdx.x & 31;
  const int c4 = C >> 2;
  const long long rows = (long long)outer * n_out * inner;
  for (long long row = blockIdx.x * 8LL + (threadIdx.x >> 5); row < rows; row += gridDim.x * 8LL) {
    const int r = (int)(row % inner);
    const long long t = row / inner;
    const int i = (int)(t % n_out);
    const long long a = t / n_out;
    int lo, hi;
    tl_range(i, scale, n_in, lo, hi);
    for (int cg = lane; cg < c4; cg += 32) {
      float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
      for (int o = lo; o <= hi; ++o) {
        int a0, a1;
        float la;
        tl_src(o, scale, n_out, a0, a1, la);
        float w = (a0 == i ? 1.f - la : 0.f) + (a1 == i ? la : 0.f);
        if (w == 0.f) continue;
        const long long irow = (a * n_in + o) * inner + r;
        if (vw != nullptr) w *= vw[irow * ldw];
        const float4 g = load4(in + irow * ldi + cg * 4);
        acc.x += w * g.x; acc.y += w * g.y; acc.z += w * g.z; acc.w += w * g.w;
      }
      store4(out + row * ldo + cg * 4, acc);
    }
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

// ---------------------------------------------------------------------------------------------
// Multi-level mix: out[v,:] = base[v,:] (opt) + sum_{l < nlev} w_l(v) * interp_l(src_l)[v,:], with
// w_l(v) = wts[v*ldw + l] (or 1).  One kernel for the whole OccHead level fusion
// (P/coocc/dense_heads/occ_head.py:161-165) instead of one read-modify-write pass per level: the
// [V, C] output is written once and the small coarse levels stay L2-resident.
// One warp per output voxel, lanes over channel groups of 4; the index arithmetic (32-bit) is done
// once per voxel and level, zero-weight corners (e.g. every corner but one of a same-size level) are
// skipped warp-uniformly.
// ---------------------------------------------------------------------------------------------
constexpr int kTlLevels = 4;
constexpr int kTileXY = 4;
struct TlMix {
  const void* src[kTlLevels];
  void* dsrc[kTlLevels];        // backward only: gradient of same-size levels (else nullptr)
  long long lds[kTlLevels];
  int sX[kTlLevels], sY[kTlLevels], sZ[kTlLevels];
  float fx[kTlLevels], fy[kTlLevels], fz[kTlLevels];
  int nlev;
};

template <typename T, int G>
__global__ void __launch_bounds__(256) trilinear_mix_fwd_kernel(TlMix m, int oX, int oY, int oZ, int C,
                                                                const T* __restrict__ base, long long ldb,
                                                                const float* __restrict__ wts, long long ldw,
                                                                T* __restrict__ out, long long ldo) {
  const int lane = threadIdx.x & 31;
  const int c4 = C >> 2;
  // blocks walk compact (kTileXY x kTileXY x oZ) bricks of the output so that the corner rows of the
  // coarse levels, shared by neighbouring output voxels, are re-read from L1 instead of L2
  const int tilesY = (oY + kTileXY - 1) / kTileXY;
  const int ntile = ((oX + kTileXY - 1) / kTileXY) * tilesY;
  const int per_tile = kTileXY * kTileXY * oZ;
  for (int tile = blockIdx.x; tile < ntile; tile += gridDim.x)
  for (int idx = threadIdx.x >> 5; idx < per_tile; idx += 8) {
    const int z = idx % oZ, cxy = idx / oZ;
    const int x = (tile / tilesY) * kTileXY + cxy / kTileXY, y = (tile % tilesY) * kTileXY + cxy % kTileXY;
    if (x >= oX || y >= oY) continue;
    const int v = (x * oY + y) * oZ + z;
    float4 acc[G];
#pragma unroll
    for (int g = 0; g < G; ++g) acc[g] = make_float4(0.f, 0.f, 0.f, 0.f);
    if (base != nullptr) {
#pragma unroll
      for (int g = 0; g < G; ++g)
        if (lane + 32 * g < c4) acc[g] = load4(base + (long long)v * ldb + (lane + 32 * g) * 4);
    }
    for (int l = 0; l < m.nlev; ++l) {
      const T* src = reinterpret_cast<const T*>(m.src[l]);
      int x0, x1, y0, y1, z0, z1;
      float lx, ly, lz;
      tl_src(x, m.fx[l], m.sX[l], x0, x1, lx);
      tl_src(y, m.fy[l], m.sY[l], y0, y1, ly);
      tl_src(z, m.fz[l], m.sZ[l], z0, z1, lz);
      const float wl = wts ? wts[(long long)v * ldw + l] : 1.f;
#pragma unroll
      for (int k = 0; k < 8; ++k) {
        const float w = wl * ((k & 4) ? lx : 1.f - lx) * ((k & 2) ? ly : 1.f - ly) * ((k & 1) ? lz : 1.f - lz);
        if (w == 0.f) continue;
        const int xi = (k & 4) ? x1 : x0, yi = (k & 2) ? y1 : y0, zi = (k & 1) ? z1 : z0;
        const T* row = src + (long long)((xi * m.sY[l] + yi) * m.sZ[l] + zi) * m.lds[l];
#pragma unroll
        for (int g = 0; g < G; ++g) {
          if (lane + 32 * g < c4) {
            const float4 sv = load4(row + (lane + 32 * g) * 4);
            acc[g].x += w * sv.x; acc[g].y += w * sv.y; acc[g].z += w * sv.z; acc[g].w += w * sv.w;
          }
        }
      }
    }
#pragma unroll
    for (int g = 0; g < G; ++g)
      if (lane + 32 * g < c4) store4(out + (long long)v * ldo + (lane + 32 * g) * 4, acc[g]);
  }
}

// Backward companion of the mix, one pass over dout:
//   dw[v*lddw + l] = sum_c dout[v,c] * interp_l(src_l)[v,c]            (when dw != nullptr)
//   dsrc_l[v,:]    = w_l(v) * dout[v,:]   for the levels whose size equals the output size
// (the coarser levels' gradients are transposed gathers: trilinear_bwd_kernel per level).
template <typename T, int G>
__global__ void __launch_bounds__(256) trilinear_mix_bwd_kernel(TlMix m, int oX, int oY, int oZ, int C,
                                                                const T* __restrict__ dout, long long ldd,
                                                                const float* __restrict__ wts, long long ldw,
                                                                float* __restrict__ dw, long long lddw) {
  const int lane = threadIdx.x & 31;
  const int c4 = C >> 2;
  // blocks walk compact (kTileXY x kTileXY x oZ) bricks of the output so that the corner rows of the
  // coarse levels, shared by neighbouring output voxels, are re-read from L1 instead of L2
  const int tilesY = (oY + kTileXY - 1) / kTileXY;
  const int ntile = ((oX + kTileXY - 1) / kTileXY) * tilesY;
  const int per_tile = kTileXY * kTileXY * oZ;
  for (int tile = blockIdx.x; tile < ntile; tile += gridDim.x)
  for (int idx = threadIdx.x >> 5; idx < per_tile; idx += 8) {
    const int z = idx % oZ, cxy = idx / oZ;
    const int x = (tile / tilesY) * kTileXY + cxy / kTileXY, y = (tile % tilesY) * kTileXY + cxy % kTileXY;
    if (x >= oX || y >= oY) continue;
    const int v = (x * oY + y) * oZ + z;
    float4 g4[G];
#pragma unroll
    for (int g = 0; g < G; ++g)
      g4[g] = (lane + 32 * g < c4) ? load4(dout + (long long)v * ldd + (lane + 32 * g) * 4) : make_float4(0.f, 0.f, 0.f, 0.f);
    float dots[kTlLevels];
#pragma unroll
    for (int l = 0; l < kTlLevels; ++l) dots[l] = 0.f;
#pragma unroll
    for (int l = 0; l < kTlLevels; ++l) {
      if (l >= m.nlev) break;
      const float wl = wts ? wts[(long long)v * ldw + l] : 1.f;
      if (m.dsrc[l] != nullptr) {
        T* drow = reinterpret_cast<T*>(m.dsrc[l]) + (long long)v * m.lds[l];
#pragma unroll
        for (int g = 0; g < G; ++g)
          if (lane + 32 * g < c4)
            store4(drow + (lane + 32 * g) * 4, make_float4(wl * g4[g].x, wl * g4[g].y, wl * g4[g].z, wl * g4[g].w));
      }
      if (dw == nullptr) continue;
      const T* src = reinterpret_cast<const T*>(m.src[l]);
      int x0, x1, y0, y1, z0, z1;
      float lx, ly, lz;
      tl_src(x, m.fx[l], m.sX[l], x0, x1, lx);
      tl_src(y, m.fy[l], m.sY[l], y0, y1, ly);
      tl_src(z, m.fz[l], m.sZ[l], z0, z1, lz);
      float d = 0.f;
#pragma unroll
      for (int k = 0; k < 8; ++k) {
        const float w = ((k & 4) ? lx : 1.f - lx) * ((k & 2) ? ly : 1.f - ly) * ((k & 1) ? lz : 1.f - lz);
        if (w == 0.f) continue;
        const int xi = (k & 4) ? x1 : x0, yi = (k & 2) ? y1 : y0, zi = (k & 1) ? z1 : z0;
        const T* row = src + (long long)((xi * m.sY[l] + yi) * m.sZ[l] + zi) * m.lds[l];
#pragma unroll
        for (int g = 0; g < G; ++g) {
          if (lane + 32 * g < c4) {
            const float4 sv = load4(row + (lane + 32 * g) * 4);
            d += w * (sv.x * g4[g].x + sv.y * g4[g].y + sv.z * g4[g].z + sv.w * g4[g].w);
          }
        }
      }
      dots[l] = d;
    }
    if (dw != nullptr) {
#pragma unroll
      for (int l = 0; l < kTlLevels; ++l) {
#pragma unroll
        for (int o = 16; o >= 1; o >>= 1) dots[l] += __shfl_xor_sync(0xffffffffu, dots[l], o);
      }
      if (lane < m.nlev) {
        float val = dots[0];
#pragma unroll
        for (int l = 1; l < kTlLevels; ++l) val = (lane == l) ? dots[l] : val;
        dw[(long long)v * lddw + lane] = val;
      }
    }
  }
}


// ---------------------------------------------------------------------------------------------
// Column form of the mix (the one normally launched).  One warp owns an output column (x, y, all z): for a coarse
// level the four (x-corner, y-corner) rows of every source z it needs are combined once and shared by all output z of
// the column -- a 2x finer column of 16 voxels needs 4 x 5 rows per 8-voxel chunk instead of 8 rows per voxel (ncu r02:
// the per-voxel kernels above move 0.34-0.52 GB in 0.9 ms, bound by L1/L2 row gathers, not by DRAM).  Output z is
// processed in chunks of kZC = 8; a level qualifies if it has the output's size (direct rows) or is coarse enough
// that a chunk touches at most kMaxZS source planes (scale <= 1/2).  Weights w_l(v) and the base are applied per voxel.
// ---------------------------------------------------------------------------------------------
constexpr int kZC = 8;
constexpr int kMaxZS = 6;

struct ColZ {          // z interpolation of one output voxel w.r.t. one level
  int z0, z1;
  float l1;
};

template <typename T>
__global__ void __launch_bounds__(256) trilinear_mix_fwd_col_kernel(TlMix m, int oX, int oY, int oZ, int C,
                                                                    const T* __restrict__ base, long long ldb,
                                                                    const float* __restrict__ wts, long long ldw,
                                                                    T* __restrict__ out, long long ldo) {
  constexpr int G = 1;
  const int c4 = C >> 2;
  const int nslab = (c4 + 31) / 32;                      // 128-channel slabs: a warp owns (column, slab)
  const int ncol = oX * oY;
  for (int item = blockIdx.x * 8 + (threadIdx.x >> 5); item < ncol * nslab; item += gridDim.x * 8) {
    const int col = item / nslab;
    const int lane = (threadIdx.x & 31) + 32 * (item % nslab);     // channel group of this lane (4 channels)
    const int x = col / oY, y = col % oY;
    for (int zb = 0; zb < oZ; zb += kZC) {
      float4 acc[kZC][G];
#pragma unroll
      for (int zi = 0; zi < kZC; ++zi) {
        const int z = zb + zi;
#pragma unroll
        for (int g = 0; g < G; ++g) {
          acc[zi][g] = make_float4(0.f, 0.f, 0.f, 0.f);
          if (base != nullptr && z < oZ && lane + 32 * g < c4)
            acc[zi][g] = load4(base + ((long long)col * oZ + z) * ldb + (lane + 32 * g) * 4);
        }
      }
      for (int l = 0; l < m.nlev; ++l) {
        const T* src = reinterpret_cast<const T*>(m.src[l]);
        float wl[kZC];
#pragma unroll
        for (int zi = 0; zi < kZC; ++zi)
          wl[zi] = (wts != nullptr && zb + zi < oZ) ? wts[((long long)col * oZ + zb + zi) * ldw + l] : 1.f;
        if (m.sX[l] == oX && m.sY[l] == oY && m.sZ[l] == oZ) {           // same size: the voxel's own row
#pragma unroll
          for (int zi = 0; zi < kZC; ++zi) {
            if (zb + zi >= oZ) continue;
            const T* row = src + ((long long)col * oZ + zb + zi) * m.lds[l];
#pragma unroll
            for (int g = 0; g < G; ++g)
              if (lane + 32 * g < c4) {
                const float4 sv = load4(row + (lane + 32 * g) * 4);
                acc[zi][g].x += wl[zi] * sv.x; acc[zi][g].y += wl[zi] * sv.y;
                acc[zi][g].z += wl[zi] * sv.z; acc[zi][g].w += wl[zi] * sv.w;
              }
          }
          continue;
        }
        int x0, x1, y0, y1;
        float lx, ly;
        tl_src(x, m.fx[l], m.sX[l], x0, x1, lx);
        tl_src(y, m.fy[l], m.sY[l], y0, y1, ly);
        ColZ cz[kZC];
        int zlo = 1 << 30, zhi = -1;
#pragma unroll
        for (int zi = 0; zi < kZC; ++zi) {
          tl_src(min(zb + zi, oZ - 1), m.fz[l], m.sZ[l], cz[zi].z0, cz[zi].z1, cz[zi].l1);
          zlo = min(zlo, cz[zi].z0);
          zhi = max(zhi, cz[zi].z1);
        }
        const long long r00 = ((long long)x0 * m.sY[l] + y0) * m.sZ[l], r01 = ((long long)x0 * m.sY[l] + y1) * m.sZ[l];
        const long long r10 = ((long long)x1 * m.sY[l] + y0) * m.sZ[l], r11 = ((long long)x1 * m.sY[l] + y1) * m.sZ[l];
        const float w00 = (1.f - lx) * (1.f - ly), w01 = (1.f - lx) * ly, w10 = lx * (1.f - ly), w11 = lx * ly;
#pragma unroll
        for (int s = 0; s < kMaxZS; ++s) {
          const int zs = zlo + s;
          if (zs > zhi) break;
#pragma unroll
          for (int g = 0; g < G; ++g) {
            if (lane + 32 * g >= c4) continue;
            const int co = (lane + 32 * g) * 4;
            const float4 a = load4(src + (r00 + zs) * m.lds[l] + co), b = load4(src + (r01 + zs) * m.lds[l] + co);
            const float4 c = load4(src + (r10 + zs) * m.lds[l] + co), d = load4(src + (r11 + zs) * m.lds[l] + co);
            float4 r;
            r.x = w00 * a.x + w01 * b.x + w10 * c.x + w11 * d.x;
            r.y = w00 * a.y + w01 * b.y + w10 * c.y + w11 * d.y;
            r.z = w00 * a.z + w01 * b.z + w10 * c.z + w11 * d.z;
            r.w = w00 * a.w + w01 * b.w + w10 * c.w + w11 * d.w;
#pragma unroll
            for (int zi = 0; zi < kZC; ++zi) {
              const float cf = wl[zi] * ((cz[zi].z0 == zs ? 1.f - cz[zi].l1 : 0.f) + (cz[zi].z1 == zs ? cz[zi].l1 : 0.f));
              acc[zi][g].x += cf * r.x; acc[zi][g].y += cf * r.y; acc[zi][g].z += cf * r.z; acc[zi][g].w += cf * r.w;
            }
          }
        }
      }
#pragma unroll
      for (int zi = 0; zi < kZC; ++zi) {
        if (zb + zi >= oZ) continue;
#pragma unroll
        for (int g = 0; g < G; ++g)
          if (lane + 32 * g < c4) store4(out + ((long long)col * oZ + zb + zi) * ldo + (lane + 32 * g) * 4, acc[zi][g]);
      }
    }
  }
}


// Backward companion in column form (same outputs as trilinear_mix_bwd_kernel): per (column, 8-voxel chunk) the
// 8 x 4 dot products <dout[v,:], interp_l(src_l)[v,:]> are accumulated per lane and reduced across the warp with one
// 31-shuffle transpose reduction (lane i ends up with dot i) instead of 5 shuffles per dot.  dw needs all channels of a
// voxel in one warp, so this form is used for C <= 128.
template <typename T>
__global__ void __launch_bounds__(256) trilinear_mix_bwd_col_kernel(TlMix m, int oX, int oY, int oZ, int C,
                                                                    const T* __restrict__ dout, long long ldd,
                                                                    const float* __restrict__ wts, long long ldw,
                                                                    float* __restrict__ dw, long long lddw) {
  const int lane = threadIdx.x & 31;
  const int c4 = C >> 2;
  const bool act = lane < c4;
  const int ncol = oX * oY;
  for (int col = blockIdx.x * 8 + (threadIdx.x >> 5); col < ncol; col += gridDim.x * 8) {
    const int x = col / oY, y = col % oY;
    for (int zb = 0; zb < oZ; zb += kZC) {
      float4 g4[kZC];
#pragma unroll
      for (int zi = 0; zi < kZC; ++zi)
        g4[zi] = (act && zb + zi < oZ) ? load4(dout + ((long long)col * oZ + zb + zi) * ldd + lane * 4)
                                       : make_float4(0.f, 0.f, 0.f, 0.f);
      float dots[kZC * kTlLevels];
#pragma unroll
      for (int i = 0; i < kZC * kTlLevels; ++i) dots[i] = 0.f;
#pragma unroll
      for (int l = 0; l < kTlLevels; ++l) {
        if (l >= m.nlev) continue;
        const T* src = reinterpret_cast<const T*>(m.src[l]);
        const bool same = m.sX[l] == oX && m.sY[l] == oY && m.sZ[l] == oZ;
        if (m.dsrc[l] != nullptr) {                 // same-size level: d src = w_l(v) * dout
#pragma unroll
          for (int zi = 0; zi < kZC; ++zi) {
            if (!act || zb + zi >= oZ) continue;
            const long long v = (long long)col * oZ + zb + zi;
            const float wl = wts ? wts[v * ldw + l] : 1.f;
            store4(reinterpret_cast<T*>(m.dsrc[l]) + v * m.lds[l] + lane * 4,
                   make_float4(wl * g4[zi].x, wl * g4[zi].y, wl * g4[zi].z, wl * g4[zi].w));
          }
        }
        if (dw == nullptr) continue;
        if (same) {
#pragma unroll
          for (int zi = 0; zi < kZC; ++zi) {
            if (!act || zb + zi >= oZ) continue;
            const float4 sv = load4(src + ((long long)col * oZ + zb + zi) * m.lds[l] + lane * 4);
            dots[zi * kTlLevels + l] = sv.x * g4[zi].x + sv.y * g4[zi].y + sv.z * g4[zi].z + sv.w * g4[zi].w;
          }
          continue;
        }
        int x0, x1, y0, y1;
        float lx, ly;
        tl_src(x, m.fx[l], m.sX[l], x0, x1, lx);
        tl_src(y, m.fy[l], m.sY[l], y0, y1, ly);
        ColZ cz[kZC];
        int zlo = 1 << 30, zhi = -1;
#pragma unroll
        for (int zi = 0; zi < kZC; ++zi) {
          tl_src(min(zb + zi, oZ - 1), m.fz[l], m.sZ[l], cz[zi].z0, cz[zi].z1, cz[zi].l1);
          zlo = min(zlo, cz[zi].z0);
          zhi = max(zhi, cz[zi].z1);
        }
        const long long r00 = ((long long)x0 * m.sY[l] + y0) * m.sZ[l], r01 = ((long long)x0 * m.sY[l] + y1) * m.sZ[l];
        const long long r10 = ((long long)x1 * m.sY[l] + y0) * m.sZ[l], r11 = ((long long)x1 * m.sY[l] + y1) * m.sZ[l];
        const float w00 = (1.f - lx) * (1.f - ly), w01 = (1.f - lx) * ly, w10 = lx * (1.f - ly), w11 = lx * ly;
#pragma unroll
        for (int s2 = 0; s2 < kMaxZS; ++s2) {
          const int zs = zlo + s2;
          if (zs > zhi || !act) break;
          const int co = lane * 4;
          const float4 a = load4(src + (r00 + zs) * m.lds[l] + co), b = load4(src + (r01 + zs) * m.lds[l] + co);
          const float4 c = load4(src + (r10 + zs) * m.lds[l] + co), d = load4(src + (r11 + zs) * m.lds[l] + co);
          float4 r;
          r.x = w00 * a.x + w01 * b.x + w10 * c.x + w11 * d.x;
          r.y = w00 * a.y + w01 * b.y + w10 * c.y + w11 * d.y;
          r.z = w00 * a.z + w01 * b.z + w10 * c.z + w11 * d.z;
          r.w = w00 * a.w + w01 * b.w + w10 * c.w + w11 * d.w;
#pragma unroll
          for (int zi = 0; zi < kZC; ++zi) {
            const float cf = (cz[zi].z0 == zs ? 1.f - cz[zi].l1 : 0.f) + (cz[zi].z1 == zs ? cz[zi].l1 : 0.f);
            dots[zi * kTlLevels + l] += cf * (r.x * g4[zi].x + r.y * g4[zi].y + r.z * g4[zi].z + r.w * g4[zi].w);
          }
        }
      }
      if (dw != nullptr) {
        // transpose reduction: after the five steps lane i holds the warp total of dots[i]
#pragma unroll
        for (int off = 16; off >= 1; off >>= 1) {
          const bool up = (lane & off) != 0;
#pragma unroll
          for (int i = 0; i < off; ++i) {
            const float send = up ? dots[i] : dots[i + off];
            const float keep = up ? dots[i + off] : dots[i];
            dots[i] = keep + __shfl_xor_sync(0xffffffffu, send, off);
          }
        }
        const int zi = lane / kTlLevels, l = lane % kTlLevels;
        if (l < m.nlev && zb + zi < oZ) dw[((long long)col * oZ + zb + zi) * lddw + l] = dots[0];
      }
    }
  }
}

// can the column kernels serve this mix?  every level same-size or at most half the output's z resolution
static bool mix_col_ok(const TlMix& m, int oX, int oY, int oZ) {
  for (int l = 0; l < m.nlev; ++l) {
    const bool same = m.sX[l] == oX && m.sY[l] == oY && m.sZ[l] == oZ;
    if (!same && 2 * m.sZ[l] > oZ) return false;
  }
  return true;
}
static int mix_col_grid(int oX, int oY) {
  long long b = ((long long)oX * oY + 7) / 8;
  if (b > 148LL * 8) b = 148LL * 8;
  return b < 1 ? 1 : (int)b;
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
typedef __nv_bfloat16 bf16_t;
#define CK_LAUNCH() (cudaGetLastError() == cudaSuccess ? 0 : COOCC_ERR_CUDA)

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

static int tl_axis_grid(long long rows) {
  long long b = (rows + 7) / 8;
  if (b > 148LL * 32) b = 148LL * 32;
  return b < 1 ? 1 : (int)b;
}

// three 1-D transposed passes (x, then y, then z) with fp32 intermediates in stream-ordered scratch
template <typename T>
static int trilinear_bwd_separable(const T* dout, long long ldd, int oX, int oY, int oZ, int C, const float* wts,
                                   long long ldw, T* dsrc, long long lds, int sX, int sY, int sZ, cudaStream_t st) {
  const long long n1 = (long long)sX * oY * oZ, n2 = (long long)sX * sY * oZ;
  float* tmp = nullptr;
  if (cudaMallocAsync(reinterpret_cast<void**>(&tmp), (size_t)(n1 + n2) * C * sizeof(float), st) != cudaSuccess)
    return COOCC_ERR_CUDA;
  float* t1 = tmp;
  float* t2 = tmp + n1 * C;
  trilinear_bwd_axis_kernel<T, float><<<tl_axis_grid(n1), 256, 0, st>>>(dout, ldd, 1, oX, sX, oY * oZ, C,
                                                                       (float)sX / (float)oX, wts, ldw, t1, C);
  trilinear_bwd_axis_kernel<float, float><<<tl_axis_grid(n2), 256, 0, st>>>(t1, C, sX, oY, sY, oZ, C,
                                                                           (float)sY / (float)oY, nullptr, 0, t2, C);
  trilinear_bwd_axis_kernel<float, T><<<tl_axis_grid((long long)sX * sY * sZ), 256, 0, st>>>(
      t2, C, sX * sY, oZ, sZ, 1, C, (float)sZ / (float)oZ, nullptr, 0, dsrc, lds);
  const int rc = cudaGetLastError() == cudaSuccess ? 0 : COOCC_ERR_CUDA;
  cudaFreeAsync(tmp, st);
  return rc;
}

static int g_tl_separable = 1, g_tl_column = 1;
/* benchmark hook: bit 0: 0 = direct transposed gather, 1 = separable passes (default); bit 1 set = per-voxel mix
 * kernels instead of the column form */
extern "C" int coocc_trilinear_tune(int flags) {
  g_tl_separable = (flags & 1) ? 1 : 0;
  g_tl_column = (flags & 2) ? 0 : 1;
  return 0;
}

extern "C" int coocc_trilinear_bwd(const void* dout, long long ldd, int oX, int oY, int oZ, int C,
                                   const float* wts, long long ldw, void* dsrc, long long lds, int sX, int sY,
                                   int sZ, int act_bf16, void* stream) {
  if (!dout || !dsrc || (C & 3) || (lds & 3) || (ldd & 3)) return COOCC_ERR_ARG;
  // separable passes win from x4 upwards (each source voxel would gather 10^3+ output rows directly); for x2 the direct
  // gather is faster (B200, 200x200x16 <- 100x100x8: C=128 0.43 vs 0.49 ms, C=256 0.82 vs 1.15 ms; x4: 0.38 vs 0.29;
  // x8: 1.10 vs 0.56; tools/tl_axis_bench.py)
  if (g_tl_separable && (long long)oX * oY * oZ >= 4096 && (long long)oX * oY * oZ >= 27LL * sX * sY * sZ) {
    return act_bf16 ? trilinear_bwd_separable<bf16_t>((const bf16_t*)dout, ldd, oX, oY, oZ, C, wts, ldw, (bf16_t*)dsrc,
                                                      lds, sX, sY, sZ, (cudaStream_t)stream)
                    : trilinear_bwd_separable<float>((const float*)dout, ldd, oX, oY, oZ, C, wts, ldw, (float*)dsrc,
                                                     lds, sX, sY, sZ, (cudaStream_t)stream);
  }
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

// ---- multi-level mix (see trilinear_mix_fwd_kernel) ------------------------------------------
static int fill_mix(TlMix& m, int nlev, const void* const* src, const long long* lds, const int* sdims, int oX, int oY,
                    int oZ) {
  if (nlev < 1 || nlev > kTlLevels || !src || !lds || !sdims) return COOCC_ERR_ARG;
  memset(&m, 0, sizeof(m));
  m.nlev = nlev;
  for (int l = 0; l < nlev; ++l) {
    if (!src[l] || (lds[l] & 3)) return COOCC_ERR_ARG;
    m.src[l] = src[l];
    m.lds[l] = lds[l];
    m.sX[l] = sdims[3 * l]; m.sY[l] = sdims[3 * l + 1]; m.sZ[l] = sdims[3 * l + 2];
    m.fx[l] = (float)m.sX[l] / (float)oX;
    m.fy[l] = (float)m.sY[l] / (float)oY;
    m.fz[l] = (float)m.sZ[l] / (float)oZ;
  }
  return 0;
}
static int mix_grid(int oX, int oY) {
  long long b = (long long)((oX + kTileXY - 1) / kTileXY) * ((oY + kTileXY - 1) / kTileXY);
  if (b > 148LL * 8) b = 148LL * 8;
  return b < 1 ? 1 : (int)b;
}

#define TL_DISPATCH_G(KERNEL, T, ...)                                                          \
  do {                                                                                         \
    const int groups_ = ((C >> 2) + 31) / 32;                                                  \
    if (groups_ <= 1) KERNEL<T, 1><<<g, 256, 0, (cudaStream_t)stream>>>(__VA_ARGS__);          \
    else if (groups_ == 2) KERNEL<T, 2><<<g, 256, 0, (cudaStream_t)stream>>>(__VA_ARGS__);     \
    else if (groups_ <= 4) KERNEL<T, 4><<<g, 256, 0, (cudaStream_t)stream>>>(__VA_ARGS__);     \
    else return COOCC_ERR_CAPACITY;                                                            \
  } while (0)

extern "C" int coocc_trilinear_mix_fwd(int nlev, const void* const* src, const long long* lds, const int* sdims,
                                       int C, const void* base, long long ldb, const float* wts, long long ldw,
                                       void* out, long long ldo, int oX, int oY, int oZ, int act_bf16,
                                       void* stream) {
  if (!out || (C & 3) || (ldo & 3) || (base && (ldb & 3))) return COOCC_ERR_ARG;
  TlMix m;
  int rc = fill_mix(m, nlev, src, lds, sdims, oX, oY, oZ);
  if (rc) return rc;
  if (g_tl_column && m.nlev >= 2 && mix_col_ok(m, oX, oY, oZ)) {      // (a single coarse level gains nothing)
    const int g = mix_col_grid(oX, oY * (((C >> 2) + 31) / 32));
    if (act_bf16)
      trilinear_mix_fwd_col_kernel<bf16_t><<<g, 256, 0, (cudaStream_t)stream>>>(m, oX, oY, oZ, C, (const bf16_t*)base, ldb,
                                                                              wts, ldw, (bf16_t*)out, ldo);
    else
      trilinear_mix_fwd_col_kernel<float><<<g, 256, 0, (cudaStream_t)stream>>>(m, oX, oY, oZ, C, (const float*)base, ldb,
                                                                             wts, ldw, (float*)out, ldo);
    return CK_LAUNCH();
  }
  const int g = mix_grid(oX, oY);
  if (act_bf16)
    TL_DISPATCH_G(trilinear_mix_fwd_kernel, bf16_t, m, oX, oY, oZ, C, (const bf16_t*)base, ldb, wts, ldw, (bf16_t*)out,
                  ldo);
  else
    TL_DISPATCH_G(trilinear_mix_fwd_kernel, float, m, oX, oY, oZ, C, (const float*)base, ldb, wts, ldw, (float*)out,
                  ldo);
  return CK_LAUNCH();
}

extern "C" int coocc_trilinear_mix_bwd(int nlev, const void* const* src, const long long* lds, const int* sdims,
                                       void* const* dsrc_same, int C, const void* dout, long long ldd,
                                       const float* wts, long long ldw, float* dw, long long lddw, int oX, int oY,
                                       int oZ, int act_bf16, void* stream) {
  if (!dout || (C & 3) || (ldd & 3)) return COOCC_ERR_ARG;
  TlMix m;
  int rc = fill_mix(m, nlev, src, lds, sdims, oX, oY, oZ);
  if (rc) return rc;
  if (dsrc_same)
    for (int l = 0; l < nlev; ++l) {
      if (dsrc_same[l] && (m.sX[l] != oX || m.sY[l] != oY || m.sZ[l] != oZ)) return COOCC_ERR_ARG;
      m.dsrc[l] = dsrc_same[l];
    }
  static_assert(kZC * kTlLevels == 32, "the transpose reduction maps one dot product to every lane");
  if (g_tl_column && m.nlev >= 2 && mix_col_ok(m, oX, oY, oZ) && (C >> 2) <= 32) {
    const int g = mix_col_grid(oX, oY);
    if (act_bf16)
      trilinear_mix_bwd_col_kernel<bf16_t><<<g, 256, 0, (cudaStream_t)stream>>>(m, oX, oY, oZ, C, (const bf16_t*)dout, ldd,
                                                                              wts, ldw, dw, lddw);
    else
      trilinear_mix_bwd_col_kernel<float><<<g, 256, 0, (cudaStream_t)stream>>>(m, oX, oY, oZ, C, (const float*)dout, ldd,
                                                                             wts, ldw, dw, lddw);
    return CK_LAUNCH();
  }
  const int g = mix_grid(oX, oY);
  if (act_bf16)
    TL_DISPATCH_G(trilinear_mix_bwd_kernel, bf16_t, m, oX, oY, oZ, C, (const bf16_t*)dout, ldd, wts, ldw, dw, lddw);
  else
    TL_DISPATCH_G(trilinear_mix_bwd_kernel, float, m, oX, oY, oZ, C, (const float*)dout, ldd, wts, ldw, dw, lddw);
  return CK_LAUNCH();
}
