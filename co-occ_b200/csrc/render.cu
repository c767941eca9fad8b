// render.cu -- NeRF-style volume-render regulariser (reference: the inline block of
// P/coocc/detectors/coocc_ray.py:358-433).
//
// Restructuring: rgb/sigma depend only on the voxel a sample falls in, so the two
// MLP heads are evaluated once per voxel of the render box (tensor-core GEMMs, conv_tc.cu) into a
// table tab[T][4] = (rgb_raw[3], relu(sigma)); the kernels here do
//   box_gather / box_scatter : feature rows of the render box  <->  full grid
//   composite fwd / bwd      : per-ray geometry -> voxel index, alpha compositing, gradients
//   upsample16 + MSE fwd/bwd : x16 bilinear (align_corners=False) and the two render losses
// The render box is the reference's hard-coded 100x100x8 @ 1 m grid clipped to the feature grid
// (coocc_ray.py:372).
#include <cuda_runtime.h>
#include <math.h>
#include <stdint.h>

#include "../../include/coocc_b200.h"

namespace coocc {

constexpr int kBoxX = 100, kBoxY = 100, kBoxZ = 8;
constexpr int kMaxD = 128;

// rows of the render box (t = (x*by + y)*bz + z) gathered from / scattered to grid rows
__global__ void box_gather_kernel(const float* __restrict__ grid, long long ld, int Y, int Z, int bx,
                                  int by, int bz, int C, float* __restrict__ rows) {
  const long long t = blockIdx.x * (long long)(blockDim.x >> 5) + (threadIdx.x >> 5);
  if (t >= (long long)bx * by * bz) return;
  const int z = t % bz, y = (t / bz) % by, x = t / (bz * by);
  const float* s = grid + (((long long)x * Y + y) * Z + z) * ld;
  float* d = rows + t * C;
  for (int c = threadIdx.x & 31; c < C; c += 32) d[c] = s[c];
}
__global__ void box_scatter_add_kernel(const float* __restrict__ rows, int C, int Y, int Z, int bx, int by,
                                       int bz, float* __restrict__ grid, long long ld) {
  const long long t = blockIdx.x * (long long)(blockDim.x >> 5) + (threadIdx.x >> 5);
  if (t >= (long long)bx * by * bz) return;
  const int z = t % bz, y = (t / bz) % by, x = t / (bz * by);
  float* d = grid + (((long long)x * Y + y) * Z + z) * ld;
  const float* s = rows + t * C;
  for (int c = threadIdx.x & 31; c < C; c += 32) d[c] += s[c];
}

// bf16 activations: the box rows are gathered as they are stored (8 channels per 16-byte access) ...
__global__ void box_gather_bf16_kernel(const uint4* __restrict__ grid, long long ld8, int Y, int Z, int bx, int by,
                                       int bz, int C8, uint4* __restrict__ rows) {
  const long long t = blockIdx.x * (long long)(blockDim.x >> 5) + (threadIdx.x >> 5);
  if (t >= (long long)bx * by * bz) return;
  const int z = t % bz, y = (t / bz) % by, x = t / (bz * by);
  const uint4* s = grid + (((long long)x * Y + y) * Z + z) * ld8;
  uint4* d = rows + t * C8;
  for (int c = threadIdx.x & 31; c < C8; c += 32) d[c] = s[c];
}
// ... and the gradient of the whole grid is written in one pass: the box rows' gradients inside the box, zeros elsewhere
// (no separate zero fill, no fp32 round trip)
__global__ void box_scatter_bf16_kernel(const uint4* __restrict__ rows, int C8, int X, int Y, int Z, int bx, int by,
                                        int bz, uint4* __restrict__ grid, long long ld8) {
  const long long total = (long long)X * Y * Z * C8;
  for (long long i = blockIdx.x * 256LL + threadIdx.x; i < total; i += (long long)gridDim.x * 256LL) {
    const long long v = i / C8;
    const int c = (int)(i - v * C8);
    const int z = v % Z, y = (v / Z) % Y, x = v / ((long long)Z * Y);
    const bool in = x < bx && y < by && z < bz;
    grid[v * ld8 + c] = in ? rows[(((long long)x * by + y) * bz + z) * C8 + c] : make_uint4(0u, 0u, 0u, 0u);
  }
}

// torch.linspace(0, D, D)[i] (ATen's symmetric evaluation)
__device__ __forceinline__ float zval(int i, int D) {
  const float step = (float)D / (float)(D - 1);
  return i < D / 2 ? step * (float)i : (float)D - step * (float)(D - 1 - i);
}

// sample -> packed voxel coords (x | y<<8 | z<<16 | inside<<31), coocc_ray.py:377-384
__device__ __forceinline__ uint32_t sample_voxel(const float* __restrict__ g) {
  const float gx = (g[0] - (-50.0f)) / 1.0f;
  const float gy = (g[1] - (-50.0f)) / 1.0f;
  const float gz = (g[2] - (-5.0f)) / 1.0f;
  const bool in = gx >= 0.f && gx < (float)kBoxX && gy >= 0.f && gy < (float)kBoxY && gz >= 0.f &&
                  gz < (float)kBoxZ;
  if (!in) return 0u;
  return (uint32_t)(int)gx | ((uint32_t)(int)gy << 8) | ((uint32_t)(int)gz << 16) | 0x80000000u;
}

struct RayCtx {
  // per-lane segment state for one ray; a lane owns samples [lane*S, lane*S+S)
  float alpha[4], trans_local[4], rgb[4][3], sig[4], dist[4];
  int tix[4];
  bool in[4];
};

// Shared evaluation of one ray's samples (used by forward and backward).
// vox: smem [D][32] packed voxels of the block's rays; r = ray slot in the block.
// Returns per-lane: alpha, local transmittance prefix (product of t_j before the sample within
// the segment), and via Tpre the product of all t_j of earlier lanes.
__device__ __forceinline__ void eval_ray(const uint32_t* __restrict__ vox, int r, int D, int S,
                                         const float4* __restrict__ tab, int bx, int by, int bz,
                                         RayCtx& c, float& Tpre, int* __restrict__ err) {
  const int lane = threadIdx.x & 31;
  float prod = 1.f;
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int d = lane * S + i;
    c.alpha[i] = 0.f; c.trans_local[i] = prod; c.in[i] = false; c.tix[i] = -1; c.sig[i] = 0.f;
    c.dist[i] = 0.f; c.rgb[i][0] = c.rgb[i][1] = c.rgb[i][2] = 0.f;
    if (i < S && d < D) {
      const uint32_t p = vox[d * 32 + r];
      const int x = p & 255u, y = (p >> 8) & 255u, z = (p >> 16) & 255u;
      c.in[i] = (p >> 31) != 0u;
      // the reference indexes voxel_feats[:, x, y, z] directly: outside samples read voxel (0,0,0)
      if (x >= bx || y >= by || z >= bz) {
        if (err) atomicExch(err, 1);   // the reference raises IndexError here
        c.tix[i] = 0;
      } else {
        c.tix[i] = (x * by + y) * bz + z;
      }
      const float4 t = tab[c.tix[i]];
      c.sig[i] = t.w;
      const float m = c.in[i] ? 1.f : 0.f;
      c.rgb[i][0] = 1.f / (1.f + expf(-(t.x * m)));
      c.rgb[i][1] = 1.f / (1.f + expf(-(t.y * m)));
      c.rgb[i][2] = 1.f / (1.f + expf(-(t.z * m)));
      float dist = 1e10f;
      if (d + 1 < D) {
        const uint32_t pn = vox[(d + 1) * 32 + r];
        const float ddx = (float)((int)(pn & 255u) - x), ddy = (float)((int)((pn >> 8) & 255u) - y),
                    ddz = (float)((int)((pn >> 16) & 255u) - z);
        dist = sqrtf(ddx * ddx + ddy * ddy + ddz * ddz);
      }
      c.dist[i] = dist;
      const float a = 1.f - expf(-fmaxf(c.sig[i] * dist, 0.f));
      c.alpha[i] = a;
      prod *= (1.f - a + 1e-10f);
    }
  }
  // exclusive multiplicative scan of the segment products across lanes
  float incl = prod;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    const float t = __shfl_up_sync(0xffffffffu, incl, o);
    if (lane >= o) incl *= t;
  }
  Tpre = __shfl_up_sync(0xffffffffu, incl, 1);
  if (lane == 0) Tpre = 1.f;
}

// geom [Ncam][D][HW][3]; one block = 32 consecutive rays of one camera, 256 threads
__global__ void __launch_bounds__(256) composite_fwd_kernel(const float* __restrict__ geom, int D, int HW,
                                                            const float4* __restrict__ tab, int bx, int by,
                                                            int bz, float* __restrict__ rgb_map,
                                                            float* __restrict__ depth_map,
                                                            int* __restrict__ err) {
  __shared__ uint32_t vox[kMaxD * 32];
  const int cam = blockIdx.y;
  const int ray0 = blockIdx.x * 32;
  const float* gcam = geom + (long long)cam * D * HW * 3;
  for (int e = threadIdx.x; e < D * 32; e += 256) {
    const int d = e >> 5, r = e & 31;
    const int ray = ray0 + r;
    vox[e] = ray < HW ? sample_voxel(gcam + ((long long)d * HW + ray) * 3) : 0u;
  }
  __syncthreads();
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int S = (D + 31) / 32;
  for (int r = warp; r < 32; r += 8) {
    const int ray = ray0 + r;
    if (ray >= HW) break;
    RayCtx c;
    float Tpre;
    eval_ray(vox, r, D, S, tab, bx, by, bz, c, Tpre, err);
    float acc[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const int d = lane * S + i;
      if (i < S && d < D) {
        const float w = c.alpha[i] * (Tpre * c.trans_local[i]);
        acc[0] += w * c.rgb[i][0];
        acc[1] += w * c.rgb[i][1];
        acc[2] += w * c.rgb[i][2];
        acc[3] += w * zval(d, D);
      }
    }
#pragma unroll
    for (int o = 16; o >= 1; o >>= 1) {
#pragma unroll
      for (int j = 0; j < 4; ++j) acc[j] += __shfl_xor_sync(0xffffffffu, acc[j], o);
    }
    if (lane == 0) {
      const long long o = (long long)cam * HW + ray;
      rgb_map[o * 3 + 0] = acc[0];
      rgb_map[o * 3 + 1] = acc[1];
      rgb_map[o * 3 + 2] = acc[2];
      depth_map[o] = acc[3];
    }
  }
}

// d_tab[t] += (d rgb_raw[3], d sigma) ; sigma is post-ReLU in tab, its ReLU mask is sigma > 0
__global__ void __launch_bounds__(256) composite_bwd_kernel(const float* __restrict__ geom, int D, int HW,
                                                            const float4* __restrict__ tab, int bx, int by,
                                                            int bz, const float* __restrict__ g_rgb,
                                                            const float* __restrict__ g_depth,
                                                            float* __restrict__ d_tab) {
  __shared__ uint32_t vox[kMaxD * 32];
  const int cam = blockIdx.y;
  const int ray0 = blockIdx.x * 32;
  const float* gcam = geom + (long long)cam * D * HW * 3;
  for (int e = threadIdx.x; e < D * 32; e += 256) {
    const int d = e >> 5, r = e & 31;
    const int ray = ray0 + r;
    vox[e] = ray < HW ? sample_voxel(gcam + ((long long)d * HW + ray) * 3) : 0u;
  }
  __syncthreads();
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int S = (D + 31) / 32;
  for (int r = warp; r < 32; r += 8) {
    const int ray = ray0 + r;
    if (ray >= HW) break;
    RayCtx c;
    float Tpre;
    eval_ray(vox, r, D, S, tab, bx, by, bz, c, Tpre, nullptr);
    const long long o = (long long)cam * HW + ray;
    const float gr = g_rgb[o * 3 + 0], gg = g_rgb[o * 3 + 1], gb = g_rgb[o * 3 + 2], gd = g_depth[o];
    // G_d = dL/dw_d ; suffix sums of G_m w_m over later samples
    float G[4], w[4], T[4];
    float local = 0.f;
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const int d = lane * S + i;
      G[i] = 0.f; w[i] = 0.f; T[i] = 0.f;
      if (i < S && d < D) {
        T[i] = Tpre * c.trans_local[i];
        w[i] = c.alpha[i] * T[i];
        G[i] = gr * c.rgb[i][0] + gg * c.rgb[i][1] + gb * c.rgb[i][2] + gd * zval(d, D);
        local += G[i] * w[i];
      }
    }
    // exclusive suffix sum across lanes (sum over lanes > lane)
    float incl = local;
#pragma unroll
    for (int off = 1; off < 32; off <<= 1) {
      const float t = __shfl_down_sync(0xffffffffu, incl, off);
      if (lane + off < 32) incl += t;
    }
    float suffix = incl - local;     // contributions of later lanes
#pragma unroll
    for (int i = 3; i >= 0; --i) {
      const int d = lane * S + i;
      if (i < S && d < D) {
        // here `suffix` = sum_{m > d} G_m w_m
        const float tj = 1.f - c.alpha[i] + 1e-10f;
        const float dalpha = G[i] * T[i] - suffix / tj;
        const float sd = c.sig[i] * c.dist[i];
        // alpha = 1 - exp(-relu(sigma*dist)), sigma = relu(raw)
        float dsig = 0.f;
        if (sd > 0.f && c.sig[i] > 0.f) dsig = dalpha * c.dist[i] * expf(-sd);
        float4 gt = make_float4(0.f, 0.f, 0.f, dsig);
        if (c.in[i] && w[i] != 0.f) {
          gt.x = gr * w[i] * c.rgb[i][0] * (1.f - c.rgb[i][0]);
          gt.y = gg * w[i] * c.rgb[i][1] * (1.f - c.rgb[i][1]);
          gt.z = gb * w[i] * c.rgb[i][2] * (1.f - c.rgb[i][2]);
        }
        if (gt.x != 0.f || gt.y != 0.f || gt.z != 0.f || gt.w != 0.f) {
          float* p = d_tab + (long long)c.tix[i] * 4;
          asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(p), "f"(gt.x), "f"(gt.y),
                       "f"(gt.z), "f"(gt.w)
                       : "memory");
        }
        suffix += G[i] * w[i];
      }
    }
  }
}

// ------------------------------------------------------------------------------------------
// x16 bilinear upsample (align_corners=False) fused with the two MSE losses
// ------------------------------------------------------------------------------------------
__device__ __forceinline__ void src_index(int dst, int in_size, int& i0, int& i1, float& l1) {
  float s = (1.0f / 16.0f) * ((float)dst + 0.5f) - 0.5f;
  if (s < 0.f) s = 0.f;
  i0 = (int)s;
  i1 = i0 + (i0 < in_size - 1 ? 1 : 0);
  l1 = s - (float)i0;
}

// acc[0] += sum (rgb - gt)^2, acc[1] += sum_fg (depth/D - gt'/D)^2, acc[2] += |fg|
__global__ void __launch_bounds__(256) upsample_loss_fwd_kernel(
    const float* __restrict__ rgb_map, const float* __restrict__ depth_map, int ncam, int H, int W, int D,
    const float* __restrict__ gt_img, const float* __restrict__ gt_depth, float* __restrict__ rgbs,
    float* __restrict__ depths, float* __restrict__ acc) {
  const int HH = 16 * H, WW = 16 * W;
  const long long total = (long long)ncam * HH * WW;
  float e_rgb = 0.f, e_d = 0.f, n_fg = 0.f;
  for (long long p = blockIdx.x * 256LL + threadIdx.x; p < total; p += gridDim.x * 256LL) {
    const int x = p % WW, y = (p / WW) % HH, n = p / ((long long)WW * HH);
    int y0, y1, x0, x1;
    float ly, lx;
    src_index(y, H, y0, y1, ly);
    src_index(x, W, x0, x1, lx);
    const float hy = 1.f - ly, hx = 1.f - lx;
    const long long b00 = ((long long)n * H + y0) * W + x0, b01 = ((long long)n * H + y0) * W + x1,
                    b10 = ((long long)n * H + y1) * W + x0, b11 = ((long long)n * H + y1) * W + x1;
    const float dv = hy * (hx * depth_map[b00] + lx * depth_map[b01]) +
                     ly * (hx * depth_map[b10] + lx * depth_map[b11]);
    depths[p] = dv;
#pragma unroll
    for (int c = 0; c < 3; ++c) {
      const float v = hy * (hx * rgb_map[b00 * 3 + c] + lx * rgb_map[b01 * 3 + c]) +
                      ly * (hx * rgb_map[b10 * 3 + c] + lx * rgb_map[b11 * 3 + c]);
      rgbs[p * 3 + c] = v;
      const float g = gt_img[(((long long)n * 3 + c) * HH + y) * WW + x];
      e_rgb += (v - g) * (v - g);
    }
    float gd = (gt_depth[p] - (2.0f - 0.5f / 2.0f)) / 0.5f;     // coocc_ray.py:423-425
    gd = fminf(fmaxf(gd, 0.f), (float)D);
    if (gd > 0.f) {
      const float df = dv / (float)D - gd / (float)D;
      e_d += df * df;
      n_fg += 1.f;
    }
  }
#pragma unroll
  for (int o = 16; o >= 1; o >>= 1) {
    e_rgb += __shfl_xor_sync(0xffffffffu, e_rgb, o);
    e_d += __shfl_xor_sync(0xffffffffu, e_d, o);
    n_fg += __shfl_xor_sync(0xffffffffu, n_fg, o);
  }
  __shared__ float ws[3][8];
  if ((threadIdx.x & 31) == 0) {
    ws[0][threadIdx.x >> 5] = e_rgb; ws[1][threadIdx.x >> 5] = e_d; ws[2][threadIdx.x >> 5] = n_fg;
  }
  __syncthreads();
  if (threadIdx.x < 3) {
    float s = 0.f;
    for (int i = 0; i < 8; ++i) s += ws[threadIdx.x][i];
    atomicAdd(&acc[threadIdx.x], s);
  }
}

// losses[0] = loss_depth_render, losses[1] = loss_rgb
__global__ void loss_finalize_kernel(const float* __restrict__ acc, float n_rgb, float* __restrict__ losses) {
  losses[0] = acc[1] / acc[2];
  losses[1] = acc[0] / n_rgb;
}

// d rgb_map / d depth_map: one thread per low-res pixel gathers over the high-res pixels it feeds
__global__ void __launch_bounds__(128) upsample_loss_bwd_kernel(
    const float* __restrict__ rgbs, const float* __restrict__ depths, int ncam, int H, int W, int D,
    const float* __restrict__ gt_img, const float* __restrict__ gt_depth, const float* __restrict__ acc,
    const float* __restrict__ g_losses /* [2]: d/d loss_depth, d/d loss_rgb */, float* __restrict__ g_rgb_map,
    float* __restrict__ g_depth_map) {
  const int HH = 16 * H, WW = 16 * W;
  const long long lp = blockIdx.x;                    // low-res pixel, one block each
  const int wx = lp % W, hy = (lp / W) % H, n = lp / ((long long)W * H);
  const float s_rgb = g_losses[1] * 2.f / ((float)ncam * HH * WW * 3.f);
  const float s_dep = g_losses[0] * 2.f / (acc[2] * (float)D);
  float a[4] = {0.f, 0.f, 0.f, 0.f};
  const int ylo = max(0, hy * 16 - 8), yhi = min(HH - 1, hy * 16 + 23);
  const int xlo = max(0, wx * 16 - 8), xhi = min(WW - 1, wx * 16 + 23);
  const int nx = xhi - xlo + 1, ny = yhi - ylo + 1;
  for (int e = threadIdx.x; e < nx * ny; e += 128) {
    const int x = xlo + e % nx, y = ylo + e / nx;
    int y0, y1, x0, x1;
    float ly, lx;
    src_index(y, H, y0, y1, ly);
    src_index(x, W, x0, x1, lx);
    float wgt = 0.f;
    if (y0 == hy) wgt += (1.f - ly) * ((x0 == wx ? 1.f - lx : 0.f) + (x1 == wx ? lx : 0.f));
    if (y1 == hy) wgt += ly * ((x0 == wx ? 1.f - lx : 0.f) + (x1 == wx ? lx : 0.f));
    if (wgt == 0.f) continue;
    const long long p = ((long long)n * HH + y) * WW + x;
#pragma unroll
    for (int c = 0; c < 3; ++c) {
      const float g = gt_img[(((long long)n * 3 + c) * HH + y) * WW + x];
      a[c] += wgt * s_rgb * (rgbs[p * 3 + c] - g);
    }
    float gd = (gt_depth[p] - (2.0f - 0.5f / 2.0f)) / 0.5f;
    gd = fminf(fmaxf(gd, 0.f), (float)D);
    if (gd > 0.f) a[3] += wgt * s_dep * (depths[p] / (float)D - gd / (float)D);
  }
#pragma unroll
  for (int o = 16; o >= 1; o >>= 1)
#pragma unroll
    for (int j = 0; j < 4; ++j) a[j] += __shfl_xor_sync(0xffffffffu, a[j], o);
  __shared__ float ws[4][4];
  if ((threadIdx.x & 31) == 0)
    for (int j = 0; j < 4; ++j) ws[j][threadIdx.x >> 5] = a[j];
  __syncthreads();
  if (threadIdx.x < 4) {
    const float s = ws[threadIdx.x][0] + ws[threadIdx.x][1] + ws[threadIdx.x][2] + ws[threadIdx.x][3];
    if (threadIdx.x < 3) g_rgb_map[lp * 3 + threadIdx.x] = s;
    else g_depth_map[lp] = s;
  }
}

}  // namespace coocc

using namespace coocc;
#define CK_LAUNCH() (cudaGetLastError() == cudaSuccess ? 0 : COOCC_ERR_CUDA)

extern "C" int coocc_render_box(int X, int Y, int Z, int* bx, int* by, int* bz) {
  if (!bx || !by || !bz) return COOCC_ERR_ARG;
  *bx = X < kBoxX ? X : kBoxX;
  *by = Y < kBoxY ? Y : kBoxY;
  *bz = Z < kBoxZ ? Z : kBoxZ;
  return 0;
}

extern "C" int coocc_render_box_gather(const float* grid, long long ld, int X, int Y, int Z, int C,
                                       float* rows, void* stream) {
  if (!grid || !rows) return COOCC_ERR_ARG;
  int bx, by, bz;
  coocc_render_box(X, Y, Z, &bx, &by, &bz);
  const long long T = (long long)bx * by * bz;
  box_gather_kernel<<<(unsigned)((T + 7) / 8), 256, 0, (cudaStream_t)stream>>>(grid, ld, Y, Z, bx, by, bz, C, rows);
  return CK_LAUNCH();
}

extern "C" int coocc_render_box_scatter_add(const float* rows, int C, int X, int Y, int Z, float* grid,
                                            long long ld, void* stream) {
  if (!grid || !rows) return COOCC_ERR_ARG;
  int bx, by, bz;
  coocc_render_box(X, Y, Z, &bx, &by, &bz);
  const long long T = (long long)bx * by * bz;
  box_scatter_add_kernel<<<(unsigned)((T + 7) / 8), 256, 0, (cudaStream_t)stream>>>(rows, C, Y, Z, bx, by, bz, grid, ld);
  return CK_LAUNCH();
}

extern "C" int coocc_render_box_gather_bf16(const void* grid, long long ld, int X, int Y, int Z, int C, void* rows,
                                            void* stream) {
  if (!grid || !rows || (C & 7) || (ld & 7)) return COOCC_ERR_ARG;
  int bx, by, bz;
  coocc_render_box(X, Y, Z, &bx, &by, &bz);
  const long long T = (long long)bx * by * bz;
  box_gather_bf16_kernel<<<(unsigned)((T + 7) / 8), 256, 0, (cudaStream_t)stream>>>(
      reinterpret_cast<const uint4*>(grid), ld / 8, Y, Z, bx, by, bz, C / 8, reinterpret_cast<uint4*>(rows));
  return CK_LAUNCH();
}

extern "C" int coocc_render_box_scatter_bf16(const void* rows, int C, int X, int Y, int Z, void* grid, long long ld,
                                             void* stream) {
  if (!grid || !rows || (C & 7) || (ld & 7)) return COOCC_ERR_ARG;
  int bx, by, bz;
  coocc_render_box(X, Y, Z, &bx, &by, &bz);
  const long long V = (long long)X * Y * Z;
  long long blocks = (V * (C / 8) + 255) / 256;
  if (blocks > 148 * 32) blocks = 148 * 32;
  box_scatter_bf16_kernel<<<(unsigned)blocks, 256, 0, (cudaStream_t)stream>>>(
      reinterpret_cast<const uint4*>(rows), C / 8, X, Y, Z, bx, by, bz, reinterpret_cast<uint4*>(grid), ld / 8);
  return CK_LAUNCH();
}

extern "C" int coocc_render_composite_fwd(const float* geom, int ncam, int D, int H, int W, const float* tab,
                                          int X, int Y, int Z, float* rgb_map, float* depth_map, int* err,
                                          void* stream) {
  if (!geom || !tab || !rgb_map || !depth_map || D < 2 || D > kMaxD || ncam < 1) return COOCC_ERR_ARG;
  int bx, by, bz;
  coocc_render_box(X, Y, Z, &bx, &by, &bz);
  dim3 grid((H * W + 31) / 32, ncam);
  composite_fwd_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(geom, D, H * W, (const float4*)tab, bx, by, bz,
                                                              rgb_map, depth_map, err);
  return CK_LAUNCH();
}

extern "C" int coocc_render_composite_bwd(const float* geom, int ncam, int D, int H, int W, const float* tab,
                                          int X, int Y, int Z, const float* g_rgb_map,
                                          const float* g_depth_map, float* d_tab, void* stream) {
  if (!geom || !tab || !g_rgb_map || !g_depth_map || !d_tab || D < 2 || D > kMaxD) return COOCC_ERR_ARG;
  int bx, by, bz;
  coocc_render_box(X, Y, Z, &bx, &by, &bz);
  dim3 grid((H * W + 31) / 32, ncam);
  composite_bwd_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(geom, D, H * W, (const float4*)tab, bx, by, bz,
                                                              g_rgb_map, g_depth_map, d_tab);
  return CK_LAUNCH();
}

extern "C" int coocc_render_upsample_loss_fwd(const float* rgb_map, const float* depth_map, int ncam, int H,
                                              int W, int D, const float* gt_img, const float* gt_depth,
                                              float* rgbs, float* depths, float* acc3, float* losses2,
                                              void* stream) {
  if (!rgb_map || !depth_map || !gt_img || !gt_depth || !rgbs || !depths || !acc3 || !losses2) return COOCC_ERR_ARG;
  cudaStream_t st = (cudaStream_t)stream;
  cudaMemsetAsync(acc3, 0, 3 * sizeof(float), st);
  const long long total = (long long)ncam * 256 * H * W;
  int blocks = (int)((total + 255) / 256);
  if (blocks > 148 * 16) blocks = 148 * 16;
  upsample_loss_fwd_kernel<<<blocks, 256, 0, st>>>(rgb_map, depth_map, ncam, H, W, D, gt_img, gt_depth, rgbs,
                                                  depths, acc3);
  loss_finalize_kernel<<<1, 1, 0, st>>>(acc3, (float)total * 3.f, losses2);
  return CK_LAUNCH();
}

extern "C" int coocc_render_upsample_loss_bwd(const float* rgbs, const float* depths, int ncam, int H, int W,
                                              int D, const float* gt_img, const float* gt_depth,
                                              const float* acc3, const float* g_losses2, float* g_rgb_map,
                                              float* g_depth_map, void* stream) {
  if (!rgbs || !depths || !gt_img || !gt_depth || !acc3 || !g_losses2 || !g_rgb_map || !g_depth_map) return COOCC_ERR_ARG;
  upsample_loss_bwd_kernel<<<ncam * H * W, 128, 0, (cudaStream_t)stream>>>(
      rgbs, depths, ncam, H, W, D, gt_img, gt_depth, acc3, g_losses2, g_rgb_map, g_depth_map);
  return CK_LAUNCH();
}
