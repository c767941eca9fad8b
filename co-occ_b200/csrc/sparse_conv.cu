// sparse_conv.cu -- index side of the sparse LiDAR encoder:
// SparseLiDAREnc8x, P/coocc/voxel_encoder/sparse_lidar_enc.py:125-177, which the reference runs on the third-party
// spconv 2.3.6 (SubMConv3d / SparseConv3d, not vendored under /root/reference -- the algorithm restated here is
// spconv's published one: rulebook of (input row, kernel offset, output row) pairs + gather-GEMM-scatter).
//
// B200 formulation: a sparse convolution over N active voxels is ONE dense GEMM on an explicit sparse im2col matrix
//   cols[i, k*Cin + c] = feats[nbr[i, k], c]   (0 where the neighbour is inactive),   out = cols @ W^T
// with W = the spconv weight [Cout, kz, ky, kx, Cin] read in place as [Cout, 27*Cin].  N is ~1e5, so the im2col matrix
// is a few hundred MB at most and the GEMM runs on the tcgen05 conv kernel (conv_tc.cu as a 1x1x1 convolution over N
// "voxels", with its BatchNorm-statistics epilogue).  This file provides the parts around that GEMM:
//   sp_flag_outputs   active output sites of a strided SparseConv3d (every site reached by an active input),
//   sp_neighbors      the [N_out, 27] neighbour table (rulebook in output-stationary form) through a dense index grid
//                     (one int32 per cell of the level's grid -- 164 MB at 800x800x64, nothing on a 180 GB part),
//   sp_gather_cols    the im2col gather, one row of Cin floats per 16-byte-vectorised copy,
//   sp_scatter_cols   its transpose (gradient w.r.t. the input features), atomics.
// Coordinates are spconv's (batch, z, y, x) int32 rows; batch must be 0 (the path is batch-1, sparse_lidar_enc.py:166).
#include <cuda_runtime.h>
#include <stdint.h>

#include "../../include/coocc_b200.h"

namespace coocc {
namespace sp {

// flags[o] = 1 for every output site o = (i + pad - k) / stride (exact division, in range) of an active input i
__global__ void __launch_bounds__(256) flag_outputs_kernel(const int* __restrict__ coords, int n, int stride, int pad,
                                                           int oD, int oH, int oW, unsigned char* __restrict__ flags) {
  const long long total = (long long)n * 27;
  for (long long t = blockIdx.x * 256LL + threadIdx.x; t < total; t += gridDim.x * 256LL) {
    const int i = (int)(t / 27), k = (int)(t % 27);
    const int kz = k / 9, ky = (k / 3) % 3, kx = k % 3;
    const int z = coords[i * 4 + 1] + pad - kz, y = coords[i * 4 + 2] + pad - ky, x = coords[i * 4 + 3] + pad - kx;
    if (z < 0 || y < 0 || x < 0 || z % stride || y % stride || x % stride) continue;
    const int oz = z / stride, oy = y / stride, ox = x / stride;
    if (oz >= oD || oy >= oH || ox >= oW) continue;
    flags[((long long)oz * oH + oy) * oW + ox] = 1;
  }
}

// nbr[o, k] = row of the input voxel at out_coord * stride - pad + k, or -1
__global__ void __launch_bounds__(256) neighbors_kernel(const int* __restrict__ out_coords, int n_out, int stride, int pad,
                                                        const int* __restrict__ grid, int D, int H, int W,
                                                        int* __restrict__ nbr) {
  const long long total = (long long)n_out * 27;
  for (long long t = blockIdx.x * 256LL + threadIdx.x; t < total; t += gridDim.x * 256LL) {
    const int o = (int)(t / 27), k = (int)(t % 27);
    const int kz = k / 9, ky = (k / 3) % 3, kx = k % 3;
    const int z = out_coords[o * 4 + 1] * stride - pad + kz, y = out_coords[o * 4 + 2] * stride - pad + ky,
              x = out_coords[o * 4 + 3] * stride - pad + kx;
    int r = -1;
    if (z >= 0 && y >= 0 && x >= 0 && z < D && y < H && x < W) r = grid[((long long)z * H + y) * W + x];
    nbr[t] = r;
  }
}

// cols[o, k*C + c] = feats[nbr[o,k], c]; one thread per 16-byte vector of a (row, offset) pair
__global__ void __launch_bounds__(256) gather_cols_kernel(const float* __restrict__ feats, long long ldf, int C,
                                                          const int* __restrict__ nbr, int n_out,
                                                          float* __restrict__ cols, long long ldc) {
  const int c4 = C >> 2;
  const long long total = (long long)n_out * 27 * c4;
  for (long long t = blockIdx.x * 256LL + threadIdx.x; t < total; t += gridDim.x * 256LL) {
    const int v = (int)(t % c4);
    const long long ok = t / c4;
    const int k = (int)(ok % 27);
    const long long o = ok / 27;
    const int r = nbr[ok];
    float4 val = make_float4(0.f, 0.f, 0.f, 0.f);
    if (r >= 0) val = *reinterpret_cast<const float4*>(feats + (long long)r * ldf + v * 4);
    *reinterpret_cast<float4*>(cols + o * ldc + (long long)k * C + v * 4) = val;
  }
}

// dfeats[nbr[o,k], c] += dcols[o, k*C + c]
__global__ void __launch_bounds__(256) scatter_cols_kernel(const float* __restrict__ dcols, long long ldc, int C,
                                                           const int* __restrict__ nbr, int n_out,
                                                           float* __restrict__ dfeats, long long ldf) {
  const int c4 = C >> 2;
  const long long total = (long long)n_out * 27 * c4;
  for (long long t = blockIdx.x * 256LL + threadIdx.x; t < total; t += gridDim.x * 256LL) {
    const int v = (int)(t % c4);
    const long long ok = t / c4;
    const int k = (int)(ok % 27);
    const long long o = ok / 27;
    const int r = nbr[ok];
    if (r < 0) continue;
    const float4 g = *reinterpret_cast<const float4*>(dcols + o * ldc + (long long)k * C + v * 4);
    float* d = dfeats + (long long)r * ldf + v * 4;
    asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(d), "f"(g.x), "f"(g.y), "f"(g.z), "f"(g.w)
                 : "memory");
  }
}

static int grid_for(long long total) {
  long long b = (total + 255) / 256;
  if (b > 148LL * 32) b = 148LL * 32;
  return (int)(b < 1 ? 1 : b);
}

}  // namespace sp
}  // namespace coocc

using namespace coocc::sp;
#define CK_LAUNCH() (cudaGetLastError() == cudaSuccess ? 0 : COOCC_ERR_CUDA)

// coords: int32 [n][4] = (batch, z, y, x) of the active inputs; flags: uint8 [oD*oH*oW], zeroed by the caller.
extern "C" int coocc_sp_flag_outputs(const int* coords, int n, int stride, int pad, int oD, int oH, int oW,
                                     unsigned char* flags, void* stream) {
  if (!coords || !flags || n < 0 || stride < 1 || pad < 0 || oD < 1 || oH < 1 || oW < 1) return COOCC_ERR_ARG;
  if (n == 0) return 0;
  flag_outputs_kernel<<<grid_for((long long)n * 27), 256, 0, (cudaStream_t)stream>>>(coords, n, stride, pad, oD, oH, oW, flags);
  return CK_LAUNCH();
}

// out_coords: int32 [n_out][4]; grid: int32 [D*H*W] of the INPUT level (row id or -1); nbr: int32 [n_out][27],
// offset k = (kz*3 + ky)*3 + kx <-> input site out*stride - pad + (kz,ky,kx).  SubMConv3d: stride 1, pad 1,
// out_coords = the input coordinates; SparseConv3d(3, stride 2, padding 1): stride 2, pad 1.
extern "C" int coocc_sp_neighbors(const int* out_coords, int n_out, int stride, int pad, const int* grid, int D, int H, int W,
                                  int* nbr, void* stream) {
  if (!out_coords || !grid || !nbr || n_out < 0 || stride < 1 || pad < 0 || D < 1 || H < 1 || W < 1) return COOCC_ERR_ARG;
  if (n_out == 0) return 0;
  neighbors_kernel<<<grid_for((long long)n_out * 27), 256, 0, (cudaStream_t)stream>>>(out_coords, n_out, stride, pad, grid,
                                                                                     D, H, W, nbr);
  return CK_LAUNCH();
}

// feats [n_in][ldf] fp32 (C % 4 == 0, ldf % 4 == 0); cols [n_out][ldc] with ldc >= 27*C, ldc % 4 == 0
extern "C" int coocc_sp_gather_cols(const float* feats, long long ldf, int C, const int* nbr, int n_out, float* cols,
                                    long long ldc, void* stream) {
  if (!feats || !nbr || !cols || C < 4 || (C & 3) || (ldf & 3) || (ldc & 3) || ldc < 27LL * C || n_out < 0)
    return COOCC_ERR_ARG;
  if (n_out == 0) return 0;
  gather_cols_kernel<<<grid_for((long long)n_out * 27 * (C >> 2)), 256, 0, (cudaStream_t)stream>>>(feats, ldf, C, nbr,
                                                                                                   n_out, cols, ldc);
  return CK_LAUNCH();
}

// dfeats [n_in][ldf] += transpose of the gather (zeroed by the caller)
extern "C" int coocc_sp_scatter_cols(const float* dcols, long long ldc, int C, const int* nbr, int n_out, float* dfeats,
                                     long long ldf, void* stream) {
  if (!dcols || !nbr || !dfeats || C < 4 || (C & 3) || (ldf & 3) || (ldc & 3) || ldc < 27LL * C || n_out < 0)
    return COOCC_ERR_ARG;
  if (n_out == 0) return 0;
  scatter_cols_kernel<<<grid_for((long long)n_out * 27 * (C >> 2)), 256, 0, (cudaStream_t)stream>>>(dcols, ldc, C, nbr,
                                                                                                    n_out, dfeats, ldf);
  return CK_LAUNCH();
}
