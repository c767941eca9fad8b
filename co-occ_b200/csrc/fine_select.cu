// fine_select.cu -- device-side point selection of the OccHead fine / cascade stage
// (P/coocc/dense_heads/occ_head.py:182-205, P/utils/coordinate_transform.py:3-21), without a host round trip so the
// whole step can live in one CUDA graph.
//
// Reference: mask = argmax(coarse logits) != empty; coords = nonzero(mask); if N >= topk keep a random subset of topk
// parents (torch.randperm on the host); every kept parent expands to its ratio^3 children.  N is data dependent, so
// the reference synchronises (nonzero) and draws from torch's CPU generator.  Here:
//   1. keys:   one thread per coarse voxel, argmax over C logits (first maximum), occupied voxels get a 30-bit
//              pseudo-random key hash(seed, draw, voxel), free voxels the maximum key; N is counted on the fly;
//   2. sort:   the shared LSD radix sort (radix_sort.cuh) of (key, voxel) over all V voxels -- the first
//              P = min(N, topk) entries are a uniform random subset of the occupied voxels (all of them if N <= topk);
//   3. expand: slot j < P -> children (parent * ratio + offset), child index o * topk + j (offset-major like the
//              reference's [r^3, 3, P] -> [3, r^3 * P] reshape); slots j >= P are padding with coordinates 0;
//   4. labels (coocc_fine_gather_labels, called from loss_point, occ_head.py:298): label-grid entry of every child,
//              `ignore` for padding slots -- every loss of loss_point masks those out.
// The buffers have the fixed capacity topk, the kernels read N from device memory: one graph serves every scene.
// The eager module path (modules.OccHead.forward_fine, host randperm) stays bit-identical to the reference's draw;
// this path draws a different (equally distributed) subset -- tests/test_gpu_fine.py checks the set properties.
#include <cuda_runtime.h>
#include <stdint.h>

#include "../../include/coocc_b200.h"
#include "radix_sort.cuh"

namespace coocc {
namespace fine {

constexpr uint32_t kFreeKey = 0x3FFFFFFFu;        // 30 significant bits (3 radix passes of 10)

__device__ __forceinline__ uint32_t mix32(uint32_t x) {       // murmur3 finaliser
  x ^= x >> 16; x *= 0x85EBCA6Bu; x ^= x >> 13; x *= 0xC2B2AE35u; x ^= x >> 16;
  return x;
}

struct SelWs {
  uint32_t *kA, *vA, *kB, *vB;
  int* counters;
};

static size_t align_up(size_t x) { return (x + 255) & ~(size_t)255; }

static size_t ws_layout(char* base, int V, SelWs* w) {
  size_t off = 0;
  auto take = [&](size_t bytes) { char* p = base ? base + off : nullptr; off += align_up(bytes); return p; };
  w->kA = reinterpret_cast<uint32_t*>(take((size_t)V * 4));
  w->vA = reinterpret_cast<uint32_t*>(take((size_t)V * 4));
  w->kB = reinterpret_cast<uint32_t*>(take((size_t)V * 4));
  w->vB = reinterpret_cast<uint32_t*>(take((size_t)V * 4));
  w->counters = reinterpret_cast<int*>(take(radix_counters_bytes(V, 1)));
  return off;
}

__global__ void __launch_bounds__(256) select_keys_kernel(const float* __restrict__ logits, long long ld, int V, int C,
                                                          int empty_idx, const unsigned long long* __restrict__ state,
                                                          uint32_t* __restrict__ keys, uint32_t* __restrict__ vals,
                                                          int* __restrict__ nsel) {
  __shared__ int s_cnt;
  if (threadIdx.x == 0) s_cnt = 0;
  __syncthreads();
  const uint32_t seed = (uint32_t)state[0] ^ mix32((uint32_t)(state[0] >> 32) + 0x9E3779B9u);
  const uint32_t draw = mix32((uint32_t)state[1] * 0x9E3779B1u + 0x7F4A7C15u);
  int mine = 0;
  for (int v = blockIdx.x * 256 + threadIdx.x; v < V; v += gridDim.x * 256) {
    const float* row = logits + (long long)v * ld;
    float best = row[0];
    int arg = 0;
    for (int c = 1; c < C; ++c) {
      const float x = row[c];
      if (x > best) { best = x; arg = c; }
    }
    const bool occ = arg != empty_idx;
    uint32_t key = kFreeKey;
    if (occ) {
      key = mix32(mix32((uint32_t)v ^ seed) + draw) & 0x3FFFFFFFu;
      if (key == kFreeKey) key = kFreeKey - 1;
      ++mine;
    }
    keys[v] = key;
    vals[v] = (uint32_t)v;
  }
  if (mine) atomicAdd(&s_cnt, mine);
  __syncthreads();
  if (threadIdx.x == 0 && s_cnt) atomicAdd(&nsel[0], s_cnt);
}

__global__ void __launch_bounds__(256) select_expand_kernel(const uint32_t* __restrict__ sorted_vals, int Y, int Z,
                                                            int ratio, int topk, int* __restrict__ coords,
                                                            int* __restrict__ nsel,
                                                            unsigned long long* __restrict__ state) {
  const int r3 = ratio * ratio * ratio;
  const long long M = (long long)r3 * topk;
  const int N = nsel[0];
  const int P = N < topk ? N : topk;
  for (long long i = blockIdx.x * 256LL + threadIdx.x; i < M; i += gridDim.x * 256LL) {
    const int o = (int)(i / topk), j = (int)(i % topk);
    int cx = 0, cy = 0, cz = 0;
    if (j < P) {
      const int v = (int)sorted_vals[j];
      const int z = v % Z, y = (v / Z) % Y, x = v / (Z * Y);
      const int oz = o % ratio, oy = (o / ratio) % ratio, ox = o / (ratio * ratio);   // meshgrid 'ij' order
      cx = x * ratio + ox; cy = y * ratio + oy; cz = z * ratio + oz;
    }
    coords[i] = cx; coords[M + i] = cy; coords[2 * M + i] = cz;
  }
  if (blockIdx.x == 0 && threadIdx.x == 0) {
    nsel[1] = P;
    state[1] += 1;              // next replay draws a new subset
  }
}

__global__ void __launch_bounds__(256) gather_labels_kernel(const int* __restrict__ coords, long long M, int topk,
                                                            const int* __restrict__ nsel, const void* __restrict__ gt,
                                                            int gt_bytes, int GY, int GZ, int ignore,
                                                            int* __restrict__ labels) {
  const int P = nsel ? nsel[1] : topk;
  for (long long i = blockIdx.x * 256LL + threadIdx.x; i < M; i += gridDim.x * 256LL) {
    int lab = ignore;
    if ((int)(i % topk) < P) {
      const long long gi = ((long long)coords[i] * GY + coords[M + i]) * GZ + coords[2 * M + i];
      lab = gt_bytes == 8 ? (int)reinterpret_cast<const long long*>(gt)[gi]
                          : gt_bytes == 4 ? reinterpret_cast<const int*>(gt)[gi]
                                          : (int)reinterpret_cast<const unsigned char*>(gt)[gi];
    }
    labels[i] = lab;
  }
}

}  // namespace fine
}  // namespace coocc

using namespace coocc;
using namespace coocc::fine;

extern "C" long long coocc_fine_select_workspace(int V) {
  if (V < 1) return -1;
  SelWs w;
  return (long long)ws_layout(nullptr, V, &w);
}

// logits [V, C] fp32 rows (row stride ld) of the coarse prediction on an X x Y x Z grid; state = device uint64[2]
// (seed, draw counter -- incremented by this call); coords = int32 [3, ratio^3 * topk];
// nsel = device int32[2] <- (N occupied coarse voxels, P = min(N, topk) selected parents).
extern "C" int coocc_fine_select(const float* logits, long long ld, int X, int Y, int Z, int C, int empty_idx, int ratio,
                                 int topk, unsigned long long* state, int* coords, int* nsel, void* workspace,
                                 void* stream) {
  if (!logits || !state || !coords || !nsel || !workspace || X < 1 || Y < 1 || Z < 1 || C < 2 || ratio < 1 || topk < 1 ||
      ld < C)
    return COOCC_ERR_ARG;
  const long long Vl = (long long)X * Y * Z;
  if (Vl > (1LL << 30)) return COOCC_ERR_CAPACITY;
  const int V = (int)Vl;
  cudaStream_t st = (cudaStream_t)stream;
  SelWs w;
  ws_layout(reinterpret_cast<char*>(workspace), V, &w);
  if (cudaMemsetAsync(nsel, 0, 2 * sizeof(int), st) != cudaSuccess) return COOCC_ERR_CUDA;
  int blocks = (V + 255) / 256;
  if (blocks > 148 * 8) blocks = 148 * 8;
  select_keys_kernel<<<blocks, 256, 0, st>>>(logits, ld, V, C, empty_idx, state, w.kA, w.vA, nsel);
  uint32_t *ko = nullptr, *vo = nullptr;
  if (radix_sort_pairs(w.kA, w.vA, w.kB, w.vB, V, 1, 30, w.counters, st, &ko, &vo) != 0) return COOCC_ERR_CUDA;
  const long long M = (long long)ratio * ratio * ratio * topk;
  long long b = (M + 255) / 256;
  if (b > 148LL * 8) b = 148LL * 8;
  select_expand_kernel<<<(unsigned)b, 256, 0, st>>>(vo, Y, Z, ratio, topk, coords, nsel, state);
  return cudaGetLastError() == cudaSuccess ? 0 : COOCC_ERR_CUDA;
}

// labels[i] = gt[coords[:, i]] for the M = r^3 * topk slots of coocc_fine_select (slot i belongs to parent i % topk),
// `ignore` for slots of parents >= nsel[1].  gt = [GX, GY, GZ] integers of gt_bytes (1, 4 or 8) bytes.
// nsel == NULL: every slot is valid (coordinates from the host path).
extern "C" int coocc_fine_gather_labels(const int* coords, long long M, int topk, const int* nsel, const void* gt,
                                        int gt_bytes, int GX, int GY, int GZ, int ignore, int* labels, void* stream) {
  if (!coords || !gt || !labels || M < 0 || topk < 1 || GX < 1 || GY < 1 || GZ < 1 ||
      (gt_bytes != 1 && gt_bytes != 4 && gt_bytes != 8))
    return COOCC_ERR_ARG;
  if (M == 0) return 0;
  long long b = (M + 255) / 256;
  if (b > 148LL * 8) b = 148LL * 8;
  gather_labels_kernel<<<(unsigned)b, 256, 0, (cudaStream_t)stream>>>(coords, M, topk, nsel, gt, gt_bytes, GY, GZ, ignore,
                                                                      labels);
  return cudaGetLastError() == cudaSuccess ? 0 : COOCC_ERR_CUDA;
}
