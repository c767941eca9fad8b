// occ_loss.cu -- the occupancy head's voxel losses (coarse level):
//   OccHead.loss_voxel   P/coocc/dense_heads/occ_head.py:267-293
//     label majority-vote downsample                       :269-280 (torch.mode, ties -> smallest value)
//     CE_ssc_loss (class-weighted CE, ignore 255)          P/utils/semkitti.py:139-149
//     sem_scal_loss (per-class precision/recall/specificity BCE)   semkitti.py:92-136
//     geo_scal_loss (empty / non-empty precision/recall/specificity) semkitti.py:62-89
//     lovasz_softmax (classes='present', ignore 255)       P/coocc/dense_heads/lovasz_softmax.py:20-34,156-203
//
// One pass over the logits produces softmax, the CE terms, every per-class sum the two "scal" losses
// need and the Lovasz errors |fg - p_c| (class-major); a segmented LSD radix sort (3 x 10 bits on the
// 30 significant bits of an error in [0,1]) orders the errors of every class at once; a chunked scan
// turns sorted foreground flags into the Lovasz gradient and the loss; the backward is one pass that
// recomputes the softmax and applies  d logit = p (dp - <p, dp>) + CE term  with
//   dp_c = sem_a[c] [t=c] + sem_b[c] + (c==empty)(geo_a [t=empty] + geo_b) + G[c][v].
// All HBM-bound; per launch algorithmic bytes are given in DESIGN.md.
#include <cuda_runtime.h>
#include <math.h>
#include <stdint.h>

#include "../../include/coocc_b200.h"
#include "radix_sort.cuh"

namespace coocc {

constexpr int kMaxCls = 32;
constexpr int kRadixBits = 10;
constexpr int kBins = 1 << kRadixBits;     // 1024
constexpr int kChunk = 4096;               // items per warp chunk (sort and scan)
constexpr int kWarpsPerBlock = 4;
constexpr uint32_t kOneBits = 0x3F800000u; // bits of 1.0f

struct OccAcc {
  double sp[kMaxCls];    // sum over valid voxels of p_c
  double spt[kMaxCls];   // sum over valid voxels with t == c of p_c
  double nt[kMaxCls];    // number of valid voxels with t == c
  double lov[kMaxCls];   // Lovasz loss of class c
  double ce_num, ce_den, nvalid, pad;
};
struct OccCoef {
  float sem_a[kMaxCls], sem_b[kMaxCls];   // d sem / d nominator_c , d sem / d sum_p_c
  float geo_a, geo_b;                     // d geo / d nominator_empty , d geo / d sum_p_empty
  float ce_scale;                         // 1 / sum of class weights over valid voxels
  float lov_scale;                        // 1 / number of present classes
};

struct OccWs {
  OccAcc* acc;
  OccCoef* coef;
  float* G;            // [C][V] d lovasz / d p_c(v)
  uint32_t* keyA; uint32_t* keyB; uint32_t* valA; uint32_t* valB;   // [C][V]
  int* counters;       // [C][nchunks][kBins]
  int* lcnt;           // [C][nchunks]
  int nchunks;
};

static size_t align_up(size_t x) { return (x + 255) / 256 * 256; }

static size_t ws_layout(char* base, int V, int C, OccWs* w) {
  const int nchunks = (V + kChunk - 1) / kChunk;
  size_t off = 0;
  auto take = [&](size_t bytes) { char* p = base ? base + off : nullptr; off += align_up(bytes); return p; };
  OccAcc* acc = reinterpret_cast<OccAcc*>(take(sizeof(OccAcc)));
  OccCoef* coef = reinterpret_cast<OccCoef*>(take(sizeof(OccCoef)));
  float* G = reinterpret_cast<float*>(take((size_t)C * V * 4));
  uint32_t* kA = reinterpret_cast<uint32_t*>(take((size_t)C * V * 4));
  uint32_t* kB = reinterpret_cast<uint32_t*>(take((size_t)C * V * 4));
  uint32_t* vA = reinterpret_cast<uint32_t*>(take((size_t)C * V * 4));
  uint32_t* vB = reinterpret_cast<uint32_t*>(take((size_t)C * V * 4));
  int* counters = reinterpret_cast<int*>(take(radix_counters_bytes(V, C)));
  int* lcnt = reinterpret_cast<int*>(take((size_t)C * nchunks * 4));
  if (w) *w = OccWs{acc, coef, G, kA, kB, vA, vB, counters, lcnt, nchunks};
  return off;
}

// ------------------------------------------------------------------------------------------
// label downsample: mode over ratio^3 fine labels; in a non-empty cell (label sum != empty_idx) every
// zero label counts as a distinct negative value (occ_head.py:274-277), so it can only win a tie of
// singletons, in which case the cell becomes 255; ties go to the smallest value (torch.mode).
// ------------------------------------------------------------------------------------------
template <typename T>
__global__ void label_mode_kernel(const T* __restrict__ lab, int X, int Y, int Z, int r, int empty_idx,
                                  int* __restrict__ out) {
  const int v = blockIdx.x * blockDim.x + threadIdx.x;
  if (v >= X * Y * Z) return;
  const int z = v % Z, y = (v / Z) % Y, x = v / (Z * Y);
  const int n = r * r * r;
  int vals[64];
  long long sum = 0;
  int nzero = 0;
  for (int i = 0; i < n; ++i) {
    const int dz = i % r, dy = (i / r) % r, dx = i / (r * r);
    const long long idx = (((long long)(x * r + dx) * (Y * r)) + (y * r + dy)) * (Z * r) + (z * r + dz);
    const int l = (int)lab[idx];
    vals[i] = l;
    sum += l;
    nzero += (l == 0);
  }
  int res;
  if (sum == empty_idx) {
    // empty cell: plain mode of the raw labels
    int best = 0, bestc = 0;
    for (int i = 0; i < n; ++i) {
      int c = 0;
      for (int j = 0; j < n; ++j) c += (vals[j] == vals[i]);
      if (c > bestc || (c == bestc && vals[i] < best)) { best = vals[i]; bestc = c; }
    }
    res = best;
  } else {
    int best = 0, bestc = 0;
    for (int i = 0; i < n; ++i) {
      if (vals[i] == 0) continue;
      int c = 0;
      for (int j = 0; j < n; ++j) c += (vals[j] == vals[i]);
      if (c > bestc || (c == bestc && vals[i] < best)) { best = vals[i]; bestc = c; }
    }
    // every zero is a singleton negative value: it wins (smallest value) iff the best count is 1
    res = (nzero > 0 && bestc <= 1) ? -1 : best;
  }
  out[v] = res < 0 ? 255 : res;
}

// ------------------------------------------------------------------------------------------
// pass 1: softmax, sums, Lovasz errors -> sort keys
// ------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) occ_stats_kernel(const float* __restrict__ logits, long long ld,
                                                        const int* __restrict__ labels, int V, int C,
                                                        const float* __restrict__ class_w, int ignore,
                                                        OccAcc* __restrict__ acc, uint32_t* __restrict__ keys,
                                                        uint32_t* __restrict__ vals) {
  __shared__ double s_sp[kMaxCls], s_spt[kMaxCls], s_nt[kMaxCls], s_misc[3];
  if (threadIdx.x < kMaxCls) { s_sp[threadIdx.x] = 0; s_spt[threadIdx.x] = 0; s_nt[threadIdx.x] = 0; }
  if (threadIdx.x < 3) s_misc[threadIdx.x] = 0;
  __syncthreads();
  float l_sp[kMaxCls];
#pragma unroll
  for (int c = 0; c < kMaxCls; ++c) l_sp[c] = 0.f;
  float ce_num = 0.f, ce_den = 0.f;
  int nvalid = 0;
  for (int v = blockIdx.x * blockDim.x + threadIdx.x; v < V; v += gridDim.x * blockDim.x) {
    const int t = labels[v];
    const bool valid = t != ignore;
    const float* row = logits + (long long)v * ld;
    float x[kMaxCls];
    float m = -INFINITY;
#pragma unroll
    for (int c = 0; c < kMaxCls; ++c)
      if (c < C) { x[c] = row[c]; m = fmaxf(m, x[c]); }
    float s = 0.f;
#pragma unroll
    for (int c = 0; c < kMaxCls; ++c)
      if (c < C) { x[c] = expf(x[c] - m); s += x[c]; }
    const float inv = 1.f / s;
    // log-softmax of the target class the way F.cross_entropy evaluates it: (x_t - max) - log(sum)
    const float logp_t = (valid && t >= 0 && t < C) ? (row[t] - m) - logf(s) : 0.f;
#pragma unroll
    for (int c = 0; c < kMaxCls; ++c) {
      if (c < C) {
        const float p = x[c] * inv;
        const bool fg = valid && t == c;
        float e = 0.f;                      // ignored voxels: e = 0, fg = 0 (contribute nothing)
        if (valid) {
          l_sp[c] += p;
          e = fg ? 1.f - p : p;             // |fg - p|
        }
        // descending order of e == ascending order of (bits(1.0) - bits(e)); e in [0,1]
        keys[(size_t)c * V + v] = kOneBits - __float_as_uint(fminf(fmaxf(e, 0.f), 1.f));
        vals[(size_t)c * V + v] = (uint32_t)v;
        if (fg) {
          atomicAdd(&s_spt[c], (double)p);
          atomicAdd(&s_nt[c], 1.0);
          const float w = class_w ? class_w[c] : 1.f;
          ce_num += w * -logp_t;
          ce_den += w;
        }
      }
    }
    nvalid += valid ? 1 : 0;
  }
  // block reduction of the per-thread sums
#pragma unroll
  for (int c = 0; c < kMaxCls; ++c) {
    if (c < C) {
      float t = l_sp[c];
#pragma unroll
      for (int o = 16; o >= 1; o >>= 1) t += __shfl_xor_sync(0xffffffffu, t, o);
      if ((threadIdx.x & 31) == 0) atomicAdd(&s_sp[c], (double)t);
    }
  }
  float a = ce_num, b = ce_den, n = (float)nvalid;
#pragma unroll
  for (int o = 16; o >= 1; o >>= 1) {
    a += __shfl_xor_sync(0xffffffffu, a, o);
    b += __shfl_xor_sync(0xffffffffu, b, o);
    n += __shfl_xor_sync(0xffffffffu, n, o);
  }
  if ((threadIdx.x & 31) == 0) {
    atomicAdd(&s_misc[0], (double)a);
    atomicAdd(&s_misc[1], (double)b);
    atomicAdd(&s_misc[2], (double)n);
  }
  __syncthreads();
  if (threadIdx.x < C) {
    atomicAdd(&acc->sp[threadIdx.x], s_sp[threadIdx.x]);
    atomicAdd(&acc->spt[threadIdx.x], s_spt[threadIdx.x]);
    atomicAdd(&acc->nt[threadIdx.x], s_nt[threadIdx.x]);
  }
  if (threadIdx.x == 0) {
    atomicAdd(&acc->ce_num, s_misc[0]);
    atomicAdd(&acc->ce_den, s_misc[1]);
    atomicAdd(&acc->nvalid, s_misc[2]);
  }
}

// ------------------------------------------------------------------------------------------
// segmented LSD radix sort, one warp per chunk of kChunk consecutive items of one class
// ------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(32 * kWarpsPerBlock) radix_hist_kernel(const uint32_t* __restrict__ keys, int V,
                                                                        int nchunks, int shift,
                                                                        int* __restrict__ counters) {
  __shared__ int hist[kWarpsPerBlock][kBins];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int chunk = blockIdx.x * kWarpsPerBlock + warp;
  const int c = blockIdx.y;
  for (int i = lane; i < kBins; i += 32) hist[warp][i] = 0;
  __syncwarp();
  if (chunk < nchunks) {
    const uint32_t* k = keys + (size_t)c * V;
    const int beg = chunk * kChunk, end = min(V, beg + kChunk);
    for (int i = beg + lane; i < end; i += 32) atomicAdd(&hist[warp][(k[i] >> shift) & (kBins - 1)], 1);
    __syncwarp();
    int* dst = counters + ((size_t)c * nchunks + chunk) * kBins;          // [class][chunk][digit]
    for (int d = lane; d < kBins; d += 32) dst[d] = hist[warp][d];
  }
}

// Exclusive scan of counters[seg] in (digit, chunk) order, chunk-major [chunk][digit] storage so that every
// step is one coalesced 4 KB row.  The chunk range is cut into kScanSlices slices handled by different
// blocks: (1) per-slice digit totals, (2) one block per segment turns them into slice bases, (3) every slice
// writes its exclusive prefixes.
constexpr int kScanSlices = 32;

__global__ void __launch_bounds__(kBins) radix_scan1_kernel(const int* __restrict__ counters, int nchunks,
                                                            int* __restrict__ slice_tot) {
  const int seg = blockIdx.y, s = blockIdx.x, d = threadIdx.x;
  const int per = (nchunks + kScanSlices - 1) / kScanSlices;
  const int beg = s * per, end = min(nchunks, beg + per);
  const int* col = counters + (size_t)seg * nchunks * kBins + d;
  int tot = 0;
  for (int i = beg; i < end; ++i) tot += col[(size_t)i * kBins];
  slice_tot[((size_t)seg * kScanSlices + s) * kBins + d] = tot;
}

__global__ void __launch_bounds__(kBins) radix_scan2_kernel(int* __restrict__ slice_tot) {
  __shared__ int wsum[kBins / 32];
  int* st = slice_tot + (size_t)blockIdx.x * kScanSlices * kBins + threadIdx.x;
  int tot = 0;
  for (int s = 0; s < kScanSlices; ++s) tot += st[(size_t)s * kBins];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  int incl = tot;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    const int t = __shfl_up_sync(0xffffffffu, incl, o);
    if (lane >= o) incl += t;
  }
  if (lane == 31) wsum[warp] = incl;
  __syncthreads();
  if (warp == 0) {
    int w = wsum[lane];
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      const int t = __shfl_up_sync(0xffffffffu, w, o);
      if (lane >= o) w += t;
    }
    wsum[lane] = w;
  }
  __syncthreads();
  int base = incl - tot + (warp > 0 ? wsum[warp - 1] : 0);
  for (int s = 0; s < kScanSlices; ++s) {
    const int t = st[(size_t)s * kBins];
    st[(size_t)s * kBins] = base;
    base += t;
  }
}

__global__ void __launch_bounds__(kBins) radix_scan3_kernel(int* __restrict__ counters, int nchunks,
                                                            const int* __restrict__ slice_tot) {
  const int seg = blockIdx.y, s = blockIdx.x, d = threadIdx.x;
  const int per = (nchunks + kScanSlices - 1) / kScanSlices;
  const int beg = s * per, end = min(nchunks, beg + per);
  int* col = counters + (size_t)seg * nchunks * kBins + d;
  int base = slice_tot[((size_t)seg * kScanSlices + s) * kBins + d];
  for (int i = beg; i < end; ++i) {
    const int t = col[(size_t)i * kBins];
    col[(size_t)i * kBins] = base;
    base += t;
  }
}

__global__ void __launch_bounds__(32 * kWarpsPerBlock) radix_scatter_kernel(const uint32_t* __restrict__ keys_in,
                                                                           const uint32_t* __restrict__ vals_in,
                                                                           int V, int nchunks, int shift,
                                                                           const int* __restrict__ counters,
                                                                           uint32_t* __restrict__ keys_out,
                                                                           uint32_t* __restrict__ vals_out) {
  __shared__ int off[kWarpsPerBlock][kBins];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int chunk = blockIdx.x * kWarpsPerBlock + warp;
  const int c = blockIdx.y;
  if (chunk >= nchunks) return;
  const int* src = counters + ((size_t)c * nchunks + chunk) * kBins;
  for (int d = lane; d < kBins; d += 32) off[warp][d] = src[d];
  __syncwarp();
  const uint32_t* k = keys_in + (size_t)c * V;
  const uint32_t* vv = vals_in + (size_t)c * V;
  uint32_t* ko = keys_out + (size_t)c * V;
  uint32_t* vo = vals_out + (size_t)c * V;
  const int beg = chunk * kChunk, end = min(V, beg + kChunk);
  for (int i0 = beg; i0 < end; i0 += 32) {
    const int i = i0 + lane;
    const bool act = i < end;
    uint32_t key = 0, val = 0;
    int d = kBins;                       // inactive lanes share a private pseudo-digit
    if (act) { key = k[i]; val = vv[i]; d = (key >> shift) & (kBins - 1); }
    const uint32_t peers = __match_any_sync(0xffffffffu, d);
    const int rank = __popc(peers & ((1u << lane) - 1u));
    int base = 0;
    if (act) base = off[warp][d];
    __syncwarp();
    if (act && rank == 0) off[warp][d] = base + __popc(peers);
    __syncwarp();
    if (act) { ko[base + rank] = key; vo[base + rank] = val; }
  }
}

// Sorts `nseg` independent segments of n (key, value) pairs (segment s at offset s*n) ascending by the
// low `bits` key bits (rounded up to a multiple of 10), stable.  Ping-pongs between (kA, vA) and (kB, vB);
// *k_out / *v_out receive the buffers that hold the result.  counters: radix_counters_bytes(n, nseg).
size_t radix_counters_bytes(int n, int nseg) {
  const int nchunks = (n + kChunk - 1) / kChunk;
  return ((size_t)nseg * nchunks * kBins + (size_t)nseg * kScanSlices * kBins) * sizeof(int);
}
int radix_sort_pairs(uint32_t* kA, uint32_t* vA, uint32_t* kB, uint32_t* vB, int n, int nseg, int bits, int* counters,
                     cudaStream_t st, uint32_t** k_out, uint32_t** v_out) {
  const int nchunks = (n + kChunk - 1) / kChunk;
  const dim3 grid((nchunks + kWarpsPerBlock - 1) / kWarpsPerBlock, nseg);
  uint32_t *kin = kA, *vin = vA, *kout = kB, *vout = vB;
  int* slice_tot = counters + (size_t)nseg * nchunks * kBins;
  for (int shift = 0; shift < bits; shift += kRadixBits) {
    radix_hist_kernel<<<grid, 32 * kWarpsPerBlock, 0, st>>>(kin, n, nchunks, shift, counters);
    radix_scan1_kernel<<<dim3(kScanSlices, nseg), kBins, 0, st>>>(counters, nchunks, slice_tot);
    radix_scan2_kernel<<<nseg, kBins, 0, st>>>(slice_tot);
    radix_scan3_kernel<<<dim3(kScanSlices, nseg), kBins, 0, st>>>(counters, nchunks, slice_tot);
    radix_scatter_kernel<<<grid, 32 * kWarpsPerBlock, 0, st>>>(kin, vin, n, nchunks, shift, counters, kout, vout);
    uint32_t* t = kin; kin = kout; kout = t;
    t = vin; vin = vout; vout = t;
  }
  *k_out = kin;
  *v_out = vin;
  return cudaGetLastError() == cudaSuccess ? 0 : -1;
}

// ------------------------------------------------------------------------------------------
// Lovasz gradient over the sorted order: chunk foreground counts, scan, apply
// ------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(32 * kWarpsPerBlock) lovasz_count_kernel(const uint32_t* __restrict__ vals,
                                                                          const int* __restrict__ labels, int V,
                                                                          int nchunks, int* __restrict__ lcnt) {
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int chunk = blockIdx.x * kWarpsPerBlock + warp;
  const int c = blockIdx.y;
  if (chunk >= nchunks) return;
  const uint32_t* vv = vals + (size_t)c * V;
  const int beg = chunk * kChunk, end = min(V, beg + kChunk);
  int n = 0;
  for (int i = beg + lane; i < end; i += 32) n += (labels[vv[i]] == c);
#pragma unroll
  for (int o = 16; o >= 1; o >>= 1) n += __shfl_xor_sync(0xffffffffu, n, o);
  if (lane == 0) lcnt[(size_t)c * nchunks + chunk] = n;
}

__global__ void lovasz_scan_kernel(int* __restrict__ lcnt, int nchunks) {   // <<<C, 32>>>
  int* row = lcnt + (size_t)blockIdx.x * nchunks;
  const int lane = threadIdx.x;
  int carry = 0;
  for (int i0 = 0; i0 < nchunks; i0 += 32) {
    const int i = i0 + lane;
    const int x = i < nchunks ? row[i] : 0;
    int incl = x;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      const int t = __shfl_up_sync(0xffffffffu, incl, o);
      if (lane >= o) incl += t;
    }
    if (i < nchunks) row[i] = carry + incl - x;
    carry += __shfl_sync(0xffffffffu, incl, 31);
  }
}

// jaccard(i) with cf = inclusive foreground count up to i (lovasz_softmax.py:20-34, fp32 like torch)
__device__ __forceinline__ float jaccard_at(float gts, int i, float cf) {
  const float inter = gts - cf;
  const float uni = gts + ((float)(i + 1) - cf);
  return 1.f - inter / uni;
}

__global__ void __launch_bounds__(32 * kWarpsPerBlock) lovasz_apply_kernel(const uint32_t* __restrict__ keys,
                                                                          const uint32_t* __restrict__ vals,
                                                                          const int* __restrict__ labels, int V,
                                                                          int nchunks, const int* __restrict__ lcnt,
                                                                          OccAcc* __restrict__ acc,
                                                                          float* __restrict__ G) {
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int chunk = blockIdx.x * kWarpsPerBlock + warp;
  const int c = blockIdx.y;
  if (chunk >= nchunks) return;
  const float gts = (float)acc->nt[c];
  const uint32_t* kk = keys + (size_t)c * V;
  const uint32_t* vv = vals + (size_t)c * V;
  float* g = G + (size_t)c * V;
  const int beg = chunk * kChunk, end = min(V, beg + kChunk);
  if (gts <= 0.f) {                       // class not present: no loss term, zero gradient
    for (int i = beg + lane; i < end; i += 32) g[vv[i]] = 0.f;
    return;
  }
  int carry = lcnt[(size_t)c * nchunks + chunk];
  double dot = 0.0;
  for (int i0 = beg; i0 < end; i0 += 32) {
    const int i = i0 + lane;
    const bool act = i < end;
    uint32_t v = 0;
    int fg = 0;
    float e = 0.f;
    if (act) {
      v = vv[i];
      const int t = labels[v];
      fg = (t == c);
      e = __uint_as_float(kOneBits - kk[i]);
    }
    int incl = fg;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      const int t = __shfl_up_sync(0xffffffffu, incl, o);
      if (lane >= o) incl += t;
    }
    const int cf = carry + incl;
    if (act) {
      float gi = jaccard_at(gts, i, (float)cf);
      if (i > 0) gi -= jaccard_at(gts, i - 1, (float)(cf - fg));
      dot += (double)e * (double)gi;
      // d e / d p = -1 on foreground (e = 1 - p), +1 otherwise (e = p); ignored voxels carry e = 0, fg = 0
      g[v] = fg ? -gi : gi;
    }
    carry += __shfl_sync(0xffffffffu, incl, 31);
  }
#pragma unroll
  for (int o = 16; o >= 1; o >>= 1) dot += __shfl_xor_sync(0xffffffffu, dot, o);
  if (lane == 0) atomicAdd(&acc->lov[c], dot);
}

// ------------------------------------------------------------------------------------------
// scalars: the four losses and the coefficients of the backward pass
// ------------------------------------------------------------------------------------------
__device__ __forceinline__ double bce1(double x) { return -fmax(log(x), -100.0); }      // F.binary_cross_entropy(x, 1)
__device__ __forceinline__ double dbce1(double x) { return log(x) > -100.0 ? -1.0 / x : 0.0; }

__global__ void occ_finalize_kernel(OccAcc* __restrict__ acc, OccCoef* __restrict__ coef, int C, int empty_idx,
                                    float* __restrict__ losses) {
  if (threadIdx.x != 0 || blockIdx.x != 0) return;
  const double nv = acc->nvalid;
  // CE (semkitti.py:139-149): weighted mean over the valid voxels
  losses[0] = (float)(acc->ce_den > 0 ? acc->ce_num / acc->ce_den : 0.0 / 0.0);
  coef->ce_scale = (float)(acc->ce_den > 0 ? 1.0 / acc->ce_den : 0.0);
  // sem_scal (semkitti.py:92-136)
  double loss = 0.0, count = 0.0;
  for (int c = 0; c < C; ++c) count += acc->nt[c] > 0 ? 1.0 : 0.0;
  for (int c = 0; c < C; ++c) {
    double a = 0.0, b = 0.0;
    const double nt = acc->nt[c], sp = acc->sp[c], nom = acc->spt[c];
    if (nt > 0) {
      if (sp > 0) {                               // precision = nom / sum p
        const double pr = nom / sp;
        loss += bce1(pr);
        a += dbce1(pr) / sp;
        b += dbce1(pr) * (-nom / (sp * sp));
      }
      {                                           // recall = nom / n_t
        const double rc = nom / nt;
        loss += bce1(rc);
        a += dbce1(rc) / nt;
      }
      const double nneg = nv - nt;
      if (nneg > 0) {                             // specificity = sum (1-p)(1-[t=c]) / sum (1-[t=c])
        const double num = nneg - (sp - nom);
        const double spc = num / nneg;
        loss += bce1(spc);
        a += dbce1(spc) / nneg;                   // d num / d nom = +1
        b += dbce1(spc) * (-1.0 / nneg);          // d num / d sp  = -1
      }
    }
    coef->sem_a[c] = count > 0 ? (float)(a / count) : 0.f;
    coef->sem_b[c] = count > 0 ? (float)(b / count) : 0.f;
  }
  losses[1] = (float)(loss / count);
  // geo_scal (semkitti.py:62-89), eps = 1e-5
  {
    const double eps = 1e-5;
    const int e = empty_idx;
    const double sp0 = acc->sp[e], nom0 = acc->spt[e], nt0 = acc->nt[e];
    const double inter = (nv - nt0) - (sp0 - nom0);          // sum_{t != empty} (1 - p_empty)
    const double d1 = (nv - sp0) + eps, d2 = (nv - nt0) + eps, d3 = nt0 + eps;
    const double pr = inter / d1, rc = inter / d2, spc = nom0 / d3;
    losses[2] = (float)(bce1(pr) + bce1(rc) + bce1(spc));
    // d inter / d nom0 = +1, d inter / d sp0 = -1, d d1 / d sp0 = -1
    coef->geo_a = (float)(dbce1(pr) / d1 + dbce1(rc) / d2 + dbce1(spc) / d3);
    coef->geo_b = (float)(dbce1(pr) * (-1.0 / d1 + inter / (d1 * d1)) + dbce1(rc) * (-1.0 / d2));
  }
  // Lovasz: mean over the present classes (lovasz_softmax.py:176-203)
  {
    double s = 0.0, n = 0.0;
    for (int c = 0; c < C; ++c)
      if (acc->nt[c] > 0) { s += acc->lov[c]; n += 1.0; }
    losses[3] = (float)(n > 0 ? s / n : 0.0);
    coef->lov_scale = (float)(n > 0 ? 1.0 / n : 0.0);
  }
}

// ------------------------------------------------------------------------------------------
// backward
// ------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) occ_bwd_kernel(const float* __restrict__ logits, long long ld,
                                                      const int* __restrict__ labels, int V, int C,
                                                      const float* __restrict__ class_w, int ignore, int empty_idx,
                                                      const OccCoef* __restrict__ coef, const float* __restrict__ G,
                                                      const float* __restrict__ gl, float* __restrict__ dlogits,
                                                      long long ldd) {
  __shared__ OccCoef sc;
  if (threadIdx.x == 0) sc = *coef;
  __syncthreads();
  const float g_ce = gl[0], g_sem = gl[1], g_geo = gl[2], g_lov = gl[3] * sc.lov_scale;
  for (int v = blockIdx.x * blockDim.x + threadIdx.x; v < V; v += gridDim.x * blockDim.x) {
    const int t = labels[v];
    float* out = dlogits + (long long)v * ldd;
    if (t == ignore) {
      for (int c = 0; c < C; ++c) out[c] = 0.f;
      continue;
    }
    const float* row = logits + (long long)v * ld;
    float p[kMaxCls], dp[kMaxCls];
    float m = -INFINITY;
#pragma unroll
    for (int c = 0; c < kMaxCls; ++c)
      if (c < C) { p[c] = row[c]; m = fmaxf(m, p[c]); }
    float s = 0.f;
#pragma unroll
    for (int c = 0; c < kMaxCls; ++c)
      if (c < C) { p[c] = expf(p[c] - m); s += p[c]; }
    const float inv = 1.f / s;
    float dotp = 0.f;
#pragma unroll
    for (int c = 0; c < kMaxCls; ++c) {
      if (c < C) {
        p[c] *= inv;
        float d = g_sem * (sc.sem_b[c] + (t == c ? sc.sem_a[c] : 0.f)) + g_lov * G[(size_t)c * V + v];
        if (c == empty_idx) d += g_geo * (sc.geo_b + (t == c ? sc.geo_a : 0.f));
        dp[c] = d;
        dotp += p[c] * d;
      }
    }
    const float wce = g_ce * sc.ce_scale * (class_w ? class_w[t] : 1.f);
#pragma unroll
    for (int c = 0; c < kMaxCls; ++c)
      if (c < C) out[c] = p[c] * (dp[c] - dotp) + wce * (p[c] - (t == c ? 1.f : 0.f));
  }
}

}  // namespace coocc

using namespace coocc;

extern "C" long long coocc_occ_loss_workspace(int V, int C) {
  if (V < 1 || C < 1 || C > kMaxCls) return -1;
  return (long long)ws_layout(nullptr, V, C, nullptr);
}

extern "C" int coocc_occ_label_mode(const void* labels, int label_bytes, int X, int Y, int Z, int ratio,
                                    int empty_idx, int* out, void* stream) {
  if (!labels || !out || X < 1 || Y < 1 || Z < 1 || ratio < 1) return COOCC_ERR_ARG;
  if (ratio * ratio * ratio > 64) return COOCC_ERR_CAPACITY;
  const int V = X * Y * Z;
  cudaStream_t st = (cudaStream_t)stream;
  const int blocks = (V + 127) / 128;
  if (label_bytes == 8)
    label_mode_kernel<long long><<<blocks, 128, 0, st>>>((const long long*)labels, X, Y, Z, ratio, empty_idx, out);
  else if (label_bytes == 4)
    label_mode_kernel<int><<<blocks, 128, 0, st>>>((const int*)labels, X, Y, Z, ratio, empty_idx, out);
  else if (label_bytes == 1)
    label_mode_kernel<unsigned char><<<blocks, 128, 0, st>>>((const unsigned char*)labels, X, Y, Z, ratio, empty_idx, out);
  else
    return COOCC_ERR_ARG;
  return cudaGetLastError() == cudaSuccess ? 0 : COOCC_ERR_CUDA;
}

extern "C" int coocc_occ_loss_fwd(const float* logits, long long ld, const int* labels, int V, int C,
                                  const float* class_w, int ignore, int empty_idx, void* workspace,
                                  float* losses4, void* stream) {
  if (!logits || !labels || !workspace || !losses4 || V < 1 || C < 2 || C > kMaxCls || ld < C || empty_idx < 0 ||
      empty_idx >= C)
    return COOCC_ERR_ARG;
  cudaStream_t st = (cudaStream_t)stream;
  OccWs w;
  ws_layout(reinterpret_cast<char*>(workspace), V, C, &w);
  if (cudaMemsetAsync(w.acc, 0, sizeof(OccAcc), st) != cudaSuccess) return COOCC_ERR_CUDA;
  int blocks = (V + 255) / 256;
  if (blocks > 148 * 8) blocks = 148 * 8;
  occ_stats_kernel<<<blocks, 256, 0, st>>>(logits, ld, labels, V, C, class_w, ignore, w.acc, w.keyA, w.valA);
  // 3 x 10-bit LSD passes over the 30 significant key bits
  const dim3 grid((w.nchunks + kWarpsPerBlock - 1) / kWarpsPerBlock, C);
  uint32_t *kin = nullptr, *vin = nullptr;
  if (radix_sort_pairs(w.keyA, w.valA, w.keyB, w.valB, V, C, 30, w.counters, st, &kin, &vin) != 0) return COOCC_ERR_CUDA;
  lovasz_count_kernel<<<grid, 32 * kWarpsPerBlock, 0, st>>>(vin, labels, V, w.nchunks, w.lcnt);
  lovasz_scan_kernel<<<C, 32, 0, st>>>(w.lcnt, w.nchunks);
  lovasz_apply_kernel<<<grid, 32 * kWarpsPerBlock, 0, st>>>(kin, vin, labels, V, w.nchunks, w.lcnt, w.acc, w.G);
  occ_finalize_kernel<<<1, 32, 0, st>>>(w.acc, w.coef, C, empty_idx, losses4);
  return cudaGetLastError() == cudaSuccess ? 0 : COOCC_ERR_CUDA;
}

extern "C" int coocc_occ_loss_bwd(const float* logits, long long ld, const int* labels, int V, int C,
                                  const float* class_w, int ignore, int empty_idx, const void* workspace,
                                  const float* g_losses4, float* dlogits, long long ldd, void* stream) {
  if (!logits || !labels || !workspace || !g_losses4 || !dlogits || V < 1 || C < 2 || C > kMaxCls || ld < C || ldd < C)
    return COOCC_ERR_ARG;
  OccWs w;
  ws_layout(reinterpret_cast<char*>(const_cast<void*>(workspace)), V, C, &w);
  int blocks = (V + 255) / 256;
  if (blocks > 148 * 8) blocks = 148 * 8;
  occ_bwd_kernel<<<blocks, 256, 0, (cudaStream_t)stream>>>(logits, ld, labels, V, C, class_w, ignore, empty_idx, w.coef,
                                                          w.G, g_losses4, dlogits, ldd);
  return cudaGetLastError() == cudaSuccess ? 0 : COOCC_ERR_CUDA;
}
