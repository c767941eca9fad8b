// conv_tc.cu -- 3D convolution (1x1x1 / 3x3x3, stride 1 / 2) and plain GEMM on the B200 5th-gen
// tensor cores: TMA (im2col mode) -> shared memory -> tcgen05.mma -> TMEM -> epilogue.
//
// Replaces, for the fused-voxel hot path, the cuDNN Conv3d / cuBLAS Linear calls of the reference
// (P/coocc/fuser/bifuser_n.py:23-30, P/coocc/backbones/resnet3d.py:16-31, P/coocc/necks/fpn3d.py:48-67,
//  P/coocc/dense_heads/occ_head.py:102-132, P/utils/nerf_mlp.py:92-105).
//
// Data layout: activations NDHWC (= torch channels_last_3d of [1,C,X,Y,Z]) i.e. a row-major
// [V = X*Y*Z, C] matrix with row stride `ld`; weights [Cout][kx][ky][kz][Cin] (= channels_last_3d of
// the reference's [Cout,Cin,3,3,3] parameter) i.e. row-major [Cout, taps*Cin].
//
// One kernel template, three uses:
//   fprop : D[v, co] = sum_{tap,ci} X[v*s + tap - pad, ci] * W[co, tap, ci]
//           A = X, K-major, TMA im2col;  B = W rows, K-major, TMA tiled.
//   dgrad : dX[v, ci] = sum_{tap,co} dY[v - tap + pad, co] * W[co, tap, ci]      (stride 1)
//           A = dY, K-major, TMA im2col with mirrored taps;  B = W read *in place* as an
//           MN-major operand (ci contiguous), so no transposed weight copy is ever made.
//   wgrad : dW[co, tap, ci] = sum_v dY[v, co] * X[v*s + tap - pad, ci]
//           A = dY MN-major (TMA tiled), B = X MN-major (TMA im2col), split-K over voxel blocks,
//           fp32 red.global.add epilogue.
//
// CTA = 384 threads: warp 0 TMA producer, warp 1 MMA issuer, warp 2 TMEM allocator, warps 4-11
// epilogue (two per TMEM lane quadrant; a single warp per scheduler was issue-bound: the 1x1x1 layers spent 3.7 us
// per tile in the epilogue against 0.2 us of MMAs, profiles/r02_conv_k1_ncu.txt).  Persistent grid (<= #SMs), 4-stage smem ring, 2 TMEM accumulator buffers so the
// epilogue of tile i overlaps the MMAs of tile i+1.
#include <cuda.h>
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include <cuda_bf16.h>

#include "../../include/coocc_b200.h"
#include "tc_common.cuh"

namespace coocc {

constexpr int kStages = 4;
constexpr int kBM = 128;              // MMA M (TMEM lanes)
constexpr int kMaxBN = 256;           // MMA N upper bound (TMEM columns per accumulator buffer)
constexpr int kABytes = kBM * 128;    // 16 KiB per stage
constexpr int kBBytesMax = kMaxBN * 128;
constexpr int kStageBytes = kABytes + kBBytesMax;
constexpr int kRingBytes = 216 * 1024;                  // operand ring(s), cut into `nstages` stages (every ring layout
                                                        // in use fits: ky 3 x 40 + 6 x 16 KiB, ky wgrad 3 x 72 KiB)
constexpr int kMaxBRing = 8;          // ky-fused path: weight-tile ring depth bound
constexpr int kMaxARing = 4;          // ky-fused path: extended activation-tile ring depth bound
constexpr int kMaxAcc = 4;            // TMEM accumulator buffers (512 columns / acc_stride)
constexpr int kSched = 4;             // dynamic tile scheduler: ring of published tile indices
constexpr int kStatAccBytes = 8192;   // per-CTA running BatchNorm sums: [epilogue warp][column chunk][lane] float2
constexpr int kSmemBytes = kRingBytes + 1024 /*align*/ + 512 /*barriers*/ + kStatAccBytes;
static_assert(kSmemBytes <= 232448, "shared memory");
static_assert(kStages * kStageBytes <= kRingBytes, "ring");
constexpr int kEpiWarps = 8;          // two epilogue warps per TMEM lane quadrant (they split the 32-column chunks)
constexpr int kThreads = 128 + 32 * kEpiWarps;
constexpr int kStatRow = 36;          // words per row of the statistics staging tile (conflict-free 16 B stores / column reads)
constexpr int kStatBytes = kEpiWarps * 32 * kStatRow * 4;

enum { MODE_FPROP = 0, MODE_DGRAD = 1, MODE_WGRAD = 2 };

struct TcParams {
  CUtensorMap tmA;
  CUtensorMap tmB;
  int mode;
  int M, N;            // output rows / cols (fprop: V_out, Cout; dgrad: V, Cin; wgrad: Cout, Cin)
  int BN;              // N tile (multiple of 16, <= 256)
  int MT;              // 128-row MMA tiles per CTA tile (1, 2 or 4; MT*BN <= 512 TMEM columns)
  int NT;              // wgrad: filter taps per CTA tile sharing one dY tile (1 or 3)
  int nstages;         // smem ring depth (4, or 3 when a stage is 64 KiB)
  int nacc;            // TMEM accumulator buffers (2, or 1 when MT*BN > 256: epilogue not overlapped; 4 for the narrow
                       // 1x1x1 tiles, whose per-tile load -> MMA -> epilogue chain is latency-bound)
  int acc_stride;      // TMEM columns between accumulator buffers
  int hstride;         // TMEM column stride between the MT halves
  int es;              // operand element size in bytes (host bookkeeping)
  int mc;              // 1: CTA pairs (cluster of 2) on adjacent M tiles share the B tile by TMA multicast
  int taps;            // 1 or 27
  int Kc;              // reduction channels per tap (fprop: Cin, dgrad: Cout); wgrad: unused
  int a_im2col;        // A (fprop/dgrad) or B (wgrad) loaded through im2col
  int oX, oY, oZ;      // spatial dims of the *pixel* index space of the im2col operand
  int cstride;         // conv stride (im2col traversal stride)
  int lo;              // im2col lower corner (= -pad)
  int wK;              // row length of the weight matrix in elements (taps * Cin)  [dgrad B coords]
  int Cin;             // weight inner channel count (B K-offset per tap)
  // epilogue
  float* out;
  long long ldc;
  const float* bias;
  int relu;
  int accum;          // epilogue adds the previous contents of out (multi-pass precision split)
  int out_bf16;       // fprop / dgrad: `out` is a bf16 matrix (row stride ldc in bf16 elements)
  float* stats;        // [2][N] sum / sum of squares over rows, or nullptr
  // wgrad
  int nvb;             // number of voxel (K) blocks
  int ksplit;
  // ky-fused 3x3x3 stride-1 path (fprop / dgrad): one y-extended activation tile feeds the three
  // ky taps of a (kx, kz) pair, see the comment above tc_conv_kernel
  int ky;              // 1: enabled
  int gX, gYZ;         // grid: x extent, Y*Z
  int gZ;              // z extent (rows per y step)
  int ny, nyb;         // y rows per 128-pixel block (128 / Z), y blocks per x (ceil(Y / ny))
  int ext_bytes;       // bytes of one extended tile ((ny + 2) * Z rows of 128 B, rounded to 1 KiB)
  int NA, NB;          // ring depths: extended A tiles (MT per entry), B tiles
  // parity-class launch of a stride-2 data gradient (dgrad_impl, `cls`): an explicit tap list -- im2col offsets and
  // weight taps are no longer tied to each other -- and output rows scattered to the fine grid
  int* sched;                        // dynamic tile scheduler: {next tile, finished CTAs} in global memory, or nullptr
  const void* addend;                // lean bf16 epilogue: out = acc + addend (bf16 rows, the gradient arriving through
  long long ld_add;                  //   a skip connection; saves the separate add pass over the tensor)
  int b_kmajor;                      // dgrad: B is a transposed weight copy [Cin][taps][Cout] (K-major, like fprop)
  int epi;                           // epilogue warps: 4 (one per TMEM lane quadrant) or 8 (two, splitting the column chunks)
  int stat_off;                      // byte offset of the statistics staging tiles behind the operand ring (0: none)
  int tapmode;                       // 1: tap t reads im2col offset tap_off[t] and weight tap tap_w[t]
  unsigned char tap_off[8];          // kx | ky << 2 | kz << 4   (offsets 0 / 1)
  unsigned char tap_w[8];            // (kx * 3 + ky) * 3 + kz of the weight tensor
  // all eight parity classes in ONE launch (tile = (class, mt, nt), class-major, classes ordered by tap count so the
  // long tiles are dealt first): eight separate launches left the small grids on 4-20 SMs each
  int ncls;                          // 0: single class (tap_off / tap_w / opx.. above), 8: class table below
  unsigned char cls_ntaps[8];
  unsigned char cls_par[8];          // px << 2 | py << 1 | pz
  unsigned char cls_off[8][8];
  unsigned char cls_w[8][8];
  int omap;                          // 1: output row (i, j, k) of the coarse grid -> fine voxel (2i+opx, 2j+opy, 2k+opz)
  int opx, opy, opz, fX, fY, fZ;
};

template <int ES> struct Elt {
  static constexpr int BKE = 128 / ES;       // K elements per stage for K-major operands
  static constexpr int CH = 128 / ES;        // MN elements per 128B chunk for MN-major operands
  static constexpr int KR = 128 / ES;        // K rows per stage for MN-major operands
  static constexpr int UK = 32 / ES;         // K elements per tcgen05.mma
  static constexpr int NUK = BKE / UK;       // MMAs per stage (= 4)
  static constexpr bool TF32 = (ES == 4);
  static constexpr uint32_t FMT = (ES == 4) ? 2u : 1u;                // TF32 : BF16
  static constexpr uint32_t MN_LAYOUT = (ES == 4) ? 1u : 2u;          // SW128_BASE32B : SW128
  static constexpr uint32_t MN_SBO = (ES == 4) ? 512u : 1024u;        // K-atom stride (4 / 8 rows)
  static constexpr uint32_t MN_CHUNK = KR * 128;                      // bytes per MN chunk
  static constexpr uint32_t MN_KSTEP = UK * 128;                      // bytes per MMA K step
};

__device__ __forceinline__ void pixel_to_whd(const TcParams& p, int pix, int& w, int& h, int& d) {
  const int z = pix % p.oZ;
  const int t = pix / p.oZ;
  const int y = t % p.oY;
  const int x = t / p.oY;
  w = p.lo + z * p.cstride;
  h = p.lo + y * p.cstride;
  d = p.lo + x * p.cstride;
}

// Explicit shared-state-space accesses (32-bit addresses): the kernel's smem base goes through an integer alignment
// round trip, so plain pointer dereferences compile to *generic* LD/ST -- measured as the top long-scoreboard stalls of
// the 1x1x1 epilogue (profiles/r02_conv_k1_ncu.txt).
__device__ __forceinline__ float lds_f32(uint32_t a) {
  float f;
  asm volatile("ld.shared.f32 %0, [%1];" : "=f"(f) : "r"(a));
  return f;
}
__device__ __forceinline__ float2 lds_f32x2(uint32_t a) {
  float2 f;
  asm volatile("ld.shared.v2.f32 {%0, %1}, [%2];" : "=f"(f.x), "=f"(f.y) : "r"(a));
  return f;
}
__device__ __forceinline__ void sts_f32x2(uint32_t a, float x, float y) {
  asm volatile("st.shared.v2.f32 [%0], {%1, %2};" ::"r"(a), "f"(x), "f"(y) : "memory");
}
__device__ __forceinline__ uint4 lds_u32x4(uint32_t a) {
  uint4 v;
  asm volatile("ld.shared.v4.b32 {%0, %1, %2, %3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "r"(a));
  return v;
}
__device__ __forceinline__ void sts_u32x4(uint32_t a, uint32_t x, uint32_t y, uint32_t z, uint32_t w) {
  asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(a), "r"(x), "r"(y), "r"(z), "r"(w) : "memory");
}

// a + b on 8 packed bf16 values, evaluated in fp32 and rounded once (what torch's bf16 add does)
__device__ __forceinline__ uint32_t bf16x2_add(uint32_t a, uint32_t b) {
  const float lo = __uint_as_float(a << 16) + __uint_as_float(b << 16);
  const float hi = __uint_as_float(a & 0xFFFF0000u) + __uint_as_float(b & 0xFFFF0000u);
  const __nv_bfloat162 h = __floats2bfloat162_rn(lo, hi);
  return *reinterpret_cast<const uint32_t*>(&h);
}
__device__ __forceinline__ uint4 bf16x8_add(uint4 a, uint4 b) {
  return make_uint4(bf16x2_add(a.x, b.x), bf16x2_add(a.y, b.y), bf16x2_add(a.z, b.z), bf16x2_add(a.w, b.w));
}

// Epilogue helper: per-column sum / sum of squares of a 32 x 32 accumulator chunk (row = lane) over the valid rows,
// added to the warp's running cell (shared address `cell`) of column c0 + lane.
//   tile != 0: shared-memory transpose -- every lane stores its row (8 x 16 B, row stride 36 words: conflict-free),
//     then sums one column; ~110 instructions
//   else: register butterfly transpose-reduce; ~310 instructions (when the operand ring leaves no room for the tiles)
__device__ __forceinline__ void epi_stats(const uint32_t (&v)[32], bool row_ok, uint32_t tile, uint32_t cell, int lane) {
  float st, qt;
  if (tile != 0u) {
    const int nvalid = __popc(__ballot_sync(0xffffffffu, row_ok));    // valid rows are a prefix of the warp
    const uint32_t dst = tile + (uint32_t)(lane * kStatRow * 4);
#pragma unroll
    for (int i = 0; i < 8; ++i) sts_u32x4(dst + 16u * i, v[4 * i], v[4 * i + 1], v[4 * i + 2], v[4 * i + 3]);
    __syncwarp();
    float s0 = 0.f, q0 = 0.f, s1 = 0.f, q1 = 0.f;
    const uint32_t src = tile + (uint32_t)(lane * 4);
    if (nvalid == 32) {
#pragma unroll
      for (int r = 0; r < 32; r += 2) {
        const float f = lds_f32(src + (uint32_t)(r * kStatRow * 4)), g = lds_f32(src + (uint32_t)((r + 1) * kStatRow * 4));
        s0 += f; q0 = fmaf(f, f, q0);
        s1 += g; q1 = fmaf(g, g, q1);
      }
    } else {
      for (int r = 0; r < nvalid; ++r) {
        const float f = lds_f32(src + (uint32_t)(r * kStatRow * 4));
        s0 += f; q0 = fmaf(f, f, q0);
      }
    }
    __syncwarp();
    st = s0 + s1; qt = q0 + q1;
  } else {
    float s[32], ss[32];
#pragma unroll
    for (int i = 0; i < 32; ++i) {
      const float f = row_ok ? __uint_as_float(v[i]) : 0.f;
      s[i] = f;
      ss[i] = f * f;
    }
#pragma unroll
    for (int off = 16; off >= 1; off >>= 1) {
      const bool up = (lane & off) != 0;
#pragma unroll
      for (int i = 0; i < off; ++i) {
        const float send_s = up ? s[i] : s[i + off];
        const float send_q = up ? ss[i] : ss[i + off];
        const float recv_s = __shfl_xor_sync(0xffffffffu, send_s, off);
        const float recv_q = __shfl_xor_sync(0xffffffffu, send_q, off);
        s[i] = (up ? s[i + off] : s[i]) + recv_s;
        ss[i] = (up ? ss[i + off] : ss[i]) + recv_q;
      }
    }
    st = s[0]; qt = ss[0];
  }
  const float2 a = lds_f32x2(cell);
  sts_f32x2(cell, a.x + st, a.y + qt);
}

template <int ES>
__global__ void __launch_bounds__(kThreads, 1) tc_conv_kernel(const __grid_constant__ TcParams p) {
  using E = Elt<ES>;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) &
                                             ~static_cast<uintptr_t>(1023));
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(smem + kRingBytes);   // operand ring (ky path: B ring)
  uint64_t* empty_bar = full_bar + kMaxBRing;
  uint64_t* fulla_bar = empty_bar + kMaxBRing;  // ky path: extended A tiles
  uint64_t* emptya_bar = fulla_bar + kMaxARing;
  uint64_t* tfull_bar = emptya_bar + kMaxARing; // [kMaxAcc] accumulator ready
  uint64_t* tempty_bar = tfull_bar + kMaxAcc;   // [kMaxAcc] accumulator drained
  uint64_t* sfull_bar = tempty_bar + kMaxAcc;   // [kSched] scheduled tile published
  uint64_t* sempty_bar = sfull_bar + kSched;    // [kSched] ... and read by every consumer
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(sempty_bar + kSched);
  int* sched_tile = reinterpret_cast<int*>(tmem_slot + 1);      // [kSched]

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const bool wgrad = (p.mode == MODE_WGRAD);
  const bool a_mn = wgrad;
  const bool b_mn = (p.mode != MODE_FPROP) && !p.b_kmajor;

  // ---- tile bookkeeping (identical in every role) ---------------------------------------
  const int ntn = (p.N + p.BN - 1) / p.BN;
  // ky path: M is cut into 128-pixel blocks (x, y-block) that never straddle an x row
  const int nmb = (p.ky && !wgrad) ? p.gX * p.nyb : (p.M + kBM - 1) / kBM;
  const int ntm = (nmb + p.MT - 1) / p.MT;
  // with multicast a "tile" of the scheduler is a PAIR of M tiles (one per CTA of the cluster)
  const int crank = p.mc ? (int)cluster_ctarank() : 0;
  const int cidx = p.mc ? (int)(blockIdx.x >> 1) : (int)blockIdx.x;
  const int cnum = p.mc ? (int)(gridDim.x >> 1) : (int)gridDim.x;
  const int ntm_sched = p.mc ? (ntm + 1) / 2 : ntm;
  const int ntg = (p.taps + p.NT - 1) / p.NT;       // wgrad: tap groups
  const int ntiles = wgrad ? ntm * ntn * ntg * p.ksplit : ntm_sched * ntn * p.ksplit * (p.ncls ? p.ncls : 1);
  const int kb_per_tap = (p.Kc + E::BKE - 1) / E::BKE;

  if (threadIdx.x == 0) {
    tma_prefetch_desc(&p.tmA);
    tma_prefetch_desc(&p.tmB);
    for (int s = 0; s < kMaxBRing; ++s) {
      mbar_init(&full_bar[s], 1);
      mbar_init(&empty_bar[s], p.mc ? 2 : 1);   // multicast: both CTAs' MMAs must release the stage
    }
    for (int s = 0; s < kMaxARing; ++s) {
      mbar_init(&fulla_bar[s], 1);
      mbar_init(&emptya_bar[s], 1);
    }
    for (int s = 0; s < kSched; ++s) {
      mbar_init(&sfull_bar[s], 1);
      mbar_init(&sempty_bar[s], 1u + (uint32_t)p.epi);     // the MMA thread + one arrive per epilogue warp
    }
    for (int s = 0; s < kMaxAcc; ++s) {
      mbar_init(&tfull_bar[s], 1);
      mbar_init(&tempty_bar[s], (uint32_t)p.epi);   // one arrive per epilogue warp
    }
    fence_barrier_init();
  }
  if (warp == 2) tmem_alloc(tmem_slot, 512);
  tc_fence_before();
  __syncthreads();
  if (p.mc) cluster_sync_all();       // peer barriers must exist before any multicast / remote arrive
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  // ---- tile scheduler --------------------------------------------------------------------
  // static (p.sched == nullptr): CTA c takes tiles c, c + grid, ...  Dynamic: tiles are handed out by an atomic counter,
  // so a CTA that gets its SM late (another kernel -- the FPS side branch of the step runner, NCCL -- holds it when
  // the grid launches) finds the remaining tiles or none at all instead of a static share that would run as a
  // second wave.  The producer thread fetches (one tile ahead, the atomic's latency is off the critical path) and
  // publishes the index through a small shared-memory ring; the MMA thread and the epilogue warps read it there.
  // The last CTA to see the end resets the counters for the next launch that uses them.
  const bool dyn = p.sched != nullptr;
  int sched_next = 0;                                       // producer thread only
  auto tile_produce = [&](int it_) -> int {
    if (!dyn) {
      const int t = cidx + it_ * cnum;
      return t < ntiles ? t : -1;
    }
    const int slot = it_ % kSched;
    mbar_wait(&sempty_bar[slot], (uint32_t)((it_ / kSched) & 1) ^ 1u);
    int t = (it_ == 0) ? atomicAdd(&p.sched[0], 1) : sched_next;
    if (t >= ntiles) t = -1;
    sched_tile[slot] = t;
    mbar_arrive(&sfull_bar[slot]);
    if (t >= 0) {
      sched_next = atomicAdd(&p.sched[0], 1);
    } else if (atomicAdd(&p.sched[1], 1) == (int)gridDim.x - 1) {
      p.sched[0] = 0;
      p.sched[1] = 0;
      __threadfence();
    }
    return t;
  };
  auto tile_consume = [&](int it_, bool whole_warp) -> int {
    if (!dyn) {
      const int t = cidx + it_ * cnum;
      return t < ntiles ? t : -1;
    }
    const int slot = it_ % kSched;
    mbar_wait(&sfull_bar[slot], (uint32_t)((it_ / kSched) & 1));
    const int t = *reinterpret_cast<volatile int*>(&sched_tile[slot]);
    if (whole_warp) {
      __syncwarp();
      if (lane == 0) mbar_arrive(&sempty_bar[slot]);
    } else {
      mbar_arrive(&sempty_bar[slot]);
    }
    return t;
  };

  const uint32_t a_bytes = kABytes * (uint32_t)p.MT;
  const uint32_t b_bytes = (uint32_t)p.BN * 128u;
  const uint32_t stage_bytes = a_bytes + (uint32_t)p.NT * ((b_bytes + 1023u) / 1024u * 1024u);
  const int nstages = p.nstages;

  if (warp == 0) {
    // =========================== TMA producer ===========================================
    if (p.ky && wgrad) {
      if (elect_one()) {
        // ---- ky-fused wgrad producer: K blocks are the 128-pixel (x, y-block) blocks; per block the
        //      dY^T tile (128 couts) and ONE y-extended X tile that serves the three ky taps of (kx, kz)
        const uint32_t a_chunk = 128u * 128u;
        const int nchA = kBM / E::CH, nchB = p.BN / E::CH;
        const uint32_t ext_tx = (uint32_t)((p.ny + 2) * p.gZ) * 128u;
        const uint32_t stage_b = (uint32_t)nchA * a_chunk + (uint32_t)nchB * (uint32_t)p.ext_bytes;
        int stage = 0;
        uint32_t phase = 0;
        for (int it_ = 0;; ++it_) {
          const int tile = tile_produce(it_);
          if (tile < 0) break;
          int t = tile;
          const int sp = t % p.ksplit; t /= p.ksplit;
          const int nt = t % ntn; t /= ntn;
          const int mt = t % ntm; t /= ntm;
          const int kx = t / 3, kz = t % 3;
          const int it0 = (int)((long long)p.nvb * sp / p.ksplit);
          const int it1 = (int)((long long)p.nvb * (sp + 1) / p.ksplit);
          for (int it = it0; it < it1; ++it) {
            const int x = it / p.nyb, yb = it - x * p.nyb;
            mbar_wait(&empty_bar[stage], phase ^ 1);
            uint8_t* sa = smem + stage * stage_b;
            uint8_t* sb = sa + nchA * a_chunk;
            mbar_expect_tx(&full_bar[stage], (uint32_t)nchA * a_chunk + (uint32_t)nchB * ext_tx);
            for (int j = 0; j < nchA; ++j)
              tma_load_3d(sa + j * a_chunk, &p.tmA, &full_bar[stage], mt * kBM + j * E::CH, yb * kBM, x);
            for (int j = 0; j < nchB; ++j)
              tma_load_im2col_5d(sb + j * p.ext_bytes, &p.tmB, &full_bar[stage], nt * p.BN + j * E::CH, p.lo,
                                 p.lo + yb * p.ny, p.lo + x, 0, (uint16_t)kz, (uint16_t)0, (uint16_t)kx);
            if (++stage == nstages) { stage = 0; phase ^= 1; }
          }
        }
      }
    } else if (p.ky) {
      if (elect_one()) {
        // ---- ky-fused producer: per (kx, kz, k-block) one extended A tile per 128-pixel block, then
        //      the three weight tiles of ky = 0..2
        const uint32_t a_entry = (uint32_t)p.MT * (uint32_t)p.ext_bytes;
        const uint32_t b_entry = (b_bytes + 1023u) / 1024u * 1024u;
        uint8_t* const ringB = smem + (uint32_t)p.NA * a_entry;
        const uint32_t ext_tx = (uint32_t)((p.ny + 2) * p.gZ) * 128u;
        int sa_i = 0, sb_i = 0;
        uint32_t pa = 0, pb = 0;
        const int ngroups = 9 * kb_per_tap;
        for (int it_ = 0;; ++it_) {
          const int tile = tile_produce(it_);
          if (tile < 0) break;
          const int nt = tile % ntn;
          const int mt = tile / ntn;
          const int nhalf = max(0, min(p.MT, nmb - mt * p.MT));
          for (int g = 0; g < ngroups; ++g) {
            const int kb = g % kb_per_tap;
            const int t2 = g / kb_per_tap;
            const int kz = t2 % 3, kx = t2 / 3;
            mbar_wait(&emptya_bar[sa_i], pa ^ 1);
            mbar_expect_tx(&fulla_bar[sa_i], (uint32_t)nhalf * ext_tx);
            for (int hh = 0; hh < nhalf; ++hh) {
              const int mb = mt * p.MT + hh;
              const int x = mb / p.nyb, yb = mb - x * p.nyb;
              tma_load_im2col_5d(smem + sa_i * a_entry + hh * p.ext_bytes, &p.tmA, &fulla_bar[sa_i], kb * E::BKE,
                                 p.lo, p.lo + yb * p.ny, p.lo + x, 0, (uint16_t)kz, (uint16_t)0, (uint16_t)kx);
            }
            if (++sa_i == p.NA) { sa_i = 0; pa ^= 1; }
            for (int ky = 0; ky < 3; ++ky) {
              const int tap = kx * 9 + ky * 3 + kz;
              const int wtap = (p.mode == MODE_DGRAD) ? (26 - tap) : tap;
              mbar_wait(&empty_bar[sb_i], pb ^ 1);
              mbar_expect_tx(&full_bar[sb_i], b_bytes);
              uint8_t* sb = ringB + sb_i * b_entry;
              if (!b_mn) {
                tma_load_2d(sb, &p.tmB, &full_bar[sb_i], wtap * p.Cin + kb * E::BKE, nt * p.BN);
              } else {
                const int nch = p.BN / E::CH;
                for (int j = 0; j < nch; ++j)
                  tma_load_2d(sb + j * E::MN_CHUNK, &p.tmB, &full_bar[sb_i], wtap * p.Cin + nt * p.BN + j * E::CH,
                              kb * E::KR);
              }
              if (++sb_i == p.NB) { sb_i = 0; pb ^= 1; }
            }
          }
        }
      }
    } else if (elect_one()) {
      int stage = 0;
      uint32_t phase = 0;
      for (int it_ = 0;; ++it_) {
          const int tile = tile_produce(it_);
          if (tile < 0) break;
        int mt, nt, it0, it1;          // it = flattened (tap, k-block) iteration
        int wg_tap = 0, cls = 0;
        if (!wgrad) {
          int t = tile;
          const int sp = t % p.ksplit; t /= p.ksplit;
          nt = t % ntn;
          mt = p.mc ? 2 * (t / ntn) + crank : t / ntn;
          if (p.ncls) { cls = mt / ntm; mt -= cls * ntm; }
          const int nk_total = (p.ncls ? (int)p.cls_ntaps[cls] : p.taps) * kb_per_tap;
          it0 = (int)((long long)nk_total * sp / p.ksplit);
          it1 = (int)((long long)nk_total * (sp + 1) / p.ksplit);
        } else {
          int t = tile;
          const int sp = t % p.ksplit; t /= p.ksplit;
          nt = t % ntn; t /= ntn;
          mt = t % ntm; t /= ntm;
          wg_tap = t;
          it0 = (int)((long long)p.nvb * sp / p.ksplit);
          it1 = (int)((long long)p.nvb * (sp + 1) / p.ksplit);
        }
        for (int it = it0; it < it1; ++it) {
          const int tap = wgrad ? wg_tap * p.NT : it / kb_per_tap;    // wgrad: first tap of the group
          const int kb = wgrad ? it : it % kb_per_tap;
          // filter offsets of this tap in the im2col operand (W<->z, H<->y, D<->x)
          int kx = (p.taps == 1) ? 0 : tap / 9, ky = (p.taps == 1) ? 0 : (tap / 3) % 3,
              kz = (p.taps == 1) ? 0 : tap % 3;
          // dgrad walks the mirrored tap of the weight tensor
          int wtap = (p.mode == MODE_DGRAD) ? (p.taps - 1 - tap) : tap;
          if (p.tapmode) {
            const int o = p.ncls ? p.cls_off[cls][tap] : p.tap_off[tap];
            kx = o & 3; ky = (o >> 2) & 3; kz = (o >> 4) & 3;
            wtap = p.ncls ? p.cls_w[cls][tap] : p.tap_w[tap];
          }
          {
            mbar_wait(&empty_bar[stage], phase ^ 1);
            uint8_t* sa = smem + stage * stage_bytes;
            uint8_t* sb = sa + a_bytes;
            // a 128-row half that starts past the last pixel is not loaded at all (an im2col load whose
            // base pixel lies outside the tensor faults); its rows are masked in the epilogue
            const int nhalf = max(0, min(p.MT, (p.M - mt * p.MT * kBM + kBM - 1) / kBM));
            const int ntap = wgrad ? min(p.NT, p.taps - tap) : 1;
            mbar_expect_tx(&full_bar[stage], (uint32_t)nhalf * kABytes + (uint32_t)ntap * b_bytes);
            if (!wgrad) {
              // ---- A: MT x [128 pixels x BKE channels], K-major
              for (int hh = 0; hh < nhalf; ++hh) {
                const int pix = (mt * p.MT + hh) * kBM;
                if (p.a_im2col) {
                  int w, h, d;
                  pixel_to_whd(p, pix, w, h, d);
                  tma_load_im2col_5d(sa + hh * kABytes, &p.tmA, &full_bar[stage], kb * E::BKE, w, h, d, 0,
                                     (uint16_t)kz, (uint16_t)ky, (uint16_t)kx);
                } else {
                  tma_load_2d(sa + hh * kABytes, &p.tmA, &full_bar[stage], kb * E::BKE, pix);
                }
              }
              // ---- B (with multicast each CTA of the pair fetches half of the tile for both)
              if (!b_mn) {   // weights rows, K-major: box {BKE, BN} (or {BKE, BN/2} per CTA)
                if (p.mc) {
                  const int half = p.BN >> 1;
                  tma_load_2d_mc(sb + crank * half * 128, &p.tmB, &full_bar[stage], wtap * p.Cin + kb * E::BKE,
                                 nt * p.BN + crank * half, (uint16_t)3);
                } else {
                  tma_load_2d(sb, &p.tmB, &full_bar[stage], wtap * p.Cin + kb * E::BKE, nt * p.BN);
                }
              } else {       // dgrad: W in place, MN-major chunks of CH input channels x KR couts
                const int nch = p.BN / E::CH;
                if (p.mc) {
                  for (int j = crank * (nch >> 1); j < (crank + 1) * (nch >> 1); ++j)
                    tma_load_2d_mc(sb + j * E::MN_CHUNK, &p.tmB, &full_bar[stage],
                                   wtap * p.Cin + nt * p.BN + j * E::CH, kb * E::KR, (uint16_t)3);
                } else {
                  for (int j = 0; j < nch; ++j)
                    tma_load_2d(sb + j * E::MN_CHUNK, &p.tmB, &full_bar[stage],
                                wtap * p.Cin + nt * p.BN + j * E::CH, kb * E::KR);
                }
              }
            } else {
              // ---- A: MT dY^T tiles, MN-major: chunks of CH couts x KR voxels
              for (int hh = 0; hh < nhalf; ++hh)
                for (int j = 0; j < kBM / E::CH; ++j)
                  tma_load_2d(sa + hh * kABytes + j * E::MN_CHUNK, &p.tmA, &full_bar[stage],
                              (mt * p.MT + hh) * kBM + j * E::CH, kb * E::KR);
              // ---- B: NT X tiles (one per tap of the group), MN-major: chunks of CH cins x KR pixels
              const int nch = p.BN / E::CH;
              const uint32_t b_stride = (b_bytes + 1023u) / 1024u * 1024u;
              int w = 0, h = 0, d = 0;
              if (p.a_im2col) pixel_to_whd(p, kb * E::KR, w, h, d);
              for (int tt = 0; tt < ntap; ++tt) {
                const int tp = tap + tt;
                const int tx = (p.taps == 1) ? 0 : tp / 9, ty = (p.taps == 1) ? 0 : (tp / 3) % 3,
                          tz = (p.taps == 1) ? 0 : tp % 3;
                for (int j = 0; j < nch; ++j) {
                  if (p.a_im2col)
                    tma_load_im2col_5d(sb + tt * b_stride + j * E::MN_CHUNK, &p.tmB, &full_bar[stage],
                                       nt * p.BN + j * E::CH, w, h, d, 0, (uint16_t)tz, (uint16_t)ty,
                                       (uint16_t)tx);
                  else
                    tma_load_2d(sb + tt * b_stride + j * E::MN_CHUNK, &p.tmB, &full_bar[stage],
                                nt * p.BN + j * E::CH, kb * E::KR);
                }
              }
            }
            if (++stage == nstages) { stage = 0; phase ^= 1; }
          }
        }
      }
    }
  } else if (warp == 1) {
    // =========================== MMA issuer =============================================
    if (p.ky && wgrad) {
      if (elect_one()) {
        const uint32_t idesc = make_idesc(E::FMT, 1u, 1u, kBM, (uint32_t)p.BN);
        const uint32_t a_chunk = 128u * 128u;
        const int nchA = kBM / E::CH, nchB = p.BN / E::CH;
        const uint32_t stage_b = (uint32_t)nchA * a_chunk + (uint32_t)nchB * (uint32_t)p.ext_bytes;
        const uint32_t ky_step = (uint32_t)p.gZ * 128u;
        int stage = 0;
        uint32_t phase = 0;
        for (int it_ = 0;; ++it_) {
          const int tile = tile_consume(it_, false);
          if (tile < 0) break;
          const int sp = tile % p.ksplit;
          const int nk = (int)((long long)p.nvb * (sp + 1) / p.ksplit) - (int)((long long)p.nvb * sp / p.ksplit);
          mbar_wait(&tempty_bar[0], (uint32_t)(it_ & 1) ^ 1);
          tc_fence_after();
          for (int k = 0; k < nk; ++k) {
            mbar_wait(&full_bar[stage], phase);
            tc_fence_after();
            const uint32_t sa = smem_u32(smem + stage * stage_b);
            const uint32_t sb = sa + nchA * a_chunk;
            for (int tt = 0; tt < 3; ++tt) {
              const uint32_t dcol = tmem_base + (uint32_t)(tt * p.hstride);
#pragma unroll
              for (int ks = 0; ks < kBM / E::UK; ++ks) {
                const uint64_t ad = make_smem_desc(sa + ks * E::MN_KSTEP, a_chunk, E::MN_SBO, E::MN_LAYOUT);
                const uint64_t bd = make_smem_desc(sb + tt * ky_step + ks * E::MN_KSTEP, (uint32_t)p.ext_bytes,
                                                   E::MN_SBO, E::MN_LAYOUT);
                umma<E::TF32>(dcol, ad, bd, idesc, (k | ks) ? 1u : 0u);
              }
            }
            umma_commit(&empty_bar[stage]);
            if (++stage == nstages) { stage = 0; phase ^= 1; }
          }
          if (nk > 0) umma_commit(&tfull_bar[0]);
          else mbar_arrive(&tfull_bar[0]);
        }
      }
    } else if (p.ky) {
      if (elect_one()) {
        const uint32_t idesc = make_idesc(E::FMT, 0u, b_mn ? 1u : 0u, kBM, (uint32_t)p.BN);
        const uint32_t a_entry = (uint32_t)p.MT * (uint32_t)p.ext_bytes;
        const uint32_t b_entry = (b_bytes + 1023u) / 1024u * 1024u;
        const uint32_t ringA = smem_u32(smem);
        const uint32_t ringB = ringA + (uint32_t)p.NA * a_entry;
        const uint32_t ky_step = (uint32_t)p.gZ * 128u;       // one y step inside the extended tile
        int sa_i = 0, sb_i = 0;
        uint32_t pa = 0, pb = 0;
        int acc = 0;
        uint32_t acc_phase = 0;
        const int ngroups = 9 * kb_per_tap;
        for (int it_ = 0;; ++it_) {
          const int tile = tile_consume(it_, false);
          if (tile < 0) break;
          mbar_wait(&tempty_bar[acc], acc_phase ^ 1);
          tc_fence_after();
          const uint32_t d_tmem = tmem_base + (uint32_t)(acc * p.acc_stride);
          for (int g = 0; g < ngroups; ++g) {
            mbar_wait(&fulla_bar[sa_i], pa);
            const uint32_t sa = ringA + sa_i * a_entry;
            for (int ky = 0; ky < 3; ++ky) {
              mbar_wait(&full_bar[sb_i], pb);
              tc_fence_after();
              const uint32_t sbt = ringB + sb_i * b_entry;
              for (int hh = 0; hh < p.MT; ++hh) {
                const uint32_t sah = sa + hh * p.ext_bytes + ky * ky_step;
                const uint32_t dcol = d_tmem + (uint32_t)(hh * p.hstride);
#pragma unroll
                for (int j = 0; j < E::NUK; ++j) {
                  const uint64_t ad = make_smem_desc(sah + j * 32, 16, 1024, 2);
                  const uint64_t bd = b_mn ? make_smem_desc(sbt + j * E::MN_KSTEP, E::MN_CHUNK, E::MN_SBO, E::MN_LAYOUT)
                                           : make_smem_desc(sbt + j * 32, 16, 1024, 2);
                  umma<E::TF32>(dcol, ad, bd, idesc, (g | ky | j) ? 1u : 0u);
                }
              }
              umma_commit(&empty_bar[sb_i]);
              if (++sb_i == p.NB) { sb_i = 0; pb ^= 1; }
            }
            umma_commit(&emptya_bar[sa_i]);
            if (++sa_i == p.NA) { sa_i = 0; pa ^= 1; }
          }
          umma_commit(&tfull_bar[acc]);
          if (++acc == p.nacc) { acc = 0; acc_phase ^= 1; }
        }
      }
    } else if (elect_one()) {
      const uint32_t idesc = make_idesc(E::FMT, a_mn ? 1u : 0u, b_mn ? 1u : 0u, kBM, (uint32_t)p.BN);
      int stage = 0;
      uint32_t phase = 0;
      int acc = 0;
      uint32_t acc_phase = 0;
      for (int it_ = 0;; ++it_) {
          const int tile = tile_consume(it_, false);
          if (tile < 0) break;
        int nk, ntap = 1;
        {
          const int sp = tile % p.ksplit;
          const int ntaps_t = p.ncls ? (int)p.cls_ntaps[tile / p.ksplit / ntn / ntm] : p.taps;
          const int tot = wgrad ? p.nvb : ntaps_t * kb_per_tap;
          nk = (int)((long long)tot * (sp + 1) / p.ksplit) - (int)((long long)tot * sp / p.ksplit);
          if (wgrad) {
            const int tg = tile / p.ksplit / ntn / ntm;
            ntap = min(p.NT, p.taps - tg * p.NT);
          }
        }
        mbar_wait(&tempty_bar[acc], acc_phase ^ 1);
        tc_fence_after();
        const uint32_t d_tmem = tmem_base + (uint32_t)(acc * p.acc_stride);   // nacc == 1 -> acc stays 0
        for (int k = 0; k < nk; ++k) {
          mbar_wait(&full_bar[stage], phase);
          tc_fence_after();
          const uint32_t sa = smem_u32(smem + stage * stage_bytes);
          const uint32_t sb = sa + a_bytes;
          const uint32_t b_stride = (b_bytes + 1023u) / 1024u * 1024u;
          for (int hh = 0; hh < p.MT; ++hh) {
            const uint32_t sah = sa + hh * kABytes;
            for (int tt = 0; tt < ntap; ++tt) {
              const uint32_t sbt = sb + tt * b_stride;
              const uint32_t dcol = d_tmem + (uint32_t)((hh * p.NT + tt) * p.hstride);
#pragma unroll
              for (int j = 0; j < E::NUK; ++j) {
                const uint64_t ad = a_mn ? make_smem_desc(sah + j * E::MN_KSTEP, E::MN_CHUNK, E::MN_SBO, E::MN_LAYOUT)
                                         : make_smem_desc(sah + j * 32, 16, 1024, 2);
                const uint64_t bd = b_mn ? make_smem_desc(sbt + j * E::MN_KSTEP, E::MN_CHUNK, E::MN_SBO, E::MN_LAYOUT)
                                         : make_smem_desc(sbt + j * 32, 16, 1024, 2);
                umma<E::TF32>(dcol, ad, bd, idesc, (k | j) ? 1u : 0u);
              }
            }
          }
          if (p.mc) umma_commit_mc(&empty_bar[stage], (uint16_t)3);   // frees the slot in both CTAs
          else umma_commit(&empty_bar[stage]);     // frees the smem slot when the MMAs retire
          if (++stage == nstages) { stage = 0; phase ^= 1; }
        }
        if (nk > 0) {
          umma_commit(&tfull_bar[acc]);            // accumulator complete -> epilogue
        } else {
          mbar_arrive(&tfull_bar[acc]);
        }
        if (++acc == p.nacc) { acc = 0; acc_phase ^= 1; }
      }
    }
  } else if (warp >= 4) {
    // =========================== epilogue (4 warps = 128 TMEM lanes) =====================
    const int q = warp & 3;                 // TMEM lane quadrant of this warp
    const int csel = (warp - 4) >> 2;       // which of the quadrant's two warps: chunks csel, csel + 2, ...
    // staging tile of this warp (shared address; 0: none): statistics transpose + output rows for coalesced stores
    const uint32_t stat_tile = p.stat_off ? smem_u32(smem + p.stat_off) + (uint32_t)((warp - 4) * (32 * kStatRow * 4)) : 0u;
    // Running per-column sums of this warp over all the tiles of the CTA (cell = [column chunk of this warp][lane]),
    // flushed with global atomics once at the end (or when the N tile changes).  One atomic pair per tile and column
    // chunk put 20 000 same-line requests on each L2 slice holding `stats` -- the 1x1x1 layers ran at the speed of
    // those atomics (122 us for 246 MB), not of their loads.
    const int cshift = p.epi == 8 ? 6 : 5;                                  // columns between two chunks of this warp
    const uint32_t stat_acc = smem_u32(smem + kRingBytes + 512) + (uint32_t)((((warp - 4) * (32 / p.epi)) * 32 + lane) * 8);
    const int nslots = (p.BN - csel * 32 + (1 << cshift) - 1) >> cshift;    // chunks this warp handles per tile
    int acc_nt = -1;
    auto stat_flush = [&]() {
      for (int sl = 0; sl < nslots; ++sl) {
        const int cc = csel * 32 + (sl << cshift) + lane;
        const int col = acc_nt * p.BN + cc;
        const float2 a = lds_f32x2(stat_acc + (uint32_t)(sl * 256));
        if (col < p.N && cc < p.BN) {
          atomicAdd(&p.stats[col], a.x);
          atomicAdd(&p.stats[p.N + col], a.y);
        }
      }
    };
    if (p.stats != nullptr)
      for (int sl = 0; sl < nslots; ++sl) sts_f32x2(stat_acc + (uint32_t)(sl * 256), 0.f, 0.f);
    int acc = 0;
    uint32_t acc_phase = 0;
    for (int it_ = 0;; ++it_) {
      const int tile = tile_consume(it_, true);
      if (tile < 0) break;
      int mt, nt, tap = 0, nk = 1;
      int opx = p.opx, opy = p.opy, opz = p.opz;
      if (!wgrad) {
        int t = tile;
        const int sp = t % p.ksplit; t /= p.ksplit;
        nt = t % ntn;
        mt = p.mc ? 2 * (t / ntn) + crank : t / ntn;
        int ntaps_t = p.taps;
        if (p.ncls) {
          const int cls = mt / ntm;
          mt -= cls * ntm;
          ntaps_t = p.cls_ntaps[cls];
          const int par = p.cls_par[cls];
          opx = (par >> 2) & 1; opy = (par >> 1) & 1; opz = par & 1;
        }
        const int tot = ntaps_t * kb_per_tap;
        nk = (int)((long long)tot * (sp + 1) / p.ksplit) - (int)((long long)tot * sp / p.ksplit);
      } else {
        int t = tile;
        const int sp = t % p.ksplit; t /= p.ksplit;
        nt = t % ntn; t /= ntn;
        mt = t % ntm; t /= ntm;
        tap = t * p.NT;              // first tap of the group (ky path: t = kx * 3 + kz, see wtap)
        nk = (int)((long long)p.nvb * (sp + 1) / p.ksplit) - (int)((long long)p.nvb * sp / p.ksplit);
      }
      if (p.stats != nullptr && nt != acc_nt) {
        if (acc_nt >= 0) {
          stat_flush();
          for (int sl = 0; sl < nslots; ++sl) sts_f32x2(stat_acc + (uint32_t)(sl * 256), 0.f, 0.f);
        }
        acc_nt = nt;
      }
      mbar_wait(&tfull_bar[acc], acc_phase);
      tc_fence_after();
      const int nsub = p.MT * p.NT;
      for (int u = 0; u < nsub; ++u) {
      const int hh = u / p.NT, tt = u % p.NT;
      // wgrad: tap of this sub-accumulator (ky path: the group is (kx, 0..2, kz))
      const int wtap = (wgrad && p.ky) ? ((tap / 9) * 9 + tt * 3 + (tap / 3) % 3) : tap + tt;
      if (wgrad && wtap >= p.taps) continue;
      int row = (mt * p.MT + hh) * kBM + q * 32 + lane;
      bool row_ok = row < p.M;
      if (p.omap && row_ok) {
        const int z = row % p.oZ, t = row / p.oZ;
        const int y = t % p.oY, x = t / p.oY;
        const int fx = 2 * x + opx, fy = 2 * y + opy, fz = 2 * z + opz;
        row_ok = fx < p.fX && fy < p.fY && fz < p.fZ;
        row = (fx * p.fY + fy) * p.fZ + fz;
      }
      if (p.ky && !wgrad) {
        // block (x, yb): 128 consecutive pixels of x row `x` starting at yb * 128
        const int mb = mt * p.MT + hh;
        const int x = mb / p.nyb, in_x = (mb - x * p.nyb) * kBM + q * 32 + lane;
        row = x * p.gYZ + in_x;
        row_ok = (mb < nmb) && (in_x < p.gYZ);
      }
      const uint32_t t_base = tmem_base + (uint32_t)(acc * p.acc_stride) + (uint32_t)(u * p.hstride) + ((uint32_t)(q * 32) << 16);
      if (!wgrad && p.ksplit == 1 && p.out_bf16 && !p.accum && nk > 0 && (p.ldc & 7) == 0 && (p.BN & 31) == 0 &&
          (nt + 1) * p.BN <= p.N && (reinterpret_cast<uintptr_t>(p.bias) & 15) == 0) {
        // common case (bf16 activations, whole chunks): a lean loop -- the generic one below re-tests its options per
        // chunk and per value (~350 instructions per chunk besides the statistics)
        __nv_bfloat16* orow = reinterpret_cast<__nv_bfloat16*>(p.out) + (long long)row * p.ldc + nt * p.BN;
        // Row-per-lane 16-byte stores fill half a 32-byte sector per request and the write-heavy 1x1x1 layers ran at
        // the L2's sector rate (128->256 on 640 k voxels: 237 us for 492 MB).  With a staging tile the warp transposes
        // the packed rows through shared memory (XOR-swizzled 16-byte pieces, conflict-free both ways) and every
        // store instruction writes 8 rows x 64 contiguous bytes.
        const bool staged = stat_tile != 0u && !p.omap;
        const int nvalid = __popc(__ballot_sync(0xffffffffu, row_ok));          // valid rows are a prefix
        __nv_bfloat16* const obase = orow - (long long)lane * p.ldc;            // row 0 of this warp
        // gradient of a skip connection added on the way out (same row index as the output, own row stride)
        const __nv_bfloat16* const arow =
            p.addend ? reinterpret_cast<const __nv_bfloat16*>(p.addend) + (long long)row * p.ld_add + nt * p.BN : nullptr;
        const __nv_bfloat16* const abase = arow ? arow - (long long)lane * p.ld_add : nullptr;
        for (int c0 = csel * 32; c0 < p.BN; c0 += 1 << cshift) {
          uint32_t v[32];
          tmem_ld_32x32(t_base + (uint32_t)c0, v);
          tmem_ld_wait();
          if (p.stats != nullptr) epi_stats(v, row_ok, stat_tile, stat_acc + (uint32_t)((c0 >> cshift) * 256), lane);
          if (p.bias != nullptr) {
            const float4* b4 = reinterpret_cast<const float4*>(p.bias + nt * p.BN + c0);
#pragma unroll
            for (int i = 0; i < 8; ++i) {
              const float4 b = __ldg(b4 + i);
              v[4 * i] = __float_as_uint(__uint_as_float(v[4 * i]) + b.x);
              v[4 * i + 1] = __float_as_uint(__uint_as_float(v[4 * i + 1]) + b.y);
              v[4 * i + 2] = __float_as_uint(__uint_as_float(v[4 * i + 2]) + b.z);
              v[4 * i + 3] = __float_as_uint(__uint_as_float(v[4 * i + 3]) + b.w);
            }
          }
          if (p.relu) {
#pragma unroll
            for (int i = 0; i < 32; ++i) v[i] = __float_as_uint(fmaxf(__uint_as_float(v[i]), 0.f));
          }
          if (staged) {
            const int sw = (lane >> 1) & 3;
#pragma unroll
            for (int pc = 0; pc < 4; ++pc) {
              uint32_t w4[4];
#pragma unroll
              for (int j = 0; j < 4; ++j) {
                const __nv_bfloat162 h =
                    __floats2bfloat162_rn(__uint_as_float(v[8 * pc + 2 * j]), __uint_as_float(v[8 * pc + 2 * j + 1]));
                w4[j] = *reinterpret_cast<const uint32_t*>(&h);
              }
              sts_u32x4(stat_tile + (uint32_t)(64 * lane + 16 * (pc ^ sw)), w4[0], w4[1], w4[2], w4[3]);
            }
            __syncwarp();
            const int pp = lane & 3;
#pragma unroll
            for (int j = 0; j < 4; ++j) {
              const int r = 8 * j + (lane >> 2);
              uint4 val = lds_u32x4(stat_tile + (uint32_t)(64 * r + 16 * (pp ^ ((r >> 1) & 3))));
              if (r < nvalid) {
                if (abase != nullptr)
                  val = bf16x8_add(val, __ldg(reinterpret_cast<const uint4*>(abase + (long long)r * p.ld_add + c0 + 8 * pp)));
                *reinterpret_cast<uint4*>(obase + (long long)r * p.ldc + c0 + 8 * pp) = val;
              }
            }
            __syncwarp();
          } else if (row_ok) {
#pragma unroll
            for (int i = 0; i < 32; i += 8) {
              uint32_t w4[4];
#pragma unroll
              for (int j = 0; j < 4; ++j) {
                const __nv_bfloat162 h = __floats2bfloat162_rn(__uint_as_float(v[i + 2 * j]), __uint_as_float(v[i + 2 * j + 1]));
                w4[j] = *reinterpret_cast<const uint32_t*>(&h);
              }
              uint4 val = make_uint4(w4[0], w4[1], w4[2], w4[3]);
              if (arow != nullptr) val = bf16x8_add(val, __ldg(reinterpret_cast<const uint4*>(arow + c0 + i)));
              *reinterpret_cast<uint4*>(orow + c0 + i) = val;
            }
          }
        }
        continue;
      }
      for (int c0 = csel * 32; c0 < p.BN; c0 += 1 << cshift) {
        uint32_t v[32];
        if (nk > 0) {
          tmem_ld_32x32(t_base + (uint32_t)c0, v);
          tmem_ld_wait();
        } else {
#pragma unroll
          for (int i = 0; i < 32; ++i) v[i] = 0u;
        }
        const int col0 = nt * p.BN + c0;
        if (!wgrad && p.ksplit > 1) {
          // split-K partial of an under-filled layer: accumulate atomically into the zeroed output
          if (row_ok && nk > 0) {
            float* o = p.out + (long long)row * p.ldc + col0;
            if (col0 + 32 <= p.N && (p.ldc & 3) == 0) {
#pragma unroll
              for (int i = 0; i < 32; i += 4) {
                asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(o + i),
                             "f"(__uint_as_float(v[i])), "f"(__uint_as_float(v[i + 1])),
                             "f"(__uint_as_float(v[i + 2])), "f"(__uint_as_float(v[i + 3]))
                             : "memory");
              }
            } else {
#pragma unroll
              for (int i = 0; i < 32; ++i)
                if (col0 + i < p.N) atomicAdd(o + i, __uint_as_float(v[i]));
            }
          }
        } else if (!wgrad) {
          float* o = p.out + (long long)row * p.ldc + col0;
          const bool vec = (col0 + 32 <= p.N) && ((p.ldc & 3) == 0);
          if (p.accum && row_ok) {
            // out += acc : second / third pass of the 3xTF32 split
            if (vec) {
#pragma unroll
              for (int i = 0; i < 32; i += 4) {
                const float4 f = *reinterpret_cast<const float4*>(o + i);
                v[i] = __float_as_uint(__uint_as_float(v[i]) + f.x);
                v[i + 1] = __float_as_uint(__uint_as_float(v[i + 1]) + f.y);
                v[i + 2] = __float_as_uint(__uint_as_float(v[i + 2]) + f.z);
                v[i + 3] = __float_as_uint(__uint_as_float(v[i + 3]) + f.w);
              }
            } else {
#pragma unroll
              for (int i = 0; i < 32; ++i)
                if (col0 + i < p.N) v[i] = __float_as_uint(__uint_as_float(v[i]) + o[i]);
            }
          }
          if (p.stats != nullptr) epi_stats(v, row_ok, stat_tile, stat_acc + (uint32_t)((c0 >> cshift) * 256), lane);
          if (row_ok && p.out_bf16) {
            // bf16 activations: the next convolution's operand type, written straight from the epilogue
            __nv_bfloat16* ob = reinterpret_cast<__nv_bfloat16*>(p.out) + (long long)row * p.ldc + col0;
            if (col0 + 32 <= p.N && (p.ldc & 7) == 0) {
#pragma unroll
              for (int i = 0; i < 32; i += 8) {
                uint32_t w4[4];
#pragma unroll
                for (int j = 0; j < 4; ++j) {
                  float a = __uint_as_float(v[i + 2 * j]), b = __uint_as_float(v[i + 2 * j + 1]);
                  if (p.bias) { a += p.bias[col0 + i + 2 * j]; b += p.bias[col0 + i + 2 * j + 1]; }
                  if (p.relu) { a = fmaxf(a, 0.f); b = fmaxf(b, 0.f); }
                  const __nv_bfloat162 h = __floats2bfloat162_rn(a, b);
                  w4[j] = *reinterpret_cast<const uint32_t*>(&h);
                }
                *reinterpret_cast<uint4*>(ob + i) = make_uint4(w4[0], w4[1], w4[2], w4[3]);
              }
            } else {
#pragma unroll
              for (int i = 0; i < 32; ++i) {
                if (col0 + i < p.N) {
                  float f = __uint_as_float(v[i]);
                  if (p.bias) f += p.bias[col0 + i];
                  if (p.relu) f = fmaxf(f, 0.f);
                  ob[i] = __float2bfloat16_rn(f);
                }
              }
            }
          } else if (row_ok) {
            if (vec) {
#pragma unroll
              for (int i = 0; i < 32; i += 4) {
                float4 f;
                f.x = __uint_as_float(v[i]); f.y = __uint_as_float(v[i + 1]);
                f.z = __uint_as_float(v[i + 2]); f.w = __uint_as_float(v[i + 3]);
                if (p.bias) {
                  f.x += p.bias[col0 + i]; f.y += p.bias[col0 + i + 1];
                  f.z += p.bias[col0 + i + 2]; f.w += p.bias[col0 + i + 3];
                }
                if (p.relu) {
                  f.x = fmaxf(f.x, 0.f); f.y = fmaxf(f.y, 0.f);
                  f.z = fmaxf(f.z, 0.f); f.w = fmaxf(f.w, 0.f);
                }
                *reinterpret_cast<float4*>(o + i) = f;
              }
            } else {
#pragma unroll
              for (int i = 0; i < 32; ++i) {
                if (col0 + i < p.N) {
                  float f = __uint_as_float(v[i]);
                  if (p.bias) f += p.bias[col0 + i];
                  if (p.relu) f = fmaxf(f, 0.f);
                  o[i] = f;
                }
              }
            }
          }
        } else if (row_ok && nk > 0) {
          // wgrad: accumulate the split-K partial into dW[row][tap][col]
          float* o = p.out + ((long long)row * p.taps + wtap) * p.ldc + col0;
          if (col0 + 32 <= p.N && (p.ldc & 3) == 0) {
#pragma unroll
            for (int i = 0; i < 32; i += 4) {
              asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(o + i),
                           "f"(__uint_as_float(v[i])), "f"(__uint_as_float(v[i + 1])),
                           "f"(__uint_as_float(v[i + 2])), "f"(__uint_as_float(v[i + 3]))
                           : "memory");
            }
          } else {
#pragma unroll
            for (int i = 0; i < 32; ++i)
              if (col0 + i < p.N) atomicAdd(o + i, __uint_as_float(v[i]));
          }
        }
      }
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&tempty_bar[acc]);
      if (++acc == p.nacc) { acc = 0; acc_phase ^= 1; }
    }
    if (p.stats != nullptr && acc_nt >= 0) stat_flush();
  }

  tc_fence_before();
  __syncthreads();
  if (p.mc) cluster_sync_all();       // the peer may still multicast into / arrive on this CTA
  if (warp == 2) {
    tc_fence_after();
    tmem_dealloc(tmem_base, 512);
  }
}

// ------------------------------------------------------------------------------------------
// host side
// ------------------------------------------------------------------------------------------
typedef CUresult (*PFN_encodeTiled)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*,
                                    const cuuint64_t*, const cuuint64_t*, const cuuint32_t*,
                                    const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                    CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
typedef CUresult (*PFN_encodeIm2col)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*,
                                     const cuuint64_t*, const cuuint64_t*, const int*, const int*,
                                     cuuint32_t, cuuint32_t, const cuuint32_t*,
                                     CUtensorMapInterleave, CUtensorMapSwizzle,
                                     CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static PFN_encodeTiled g_encodeTiled = nullptr;
static PFN_encodeIm2col g_encodeIm2col = nullptr;
static int g_num_sms = 0;
static int g_driver_version = 0;

static int init_driver_api() {
  if (g_encodeTiled && g_encodeIm2col) return 0;
  cudaDriverEntryPointQueryResult qres;
  void* fn = nullptr;
  if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &qres) != cudaSuccess || !fn)
    return COOCC_ERR_DRIVER;
  g_encodeTiled = reinterpret_cast<PFN_encodeTiled>(fn);
  fn = nullptr;
  if (cudaGetDriverEntryPoint("cuTensorMapEncodeIm2col", &fn, cudaEnableDefault, &qres) != cudaSuccess || !fn)
    return COOCC_ERR_DRIVER;
  g_encodeIm2col = reinterpret_cast<PFN_encodeIm2col>(fn);
  int dev = 0;
  cudaGetDevice(&dev);
  cudaDeviceGetAttribute(&g_num_sms, cudaDevAttrMultiProcessorCount, dev);
  cudaDriverGetVersion(&g_driver_version);
  return 0;
}

static CUtensorMapDataType tm_dtype(int es) {
  // TFLOAT32: the TMA unit rounds fp32 to tf32 (round-to-nearest) on its way to shared memory,
  // so the MMA consumes correctly rounded operands instead of truncated ones.
  return es == 4 ? CU_TENSOR_MAP_DATA_TYPE_TFLOAT32 : CU_TENSOR_MAP_DATA_TYPE_BFLOAT16;
}

// 2-D row-major matrix [rows][cols] with row stride ld (elements); box {box_c, box_r}
static int make_tm_2d(CUtensorMap* tm, const void* ptr, int es, long long rows, long long cols,
                      long long ld, int box_c, int box_r, bool mn_major) {
  cuuint64_t gdim[2] = {(cuuint64_t)cols, (cuuint64_t)rows};
  cuuint64_t gstr[1] = {(cuuint64_t)ld * es};
  cuuint32_t box[2] = {(cuuint32_t)box_c, (cuuint32_t)box_r};
  cuuint32_t estr[2] = {1, 1};
  CUtensorMapSwizzle sw = (mn_major && es == 4) ? CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B
                                                : CU_TENSOR_MAP_SWIZZLE_128B;
  CUresult r = g_encodeTiled(tm, tm_dtype(es), 2, const_cast<void*>(ptr), gdim, gstr, box, estr,
                             CU_TENSOR_MAP_INTERLEAVE_NONE, sw, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                             CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    fprintf(stderr, "[coocc_b200] cuTensorMapEncodeTiled failed: %d (rows=%lld cols=%lld ld=%lld box=%d,%d)\n",
            (int)r, rows, cols, ld, box_c, box_r);
    return COOCC_ERR_TENSORMAP;
  }
  return 0;
}

// [X][rows][cols] tensor (row stride ld elements, x stride rows*ld), box {box_c, box_r, 1}: rows past
// the end of an x slab are zero-filled instead of running into the next slab
static int make_tm_3d(CUtensorMap* tm, const void* ptr, int es, long long X, long long rows, long long cols,
                      long long ld, int box_c, int box_r, bool mn_major) {
  cuuint64_t gdim[3] = {(cuuint64_t)cols, (cuuint64_t)rows, (cuuint64_t)X};
  cuuint64_t gstr[2] = {(cuuint64_t)ld * es, (cuuint64_t)ld * es * rows};
  cuuint32_t box[3] = {(cuuint32_t)box_c, (cuuint32_t)box_r, 1};
  cuuint32_t estr[3] = {1, 1, 1};
  CUtensorMapSwizzle sw = (mn_major && es == 4) ? CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B
                                                : CU_TENSOR_MAP_SWIZZLE_128B;
  CUresult r = g_encodeTiled(tm, tm_dtype(es), 3, const_cast<void*>(ptr), gdim, gstr, box, estr,
                             CU_TENSOR_MAP_INTERLEAVE_NONE, sw, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                             CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    fprintf(stderr, "[coocc_b200] cuTensorMapEncodeTiled(3d) failed: %d\n", (int)r);
    return COOCC_ERR_TENSORMAP;
  }
  return 0;
}

// NDHWC tensor [1][X][Y][Z][C] with pixel stride ld (elements), im2col box {chan, pixels}
static int make_tm_im2col(CUtensorMap* tm, const void* ptr, int es, int X, int Y, int Z, int C,
                          long long ld, int lo, int hi, int stride, int chan, int pixels,
                          bool mn_major, int hi_h = -1000) {
  cuuint64_t gdim[5] = {(cuuint64_t)C, (cuuint64_t)Z, (cuuint64_t)Y, (cuuint64_t)X, 1};
  cuuint64_t gstr[4] = {(cuuint64_t)ld * es, (cuuint64_t)ld * es * Z, (cuuint64_t)ld * es * Z * Y,
                        (cuuint64_t)ld * es * Z * Y * X};
  int lower[3] = {lo, lo, lo};
  int upper[3] = {hi, hi_h == -1000 ? hi : hi_h, hi};     // (W, H, D) = (z, y, x)
  cuuint32_t estr[5] = {1, (cuuint32_t)stride, (cuuint32_t)stride, (cuuint32_t)stride, 1};
  CUtensorMapSwizzle sw = (mn_major && es == 4) ? CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B
                                                : CU_TENSOR_MAP_SWIZZLE_128B;
  CUresult r = g_encodeIm2col(tm, tm_dtype(es), 5, const_cast<void*>(ptr), gdim, gstr, lower, upper,
                              (cuuint32_t)chan, (cuuint32_t)pixels, estr,
                              CU_TENSOR_MAP_INTERLEAVE_NONE, sw, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                              CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    fprintf(stderr, "[coocc_b200] cuTensorMapEncodeIm2col failed: %d (X=%d Y=%d Z=%d C=%d ld=%lld lo=%d hi=%d s=%d box=%d,%d)\n",
            (int)r, X, Y, Z, C, ld, lo, hi, stride, chan, pixels);
    return COOCC_ERR_TENSORMAP;
  }
  // Same workaround CUTLASS applies (cute/atom/copy_traits_sm90_im2col.hpp): drivers <= 13.1
  // mis-encode im2col maps of tensors smaller than 128 KiB; clear bit 21 of the second word.
  if (g_driver_version <= 13010) {
    const long long bytes = (long long)ld * es * Z * Y * X;
    if (bytes < 131072) reinterpret_cast<uint64_t*>(tm)[1] &= ~(1ull << 21);
  }
  return 0;
}

static int out_dim(int in, int ksize, int stride) {
  const int pad = ksize / 2;
  return (in + 2 * pad - ksize) / stride + 1;
}

// epilogue warps of a launch (COOCC_CONV_EPI=4|8 forces one setting)
static int pick_epi(const TcParams& p) {
  static int forced = -1;
  if (forced < 0) {
    const char* e = getenv("COOCC_CONV_EPI");
    forced = e ? atoi(e) : 0;
  }
  if (forced == 4 || forced == 8) return forced;
  return kEpiWarps;
}

// SMs the persistent grids may use (0 = all).  graph.GraphedStep lowers it while it captures the convolutions that
// run next to the FPS side branch (32 SMs for its two 16-CTA clusters): a persistent grid of 148 CTAs with 32 SMs
// taken would run its last 32 CTAs as a second wave.
static int g_sm_budget = 0;
extern "C" int coocc_conv_set_sm_budget(int n) {
  g_sm_budget = n > 0 ? n : 0;
  return 0;
}

// Dynamic tile scheduling (see the kernel): counters come from a ring of 64 pairs, one per launch in turn, so kernels
// of different streams that overlap do not share one; the last CTA of a launch resets its pair.
__device__ int g_tile_sched[64][2];
static int g_dyn_sched = 0;
static unsigned g_sched_seq = 0;
extern "C" int coocc_conv_set_dynamic(int mode) {
  g_dyn_sched = (mode == 1 || mode == 2) ? mode : 0;      // 1: every launch, 2: launches with many long tiles only
  return 0;
}

template <int ES>
static int launch(const TcParams& p_in, int ntiles, cudaStream_t st) {
  TcParams p = p_in;
  p.epi = pick_epi(p);
  p.sched = nullptr;
  static int dyn_always = -1;
  if (dyn_always < 0) {
    const char* e = getenv("COOCC_CONV_DYNAMIC");
    dyn_always = (e && e[0] == '2') ? 1 : 0;          // 2: every launch (experiments); else only where the runner asks
  }
  bool want_dyn = g_dyn_sched == 1 || dyn_always;
  if (g_dyn_sched == 2 && p.ksplit <= 1 && ntiles >= 2 * g_num_sms) {
    // long tiles only: a short tile would wait for its scheduler atomic (+3 % over a whole step when used everywhere)
    const int bke = 128 / (p.es ? p.es : 2);
    const int kb = p.mode == MODE_WGRAD ? p.nvb : (p.ky ? 27 : (p.ncls ? 4 : p.taps)) * ((p.Kc + bke - 1) / bke);
    want_dyn = kb >= 27;
  }
  if (want_dyn && !p.mc) {
    static int* base = nullptr;
    if (!base && cudaGetSymbolAddress(reinterpret_cast<void**>(&base), g_tile_sched) != cudaSuccess) return COOCC_ERR_CUDA;
    p.sched = base + 2 * (g_sched_seq++ % 64u);
  }
  if (p.acc_stride == 0) p.acc_stride = kMaxBN;
  static int stat_smem = -1;
  if (stat_smem < 0) {
    const char* e = getenv("COOCC_CONV_STATSMEM");
    stat_smem = (e && e[0] == '0') ? 0 : 1;
  }
  if (!stat_smem) p.stat_off = 0;
  static bool attr_set = false;
  if (!attr_set) {
    cudaError_t e = cudaFuncSetAttribute(tc_conv_kernel<ES>, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmemBytes);
    if (e != cudaSuccess) return COOCC_ERR_CUDA;
    attr_set = true;
  }
  if (ntiles < 1) return 0;
  cudaLaunchConfig_t cfg;
  memset(&cfg, 0, sizeof(cfg));
  cudaLaunchAttribute attr[1];
  if (p.mc) {
    // ntiles counts CTA pairs
    const int sms = (g_sm_budget > 0 && g_sm_budget < g_num_sms) ? g_sm_budget : g_num_sms;
    const int pairs = ntiles < sms / 2 ? ntiles : sms / 2;
    cfg.gridDim = dim3(2 * pairs, 1, 1);
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = 2;
    attr[0].val.clusterDim.y = 1;
    attr[0].val.clusterDim.z = 1;
    cfg.attrs = attr;
    cfg.numAttrs = 1;
  } else {
    const int sms = (g_sm_budget > 0 && g_sm_budget < g_num_sms && !p.sched) ? g_sm_budget : g_num_sms;
    cfg.gridDim = dim3(ntiles < sms ? ntiles : sms, 1, 1);
  }
  cfg.blockDim = dim3(128 + 32 * p.epi, 1, 1);
  cfg.dynamicSmemBytes = kSmemBytes;
  cfg.stream = st;
  return cudaLaunchKernelEx(&cfg, tc_conv_kernel<ES>, p) == cudaSuccess ? 0 : COOCC_ERR_CUDA;
}

// column sums / sums of squares of a small [M, N] matrix (BN statistics of split-K layers)
__global__ void __launch_bounds__(256) colstats_kernel(const float* __restrict__ y, long long ld, int M, int N,
                                                       float* __restrict__ stats) {
  const int c = blockIdx.x * 32 + (threadIdx.x & 31);
  float s = 0.f, q = 0.f;
  if (c < N)
    for (int r = blockIdx.y * 8 + (threadIdx.x >> 5); r < M; r += gridDim.y * 8) {
      const float v = y[(long long)r * ld + c];
      s += v;
      q += v * v;
    }
  __shared__ float sh[2][8][32];
  sh[0][threadIdx.x >> 5][threadIdx.x & 31] = s;
  sh[1][threadIdx.x >> 5][threadIdx.x & 31] = q;
  __syncthreads();
  if (threadIdx.x < 32 && c < N) {
    for (int j = 1; j < 8; ++j) { s += sh[0][j][threadIdx.x]; q += sh[1][j][threadIdx.x]; }
    atomicAdd(&stats[c], s);
    atomicAdd(&stats[N + c], q);
  }
}

// split the reduction of an under-filled fprop/dgrad launch over several CTAs
static int pick_ksplit(int ntiles, int nk_total) {
  static int model = -1;
  if (model < 0) {
    const char* e = getenv("COOCC_CONV_KSPLIT_MODEL");
    model = (e && e[0] == '0') ? 0 : 1;
  }
  if (!model) {
    if (ntiles * 2 > g_num_sms || nk_total < 8) return 1;
    int ks = g_num_sms / ntiles;
    if (ks > nk_total / 4) ks = nk_total / 4;
    return ks < 1 ? 1 : ks;
  }
  // Wave-quantisation model, in k-block units (one k-block of a 128 x 256 tile ~ 0.4 us): the persistent CTAs take
  // ceil(tiles * ks / SMs) rounds of nk / ks k-blocks plus a fixed cost per tile (pipeline fill, epilogue -- atomic
  // for ks > 1), and a split launch pays for its scratch memset, the bf16 conversion and the statistics pass.
  // 79 or 158 tiles on 148 SMs (the 50x50x4 levels) run at 53 % unsplit; 3-7 splits bring that to ~85 %.
  if (nk_total < 16 || ntiles >= 4 * g_num_sms) return 1;
  double best = 1e30;
  int best_ks = 1;
  for (int ks = 1; ks <= 16 && ks * 8 <= nk_total; ++ks) {
    const long long waves = ((long long)ntiles * ks + g_num_sms - 1) / g_num_sms;
    const double t = (double)waves * ((double)nk_total / ks + (ks > 1 ? 16.0 : 10.0)) + (ks > 1 ? 64.0 : 0.0);
    if (t < best * 0.97) { best = t; best_ks = ks; }      // (a larger split has to win by 3 %)
  }
  return best_ks;
}

static int prepare_split(TcParams& p, int nk_total, cudaStream_t st, bool zero_out = true) {
  // Several 128-row MMA tiles per CTA tile share one B tile: fewer operand bytes per MAC have to
  // be pulled from L2 into the SM (the measured limiter, profiles/r01_ncu_full_tc_conv_bf16.summary.txt).
  //   BN <= 128: MT = 2, two TMEM buffers (epilogue overlaps the next tile)
  //   BN == 256: MT = 2 uses all 512 TMEM columns -> single buffer, only for long reductions where the
  //              exposed epilogue is a few % of the tile; 64 KiB stages -> 3-deep ring
  p.MT = 1;
  static int mt_enabled = -1;
  if (mt_enabled < 0) {
    const char* e = getenv("COOCC_CONV_MT2");
    mt_enabled = e ? atoi(e) : 2;       // 0: off, 1: narrow tiles only, 2: also 256-wide tiles
  }
  const int ntn_ = (p.N + p.BN - 1) / p.BN;
  if (mt_enabled >= 1 && p.BN <= 128 && ((p.M + 2 * kBM - 1) / (2 * kBM)) * ntn_ >= g_num_sms) p.MT = 2;
  if (mt_enabled >= 2 && p.BN > 128 && nk_total >= 32 && ((p.M + 2 * kBM - 1) / (2 * kBM)) * ntn_ >= g_num_sms)
    p.MT = 2;
  p.NT = 1;
  p.hstride = p.BN <= 128 ? 128 : 256;
  p.nacc = (p.MT * p.hstride <= 256) ? 2 : 1;
  {
    static int acc4 = -1;
    if (acc4 < 0) {
      const char* e = getenv("COOCC_CONV_ACC4");
      acc4 = (e && e[0] == '1') ? 1 : 0;      // measured slower than 256-row tiles with two buffers (91 vs 75 us): opt-in
    }
    if (acc4 && p.taps == 1 && p.BN <= 128) { p.MT = 1; p.nacc = 4; p.acc_stride = 128; }
  }
  {
    const int stage = kABytes * p.MT + (p.BN * 128 + 1023) / 1024 * 1024;
    // (COOCC_CONV_DEEP_RING=1: up to 8 stages for the 1x1x1 layers.  Measured: no change -- 75 / 99 / 172 us for
    // 128->64 / 128 / 256 on 640 k voxels either way; those layers are bound by the per-tile load -> MMA -> epilogue
    // latency chain, not by bytes in flight -- so it stays off)
    static int deep = -1;
    if (deep < 0) {
      const char* e = getenv("COOCC_CONV_DEEP_RING");
      deep = (e && e[0] == '1') ? 1 : 0;
    }
    const int cap = (p.taps == 1 && deep) ? kMaxBRing : kStages;
    p.nstages = kRingBytes / stage;
    if (p.nstages > cap) p.nstages = cap;
    // statistics staging tiles go behind the ring when it leaves room (the 1x1x1 and narrow layers: exactly the
    // ones whose epilogue is exposed); they also stage the bf16 output rows for coalesced stores, for which a 1x1x1
    // layer gives up ring stages (down to 3)
    if (p.taps == 1)
      while (p.nstages > 3 && kRingBytes - p.nstages * stage < kStatBytes) --p.nstages;
    p.stat_off = (kRingBytes - p.nstages * stage >= kStatBytes) ? p.nstages * stage : 0;
  }
  const int ntm = (p.M + kBM * p.MT - 1) / (kBM * p.MT);
  const int ntiles = ntm * ((p.N + p.BN - 1) / p.BN);
  // CTA pairs sharing the weight tile by TMA multicast (halves the B traffic out of L2 per CTA)
  static int mc_enabled = -1;
  if (mc_enabled < 0) {
    // measured on B200 (profiles/): no gain -- the limiter is the per-SM operand ingest, which
    // multicast does not reduce -- so the pairing is opt-in (COOCC_CONV_MC=1)
    const char* e = getenv("COOCC_CONV_MC");
    mc_enabled = (e && e[0] == '1') ? 1 : 0;
  }
  const int es = p.es;
  const int ch = 128 / es;
  p.mc = 0;
  if (mc_enabled && ntm >= 2) {
    if (p.mode == MODE_FPROP && ((p.BN / 2) % 8) == 0) p.mc = 1;
    if (p.mode == MODE_DGRAD && ((p.BN / ch) % 2) == 0) p.mc = 1;
  }
  p.ksplit = 1;
  if (p.bias == nullptr && !p.relu) p.ksplit = pick_ksplit(ntiles, nk_total);
  if (p.ksplit > 1 && !p.accum && zero_out) {
    if (cudaMemset2DAsync(p.out, (size_t)p.ldc * sizeof(float), 0, (size_t)p.N * sizeof(float), (size_t)p.M, st) !=
        cudaSuccess)
      return COOCC_ERR_CUDA;
  }
  return 0;
}

// fp32 [M, N] (row stride lds) -> bf16 (row stride ldd): epilogue of split-K launches whose output is bf16
__global__ void __launch_bounds__(256) f32_to_bf16_rows_kernel(const float* __restrict__ src, long long lds, int M,
                                                               int N, __nv_bfloat16* __restrict__ dst, long long ldd) {
  const long long total = (long long)M * N;
  for (long long i = blockIdx.x * 256LL + threadIdx.x; i < total; i += gridDim.x * 256LL) {
    const long long r = i / N;
    const int c = (int)(i % N);
    dst[r * ldd + c] = __float2bfloat16_rn(src[r * lds + c]);
  }
}

// Split-K launches accumulate with fp32 atomics; when the caller wants bf16 the partial sums go to
// a stream-ordered fp32 scratch matrix that is converted afterwards (only the small, under-filled
// layers take this route).
struct SplitScratch {
  float* tmp = nullptr;
  void* real_out = nullptr;
  long long real_ld = 0;
};
static int split_scratch_begin(TcParams& p, SplitScratch& sc, cudaStream_t st) {
  if (p.ksplit <= 1) return 0;
  if (p.out_bf16) {
    const long long ld = ((long long)p.N + 3) / 4 * 4;
    if (cudaMallocAsync(reinterpret_cast<void**>(&sc.tmp), (size_t)p.M * ld * sizeof(float), st) != cudaSuccess)
      return COOCC_ERR_CUDA;
    sc.real_out = p.out; sc.real_ld = p.ldc;
    p.out = sc.tmp; p.ldc = ld; p.out_bf16 = 0;
  }
  if (!p.accum &&
      cudaMemset2DAsync(p.out, (size_t)p.ldc * sizeof(float), 0, (size_t)p.N * sizeof(float), (size_t)p.M, st) != cudaSuccess)
    return COOCC_ERR_CUDA;
  return 0;
}
static int split_scratch_end(TcParams& p, SplitScratch& sc, cudaStream_t st) {
  if (!sc.tmp) return 0;
  long long blocks = ((long long)p.M * p.N + 255) / 256;
  if (blocks > 148 * 16) blocks = 148 * 16;
  f32_to_bf16_rows_kernel<<<(int)blocks, 256, 0, st>>>(sc.tmp, p.ldc, p.M, p.N,
                                                      reinterpret_cast<__nv_bfloat16*>(sc.real_out), sc.real_ld);
  const int rc = cudaGetLastError() == cudaSuccess ? 0 : COOCC_ERR_CUDA;
  cudaFreeAsync(sc.tmp, st);
  return rc;
}

// ky-fused path for 3x3x3 stride-1 fprop / dgrad.  The im2col A operand of the three taps
// (kx, 0..2, kz) is the same set of input rows shifted by one y step = Z rows of 128 B, so one
// y-extended tile ((ny + 2) * Z pixel rows, loaded by a single im2col request whose H bounding box
// is widened to y in [-1, nyb*ny]) serves all three: the MMA descriptors of tap ky start
// ky * Z * 128 bytes into the tile (a multiple of the 1 KiB swizzle atom for Z % 8 == 0).  A bytes
// pulled out of L2 per MAC drop 2.4x (Z = 16) -- the measured limiter of the plain im2col path.
// M tiles become (x, y-block) blocks of 128 pixels so that a block never straddles an x row.
static int ky_enabled = -1, ky_mt = 2;
static bool try_ky(TcParams& p, int X, int Y, int Z, int ksize, int stride, int* upper_h) {
  if (ky_enabled < 0) {
    const char* e = getenv("COOCC_CONV_KY");
    ky_enabled = (e && e[0] == '0') ? 0 : 1;
    const char* m = getenv("COOCC_CONV_KY_MT");
    if (m) ky_mt = atoi(m);
  }
  if (!ky_enabled || ksize != 3 || stride != 1) return false;
  if (Z != 8 && Z != 16 && Z != 32 && Z != 64) return false;
  const int ny = 128 / Z;
  const int nyb = (Y + ny - 1) / ny;
  const int up = nyb * ny - Y + 1;
  if (up > 15) return false;
  if ((long long)nyb * ny * 100 > (long long)Y * 115) return false;       // > 15 % padded rows: not worth it
  const int ntn = (p.N + p.BN - 1) / p.BN;
  const int nmb = X * nyb;
  int MT = p.BN <= 128 ? ky_mt : 2;
  if (MT * (p.BN <= 128 ? 128 : 256) > 512) MT = 512 / (p.BN <= 128 ? 128 : 256);
  while (MT > 1 && ((nmb + MT - 1) / MT) * ntn < g_num_sms) MT >>= 1;
  if (nmb * ntn < g_num_sms) return false;                                 // under-filled: split-K path
  const int ext_bytes = ((ny + 2) * Z * 128 + 1023) / 1024 * 1024;
  const int a_entry = MT * ext_bytes;
  const int b_entry = (p.BN * 128 + 1023) / 1024 * 1024;
  int NA = 2;
  int NB = (kRingBytes - NA * a_entry) / b_entry;
  if (NB > 6) NB = 6;
  if (NB < 3) return false;
  if (kRingBytes - NA * a_entry - NB * b_entry >= a_entry) NA = 3;
  p.ky = 1;
  p.gX = X; p.gYZ = Y * Z; p.gZ = Z; p.ny = ny; p.nyb = nyb;
  p.ext_bytes = ext_bytes; p.NA = NA; p.NB = NB;
  p.MT = MT; p.NT = 1;
  p.hstride = p.BN <= 128 ? 128 : 256;
  p.nacc = (MT * p.hstride <= 256) ? 2 : 1;
  p.nstages = kStages;
  p.mc = 0;
  p.ksplit = 1;
  *upper_h = up;
  return true;
}

static int pick_bn(int n) {
  int bn = ((n + 15) / 16) * 16;
  if (bn <= kMaxBN) return bn;
  // largest multiple-of-32 tile <= 256 that divides n, else 256 with a masked tail
  for (int t = 256; t >= 64; t -= 32)
    if (n % t == 0) return t;
  return 256;
}

}  // namespace coocc

using namespace coocc;

static int check_desc(const coocc_conv_desc* d) {
  if (!d) return COOCC_ERR_ARG;
  if (d->ksize != 1 && d->ksize != 3) return COOCC_ERR_ARG;
  if (d->stride != 1 && d->stride != 2) return COOCC_ERR_ARG;
  if (d->dtype != COOCC_DTYPE_TF32 && d->dtype != COOCC_DTYPE_BF16) return COOCC_ERR_ARG;   // X3 is expanded by the callers
  if (d->X < 1 || d->Y < 1 || d->Z < 1 || d->Cin < 1 || d->Cout < 1) return COOCC_ERR_ARG;
  const int es = d->dtype == COOCC_DTYPE_TF32 ? 4 : 2;
  if (((long long)d->ldx * es) % 16 || ((long long)d->ldy * es) % 16) return COOCC_ERR_ALIGN;
  if (((long long)d->Cin * es) % 16) return COOCC_ERR_ALIGN;   // weight rows (taps*Cin) must be 16B multiples
  if (d->ldx < d->Cin || d->ldy < d->Cout) return COOCC_ERR_ARG;
  return 0;
}

static int fwd_impl(const coocc_conv_desc* d, const void* x, const void* w, void* yv, long long ldo,
                    const float* bias, int relu, float* stats, int accum, void* stream) {
  int rc = check_desc(d);
  if (rc) return rc;
  if ((rc = init_driver_api())) return rc;
  const int es = d->dtype == COOCC_DTYPE_TF32 ? 4 : 2;
  const int taps = d->ksize == 3 ? 27 : 1;
  const int oX = out_dim(d->X, d->ksize, d->stride), oY = out_dim(d->Y, d->ksize, d->stride),
            oZ = out_dim(d->Z, d->ksize, d->stride);
  float* y = reinterpret_cast<float*>(yv);
  TcParams p;
  memset(&p, 0, sizeof(p));
  p.mode = MODE_FPROP;
  p.M = oX * oY * oZ;
  p.N = d->Cout;
  p.BN = pick_bn(d->Cout);
  p.taps = taps;
  p.Kc = d->Cin;
  p.Cin = d->Cin;
  p.oX = oX; p.oY = oY; p.oZ = oZ;
  p.cstride = d->stride;
  p.lo = -(d->ksize / 2);
  p.out = y; p.ldc = ldo; p.bias = bias; p.relu = relu; p.stats = stats; p.accum = accum;
  p.out_bf16 = d->out_bf16 ? 1 : 0;
  if (p.out_bf16 && accum) return COOCC_ERR_ARG;
  p.ksplit = 1;
  const int bke = 128 / es;
  const bool plain = (d->ksize == 1 && d->stride == 1);
  p.a_im2col = plain ? 0 : 1;
  p.es = es;
  int up_h = 0;
  if (!false && try_ky(p, d->X, d->Y, d->Z, d->ksize, d->stride, &up_h)) {
    rc = make_tm_im2col(&p.tmA, x, es, d->X, d->Y, d->Z, d->Cin, d->ldx, -1, -1, 1, bke, (p.ny + 2) * d->Z, false,
                        up_h);
    if (rc) return rc;
    rc = make_tm_2d(&p.tmB, w, es, d->Cout, (long long)taps * d->Cin, (long long)taps * d->Cin, bke, p.BN, false);
    if (rc) return rc;
    const int ntiles = ((p.gX * p.nyb + p.MT - 1) / p.MT) * ((p.N + p.BN - 1) / p.BN);
    return es == 4 ? launch<4>(p, ntiles, (cudaStream_t)stream) : launch<2>(p, ntiles, (cudaStream_t)stream);
  }
  if (plain) {
    rc = make_tm_2d(&p.tmA, x, es, p.M, d->Cin, d->ldx, bke, kBM, false);
  } else {
    const int pad = d->ksize / 2;
    rc = make_tm_im2col(&p.tmA, x, es, d->X, d->Y, d->Z, d->Cin, d->ldx, -pad, pad - (d->ksize - 1),
                        d->stride, bke, kBM, false);
  }
  if (rc) return rc;
  float* stats_after = nullptr;
  if ((rc = prepare_split(p, taps * ((d->Cin + bke - 1) / bke), (cudaStream_t)stream, false))) return rc;
  SplitScratch sc;
  if ((rc = split_scratch_begin(p, sc, (cudaStream_t)stream))) return rc;
  // weight rows; with multicast each CTA of a pair fetches half of the N tile
  rc = make_tm_2d(&p.tmB, w, es, d->Cout, (long long)taps * d->Cin, (long long)taps * d->Cin, bke,
                  p.mc ? p.BN / 2 : p.BN, false);
  if (rc) return rc;
  if (p.ksplit > 1 && p.stats) { stats_after = p.stats; p.stats = nullptr; }
  const int ntm_ = (p.M + kBM * p.MT - 1) / (kBM * p.MT);
  const int ntiles = (p.mc ? (ntm_ + 1) / 2 : ntm_) * ((p.N + p.BN - 1) / p.BN) * p.ksplit;
  rc = es == 4 ? launch<4>(p, ntiles, (cudaStream_t)stream) : launch<2>(p, ntiles, (cudaStream_t)stream);
  if (!rc && stats_after) {
    dim3 grid((p.N + 31) / 32, p.M >= 4096 ? 64 : (p.M + 63) / 64);
    colstats_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(p.out, p.ldc, p.M, p.N, stats_after);
    rc = cudaGetLastError() == cudaSuccess ? 0 : COOCC_ERR_CUDA;
  }
  const int rc2 = split_scratch_end(p, sc, (cudaStream_t)stream);
  return rc ? rc : rc2;
}

// W[Cout][T][Cin] -> Wt[Cin][T][Cout] (bf16), 32 x 32 tiles through shared memory
__global__ void __launch_bounds__(256) transpose_w_kernel(const uint16_t* __restrict__ w, uint16_t* __restrict__ wt,
                                                          int Cout, int T, int Cin) {
  __shared__ uint16_t tile[32][34];
  const int t = blockIdx.z;
  const int ci0 = blockIdx.x * 32, co0 = blockIdx.y * 32;
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
  for (int r = ty; r < 32; r += 8) {
    const int co = co0 + r, ci = ci0 + tx;
    tile[r][tx] = (co < Cout && ci < Cin) ? w[((long long)co * T + t) * Cin + ci] : (uint16_t)0;
  }
  __syncthreads();
  for (int r = ty; r < 32; r += 8) {
    const int ci = ci0 + r, co = co0 + tx;
    if (ci < Cin && co < Cout) wt[((long long)ci * T + t) * Cout + co] = tile[tx][r];
  }
}

// dX[v, ci] = sum_{tap,co} dY[v - tap + pad, co] W[co, tap, ci]; stride-1 convolutions only
// (for stride 2 the caller scatters dY onto the input lattice first, see conv3d_dgrad docs).
// dx[r][c] += addend[r][c] on bf16 rows (fallback of the fused skip-gradient add for launches that cannot take the
// lean epilogue: split-K layers, ragged N tiles, strided data gradients)
__global__ void __launch_bounds__(256) add_rows_bf16_kernel(__nv_bfloat16* __restrict__ dx, long long ldo,
                                                            const __nv_bfloat16* __restrict__ addend, long long ld_add,
                                                            long long M, int C) {
  const long long total = M * C;
  for (long long i = blockIdx.x * 256LL + threadIdx.x; i < total; i += (long long)gridDim.x * 256LL) {
    const long long r = i / C;
    const int c = (int)(i - r * C);
    dx[r * ldo + c] = __float2bfloat16_rn(__bfloat162float(dx[r * ldo + c]) + __bfloat162float(addend[r * ld_add + c]));
  }
}
static int add_rows_bf16(void* dx, long long ldo, const void* addend, long long ld_add, long long M, int C,
                         cudaStream_t st) {
  long long blocks = (M * C + 255) / 256;
  if (blocks > 148 * 16) blocks = 148 * 16;
  add_rows_bf16_kernel<<<(int)blocks, 256, 0, st>>>(reinterpret_cast<__nv_bfloat16*>(dx), ldo,
                                                    reinterpret_cast<const __nv_bfloat16*>(addend), ld_add, M, C);
  return cudaGetLastError() == cudaSuccess ? 0 : COOCC_ERR_CUDA;
}

static int dgrad_impl(const coocc_conv_desc* d, const void* dy, const void* w, void* dxv, long long ldo,
                      int accum, void* stream, int cls = -1, const void* addend = nullptr, long long ld_add = 0) {
  int rc = check_desc(d);
  if (rc) return rc;
  if ((d->stride != 1) != (cls >= 0)) return COOCC_ERR_ARG;
  if ((rc = init_driver_api())) return rc;
  if (cls >= 0) {
    // Stride-2 data gradient, parity class cls = px*4 + py*2 + pz of the fine grid (d->X/Y/Z):
    //   dx[2i + p] = sum_t dy[(2i + p + 1 - t) / 2] W_t  over the taps t of matching parity, per axis
    //     p = 0: t = 1 (dy[i]);   p = 1: t = 2 (dy[i]) and t = 0 (dy[i + 1])      (3x3x3, padding 1)
    //     1x1x1: class 0 only, dx[2i] = dy[i] W_0
    // i.e. a small convolution of dy on its own (coarse) grid with 1 / 2 / 4 / 8 of the 27 taps whose output rows
    // are the class's voxels of dx: the eight classes together do the forward's FLOPs, where zero insertion
    // (coocc_dilate2 + a stride-1 dgrad, round 1) did eight times as many.
    const int es = d->dtype == COOCC_DTYPE_TF32 ? 4 : 2;
    const int ch = 128 / es;
    const int px = (cls >> 2) & 1, py = (cls >> 1) & 1, pz = cls & 1;
    if (d->ksize == 1 && cls != 0) return COOCC_ERR_ARG;
    if (cls > 8) return COOCC_ERR_ARG;
    const int oX = out_dim(d->X, d->ksize, 2), oY = out_dim(d->Y, d->ksize, 2), oZ = out_dim(d->Z, d->ksize, 2);
    TcParams p;
    memset(&p, 0, sizeof(p));
    p.mode = MODE_DGRAD;
    p.M = oX * oY * oZ;
    p.N = d->Cin;
    p.BN = ((d->Cin + ch - 1) / ch) * ch;
    if (p.BN > kMaxBN) {
      p.BN = kMaxBN;
      for (int t = 256; t >= ch; t -= ch)
        if (d->Cin % t == 0) { p.BN = t; break; }
    }
    p.tapmode = 1;
    int nt = 0;
    // tap list of one parity class: im2col offsets and weight taps
    const auto class_taps = [](int cx, int cy, int cz, unsigned char* off, unsigned char* wt) {
      const int par[3] = {cx, cy, cz};
      int offs[3][2], wts[3][2], cnt[3], n = 0;
      for (int a = 0; a < 3; ++a) {
        if (par[a] == 0) { cnt[a] = 1; offs[a][0] = 0; wts[a][0] = 1; }
        else { cnt[a] = 2; offs[a][0] = 0; wts[a][0] = 2; offs[a][1] = 1; wts[a][1] = 0; }
      }
      for (int a = 0; a < cnt[0]; ++a)
        for (int b = 0; b < cnt[1]; ++b)
          for (int c = 0; c < cnt[2]; ++c) {
            off[n] = (unsigned char)(offs[0][a] | (offs[1][b] << 2) | (offs[2][c] << 4));
            wt[n] = (unsigned char)((wts[0][a] * 3 + wts[1][b]) * 3 + wts[2][c]);
            ++n;
          }
      return n;
    };
    if (d->ksize == 1) {
      p.tap_off[0] = 0; p.tap_w[0] = 0; nt = 1;
    } else if (cls == 8) {
      // every class in one launch, most taps first: 7 | 3 5 6 | 1 2 4 | 0
      static const int order[8] = {7, 3, 5, 6, 1, 2, 4, 0};
      p.ncls = 8;
      for (int i = 0; i < 8; ++i) {
        const int c = order[i];
        p.cls_par[i] = (unsigned char)c;
        p.cls_ntaps[i] = (unsigned char)class_taps((c >> 2) & 1, (c >> 1) & 1, c & 1, p.cls_off[i], p.cls_w[i]);
      }
      nt = 8;                       // (upper bound: sizes the split / MT heuristics; tiles use cls_ntaps)
    } else {
      nt = class_taps(px, py, pz, p.tap_off, p.tap_w);
    }
    p.taps = nt;
    p.Kc = d->Cout;
    p.Cin = d->Cin;
    p.oX = oX; p.oY = oY; p.oZ = oZ;
    p.cstride = 1;
    p.lo = 0;
    p.omap = 1; p.opx = px; p.opy = py; p.opz = pz; p.fX = d->X; p.fY = d->Y; p.fZ = d->Z;
    p.out = reinterpret_cast<float*>(dxv); p.ldc = ldo; p.accum = accum;
    p.out_bf16 = d->out_bf16 ? 1 : 0;
    if (p.out_bf16 && accum) return COOCC_ERR_ARG;
    p.es = es;
    const int wtaps = d->ksize == 3 ? 27 : 1;
    if (d->ksize == 1) {
      p.a_im2col = 0;
      rc = make_tm_2d(&p.tmA, dy, es, p.M, d->Cout, d->ldy, ch, kBM, false);
    } else {
      p.a_im2col = 1;
      rc = make_tm_im2col(&p.tmA, dy, es, oX, oY, oZ, d->Cout, d->ldy, 0, 0, 1, ch, kBM, false);
    }
    if (rc) return rc;
    rc = make_tm_2d(&p.tmB, w, es, d->Cout, (long long)wtaps * d->Cin, (long long)wtaps * d->Cin, ch, ch, true);
    if (rc) return rc;
    if ((rc = prepare_split(p, nt * ((d->Cout + ch - 1) / ch), (cudaStream_t)stream, false))) return rc;
    p.ksplit = 1;            // (split-K accumulates into a contiguous zeroed [M, N] block; the class rows are scattered)
    p.mc = 0;
    const int ntm_ = (p.M + kBM * p.MT - 1) / (kBM * p.MT);
    const int ntiles = ntm_ * ((p.N + p.BN - 1) / p.BN) * (p.ncls ? p.ncls : 1);
    return es == 4 ? launch<4>(p, ntiles, (cudaStream_t)stream) : launch<2>(p, ntiles, (cudaStream_t)stream);
  }
  const int es = d->dtype == COOCC_DTYPE_TF32 ? 4 : 2;
  const int taps = d->ksize == 3 ? 27 : 1;
  const int ch = 128 / es;
  float* dx = reinterpret_cast<float*>(dxv);
  TcParams p;
  memset(&p, 0, sizeof(p));
  p.mode = MODE_DGRAD;
  p.M = d->X * d->Y * d->Z;
  p.N = d->Cin;
  // MN-major B tiles are loaded in chunks of `ch` input channels
  p.BN = ((d->Cin + ch - 1) / ch) * ch;
  if (p.BN > kMaxBN) {
    p.BN = kMaxBN;
    for (int t = 256; t >= ch; t -= ch)
      if (d->Cin % t == 0) { p.BN = t; break; }
  }
  p.taps = taps;
  p.Kc = d->Cout;
  p.Cin = d->Cin;
  p.oX = d->X; p.oY = d->Y; p.oZ = d->Z;
  p.cstride = 1;
  p.lo = -(d->ksize / 2);
  p.out = dx; p.ldc = ldo; p.bias = nullptr; p.relu = 0; p.stats = nullptr; p.accum = accum;
  p.out_bf16 = d->out_bf16 ? 1 : 0;
  if (p.out_bf16 && accum) return COOCC_ERR_ARG;
  p.ksplit = 1;
  const bool plain = (d->ksize == 1);
  p.a_im2col = plain ? 0 : 1;
  p.es = es;
  int up_h = 0;
  if (try_ky(p, d->X, d->Y, d->Z, d->ksize, 1, &up_h)) {
    rc = make_tm_im2col(&p.tmA, dy, es, d->X, d->Y, d->Z, d->Cout, d->ldy, -1, -1, 1, ch, (p.ny + 2) * d->Z, false,
                        up_h);
    if (rc) return rc;
    const int ntiles = ((p.gX * p.nyb + p.MT - 1) / p.MT) * ((p.N + p.BN - 1) / p.BN);
    const bool fuse_add = addend && p.out_bf16 && !accum && (ldo % 8) == 0 && (ld_add % 8) == 0 && (p.BN % 32) == 0 &&
                          (p.N % p.BN) == 0;
    if (fuse_add) { p.addend = addend; p.ld_add = ld_add; }
    const auto finish = [&](int rc_) {
      if (!rc_ && addend && !fuse_add) rc_ = add_rows_bf16(dxv, ldo, addend, ld_add, p.M, d->Cin, (cudaStream_t)stream);
      return rc_;
    };
    // bf16: run the big-grid data gradients on a transposed copy of the weights (K-major B operand, the forward's
    // configuration).  Reading W in place as an MN-major operand kept the tensor pipe at 67 % active against 79 % for
    // the forward of the same layer (profiles/r02_conv_full.summary.txt); the copy is a few MB per layer.
    static int wt_enabled = -1;
    if (wt_enabled < 0) {
      const char* e = getenv("COOCC_DGRAD_WT");
      wt_enabled = (e && e[0] == '0') ? 0 : 1;
    }
    const long long welts = (long long)d->Cout * taps * d->Cin;
    // (measured: 128->128 on the 200x200x16 grid 1.71 -> 1.53 ms for four launches; no gain for 256-wide N tiles)
    if (es == 2 && wt_enabled && p.BN <= 128 && welts <= (8ll << 20) && (d->Cout % 8) == 0) {
      uint16_t* wt = nullptr;
      if (cudaMallocAsync(reinterpret_cast<void**>(&wt), (size_t)welts * 2, (cudaStream_t)stream) != cudaSuccess)
        return COOCC_ERR_CUDA;
      dim3 grid((d->Cin + 31) / 32, (d->Cout + 31) / 32, taps);
      transpose_w_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(reinterpret_cast<const uint16_t*>(w), wt, d->Cout, taps,
                                                                d->Cin);
      p.b_kmajor = 1;
      p.Cin = d->Cout;                   // B K-offset per tap in the transposed copy
      rc = make_tm_2d(&p.tmB, wt, es, d->Cin, (long long)taps * d->Cout, (long long)taps * d->Cout, ch, p.BN, false);
      if (!rc) rc = launch<2>(p, ntiles, (cudaStream_t)stream);
      cudaFreeAsync(wt, (cudaStream_t)stream);
      return finish(rc);
    }
    rc = make_tm_2d(&p.tmB, w, es, d->Cout, (long long)taps * d->Cin, (long long)taps * d->Cin, ch, ch, true);
    if (rc) return rc;
    return finish(es == 4 ? launch<4>(p, ntiles, (cudaStream_t)stream) : launch<2>(p, ntiles, (cudaStream_t)stream));
  }
  if (plain) {
    rc = make_tm_2d(&p.tmA, dy, es, p.M, d->Cout, d->ldy, ch, kBM, false);
  } else {
    const int pad = d->ksize / 2;
    rc = make_tm_im2col(&p.tmA, dy, es, d->X, d->Y, d->Z, d->Cout, d->ldy, -pad,
                        pad - (d->ksize - 1), 1, ch, kBM, false);
  }
  if (rc) return rc;
  rc = make_tm_2d(&p.tmB, w, es, d->Cout, (long long)taps * d->Cin, (long long)taps * d->Cin, ch, ch, true);
  if (rc) return rc;
  p.es = es;
  if ((rc = prepare_split(p, taps * ((d->Cout + ch - 1) / ch), (cudaStream_t)stream, false))) return rc;
  const bool fuse_add = addend && p.ksplit == 1 && p.out_bf16 && !accum && (ldo % 8) == 0 && (ld_add % 8) == 0 &&
                        (p.BN % 32) == 0 && (p.N % p.BN) == 0;
  if (fuse_add) { p.addend = addend; p.ld_add = ld_add; }
  SplitScratch sc;
  if ((rc = split_scratch_begin(p, sc, (cudaStream_t)stream))) return rc;
  const int ntm_ = (p.M + kBM * p.MT - 1) / (kBM * p.MT);
  const int ntiles = (p.mc ? (ntm_ + 1) / 2 : ntm_) * ((p.N + p.BN - 1) / p.BN) * p.ksplit;
  rc = es == 4 ? launch<4>(p, ntiles, (cudaStream_t)stream) : launch<2>(p, ntiles, (cudaStream_t)stream);
  const int rc2 = split_scratch_end(p, sc, (cudaStream_t)stream);
  if (!rc && !rc2 && addend && !fuse_add) rc = add_rows_bf16(dxv, ldo, addend, ld_add, p.M, d->Cin, (cudaStream_t)stream);
  return rc ? rc : rc2;
}

// split-K factor of a wgrad launch: tiles are dealt round-robin to the persistent CTAs, so pick the
// factor whose tile count fills whole waves (297 tiles on 148 SMs cost 3 waves, 296 cost 2)
static int pick_wgrad_ksplit(int base_tiles, int nvb) {
  int ks = 1;
  // every split keeps >= 8 k-blocks so the pipeline fill / atomic epilogue stay amortised
  int max_ks = nvb / 8;
  // (1x1x1 layers have a single base tile: 48 splits left 100 SMs idle on an HBM-bound reduction over 640 k voxels,
  // ncu r02: grid 48, 129 us for 246 MB)
  int cap = 48;
  if ((long long)base_tiles * cap < g_num_sms) cap = (g_num_sms + base_tiles - 1) / base_tiles;
  if (max_ks > cap) max_ks = cap;
  if (max_ks < 1) max_ks = 1;
  double best = -1.0;
  for (int c = 1; c <= max_ks; ++c) {
    const long long tiles = (long long)base_tiles * c;
    const long long waves = (tiles + g_num_sms - 1) / g_num_sms;
    double eff = (double)tiles / (double)(waves * g_num_sms);
    if (tiles < g_num_sms) eff *= 0.9;                 // prefer filling the machine at least once
    eff -= 0.002 * c;                                  // slight preference for fewer atomic passes
    if (eff > best) { best = eff; ks = c; }
  }
  return ks;
}

// dW[co, tap, ci] += sum_v dY[v, co] X[v*s + tap - pad, ci]   (dw must be zero-filled by the caller)
static int wgrad_impl(const coocc_conv_desc* d, const void* x, const void* dy, float* dw, void* stream) {
  int rc = check_desc(d);
  if (rc) return rc;
  if ((rc = init_driver_api())) return rc;
  const int es = d->dtype == COOCC_DTYPE_TF32 ? 4 : 2;
  const int taps = d->ksize == 3 ? 27 : 1;
  const int ch = 128 / es;
  const int oX = out_dim(d->X, d->ksize, d->stride), oY = out_dim(d->Y, d->ksize, d->stride),
            oZ = out_dim(d->Z, d->ksize, d->stride);
  const long long Vo = (long long)oX * oY * oZ;
  TcParams p;
  memset(&p, 0, sizeof(p));
  p.mode = MODE_WGRAD;
  p.mc = 0;
  p.es = es;
  p.M = d->Cout;
  p.N = d->Cin;
  p.taps = taps;
  p.Cin = d->Cin;
  p.oX = oX; p.oY = oY; p.oZ = oZ;
  p.cstride = d->stride;
  p.lo = -(d->ksize / 2);
  p.out = dw; p.ldc = d->Cin;
  p.nvb = (int)((Vo + ch - 1) / ch);
  // ky-fused wgrad (bf16): 128 couts x (3 ky taps x <= 128 cins) accumulators; per 128-pixel block one
  // dY^T tile and one y-extended X tile (see try_ky) -> 170 MACs per operand element pulled from L2
  // versus 96 (NT = 3) / 128 (MT = 2) of the plain shapes below
  {
    TcParams q = p;
    q.BN = ((d->Cin + ch - 1) / ch) * ch;
    if (q.BN > 128) q.BN = 128;
    int up_h = 0;
    if (es == 2 && try_ky(q, d->X, d->Y, d->Z, d->ksize, d->stride, &up_h)) {
      p = q;
      p.MT = 1; p.NT = 3; p.hstride = 128; p.nacc = 1;
      p.nvb = p.gX * p.nyb;
      const int stage = (kBM / ch) * 128 * 128 + (p.BN / ch) * p.ext_bytes;
      p.nstages = kRingBytes / stage;
      if (p.nstages > kStages) p.nstages = kStages;
      const int base_tiles = ((p.M + kBM - 1) / kBM) * ((p.N + p.BN - 1) / p.BN) * 9;
      p.ksplit = pick_wgrad_ksplit(base_tiles, p.nvb);
      rc = make_tm_3d(&p.tmA, dy, es, d->X, (long long)d->Y * d->Z, d->Cout, d->ldy, ch, 128, true);
      if (rc) return rc;
      rc = make_tm_im2col(&p.tmB, x, es, d->X, d->Y, d->Z, d->Cin, d->ldx, -1, -1, 1, ch, (p.ny + 2) * d->Z, true,
                          up_h);
      if (rc) return rc;
      p.a_im2col = 1;
      return launch<2>(p, base_tiles * p.ksplit, (cudaStream_t)stream);
    }
  }
  // CTA tile shape (fewer operand bytes per MAC, like fprop):
  //   Cout >= 256          : two cout tiles share one X tile        (MT = 2, BN <= 256)
  //   Cout <= 128, 27 taps : three taps share one dY^T tile          (NT = 3, BN <= 128)
  static int wg_shape = -1;
  if (wg_shape < 0) {
    const char* e = getenv("COOCC_WGRAD_SHAPES");
    wg_shape = (e && e[0] == '0') ? 0 : 1;
  }
  p.MT = 1;
  p.NT = 1;
  int bn_cap = kMaxBN;
  if (wg_shape && p.nvb >= 64) {
    if (d->Cout >= 256) p.MT = 2;
    else if (taps == 27) { p.NT = 3; bn_cap = 128; }
  }
  p.BN = ((d->Cin + ch - 1) / ch) * ch;
  if (p.BN > bn_cap) {
    p.BN = bn_cap;
    for (int t = bn_cap; t >= ch; t -= ch)
      if (d->Cin % t == 0) { p.BN = t; break; }
  }
  p.hstride = p.BN <= 128 ? 128 : 256;
  p.nacc = (p.MT * p.NT * p.hstride <= 256) ? 2 : 1;
  {
    const int stage = kABytes * p.MT + p.NT * ((p.BN * 128 + 1023) / 1024 * 1024);
    p.nstages = kRingBytes / stage;
    const char* e = getenv("COOCC_CONV_DEEP_RING");
    const int cap = (taps == 1 && e && e[0] == '1') ? kMaxBRing : kStages;        // (opt-in, see prepare_split)
    if (p.nstages > cap) p.nstages = cap;
  }
  const int base_tiles = ((p.M + kBM * p.MT - 1) / (kBM * p.MT)) * ((p.N + p.BN - 1) / p.BN) * ((taps + p.NT - 1) / p.NT);
  // split-K factor: tiles are dealt round-robin to the persistent CTAs, so pick the factor whose
  // tile count fills whole waves (297 tiles on 148 SMs cost 3 waves, 296 cost 2)
  const int ks = pick_wgrad_ksplit(base_tiles, p.nvb);
  p.ksplit = ks;
  const bool plain = (d->ksize == 1 && d->stride == 1);
  p.a_im2col = plain ? 0 : 1;
  rc = make_tm_2d(&p.tmA, dy, es, Vo, d->Cout, d->ldy, ch, ch, true);
  if (rc) return rc;
  if (plain) {
    rc = make_tm_2d(&p.tmB, x, es, Vo, d->Cin, d->ldx, ch, ch, true);
  } else {
    const int pad = d->ksize / 2;
    rc = make_tm_im2col(&p.tmB, x, es, d->X, d->Y, d->Z, d->Cin, d->ldx, -pad, pad - (d->ksize - 1),
                        d->stride, ch, ch, true);
  }
  if (rc) return rc;
  const int ntiles = base_tiles * ks;
  return es == 4 ? launch<4>(p, ntiles, (cudaStream_t)stream) : launch<2>(p, ntiles, (cudaStream_t)stream);
}

// ------------------------------------------------------------------------------------------
// COOCC_DTYPE_TF32X3: fp32-accurate mode.  Every fp32 operand is split as x = hi + lo with
// hi = round-to-nearest tf32(x) and lo = x - hi (exact in fp32); the product is evaluated as
// hi*hi + hi*lo + lo*hi in three tensor-core passes accumulating in fp32 (the dropped lo*lo term
// is ~2^-22 relative).  Used for the 1e-3 parity runs; the fast modes are single-pass.
// ------------------------------------------------------------------------------------------
namespace coocc {
__global__ void split_tf32_kernel(const float* __restrict__ x, long long ldx, long long rows, int cols,
                                  float* __restrict__ hi, float* __restrict__ lo, long long ldo) {
  const long long n = rows * ldo;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n;
       i += (long long)gridDim.x * blockDim.x) {
    const long long r = i / ldo;
    const int c = (int)(i % ldo);
    float v = 0.f;
    if (c < cols) v = x[r * ldx + c];
    uint32_t h;
    asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(h) : "f"(v));
    const float hf = __uint_as_float(h);
    hi[i] = hf;
    lo[i] = v - hf;
  }
}

struct SplitBuf {
  float* hi = nullptr;
  float* lo = nullptr;
  long long ld = 0;
};

static int split_rows(const void* x, long long ldx, long long rows, int cols, SplitBuf* out, cudaStream_t st) {
  const long long ld = ((long long)cols + 3) / 4 * 4;
  const size_t bytes = (size_t)rows * ld * sizeof(float);
  float* base = nullptr;
  if (cudaMallocAsync(reinterpret_cast<void**>(&base), 2 * bytes, st) != cudaSuccess) return COOCC_ERR_CUDA;
  out->hi = base;
  out->lo = base + rows * ld;
  out->ld = ld;
  const long long n = rows * ld;
  int blocks = (int)((n + 255) / 256);
  if (blocks > 148 * 16) blocks = 148 * 16;
  if (blocks < 1) blocks = 1;
  split_tf32_kernel<<<blocks, 256, 0, st>>>(reinterpret_cast<const float*>(x), ldx, rows, cols, out->hi, out->lo, ld);
  return cudaGetLastError() == cudaSuccess ? 0 : COOCC_ERR_CUDA;
}
static void split_free(SplitBuf* b, cudaStream_t st) {
  if (b->hi) cudaFreeAsync(b->hi, st);
  b->hi = b->lo = nullptr;
}
}  // namespace coocc

// tuning hook (benchmarks / A-B tests): ky = 0 / 1 disables / enables the ky-fused path,
// ky_mt = 128-pixel blocks per CTA tile for N <= 128 layers (1, 2 or 4); negative = keep
extern "C" int coocc_conv_tune(int ky, int ky_mt_) {
  if (coocc::ky_enabled < 0) { coocc::ky_enabled = 1; }
  if (ky >= 0) coocc::ky_enabled = ky ? 1 : 0;
  if (ky_mt_ == 1 || ky_mt_ == 2 || ky_mt_ == 4) coocc::ky_mt = ky_mt_;
  return 0;
}

extern "C" int coocc_conv3d_fwd(const coocc_conv_desc* d, const void* x, const void* w, void* y,
                                long long ldo, const float* bias, int relu, float* stats,
                                void* stream) {
  if (!d) return COOCC_ERR_ARG;
  if (d->dtype != COOCC_DTYPE_TF32X3) return fwd_impl(d, x, w, y, ldo, bias, relu, stats, 0, stream);
  cudaStream_t st = (cudaStream_t)stream;
  const int taps = d->ksize == 3 ? 27 : 1;
  SplitBuf xs, ws;
  int rc = split_rows(x, d->ldx, (long long)d->X * d->Y * d->Z, d->Cin, &xs, st);
  if (!rc) rc = split_rows(w, (long long)taps * d->Cin, d->Cout, taps * d->Cin, &ws, st);
  coocc_conv_desc t = *d;
  t.dtype = COOCC_DTYPE_TF32;
  t.ldx = xs.ld;
  if (!rc) rc = fwd_impl(&t, xs.hi, ws.hi, y, ldo, nullptr, 0, nullptr, 0, stream);
  if (!rc) rc = fwd_impl(&t, xs.hi, ws.lo, y, ldo, nullptr, 0, nullptr, 1, stream);
  if (!rc) rc = fwd_impl(&t, xs.lo, ws.hi, y, ldo, bias, relu, stats, 1, stream);
  split_free(&xs, st);
  split_free(&ws, st);
  return rc;
}

extern "C" int coocc_conv3d_dgrad_add(const coocc_conv_desc* d, const void* dy, const void* w, void* dx, long long ldo,
                                      const void* addend, long long ld_add, void* stream) {
  if (!d || !addend) return COOCC_ERR_ARG;
  if (d->dtype != COOCC_DTYPE_BF16 || !d->out_bf16 || d->stride != 1) return COOCC_ERR_ARG;
  return dgrad_impl(d, dy, w, dx, ldo, 0, stream, -1, addend, ld_add);
}

extern "C" int coocc_conv3d_dgrad(const coocc_conv_desc* d, const void* dy, const void* w, void* dx,
                                  long long ldo, void* stream) {
  if (!d) return COOCC_ERR_ARG;
  cudaStream_t st = (cudaStream_t)stream;
  const int taps = d->ksize == 3 ? 27 : 1;
  if (d->stride == 2) {
    // d->X/Y/Z = the (fine) input grid, dy lives on the strided output grid: one launch per parity class
    const int ncls = d->ksize == 3 ? 8 : 1;
    const long long Vf = (long long)d->X * d->Y * d->Z;
    const long long Vo = (long long)out_dim(d->X, d->ksize, 2) * out_dim(d->Y, d->ksize, 2) * out_dim(d->Z, d->ksize, 2);
    int rc = 0;
    if (d->ksize == 1) {       // only the even voxels receive a gradient
      const size_t eb = d->out_bf16 ? 2 : 4;
      if (cudaMemset2DAsync(dx, (size_t)ldo * eb, 0, (size_t)d->Cin * eb, (size_t)Vf, st) != cudaSuccess) return COOCC_ERR_CUDA;
    }
    static int one_launch = -1;
    if (one_launch < 0) {
      const char* e = getenv("COOCC_DGRAD_S2_ONE");
      one_launch = (e && e[0] == '0') ? 0 : 1;
    }
    if (d->dtype != COOCC_DTYPE_TF32X3) {
      if (ncls == 8 && one_launch) return dgrad_impl(d, dy, w, dx, ldo, 0, stream, 8);
      for (int c = 0; c < ncls && !rc; ++c) rc = dgrad_impl(d, dy, w, dx, ldo, 0, stream, c);
      return rc;
    }
    SplitBuf gs, ws;
    rc = split_rows(dy, d->ldy, Vo, d->Cout, &gs, st);
    if (!rc) rc = split_rows(w, (long long)taps * d->Cin, d->Cout, taps * d->Cin, &ws, st);
    coocc_conv_desc t = *d;
    t.dtype = COOCC_DTYPE_TF32;
    t.ldy = gs.ld;
    for (int c = (ncls == 8 && one_launch) ? 8 : 0; c < (ncls == 8 && one_launch ? 9 : ncls) && !rc; ++c) {
      rc = dgrad_impl(&t, gs.hi, ws.hi, dx, ldo, 0, stream, c);
      if (!rc) rc = dgrad_impl(&t, gs.hi, ws.lo, dx, ldo, 1, stream, c);
      if (!rc) rc = dgrad_impl(&t, gs.lo, ws.hi, dx, ldo, 1, stream, c);
    }
    split_free(&gs, st);
    split_free(&ws, st);
    return rc;
  }
  if (d->dtype != COOCC_DTYPE_TF32X3) return dgrad_impl(d, dy, w, dx, ldo, 0, stream);
  SplitBuf gs, ws;
  int rc = split_rows(dy, d->ldy, (long long)d->X * d->Y * d->Z, d->Cout, &gs, st);
  if (!rc) rc = split_rows(w, (long long)taps * d->Cin, d->Cout, taps * d->Cin, &ws, st);
  coocc_conv_desc t = *d;
  t.dtype = COOCC_DTYPE_TF32;
  t.ldy = gs.ld;
  if (!rc) rc = dgrad_impl(&t, gs.hi, ws.hi, dx, ldo, 0, stream);
  if (!rc) rc = dgrad_impl(&t, gs.hi, ws.lo, dx, ldo, 1, stream);
  if (!rc) rc = dgrad_impl(&t, gs.lo, ws.hi, dx, ldo, 1, stream);
  split_free(&gs, st);
  split_free(&ws, st);
  return rc;
}

extern "C" int coocc_conv3d_wgrad(const coocc_conv_desc* d, const void* x, const void* dy, float* dw,
                                  void* stream) {
  if (!d) return COOCC_ERR_ARG;
  if (d->dtype != COOCC_DTYPE_TF32X3) return wgrad_impl(d, x, dy, dw, stream);
  cudaStream_t st = (cudaStream_t)stream;
  const int pad = d->ksize / 2;
  const long long Vo = (long long)((d->X + 2 * pad - d->ksize) / d->stride + 1) *
                       ((d->Y + 2 * pad - d->ksize) / d->stride + 1) *
                       ((d->Z + 2 * pad - d->ksize) / d->stride + 1);
  SplitBuf xs, gs;
  int rc = split_rows(x, d->ldx, (long long)d->X * d->Y * d->Z, d->Cin, &xs, st);
  if (!rc) rc = split_rows(dy, d->ldy, Vo, d->Cout, &gs, st);
  coocc_conv_desc t = *d;
  t.dtype = COOCC_DTYPE_TF32;
  t.ldx = xs.ld;
  t.ldy = gs.ld;
  if (!rc) rc = wgrad_impl(&t, xs.hi, gs.hi, dw, stream);
  if (!rc) rc = wgrad_impl(&t, xs.hi, gs.lo, dw, stream);
  if (!rc) rc = wgrad_impl(&t, xs.lo, gs.hi, dw, stream);
  split_free(&xs, st);
  split_free(&gs, st);
  return rc;
}
