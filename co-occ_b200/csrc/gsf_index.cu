// gsf_index.cu -- GSFusion index pipeline (integer-exact, no autograd):
//   pack      : strided [C,X,Y,Z] fp32 -> NDHWC rows + occupancy flag (sum over channels != 0)
//   compact   : flags -> ordered occupied-voxel list + voxel->rank table (torch.nonzero order)
//   fps       : furthest point sampling of the occupied voxels, one thread-block cluster
//   rep_topk  : exact K nearest keys of every representative, window search on the voxel lattice
//   ball_assign: ball-query propagation + "last writer wins" assignment as an atomicMax
//
// Replaces P/coocc/fuser/bifuser_n.py:38-125,129-135 and the two CUDA ops it calls
// (mmdet3d/ops/furthest_point_sample/src/furthest_point_sample_cuda.cu:26-141,
//  mmdet3d/ops/ball_query/src/ball_query_cuda.cu:11-54).
//
// All coordinates are integer voxel indices, so every distance is an exact integer:
// `val < 13.3` <=> d2 <= 176 and `d2 < 6*6` are evaluated in int32.
#include <cooperative_groups.h>
#include <cuda_runtime.h>
#include <stdint.h>

#include "../../include/coocc_b200.h"

namespace cg = cooperative_groups;

namespace coocc {

// ------------------------------------------------------------------------------------------
// pack: transpose a strided channel-major grid into NDHWC rows and flag occupied voxels
// ------------------------------------------------------------------------------------------
// grid: (tiles along the fast axis, other axis a, other axis b); block 32x8.
// Each block owns 32 voxels along the axis with the smallest stride and loops over all channels.
__global__ void __launch_bounds__(256) pack_kernel(const float* __restrict__ src, long long sC,
                                                   long long sX, long long sY, long long sZ, int C,
                                                   int X, int Y, int Z, int fast, float* __restrict__ dst,
                                                   long long ldo, uint8_t* __restrict__ flags) {
  __shared__ float tile[32][33];
  __shared__ float rowsum[8][32];
  const int tx = threadIdx.x, ty = threadIdx.y;
  // decode block -> (fast-axis start, a, b)
  const int dims[3] = {X, Y, Z};
  const int a_ax = (fast == 0) ? 1 : 0;
  const int b_ax = (fast == 2) ? 1 : 2;
  const int f0 = blockIdx.x * 32;
  const int ca = blockIdx.y, cb = blockIdx.z;
  int c3[3];
  c3[a_ax] = ca;
  c3[b_ax] = cb;
  const long long strides[3] = {sX, sY, sZ};
  const int nf = dims[fast];
  // voxel handled by lane `l` of this block: fast coordinate f0 + l
  float acc[4] = {0.f, 0.f, 0.f, 0.f};   // ty handles voxels ty*4 .. ty*4+3 in the write phase
  for (int cbase = 0; cbase < C; cbase += 32) {
    // read: warp ty reads channels cbase + ty*4 + j, lane = position along the fast axis
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int c = cbase + ty * 4 + j;
      const int f = f0 + tx;
      float v = 0.f;
      if (c < C && f < nf) {
        c3[fast] = f;
        v = src[c * sC + c3[0] * strides[0] + c3[1] * strides[1] + c3[2] * strides[2]];
      }
      tile[ty * 4 + j][tx] = v;
    }
    __syncthreads();
    // write: warp ty writes voxel rows ty*4 + j, lane = channel
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int l = ty * 4 + j;
      const int f = f0 + l;
      const float v = tile[tx][l];
      if (f < nf && cbase + tx < C) {
        c3[fast] = f;
        const long long vox = ((long long)c3[0] * Y + c3[1]) * Z + c3[2];
        dst[vox * ldo + cbase + tx] = v;
      }
      float s = v;
#pragma unroll
      for (int o = 16; o >= 1; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
      acc[j] += s;
    }
    __syncthreads();
  }
  if (tx == 0) {
#pragma unroll
    for (int j = 0; j < 4; ++j) rowsum[ty][j] = acc[j];
  }
  __syncthreads();
  if (ty == 0) {
    const int f = f0 + tx;
    if (f < nf) {
      c3[fast] = f;
      const long long vox = ((long long)c3[0] * Y + c3[1]) * Z + c3[2];
      flags[vox] = rowsum[tx >> 2][tx & 3] != 0.f ? 1 : 0;
    }
  }
}


// Second-generation pack (the one normally launched): 32 voxels x 128 channels per block pass -- every thread has 16
// independent 128-byte-row loads in flight (the first version: 4) -- and, when the fast axis and the next one are
// contiguous in memory (both upstream layouts: [C,Z,X,Y] and [C,Z,Y,X]), the two are walked as one flat axis so that
// no tile is partly empty (200 = 6 x 32 + 8).  fast / mid / slow = spatial axes by increasing stride; a block handles
// flat positions f0 .. f0+31 of (mid, fast) at one `slow` coordinate.
__global__ void __launch_bounds__(256) pack_flat_kernel(const float* __restrict__ src, long long sC, long long sF,
                                                        long long sS, int C, int nF, int nFast, int fast, int mid,
                                                        int slow, int Y, int Z, float* __restrict__ dst, long long ldo,
                                                        uint8_t* __restrict__ flags) {
  __shared__ float tile[128][33];
  __shared__ float part[8][32];
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
  const int f = blockIdx.x * 32 + tx;                  // this lane's flat position in the read phase
  const long long base = (long long)blockIdx.y * sS;
  float sum = 0.f;
  for (int cbase = 0; cbase < C; cbase += 128) {
    float v[16];
#pragma unroll
    for (int j = 0; j < 16; ++j) {
      const int c = cbase + ty * 16 + j;
      v[j] = (c < C && f < nF) ? src[(long long)c * sC + base + (long long)f * sF] : 0.f;
    }
#pragma unroll
    for (int j = 0; j < 16; ++j) {
      tile[ty * 16 + j][tx] = v[j];
      sum += v[j];
    }
    __syncthreads();
    // write: warp ty writes voxel rows ty*4 .. ty*4+3, lane = channel (+32, +64, +96)
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int l = ty * 4 + j;
      const int fl = blockIdx.x * 32 + l;
      if (fl < nF) {
        int c3[3];
        c3[fast] = fl % nFast;
        c3[mid] = fl / nFast;
        c3[slow] = blockIdx.y;
        float* row = dst + (((long long)c3[0] * Y + c3[1]) * Z + c3[2]) * ldo + cbase;
#pragma unroll
        for (int k = 0; k < 4; ++k)
          if (cbase + tx + 32 * k < C) row[tx + 32 * k] = tile[tx + 32 * k][l];
      }
    }
    __syncthreads();
  }
  part[ty][tx] = sum;
  __syncthreads();
  if (ty == 0 && f < nF) {
    float t = 0.f;
#pragma unroll
    for (int j = 0; j < 8; ++j) t += part[j][tx];
    int c3[3];
    c3[fast] = f % nFast;
    c3[mid] = f / nFast;
    c3[slow] = blockIdx.y;
    flags[((long long)c3[0] * Y + c3[1]) * Z + c3[2]] = t != 0.f ? 1 : 0;
  }
}

// ------------------------------------------------------------------------------------------
// compact: ordered stream compaction of the flags (3 passes, deterministic)
// ------------------------------------------------------------------------------------------
constexpr int kChunk = 2048;   // voxels per block in passes 1 and 3 (256 threads x 8)

__global__ void __launch_bounds__(256) count_kernel(const uint8_t* __restrict__ flags, int V,
                                                    int* __restrict__ chunk_cnt) {
  const int base = blockIdx.x * kChunk;
  int c = 0;
  for (int i = threadIdx.x; i < kChunk; i += 256) {
    const int v = base + i;
    c += (v < V && flags[v]) ? 1 : 0;
  }
  c = __reduce_add_sync(0xffffffffu, c);
  __shared__ int ws[8];
  if ((threadIdx.x & 31) == 0) ws[threadIdx.x >> 5] = c;
  __syncthreads();
  if (threadIdx.x == 0) {
    int s = 0;
    for (int i = 0; i < 8; ++i) s += ws[i];
    chunk_cnt[blockIdx.x] = s;
  }
}

// single block: exclusive scan of the chunk counts in place; total -> *count
__global__ void __launch_bounds__(1024) scan_kernel(int* __restrict__ chunk_cnt, int nchunks,
                                                    int* __restrict__ count) {
  __shared__ int ws[32];
  __shared__ int carry;
  if (threadIdx.x == 0) carry = 0;
  __syncthreads();
  for (int base = 0; base < nchunks; base += 1024) {
    const int i = base + threadIdx.x;
    const int v = i < nchunks ? chunk_cnt[i] : 0;
    int s = v;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      const int t = __shfl_up_sync(0xffffffffu, s, o);
      if ((threadIdx.x & 31) >= o) s += t;
    }
    if ((threadIdx.x & 31) == 31) ws[threadIdx.x >> 5] = s;
    __syncthreads();
    if (threadIdx.x < 32) {
      int w = ws[threadIdx.x];
#pragma unroll
      for (int o = 1; o < 32; o <<= 1) {
        const int t = __shfl_up_sync(0xffffffffu, w, o);
        if (threadIdx.x >= o) w += t;
      }
      ws[threadIdx.x] = w;
    }
    __syncthreads();
    const int warp_off = (threadIdx.x >> 5) ? ws[(threadIdx.x >> 5) - 1] : 0;
    const int incl = s + warp_off + carry;
    if (i < nchunks) chunk_cnt[i] = incl - v;
    __syncthreads();
    if (threadIdx.x == 1023) carry = incl;
    __syncthreads();
  }
  if (threadIdx.x == 0) *count = carry;
}

__global__ void __launch_bounds__(256) scatter_kernel(const uint8_t* __restrict__ flags, int V,
                                                      const int* __restrict__ chunk_off,
                                                      int* __restrict__ list, int* __restrict__ rank) {
  // thread t owns 8 consecutive voxels so that ranks are assigned in voxel order
  const int base = blockIdx.x * kChunk + threadIdx.x * 8;
  int f[8], c = 0;
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    const int v = base + i;
    f[i] = (v < V && flags[v]) ? 1 : 0;
    c += f[i];
  }
  int s = c;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    const int t = __shfl_up_sync(0xffffffffu, s, o);
    if ((threadIdx.x & 31) >= o) s += t;
  }
  __shared__ int ws[8];
  if ((threadIdx.x & 31) == 31) ws[threadIdx.x >> 5] = s;
  __syncthreads();
  int off = chunk_off[blockIdx.x] + s - c;
  for (int w = 0; w < (threadIdx.x >> 5); ++w) off += ws[w];
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    const int v = base + i;
    if (v < V) {
      if (f[i]) {
        list[off] = v;
        rank[v] = off;
        ++off;
      } else {
        rank[v] = -1;
      }
    }
  }
}

// ------------------------------------------------------------------------------------------
// FPS: one thread-block cluster, running min-distance in registers, packed coords in smem
// ------------------------------------------------------------------------------------------
constexpr int kFpsThreads = 512;
constexpr uint32_t kInfDist = 0x7fffffffu;   // stands in for the reference's 1e10 initial value

struct FpsCand {
  uint32_t d;      // running min squared distance of the candidate (maximised)
  uint32_t t;      // tie key (minimised): bitrev(k mod bs) << 22 | k / bs
  uint32_t c;      // packed coordinate x | y << 10 | z << 20
  uint32_t pad;
};

__device__ __forceinline__ uint32_t pack_xyz(int v, int Y, int Z) {
  const int z = v % Z;
  const int t = v / Z;
  const int y = t % Y;
  const int x = t / Y;
  return (uint32_t)x | ((uint32_t)y << 10) | ((uint32_t)z << 20);
}
__device__ __forceinline__ int d2_packed(uint32_t a, uint32_t b) {
  const int dx = (int)(a & 1023u) - (int)(b & 1023u);
  const int dy = (int)((a >> 10) & 1023u) - (int)((b >> 10) & 1023u);
  const int dz = (int)(a >> 20) - (int)(b >> 20);
  return dx * dx + dy * dy + dz * dz;
}

// better(a, b): a beats b  (larger distance, then smaller tie key)
__device__ __forceinline__ void warp_best(uint32_t& d, uint32_t& t, uint32_t& c) {
  const uint32_t md = __reduce_max_sync(0xffffffffu, d);
  const uint32_t mt = __reduce_min_sync(0xffffffffu, d == md ? t : 0xffffffffu);
  const uint32_t src = __ffs(__ballot_sync(0xffffffffu, d == md && t == mt)) - 1;
  c = __shfl_sync(0xffffffffu, c, src);
  d = md;
  t = mt;
}

struct FpsJob {
  const int* list;   // occupied voxel ids (ascending), the points
  const int* count;  // device pointer to N
  int* out;          // [m] sampled indices into list
};

// --- cluster exchange ---------------------------------------------------------------------
// Two levels per round: the 16 warps of a CTA reduce their candidates through shared memory (one
// block barrier), then warp 0 sends the CTA's candidate into the inbox of every CTA of the cluster
// with st.async (remote shared-memory store that completes bytes on the *receiver's* mbarrier) and
// every warp waits on the local mbarrier -- no barrier.cluster.  (Round 1 sent every warp's
// candidate to every CTA: 16 x CS messages and mbarrier transactions per CTA and round, which is
// what bounded the round at 2.4 us; now CS messages.)
__device__ __forceinline__ uint32_t mapa_u32(uint32_t local_smem_addr, uint32_t cta_rank) {
  uint32_t r;
  asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(local_smem_addr), "r"(cta_rank));
  return r;
}
__device__ __forceinline__ void st_async_v4(uint32_t remote_addr, uint32_t remote_mbar, uint32_t a, uint32_t b,
                                            uint32_t c, uint32_t d) {
  asm volatile(
      "st.async.weak.shared::cluster.mbarrier::complete_tx::bytes.v4.b32 [%0], {%1, %2, %3, %4}, [%5];" ::"r"(
          remote_addr),
      "r"(a), "r"(b), "r"(c), "r"(d), "r"(remote_mbar)
      : "memory");
}
__device__ __forceinline__ void fps_mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"((uint32_t)__cvta_generic_to_shared(bar)), "r"(count));
}
__device__ __forceinline__ void fps_mbar_arm(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(
                   (uint32_t)__cvta_generic_to_shared(bar)),
               "r"(bytes)
               : "memory");
}
// Default (.acquire.cta) semantics on purpose: the payload arrives through st.async, whose bytes are made visible by
// the transaction count of this very mbarrier; an .acquire.cluster wait makes ptxas emit CCTL.IVALL (L1 invalidate)
// after every wait -- measured at 65 % of all stall samples of the kernel (profiles/r02_fps_ncu.txt).
__device__ __forceinline__ void fps_mbar_wait(uint64_t* bar, uint32_t parity) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "FPS_WAIT:\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t"
      "@p bra FPS_DONE;\n\t"
      "bra FPS_WAIT;\n\t"
      "FPS_DONE:\n\t}\n" ::"r"((uint32_t)__cvta_generic_to_shared(bar)),
      "r"(parity)
      : "memory");
}

constexpr int kFpsWarps = kFpsThreads / 32;

// Optional phase trace (tools/fps_trace.py): clock64 of CTA 0 / warp 0 at six points of rounds 1024..1087 of job 0.
#ifdef COOCC_FPS_TRACE
__device__ long long g_fps_trace[16 * 64 * 8];
#define FPS_T(slot) do { if (blockIdx.y == 0 && tid == 0 && j >= 1024 && j < 1088) \
                           g_fps_trace[(rank * 64 + (j - 1024)) * 8 + (slot)] = clock64(); } while (0)
#else
#define FPS_T(slot) do { } while (0)
#endif

// Points are dealt to the warps of a CTA in groups ("slots") of 32 consecutive list entries, round
// robin: slot s of warp w holds entries base + (s*16 + w)*32 + lane.  The list is sorted by voxel id,
// so a slot is a short run of a grid row with a tight bounding box.  Per round a warp
//   1. tests the new sample against the bounding box of each of its slots (lane-parallel),
//   2. updates only the slots the sample can reach (one point per lane) and refreshes their cached
//      maxima,
//   3. reduces its slot maxima to the warp candidate and sends it to every CTA (one hop).
// A sample only reaches slots within sqrt(current max-min distance), so after the first rounds a
// warp touches 0-2 slots per round however many points it owns: the round latency no longer grows
// with N.  Running distances and coordinates live in shared memory.
__global__ void __launch_bounds__(kFpsThreads, 1)
fps_kernel(FpsJob job0, FpsJob job1, int m, int Y, int Z, int log2bs_override, int S, int* started) {
  cg::cluster_group cluster = cg::this_cluster();
  const unsigned CS = cluster.num_blocks();
  const unsigned rank = cluster.block_rank();
  const FpsJob job = (blockIdx.y == 0) ? job0 : job1;
  const int n = *job.count;
  extern __shared__ uint32_t dyn[];
  uint32_t* dist = dyn;                                  // [S * 512]
  uint32_t* coords = dist + (size_t)S * kFpsThreads;     // [S * 512]
  uint32_t* blo = coords + (size_t)S * kFpsThreads;      // [16][S] slot bounding box (packed min)
  uint32_t* bhi = blo + kFpsWarps * S;                   // [16][S] packed max
  uint32_t* sd = bhi + kFpsWarps * S;                    // [16][S] slot best distance
  uint32_t* stie = sd + kFpsWarps * S;                   // [16][S] slot best tie key
  uint32_t* sc = stie + kFpsWarps * S;                   // [16][S] slot best coordinate
  __shared__ __align__(16) FpsCand inbox[2][16];               // [parity][sender cta]
  __shared__ __align__(16) FpsCand wcand[2][kFpsWarps];        // [parity][warp] candidates of this CTA's warps
  __shared__ __align__(8) uint64_t mbar[2];
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const uint32_t nmsg = CS;

  // block-size rule of the reference launcher (opt_n_threads): 2^floor(log2 n), capped at 1024
  int log2bs = 0;
  if (log2bs_override >= 0) {
    log2bs = log2bs_override;
  } else {
    while ((2 << log2bs) <= n && log2bs < 10) ++log2bs;
  }
  const uint32_t bsmask = (1u << log2bs) - 1u;
  // slot g = (sl * 16 + warp) * CS + rank holds list entries 32 g .. 32 g + 31: consecutive slots go to consecutive
  // CTAs, then warps.  The slots a new sample can reach are neighbours in space, i.e. short runs of consecutive
  // slots a few grid rows apart; dealt this way they spread over all CTAs and warps of the cluster (with one
  // contiguous range per CTA they all landed in one CTA, whose warps updated 3-6 slots each while the other 15
  // CTAs waited: 2400 of the 4000 cycles of a round, profiles/r02_fps_trace.txt).
  const int slot_stride = (int)CS * 32;
  const int slot_base = (int)rank * 32;

  // ---- load points, build slot bounding boxes ---------------------------------------------
  for (int sl = 0; sl < S; ++sl) {
    const int k = (sl * kFpsWarps + warp) * slot_stride + slot_base + lane;
    uint32_t c = 0u;
    int xl = 1 << 20, yl = 1 << 20, zl = 1 << 20, xh = -1, yh = -1, zh = -1;
    if (k < n) {
      c = pack_xyz(job.list[k], Y, Z);
      xl = xh = c & 1023u; yl = yh = (c >> 10) & 1023u; zl = zh = c >> 20;
    }
    coords[sl * kFpsThreads + tid] = c;
    dist[sl * kFpsThreads + tid] = kInfDist;
    xl = __reduce_min_sync(0xffffffffu, xl); xh = __reduce_max_sync(0xffffffffu, xh);
    yl = __reduce_min_sync(0xffffffffu, yl); yh = __reduce_max_sync(0xffffffffu, yh);
    zl = __reduce_min_sync(0xffffffffu, zl); zh = __reduce_max_sync(0xffffffffu, zh);
    if (lane == 0) {
      const bool has = xh >= 0;
      blo[warp * S + sl] = has ? ((uint32_t)xl | ((uint32_t)yl << 10) | ((uint32_t)zl << 20)) : 0xffffffffu;
      bhi[warp * S + sl] = has ? ((uint32_t)xh | ((uint32_t)yh << 10) | ((uint32_t)zh << 20)) : 0u;
      sd[warp * S + sl] = has ? kInfDist : 0u;          // INF forces the first update
      stie[warp * S + sl] = 0xffffffffu;
      sc[warp * S + sl] = 0u;
    }
  }
  uint32_t wd = 0u, wt = 0xffffffffu, wc = 0u;           // cached warp candidate

  uint32_t cur = pack_xyz(job.list[0], Y, Z);            // start index 0 (furthest_point_sample.py)
  if (tid == 0) {
    if (rank == 0 && m > 0) job.out[0] = 0;
    fps_mbar_init(&mbar[0], 1);
    fps_mbar_init(&mbar[1], 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    fps_mbar_arm(&mbar[1], nmsg * 16);                   // round 1
    fps_mbar_arm(&mbar[0], nmsg * 16);                   // round 2
  }
  uint32_t r_slot[2] = {0, 0}, r_bar[2] = {0, 0};
  if (lane < CS) {
    for (int p2 = 0; p2 < 2; ++p2) {
      r_slot[p2] = mapa_u32((uint32_t)__cvta_generic_to_shared(&inbox[p2][rank]), lane);
      r_bar[p2] = mapa_u32((uint32_t)__cvta_generic_to_shared(&mbar[p2]), lane);
    }
  }
  __syncthreads();
  cluster.sync();
  // every CTA of this cluster is resident: tell a gate kernel on another stream (coocc_gsf_fps_gate) that the
  // convolution it holds back may now take the remaining SMs
  if (started != nullptr && rank == 0 && tid == 0) atomicAdd(started, 1);

  for (int j = 1; j < m; ++j) {
    const int par = j & 1;
    const int px = cur & 1023u, py = (cur >> 10) & 1023u, pz = cur >> 20;
    bool changed = false;
    FPS_T(0);
    for (int s0 = 0; s0 < S; s0 += 32) {
      // 1. which of this warp's slots can the new sample reach?
      const int sl = s0 + lane;
      bool need = false;
      if (sl < S) {
        const uint32_t lo = blo[warp * S + sl], hi = bhi[warp * S + sl];
        if (lo != 0xffffffffu) {
          const int ddx = max(0, max((int)(lo & 1023u) - px, px - (int)(hi & 1023u)));
          const int ddy = max(0, max((int)((lo >> 10) & 1023u) - py, py - (int)((hi >> 10) & 1023u)));
          const int ddz = max(0, max((int)(lo >> 20) - pz, pz - (int)(hi >> 20)));
          need = (uint32_t)(ddx * ddx + ddy * ddy + ddz * ddz) < sd[warp * S + sl];
        }
      }
      uint32_t todo = __ballot_sync(0xffffffffu, need);
      // 2. update those slots, one point per lane
      while (todo) {
        const int u = s0 + __ffs(todo) - 1;
        todo &= todo - 1;
        const int k0 = (u * kFpsWarps + warp) * slot_stride + slot_base;
        const int k = k0 + lane;
        uint32_t bd = 0u, bt = 0xffffffffu, bc = 0u;
        if (k < n) {
          const uint32_t c = coords[u * kFpsThreads + tid];
          const uint32_t d = min((uint32_t)d2_packed(c, cur), dist[u * kFpsThreads + tid]);
          dist[u * kFpsThreads + tid] = d;
          bd = d;
          bc = c;
        }
        if (log2bs >= 5) {
          // The 32 entries of a slot differ only in the low five bits of k, which the bit reversal turns into the
          // top five bits of the tie key: inside a slot the tie order is bitrev5(lane).  One packed max-reduction
          // yields the slot's best distance and the winning lane.
          const uint32_t key = (bd << 5) | (31u - (__brev((uint32_t)lane) >> 27));
          const uint32_t mk = __reduce_max_sync(0xffffffffu, k < n ? key : 0u);
          const uint32_t wl = __brev(31u - (mk & 31u)) >> 27;
          bd = mk >> 5;
          bc = __shfl_sync(0xffffffffu, bc, wl);
          const uint32_t kw = (uint32_t)k0 + wl;
          bt = (__brev(kw & bsmask) >> (32 - log2bs) << 22) | (kw >> log2bs);
        } else {
          if (k < n) bt = (log2bs ? (__brev((uint32_t)k & bsmask) >> (32 - log2bs) << 22) : 0u) | ((uint32_t)k >> log2bs);
          warp_best(bd, bt, bc);
        }
        if (lane == 0) {
          sd[warp * S + u] = bd; stie[warp * S + u] = bt; sc[warp * S + u] = bc;
        }
        changed = true;
      }
    }
    FPS_T(1);
    // 3. warp candidate = best slot maximum (only when a slot changed)
    if (changed) {
      __syncwarp();
      uint32_t bd = 0u, bt = 0xffffffffu, bc = 0u;
      for (int sl = lane; sl < S; sl += 32) {
        const uint32_t d = sd[warp * S + sl], t = stie[warp * S + sl];
        if (d > bd || (d == bd && t < bt)) { bd = d; bt = t; bc = sc[warp * S + sl]; }
      }
      warp_best(bd, bt, bc);
      wd = bd; wt = bt; wc = bc;
    }
    FPS_T(2);
    // level 1: the CTA's best candidate (warp candidates through shared memory, one block barrier)
    if (lane == 0) {
      FpsCand w;
      w.d = wd; w.t = wt; w.c = wc; w.pad = 0u;
      wcand[par][warp] = w;
    }
    __syncthreads();
    FPS_T(3);
    if (warp == 0) {
      uint32_t bd = 0u, bt = 0xffffffffu, bc = 0u;
      if (lane < kFpsWarps) {
        const FpsCand cnd = wcand[par][lane];
        bd = cnd.d; bt = cnd.t; bc = cnd.c;
      }
      warp_best(bd, bt, bc);
      FPS_T(6);
      // level 2: lane i -> CTA i, completing 16 bytes on the receiver's mbarrier
      if (lane < CS) st_async_v4(r_slot[par], r_bar[par], bd, bt, bc, 0u);
    }
    FPS_T(4);
    fps_mbar_wait(&mbar[par], ((j - 1) >> 1) & 1);
    FPS_T(5);
    uint32_t bd = 0u, bt = 0xffffffffu, bc = 0u;
    if (lane < nmsg) {
      const FpsCand cnd = inbox[par][lane];
      bd = cnd.d; bt = cnd.t; bc = cnd.c;
    }
    warp_best(bd, bt, bc);
    cur = bc;
    FPS_T(7);
    // Buffer reuse needs no further barrier.  wcand[par] is rewritten in round j+2; a warp gets there only
    // through the block barrier of round j+1, which warp 0 reaches after its reads above.  inbox[par] is
    // overwritten by round j+2 messages; a CTA sends those only after it received *every* CTA's round j+1
    // message, and this CTA's is sent behind the round j+1 block barrier, i.e. after all its warps finished
    // the reads above (program order).  Re-arming for round j+2 happens here, before that barrier, hence
    // before any round j+2 byte can arrive; slower local warps still waiting on the completed phase are
    // unaffected (parity wait).
    // (Letting warp 0 alone run the exchange while the others park in a second block barrier was measured too:
    // 2.86 ms instead of 2.75 ms on the 200x200x16 grid.)
    if (tid == 0) {
      if (j + 2 < m) fps_mbar_arm(&mbar[par], nmsg * 16);
      if (rank == 0) {
        const uint32_t kmod = __brev(bt >> 22) >> (32 - log2bs);
        job.out[j] = (int)(((bt & 0x3fffffu) << log2bs) | kmod);
      }
    }
  }
  cluster.sync();
}

// ------------------------------------------------------------------------------------------
// rep_topk: K nearest keys (d2 <= 176) of every representative, ordered (d2 asc, key asc)
// ------------------------------------------------------------------------------------------
constexpr int kMaxK = 8;
constexpr int kTopkR = 13;            // floor(sqrt(176))
constexpr int kTopkD2 = 176;          // val < 13.3  <=>  d2 <= 176 on integer coordinates

__global__ void __launch_bounds__(128) rep_topk_kernel(const int* __restrict__ rep_idx, int nrep,
                                                       const int* __restrict__ qlist,
                                                       const int* __restrict__ key_rank, int X, int Y,
                                                       int Z, int K, int* __restrict__ out_idx,
                                                       int* __restrict__ out_d2) {
  const int r = blockIdx.x * 4 + (threadIdx.x >> 5);
  const int lane = threadIdx.x & 31;
  if (r >= nrep) return;
  const int v0 = qlist[rep_idx[r]];
  const int z0 = v0 % Z, y0 = (v0 / Z) % Y, x0 = v0 / (Z * Y);
  uint32_t bd[kMaxK], bv[kMaxK];
#pragma unroll
  for (int i = 0; i < kMaxK; ++i) { bd[i] = 0xffffffffu; bv[i] = 0xffffffffu; }
  const int W = 2 * kTopkR + 1;
  const int zlo = max(0, z0 - kTopkR), zhi = min(Z - 1, z0 + kTopkR);
  const int nz = zhi - zlo + 1;
  const int total = W * W * nz;
  for (int e = lane; e < total; e += 32) {
    const int z = zlo + e % nz;
    const int t = e / nz;
    const int dy = t % W - kTopkR, dx = t / W - kTopkR;
    const int x = x0 + dx, y = y0 + dy;
    if (x < 0 || x >= X || y < 0 || y >= Y) continue;
    const int dz = z - z0;
    const uint32_t d2 = dx * dx + dy * dy + dz * dz;
    if (d2 > (uint32_t)kTopkD2) continue;
    const int v = (x * Y + y) * Z + z;
    if (key_rank[v] < 0) continue;
    // insert (d2, v) into the lane-local sorted list
    uint32_t cd = d2, cv = (uint32_t)v;
#pragma unroll
    for (int i = 0; i < kMaxK; ++i) {
      if (i < K) {
        const bool lt = cd < bd[i] || (cd == bd[i] && cv < bv[i]);
        if (lt) {
          const uint32_t td = bd[i], tv = bv[i];
          bd[i] = cd; bv[i] = cv; cd = td; cv = tv;
        }
      }
    }
  }
  // K rounds of warp arg-min; the winning lane pops its head
  for (int k = 0; k < K; ++k) {
    const uint32_t md = __reduce_min_sync(0xffffffffu, bd[0]);
    const uint32_t mv = __reduce_min_sync(0xffffffffu, bd[0] == md ? bv[0] : 0xffffffffu);
    const bool win = (bd[0] == md && bv[0] == mv && md != 0xffffffffu);
    if (win) {
#pragma unroll
      for (int i = 0; i + 1 < kMaxK; ++i) { bd[i] = bd[i + 1]; bv[i] = bv[i + 1]; }
      bd[kMaxK - 1] = 0xffffffffu; bv[kMaxK - 1] = 0xffffffffu;
    }
    if (lane == 0) {
      const bool valid = md != 0xffffffffu;
      out_idx[r * K + k] = valid ? key_rank[mv] : -1;
      out_d2[r * K + k] = valid ? (int)md : -1;
    }
  }
}

// ------------------------------------------------------------------------------------------
// ball_assign: for every representative, the first `nsample` occupied query voxels (ascending
// voxel order) with d2 < radius^2 inherit its neighbours; the reference's duplicate-index
// scatter is resolved as "largest representative position wins" (= last writer in
// (representative, slot) order) with an atomicMax per (k, query).
// ------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(128) ball_assign_kernel(const int* __restrict__ rep_idx, int nrep,
                                                          const int* __restrict__ qlist,
                                                          const int* __restrict__ q_rank,
                                                          const int* __restrict__ topk_idx, int X,
                                                          int Y, int Z, int K, int radius, int nsample,
                                                          int nq_stride, int* __restrict__ winner,
                                                          int* __restrict__ group_out) {
  const int r = blockIdx.x * 4 + (threadIdx.x >> 5);
  const int lane = threadIdx.x & 31;
  if (r >= nrep) return;
  const int v0 = qlist[rep_idx[r]];
  const int z0 = v0 % Z, y0 = (v0 / Z) % Y, x0 = v0 / (Z * Y);
  uint32_t vmask = 0;
  for (int k = 0; k < K; ++k) vmask |= (topk_idx[r * K + k] >= 0 ? 1u : 0u) << k;
  if (vmask == 0 && group_out == nullptr) return;
  const int W = 2 * radius - 1;            // |d| <= radius-1 suffices for d2 < radius^2
  const int R = radius - 1;
  const int r2 = radius * radius;
  const int total = W * W * W;
  int cnt = 0;
  int first = -1;
  for (int e0 = 0; e0 < total && cnt < nsample; e0 += 32) {
    const int e = e0 + lane;
    int q = -1;
    if (e < total) {
      const int dz = e % W - R, dy = (e / W) % W - R, dx = e / (W * W) - R;
      const int x = x0 + dx, y = y0 + dy, z = z0 + dz;
      if (x >= 0 && x < X && y >= 0 && y < Y && z >= 0 && z < Z &&
          dx * dx + dy * dy + dz * dz < r2)
        q = q_rank[(x * Y + y) * Z + z];
    }
    const uint32_t hit = __ballot_sync(0xffffffffu, q >= 0);
    const int pos = cnt + __popc(hit & ((1u << lane) - 1u));
    if (q >= 0 && pos < nsample) {
      for (int k = 0; k < K; ++k)
        if (vmask & (1u << k)) atomicMax(&winner[k * nq_stride + q], r);
      if (group_out) group_out[r * nsample + pos] = q;
    }
    if (first < 0 && hit) first = __shfl_sync(0xffffffffu, q, __ffs(hit) - 1);
    cnt += __popc(hit);
  }
  if (group_out) {
    // ball_query pads the unused slots with the first hit (ball_query_cuda.cu:43-47)
    const int c = min(cnt, nsample);
    for (int s = c + lane; s < nsample; s += 32) group_out[r * nsample + s] = first < 0 ? 0 : first;
  }
}

// K == 1 brute-force branch of the reference (N_q <= 2048, bifuser_n.py:55-60): every query is
// its own representative; nn[q] = nearest key within the threshold or -1.
__global__ void __launch_bounds__(128) direct_nn_kernel(const int* __restrict__ qlist,
                                                        const int* __restrict__ qcount,
                                                        const int* __restrict__ key_rank, int X, int Y,
                                                        int Z, int* __restrict__ nn) {
  const int q = blockIdx.x * 4 + (threadIdx.x >> 5);
  const int lane = threadIdx.x & 31;
  if (q >= *qcount) return;
  const int v0 = qlist[q];
  const int z0 = v0 % Z, y0 = (v0 / Z) % Y, x0 = v0 / (Z * Y);
  const int W = 2 * kTopkR + 1;
  const int zlo = max(0, z0 - kTopkR), zhi = min(Z - 1, z0 + kTopkR);
  const int nz = zhi - zlo + 1;
  uint32_t bd = 0xffffffffu, bv = 0xffffffffu;
  for (int e = lane; e < W * W * nz; e += 32) {
    const int z = zlo + e % nz;
    const int t = e / nz;
    const int dy = t % W - kTopkR, dx = t / W - kTopkR;
    const int x = x0 + dx, y = y0 + dy;
    if (x < 0 || x >= X || y < 0 || y >= Y) continue;
    const int dz = z - z0;
    const uint32_t d2 = dx * dx + dy * dy + dz * dz;
    if (d2 > (uint32_t)kTopkD2) continue;
    const int v = (x * Y + y) * Z + z;
    if (key_rank[v] < 0) continue;
    if (d2 < bd || (d2 == bd && (uint32_t)v < bv)) { bd = d2; bv = (uint32_t)v; }
  }
  const uint32_t md = __reduce_min_sync(0xffffffffu, bd);
  const uint32_t mv = __reduce_min_sync(0xffffffffu, bd == md ? bv : 0xffffffffu);
  if (lane == 0) nn[q] = md != 0xffffffffu ? key_rank[mv] : -1;
}

}  // namespace coocc

using namespace coocc;

extern "C" int coocc_gsf_pack(const float* src, long long sC, long long sX, long long sY, long long sZ,
                              int C, int X, int Y, int Z, float* dst, long long ldo,
                              unsigned char* flags, void* stream) {
  if (!src || !dst || !flags || C < 1 || X < 1 || Y < 1 || Z < 1) return COOCC_ERR_ARG;
  // fast axis = spatial axis with the smallest stride (coalesced reads along it)
  int fast = 2;
  long long best = sZ;
  if (sY < best) { best = sY; fast = 1; }
  if (sX < best) { best = sX; fast = 0; }
  const int dims[3] = {X, Y, Z};
  const int a_ax = (fast == 0) ? 1 : 0;
  const int b_ax = (fast == 2) ? 1 : 2;
  {
    // flat walk over (mid, fast) when the two axes are contiguous in memory
    const long long st3[3] = {sX, sY, sZ};
    const int mid = st3[a_ax] <= st3[b_ax] ? a_ax : b_ax;
    const int slow = mid == a_ax ? b_ax : a_ax;
    if (st3[mid] == st3[fast] * dims[fast] && dims[slow] <= 65535 && (long long)dims[mid] * dims[fast] < (1LL << 30)) {
      const int nF = dims[mid] * dims[fast];
      dim3 grid((nF + 31) / 32, dims[slow]);
      pack_flat_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(src, sC, st3[fast], st3[slow], C, nF, dims[fast], fast, mid,
                                                              slow, Y, Z, dst, ldo, flags);
      return cudaGetLastError() == cudaSuccess ? 0 : COOCC_ERR_CUDA;
    }
  }
  if (dims[a_ax] > 65535 || dims[b_ax] > 65535) return COOCC_ERR_CAPACITY;
  dim3 grid((dims[fast] + 31) / 32, dims[a_ax], dims[b_ax]);
  pack_kernel<<<grid, dim3(32, 8), 0, (cudaStream_t)stream>>>(src, sC, sX, sY, sZ, C, X, Y, Z, fast, dst,
                                                            ldo, flags);
  return cudaGetLastError() == cudaSuccess ? 0 : COOCC_ERR_CUDA;
}

extern "C" long long coocc_gsf_compact_workspace(int V) {
  return (long long)((V + kChunk - 1) / kChunk) * sizeof(int);
}

extern "C" int coocc_gsf_compact(const unsigned char* flags, int V, int* list, int* rank, int* count,
                                 void* workspace, void* stream) {
  if (!flags || !list || !rank || !count || !workspace || V < 1) return COOCC_ERR_ARG;
  const int nchunks = (V + kChunk - 1) / kChunk;
  int* chunk = reinterpret_cast<int*>(workspace);
  cudaStream_t st = (cudaStream_t)stream;
  count_kernel<<<nchunks, 256, 0, st>>>(flags, V, chunk);
  scan_kernel<<<1, 1024, 0, st>>>(chunk, nchunks, count);
  scatter_kernel<<<nchunks, 256, 0, st>>>(flags, V, chunk, list, rank);
  return cudaGetLastError() == cudaSuccess ? 0 : COOCC_ERR_CUDA;
}

static int g_fps_cs = 0, g_fps_flags = 0, g_fps_signal = 0;
__device__ int g_fps_started = 0;

static size_t fps_smem_bytes(int S) {
  return ((size_t)2 * S * kFpsThreads + (size_t)5 * kFpsWarps * S) * sizeof(uint32_t);
}

// Runs one or two independent FPS problems (both GSFusion directions) in one launch, one cluster
// each.  n_max = upper bound of the point counts (sizes the cluster); the exact counts are read
// on the device from count0/count1.
extern "C" int coocc_gsf_fps(const int* list0, const int* count0, int* out0, const int* list1,
                             const int* count1, int* out1, int n_max, int m, int Y, int Z,
                             void* stream) {
  if (!list0 || !count0 || !out0 || n_max < 1 || m < 1) return COOCC_ERR_ARG;
  if (Y > 1023 || Z > 1023) return COOCC_ERR_CAPACITY;
  const int njobs = list1 ? 2 : 1;
  FpsJob j0{list0, count0, out0}, j1{list1 ? list1 : list0, list1 ? count1 : count0, list1 ? out1 : out0};
  // smallest cluster with <= 24 slots per warp (measured on B200: 48k points run fastest on 4 CTAs;
  // the per-round cost is the one-hop exchange, which grows with the cluster size)
  int cs = 1;
  while (cs < 16 && (n_max + cs * kFpsThreads - 1) / (cs * kFpsThreads) > 24) cs *= 2;
  if (g_fps_cs > 0) cs = g_fps_cs;
  const int S = (n_max + cs * kFpsThreads - 1) / (cs * kFpsThreads);
  const size_t smem = fps_smem_bytes(S);
  if (smem > 217 * 1024) return COOCC_ERR_CAPACITY;
  cudaStream_t st = (cudaStream_t)stream;
  cudaError_t e = cudaFuncSetAttribute(fps_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  if (e != cudaSuccess) return COOCC_ERR_CUDA;
  if (cs > 8) {
    e = cudaFuncSetAttribute(fps_kernel, cudaFuncAttributeNonPortableClusterSizeAllowed, 1);
    if (e != cudaSuccess) return COOCC_ERR_CUDA;
  }
  cudaLaunchConfig_t cfg;
  memset(&cfg, 0, sizeof(cfg));
  cfg.gridDim = dim3(cs, njobs, 1);
  cfg.blockDim = dim3(kFpsThreads, 1, 1);
  cfg.dynamicSmemBytes = smem;
  cfg.stream = st;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = cs;
  attr[0].val.clusterDim.y = 1;
  attr[0].val.clusterDim.z = 1;
  cfg.attrs = attr;
  cfg.numAttrs = 1;
  int* started = nullptr;
  if (g_fps_signal && cudaGetSymbolAddress(reinterpret_cast<void**>(&started), g_fps_started) != cudaSuccess)
    return COOCC_ERR_CUDA;
  e = cudaLaunchKernelEx(&cfg, fps_kernel, j0, j1, m, Y, Z, -1, S, started);
  return e == cudaSuccess ? 0 : COOCC_ERR_CUDA;
}

// Index pipelining (co-occ_b200/graph.py): the FPS clusters of the NEXT step run on a side branch next to the last
// convolution of this step's backward, which is captured on the SMs they leave free.  Which of the two kernels the
// block scheduler places first when both become ready is not defined (stream priorities are not honoured inside
// an instantiated graph without per-node priorities), and a convolution placed first would scatter over all GPCs
// and leave no 16 free SMs in one GPC for a cluster.  So the order is made explicit: launches issued while the signal
// is on count their resident clusters in g_fps_started, and a one-thread gate kernel in front of the convolution
// waits for them (and resets the counter).  The clusters do not depend on the gate, so this cannot deadlock.
extern "C" int coocc_gsf_fps_signal(int on) {
  g_fps_signal = on ? 1 : 0;
  return 0;
}
__global__ void fps_gate_kernel(int* started, int nclusters) {
  while (atomicAdd(started, 0) < nclusters) __nanosleep(200);
  *started = 0;
  __threadfence();
}
extern "C" int coocc_gsf_fps_gate(int nclusters, void* stream) {
  int* started = nullptr;
  if (cudaGetSymbolAddress(reinterpret_cast<void**>(&started), g_fps_started) != cudaSuccess) return COOCC_ERR_CUDA;
  fps_gate_kernel<<<1, 1, 0, (cudaStream_t)stream>>>(started, nclusters);
  return cudaGetLastError() == cudaSuccess ? 0 : COOCC_ERR_CUDA;
}

// tuning hook (benchmarks only): force the cluster size (0 = automatic) and exchange variant flags
#ifdef COOCC_FPS_TRACE
extern "C" int coocc_gsf_fps_trace(long long* host_out) {
  return cudaMemcpyFromSymbol(host_out, g_fps_trace, sizeof(long long) * 16 * 64 * 8) == cudaSuccess ? 0 : COOCC_ERR_CUDA;
}
#endif

extern "C" int coocc_gsf_fps_tune(int cluster_size, int flags) {
  g_fps_cs = cluster_size;
  g_fps_flags = flags;
  return 0;
}

extern "C" int coocc_gsf_rep_topk(const int* rep_idx, int nrep, const int* qlist, const int* key_rank,
                                  int X, int Y, int Z, int K, int* out_idx, int* out_d2, void* stream) {
  if (!rep_idx || !qlist || !key_rank || !out_idx || !out_d2 || K < 1 || K > kMaxK) return COOCC_ERR_ARG;
  rep_topk_kernel<<<(nrep + 3) / 4, 128, 0, (cudaStream_t)stream>>>(rep_idx, nrep, qlist, key_rank, X, Y,
                                                                   Z, K, out_idx, out_d2);
  return cudaGetLastError() == cudaSuccess ? 0 : COOCC_ERR_CUDA;
}

extern "C" int coocc_gsf_ball_assign(const int* rep_idx, int nrep, const int* qlist, const int* q_rank,
                                     const int* topk_idx, int X, int Y, int Z, int K, int radius,
                                     int nsample, int nq_stride, int* winner, int* group_out,
                                     void* stream) {
  if (!rep_idx || !qlist || !q_rank || !topk_idx || !winner || K < 1 || K > kMaxK || radius < 1)
    return COOCC_ERR_ARG;
  ball_assign_kernel<<<(nrep + 3) / 4, 128, 0, (cudaStream_t)stream>>>(
      rep_idx, nrep, qlist, q_rank, topk_idx, X, Y, Z, K, radius, nsample, nq_stride, winner, group_out);
  return cudaGetLastError() == cudaSuccess ? 0 : COOCC_ERR_CUDA;
}

extern "C" int coocc_gsf_direct_nn(const int* qlist, const int* qcount, int nq_max, const int* key_rank,
                                   int X, int Y, int Z, int* nn, void* stream) {
  if (!qlist || !qcount || !key_rank || !nn || nq_max < 1) return COOCC_ERR_ARG;
  direct_nn_kernel<<<(nq_max + 3) / 4, 128, 0, (cudaStream_t)stream>>>(qlist, qcount, key_rank, X, Y, Z, nn);
  return cudaGetLastError() == cudaSuccess ? 0 : COOCC_ERR_CUDA;
}
