// peer_reduce.cu -- latency-bound all-reduce (sum) of a few thousand floats over NVLink peer memory, for the SyncBN
// statistics of the data-parallel path.
//
// The reference converts every BatchNorm of the model to torch.nn.SyncBatchNorm (tools/train.py:222-223, sync_bn=True
// in the configs): each of the 36 BatchNorms of the hot path all-gathers its batch statistics in forward and
// all-reduces two sums in backward -- 72 collectives of <= 8 KB per step, every one on the critical path.  Through
// NCCL each costs ~20 us of launch + protocol latency; here it is one single-CTA kernel over peer-mapped buffers:
//
//   push   every rank stores its n values into its own row of the call's slot in EVERY rank's exchange buffer, each
//          value as one 8-byte word {float bits, epoch} (plain stores to peer memory over NVLink / NVSwitch; an aligned
//          8-byte store is single-copy atomic, so the epoch travels with the value: no fence, no separate flag -- the
//          low-latency protocol NCCL calls LL);
//   wait   a rank spins on its LOCAL words until a word carries the slot's current epoch;
//   sum    it adds the rows in rank order (identical order on every rank => bit-identical replicas) and writes the
//          result over its input.
// Measured on 2 x B200: see profiles/ (first version with a separate flag and two system fences: 12.7 us per call
// against 18.9 us for the NCCL all-reduce of the same 8 KB).
//
// Buffer of a rank (all ranks use the same layout; allocation and pointer exchange are the host's job -- the Python
// side uses torch.distributed._symmetric_memory):  uint2 rows[nslots][world][slot_floats].  Slots are used round robin
// by the sequence of calls of a step (every rank issues the same sequence); `epoch[slot]` (device memory of the calling
// rank) counts the uses of a slot, so nothing baked into a CUDA graph changes between replays.  Two slots already
// exclude reuse hazards: a peer can write epoch e+1 of a slot only after it received this rank's words of the call
// *after* epoch e, which this rank sends after it has consumed epoch e.
#include <cuda_runtime.h>
#include <stdint.h>

#include "../../include/coocc_b200.h"

namespace coocc {

constexpr int kPeerMax = 16;

struct PeerBufs {
  uint2* rows[kPeerMax];        // rank r's exchange buffer as seen from this device
};

__device__ __forceinline__ uint2 ld_word(const uint2* p) {
  uint2 v;
  asm volatile("ld.relaxed.sys.global.v2.u32 {%0, %1}, [%2];" : "=r"(v.x), "=r"(v.y) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ void st_word(uint2* p, uint32_t a, uint32_t b) {
  asm volatile("st.relaxed.sys.global.v2.u32 [%0], {%1, %2};" ::"l"(p), "r"(a), "r"(b) : "memory");
}

__global__ void __launch_bounds__(1024) peer_allreduce_kernel(float* __restrict__ data, int n, PeerBufs pb, int rank,
                                                              int world, int slot, int slot_floats,
                                                              unsigned* __restrict__ epoch) {
  const unsigned e = epoch[slot] + 1u;
  const size_t row_off = ((size_t)slot * world + rank) * slot_floats;
  // push my row into every rank's buffer (my own included)
  for (int i = threadIdx.x; i < n; i += blockDim.x) {
    const uint32_t v = __float_as_uint(data[i]);
    for (int r = 0; r < world; ++r) st_word(pb.rows[r] + row_off + i, v, e);
  }
  // every thread waits for the words it sums: no block barrier, no fence
  const uint2* mine = pb.rows[rank] + (size_t)slot * world * slot_floats;
  for (int i = threadIdx.x; i < n; i += blockDim.x) {
    float s = 0.f;
    for (int r = 0; r < world; ++r) {
      const uint2* w = mine + (size_t)r * slot_floats + i;
      uint2 v = ld_word(w);
      while (v.y != e) v = ld_word(w);
      s += __uint_as_float(v.x);
    }
    data[i] = s;
  }
  __syncthreads();
  if (threadIdx.x == 0) epoch[slot] = e;
}

}  // namespace coocc

using namespace coocc;

// Bytes of one rank's exchange buffer.
extern "C" long long coocc_peer_buffer_bytes(int world, int nslots, int slot_floats) {
  if (world < 1 || world > kPeerMax || nslots < 2 || slot_floats < 1) return -1;
  return (long long)nslots * world * slot_floats * 8;
}

// data: n <= slot_floats floats on this device, summed over all ranks in place.  peer_bufs: HOST array of `world`
// device pointers, entry r = rank r's exchange buffer mapped into this process (zero-initialised once, before the first
// call on any rank).  slot in [0, nslots): the call's position in the step's call sequence modulo nslots (identical on
// every rank).  epoch: device unsigned[nslots] of this rank, zero-initialised.
extern "C" int coocc_peer_allreduce(float* data, int n, void* const* peer_bufs, int rank, int world, int slot, int nslots,
                                    int slot_floats, unsigned* epoch, void* stream) {
  if (!data || !peer_bufs || !epoch || n < 1 || n > slot_floats || world < 1 || world > kPeerMax || rank < 0 ||
      rank >= world || slot < 0 || slot >= nslots)
    return COOCC_ERR_ARG;
  PeerBufs pb{};
  for (int r = 0; r < world; ++r) {
    if (!peer_bufs[r]) return COOCC_ERR_ARG;
    if (reinterpret_cast<uintptr_t>(peer_bufs[r]) & 7) return COOCC_ERR_ALIGN;
    pb.rows[r] = reinterpret_cast<uint2*>(peer_bufs[r]);
  }
  int threads = n < 1024 ? ((n + 31) / 32) * 32 : 1024;
  if (threads < 32) threads = 32;
  peer_allreduce_kernel<<<1, threads, 0, (cudaStream_t)stream>>>(data, n, pb, rank, world, slot, slot_floats, epoch);
  return cudaGetLastError() == cudaSuccess ? 0 : COOCC_ERR_CUDA;
}
