// comm_kernels.cuh -- one-shot "push" all-reduce over NVLink peer memory for the tensor-parallel layers.
//
// The row-parallel projections (o_proj, down_proj) of a decode step leave an FP16 [batch, hidden] partial on every rank
// (512 KiB at Llama-65B, batch 32) that must be summed across the ranks: a latency-bound collective.
// The reference has no multi-GPU code at all; round 1 called ncclAllReduce.  This kernel does the exchange itself:
//   1. every CTA pushes its slice of the local partial into slot [parity][my rank] of EVERY rank's receive buffer (plain
//      16-byte stores through the NVLink peer mappings: fire and forget, no read round trip),
//   2. fences, raises one flag per peer (st.release.sys) and waits for the peers' flags of the same CTA index
//      (ld.acquire.sys on local memory),
//   3. sums the `world` slots of its slice from LOCAL memory in rank order (FP32, identical on every rank) and writes the
//      FP16 result.
// Receive buffers and flags are double-buffered by the parity of a device-side epoch counter, which makes a closing
// barrier unnecessary (a rank can only reach epoch e+2 after every peer has signalled e+1, i.e. finished reading e) and
// keeps the kernel CUDA-graph capturable: no host-side state changes between replays.
#pragma once
#include "ptx_sm100.cuh"

namespace atom {

constexpr int AR_CTAS = 64, AR_THREADS = 256;

__device__ __forceinline__ void st_release_sys_u32(uint32_t* p, uint32_t v) {
  asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
__device__ __forceinline__ uint32_t ld_acquire_sys_u32(const uint32_t* p) {
  uint32_t v;
  asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ uint4 ld_volatile_v4(const void* p) {
  uint4 v;
  asm volatile("ld.volatile.global.v4.u32 {%0, %1, %2, %3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "l"(p) : "memory");
  return v;
}

// in, out: f16 [numel] local; bufs[r]: rank r's receive buffer, f16 [2][world][slot_elems]; flags[r]: u32 [2][AR_CTAS][world];
// epoch: u32 [AR_CTAS] local, zero-initialised once.  numel % 8 == 0, numel <= slot_elems.
__global__ void __launch_bounds__(AR_THREADS)
allreduce_push_kernel(const uint4* __restrict__ in, uint4* __restrict__ out, uint4* const* __restrict__ bufs,
                      uint32_t* const* __restrict__ flags, uint32_t* __restrict__ epoch, long long nchunks, long long slot_chunks,
                      int rank, int world) {
  __shared__ uint32_t e_s;
  const int tid = threadIdx.x, bid = blockIdx.x;
  if (tid == 0) e_s = epoch[bid] + 1;
  __syncthreads();
  const uint32_t e = e_s, par = e & 1u;
  const long long per = (nchunks + gridDim.x - 1) / gridDim.x;
  const long long lo = min((long long)bid * per, nchunks), hi = min(lo + per, nchunks);
  // 1. push my slice into slot [par][rank] of every rank (my own included: the reduction below is then uniform)
  for (long long i = lo + tid; i < hi; i += AR_THREADS) {
    const uint4 v = in[i];
    for (int r = 0; r < world; ++r) bufs[r][((long long)par * world + rank) * slot_chunks + i] = v;
  }
  __threadfence_system();
  __syncthreads();
  // 2. one flag per peer; wait for the same CTA of every peer
  if (tid < world) {
    st_release_sys_u32(flags[tid] + ((size_t)par * gridDim.x + bid) * world + rank, e);
    const uint32_t* mine = flags[rank] + ((size_t)par * gridDim.x + bid) * world + tid;
    // bounded: a peer that never arrives (crashed rank, mismatched launch sequence) traps after ~4 s instead of hanging the GPU
    unsigned long long t0 = 0;
    for (uint32_t spins = 0; ld_acquire_sys_u32(mine) != e; ++spins) {
      if ((spins & 0xFFFu) == 0xFFFu) {
        unsigned long long now;
        asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(now));
        if (t0 == 0) t0 = now;
        else if (now - t0 > 4000000000ull) asm volatile("trap;");
      }
    }
  }
  __syncthreads();
  // 3. local reduction in rank order
  const uint4* local = bufs[rank] + (long long)par * world * slot_chunks;
  for (long long i = lo + tid; i < hi; i += AR_THREADS) {
    float acc[8];
#pragma unroll
    for (int k = 0; k < 8; ++k) acc[k] = 0.f;
    for (int r = 0; r < world; ++r) {
      const uint4 v = ld_volatile_v4(local + (long long)r * slot_chunks + i);     // written remotely: do not trust L1
      const __half2* h = reinterpret_cast<const __half2*>(&v);
#pragma unroll
      for (int k = 0; k < 4; ++k) { const float2 f = __half22float2(h[k]); acc[2 * k] += f.x; acc[2 * k + 1] += f.y; }
    }
    uint4 o;
    __half2* oh = reinterpret_cast<__half2*>(&o);
#pragma unroll
    for (int k = 0; k < 4; ++k) oh[k] = __floats2half2_rn(acc[2 * k], acc[2 * k + 1]);
    out[i] = o;
  }
  if (tid == 0) epoch[bid] = e;
}

}  // namespace atom
