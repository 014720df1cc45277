// device_common.cuh — shared device-side building blocks (sm_100a).
//   * mbarrier + 1-D bulk async copy (TMA engine, SASS UBLKCP) wrappers
//   * epoch-stamped decoupled look-back status words (no per-frame zeroing)
//   * warp / block scan helpers
#pragma once
#include <cstdint>
#include <cuda_runtime.h>

namespace vkgs {

constexpr unsigned FULL_MASK = 0xffffffffu;

// ---------------------------------------------------------------------------------------------
// mbarrier + cp.async.bulk (global -> shared::cta). Size and both addresses must be multiples of 16.

__device__ __forceinline__ uint32_t smem_u32(const void* p)
{
  return static_cast<uint32_t>(__cvta_generic_to_shared(p));
}

__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count)
{
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}

__device__ __forceinline__ void mbar_fence_init()
{
  // make the initialised barrier visible to the async proxy (TMA engine)
  asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}

__device__ __forceinline__ void mbar_arrive_expect_tx(uint64_t* bar, uint32_t bytes)
{
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}

__device__ __forceinline__ void bulk_copy_g2s(void* smem_dst, const void* gmem_src, uint32_t bytes, uint64_t* bar)
{
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smem_u32(smem_dst)),
               "l"(gmem_src), "r"(bytes), "r"(smem_u32(bar))
               : "memory");
}

__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity)
{
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "WAIT_%=:\n"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
      "@p bra DONE_%=;\n"
      "bra WAIT_%=;\n"
      "DONE_%=:\n"
      "}\n" ::"r"(smem_u32(bar)),
      "r"(parity)
      : "memory");
}

// ---------------------------------------------------------------------------------------------
// Decoupled look-back status words.
//   bits 63..34 : epoch (unique per kernel launch that uses the array; 0 is never used)
//   bits 33..32 : state (1 = block aggregate available, 2 = inclusive prefix available)
//   bits 31..0  : value
// A word whose epoch does not match is "not ready", so arrays are never cleared between frames.

constexpr uint64_t LB_AGGREGATE = 1ull;
constexpr uint64_t LB_INCLUSIVE = 2ull;

__device__ __forceinline__ uint64_t lb_pack(uint32_t epoch, uint64_t state, uint32_t value)
{
  return (static_cast<uint64_t>(epoch) << 34) | (state << 32) | static_cast<uint64_t>(value);
}

__device__ __forceinline__ void lb_store(uint64_t* p, uint64_t v)
{
  asm volatile("st.relaxed.gpu.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}

__device__ __forceinline__ uint64_t lb_load(const uint64_t* p)
{
  uint64_t v;
  asm volatile("ld.relaxed.gpu.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
  return v;
}

// Walk predecessors tile-1, tile-2, ... of a single scalar chain until an inclusive prefix is
// found. `stride` is the distance in words between consecutive tiles of the same chain.
__device__ __forceinline__ uint32_t lb_lookback(const uint64_t* status, int64_t tile, int64_t stride, uint32_t epoch)
{
  uint32_t exclusive = 0;
  for(int64_t p = tile - 1; p >= 0; --p)
  {
    uint64_t w;
    do
    {
      w = lb_load(status + p * stride);
    } while(static_cast<uint32_t>(w >> 34) != epoch);
    exclusive += static_cast<uint32_t>(w);
    if(((w >> 32) & 3ull) == LB_INCLUSIVE)
      break;
  }
  return exclusive;
}

// ---------------------------------------------------------------------------------------------
// scans

__device__ __forceinline__ uint32_t warp_inclusive_scan(uint32_t v, unsigned lane)
{
#pragma unroll
  for(int o = 1; o < 32; o <<= 1)
  {
    const uint32_t n = __shfl_up_sync(FULL_MASK, v, o);
    if(lane >= static_cast<unsigned>(o))
      v += n;
  }
  return v;
}

// Exclusive scan over a block of NWARPS*32 threads; returns the exclusive prefix of `v` and the
// block total through `total`. `s_warp` must hold NWARPS+1 words. Contains two __syncthreads().
template <int NWARPS>
__device__ __forceinline__ uint32_t block_exclusive_scan(uint32_t v, uint32_t* s_warp, uint32_t& total)
{
  const unsigned lane = threadIdx.x & 31u, warp = threadIdx.x >> 5;
  const uint32_t inc  = warp_inclusive_scan(v, lane);
  if(lane == 31)
    s_warp[warp] = inc;
  __syncthreads();
  if(warp == 0)
  {
    uint32_t w = lane < NWARPS ? s_warp[lane] : 0u;
    uint32_t s = warp_inclusive_scan(w, lane);
    if(lane < NWARPS)
      s_warp[lane] = s - w;
    if(lane == NWARPS - 1)
      s_warp[NWARPS] = s;
  }
  __syncthreads();
  total = s_warp[NWARPS];
  return s_warp[warp] + inc - v;
}

}  // namespace vkgs
