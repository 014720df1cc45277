// device_common.cuh — shared device-side building blocks (sm_100a).
//   * mbarrier + 1-D bulk async copy (TMA engine, SASS UBLKCP) wrappers
//   * epoch-stamped decoupled look-back status words (no per-frame zeroing)
//   * warp / block scan helpers
#pragma once
#include <cmath>
#include <cstdint>
#include <cstring>
#include <cuda_runtime.h>

namespace vkgs {

constexpr unsigned FULL_MASK = 0xffffffffu;

// Optional in-kernel timeline (tools/kbench.cu defines VKGS_TIMELINE): thread 0 of a block stamps
// the SM clock at phase boundaries. Compiles to nothing in the product build.
#ifdef VKGS_TIMELINE
static __device__ long long* g_vkgsTimeline = nullptr;  // [block][16] (single-TU harness only)
#define VKGS_TL(block, k)                                                                                                      \
  do                                                                                                                           \
  {                                                                                                                            \
    if(threadIdx.x == 0 && g_vkgsTimeline)                                                                                     \
      g_vkgsTimeline[(block)*16 + (k)] = clock64();                                                                            \
  } while(0)
#else
#define VKGS_TL(block, k) ((void)0)
#endif

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
// Programmatic dependent launch (griddepcontrol). The kernels of a frame's front end are a chain of short, latency-bound
// launches on one stream: launched with the programmatic-stream-serialization attribute, a kernel's CTAs may become
// resident while its predecessor drains, run the part of their prologue that does not depend on it (shared-memory
// clears, barrier inits, the ticket draw) and block in pdl_wait() until the predecessor has completed and its writes are
// visible. Both instructions are no-ops in a launch without the attribute.
__device__ __forceinline__ void pdl_wait()
{
  asm volatile("griddepcontrol.wait;" ::: "memory");
}
__device__ __forceinline__ void pdl_launch_dependents()
{
  asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
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

// Per-thread look-back over one chain (tile-1, tile-2, ...; `stride` words between consecutive
// tiles of the chain) until an inclusive prefix is found. Predecessors are fetched BATCH at a time
// with independent loads, so the serial dependency is one L2 round trip per BATCH tiles instead of
// one per tile (all resident tiles publish their aggregates at about the same time, so the walk
// can be hundreds of tiles long in the first wave).
template <int BATCH>
__device__ __forceinline__ uint32_t lb_lookback(const uint64_t* status, int64_t tile, int64_t stride, uint32_t epoch)
{
  uint32_t exclusive = 0;
  int64_t  p         = tile - 1;
  while(p >= 0)
  {
    uint64_t w[BATCH];
#pragma unroll
    for(int b = 0; b < BATCH; b++)
      w[b] = (p - b >= 0) ? lb_load(status + (p - b) * stride) : 0ull;
#pragma unroll
    for(int b = 0; b < BATCH; b++)
    {
      if(p - b < 0)
        return exclusive;
      while(static_cast<uint32_t>(w[b] >> 34) != epoch)
        w[b] = lb_load(status + (p - b) * stride);
      exclusive += static_cast<uint32_t>(w[b]);
      if(((w[b] >> 32) & 3ull) == LB_INCLUSIVE)
        return exclusive;
    }
    p -= BATCH;
  }
  return exclusive;
}

// Warp-cooperative look-back over a scalar chain (stride 1): each step inspects a window of
// 32*W predecessors at once (W independent loads per lane). Must be called by all 32 lanes of one
// warp; every lane gets the result.
template <int W = 1>
__device__ __forceinline__ uint32_t lb_lookback_warp(const uint64_t* status, int64_t tile, uint32_t epoch)
{
  const unsigned lane      = threadIdx.x & 31u;
  uint32_t       exclusive = 0;
  for(int64_t base = tile - 1; base >= 0; base -= 32 * W)
  {
    uint64_t w[W];
#pragma unroll
    for(int k = 0; k < W; k++)
    {
      const int64_t p = base - 32 * k - lane;
      w[k]            = p >= 0 ? lb_load(status + p) : 0ull;
    }
    bool found = false;
#pragma unroll
    for(int k = 0; k < W; k++)
    {
      const int64_t p = base - 32 * k - lane;
      if(p >= 0)
        while(static_cast<uint32_t>(w[k] >> 34) != epoch)
          w[k] = lb_load(status + p);
      const bool     inclusive = p >= 0 && ((w[k] >> 32) & 3ull) == LB_INCLUSIVE;
      const unsigned incMask   = __ballot_sync(FULL_MASK, inclusive);
      // lanes up to and including the nearest inclusive predecessor contribute
      const unsigned upto = incMask ? (__ffs(incMask) - 1) : 31u;
      uint32_t       v    = (p >= 0 && lane <= upto) ? static_cast<uint32_t>(w[k]) : 0u;
#pragma unroll
      for(int o = 16; o > 0; o >>= 1)
        v += __shfl_xor_sync(FULL_MASK, v, o);
      exclusive += v;
      if(incMask)
      {
        found = true;
        break;
      }
    }
    if(found)
      break;
  }
  return exclusive;
}

// ---------------------------------------------------------------------------------------------
// Lanes of `active` holding the same BITS-bit digit as the caller. MATCH.ANY is serviced once per
// DISTINCT value in the warp (about 800 cycles for 32 random 8-bit digits, measured on B200), so
// the peer mask is built from BITS ballots instead: cost independent of the digit distribution.
template <int BITS>
__device__ __forceinline__ unsigned match_digit(unsigned active, uint32_t digit)
{
  unsigned peers = active;
#pragma unroll
  for(int b = 0; b < BITS; b++)
  {
    const bool     bit = (digit >> b) & 1u;
    const unsigned v   = __ballot_sync(active, bit);
    peers &= bit ? v : ~v;
  }
  return peers;
}

// ---- exact fp32 building blocks ---------------------------------------------------------------------------------
// The decision-critical elementary functions below run on the device with explicit round-to-nearest intrinsics (immune
// to FMA contraction, whatever the file's -fmad setting) and on the host with the plain operators (host code of this
// library is compiled with -ffp-contract=off): the same IEEE operations in the same order. The host instantiations are
// exported through vkgs_exact_math_host so that tests/test_oracle_kat.py can pin the functions the kernels run against
// the oracle bit for bit on a machine without a GPU.
#ifdef __CUDA_ARCH__
#define VKGS_FMA(a, b, c) __fmaf_rn(a, b, c)
#define VKGS_MUL(a, b) __fmul_rn(a, b)
#define VKGS_ADD(a, b) __fadd_rn(a, b)
#define VKGS_SUB(a, b) __fsub_rn(a, b)
#define VKGS_DIV(a, b) __fdiv_rn(a, b)
#define VKGS_SQRT(a) __fsqrt_rn(a)
#define VKGS_U2F(u) __uint_as_float(u)
#else
__host__ inline float vkgsHostU2F(uint32_t u)
{
  float f;
  memcpy(&f, &u, 4);
  return f;
}
#define VKGS_FMA(a, b, c) fmaf(a, b, c)
#define VKGS_MUL(a, b) ((a) * (b))
#define VKGS_ADD(a, b) ((a) + (b))
#define VKGS_SUB(a, b) ((a) - (b))
#define VKGS_DIV(a, b) ((a) / (b))
#define VKGS_SQRT(a) sqrtf(a)
#define VKGS_U2F(u) vkgsHostU2F(u)
#endif

// Same operation sequence as orc_expf (oracle/vkgs_oracle.c): Cody-Waite + Cephes polynomial.
__host__ __device__ __forceinline__ float expfExact(float x)
{
  if(x != x)  // exp(NaN) = NaN, like the intrinsic this stands in for (fminf / fmaxf would turn it into -87)
    return x;
  x              = fminf(fmaxf(x, -87.0f), 88.0f);
  const float kf = rintf(VKGS_MUL(x, 1.44269504088896341f));
  float       r  = VKGS_FMA(-kf, 0.693359375f, x);
  r              = VKGS_FMA(-kf, -2.12194440e-4f, r);
  float p        = 1.9875691500e-4f;
  p              = VKGS_FMA(p, r, 1.3981999507e-3f);
  p              = VKGS_FMA(p, r, 8.3334519073e-3f);
  p              = VKGS_FMA(p, r, 4.1665795894e-2f);
  p              = VKGS_FMA(p, r, 1.6666665459e-1f);
  p              = VKGS_FMA(p, r, 5.0000001201e-1f);
  const float r2 = VKGS_MUL(r, r);
  float       e  = VKGS_FMA(p, r2, r);
  e              = VKGS_ADD(e, 1.0f);
  const int   k  = static_cast<int>(kf);
  return VKGS_MUL(e, VKGS_U2F(static_cast<uint32_t>(k + 127) << 23));
}

// QUANTIZE_NORMALS: encodeNormalOctahedral -> decodeNormalOctahedral (shaders/octahedral_normal.h.slang:28-85), the
// operation order of orc_oct_quantize_normal (oracle). Host + device: tests/test_abi.py pins the host instantiation
// against the oracle bit for bit (callers are compiled without FMA contraction).
__host__ __device__ inline void octQuantizeNormal(float n[3])
{
  // octEncode: project to the octahedron, wrap the bottom hemisphere
  const float inv = 1.0f / ((fabsf(n[0]) + fabsf(n[1])) + fabsf(n[2]));
  float       px = n[0] * inv, py = n[1] * inv;
  if(n[2] < 0.0f)
  {
    const float wx = (1.0f - fabsf(py)) * (px >= 0.0f ? 1.0f : -1.0f), wy = (1.0f - fabsf(px)) * (py >= 0.0f ? 1.0f : -1.0f);
    px = wx, py = wy;
  }
  // octPack / octUnpack: 16 bits per component
  const float    sx = (px * 0.5f + 0.5f) * 65535.0f, sy = (py * 0.5f + 0.5f) * 65535.0f;
  const uint32_t ux = static_cast<uint32_t>(fminf(fmaxf(sx, 0.0f), 65535.0f)), uy = static_cast<uint32_t>(fminf(fmaxf(sy, 0.0f), 65535.0f));
  const float    fx = static_cast<float>(ux) / 65535.0f * 2.0f - 1.0f, fy = static_cast<float>(uy) / 65535.0f * 2.0f - 1.0f;
  // octDecode
  float x = fx, y = fy;
  const float z = (1.0f - fabsf(fx)) - fabsf(fy);
  if(z < 0.0f)
  {
    x = (1.0f - fabsf(fy)) * (fx >= 0.0f ? 1.0f : -1.0f);
    y = (1.0f - fabsf(fx)) * (fy >= 0.0f ? 1.0f : -1.0f);
  }
  const float rn = 1.0f / sqrtf((x * x + y * y) + z * z);
  n[0] = x * rn, n[1] = y * rn, n[2] = z * rn;
}

// Fixed-sequence fp32 trigonometry of the fisheye camera (same operation sequences as orc_atan2f_ypos / orc_acosf /
// orc_sincosf in oracle/vkgs_oracle.c: Cephes single-precision range reductions and polynomials).
__host__ __device__ __forceinline__ float atanNonnegExact(float x)
{
  float y0 = 0.0f;
  if(x > 2.414213562373095f)
  {
    y0 = 1.57079632679489661923f;
    x  = -VKGS_DIV(1.0f, x);
  }
  else if(x > 0.4142135623730950f)
  {
    y0 = 0.78539816339744830962f;
    x  = VKGS_DIV(VKGS_SUB(x, 1.0f), VKGS_ADD(x, 1.0f));
  }
  const float z = VKGS_MUL(x, x);
  float       p = 8.05374449538e-2f;
  p             = VKGS_FMA(p, z, -1.38776856032e-1f);
  p             = VKGS_FMA(p, z, 1.99777106478e-1f);
  p             = VKGS_FMA(p, z, -3.33329491539e-1f);
  return VKGS_ADD(y0, VKGS_FMA(VKGS_MUL(p, z), x, x));
}

__host__ __device__ __forceinline__ float atan2fYposExact(float y, float x)  // y > 0
{
  if(x == 0.0f)
    return 1.57079632679489661923f;
  const float t = VKGS_DIV(y, x);
  return x > 0.0f ? atanNonnegExact(t) : VKGS_SUB(3.14159265358979323846f, atanNonnegExact(-t));
}

__host__ __device__ __forceinline__ float asinSmallExact(float x)  // |x| <= 0.5
{
  const float z = VKGS_MUL(x, x);
  float       p = 4.2163199048e-2f;
  p             = VKGS_FMA(p, z, 2.4181311049e-2f);
  p             = VKGS_FMA(p, z, 4.5470025998e-2f);
  p             = VKGS_FMA(p, z, 7.4953002686e-2f);
  p             = VKGS_FMA(p, z, 1.6666752422e-1f);
  return VKGS_FMA(VKGS_MUL(p, z), x, x);
}

__host__ __device__ __forceinline__ float acosfExact(float x)  // x in [-1, 1]
{
  if(x > 0.5f)
    return VKGS_MUL(2.0f, asinSmallExact(VKGS_SQRT(VKGS_MUL(0.5f, VKGS_SUB(1.0f, x)))));
  if(x < -0.5f)
    return VKGS_SUB(3.14159265358979323846f, VKGS_MUL(2.0f, asinSmallExact(VKGS_SQRT(VKGS_MUL(0.5f, VKGS_ADD(1.0f, x))))));
  return VKGS_SUB(1.57079632679489661923f, asinSmallExact(x));
}

__host__ __device__ __forceinline__ void sincosfExact(float xx, float& sOut, float& cOut)  // |xx| < 8192
{
  float    x = fabsf(xx);
  uint32_t j = static_cast<uint32_t>(VKGS_MUL(x, 1.27323954473516f));
  j          = (j + 1u) & ~1u;
  const float y = static_cast<float>(j);
  x             = VKGS_FMA(-y, 0.78515625f, x);
  x             = VKGS_FMA(-y, 2.4187564849853515625e-4f, x);
  x             = VKGS_FMA(-y, 3.77489497744594108e-8f, x);
  const float z = VKGS_MUL(x, x);
  float       ps = -1.9515295891e-4f;
  ps             = VKGS_FMA(ps, z, 8.3321608736e-3f);
  ps             = VKGS_FMA(ps, z, -1.6666654611e-1f);
  const float sp = VKGS_FMA(VKGS_MUL(ps, z), x, x);
  float       pc = 2.443315711809948e-5f;
  pc             = VKGS_FMA(pc, z, -1.388731625493765e-3f);
  pc             = VKGS_FMA(pc, z, 4.166664568298827e-2f);
  const float cp = VKGS_FMA(pc, VKGS_MUL(z, z), VKGS_FMA(-0.5f, z, 1.0f));
  float       s, c;
  switch((j >> 1) & 3u)
  {
    case 0: s = sp, c = cp; break;
    case 1: s = cp, c = -sp; break;
    case 2: s = -sp, c = -cp; break;
    default: s = -cp, c = sp; break;
  }
  sOut = xx < 0.0f ? -s : s;
  cOut = c;
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
