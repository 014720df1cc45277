// k_blend.cu — tiled software rasterizer: per-pixel gaussian evaluation + ordered alpha compositing.
//
// Replaces the hardware raster + fragment shader + ROP blend of the reference:
//   threedgs_raster.frag.slang:223-311  A = dot(fragPos,fragPos); discard A > 8;
//                                       opacity = exp(-A/2) * a; discard opacity <= 1/255
//   src/gaussian_splatting.cpp:2066-2087 blend state: back-to-front "over" with additive alpha, or
//                                       front-to-back "under" with premultiplied colour
//   colour target cleared to 0 (src/gaussian_splatting.cpp:582), fp32 RGBA here.
// One CTA per 16x16 tile, one thread per pixel, a warp covers an 8x4 pixel block. The tile's
// depth-ordered splat list is consumed in batches of 256: each thread gathers one 48-byte record
// (prefetched into registers one batch ahead), computes which of the 8 warp blocks the splat's
// pixel bounding box touches, and parks both in shared memory; every warp then ballots the batch
// 32 entries at a time and evaluates only the splats that touch its block, in list order — the
// per-pixel blend order is exactly the sorted order, like the ROP.
//
// Exactness: fragPos / A are evaluated with explicit fp32 mul/fma in the oracle's operation order,
// so the `A > 8` discard is bit-exact. opacity uses the SFU ex2 for speed; whenever that value is
// within a guard band of the 1/255 discard threshold it is recomputed with the same fixed-sequence
// expf the oracle uses, so the discard decision is exact too and values differ by a few ulp only.
#include <cuda_fp16.h>

#include "device_common.cuh"
#include "kernels.hpp"

namespace vkgs {

namespace {

// Same operation sequence as orc_expf (oracle/vkgs_oracle.c): Cody-Waite + Cephes polynomial.
__device__ __forceinline__ float expfExact(float x)
{
  x              = fminf(fmaxf(x, -87.0f), 88.0f);
  const float kf = rintf(__fmul_rn(x, 1.44269504088896341f));
  float       r  = __fmaf_rn(-kf, 0.693359375f, x);
  r              = __fmaf_rn(-kf, -2.12194440e-4f, r);
  float p        = 1.9875691500e-4f;
  p              = __fmaf_rn(p, r, 1.3981999507e-3f);
  p              = __fmaf_rn(p, r, 8.3334519073e-3f);
  p              = __fmaf_rn(p, r, 4.1665795894e-2f);
  p              = __fmaf_rn(p, r, 1.6666665459e-1f);
  p              = __fmaf_rn(p, r, 5.0000001201e-1f);
  const float r2 = __fmul_rn(r, r);
  float       e  = __fmaf_rn(p, r2, r);
  e              = __fadd_rn(e, 1.0f);
  const int   k  = static_cast<int>(kf);
  return __fmul_rn(e, __uint_as_float(static_cast<uint32_t>(k + 127) << 23));
}

// Slow path of the blend loop: the SFU opacity landed within the guard band of the discard
// threshold, so the fragment is re-evaluated with the oracle's exp. `negAlpha` is MINUS the splat
// alpha; returns MINUS the fragment opacity, or 0 when the fragment is discarded.
__device__ __noinline__ float exactNegOpacity(float A, float negAlpha)
{
  const float op = __fmul_rn(expfExact(__fmul_rn(-0.5f, A)), -negAlpha);
  return op > 1.0f / 255.0f ? -op : 0.0f;
}

__device__ __forceinline__ float ex2Approx(float x)
{
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}

// shared-memory accesses through explicit 32-bit addresses (keeps address arithmetic out of the
// inner loop: one IMAD per splat)
__device__ __forceinline__ float4 ldsV4(uint32_t addr)
{
  float4 v;
  asm volatile("ld.shared.v4.f32 {%0,%1,%2,%3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "r"(addr));
  return v;
}
__device__ __forceinline__ float2 ldsV2(uint32_t addr)
{
  float2 v;
  asm volatile("ld.shared.v2.f32 {%0,%1}, [%2];" : "=f"(v.x), "=f"(v.y) : "r"(addr));
  return v;
}
__device__ __forceinline__ uint32_t ldsU32(uint32_t addr)
{
  uint32_t v;
  asm volatile("ld.shared.u32 %0, [%1];" : "=r"(v) : "r"(addr));
  return v;
}
__device__ __forceinline__ void stsU32(uint32_t addr, uint32_t v)
{
  asm volatile("st.shared.u32 [%0], %1;" ::"r"(addr), "r"(v) : "memory");
}
// shared window address of the dynamic shared segment, opaque to the optimiser (a plain cvta gets
// re-materialised from SR_CgaCtaId inside the inner loop)
__device__ __forceinline__ uint32_t smemBaseOpaque(const void* p)
{
  uint32_t a;
  asm volatile("{ .reg .u64 t; cvta.to.shared.u64 t, %1; cvt.u32.u64 %0, t; }" : "=r"(a) : "l"(p));
  return a;
}

// ---- packed fp32 pairs (sm_100 FFMA2 / FMUL2 / FADD2): one issue slot, two IEEE-rounded results ----
// (the float2 intrinsics, not inline PTX on b64 registers: ptxas only accumulates in place —
//  FFMA2 Rd = Ra * Rb + Rd — when it sees the value as a float pair)
typedef float2 f32x2;
__device__ __forceinline__ f32x2 pk(float lo, float hi)
{
  return make_float2(lo, hi);
}
__device__ __forceinline__ void upk(f32x2 v, float& lo, float& hi)
{
  lo = v.x, hi = v.y;
}
__device__ __forceinline__ f32x2 fma2(f32x2 a, f32x2 b, f32x2 c)
{
  return __ffma2_rn(a, b, c);
}
__device__ __forceinline__ f32x2 mul2(f32x2 a, f32x2 b)
{
  return __fmul2_rn(a, b);
}
__device__ __forceinline__ f32x2 add2(f32x2 a, f32x2 b)
{
  return __fadd2_rn(a, b);
}
__device__ __forceinline__ void fma2acc(f32x2& c, f32x2 a, f32x2 b)
{
  c = __ffma2_rn(a, b, c);
}
__device__ __forceinline__ void add2acc(f32x2& c, f32x2 a)
{
  c = __fadd2_rn(c, a);
}

// 16-byte asynchronous global -> shared copy (LDGSTS), per-thread addresses
__device__ __forceinline__ void cpAsync16(uint32_t smemDst, const void* gmemSrc)
{
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(smemDst), "l"(gmemSrc) : "memory");
}
__device__ __forceinline__ void cpAsyncCommit()
{
  asm volatile("cp.async.commit_group;" ::: "memory");
}
__device__ __forceinline__ void cpAsyncWaitAll()
{
  asm volatile("cp.async.wait_group 0;" ::: "memory");
}

constexpr uint32_t REC_BYTES   = RECORD_WORDS * 4;  // record, 48 B: cx cy w1x w1y | w2x w2y r g | b a bbox bbox
constexpr int      BLEND_WARPS = BLEND_THREADS / 32;
constexpr int      EPT         = 1;                  // list entries gathered + classified per thread per round
constexpr int      BATCH       = EPT * BLEND_THREADS;  // list entries staged per round
constexpr uint32_t SMEM_REC    = BATCH * REC_BYTES; // bytes of one record buffer
constexpr uint32_t SMEM_HIT    = 2 * SMEM_REC;      // hit masks: [2 buffers][BLEND_WARPS warps][BATCH/32 words]
static_assert(BLEND_WARPS == 4, "the tile is split into 2x2 warp blocks of 8x8 pixels");

// One CTA (4 warps) per 16x16 tile; a warp owns an 8x8 pixel block and every thread TWO pixels of
// it (same column, rows ly and ly+4), evaluated together with packed fp32 instructions: the loads,
// the loop control and the x-dependent products are shared by the pair and every FFMA2 retires two
// IEEE-rounded results in one issue slot (the kernel is issue/latency bound, not FMA-pipe bound).
//
// The tile's list is consumed in batches of 128 entries through a double-buffered shared ring:
// every thread gathers ONE 48-byte record of the next batch straight into shared memory with
// cp.async (no staging registers) while the current batch is blended, then classifies its entry —
// which of the four warp blocks can the splat touch (pixel bbox, then a separating-axis test along
// the splat's own axes against the opacity-limited radius) — and the warp ballots of those bits
// become per-warp hit masks, so the blending warps iterate set bits only. One barrier per batch.
//
// Inner loop: two list entries are evaluated per trip, branch-free up to the blend (discards are
// zeros), so their shared loads, SFU ex2 and dependent FMA chains interleave; only the final
// transmittance update is ordered. Signs are arranged so that no negation is needed per hit:
// the loop works on c - p (A is even in it), carries MINUS the opacity, and accumulates MINUS the
// colour.
template <bool FTB, bool NOGAUSS, bool COUNT, bool SURF>
__global__ void __launch_bounds__(BLEND_THREADS, SURF ? 6 : 10) k_blend(const __grid_constant__ BlendArgs a)
{
  // records ring | hit masks | (surface info only) per-entry (normal, NDC depth) ring | splat-id ring
  constexpr uint32_t SMEM_SURF = SMEM_HIT + 2 * BLEND_WARPS * (BATCH / 32) * 4;
  constexpr uint32_t SMEM_SID  = SMEM_SURF + 2 * BATCH * 16;
  __shared__ __align__(16) unsigned char s_raw[SURF ? SMEM_SID + 2 * BATCH * 4 : SMEM_SURF];
  const uint32_t sbase = smemBaseOpaque(s_raw);

  const unsigned tid = threadIdx.x, lane = tid & 31u, warp = tid >> 5;
  const uint32_t tile = blockIdx.x;
  const uint32_t tx = tile % a.tilesX, ty = tile / a.tilesX;
  const uint32_t tileX0 = tx * TILE_W, tileY0 = ty * TILE_H;
  const uint32_t px = tileX0 + (warp & 1u) * 8u + (lane & 7u), pyA = tileY0 + (warp >> 1) * 8u + (lane >> 3), pyB = pyA + 4u;
  const bool     insideA = px < a.width && pyA < a.height, insideB = px < a.width && pyB < a.height;
  const float    nfx = -(static_cast<float>(px) + 0.5f);
  const f32x2    nfy2 = pk(-(static_cast<float>(pyA) + 0.5f), -(static_cast<float>(pyB) + 0.5f));
  const float    tileCx = static_cast<float>(tileX0) + 4.0f, tileCy = static_cast<float>(tileY0) + 4.0f;  // centre of warp block 0

  const uint2 range = make_uint2(a.rangeBegin[tile], a.rangeEnd[tile]);  // empty tile: begin > end
  f32x2       c0 = pk(0.f, 0.f), c1 = c0, c2 = c0;            // MINUS the colour accumulators of the two pixels
  f32x2       acc = FTB ? pk(1.0f, 1.0f) : pk(0.f, 0.f);      // FTB: transmittance T = 1 - A_dst;  BTF: MINUS the sum of alphas
  const f32x2 one2 = pk(1.0f, 1.0f);
  const f32x2 kExp = pk(-0.72134752044448170368f, -0.72134752044448170368f);  // exp(-A/2) = 2^(-A/2 * log2 e)
  const float THRESHOLD = 1.0f / 255.0f;
  const float BAND      = THRESHOLD * 4e-6f;  // ex2.approx + the two roundings are within 1e-6 relative
  const float eps       = a.transmittanceEpsilon;
  bool        warpDone  = __all_sync(FULL_MASK, !insideA && !insideB);
  uint32_t    nEvaluated = 0, nBlended = 0;  // COUNT only
  // SURF only: MINUS the integrated normals, picked depth and last blended splat id of the two pixels
  f32x2    sn0 = pk(0.f, 0.f), sn1 = sn0, sn2 = sn0;
  float    depthA = 0.0f, depthB = 0.0f;
  uint32_t sidA = 0xffffffffu, sidB = 0xffffffffu;
  uint32_t surfBuf = 0;  // which ring buffer the blending loop is reading

  // asynchronous gather of this thread's entry of a batch into record buffer `buf`
  auto gather = [&](uint32_t id, uint32_t buf, uint32_t slot) {
    const unsigned char* src = reinterpret_cast<const unsigned char*>(a.records + static_cast<uint64_t>(id) * RECORD_WORDS);
    const uint32_t       dst = sbase + buf * SMEM_REC + slot * REC_BYTES;
    cpAsync16(dst, src);
    cpAsync16(dst + 16, src + 16);
    cpAsync16(dst + 32, src + 32);
    if(SURF)
    {
      cpAsync16(sbase + SMEM_SURF + (buf * BATCH + slot) * 16u, a.surface + id);
      stsU32(sbase + SMEM_SID + (buf * BATCH + slot) * 4u, id);
    }
  };
  // classify this thread's (landed) entry: per-warp-block hit bits -> per-warp hit masks
  auto classify = [&](bool have, uint32_t buf, uint32_t slot) {
    uint32_t bits = 0;
    if(have)
    {
      const uint32_t src = sbase + buf * SMEM_REC + slot * REC_BYTES;
      const float4   r0 = ldsV4(src), r1 = ldsV4(src + 16), r2 = ldsV4(src + 32);
      const uint32_t bb0 = __float_as_uint(r2.z), bb1 = __float_as_uint(r2.w);
      const uint32_t x0 = bb0 & 0xffffu, y0 = bb0 >> 16, x1 = bb1 & 0xffffu, y1 = bb1 >> 16;
      const uint32_t colL = (x0 <= tileX0 + 7u && x1 >= tileX0) ? 0x5u : 0u;        // warps 0,2
      const uint32_t colR = (x0 <= tileX0 + 15u && x1 >= tileX0 + 8u) ? 0xau : 0u;  // warps 1,3
      const uint32_t rowT = (y0 <= tileY0 + 7u && y1 >= tileY0) ? 0x3u : 0u;        // warps 0,1
      const uint32_t rowB = (y0 <= tileY0 + 15u && y1 >= tileY0 + 8u) ? 0xcu : 0u;  // warps 2,3
      bits                = (colL | colR) & (rowT | rowB);
      // Separating-axis test along the splat's own axes: over an 8x8 block of pixel centres (half
      // extents 3.5) the fragPos component f_i = dot(p - c, w_i) stays within f_i(centre) +- e_i, and
      // |f_i| > L everywhere means A > L^2 everywhere. A fragment survives only if A <= 8 and
      // exp(-A/2) * alpha > 1/255, i.e. A < 2 ln(255 alpha): L^2 = min(8, 2 ln(255 alpha)), with margins
      // for the approximate log / sqrt and the rounding of f_i.
      float lim = 2.8292f;
      if(!NOGAUSS)
      {
        const float amax = 1.3862943611f * __log2f(255.0f * r2.y) * 1.0001f + 1e-3f;
        lim              = amax > 0.0f ? __fsqrt_rn(fminf(amax, 8.0f)) * 1.0002f + 2e-4f : -1.0f;
      }
      const float e1 = 3.5f * (fabsf(r0.z) + fabsf(r0.w)), e2 = 3.5f * (fabsf(r1.x) + fabsf(r1.y));
#pragma unroll
      for(uint32_t b = 0; b < 4; b++)
      {
        const float ddx = tileCx + static_cast<float>(8u * (b & 1u)) - r0.x, ddy = tileCy + static_cast<float>(8u * (b >> 1)) - r0.y;
        const float f1 = fabsf(ddx * r0.z + ddy * r0.w) - e1, f2 = fabsf(ddx * r1.x + ddy * r1.y) - e2;
        if(!(fmaxf(f1, f2) <= lim))
          bits &= ~(1u << b);
      }
    }
    const unsigned m0 = __ballot_sync(FULL_MASK, bits & 1u), m1 = __ballot_sync(FULL_MASK, bits & 2u),
                   m2 = __ballot_sync(FULL_MASK, bits & 4u), m3 = __ballot_sync(FULL_MASK, bits & 8u);
    if(lane < 4)  // hit[buf][blend warp = lane][word = slot / 32]
      stsU32(sbase + SMEM_HIT + ((buf * BLEND_WARPS + lane) * (BATCH / 32) + (slot >> 5)) * 4u, lane == 0 ? m0 : (lane == 1 ? m1 : (lane == 2 ? m2 : m3)));
  };

  // One list entry against this thread's two pixels, up to (not including) the ordered blend.
  // Returns MINUS the fragment opacities (0 = discarded); `gmin` = distance of the nearer of the two
  // SFU opacities to the discard threshold (the caller sends near misses to the exact path).
  struct Frag
  {
    f32x2  A2, n2;  // A of the two pixels; minus opacity (masked)
    float  gmin;
    float  r, g, b, alpha;
    uint32_t addr;  // shared address of the staged record (SURF: locates the entry's normal / id)
  };
  auto evalFrag = [&](uint32_t addr) {
    Frag         f;
    f.addr = addr;
    const float4 ra  = ldsV4(addr);       // cx cy w1x w1y
    const float4 rb  = ldsV4(addr + 16);  // w2x w2y r g
    const float2 rc  = ldsV2(addr + 32);  // b a
    const float  ndx = __fadd_rn(ra.x, nfx);            // -(px - cx): A is even in (dx,dy)
    const f32x2  ndy2 = add2(pk(ra.y, ra.y), nfy2);
    const float  t = __fmul_rn(ndx, ra.z), u = __fmul_rn(ndx, rb.x);
    const f32x2  fpx2 = fma2(ndy2, pk(ra.w, ra.w), pk(t, t));
    const f32x2  fpy2 = fma2(ndy2, pk(rb.y, rb.y), pk(u, u));
    f.A2              = fma2(fpy2, fpy2, mul2(fpx2, fpx2));
    f.r = rb.z, f.g = rb.w, f.b = rc.x, f.alpha = rc.y;
    float Alo, Ahi;
    upk(f.A2, Alo, Ahi);
    const bool vA = !(Alo > 8.0f), vB = !(Ahi > 8.0f);
    if(NOGAUSS)
    {
      f.n2   = pk(vA ? -1.0f : 0.0f, vB ? -1.0f : 0.0f);
      f.gmin = 1.0f;
    }
    else
    {
      float xlo, xhi, nlo, nhi, glo, ghi;
      upk(mul2(f.A2, kExp), xlo, xhi);
      const float na = -rc.y;
      const f32x2 n2 = mul2(pk(ex2Approx(xlo), ex2Approx(xhi)), pk(na, na));
      upk(n2, nlo, nhi);
      upk(add2(n2, pk(THRESHOLD, THRESHOLD)), glo, ghi);
      f.n2   = pk((vA && nlo < -THRESHOLD) ? nlo : 0.0f, (vB && nhi < -THRESHOLD) ? nhi : 0.0f);
      f.gmin = fminf(fabsf(glo), fabsf(ghi));
    }
    return f;
  };
  // within the guard band of the 1/255 discard threshold (rare): decide exactly, like the oracle
  auto fixFrag = [&](Frag& f) {
    float Alo, Ahi, mlo, mhi;
    upk(f.A2, Alo, Ahi);
    upk(f.n2, mlo, mhi);
    const float opLo = ex2Approx(Alo * -0.72134752044448170368f) * f.alpha, opHi = ex2Approx(Ahi * -0.72134752044448170368f) * f.alpha;
    if(!(Alo > 8.0f) && fabsf(opLo - THRESHOLD) <= BAND)
      mlo = exactNegOpacity(Alo, -f.alpha);
    if(!(Ahi > 8.0f) && fabsf(opHi - THRESHOLD) <= BAND)
      mhi = exactNegOpacity(Ahi, -f.alpha);
    f.n2 = pk(mlo, mhi);
  };
  auto blendFrag = [&](const Frag& f) {
    if(COUNT)
    {
      float mlo, mhi;
      upk(f.n2, mlo, mhi);
      nEvaluated += 1u;
      nBlended += (mlo != 0.0f && insideA ? 1u : 0u) + (mhi != 0.0f && insideB ? 1u : 0u);
    }
    if(FTB)
    {
      const f32x2 nw2 = mul2(f.n2, acc);  // -(opacity * T)
      fma2acc(c0, nw2, pk(f.r, f.r));
      fma2acc(c1, nw2, pk(f.g, f.g));
      fma2acc(c2, nw2, pk(f.b, f.b));
      add2acc(acc, nw2);
      if(SURF)
      {
        // threedgs_raster.frag.slang:316-350: normal * opacity under the same operator; first depth at
        // which the transmittance drops below the iso threshold; id of the last fragment that was kept
        const uint32_t slot = (f.addr - (sbase + surfBuf * SMEM_REC)) / REC_BYTES;
        const float4   sv   = ldsV4(sbase + SMEM_SURF + (surfBuf * BATCH + slot) * 16u);
        const uint32_t id   = ldsU32(sbase + SMEM_SID + (surfBuf * BATCH + slot) * 4u);
        fma2acc(sn0, nw2, pk(sv.x, sv.x));
        fma2acc(sn1, nw2, pk(sv.y, sv.y));
        fma2acc(sn2, nw2, pk(sv.z, sv.z));
        float mlo, mhi, Tlo, Thi;
        upk(f.n2, mlo, mhi);
        upk(acc, Tlo, Thi);
        if(mlo != 0.0f)
        {
          sidA = id;
          if(depthA == 0.0f && Tlo < a.depthIsoThreshold)
            depthA = sv.w;
        }
        if(mhi != 0.0f)
        {
          sidB = id;
          if(depthB == 0.0f && Thi < a.depthIsoThreshold)
            depthB = sv.w;
        }
      }
    }
    else
    {
      const f32x2 t2 = add2(one2, f.n2);  // 1 - opacity
      c0             = mul2(c0, t2);
      c1             = mul2(c1, t2);
      c2             = mul2(c2, t2);
      fma2acc(c0, f.n2, pk(f.r, f.r));
      fma2acc(c1, f.n2, pk(f.g, f.g));
      fma2acc(c2, f.n2, pk(f.b, f.b));
      add2acc(acc, f.n2);
    }
  };

  if(range.x < range.y)
  {
    // prologue: batch 0 lands and is classified; the list index of batch 1 is already on its way
    uint32_t idNext[EPT];
    {
#pragma unroll
      for(int e = 0; e < EPT; e++)
        if(range.x + e * BLEND_THREADS + tid < range.y)
          gather(a.tileVals[range.x + e * BLEND_THREADS + tid], 0, e * BLEND_THREADS + tid);
      cpAsyncCommit();
#pragma unroll
      for(int e = 0; e < EPT; e++)
        idNext[e] = (range.x + BATCH + e * BLEND_THREADS + tid < range.y) ? a.tileVals[range.x + BATCH + e * BLEND_THREADS + tid] : 0u;
      cpAsyncWaitAll();
#pragma unroll
      for(int e = 0; e < EPT; e++)
        classify(range.x + e * BLEND_THREADS + tid < range.y, 0, e * BLEND_THREADS + tid);
    }
    __syncthreads();
    for(uint32_t base = range.x, buf = 0;; base += BATCH, buf ^= 1u)
    {
      const bool more = base + BATCH < range.y;
      if(more)
      {
        // records of the next batch fly into the other buffer while this one is blended
#pragma unroll
        for(int e = 0; e < EPT; e++)
          if(base + BATCH + e * BLEND_THREADS + tid < range.y)
            gather(idNext[e], buf ^ 1u, e * BLEND_THREADS + tid);
        cpAsyncCommit();
#pragma unroll
        for(int e = 0; e < EPT; e++)
          if(base + 2 * BATCH + e * BLEND_THREADS + tid < range.y)
            idNext[e] = a.tileVals[base + 2 * BATCH + e * BLEND_THREADS + tid];
      }

      if(!warpDone)
      {
        const uint32_t recBase = sbase + buf * SMEM_REC;
        surfBuf                = buf;
        const uint32_t hitBase = sbase + SMEM_HIT + (buf * BLEND_WARPS + warp) * (BATCH / 32) * 4u;
#pragma unroll 1
        for(uint32_t chunk = 0; chunk < BATCH / 32; chunk++)
        {
          unsigned       m         = ldsU32(hitBase + chunk * 4u);
          const uint32_t chunkAddr = recBase + chunk * 32u * REC_BYTES;
          while(m)
          {
            const uint32_t addr0 = chunkAddr + (__ffs(m) - 1) * REC_BYTES;
            m &= m - 1;
            if(m)
            {
              const uint32_t addr1 = chunkAddr + (__ffs(m) - 1) * REC_BYTES;
              m &= m - 1;
              Frag f0 = evalFrag(addr0), f1 = evalFrag(addr1);
              if(fminf(f0.gmin, f1.gmin) <= BAND)
              {
                fixFrag(f0);
                fixFrag(f1);
              }
              blendFrag(f0);
              blendFrag(f1);
            }
            else
            {
              Frag f0 = evalFrag(addr0);
              if(f0.gmin <= BAND)
                fixFrag(f0);
              blendFrag(f0);
            }
          }
          if(FTB)
          {
            // all pixels of the block saturated (remaining transmittance below eps) -> stop reading the list
            float Tlo, Thi;
            upk(acc, Tlo, Thi);
            if(__all_sync(FULL_MASK, (Tlo < eps || !insideA) && (Thi < eps || !insideB)))
            {
              warpDone = true;
              break;
            }
          }
        }
      }
      if(more)
      {
        cpAsyncWaitAll();
#pragma unroll
        for(int e = 0; e < EPT; e++)
          classify(base + BATCH + e * BLEND_THREADS + tid < range.y, buf ^ 1u, e * BLEND_THREADS + tid);
      }
      // one barrier per batch: publishes the next batch, retires this one, and votes on whether any
      // warp block of the tile still needs the rest of the list
      const int active = __syncthreads_or(!warpDone);
      if(!more || !active)
        break;
    }
  }

  if(COUNT)
  {
    // list entries this warp block evaluated (one count per warp) and fragments blended (all lanes)
    const uint32_t blended = __reduce_add_sync(FULL_MASK, nBlended);
    if(lane == 0)
    {
      atomicAdd(a.fragmentCounters + 0, static_cast<unsigned long long>(nEvaluated));
      atomicAdd(a.fragmentCounters + 1, static_cast<unsigned long long>(blended));
    }
  }

  // the colour target is rounded ONCE from the fp32 accumulators (the reference's ROP rounds after
  // every blend in the target format)
  float ca[2][4];
  upk(c0, ca[0][0], ca[1][0]);
  upk(c1, ca[0][1], ca[1][1]);
  upk(c2, ca[0][2], ca[1][2]);
  upk(acc, ca[0][3], ca[1][3]);
#pragma unroll
  for(int p = 0; p < 2; p++)
  {
    if(!(p ? insideB : insideA))
      continue;
    const uint64_t o  = static_cast<uint64_t>(p ? pyB : pyA) * a.width + px;
    // (0 - x, not -x: an untouched pixel must come out as +0)
    const float    c0f = __fsub_rn(0.0f, ca[p][0]), c1f = __fsub_rn(0.0f, ca[p][1]), c2f = __fsub_rn(0.0f, ca[p][2]);
    const float    al = FTB ? 1.0f - ca[p][3] : __fsub_rn(0.0f, ca[p][3]);
    if(SURF)
    {
      float nx[2], ny[2], nz[2];
      upk(sn0, nx[0], nx[1]);
      upk(sn1, ny[0], ny[1]);
      upk(sn2, nz[0], nz[1]);
      a.outNormals[o] = make_float4(__fsub_rn(0.0f, nx[p]), __fsub_rn(0.0f, ny[p]), __fsub_rn(0.0f, nz[p]), al);
      a.outDepthT[o]  = make_float2(p ? depthB : depthA, ca[p][3]);
      a.outSplatId[o] = p ? sidB : sidA;
    }
    if(a.targetFormat == VKGS_FORMAT_FLOAT32)
      static_cast<float4*>(a.image)[o] = make_float4(c0f, c1f, c2f, al);
    else if(a.targetFormat == VKGS_FORMAT_FLOAT16)
    {
      const __half2 lo = __floats2half2_rn(c0f, c1f), hi = __floats2half2_rn(c2f, al);
      static_cast<uint2*>(a.image)[o] = make_uint2(*reinterpret_cast<const uint32_t*>(&lo), *reinterpret_cast<const uint32_t*>(&hi));
    }
    else
    {
      // R8G8B8A8_UNORM: clamp to [0,1], scale, round to nearest
      const uint32_t r = __float2uint_rn(__saturatef(c0f) * 255.0f), gch = __float2uint_rn(__saturatef(c1f) * 255.0f),
                     b = __float2uint_rn(__saturatef(c2f) * 255.0f), aa = __float2uint_rn(__saturatef(al) * 255.0f);
      static_cast<uint32_t*>(a.image)[o] = r | (gch << 8) | (b << 16) | (aa << 24);
    }
  }
}

}  // namespace

void launchBlend(const BlendArgs& args, cudaStream_t stream)
{
  const uint32_t tiles = args.tilesX * args.tilesY;
  const bool count = args.fragmentCounters != nullptr;
  if(args.outNormals)
  {
    // surface-info variant: front to back only (the context rejects other combinations)
    if(args.disableOpacityGaussian)
      k_blend<true, true, false, true><<<tiles, BLEND_THREADS, 0, stream>>>(args);
    else
      k_blend<true, false, false, true><<<tiles, BLEND_THREADS, 0, stream>>>(args);
    return;
  }
#define VKGS_BLEND_LAUNCH(F, G)                                                                                                  \
  do                                                                                                                             \
  {                                                                                                                              \
    if(count)                                                                                                                    \
      k_blend<F, G, true, false><<<tiles, BLEND_THREADS, 0, stream>>>(args);                                                            \
    else                                                                                                                         \
      k_blend<F, G, false, false><<<tiles, BLEND_THREADS, 0, stream>>>(args);                                                           \
  } while(0)
  if(args.frontToBack)
  {
    if(args.disableOpacityGaussian)
      VKGS_BLEND_LAUNCH(true, true);
    else
      VKGS_BLEND_LAUNCH(true, false);
  }
  else
  {
    if(args.disableOpacityGaussian)
      VKGS_BLEND_LAUNCH(false, true);
    else
      VKGS_BLEND_LAUNCH(false, false);
  }
#undef VKGS_BLEND_LAUNCH
}

}  // namespace vkgs
