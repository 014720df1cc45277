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
__device__ __forceinline__ void stsV4(uint32_t addr, float4 v)
{
  asm volatile("st.shared.v4.f32 [%0], {%1,%2,%3,%4};" ::"r"(addr), "f"(v.x), "f"(v.y), "f"(v.z), "f"(v.w) : "memory");
}
__device__ __forceinline__ void stsV2(uint32_t addr, float2 v)
{
  asm volatile("st.shared.v2.f32 [%0], {%1,%2};" ::"r"(addr), "f"(v.x), "f"(v.y) : "memory");
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

constexpr uint32_t REC_BYTES   = RECORD_WORDS * 4;  // staged record, 48 B: cx -cy w1x w1y | w2x w2y -r -g | -b -a . .
constexpr int      BLEND_WARPS = BLEND_THREADS / 32;
constexpr int      BATCH       = BLEND_THREADS;     // list entries staged per round (one per thread)
constexpr uint32_t SMEM_REC    = BATCH * REC_BYTES; // bytes of one record buffer
constexpr uint32_t SMEM_HIT    = 2 * SMEM_REC;      // hit masks: [2 buffers][BLEND_WARPS warps][BATCH/32 words]
static_assert(BLEND_WARPS == 4 && BATCH == 128, "the tile is split into 2x2 warp blocks of 8x8 pixels");

// One CTA (4 warps) per 16x16 tile; a warp owns an 8x8 pixel block and every thread TWO pixels of
// it (same column, rows ly and ly+4), evaluated together with packed fp32 instructions: the loads,
// the loop control and the x-dependent products are shared by the pair and every FFMA2 retires two
// IEEE-rounded results in one issue slot (the kernel is issue-bound, not FMA-pipe bound).
// The list is consumed in batches of 128 entries through a double-buffered shared staging area
// (one barrier per batch). The staging thread of an entry also decides which of the four warp
// blocks the splat can touch (pixel bbox, then a separating-axis test along the splat's own axes
// against the opacity-limited radius) and the warp ballots of those bits become per-warp hit masks:
// the blending warps iterate set bits only.
template <bool FTB, bool NOGAUSS>
__global__ void __launch_bounds__(BLEND_THREADS, 8) k_blend(const __grid_constant__ BlendArgs a)
{
  __shared__ __align__(16) unsigned char s_raw[2 * SMEM_REC + 2 * BLEND_WARPS * (BATCH / 32) * 4];
  const uint32_t sbase = smemBaseOpaque(s_raw);

  const unsigned tid = threadIdx.x, lane = tid & 31u, warp = tid >> 5;
  const uint32_t tile = blockIdx.x;
  const uint32_t tx = tile % a.tilesX, ty = tile / a.tilesX;
  const uint32_t tileX0 = tx * TILE_W, tileY0 = ty * TILE_H;
  const uint32_t px = tileX0 + (warp & 1u) * 8u + (lane & 7u), pyA = tileY0 + (warp >> 1) * 8u + (lane >> 3), pyB = pyA + 4u;
  const bool     insideA = px < a.width && pyA < a.height, insideB = px < a.width && pyB < a.height;
  const float    fx = static_cast<float>(px) + 0.5f;
  const f32x2    fy2 = pk(static_cast<float>(pyA) + 0.5f, static_cast<float>(pyB) + 0.5f);
  const float    tileCx = static_cast<float>(tileX0) + 4.0f, tileCy = static_cast<float>(tileY0) + 4.0f;  // centre of warp block 0

  const uint2 range = make_uint2(a.rangeBegin[tile], a.rangeEnd[tile]);  // empty tile: begin > end
  f32x2       c0 = pk(0.f, 0.f), c1 = c0, c2 = c0;          // colour accumulators of the two pixels
  f32x2       acc = FTB ? pk(1.0f, 1.0f) : pk(0.f, 0.f);        // FTB: transmittance T = 1 - A_dst;  BTF: MINUS the sum of alphas
  const f32x2 one2 = pk(1.0f, 1.0f);
  const f32x2 kExp = pk(-0.72134752044448170368f, -0.72134752044448170368f);  // exp(-A/2) = 2^(-A/2 * log2 e)
  const float THRESHOLD = 1.0f / 255.0f;
  const float BAND      = THRESHOLD * 4e-6f;  // ex2.approx + the two roundings are within 1e-6 relative
  const float eps       = a.transmittanceEpsilon;
  bool        warpDone  = __all_sync(FULL_MASK, !insideA && !insideB);

  // gather of this thread's entry of a batch (record, prefetched one batch ahead)
  float4 r0, r1, r2;
  auto   fetch = [&](uint32_t base) {
    if(base + tid < range.y)
    {
      const uint32_t id  = a.tileVals[base + tid];
      const float4*  rec = reinterpret_cast<const float4*>(a.records + static_cast<uint64_t>(id) * RECORD_WORDS);
      r0                 = __ldg(rec + 0);
      r1                 = __ldg(rec + 1);
      r2                 = __ldg(rec + 2);
    }
  };
  // park the entry in shared memory (sign conventions of the inner loop) + per-warp-block hit bits
  auto stage = [&](uint32_t base, uint32_t buf) {
    uint32_t bits = 0;
    if(base + tid < range.y)
    {
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
      if(bits)
      {
        const uint32_t dst = sbase + buf * SMEM_REC + tid * REC_BYTES;
        stsV4(dst, make_float4(r0.x, -r0.y, r0.z, r0.w));
        stsV4(dst + 16, make_float4(r1.x, r1.y, -r1.z, -r1.w));
        stsV2(dst + 32, make_float2(-r2.x, -r2.y));
      }
    }
    const unsigned m0 = __ballot_sync(FULL_MASK, bits & 1u), m1 = __ballot_sync(FULL_MASK, bits & 2u),
                   m2 = __ballot_sync(FULL_MASK, bits & 4u), m3 = __ballot_sync(FULL_MASK, bits & 8u);
    if(lane < 4)  // hit[buf][blend warp = lane][word = this staging warp]
      stsU32(sbase + SMEM_HIT + ((buf * BLEND_WARPS + lane) * (BATCH / 32) + warp) * 4u, lane == 0 ? m0 : (lane == 1 ? m1 : (lane == 2 ? m2 : m3)));
  };

  if(range.x < range.y)
  {
    fetch(range.x);
    stage(range.x, 0);
    __syncthreads();
    for(uint32_t base = range.x, buf = 0;; base += BATCH, buf ^= 1u)
    {
      const bool more = base + BATCH < range.y;
      if(more)
        fetch(base + BATCH);  // global gathers of the next batch fly while this one is blended

      if(!warpDone)
      {
        const uint32_t recBase = sbase + buf * SMEM_REC;
        const uint32_t hitBase = sbase + SMEM_HIT + (buf * BLEND_WARPS + warp) * (BATCH / 32) * 4u;
        for(uint32_t chunk = 0; chunk < BATCH / 32; chunk++)
        {
          unsigned       m         = ldsU32(hitBase + chunk * 4u);
          const uint32_t chunkAddr = recBase + chunk * 32u * REC_BYTES;
          while(m)
          {
            const uint32_t addr = chunkAddr + (__ffs(m) - 1) * REC_BYTES;
            m &= m - 1;
            const float4 ra  = ldsV4(addr);       // cx -cy w1x w1y
            const float4 rb  = ldsV4(addr + 16);  // w2x w2y -r -g
            const float  dx  = __fsub_rn(fx, ra.x);
            const f32x2  dy2 = add2(fy2, pk(ra.y, ra.y));
            const float  t = __fmul_rn(dx, ra.z), u = __fmul_rn(dx, rb.x);
            const f32x2  fpx2 = fma2(dy2, pk(ra.w, ra.w), pk(t, t));
            const f32x2  fpy2 = fma2(dy2, pk(rb.y, rb.y), pk(u, u));
            const f32x2  A2   = fma2(fpy2, fpy2, mul2(fpx2, fpx2));
            float        Alo, Ahi;
            upk(A2, Alo, Ahi);
            const bool vA = !(Alo > 8.0f), vB = !(Ahi > 8.0f);
            if(!(vA || vB))
              continue;
            const float2 rc = ldsV2(addr + 32);  // -b -a
            // MINUS the fragment opacity of the two pixels, 0 for a discarded fragment
            float mlo = vA ? -1.0f : 0.0f, mhi = vB ? -1.0f : 0.0f;
            if(!NOGAUSS)
            {
              float xlo, xhi, nlo, nhi, glo, ghi;
              upk(mul2(A2, kExp), xlo, xhi);
              const f32x2 n2 = mul2(pk(ex2Approx(xlo), ex2Approx(xhi)), pk(rc.y, rc.y));
              upk(n2, nlo, nhi);
              upk(add2(n2, pk(THRESHOLD, THRESHOLD)), glo, ghi);
              mlo = (vA && nlo < -THRESHOLD) ? nlo : 0.0f;
              mhi = (vB && nhi < -THRESHOLD) ? nhi : 0.0f;
              if(fminf(fabsf(glo), fabsf(ghi)) <= BAND)
              {
                // within the guard band of the 1/255 discard threshold (rare): decide exactly
                if(vA && fabsf(glo) <= BAND)
                  mlo = exactNegOpacity(Alo, rc.y);
                if(vB && fabsf(ghi) <= BAND)
                  mhi = exactNegOpacity(Ahi, rc.y);
              }
            }
            const f32x2 nop2 = pk(mlo, mhi);
            if(FTB)
            {
              const f32x2 nw2 = mul2(nop2, acc);  // -(opacity * T)
              fma2acc(c0, nw2, pk(rb.z, rb.z));
              fma2acc(c1, nw2, pk(rb.w, rb.w));
              fma2acc(c2, nw2, pk(rc.x, rc.x));
              add2acc(acc, nw2);
            }
            else
            {
              const f32x2 t2 = add2(one2, nop2);  // 1 - opacity
              c0             = mul2(c0, t2);
              c1             = mul2(c1, t2);
              c2             = mul2(c2, t2);
              fma2acc(c0, nop2, pk(rb.z, rb.z));
              fma2acc(c1, nop2, pk(rb.w, rb.w));
              fma2acc(c2, nop2, pk(rc.x, rc.x));
              add2acc(acc, nop2);
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
        stage(base + BATCH, buf ^ 1u);
      // one barrier per batch: publishes the next staged batch, retires this one, and votes on
      // whether any warp block of the tile still needs the rest of the list
      const int active = __syncthreads_or(!warpDone);
      if(!more || !active)
        break;
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
    const float    c0f = ca[p][0], c1f = ca[p][1], c2f = ca[p][2];
    const float    al = FTB ? 1.0f - ca[p][3] : -ca[p][3];
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
  if(args.frontToBack)
  {
    if(args.disableOpacityGaussian)
      k_blend<true, true><<<tiles, BLEND_THREADS, 0, stream>>>(args);
    else
      k_blend<true, false><<<tiles, BLEND_THREADS, 0, stream>>>(args);
  }
  else
  {
    if(args.disableOpacityGaussian)
      k_blend<false, true><<<tiles, BLEND_THREADS, 0, stream>>>(args);
    else
      k_blend<false, false><<<tiles, BLEND_THREADS, 0, stream>>>(args);
  }
}

}  // namespace vkgs
