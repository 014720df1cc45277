// k_blend.cu — tiled software rasterizer: per-pixel gaussian evaluation + ordered alpha compositing.
//
// Replaces the hardware raster + fragment shader + ROP blend of the reference:
//   threedgs_raster.frag.slang:223-311  A = dot(fragPos,fragPos); discard A > 8;
//                                       opacity = exp(-A/2) * a; discard opacity <= 1/255
//   src/gaussian_splatting.cpp:2066-2087 blend state: back-to-front "over" with additive alpha, or
//                                       front-to-back "under" with premultiplied colour
//   colour target cleared to 0 (src/gaussian_splatting.cpp:582), fp32 RGBA here.
// Binning tiles are 32x32 pixels; each is blended by TILE_H / BLEND_H CTAs (32x16 bands) that read the
// tile's depth-ordered splat list; a warp owns an 8x8 pixel block and every thread two pixels of it
// (packed fp32). The list is consumed in batches of 128 through a shared ring filled with cp.async
// gathers of the 48-byte records; the gathering thread classifies its entry against the CTA's warp
// blocks and every warp then evaluates only the entries that can touch its block, in list order — the
// per-pixel blend order is exactly the sorted order, like the ROP. Variants (template flags): fragment
// counters (profiling), the surface-info side outputs, and the VK3DGUT fragment stage.
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

__device__ __forceinline__ float rcpApprox(float x)
{
  float y;
  asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
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

// record, 48 B: cx cy w1x w1y | w2x w2y r g | b a bbox bbox   (3DGUT: 96 B, see k_preprocess.cu)
constexpr int      BLEND_WARPS = BLEND_THREADS / 32;
constexpr int      BATCH       = 128;  // list entries staged per round, one per thread of the first BATCH/32 warps. Smaller than
                                       // the CTA on purpose: with early termination most tiles finish inside their first
                                       // batches, and whatever is classified beyond that point is wasted
static_assert(BATCH % 32 == 0 && BATCH <= BLEND_THREADS, "whole staging warps");
constexpr int      BLOCKS_X    = TILE_W / 8;  // the tile is split into BLOCKS_X x BLOCKS_Y warp blocks of 8x8 pixels
constexpr int      BLOCKS_Y    = BLEND_H / 8;
constexpr int      BANDS       = TILE_H / BLEND_H;  // CTAs per tile (consecutive block indices: they share the list in L2)
static_assert(BLOCKS_X * BLOCKS_Y == BLEND_WARPS && TILE_W % 8 == 0 && BLEND_H % 8 == 0 && TILE_H % BLEND_H == 0, "one warp per 8x8 pixel block");

// dynamic shared memory of a blend CTA: records ring, hit masks, (surface info) normal + id rings, (3DGUT) instance index
// ring + the world-space ray directions of every thread's two pixels (multi-instance scenes only)
__host__ __device__ constexpr uint32_t blendSmemBytes(bool surf, bool gut)
{
  return 2u * BATCH * (gut ? GUT_RECORD_WORDS * 4u + 48u : RECORD_WORDS * 4u) + 2u * BLEND_WARPS * (BATCH / 32) * 4u + (surf ? 2u * BATCH * 20u : 0u)
         + (gut ? 2u * BATCH * 4u + BLEND_THREADS * 32u : 0u);
}

// ---- 3DGUT fragment helpers -----------------------------------------------------------------------------------------
// pixel centre inside the quad: |p - c| <= extent (EXTENT_CONIC), or |dot(p - c, w_i)| <= 1 (EXTENT_EIGEN; q0.zw = w1,
// k = |w2| / |w1|, w2 = k (w1.y, -w1.x)). A pixel outside the fisheye field of view carries y = 1e30 and fails here.
__device__ __forceinline__ bool gutInsideQuad(bool eigen, const float4& q0, float k, float pxc, float pyc)
{
  const float ddx = __fsub_rn(pxc, q0.x), ddy = __fsub_rn(pyc, q0.y);
  if(!eigen)
    return fabsf(ddx) <= q0.z && fabsf(ddy) <= q0.w;
  const float w2x = q0.w * k, w2y = -q0.z * k;
  return fabsf(ddx * q0.z + ddy * q0.w) <= 1.0f && fabsf(ddx * w2x + ddy * w2y) <= 1.0f;
}

// Exact evaluation of one pixel against the record at shared address `addr` (the oracle's operation order and exp;
// dm = ray direction in the model space of the entry's instance): the reference path for every kernel degree, and the
// arbiter of the quadratic kernel's fast path. Returns MINUS the opacity, 0 when the fragment is discarded.
// Not inlined: it is rare on the default path, and eight inlined copies (two hits x two pixels x two call sites)
// would push the blend loop out of the instruction cache.
template <bool NOGAUSS>
__device__ __noinline__ float gutExactPixel(const GutFrameConstants& g, uint32_t addr, float pxc, float pyc, float dm0, float dm1, float dm2)
{
  const float4 q0 = ldsV4(addr), q1 = ldsV4(addr + 16), q2 = ldsV4(addr + 32), q3 = ldsV4(addr + 48), q4 = ldsV4(addr + 64), q5 = ldsV4(addr + 80);
  bool  ok = gutInsideQuad(g.extentEigen != 0u, q0, q2.w, pxc, pyc) && !(q1.w <= g.alphaCullThreshold);
  const float rd0 = __fmul_rn(q3.x, __fadd_rn(__fadd_rn(__fmul_rn(dm0, q3.w), __fmul_rn(dm1, q4.z)), __fmul_rn(dm2, q5.y)));
  const float rd1 = __fmul_rn(q3.y, __fadd_rn(__fadd_rn(__fmul_rn(dm0, q4.x), __fmul_rn(dm1, q4.w)), __fmul_rn(dm2, q5.z)));
  const float rd2 = __fmul_rn(q3.z, __fadd_rn(__fadd_rn(__fmul_rn(dm0, q4.y), __fmul_rn(dm1, q5.x)), __fmul_rn(dm2, q5.w)));
  const float rn  = __fdiv_rn(1.0f, __fsqrt_rn(__fadd_rn(__fadd_rn(__fmul_rn(rd0, rd0), __fmul_rn(rd1, rd1)), __fmul_rn(rd2, rd2))));
  const float d0 = __fmul_rn(rd0, rn), d1 = __fmul_rn(rd1, rn), d2 = __fmul_rn(rd2, rn);
  const float c0x = __fsub_rn(__fmul_rn(d1, q2.z), __fmul_rn(d2, q2.y)), c1x = __fsub_rn(__fmul_rn(d2, q2.x), __fmul_rn(d0, q2.z)),
              c2x = __fsub_rn(__fmul_rn(d0, q2.y), __fmul_rn(d1, q2.x));
  const float dist = __fadd_rn(__fadd_rn(__fmul_rn(c0x, c0x), __fmul_rn(c1x, c1x)), __fmul_rn(c2x, c2x));
  float       resp;
  switch(g.kernelDegree)
  {
    case 8: {
      const float d2s = __fmul_rn(dist, dist);
      resp            = expfExact(__fmul_rn(__fmul_rn(-0.000685871056241f, d2s), d2s));
      break;
    }
    case 5:
      resp = expfExact(__fmul_rn(__fmul_rn(__fmul_rn(-0.0185185185185f, dist), dist), __fsqrt_rn(dist)));
      break;
    case 4:
      resp = expfExact(__fmul_rn(__fmul_rn(-0.0555555555556f, dist), dist));
      break;
    case 3:
      resp = expfExact(__fmul_rn(__fmul_rn(-0.166666666667f, dist), __fsqrt_rn(dist)));
      break;
    case 1:
      resp = expfExact(__fmul_rn(-1.5f, __fsqrt_rn(dist)));
      break;
    case 0:
      resp = fmaxf(__fadd_rn(1.0f, __fmul_rn(-0.329630334487f, __fsqrt_rn(dist))), 0.0f);
      break;
    default:
      resp = expfExact(__fmul_rn(-0.5f, dist));
      break;
  }
  const float alpha = fminf(g.alphaClamp, __fmul_rn(resp, q1.w));
  ok                = ok && alpha > 1.0f / 255.0f && resp > g.kernelMinResponse;
  return ok ? (NOGAUSS ? -1.0f : -alpha) : 0.0f;
}

// One CTA (BLEND_WARPS warps) per band of a tile; a warp owns an 8x8 pixel block and every thread TWO pixels of
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
//
// Order. Every pixel sees its fragments in list order. Back-to-front frames (the reference default) are
// composited from the END of the list with the same "under" update as front-to-back frames — identical
// in exact arithmetic to the reference's "over" blend — and keep the reference's additive alpha (the sum of
// the fragment opacities); so both orders can stop at a saturated pixel (transmittance_epsilon).
//
// Discards. The fragment survives iff A <= 8 and exp(-A/2) * alpha > 1/255, i.e. A < A* = 2 ln(255 alpha).
// The staging thread leaves cut = min(8, A* - band) and A* in the staged record (over the bbox words, which
// only the classification reads): A <= cut is a sure keep, and only a pixel whose A lies within the band of
// A* is sent to the exact path (the oracle's fixed-sequence exp), so both discards cost one comparison.
#ifndef VKGS_BLEND_RESIDENT_THREADS
#define VKGS_BLEND_RESIDENT_THREADS 1280
#endif
#ifndef VKGS_GUT_RESIDENT_THREADS
#define VKGS_GUT_RESIDENT_THREADS 1024
#endif
template <bool FTB, bool NOGAUSS, bool COUNT, bool SURF, bool GUT, bool GUTX = false>
__global__ void __launch_bounds__(BLEND_THREADS, (GUT ? VKGS_GUT_RESIDENT_THREADS : SURF ? 768 : VKGS_BLEND_RESIDENT_THREADS) / BLEND_THREADS) k_blend(const __grid_constant__ BlendArgs a)
{
  // records ring | hit masks | (surface info only) per-entry (normal, NDC depth) ring | splat-id ring
  constexpr uint32_t REC_BYTES = (GUT ? GUT_RECORD_WORDS : RECORD_WORDS) * 4;
  // a staged entry: the record; 3DGUT: + 48 bytes the staging thread computes for the quadratic kernel's fast path (the
  // ray-to-canonical-space matrix with the scale folded in, the discard threshold in the particle's squared distance
  // and its guard band, see classify / evalFrag)
  constexpr uint32_t SLOT_BYTES = GUT ? REC_BYTES + 48 : REC_BYTES;
  constexpr uint32_t SMEM_REC  = BATCH * SLOT_BYTES;  // bytes of one record buffer
  constexpr uint32_t SMEM_HIT  = 2 * SMEM_REC;       // hit masks: [2 buffers][BLEND_WARPS warps][BATCH/32 words]
  constexpr uint32_t SMEM_SURF = SMEM_HIT + 2 * BLEND_WARPS * (BATCH / 32) * 4;
  constexpr uint32_t SMEM_SID  = SMEM_SURF + 2 * BATCH * 16;
  constexpr uint32_t SMEM_INST = SURF ? SMEM_SID + 2 * BATCH * 4 : SMEM_SURF;  // 3DGUT: instance index of every staged entry
  constexpr uint32_t SMEM_WORLD = SMEM_INST + 2 * BATCH * 4;  // 3DGUT: world-space ray directions, 32 bytes per thread
  static_assert(SMEM_INST + (GUT ? 2 * BATCH * 4 + BLEND_THREADS * 32 : 0) == blendSmemBytes(SURF, GUT), "launch-side size");
  extern __shared__ __align__(16) unsigned char s_raw[];
  const uint32_t sbase = smemBaseOpaque(s_raw);

  const unsigned tid = threadIdx.x, lane = tid & 31u, warp = tid >> 5;
  const uint32_t tile = a.firstTile + blockIdx.x / BANDS, band = blockIdx.x % BANDS;
  const uint32_t tx = tile % a.tilesX, ty = tile / a.tilesX;
  const uint32_t tileX0 = tx * TILE_W, tileY0 = ty * TILE_H + band * BLEND_H;  // origin of this CTA's band
  const uint32_t px = tileX0 + (warp % BLOCKS_X) * 8u + (lane & 7u), pyA = tileY0 + (warp / BLOCKS_X) * 8u + (lane >> 3), pyB = pyA + 4u;
  const bool     insideA = px < a.width && pyA < a.height, insideB = px < a.width && pyB < a.height;
  const float    nfx = -(static_cast<float>(px) + 0.5f);
  const f32x2    nfy2 = pk(-(static_cast<float>(pyA) + 0.5f), -(static_cast<float>(pyB) + 0.5f));
  const float    tileCx = static_cast<float>(tileX0) + 4.0f, tileCy = static_cast<float>(tileY0) + 4.0f;  // centre of warp block 0
  // 3DGUT: the model-space ray directions of this thread's two pixels (generatePinholeRay, cameras.h.slang:27-44,
  // called with SV_Position.xy AND a sub-pixel offset of 0.5, threedgut_raster.frag.slang:92 — restated as written;
  // then threedgut_raster.frag.slang:117-121), in the operation order of orc_gut_fragment
  float gutDirA[3] = {0.f, 0.f, 0.f}, gutDirB[3] = {0.f, 0.f, 0.f};      // model space of instance 0
  float gutPyA = static_cast<float>(pyA) + 0.5f, gutPyB = static_cast<float>(pyB) + 0.5f;
  bool  gutFovA = true, gutFovB = true;  // fisheye: pixel inside the field of view (else every fragment is discarded)
  if(GUT)
  {
#pragma unroll
    for(int p = 0; p < 2; p++)
    {
      float tgt[4], dir[4];
      if(GUTX && a.gut.fisheye)
      {
        // generateFisheyeRay(SV_Position.xy, viewport, fovRad, 0, viewInverse) (cameras.h.slang:47-82), in the
        // operation order of orc_gut_fragment; direction goes through viewInverse below like the pinhole target
        const float fpx = static_cast<float>(px) + 0.5f, fpy = p ? gutPyB : gutPyA;
        const float u = __fsub_rn(__fmul_rn(__fdiv_rn(fpx, __fsub_rn(a.gut.viewport[0], 1.0f)), 2.0f), 1.0f);
        const float v = __fsub_rn(__fmul_rn(__fdiv_rn(fpy, __fsub_rn(a.gut.viewport[1], 1.0f)), 2.0f), 1.0f);
        const float r = __fsqrt_rn(__fadd_rn(__fmul_rn(u, u), __fmul_rn(v, v)));
        (p ? gutFovB : gutFovA) = !(r > 1.0f);
        float phiCos = fabsf(r) > 1e-9f ? __fdiv_rn(u, r) : 0.0f;
        phiCos       = fminf(fmaxf(phiCos, -1.0f), 1.0f);
        float phi    = acosfExact(phiCos);
        phi          = v < 0.0f ? -phi : phi;
        const float theta = __fmul_rn(__fmul_rn(r, a.gut.fovRad), 0.5f);
        float       sphi, cphi, sth, cth;
        sincosfExact(phi, sphi, cphi);
        sincosfExact(theta, sth, cth);
        tgt[0] = __fmul_rn(cphi, sth), tgt[1] = __fmul_rn(-sphi, sth), tgt[2] = -cth, tgt[3] = 0.0f;
      }
      else
      {
        const float pcx = __fadd_rn(static_cast<float>(px) + 0.5f, 0.5f), pcy = __fadd_rn(p ? gutPyB : gutPyA, 0.5f);
        const float dx = __fsub_rn(__fmul_rn(__fdiv_rn(pcx, a.gut.viewport[0]), 2.0f), 1.0f);
        const float dy = __fsub_rn(__fmul_rn(__fdiv_rn(pcy, a.gut.viewport[1]), 2.0f), 1.0f);
        const float t4[4] = {dx, dy, 1.0f, 1.0f};
#pragma unroll
        for(int j = 0; j < 4; j++)
          tgt[j] = __fadd_rn(__fadd_rn(__fadd_rn(__fmul_rn(t4[0], a.gut.projInverse[0 + j]), __fmul_rn(t4[1], a.gut.projInverse[4 + j])),
                                       __fmul_rn(t4[2], a.gut.projInverse[8 + j])),
                             __fmul_rn(t4[3], a.gut.projInverse[12 + j]));
      }
#pragma unroll
      for(int j = 0; j < 4; j++)
        dir[j] = __fadd_rn(__fadd_rn(__fadd_rn(__fmul_rn(tgt[0], a.gut.viewInverse[0 + j]), __fmul_rn(tgt[1], a.gut.viewInverse[4 + j])),
                                     __fmul_rn(tgt[2], a.gut.viewInverse[8 + j])),
                           __fmul_rn(0.0f, a.gut.viewInverse[12 + j]));
      const float dn = __fdiv_rn(1.0f, __fsqrt_rn(__fadd_rn(__fadd_rn(__fadd_rn(__fmul_rn(dir[0], dir[0]), __fmul_rn(dir[1], dir[1])), __fmul_rn(dir[2], dir[2])),
                                                           __fmul_rn(dir[3], dir[3]))));
      const float rd[3] = {__fmul_rn(dir[0], dn), __fmul_rn(dir[1], dn), __fmul_rn(dir[2], dn)};
      // multi-instance scenes re-derive the model-space direction per entry from the world-space one: parked in shared
      // memory (six more live registers would spill the blend loop of the common single-instance case)
      if(GUTX && a.gut.instanceCount > 1u)
      {
#pragma unroll
        for(int j = 0; j < 3; j++)
          stsU32(sbase + SMEM_WORLD + tid * 32u + (3u * p + j) * 4u, __float_as_uint(rd[j]));
      }
      float       dm[3];
#pragma unroll
      for(int j = 0; j < 3; j++)
        dm[j] = __fadd_rn(__fadd_rn(__fmul_rn(rd[0], a.gut.modelInverse[0 + j]), __fmul_rn(rd[1], a.gut.modelInverse[4 + j])),
                          __fmul_rn(rd[2], a.gut.modelInverse[8 + j]));
      const float dmn = __fdiv_rn(1.0f, __fsqrt_rn(__fadd_rn(__fadd_rn(__fmul_rn(dm[0], dm[0]), __fmul_rn(dm[1], dm[1])), __fmul_rn(dm[2], dm[2]))));
#pragma unroll
      for(int j = 0; j < 3; j++)
        (p ? gutDirB : gutDirA)[j] = __fmul_rn(dm[j], dmn);
    }
    // outside the field of view every fragment is discarded (threedgut_raster.frag.slang:96-100): such a pixel fails
    // every quad test from here on
    if(GUTX && !gutFovA)
      gutPyA = 1e30f;
    if(GUTX && !gutFovB)
      gutPyB = 1e30f;
  }

  const uint2 range = make_uint2(a.rangeBegin[tile], a.rangeEnd[tile]);  // empty tile: begin > end
  f32x2       c0 = pk(0.f, 0.f), c1 = c0, c2 = c0;            // MINUS the colour accumulators of the two pixels
  f32x2       acc = pk(1.0f, 1.0f);                           // transmittance T of the two pixels (both compositing orders)
  f32x2       asum = pk(0.f, 0.f);                            // back-to-front frames: MINUS the sum of the fragment opacities
  const f32x2 kExp = pk(-0.72134752044448170368f, -0.72134752044448170368f);  // exp(-A/2) = 2^(-A/2 * log2 e)
  const float BAND      = 3e-5f;  // half width of the band of A around A* = 2 ln(255 alpha) that the exact path decides
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
    const unsigned char* src = reinterpret_cast<const unsigned char*>(a.records) + static_cast<uint64_t>(id) * REC_BYTES;
    const uint32_t       dst = sbase + buf * SMEM_REC + slot * SLOT_BYTES;
#pragma unroll
    for(uint32_t k = 0; k < REC_BYTES; k += 16)
      cpAsync16(dst + k, src + k);
    if(SURF)
    {
      cpAsync16(sbase + SMEM_SURF + (buf * BATCH + slot) * 16u, a.surface + id);
      stsU32(sbase + SMEM_SID + (buf * BATCH + slot) * 4u, id);
    }
    if(GUT && GUTX)
    {
      // which instance the global id belongs to (the global index table, implicit in the offsets)
      uint32_t inst = 0;
#pragma unroll
      for(int k = 1; k < GUT_MAX_INSTANCES; k++)
        inst += (static_cast<uint32_t>(k) < a.gut.instanceCount && id >= a.gut.instanceOffset[k]) ? 1u : 0u;
      stsU32(sbase + SMEM_INST + (buf * BATCH + slot) * 4u, inst);
    }
  };
  // classify this thread's (landed) entry: per-warp-block hit bits -> per-warp hit masks
  auto classify = [&](bool have, uint32_t buf, uint32_t slot) {
    uint32_t bits = 0;
    if(have && GUT)
    {
      // 3DGUT quad: axis-aligned rectangle centre +- extent (EXTENT_CONIC); a block of pixel centres
      // [x0+0.5, x0+7.5] overlaps it iff |block centre - c| <= extent + 3.5 on both axes
      float4 r0 = ldsV4(sbase + buf * SMEM_REC + slot * SLOT_BYTES);  // cx cy ex ey
      float  w1x = 0.f, w1y = 0.f, w2x = 0.f, w2y = 0.f, lim1 = 3.0e38f, lim2 = 3.0e38f;
      if(GUTX && a.gut.extentEigen)
      {
        // words 2,3 hold w1 = b1 / |b1|^2, word 11 k = |w2| / |w1|, w2 = k (w1.y, -w1.x): bounding box of centre +- b1 +- b2,
        // and a separating-axis test along the quad's own axes: over a block of pixel centres (half extents 3.5)
        // dot(p - c, w_i) stays within its value at the block centre +- 3.5 (|w_i.x| + |w_i.y|)
        const float k   = __uint_as_float(ldsU32(sbase + buf * SMEM_REC + slot * SLOT_BYTES + 44));
        const float i1  = 1.0f / (r0.z * r0.z + r0.w * r0.w), i2 = i1 / k;  // 1/|w1|^2, and b2 = w2 / |w2|^2 = (w1.y, -w1.x) / (k |w1|^2)
        const float b1x = r0.z * i1, b1y = r0.w * i1, b2x = r0.w * i2, b2y = -r0.z * i2;
        w1x = r0.z, w1y = r0.w, w2x = r0.w * k, w2y = -r0.z * k;
        lim1 = 1.001f + 3.501f * (fabsf(w1x) + fabsf(w1y)), lim2 = 1.001f + 3.501f * (fabsf(w2x) + fabsf(w2y));
        r0.z = (fabsf(b1x) + fabsf(b2x)) * 1.0001f, r0.w = (fabsf(b1y) + fabsf(b2y)) * 1.0001f;
      }
#pragma unroll
      for(uint32_t b = 0; b < BLEND_WARPS; b++)
      {
        const float ddx = tileCx + static_cast<float>(8u * (b % BLOCKS_X)) - r0.x, ddy = tileCy + static_cast<float>(8u * (b / BLOCKS_X)) - r0.y;
        bool        hit = fabsf(ddx) <= r0.z + 3.501f && fabsf(ddy) <= r0.w + 3.501f;
        if(GUTX)
          hit = hit && fabsf(ddx * w1x + ddy * w1y) <= lim1 && fabsf(ddx * w2x + ddy * w2y) <= lim2;
        if(hit)
          bits |= 1u << b;
      }
      // Fast path of the quadratic kernel, per entry: M = diag(1/scale) R^T (rows M0..M2, so r = M dm), and the discard
      // threshold in the particle's squared distance d (response = exp(-d/2)):
      //   alpha = min(clamp, response * density) > 1/255   <=>  d < D2 = 2 ln(255 density)      (clamp > 1/255)
      //   response > kernelMinResponse                       <=>  d < D1 = -2 ln(kernelMinResponse)
      // so a fragment is kept iff d < cut = min(D1, D2), and a pixel whose d lies within `bd` of cut is re-evaluated
      // exactly (a d further than bd from the smaller threshold is decided for both). bd: the canonical origin is
      // hundreds of units long (distance / scale), so the cross product cancels and the fast and the exact evaluation
      // orders differ by up to about 2.6e-6 |ro| sqrt(d) in d for |ro| < 400 and less than 3e-3 beyond (fp32 emulation
      // of both orders over 3.3M random particle / ray pairs with d in 2..14, |ro| up to 4000); the band below is more
      // than twice that envelope everywhere.
      {
        const uint32_t rec     = sbase + buf * SMEM_REC + slot * SLOT_BYTES;
        const float4   q1      = ldsV4(rec + 16), q2 = ldsV4(rec + 32), q3 = ldsV4(rec + 48), q4 = ldsV4(rec + 64), q5 = ldsV4(rec + 80);
        const float    density = q1.w;
        const float    roLen   = (GUTX && a.gut.extentEigen) ? sqrtf(q2.x * q2.x + q2.y * q2.y + q2.z * q2.z) : q2.w;
        const bool     alive   = !(density <= a.gut.alphaCullThreshold) && a.gut.alphaClamp > 1.0f / 255.0f;
        const float    d2      = 1.3862943611198906f * __log2f(255.0f * density);
        const float    d1      = a.gut.kernelMinResponse > 0.0f ? -1.3862943611198906f * __log2f(a.gut.kernelMinResponse) : 3.0e38f;
        // Degenerate particles (a scale below 1e-18 or so: |r|^2 or |r x ro|^2 overflow in the un-normalised fast path, while
        // the oracle's normalised order stays finite) take the exact evaluation for every pixel: an infinite band.
        const float mMax = fmaxf(fmaxf(fmaxf(fabsf(q3.x * q3.w), fabsf(q3.x * q4.z)), fmaxf(fabsf(q3.x * q5.y), fabsf(q3.y * q4.x))),
                                 fmaxf(fmaxf(fabsf(q3.y * q4.w), fabsf(q3.y * q5.z)), fmaxf(fmaxf(fabsf(q3.z * q4.y), fabsf(q3.z * q5.x)), fabsf(q3.z * q5.w))));
        const float roMax = fmaxf(fmaxf(fabsf(q2.x), fabsf(q2.y)), fmaxf(fabsf(q2.z), 1.0f));
        const bool  tame  = mMax * roMax < 1e18f;  // (false for NaN)
        const float bd    = !alive ? -1.0f : (tame ? 2.2f * (2e-3f + 4e-7f * roLen) + 1e-5f : __uint_as_float(0x7f800000u));
        const float    cut     = alive ? fminf(d1, d2) : -3.0e38f;
        asm volatile("st.shared.v4.f32 [%0], {%1,%2,%3,%4};" ::"r"(rec + REC_BYTES), "f"(q3.x * q3.w), "f"(q3.x * q4.z), "f"(q3.x * q5.y), "f"(cut) : "memory");
        asm volatile("st.shared.v4.f32 [%0], {%1,%2,%3,%4};" ::"r"(rec + REC_BYTES + 16), "f"(q3.y * q4.x), "f"(q3.y * q4.w), "f"(q3.y * q5.z), "f"(bd) : "memory");
        asm volatile("st.shared.v4.f32 [%0], {%1,%2,%3,%4};" ::"r"(rec + REC_BYTES + 32), "f"(q3.z * q4.y), "f"(q3.z * q5.x), "f"(q3.z * q5.w), "f"(0.0f) : "memory");
      }
    }
    else if(have)
    {
      const uint32_t src = sbase + buf * SMEM_REC + slot * SLOT_BYTES;
      const float4   r0 = ldsV4(src), r1 = ldsV4(src + 16), r2 = ldsV4(src + 32);
      const uint32_t bb0 = __float_as_uint(r2.z), bb1 = __float_as_uint(r2.w);
      const uint32_t x0 = bb0 & 0xffffu, y0 = bb0 >> 16, x1 = bb1 & 0xffffu, y1 = bb1 >> 16;
      // Warp blocks (bit b = block (b % BLOCKS_X, b / BLOCKS_X)) the splat can touch: the pixel bbox must overlap the block,
      // and a separating-axis test along the splat's own axes must not exclude it. Over an 8x8 block of pixel centres
      // (half extents 3.5) the fragPos component f_i = dot(p - c, w_i) stays within f_i(centre) +- e_i, and |f_i| > L
      // everywhere means A > L^2 everywhere. A fragment survives only if A <= 8 and exp(-A/2) * alpha > 1/255, i.e.
      // A < 2 ln(255 alpha): L^2 = min(8, 2 ln(255 alpha)), with margins for the approximate log / sqrt and the rounding
      // of f_i. f_i(centre of block (bx, by)) = gx_i[bx] + gy_i[by]: one product per block column / row and axis.
      float lim = 2.8292f;
      if(!NOGAUSS)
      {
        const float amax = 1.3862943611f * __log2f(255.0f * r2.y) * 1.0001f + 1e-3f;
        lim              = amax > 0.0f ? __fsqrt_rn(fminf(amax, 8.0f)) * 1.0002f + 2e-4f : -1.0f;
      }
      const float l1 = lim + 3.5f * (fabsf(r0.z) + fabsf(r0.w)), l2 = lim + 3.5f * (fabsf(r1.x) + fabsf(r1.y));
      float       gx1[BLOCKS_X], gx2[BLOCKS_X];
      bool        colHit[BLOCKS_X];
#pragma unroll
      for(uint32_t bx = 0; bx < BLOCKS_X; bx++)
      {
        const float    ddx = tileCx + static_cast<float>(8u * bx) - r0.x;
        const uint32_t X   = tileX0 + 8u * bx;
        gx1[bx] = ddx * r0.z, gx2[bx] = ddx * r1.x;
        colHit[bx] = x0 <= X + 7u && x1 >= X;
      }
#pragma unroll
      for(uint32_t by = 0; by < BLOCKS_Y; by++)
      {
        const float    ddy = tileCy + static_cast<float>(8u * by) - r0.y;
        const uint32_t Y   = tileY0 + 8u * by;
        const float    gy1 = ddy * r0.w, gy2 = ddy * r1.y;
        const bool     rowHit = y0 <= Y + 7u && y1 >= Y;
#pragma unroll
        for(uint32_t bx = 0; bx < BLOCKS_X; bx++)
          if(rowHit && colHit[bx] && fabsf(gx1[bx] + gy1) <= l1 && fabsf(gx2[bx] + gy2) <= l2)
            bits |= 1u << (by * BLOCKS_X + bx);
      }
      // the bbox words of the staged record have served their purpose: they become the fragment stage's discard
      // thresholds in A (see evalFrag). A* through the SFU log (error < 3e-6 in A); the oracle's own decision
      // exp(-A/2) * alpha > 1/255 flips within 1e-6 of A*, so anything farther than the 3e-5 band from A* is decided
      // the same way by both, and the band itself goes to the exact path.
      const float aStar = NOGAUSS ? 3.0e38f : 1.3862943611198906f * __log2f(255.0f * r2.y);
      stsU32(src + 40, __float_as_uint(fminf(8.0f, aStar - 3e-5f)));
      stsU32(src + 44, __float_as_uint(aStar));
    }
#pragma unroll
    for(uint32_t b = 0; b < BLEND_WARPS; b++)
    {
      const unsigned mb = __ballot_sync(FULL_MASK, (bits >> b) & 1u);
      if(lane == b)  // hit[buf][blend warp b][word = slot / 32]
        stsU32(sbase + SMEM_HIT + ((buf * BLEND_WARPS + b) * (BATCH / 32) + (slot >> 5)) * 4u, mb);
    }
  };

  // One list entry against this thread's two pixels, up to (not including) the ordered blend.
  // Returns MINUS the fragment opacities (0 = discarded); `gmin` = distance of the nearer of the two
  // SFU opacities to the discard threshold (the caller sends near misses to the exact path).
  struct Frag
  {
    f32x2  A2, n2;  // A of the two pixels; minus opacity (masked)
    float  gmin;    // distance of the nearer of the two A values to A* (near misses go to the exact path)
    float  astar;
    float  r, g, b, alpha;
    uint32_t addr;  // shared address of the staged record (SURF: locates the entry's normal / id)
  };
  auto evalFrag = [&](uint32_t addr) {
    Frag         f;
    f.addr = addr;
    if(GUT)
    {
      // threedgut_raster.frag.slang:87-191 + particleProcessHitGut (threedgrt.h.slang:238-278), in the
      // operation order of orc_gut_fragment; exp through the oracle's fixed sequence: decisions are exact
      const float4 q0 = ldsV4(addr);       // cx cy ex ey
      const float4 q1 = ldsV4(addr + 16);  // r g b a
      const float4 q2 = ldsV4(addr + 32);  // canonical ray origin
      f.r = q1.x, f.g = q1.y, f.b = q1.z, f.alpha = q1.w;
      f.gmin = 1.0f;
      f.A2   = pk(0.f, 0.f);
      // ray directions in the model space of the entry's instance (threedgut_raster.frag.slang:117-121)
      float dmA[3] = {gutDirA[0], gutDirA[1], gutDirA[2]}, dmB[3] = {gutDirB[0], gutDirB[1], gutDirB[2]};
      if(GUTX && a.gut.instanceCount > 1u)
      {
        const uint32_t slot = (addr - (sbase + surfBuf * SMEM_REC)) / SLOT_BYTES;
        const float*   mi   = a.gut.instanceInverse[ldsU32(sbase + SMEM_INST + (surfBuf * BATCH + slot) * 4u)];
#pragma unroll
        for(int p = 0; p < 2; p++)
        {
          const uint32_t wAddr = sbase + SMEM_WORLD + tid * 32u + 12u * p;
          const float    rd[3] = {__uint_as_float(ldsU32(wAddr)), __uint_as_float(ldsU32(wAddr + 4u)), __uint_as_float(ldsU32(wAddr + 8u))};
          float          dm[3];
#pragma unroll
          for(int j = 0; j < 3; j++)
            dm[j] = __fadd_rn(__fadd_rn(__fmul_rn(rd[0], mi[0 + j]), __fmul_rn(rd[1], mi[3 + j])), __fmul_rn(rd[2], mi[6 + j]));
          const float dmn = __fdiv_rn(1.0f, __fsqrt_rn(__fadd_rn(__fadd_rn(__fmul_rn(dm[0], dm[0]), __fmul_rn(dm[1], dm[1])), __fmul_rn(dm[2], dm[2]))));
#pragma unroll
          for(int j = 0; j < 3; j++)
            (p ? dmB : dmA)[j] = __fmul_rn(dm[j], dmn);
        }
      }
      float nOp[2];
      if(a.gut.kernelDegree != 2u)
      {
        nOp[0] = gutExactPixel<NOGAUSS>(a.gut, addr, -nfx, gutPyA, dmA[0], dmA[1], dmA[2]);
        nOp[1] = gutExactPixel<NOGAUSS>(a.gut, addr, -nfx, gutPyB, dmB[0], dmB[1], dmB[2]);
      }
      else
      {
        // Fast path of the default (quadratic = Gaussian) kernel, both pixels with packed fp32: d = |r x ro|^2 / |r|^2
        // with r = M dm not normalised, reciprocal and exp on the SFU; accept / reject is one comparison of d with the
        // entry's threshold, and a pixel inside the guard band around it is re-evaluated exactly (classify), so the
        // decisions never differ from the oracle's.
        const float4 g0 = ldsV4(addr + REC_BYTES);       // M00 M01 M02 cut
        const float4 g1 = ldsV4(addr + REC_BYTES + 16);  // M10 M11 M12 bd
        const float4 g2 = ldsV4(addr + REC_BYTES + 32);  // M20 M21 M22
        const f32x2 m0 = pk(dmA[0], dmB[0]), m1 = pk(dmA[1], dmB[1]), m2 = pk(dmA[2], dmB[2]);
        const f32x2 r0 = fma2(m2, pk(g0.z, g0.z), fma2(m1, pk(g0.y, g0.y), mul2(m0, pk(g0.x, g0.x))));
        const f32x2 r1 = fma2(m2, pk(g1.z, g1.z), fma2(m1, pk(g1.y, g1.y), mul2(m0, pk(g1.x, g1.x))));
        const f32x2 r2 = fma2(m2, pk(g2.z, g2.z), fma2(m1, pk(g2.y, g2.y), mul2(m0, pk(g2.x, g2.x))));
        const float nox = -q2.x, noy = -q2.y, noz = -q2.z;
        const f32x2 cx = fma2(r1, pk(q2.z, q2.z), mul2(r2, pk(noy, noy)));
        const f32x2 cy = fma2(r2, pk(q2.x, q2.x), mul2(r0, pk(noz, noz)));
        const f32x2 cz = fma2(r0, pk(q2.y, q2.y), mul2(r1, pk(nox, nox)));
        const f32x2 num = fma2(cz, cz, fma2(cy, cy, mul2(cx, cx)));
        const f32x2 den = fma2(r2, r2, fma2(r1, r1, mul2(r0, r0)));
        float       nlo, nhi, dlo, dhi;
        upk(num, nlo, nhi);
        upk(den, dlo, dhi);
        const f32x2 dist2 = mul2(num, pk(rcpApprox(dlo), rcpApprox(dhi)));
        const f32x2 earg2 = mul2(dist2, pk(-0.72134752044448170368f, -0.72134752044448170368f));
        float       distA, distB, eA, eB, alA, alB;
        upk(dist2, distA, distB);
        upk(earg2, eA, eB);
        upk(mul2(pk(ex2Approx(eA), ex2Approx(eB)), pk(q1.w, q1.w)), alA, alB);
        alA = fminf(a.gut.alphaClamp, alA), alB = fminf(a.gut.alphaClamp, alB);
        // outside the quad: d = +big (rows) or the threshold = -big (the column both pixels share), beyond every band
        float tA, tB;
        if(GUTX && a.gut.extentEigen)
        {
          const bool inA = gutInsideQuad(true, q0, q2.w, -nfx, gutPyA), inB = gutInsideQuad(true, q0, q2.w, -nfx, gutPyB);
          tA = (inA ? distA : 3.0e38f) - g0.w, tB = (inB ? distB : 3.0e38f) - g0.w;
        }
        else
        {
          const float cutX = fabsf(__fsub_rn(-nfx, q0.x)) <= q0.z ? g0.w : -3.0e38f;
          tA = (fabsf(__fsub_rn(gutPyA, q0.y)) <= q0.w ? distA : 3.0e38f) - cutX;
          tB = (fabsf(__fsub_rn(gutPyB, q0.y)) <= q0.w ? distB : 3.0e38f) - cutX;
        }
        nOp[0] = tA < 0.0f ? (NOGAUSS ? -1.0f : -alA) : 0.0f;
        nOp[1] = tB < 0.0f ? (NOGAUSS ? -1.0f : -alB) : 0.0f;
        const bool nearA = !(fabsf(tA) > g1.w), nearB = !(fabsf(tB) > g1.w);  // (a NaN distance is re-evaluated too)
        if(nearA || nearB)
        {
          if(nearA)
            nOp[0] = gutExactPixel<NOGAUSS>(a.gut, addr, -nfx, gutPyA, dmA[0], dmA[1], dmA[2]);
          if(nearB)
            nOp[1] = gutExactPixel<NOGAUSS>(a.gut, addr, -nfx, gutPyB, dmB[0], dmB[1], dmB[2]);
        }
      }
      f.n2 = pk(nOp[0], nOp[1]);
      return f;
    }
    const float4 ra  = ldsV4(addr);       // cx cy w1x w1y
    const float4 rb  = ldsV4(addr + 16);  // w2x w2y r g
    const float4 rc  = ldsV4(addr + 32);  // b a | cut, A* (written over the bbox words by the staging thread)
    const float  ndx = __fadd_rn(ra.x, nfx);            // -(px - cx): A is even in (dx,dy)
    const f32x2  ndy2 = add2(pk(ra.y, ra.y), nfy2);
    const float  t = __fmul_rn(ndx, ra.z), u = __fmul_rn(ndx, rb.x);
    const f32x2  fpx2 = fma2(ndy2, pk(ra.w, ra.w), pk(t, t));
    const f32x2  fpy2 = fma2(ndy2, pk(rb.y, rb.y), pk(u, u));
    f.A2              = fma2(fpy2, fpy2, mul2(fpx2, fpx2));
    f.r = rb.z, f.g = rb.w, f.b = rc.x, f.alpha = rc.y;
    float Alo, Ahi, dlo, dhi;
    upk(f.A2, Alo, Ahi);
    const bool vA = Alo <= rc.z, vB = Ahi <= rc.z;
    upk(add2(f.A2, pk(-rc.w, -rc.w)), dlo, dhi);
    f.gmin  = fminf(fabsf(dlo), fabsf(dhi));
    f.astar = rc.w;
    if(NOGAUSS)
      f.n2 = pk(vA ? -1.0f : 0.0f, vB ? -1.0f : 0.0f);
    else
    {
      float xlo, xhi, nlo, nhi;
      upk(mul2(f.A2, kExp), xlo, xhi);
      const float na = -rc.y;
      const f32x2 n2 = mul2(pk(ex2Approx(xlo), ex2Approx(xhi)), pk(na, na));
      upk(n2, nlo, nhi);
      f.n2 = pk(vA ? nlo : 0.0f, vB ? nhi : 0.0f);
    }
    return f;
  };
  // within the guard band of the 1/255 discard threshold (rare): decide exactly, like the oracle
  auto fixFrag = [&](Frag& f) {
    if(GUT)
      return;  // (the 3DGUT evaluation arbitrates its own near misses)
    float Alo, Ahi, mlo, mhi;
    upk(f.A2, Alo, Ahi);
    upk(f.n2, mlo, mhi);
    if(fabsf(Alo - f.astar) <= BAND)
      mlo = !(Alo > 8.0f) ? (NOGAUSS ? -1.0f : exactNegOpacity(Alo, -f.alpha)) : 0.0f;
    if(fabsf(Ahi - f.astar) <= BAND)
      mhi = !(Ahi > 8.0f) ? (NOGAUSS ? -1.0f : exactNegOpacity(Ahi, -f.alpha)) : 0.0f;
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
    // "under" update in traversal (front-to-back) order: C += c * opacity * T, T -= opacity * T
    const f32x2 nw2 = mul2(f.n2, acc);  // -(opacity * T)
    fma2acc(c0, nw2, pk(f.r, f.r));
    fma2acc(c1, nw2, pk(f.g, f.g));
    fma2acc(c2, nw2, pk(f.b, f.b));
    add2acc(acc, nw2);
    if(!FTB)
      add2acc(asum, f.n2);  // back-to-front blend state: alpha = sum of the fragment opacities (src/gaussian_splatting.cpp:2078-2087)
    if(SURF)
    {
      // threedgs_raster.frag.slang:316-350: normal * opacity under the same operator; first depth at
      // which the transmittance drops below the iso threshold; id of the last fragment that was kept
      const uint32_t slot = (f.addr - (sbase + surfBuf * SMEM_REC)) / SLOT_BYTES;
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
  };

  if(range.x < range.y)
  {
    // prologue: batch 0 lands and is classified; the list index of batch 1 is already on its way
    const bool stager = tid < BATCH;  // (warp-uniform)
    uint32_t   idNext = 0;
    // list position of traversal position p (p = range.x + k): front-to-back frames read the list forwards,
    // back-to-front frames from its end (nearest fragment first)
    auto listAt = [&](uint32_t p) -> uint32_t { return FTB ? p : range.y - 1u - (p - range.x); };
    if(stager)
    {
      if(range.x + tid < range.y)
        gather(a.tileVals[listAt(range.x + tid)], 0, tid);
      cpAsyncCommit();
      idNext = (range.x + BATCH + tid < range.y) ? a.tileVals[listAt(range.x + BATCH + tid)] : 0u;
      cpAsyncWaitAll();
      classify(range.x + tid < range.y, 0, tid);
    }
    __syncthreads();
    for(uint32_t base = range.x, buf = 0;; base += BATCH, buf ^= 1u)
    {
      const bool more = base + BATCH < range.y;
      if(more)
      {
        // records of the next batch fly into the other buffer while this one is blended
        if(stager)
        {
          if(base + BATCH + tid < range.y)
            gather(idNext, buf ^ 1u, tid);
          cpAsyncCommit();
          if(base + 2 * BATCH + tid < range.y)
            idNext = a.tileVals[listAt(base + 2 * BATCH + tid)];
        }
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
          const uint32_t chunkAddr = recBase + chunk * 32u * SLOT_BYTES;
          while(m)
          {
            const uint32_t addr0 = chunkAddr + (__ffs(m) - 1) * SLOT_BYTES;
            m &= m - 1;
#ifdef VKGS_BLEND_SINGLE
            if(false)
#else
            if(m)
#endif
            {
              const uint32_t addr1 = chunkAddr + (__ffs(m) - 1) * SLOT_BYTES;
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
      if(more && stager)
      {
        cpAsyncWaitAll();
        classify(base + BATCH + tid < range.y, buf ^ 1u, tid);
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
  float as[2];
  upk(asum, as[0], as[1]);
#pragma unroll
  for(int p = 0; p < 2; p++)
  {
    if(!(p ? insideB : insideA))
      continue;
    const uint64_t o  = static_cast<uint64_t>(p ? pyB : pyA) * a.width + px;
    // (0 - x, not -x: an untouched pixel must come out as +0)
    const float    c0f = __fsub_rn(0.0f, ca[p][0]), c1f = __fsub_rn(0.0f, ca[p][1]), c2f = __fsub_rn(0.0f, ca[p][2]);
    const float    al = FTB ? 1.0f - ca[p][3] : __fsub_rn(0.0f, as[p]);
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

template <bool FTB, bool NOGAUSS, bool COUNT, bool SURF, bool GUT, bool GUTX = false>
static void allowSmem()
{
  cudaFuncSetAttribute(k_blend<FTB, NOGAUSS, COUNT, SURF, GUT, GUTX>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                       static_cast<int>(blendSmemBytes(SURF, GUT)));
  if(GUT)  // four resident CTAs of 144-byte slots: ask for the carveout instead of leaving it to the launch heuristic
    cudaFuncSetAttribute(k_blend<FTB, NOGAUSS, COUNT, SURF, GUT, GUTX>, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared);
}

void initBlendKernels()
{
  // dynamic shared memory above the 48 KB default (wide tiles, 96-byte 3DGUT records)
  allowSmem<true, true, false, false, false>(), allowSmem<true, false, false, false, false>();
  allowSmem<false, true, false, false, false>(), allowSmem<false, false, false, false, false>();
  allowSmem<true, true, true, false, false>(), allowSmem<true, false, true, false, false>();
  allowSmem<false, true, true, false, false>(), allowSmem<false, false, true, false, false>();
  allowSmem<true, true, false, true, false>(), allowSmem<true, false, false, true, false>();
  allowSmem<true, true, false, false, true>(), allowSmem<true, false, false, false, true>();
  allowSmem<false, true, false, false, true>(), allowSmem<false, false, false, false, true>();
  allowSmem<true, true, false, false, true, true>(), allowSmem<true, false, false, false, true, true>();
  allowSmem<false, true, false, false, true, true>(), allowSmem<false, false, false, false, true, true>();
}

void launchBlend(const BlendArgs& args, cudaStream_t stream)
{
  const uint32_t tiles = (args.tileCount ? args.tileCount : args.tilesX * args.tilesY) * BANDS;  // CTAs
  const bool count = args.fragmentCounters != nullptr;
  if(args.outNormals)
  {
    // surface-info variant: front to back only (the context rejects other combinations)
    constexpr uint32_t SMEM = blendSmemBytes(true, false);
    if(args.disableOpacityGaussian)
      k_blend<true, true, false, true, false><<<tiles, BLEND_THREADS, SMEM, stream>>>(args);
    else
      k_blend<true, false, false, true, false><<<tiles, BLEND_THREADS, SMEM, stream>>>(args);
    return;
  }
  if(args.gut.enabled)
  {
    // VK3DGUT fragment stage (no fragment counters / surface info in this variant)
    constexpr uint32_t SMEM = blendSmemBytes(false, true);
    // GUTX: the general instantiation (multi-instance scenes, EXTENT_EIGEN quads, fisheye rays); the common case
    // (one instance, EXTENT_CONIC, pinhole) runs an instantiation with those branches compiled out
    const bool general = args.gut.instanceCount > 1u || args.gut.extentEigen || args.gut.fisheye;
#define VKGS_GUT_LAUNCH(F, G)                                                                                                    \
  do                                                                                                                             \
  {                                                                                                                              \
    if(general)                                                                                                                  \
      k_blend<F, G, false, false, true, true><<<tiles, BLEND_THREADS, SMEM, stream>>>(args);                                      \
    else                                                                                                                         \
      k_blend<F, G, false, false, true, false><<<tiles, BLEND_THREADS, SMEM, stream>>>(args);                                     \
  } while(0)
    if(args.frontToBack)
    {
      if(args.disableOpacityGaussian)
        VKGS_GUT_LAUNCH(true, true);
      else
        VKGS_GUT_LAUNCH(true, false);
    }
    else
    {
      if(args.disableOpacityGaussian)
        VKGS_GUT_LAUNCH(false, true);
      else
        VKGS_GUT_LAUNCH(false, false);
    }
#undef VKGS_GUT_LAUNCH
    return;
  }
  constexpr uint32_t SMEM = blendSmemBytes(false, false);
#define VKGS_BLEND_LAUNCH(F, G)                                                                                                  \
  do                                                                                                                             \
  {                                                                                                                              \
    if(count)                                                                                                                    \
      k_blend<F, G, true, false, false><<<tiles, BLEND_THREADS, SMEM, stream>>>(args);                                                            \
    else                                                                                                                         \
      k_blend<F, G, false, false, false><<<tiles, BLEND_THREADS, SMEM, stream>>>(args);                                                           \
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
