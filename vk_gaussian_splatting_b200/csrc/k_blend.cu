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
__device__ __noinline__ float expfExact(float x)
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

constexpr uint32_t REC_BYTES = RECORD_WORDS * 4;  // 48: cx cy w1x w1y | w2x w2y r g | b a mask -

template <bool FTB>
__global__ void __launch_bounds__(BLEND_THREADS) k_blend(const __grid_constant__ BlendArgs a)
{
  __shared__ __align__(16) float s_rec[BLEND_THREADS * RECORD_WORDS];
  const uint32_t sbase = smem_u32(s_rec);

  const unsigned tid = threadIdx.x, lane = tid & 31u, warp = tid >> 5;
  const uint32_t tile = blockIdx.x;
  const uint32_t tx = tile % a.tilesX, ty = tile / a.tilesX;
  const uint32_t tileX0 = tx * TILE_W, tileY0 = ty * TILE_H;
  const uint32_t px = tileX0 + (warp & 1u) * 8u + (lane & 7u), py = tileY0 + (warp >> 1) * 4u + (lane >> 3);
  const bool     inside = px < a.width && py < a.height;
  const float    fx = static_cast<float>(px) + 0.5f, fy = static_cast<float>(py) + 0.5f;
  // centre of this warp's 8x4 block of pixel centres (half extents 3.5 x 1.5)
  const float    blockCx = static_cast<float>(tileX0 + (warp & 1u) * 8u) + 4.0f, blockCy = static_cast<float>(tileY0 + (warp >> 1) * 4u) + 2.0f;

  const uint2 range = make_uint2(a.rangeBegin[tile], a.rangeEnd[tile]);  // empty tile: begin > end
  float       c0 = 0.f, c1 = 0.f, c2 = 0.f;
  float       acc = FTB ? 1.0f : 0.0f;  // FTB: transmittance T = 1 - A_dst;  BTF: sum of alphas
  bool        done = !inside;
  const float THRESHOLD = 1.0f / 255.0f;
  const float THR_HI    = THRESHOLD * (1.0f + 4e-6f);
  const float THR_LO    = THRESHOLD * (1.0f - 4e-6f);
  const float eps       = a.transmittanceEpsilon;

  // gather of this thread's entry of a batch: record + warp-block mask
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
  fetch(range.x);

  for(uint32_t base = range.x; base < range.y; base += BLEND_THREADS)
  {
    const uint32_t n = min(static_cast<uint32_t>(BLEND_THREADS), range.y - base);
    // all pixels of the tile saturated (front-to-back only) -> stop reading the list.
    // (the barrier also protects the shared batch against the previous round's readers)
    const int active = __syncthreads_count(!done);
    if(FTB && active == 0)
      break;
    if(tid < n)
    {
      // which of the 8 warp blocks (2 columns x 4 rows of 8x4 pixels) does the pixel bbox touch?
      const uint32_t bb0 = __float_as_uint(r2.z), bb1 = __float_as_uint(r2.w);
      const uint32_t x0 = bb0 & 0xffffu, y0 = bb0 >> 16, x1 = bb1 & 0xffffu, y1 = bb1 >> 16;
      const uint32_t colL = (x0 <= tileX0 + 7u && x1 >= tileX0) ? 0x55u : 0u;       // warps 0,2,4,6
      const uint32_t colR = (x0 <= tileX0 + 15u && x1 >= tileX0 + 8u) ? 0xaau : 0u;  // warps 1,3,5,7
      uint32_t       rows = 0;
#pragma unroll
      for(uint32_t r = 0; r < 4; r++)
        rows |= (y0 <= tileY0 + 4u * r + 3u && y1 >= tileY0 + 4u * r) ? (3u << (2u * r)) : 0u;
      r2.z = __uint_as_float((colL | colR) & rows);
      const uint32_t dst = sbase + tid * REC_BYTES;
      stsV4(dst, r0);
      stsV4(dst + 16, r1);
      stsV4(dst + 32, r2);
    }
    __syncthreads();
    fetch(base + BLEND_THREADS);  // prefetch the next batch while this one is blended
    if(__all_sync(FULL_MASK, done))
      continue;

    for(uint32_t chunk = 0; chunk < n; chunk += 32)
    {
      // Lane l tests splat chunk+l against THIS warp's 8x4 pixel block: first the precomputed bbox
      // mask, then a separating-axis test along the splat's own axes — over the block the
      // interpolated fragPos component f_i = dot(p - c, w_i) stays within f_i(centre) +- e_i, and
      // |f_i| > sqrt(8) everywhere means A > 8 everywhere. Conservative, and amortised 32x.
      const uint32_t j   = chunk + lane;
      bool           hit = j < n && ((ldsU32(sbase + j * REC_BYTES + 40) >> warp) & 1u);
      if(hit)
      {
        const float4 qa = ldsV4(sbase + j * REC_BYTES);
        const float2 qb = ldsV2(sbase + j * REC_BYTES + 16);
        const float  ddx = blockCx - qa.x, ddy = blockCy - qa.y;
        const float  f1 = fabsf(ddx * qa.z + ddy * qa.w) - (3.5f * fabsf(qa.z) + 1.5f * fabsf(qa.w));
        const float  f2 = fabsf(ddx * qb.x + ddy * qb.y) - (3.5f * fabsf(qb.x) + 1.5f * fabsf(qb.y));
        hit = fmaxf(f1, f2) <= 2.829f;  // sqrt(8) = 2.82843 plus a safety margin for rounding
      }
      unsigned       m   = __ballot_sync(FULL_MASK, hit);
      const uint32_t chunkAddr = sbase + chunk * REC_BYTES;
      while(m)
      {
        const uint32_t addr = chunkAddr + (__ffs(m) - 1) * REC_BYTES;
        m &= m - 1;
        const float4 ra = ldsV4(addr);
        const float4 rb = ldsV4(addr + 16);
        const float  dx = __fsub_rn(fx, ra.x), dy = __fsub_rn(fy, ra.y);
        const float  fpx = __fmaf_rn(dy, ra.w, __fmul_rn(dx, ra.z));
        const float  fpy = __fmaf_rn(dy, rb.y, __fmul_rn(dx, rb.x));
        const float  A   = __fmaf_rn(fpy, fpy, __fmul_rn(fpx, fpx));
        if(A > 8.0f || done)
          continue;
        const float2 rc = ldsV2(addr + 32);
        float        op;
        if(a.disableOpacityGaussian)
          op = 1.0f;
        else
        {
          op = ex2Approx(A * -0.72134752044448170368f) * rc.y;  // exp(-A/2) = 2^(-A/2 * log2 e)
          if(op <= THR_HI)
          {
            if(op < THR_LO)
              continue;
            op = __fmul_rn(expfExact(__fmul_rn(-0.5f, A)), rc.y);  // within the guard band: decide exactly
            if(op <= THRESHOLD)
              continue;
          }
        }
        if(FTB)
        {
          const float w = op * acc;
          c0            = fmaf(rb.z, w, c0);
          c1            = fmaf(rb.w, w, c1);
          c2            = fmaf(rc.x, w, c2);
          acc -= w;
          done = acc < eps;
        }
        else
        {
          const float t = 1.0f - op;
          c0            = fmaf(rb.z, op, c0 * t);
          c1            = fmaf(rb.w, op, c1 * t);
          c2            = fmaf(rc.x, op, c2 * t);
          acc += op;
        }
      }
      if(FTB && __all_sync(FULL_MASK, done))
        break;
    }
  }
  if(inside)
  {
    // the colour target is rounded ONCE from the fp32 accumulators (the reference's ROP rounds after
    // every blend in the target format)
    const uint64_t o  = static_cast<uint64_t>(py) * a.width + px;
    const float    al = FTB ? 1.0f - acc : acc;
    if(a.targetFormat == VKGS_FORMAT_FLOAT32)
      static_cast<float4*>(a.image)[o] = make_float4(c0, c1, c2, al);
    else if(a.targetFormat == VKGS_FORMAT_FLOAT16)
    {
      const __half2 lo = __floats2half2_rn(c0, c1), hi = __floats2half2_rn(c2, al);
      static_cast<uint2*>(a.image)[o] = make_uint2(*reinterpret_cast<const uint32_t*>(&lo), *reinterpret_cast<const uint32_t*>(&hi));
    }
    else
    {
      // R8G8B8A8_UNORM: clamp to [0,1], scale, round to nearest
      const uint32_t r = __float2uint_rn(__saturatef(c0) * 255.0f), gch = __float2uint_rn(__saturatef(c1) * 255.0f),
                     b = __float2uint_rn(__saturatef(c2) * 255.0f), aa = __float2uint_rn(__saturatef(al) * 255.0f);
      static_cast<uint32_t*>(a.image)[o] = r | (gch << 8) | (b << 16) | (aa << 24);
    }
  }
}

}  // namespace

void launchBlend(const BlendArgs& args, cudaStream_t stream)
{
  const uint32_t tiles = args.tilesX * args.tilesY;
  if(args.frontToBack)
    k_blend<true><<<tiles, BLEND_THREADS, 0, stream>>>(args);
  else
    k_blend<false><<<tiles, BLEND_THREADS, 0, stream>>>(args);
}

}  // namespace vkgs
