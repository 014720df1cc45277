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
// into shared memory, then every warp ballots which of the batch's splats overlap its 8x4 block
// (bounding-box test) and evaluates only those, in list order — per-pixel blend order is exactly
// the sorted order, like the ROP.
//
// Exactness: fragPos / A are evaluated with explicit fp32 mul/fma in the oracle's operation order,
// so the `A > 8` discard is bit-exact. opacity uses the SFU ex2 for speed; whenever that value is
// within a guard band of the 1/255 discard threshold it is recomputed with the same fixed-sequence
// expf the oracle uses, so the discard decision is exact too and values differ by a few ulp only.
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

__device__ __forceinline__ float ex2Approx(float x)
{
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}

template <bool FTB>
__global__ void __launch_bounds__(BLEND_THREADS) k_blend(const __grid_constant__ BlendArgs a)
{
  __shared__ float4 s_a[BLEND_THREADS];  // cx, cy, w1x, w1y
  __shared__ float4 s_b[BLEND_THREADS];  // w2x, w2y, r, g
  __shared__ float4 s_c[BLEND_THREADS];  // b, a, bbox0, bbox1

  const unsigned tid = threadIdx.x, lane = tid & 31u, warp = tid >> 5;
  const uint32_t tile = blockIdx.x;
  const uint32_t tx = tile % a.tilesX, ty = tile / a.tilesX;
  const uint32_t wx0 = tx * TILE_W + (warp & 1u) * 8u, wy0 = ty * TILE_H + (warp >> 1) * 4u;
  const uint32_t px = wx0 + (lane & 7u), py = wy0 + (lane >> 3);
  const bool     inside = px < a.width && py < a.height;
  const float    fx = static_cast<float>(px) + 0.5f, fy = static_cast<float>(py) + 0.5f;

  const uint2 range = a.ranges[tile];
  float       c0 = 0.f, c1 = 0.f, c2 = 0.f, alpha = 0.f;
  bool        done = !inside;
  const float THRESHOLD = 1.0f / 255.0f;
  const float GUARD     = THRESHOLD * 4e-6f;

  for(uint32_t base = range.x; base < range.y; base += BLEND_THREADS)
  {
    const uint32_t n = min(static_cast<uint32_t>(BLEND_THREADS), range.y - base);
    // all pixels of the tile saturated (front-to-back only) -> stop reading the list
    const int active = __syncthreads_count(!done);
    if(FTB && active == 0)
      break;
    if(tid < n)
    {
      const uint32_t id  = a.tileVals[base + tid];
      const float4*  rec = reinterpret_cast<const float4*>(a.records + static_cast<uint64_t>(id) * RECORD_WORDS);
      s_a[tid]           = __ldg(rec + 0);
      s_b[tid]           = __ldg(rec + 1);
      s_c[tid]           = __ldg(rec + 2);
    }
    __syncthreads();
    if(__all_sync(FULL_MASK, done))
      continue;

    for(uint32_t chunk = 0; chunk < n; chunk += 32)
    {
      const uint32_t j   = chunk + lane;
      bool           hit = false;
      if(j < n)
      {
        const uint32_t bb0 = __float_as_uint(s_c[j].z), bb1 = __float_as_uint(s_c[j].w);
        hit = (bb0 & 0xffffu) <= wx0 + 7u && (bb1 & 0xffffu) >= wx0 && (bb0 >> 16) <= wy0 + 3u && (bb1 >> 16) >= wy0;
      }
      unsigned m = __ballot_sync(FULL_MASK, hit);
      while(m)
      {
        const uint32_t jj = chunk + __ffs(m) - 1;
        m &= m - 1;
        if(done)
          continue;
        const float4 ra = s_a[jj];
        const float  dx = __fsub_rn(fx, ra.x), dy = __fsub_rn(fy, ra.y);
        const float4 rb = s_b[jj];
        const float  fpx = __fmaf_rn(dy, ra.w, __fmul_rn(dx, ra.z));
        const float  fpy = __fmaf_rn(dy, rb.y, __fmul_rn(dx, rb.x));
        const float  A   = __fmaf_rn(fpy, fpy, __fmul_rn(fpx, fpx));
        if(A > 8.0f)
          continue;
        const float4 rc = s_c[jj];
        float        op;
        if(a.disableOpacityGaussian)
          op = 1.0f;
        else
        {
          op = ex2Approx(A * -0.72134752044448170368f) * rc.y;  // exp(-A/2) = 2^(-A/2 * log2 e)
          if(fabsf(op - THRESHOLD) <= GUARD)
            op = __fmul_rn(expfExact(__fmul_rn(-0.5f, A)), rc.y);
        }
        if(op <= THRESHOLD)
          continue;
        if(FTB)
        {
          const float t = 1.0f - alpha;
          const float w = op * t;
          c0            = fmaf(rb.z, w, c0);
          c1            = fmaf(rb.w, w, c1);
          c2            = fmaf(rc.x, w, c2);
          alpha += w;
          if(1.0f - alpha < a.transmittanceEpsilon)
            done = true;
        }
        else
        {
          const float t = 1.0f - op;
          c0            = fmaf(rb.z, op, c0 * t);
          c1            = fmaf(rb.w, op, c1 * t);
          c2            = fmaf(rc.x, op, c2 * t);
          alpha += op;
        }
      }
      if(FTB && __all_sync(FULL_MASK, done))
        break;
    }
  }
  if(inside)
    a.image[static_cast<uint64_t>(py) * a.width + px] = make_float4(c0, c1, c2, alpha);
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
