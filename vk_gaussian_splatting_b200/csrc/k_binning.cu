// k_binning.cu — per-tile splat lists for the software rasterizer.
//
// The reference hands one screen-aligned quad per splat to the hardware rasterizer in sorted order
// (drawSplatPrimitives, src/gaussian_splatting.cpp:1369-1465) and the ROPs blend fragments in
// primitive order. The CUDA rasterizer reproduces that order per pixel with per-tile lists:
//   k_bin_emit   walks the depth-sorted ids in order, looks up each splat's pixel bounding box
//                (written by the preprocess kernel), and appends one (tile id, splat id) pair per
//                covered TILE_W x TILE_H tile at an offset given by a decoupled look-back prefix sum — so
//                pairs are emitted in depth order;
//   a stable 2-pass (16-bit) radix sort on the tile id (k_radix_sort.cu) groups them per tile and
//                keeps the depth order inside every tile;
//   the last sort pass also records each tile's [begin,end) while it scatters (atomicMin/Max).
// The digit histograms for the tile sort are accumulated by k_bin_emit while it drains its staging
// buffer.
#include "device_common.cuh"
#include "kernels.hpp"

namespace vkgs {

namespace {

constexpr int BIN_WARPS    = BIN_THREADS / 32;
constexpr int BIN_ITEMS = 4;                        // depth ranks per thread (one 16-byte id load)
constexpr int BIN_PART  = BIN_THREADS * BIN_ITEMS;  // ranks per partition
#ifndef VKGS_BIN_CHUNK
#define VKGS_BIN_CHUNK 2048
#endif
constexpr int BIN_CHUNK = VKGS_BIN_CHUNK;           // pairs staged in shared memory per round
constexpr uint32_t BIN_BIG = 32;                    // splats covering more tiles than this are expanded warp-cooperatively
constexpr uint32_t BIN_HUGE = 256;                  // ... and more than this by a whole CTA of k_bin_big (see below)

__global__ void __launch_bounds__(BIN_THREADS) k_bin_emit(const __grid_constant__ BinArgs a)
{
  __shared__ uint32_t s_keys[BIN_CHUNK];
  __shared__ uint32_t s_vals[BIN_CHUNK];
  __shared__ uint32_t s_whist[BIN_WARPS][2][256];  // interleaved copies of the digit tables of the tile ids
  __shared__ uint32_t s_scan[BIN_WARPS + 1];
  __shared__ uint32_t s_part, s_base;
  const unsigned      tid = threadIdx.x, lane = tid & 31u, warp = tid >> 5;

  if(tid == 0)
    s_part = atomicAdd(&a.counters->ticket[a.ticketSlot], 1u);
  for(int i = tid; i < BIN_WARPS * 2 * 256; i += BIN_THREADS)
    (&s_whist[0][0][0])[i] = 0u;
  pdl_wait();  // (everything above is independent of the depth sort before this kernel)
  pdl_launch_dependents();
  const uint32_t count = a.counters->visible;
  const uint32_t parts = (count + BIN_PART - 1) / BIN_PART;
  const uint32_t* __restrict__ sortedIds = a.sortedIds[a.sortedSel ? *a.sortedSel : 0u];
  __syncthreads();
  const uint32_t part = s_part;
  if(part >= parts)
    return;
  VKGS_TL(part, 0);

  // thread t owns ranks r0 .. r0+3 (blocked: emission order == rank order)
  const uint32_t r0 = part * BIN_PART + tid * BIN_ITEMS;
  uint32_t       id[BIN_ITEMS];
  if(r0 + BIN_ITEMS <= count)
  {
    const uint4 v = *reinterpret_cast<const uint4*>(sortedIds + r0);
    id[0] = v.x, id[1] = v.y, id[2] = v.z, id[3] = v.w;
  }
  else
  {
#pragma unroll
    for(int i = 0; i < BIN_ITEMS; i++)
      id[i] = (r0 + i < count) ? sortedIds[r0 + i] : 0xffffffffu;
  }
  // four independent 8-byte gathers of the pixel bounding boxes
  uint2 bb[BIN_ITEMS];
#pragma unroll
  for(int i = 0; i < BIN_ITEMS; i++)
    bb[i] = (id[i] != 0xffffffffu) ? __ldg(a.bboxes + id[i]) : make_uint2(1u, 0u);
  uint32_t x0[BIN_ITEMS], nx[BIN_ITEMS], y0[BIN_ITEMS], n[BIN_ITEMS], mine = 0;
#pragma unroll
  for(int i = 0; i < BIN_ITEMS; i++)
  {
    const uint32_t px0 = bb[i].x & 0xffffu, py0 = bb[i].x >> 16, px1 = bb[i].y & 0xffffu, py1 = bb[i].y >> 16;
    const bool     ok  = px1 >= px0 && py1 >= py0;
    x0[i]              = px0 / TILE_W;
    y0[i]              = py0 / TILE_H;
    nx[i]              = ok ? (px1 / TILE_W - x0[i] + 1) : 0u;
    n[i]               = ok ? nx[i] * (py1 / TILE_H - y0[i] + 1) : 0u;
    mine += n[i];
  }

  VKGS_TL(part, 1);
  uint32_t       total;
  const uint32_t local = block_exclusive_scan<BIN_WARPS>(mine, s_scan, total);
  VKGS_TL(part, 2);

  // Publish this partition's pair count at once (successors only need the aggregate) ...
  if(tid == 0)
    lb_store(a.status + part, lb_pack(a.epoch, part == 0 ? LB_INCLUSIVE : LB_AGGREGATE, total));

  // resolve the exclusive prefix of the pair counts (warp 0 only; result in s_base after the next barrier)
  auto resolvePrefix = [&]() {
    uint32_t excl = 0;
    if(part != 0)
    {
      if(a.debugFlags & 32u)
      {
        if(tid == 0)
          excl = atomicAdd(&a.counters->tilePairs, total);
        excl = __shfl_sync(FULL_MASK, excl, 0);
      }
      else
        excl = lb_lookback_warp<4>(a.status, part, a.epoch);
      if(tid == 0)
        lb_store(a.status + part, lb_pack(a.epoch, LB_INCLUSIVE, excl + total));
    }
    if(tid == 0)
    {
      s_base = excl;
      // 32-bit pair positions: a frame that wants 2^32 pairs or more is an overflow, not a wrap-around (every partition
      // checks its own end, so a carry in the middle of the frame is seen too)
      const uint64_t d64 = static_cast<uint64_t>(excl) + total;
      const uint32_t d   = d64 > 0xffffffffull ? 0xffffffffu : static_cast<uint32_t>(d64);
      if(d64 > a.capacity)
      {
        a.counters->overflow       = 1u;
        a.counters->stickyOverflow = 1u;
        atomicMax(&a.counters->stickyPairs, d);
      }
      if(part == parts - 1)
      {
        a.counters->tilePairs        = d;
        a.counters->tilePairsClamped = d < a.capacity ? d : a.capacity;
      }
    }
  };

  // Partitions holding HUGE splats (camera close to / inside the scene: a splat can cover the whole screen,
  // and depth order puts all of them into the same few partitions) take a different route: no staging rounds;
  // huge splats are handed to k_bin_big with their absolute output offset (a whole CTA expands each), the rest
  // is written straight to its final position. Rare, so the common path below stays as it is.
  bool hasHuge = false;
#pragma unroll
  for(int i = 0; i < BIN_ITEMS; i++)
    hasHuge = hasHuge || n[i] > BIN_HUGE;
  if(__syncthreads_or(hasHuge))
  {
    if(tid < 32)
      resolvePrefix();
    __syncthreads();
    const uint32_t base = s_base;
    uint32_t       off  = local;
#pragma unroll
    for(int i = 0; i < BIN_ITEMS; i++)
    {
      const uint64_t g0 = static_cast<uint64_t>(base) + off;
      if(n[i] > BIN_HUGE)
      {
        const uint32_t slot = atomicAdd(&a.counters->bigCount, 1u);
        if(slot < a.bigCapacity)
          // (an entry whose pairs would end beyond 2^32 is unusable: the claimed slot gets a zero-size entry, never a stale one)
          a.bigList[slot] = (g0 + n[i] <= 0xffffffffull) ? make_uint4(static_cast<uint32_t>(g0), id[i], x0[i] | (y0[i] << 16), nx[i] | ((n[i] / nx[i]) << 16))
                                                         : make_uint4(0u, 0u, 0u, 0u);
        else
        {
          // list full: expand it here after all (slow, still correct)
          for(uint32_t j = 0; j < n[i]; j++)
            if(g0 + j < a.capacity)
            {
              const uint32_t key = (y0[i] + j / nx[i]) * a.tilesX + x0[i] + j % nx[i];
              a.tileKeys[g0 + j] = key, a.tileVals[g0 + j] = id[i];
              atomicAdd(&s_whist[lane & (BIN_WARPS - 1)][0][key & 0xffu], 1u);
              atomicAdd(&s_whist[lane & (BIN_WARPS - 1)][1][(key >> 8) & 0xffu], 1u);
            }
        }
      }
      // everything else: by the owning warp, 32 pairs per step
      unsigned mask = __ballot_sync(FULL_MASK, n[i] != 0u && n[i] <= BIN_HUGE);
      while(mask)
      {
        const int src = __ffs(mask) - 1;
        mask &= mask - 1;
        const uint64_t bG0 = __shfl_sync(FULL_MASK, g0, src);
        const uint32_t bN = __shfl_sync(FULL_MASK, n[i], src), bNx = __shfl_sync(FULL_MASK, nx[i], src);
        const uint32_t bX0 = __shfl_sync(FULL_MASK, x0[i], src), bY0 = __shfl_sync(FULL_MASK, y0[i], src), bId = __shfl_sync(FULL_MASK, id[i], src);
        for(uint32_t j = lane; j < bN; j += 32)
          if(bG0 + j < a.capacity)
          {
            const uint32_t ty = j / bNx, tx = j - ty * bNx, key = (bY0 + ty) * a.tilesX + bX0 + tx;
            a.tileKeys[bG0 + j] = key, a.tileVals[bG0 + j] = bId;
            atomicAdd(&s_whist[lane & (BIN_WARPS - 1)][0][key & 0xffu], 1u);
            atomicAdd(&s_whist[lane & (BIN_WARPS - 1)][1][(key >> 8) & 0xffu], 1u);
          }
      }
      off += n[i];
    }
    __syncthreads();
  }
  else
  {
  // ... emit through shared memory so global writes are fully coalesced: every round stages up
  // to BIN_CHUNK pairs of the block's contiguous output range and drains them row by row (the drain
  // also counts the two 8-bit digits of each tile id). The first round is staged BEFORE the
  // look-back is resolved: it only needs block-local offsets, and by the time it is done the
  // predecessors have published.
  uint32_t blockBase = 0;
  for(uint32_t w = 0; (w < total || w == 0) && !(a.debugFlags & 16u); w += BIN_CHUNK)
  {
    const uint32_t lim = min(total, w + BIN_CHUNK);
    uint32_t       off = local;
#pragma unroll
    for(int i = 0; i < BIN_ITEMS; i++)
    {
      const uint32_t lo = max(off, w), hi = min(off + n[i], lim);
      const bool     big = n[i] > BIN_BIG;
      if(lo < hi && !big)
      {
        // (the integer division only when the splat's pairs straddle a staging round; the key advances incrementally:
        //  +1 along a tile row, + tilesX - nx + 1 at its end)
        const uint32_t j  = lo - off;
        uint32_t       ty = 0, tx = 0;
        if(j != 0u)
          ty = j / nx[i], tx = j - ty * nx[i];
        uint32_t key = (y0[i] + ty) * a.tilesX + x0[i] + tx;
        for(uint32_t p = lo; p < hi; p++)
        {
          s_keys[p - w] = key;
          s_vals[p - w] = id[i];
          key++;
          if(++tx == nx[i])
            tx = 0, key += a.tilesX - nx[i];
        }
      }
      // a splat covering many tiles (close to the camera: up to the whole screen) is expanded by its
      // whole warp, 32 pairs per step, instead of one thread walking thousands of tiles
      unsigned bigMask = __ballot_sync(FULL_MASK, big && lo < hi);
      while(bigMask)
      {
        const int src = __ffs(bigMask) - 1;
        bigMask &= bigMask - 1;
        const uint32_t bOff = __shfl_sync(FULL_MASK, off, src), bLo = __shfl_sync(FULL_MASK, lo, src), bHi = __shfl_sync(FULL_MASK, hi, src);
        const uint32_t bNx = __shfl_sync(FULL_MASK, nx[i], src), bX0 = __shfl_sync(FULL_MASK, x0[i], src), bY0 = __shfl_sync(FULL_MASK, y0[i], src);
        const uint32_t bId = __shfl_sync(FULL_MASK, id[i], src);
        for(uint32_t p = bLo + lane; p < bHi; p += 32)
        {
          const uint32_t j = p - bOff, ty = j / bNx, tx = j - ty * bNx;
          s_keys[p - w]    = (bY0 + ty) * a.tilesX + bX0 + tx;
          s_vals[p - w]    = bId;
        }
      }
      off += n[i];
    }
    if(w == 0)
    {
      // resolve the exclusive prefix of the pair counts (warp 0), everybody else is still staging
      if(tid < 32)
        resolvePrefix();
    }
    __syncthreads();
    if(w == 0)
    {
      blockBase = s_base;
      VKGS_TL(part, 3);
    }
    const uint32_t cnt = lim - w;
    for(uint32_t qb = warp * 32; qb < cnt; qb += BIN_THREADS)
    {
      const uint32_t q  = qb + lane;
      const uint64_t g  = static_cast<uint64_t>(blockBase) + w + q;
      if(q < cnt && g < a.capacity)
      {
        const uint32_t key = s_keys[q];
        a.tileKeys[g]      = key;
        a.tileVals[g]      = s_vals[q];
        // digit histograms for the tile sort: shared-memory atomics on a lane-interleaved copy
        // (measured on B200: > 7 shared atomics per cycle per SM, far cheaper than a ballot multi-split)
        if(!(a.debugFlags & 64u))
        {
          atomicAdd(&s_whist[lane & (BIN_WARPS - 1)][0][key & 0xffu], 1u);
          atomicAdd(&s_whist[lane & (BIN_WARPS - 1)][1][(key >> 8) & 0xffu], 1u);
        }
      }
    }
    __syncthreads();
  }
  }  // common path
  VKGS_TL(part, 4);
  for(int i = tid; i < 2 * 256; i += BIN_THREADS)
  {
    uint32_t v = 0;
#pragma unroll
    for(int wv = 0; wv < BIN_WARPS; wv++)
      v += (&s_whist[wv][0][0])[i];
    if(v)
      atomicAdd(&a.counters->tileHist[0][0] + i, v);
  }
}

// Huge splats handed over by k_bin_emit: entry = (absolute output offset, splat id, x0 | y0 << 16, nx | ny << 16) in
// tiles. One CTA per entry (grid-strided) writes its nx * ny (tile id, splat id) pairs, coalesced, and counts
// their digits for the tile sort.
__global__ void __launch_bounds__(BIN_THREADS) k_bin_big(const __grid_constant__ BinArgs a)
{
  __shared__ uint32_t s_whist[BIN_WARPS][2][256];
  const unsigned      tid = threadIdx.x, lane = tid & 31u;
  pdl_wait();
  pdl_launch_dependents();
  const uint32_t      count = min(a.counters->bigCount, a.bigCapacity);
  if(blockIdx.x >= count)
    return;
  for(int i = tid; i < BIN_WARPS * 2 * 256; i += BIN_THREADS)
    (&s_whist[0][0][0])[i] = 0u;
  __syncthreads();
  for(uint32_t e = blockIdx.x; e < count; e += gridDim.x)
  {
    const uint4    it = a.bigList[e];
    const uint32_t x0 = it.z & 0xffffu, y0 = it.z >> 16, nx = it.w & 0xffffu, n = nx * (it.w >> 16);
    for(uint32_t j = tid; j < n; j += BIN_THREADS)
    {
      const uint64_t g = static_cast<uint64_t>(it.x) + j;
      if(g < a.capacity)
      {
        const uint32_t ty = j / nx, tx = j - ty * nx, key = (y0 + ty) * a.tilesX + x0 + tx;
        a.tileKeys[g] = key, a.tileVals[g] = it.y;
        atomicAdd(&s_whist[lane & (BIN_WARPS - 1)][0][key & 0xffu], 1u);
        atomicAdd(&s_whist[lane & (BIN_WARPS - 1)][1][(key >> 8) & 0xffu], 1u);
      }
    }
  }
  __syncthreads();
  for(int i = tid; i < 2 * 256; i += BIN_THREADS)
  {
    uint32_t v = 0;
#pragma unroll
    for(int wv = 0; wv < BIN_WARPS; wv++)
      v += (&s_whist[wv][0][0])[i];
    if(v)
      atomicAdd(&a.counters->tileHist[0][0] + i, v);
  }
}

}  // namespace

void launchBinBig(const BinArgs& args, cudaStream_t stream, bool pdl)
{
  // (launched every frame: the list length is only known on the device; CTAs beyond it exit at once)
  launchKernelPdl(k_bin_big, 148 * 2, BIN_THREADS, 0, stream, args, pdl);
}

void launchBinEmit(const BinArgs& args, cudaStream_t stream, bool pdl)
{
  const uint32_t parts = (args.maxCount + BIN_PART - 1) / BIN_PART;
  if(parts == 0)
    return;
  launchKernelPdl(k_bin_emit, parts, BIN_THREADS, 0, stream, args, pdl);
}

}  // namespace vkgs
