// k_binning.cu — per-tile splat lists for the software rasterizer.
//
// The reference hands one screen-aligned quad per splat to the hardware rasterizer in sorted order
// (drawSplatPrimitives, src/gaussian_splatting.cpp:1369-1465) and the ROPs blend fragments in
// primitive order. The CUDA rasterizer reproduces that order per pixel with per-tile lists:
//   k_bin_emit   walks the depth-sorted ids in order, looks up each splat's pixel bounding box
//                (written by the preprocess kernel), and appends one (tile id, splat id) pair per
//                covered 16x16 tile at an offset given by a decoupled look-back prefix sum — so
//                pairs are emitted in depth order;
//   a stable 2-pass (16-bit) radix sort on the tile id (k_radix_sort.cu) groups them per tile and
//                keeps the depth order inside every tile;
//   k_tile_ranges finds each tile's [begin,end) in the sorted list.
// The digit histograms for the tile sort are accumulated by k_bin_emit.
#include "device_common.cuh"
#include "kernels.hpp"

namespace vkgs {

namespace {

constexpr int NWARPS = BIN_THREADS / 32;

__global__ void __launch_bounds__(BIN_THREADS) k_bin_emit(const __grid_constant__ BinArgs a)
{
  __shared__ uint32_t s_hist[2][256];
  __shared__ uint32_t s_scan[NWARPS + 1];
  __shared__ uint32_t s_part, s_base;
  const unsigned      tid = threadIdx.x;

  const uint32_t count = a.counters->visible;
  const uint32_t parts = (count + BIN_THREADS - 1) / BIN_THREADS;
  if(tid == 0)
    s_part = atomicAdd(&a.counters->ticket[a.ticketSlot], 1u);
  s_hist[0][tid] = 0u;
  s_hist[1][tid] = 0u;
  __syncthreads();
  const uint32_t part = s_part;
  if(part >= parts)
    return;

  const uint32_t r  = part * BIN_THREADS + tid;  // depth rank
  uint32_t       id = 0, x0 = 1, x1 = 0, y0 = 1, y1 = 0;
  if(r < count)
  {
    id             = a.sortedIds[r];
    const uint2 bb = *reinterpret_cast<const uint2*>(a.records + static_cast<uint64_t>(id) * RECORD_WORDS + 10);
    if((bb.y & 0xffffu) >= (bb.x & 0xffffu))
    {
      x0 = (bb.x & 0xffffu) / TILE_W, y0 = (bb.x >> 16) / TILE_H;
      x1 = (bb.y & 0xffffu) / TILE_W, y1 = (bb.y >> 16) / TILE_H;
    }
  }
  const uint32_t nTiles = (x1 >= x0 && y1 >= y0) ? (x1 - x0 + 1) * (y1 - y0 + 1) : 0u;

  uint32_t       total;
  const uint32_t local = block_exclusive_scan<NWARPS>(nTiles, s_scan, total);

  if(tid == 0)
  {
    uint64_t* st = a.status + part;
    uint32_t  excl = 0;
    if(part == 0)
      lb_store(st, lb_pack(a.epoch, LB_INCLUSIVE, total));
    else
    {
      lb_store(st, lb_pack(a.epoch, LB_AGGREGATE, total));
      excl = lb_lookback(a.status, part, 1, a.epoch);
      lb_store(st, lb_pack(a.epoch, LB_INCLUSIVE, excl + total));
    }
    s_base = excl;
    if(part == parts - 1)
    {
      const uint32_t d              = excl + total;
      a.counters->tilePairs         = d;
      a.counters->tilePairsClamped  = d < a.capacity ? d : a.capacity;
      if(d > a.capacity)
        a.counters->overflow = 1u;
    }
  }
  __syncthreads();

  uint32_t off = s_base + local;
  for(uint32_t ty = y0; ty <= y1 && nTiles; ty++)
    for(uint32_t tx = x0; tx <= x1; tx++)
    {
      const uint32_t key = ty * a.tilesX + tx;
      if(off < a.capacity)
      {
        a.tileKeys[off] = key;
        a.tileVals[off] = id;
        atomicAdd(&s_hist[0][key & 0xffu], 1u);
        atomicAdd(&s_hist[1][(key >> 8) & 0xffu], 1u);
      }
      off++;
    }
  __syncthreads();
  for(int i = tid; i < 2 * 256; i += BIN_THREADS)
  {
    const uint32_t v = (&s_hist[0][0])[i];
    if(v)
      atomicAdd(&a.counters->tileHist[0][0] + i, v);
  }
}

__global__ void __launch_bounds__(256) k_tile_ranges(const uint32_t* __restrict__ tileKeys, const FrameCounters* counters, uint2* ranges)
{
  const uint32_t count = counters->tilePairsClamped;
  for(uint64_t i = static_cast<uint64_t>(blockIdx.x) * 256 + threadIdx.x; i < count; i += static_cast<uint64_t>(gridDim.x) * 256)
  {
    const uint32_t k = tileKeys[i];
    if(i == 0 || tileKeys[i - 1] != k)
      ranges[k].x = static_cast<uint32_t>(i);
    if(i + 1 == count || tileKeys[i + 1] != k)
      ranges[k].y = static_cast<uint32_t>(i + 1);
  }
}

}  // namespace

void launchBinEmit(const BinArgs& args, cudaStream_t stream)
{
  const uint32_t parts = (args.maxCount + BIN_THREADS - 1) / BIN_THREADS;
  if(parts == 0)
    return;
  k_bin_emit<<<parts, BIN_THREADS, 0, stream>>>(args);
}

void launchTileRanges(const uint32_t* tileKeys, const FrameCounters* counters, uint32_t capacity, uint2* ranges, cudaStream_t stream)
{
  uint32_t blocks = (capacity + 256 * 8 - 1) / (256 * 8);
  blocks          = blocks < 1 ? 1 : (blocks > 148 * 16 ? 148 * 16 : blocks);
  k_tile_ranges<<<blocks, 256, 0, stream>>>(tileKeys, counters, ranges);
}

}  // namespace vkgs
