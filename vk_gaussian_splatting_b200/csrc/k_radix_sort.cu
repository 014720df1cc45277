// k_radix_sort.cu — stable LSD radix sort of (u32 key, u32 value) pairs, 8 bits per pass.
//
// Replaces the vrdx key-value sort the reference records each frame
// (vrdxCmdSortKeyValueIndirect, 3rdparty/vrdx/src/vk_radix_sort.cc:249-258,262-416; shaders
// upsweep/spine/downsweep.slang). Same contract: ascending, stable, pair count read on the device.
// vrdx is reduce-then-scan (3 dispatches and ~20 B/pair per pass). This is a single-pass
// ("onesweep") design instead: the digit histograms of all passes are produced by whichever
// kernel wrote the keys, and each pass is ONE kernel that ranks a 4096-pair partition in shared
// memory, resolves its global digit offsets with a decoupled look-back over earlier partitions
// (epoch-stamped 64-bit status words: nothing is cleared between passes or frames) and scatters
// through shared memory so global writes leave in digit-contiguous runs. Traffic per pass is one
// read + one write of the pairs (16 B/pair), i.e. 64 B/pair for 32-bit keys.
//
// Stability: a partition is ranked in (warp, row, lane) order, which is the input order because
// each warp loads a contiguous chunk in a warp-striped arrangement; partitions are ordered by a
// ticket taken at block start, which also guarantees the look-back never waits on a block that
// has not been scheduled.
#include "device_common.cuh"
#include "kernels.hpp"

namespace vkgs {

namespace {

constexpr int NWARPS = SORT_THREADS / 32;
static_assert(SORT_THREADS == 256, "one thread per digit: the publication, the scans and the look-back are written for 256 threads");
static_assert(SORT_ITEMS % 2 == 0, "rows are ranked two at a time");

// Shared memory of a partition. The ranking tables are dead once every key has its rank, and the staging area is only
// written after that point, so the two alias.
struct SortSmem
{
  union
  {
    struct
    {
      // mask[w][p][r][d]: lanes of warp w whose key of row r of the current row pair has digit d (p = parity of the pair:
      // a pair's masks are cleared while the next pair already fills the other copy)
      uint32_t mask[NWARPS][2][2][256];
      uint32_t cnt[NWARPS][256];  // keys with digit d in the rows warp w has ranked so far; then its exclusive offset over warps
    } rank;
    struct
    {
      uint32_t keys[SORT_PART];  // block-sorted pairs
      uint32_t vals[SORT_PART];
    } stage;
  };
  uint32_t digitCount[256];  // partition digit counts (counted right after the load), then each digit's first block-local position
  uint32_t delta[256];       // global position of the digit's first key of this partition MINUS its block-local position
  uint32_t scan[NWARPS + 1];
  uint32_t warpTot[NWARPS];
  uint32_t part;
};

// One 8-bit pass over one 4096-pair partition.
//
// Ranking. A key's position inside the partition is (keys with a smaller digit) + (keys with the same digit earlier in
// the input). Each warp owns 512 consecutive pairs as 16 rows of 32 (lane = consecutive pair), so "earlier" means an
// earlier row of the same warp, a lower lane of the same row, or a lower warp. Within a row the lanes that share a
// digit find each other through shared memory: every lane ORs its lane bit into the row's mask word of its digit, and
// after a warp barrier reads the word back — the peer mask — at a cost that does not depend on how many distinct digits
// the row holds (8 ballots + 8 selects per key before; MATCH.ANY is serviced once per distinct value). The lowest peer
// adds the group size to the warp's running digit counter and clears the word. Two rows are ranked per step (two mask
// copies; the second row also reads the first row's mask of ITS digit), which halves the warp barriers.
template <bool RANGES>
__global__ void __launch_bounds__(SORT_THREADS, 4) k_sort_pass(const __grid_constant__ SortPassArgs a)
{
  extern __shared__ __align__(16) unsigned char smemRaw[];
  SortSmem&      sm   = *reinterpret_cast<SortSmem*>(smemRaw);
  const unsigned tid  = threadIdx.x;
  const unsigned lane = tid & 31u, warp = tid >> 5;

#ifdef VKGS_TIMELINE
  const long long tl0 = clock64();
#endif
  // what does not depend on the kernel before this one: the ticket (the counter was cleared at the start of the frame) and
  // the shared-memory tables
  if(tid == 0)
    sm.part = atomicAdd(a.ticket, 1u);
  {
    uint4* z = reinterpret_cast<uint4*>(&sm.rank);
#pragma unroll
    for(int i = 0; i < static_cast<int>(sizeof(sm.rank) / 16 / SORT_THREADS); i++)
      z[i * SORT_THREADS + tid] = make_uint4(0u, 0u, 0u, 0u);
    static_assert(sizeof(sm.rank) % (16 * SORT_THREADS) == 0, "whole uint4 rounds");
  }
  sm.digitCount[tid] = 0u;
  pdl_wait();
  pdl_launch_dependents();
#ifdef VKGS_NO_COUNT_CLAMP  // (A/B builds only)
  const uint32_t count = *a.countPtr;
#else
  const uint32_t count = min(*a.countPtr, a.maxCount);  // (a device-side count beyond the host bound is the caller's bug: no reads past the buffers)
#endif
  const uint32_t parts = (count + SORT_PART - 1) / SORT_PART;
  uint32_t       cur   = a.srcSelIn ? *a.srcSelIn : 0u;
  if(a.srcSelOut)
  {
    // identity pass (every key has the same digit, e.g. the exponent byte of NDC depths): skip it
    const int same = __syncthreads_or(count == 0u || a.histogram[tid] == count);
    if(same)
    {
      if(blockIdx.x == 0 && tid == 0)
        *a.srcSelOut = cur;
      return;
    }
  }
  const uint32_t* __restrict__ keysIn  = a.keys[cur];
  const uint32_t* __restrict__ valsIn  = a.vals[cur];
  uint32_t* __restrict__       keysOut = a.keys[cur ^ 1u];
  uint32_t* __restrict__       valsOut = a.vals[cur ^ 1u];
  __syncthreads();
  const uint32_t part = sm.part;
  if(part >= parts)
    return;
#ifdef VKGS_TIMELINE
  if(tid == 0 && g_vkgsTimeline)
    g_vkgsTimeline[part * 16 + 0] = tl0;
#endif
  VKGS_TL(part, 1);

  // ---- load keys (warp-striped: warp w owns SORT_ITEMS*32 consecutive pairs) ------------------
  const uint32_t partBase = part * SORT_PART;
  const uint32_t warpBase = partBase + warp * (SORT_ITEMS * 32);
  const bool     full     = partBase + SORT_PART <= count;  // (block-uniform)
  uint32_t       keys[SORT_ITEMS];
  if(full)
  {
#pragma unroll
    for(int i = 0; i < SORT_ITEMS; i++)
      keys[i] = keysIn[warpBase + i * 32 + lane];
  }
  else
  {
#pragma unroll
    for(int i = 0; i < SORT_ITEMS; i++)
    {
      const uint32_t idx = warpBase + i * 32 + lane;
      keys[i]            = idx < count ? keysIn[idx] : 0xffffffffu;  // padding sorts last, is never written back
    }
  }
  VKGS_TL(part, 2);

  // ---- count the partition's digits FIRST (shared atomics, a few hundred cycles) and publish them:
  // successors' look-backs only need these counts, so they must not wait for the ranking below ------
#pragma unroll
  for(int i = 0; i < SORT_ITEMS; i++)
    atomicAdd(&sm.digitCount[(keys[i] >> a.shift) & 0xffu], 1u);
  __syncthreads();
  uint32_t  realCount = sm.digitCount[tid];
  uint64_t* mine      = a.status + static_cast<uint64_t>(part) * 256 + tid;
  {
    // padding keys (digit 0xff of the last partition) are not counted
    if(tid == 255 && !full)
      realCount -= (partBase + SORT_PART - count);
    if(part == 0)
    {
      // exclusive scan of the global digit histogram = first global position of each digit.
      // (only partition 0 needs it; everyone else inherits it through the look-back chain)
      const uint32_t h   = a.histogram[tid];
      const uint32_t inc = warp_inclusive_scan(h, lane);
      if(lane == 31)
        sm.warpTot[warp] = inc;
      __syncthreads();  // (block-uniform branch)
      uint32_t add = 0;
      for(unsigned w = 0; w < warp; w++)
        add += sm.warpTot[w];
      const uint32_t excl = inc - h + add;
      lb_store(mine, lb_pack(a.epoch, LB_INCLUSIVE, excl + realCount));
      sm.delta[tid] = excl;
      if(tid == 0 && a.srcSelOut)
        *a.srcSelOut = cur ^ 1u;
    }
    else
      lb_store(mine, lb_pack(a.epoch, LB_AGGREGATE, realCount));
  }

  // ---- rank within the warp --------------------------------------------------------------------
  uint32_t rank2[SORT_ITEMS / 2];  // two 16-bit ranks per word (a warp holds 512 keys)
  {
    uint32_t*      cnt     = sm.rank.cnt[warp];
    const uint32_t laneBit = 1u << lane, lower = laneBit - 1u;
#pragma unroll
    for(int ii = 0; ii < SORT_ITEMS / 2; ii++)
    {
      uint32_t*      m0 = sm.rank.mask[warp][ii & 1][0];
      uint32_t*      m1 = sm.rank.mask[warp][ii & 1][1];
      const uint32_t d0 = (keys[2 * ii] >> a.shift) & 0xffu, d1 = (keys[2 * ii + 1] >> a.shift) & 0xffu;
      atomicOr(&m0[d0], laneBit);
      atomicOr(&m1[d1], laneBit);
      __syncwarp();
      const uint32_t peers0 = m0[d0], peers1 = m1[d1], row0same = m0[d1];
      const uint32_t b0 = cnt[d0], b1 = cnt[d1];
      __syncwarp();
      const uint32_t r0 = b0 + __popc(peers0 & lower);
      const uint32_t r1 = b1 + __popc(row0same) + __popc(peers1 & lower);
      rank2[ii]         = r0 | (r1 << 16);
      if((peers0 & lower) == 0u)  // lowest lane of the group (two groups of a pair may share a digit: atomic add)
      {
        atomicAdd(&cnt[d0], static_cast<uint32_t>(__popc(peers0)));
        m0[d0] = 0u;
      }
      if((peers1 & lower) == 0u)
      {
        atomicAdd(&cnt[d1], static_cast<uint32_t>(__popc(peers1)));
        m1[d1] = 0u;
      }
    }
  }
  __syncthreads();
  VKGS_TL(part, 3);

  // ---- per digit: exclusive scan over warps; exclusive scan over digits --------------------------
  {
    uint32_t run = 0;
#pragma unroll
    for(int w = 0; w < NWARPS; w++)
    {
      const uint32_t c     = sm.rank.cnt[w][tid];
      sm.rank.cnt[w][tid]  = run;
      run += c;
    }
    uint32_t       total;
    const uint32_t excl = block_exclusive_scan<NWARPS>(run, sm.scan, total);
    sm.digitCount[tid]  = excl;  // from here on: first block-local position of the digit
  }
  __syncthreads();
  VKGS_TL(part, 4);

  // ---- block-local positions; stage pairs in sorted order --------------------------------------
  uint32_t pos[SORT_ITEMS];
#pragma unroll
  for(int i = 0; i < SORT_ITEMS; i++)
  {
    const uint32_t digit = (keys[i] >> a.shift) & 0xffu;
    pos[i]               = sm.digitCount[digit] + sm.rank.cnt[warp][digit] + ((rank2[i >> 1] >> (16 * (i & 1))) & 0xffffu);
  }
  __syncthreads();  // the ranking tables are dead from here: their storage becomes the staging area
#pragma unroll
  for(int i = 0; i < SORT_ITEMS; i++)
    sm.stage.keys[pos[i]] = keys[i];
  // values go straight from global to their staged position
  if(full)
  {
    uint32_t vals[SORT_ITEMS];
#pragma unroll
    for(int i = 0; i < SORT_ITEMS; i++)
      vals[i] = valsIn[warpBase + i * 32 + lane];
#pragma unroll
    for(int i = 0; i < SORT_ITEMS; i++)
      sm.stage.vals[pos[i]] = vals[i];
  }
  else
  {
#pragma unroll
    for(int i = 0; i < SORT_ITEMS; i++)
    {
      const uint32_t idx    = warpBase + i * 32 + lane;
      sm.stage.vals[pos[i]] = idx < count ? valsIn[idx] : 0u;
    }
  }
  VKGS_TL(part, 5);

  // ---- global digit offsets: decoupled look-back, one chain per digit (= per thread) -------------
  {
    uint32_t excl;
    if(part != 0)
    {
      excl = lb_lookback<16>(a.status + tid, part, 256, a.epoch);
      lb_store(mine, lb_pack(a.epoch, LB_INCLUSIVE, excl + realCount));
    }
    else
      excl = sm.delta[tid];
    sm.delta[tid] = excl - sm.digitCount[tid];
  }
  VKGS_TL(part, 6);
  __syncthreads();
  VKGS_TL(part, 7);

  // ---- scatter: consecutive threads write consecutive positions of a digit run -----------------
  const uint32_t valid = min(static_cast<uint32_t>(SORT_PART), count - partBase);
#pragma unroll
  for(int i = 0; i < SORT_ITEMS; i++)
  {
    const uint32_t j = i * SORT_THREADS + tid;
    if(j < valid)
    {
      const uint32_t k   = sm.stage.keys[j];
      const uint32_t dst = j + sm.delta[(k >> a.shift) & 0xffu];
      keysOut[dst]       = k;
      valsOut[dst]       = sm.stage.vals[j];
      if(RANGES)
      {
        // Final pass of the tile sort: equal keys are contiguous in the staged (block-sorted) order — the input of the
        // last pass is ordered by the low digit, so inside every high-digit run the full key is non-decreasing
        if(j == 0 || sm.stage.keys[j - 1] != k)
          atomicMin(a.rangeBegin + k, dst);
        if(j + 1 == valid || sm.stage.keys[j + 1] != k)
          atomicMax(a.rangeEnd + k, dst + 1u);
      }
    }
  }
  VKGS_TL(part, 8);
}

// Digit histograms of all passes in one sweep over the keys (stand-alone sort only: the frame
// pipeline produces its histograms inside the kernels that write the keys). Each thread reads 16
// keys with 16-byte loads and counts every digit with a shared-memory atomic into one of
// HIST_COPIES block-private copies of the table (lane-interleaved, which spreads same-digit
// collisions); copies are reduced at the end and flushed with one global atomic per non-empty bin.
constexpr int HIST_THREADS = 512;
constexpr int HIST_COPIES  = 4;

template <int PASSES>
__global__ void __launch_bounds__(HIST_THREADS) k_histogram(const uint32_t* __restrict__ keys, const uint32_t* countPtr, uint32_t maxCount,
                                                            uint32_t* hist, int firstShift)
{
  __shared__ uint32_t s[HIST_COPIES][PASSES][256];
  const unsigned      tid = threadIdx.x;
  for(int i = tid; i < HIST_COPIES * PASSES * 256; i += HIST_THREADS)
    (&s[0][0][0])[i] = 0u;
  __syncthreads();
  const uint32_t count = min(*countPtr, maxCount);
  uint32_t (*mine)[256] = s[tid % HIST_COPIES];
  const uint64_t vecs   = count / 4;
  const uint4*   k4     = reinterpret_cast<const uint4*>(keys);
  for(uint64_t i = static_cast<uint64_t>(blockIdx.x) * HIST_THREADS + tid; i < vecs; i += static_cast<uint64_t>(gridDim.x) * HIST_THREADS)
  {
    const uint4    v    = k4[i];
    const uint32_t k[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
    for(int j = 0; j < 4; j++)
#pragma unroll
      for(int p = 0; p < PASSES; p++)
        atomicAdd(&mine[p][(k[j] >> (firstShift + 8 * p)) & 0xffu], 1u);
  }
  if(blockIdx.x == 0 && tid < (count & 3u))
  {
    const uint32_t kk = keys[vecs * 4 + tid];
    for(int p = 0; p < PASSES; p++)
      atomicAdd(&mine[p][(kk >> (firstShift + 8 * p)) & 0xffu], 1u);
  }
  __syncthreads();
  for(int i = tid; i < PASSES * 256; i += HIST_THREADS)
  {
    uint32_t v = 0;
#pragma unroll
    for(int c = 0; c < HIST_COPIES; c++)
      v += (&s[c][0][0])[i];
    if(v)
      atomicAdd(hist + i, v);
  }
}

}  // namespace

void launchSortPass(const SortPassArgs& args, cudaStream_t stream, bool pdl)
{
  const uint32_t parts = (args.maxCount + SORT_PART - 1) / SORT_PART;
  if(parts == 0)
    return;
  if(args.rangeBegin)
    launchKernelPdl(k_sort_pass<true>, parts, SORT_THREADS, sizeof(SortSmem), stream, args, pdl);
  else
    launchKernelPdl(k_sort_pass<false>, parts, SORT_THREADS, sizeof(SortSmem), stream, args, pdl);
}

void launchHistogram(const uint32_t* keys, const uint32_t* countPtr, uint32_t maxCount, uint32_t* hist, int firstShift, int passes,
                     cudaStream_t stream)
{
  uint32_t blocks = (maxCount + HIST_THREADS * 16 - 1) / (HIST_THREADS * 16);
  blocks          = blocks < 1 ? 1 : (blocks > 148 * 4 ? 148 * 4 : blocks);
  if(passes == 4)
    k_histogram<4><<<blocks, HIST_THREADS, 0, stream>>>(keys, countPtr, maxCount, hist, firstShift);
  else if(passes == 2)
    k_histogram<2><<<blocks, HIST_THREADS, 0, stream>>>(keys, countPtr, maxCount, hist, firstShift);
  else
    k_histogram<1><<<blocks, HIST_THREADS, 0, stream>>>(keys, countPtr, maxCount, hist, firstShift);
}

void initSortKernels()
{
  cudaFuncSetAttribute(k_sort_pass<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(sizeof(SortSmem)));
  cudaFuncSetAttribute(k_sort_pass<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(sizeof(SortSmem)));
}

}  // namespace vkgs
