// k_radix_sort.cu — stable LSD radix sort of (u32 key, u32 value) pairs, 8 bits per pass.
//
// Replaces the vrdx key-value sort the reference records each frame
// (vrdxCmdSortKeyValueIndirect, 3rdparty/vrdx/src/vk_radix_sort.cc:249-258,262-416; shaders
// upsweep/spine/downsweep.slang). Same contract: ascending, stable, pair count read on the device.
// vrdx is reduce-then-scan (3 dispatches and ~20 B/pair per pass). This is a single-pass
// ("onesweep") design instead: the digit histograms of all passes are produced by whichever
// kernel wrote the keys, and each pass is ONE kernel that ranks an 8192-pair partition in shared
// memory, resolves its global digit offsets with a decoupled look-back over earlier partitions
// (epoch-stamped 64-bit status words: nothing is cleared between passes or frames) and scatters
// through shared memory so global writes leave in digit-contiguous runs. Traffic per pass is one
// read + one write of the pairs (16 B/pair), i.e. 64 B/pair for 32-bit keys.
//
// Stability: a partition is ranked in (warp, item, lane) order, which is the input order because
// each warp loads a contiguous chunk in a warp-striped arrangement; partitions are ordered by a
// ticket taken at block start, which also guarantees the look-back never waits on a block that
// has not been scheduled.
#include "device_common.cuh"
#include "kernels.hpp"

namespace vkgs {

namespace {

constexpr int NWARPS = SORT_THREADS / 32;

struct SortSmem
{
  union
  {
    uint32_t warpHist[NWARPS][256];  // per-warp digit counters, then per-warp exclusive offsets
    uint32_t stageKeys[SORT_PART];   // (re-used after ranking) block-sorted keys
  };
  uint32_t stageVals[SORT_PART];
  uint32_t digitStart[256];   // first block-local position of each digit
  uint32_t globalBase[256];   // global position of the first key of each digit of this partition
  uint32_t scan[NWARPS + 1];
  uint32_t part;
};
static_assert(sizeof(uint32_t) * NWARPS * 256 <= sizeof(uint32_t) * SORT_PART, "warpHist must fit in the key staging area");

template <int BITS>
__global__ void __launch_bounds__(SORT_THREADS, 2) k_sort_pass(const __grid_constant__ SortPassArgs a)
{
  extern __shared__ __align__(16) unsigned char smemRaw[];
  SortSmem&      sm   = *reinterpret_cast<SortSmem*>(smemRaw);
  const unsigned tid  = threadIdx.x;
  const unsigned lane = tid & 31u, warp = tid >> 5;

#ifdef VKGS_TIMELINE
  const long long tl0 = clock64();
#endif
  const uint32_t count = *a.countPtr;
  const uint32_t parts = (count + SORT_PART - 1) / SORT_PART;
  uint32_t       cur   = a.srcSelIn ? *a.srcSelIn : 0u;
  if(a.srcSelOut)
  {
    // identity pass (every key has the same digit, e.g. the exponent byte of NDC depths): skip it
    const int same = __syncthreads_or(count == 0u || a.histogram[tid & 255u] == count);
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
  if(tid == 0)
    sm.part = atomicAdd(a.ticket, 1u);
  for(int i = tid; i < NWARPS * 256; i += SORT_THREADS)
    (&sm.warpHist[0][0])[i] = 0u;
  if(tid < 256)
    sm.digitStart[tid] = 0u;  // (first used as the partition's digit counters, see below)
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
  uint32_t       keys[SORT_ITEMS];
#pragma unroll
  for(int i = 0; i < SORT_ITEMS; i++)
  {
    const uint32_t idx = warpBase + i * 32 + lane;
    keys[i]            = idx < count ? keysIn[idx] : 0xffffffffu;  // padding sorts last, is never written back
  }
  VKGS_TL(part, 2);

  // ---- count the partition's digits FIRST (shared atomics, a few hundred cycles) and publish them:
  // successors' look-backs only need these counts, so they must not wait for the ranking below
  // (several thousand cycles, twice that on an SM shared by two partitions) ------------------------
#pragma unroll
  for(int i = 0; i < SORT_ITEMS; i++)
    atomicAdd(&sm.digitStart[(keys[i] >> a.shift) & 0xffu], 1u);
  __syncthreads();
  uint32_t  realCount = 0;
  uint64_t* mine      = a.status + static_cast<uint64_t>(part) * 256 + (tid & 255u);
  if(tid < 256)
  {
    // padding keys (digit 0xff of the last partition) are not counted
    realCount = sm.digitStart[tid];
    if(tid == 255 && partBase + SORT_PART > count)
      realCount -= (partBase + SORT_PART - count);
    if(part == 0)
    {
      // exclusive scan of the global digit histogram = first global position of each digit.
      // (only partition 0 needs it; everyone else inherits it through the look-back chain)
      const uint32_t h   = a.histogram[tid];
      uint32_t       inc = warp_inclusive_scan(h, lane);
      __shared__ uint32_t warpTot[8];
      if(lane == 31)
        warpTot[warp] = inc;
      asm volatile("bar.sync 1, 256;");
      uint32_t add = 0;
      for(unsigned w = 0; w < warp; w++)
        add += warpTot[w];
      const uint32_t excl = inc - h + add;
      lb_store(mine, lb_pack(a.epoch, LB_INCLUSIVE, excl + realCount));
      sm.globalBase[tid] = excl;
      if(tid == 0 && a.srcSelOut)
        *a.srcSelOut = cur ^ 1u;
    }
    else
      lb_store(mine, lb_pack(a.epoch, LB_AGGREGATE, realCount));
  }

  // ---- rank within the warp (ballot multi-split) ------------------------------------------------
  // peer masks first (see match_digit), then the dependent counter updates
  uint32_t rank[SORT_ITEMS];
  {
    unsigned peers[SORT_ITEMS];
#pragma unroll
    for(int i = 0; i < SORT_ITEMS; i++)
    {
      const uint32_t digit = (keys[i] >> a.shift) & 0xffu;
      peers[i]             = match_digit<BITS>(FULL_MASK, digit);
      if(BITS < 8)
      {
        // real digits are < 2^BITS; the padding digit 0xff is told apart by its top bit
        const bool     top = (digit >> 7) & 1u;
        const unsigned v   = __ballot_sync(FULL_MASK, top);
        peers[i] &= top ? v : ~v;
      }
    }
#pragma unroll
    for(int i = 0; i < SORT_ITEMS; i++)
    {
      const uint32_t digit  = (keys[i] >> a.shift) & 0xffu;
      const unsigned leader = __ffs(peers[i]) - 1;
      uint32_t       before = 0;
      if(lane == leader)
      {
        before                   = sm.warpHist[warp][digit];
        sm.warpHist[warp][digit] = before + __popc(peers[i]);
      }
      before  = __shfl_sync(FULL_MASK, before, leader);
      rank[i] = before + __popc(peers[i] & ((1u << lane) - 1u));
      __syncwarp();
    }
  }
  // values are only needed for staging: issue their loads now so the latency hides behind the scans
  uint32_t vals[SORT_ITEMS];
#pragma unroll
  for(int i = 0; i < SORT_ITEMS; i++)
  {
    const uint32_t idx = warpBase + i * 32 + lane;
    vals[i]            = idx < count ? valsIn[idx] : 0u;
  }
  __syncthreads();
  VKGS_TL(part, 3);

  // ---- per-digit: exclusive scan over warps, block totals ------------------------------------------
  uint32_t digitCount = 0;
  if(tid < 256)
  {
    uint32_t run = 0;
#pragma unroll
    for(int w = 0; w < NWARPS; w++)
    {
      const uint32_t c    = sm.warpHist[w][tid];
      sm.warpHist[w][tid] = run;
      run += c;
    }
    digitCount = run;
  }
  // exclusive scan of the 256 digit totals -> block-local start of each digit
  {
    uint32_t       total;
    const uint32_t excl = block_exclusive_scan<NWARPS>(tid < 256 ? digitCount : 0u, sm.scan, total);
    if(tid < 256)
      sm.digitStart[tid] = excl;
  }
  __syncthreads();
  VKGS_TL(part, 4);

  // ---- block-local positions; stage pairs in sorted order --------------------------------------
  uint32_t pos[SORT_ITEMS];
#pragma unroll
  for(int i = 0; i < SORT_ITEMS; i++)
  {
    const uint32_t digit = (keys[i] >> a.shift) & 0xffu;
    pos[i]               = sm.digitStart[digit] + sm.warpHist[warp][digit] + rank[i];
  }
  __syncthreads();  // warpHist is dead from here: its storage becomes stageKeys
#pragma unroll
  for(int i = 0; i < SORT_ITEMS; i++)
  {
    sm.stageKeys[pos[i]] = keys[i];
    sm.stageVals[pos[i]] = vals[i];
  }
  VKGS_TL(part, 5);

  // ---- global digit offsets: decoupled look-back, one chain per digit --------------------------
  if(tid < 256 && part != 0)
  {
    const uint32_t excl = lb_lookback<16>(a.status + tid, part, 256, a.epoch);
    lb_store(mine, lb_pack(a.epoch, LB_INCLUSIVE, excl + realCount));
    sm.globalBase[tid] = excl;
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
      const uint32_t k     = sm.stageKeys[j];
      const uint32_t digit = (k >> a.shift) & 0xffu;
      const uint32_t dst   = sm.globalBase[digit] + (j - sm.digitStart[digit]);
      keysOut[dst]         = k;
      valsOut[dst]         = sm.stageVals[j];
      if(a.rangeBegin)
      {
        // equal keys are contiguous in the staged (block-sorted) order: the input of the last pass is
        // ordered by the low digit, so inside every high-digit run the full key is non-decreasing
        if(j == 0 || sm.stageKeys[j - 1] != k)
          atomicMin(a.rangeBegin + k, dst);
        if(j + 1 == valid || sm.stageKeys[j + 1] != k)
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
__global__ void __launch_bounds__(HIST_THREADS) k_histogram(const uint32_t* __restrict__ keys, const uint32_t* countPtr,
                                                            uint32_t* hist, int firstShift)
{
  __shared__ uint32_t s[HIST_COPIES][PASSES][256];
  const unsigned      tid = threadIdx.x;
  for(int i = tid; i < HIST_COPIES * PASSES * 256; i += HIST_THREADS)
    (&s[0][0][0])[i] = 0u;
  __syncthreads();
  const uint32_t count = *countPtr;
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

void launchSortPass(const SortPassArgs& args, cudaStream_t stream)
{
  const uint32_t parts = (args.maxCount + SORT_PART - 1) / SORT_PART;
  if(parts == 0)
    return;
  if(args.digitBits > 0 && args.digitBits <= 5)
    k_sort_pass<5><<<parts, SORT_THREADS, sizeof(SortSmem), stream>>>(args);
  else
    k_sort_pass<8><<<parts, SORT_THREADS, sizeof(SortSmem), stream>>>(args);
}

void launchHistogram(const uint32_t* keys, const uint32_t* countPtr, uint32_t maxCount, uint32_t* hist, int firstShift, int passes,
                     cudaStream_t stream)
{
  uint32_t blocks = (maxCount + HIST_THREADS * 16 - 1) / (HIST_THREADS * 16);
  blocks          = blocks < 1 ? 1 : (blocks > 148 * 4 ? 148 * 4 : blocks);
  if(passes == 4)
    k_histogram<4><<<blocks, HIST_THREADS, 0, stream>>>(keys, countPtr, hist, firstShift);
  else if(passes == 2)
    k_histogram<2><<<blocks, HIST_THREADS, 0, stream>>>(keys, countPtr, hist, firstShift);
  else
    k_histogram<1><<<blocks, HIST_THREADS, 0, stream>>>(keys, countPtr, hist, firstShift);
}

void initSortKernels()
{
  cudaFuncSetAttribute(k_sort_pass<8>, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(sizeof(SortSmem)));
  cudaFuncSetAttribute(k_sort_pass<5>, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(sizeof(SortSmem)));
}

}  // namespace vkgs
