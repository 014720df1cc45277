// kernels.hpp — host-callable launchers of the sm_100a kernels (one translation unit per stage).
#pragma once
#include <cstdint>
#include <cuda_runtime.h>

#include "vkgs_b200.h"

namespace vkgs {

// ---- geometry of the per-frame pipeline ------------------------------------------------------
constexpr int      PRE_TILE        = 256;   // splats per preprocess tile (= block size)
constexpr int      RECORD_WORDS    = 12;    // per-splat record, 48 B
constexpr int      GUT_RECORD_WORDS = 24;   // 3DGUT pipeline: 96 B, see k_preprocess.cu
constexpr int      SORT_THREADS    = 256;
#ifndef VKGS_SORT_ITEMS
#define VKGS_SORT_ITEMS 16
#endif
constexpr int      SORT_ITEMS      = VKGS_SORT_ITEMS;  // keys per thread
constexpr int      SORT_PART       = SORT_THREADS * SORT_ITEMS;  // 4096 pairs per partition, four partitions resident per SM
constexpr int      BIN_THREADS     = 256;
constexpr int      TILE_W          = 32;
constexpr int      TILE_H          = 32;    // binning / tile-sort granularity: 32x32 pixels (fewest (tile, splat) pairs) ...
#ifndef VKGS_BLEND_H
#define VKGS_BLEND_H 16
#endif
constexpr int      BLEND_H         = VKGS_BLEND_H;  // ... blended by TILE_H / BLEND_H CTAs per tile, each a 32 x BLEND_H band reading the same list
constexpr int      BLEND_THREADS   = TILE_W * BLEND_H / 2;  // one warp per 8x8 pixel block of the band, two pixels per thread

// Small per-frame control block in HBM, cleared with one memset at the start of every frame.
struct FrameCounters
{
  // NOT cleared per frame (the per-frame memset starts at `visible`): set by any frame of the slot whose tile lists
  // overflowed since the host last looked (vkgs_sync), with the largest pair count wanted
  uint32_t stickyOverflow;
  uint32_t stickyPairs;
  uint32_t visible;            // V: splats that passed the dist-stage cull (IndirectParams.instanceCount)
  uint32_t tilePairs;          // D: (splat,tile) pairs the binning wanted to emit
  uint32_t tilePairsClamped;   // min(D, capacity): what the tile sort actually processes
  uint32_t overflow;           // set when D exceeded the tile-list capacity
  uint32_t sortSrc[4];         // [p]: which (key,id) buffer holds the output of depth-sort pass p (a pass
                               // whose digit is constant over all keys is skipped and does not flip it)
  unsigned long long fragments[2];  // profiling only (vkgs_options._reserved[0] & 128): list entries evaluated by a
                                    // warp block (x64 pixels), and fragments that passed both discards and were blended
  uint32_t ticket[12];         // dynamic tile / partition tickets, one per kernel launch
  uint32_t bigCount;           // huge splats handed from k_bin_emit to k_bin_big this frame
  uint32_t depthHist[4][256];  // digit histograms of the depth keys (filled by the preprocess kernel)
  uint32_t tileHist[2][256];   // digit histograms of the tile ids (filled by the binning kernel)
};

struct DeviceSplatSet
{
  const float* centers;  // 3 x f32, padded to PRE_TILE rows
  const float* cov6;     // 6 x f32
  const float* scales;   // 3 x f32 (log), size culling and surface-info normals
  const float* rotations;// 4 x f32 raw quaternion (w,x,y,z), surface-info normals only (else null)
  const void*  rgba;     // 4 x {f32,f16,u8}
  const void*  sh;       // 45 x {f32,f16,u8} or nullptr
  uint32_t     count;
  uint32_t     shDegree;
  uint32_t     shFormat;
  uint32_t     rgbaFormat;
};

struct PreprocessArgs
{
  DeviceSplatSet    set;
  vkgs_frame_params fp;
  vkgs_options      opt;
  float             mv[16];      // mul(transform, viewMatrix), evaluated once per frame on the host
  float             camModel[3]; // camera position in model space
  float             gutOrigin[3];// 3DGUT: ray origin (viewInverse translation) in model space
  uint32_t*         keys;        // [V] compacted, ascending splat id
  uint32_t*         ids;         // [V]
  uint32_t*         records;     // [N][RECORD_WORDS] (3DGUT pipeline: [N][GUT_RECORD_WORDS]), indexed by splat id
  uint2*            bboxes;      // [N] copy of the two pixel-bbox words of the record: the binning gathers 8 B per splat
                                 // by sorted id, and a compact array stays L2-resident where the 48-byte records do not
  FrameCounters*    counters;
  uint64_t*         status;      // look-back chain, one word per tile (of THIS launch)
  uint32_t          epoch;
  uint32_t          ticketSlot;
  // Multi-instance scenes (one launch per splat-set instance, in global-id order; replaces the
  // reference's global index table, src/splat_set_manager_vk.cpp:2304-2360): global id = idBase +
  // local id; tickets already drawn by earlier launches of the frame (tiles + CTAs of each); and whether the append
  // continues after the pairs of earlier instances (base = counters->visible).
  float4*           surface;     // [N] surface-info only (else null): world normal of the splat (xyz), NDC depth (w)
  uint32_t          idBase;
  uint32_t          ticketBase;
  uint32_t          chained;
  // Resident CTAs per SM of the persistent launch (0 = as many as fit). With several frames in flight
  // the context asks for ONE: the kernel then takes longer on its own but leaves two thirds of every
  // SM to the other frames' kernels, and frame throughput goes up (measured 3520 -> 3670 fps, config 2).
  uint32_t          ctasPerSm;
};

void launchPreprocess(const PreprocessArgs& args, cudaStream_t stream);
// CTAs of the (persistent) preprocess launch for these arguments; each CTA draws one ticket per tile
// it processes plus one final ticket, so a launch consumes tiles + grid tickets.
uint32_t preprocessGrid(const PreprocessArgs& args);

struct SortPassArgs
{
  // ping-pong buffers; the pass reads buffer `cur` and writes buffer `cur^1`, where cur = *srcSelIn
  // (0 when srcSelIn is null). With srcSelOut set, a pass whose digit histogram has a single
  // non-empty bin is skipped (it would be the identity) and *srcSelOut tells later kernels where
  // the data is.
  uint32_t*       keys[2];
  uint32_t*       vals[2];
  const uint32_t* srcSelIn;
  uint32_t*       srcSelOut;
  const uint32_t* countPtr;   // device-side number of pairs
  uint32_t        maxCount;   // host-side upper bound (sizes the grid)
  const uint32_t* histogram;  // 256 digit counts of this pass (not yet scanned)
  uint64_t*       status;     // [partitions][256] look-back words
  uint32_t*       ticket;
  uint32_t        epoch;
  int             shift;
  int             digitBits;  // number of significant bits of this pass's digit (0 or 8 = all eight)
  // Final pass of the tile sort only: every run of equal (full) keys in the output is a tile's list;
  // the scatter records its [begin,end) with atomicMin/atomicMax (arrays pre-set to 0xffffffff / 0).
  uint32_t*       rangeBegin;
  uint32_t*       rangeEnd;
};

// `pdl`: launch with the programmatic-stream-serialization attribute (the preceding work on the stream is a kernel of this
// library that calls pdl_launch_dependents / exits; see device_common.cuh)
void launchSortPass(const SortPassArgs& args, cudaStream_t stream, bool pdl = false);

// cudaLaunchKernelEx with (or without) the programmatic-stream-serialization attribute
template <typename Kernel, typename Args>
inline void launchKernelPdl(Kernel kernel, unsigned grid, unsigned block, size_t smem, cudaStream_t stream, const Args& args, bool pdl)
{
  cudaLaunchConfig_t cfg{};
  cfg.gridDim          = dim3(grid, 1, 1);
  cfg.blockDim         = dim3(block, 1, 1);
  cfg.dynamicSmemBytes = smem;
  cfg.stream           = stream;
  cudaLaunchAttribute attr[1];
  attr[0].id                                         = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = pdl ? 1 : 0;
  cfg.attrs                                          = attr;
  cfg.numAttrs                                       = 1;
  cudaLaunchKernelEx(&cfg, kernel, args);
}

// Digit histograms for `passes` 8-bit digits starting at bit `firstShift` (stand-alone sort only;
// the frame pipeline fuses its histograms into the producing kernels).
void launchHistogram(const uint32_t* keys, const uint32_t* countPtr, uint32_t maxCount, uint32_t* hist /*[passes][256]*/,
                     int firstShift, int passes, cudaStream_t stream);

struct BinArgs
{
  const uint32_t* sortedIds[2];  // [V] depth-sorted splat ids: buffer *sortedSel holds them
  const uint32_t* sortedSel;
  const uint2*    bboxes;      // [N] pixel bounding boxes (x0 | y0<<16, x1 | y1<<16), indexed by splat id
  FrameCounters*  counters;
  uint32_t*       tileKeys;    // [capacity]
  uint32_t*       tileVals;    // [capacity] splat id
  uint32_t        capacity;
  uint32_t        maxCount;    // upper bound of V
  uint32_t        tilesX, tilesY;
  uint64_t*       status;
  uint32_t        epoch;
  uint32_t        ticketSlot;
  uint32_t        debugFlags;  // profiling ablations (0 in production)
  uint4*          bigList;     // [bigCapacity] huge splats: (output offset, splat id, x0 | y0 << 16, nx | ny << 16)
  uint32_t        bigCapacity;
};

void launchBinEmit(const BinArgs& args, cudaStream_t stream, bool pdl = false);
void launchBinBig(const BinArgs& args, cudaStream_t stream, bool pdl = false);  // expands the huge splats k_bin_emit set aside


// per-frame constants of the VK3DGUT fragment stage (FrameInfo + SplatSetDesc fields it reads)
constexpr int GUT_MAX_INSTANCES = 8;  // splat-set instances of a 3DGUT scene (their inverse transforms travel as kernel arguments)

struct GutFrameConstants
{
  uint32_t enabled;
  uint32_t kernelDegree;
  uint32_t extentEigen;  // EXTENT_METHOD == EXTENT_EIGEN: quad = centre +- b1 +- b2 instead of an axis-aligned rectangle
  uint32_t fisheye;      // CAMERA_TYPE == CAMERA_FISHEYE: generateFisheyeRay instead of generatePinholeRay
  float    fovRad;
  float    viewInverse[16], projInverse[16], modelInverse[16];  // modelInverse: instance 0
  // multi-instance scenes: first global splat id of each instance and the upper-left 3x3 of its
  // transformInverse (row-vector convention, [3*i + j] = transformInverse[4*i + j])
  uint32_t instanceCount;
  uint32_t instanceOffset[GUT_MAX_INSTANCES];
  float    instanceInverse[GUT_MAX_INSTANCES][9];
  float    viewport[2];
  float    alphaClamp, kernelMinResponse, alphaCullThreshold;
};

struct BlendArgs
{
  const uint32_t* tileVals;  // tile-sorted splat ids
  const uint32_t* rangeBegin;  // [tiles] first entry of the tile's list in tileVals (0xffffffff = empty)
  const uint32_t* rangeEnd;    // [tiles] one past the last entry (0 = empty)
  const uint32_t* records;
  void*           image;     // [H][W] RGBA in targetFormat (float4 / half4 / uchar4)
  uint32_t        targetFormat;
  uint32_t        width, height, tilesX, tilesY;
  uint32_t        firstTile = 0, tileCount = 0;  // the launch covers tiles [firstTile, firstTile + tileCount) in row-major order (0 = all)
  uint32_t        frontToBack;
  uint32_t        disableOpacityGaussian;
  float           transmittanceEpsilon;
  unsigned long long* fragmentCounters;  // null in production: see FrameCounters::fragments
  // surface-info side outputs (front to back only; all null when options.surface_info is off)
  const float4*   surface;         // [N] per splat: world normal, NDC depth (written by the preprocess kernel)
  float4*         outNormals;      // [H][W] integrated normal * opacity, a = 1 - T
  float2*         outDepthT;       // [H][W] picked depth, transmittance
  uint32_t*       outSplatId;      // [H][W] id of the last blended fragment
  float           depthIsoThreshold;
  GutFrameConstants gut;           // gut.enabled = 0 for the 3DGS pipeline
};

void launchBlend(const BlendArgs& args, cudaStream_t stream);

}  // namespace vkgs
