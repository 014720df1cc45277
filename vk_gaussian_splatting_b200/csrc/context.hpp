// context.hpp — the context object behind the C ABI (shared by context.cu and sort_api.cu)
#pragma once
#include <string>
#include <vector>

#include "kernels.hpp"
#include "vkgs_b200.h"

namespace vkgs {

#ifndef VKGS_MAX_FIF
#define VKGS_MAX_FIF 4
#endif
constexpr int MAX_FRAMES_IN_FLIGHT = VKGS_MAX_FIF;

// Everything one in-flight frame owns. Two slots let frame N+1's front end (preprocess, sorts,
// binning — latency-bound kernels that leave most issue slots idle) overlap frame N's blend on the
// same GPU, like the frames-in-flight of the reference's swapchain loop.
struct FrameSlot
{
  cudaStream_t   stream = nullptr;       // front end (preprocess, sorts, binning): HIGH priority; frame completion is visible here
  cudaStream_t   streamBlend = nullptr;  // blend + copies to host: LOW priority, so another frame's latency-bound front end is
                                         // scheduled ahead of the remaining blend CTAs and fills the issue slots they leave idle
  cudaStream_t   streamCopy = nullptr;   // synchronous frames to host memory: the frame leaves in strips of tile rows, each copied
                                         // while the next one is blended
  static constexpr int COPY_STRIPS = 4;
  cudaEvent_t    evStrip[COPY_STRIPS] = {nullptr, nullptr, nullptr, nullptr};
  cudaEvent_t    evFront = nullptr, evBlend = nullptr;
  uint32_t *     dKeys[2] = {nullptr, nullptr}, *dIds[2] = {nullptr, nullptr};
  uint32_t*      dRecords    = nullptr;
  uint2*         dBboxes     = nullptr;
  uint4*         dBigList    = nullptr;  // huge splats of the frame (k_bin_emit -> k_bin_big)
  float4*        dSurface    = nullptr;  // surface-info only: per-splat world normal + NDC depth
  float4*        dOutNormals = nullptr;  // surface-info only: side outputs of the frame
  float2*        dOutDepthT  = nullptr;
  uint32_t*      dOutSplatId = nullptr;
  FrameCounters* dCounters   = nullptr;
  FrameCounters* hCounters   = nullptr;  // pinned; the first 48 bytes (through fragments[]) are copied back every frame
  uint64_t *     dPreStatus = nullptr, *dSortStatus = nullptr, *dBinStatus = nullptr, *dTileSortStatus = nullptr;
  uint32_t *     dTileKeys[2] = {nullptr, nullptr}, *dTileVals[2] = {nullptr, nullptr};
  uint64_t       tileCapacity = 0;
  uint32_t*      dRanges      = nullptr;  // [2][tiles]: list begin, list end
  void*          dImage       = nullptr;  // [H][W] RGBA in the target format (allocated for fp32, the largest)
  uint32_t       imgW = 0, imgH = 0;
  cudaEvent_t    ev[VKGS_K_COUNT + 1]{};
  cudaEvent_t    evDone     = nullptr;
  bool           evRecorded = false;
  bool           haveFrame  = false;
  vkgs_frame_params lastFp{};
  void*          lastHost      = nullptr;  // host destination of the slot's last frame (re-render after a tile-list overflow)
  bool           lastThin      = false;
  bool           lastPresorted = false;
  uint32_t       framesSinceSync = 0;        // frames enqueued on this slot since the host last checked for overflow
};

}  // namespace vkgs

struct vkgs_ctx
{
  int          device     = 0;
  cudaStream_t userStream = nullptr;  // optional: completion of every frame is made visible on it
  std::string  lastError;
  uint64_t     launches  = 0;
  uint32_t     epoch     = 0;
  bool         profiling = false;
  int          framesInFlight = vkgs::MAX_FRAMES_IN_FLIGHT < 4 ? vkgs::MAX_FRAMES_IN_FLIGHT : 4;
  int          nextSlot  = 0;
  int          lastSlot  = -1;  // slot of the most recently enqueued frame

  // scene (shared by all slots): splat sets in HBM + the instances that place them in the world.
  // Global splat id = instance.globalOffset + local id, instances in creation order — the layout of
  // the reference's global index table (SplatSetManagerVk::rebuildGlobalIndexTables,
  // src/splat_set_manager_vk.cpp:2304-2360), here implicit: one preprocess launch per instance.
  struct SetStorage
  {
    vkgs::DeviceSplatSet view{};
    void *               dCenters = nullptr, *dCov = nullptr, *dScales = nullptr, *dRgba = nullptr, *dSh = nullptr, *dRotations = nullptr;
  };
  struct Instance
  {
    uint32_t setIndex     = 0;
    uint32_t globalOffset = 0;
    uint32_t tileOffset   = 0;     // first preprocess tile (ticket / look-back status index) of the instance
    bool     frameModel   = false; // transform comes from vkgs_frame_params.model (single-set vkgs_upload)
    float    transform[16]{}, transformInverse[16]{};
  };
  bool                  uploaded = false;
  vkgs_options          opt{};
  std::vector<SetStorage> sets;
  std::vector<Instance>   instances;
  uint32_t              totalSplats = 0;  // sum over instances
  uint32_t              totalTiles  = 0;  // preprocess tiles over all instances

  vkgs::FrameSlot slots[vkgs::MAX_FRAMES_IN_FLIGHT];

  // image comparison: the captured frame (ImageCompare's capture image, src/image_compare.cpp)
  void*    dCapture = nullptr;
  uint32_t captureW = 0, captureH = 0;
};

#define CU_TRY(ctx, expr)                                                                                                      \
  do                                                                                                                           \
  {                                                                                                                            \
    cudaError_t e_ = (expr);                                                                                                   \
    if(e_ != cudaSuccess)                                                                                                      \
    {                                                                                                                          \
      (ctx)->lastError = std::string(#expr) + ": " + cudaGetErrorString(e_);                                                   \
      return VKGS_ERR_CUDA;                                                                                                    \
    }                                                                                                                          \
  } while(0)

namespace vkgs {
template <typename T>
inline void freeDev(T*& p)
{
  if(p)
    cudaFree(p);
  p = nullptr;
}
}  // namespace vkgs
