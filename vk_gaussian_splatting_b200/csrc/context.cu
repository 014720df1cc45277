// context.cu — the C ABI (include/vkgs_b200.h): context lifetime, scene upload, frame orchestration.
//
// One frame = the reference's processSortingOnGPU + drawSplatPrimitives
// (src/gaussian_splatting.cpp:1298-1367, 1369-1465) as a fixed sequence of stream-ordered launches
// with every data-dependent size (V, tile-pair count) read on the device — no host round trip
// inside a frame:
//   memset(control block) -> preprocess (one launch per splat-set instance) -> <=4 x sort pass ->
//   bin emit -> bin big -> 2 x tile sort pass (the second also records the per-tile list ranges) -> blend
// Up to four frames are in flight (FrameSlot), each with a high-priority stream for its front end and
// a low-priority one for the blend + copy to host, so a frame's latency-bound front end runs beside
// the previous frames' blends.
#include <algorithm>
#include <cmath>
#include <cstddef>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <vector>

#include "context.hpp"
#include "host_pack.hpp"

namespace vkgs {
void initSortKernels();
void initPreprocessKernels();
void initBlendKernels();
}  // namespace vkgs

using namespace vkgs;

namespace {

constexpr uint32_t BIG_LIST_CAPACITY = 1u << 16;  // huge splats per frame handed to k_bin_big (beyond: expanded in place)

int fail(vkgs_ctx* ctx, int code, const char* msg)
{
  if(ctx)
    ctx->lastError = msg;
  return code;
}

// cudaMemset / cudaMemcpy from pageable memory run on the legacy default stream and may return before the device has finished
// them; the frame streams are non-blocking (not ordered behind the legacy stream), so every group of such calls is closed
// with this barrier before anything on a frame stream can touch the memory. (All callers are rare: upload, regrowth, repair.)
void legacyStreamBarrier()
{
  cudaStreamSynchronize(cudaStreamLegacy);
}

void freeSlotScene(FrameSlot& s)
{
  for(int i = 0; i < 2; i++)
    freeDev(s.dKeys[i]), freeDev(s.dIds[i]), freeDev(s.dTileKeys[i]), freeDev(s.dTileVals[i]);
  freeDev(s.dRecords), freeDev(s.dBboxes), freeDev(s.dBigList), freeDev(s.dSurface), freeDev(s.dPreStatus), freeDev(s.dSortStatus), freeDev(s.dBinStatus), freeDev(s.dTileSortStatus);
  s.tileCapacity = 0;
  s.haveFrame    = false;
  // nothing of the old scene is pending any more: no stale overflow flag may trigger a repair of a frame that is gone
  s.framesSinceSync = 0;
  if(s.hCounters)
    s.hCounters->overflow = s.hCounters->stickyOverflow = s.hCounters->stickyPairs = 0;
  if(s.dCounters)
  {
    cudaMemset(s.dCounters, 0, offsetof(FrameCounters, visible));
    legacyStreamBarrier();
  }
}

void freeScene(vkgs_ctx* c)
{
  for(auto& st : c->sets)
    freeDev(st.dCenters), freeDev(st.dCov), freeDev(st.dScales), freeDev(st.dRgba), freeDev(st.dSh), freeDev(st.dRotations);
  c->sets.clear();
  c->instances.clear();
  c->totalSplats = c->totalTiles = 0;
  for(auto& s : c->slots)
    freeSlotScene(s);
  c->uploaded = false;
  c->lastSlot = -1;
}

int allocTileLists(vkgs_ctx* c, FrameSlot& s, uint64_t capacity)
{
  capacity = std::min<uint64_t>(capacity, 0xfffff000ull);
  // allocate the new buffers first: if any cudaMalloc fails (the regrow after an overflow is the likely moment
  // for an out-of-memory), the slot keeps its old, consistent lists and capacity
  uint32_t *nk[2] = {nullptr, nullptr}, *nv[2] = {nullptr, nullptr};
  uint64_t* nst   = nullptr;
  const uint64_t parts = (capacity + SORT_PART - 1) / SORT_PART;
  cudaError_t    e     = cudaSuccess;
  for(int i = 0; i < 2 && e == cudaSuccess; i++)
  {
    e = cudaMalloc(&nk[i], capacity * sizeof(uint32_t));
    if(e == cudaSuccess)
      e = cudaMalloc(&nv[i], capacity * sizeof(uint32_t));
  }
  if(e == cudaSuccess)
    e = cudaMalloc(&nst, parts * 256 * sizeof(uint64_t));
  if(e == cudaSuccess)
    e = cudaMemset(nst, 0, parts * 256 * sizeof(uint64_t));
  if(e == cudaSuccess)
    e = cudaStreamSynchronize(cudaStreamLegacy);
  if(e != cudaSuccess)
  {
    for(int i = 0; i < 2; i++)
      freeDev(nk[i]), freeDev(nv[i]);
    freeDev(nst);
    c->lastError = std::string("tile-list allocation: ") + cudaGetErrorString(e);
    cudaGetLastError();  // (clear the sticky-free error state of the failed cudaMalloc)
    return e == cudaErrorMemoryAllocation ? VKGS_ERR_OUT_OF_MEMORY : VKGS_ERR_CUDA;
  }
  for(int i = 0; i < 2; i++)
  {
    freeDev(s.dTileKeys[i]), freeDev(s.dTileVals[i]);
    s.dTileKeys[i] = nk[i], s.dTileVals[i] = nv[i];
  }
  freeDev(s.dTileSortStatus);
  s.dTileSortStatus = nst;
  s.tileCapacity    = capacity;
  return VKGS_OK;
}

int allocSlotScene(vkgs_ctx* c, FrameSlot& s, uint64_t n, uint64_t preTiles)
{
  for(int i = 0; i < 2; i++)
  {
    CU_TRY(c, cudaMalloc(&s.dKeys[i], n * sizeof(uint32_t)));
    CU_TRY(c, cudaMalloc(&s.dIds[i], n * sizeof(uint32_t)));
  }
  const uint64_t recordWords = c->opt.pipeline == VKGS_PIPELINE_3DGUT ? GUT_RECORD_WORDS : RECORD_WORDS;
  CU_TRY(c, cudaMalloc(&s.dRecords, n * recordWords * sizeof(uint32_t)));
  CU_TRY(c, cudaMalloc(&s.dBboxes, n * sizeof(uint2)));
  CU_TRY(c, cudaMalloc(&s.dBigList, BIG_LIST_CAPACITY * sizeof(uint4)));
  if(c->opt.surface_info)
    CU_TRY(c, cudaMalloc(&s.dSurface, n * sizeof(float4)));
  const uint64_t binParts = (n + 255) / 256, sortParts = (n + SORT_PART - 1) / SORT_PART;
  CU_TRY(c, cudaMalloc(&s.dPreStatus, preTiles * sizeof(uint64_t)));
  CU_TRY(c, cudaMalloc(&s.dBinStatus, binParts * sizeof(uint64_t)));
  CU_TRY(c, cudaMalloc(&s.dSortStatus, sortParts * 256 * sizeof(uint64_t)));
  CU_TRY(c, cudaMemset(s.dPreStatus, 0, preTiles * sizeof(uint64_t)));
  CU_TRY(c, cudaMemset(s.dBinStatus, 0, binParts * sizeof(uint64_t)));
  CU_TRY(c, cudaMemset(s.dSortStatus, 0, sortParts * 256 * sizeof(uint64_t)));
  legacyStreamBarrier();
  return allocTileLists(c, s, std::max<uint64_t>(8 * n, 1u << 20));
}

int ensureTargets(vkgs_ctx* c, FrameSlot& s, uint32_t w, uint32_t h)
{
  if(w == 0 || h == 0 || w > 65535 || h > 65535)
    return fail(c, VKGS_ERR_INVALID_ARGUMENT, "viewport must be within 1..65535 pixels per side");
  const uint32_t tx = (w + TILE_W - 1) / TILE_W, ty = (h + TILE_H - 1) / TILE_H;
  if(tx * ty > 65536)
    return fail(c, VKGS_ERR_UNSUPPORTED, "more than 65536 tiles (tile ids are sorted on 16 bits)");
  if(s.imgW != w || s.imgH != h || (c->opt.surface_info && !s.dOutNormals))
  {
    CU_TRY(c, cudaStreamSynchronize(s.stream));
    freeDev(s.dImage);
    freeDev(s.dRanges);
    freeDev(s.dOutNormals), freeDev(s.dOutDepthT), freeDev(s.dOutSplatId);
    if(c->opt.surface_info)
    {
      CU_TRY(c, cudaMalloc(&s.dOutNormals, sizeof(float4) * static_cast<size_t>(w) * h));
      CU_TRY(c, cudaMalloc(&s.dOutDepthT, sizeof(float2) * static_cast<size_t>(w) * h));
      CU_TRY(c, cudaMalloc(&s.dOutSplatId, sizeof(uint32_t) * static_cast<size_t>(w) * h));
    }
    CU_TRY(c, cudaMalloc(&s.dImage, sizeof(float4) * static_cast<size_t>(w) * h));
    CU_TRY(c, cudaMalloc(&s.dRanges, 2 * sizeof(uint32_t) * tx * ty));
    s.imgW = w, s.imgH = h;
  }
  return VKGS_OK;
}

uint32_t nextEpoch(vkgs_ctx* c)
{
  c->epoch++;
  if(c->epoch >= (1u << 30))
  {
    // epoch space exhausted (2^30 launches): clear the status arrays once and restart
    const uint64_t n = c->totalSplats;
    for(auto& s : c->slots)
    {
      cudaStreamSynchronize(s.stream);
      if(!s.dPreStatus)
        continue;
      cudaMemset(s.dPreStatus, 0, c->totalTiles * sizeof(uint64_t));
      cudaMemset(s.dBinStatus, 0, ((n + 255) / 256) * sizeof(uint64_t));
      cudaMemset(s.dSortStatus, 0, ((n + SORT_PART - 1) / SORT_PART) * 256 * sizeof(uint64_t));
      cudaMemset(s.dTileSortStatus, 0, ((s.tileCapacity + SORT_PART - 1) / SORT_PART) * 256 * sizeof(uint64_t));
    }
    legacyStreamBarrier();
    c->epoch = 1;
  }
  return c->epoch;
}

// host-side per-frame constants, evaluated in the oracle's operation order
void frameConstants(const vkgs_frame_params& fp, float mv[16], float camModel[3], float gutOrigin[3])
{
  // 3DGUT ray origin: mul(float4(0,0,0,1), viewInverse).xyz = viewInverse[3].xyz, then into model space
  // (cameras.h.slang:40, threedgut_raster.frag.slang:117)
  {
    const float o[4] = {fp.view_inverse[12], fp.view_inverse[13], fp.view_inverse[14], 1.0f};
    for(int j = 0; j < 3; j++)
      gutOrigin[j] = ((o[0] * fp.model_inverse[0 + j] + o[1] * fp.model_inverse[4 + j]) + o[2] * fp.model_inverse[8 + j])
                     + o[3] * fp.model_inverse[12 + j];
  }
  // (fp.model / fp.model_inverse hold the transform of the instance being launched)
  for(int i = 0; i < 4; i++)
    for(int j = 0; j < 4; j++)
      mv[4 * i + j] = ((fp.model[4 * i + 0] * fp.view[0 + j] + fp.model[4 * i + 1] * fp.view[4 + j]) + fp.model[4 * i + 2] * fp.view[8 + j])
                      + fp.model[4 * i + 3] * fp.view[12 + j];
  const float cp[4] = {fp.camera_position[0], fp.camera_position[1], fp.camera_position[2], 1.0f};
  for(int j = 0; j < 3; j++)
    camModel[j] = ((cp[0] * fp.model_inverse[0 + j] + cp[1] * fp.model_inverse[4 + j]) + cp[2] * fp.model_inverse[8 + j])
                  + cp[3] * fp.model_inverse[12 + j];
}

// What a frame was asked to do besides rendering into the slot's framebuffer (kept per slot so that a frame whose
// tile lists overflowed can be enqueued again by vkgs_sync without the caller's help).
struct FrameRequest
{
  void*           hostRgba       = nullptr;  // pinned host destination of the finished frame (or null)
  bool            throughputMode = false;    // thin co-running front end (asynchronous API with several frames in flight)
  int             forceSlot      = -1;       // re-render: the slot the frame was first enqueued on
  const uint32_t* presortedIds   = nullptr;  // CPU-sorting mode: HOST ids to draw in this order (no dist cull, no sort)
  uint32_t        presortedCount = 0;
};

int repairSlotBeforeReuse(vkgs_ctx* c, int si);

// Enqueue one frame on the next slot; optionally a device->host copy of the finished frame.
int enqueueFrame(vkgs_ctx* c, const vkgs_frame_params& fp, const FrameRequest& rq, int* slotOut)
{
  void* const hostRgba       = rq.hostRgba;
  bool        throughputMode = rq.throughputMode;
  const bool  presorted      = rq.presortedIds != nullptr;
  if(!c->uploaded)
    return fail(c, VKGS_ERR_NOT_UPLOADED, "vkgs_render before vkgs_upload");
  if(fp.width == 0 || fp.height == 0)
    return fail(c, VKGS_ERR_INVALID_ARGUMENT, "zero-sized viewport");
  // The quad offsets of the reference are scaled by basisViewport * 2 * inverseFocalAdjustment
  // (threedgs_raster.mesh.slang:284, threedgut_raster.mesh.slang:248); the kernels are built for the values
  // updateAndUploadFrameInfoUBO writes for a perspective camera at devicePixelRatio 1: (1/W, 1/H) and 1.
  if(fp.inverse_focal_adjustment != 1.0f || fp.basis_viewport[0] != 1.0f / fp.viewport[0] || fp.basis_viewport[1] != 1.0f / fp.viewport[1])
    return fail(c, VKGS_ERR_UNSUPPORTED, "inverse_focal_adjustment must be 1 and basis_viewport (1/W, 1/H) (orthographic focal adjustment / devicePixelRatio != 1 are outside the built path)");
  CU_TRY(c, cudaSetDevice(c->device));
  const int  si = rq.forceSlot >= 0 ? rq.forceSlot : c->nextSlot;
  FrameSlot& s  = c->slots[si];
  // Reusing a slot: the caller is `frames in flight` frames ahead of the frame that still lives here. Wait for it (the
  // back-pressure of the asynchronous API) and, should its tile lists have overflowed, repair it now — grow the lists,
  // render it again into the same destination — so that no frame is ever lost to a reused slot.
  static const bool reuseCheck = []() { const char* e = getenv("VKGS_NO_REUSE_CHECK"); return !(e && *e == '1'); }();  // (A/B measurements)
  if(reuseCheck && rq.forceSlot < 0 && s.haveFrame && s.framesSinceSync > 0)
    if(int rc = repairSlotBeforeReuse(c, si))
      return rc;
  // A frame of the asynchronous path that finds the device idle (the previous frame has already left it: the first frame
  // of a burst, or a caller slower than the GPU) has nothing to share the SMs with: it takes the full-width, latency-
  // oriented launches of the synchronous path. Same bits either way (test_four_frames_in_flight_...).
  static const bool idleFull = []() { const char* e = getenv("VKGS_NO_IDLE_FULL"); return !(e && *e == '1'); }();
  if(throughputMode && idleFull && rq.forceSlot < 0 && !c->profiling)
  {
    if(c->lastSlot < 0 || !c->slots[c->lastSlot].haveFrame || cudaEventQuery(c->slots[c->lastSlot].evBlend) == cudaSuccess)
      throughputMode = false;
    else
      (void)cudaGetLastError();  // cudaErrorNotReady is not an error
  }
  if(int rc = ensureTargets(c, s, fp.width, fp.height))
    return rc;
  const uint32_t n  = c->totalSplats;
  const uint32_t tx = (fp.width + TILE_W - 1) / TILE_W, ty = (fp.height + TILE_H - 1) / TILE_H;
  cudaStream_t   st = s.stream;
  auto           mark = [&](int slot) {
    if(c->profiling)
      cudaEventRecord(s.ev[slot], st);
  };
  // Programmatic dependent launch along the front-end chain (kernel -> kernel on one stream), for the synchronous call
  // only: one frame alone is a chain of latency-bound launches and the pre-launched prologues take 17 us off it (cfg2:
  // 357 -> 340 us); with several frames in flight the pre-launched CTAs hold registers and shared memory the other
  // frames' kernels could use (measured 4720 -> 4510 frames/s). Off while per-kernel events are recorded between the
  // launches (the timings would overlap), and by VKGS_NO_PDL=1 (A/B measurements).
  static const bool pdlEnv = []() { const char* e = getenv("VKGS_NO_PDL"); return !(e && *e == '1'); }();
  const bool        pdl    = pdlEnv && !c->profiling && !throughputMode;
  // VKGS_NO_STRIPS=1: a synchronous frame to host memory is blended and copied in one piece (A/B measurements)
  static const bool stripsOn = []() { const char* e = getenv("VKGS_NO_STRIPS"); return !(e && *e == '1'); }();

  CU_TRY(c, cudaMemsetAsync(&s.dCounters->visible, 0, sizeof(FrameCounters) - offsetof(FrameCounters, visible), st));
  // tile list ranges: begin = 0xffffffff, end = 0 (two arrays, two byte-pattern memsets)
  CU_TRY(c, cudaMemsetAsync(s.dRanges, 0xff, sizeof(uint32_t) * tx * ty, st));
  CU_TRY(c, cudaMemsetAsync(s.dRanges + tx * ty, 0, sizeof(uint32_t) * tx * ty, st));
  mark(0);

  // ---- "GPU Dist" (+ the per-splat half of "Rasterization", fused) -----------------------------
  // One launch per splat-set instance, in global-id order (dist.comp.slang:53 resolves the global id
  // through the global index table; here the table is implicit in the launch sequence). The launches
  // chain their deterministic append through counters->visible.
  uint32_t ticketsDrawn = 0;
  for(size_t k = 0; k < c->instances.size(); k++)
  {
    const vkgs_ctx::Instance& inst = c->instances[k];
    PreprocessArgs            pa{};
    pa.set = c->sets[inst.setIndex].view;
    pa.fp  = fp;
    if(!inst.frameModel)
    {
      std::memcpy(pa.fp.model, inst.transform, sizeof(pa.fp.model));
      std::memcpy(pa.fp.model_inverse, inst.transformInverse, sizeof(pa.fp.model_inverse));
    }
    pa.opt = c->opt;
    if(presorted)
    {
      // CPU-sorting mode: no dist shader, so no dist-stage cull — every splat of the scene gets its record, the frustum
      // test moves to the raster stage and size culling is off (src/gaussian_splatting_ui.cpp:1468-1490)
      if(pa.opt.frustum_culling_mode == VKGS_FRUSTUM_CULLING_AT_DIST)
        pa.opt.frustum_culling_mode = VKGS_FRUSTUM_CULLING_AT_RASTER;
      pa.opt.size_culling_mode = VKGS_SIZE_CULLING_DISABLED;
    }
    frameConstants(pa.fp, pa.mv, pa.camModel, pa.gutOrigin);
    pa.keys       = s.dKeys[0];
    pa.ids        = s.dIds[0];
    pa.records    = s.dRecords;
    pa.bboxes     = s.dBboxes;
    pa.surface    = c->opt.surface_info ? s.dSurface : nullptr;
    pa.counters   = s.dCounters;
    pa.status     = s.dPreStatus + inst.tileOffset;
    pa.epoch      = nextEpoch(c);
    pa.ticketSlot = 0;
    pa.idBase     = inst.globalOffset;
    pa.ticketBase = ticketsDrawn;
    pa.chained    = k > 0;
    pa.ctasPerSm  = (throughputMode && c->framesInFlight > 1) ? 1u : 0u;  // thin launch that co-runs with other frames
    ticketsDrawn += (pa.set.count + PRE_TILE - 1) / PRE_TILE + preprocessGrid(pa);
    launchPreprocess(pa, st);
    c->launches++;
  }
  mark(VKGS_K_PREPROCESS + 1);
  mark(VKGS_K_SORT_SCAN + 1);  // (depth-key digit histograms are fused into the preprocess kernel)

  // ---- "GPU Sort": up to 4 x 8-bit stable passes over (key,id) -----------------------------------
  if(presorted)
  {
    // the caller's order replaces the sort (tryConsumeAndUploadCpuSortingResult: memcpy + vkCmdCopyBuffer of 4N bytes,
    // src/splat_set_manager_vk.cpp:3396-3414); V = the number of ids handed in. sortSrc[3] stays 0: buffer 0.
    CU_TRY(c, cudaMemcpyAsync(s.dIds[0], rq.presortedIds, sizeof(uint32_t) * rq.presortedCount, cudaMemcpyHostToDevice, st));
    CU_TRY(c, cudaMemcpyAsync(&s.dCounters->visible, &rq.presortedCount, sizeof(uint32_t), cudaMemcpyHostToDevice, st));
    for(int p = 0; p < 4; p++)
      mark(VKGS_K_SORT_PASS0 + p + 1);
  }
  for(int p = 0; p < 4 && !presorted; p++)
  {
    SortPassArgs sa{};
    sa.keys[0] = s.dKeys[0], sa.keys[1] = s.dKeys[1];
    sa.vals[0] = s.dIds[0], sa.vals[1] = s.dIds[1];
    sa.srcSelIn  = p ? &s.dCounters->sortSrc[p - 1] : nullptr;
    sa.srcSelOut = &s.dCounters->sortSrc[p];
    sa.countPtr  = &s.dCounters->visible;
    sa.maxCount  = n;
    sa.histogram = &s.dCounters->depthHist[p][0];
    sa.status    = s.dSortStatus;
    sa.ticket    = &s.dCounters->ticket[1 + p];
    sa.epoch     = nextEpoch(c);
    sa.shift     = 8 * p;
    launchSortPass(sa, st, pdl && !(presorted));
    c->launches++;
    mark(VKGS_K_SORT_PASS0 + p + 1);
  }
  // the sorted pairs are in buffer sortSrc[3] (0 unless an odd number of passes was skipped)

  // ---- "Rasterization": binning, tile sort, blend -------------------------------------------------
  BinArgs ba{};
  ba.sortedIds[0] = s.dIds[0], ba.sortedIds[1] = s.dIds[1];
  ba.sortedSel  = &s.dCounters->sortSrc[3];
  ba.bboxes     = s.dBboxes;
  ba.counters   = s.dCounters;
  ba.tileKeys   = s.dTileKeys[0];
  ba.tileVals   = s.dTileVals[0];
  ba.capacity   = static_cast<uint32_t>(s.tileCapacity);
  ba.maxCount   = n;
  ba.tilesX     = tx;
  ba.tilesY     = ty;
  ba.status     = s.dBinStatus;
  ba.epoch      = nextEpoch(c);
  ba.ticketSlot = 5;
  ba.debugFlags = c->opt._reserved[0];
  ba.bigList     = s.dBigList;
  ba.bigCapacity = BIG_LIST_CAPACITY;
  launchBinEmit(ba, st, pdl && !presorted);  // (a presorted frame's binning follows the copy of the caller's ids, not a kernel)
  launchBinBig(ba, st, pdl);
  c->launches += 2;
  mark(VKGS_K_BIN_EMIT + 1);
  mark(VKGS_K_TILE_HIST + 1);  // (tile-id digit histograms are fused into the emit kernel)

  for(int p = 0; p < 2; p++)
  {
    SortPassArgs sa{};
    // fixed ping-pong (no pass skipping): pass 0 reads buffer 0, pass 1 reads buffer 1
    sa.keys[0] = s.dTileKeys[p & 1], sa.keys[1] = s.dTileKeys[(p + 1) & 1];
    sa.vals[0] = s.dTileVals[p & 1], sa.vals[1] = s.dTileVals[(p + 1) & 1];
    sa.countPtr  = &s.dCounters->tilePairsClamped;
    sa.maxCount  = static_cast<uint32_t>(s.tileCapacity);
    sa.histogram = &s.dCounters->tileHist[p][0];
    sa.status    = s.dTileSortStatus;
    sa.ticket    = &s.dCounters->ticket[6 + p];
    sa.epoch     = nextEpoch(c);
    sa.shift     = 8 * p;
    if(p == 1)
    {
      // high digit of a tile id: only ceil(log2(tiles)) - 8 significant bits
      uint32_t bits = 0;
      while((1u << (8 + bits)) < tx * ty)
        bits++;
      sa.digitBits  = static_cast<int>(bits ? bits : 1);
      sa.rangeBegin = s.dRanges;
      sa.rangeEnd   = s.dRanges + tx * ty;
    }
    launchSortPass(sa, st, pdl);
    c->launches++;
    mark(VKGS_K_TILE_SORT0 + p + 1);
  }

  mark(VKGS_K_TILE_RANGES + 1);  // (tile ranges are recorded by the last tile-sort pass)

  BlendArgs bl{};
  bl.tileVals               = s.dTileVals[0];
  bl.rangeBegin             = s.dRanges;
  bl.rangeEnd               = s.dRanges + tx * ty;
  bl.records                = s.dRecords;
  bl.image                  = s.dImage;
  bl.targetFormat           = c->opt.target_format;
  bl.width                  = fp.width;
  bl.height                 = fp.height;
  bl.tilesX                 = tx;
  bl.tilesY                 = ty;
  bl.frontToBack            = c->opt.front_to_back;
  bl.disableOpacityGaussian = c->opt.disable_opacity_gaussian;
  bl.transmittanceEpsilon   = c->opt.transmittance_epsilon;
  bl.fragmentCounters       = (c->opt._reserved[0] & 128u) ? &s.dCounters->fragments[0] : nullptr;
  if(c->opt.pipeline == VKGS_PIPELINE_3DGUT)
  {
    const vkgs_ctx::Instance& inst = c->instances[0];
    bl.gut.enabled      = 1;
    bl.gut.kernelDegree = c->opt.kernel_degree;
    bl.gut.extentEigen  = c->opt.extent_projection == VKGS_EXTENT_EIGEN;
    bl.gut.fisheye      = c->opt.camera_model == VKGS_CAMERA_FISHEYE;
    bl.gut.fovRad       = fp.fov_rad;
    std::memcpy(bl.gut.viewInverse, fp.view_inverse, sizeof(bl.gut.viewInverse));
    std::memcpy(bl.gut.projInverse, fp.proj_inverse, sizeof(bl.gut.projInverse));
    std::memcpy(bl.gut.modelInverse, inst.frameModel ? fp.model_inverse : inst.transformInverse, sizeof(bl.gut.modelInverse));
    bl.gut.instanceCount = static_cast<uint32_t>(c->instances.size());
    for(size_t k = 0; k < c->instances.size() && k < GUT_MAX_INSTANCES; k++)
    {
      const float* mi             = c->instances[k].frameModel ? fp.model_inverse : c->instances[k].transformInverse;
      bl.gut.instanceOffset[k]    = c->instances[k].globalOffset;
      for(int i = 0; i < 3; i++)
        for(int j = 0; j < 3; j++)
          bl.gut.instanceInverse[k][3 * i + j] = mi[4 * i + j];
    }
    bl.gut.viewport[0] = fp.viewport[0], bl.gut.viewport[1] = fp.viewport[1];
    bl.gut.alphaClamp         = fp.alpha_clamp;
    bl.gut.kernelMinResponse  = fp.kernel_min_response;
    bl.gut.alphaCullThreshold = fp.alpha_cull_threshold;
  }
  if(c->opt.surface_info)
  {
    bl.surface           = s.dSurface;
    bl.outNormals        = s.dOutNormals;
    bl.outDepthT         = s.dOutDepthT;
    bl.outSplatId        = s.dOutSplatId;
    bl.depthIsoThreshold = fp.depth_iso_threshold;
  }
  // the blend (and the copies to host) run on the slot's low-priority stream; the slot's main stream
  // waits for them, so frame completion / buffer reuse are still ordered on `st`
  CU_TRY(c, cudaEventRecord(s.evFront, st));
  CU_TRY(c, cudaStreamWaitEvent(s.streamBlend, s.evFront, 0));
  // A synchronous frame to host memory has nothing else to overlap its copy with: it is blended in strips of whole tile
  // rows (one launch each) and every strip is copied out on a second stream while the next one is blended. Frames of the
  // asynchronous path keep one launch and one copy: their copy overlaps the next frame. Measured (1 M splats, RGBA16F):
  // 647 -> 581 us per call at 1080p, 1682 -> 1489 us at 4K; an 8 MB RGBA8 frame is better off in one piece (492 vs 507 us).
  const size_t   rowBytes = 4ull * formatSize(c->opt.target_format) * fp.width;
  const uint32_t strips   = (hostRgba && !throughputMode && !c->profiling && stripsOn && bl.tilesY >= 2u * FrameSlot::COPY_STRIPS
                           && rowBytes * fp.height >= (12u << 20))
                                ? FrameSlot::COPY_STRIPS
                                : 1u;
  if(strips == 1u)
  {
    launchBlend(bl, s.streamBlend);
    c->launches++;
    if(c->profiling)
      cudaEventRecord(s.ev[VKGS_K_BLEND + 1], s.streamBlend);
    CU_TRY(c, cudaMemcpyAsync(s.hCounters, s.dCounters, offsetof(FrameCounters, ticket), cudaMemcpyDeviceToHost, s.streamBlend));
    if(hostRgba)
      CU_TRY(c, cudaMemcpyAsync(hostRgba, s.dImage, rowBytes * fp.height, cudaMemcpyDeviceToHost, s.streamBlend));
    CU_TRY(c, cudaEventRecord(s.evBlend, s.streamBlend));
  }
  else
  {
    // all launches first: a copy into pageable memory blocks the caller until it is done
    for(uint32_t k = 0; k < strips; k++)
    {
      const uint32_t ty0 = bl.tilesY * k / strips, ty1 = bl.tilesY * (k + 1) / strips;
      bl.firstTile = ty0 * bl.tilesX, bl.tileCount = (ty1 - ty0) * bl.tilesX;
      launchBlend(bl, s.streamBlend);
      c->launches++;
      CU_TRY(c, cudaEventRecord(s.evStrip[k], s.streamBlend));
    }
    for(uint32_t k = 0; k < strips; k++)
    {
      const uint32_t ty0 = bl.tilesY * k / strips, ty1 = bl.tilesY * (k + 1) / strips;
      CU_TRY(c, cudaStreamWaitEvent(s.streamCopy, s.evStrip[k], 0));
      const size_t y0 = static_cast<size_t>(ty0) * TILE_H, y1 = std::min<size_t>(static_cast<size_t>(ty1) * TILE_H, fp.height);
      CU_TRY(c, cudaMemcpyAsync(static_cast<char*>(hostRgba) + y0 * rowBytes, static_cast<const char*>(s.dImage) + y0 * rowBytes,
                                (y1 - y0) * rowBytes, cudaMemcpyDeviceToHost, s.streamCopy));
    }
    CU_TRY(c, cudaMemcpyAsync(s.hCounters, s.dCounters, offsetof(FrameCounters, ticket), cudaMemcpyDeviceToHost, s.streamCopy));
    CU_TRY(c, cudaEventRecord(s.evBlend, s.streamCopy));
  }
  s.evRecorded = c->profiling;
  CU_TRY(c, cudaStreamWaitEvent(st, s.evBlend, 0));
  if(c->userStream)
  {
    // completion of this frame becomes visible on the caller's stream, in submission order
    CU_TRY(c, cudaEventRecord(s.evDone, st));
    CU_TRY(c, cudaStreamWaitEvent(c->userStream, s.evDone, 0));
  }
  CU_TRY(c, cudaGetLastError());
  s.lastFp        = fp;
  s.lastHost      = hostRgba;
  s.lastThin      = throughputMode;
  s.lastPresorted = presorted;
  s.haveFrame     = true;
  s.framesSinceSync++;
  c->lastSlot = si;
  if(rq.forceSlot < 0)
    c->nextSlot = (si + 1) % std::max(1, c->framesInFlight);
  if(slotOut)
    *slotOut = si;
  return VKGS_OK;
}

// After a sync: if a frame of a slot overflowed its tile lists since the last look, grow the lists of every slot to
// what that frame wanted. flagged[i] = the LAST frame of slot i must be rendered again; *lost = an earlier frame of some
// slot overflowed too (several frames were enqueued on the slot between two syncs) and its result — already overwritten
// in the slot, possibly copied to the caller's host buffer — came from truncated lists.
// Returns VKGS_ERR_OVERFLOW when something was grown.
int checkOverflow(vkgs_ctx* c, bool* flagged /*[MAX_FRAMES_IN_FLIGHT], optional*/, bool* lost = nullptr)
{
  bool grown = false;
  for(int i = 0; i < MAX_FRAMES_IN_FLIGHT; i++)
  {
    FrameSlot& s = c->slots[i];
    if(flagged)
      flagged[i] = false;
    const uint32_t frames = s.framesSinceSync;
    s.framesSinceSync     = 0;
    if(s.haveFrame && (s.hCounters->overflow || s.hCounters->stickyOverflow))
    {
      const uint64_t pairs = std::max(s.hCounters->tilePairs, s.hCounters->stickyPairs);
      const uint64_t want  = pairs * 5 / 4 + 65536;
      for(auto& t : c->slots)
        if(t.tileCapacity && t.tileCapacity < want)
        {
          cudaStreamSynchronize(t.stream);
          if(int rc = allocTileLists(c, t, want))
            return rc;
        }
      if(lost && frames > 1)
        *lost = true;
      if(flagged)
        flagged[i] = s.hCounters->overflow != 0;
      s.hCounters->overflow = s.hCounters->stickyOverflow = s.hCounters->stickyPairs = 0;
      cudaMemset(s.dCounters, 0, offsetof(FrameCounters, visible));  // clear the sticky words
      legacyStreamBarrier();
      grown = true;
    }
  }
  return grown ? fail(c, VKGS_ERR_OVERFLOW, "tile lists overflowed; capacity was grown, render the frame again") : VKGS_OK;
}

// See enqueueFrame: slot `si` is about to be reused while it still holds a frame that has not been checked.
int repairSlotBeforeReuse(vkgs_ctx* c, int si)
{
  FrameSlot& s = c->slots[si];
  for(int attempt = 0; attempt < 4; attempt++)
  {
    CU_TRY(c, cudaStreamSynchronize(s.stream));
    if(!(s.hCounters->overflow || s.hCounters->stickyOverflow))
    {
      s.framesSinceSync = 0;
      return VKGS_OK;
    }
    const bool     truncated = s.hCounters->overflow != 0;
    const uint64_t pairs     = std::max(s.hCounters->tilePairs, s.hCounters->stickyPairs);
    const uint64_t want      = pairs * 5 / 4 + 65536;
    for(auto& t : c->slots)
      if(t.tileCapacity && t.tileCapacity < want)
      {
        CU_TRY(c, cudaStreamSynchronize(t.stream));
        if(int rc = allocTileLists(c, t, want))
          return rc;
      }
    s.hCounters->overflow = s.hCounters->stickyOverflow = s.hCounters->stickyPairs = 0;
    CU_TRY(c, cudaMemset(s.dCounters, 0, offsetof(FrameCounters, visible)));  // clear the sticky words
    legacyStreamBarrier();
    s.framesSinceSync = 0;
    if(!truncated)
      return VKGS_OK;  // (a sticky flag of an earlier, already repaired frame)
    if(s.lastPresorted)
      return fail(c, VKGS_ERR_OVERFLOW, "tile lists overflowed in a presorted frame; capacity was grown, render it again");
    FrameRequest rq;
    rq.hostRgba       = s.lastHost;
    rq.throughputMode = s.lastThin;
    rq.forceSlot      = si;
    const vkgs_frame_params fp   = s.lastFp;
    const int               last = c->lastSlot;
    if(int rc = enqueueFrame(c, fp, rq, nullptr))
      return rc;
    c->lastSlot = last;
  }
  return fail(c, VKGS_ERR_OVERFLOW, "tile lists still overflow after regrowing");
}

int syncAll(vkgs_ctx* c)
{
  for(auto& s : c->slots)
    if(s.stream)
      CU_TRY(c, cudaStreamSynchronize(s.stream));
  if(c->userStream)
    CU_TRY(c, cudaStreamSynchronize(c->userStream));
  return VKGS_OK;
}

void fillStats(vkgs_ctx* c, FrameSlot& s, vkgs_outputs* out)
{
  out->visible_count = s.hCounters->visible;
  out->tile_pairs    = s.hCounters->tilePairs;
  out->list_entries_evaluated = s.hCounters->fragments[0];
  out->fragments_blended      = s.hCounters->fragments[1];
  const uint64_t n = c->totalSplats, v = out->visible_count, p = static_cast<uint64_t>(s.lastFp.width) * s.lastFp.height;
  // SH bytes per visible splat: exact for one set, splat-count weighted over the instances otherwise
  double shB = 0.0;
  for(const auto& inst : c->instances)
  {
    const DeviceSplatSet& set = c->sets[inst.setIndex].view;
    const uint32_t        deg = std::min(set.shDegree, s.lastFp.sh_degree);
    shB += 12.0 * ((deg + 1) * (deg + 1) - 1) * set.count;
  }
  shB                    = n ? shB / static_cast<double>(n) : 0.0;
  out->bytes_algorithmic = 12 * n + static_cast<uint64_t>((132.0 + shB) * static_cast<double>(v)) + 16 * p;
  std::memset(out->ms_kernel, 0, sizeof(out->ms_kernel));
  out->ms_dist = out->ms_sort = out->ms_raster = out->ms_total = 0.0f;
  if(s.evRecorded)
  {
    for(int k = 0; k < VKGS_K_COUNT; k++)
      cudaEventElapsedTime(&out->ms_kernel[k], s.ev[k], s.ev[k + 1]);
    out->ms_dist = out->ms_kernel[VKGS_K_PREPROCESS];
    for(int k = VKGS_K_SORT_SCAN; k < VKGS_K_BIN_EMIT; k++)
      out->ms_sort += out->ms_kernel[k];
    for(int k = VKGS_K_BIN_EMIT; k < VKGS_K_COUNT; k++)
      out->ms_raster += out->ms_kernel[k];
    cudaEventElapsedTime(&out->ms_total, s.ev[0], s.ev[VKGS_K_COUNT]);
  }
}

}  // namespace

// ----------------------------------------------------------------------------------------------
extern "C" {

const char* vkgs_version(void)
{
  return "vkgs_b200 0.2.0 (sm_100a)";
}

uint32_t vkgs_abi_struct_size(int which)
{
  switch(which)
  {
    case 0:
      return sizeof(vkgs_splat_set_view);
    case 1:
      return sizeof(vkgs_options);
    case 2:
      return sizeof(vkgs_frame_params);
    case 3:
      return sizeof(vkgs_camera);
    case 4:
      return sizeof(vkgs_outputs);
    case 5:
      return sizeof(vkgs_instance);
    case 6:
      return sizeof(vkgs_image_metrics);
    default:
      return 0;
  }
}

int vkgs_pack_host(const vkgs_splat_set_view* set, const vkgs_options* optIn, float* centers, float* cov6, void* rgba, void* sh)
{
  if(!set)
    return VKGS_ERR_INVALID_ARGUMENT;
  vkgs_options opt;
  if(optIn)
    opt = *optIn;
  else
    vkgs_default_options(&opt);
  try
  {
    PackedSplatSet packed;
    if(int rc = packSplatSet(*set, opt, 1, packed))
      return rc;
    const uint64_t n = packed.count;
    if(centers)
      std::memcpy(centers, packed.centers.data(), n * 12);
    if(cov6)
      std::memcpy(cov6, packed.cov6.data(), n * 24);
    if(rgba)
      std::memcpy(rgba, packed.rgba.data(), n * 4 * formatSize(opt.rgba_format));
    if(sh && packed.shDegree)
      std::memcpy(sh, packed.sh.data(), n * 45 * formatSize(opt.sh_format));
  }
  catch(const std::exception&)
  {
    return VKGS_ERR_OUT_OF_MEMORY;
  }
  return VKGS_OK;
}

int vkgs_create(int device, vkgs_ctx** out)
{
  if(!out)
    return VKGS_ERR_INVALID_ARGUMENT;
  *out    = nullptr;
  int cnt = 0;
  if(cudaGetDeviceCount(&cnt) != cudaSuccess || cnt == 0 || device < 0 || device >= cnt)
    return VKGS_ERR_NO_DEVICE;
  cudaDeviceProp prop{};
  if(cudaGetDeviceProperties(&prop, device) != cudaSuccess || prop.major != 10)
    return VKGS_ERR_NO_DEVICE;  // kernels are built for sm_100a only
  if(cudaSetDevice(device) != cudaSuccess)
    return VKGS_ERR_NO_DEVICE;
  vkgs_ctx* c = new vkgs_ctx();
  c->device   = device;
  bool ok     = true;
  int  prioLeast = 0, prioGreatest = 0;
  cudaDeviceGetStreamPriorityRange(&prioLeast, &prioGreatest);
  for(auto& s : c->slots)
  {
    ok = ok && cudaStreamCreateWithPriority(&s.stream, cudaStreamNonBlocking, prioGreatest) == cudaSuccess;
    ok = ok && cudaStreamCreateWithPriority(&s.streamBlend, cudaStreamNonBlocking, prioLeast) == cudaSuccess;
    ok = ok && cudaStreamCreateWithPriority(&s.streamCopy, cudaStreamNonBlocking, prioLeast) == cudaSuccess;
    for(auto& e : s.evStrip)
      ok = ok && cudaEventCreateWithFlags(&e, cudaEventDisableTiming) == cudaSuccess;
    ok = ok && cudaEventCreateWithFlags(&s.evFront, cudaEventDisableTiming) == cudaSuccess;
    ok = ok && cudaEventCreateWithFlags(&s.evBlend, cudaEventDisableTiming) == cudaSuccess;
    for(auto& e : s.ev)
      ok = ok && cudaEventCreate(&e) == cudaSuccess;
    ok = ok && cudaEventCreateWithFlags(&s.evDone, cudaEventDisableTiming) == cudaSuccess;
    ok = ok && cudaMalloc(&s.dCounters, sizeof(FrameCounters)) == cudaSuccess;
    ok = ok && cudaMemset(s.dCounters, 0, sizeof(FrameCounters)) == cudaSuccess;  // (the sticky words are never cleared per frame)
    ok = ok && cudaMallocHost(&s.hCounters, sizeof(FrameCounters)) == cudaSuccess;
    if(ok)
      std::memset(s.hCounters, 0, sizeof(FrameCounters));
  }
  legacyStreamBarrier();  // (the memsets above)
  initSortKernels();
  initPreprocessKernels();
  initBlendKernels();
  if(!ok || cudaGetLastError() != cudaSuccess)
  {
    vkgs_destroy(c);
    return VKGS_ERR_CUDA;
  }
  *out = c;
  return VKGS_OK;
}

int vkgs_destroy(vkgs_ctx* c)
{
  if(!c)
    return VKGS_ERR_INVALID_ARGUMENT;
  cudaSetDevice(c->device);
  for(auto& s : c->slots)
    if(s.stream)
      cudaStreamSynchronize(s.stream);
  freeScene(c);
  freeDev(c->dCapture);
  for(auto& s : c->slots)
  {
    freeDev(s.dImage), freeDev(s.dRanges), freeDev(s.dCounters);
    freeDev(s.dOutNormals), freeDev(s.dOutDepthT), freeDev(s.dOutSplatId);
    if(s.hCounters)
      cudaFreeHost(s.hCounters);
    for(auto& e : s.ev)
      if(e)
        cudaEventDestroy(e);
    if(s.evDone)
      cudaEventDestroy(s.evDone);
    if(s.evFront)
      cudaEventDestroy(s.evFront);
    if(s.evBlend)
      cudaEventDestroy(s.evBlend);
    if(s.streamBlend)
      cudaStreamDestroy(s.streamBlend);
    if(s.streamCopy)
      cudaStreamDestroy(s.streamCopy);
    for(auto& e : s.evStrip)
      if(e)
        cudaEventDestroy(e);
    if(s.stream)
      cudaStreamDestroy(s.stream);
  }
  delete c;
  return VKGS_OK;
}

int vkgs_set_stream(vkgs_ctx* c, void* cuda_stream)
{
  if(!c)
    return VKGS_ERR_INVALID_ARGUMENT;
  if(int rc = syncAll(c))
    return rc;
  c->userStream = static_cast<cudaStream_t>(cuda_stream);
  return VKGS_OK;
}

int vkgs_set_target_format(vkgs_ctx* c, uint32_t target_format)
{
  if(!c || target_format > VKGS_FORMAT_UINT8)
    return VKGS_ERR_INVALID_ARGUMENT;
  // (vkgs_sync, not a plain wait: frames still in flight are completed — an overflowed one repaired — in the OLD format,
  //  into destinations sized for it)
  if(int rc = c->uploaded ? vkgs_sync(c) : syncAll(c))
    return rc;
  c->opt.target_format = target_format;
  return VKGS_OK;
}

int vkgs_set_frames_in_flight(vkgs_ctx* c, int frames)
{
  if(!c || frames < 1 || frames > MAX_FRAMES_IN_FLIGHT)
    return VKGS_ERR_INVALID_ARGUMENT;
  if(int rc = c->uploaded ? vkgs_sync(c) : syncAll(c))
    return rc;
  c->framesInFlight = frames;
  c->nextSlot       = 0;
  return VKGS_OK;
}

const char* vkgs_last_error(const vkgs_ctx* c)
{
  return c ? c->lastError.c_str() : "null context";
}

uint64_t vkgs_launch_count(const vkgs_ctx* c)
{
  return c ? c->launches : 0;
}

int vkgs_set_profiling(vkgs_ctx* c, int enabled)
{
  if(!c)
    return VKGS_ERR_INVALID_ARGUMENT;
  c->profiling = enabled != 0;
  return VKGS_OK;
}

}  // extern "C"

namespace {

int uploadScene(vkgs_ctx* c, const vkgs_splat_set_view* sets, uint32_t setCount, const vkgs_instance* instances,
                uint32_t instanceCount, const vkgs_options* optIn)
{
  vkgs_options opt;
  if(optIn)
    opt = *optIn;
  else
    vkgs_default_options(&opt);
  if(setCount == 0 || (instances && instanceCount == 0))
    return fail(c, VKGS_ERR_INVALID_ARGUMENT, "a scene needs at least one splat set and one instance");
  if(opt.frustum_culling_mode > VKGS_FRUSTUM_CULLING_AT_RASTER)
    return fail(c, VKGS_ERR_INVALID_ARGUMENT, "bad frustum_culling_mode");
  if(opt.target_format > VKGS_FORMAT_UINT8)
    return fail(c, VKGS_ERR_INVALID_ARGUMENT, "bad target_format");
  if(opt.pipeline > VKGS_PIPELINE_3DGUT)
    return fail(c, VKGS_ERR_INVALID_ARGUMENT, "bad pipeline");
  if(opt.pipeline == VKGS_PIPELINE_3DGUT
     && (opt.extent_projection > VKGS_EXTENT_CONIC || opt.surface_info || (instances && instanceCount > GUT_MAX_INSTANCES)))
    return fail(c, VKGS_ERR_UNSUPPORTED, "the 3DGUT pipeline is built for at most 8 instances and without surface info");
  if(opt.camera_model > VKGS_CAMERA_FISHEYE || (opt.camera_model == VKGS_CAMERA_FISHEYE && opt.pipeline != VKGS_PIPELINE_3DGUT))
    return fail(c, VKGS_ERR_UNSUPPORTED, "the fisheye camera model needs the 3DGUT pipeline (the 3DGS raster pipelines are pinhole only)");
  if(opt.surface_info && !opt.front_to_back)
    return fail(c, VKGS_ERR_UNSUPPORTED, "surface_info needs front_to_back (the reference only produces it in its FTB pass)");
  uint64_t total = 0;
  for(uint32_t k = 0; k < (instances ? instanceCount : 1u); k++)
  {
    const uint32_t si = instances ? instances[k].splat_set_index : 0u;
    if(si >= setCount)
      return fail(c, VKGS_ERR_INVALID_ARGUMENT, "instance refers to a splat set that does not exist");
    if(sets[si].count == 0)
      return fail(c, VKGS_ERR_INVALID_ARGUMENT, "splat count must be in 1..2^31-1");
    total += sets[si].count;
  }
  if(total > 0x7fffffffull)
    return fail(c, VKGS_ERR_INVALID_ARGUMENT, "splat count must be in 1..2^31-1");
  CU_TRY(c, cudaSetDevice(c->device));
  // frames of the previous scene that are still in flight are completed first (an overflowed one repaired), so that every
  // host destination handed to vkgs_render_to_host_async holds its frame before the scene it was rendered from goes away
  if(c->uploaded)
  {
    const int rc = vkgs_sync(c);
    if(rc != VKGS_OK && rc != VKGS_ERR_OVERFLOW)
      return rc;
  }
  else if(int rc = syncAll(c))
    return rc;

  freeScene(c);
  auto up = [&](void*& dst, const void* src, size_t bytes) -> cudaError_t {
    cudaError_t e = cudaMalloc(&dst, bytes);
    if(e != cudaSuccess)
      return e;
    return cudaMemcpy(dst, src, bytes, cudaMemcpyHostToDevice);
  };
  c->sets.resize(setCount);
  for(uint32_t i = 0; i < setCount; i++)
  {
    PackedSplatSet packed;
    if(int rc = packSplatSet(sets[i], opt, PRE_TILE, packed))
    {
      freeScene(c);
      return fail(c, rc, "packSplatSet failed (null array, or f_rest_per_splat not 0/45)");
    }
    vkgs_ctx::SetStorage& st = c->sets[i];
    CU_TRY(c, up(st.dCenters, packed.centers.data(), packed.centers.size() * 4));
    CU_TRY(c, up(st.dCov, packed.cov6.data(), packed.cov6.size() * 4));
    CU_TRY(c, up(st.dScales, packed.scales.data(), packed.scales.size() * 4));
    CU_TRY(c, up(st.dRgba, packed.rgba.data(), packed.rgba.size()));
    if(packed.shDegree)
      CU_TRY(c, up(st.dSh, packed.sh.data(), packed.sh.size()));
    if(!packed.rotations.empty())
      CU_TRY(c, up(st.dRotations, packed.rotations.data(), packed.rotations.size() * 4));
    st.view.rotations  = static_cast<const float*>(st.dRotations);
    st.view.centers    = static_cast<const float*>(st.dCenters);
    st.view.cov6       = static_cast<const float*>(st.dCov);
    st.view.scales     = static_cast<const float*>(st.dScales);
    st.view.rgba       = st.dRgba;
    st.view.sh         = st.dSh;
    st.view.count      = static_cast<uint32_t>(packed.count);
    st.view.shDegree   = packed.shDegree;
    st.view.shFormat   = packed.shFormat;
    st.view.rgbaFormat = packed.rgbaFormat;
  }
  // instances in creation order: global id = globalOffset + local id (rebuildGlobalIndexTables)
  uint32_t offset = 0, tiles = 0;
  for(uint32_t k = 0; k < (instances ? instanceCount : 1u); k++)
  {
    vkgs_ctx::Instance inst;
    inst.setIndex     = instances ? instances[k].splat_set_index : 0u;
    inst.globalOffset = offset;
    inst.tileOffset   = tiles;
    inst.frameModel   = instances == nullptr;
    if(instances)
    {
      std::memcpy(inst.transform, instances[k].transform, sizeof(inst.transform));
      std::memcpy(inst.transformInverse, instances[k].transform_inverse, sizeof(inst.transformInverse));
    }
    const uint32_t cnt = c->sets[inst.setIndex].view.count;
    offset += cnt;
    tiles += (cnt + PRE_TILE - 1) / PRE_TILE;
    c->instances.push_back(inst);
  }
  c->totalSplats = offset;
  c->totalTiles  = tiles;
  c->opt         = opt;
  for(auto& s : c->slots)  // (re)allocated with the next frame if the new options need them
    freeDev(s.dOutNormals), freeDev(s.dOutDepthT), freeDev(s.dOutSplatId);

  // sorting / raster buffers (the reference allocates its sorting buffers with the splat set too,
  // src/splat_set_manager_vk.cpp:2426-2517), one set per frame in flight
  for(auto& s : c->slots)
    if(int rc = allocSlotScene(c, s, offset, tiles))
      return rc;
  // the scene arrays were copied from pageable host memory on the legacy stream: the copies may still be in flight when
  // cudaMemcpy returns, and the frame streams are not ordered behind them
  CU_TRY(c, cudaStreamSynchronize(cudaStreamLegacy));
  c->uploaded = true;
  c->nextSlot = 0;
  return VKGS_OK;
}

}  // namespace

extern "C" {

int vkgs_upload(vkgs_ctx* c, const vkgs_splat_set_view* set, const vkgs_options* optIn)
{
  if(!c || !set)
    return VKGS_ERR_INVALID_ARGUMENT;
  try
  {
    return uploadScene(c, set, 1, nullptr, 0, optIn);
  }
  catch(const std::exception& e)
  {
    return fail(c, VKGS_ERR_OUT_OF_MEMORY, (std::string("host allocation failed while packing the scene: ") + e.what()).c_str());
  }
}

int vkgs_upload_scene(vkgs_ctx* c, const vkgs_splat_set_view* sets, uint32_t set_count, const vkgs_instance* instances,
                      uint32_t instance_count, const vkgs_options* optIn)
{
  if(!c || !sets || !instances)
    return VKGS_ERR_INVALID_ARGUMENT;
  try
  {
    return uploadScene(c, sets, set_count, instances, instance_count, optIn);
  }
  catch(const std::exception& e)
  {
    return fail(c, VKGS_ERR_OUT_OF_MEMORY, (std::string("host allocation failed while packing the scene: ") + e.what()).c_str());
  }
}

int vkgs_set_instance_transform(vkgs_ctx* c, uint32_t instance, const float* transform, const float* transform_inverse)
{
  if(!c || !transform || !transform_inverse)
    return VKGS_ERR_INVALID_ARGUMENT;
  if(!c->uploaded || instance >= c->instances.size() || c->instances[instance].frameModel)
    return fail(c, VKGS_ERR_INVALID_ARGUMENT, "no such instance (scenes uploaded with vkgs_upload take their transform from the frame parameters)");
  // frames already enqueued captured the old transform by value (kernel arguments): no sync needed
  std::memcpy(c->instances[instance].transform, transform, 16 * sizeof(float));
  std::memcpy(c->instances[instance].transformInverse, transform_inverse, 16 * sizeof(float));
  return VKGS_OK;
}

int vkgs_global_index_table(const vkgs_ctx* c, uint32_t* instance_index, uint32_t* splat_index, uint64_t capacity, uint64_t* total)
{
  if(!c || !c->uploaded)
    return VKGS_ERR_INVALID_ARGUMENT;
  if(total)
    *total = c->totalSplats;
  for(size_t k = 0; k < c->instances.size(); k++)
  {
    const vkgs_ctx::Instance& inst = c->instances[k];
    const uint32_t            cnt  = c->sets[inst.setIndex].view.count;
    for(uint32_t i = 0; i < cnt && inst.globalOffset + i < capacity; i++)
    {
      if(instance_index)
        instance_index[inst.globalOffset + i] = static_cast<uint32_t>(k);
      if(splat_index)
        splat_index[inst.globalOffset + i] = i;
    }
  }
  return VKGS_OK;
}

int vkgs_render_async(vkgs_ctx* c, const vkgs_frame_params* fp)
{
  if(!c || !fp)
    return VKGS_ERR_INVALID_ARGUMENT;
  FrameRequest rq;
  rq.throughputMode = true;
  return enqueueFrame(c, *fp, rq, nullptr);
}

int vkgs_render_to_host_async(vkgs_ctx* c, const vkgs_frame_params* fp, void* host_rgba)
{
  if(!c || !fp || !host_rgba)
    return VKGS_ERR_INVALID_ARGUMENT;
  FrameRequest rq;
  rq.hostRgba       = host_rgba;
  rq.throughputMode = true;
  return enqueueFrame(c, *fp, rq, nullptr);
}

int vkgs_sync(vkgs_ctx* c)
{
  if(!c)
    return VKGS_ERR_INVALID_ARGUMENT;
  // Tile-list overflow is handled here, out of the caller's sight: the lists of every slot are grown to what the
  // overflowing frame wanted and the frames that were produced from truncated lists are enqueued again on their own
  // slots (same parameters, same host destination), so that after a successful vkgs_sync every frame in flight is
  // complete. Frames of OTHER slots are untouched: a slot's lists are private to it.
  for(int attempt = 0; attempt < 4; attempt++)
  {
    if(int rc = syncAll(c))
      return rc;
    bool      flagged[MAX_FRAMES_IN_FLIGHT];
    bool      lost = false;
    const int rc   = checkOverflow(c, flagged, &lost);
    if(rc != VKGS_ERR_OVERFLOW)
      return rc;
    if(lost)
      return fail(c, VKGS_ERR_OVERFLOW, "tile lists overflowed in a frame whose slot was reused before vkgs_sync; capacity was grown, "
                                        "render the frames since the previous vkgs_sync again");
    const int last = c->lastSlot;
    for(int i = 0; i < MAX_FRAMES_IN_FLIGHT; i++)
      if(flagged[i])
      {
        FrameSlot& s = c->slots[i];
        if(s.lastPresorted)
          return fail(c, VKGS_ERR_OVERFLOW, "tile lists overflowed in a presorted frame; capacity was grown, render it again");
        FrameRequest rq;
        rq.hostRgba       = s.lastHost;
        rq.throughputMode = s.lastThin;
        rq.forceSlot      = i;
        const vkgs_frame_params fp = s.lastFp;
        if(int rc2 = enqueueFrame(c, fp, rq, nullptr))
          return rc2;
      }
    c->lastSlot = last;
  }
  return fail(c, VKGS_ERR_OVERFLOW, "tile lists still overflow after regrowing");
}

int vkgs_last_frame_stats(vkgs_ctx* c, vkgs_outputs* out)
{
  if(!c || !out)
    return VKGS_ERR_INVALID_ARGUMENT;
  if(c->lastSlot < 0 || !c->slots[c->lastSlot].haveFrame)
    return fail(c, VKGS_ERR_INVALID_ARGUMENT, "no frame rendered yet");
  FrameSlot& s = c->slots[c->lastSlot];
  CU_TRY(c, cudaStreamSynchronize(s.stream));
  fillStats(c, s, out);
  return VKGS_OK;
}

const void* vkgs_device_framebuffer(const vkgs_ctx* c)
{
  return (c && c->lastSlot >= 0) ? c->slots[c->lastSlot].dImage : nullptr;
}

static int renderSync(vkgs_ctx* c, const vkgs_frame_params* fp, vkgs_outputs* out, const uint32_t* presortedIds, uint32_t presortedCount)
{
  for(int attempt = 0; attempt < 3; attempt++)
  {
    int          si = 0;
    FrameRequest rq;
    rq.hostRgba       = out->rgba;  // synchronous call: nothing to overlap with, full-occupancy front end
    rq.presortedIds   = presortedIds;
    rq.presortedCount = presortedCount;
    if(int rc = enqueueFrame(c, *fp, rq, &si))
      return rc;
    FrameSlot& s = c->slots[si];
    if(int rc = syncAll(c))  // (every slot: the overflow check below reads the counters of all of them)
      return rc;
    const int rc = checkOverflow(c, nullptr);
    if(rc == VKGS_ERR_OVERFLOW)
      continue;  // lists were regrown: run the frame again
    if(rc)
      return rc;
    fillStats(c, s, out);
    const uint64_t v   = std::min<uint64_t>(out->visible_count, out->sorted_ids_capacity);
    const uint32_t sel = s.hCounters->sortSrc[3] & 1u;
    if(out->sorted_ids && v)
      CU_TRY(c, cudaMemcpy(out->sorted_ids, s.dIds[sel], v * sizeof(uint32_t), cudaMemcpyDeviceToHost));
    if(out->sorted_keys && v && !presortedIds)
      CU_TRY(c, cudaMemcpy(out->sorted_keys, s.dKeys[sel], v * sizeof(uint32_t), cudaMemcpyDeviceToHost));
    return VKGS_OK;
  }
  return fail(c, VKGS_ERR_OVERFLOW, "tile lists still overflow after regrowing");
}

int vkgs_render(vkgs_ctx* c, const vkgs_frame_params* fp, vkgs_outputs* out)
{
  if(!c || !fp || !out)
    return VKGS_ERR_INVALID_ARGUMENT;
  return renderSync(c, fp, out, nullptr, 0);
}

int vkgs_render_presorted(vkgs_ctx* c, const vkgs_frame_params* fp, const uint32_t* ids, uint64_t count, vkgs_outputs* out)
{
  if(!c || !fp || !out || (!ids && count))
    return VKGS_ERR_INVALID_ARGUMENT;
  if(!c->uploaded)
    return fail(c, VKGS_ERR_NOT_UPLOADED, "vkgs_render_presorted before vkgs_upload");
  if(count > c->totalSplats)
    return fail(c, VKGS_ERR_INVALID_ARGUMENT, "more ids than splats in the scene");
  for(uint64_t i = 0; i < count; i++)
    if(ids[i] >= c->totalSplats)
      return fail(c, VKGS_ERR_INVALID_ARGUMENT, "presorted id out of range");
  static const uint32_t none = 0;
  return renderSync(c, fp, out, count ? ids : &none, static_cast<uint32_t>(count));
}

int vkgs_read_records(vkgs_ctx* c, uint32_t* records12, uint64_t first, uint64_t count)
{
  if(!c || !records12 || !c->uploaded || c->lastSlot < 0 || first + count > c->totalSplats)
    return VKGS_ERR_INVALID_ARGUMENT;
  if(c->opt.pipeline == VKGS_PIPELINE_3DGUT)
    return fail(c, VKGS_ERR_UNSUPPORTED, "vkgs_read_records: the 3DGUT pipeline keeps a different (24-word) per-splat record");
  FrameSlot& s = c->slots[c->lastSlot];
  CU_TRY(c, cudaStreamSynchronize(s.stream));
  CU_TRY(c, cudaMemcpy(records12, s.dRecords + first * RECORD_WORDS, count * RECORD_WORDS * sizeof(uint32_t), cudaMemcpyDeviceToHost));
  return VKGS_OK;
}

int vkgs_read_surface_info(vkgs_ctx* c, float* normals, float* depth_transmittance, uint32_t* splat_id)
{
  if(!c)
    return VKGS_ERR_INVALID_ARGUMENT;
  if(!c->uploaded || !c->opt.surface_info || c->lastSlot < 0 || !c->slots[c->lastSlot].dOutNormals)
    return fail(c, VKGS_ERR_INVALID_ARGUMENT, "no frame with options.surface_info rendered yet");
  FrameSlot& s = c->slots[c->lastSlot];
  CU_TRY(c, cudaStreamSynchronize(s.stream));
  const size_t px = static_cast<size_t>(s.imgW) * s.imgH;
  if(normals)
    CU_TRY(c, cudaMemcpy(normals, s.dOutNormals, px * sizeof(float4), cudaMemcpyDeviceToHost));
  if(depth_transmittance)
    CU_TRY(c, cudaMemcpy(depth_transmittance, s.dOutDepthT, px * sizeof(float2), cudaMemcpyDeviceToHost));
  if(splat_id)
    CU_TRY(c, cudaMemcpy(splat_id, s.dOutSplatId, px * sizeof(uint32_t), cudaMemcpyDeviceToHost));
  return VKGS_OK;
}

int vkgs_read_packed(vkgs_ctx* c, float* centers, float* cov6, float* rgba, float* sh)
{
  if(!c || !c->uploaded)
    return VKGS_ERR_INVALID_ARGUMENT;
  const vkgs_ctx::SetStorage& st = c->sets[0];  // (splat set 0)
  if(st.view.shFormat != VKGS_FORMAT_FLOAT32 || st.view.rgbaFormat != VKGS_FORMAT_FLOAT32)
    return fail(c, VKGS_ERR_UNSUPPORTED, "vkgs_read_packed needs fp32 formats");
  const uint64_t n = st.view.count;
  if(centers)
    CU_TRY(c, cudaMemcpy(centers, st.dCenters, n * 12, cudaMemcpyDeviceToHost));
  if(cov6)
    CU_TRY(c, cudaMemcpy(cov6, st.dCov, n * 24, cudaMemcpyDeviceToHost));
  if(rgba)
    CU_TRY(c, cudaMemcpy(rgba, st.dRgba, n * 16, cudaMemcpyDeviceToHost));
  if(sh && st.dSh)
    CU_TRY(c, cudaMemcpy(sh, st.dSh, n * 180, cudaMemcpyDeviceToHost));
  return VKGS_OK;
}

}  // extern "C"
