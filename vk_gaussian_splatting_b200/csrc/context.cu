// context.cu — the C ABI (include/vkgs_b200.h): context lifetime, scene upload, frame orchestration.
//
// One frame = the reference's processSortingOnGPU + drawSplatPrimitives
// (src/gaussian_splatting.cpp:1298-1367, 1369-1465) as a fixed sequence of stream-ordered launches
// with every data-dependent size (V, tile-pair count) read on the device — no host round trip
// inside a frame:
//   memset(control block) -> preprocess -> 4 x sort pass -> bin emit -> 2 x tile sort pass
//   -> tile ranges -> blend
#include <algorithm>
#include <cmath>
#include <cstdio>
#include <cstring>
#include <string>
#include <vector>

#include "host_pack.hpp"
#include "kernels.hpp"
#include "vkgs_b200.h"

namespace vkgs {
void initSortKernels();
void initPreprocessKernels();
}  // namespace vkgs

using namespace vkgs;

struct vkgs_ctx
{
  int          device      = 0;
  cudaStream_t ownStream   = nullptr;
  cudaStream_t stream      = nullptr;
  std::string  lastError;
  uint64_t     launches    = 0;
  uint32_t     epoch       = 0;
  bool         profiling   = false;

  // scene
  bool           uploaded = false;
  vkgs_options   opt{};
  DeviceSplatSet set{};
  uint64_t       paddedCount = 0;
  void *         dCenters = nullptr, *dCov = nullptr, *dScales = nullptr, *dRgba = nullptr, *dSh = nullptr;

  // per-frame buffers
  uint32_t *     dKeys[2] = {nullptr, nullptr}, *dIds[2] = {nullptr, nullptr};
  uint32_t*      dRecords    = nullptr;
  FrameCounters* dCounters   = nullptr;
  uint64_t *     dPreStatus = nullptr, *dSortStatus = nullptr, *dBinStatus = nullptr, *dTileSortStatus = nullptr;
  uint32_t *     dTileKeys[2] = {nullptr, nullptr}, *dTileVals[2] = {nullptr, nullptr};
  uint64_t       tileCapacity = 0;
  uint2*         dRanges      = nullptr;
  uint32_t       rangesTiles  = 0;
  float4*        dImage       = nullptr;
  uint32_t       imgW = 0, imgH = 0;
  FrameCounters* hCounters = nullptr;  // pinned

  // last frame
  vkgs_frame_params lastFp{};
  bool              haveFrame = false;
  cudaEvent_t       ev[VKGS_K_COUNT + 1]{};
  bool              evRecorded = false;
};

namespace {

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

int fail(vkgs_ctx* ctx, int code, const char* msg)
{
  if(ctx)
    ctx->lastError = msg;
  return code;
}

template <typename T>
void freeDev(T*& p)
{
  if(p)
    cudaFree(p);
  p = nullptr;
}

void freeScene(vkgs_ctx* c)
{
  freeDev(c->dCenters), freeDev(c->dCov), freeDev(c->dScales), freeDev(c->dRgba), freeDev(c->dSh);
  for(int i = 0; i < 2; i++)
    freeDev(c->dKeys[i]), freeDev(c->dIds[i]), freeDev(c->dTileKeys[i]), freeDev(c->dTileVals[i]);
  freeDev(c->dRecords), freeDev(c->dPreStatus), freeDev(c->dSortStatus), freeDev(c->dBinStatus), freeDev(c->dTileSortStatus);
  c->tileCapacity = 0;
  c->uploaded     = false;
}

int allocTileLists(vkgs_ctx* c, uint64_t capacity)
{
  for(int i = 0; i < 2; i++)
    freeDev(c->dTileKeys[i]), freeDev(c->dTileVals[i]);
  freeDev(c->dTileSortStatus);
  capacity = std::min<uint64_t>(capacity, 0xfffff000ull);
  for(int i = 0; i < 2; i++)
  {
    CU_TRY(c, cudaMalloc(&c->dTileKeys[i], capacity * sizeof(uint32_t)));
    CU_TRY(c, cudaMalloc(&c->dTileVals[i], capacity * sizeof(uint32_t)));
  }
  const uint64_t parts = (capacity + SORT_PART - 1) / SORT_PART;
  CU_TRY(c, cudaMalloc(&c->dTileSortStatus, parts * 256 * sizeof(uint64_t)));
  CU_TRY(c, cudaMemset(c->dTileSortStatus, 0, parts * 256 * sizeof(uint64_t)));
  c->tileCapacity = capacity;
  return VKGS_OK;
}

int ensureTargets(vkgs_ctx* c, uint32_t w, uint32_t h)
{
  if(w == 0 || h == 0 || w > 65535 || h > 65535)
    return fail(c, VKGS_ERR_INVALID_ARGUMENT, "viewport must be within 1..65535 pixels per side");
  const uint32_t tx = (w + TILE_W - 1) / TILE_W, ty = (h + TILE_H - 1) / TILE_H;
  if(tx * ty > 65536)
    return fail(c, VKGS_ERR_UNSUPPORTED, "more than 65536 tiles (tile ids are sorted on 16 bits)");
  if(c->imgW != w || c->imgH != h)
  {
    freeDev(c->dImage);
    freeDev(c->dRanges);
    CU_TRY(c, cudaMalloc(&c->dImage, sizeof(float4) * static_cast<size_t>(w) * h));
    CU_TRY(c, cudaMalloc(&c->dRanges, sizeof(uint2) * tx * ty));
    c->imgW = w, c->imgH = h, c->rangesTiles = tx * ty;
  }
  return VKGS_OK;
}

uint32_t nextEpoch(vkgs_ctx* c)
{
  c->epoch++;
  if(c->epoch >= (1u << 30))
  {
    // epoch space exhausted (2^30 launches): clear the status arrays once and restart
    cudaStreamSynchronize(c->stream);
    const uint64_t n = c->set.count;
    cudaMemset(c->dPreStatus, 0, ((n + PRE_TILE - 1) / PRE_TILE) * sizeof(uint64_t));
    cudaMemset(c->dBinStatus, 0, ((n + BIN_THREADS - 1) / BIN_THREADS) * sizeof(uint64_t));
    cudaMemset(c->dSortStatus, 0, ((n + SORT_PART - 1) / SORT_PART) * 256 * sizeof(uint64_t));
    cudaMemset(c->dTileSortStatus, 0, ((c->tileCapacity + SORT_PART - 1) / SORT_PART) * 256 * sizeof(uint64_t));
    c->epoch = 1;
  }
  return c->epoch;
}

// host-side per-frame constants, evaluated in the oracle's operation order
void frameConstants(const vkgs_frame_params& fp, float mv[16], float camModel[3])
{
  for(int i = 0; i < 4; i++)
    for(int j = 0; j < 4; j++)
      mv[4 * i + j] = ((fp.model[4 * i + 0] * fp.view[0 + j] + fp.model[4 * i + 1] * fp.view[4 + j]) + fp.model[4 * i + 2] * fp.view[8 + j])
                      + fp.model[4 * i + 3] * fp.view[12 + j];
  const float cp[4] = {fp.camera_position[0], fp.camera_position[1], fp.camera_position[2], 1.0f};
  for(int j = 0; j < 3; j++)
    camModel[j] = ((cp[0] * fp.model_inverse[0 + j] + cp[1] * fp.model_inverse[4 + j]) + cp[2] * fp.model_inverse[8 + j])
                  + cp[3] * fp.model_inverse[12 + j];
}

void mark(vkgs_ctx* c, int slot)
{
  if(c->profiling)
    cudaEventRecord(c->ev[slot], c->stream);
}

int enqueueFrame(vkgs_ctx* c, const vkgs_frame_params& fp)
{
  if(!c->uploaded)
    return fail(c, VKGS_ERR_NOT_UPLOADED, "vkgs_render before vkgs_upload");
  if(fp.width == 0 || fp.height == 0)
    return fail(c, VKGS_ERR_INVALID_ARGUMENT, "zero-sized viewport");
  if(int rc = ensureTargets(c, fp.width, fp.height))
    return rc;
  CU_TRY(c, cudaSetDevice(c->device));
  const uint32_t n  = c->set.count;
  const uint32_t tx = (fp.width + TILE_W - 1) / TILE_W, ty = (fp.height + TILE_H - 1) / TILE_H;

  CU_TRY(c, cudaMemsetAsync(c->dCounters, 0, sizeof(FrameCounters), c->stream));
  CU_TRY(c, cudaMemsetAsync(c->dRanges, 0, sizeof(uint2) * tx * ty, c->stream));
  mark(c, 0);

  // ---- "GPU Dist" (+ the per-splat half of "Rasterization", fused) -----------------------------
  PreprocessArgs pa{};
  pa.set = c->set;
  pa.fp  = fp;
  pa.opt = c->opt;
  frameConstants(fp, pa.mv, pa.camModel);
  pa.keys       = c->dKeys[0];
  pa.ids        = c->dIds[0];
  pa.records    = c->dRecords;
  pa.counters   = c->dCounters;
  pa.status     = c->dPreStatus;
  pa.epoch      = nextEpoch(c);
  pa.ticketSlot = 0;
  launchPreprocess(pa, c->stream);
  c->launches++;
  mark(c, VKGS_K_PREPROCESS + 1);
  mark(c, VKGS_K_SORT_SCAN + 1);  // (depth-key digit histograms are fused into the preprocess kernel)

  // ---- "GPU Sort": 4 x 8-bit stable passes over (key,id) ----------------------------------------
  for(int p = 0; p < 4; p++)
  {
    SortPassArgs sa{};
    sa.keys[0] = c->dKeys[0], sa.keys[1] = c->dKeys[1];
    sa.vals[0] = c->dIds[0], sa.vals[1] = c->dIds[1];
    sa.srcSelIn  = p ? &c->dCounters->sortSrc[p - 1] : nullptr;
    sa.srcSelOut = &c->dCounters->sortSrc[p];
    sa.countPtr  = &c->dCounters->visible;
    sa.maxCount  = n;
    sa.histogram = &c->dCounters->depthHist[p][0];
    sa.status    = c->dSortStatus;
    sa.ticket    = &c->dCounters->ticket[1 + p];
    sa.epoch     = nextEpoch(c);
    sa.shift     = 8 * p;
    launchSortPass(sa, c->stream);
    c->launches++;
    mark(c, VKGS_K_SORT_PASS0 + p + 1);
  }
  // the sorted pairs are in buffer sortSrc[3] (0 unless an odd number of passes was skipped)

  // ---- "Rasterization": binning, tile sort, blend -------------------------------------------------
  BinArgs ba{};
  ba.sortedIds[0] = c->dIds[0], ba.sortedIds[1] = c->dIds[1];
  ba.sortedSel  = &c->dCounters->sortSrc[3];
  ba.records    = c->dRecords;
  ba.counters   = c->dCounters;
  ba.tileKeys   = c->dTileKeys[0];
  ba.tileVals   = c->dTileVals[0];
  ba.capacity   = static_cast<uint32_t>(c->tileCapacity);
  ba.maxCount   = n;
  ba.tilesX     = tx;
  ba.tilesY     = ty;
  ba.status     = c->dBinStatus;
  ba.epoch      = nextEpoch(c);
  ba.ticketSlot = 5;
  ba.debugFlags = c->opt._reserved[5];
  launchBinEmit(ba, c->stream);
  c->launches++;
  mark(c, VKGS_K_BIN_EMIT + 1);
  mark(c, VKGS_K_TILE_HIST + 1);  // (tile-id digit histograms are fused into the emit kernel)

  for(int p = 0; p < 2; p++)
  {
    SortPassArgs sa{};
    // fixed ping-pong (no pass skipping): pass 0 reads buffer 0, pass 1 reads buffer 1
    sa.keys[0] = c->dTileKeys[p & 1], sa.keys[1] = c->dTileKeys[(p + 1) & 1];
    sa.vals[0] = c->dTileVals[p & 1], sa.vals[1] = c->dTileVals[(p + 1) & 1];
    sa.countPtr  = &c->dCounters->tilePairsClamped;
    sa.maxCount  = static_cast<uint32_t>(c->tileCapacity);
    sa.histogram = &c->dCounters->tileHist[p][0];
    sa.status    = c->dTileSortStatus;
    sa.ticket    = &c->dCounters->ticket[6 + p];
    sa.epoch     = nextEpoch(c);
    sa.shift     = 8 * p;
    launchSortPass(sa, c->stream);
    c->launches++;
    mark(c, VKGS_K_TILE_SORT0 + p + 1);
  }

  launchTileRanges(c->dTileKeys[0], c->dCounters, static_cast<uint32_t>(c->tileCapacity), c->dRanges, c->stream);
  c->launches++;
  mark(c, VKGS_K_TILE_RANGES + 1);

  BlendArgs bl{};
  bl.tileVals               = c->dTileVals[0];
  bl.ranges                 = c->dRanges;
  bl.records                = c->dRecords;
  bl.image                  = c->dImage;
  bl.width                  = fp.width;
  bl.height                 = fp.height;
  bl.tilesX                 = tx;
  bl.tilesY                 = ty;
  bl.frontToBack            = c->opt.front_to_back;
  bl.disableOpacityGaussian = c->opt.disable_opacity_gaussian;
  bl.transmittanceEpsilon   = c->opt.front_to_back ? c->opt.transmittance_epsilon : 0.0f;
  launchBlend(bl, c->stream);
  c->launches++;
  mark(c, VKGS_K_BLEND + 1);
  c->evRecorded = c->profiling;

  CU_TRY(c, cudaMemcpyAsync(c->hCounters, c->dCounters, 32, cudaMemcpyDeviceToHost, c->stream));
  CU_TRY(c, cudaGetLastError());
  c->lastFp    = fp;
  c->haveFrame = true;
  return VKGS_OK;
}

// After a sync: if the tile lists overflowed, grow them so the caller can re-render.
int checkOverflow(vkgs_ctx* c)
{
  if(c->haveFrame && c->hCounters->overflow)
  {
    const uint64_t want = static_cast<uint64_t>(c->hCounters->tilePairs) * 5 / 4 + 65536;
    if(int rc = allocTileLists(c, want))
      return rc;
    return fail(c, VKGS_ERR_OVERFLOW, "tile lists overflowed; capacity was grown, render the frame again");
  }
  return VKGS_OK;
}

void fillStats(vkgs_ctx* c, vkgs_outputs* out)
{
  out->visible_count = c->hCounters->visible;
  out->tile_pairs    = c->hCounters->tilePairs;
  const uint64_t n = c->set.count, v = out->visible_count, p = static_cast<uint64_t>(c->lastFp.width) * c->lastFp.height;
  const uint32_t deg    = std::min(c->set.shDegree, c->lastFp.sh_degree);
  const uint64_t shB    = 12ull * ((deg + 1) * (deg + 1) - 1);
  out->bytes_algorithmic = 12 * n + (132 + shB) * v + 16 * p;
  std::memset(out->ms_kernel, 0, sizeof(out->ms_kernel));
  out->ms_dist = out->ms_sort = out->ms_raster = out->ms_total = 0.0f;
  if(c->evRecorded)
  {
    for(int k = 0; k < VKGS_K_COUNT; k++)
      cudaEventElapsedTime(&out->ms_kernel[k], c->ev[k], c->ev[k + 1]);
    out->ms_dist = out->ms_kernel[VKGS_K_PREPROCESS];
    for(int k = VKGS_K_SORT_SCAN; k < VKGS_K_BIN_EMIT; k++)
      out->ms_sort += out->ms_kernel[k];
    for(int k = VKGS_K_BIN_EMIT; k < VKGS_K_COUNT; k++)
      out->ms_raster += out->ms_kernel[k];
    cudaEventElapsedTime(&out->ms_total, c->ev[0], c->ev[VKGS_K_COUNT]);
  }
}

}  // namespace

// ----------------------------------------------------------------------------------------------
extern "C" {

const char* vkgs_version(void)
{
  return "vkgs_b200 0.1.0 (sm_100a)";
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
  if(cudaStreamCreateWithFlags(&c->ownStream, cudaStreamNonBlocking) != cudaSuccess)
  {
    delete c;
    return VKGS_ERR_CUDA;
  }
  c->stream = c->ownStream;
  for(auto& e : c->ev)
    cudaEventCreate(&e);
  cudaMalloc(&c->dCounters, sizeof(FrameCounters));
  cudaMallocHost(&c->hCounters, sizeof(FrameCounters));
  std::memset(c->hCounters, 0, sizeof(FrameCounters));
  initSortKernels();
  initPreprocessKernels();
  if(cudaGetLastError() != cudaSuccess)
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
  cudaStreamSynchronize(c->stream);
  freeScene(c);
  freeDev(c->dImage), freeDev(c->dRanges), freeDev(c->dCounters);
  if(c->hCounters)
    cudaFreeHost(c->hCounters);
  for(auto& e : c->ev)
    if(e)
      cudaEventDestroy(e);
  if(c->ownStream)
    cudaStreamDestroy(c->ownStream);
  delete c;
  return VKGS_OK;
}

int vkgs_set_stream(vkgs_ctx* c, void* cuda_stream)
{
  if(!c)
    return VKGS_ERR_INVALID_ARGUMENT;
  cudaStreamSynchronize(c->stream);
  c->stream = cuda_stream ? static_cast<cudaStream_t>(cuda_stream) : c->ownStream;
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

int vkgs_upload(vkgs_ctx* c, const vkgs_splat_set_view* set, const vkgs_options* optIn)
{
  if(!c || !set)
    return VKGS_ERR_INVALID_ARGUMENT;
  vkgs_options opt;
  if(optIn)
    opt = *optIn;
  else
    vkgs_default_options(&opt);
  if(set->count == 0 || set->count > 0x7fffffffull)
    return fail(c, VKGS_ERR_INVALID_ARGUMENT, "splat count must be in 1..2^31-1");
  if(opt.frustum_culling_mode > VKGS_FRUSTUM_CULLING_AT_RASTER)
    return fail(c, VKGS_ERR_INVALID_ARGUMENT, "bad frustum_culling_mode");
  CU_TRY(c, cudaSetDevice(c->device));
  CU_TRY(c, cudaStreamSynchronize(c->stream));

  PackedSplatSet packed;
  if(int rc = packSplatSet(*set, opt, PRE_TILE, packed))
    return fail(c, rc, "packSplatSet failed (null array, or f_rest_per_splat not 0/45)");

  freeScene(c);
  const uint64_t n = packed.count, pad = packed.paddedCount;
  auto           up = [&](void*& dst, const void* src, size_t bytes) -> cudaError_t {
    cudaError_t e = cudaMalloc(&dst, bytes);
    if(e != cudaSuccess)
      return e;
    return cudaMemcpy(dst, src, bytes, cudaMemcpyHostToDevice);
  };
  CU_TRY(c, up(c->dCenters, packed.centers.data(), packed.centers.size() * 4));
  CU_TRY(c, up(c->dCov, packed.cov6.data(), packed.cov6.size() * 4));
  CU_TRY(c, up(c->dScales, packed.scales.data(), packed.scales.size() * 4));
  CU_TRY(c, up(c->dRgba, packed.rgba.data(), packed.rgba.size()));
  if(packed.shDegree)
    CU_TRY(c, up(c->dSh, packed.sh.data(), packed.sh.size()));
  c->set.centers    = static_cast<const float*>(c->dCenters);
  c->set.cov6       = static_cast<const float*>(c->dCov);
  c->set.scales     = static_cast<const float*>(c->dScales);
  c->set.rgba       = c->dRgba;
  c->set.sh         = c->dSh;
  c->set.count      = static_cast<uint32_t>(n);
  c->set.shDegree   = packed.shDegree;
  c->set.shFormat   = packed.shFormat;
  c->set.rgbaFormat = packed.rgbaFormat;
  c->paddedCount    = pad;
  c->opt            = opt;

  // sorting / raster buffers (the reference allocates its sorting buffers with the splat set too,
  // src/splat_set_manager_vk.cpp:2426-2517)
  for(int i = 0; i < 2; i++)
  {
    CU_TRY(c, cudaMalloc(&c->dKeys[i], n * sizeof(uint32_t)));
    CU_TRY(c, cudaMalloc(&c->dIds[i], n * sizeof(uint32_t)));
  }
  CU_TRY(c, cudaMalloc(&c->dRecords, n * RECORD_WORDS * sizeof(uint32_t)));
  const uint64_t preTiles = (n + PRE_TILE - 1) / PRE_TILE, binParts = (n + BIN_THREADS - 1) / BIN_THREADS,
                 sortParts = (n + SORT_PART - 1) / SORT_PART;
  CU_TRY(c, cudaMalloc(&c->dPreStatus, preTiles * sizeof(uint64_t)));
  CU_TRY(c, cudaMalloc(&c->dBinStatus, binParts * sizeof(uint64_t)));
  CU_TRY(c, cudaMalloc(&c->dSortStatus, sortParts * 256 * sizeof(uint64_t)));
  CU_TRY(c, cudaMemset(c->dPreStatus, 0, preTiles * sizeof(uint64_t)));
  CU_TRY(c, cudaMemset(c->dBinStatus, 0, binParts * sizeof(uint64_t)));
  CU_TRY(c, cudaMemset(c->dSortStatus, 0, sortParts * 256 * sizeof(uint64_t)));
  if(int rc = allocTileLists(c, std::max<uint64_t>(8 * n, 1u << 20)))
    return rc;
  c->uploaded  = true;
  c->haveFrame = false;
  return VKGS_OK;
}

int vkgs_render_async(vkgs_ctx* c, const vkgs_frame_params* fp)
{
  if(!c || !fp)
    return VKGS_ERR_INVALID_ARGUMENT;
  return enqueueFrame(c, *fp);
}

int vkgs_sync(vkgs_ctx* c)
{
  if(!c)
    return VKGS_ERR_INVALID_ARGUMENT;
  CU_TRY(c, cudaStreamSynchronize(c->stream));
  return checkOverflow(c);
}

int vkgs_last_frame_stats(vkgs_ctx* c, vkgs_outputs* out)
{
  if(!c || !out)
    return VKGS_ERR_INVALID_ARGUMENT;
  if(!c->haveFrame)
    return fail(c, VKGS_ERR_INVALID_ARGUMENT, "no frame rendered yet");
  CU_TRY(c, cudaStreamSynchronize(c->stream));
  fillStats(c, out);
  return VKGS_OK;
}

const void* vkgs_device_framebuffer(const vkgs_ctx* c)
{
  return c ? c->dImage : nullptr;
}

int vkgs_render(vkgs_ctx* c, const vkgs_frame_params* fp, vkgs_outputs* out)
{
  if(!c || !fp || !out)
    return VKGS_ERR_INVALID_ARGUMENT;
  for(int attempt = 0; attempt < 3; attempt++)
  {
    if(int rc = enqueueFrame(c, *fp))
      return rc;
    if(out->rgba)
      CU_TRY(c, cudaMemcpyAsync(out->rgba, c->dImage, sizeof(float4) * static_cast<size_t>(fp->width) * fp->height,
                                cudaMemcpyDeviceToHost, c->stream));
    CU_TRY(c, cudaStreamSynchronize(c->stream));
    const int rc = checkOverflow(c);
    if(rc == VKGS_ERR_OVERFLOW)
      continue;  // lists were regrown: run the frame again
    if(rc)
      return rc;
    fillStats(c, out);
    const uint64_t v = std::min<uint64_t>(out->visible_count, out->sorted_ids_capacity);
    const uint32_t sel = c->hCounters->sortSrc[3] & 1u;
    if(out->sorted_ids && v)
      CU_TRY(c, cudaMemcpy(out->sorted_ids, c->dIds[sel], v * sizeof(uint32_t), cudaMemcpyDeviceToHost));
    if(out->sorted_keys && v)
      CU_TRY(c, cudaMemcpy(out->sorted_keys, c->dKeys[sel], v * sizeof(uint32_t), cudaMemcpyDeviceToHost));
    return VKGS_OK;
  }
  return fail(c, VKGS_ERR_OVERFLOW, "tile lists still overflow after regrowing");
}

int vkgs_read_records(vkgs_ctx* c, uint32_t* records12, uint64_t first, uint64_t count)
{
  if(!c || !records12 || !c->uploaded || first + count > c->set.count)
    return VKGS_ERR_INVALID_ARGUMENT;
  CU_TRY(c, cudaStreamSynchronize(c->stream));
  CU_TRY(c, cudaMemcpy(records12, c->dRecords + first * RECORD_WORDS, count * RECORD_WORDS * sizeof(uint32_t), cudaMemcpyDeviceToHost));
  return VKGS_OK;
}

int vkgs_read_packed(vkgs_ctx* c, float* centers, float* cov6, float* rgba, float* sh)
{
  if(!c || !c->uploaded)
    return VKGS_ERR_INVALID_ARGUMENT;
  if(c->set.shFormat != VKGS_FORMAT_FLOAT32 || c->set.rgbaFormat != VKGS_FORMAT_FLOAT32)
    return fail(c, VKGS_ERR_UNSUPPORTED, "vkgs_read_packed needs fp32 formats");
  const uint64_t n = c->set.count;
  if(centers)
    CU_TRY(c, cudaMemcpy(centers, c->dCenters, n * 12, cudaMemcpyDeviceToHost));
  if(cov6)
    CU_TRY(c, cudaMemcpy(cov6, c->dCov, n * 24, cudaMemcpyDeviceToHost));
  if(rgba)
    CU_TRY(c, cudaMemcpy(rgba, c->dRgba, n * 16, cudaMemcpyDeviceToHost));
  if(sh && c->dSh)
    CU_TRY(c, cudaMemcpy(sh, c->dSh, n * 180, cudaMemcpyDeviceToHost));
  return VKGS_OK;
}

int vkgs_sort_pairs(vkgs_ctx* c, const uint32_t* keys, const uint32_t* values, uint64_t n, uint32_t* keysOut, uint32_t* valuesOut,
                    int repeats, float* msDevice)
{
  if(!c || (!keys && n) || (!values && n) || n > 0xfffff000ull)
    return VKGS_ERR_INVALID_ARGUMENT;
  if(msDevice)
    *msDevice = 0.0f;
  if(n == 0)
    return VKGS_OK;
  CU_TRY(c, cudaSetDevice(c->device));
  repeats = std::max(repeats, 1);
  uint32_t *dk[2] = {nullptr, nullptr}, *dv[2] = {nullptr, nullptr}, *dIn[2] = {nullptr, nullptr};
  struct Ctl
  {
    uint32_t count;
    uint32_t ticket[4];
    uint32_t hist[4][256];
  }* dCtl             = nullptr;
  uint64_t*      dSt  = nullptr;
  const uint64_t parts = (n + SORT_PART - 1) / SORT_PART;
  auto           cleanup = [&]() {
    for(int i = 0; i < 2; i++)
      freeDev(dk[i]), freeDev(dv[i]), freeDev(dIn[i]);
    freeDev(dCtl), freeDev(dSt);
  };
  cudaError_t e = cudaSuccess;
  for(int i = 0; i < 2 && e == cudaSuccess; i++)
  {
    e = cudaMalloc(&dk[i], n * 4);
    if(e == cudaSuccess)
      e = cudaMalloc(&dv[i], n * 4);
    if(e == cudaSuccess)
      e = cudaMalloc(&dIn[i], n * 4);
  }
  if(e == cudaSuccess)
    e = cudaMalloc(&dCtl, sizeof(Ctl));
  if(e == cudaSuccess)
    e = cudaMalloc(&dSt, parts * 256 * sizeof(uint64_t));
  if(e == cudaSuccess)
    e = cudaMemset(dSt, 0, parts * 256 * sizeof(uint64_t));
  if(e == cudaSuccess)
    e = cudaMemcpy(dIn[0], keys, n * 4, cudaMemcpyHostToDevice);
  if(e == cudaSuccess)
    e = cudaMemcpy(dIn[1], values, n * 4, cudaMemcpyHostToDevice);
  if(e != cudaSuccess)
  {
    cleanup();
    c->lastError = std::string("vkgs_sort_pairs alloc/upload: ") + cudaGetErrorString(e);
    return VKGS_ERR_CUDA;
  }
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0);
  cudaEventCreate(&e1);
  float total = 0.0f;
  // private epoch space: this status array is local to the call
  uint32_t epoch = 0;
  for(int rep = 0; rep < repeats; rep++)
  {
    // restore the unsorted input (not timed), then time histogram + 4 passes
    cudaMemcpyAsync(dk[0], dIn[0], n * 4, cudaMemcpyDeviceToDevice, c->stream);
    cudaMemcpyAsync(dv[0], dIn[1], n * 4, cudaMemcpyDeviceToDevice, c->stream);
    Ctl h{};
    h.count = static_cast<uint32_t>(n);
    cudaMemcpyAsync(dCtl, &h, sizeof(Ctl), cudaMemcpyHostToDevice, c->stream);
    cudaStreamSynchronize(c->stream);
    cudaEventRecord(e0, c->stream);
    launchHistogram(dk[0], &dCtl->count, static_cast<uint32_t>(n), &dCtl->hist[0][0], 0, 4, c->stream);
    c->launches++;
    for(int p = 0; p < 4; p++)
    {
      SortPassArgs sa{};
      sa.keys[0] = dk[p & 1], sa.keys[1] = dk[(p + 1) & 1];
      sa.vals[0] = dv[p & 1], sa.vals[1] = dv[(p + 1) & 1];
      sa.countPtr  = &dCtl->count;
      sa.maxCount  = static_cast<uint32_t>(n);
      sa.histogram = &dCtl->hist[p][0];
      sa.status    = dSt;
      sa.ticket    = &dCtl->ticket[p];
      sa.epoch     = ++epoch;
      sa.shift     = 8 * p;
      launchSortPass(sa, c->stream);
      c->launches++;
    }
    cudaEventRecord(e1, c->stream);
    cudaEventSynchronize(e1);
    float ms = 0.0f;
    cudaEventElapsedTime(&ms, e0, e1);
    total += ms;
  }
  cudaEventDestroy(e0);
  cudaEventDestroy(e1);
  e = cudaGetLastError();
  if(e == cudaSuccess && keysOut)
    e = cudaMemcpy(keysOut, dk[0], n * 4, cudaMemcpyDeviceToHost);
  if(e == cudaSuccess && valuesOut)
    e = cudaMemcpy(valuesOut, dv[0], n * 4, cudaMemcpyDeviceToHost);
  cleanup();
  if(e != cudaSuccess)
  {
    c->lastError = std::string("vkgs_sort_pairs: ") + cudaGetErrorString(e);
    return VKGS_ERR_CUDA;
  }
  if(msDevice)
    *msDevice = total / static_cast<float>(repeats);
  return VKGS_OK;
}

}  // extern "C"
