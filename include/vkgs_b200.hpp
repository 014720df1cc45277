/*
 * vkgs_b200.hpp — header-only C++17 host layer over the C ABI (vkgs_b200.h), with the names of the
 * reference's own interface for this path, so that a host written against the reference reads the same:
 *
 *   reference (nvpro-samples/vk_gaussian_splatting)                      here
 *   ---------------------------------------------------------------     ---------------------------------------
 *   struct SplatSet (src/splat_set.h:33-48)                              vkgs_b200::SplatSet (same member names)
 *   PlyLoaderAsync::innerLoad (src/ply_loader_async.cpp:291-453)         SplatSet::loadFromFile
 *   GaussianSplatting::onAttach / onDetach (src/gaussian_splatting.h:124-127)   GaussianSplatting::onAttach / onDetach
 *   IAppElement::onResize (nvpro_core2/nvapp/application.hpp:120-135)    GaussianSplatting::onResize
 *   SplatSetVk::initDataStorage (src/splat_set_vk.cpp:117-170)           GaussianSplatting::initDataStorage
 *   updateAndUploadFrameInfoUBO (src/gaussian_splatting.cpp:1150-1295)   GaussianSplatting::updateAndUploadFrameInfoUBO
 *   onRender -> processSortingOnGPU + drawSplatPrimitives (:335,:1298,:1369)   GaussianSplatting::onRender
 *   prmRaster / prmData / prmRtx (src/parameters.h)                      GaussianSplatting::prm (vkgs_options)
 *   prmFrame (shaderio::FrameInfo)                                       GaussianSplatting::prmFrame (vkgs_frame_params)
 *   tryConsumeAndUploadCpuSortingResult (src/splat_set_manager_vk.cpp:3334-3416)   GaussianSplatting::onRenderCpuSorted
 *   vrdxCmdSortKeyValueIndirect (3rdparty/vrdx/src/vk_radix_sort.cc:249-258)       GaussianSplatting::cmdSortKeyValueIndirect
 *
 * Error behaviour follows the reference's convention (bool + a logged message): every call returns false on
 * failure and lastError() holds the text; nothing throws. There is no CPU fallback: without an sm_100 device
 * onAttach fails with VKGS_ERR_NO_DEVICE.
 */
#ifndef VKGS_B200_HPP_
#define VKGS_B200_HPP_

#include <cstdint>
#include <cstring>
#include <string>
#include <vector>

#include "vkgs_b200.h"

namespace vkgs_b200 {

/* Host scene, the SoA layout of the reference's SplatSet (positions[3N] f_dc[3N] f_rest[45N or 0] opacity[N] (logit)
 * scale[3N] (log) rotation[4N] (w,x,y,z)). */
struct SplatSet
{
  std::vector<float> positions, f_dc, f_rest, opacity, scale, rotation;

  size_t size() const { return positions.size() / 3; }
  /* 45 f_rest values per splat = SH degree 3, none = degree 0 (the two layouts the reference's shaders read) */
  uint32_t maxShDegree() const { return (size() && f_rest.size() == 45 * size()) ? 3u : 0u; }

  vkgs_splat_set_view view() const
  {
    vkgs_splat_set_view v{};
    v.positions = positions.data(), v.f_dc = f_dc.data(), v.f_rest = f_rest.empty() ? nullptr : f_rest.data();
    v.opacity = opacity.data(), v.scale = scale.data(), v.rotation = rotation.data();
    v.count            = size();
    v.f_rest_per_splat = size() ? static_cast<uint32_t>(f_rest.size() / size()) : 0u;
    return v;
  }

  /* .ply / .spz / .splat by extension, like PlyLoaderAsync; arrays come out in RUB coordinates */
  bool loadFromFile(const std::string& path, std::string* error = nullptr)
  {
    vkgs_scene* scene = nullptr;
    if(vkgs_scene_load(path.c_str(), &scene) != VKGS_OK)
    {
      if(error)
        *error = vkgs_scene_load_error();
      return false;
    }
    vkgs_splat_set_view v{};
    vkgs_scene_view(scene, &v);
    const size_t n = v.count;
    positions.assign(v.positions, v.positions + 3 * n), f_dc.assign(v.f_dc, v.f_dc + 3 * n);
    if(v.f_rest)
      f_rest.assign(v.f_rest, v.f_rest + static_cast<size_t>(v.f_rest_per_splat) * n);
    else
      f_rest.clear();
    opacity.assign(v.opacity, v.opacity + n), scale.assign(v.scale, v.scale + 3 * n), rotation.assign(v.rotation, v.rotation + 4 * n);
    vkgs_scene_free(scene);
    return true;
  }

  /* the deterministic synthetic scene of SURVEY.md 8(d) */
  bool synthesize(uint64_t n, uint32_t shDegree, uint64_t seed)
  {
    positions.resize(3 * n), f_dc.resize(3 * n), f_rest.resize(shDegree ? 45 * n : 0), opacity.resize(n), scale.resize(3 * n), rotation.resize(4 * n);
    return vkgs_synth_scene(n, shDegree, seed, positions.data(), f_dc.data(), f_rest.empty() ? nullptr : f_rest.data(), opacity.data(),
                            scale.data(), rotation.data())
           == VKGS_OK;
  }
};

class GaussianSplatting
{
public:
  vkgs_options      prm{};       /* the shader macro set (prmRaster / prmData / prmRtx of the reference) */
  vkgs_frame_params prmFrame{};  /* shaderio::FrameInfo of the current frame */

  GaussianSplatting() { vkgs_default_options(&prm); }
  ~GaussianSplatting() { onDetach(); }
  GaussianSplatting(const GaussianSplatting&)            = delete;
  GaussianSplatting& operator=(const GaussianSplatting&) = delete;

  const std::string& lastError() const { return m_error; }
  vkgs_ctx*          context() const { return m_ctx; }

  bool onAttach(int cudaDevice)
  {
    onDetach();
    const int rc = vkgs_create(cudaDevice, &m_ctx);
    if(rc != VKGS_OK)
    {
      m_error = "vkgs_create failed with code " + std::to_string(rc) + (rc == VKGS_ERR_NO_DEVICE ? " (no sm_100 CUDA device; there is no CPU fallback)" : "");
      m_ctx   = nullptr;
      return false;
    }
    return true;
  }
  void onDetach()
  {
    if(m_ctx)
      vkgs_destroy(m_ctx);
    m_ctx = nullptr;
  }
  void onResize(uint32_t width, uint32_t height) { m_width = width, m_height = height; }

  /* RAM -> VRAM: packs like SplatSetVk::initDataBuffers and allocates the sorting buffers; `set` may be freed afterwards */
  bool initDataStorage(const SplatSet& set)
  {
    const vkgs_splat_set_view v = set.view();
    return check(vkgs_upload(m_ctx, &v, &prm));
  }
  /* several splat sets + instances (SplatSetManagerVk) */
  bool initDataStorage(const std::vector<const SplatSet*>& sets, const std::vector<vkgs_instance>& instances)
  {
    std::vector<vkgs_splat_set_view> views;
    for(const SplatSet* s : sets)
      views.push_back(s->view());
    return check(vkgs_upload_scene(m_ctx, views.data(), static_cast<uint32_t>(views.size()), instances.data(),
                                   static_cast<uint32_t>(instances.size()), &prm));
  }

  /* fills prmFrame from the camera (glm::lookAt / perspectiveRH_ZO arithmetic) at the current size; the caller may
   * then override any FrameInfo field (splat_scale, sh_degree, model, ...) before onRender */
  bool updateAndUploadFrameInfoUBO(const vkgs_camera& camera)
  {
    if(!check(vkgs_frame_params_from_camera(&camera, m_width, m_height, &prmFrame)))
      return false;
    if(prm.camera_model == VKGS_CAMERA_FISHEYE)
      vkgs_frame_params_set_fisheye(&prmFrame);
    return true;
  }

  /* one frame: dist/cull + sort (processSortingOnGPU) and projection + raster + blend (drawSplatPrimitives).
   * rgbaOut: host, W*H*4 elements of prm.target_format, may be null (frame stays on the device) */
  bool onRender(void* rgbaOut, vkgs_outputs* stats = nullptr)
  {
    vkgs_outputs o{};
    o.rgba        = static_cast<float*>(rgbaOut);
    const bool ok = check(vkgs_render(m_ctx, &prmFrame, &o));
    if(stats)
      *stats = o;
    return ok;
  }
  /* SORTING_CPU_ASYNC_MULTI: the frame drawn from an index buffer sorted on the host (SplatSorterAsync), every id in the
   * caller's order, frustum culling at the raster stage */
  bool onRenderCpuSorted(const std::vector<uint32_t>& sortedIndices, void* rgbaOut, vkgs_outputs* stats = nullptr)
  {
    vkgs_outputs o{};
    o.rgba        = static_cast<float*>(rgbaOut);
    const bool ok = check(vkgs_render_presorted(m_ctx, &prmFrame, sortedIndices.data(), sortedIndices.size(), &o));
    if(stats)
      *stats = o;
    return ok;
  }
  /* the vrdx entry point on device buffers: in place, count read on the device, caller-provided storage
   * (sortStorageBytes), stream-ordered on `cudaStream` (nullptr = the context's own stream) */
  static uint64_t sortStorageBytes(uint64_t maxCount) { return vkgs_sort_pairs_storage_bytes(maxCount); }
  bool cmdSortKeyValueIndirect(uint32_t* dKeys, uint32_t* dValues, const uint32_t* dCount, uint64_t maxCount, void* dStorage,
                               uint64_t storageBytes, void* cudaStream = nullptr)
  {
    return check(vkgs_sort_pairs_device(m_ctx, dKeys, dValues, dCount, maxCount, dStorage, storageBytes, cudaStream));
  }
  /* the frames-in-flight loop of nvapp::Application: enqueue, then sync() */
  bool onRenderAsync(void* pinnedRgbaOut = nullptr)
  {
    return check(pinnedRgbaOut ? vkgs_render_to_host_async(m_ctx, &prmFrame, pinnedRgbaOut) : vkgs_render_async(m_ctx, &prmFrame));
  }
  bool sync() { return check(vkgs_sync(m_ctx)); }
  bool lastFrameStats(vkgs_outputs* out) { return check(vkgs_last_frame_stats(m_ctx, out)); }

private:
  bool check(int rc)
  {
    if(rc == VKGS_OK)
      return true;
    const char* text = m_ctx ? vkgs_last_error(m_ctx) : nullptr;
    m_error          = (text && *text) ? text : ("vkgs error " + std::to_string(rc));
    return false;
  }

  vkgs_ctx*   m_ctx   = nullptr;
  uint32_t    m_width = 1920, m_height = 1080;
  std::string m_error;
};

}  // namespace vkgs_b200
#endif /* VKGS_B200_HPP_ */
