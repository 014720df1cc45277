/*
 * vkgs_b200.h — C ABI of the B200-native 3DGS forward rasterization path.
 *
 * Drop-in boundary for ONE path of nvpro-samples/vk_gaussian_splatting: the VK3DGSR splat
 * render call, i.e. the body of GaussianSplatting::renderHybridPipeline for PIPELINE_MESH /
 * PIPELINE_VERT (src/gaussian_splatting.cpp:494-958):
 *     updateAndUploadFrameInfoUBO  -> vkgs_frame_params (filled by the host, or by
 *                                     vkgs_frame_params_from_camera)
 *     processSortingOnGPU          -> dist/cull kernel + device radix sort   (vkgs_render)
 *     drawSplatPrimitives          -> project/SH + tile binning + tile blend (vkgs_render)
 * The reference has no FFI; its seam is C++ (nvapp::IAppElement::onRender implemented by
 * GaussianSplatting, src/gaussian_splatting.h:124-136,240-244). Every entry point below names
 * the reference interface it replaces. See INTEGRATION.md for the reference-side binding.
 *
 * Conventions: plain pointers and sizes only; return 0 (VKGS_OK) or a negative error code, never
 * an exception; one context per GPU; a context is single-caller (like the reference's render
 * thread); the context owns all device memory; output buffers are caller-owned.
 * The library REQUIRES a CUDA device for every compute entry point — there is no CPU fallback.
 */
#ifndef VKGS_B200_H_
#define VKGS_B200_H_

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#if defined(__GNUC__)
#define VKGS_API __attribute__((visibility("default")))
#else
#define VKGS_API
#endif

#define VKGS_OK 0
#define VKGS_ERR_INVALID_ARGUMENT (-1)
#define VKGS_ERR_CUDA (-2)          /* a CUDA runtime call failed (see vkgs_last_error) */
#define VKGS_ERR_NO_DEVICE (-3)     /* no usable sm_100 device: the path never falls back to CPU */
#define VKGS_ERR_NOT_UPLOADED (-4)  /* vkgs_render before vkgs_upload */
#define VKGS_ERR_OVERFLOW (-5)      /* tile-list capacity exceeded even after regrow */
#define VKGS_ERR_UNSUPPORTED (-6)   /* option outside the hot-path scope */
#define VKGS_ERR_IO (-7)            /* loader: file missing / malformed */
#define VKGS_ERR_OUT_OF_MEMORY (-8) /* a host allocation failed while packing / staging (no exception crosses the ABI) */

/* value sets = the reference's shader macros (shaders/shaderio.h:23-103) */
#define VKGS_FORMAT_FLOAT32 0
#define VKGS_FORMAT_FLOAT16 1
#define VKGS_FORMAT_UINT8 2
#define VKGS_FRUSTUM_CULLING_NONE 0
#define VKGS_FRUSTUM_CULLING_AT_DIST 1
#define VKGS_FRUSTUM_CULLING_AT_RASTER 2
#define VKGS_SIZE_CULLING_DISABLED 0
#define VKGS_SIZE_CULLING_ENABLED 1
#define VKGS_PIPELINE_3DGS 0
#define VKGS_PIPELINE_3DGUT 1
#define VKGS_CAMERA_PINHOLE 0
#define VKGS_CAMERA_FISHEYE 1
#define VKGS_EXTENT_EIGEN 0
#define VKGS_EXTENT_CONIC 1

typedef struct vkgs_ctx vkgs_ctx;

/* Host view of a splat set: the SoA layout of struct SplatSet (src/splat_set.h:33-48).
 * All pointers are HOST pointers, borrowed for the duration of the call only.
 *   positions[3N] f_dc[3N] opacity[N] (logit) scale[3N] (log) rotation[4N] (w,x,y,z)
 *   f_rest[f_rest_per_splat*N], channel-major per splat (R0..R14,G0..G14,B0..B14);
 *   f_rest_per_splat is 45 (SH degree 3) or 0 (degree 0) — the only two layouts the
 *   reference's shaders read correctly (fetchShFromBuffer uses a fixed stride of 45,
 *   shaders/threedgs_particle_buffers.h.slang:110). */
typedef struct vkgs_splat_set_view
{
  const float* positions;
  const float* f_dc;
  const float* f_rest;
  const float* opacity;
  const float* scale;
  const float* rotation;
  uint64_t     count;
  uint32_t     f_rest_per_splat;
  uint32_t     _pad;
} vkgs_splat_set_view;

/* The compile-time shader macro set of the path (src/gaussian_splatting.cpp:1653-1698). */
typedef struct vkgs_options
{
  uint32_t frustum_culling_mode;     /* FRUSTUM_CULLING_MODE, default AT_DIST (1) */
  uint32_t size_culling_mode;        /* SIZE_CULLING_MODE, default DISABLED */
  uint32_t front_to_back;            /* FRONT_TO_BACK: 0 = back-to-front "over" (reference default,
                                        alpha = sum of alphas), 1 = front-to-back "under" (alpha = 1-T) */
  uint32_t ms_antialiasing;          /* MS_ANTIALIASING (mip-splatting opacity compensation) */
  uint32_t sh_format;                /* SH_FORMAT   (VKGS_FORMAT_*) */
  uint32_t rgba_format;              /* RGBA_FORMAT (VKGS_FORMAT_*) */
  uint32_t point_cloud_mode;         /* POINT_CLOUD_MODE */
  uint32_t show_sh_only;             /* SHOW_SH_ONLY */
  uint32_t disable_opacity_gaussian; /* DISABLE_OPACITY_GAUSSIAN */
  float    transmittance_epsilon;    /* stop compositing a pixel once the remaining transmittance drops below this
                                        (0 = never, the reference's exact behaviour). Colour error bound: eps*max|rgb|;
                                        front_to_back alpha (1-T) is within eps. Back-to-front frames are composited
                                        from the near end of the list, so they can stop early too, but their alpha is
                                        the reference's SUM of opacities: with eps > 0 it only sums the fragments
                                        composited before the stop (use 0 when that channel matters). */
  uint32_t target_format;            /* colour target of the frame: VKGS_FORMAT_FLOAT32 (default; parity tests),
                                        VKGS_FORMAT_FLOAT16 (the reference's default COLOR_MAIN format,
                                        R16G16B16A16_SFLOAT, src/gaussian_splatting.h:338) or VKGS_FORMAT_UINT8
                                        (R8G8B8A8_UNORM). Blending is always fp32; the target is rounded once. */
  uint32_t surface_info;             /* NEED_SURFACE_INFO (front_to_back only, like the reference: src/gaussian_splatting.cpp:2050):
                                        the frame also produces integrated normals, picked depth + transmittance and the
                                        splat id per pixel, see vkgs_read_surface_info */
  uint32_t pipeline;                 /* VKGS_PIPELINE_3DGS (default: PIPELINE_MESH / PIPELINE_VERT, VK3DGSR) or
                                        VKGS_PIPELINE_3DGUT (PIPELINE_MESH_3DGUT, VK3DGUT: unscented-transform projection,
                                        per-fragment ray / particle evaluation; pinhole camera, <= 8 instances) */
  uint32_t extent_projection;        /* EXTENT_METHOD of the 3DGUT pipeline: VKGS_EXTENT_EIGEN or VKGS_EXTENT_CONIC (the
                                        reference's default, src/parameters.h:190); vkgs_default_options sets CONIC */
  uint32_t kernel_degree;            /* KERNEL_DEGREE of the 3DGUT particle response (shaders/shaderio.h:114-119);
                                        vkgs_default_options sets 2 (quadratic = Gaussian) */
  uint32_t camera_model;             /* CAMERA_TYPE (shaders/shaderio.h:100-101): VKGS_CAMERA_PINHOLE (default) or VKGS_CAMERA_FISHEYE
                                        (3DGUT pipeline only: equidistant fisheye projection of the sigma points, fisheye dist-stage
                                        cull, generateFisheyeRay per pixel; frame fields fov_rad and the fisheye focal, see
                                        vkgs_frame_params_set_fisheye) */
  uint32_t quantize_normals;         /* QUANTIZE_NORMALS (surface_info only; prmRaster.quantizeNormals, src/parameters.h:195, default 1):
                                        the per-splat normal travels from the mesh to the fragment stage as a 2x16-bit octahedral
                                        code (shaders/octahedral_normal.h.slang, threedgs_raster.mesh.slang:224-229, frag.slang:198-203) */
  uint32_t _reserved[1];             /* [0]: profiling flags (0 in production); bit 7 (128) = count blended fragments */
} vkgs_options;

/* Per-frame parameters: the fields of shaderio::FrameInfo the path reads
 * (shaders/shaderio.h:238-317), as filled by updateAndUploadFrameInfoUBO
 * (src/gaussian_splatting.cpp:1150-1295), plus the single splat-set instance transform
 * (SplatSetDesc.transform / transformInverse, shaders/shaderio.h:439-481).
 * Matrices are glm column-major float[16], exactly the bytes the reference uploads. */
typedef struct vkgs_frame_params
{
  float    view[16];           /* FrameInfo.viewMatrix  = glm::lookAt(eye,ctr,up) */
  float    proj[16];           /* FrameInfo.projectionMatrix = perspectiveRH_ZO, [1][1] *= -1 */
  float    model[16];          /* SplatSetDesc.transform */
  float    model_inverse[16];  /* SplatSetDesc.transformInverse */
  float    camera_position[3]; /* FrameInfo.cameraPosition (world) */
  float    focal[2];           /* FrameInfo.focal = (P00*W/2, P11*H/2); focal[1] < 0 */
  float    viewport[2];        /* FrameInfo.viewport = (W,H) */
  float    basis_viewport[2];  /* FrameInfo.basisViewport: must be (1/W,1/H) ... */
  float    inverse_focal_adjustment; /* ... and 1: other values (devicePixelRatio != 1, orthographic focal adjustment) are
                                        rejected with VKGS_ERR_UNSUPPORTED rather than silently ignored */
  float    splat_scale;              /* FrameInfo.splatScale, default 1 */
  float    frustum_dilation;         /* default 0.2 */
  float    alpha_cull_threshold;     /* default 1/255 */
  float    size_culling_min_pixels;  /* default 1 */
  uint32_t sh_degree;                /* FrameInfo.shDegree, default 3 */
  uint32_t width, height;
  float    depth_iso_threshold;      /* FrameInfo.depthIsoThreshold, default 0.7 (surface_info only) */
  float    thin_particle_threshold;  /* FrameInfo.thinParticleThreshold, default 1e-6 (surface_info only) */
  /* 3DGUT pipeline only (src/gaussian_splatting.cpp:1166-1169,1200,1254-1259; shaders/shaderio.h:241-248,271): */
  float    view_inverse[16];         /* FrameInfo.viewInverse = glm::inverse(viewMatrix) */
  float    proj_inverse[16];         /* FrameInfo.projInverse = glm::inverse(projectionMatrix) */
  float    view_quat[4];             /* FrameInfo.viewQuat = glm::quat_cast(viewMatrix) as (x,y,z,w) */
  float    view_trans[3];            /* FrameInfo.viewTrans = viewMatrix[3].xyz */
  float    near_far[2];              /* FrameInfo.nearFar = camera clip planes */
  float    alpha_clamp;              /* FrameInfo.alphaClamp, default 0.99 */
  float    kernel_min_response;      /* KERNEL_MIN_RESPONSE, default 0.0113 (src/parameters.h:216) */
  float    fov_rad;                  /* FrameInfo.fovRad = radians(vertical fov) (src/gaussian_splatting.cpp:1168); fisheye camera only */
} vkgs_frame_params;

/* The fields of struct Camera (src/camera_set.h:44-63) the pinhole path uses. */
typedef struct vkgs_camera
{
  float eye[3], ctr[3], up[3];
  float fov_deg; /* vertical */
  float znear, zfar;
} vkgs_camera;

/* Per-frame results. Mirrors what the reference exposes after a frame: COLOR_MAIN
 * (src/gaussian_splatting.h:346), the sorted index buffer, IndirectParams.instanceCount
 * (shaders/shaderio.h:343-356), and the profiler sections "GPU Dist" / "GPU Sort" /
 * "Rasterization" (src/gaussian_splatting.cpp:1324,1346,567). */
typedef struct vkgs_outputs
{
  float*    rgba;       /* HOST, W*H*4 elements of the target format (fp32 by default; fp16/u8 bit
                           patterns when options.target_format says so), row 0 = top; may be NULL */
  uint32_t* sorted_ids; /* HOST, optional, capacity sorted_ids_capacity */
  uint32_t* sorted_keys;/* HOST, optional, same capacity */
  uint64_t  sorted_ids_capacity;
  uint32_t  visible_count;   /* V = IndirectParams.instanceCount */
  uint32_t  _pad;
  uint64_t  tile_pairs;      /* (splat,tile) pairs binned this frame (implementation overhead) */
  float     ms_dist;         /* "GPU Dist"      : dist/cull + projection/SH kernel */
  float     ms_sort;         /* "GPU Sort"      : radix sort of (key,id) */
  float     ms_raster;       /* "Rasterization" : binning + tile sort + blend */
  float     ms_total;        /* first kernel to framebuffer complete (device time) */
  float     ms_kernel[16];   /* per-kernel device time, see VKGS_K_* */
  uint64_t  bytes_algorithmic; /* 12N + (132+SH(d))V + 16P, SURVEY.md §8(d) */
  /* profiling only, 0 unless options._reserved[0] & 128: (list entry, 8x8 pixel block) pairs the blend evaluated, and
   * fragments that passed both discards and were blended (the reference's ROP invocations) */
  uint64_t  list_entries_evaluated;
  uint64_t  fragments_blended;
} vkgs_outputs;

/* indices into vkgs_outputs.ms_kernel */
#define VKGS_K_PREPROCESS 0   /* dist/cull + project + SH (one fused kernel) */
#define VKGS_K_SORT_SCAN 1    /* digit histograms of the depth keys */
#define VKGS_K_SORT_PASS0 2   /* .. +3 = passes 0..3 */
#define VKGS_K_BIN_EMIT 6
#define VKGS_K_TILE_HIST 7    /* digit histograms of the tile ids */
#define VKGS_K_TILE_SORT0 8   /* .. +1 */
#define VKGS_K_TILE_RANGES 10
#define VKGS_K_BLEND 11
#define VKGS_K_COUNT 12

/* ---- lifetime (replaces GaussianSplatting::onAttach/onDetach, src/gaussian_splatting.h:124-127) */
VKGS_API int vkgs_create(int device, vkgs_ctx** out);
VKGS_API int vkgs_destroy(vkgs_ctx* ctx);
/* Make every frame's completion visible on a caller-owned CUDA stream (cudaStream_t as void*), in
 * submission order: work the caller enqueues on it afterwards sees the finished framebuffer, and
 * events recorded on it bracket the frames. NULL detaches. Frames run on the context's own streams. */
VKGS_API int vkgs_set_stream(vkgs_ctx* ctx, void* cuda_stream);
VKGS_API const char* vkgs_last_error(const vkgs_ctx* ctx);
VKGS_API const char* vkgs_version(void);
/* sizeof() of the ABI structs as compiled, for binding validation:
 * 0 vkgs_splat_set_view, 1 vkgs_options, 2 vkgs_frame_params, 3 vkgs_camera, 4 vkgs_outputs, 5 vkgs_instance, 6 vkgs_image_metrics */
VKGS_API uint32_t vkgs_abi_struct_size(int which);

/* ---- scene upload (replaces SplatSetVk::initDataStorage/initDataBuffers,
 *      src/splat_set_vk.cpp:117-170,188-480 and the sorting-buffer allocation,
 *      src/splat_set_manager_vk.cpp:2426-2517). Packs on the host exactly like the reference
 *      (cov6, clamped rgba, coefficient-major SH), copies to HBM; caller may free after return. */
VKGS_API int vkgs_upload(vkgs_ctx* ctx, const vkgs_splat_set_view* set, const vkgs_options* opt);
/* The host half of vkgs_upload alone (no GPU needed): writes the device layouts into caller
 * buffers — centers[3N] f32, cov6[6N] f32, rgba[4N] and sh[45N] in opt->rgba_format / sh_format
 * (sh may be NULL for degree-0 sets). This is SplatSetVk::initDataBuffers without the upload. */
VKGS_API int vkgs_pack_host(const vkgs_splat_set_view* set, const vkgs_options* opt, float* centers, float* cov6, void* rgba,
                            void* sh);
VKGS_API void vkgs_default_options(vkgs_options* opt);

/* ---- multi-instance scenes (replaces SplatSetManagerVk's splat sets + instances and its global
 *      index table: createInstance / rebuildGlobalIndexTables, src/splat_set_manager_vk.cpp:2304-2360,
 *      GlobalSplatIndexEntry shaders/shaderio.h:523-527, resolveGlobalSplatID
 *      shaders/threedgs_particle_storage.h.slang:34-42). An instance places one splat set in the world
 *      with its own transform (SplatSetDesc.transform / transformInverse, glm column-major, both
 *      supplied by the caller like the reference keeps both); several instances may share a set.
 *      Global splat id = (splat count of all earlier instances) + local id: the ids the sort
 *      returns and vkgs_read_records is indexed by. All instances are culled, sorted and blended
 *      together. With a scene uploaded this way vkgs_frame_params.model / model_inverse are ignored. */
typedef struct vkgs_instance
{
  uint32_t splat_set_index; /* index into the `sets` array of vkgs_upload_scene */
  uint32_t _pad;
  float    transform[16];
  float    transform_inverse[16];
} vkgs_instance;
VKGS_API int vkgs_upload_scene(vkgs_ctx* ctx, const vkgs_splat_set_view* sets, uint32_t set_count, const vkgs_instance* instances,
                               uint32_t instance_count, const vkgs_options* opt);
/* Move an instance (takes effect for frames enqueued afterwards; nothing is re-uploaded). */
VKGS_API int vkgs_set_instance_transform(vkgs_ctx* ctx, uint32_t instance, const float* transform, const float* transform_inverse);
/* The global index table as the reference builds it: for every global splat id the instance
 * index and the splat index inside that instance's set (either array may be NULL; `total`
 * receives the global splat count). */
VKGS_API int vkgs_global_index_table(const vkgs_ctx* ctx, uint32_t* instance_index, uint32_t* splat_index, uint64_t capacity,
                                     uint64_t* total);

/* ---- frame (replaces updateAndUploadFrameInfoUBO + processSortingOnGPU + drawSplatPrimitives,
 *      src/gaussian_splatting.cpp:1150,1298,1369). */
VKGS_API int vkgs_frame_params_from_camera(const vkgs_camera* cam, uint32_t width, uint32_t height, vkgs_frame_params* out);
VKGS_API void vkgs_default_camera(vkgs_camera* cam);
/* Fisheye camera on the 3DGUT pipeline: FrameInfo.focal = (1,-1) * viewport / fovRad
 * (src/gaussian_splatting.cpp:1239-1243). Call after vkgs_frame_params_from_camera. */
VKGS_API void vkgs_frame_params_set_fisheye(vkgs_frame_params* fp);
/* Synchronous: returns after the frame (and the requested copies to host) completed. */
VKGS_API int vkgs_render(vkgs_ctx* ctx, const vkgs_frame_params* fp, vkgs_outputs* out);
/* The reference's CPU-sorting mode (SORTING_CPU_ASYNC_MULTI): the frame is drawn from an index buffer sorted elsewhere
 * — SplatSorterAsync on the host in the reference (src/splat_sorter_async.cpp:92-141), consumed and uploaded by
 * SplatSetManagerVk::tryConsumeAndUploadCpuSortingResult (src/splat_set_manager_vk.cpp:3334-3416). `ids` = `count` HOST
 * global splat ids in draw order (count <= splats of the scene; borrowed for the call). No dist stage and no sort run:
 * every id is drawn, frustum culling moves to the raster stage and size culling is off, exactly as the reference's UI
 * forces them in this mode (src/gaussian_splatting_ui.cpp:1468-1490). The blend operator still follows
 * options.front_to_back. Synchronous; out->sorted_ids (if set) receives the ids back, sorted_keys is not written. */
VKGS_API int vkgs_render_presorted(vkgs_ctx* ctx, const vkgs_frame_params* fp, const uint32_t* ids, uint64_t count, vkgs_outputs* out);
/* Stream-ordered: enqueue one frame, result stays in the device framebuffer. Up to four frames are
 * in flight, each on its own pair of internal streams (the frames-in-flight of the reference's swapchain loop,
 * nvpro_core2/nvapp/application.cpp:517-548); completion order == submission order. The call returns at once unless
 * `frames in flight` frames are already queued: then it waits for the oldest one (whose slot it reuses), like the fence
 * wait of that loop. */
VKGS_API int vkgs_render_async(vkgs_ctx* ctx, const vkgs_frame_params* fp);
/* Same, plus an asynchronous copy of the finished RGBA frame (W*H*4 elements of the target
 * format) to PINNED host memory; valid after vkgs_sync (or once the caller's stream reaches it). */
VKGS_API int vkgs_render_to_host_async(vkgs_ctx* ctx, const vkgs_frame_params* fp, void* host_rgba);
/* Change the colour target format (VKGS_FORMAT_*) for subsequent frames without re-uploading. */
VKGS_API int vkgs_set_target_format(vkgs_ctx* ctx, uint32_t target_format);
/* 1 = strictly one frame at a time (full-occupancy kernels, lowest latency), 2..4 (default 4) = overlap consecutive frames. */
VKGS_API int vkgs_set_frames_in_flight(vkgs_ctx* ctx, int frames);
/* Wait for every frame in flight. A frame whose tile lists overflowed is handled out of the caller's sight — here, or
 * at the moment its slot is reused by a later vkgs_render_async / vkgs_render_to_host_async: the lists are grown and the
 * frame is rendered again on its own slot with the same parameters and host destination, so after a successful vkgs_sync
 * every frame enqueued since the previous one is complete. VKGS_ERR_OVERFLOW is left for lists that still overflow after
 * four regrowths (the 2^32-pair limit) and for caller-ordered frames. */
VKGS_API int vkgs_sync(vkgs_ctx* ctx);
/* Enable per-kernel cudaEvent timing for subsequent frames (off by default). */
VKGS_API int vkgs_set_profiling(vkgs_ctx* ctx, int enabled);
/* Stats of the most recent frame (syncs). Fills everything in vkgs_outputs except the buffers. */
VKGS_API int vkgs_last_frame_stats(vkgs_ctx* ctx, vkgs_outputs* out);
/* Device pointer of the RGBA framebuffer of the last frame: W*H*4 elements of the colour target format in use
 * (fp32 by default; half / unorm8 bit patterns after vkgs_set_target_format or options.target_format). */
VKGS_API const void* vkgs_device_framebuffer(const vkgs_ctx* ctx);
/* Number of kernels launched by this context since creation. */
VKGS_API uint64_t vkgs_launch_count(const vkgs_ctx* ctx);

/* ---- stand-alone key/value radix sort (replaces vrdxCmdSortKeyValueIndirect,
 *      3rdparty/vrdx/src/vk_radix_sort.cc:249-258): stable ascending LSD sort of n (u32 key,
 *      u32 value) pairs. Host in, host out; ms_device = device time of the sort alone,
 *      averaged over `repeats` runs (>=1) on the same input. */
VKGS_API int vkgs_sort_pairs(vkgs_ctx* ctx, const uint32_t* keys, const uint32_t* values, uint64_t n,
                    uint32_t* keys_out, uint32_t* values_out, int repeats, float* ms_device);

/* The same sort on DEVICE buffers, the way the reference records it: keys and values are sorted IN PLACE, the pair count is
 * read on the device from *d_count (<= max_count; vrdx's indirect buffer), `d_storage` is caller-provided scratch of at least
 * vkgs_sort_pairs_storage_bytes(max_count) bytes (vrdxGetSorterKeyValueStorageRequirements,
 * 3rdparty/vrdx/src/vk_radix_sort.cc:209-224). Stream-ordered on `cuda_stream` (cudaStream_t as void*; NULL = the context's
 * own stream): nothing is synchronised, like a recorded vkCmd. */
VKGS_API uint64_t vkgs_sort_pairs_storage_bytes(uint64_t max_count);
VKGS_API int vkgs_sort_pairs_device(vkgs_ctx* ctx, uint32_t* d_keys, uint32_t* d_values, const uint32_t* d_count, uint64_t max_count,
                                    void* d_storage, uint64_t storage_bytes, void* cuda_stream);

/* ---- image comparison metrics (replaces ImageCompare's metrics pass: shaders/image_compare_metric.comp.slang:84-190,
 *      371-477, dispatch src/image_compare.cpp:770-830, read-back src/image_compare.cpp:874-905).
 *      MSE over RGB with the reference's fixed-point accumulation (uint32, x1e9, normalised by W*H*3),
 *      PSNR = min(10 log10(1/MSE), 99.99) dB, and the reference's fast FLIP approximation (YCxCz colour
 *      error + Sobel feature error, Minkowski pooling q = 3) or its "reference" mode (multi-scale band-pass features). */
#define VKGS_FLIP_DISABLED 0
#define VKGS_FLIP_APPROX 1
#define VKGS_FLIP_REFERENCE 2 /* five band-pass features per image, brute-force Gaussian windows up to 131x131 taps: slow by design */
typedef struct vkgs_image_metrics
{
  float    mse, psnr, flip;
  uint32_t mse_fixed, flip_fixed; /* the raw accumulators of the reference's result buffer ([0] and [8]) */
  float    ms_device;             /* device time of the metrics kernel */
} vkgs_image_metrics;
/* Keep a copy of the last rendered frame as the "capture image" (fp32 colour target). */
VKGS_API int vkgs_capture_frame(vkgs_ctx* ctx);
/* Metrics between the capture and the last rendered frame (same size). */
VKGS_API int vkgs_compare_with_capture(vkgs_ctx* ctx, uint32_t flip_mode, vkgs_image_metrics* out);
/* The same kernel on two HOST images (W*H*4 fp32 each): reference = capture, current. */
VKGS_API int vkgs_image_metrics_host(vkgs_ctx* ctx, const float* reference, const float* current, uint32_t width, uint32_t height,
                                     uint32_t flip_mode, vkgs_image_metrics* out);

/* ---- surface-info side outputs of the last frame (options.surface_info; replaces the extra attachments of the
 *      FTB raster pass: COLOR_RASTER_NORMAL, COLOR_RASTER_DEPTH, COLOR_RASTER_SPLATID, src/gaussian_splatting.cpp:594-617,
 *      664-668; written by threedgs_raster.frag.slang:316-350 and, per splat, threedgs_raster.mesh.slang:209-233).
 *      HOST buffers, any may be NULL:
 *        normals            W*H*4 f32  sum over fragments of normalWorld*opacity under the "under" operator, a = 1 - T
 *        depth_transmittance W*H*2 f32 (NDC depth of the first fragment after which T < depth_iso_threshold, 0 if none; T)
 *        splat_id           W*H u32    global id of the last blended fragment, 0xffffffff where nothing was blended
 *      (with transmittance_epsilon > 0 the list is cut short: ids / T are exact only for epsilon = 0) */
VKGS_API int vkgs_read_surface_info(vkgs_ctx* ctx, float* normals, float* depth_transmittance, uint32_t* splat_id);
/* The QUANTIZE_NORMALS round trip (encodeNormalOctahedral -> decodeNormalOctahedral, shaders/octahedral_normal.h.slang) on
 * `count` unit normals, host arrays [3*count]; host only, no GPU needed (the kernel runs the same function). */
VKGS_API int vkgs_quantize_normals_host(const float* normals_in, float* normals_out, uint64_t count);
/* Host instantiations of the fixed-sequence fp32 elementary functions the kernels run at their discard / cull decisions
 * (the reference's are driver intrinsics): which = 0 exp(a); 1 atan2(a = y > 0, b = x); 2 acos(a); 3 sin(a) -> out,
 * cos(a) -> out2. Host only, no GPU needed; pinned bit for bit against the oracle in tests/test_oracle_kat.py. */
VKGS_API int vkgs_exact_math_host(uint32_t which, const float* a, const float* b, float* out, float* out2, uint64_t count);

/* ---- parity/debug read-backs of per-splat intermediates of the last frame ----------------
 * Per-splat record, indexed by splat id (only ids that passed the dist-stage cull are valid):
 *   12 x 32-bit words: centre px (x,y), fragPos basis w1 (x,y), w2 (x,y), rgba (r,g,b,a),
 *   pixel bbox packed (x0 | y0<<16), (x1 | y1<<16); an empty bbox (x1<x0) = rejected splat. */
VKGS_API int vkgs_read_records(vkgs_ctx* ctx, uint32_t* records12, uint64_t first, uint64_t count);
/* Packed device arrays as uploaded (fp32 formats): centers[3N], cov6[6N], rgba[4N], sh[45N]. */
VKGS_API int vkgs_read_packed(vkgs_ctx* ctx, float* centers, float* cov6, float* rgba, float* sh);

/* ---- scene files -> SplatSet layout (replaces PlyLoaderAsync::innerLoad,
 *      src/ply_loader_async.cpp:291-453). Format by extension, like the reference: ".splat"
 *      (antimatter15 32-byte records), ".spz" (Niantic, gzip), anything else is parsed as INRIA
 *      .ply (ascii / binary_little_endian / binary_big_endian). Arrays come out exactly as the
 *      reference hands them to SplatSetVk: RUB coordinates, rotation (w,x,y,z), f_rest channel-major.
 *      Host only, no GPU needed. */
typedef struct vkgs_scene vkgs_scene;
VKGS_API int vkgs_scene_load(const char* path, vkgs_scene** out);
VKGS_API int vkgs_scene_view(const vkgs_scene* scene, vkgs_splat_set_view* view); /* borrowed pointers into the scene */
VKGS_API int vkgs_scene_free(vkgs_scene* scene);
VKGS_API const char* vkgs_scene_load_error(void); /* text of the last VKGS_ERR_IO on this thread */

/* ---- deterministic synthetic scene generator; host only, no GPU needed.
 *      Arrays are caller-allocated with the vkgs_splat_set_view sizes; sh_degree is 0 or 3. */
VKGS_API int vkgs_synth_scene(uint64_t n, uint32_t sh_degree, uint64_t seed, float* positions, float* f_dc, float* f_rest,
                     float* opacity, float* scale, float* rotation);

#ifdef __cplusplus
}
#endif
#endif /* VKGS_B200_H_ */
