// host_camera.cpp — host side of the per-frame parameter block.
//
// Mirrors what the reference computes on the render thread before recording the frame:
//   CameraManipulator::updateLookatMatrix       nvpro_core2/nvutils/camera_manipulator.cpp:211
//   CameraManipulator::getPerspectiveMatrix     nvpro_core2/nvutils/camera_manipulator.hpp:217-233
//   GaussianSplatting::updateAndUploadFrameInfoUBO   src/gaussian_splatting.cpp:1150-1295
// glm is not a dependency: lookAtRH / perspectiveRH_ZO are evaluated with glm's operation order so
// the matrices are bit-identical to the ones the reference uploads (tests/golden/glm_golden.json).
#include <cmath>
#include <cstring>

#include "vkgs_b200.h"

namespace {

struct V3
{
  float x, y, z;
};
inline V3    sub(V3 a, V3 b) { return {a.x - b.x, a.y - b.y, a.z - b.z}; }
inline float dot(V3 a, V3 b) { return a.x * b.x + a.y * b.y + a.z * b.z; }
inline V3    cross(V3 a, V3 b) { return {a.y * b.z - b.y * a.z, a.z * b.x - b.z * a.x, a.x * b.y - b.x * a.y}; }
inline V3    normalize(V3 v)
{
  const float inv = 1.0f / std::sqrt(dot(v, v));
  return {v.x * inv, v.y * inv, v.z * inv};
}

void setIdentity(float* m)
{
  std::memset(m, 0, 16 * sizeof(float));
  m[0] = m[5] = m[10] = m[15] = 1.0f;
}

}  // namespace

extern "C" void vkgs_default_camera(vkgs_camera* cam)
{
  // struct Camera defaults, src/camera_set.h:48-53
  if(!cam)
    return;
  cam->eye[0] = 1.7f, cam->eye[1] = 1.5f, cam->eye[2] = 1.7f;
  cam->ctr[0] = cam->ctr[1] = cam->ctr[2] = 0.0f;
  cam->up[0] = 0.0f, cam->up[1] = 1.0f, cam->up[2] = 0.0f;
  cam->fov_deg = 60.0f;
  cam->znear   = 0.1f;
  cam->zfar    = 2000.0f;
}

extern "C" void vkgs_default_options(vkgs_options* opt)
{
  if(!opt)
    return;
  std::memset(opt, 0, sizeof(*opt));
  opt->frustum_culling_mode = VKGS_FRUSTUM_CULLING_AT_DIST;  // src/parameters.h RasterParameters default
  opt->size_culling_mode    = VKGS_SIZE_CULLING_DISABLED;
  opt->front_to_back        = 0;  // back-to-front unless surface info is needed (gaussian_splatting.cpp:509)
  opt->ms_antialiasing      = 0;
  opt->sh_format            = VKGS_FORMAT_FLOAT32;
  opt->rgba_format          = VKGS_FORMAT_FLOAT32;
  opt->transmittance_epsilon = 0.0f;
}

extern "C" int vkgs_frame_params_from_camera(const vkgs_camera* cam, uint32_t width, uint32_t height, vkgs_frame_params* out)
{
  if(!cam || !out || width == 0 || height == 0)
    return VKGS_ERR_INVALID_ARGUMENT;
  std::memset(out, 0, sizeof(*out));

  // view = glm::lookAt(eye, ctr, up)
  const V3 eye{cam->eye[0], cam->eye[1], cam->eye[2]};
  const V3 ctr{cam->ctr[0], cam->ctr[1], cam->ctr[2]};
  const V3 up{cam->up[0], cam->up[1], cam->up[2]};
  const V3 f = normalize(sub(ctr, eye));
  const V3 s = normalize(cross(f, up));
  const V3 u = cross(s, f);
  float*   V = out->view;
  setIdentity(V);
  V[0] = s.x, V[4] = s.y, V[8] = s.z;
  V[1] = u.x, V[5] = u.y, V[9] = u.z;
  V[2] = -f.x, V[6] = -f.y, V[10] = -f.z;
  V[12] = -dot(s, eye);
  V[13] = -dot(u, eye);
  V[14] = dot(f, eye);

  // proj = glm::perspectiveRH_ZO(radians(fov), W/H, near, far); proj[1][1] *= -1
  const float aspect      = static_cast<float>(width) / static_cast<float>(height);
  const float fovy        = cam->fov_deg * 0.01745329251994329576923690768489f;
  const float tanHalfFovy = std::tan(fovy / 2.0f);
  float*      P           = out->proj;
  P[0]                    = 1.0f / (aspect * tanHalfFovy);
  P[5]                    = 1.0f / (tanHalfFovy);
  P[10]                   = cam->zfar / (cam->znear - cam->zfar);
  P[11]                   = -1.0f;
  P[14]                   = -(cam->zfar * cam->znear) / (cam->zfar - cam->znear);
  P[5] *= -1.0f;

  setIdentity(out->model);
  setIdentity(out->model_inverse);
  out->camera_position[0] = eye.x, out->camera_position[1] = eye.y, out->camera_position[2] = eye.z;

  const float devicePixelRatio = 1.0f;
  const float rw = static_cast<float>(width), rh = static_cast<float>(height);
  out->focal[0]                 = P[0] * 0.5f * devicePixelRatio * rw;  // gaussian_splatting.cpp:1248
  out->focal[1]                 = P[5] * 0.5f * devicePixelRatio * rh;  // :1249
  out->viewport[0]              = rw * devicePixelRatio;
  out->viewport[1]              = rh * devicePixelRatio;
  out->basis_viewport[0]        = 1.0f / rw;
  out->basis_viewport[1]        = 1.0f / rh;
  out->inverse_focal_adjustment = 1.0f;
  // FrameInfo defaults, shaders/shaderio.h:252-259
  out->splat_scale             = 1.0f;
  out->frustum_dilation        = 0.2f;
  out->alpha_cull_threshold    = 1.0f / 255.0f;
  out->size_culling_min_pixels = 1.0f;
  out->sh_degree               = 3;
  out->width                   = width;
  out->height                  = height;
  out->depth_iso_threshold     = 0.7f;   // shaders/shaderio.h:311
  out->thin_particle_threshold = 1e-6f;  // shaders/shaderio.h:316
  return VKGS_OK;
}
