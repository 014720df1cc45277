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

// glm::inverse(mat4) with glm's operation order (glm/detail/func_matrix.inl, compute_inverse<4,4>);
// m and out are column-major float[16], m[c][r] = m[4*c + r].
void inverse4(const float* a, float* out)
{
#define M(c, r) a[4 * (c) + (r)]
  const float Coef00 = M(2, 2) * M(3, 3) - M(3, 2) * M(2, 3), Coef02 = M(1, 2) * M(3, 3) - M(3, 2) * M(1, 3), Coef03 = M(1, 2) * M(2, 3) - M(2, 2) * M(1, 3);
  const float Coef04 = M(2, 1) * M(3, 3) - M(3, 1) * M(2, 3), Coef06 = M(1, 1) * M(3, 3) - M(3, 1) * M(1, 3), Coef07 = M(1, 1) * M(2, 3) - M(2, 1) * M(1, 3);
  const float Coef08 = M(2, 1) * M(3, 2) - M(3, 1) * M(2, 2), Coef10 = M(1, 1) * M(3, 2) - M(3, 1) * M(1, 2), Coef11 = M(1, 1) * M(2, 2) - M(2, 1) * M(1, 2);
  const float Coef12 = M(2, 0) * M(3, 3) - M(3, 0) * M(2, 3), Coef14 = M(1, 0) * M(3, 3) - M(3, 0) * M(1, 3), Coef15 = M(1, 0) * M(2, 3) - M(2, 0) * M(1, 3);
  const float Coef16 = M(2, 0) * M(3, 2) - M(3, 0) * M(2, 2), Coef18 = M(1, 0) * M(3, 2) - M(3, 0) * M(1, 2), Coef19 = M(1, 0) * M(2, 2) - M(2, 0) * M(1, 2);
  const float Coef20 = M(2, 0) * M(3, 1) - M(3, 0) * M(2, 1), Coef22 = M(1, 0) * M(3, 1) - M(3, 0) * M(1, 1), Coef23 = M(1, 0) * M(2, 1) - M(2, 0) * M(1, 1);
  const float Fac0[4] = {Coef00, Coef00, Coef02, Coef03}, Fac1[4] = {Coef04, Coef04, Coef06, Coef07}, Fac2[4] = {Coef08, Coef08, Coef10, Coef11};
  const float Fac3[4] = {Coef12, Coef12, Coef14, Coef15}, Fac4[4] = {Coef16, Coef16, Coef18, Coef19}, Fac5[4] = {Coef20, Coef20, Coef22, Coef23};
  const float Vec0[4] = {M(1, 0), M(0, 0), M(0, 0), M(0, 0)}, Vec1[4] = {M(1, 1), M(0, 1), M(0, 1), M(0, 1)};
  const float Vec2[4] = {M(1, 2), M(0, 2), M(0, 2), M(0, 2)}, Vec3[4] = {M(1, 3), M(0, 3), M(0, 3), M(0, 3)};
  const float SignA[4] = {+1, -1, +1, -1}, SignB[4] = {-1, +1, -1, +1};
  float       inv[16];
  for(int k = 0; k < 4; k++)
  {
    inv[4 * 0 + k] = ((Vec1[k] * Fac0[k] - Vec2[k] * Fac1[k]) + Vec3[k] * Fac2[k]) * SignA[k];
    inv[4 * 1 + k] = ((Vec0[k] * Fac0[k] - Vec2[k] * Fac3[k]) + Vec3[k] * Fac4[k]) * SignB[k];
    inv[4 * 2 + k] = ((Vec0[k] * Fac1[k] - Vec1[k] * Fac3[k]) + Vec3[k] * Fac5[k]) * SignA[k];
    inv[4 * 3 + k] = ((Vec0[k] * Fac2[k] - Vec1[k] * Fac4[k]) + Vec2[k] * Fac5[k]) * SignB[k];
  }
  const float d0 = M(0, 0) * inv[0], d1 = M(0, 1) * inv[4], d2 = M(0, 2) * inv[8], d3 = M(0, 3) * inv[12];
  const float oneOverDet = 1.0f / ((d0 + d1) + (d2 + d3));
  for(int k = 0; k < 16; k++)
    out[k] = inv[k] * oneOverDet;
#undef M
}

// glm::quat_cast(mat3(m)) (glm/gtc/quaternion.inl:81-123), returned as (x, y, z, w)
void quatCast(const float* a, float q[4])
{
#define M(c, r) a[4 * (c) + (r)]
  const float fourX = M(0, 0) - M(1, 1) - M(2, 2), fourY = M(1, 1) - M(0, 0) - M(2, 2), fourZ = M(2, 2) - M(0, 0) - M(1, 1),
              fourW = M(0, 0) + M(1, 1) + M(2, 2);
  int   biggest = 0;
  float big     = fourW;
  if(fourX > big)
    big = fourX, biggest = 1;
  if(fourY > big)
    big = fourY, biggest = 2;
  if(fourZ > big)
    big = fourZ, biggest = 3;
  const float val = std::sqrt(big + 1.0f) * 0.5f, mult = 0.25f / val;
  float       w, x, y, z;
  switch(biggest)
  {
    case 0:
      w = val, x = (M(1, 2) - M(2, 1)) * mult, y = (M(2, 0) - M(0, 2)) * mult, z = (M(0, 1) - M(1, 0)) * mult;
      break;
    case 1:
      w = (M(1, 2) - M(2, 1)) * mult, x = val, y = (M(0, 1) + M(1, 0)) * mult, z = (M(2, 0) + M(0, 2)) * mult;
      break;
    case 2:
      w = (M(2, 0) - M(0, 2)) * mult, x = (M(0, 1) + M(1, 0)) * mult, y = val, z = (M(1, 2) + M(2, 1)) * mult;
      break;
    default:
      w = (M(0, 1) - M(1, 0)) * mult, x = (M(2, 0) + M(0, 2)) * mult, y = (M(1, 2) + M(2, 1)) * mult, z = val;
      break;
  }
  q[0] = x, q[1] = y, q[2] = z, q[3] = w;
#undef M
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
  opt->pipeline              = VKGS_PIPELINE_3DGS;
  opt->extent_projection     = VKGS_EXTENT_CONIC;  // src/parameters.h:190 (only the 3DGUT pipeline reads it)
  opt->kernel_degree         = 2;                  // KERNEL_DEGREE_QUADRATIC, src/parameters.h:215
  opt->quantize_normals      = 1;                  // prmRaster.quantizeNormals, src/parameters.h:195 (surface_info only)
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
  // 3DGUT pipeline: camera pose and inverse matrices (src/gaussian_splatting.cpp:1166-1169,1200,1254-1259)
  inverse4(out->view, out->view_inverse);
  inverse4(out->proj, out->proj_inverse);
  quatCast(out->view, out->view_quat);
  out->view_trans[0] = out->view[12], out->view_trans[1] = out->view[13], out->view_trans[2] = out->view[14];
  out->near_far[0] = cam->znear, out->near_far[1] = cam->zfar;
  out->alpha_clamp         = 0.99f;    // shaders/shaderio.h:271
  out->kernel_min_response = 0.0113f;  // src/parameters.h:216
  out->fov_rad             = fovy;     // cameraManip->getRadFov(), src/gaussian_splatting.cpp:1168
  out->depth_iso_threshold     = 0.7f;   // shaders/shaderio.h:311
  out->thin_particle_threshold = 1e-6f;  // shaders/shaderio.h:316
  return VKGS_OK;
}

// FISHEYE focal, src/gaussian_splatting.cpp:1239-1243
extern "C" VKGS_API void vkgs_frame_params_set_fisheye(vkgs_frame_params* fp)
{
  if(!fp)
    return;
  fp->focal[0] = 1.0f * fp->viewport[0] / fp->fov_rad;
  fp->focal[1] = -1.0f * fp->viewport[1] / fp->fov_rad;
}
