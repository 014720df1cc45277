// TEST INFRASTRUCTURE — golden-vector generator (not product code, never shipped).
//
// Compiles against the third-party headers the reference itself calls for the HOST
// side of the hot path, from where they lie under /root/reference (never copied):
//   * glm           (nvpro_core2/third_party/glm)   lookAt / perspectiveRH_ZO / quat / mat3
//   * spz           (3rdparty/spz/src/cc/splat-types.h) coordinateConverter
// and evaluates the exact call sequences of
//   * CameraManipulator::updateLookatMatrix / getPerspectiveMatrix
//       (nvpro_core2/nvutils/camera_manipulator.cpp:211, camera_manipulator.hpp:217-233)
//   * GaussianSplatting::updateAndUploadFrameInfoUBO focal (src/gaussian_splatting.cpp:1248-1250)
//   * SplatSetVk::initDataBuffers covariance + rgba packing (src/splat_set_vk.cpp:263-288,313-345)
//   * toUint8 / glm::packHalf1x16 quantisers (src/splat_set_vk.cpp:85-112)
//   * SplatSorterAsync::innerSort distance (src/splat_sorter_async.cpp:103-122)
//   * SplatSet::convertCoordinates flips (src/splat_set.h:78-114)
// on seeded inputs, writing inputs and outputs as IEEE-754 bit patterns to JSON.
// The committed fixture tests/golden/glm_golden.json pins oracle/vkgs_oracle.c and the
// product's host code bit-for-bit against the libraries the reference executes.
//
// Build + run: `make -C oracle golden`  (needs /root/reference; outputs to oracle/_ref/).
#include <cstdint>
#include <cstdio>
#include <cstring>
#include <cmath>
#include <algorithm>
#include <vector>

#include <glm/glm.hpp>
#include <glm/gtc/matrix_transform.hpp>
#include <glm/gtc/quaternion.hpp>
#include <glm/gtc/type_ptr.hpp>
#include <glm/gtc/packing.hpp>
#include <glm/gtx/transform.hpp>

#include "splat-types.h"

static uint32_t f2u(float f)
{
  uint32_t u;
  std::memcpy(&u, &f, 4);
  return u;
}

static uint64_t g_state = 0x3D65D00Dull;
static uint64_t splitmix64()
{
  uint64_t z = (g_state += 0x9E3779B97F4A7C15ull);
  z          = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
  z          = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
  return z ^ (z >> 31);
}
static float urand(float lo, float hi)
{
  const float u = float(splitmix64() >> 40) * (1.0f / 16777216.0f);
  return lo + (hi - lo) * u;
}

static void putArr(FILE* f, const char* name, const std::vector<uint32_t>& v, bool last = false)
{
  std::fprintf(f, "  \"%s\": [", name);
  for(size_t i = 0; i < v.size(); i++)
    std::fprintf(f, "%s%u", i ? "," : "", v[i]);
  std::fprintf(f, "]%s\n", last ? "" : ",");
}

// the quantiser exactly as written in src/splat_set_vk.cpp:85-89
static uint8_t toUint8(float v, float rangeMin, float rangeMax)
{
  float normalized = (v - rangeMin) / (rangeMax - rangeMin);
  return static_cast<uint8_t>(std::clamp(std::round(normalized * 255.0f), 0.0f, 255.0f));
}

int main(int argc, char** argv)
{
  const char* outPath = argc > 1 ? argv[1] : "glm_golden.json";
  FILE*       f       = std::fopen(outPath, "w");
  if(!f)
    return 1;
  std::fprintf(f, "{\n");

  // ---------------- cameras -------------------------------------------------
  {
    struct Cam
    {
      glm::vec3 eye, ctr, up;
      float     fov, znear, zfar;
      uint32_t  w, h;
    };
    std::vector<Cam> cams = {
        {{1.7F, 1.5F, 1.7F}, {0, 0, 0}, {0, 1, 0}, 60.0f, 0.1f, 2000.0f, 1920, 1080},  // src/camera_set.h:48-53
        {{1.7F, 1.5F, 1.7F}, {0, 0, 0}, {0, 1, 0}, 60.0f, 0.1f, 2000.0f, 3840, 2160},
        {{1.7F, 1.5F, 1.7F}, {0, 0, 0}, {0, 1, 0}, 60.0f, 0.1f, 2000.0f, 512, 512},
        {{-2.0F, 0.3F, 0.9F}, {0.1F, -0.2F, 0.05F}, {0, 1, 0}, 45.0f, 0.01f, 100.0f, 1465, 766},
        {{0.0F, 2.83F, 0.001F}, {0, 0, 0}, {0, 0, -1}, 75.0f, 0.5f, 50.0f, 640, 360},
    };
    for(int i = 0; i < 8; i++)
    {
      Cam c;
      c.eye   = {urand(-3, 3), urand(-3, 3), urand(-3, 3)};
      c.ctr   = {urand(-0.5f, 0.5f), urand(-0.5f, 0.5f), urand(-0.5f, 0.5f)};
      c.up    = {0, 1, 0};
      c.fov   = urand(20, 100);
      c.znear = urand(0.01f, 0.5f);
      c.zfar  = urand(10, 3000);
      c.w     = 64 + uint32_t(splitmix64() % 4000);
      c.h     = 64 + uint32_t(splitmix64() % 2200);
      cams.push_back(c);
    }
    std::vector<uint32_t> in, view, proj, focal, viewInv, projInv, viewQuat;
    for(const Cam& c : cams)
    {
      for(int k = 0; k < 3; k++)
        in.push_back(f2u(c.eye[k]));
      for(int k = 0; k < 3; k++)
        in.push_back(f2u(c.ctr[k]));
      for(int k = 0; k < 3; k++)
        in.push_back(f2u(c.up[k]));
      in.push_back(f2u(c.fov));
      in.push_back(f2u(c.znear));
      in.push_back(f2u(c.zfar));
      in.push_back(c.w);
      in.push_back(c.h);
      // camera_manipulator.cpp:211
      glm::mat4 V = glm::lookAt(c.eye, c.ctr, c.up);
      // camera_manipulator.hpp:229-231,267,275
      const float aspect = static_cast<float>(c.w) / static_cast<float>(c.h);
      glm::mat4   P      = glm::perspectiveRH_ZO(glm::radians(c.fov), aspect, c.znear, c.zfar);
      P[1][1] *= -1;
      // gaussian_splatting.cpp:1248-1250 (devicePixelRatio = 1)
      const float     devicePixelRatio = 1.0;
      const glm::vec2 renderSize       = glm::vec2(c.w, c.h);
      const float     fx               = P[0][0] * 0.5f * devicePixelRatio * renderSize.x;
      const float     fy               = P[1][1] * 0.5f * devicePixelRatio * renderSize.y;
      for(int k = 0; k < 16; k++)
        view.push_back(f2u(glm::value_ptr(V)[k]));
      for(int k = 0; k < 16; k++)
        proj.push_back(f2u(glm::value_ptr(P)[k]));
      focal.push_back(f2u(fx));
      focal.push_back(f2u(fy));
      // 3DGUT camera pose: gaussian_splatting.cpp:1166 (viewInverse), :1200 (projInverse), :1254-1259 (viewQuat)
      const glm::mat4 Vi = glm::inverse(V), Pi = glm::inverse(P);
      const glm::quat q  = glm::quat_cast(V);
      for(int k = 0; k < 16; k++)
        viewInv.push_back(f2u(glm::value_ptr(Vi)[k]));
      for(int k = 0; k < 16; k++)
        projInv.push_back(f2u(glm::value_ptr(Pi)[k]));
      viewQuat.push_back(f2u(q.x)), viewQuat.push_back(f2u(q.y)), viewQuat.push_back(f2u(q.z)), viewQuat.push_back(f2u(q.w));
    }
    std::fprintf(f, "  \"camera_count\": %zu,\n", cams.size());
    putArr(f, "camera_in", in);
    putArr(f, "camera_view", view);
    putArr(f, "camera_proj", proj);
    putArr(f, "camera_focal", focal);
    putArr(f, "camera_view_inverse", viewInv);
    putArr(f, "camera_proj_inverse", projInv);
    putArr(f, "camera_view_quat", viewQuat);
  }

  // ---------------- pack: covariance + rgba ----------------------------------
  {
    const int             n = 256;
    std::vector<uint32_t> in, cov, rgba, rgba_u8, rgba_f16;
    for(int i = 0; i < n; i++)
    {
      float scale[3], rotation[4], f_dc[3], opacity;
      for(float& s : scale)
        s = urand(-7.0f, 0.5f);
      for(float& r : rotation)
        r = urand(-1.5f, 1.5f);
      for(float& c : f_dc)
        c = urand(-2.5f, 2.5f);
      opacity = urand(-8.0f, 8.0f);
      if(i == 0)
      {
        rotation[0] = 1;
        rotation[1] = rotation[2] = rotation[3] = 0;
      }
      if(i == 1)
      {  // degenerate quaternion → glm::normalize returns identity
        rotation[0] = rotation[1] = rotation[2] = rotation[3] = 0;
      }
      for(float s : scale)
        in.push_back(f2u(s));
      for(float r : rotation)
        in.push_back(f2u(r));
      for(float c : f_dc)
        in.push_back(f2u(c));
      in.push_back(f2u(opacity));

      // src/splat_set_vk.cpp:263-288
      glm::vec3 scl{std::exp(scale[0]), std::exp(scale[1]), std::exp(scale[2])};
      glm::quat rot{rotation[0], rotation[1], rotation[2], rotation[3]};
      rot                                   = glm::normalize(rot);
      const glm::mat3 scaleMatrix           = glm::mat3(glm::scale(scl));
      const glm::mat3 rotationMatrix        = glm::mat3_cast(rot);
      const glm::mat3 covarianceMatrix      = rotationMatrix * scaleMatrix;
      glm::mat3       transformedCovariance = covarianceMatrix * glm::transpose(covarianceMatrix);
      const int       pick[6]               = {0, 3, 6, 4, 7, 8};
      for(int k : pick)
        cov.push_back(f2u(glm::value_ptr(transformedCovariance)[k]));

      // src/splat_set_vk.cpp:313-345
      const float SH_C0 = 0.28209479177387814f;
      const float r     = glm::clamp(0.5f + SH_C0 * f_dc[0], 0.0f, 1.0f);
      const float g     = glm::clamp(0.5f + SH_C0 * f_dc[1], 0.0f, 1.0f);
      const float b     = glm::clamp(0.5f + SH_C0 * f_dc[2], 0.0f, 1.0f);
      const float a     = glm::clamp(1.0f / (1.0f + std::exp(-opacity)), 0.0f, 1.0f);
      for(float v : {r, g, b, a})
      {
        rgba.push_back(f2u(v));
        rgba_u8.push_back(toUint8(v, 0.f, 1.f));
        rgba_f16.push_back(glm::packHalf1x16(v));
      }
    }
    std::fprintf(f, "  \"pack_count\": %d,\n", n);
    putArr(f, "pack_in", in);
    putArr(f, "pack_cov6", cov);
    putArr(f, "pack_rgba", rgba);
    putArr(f, "pack_rgba_u8", rgba_u8);
    putArr(f, "pack_rgba_f16", rgba_f16);
  }

  // ---------------- SH quantisers (storeSh, src/splat_set_vk.cpp:104-112) ------
  {
    const int             n = 512;
    std::vector<uint32_t> in, u8, f16;
    for(int i = 0; i < n; i++)
    {
      float v = urand(-1.3f, 1.3f);
      if(i == 0)
        v = 0.0f;
      if(i == 1)
        v = 1.0f;
      if(i == 2)
        v = -1.0f;
      if(i == 3)
        v = 0.5f / 255.0f * 2.0f - 1.0f;  // rounding boundary
      if(i == 4)
        v = 6.1e-5f;  // fp16 subnormal boundary
      if(i == 5)
        v = 1e-8f;
      in.push_back(f2u(v));
      u8.push_back(toUint8(v, -1., 1.));
      f16.push_back(glm::packHalf1x16(v));
    }
    std::fprintf(f, "  \"quant_count\": %d,\n", n);
    putArr(f, "quant_in", in);
    putArr(f, "quant_u8", u8);
    putArr(f, "quant_f16", f16);
  }

  // ---------------- CPU sorter distance (src/splat_sorter_async.cpp:103-122) ---
  {
    const int             n = 256;
    std::vector<uint32_t> in, out;
    float                 dir[3] = {-1.7f, -1.5f, -1.7f};
    float                 cop[3] = {1.7f, 1.5f, 1.7f};
    glm::mat4             xform  = glm::translate(glm::mat4(1.0f), glm::vec3(0.25f, -0.5f, 0.125f))
                      * glm::rotate(glm::mat4(1.0f), 0.7f, glm::normalize(glm::vec3(0.3f, 1.0f, -0.2f)))
                      * glm::scale(glm::mat4(1.0f), glm::vec3(1.5f, 0.75f, 1.25f));
    for(float d : dir)
      in.push_back(f2u(d));
    for(float c : cop)
      in.push_back(f2u(c));
    for(int k = 0; k < 16; k++)
      in.push_back(f2u(glm::value_ptr(xform)[k]));
    const glm::vec4 plane(dir[0], dir[1], dir[2], -dir[0] * cop[0] - dir[1] * cop[1] - dir[2] * cop[2]);
    const float     divider = 1.0f / std::sqrt(plane[0] * plane[0] + plane[1] * plane[1] + plane[2] * plane[2]);
    for(int i = 0; i < n; i++)
    {
      float pos[3] = {urand(-2, 2), urand(-2, 2), urand(-2, 2)};
      for(float p : pos)
        in.push_back(f2u(p));
      const glm::vec4 p    = xform * glm::vec4(pos[0], pos[1], pos[2], 1.0f);
      const float     dist = std::abs(plane[0] * p[0] + plane[1] * p[1] + plane[2] * p[2] + plane[3]) * divider;
      out.push_back(f2u(dist));
    }
    std::fprintf(f, "  \"sorter_count\": %d,\n", n);
    putArr(f, "sorter_in", in);
    putArr(f, "sorter_dist", out);
  }

  // ---------------- model-view product + transformed centre -------------------
  // glm restatement of what the shaders compute with mul(): (V*M)*p and V*(M*p)
  // (shaders/threedgs_raster.mesh.slang:173-176, shaders/dist.comp.slang:58). The GPU
  // evaluates these in SPIR-V, so this block only pins the matrix conventions (A.0),
  // compared with a tolerance, not bit-exactly.
  {
    const int             n = 64;
    std::vector<uint32_t> in, out;
    glm::mat4             V = glm::lookAt(glm::vec3(1.7F, 1.5F, 1.7F), glm::vec3(0), glm::vec3(0, 1, 0));
    glm::mat4             M = glm::translate(glm::mat4(1.0f), glm::vec3(0.1f, 0.2f, -0.3f))
                  * glm::rotate(glm::mat4(1.0f), -0.4f, glm::normalize(glm::vec3(1.0f, 0.2f, 0.1f)));
    glm::mat4 P = glm::perspectiveRH_ZO(glm::radians(60.0f), 1920.0f / 1080.0f, 0.1f, 2000.0f);
    P[1][1] *= -1;
    for(int k = 0; k < 16; k++)
      in.push_back(f2u(glm::value_ptr(V)[k]));
    for(int k = 0; k < 16; k++)
      in.push_back(f2u(glm::value_ptr(M)[k]));
    for(int k = 0; k < 16; k++)
      in.push_back(f2u(glm::value_ptr(P)[k]));
    for(int i = 0; i < n; i++)
    {
      glm::vec4 p(urand(-1, 1), urand(-1, 1), urand(-1, 1), 1.0f);
      for(int k = 0; k < 3; k++)
        in.push_back(f2u(p[k]));
      glm::vec4 view = V * (M * p);
      glm::vec4 clip = P * view;
      glm::vec4 ndc  = clip / clip.w;
      for(int k = 0; k < 4; k++)
        out.push_back(f2u(view[k]));
      for(int k = 0; k < 4; k++)
        out.push_back(f2u(ndc[k]));
    }
    std::fprintf(f, "  \"xform_count\": %d,\n", n);
    putArr(f, "xform_in", in);
    putArr(f, "xform_out", out);
  }

  // ---------------- coordinate flips (src/splat_set.h:78-114 via spz) -----------
  {
    spz::CoordinateConverter c = spz::coordinateConverter(spz::CoordinateSystem::RDF, spz::CoordinateSystem::RUB);
    std::vector<uint32_t>    v;
    for(float x : c.flipP)
      v.push_back(f2u(x));
    for(float x : c.flipQ)
      v.push_back(f2u(x));
    for(float x : c.flipSh)
      v.push_back(f2u(x));
    putArr(f, "flip_rdf_to_rub", v, true);
  }

  std::fprintf(f, "}\n");
  std::fclose(f);
  return 0;
}
