// k_preprocess.cu — fused per-splat front end (one pass over the splat arrays, in storage order).
//
// Replaces, for every splat i (reference file:line):
//   dist.comp.slang:40-171       view/clip transform, NDC depth -> sortable key, frustum / size cull,
//                                append (key, id)                       [K1]
//   threedgs_raster.mesh.slang:161-289   alpha cull, SH colour, covariance projection, extent basis [K5]
//   threedgs.h.slang:26-121, threedgs_particle_storage.h.slang:103-159
// B200 design: the reference runs K5 after the sort and gathers 232 B/splat by sorted id. Here the
// whole per-splat stage runs BEFORE the sort, in storage order, so every attribute array is
// streamed exactly once with bulk async copies (TMA engine, cp.async.bulk -> shared memory, one
// mbarrier per tile); the sort then only moves 8-byte (key,id) pairs and the raster stages gather
// a compact 48-byte record per splat. The append is deterministic: a decoupled look-back prefix
// over tiles replaces the reference's global atomic, so (key,id) pairs come out in ascending splat
// id — one of the orders the reference's atomic can legally produce. The four 8-bit digit
// histograms the radix sort needs are accumulated here as well (no separate histogram pass).
//
// Arithmetic: this file is compiled with -fmad=false and evaluates every expression in the
// operation order documented in oracle/vkgs_oracle.c, with IEEE division and square root, so keys
// and records are bit-identical to the CPU oracle.
#include <cstdlib>
#include <cuda_fp16.h>

#include "device_common.cuh"
#include "kernels.hpp"

namespace vkgs {

namespace {

constexpr int NWARPS = PRE_TILE / 32;

struct PreSmem
{
  // async-copy destinations first (16-byte aligned offsets)
  float    sh[PRE_TILE * 45];   // 46080 B (fp32 worst case)
  float    cov[PRE_TILE * 6];   //  6144 B
  float    rgba[PRE_TILE * 4];  //  4096 B
  float    center[PRE_TILE * 3];//  3072 B
  float    scale[PRE_TILE * 3]; //  3072 B (size culling only)
  uint64_t mbarA;  // centers (+ scales): all the dist/cull stage needs
  uint64_t mbarB;  // cov6, rgba, SH: the per-splat projection stage
  uint32_t hist[4][256];  // digit histograms of this tile's keys (all four sort passes)
  // uint8 storage formats: the decoded value of every byte, v / 255 * 2 - 1 (SH) and v / 255 (rgba), evaluated once per CTA
  // with the IEEE division the oracle uses (threedgs_particle_buffers.h.slang:119-131) — a table look-up per coefficient
  // instead of a ten-instruction division each
  float    u8Sh[256];
  float    u8Unit[256];
  uint32_t warpScan[NWARPS + 1];
  uint32_t tile;
  uint32_t basePrefix[2];  // exclusive prefix of the tile processed in iteration parity 0 / 1
};

__device__ __forceinline__ float clampFinite(float v)
{
  return fminf(fmaxf(v, -3.4028235e38f), 3.4028235e38f);  // NaN -> -FLT_MAX, +-inf -> +-FLT_MAX, finite values unchanged
}

__device__ __forceinline__ uint32_t encodeMinMaxFp32(float v)
{
  uint32_t bits = __float_as_uint(v);
  bits ^= static_cast<uint32_t>(static_cast<int32_t>(bits) >> 31) | 0x80000000u;
  return bits;
}

// mul(v, S) with Slang row-major S[i][j] = m[4*i+j]
__device__ __forceinline__ void mulVecMat(const float v[4], const float* m, float out[4])
{
#pragma unroll
  for(int j = 0; j < 4; j++)
    out[j] = ((v[0] * m[0 + j] + v[1] * m[4 + j]) + v[2] * m[8 + j]) + v[3] * m[12 + j];
}

template <int FMT>
__device__ __forceinline__ float loadSh(const void* base, int idx, const float* u8Table)
{
  if constexpr(FMT == VKGS_FORMAT_FLOAT32)
    return static_cast<const float*>(base)[idx];
  else if constexpr(FMT == VKGS_FORMAT_FLOAT16)
    return __half2float(static_cast<const __half*>(base)[idx]);
  else  // threedgs_particle_buffers.h.slang:119-131: v / 255 * range - halfRange, through the per-CTA table of all 256 values
    return u8Table[static_cast<const uint8_t*>(base)[idx]];
}

__device__ __forceinline__ float4 loadRgba(const void* base, int t, uint32_t fmt, const float* u8Unit)
{
  if(fmt == VKGS_FORMAT_FLOAT32)
    return static_cast<const float4*>(base)[t];
  if(fmt == VKGS_FORMAT_FLOAT16)
  {
    const __half* h = static_cast<const __half*>(base) + 4 * t;
    return make_float4(__half2float(h[0]), __half2float(h[1]), __half2float(h[2]), __half2float(h[3]));
  }
  const uchar4 u = static_cast<const uchar4*>(base)[t];
  return make_float4(u8Unit[u.x], u8Unit[u.y], u8Unit[u.z], u8Unit[u.w]);
}

// fetchViewDependentRadiance, threedgs_particle_storage.h.slang:103-159
template <int FMT>
__device__ __forceinline__ void shRadiance(const void* row, int rowBase, uint32_t degree, float x, float y, float z, float rgb[3], const float* u8Table)
{
  const float C1    = 0.4886025119029199f;
  const float C2[5] = {1.0925484f, -1.0925484f, 0.3153916f, -1.0925484f, 0.5462742f};
  const float C3[7] = {-0.5900435899266435f, 2.890611442640554f, -0.4570457994644658f, 0.3731763325901154f,
                       -0.4570457994644658f, 1.445305721320277f, -0.5900435899266435f};
#pragma unroll
  for(int c = 0; c < 3; c++)
  {
#define S(k) loadSh<FMT>(row, rowBase + 3 * (k) + c, u8Table)
    float acc = 0.0f;
    acc += C1 * (-S(0) * y + S(1) * z - S(2) * x);
    if(degree >= 2)
    {
      const float xx = x * x, yy = y * y, zz = z * z, xy = x * y, yz = y * z, xz = x * z;
      acc += (C2[0] * xy) * S(3) + (C2[1] * yz) * S(4) + (C2[2] * (2.0f * zz - xx - yy)) * S(5) + (C2[3] * xz) * S(6)
             + (C2[4] * (xx - yy)) * S(7);
      if(degree >= 3)
      {
        acc += C3[0] * S(8) * (3.0f * x * x - y * y) * y + C3[1] * S(9) * x * y * z
               + C3[2] * S(10) * (4.0f * z * z - x * x - y * y) * y
               + C3[3] * S(11) * z * (2.0f * z * z - 3.0f * x * x - 3.0f * y * y)
               + C3[4] * S(12) * x * (4.0f * z * z - x * x - y * y) + C3[5] * S(13) * (x * x - y * y) * z
               + C3[6] * S(14) * x * (x * x - 3.0f * y * y);
      }
    }
#undef S
    rgb[c] = acc;
  }
}

// ---- VK3DGUT per-splat stage (threedgut_raster.mesh.slang:97-255; pinhole camera, EXTENT_CONIC) ----------
// Same operation order as orc_gut_project_splat / gut_project_point in oracle/vkgs_oracle.c.

// projectPointFisheye on initPerfectFisheyeCamera(viewport, focal) (threedgut_camera_projections.h.slang:149-171,
// threedgut_camera_models.h.slang:85-139), in the operation order of gut_project_fisheye (oracle)
// (by-value interface: a pointer argument of a non-inlined function would push the caller's sigma-point array to local memory)
__device__ __noinline__ float3 gutProjectFisheyeV(float posX, float posY, float posZ, const vkgs_frame_params& fp)
{
  const float pos[3] = {posX, posY, posZ};
  float       projected[2];
  const float W = fp.viewport[0], H = fp.viewport[1];
  const float ppx = W / 2.0f, ppy = H / 2.0f;
  const float mdx = (ppx > 0.5f * W) ? ppx : W - ppx, mdy = (ppy > 0.5f * H) ? ppy : H - ppy;
  const float maxRadius = sqrtf(mdx * mdx + mdy * mdy);
  const float fovAngleX = 2.0f * maxRadius / fp.focal[0], fovAngleY = 2.0f * maxRadius / fp.focal[1];
  const float maxAngle  = fmaxf(fovAngleX, fovAngleY) / 2.0f;
  const float absX = fabsf(pos[0]), absY = fabsf(pos[1]);
  const float minVal = fminf(absX, absY), maxVal = fmaxf(absX, absY);
  float       norm = 0.0f;
  if(!(maxVal <= 0.0f))
  {
    const float ratio = minVal / maxVal;
    norm              = maxVal * sqrtf(1.0f + ratio * ratio);
  }
  const float rho       = fmaxf(norm, 1e-7f);
  const float thetaFull = atan2fYposExact(rho, pos[2]);
  const float theta     = fminf(thetaFull, maxAngle);
  const float theta2    = theta * theta;
  float       poly      = 0.0f;
#pragma unroll
  for(int i = 2; i >= 0; --i)
    poly = theta2 * poly + 0.0f;
  const float delta = (theta * (poly * theta2 + 1.0f)) / rho;
  projected[0]      = (fp.focal[0] * pos[0]) * delta + ppx;
  projected[1]      = (fp.focal[1] * pos[1]) * delta + ppy;
  const float tolx = W * 0.1f, toly = H * 0.1f;
  const bool  valid = (theta < maxAngle) && (projected[0] > -tolx) && (projected[1] > -toly) && (projected[0] < W + tolx) && (projected[1] < H + toly);
  return make_float3(projected[0], projected[1], valid ? 1.0f : 0.0f);
}

__device__ __forceinline__ int gutProjectFisheye(const float pos[3], const vkgs_frame_params& fp, float projected[2])
{
  const float3 r = gutProjectFisheyeV(pos[0], pos[1], pos[2], fp);
  projected[0] = r.x, projected[1] = r.y;
  return r.z != 0.0f;
}

// projectPointWithShutter (global shutter) + projectPointPinhole with zero distortion coefficients
// (threedgut_camera_projections.h.slang:87-139,186-203), or projectPointFisheye
__device__ __forceinline__ int gutProjectPoint(const float world[3], const vkgs_frame_params& fp, uint32_t cameraModel, float projected[2])
{
  const float t[3]  = {fp.view_trans[0] * 1.0f, fp.view_trans[1] * 1.0f, fp.view_trans[2] * -1.0f};
  const float q[4]  = {fp.view_quat[0] * -1.0f, fp.view_quat[1] * -1.0f, fp.view_quat[2] * 1.0f, fp.view_quat[3] * 1.0f};
  const float pf[3] = {world[0] * 1.0f, world[1] * 1.0f, world[2] * -1.0f};
  const float tt[3] = {2.0f * (q[1] * pf[2] - q[2] * pf[1]), 2.0f * (q[2] * pf[0] - q[0] * pf[2]), 2.0f * (q[0] * pf[1] - q[1] * pf[0])};
  const float r[3]  = {(pf[0] + q[3] * tt[0]) + (q[1] * tt[2] - q[2] * tt[1]), (pf[1] + q[3] * tt[1]) + (q[2] * tt[0] - q[0] * tt[2]),
                       (pf[2] + q[3] * tt[2]) + (q[0] * tt[1] - q[1] * tt[0])};
  const float pos[3] = {r[0] + t[0], r[1] + t[1], r[2] + t[2]};
  if(cameraModel == VKGS_CAMERA_FISHEYE)
    return gutProjectFisheye(pos, fp, projected);
  if(pos[2] <= 0.0f)
  {
    projected[0] = projected[1] = 0.0f;
    return 0;
  }
  const float u = pos[0] / pos[2], v = pos[1] / pos[2];
  const float u2 = u * u, v2 = v * v, r2 = u2 + v2, a1 = (2.0f * u) * v, a2 = r2 + 2.0f * u2, a3 = r2 + 2.0f * v2;
  const float icDn = 1.0f + r2 * (0.0f + r2 * (0.0f + r2 * 0.0f)), icDd = 1.0f + r2 * (0.0f + r2 * (0.0f + r2 * 0.0f));
  const float icD  = icDn / icDd;
  const float dx = (0.0f * a1 + 0.0f * a2) + r2 * (0.0f + r2 * 0.0f), dy = (0.0f * a3 + 0.0f * a1) + r2 * (0.0f + r2 * 0.0f);
  const float und = icD * u + dx, vnd = icD * v + dy;
  const int   validRadial = (icD > 0.8f) && (icD < 1.2f);
  const float ppx = fp.viewport[0] / 2.0f, ppy = fp.viewport[1] / 2.0f;
  projected[0] = und * fp.focal[0] + ppx;
  projected[1] = vnd * fp.focal[1] + ppy;
  const float tolx = fp.viewport[0] * 0.1f, toly = fp.viewport[1] * 0.1f;
  return validRadial && (projected[0] > -tolx) && (projected[1] > -toly) && (projected[0] < fp.viewport[0] + tolx)
         && (projected[1] < fp.viewport[1] + toly);
}

// Unscented-transform projection + conic extent of one splat. Writes the 24-word 3DGUT record
//   cx cy ex ey | r g b a | ro.xyz |ro| | 1/scale.xyz R00 | R01 R02 R10 R11 | R12 R20 R21 R22
//   (EXTENT_EIGEN: words 2,3 = w1 of the quad, word 11 = |w2| / |w1|)
// (R = inverse rotation, ro = canonical ray origin: the ray origin is the camera for every pixel, so
// particleCannonicalRay's origin half is evaluated once per splat) and returns the pixel bounding box.
__device__ __forceinline__ bool gutProjectSplat(const PreprocessArgs& a, const float c[4], const float4 rq, const float* scaleLog,
                                               float4 col, float4 rec[6], uint32_t& bb0, uint32_t& bb1)
{
  const vkgs_frame_params& fp = a.fp;
  const float scale[3] = {expfExact(scaleLog[0]), expfExact(scaleLog[1]), expfExact(scaleLog[2])};
  const float rinv = 1.0f / sqrtf(((rq.y * rq.y + rq.z * rq.z) + rq.w * rq.w) + rq.x * rq.x);
  const float x = rq.y * rinv, y = rq.z * rinv, z = rq.w * rinv, w = rq.x * rinv;
  const float xx = x * x, yy = y * y, zz = z * z, xy = x * y, xz = x * z, yz = y * z, wx = w * x, wy = w * y, wz = w * z;
  const float rot[3][3] = {{1.0f - 2.0f * (yy + zz), 2.0f * (xy + wz), 2.0f * (xz - wy)},
                           {2.0f * (xy - wz), 1.0f - 2.0f * (xx + zz), 2.0f * (yz + wx)},
                           {2.0f * (xz + wy), 2.0f * (yz - wx), 1.0f - 2.0f * (xx + yy)}};
  if(col.w < fp.alpha_cull_threshold)
    return false;
  const float GUT_DELTA = 1.73205080757f, GUT_LAMBDA = 0.0f, GUT_D = 3.0f;
  float       sp[7][2];
  int         nvalid = 0;
  float       world[4], pt[4];
  mulVecMat(c, fp.model, world);
  nvalid += gutProjectPoint(world, fp, a.opt.camera_model, sp[0]);
  const float w0c   = GUT_LAMBDA / (GUT_D + GUT_LAMBDA);
  float       pc[2] = {sp[0][0] * w0c, sp[0][1] * w0c};
  const float wi    = 1.0f / (2.0f * (GUT_D + GUT_LAMBDA));
#pragma unroll
  for(int i = 0; i < 3; i++)
  {
    const float dl[3] = {(GUT_DELTA * scale[i]) * rot[i][0], (GUT_DELTA * scale[i]) * rot[i][1], (GUT_DELTA * scale[i]) * rot[i][2]};
    pt[0] = c[0] + dl[0], pt[1] = c[1] + dl[1], pt[2] = c[2] + dl[2], pt[3] = 1.0f;
    mulVecMat(pt, fp.model, world);
    nvalid += gutProjectPoint(world, fp, a.opt.camera_model, sp[i + 1]);
    pc[0] += wi * sp[i + 1][0], pc[1] += wi * sp[i + 1][1];
    pt[0] = c[0] - dl[0], pt[1] = c[1] - dl[1], pt[2] = c[2] - dl[2];
    mulVecMat(pt, fp.model, world);
    nvalid += gutProjectPoint(world, fp, a.opt.camera_model, sp[i + 4]);
    pc[0] += wi * sp[i + 4][0], pc[1] += wi * sp[i + 4][1];
  }
  if(nvalid == 0)
    return false;
  float cov[3];
  {
    const float cx = sp[0][0] - pc[0], cy = sp[0][1] - pc[1];
    const float w0 = GUT_LAMBDA / (GUT_D + GUT_LAMBDA) + ((1.0f - 1.0f * 1.0f) + 2.0f);
    cov[0] = w0 * (cx * cx), cov[1] = w0 * (cx * cy), cov[2] = w0 * (cy * cy);
  }
#pragma unroll
  for(int i = 0; i < 6; i++)
  {
    const float cx = sp[i + 1][0] - pc[0], cy = sp[i + 1][1] - pc[1];
    cov[0] += wi * (cx * cx), cov[1] += wi * (cx * cy), cov[2] += wi * (cy * cy);
  }
  float ex, ey, q0z, q0w, q2w = 0.0f;  // record words 2,3 (extent, or w1 of the EIGEN quad) and 11 (|ro|, or |w2|/|w1|)
  if(a.opt.extent_projection == VKGS_EXTENT_CONIC)
  {
    // threedgutProjectedExtentConicOpacity (threedgut.h.slang:118-163)
    const float dc[3] = {cov[0] + 0.3f, cov[1], cov[2] + 0.3f};
    const float ddet  = dc[0] * dc[2] - dc[1] * dc[1];
    if(ddet == 0.0f)
      return false;
    float conicW = col.w;
    if(a.opt.ms_antialiasing)
    {
      const float det = cov[0] * cov[2] - cov[1] * cov[1];
      conicW          = col.w * sqrtf(fmaxf(0.000025f, det / ddet));
    }
    if(conicW < 0.01f)
      return false;
    const float maxPower = logf(conicW / 0.01f);
    const float ef       = fminf(3.33f, sqrtf(2.0f * maxPower));
    const float mid      = 0.5f * (dc[0] + dc[2]);
    const float lambda   = mid + sqrtf(fmaxf(0.01f, mid * mid - ddet));
    const float radius   = ef * sqrtf(lambda);
    ex = fminf(ef * sqrtf(dc[0]), radius), ey = fminf(ef * sqrtf(dc[2]), radius);
    if(!(radius > 0.0f))
      return false;
    if(a.opt.ms_antialiasing)
      col.w = conicW;
    q0z = ex, q0w = ey;
  }
  else
  {
    // threedgsProjectedExtentBasis(cov, 3.33, splatScale, opacity, b1, b2) (threedgs.h.slang:60-121): the quad is
    // centre +- b1 +- b2; the record keeps w1 = b1 / |b1|^2 and |w2| / |w1| (w2 is w1 turned by -90 degrees)
    float cv0 = cov[0], cv1 = cov[1], cv2 = cov[2], detOrig = 0.0f;
    if(a.opt.ms_antialiasing)
      detOrig = cv0 * cv2 - cv1 * cv1;
    cv0 += 0.3f, cv2 += 0.3f;
    if(a.opt.ms_antialiasing)
      col.w *= sqrtf(fmaxf(detOrig / (cv0 * cv2 - cv1 * cv1), 0.0f));
    const float D = cv0 * cv2 - cv1 * cv1, trace = cv0 + cv2, t2 = 0.5f * trace;
    const float term2 = sqrtf(fmaxf(0.1f, t2 * t2 - D));
    float       ev1 = t2 + term2, ev2 = t2 - term2;
    if(ev2 <= 0.0f)
      return false;
    if(a.opt.point_cloud_mode)
      ev1 = ev2 = 0.2f;
    float       e1x = (fabsf(cv1) < 0.001f) ? 1.0f : cv1, e1y = ev1 - cv0;
    const float einv = 1.0f / sqrtf(e1x * e1x + e1y * e1y);
    e1x *= einv, e1y *= einv;
    const float m1 = fminf(3.33f * sqrtf(ev1), 2048.0f), m2 = fminf(3.33f * sqrtf(ev2), 2048.0f);
    const float b1x = e1x * fp.splat_scale * m1, b1y = e1y * fp.splat_scale * m1;
    const float b2x = e1y * fp.splat_scale * m2, b2y = -e1x * fp.splat_scale * m2;
    const float n1 = b1x * b1x + b1y * b1y, n2 = b2x * b2x + b2y * b2y;
    if(!(n1 > 0.0f) || !(n2 > 0.0f))
      return false;
    ex = fabsf(b1x) + fabsf(b2x), ey = fabsf(b1y) + fabsf(b2y);
    q0z = b1x / n1, q0w = b1y / n1;
    q2w = sqrtf(n1 / n2);
  }
  float wc[4], vc[4], cc[4];
  mulVecMat(c, fp.model, wc);
  mulVecMat(wc, fp.view, vc);
  mulVecMat(vc, fp.proj, cc);
  const float nz = cc[2] / cc[3];
  if(!(nz >= 0.0f && nz <= 1.0f) || !isfinite(pc[0]) || !isfinite(pc[1]) || !(ex < 1e7f) || !(ey < 1e7f))
    return false;
  // canonical ray origin: giscl * mul(camModel - position, invRotation) (threedgrt.h.slang:65-69)
  const float giscl[3] = {1.0f / scale[0], 1.0f / scale[1], 1.0f / scale[2]};
  const float gposc[3] = {a.gutOrigin[0] - c[0], a.gutOrigin[1] - c[1], a.gutOrigin[2] - c[2]};
  float       ro[3];
#pragma unroll
  for(int j = 0; j < 3; j++)  // invRotation[i][j] = rot[j][i]
    ro[j] = giscl[j] * ((gposc[0] * rot[j][0] + gposc[1] * rot[j][1]) + gposc[2] * rot[j][2]);
  rec[0] = make_float4(pc[0], pc[1], q0z, q0w);
  rec[1] = col;
  // (w: length of the canonical origin — the blend's fast path scales its guard band with it, see k_blend.cu)
  if(a.opt.extent_projection == VKGS_EXTENT_CONIC)
    q2w = sqrtf((ro[0] * ro[0] + ro[1] * ro[1]) + ro[2] * ro[2]);
  rec[2] = make_float4(ro[0], ro[1], ro[2], q2w);
  rec[3] = make_float4(giscl[0], giscl[1], giscl[2], rot[0][0]);        // invRot row 0 = (rot[0][0], rot[1][0], rot[2][0])
  rec[4] = make_float4(rot[1][0], rot[2][0], rot[0][1], rot[1][1]);      // invRot[0][1..2], invRot[1][0..1]
  rec[5] = make_float4(rot[2][1], rot[0][2], rot[1][2], rot[2][2]);      // invRot[1][2], invRot[2][0..2]
  // pixel bounding box of the quad, one pixel of slack (the blend applies the exact |d| <= extent test)
  const float W = fp.viewport[0], H = fp.viewport[1];
  const float fx0 = floorf(pc[0] - ex - 1.0f), fx1 = ceilf(pc[0] + ex + 1.0f);
  const float fy0 = floorf(pc[1] - ey - 1.0f), fy1 = ceilf(pc[1] + ey + 1.0f);
  if(fx1 < 0.0f || fy1 < 0.0f || fx0 > W - 1.0f || fy0 > H - 1.0f)
    return false;
  const uint32_t x0 = static_cast<uint32_t>(fmaxf(fx0, 0.0f)), x1 = static_cast<uint32_t>(fminf(fx1, W - 1.0f));
  const uint32_t y0 = static_cast<uint32_t>(fmaxf(fy0, 0.0f)), y1 = static_cast<uint32_t>(fminf(fy1, H - 1.0f));
  bb0 = x0 | (y0 << 16);
  bb1 = x1 | (y1 << 16);
  return true;
}

template <int SHFMT, bool GUT>
__global__ void __launch_bounds__(PRE_TILE) k_preprocess(const __grid_constant__ PreprocessArgs a)
{
  extern __shared__ __align__(128) unsigned char smemRaw[];
  PreSmem&       sm   = *reinterpret_cast<PreSmem*>(smemRaw);
  const unsigned tid  = threadIdx.x;
  const unsigned lane = tid & 31u, warp = tid >> 5;

  const uint32_t shElem  = SHFMT == VKGS_FORMAT_FLOAT32 ? 4u : (SHFMT == VKGS_FORMAT_FLOAT16 ? 2u : 1u);
  const uint32_t rgbaEl  = a.set.rgbaFormat == VKGS_FORMAT_FLOAT32 ? 4u : (a.set.rgbaFormat == VKGS_FORMAT_FLOAT16 ? 2u : 1u);
  const bool     hasSh   = a.set.sh != nullptr && a.set.shDegree > 0 && a.fp.sh_degree > 0;
  const bool     sizeCul = a.opt.size_culling_mode == VKGS_SIZE_CULLING_ENABLED;
  constexpr bool gut       = GUT;  // separate instantiation: the 3DGUT code must not cost the 3DGS path registers
  const bool     needScale = sizeCul || a.surface != nullptr || gut;  // log-scales are staged with the centres
  const uint32_t ablate  = a.opt._reserved[0];
  const uint32_t tiles   = (a.set.count + PRE_TILE - 1) / PRE_TILE;

  // claim a tile (ticket order == look-back order) and start its bulk copies. Stage A: what the
  // cull needs; stage B: everything else — two barriers so the cull (and the early publication of
  // the tile's visible count) does not wait for the 46 KB of SH.
  // (thread 0 only. The tile index travels with barrier A: it is stored before the arrive (release) and
  //  read by everybody after the wait (acquire); with no tile left the barrier completes at once.)
  auto claimAndLoad = [&]() {
    const uint32_t t = atomicAdd(&a.counters->ticket[a.ticketSlot], 1u) - a.ticketBase;
    sm.tile          = t;
    if(t >= tiles)
      mbar_arrive_expect_tx(&sm.mbarA, 0u);
    else
    {
      const uint64_t f = static_cast<uint64_t>(t) * PRE_TILE;
      mbar_arrive_expect_tx(&sm.mbarA, PRE_TILE * 3 * 4 * (needScale ? 2u : 1u));
      bulk_copy_g2s(sm.center, a.set.centers + f * 3, PRE_TILE * 3 * 4, &sm.mbarA);
      if(needScale)
        bulk_copy_g2s(sm.scale, a.set.scales + f * 3, PRE_TILE * 3 * 4, &sm.mbarA);
      mbar_arrive_expect_tx(&sm.mbarB, PRE_TILE * 6 * 4 + PRE_TILE * 4 * rgbaEl + (hasSh ? PRE_TILE * 45 * shElem : 0u));
      bulk_copy_g2s(sm.cov, a.set.cov6 + f * 6, PRE_TILE * 6 * 4, &sm.mbarB);
      bulk_copy_g2s(sm.rgba, static_cast<const unsigned char*>(a.set.rgba) + f * 4 * rgbaEl, PRE_TILE * 4 * rgbaEl, &sm.mbarB);
      if(hasSh)
        bulk_copy_g2s(sm.sh, static_cast<const unsigned char*>(a.set.sh) + f * 45 * shElem, PRE_TILE * 45 * shElem, &sm.mbarB);
    }
  };

  // ---- persistent CTA: per-CTA fixed costs (barrier init, histogram zero / flush, launch) are paid
  // once, and the bulk copies of tile i+1 are issued as soon as tile i's shared data is dead, so
  // they fly during tile i's look-back / append ------------------------------------------------
  if(tid == 0)
  {
    mbar_init(&sm.mbarA, 1);
    mbar_init(&sm.mbarB, 1);
    mbar_fence_init();
  }
  for(int i = tid; i < 4 * 256; i += PRE_TILE)
    (&sm.hist[0][0])[i] = 0u;
  sm.u8Sh[tid]   = static_cast<float>(tid) / 255.0f * 2.0f - 1.0f;
  sm.u8Unit[tid] = static_cast<float>(tid) / 255.0f;
  static_assert(PRE_TILE == 256, "one table entry per thread");
  __syncthreads();
  if(tid == 0)
    claimAndLoad();
  // the append of a tile is finished one iteration later (see below): what this thread still owes
  bool     pendKeep = false;
  uint32_t pendKey = 0, pendId = 0, pendSlot = 0;
  uint32_t phase = 0;
  for(;; phase ^= 1u)
  {
  mbar_wait(&sm.mbarA, phase);
  const uint32_t tile = sm.tile;
  if(tile >= tiles)
    break;
  const uint64_t first = static_cast<uint64_t>(tile) * PRE_TILE;

  const uint64_t id    = first + tid;
  const bool     inSet = id < a.set.count;

  // ---- K1: dist.comp.slang:55-167 ------------------------------------------------------------
  bool     keep = inSet;
  uint32_t key  = 0;
  float    c[4] = {sm.center[3 * tid + 0], sm.center[3 * tid + 1], sm.center[3 * tid + 2], 1.0f};
  if(keep)
  {
    float t[4], view[4], ndc[4];
    mulVecMat(c, a.fp.model, t);
    mulVecMat(t, a.fp.view, view);
    mulVecMat(view, a.fp.proj, ndc);
    const float w = ndc[3];
    ndc[0] = ndc[0] / w, ndc[1] = ndc[1] / w, ndc[2] = ndc[2] / w;
    const float depth = ndc[2];
    if(GUT && a.opt.frustum_culling_mode == VKGS_FRUSTUM_CULLING_AT_DIST && a.opt.camera_model == VKGS_CAMERA_FISHEYE)
    {
      // dist.comp.slang:75-90: fisheye projection of the view-space centre (z flipped), then the depth range
      const float pos[3] = {1.0f * view[0], 1.0f * view[1], -1.0f * view[2]};
      float       projected[2];
      if(!gutProjectFisheye(pos, a.fp, projected))
        keep = false;
      if(ndc[2] < 0.f - a.fp.frustum_dilation || ndc[2] > 1.0f)
        keep = false;
    }
    else if(a.opt.frustum_culling_mode == VKGS_FRUSTUM_CULLING_AT_DIST)
    {
      const float clip = 1.0f + a.fp.frustum_dilation;
      if(fabsf(ndc[0]) > clip || fabsf(ndc[1]) > clip || ndc[2] < 0.f - a.fp.frustum_dilation || ndc[2] > 1.0f)
        keep = false;
    }
    if(keep && sizeCul)
    {
      // (fixed-sequence exp shared with the oracle: the cull decision is bit-exact at the sizeCullingMinPixels threshold)
      const float sx = expfExact(sm.scale[3 * tid + 0]) * a.fp.splat_scale, sy = expfExact(sm.scale[3 * tid + 1]) * a.fp.splat_scale,
                  sz = expfExact(sm.scale[3 * tid + 2]) * a.fp.splat_scale;
      const float  radius = fmaxf(sx, fmaxf(sy, sz));
      float        extent = radius * 2.8284271247f * 2.0f;
      const float* m      = a.fp.model;
      const float  l0 = sqrtf(m[0] * m[0] + m[1] * m[1] + m[2] * m[2]), l1 = sqrtf(m[4] * m[4] + m[5] * m[5] + m[6] * m[6]),
                  l2 = sqrtf(m[8] * m[8] + m[9] * m[9] + m[10] * m[10]);
      extent *= fmaxf(l0, fmaxf(l1, l2));
      const float viewDist = fabsf(view[2]);
      if(viewDist > 0.0001f)
      {
        const float maxFocal        = fmaxf(fabsf(a.fp.focal[0]), fabsf(a.fp.focal[1]));
        const float projectedPixels = (extent * maxFocal) / viewDist;
        if(projectedPixels < a.fp.size_culling_min_pixels)
          keep = false;
      }
    }
    // A NaN depth (NaN / inf position: every cull comparison is false, the splat is kept) takes the canonical quiet NaN
    // 0x7FFFFFFF an arithmetic instruction of this GPU returns, AFTER the negation of the back-to-front order: its key is
    // 0xFFFFFFFF and it sorts last in both orders, whatever instruction the compiler picks for the negation.
    float kd = a.opt.front_to_back ? depth : -depth;
    if(kd != kd)
      kd = __uint_as_float(0x7fffffffu);
    key = encodeMinMaxFp32(kd);
  }

  // ---- deterministic append, part 1: publish this tile's visible count as early as possible ----
  // (decoupled look-back: successors only need the aggregate; publishing it before the heavy
  // per-splat stage means nobody ever waits on this tile's projection / SH work)
  const unsigned ballot   = __ballot_sync(FULL_MASK, keep);
  const uint32_t warpRank = __popc(ballot & ((1u << lane) - 1u));
  if(lane == 0)
    sm.warpScan[warp] = __popc(ballot);
  if(keep)
  {
    atomicAdd(&sm.hist[0][key & 0xffu], 1u);
    atomicAdd(&sm.hist[1][(key >> 8) & 0xffu], 1u);
    atomicAdd(&sm.hist[2][(key >> 16) & 0xffu], 1u);
    atomicAdd(&sm.hist[3][key >> 24], 1u);
  }
  __syncthreads();
  uint32_t tileTotal = 0, chainBase = 0;
  if(warp == 0)
  {
    const uint32_t cnt = lane < NWARPS ? sm.warpScan[lane] : 0u;
    const uint32_t inc = warp_inclusive_scan(cnt, lane);
    if(lane < NWARPS)
      sm.warpScan[lane] = inc - cnt;
    tileTotal = __shfl_sync(FULL_MASK, inc, NWARPS - 1);
    // (tile 0 of a chained launch continues after the pairs appended by the earlier instances: those
    //  launches completed, stream order, so counters->visible is final)
    if(tile == 0 && a.chained)
      chainBase = a.counters->visible;
    if(lane == 0 && !(ablate & 1u))
      lb_store(a.status + tile, lb_pack(a.epoch, tile == 0 ? LB_INCLUSIVE : LB_AGGREGATE, chainBase + tileTotal));
  }

  // ---- K5: per-splat projection + colour (threedgs_raster.mesh.slang:161-289) -------------------
  // finish the previous tile's append: its exclusive prefix was resolved by warp 0 while this
  // tile's centers were in flight (nobody waits on a look-back)
  if(pendKeep)
  {
    const uint32_t slot = sm.basePrefix[phase ^ 1u] + pendSlot;
    a.keys[slot]        = pendKey;
    a.ids[slot]         = pendId;
  }
  mbar_wait(&sm.mbarB, phase);
  if(GUT && keep)
  {
    // VK3DGUT: colour (+ SH) first, then the unscented-transform projection (threedgut_raster.mesh.slang:115-218)
    float4 col = loadRgba(sm.rgba, tid, a.set.rgbaFormat, sm.u8Unit);
    if(a.opt.show_sh_only)
      col.x = col.y = col.z = 0.5f;
    if(hasSh)
    {
      float       d[3] = {c[0] - a.camModel[0], c[1] - a.camModel[1], c[2] - a.camModel[2]};
      const float dinv = 1.0f / sqrtf(d[0] * d[0] + d[1] * d[1] + d[2] * d[2]);
      d[0] *= dinv, d[1] *= dinv, d[2] *= dinv;
      float rad[3];
      shRadiance<SHFMT>(sm.sh, 45 * static_cast<int>(tid), min(a.set.shDegree, a.fp.sh_degree), d[0], d[1], d[2], rad, sm.u8Sh);
      col.x += rad[0], col.y += rad[1], col.z += rad[2];
    }
    float4         rec6[6];
    uint32_t       bb0 = 1u, bb1 = 0u;
    const float4   rq  = *reinterpret_cast<const float4*>(a.set.rotations + 4 * id);
    const bool     ok  = gutProjectSplat(a, c, rq, sm.scale + 3 * tid, col, rec6, bb0, bb1);
    if(!ok)
      bb0 = 1u, bb1 = 0u;
    else
    {
      float4* rec = reinterpret_cast<float4*>(a.records + (a.idBase + id) * GUT_RECORD_WORDS);
      rec6[1].x = clampFinite(rec6[1].x), rec6[1].y = clampFinite(rec6[1].y), rec6[1].z = clampFinite(rec6[1].z);  // (see the 3DGS record)
#pragma unroll
      for(int k = 0; k < 6; k++)
        rec[k] = rec6[k];
    }
    a.bboxes[a.idBase + id] = make_uint2(bb0, bb1);
  }
  else if(!GUT && keep)
  {
    float4   col   = loadRgba(sm.rgba, tid, a.set.rgbaFormat, sm.u8Unit);
    bool     valid = !(col.w < a.fp.alpha_cull_threshold);  // mesh.slang:165
    float    cx = 0.f, cy = 0.f, w1x = 0.f, w1y = 0.f, w2x = 0.f, w2y = 0.f, ndcDepth = 0.f;
    uint32_t bb0 = 1u, bb1 = 0u;  // empty: x1 < x0
    if(valid)
    {
      float view[4], clip[4];
      mulVecMat(c, a.mv, view);
      mulVecMat(view, a.fp.proj, clip);
      if(a.opt.frustum_culling_mode == VKGS_FRUSTUM_CULLING_AT_RASTER)
      {
        const float lim = (1.0f + a.fp.frustum_dilation) * clip[3];
        if(fabsf(clip[0]) > lim || fabsf(clip[1]) > lim || clip[2] < (0.0f - a.fp.frustum_dilation) * clip[3] || clip[2] > clip[3])
          valid = false;
      }
      if(valid)
      {
        if(a.opt.show_sh_only)
          col.x = col.y = col.z = 0.5f;
        if(hasSh && !(ablate & 8u))
        {
          float       d[3] = {c[0] - a.camModel[0], c[1] - a.camModel[1], c[2] - a.camModel[2]};
          const float dinv = 1.0f / sqrtf(d[0] * d[0] + d[1] * d[1] + d[2] * d[2]);
          d[0] *= dinv, d[1] *= dinv, d[2] *= dinv;
          const uint32_t degree = min(a.set.shDegree, a.fp.sh_degree);
          float          rad[3];
          shRadiance<SHFMT>(sm.sh, 45 * static_cast<int>(tid), degree, d[0], d[1], d[2], rad, sm.u8Sh);
          col.x += rad[0], col.y += rad[1], col.z += rad[2];
        }
        // threedgsCovarianceProjection, threedgs.h.slang:26-56
        const float* s6      = sm.cov + 6 * tid;
        const float  C[3][3] = {{s6[0], s6[1], s6[2]}, {s6[1], s6[3], s6[4]}, {s6[2], s6[4], s6[5]}};
        const float  fx = a.fp.focal[0], fy = a.fp.focal[1];
        const float  s   = 1.0f / (view[2] * view[2]);
        const float  J00 = fx / view[2], J02 = -(fx * view[0]) * s;
        const float  J11 = fy / view[2], J12 = -(fy * view[1]) * s;
        float        T[2][3];
#pragma unroll
        for(int j = 0; j < 3; j++)
        {
          T[0][j] = J00 * a.mv[4 * j + 0] + J02 * a.mv[4 * j + 2];
          T[1][j] = J11 * a.mv[4 * j + 1] + J12 * a.mv[4 * j + 2];
        }
        float TC[2][3];
#pragma unroll
        for(int i = 0; i < 2; i++)
#pragma unroll
          for(int j = 0; j < 3; j++)
            TC[i][j] = (T[i][0] * C[0][j] + T[i][1] * C[1][j]) + T[i][2] * C[2][j];
        float cov0 = (TC[0][0] * T[0][0] + TC[0][1] * T[0][1]) + TC[0][2] * T[0][2];
        float cov1 = (TC[0][0] * T[1][0] + TC[0][1] * T[1][1]) + TC[0][2] * T[1][2];
        float cov2 = (TC[1][0] * T[1][0] + TC[1][1] * T[1][1]) + TC[1][2] * T[1][2];

        // threedgsProjectedExtentBasis, threedgs.h.slang:60-121
        float detOrig = 0.0f;
        if(a.opt.ms_antialiasing)
          detOrig = cov0 * cov2 - cov1 * cov1;
        cov0 += 0.3f;
        cov2 += 0.3f;
        if(a.opt.ms_antialiasing)
        {
          const float detBlur = cov0 * cov2 - cov1 * cov1;
          col.w *= sqrtf(fmaxf(detOrig / detBlur, 0.0f));
        }
        const float D          = cov0 * cov2 - cov1 * cov1;
        const float trace      = cov0 + cov2;
        const float traceOver2 = 0.5f * trace;
        const float term2      = sqrtf(fmaxf(0.1f, traceOver2 * traceOver2 - D));
        float       ev1        = traceOver2 + term2;
        float       ev2        = traceOver2 - term2;
        if(ev2 <= 0.0f)
          valid = false;
        if(valid)
        {
          if(a.opt.point_cloud_mode)
            ev1 = ev2 = 0.2f;
          float       e1x  = (fabsf(cov1) < 0.001f) ? 1.0f : cov1;
          float       e1y  = ev1 - cov0;
          const float einv = 1.0f / sqrtf(e1x * e1x + e1y * e1y);
          e1x *= einv, e1y *= einv;
          const float e2x = e1y, e2y = -e1x;
          const float sqrt8 = 2.8284271247461903f;
          const float m1 = fminf(sqrt8 * sqrtf(ev1), 2048.0f), m2 = fminf(sqrt8 * sqrtf(ev2), 2048.0f);
          const float b1x = e1x * a.fp.splat_scale * m1, b1y = e1y * a.fp.splat_scale * m1;
          const float b2x = e2x * a.fp.splat_scale * m2, b2y = e2y * a.fp.splat_scale * m2;

          // quad centre in pixels + affine fragPos basis (mesh.slang:201,276-289)
          const float nz = clip[2] / clip[3];
          ndcDepth       = nz;
          cx             = ((clip[0] / clip[3]) * 0.5f + 0.5f) * a.fp.viewport[0];
          cy             = ((clip[1] / clip[3]) * 0.5f + 0.5f) * a.fp.viewport[1];
          const float n1 = b1x * b1x + b1y * b1y, n2 = b2x * b2x + b2y * b2y;
          const float k1 = sqrt8 / n1, k2 = sqrt8 / n2;
          w1x = b1x * k1, w1y = b1y * k1, w2x = b2x * k2, w2y = b2y * k2;
          valid = (nz >= 0.0f && nz <= 1.0f);
          if(!(n1 > 0.0f) || !(n2 > 0.0f) || !isfinite(k1) || !isfinite(k2) || !isfinite(cx) || !isfinite(cy))
            valid = false;

          if(valid)
          {
            // Conservative pixel bounding box of {A <= 8 and opacity > 1/255}: the ellipse with
            // semi-axes b1,b2, shrunk when the splat's own alpha reaches 1/255 before A = 8.
            float hx = sqrtf(b1x * b1x + b2x * b2x), hy = sqrtf(b1y * b1y + b2y * b2y);
            if(!a.opt.disable_opacity_gaussian)
            {
              const float amax = 2.0f * logf(255.0f * col.w) * 1.00001f + 1e-4f;  // A < 2 ln(255 a)
              if(!(amax > 0.0f))
                valid = false;
              else if(amax < 8.0f)
              {
                const float r = sqrtf(amax * 0.125f);
                hx *= r, hy *= r;
              }
            }
            hx = hx * 1.00001f + 0.01f, hy = hy * 1.00001f + 0.01f;
            const float W = a.fp.viewport[0], H = a.fp.viewport[1];
            const float fx0 = ceilf(cx - hx - 0.5f), fx1 = floorf(cx + hx - 0.5f);
            const float fy0 = ceilf(cy - hy - 0.5f), fy1 = floorf(cy + hy - 0.5f);
            if(!valid || fx1 < 0.0f || fy1 < 0.0f || fx0 > W - 1.0f || fy0 > H - 1.0f || fx1 < fx0 || fy1 < fy0)
              valid = false;
            else
            {
              const uint32_t x0 = static_cast<uint32_t>(fmaxf(fx0, 0.0f)), x1 = static_cast<uint32_t>(fminf(fx1, W - 1.0f));
              const uint32_t y0 = static_cast<uint32_t>(fmaxf(fy0, 0.0f)), y1 = static_cast<uint32_t>(fminf(fy1, H - 1.0f));
              bb0 = x0 | (y0 << 16);
              bb1 = x1 | (y1 << 16);
            }
          }
        }
      }
    }
    if(!valid)
      bb0 = 1u, bb1 = 0u;
    float4* rec = reinterpret_cast<float4*>(a.records + (a.idBase + id) * RECORD_WORDS);
    if(!(ablate & 4u))
    {
    // The blend composites a discarded fragment as 0 * colour: a non-finite colour (corrupt input, outside the parity
    // contract) is clamped to +-FLT_MAX so that it cannot turn pixels the splat does not cover into NaN. Finite colours
    // (everything the parity contract covers) pass through bit for bit.
    rec[0]      = make_float4(cx, cy, w1x, w1y);
    rec[1]      = make_float4(w2x, w2y, clampFinite(col.x), clampFinite(col.y));
    rec[2]      = make_float4(clampFinite(col.z), col.w, __uint_as_float(bb0), __uint_as_float(bb1));
    a.bboxes[a.idBase + id] = make_uint2(bb0, bb1);
    }
    if(a.surface && valid)
    {
      // NEED_SURFACE_INFO: world normal of the max-density plane (threedgs_raster.mesh.slang:209-233,
      // computeEllipsoidNormalMaxDensityPlane threedgrt.h.slang:358-418), same operation order as
      // orc_splat_normal; the NDC depth rides along for the depth pick of the fragment stage
      const float4 rq   = *reinterpret_cast<const float4*>(a.set.rotations + 4 * id);  // w x y z
      const float  scl[3] = {expf(sm.scale[3 * tid + 0]), expf(sm.scale[3 * tid + 1]), expf(sm.scale[3 * tid + 2])};
      const float  rinv = 1.0f / sqrtf(((rq.x * rq.x + rq.y * rq.y) + rq.z * rq.z) + rq.w * rq.w);
      const float  x = rq.y * rinv, y = rq.z * rinv, z = rq.w * rinv, w = rq.x * rinv;
      const float  xx = x * x, yy = y * y, zz = z * z, xy = x * y, xz = x * z, yz = y * z, wx = w * x, wy = w * y, wz = w * z;
      const float  inv[3][3] = {{1.0f - 2.0f * (yy + zz), 2.0f * (xy - wz), 2.0f * (xz + wy)},
                                {2.0f * (xy + wz), 1.0f - 2.0f * (xx + zz), 2.0f * (yz - wx)},
                                {2.0f * (xz - wy), 2.0f * (yz + wx), 1.0f - 2.0f * (xx + yy)}};
      float dir[3] = {c[0] - a.camModel[0], c[1] - a.camModel[1], c[2] - a.camModel[2]};
      const float dinv = 1.0f / sqrtf((dir[0] * dir[0] + dir[1] * dir[1]) + dir[2] * dir[2]);
      dir[0] *= dinv, dir[1] *= dinv, dir[2] *= dinv;
      const float local[3] = {a.camModel[0] - c[0], a.camModel[1] - c[1], a.camModel[2] - c[2]};
      const float thr      = a.fp.thin_particle_threshold;
      const int   s0 = scl[0] < thr, s1 = scl[1] < thr, s2 = scl[2] < thr, nsmall = s0 + s1 + s2;
      float       n[3];
      if(nsmall == 0)
      {
        float canon[3], scaled[3];
#pragma unroll
        for(int j = 0; j < 3; j++)
          canon[j] = (local[0] * inv[0][j] + local[1] * inv[1][j]) + local[2] * inv[2][j];
#pragma unroll
        for(int j = 0; j < 3; j++)
          scaled[j] = canon[j] * (1.0f / (scl[j] * scl[j]));
#pragma unroll
        for(int j = 0; j < 3; j++)
          n[j] = (scaled[0] * inv[j][0] + scaled[1] * inv[j][1]) + scaled[2] * inv[j][2];
        const float rs = 1.0f / sqrtf((n[0] * n[0] + n[1] * n[1]) + n[2] * n[2]);
        n[0] *= rs, n[1] *= rs, n[2] *= rs;
        if((n[0] * local[0] + n[1] * local[1]) + n[2] * local[2] < 0.0f)
          n[0] = -n[0], n[1] = -n[1], n[2] = -n[2];
      }
      else if(nsmall == 1)
      {
        // row (axis) of the rotation matrix = column of its transpose; selects, not dynamic indexing
        n[0] = s0 ? inv[0][0] : (s1 ? inv[0][1] : inv[0][2]);
        n[1] = s0 ? inv[1][0] : (s1 ? inv[1][1] : inv[1][2]);
        n[2] = s0 ? inv[2][0] : (s1 ? inv[2][1] : inv[2][2]);
        if((n[0] * local[0] + n[1] * local[1]) + n[2] * local[2] < 0.0f)
          n[0] = -n[0], n[1] = -n[1], n[2] = -n[2];
      }
      else
        n[0] = -dir[0], n[1] = -dir[1], n[2] = -dir[2];
      const float n4[4] = {n[0], n[1], n[2], 0.0f};
      float       nw[4];
      mulVecMat(n4, a.fp.model, nw);
      const float winv = 1.0f / sqrtf((nw[0] * nw[0] + nw[1] * nw[1]) + nw[2] * nw[2]);
      float       nq[3] = {nw[0] * winv, nw[1] * winv, nw[2] * winv};
      if(a.opt.quantize_normals)  // QUANTIZE_NORMALS: the fragment stage sees the normal through its 2x16-bit octahedral code
        octQuantizeNormal(nq);
      a.surface[a.idBase + id] = make_float4(nq[0], nq[1], nq[2], ndcDepth);
    }
  }


  // every thread is done with this tile's shared data: claim the next tile and start its copies now
  __syncthreads();
  if(tid == 0)
    claimAndLoad();
  pendKeep = keep, pendKey = key, pendId = a.idBase + static_cast<uint32_t>(id), pendSlot = sm.warpScan[warp] + warpRank;

  // ---- deterministic append, part 2: resolve the exclusive prefix (predecessors published long ago).
  // Warp 0 only, and nobody waits for it: the (key,id) writes of this tile happen in the next
  // iteration, after that iteration's first barrier, so the look-back's L2 round trips overlap the
  // bulk copies of the next tile.
  if(warp == 0)
  {
    uint32_t excl = chainBase;  // (non-zero for tile 0 of a chained launch only)
    if(ablate & 1u)
    {
      if(lane == 0)
        excl = atomicAdd(&a.counters->visible, tileTotal);
    }
    else if(tile != 0)
    {
      excl = lb_lookback_warp(a.status, tile, a.epoch);
      if(lane == 0)
        lb_store(a.status + tile, lb_pack(a.epoch, LB_INCLUSIVE, excl + tileTotal));
    }
    if(lane == 0)
    {
      sm.basePrefix[phase] = excl;
      // the tile holding the last splat knows V once its prefix is resolved
      if(first + PRE_TILE >= a.set.count && !(ablate & 1u))
        a.counters->visible = excl + tileTotal;
    }
  }
  }  // persistent tile loop

  // this CTA has no tile left: once every CTA of the launch is here, the first sort pass may start its prologue
  pdl_launch_dependents();
  // the last tile's append
  __syncthreads();
  if(pendKeep)
  {
    const uint32_t slot = sm.basePrefix[phase ^ 1u] + pendSlot;
    a.keys[slot]        = pendKey;
    a.ids[slot]         = pendId;
  }
  // flush the digit histograms of the four sort passes (only bins this CTA touched)
  for(int i = tid; i < 4 * 256; i += PRE_TILE)
  {
    const uint32_t v = (&sm.hist[0][0])[i];
    if(v && !(ablate & 2u))
      atomicAdd(&a.counters->depthHist[0][0] + i, v);
  }
}

}  // namespace

namespace {
int g_preSms = 0, g_prePerSm[3] = {0, 0, 0};  // SM count, resident CTAs of k_preprocess<fmt> per SM
int g_preCapEnv = 0;                          // VKGS_PRE_CTAS_PER_SM (tuning experiments)
}

void initPreprocessKernels()
{
  const int smem = static_cast<int>(sizeof(PreSmem));
  cudaFuncSetAttribute(k_preprocess<VKGS_FORMAT_FLOAT32, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
  cudaFuncSetAttribute(k_preprocess<VKGS_FORMAT_FLOAT16, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
  cudaFuncSetAttribute(k_preprocess<VKGS_FORMAT_UINT8, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
  cudaFuncSetAttribute(k_preprocess<VKGS_FORMAT_FLOAT32, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
  cudaFuncSetAttribute(k_preprocess<VKGS_FORMAT_FLOAT16, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
  cudaFuncSetAttribute(k_preprocess<VKGS_FORMAT_UINT8, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
  int dev = 0;
  cudaGetDevice(&dev);
  cudaDeviceGetAttribute(&g_preSms, cudaDevAttrMultiProcessorCount, dev);
  cudaOccupancyMaxActiveBlocksPerMultiprocessor(&g_prePerSm[0], k_preprocess<VKGS_FORMAT_FLOAT32, false>, PRE_TILE, smem);
  cudaOccupancyMaxActiveBlocksPerMultiprocessor(&g_prePerSm[1], k_preprocess<VKGS_FORMAT_FLOAT16, false>, PRE_TILE, smem);
  cudaOccupancyMaxActiveBlocksPerMultiprocessor(&g_prePerSm[2], k_preprocess<VKGS_FORMAT_UINT8, false>, PRE_TILE, smem);
  if(const char* e = getenv("VKGS_PRE_CTAS_PER_SM"))
    g_preCapEnv = atoi(e);
}

uint32_t preprocessGrid(const PreprocessArgs& args)
{
  const uint32_t tiles = (args.set.count + PRE_TILE - 1) / PRE_TILE;
  const uint32_t fmt   = args.set.shFormat <= VKGS_FORMAT_UINT8 ? args.set.shFormat : 0u;
  uint32_t       per   = g_prePerSm[fmt] > 0 ? static_cast<uint32_t>(g_prePerSm[fmt]) : 3u;
  const uint32_t want  = g_preCapEnv > 0 ? static_cast<uint32_t>(g_preCapEnv) : args.ctasPerSm;
  if(want > 0 && want < per)
    per = want;
  const uint32_t cap = static_cast<uint32_t>(g_preSms > 0 ? g_preSms : 148) * per;
  return tiles < cap ? tiles : cap;
}

void launchPreprocess(const PreprocessArgs& args, cudaStream_t stream)
{
  const uint32_t grid = preprocessGrid(args);
  if(grid == 0)
    return;
  const size_t smem = sizeof(PreSmem);
  const bool   gut  = args.opt.pipeline == VKGS_PIPELINE_3DGUT;
  switch(args.set.shFormat)
  {
    case VKGS_FORMAT_FLOAT16:
      if(gut)
        k_preprocess<VKGS_FORMAT_FLOAT16, true><<<grid, PRE_TILE, smem, stream>>>(args);
      else
        k_preprocess<VKGS_FORMAT_FLOAT16, false><<<grid, PRE_TILE, smem, stream>>>(args);
      break;
    case VKGS_FORMAT_UINT8:
      if(gut)
        k_preprocess<VKGS_FORMAT_UINT8, true><<<grid, PRE_TILE, smem, stream>>>(args);
      else
        k_preprocess<VKGS_FORMAT_UINT8, false><<<grid, PRE_TILE, smem, stream>>>(args);
      break;
    default:
      if(gut)
        k_preprocess<VKGS_FORMAT_FLOAT32, true><<<grid, PRE_TILE, smem, stream>>>(args);
      else
        k_preprocess<VKGS_FORMAT_FLOAT32, false><<<grid, PRE_TILE, smem, stream>>>(args);
      break;
  }
}

}  // namespace vkgs

// Host instantiation of the normal quantiser the kernel runs (parity pin without a GPU).
extern "C" VKGS_API int vkgs_quantize_normals_host(const float* normals_in, float* normals_out, uint64_t count)
{
  if(!normals_in || !normals_out)
    return VKGS_ERR_INVALID_ARGUMENT;
  for(uint64_t i = 0; i < count; i++)
  {
    float n[3] = {normals_in[3 * i], normals_in[3 * i + 1], normals_in[3 * i + 2]};
    vkgs::octQuantizeNormal(n);
    normals_out[3 * i] = n[0], normals_out[3 * i + 1] = n[1], normals_out[3 * i + 2] = n[2];
  }
  return VKGS_OK;
}

// Host instantiations of the fixed-sequence elementary functions the kernels run (parity pin without a GPU):
// which = 0 expfExact(a), 1 atan2fYposExact(a = y > 0, b = x), 2 acosfExact(a), 3 sincosfExact(a) -> out = sin, out2 = cos.
extern "C" VKGS_API int vkgs_exact_math_host(uint32_t which, const float* a, const float* b, float* out, float* out2, uint64_t count)
{
  if(!a || !out || which > 3u || (which == 1u && !b) || (which == 3u && !out2))
    return VKGS_ERR_INVALID_ARGUMENT;
  for(uint64_t i = 0; i < count; i++)
  {
    switch(which)
    {
      case 0: out[i] = vkgs::expfExact(a[i]); break;
      case 1: out[i] = vkgs::atan2fYposExact(a[i], b[i]); break;
      case 2: out[i] = vkgs::acosfExact(a[i]); break;
      default: vkgs::sincosfExact(a[i], out[i], out2[i]); break;
    }
  }
  return VKGS_OK;
}
