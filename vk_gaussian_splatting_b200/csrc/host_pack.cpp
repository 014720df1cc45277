// host_pack.cpp — RAM -> VRAM layout packing on the host.
//
// Replaces SplatSetVk::initDataBuffers (src/splat_set_vk.cpp:188-480): the reference builds the
// device arrays with host parallel loops (START_PAR_LOOP, src/utilities.h:52-59) and uploads
// them; so does this. Layouts:
//   centers  3 x f32                              (:228-232)
//   cov6     upper triangle of (R S)(R S)^T       (:263-288), R = mat3(normalize(q)), S = exp(scale)
//   rgba     clamp(0.5 + C0*f_dc), clamp(sigmoid(opacity))   (:313-345), f32 / f16 / u8
//   sh       45 elements, coefficient-major, RGB inner: dst[3k+c] = f_rest[15c+k]   (:396-435)
// Arithmetic follows glm's operation order (quat normalize, mat3_cast, mat3*mat3) so results are
// bit-identical to the reference's host code; pinned by tests/golden/glm_golden.json.
#include "host_pack.hpp"

#include <algorithm>
#include <cmath>
#include <cstring>
#include <thread>

namespace vkgs {

uint32_t formatSize(uint32_t format)
{
  switch(format)
  {
    case VKGS_FORMAT_FLOAT32:
      return 4;
    case VKGS_FORMAT_FLOAT16:
      return 2;
    case VKGS_FORMAT_UINT8:
      return 1;
    default:
      return 0;
  }
}

uint8_t toUint8(float v, float rangeMin, float rangeMax)
{
  const float normalized = (v - rangeMin) / (rangeMax - rangeMin);
  const float q          = std::clamp(std::round(normalized * 255.0f), 0.0f, 255.0f);
  // (a NaN survives the clamp and its conversion is undefined behaviour in the reference's expression; x86 yields 0: stated)
  return q == q ? static_cast<uint8_t>(q) : uint8_t(0);
}

// glm::detail::toFloat16: round-half-up on the 13 dropped mantissa bits, denormals by shifting.
uint16_t packHalf(float f)
{
  uint32_t bits;
  std::memcpy(&bits, &f, 4);
  const int i    = static_cast<int>(bits);
  const int sign = (i >> 16) & 0x8000;
  int       exp  = ((i >> 23) & 0xff) - 112;
  int       man  = i & 0x007fffff;
  if(exp <= 0)
  {
    if(exp < -10)
      return static_cast<uint16_t>(sign);
    man = (man | 0x00800000) >> (1 - exp);
    if(man & 0x1000)
      man += 0x2000;
    return static_cast<uint16_t>(sign | (man >> 13));
  }
  if(exp == 0xff - 112)
  {
    if(man == 0)
      return static_cast<uint16_t>(sign | 0x7c00);
    man >>= 13;
    return static_cast<uint16_t>(sign | 0x7c00 | man | (man == 0));
  }
  if(man & 0x1000)
  {
    man += 0x2000;
    if(man & 0x00800000)
    {
      man = 0;
      exp += 1;
    }
  }
  if(exp > 30)
    return static_cast<uint16_t>(sign | 0x7c00);
  return static_cast<uint16_t>(sign | (exp << 10) | (man >> 13));
}

namespace {

template <typename F>
void parallelFor(uint64_t n, F&& fn)
{
  const uint64_t grain = 8192;  // same batch size as START_PAR_LOOP
  unsigned       nt    = std::max(1u, std::thread::hardware_concurrency());
  if(n <= grain || nt == 1)
  {
    fn(uint64_t(0), n);
    return;
  }
  const uint64_t batches = (n + grain - 1) / grain;
  nt                     = static_cast<unsigned>(std::min<uint64_t>(nt, batches));
  std::vector<std::thread> pool;
  pool.reserve(nt);
  for(unsigned t = 0; t < nt; t++)
  {
    const uint64_t b0 = batches * t / nt, b1 = batches * (t + 1) / nt;
    pool.emplace_back([&fn, b0, b1, grain, n]() { fn(b0 * grain, std::min(n, b1 * grain)); });
  }
  for(auto& th : pool)
    th.join();
}

inline void storeElem(uint32_t format, void* dst, uint64_t index, float v, float lo, float hi)
{
  if(format == VKGS_FORMAT_FLOAT32)
    static_cast<float*>(dst)[index] = v;
  else if(format == VKGS_FORMAT_FLOAT16)
    static_cast<uint16_t*>(dst)[index] = packHalf(v);
  else
    static_cast<uint8_t*>(dst)[index] = toUint8(v, lo, hi);
}

inline void covariance6(const float* scaleLog, const float* q, float* out)
{
  const float sx = std::exp(scaleLog[0]), sy = std::exp(scaleLog[1]), sz = std::exp(scaleLog[2]);
  float       w = q[0], x = q[1], y = q[2], z = q[3];
  const float len = std::sqrt((w * w + x * x) + (y * y + z * z));
  if(len <= 0.0f)
  {
    w = 1.0f, x = y = z = 0.0f;
  }
  else
  {
    const float inv = 1.0f / len;
    w *= inv, x *= inv, y *= inv, z *= inv;
  }
  const float xx = x * x, yy = y * y, zz = z * z, xz = x * z, xy = x * y, yz = y * z, wx = w * x, wy = w * y, wz = w * z;
  // columns of R*S
  const float m00 = (1.0f - 2.0f * (yy + zz)) * sx, m01 = (2.0f * (xy + wz)) * sx, m02 = (2.0f * (xz - wy)) * sx;
  const float m10 = (2.0f * (xy - wz)) * sy, m11 = (1.0f - 2.0f * (xx + zz)) * sy, m12 = (2.0f * (yz + wx)) * sy;
  const float m20 = (2.0f * (xz + wy)) * sz, m21 = (2.0f * (yz - wx)) * sz, m22 = (1.0f - 2.0f * (xx + yy)) * sz;
  // (RS)(RS)^T, summed over the three columns in order
  out[0] = m00 * m00 + m10 * m10 + m20 * m20;
  out[1] = m00 * m01 + m10 * m11 + m20 * m21;
  out[2] = m00 * m02 + m10 * m12 + m20 * m22;
  out[3] = m01 * m01 + m11 * m11 + m21 * m21;
  out[4] = m01 * m02 + m11 * m12 + m21 * m22;
  out[5] = m02 * m02 + m12 * m12 + m22 * m22;
}

}  // namespace

int packSplatSet(const vkgs_splat_set_view& set, const vkgs_options& opt, uint64_t padTo, PackedSplatSet& out)
{
  if(!set.positions || !set.f_dc || !set.opacity || !set.scale || !set.rotation || set.count == 0)
    return VKGS_ERR_INVALID_ARGUMENT;
  if(set.f_rest_per_splat != 0 && set.f_rest_per_splat != 45)
    return VKGS_ERR_UNSUPPORTED;  // the reference's fetchers assume a stride of 45 (see header)
  if(set.f_rest_per_splat == 45 && !set.f_rest)
    return VKGS_ERR_INVALID_ARGUMENT;
  if(formatSize(opt.sh_format) == 0 || formatSize(opt.rgba_format) == 0)
    return VKGS_ERR_INVALID_ARGUMENT;

  const uint64_t n   = set.count;
  const uint64_t pad = (n + padTo - 1) / padTo * padTo;
  out.count          = n;
  out.paddedCount    = pad;
  out.shDegree       = set.f_rest_per_splat == 45 ? 3 : 0;
  out.shFormat       = opt.sh_format;
  out.rgbaFormat     = opt.rgba_format;
  out.centers.assign(3 * pad, 0.0f);
  out.cov6.assign(6 * pad, 0.0f);
  out.scales.assign(3 * pad, 0.0f);
  const bool needRot = opt.surface_info || opt.pipeline == VKGS_PIPELINE_3DGUT;
  out.rotations.assign(needRot ? 4 * pad : 0, 0.0f);
  out.rgba.assign(4 * pad * formatSize(opt.rgba_format), 0);
  out.sh.assign(out.shDegree ? 45 * pad * formatSize(opt.sh_format) : 0, 0);

  parallelFor(n, [&](uint64_t begin, uint64_t end) {
    const float SH_C0 = 0.28209479177387814f;
    for(uint64_t i = begin; i < end; i++)
    {
      std::memcpy(&out.centers[3 * i], set.positions + 3 * i, 3 * sizeof(float));
      std::memcpy(&out.scales[3 * i], set.scale + 3 * i, 3 * sizeof(float));
      if(needRot)
        std::memcpy(&out.rotations[4 * i], set.rotation + 4 * i, 4 * sizeof(float));
      covariance6(set.scale + 3 * i, set.rotation + 4 * i, &out.cov6[6 * i]);
      const float* dc = set.f_dc + 3 * i;
      const float  c[4] = {std::clamp(0.5f + SH_C0 * dc[0], 0.0f, 1.0f), std::clamp(0.5f + SH_C0 * dc[1], 0.0f, 1.0f),
                           std::clamp(0.5f + SH_C0 * dc[2], 0.0f, 1.0f),
                           std::clamp(1.0f / (1.0f + std::exp(-set.opacity[i])), 0.0f, 1.0f)};
      for(int k = 0; k < 4; k++)
        storeElem(opt.rgba_format, out.rgba.data(), 4 * i + k, c[k], 0.0f, 1.0f);
      if(out.shDegree)
      {
        const float* src = set.f_rest + 45 * i;
        for(int k = 0; k < 15; k++)
          for(int ch = 0; ch < 3; ch++)
            storeElem(opt.sh_format, out.sh.data(), 45 * i + 3 * k + ch, src[15 * ch + k], -1.0f, 1.0f);
      }
    }
  });
  return VKGS_OK;
}

}  // namespace vkgs
