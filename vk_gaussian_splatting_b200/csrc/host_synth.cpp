// host_synth.cpp — deterministic synthetic scenes in SplatSet layout.
//
// Counter-based splitmix64 so generation is order independent and multi-threaded:
//   positions ~ U([-1,1]^3); log-scale per axis ~ U(ln s0, ln 10 s0), s0 = 0.002 (1e6/N)^(1/3);
//   rotation = N(0,1)^4 (w,x,y,z, un-normalised like a .ply); opacity logit ~ U(-2,4);
//   f_dc ~ U(-1.5,1.5); f_rest ~ N(0, 0.1^2) x 45 (absent for degree 0).
// Scenes are produced directly in RUB (already converted) coordinates.
#include <algorithm>
#include <cmath>
#include <cstdint>
#include <thread>
#include <vector>

#include "vkgs_b200.h"

namespace {

inline uint64_t mix64(uint64_t z)
{
  z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
  z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
  return z ^ (z >> 31);
}

struct Stream
{
  uint64_t base;
  Stream(uint64_t seed, uint64_t stream)
      : base(mix64(seed ^ (stream * 0xD1342543DE82EF95ull)))
  {
  }
  uint64_t bits(uint64_t idx) const { return mix64(base + (idx + 1) * 0x9E3779B97F4A7C15ull); }
  // uniform in [0,1) with 24 bits
  float uniform(uint64_t idx) const { return static_cast<float>(bits(idx) >> 40) * (1.0f / 16777216.0f); }
  // standard normal (Box-Muller on two counters, evaluated in double)
  float normal(uint64_t idx) const
  {
    const double u1 = (static_cast<double>(bits(2 * idx) >> 11) + 1.0) * (1.0 / 9007199254740992.0);
    const double u2 = static_cast<double>(bits(2 * idx + 1) >> 11) * (1.0 / 9007199254740992.0);
    return static_cast<float>(std::sqrt(-2.0 * std::log(u1)) * std::cos(6.283185307179586476925 * u2));
  }
};

}  // namespace

extern "C" int vkgs_synth_scene(uint64_t n, uint32_t sh_degree, uint64_t seed, float* positions, float* f_dc, float* f_rest,
                                float* opacity, float* scale, float* rotation)
{
  if(n == 0 || !positions || !f_dc || !opacity || !scale || !rotation)
    return VKGS_ERR_INVALID_ARGUMENT;
  if(sh_degree != 0 && sh_degree != 3)
    return VKGS_ERR_UNSUPPORTED;
  if(sh_degree == 3 && !f_rest)
    return VKGS_ERR_INVALID_ARGUMENT;

  const Stream sPos(seed, 1), sScale(seed, 2), sRot(seed, 3), sOp(seed, 4), sDc(seed, 5), sRest(seed, 6);
  const double s0     = 0.002 * std::cbrt(1.0e6 / static_cast<double>(n));
  const float  lnLo   = static_cast<float>(std::log(s0));
  const float  lnSpan = static_cast<float>(std::log(10.0));

  unsigned       nt    = std::max(1u, std::thread::hardware_concurrency());
  const uint64_t chunk = (n + nt - 1) / nt;
  std::vector<std::thread> pool;
  for(unsigned t = 0; t < nt; t++)
  {
    const uint64_t b = std::min(n, chunk * t), e = std::min(n, chunk * (t + 1));
    if(b >= e)
      break;
    pool.emplace_back([=]() {
      for(uint64_t i = b; i < e; i++)
      {
        for(int k = 0; k < 3; k++)
        {
          positions[3 * i + k] = 2.0f * sPos.uniform(3 * i + k) - 1.0f;
          scale[3 * i + k]     = lnLo + lnSpan * sScale.uniform(3 * i + k);
          f_dc[3 * i + k]      = 3.0f * sDc.uniform(3 * i + k) - 1.5f;
        }
        for(int k = 0; k < 4; k++)
          rotation[4 * i + k] = sRot.normal(4 * i + k);
        opacity[i] = 6.0f * sOp.uniform(i) - 2.0f;
        if(sh_degree == 3)
          for(int k = 0; k < 45; k++)
            f_rest[45 * i + k] = 0.1f * sRest.normal(45 * i + k);
      }
    });
  }
  for(auto& th : pool)
    th.join();
  return VKGS_OK;
}
