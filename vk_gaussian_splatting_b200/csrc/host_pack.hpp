// host_pack.hpp — RAM -> VRAM layout packing (host side), see host_pack.cpp
#pragma once
#include <cstdint>
#include <vector>

#include "vkgs_b200.h"

namespace vkgs {

// Device-layout arrays produced on the host, in the formats selected by vkgs_options.
struct PackedSplatSet
{
  uint64_t             count       = 0;
  uint32_t             shDegree    = 0;  // 0 or 3
  uint32_t             shFormat    = VKGS_FORMAT_FLOAT32;
  uint32_t             rgbaFormat  = VKGS_FORMAT_FLOAT32;
  uint64_t             paddedCount = 0;  // count rounded up to the preprocess tile (rows are zero padded)
  std::vector<float>   centers;          // 3 * padded
  std::vector<float>   cov6;             // 6 * padded
  std::vector<float>   scales;           // 3 * padded (log-space, read by size culling and the surface-info normals)
  std::vector<float>   rotations;        // 4 * padded raw quaternions (w,x,y,z), filled for options.surface_info / the 3DGUT pipeline only
  std::vector<uint8_t> rgba;             // 4 * padded * formatSize
  std::vector<uint8_t> sh;               // 45 * padded * formatSize (empty for degree 0)
};

uint32_t formatSize(uint32_t format);
uint16_t packHalf(float f);  // glm::packHalf1x16 semantics
uint8_t  toUint8(float v, float rangeMin, float rangeMax);

// Returns VKGS_OK or an error code.
int packSplatSet(const vkgs_splat_set_view& set, const vkgs_options& opt, uint64_t padTo, PackedSplatSet& out);

}  // namespace vkgs
