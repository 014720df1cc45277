// TEST INFRASTRUCTURE — runs the REFERENCE's own loader code on a scene file.
//
// Compiled (by `make -C oracle ref_loader`) together with the reference's third-party sources where
// they lie under /root/reference — 3rdparty/miniply/miniply.cpp and 3rdparty/spz/src/cc/*.cc — plus
// src/splat_set.h. Nothing is copied into this repository; the binary goes to oracle/_ref/.
// main() below repeats the call sequence of PlyLoaderAsync::innerLoad (src/ply_loader_async.cpp:291-453)
// for .ply and .spz (the reference file itself drags in Vulkan headers and cannot be compiled here);
// the .splat branch of the reference is a free function in an anonymous namespace of that file and is
// therefore pinned by the restated expected values instead (tests/test_loaders.py).
//
//   ref_loader dump <in.ply|in.spz> <out.bin>     SplatSet arrays as: u32 n, u32 restPerSplat, then
//                                                  positions, f_dc, f_rest, opacity, scale, rotation (f32)
//   ref_loader makespz <in.ply> <out.spz> <ver>   writes an .spz with the reference's spz library
#include <cstdint>
#include <cstdio>
#include <cstring>
#include <fstream>
#include <string>
#include <vector>

#include "miniply.h"
#include "load-spz.h"
#include "splat_set.h"

using vk_gaussian_splatting::SplatSet;

static bool loadPlyLikeReference(const std::string& filename, SplatSet& output)
{
  miniply::PLYReader reader(filename.c_str());
  if(!reader.valid())
    return false;
  uint32_t indices[45];
  bool     gsFound = false;
  while(reader.has_element() && !gsFound)
  {
    if(reader.element_is(miniply::kPLYVertexElement) && reader.load_element())
    {
      const uint32_t numVerts = reader.num_rows();
      if(numVerts == 0)
        continue;
      if(reader.find_properties(indices, 45, "f_rest_0", "f_rest_1", "f_rest_2", "f_rest_3", "f_rest_4", "f_rest_5",
                                "f_rest_6", "f_rest_7", "f_rest_8", "f_rest_9", "f_rest_10", "f_rest_11", "f_rest_12",
                                "f_rest_13", "f_rest_14", "f_rest_15", "f_rest_16", "f_rest_17", "f_rest_18",
                                "f_rest_19", "f_rest_20", "f_rest_21", "f_rest_22", "f_rest_23", "f_rest_24",
                                "f_rest_25", "f_rest_26", "f_rest_27", "f_rest_28", "f_rest_29", "f_rest_30", "f_rest_31",
                                "f_rest_32", "f_rest_33", "f_rest_34", "f_rest_35", "f_rest_36", "f_rest_37", "f_rest_38",
                                "f_rest_39", "f_rest_40", "f_rest_41", "f_rest_42", "f_rest_43", "f_rest_44"))
      {
        output.f_rest.resize(numVerts * 45);
        reader.extract_properties(indices, 45, miniply::PLYPropertyType::Float, output.f_rest.data());
      }
      if(reader.find_properties(indices, 3, "x", "y", "z"))
      {
        output.positions.resize(numVerts * 3);
        reader.extract_properties(indices, 3, miniply::PLYPropertyType::Float, output.positions.data());
      }
      if(reader.find_properties(indices, 1, "opacity"))
      {
        output.opacity.resize(numVerts);
        reader.extract_properties(indices, 1, miniply::PLYPropertyType::Float, output.opacity.data());
      }
      if(reader.find_properties(indices, 3, "scale_0", "scale_1", "scale_2"))
      {
        output.scale.resize(numVerts * 3);
        reader.extract_properties(indices, 3, miniply::PLYPropertyType::Float, output.scale.data());
      }
      if(reader.find_properties(indices, 4, "rot_0", "rot_1", "rot_2", "rot_3"))
      {
        output.rotation.resize(numVerts * 4);
        reader.extract_properties(indices, 4, miniply::PLYPropertyType::Float, output.rotation.data());
      }
      if(reader.find_properties(indices, 3, "f_dc_0", "f_dc_1", "f_dc_2"))
      {
        output.f_dc.resize(numVerts * 3);
        reader.extract_properties(indices, 3, miniply::PLYPropertyType::Float, output.f_dc.data());
      }
      gsFound = true;
    }
    reader.next_element();
  }
  if(gsFound)
    output.convertCoordinates(spz::CoordinateSystem::RDF, spz::CoordinateSystem::RUB);
  return gsFound;
}

static bool loadSpzLikeReference(const std::string& filename, SplatSet& output)
{
  spz::UnpackOptions options{.to = spz::CoordinateSystem::RUB};
  spz::GaussianCloud cloud = spz::loadSpz(filename, options);
  output.positions.swap(cloud.positions);
  output.rotation.resize(cloud.rotations.size());
  const uint32_t numSplats = uint32_t(output.positions.size() / 3);
  for(uint32_t i = 0; i < numSplats; i++)
  {
    const uint32_t offset       = i * 4;
    output.rotation[offset + 0] = cloud.rotations[offset + 3];
    output.rotation[offset + 1] = cloud.rotations[offset + 0];
    output.rotation[offset + 2] = cloud.rotations[offset + 1];
    output.rotation[offset + 3] = cloud.rotations[offset + 2];
  }
  output.scale.swap(cloud.scales);
  output.opacity.swap(cloud.alphas);
  output.f_dc = cloud.colors;
  if(numSplats == 0)
    return false;
  const size_t shCoefsCount = cloud.sh.size() / numSplats / 3;
  output.f_rest.resize(cloud.sh.size());
  for(size_t i = 0; i < numSplats; i++)
  {
    const size_t offset = i * shCoefsCount * 3;
    for(size_t j = 0; j < shCoefsCount; j++)
      output.f_rest[offset + j] = cloud.sh[(i * shCoefsCount + j) * 3];
    for(size_t j = 0; j < shCoefsCount; j++)
      output.f_rest[offset + shCoefsCount + j] = cloud.sh[(i * shCoefsCount + j) * 3 + 1];
    for(size_t j = 0; j < shCoefsCount; j++)
      output.f_rest[offset + shCoefsCount * 2 + j] = cloud.sh[(i * shCoefsCount + j) * 3 + 2];
  }
  return cloud.numPoints != 0;
}

static bool endsWith(const std::string& s, const char* e)
{
  const size_t n = std::strlen(e);
  return s.size() >= n && s.compare(s.size() - n, n, e) == 0;
}

int main(int argc, char** argv)
{
  if(argc < 4)
    return 2;
  const std::string mode = argv[1], in = argv[2], out = argv[3];
  if(mode == "dump")
  {
    SplatSet s;
    const bool ok = endsWith(in, ".spz") ? loadSpzLikeReference(in, s) : loadPlyLikeReference(in, s);
    if(!ok)
      return 1;
    FILE* f = std::fopen(out.c_str(), "wb");
    const uint32_t n = uint32_t(s.size()), rest = uint32_t(s.f_rest.size() / s.size());
    std::fwrite(&n, 4, 1, f);
    std::fwrite(&rest, 4, 1, f);
    for(const std::vector<float>* v : {&s.positions, &s.f_dc, &s.f_rest, &s.opacity, &s.scale, &s.rotation})
      std::fwrite(v->data(), 4, v->size(), f);
    std::fclose(f);
    return 0;
  }
  if(mode == "makespz")
  {
    // spz's own PLY reader + packer (3rdparty/spz/src/cc/load-spz.cc); PLY coordinates are RDF
    spz::GaussianCloud cloud = spz::loadSplatFromPly(in, spz::UnpackOptions{.to = spz::CoordinateSystem::RUB});
    if(cloud.numPoints == 0)
      return 1;
    spz::PackOptions po{.from = spz::CoordinateSystem::RUB};
    return spz::saveSpz(cloud, po, out) ? 0 : 1;
  }
  return 2;
}
