// C++ host driver over the C ABI, mirroring the reference's frame loop for the VK3DGSR path
// (nvapp::Application::drawFrame -> GaussianSplatting::onRender, src/gaussian_splatting.cpp:335,494):
//   load or synthesize a splat set -> initDataStorage -> per frame updateAndUploadFrameInfoUBO + onRender
//   -> the reference's profiler lines ("GPU Dist" / "GPU Sort" / "Rasterization") and an optional PPM of the frame.
// Build:  g++ -std=c++17 -Iinclude examples/render_host.cpp -Lvk_gaussian_splatting_b200/lib -lvkgs_b200
//             -Wl,-rpath,$PWD/vk_gaussian_splatting_b200/lib -o render_host
// Usage:  render_host [scene.ply|.spz|.splat | --synth N] [--size WxH] [--frames K] [--ftb] [--3dgut] [--fisheye]
//                     [--ppm out.ppm]
// Exit codes: 0 ok, 2 no device / render error (there is no CPU fallback), 3 bad arguments or unreadable scene.
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <vector>

#include "vkgs_b200.hpp"

int main(int argc, char** argv)
{
  std::string scenePath, ppmPath;
  uint64_t    synthCount = 100000;
  uint32_t    width = 1920, height = 1080;
  int         frames = 8;
  vkgs_b200::GaussianSplatting gs;
  for(int i = 1; i < argc; i++)
  {
    const std::string a = argv[i];
    if(a == "--synth" && i + 1 < argc)
      synthCount = std::strtoull(argv[++i], nullptr, 10);
    else if(a == "--size" && i + 1 < argc)
    {
      if(std::sscanf(argv[++i], "%ux%u", &width, &height) != 2 || !width || !height)
        return std::fprintf(stderr, "bad --size\n"), 3;
    }
    else if(a == "--frames" && i + 1 < argc)
      frames = std::atoi(argv[++i]);
    else if(a == "--ppm" && i + 1 < argc)
      ppmPath = argv[++i];
    else if(a == "--ftb")
      gs.prm.front_to_back = 1;
    else if(a == "--3dgut")
      gs.prm.pipeline = VKGS_PIPELINE_3DGUT;
    else if(a == "--fisheye")
      gs.prm.pipeline = VKGS_PIPELINE_3DGUT, gs.prm.camera_model = VKGS_CAMERA_FISHEYE;
    else if(a[0] != '-')
      scenePath = a;
    else
      return std::fprintf(stderr, "unknown argument %s\n", a.c_str()), 3;
  }

  vkgs_b200::SplatSet set;
  if(!scenePath.empty())
  {
    std::string err;
    if(!set.loadFromFile(scenePath, &err))
      return std::fprintf(stderr, "cannot load %s: %s\n", scenePath.c_str(), err.c_str()), 3;
  }
  else if(!set.synthesize(synthCount, 3, 0x3D650001ull))
    return std::fprintf(stderr, "cannot synthesize %llu splats\n", static_cast<unsigned long long>(synthCount)), 3;
  std::printf("%s: %zu splats, SH degree %u, library %s\n", scenePath.empty() ? "synthetic scene" : scenePath.c_str(), set.size(),
              set.maxShDegree(), vkgs_version());

  if(!gs.onAttach(0))
    return std::fprintf(stderr, "onAttach: %s\n", gs.lastError().c_str()), 2;
  gs.onResize(width, height);
  if(!gs.initDataStorage(set))
    return std::fprintf(stderr, "initDataStorage: %s\n", gs.lastError().c_str()), 2;

  vkgs_camera camera;
  vkgs_default_camera(&camera);  // the reference's default camera, src/camera_set.h:48-53
  std::vector<float> rgba(static_cast<size_t>(width) * height * 4);
  vkgs_set_profiling(gs.context(), 1);
  struct Timer
  {
    const char* name;
    double      sum = 0, mn = 1e30, mx = 0, last = 0;
    void        add(double us) { sum += us, mn = us < mn ? us : mn, mx = us > mx ? us : mx, last = us; }
  } timers[3] = {{"GPU Dist"}, {"GPU Sort"}, {"Rasterization"}};
  vkgs_outputs stats{};
  for(int f = 0; f < frames; f++)
  {
    if(!gs.updateAndUploadFrameInfoUBO(camera) || !gs.onRender(rgba.data(), &stats))
      return std::fprintf(stderr, "onRender: %s\n", gs.lastError().c_str()), 2;
    timers[0].add(1000.0 * stats.ms_dist), timers[1].add(1000.0 * stats.ms_sort), timers[2].add(1000.0 * stats.ms_raster);
  }
  // the timer lines of the reference's benchmark mode (nvpro_core2/nvutils/profiler.cpp:55; parsed by its benchmark.py:19-76)
  for(const Timer& t : timers)
    std::printf("Timeline \"Frame\"; level 1; Timer \"%s\"; GPU; avg %.0f; min %.0f; max %.0f; last %.0f; CPU; avg 0; min 0; max 0; last 0; samples %d;\n",
                t.name, t.sum / frames, t.mn, t.mx, t.last, frames);
  std::printf("Rasterized splats: %u of %zu, %llu (tile, splat) pairs\n", stats.visible_count, set.size(),
              static_cast<unsigned long long>(stats.tile_pairs));

  if(!ppmPath.empty())
  {
    FILE* fp = std::fopen(ppmPath.c_str(), "wb");
    if(!fp)
      return std::fprintf(stderr, "cannot write %s\n", ppmPath.c_str()), 3;
    std::fprintf(fp, "P6\n%u %u\n255\n", width, height);
    for(size_t p = 0; p < static_cast<size_t>(width) * height; p++)
      for(int c = 0; c < 3; c++)
      {
        const float v = rgba[4 * p + c];
        std::fputc(static_cast<int>((v < 0.f ? 0.f : v > 1.f ? 1.f : v) * 255.0f + 0.5f), fp);
      }
    std::fclose(fp);
    std::printf("wrote %s\n", ppmPath.c_str());
  }
  return 0;
}
