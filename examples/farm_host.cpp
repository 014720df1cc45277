// C++ multi-view farm over the C ABI: the one way this path shards — frames of different cameras are
// independent, so every GPU holds the whole splat set and renders its own view; there is no data-path collective.
// One host thread and one context per GPU (a context is single-caller, like the reference's render thread); the views
// are the eight orbit cameras of bench.py (default eye rotated about +Y). After the timed run GPU 0 renders every view
// itself and the frames are compared bit for bit with what the other GPUs produced.
// (bench.py runs the same farm as one PROCESS per GPU under torch.distributed; this is the host-language counterpart.)
// Build:  g++ -std=c++17 -pthread -Iinclude examples/farm_host.cpp -Lvk_gaussian_splatting_b200/lib -lvkgs_b200
//             -Wl,-rpath,$PWD/vk_gaussian_splatting_b200/lib -o farm_host
// Usage:  farm_host [--gpus G] [--synth N] [--size WxH] [--frames K]
// Exit codes: 0 ok, 2 no device / render error (there is no CPU fallback), 3 bad arguments, 4 cross-GPU frame mismatch.
#include <atomic>
#include <chrono>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <thread>
#include <vector>

#include "vkgs_b200.hpp"

static vkgs_camera orbitCamera(int view, int views)
{
  vkgs_camera cam;
  vkgs_default_camera(&cam);
  const double ang = 2.0 * 3.14159265358979323846 * view / views;
  const double x = cam.eye[0], z = cam.eye[2];
  cam.eye[0] = static_cast<float>(std::cos(ang) * x + std::sin(ang) * z);
  cam.eye[2] = static_cast<float>(-std::sin(ang) * x + std::cos(ang) * z);
  return cam;
}

int main(int argc, char** argv)
{
  int      gpus = 1, frames = 200;
  uint64_t synthCount = 1000000;
  uint32_t width = 1920, height = 1080;
  for(int i = 1; i < argc; i++)
  {
    const std::string a = argv[i];
    if(a == "--gpus" && i + 1 < argc)
      gpus = std::atoi(argv[++i]);
    else if(a == "--synth" && i + 1 < argc)
      synthCount = std::strtoull(argv[++i], nullptr, 10);
    else if(a == "--frames" && i + 1 < argc)
      frames = std::atoi(argv[++i]);
    else if(a == "--size" && i + 1 < argc)
    {
      if(std::sscanf(argv[++i], "%ux%u", &width, &height) != 2 || !width || !height)
        return std::fprintf(stderr, "bad --size\n"), 3;
    }
    else
      return std::fprintf(stderr, "unknown argument %s\n", a.c_str()), 3;
  }
  if(gpus < 1 || gpus > 8 || frames < 1)
    return std::fprintf(stderr, "--gpus 1..8, --frames >= 1\n"), 3;

  vkgs_b200::SplatSet set;
  if(!set.synthesize(synthCount, 3, 0x3D650001ull))
    return std::fprintf(stderr, "cannot synthesize the scene\n"), 3;

  const size_t                       pixels = static_cast<size_t>(width) * height;
  std::vector<std::vector<float>>    frame(gpus, std::vector<float>(pixels * 4));
  std::vector<double>                seconds(gpus, 0.0);
  std::vector<std::string>           error(gpus);
  std::atomic<int>                   ready{0};
  std::atomic<bool>                  go{false};
  std::vector<std::thread>           pool;
  for(int g = 0; g < gpus; g++)
    pool.emplace_back([&, g]() {
      vkgs_b200::GaussianSplatting gs;
      gs.prm.front_to_back         = 1;
      gs.prm.transmittance_epsilon = 1.0f / 32768.0f;
      gs.onResize(width, height);
      bool ok = gs.onAttach(g) && gs.initDataStorage(set) && gs.updateAndUploadFrameInfoUBO(orbitCamera(g, 8));
      for(int f = 0; ok && f < 8; f++)  // warm-up
        ok = gs.onRenderAsync();
      ok = ok && gs.sync();
      ready++;
      while(!go.load())  // all GPUs start their timed run together
        std::this_thread::yield();
      const auto t0 = std::chrono::steady_clock::now();
      for(int f = 0; ok && f < frames; f++)
        ok = gs.onRenderAsync();  // four frames in flight, the frame stays on the device
      ok         = ok && gs.sync();
      seconds[g] = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
      ok         = ok && gs.onRender(frame[g].data());  // the view once more, to the host, for the cross-GPU check
      if(!ok)
        error[g] = gs.lastError().empty() ? "render failed" : gs.lastError();
    });
  while(ready.load() < gpus)
    std::this_thread::yield();
  go = true;
  for(auto& t : pool)
    t.join();
  double slowest = 0.0;
  for(int g = 0; g < gpus; g++)
  {
    if(!error[g].empty())
      return std::fprintf(stderr, "GPU %d: %s\n", g, error[g].c_str()), 2;
    slowest = seconds[g] > slowest ? seconds[g] : slowest;
    std::printf("GPU %d: view %d, %d frames in %.3f ms = %.1f frames/s\n", g, g, frames, 1e3 * seconds[g], frames / seconds[g]);
  }
  std::printf("farm: %d GPUs x %d frames over the slowest GPU's %.3f ms = %.1f frames/s\n", gpus, frames, 1e3 * slowest, gpus * frames / slowest);

  // cross-GPU check: every view rendered again on GPU 0 must equal the frame its own GPU produced, bit for bit
  vkgs_b200::GaussianSplatting gs0;
  gs0.prm.front_to_back         = 1;
  gs0.prm.transmittance_epsilon = 1.0f / 32768.0f;
  gs0.onResize(width, height);
  if(!gs0.onAttach(0) || !gs0.initDataStorage(set))
    return std::fprintf(stderr, "GPU 0: %s\n", gs0.lastError().c_str()), 2;
  std::vector<float> again(pixels * 4);
  for(int g = 0; g < gpus; g++)
  {
    if(!gs0.updateAndUploadFrameInfoUBO(orbitCamera(g, 8)) || !gs0.onRender(again.data()))
      return std::fprintf(stderr, "GPU 0: %s\n", gs0.lastError().c_str()), 2;
    if(std::memcmp(again.data(), frame[g].data(), again.size() * sizeof(float)) != 0)
      return std::fprintf(stderr, "view %d rendered on GPU %d differs from GPU 0\n", g, g), 4;
  }
  std::printf("cross-GPU check: %d views bit-identical to GPU 0\n", gpus);
  return 0;
}
