// cpu_sorter.cpp — TEST INFRASTRUCTURE / CPU BASELINE. Not product code.
//
// Behaviour-exact restatement of the reference's CPU sorting path:
//   SplatSorterAsync::innerSort           (src/splat_sorter_async.cpp:92-141)
//   START_PAR_LOOP / END_PAR_LOOP         (src/utilities.h:52-59)
//   nvutils::parallel_batches_pooled<8192> (nvpro_core2/nvutils/parallel_work.hpp:215-268)
// The reference file itself cannot be compiled here (it pulls nvvk/profiler_vk.hpp -> vulkan_core.h).
//
// "CPU Dist": dist[i] = |plane . (M p_i)| * divider, idx[i] = i, in batches of 8192 splats spread
//             over `threads` host threads (serial when n <= 8192, like the reference).
// "CPU Sort": std::sort(std::execution::par_unseq, idx, cmp) with cmp = dist[i] > dist[j]
//             (back-to-front) or < (front-to-back). libstdc++ without TBB runs par_unseq on its
//             serial backend, so mode 0 == std::sort on one core; mode 1 = __gnu_parallel::sort
//             (OpenMP) is reported as the honest all-cores figure.
#include <algorithm>
#include <chrono>
#include <cmath>
#include <cstdint>
#include <thread>
#include <vector>
#include <parallel/algorithm>
#include <omp.h>

namespace {
constexpr uint64_t BATCHSIZE = 8192;

inline double now_ms()
{
  using namespace std::chrono;
  return duration<double, std::milli>(steady_clock::now().time_since_epoch()).count();
}
}  // namespace

extern "C" __attribute__((visibility("default"))) int orc_cpu_sort(const float* positions, uint64_t n, const float model[16],
                                                                   const float dir[3], const float cop[3], int front_to_back,
                                                                   int mode, int threads, uint32_t* indices, float* distances,
                                                                   double* ms_dist, double* ms_sort)
{
  if(!positions || !indices || !distances || threads < 1)
    return -1;
  const double t0 = now_ms();
  // splat_sorter_async.cpp:103-106
  const float plane[4] = {dir[0], dir[1], dir[2], -dir[0] * cop[0] - dir[1] * cop[1] - dir[2] * cop[2]};
  const float divider  = 1.0f / std::sqrt(plane[0] * plane[0] + plane[1] * plane[1] + plane[2] * plane[2]);

  auto item = [&](uint64_t i) {
    // glm mat4 * vec4: (m[0]*x + m[1]*y) + (m[2]*z + m[3]*w)   (glm/detail/type_mat4x4.inl)
    const float x = positions[i * 3], y = positions[i * 3 + 1], z = positions[i * 3 + 2], w = 1.0f;
    float       p[4];
    for(int r = 0; r < 4; r++)
      p[r] = (model[0 + r] * x + model[4 + r] * y) + (model[8 + r] * z + model[12 + r] * w);
    const float dist = std::abs(plane[0] * p[0] + plane[1] * p[1] + plane[2] * p[2] + plane[3]) * divider;
    distances[i]     = dist;
    indices[i]       = (uint32_t)i;
  };

  if(n <= BATCHSIZE || threads == 1)
  {
    for(uint64_t i = 0; i < n; i++)
      item(i);
  }
  else
  {
    const uint64_t numBatches = (n + BATCHSIZE - 1) / BATCHSIZE;
    // BS::thread_pool::submit_loop with num_blocks = 0 -> one contiguous block of batches per thread
    const uint64_t           nt = std::min<uint64_t>((uint64_t)threads, numBatches);
    std::vector<std::thread> pool;
    for(uint64_t t = 0; t < nt; t++)
    {
      const uint64_t b0 = numBatches * t / nt, b1 = numBatches * (t + 1) / nt;
      pool.emplace_back([&, b0, b1]() {
        for(uint64_t b = b0; b < b1; b++)
        {
          const uint64_t start = BATCHSIZE * b;
          const uint64_t end   = std::min(n, start + BATCHSIZE);
          for(uint64_t i = start; i < end; i++)
            item(i);
        }
      });
    }
    for(auto& th : pool)
      th.join();
  }
  const double t1 = now_ms();

  // splat_sorter_async.cpp:132-136
  auto compare = [&](uint32_t i, uint32_t j) {
    return front_to_back ? (distances[i] < distances[j]) : (distances[i] > distances[j]);
  };
  if(mode == 0)
  {
    std::sort(indices, indices + n, compare);
  }
  else
  {
    omp_set_num_threads(threads);
    __gnu_parallel::sort(indices, indices + n, compare);
  }
  const double t2 = now_ms();
  if(ms_dist)
    *ms_dist = t1 - t0;
  if(ms_sort)
    *ms_sort = t2 - t1;
  return 0;
}

extern "C" __attribute__((visibility("default"))) int orc_hardware_concurrency()
{
  return (int)std::thread::hardware_concurrency();
}
