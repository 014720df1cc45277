// tools/kbench.cu — stand-alone micro-benchmark / timeline harness for individual kernels.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -DVKGS_TIMELINE -Iinclude
//        -Ivk_gaussian_splatting_b200/csrc tools/kbench.cu -o gpurun_out/kbench
#include <cstdio>
#include <cstdlib>
#include <vector>
#include <algorithm>
#include <cstring>
#include "device_common.cuh"
#include "k_radix_sort.cu"
#include "k_binning.cu"

using namespace vkgs;

#define CK(x) do { cudaError_t e = (x); if(e != cudaSuccess) { printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e), __FILE__, __LINE__); exit(1);} } while(0)

struct Ctl { uint32_t count; uint32_t ticket[8]; uint32_t hist[4][256]; };

static void benchBin(uint32_t n)
{
  // synthetic records: bbox covering ~2x2 tiles at random screen positions (1920x1080), ids = random permutation
  std::vector<uint32_t> rec((size_t)n * RECORD_WORDS, 0), ids(n);
  uint64_t st = 0x9E3779B97F4A7C15ull;
  auto rnd = [&]() { st ^= st << 13; st ^= st >> 7; st ^= st << 17; return (uint32_t)(st >> 11); };
  for(uint32_t i = 0; i < n; i++)
  {
    uint32_t x0 = rnd() % 1900, y0 = rnd() % 1060, w = 4 + rnd() % 30, h = 4 + rnd() % 30;
    uint32_t x1 = std::min(1919u, x0 + w), y1 = std::min(1079u, y0 + h);
    rec[(size_t)i * RECORD_WORDS + 10] = x0 | (y0 << 16);
    rec[(size_t)i * RECORD_WORDS + 11] = x1 | (y1 << 16);
    ids[i] = i;
  }
  for(uint32_t i = n - 1; i > 0; i--) std::swap(ids[i], ids[rnd() % (i + 1)]);
  uint32_t *drec, *dids, *dk, *dv; FrameCounters* dc; uint64_t* dst; long long* dtl;
  const uint32_t cap = 8 * n, parts = (n + 1023) / 1024;
  CK(cudaMalloc(&drec, rec.size() * 4)); CK(cudaMalloc(&dids, n * 4)); CK(cudaMalloc(&dk, (size_t)cap * 4)); CK(cudaMalloc(&dv, (size_t)cap * 4));
  CK(cudaMalloc(&dc, sizeof(FrameCounters))); CK(cudaMalloc(&dst, parts * 8)); CK(cudaMemset(dst, 0, parts * 8));
  CK(cudaMalloc(&dtl, (size_t)parts * 16 * 8)); CK(cudaMemset(dtl, 0, (size_t)parts * 16 * 8));
  CK(cudaMemcpy(drec, rec.data(), rec.size() * 4, cudaMemcpyHostToDevice));
  std::vector<uint2> hbb(n); for(uint32_t i = 0; i < n; i++) hbb[i] = make_uint2(rec[(size_t)i * RECORD_WORDS + 10], rec[(size_t)i * RECORD_WORDS + 11]);
  uint2* dbb; CK(cudaMalloc(&dbb, (size_t)n * 8)); CK(cudaMemcpy(dbb, hbb.data(), (size_t)n * 8, cudaMemcpyHostToDevice)); CK(cudaMemcpy(dids, ids.data(), n * 4, cudaMemcpyHostToDevice));
  CK(cudaMemcpyToSymbol(g_vkgsTimeline, &dtl, sizeof(dtl)));
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  float best = 1e9f; FrameCounters hc{};
  for(int rep = 0; rep < 10; rep++)
  {
    FrameCounters h{}; h.visible = n; CK(cudaMemcpy(dc, &h, sizeof(h), cudaMemcpyHostToDevice));
    BinArgs ba{}; ba.sortedIds[0] = dids; ba.sortedIds[1] = dids; ba.sortedSel = nullptr; ba.bboxes = dbb; ba.counters = dc; ba.tileKeys = dk; ba.tileVals = dv; ba.capacity = cap; ba.maxCount = n;
    ba.tilesX = 120; ba.tilesY = 68; ba.status = dst; ba.epoch = rep + 1; ba.ticketSlot = 5; ba.debugFlags = 0;
    CK(cudaDeviceSynchronize());
    cudaEventRecord(e0); launchBinEmit(ba, 0); cudaEventRecord(e1); CK(cudaDeviceSynchronize());
    float ms; cudaEventElapsedTime(&ms, e0, e1); best = std::min(best, ms);
    CK(cudaMemcpy(&hc, dc, sizeof(hc), cudaMemcpyDeviceToHost));
  }
  printf("bin_emit n=%u pairs=%u : %.1f us\n", n, hc.tilePairs, best * 1000);
  std::vector<long long> tl((size_t)parts * 16);
  CK(cudaMemcpy(tl.data(), dtl, tl.size() * 8, cudaMemcpyDeviceToHost));
  const char* names[] = {"start", "ids+gather", "scan", "lookback", "emit"};
  for(uint32_t p : {0u, parts / 4, parts / 2, parts - 1})
  {
    printf("part %5u:", p);
    for(int k = 1; k < 5; k++) printf(" %s+%lld", names[k], tl[p * 16 + k] - tl[p * 16 + k - 1]);
    printf("\n");
  }
}

int main(int argc, char** argv)
{
  if(argc > 1 && !strcmp(argv[1], "bin")) { benchBin(argc > 2 ? atoi(argv[2]) : 1000000); return 0; }
  const uint32_t n = argc > 1 ? atoi(argv[1]) : 1000000;
  const int keyBits = argc > 2 ? atoi(argv[2]) : 32;
  std::vector<uint32_t> hk(n), hv(n);
  uint64_t st = 88172645463325252ull;
  for(uint32_t i = 0; i < n; i++) { st ^= st << 13; st ^= st >> 7; st ^= st << 17; hk[i] = keyBits >= 32 ? (uint32_t)st : (uint32_t)(st & ((1ull << keyBits) - 1)); hv[i] = i; }
  uint32_t *dk[2], *dv[2], *din;
  for(int i = 0; i < 2; i++) { CK(cudaMalloc(&dk[i], n * 4)); CK(cudaMalloc(&dv[i], n * 4)); }
  CK(cudaMalloc(&din, n * 4));
  CK(cudaMemcpy(din, hk.data(), n * 4, cudaMemcpyHostToDevice));
  CK(cudaMemcpy(dv[0], hv.data(), n * 4, cudaMemcpyHostToDevice));
  Ctl* dctl; CK(cudaMalloc(&dctl, sizeof(Ctl)));
  const uint32_t parts = (n + SORT_PART - 1) / SORT_PART;
  uint64_t* dst; CK(cudaMalloc(&dst, (size_t)parts * 256 * 8)); CK(cudaMemset(dst, 0, (size_t)parts * 256 * 8));
  long long* dtl; CK(cudaMalloc(&dtl, (size_t)parts * 16 * 8)); CK(cudaMemset(dtl, 0, (size_t)parts * 16 * 8));
  CK(cudaMemcpyToSymbol(g_vkgsTimeline, &dtl, sizeof(dtl)));
  initSortKernels();
  cudaEvent_t ev[8]; for(auto& e : ev) cudaEventCreate(&e);
  uint32_t epoch = 0;
  const int passes = (keyBits + 7) / 8;
  float best[8]; for(auto& b : best) b = 1e9f;
  for(int rep = 0; rep < 12; rep++)
  {
    CK(cudaMemcpy(dk[0], din, n * 4, cudaMemcpyDeviceToDevice));
    Ctl h{}; h.count = n; CK(cudaMemcpy(dctl, &h, sizeof(h), cudaMemcpyHostToDevice));
    CK(cudaDeviceSynchronize());
    cudaEventRecord(ev[0]);
    launchHistogram(dk[0], &dctl->count, n, &dctl->hist[0][0], 0, 4, 0);
    cudaEventRecord(ev[1]);
    for(int p = 0; p < passes; p++)
    {
      SortPassArgs sa{}; sa.keys[0] = dk[p & 1]; sa.vals[0] = dv[p & 1]; sa.keys[1] = dk[(p + 1) & 1]; sa.vals[1] = dv[(p + 1) & 1];
      sa.countPtr = &dctl->count; sa.maxCount = n; sa.histogram = &dctl->hist[p][0]; sa.status = dst; sa.ticket = &dctl->ticket[p];
      sa.epoch = ++epoch; sa.shift = 8 * p;
      launchSortPass(sa, 0);
      cudaEventRecord(ev[2 + p]);
    }
    CK(cudaDeviceSynchronize());
    for(int k = 0; k <= passes; k++) { float ms; cudaEventElapsedTime(&ms, ev[k], ev[k + 1]); best[k] = std::min(best[k], ms); }
  }
  printf("n=%u bits=%d  hist %.1f us |", n, keyBits, best[0] * 1000);
  for(int p = 0; p < passes; p++) printf(" pass%d %.1f us", p, best[1 + p] * 1000);
  printf("\n");
  // verify
  std::vector<uint32_t> ok(n), ov(n);
  CK(cudaMemcpy(ok.data(), dk[passes & 1], n * 4, cudaMemcpyDeviceToHost));
  std::vector<uint32_t> ref = hk; std::stable_sort(ref.begin(), ref.end());
  printf("sorted correctly: %s\n", ok == ref ? "yes" : "NO");
  // timeline of the last pass
  std::vector<long long> tl((size_t)parts * 16);
  CK(cudaMemcpy(tl.data(), dtl, tl.size() * 8, cudaMemcpyDeviceToHost));
  long long t0 = tl[0]; for(uint32_t p = 0; p < parts; p++) t0 = std::min(t0, tl[p * 16]);
  const char* names[] = {"start", "ticket", "keys-issued", "ranked", "scans", "staged", "lookback", "sync", "scattered"};
  for(uint32_t p : {0u, 1u, parts / 4, parts / 2, parts - 2, parts - 1})
  {
    if(p >= parts) continue;
    printf("part %5u:", p);
    for(int k = 1; k < 9; k++) printf(" %s+%lld", names[k], tl[p * 16 + k] - tl[p * 16 + k - 1]);
    printf("\n");
  }
  long long tend = 0; for(uint32_t p = 0; p < parts; p++) tend = std::max(tend, tl[p * 16 + 8]);
  printf("last pass span: %lld cycles\n", tend - t0);
  return 0;
}
