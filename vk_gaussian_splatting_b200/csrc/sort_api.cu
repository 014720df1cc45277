// sort_api.cu — stand-alone key/value radix sort entry point (vkgs_sort_pairs), the drop-in for
// vrdxCmdSortKeyValueIndirect (3rdparty/vrdx/src/vk_radix_sort.cc:249-258). Same kernels as the
// frame pipeline (k_radix_sort.cu); host in, host out, device time reported.
#include <algorithm>
#include <cstring>
#include <string>

#include "context.hpp"

using namespace vkgs;

extern "C" {

int vkgs_sort_pairs(vkgs_ctx* c, const uint32_t* keys, const uint32_t* values, uint64_t n, uint32_t* keysOut, uint32_t* valuesOut,
                    int repeats, float* msDevice)
{
  if(!c || (!keys && n) || (!values && n) || n > 0xfffff000ull)
    return VKGS_ERR_INVALID_ARGUMENT;
  if(msDevice)
    *msDevice = 0.0f;
  if(n == 0)
    return VKGS_OK;
  CU_TRY(c, cudaSetDevice(c->device));
  repeats = std::max(repeats, 1);
  uint32_t *dk[2] = {nullptr, nullptr}, *dv[2] = {nullptr, nullptr}, *dIn[2] = {nullptr, nullptr};
  struct Ctl
  {
    uint32_t count;
    uint32_t ticket[4];
    uint32_t hist[4][256];
  }* dCtl             = nullptr;
  uint64_t*      dSt  = nullptr;
  const uint64_t parts = (n + SORT_PART - 1) / SORT_PART;
  auto           cleanup = [&]() {
    for(int i = 0; i < 2; i++)
      freeDev(dk[i]), freeDev(dv[i]), freeDev(dIn[i]);
    freeDev(dCtl), freeDev(dSt);
  };
  cudaError_t e = cudaSuccess;
  for(int i = 0; i < 2 && e == cudaSuccess; i++)
  {
    e = cudaMalloc(&dk[i], n * 4);
    if(e == cudaSuccess)
      e = cudaMalloc(&dv[i], n * 4);
    if(e == cudaSuccess)
      e = cudaMalloc(&dIn[i], n * 4);
  }
  if(e == cudaSuccess)
    e = cudaMalloc(&dCtl, sizeof(Ctl));
  if(e == cudaSuccess)
    e = cudaMalloc(&dSt, parts * 256 * sizeof(uint64_t));
  if(e == cudaSuccess)
    e = cudaMemset(dSt, 0, parts * 256 * sizeof(uint64_t));
  if(e == cudaSuccess)
    e = cudaMemcpy(dIn[0], keys, n * 4, cudaMemcpyHostToDevice);
  if(e == cudaSuccess)
    e = cudaMemcpy(dIn[1], values, n * 4, cudaMemcpyHostToDevice);
  // The memset and the copies from pageable memory run on the legacy stream and may still be in flight when the calls
  // return; the sort runs on a non-blocking stream that is not ordered behind them (seen as corrupted VALUES — the last
  // upload — in about one sort of ten at 300 k pairs on one box).
  if(e == cudaSuccess)
    e = cudaStreamSynchronize(cudaStreamLegacy);
  if(e != cudaSuccess)
  {
    cleanup();
    c->lastError = std::string("vkgs_sort_pairs alloc/upload: ") + cudaGetErrorString(e);
    return VKGS_ERR_CUDA;
  }
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0);
  cudaEventCreate(&e1);
  float total = 0.0f;
  // private epoch space: this status array is local to the call
  uint32_t epoch = 0;
  for(int rep = 0; rep < repeats; rep++)
  {
    // restore the unsorted input (not timed), then time histogram + 4 passes
    cudaMemcpyAsync(dk[0], dIn[0], n * 4, cudaMemcpyDeviceToDevice, c->slots[0].stream);
    cudaMemcpyAsync(dv[0], dIn[1], n * 4, cudaMemcpyDeviceToDevice, c->slots[0].stream);
    Ctl h{};
    h.count = static_cast<uint32_t>(n);
    cudaMemcpyAsync(dCtl, &h, sizeof(Ctl), cudaMemcpyHostToDevice, c->slots[0].stream);
    cudaStreamSynchronize(c->slots[0].stream);
    cudaEventRecord(e0, c->slots[0].stream);
    launchHistogram(dk[0], &dCtl->count, static_cast<uint32_t>(n), &dCtl->hist[0][0], 0, 4, c->slots[0].stream);
    c->launches++;
    for(int p = 0; p < 4; p++)
    {
      SortPassArgs sa{};
      sa.keys[0] = dk[p & 1], sa.keys[1] = dk[(p + 1) & 1];
      sa.vals[0] = dv[p & 1], sa.vals[1] = dv[(p + 1) & 1];
      sa.countPtr  = &dCtl->count;
      sa.maxCount  = static_cast<uint32_t>(n);
      sa.histogram = &dCtl->hist[p][0];
      sa.status    = dSt;
      sa.ticket    = &dCtl->ticket[p];
      sa.epoch     = ++epoch;
      sa.shift     = 8 * p;
      launchSortPass(sa, c->slots[0].stream);
      c->launches++;
    }
    cudaEventRecord(e1, c->slots[0].stream);
    cudaEventSynchronize(e1);
    float ms = 0.0f;
    cudaEventElapsedTime(&ms, e0, e1);
    total += ms;
  }
  cudaEventDestroy(e0);
  cudaEventDestroy(e1);
  e = cudaGetLastError();
  if(e == cudaSuccess && keysOut)
    e = cudaMemcpy(keysOut, dk[0], n * 4, cudaMemcpyDeviceToHost);
  if(e == cudaSuccess && valuesOut)
    e = cudaMemcpy(valuesOut, dv[0], n * 4, cudaMemcpyDeviceToHost);
  cleanup();
  if(e != cudaSuccess)
  {
    c->lastError = std::string("vkgs_sort_pairs: ") + cudaGetErrorString(e);
    return VKGS_ERR_CUDA;
  }
  if(msDevice)
    *msDevice = total / static_cast<float>(repeats);
  return VKGS_OK;
}


// ---- device-buffer entry: the drop-in for vrdxCmdSortKeyValueIndirect (3rdparty/vrdx/src/vk_radix_sort.cc:249-258) ------------
// vrdx sorts a keys buffer and a values buffer in place, with the element count read from a device buffer ("indirect") and a
// caller-provided storage buffer sized by vrdxGetSorterKeyValueStorageRequirements (:209-224). Same contract here:
// stream-ordered (nothing is synchronised), stable, ascending, in place.
namespace {
struct DevSortCtl
{
  uint32_t ticket[4];
  uint32_t hist[4][256];
};
constexpr uint64_t alignUp(uint64_t v, uint64_t a)
{
  return (v + a - 1) / a * a;
}
}  // namespace

uint64_t vkgs_sort_pairs_storage_bytes(uint64_t max_count)
{
  const uint64_t parts = (max_count + SORT_PART - 1) / SORT_PART;
  return alignUp(sizeof(DevSortCtl), 256) + alignUp(parts * 256 * sizeof(uint64_t), 256) + 2 * alignUp(max_count * sizeof(uint32_t), 256);
}

int vkgs_sort_pairs_device(vkgs_ctx* c, uint32_t* dKeys, uint32_t* dValues, const uint32_t* dCount, uint64_t maxCount, void* dStorage,
                           uint64_t storageBytes, void* cudaStream)
{
  if(!c || !dKeys || !dValues || !dCount || !dStorage || maxCount > 0xfffff000ull)
    return VKGS_ERR_INVALID_ARGUMENT;
  if(storageBytes < vkgs_sort_pairs_storage_bytes(maxCount))
    return VKGS_ERR_INVALID_ARGUMENT;
  if(maxCount == 0)
    return VKGS_OK;
  CU_TRY(c, cudaSetDevice(c->device));
  cudaStream_t   st    = cudaStream ? static_cast<cudaStream_t>(cudaStream) : c->slots[0].stream;
  const uint64_t parts = (maxCount + SORT_PART - 1) / SORT_PART;
  unsigned char* base  = static_cast<unsigned char*>(dStorage);
  DevSortCtl*    ctl   = reinterpret_cast<DevSortCtl*>(base);
  uint64_t*      stat  = reinterpret_cast<uint64_t*>(base + alignUp(sizeof(DevSortCtl), 256));
  uint32_t*      tk    = reinterpret_cast<uint32_t*>(base + alignUp(sizeof(DevSortCtl), 256) + alignUp(parts * 256 * sizeof(uint64_t), 256));
  uint32_t*      tv    = reinterpret_cast<uint32_t*>(reinterpret_cast<unsigned char*>(tk) + alignUp(maxCount * sizeof(uint32_t), 256));
  // the storage is the caller's: whatever it holds must not be mistaken for look-back words of this call
  CU_TRY(c, cudaMemsetAsync(base, 0, alignUp(sizeof(DevSortCtl), 256) + parts * 256 * sizeof(uint64_t), st));
  launchHistogram(dKeys, dCount, static_cast<uint32_t>(maxCount), &ctl->hist[0][0], 0, 4, st);
  c->launches++;
  for(int p = 0; p < 4; p++)
  {
    SortPassArgs sa{};
    // four passes, fixed ping-pong (no pass is skipped): the result lands back in the caller's buffers
    sa.keys[0] = (p & 1) ? tk : dKeys, sa.keys[1] = (p & 1) ? dKeys : tk;
    sa.vals[0] = (p & 1) ? tv : dValues, sa.vals[1] = (p & 1) ? dValues : tv;
    sa.countPtr  = dCount;
    sa.maxCount  = static_cast<uint32_t>(maxCount);
    sa.histogram = &ctl->hist[p][0];
    sa.status    = stat;
    sa.ticket    = &ctl->ticket[p];
    sa.epoch     = static_cast<uint32_t>(p + 1);  // private, freshly cleared status array
    sa.shift     = 8 * p;
    launchSortPass(sa, st);
    c->launches++;
  }
  CU_TRY(c, cudaGetLastError());
  return VKGS_OK;
}

}  // extern "C"
