// k_metrics.cu — image quality metrics between a captured frame and the current frame:
// MSE / PSNR and the reference's fast FLIP approximation.
//
// Replaces shaders/image_compare_metric.comp.slang:84-190 (main), :371-477 (computeFLIPApprox), the
// colour pipeline shaders/color.h.slang:44-137, the dispatch in src/image_compare.cpp:770-830 and
// the read-back arithmetic in src/image_compare.cpp:874-905. Same accumulation contract as the
// reference: every pixel's contribution is normalised, scaled by 1e9, truncated to uint32 and added
// to an integer accumulator — integer adds commute, so the fixed-point sums are reproducible bit for
// bit whatever the reduction order (here: per-thread -> warp REDUX -> one global atomic per CTA).
// The arithmetic of the MSE term is evaluated in the oracle's operation order (this file is compiled
// with -fmad=false), so mse_fixed is bit-exact against oracle/vkgs_oracle.c.
// HBM-bound: 32 B/pixel read once (the 3x3 Sobel taps of the FLIP term hit L1/L2).
#include <cmath>
#include <cstring>
#include <string>

#include "context.hpp"
#include "device_common.cuh"

using namespace vkgs;

namespace {

struct MetricArgs
{
  const float4* reference;  // [H][W] "capture image"
  const float4* current;    // [H][W]
  uint32_t*     result;     // [4]: mse_fixed, reserved, flip_fixed, reserved (the reference's 16-byte result buffer)
  int           width, height;
  float         sampleDivider;  // W * H * 3
  uint32_t      flipMode;
  float         pixelsPerDegree;  // 67 in the reference (src/image_compare.cpp:788)
};

__device__ __forceinline__ float srgbToLinear1(float c)
{
  return (c <= 0.04045f) ? (c / 12.92f) : powf((c + 0.055f) / 1.055f, 2.4f);
}

// sRGB -> linear -> LMS (Hunt-Pointer-Estevez) -> Hunt adaptation (La = 1) -> YCxCz, color.h.slang:83-143
__device__ __forceinline__ void srgbToFlipSpace(float r, float g, float b, float out[3])
{
  const float lr = srgbToLinear1(r), lg = srgbToLinear1(g), lb = srgbToLinear1(b);
  const float L = (0.31670331f * lr + 0.70299344f * lg) + -0.01969366f * lb;
  const float M = (0.10938715f * lr + 0.87060437f * lg) + 0.01990658f * lb;
  const float S = (0.01840087f * lr + 0.10476914f * lg) + 0.87470614f * lb;
  const float k     = 5.0f * 1.0f;
  const float kcbrt = powf(k, 1.0f / 3.0f);
  const float FL    = 0.2f * kcbrt * (1.0f - expf(-0.42f * kcbrt));
  const float hl = L * FL, hm = M * FL, hs = S * FL;
  out[0] = hm;
  out[1] = hl - hm;
  out[2] = hm - hs;
}

__device__ __forceinline__ float csfLuminance(float cpd)
{
  const float s = 1.0f / sqrtf(1.0f + powf(cpd / 4.0f, 2.0f));
  return s * expf(-0.5f * cpd);
}

__device__ __forceinline__ float lum(const float4 c)
{
  return (c.x * 0.2126f + c.y * 0.7152f) + c.z * 0.0722f;
}

__device__ float sobelMagnitude(const float4* img, int x, int y, int w, int h)
{
  if(!(x > 0 && y > 0 && x < w - 1 && y < h - 1))
    return 0.0f;
  const float tl = lum(img[(y - 1) * w + x - 1]), tc = lum(img[(y - 1) * w + x]), tr = lum(img[(y - 1) * w + x + 1]);
  const float ml = lum(img[y * w + x - 1]), mr = lum(img[y * w + x + 1]);
  const float bl = lum(img[(y + 1) * w + x - 1]), bc = lum(img[(y + 1) * w + x]), br = lum(img[(y + 1) * w + x + 1]);
  const float gx = ((((-tl + tr) - 2.0f * ml) + 2.0f * mr) - bl) + br;
  const float gy = ((((-tl - 2.0f * tc) - tr) + bl) + 2.0f * bc) + br;
  return sqrtf(gx * gx + gy * gy);
}

// FLIP "reference" mode (image_compare_metric.comp.slang:231-300, 486-543): five band-pass features per
// image, |centre luminance - Gaussian-blurred luminance| * CSF at 0.5/1/2/4/8 cycles per degree, the blur a
// brute-force (2r+1)^2 window with r = ceil(3 sigma), sigma = max(ppd / (6.28 f), 0.5) — r = 65 px for the lowest
// band at the reference's 67 pixels per degree; pixels closer than r to the border keep their own luminance.
// The 1-D weights exp(-d^2 / (2 sigma^2)) are tabulated once per CTA (the shader evaluates the same two
// factors per tap); taps are summed in the shader's loop order (dx outer, dy inner).
constexpr int FLIP_BANDS      = 5;
constexpr int FLIP_MAX_RADIUS = 96;

struct FlipTables
{
  float w[FLIP_BANDS][2 * FLIP_MAX_RADIUS + 1];
  int   radius[FLIP_BANDS];
  float csf[FLIP_BANDS];
};

__device__ float flipBlurredLuminance(const float4* img, int x, int y, int w, int h, const float* wt, int radius)
{
  if(x < radius || y < radius || x >= w - radius || y >= h - radius)
    return lum(img[y * w + x]);
  float sum = 0.0f, weightSum = 0.0f;
  for(int dx = -radius; dx <= radius; dx++)
  {
    const float wx = wt[dx + radius];
    for(int dy = -radius; dy <= radius; dy++)
    {
      const float weight = wx * wt[dy + radius];
      sum += lum(img[(y + dy) * w + (x + dx)]) * weight;
      weightSum += weight;
    }
  }
  return sum / weightSum;
}

__global__ void __launch_bounds__(256) k_image_metrics(const __grid_constant__ MetricArgs a)
{
  __shared__ uint32_t   s_sum[2];
  __shared__ FlipTables s_flip;
  if(threadIdx.x < 2)
    s_sum[threadIdx.x] = 0u;
  if(a.flipMode == VKGS_FLIP_REFERENCE)
  {
    const float freq[FLIP_BANDS] = {0.5f, 1.0f, 2.0f, 4.0f, 8.0f};
    for(int b = 0; b < FLIP_BANDS; b++)
    {
      const float sigma  = fmaxf(a.pixelsPerDegree / (freq[b] * 6.28f), 0.5f);
      const int   radius = min(static_cast<int>(ceilf(3.0f * sigma)), FLIP_MAX_RADIUS);
      if(threadIdx.x == 0)
      {
        s_flip.radius[b] = radius;
        s_flip.csf[b]    = csfLuminance(freq[b]);
      }
      for(int i = threadIdx.x; i <= 2 * radius; i += blockDim.x)
      {
        const float d  = static_cast<float>(i - radius);
        s_flip.w[b][i] = expf(-(d * d) / (2.0f * sigma * sigma));
      }
    }
  }
  __syncthreads();
  // 16x16 pixels per CTA like the reference's numthreads(16,16,1); a warp covers 16x2 pixels
  const int x = blockIdx.x * 16 + (threadIdx.x & 15), y = blockIdx.y * 16 + (threadIdx.x >> 4);
  uint32_t  mseFixed = 0u, flipFixed = 0u;
  if(x < a.width && y < a.height)
  {
    const float4 ref = a.reference[y * a.width + x], cur = a.current[y * a.width + x];
    const float  dx = ref.x - cur.x, dy = ref.y - cur.y, dz = ref.z - cur.z;
    const float  squaredError = (dx * dx + dy * dy) + dz * dz;
    mseFixed                  = static_cast<uint32_t>((squaredError / a.sampleDivider) * 1000000000.0f);
    if(a.flipMode == VKGS_FLIP_REFERENCE)
    {
      float rc[3], cc[3];
      srgbToFlipSpace(ref.x, ref.y, ref.z, rc);
      srgbToFlipSpace(cur.x, cur.y, cur.z, cc);
      const float csfY = csfLuminance(1.0f), csfC = csfY * 0.4f;
      const float colorError = (fabsf(rc[0] - cc[0]) * csfY + fabsf(rc[1] - cc[1]) * csfC) + fabsf(rc[2] - cc[2]) * csfC;
      const float refLum = lum(ref), curLum = lum(cur);
      float       featureError = 0.0f;
      for(int b = 0; b < FLIP_BANDS; b++)
      {
        const float fr = fabsf(refLum - flipBlurredLuminance(a.reference, x, y, a.width, a.height, s_flip.w[b], s_flip.radius[b])) * s_flip.csf[b];
        const float fc = fabsf(curLum - flipBlurredLuminance(a.current, x, y, a.width, a.height, s_flip.w[b], s_flip.radius[b])) * s_flip.csf[b];
        featureError += fabsf(fr - fc);
      }
      const float powered    = powf(__saturatef(colorError + featureError), 3.0f);
      const float pixelCount = a.sampleDivider / 3.0f;
      flipFixed              = static_cast<uint32_t>((powered / pixelCount) * 1000000000.0f);
    }
    else if(a.flipMode == VKGS_FLIP_APPROX)
    {
      float rc[3], cc[3];
      srgbToFlipSpace(ref.x, ref.y, ref.z, rc);
      srgbToFlipSpace(cur.x, cur.y, cur.z, cc);
      const float csfY = csfLuminance(1.0f), csfC = csfY * 0.4f;
      const float colorError = (fabsf(rc[0] - cc[0]) * csfY + fabsf(rc[1] - cc[1]) * csfC) + fabsf(rc[2] - cc[2]) * csfC;
      const float refF = sobelMagnitude(a.reference, x, y, a.width, a.height);
      const float curF = sobelMagnitude(a.current, x, y, a.width, a.height);
      const float featureError = fabsf(refF - curF) * csfLuminance(4.0f);
      const float total   = colorError + featureError * 3.83f;
      const float powered = powf(__saturatef(total), 3.0f);
      const float pixelCount = a.sampleDivider / 3.0f;
      flipFixed              = static_cast<uint32_t>((powered / pixelCount) * 1000000000.0f);
    }
  }
  mseFixed  = __reduce_add_sync(FULL_MASK, mseFixed);
  flipFixed = __reduce_add_sync(FULL_MASK, flipFixed);
  if((threadIdx.x & 31u) == 0u)
  {
    atomicAdd(&s_sum[0], mseFixed);
    atomicAdd(&s_sum[1], flipFixed);
  }
  __syncthreads();
  if(threadIdx.x == 0 && s_sum[0])
    atomicAdd(a.result + 0, s_sum[0]);
  if(threadIdx.x == 1 && s_sum[1])
    atomicAdd(a.result + 2, s_sum[1]);
}

// src/image_compare.cpp:874-905
void finishMetrics(const uint32_t fixed[4], vkgs_image_metrics* out)
{
  out->mse_fixed  = fixed[0];
  out->flip_fixed = fixed[2];
  out->mse        = static_cast<float>(fixed[0]) / 1000000000.0f;
  out->psnr       = out->mse < 1e-10f ? 99.99f : std::min(10.0f * std::log10(1.0f / out->mse), 99.99f);
  out->flip       = static_cast<float>(std::pow(static_cast<double>(fixed[2]) / 1000000000.0, 1.0 / 3.0));
}

int runMetrics(vkgs_ctx* c, const float4* dRef, const float4* dCur, uint32_t w, uint32_t h, uint32_t flipMode, vkgs_image_metrics* out)
{
  uint32_t* dResult = nullptr;
  CU_TRY(c, cudaMalloc(&dResult, 16));
  cudaStream_t st = c->slots[0].stream;
  cudaEvent_t  e0, e1;
  cudaEventCreate(&e0);
  cudaEventCreate(&e1);
  MetricArgs a{dRef, dCur, dResult, static_cast<int>(w), static_cast<int>(h), static_cast<float>(w * h * 3), flipMode, 67.0f};
  cudaMemsetAsync(dResult, 0, 16, st);
  cudaEventRecord(e0, st);
  k_image_metrics<<<dim3((w + 15) / 16, (h + 15) / 16), 256, 0, st>>>(a);
  cudaEventRecord(e1, st);
  c->launches++;
  uint32_t    fixed[4] = {0, 0, 0, 0};
  cudaError_t e        = cudaMemcpyAsync(fixed, dResult, 16, cudaMemcpyDeviceToHost, st);
  if(e == cudaSuccess)
    e = cudaStreamSynchronize(st);
  float ms = 0.0f;
  cudaEventElapsedTime(&ms, e0, e1);
  cudaEventDestroy(e0);
  cudaEventDestroy(e1);
  cudaFree(dResult);
  if(e != cudaSuccess)
  {
    c->lastError = std::string("vkgs image metrics: ") + cudaGetErrorString(e);
    return VKGS_ERR_CUDA;
  }
  finishMetrics(fixed, out);
  out->ms_device = ms;
  return VKGS_OK;
}

}  // namespace

extern "C" {

int vkgs_image_metrics_host(vkgs_ctx* c, const float* reference, const float* current, uint32_t width, uint32_t height,
                            uint32_t flip_mode, vkgs_image_metrics* out)
{
  if(!c || !reference || !current || !out || width == 0 || height == 0 || flip_mode > VKGS_FLIP_REFERENCE
     || static_cast<uint64_t>(width) * height > (1ull << 28))
    return VKGS_ERR_INVALID_ARGUMENT;
  CU_TRY(c, cudaSetDevice(c->device));
  const size_t bytes = sizeof(float4) * static_cast<size_t>(width) * height;
  float4 *     dRef = nullptr, *dCur = nullptr;
  cudaError_t  e = cudaMalloc(&dRef, bytes);
  if(e == cudaSuccess)
    e = cudaMalloc(&dCur, bytes);
  if(e == cudaSuccess)
    e = cudaMemcpy(dRef, reference, bytes, cudaMemcpyHostToDevice);
  if(e == cudaSuccess)
    e = cudaMemcpy(dCur, current, bytes, cudaMemcpyHostToDevice);
  if(e == cudaSuccess)
    e = cudaStreamSynchronize(cudaStreamLegacy);  // pageable uploads may still be in flight; the metric kernels run on a non-blocking stream
  int rc = VKGS_OK;
  if(e != cudaSuccess)
  {
    c->lastError = std::string("vkgs_image_metrics_host: ") + cudaGetErrorString(e);
    rc           = VKGS_ERR_CUDA;
  }
  else
    rc = runMetrics(c, dRef, dCur, width, height, flip_mode, out);
  cudaFree(dRef);
  cudaFree(dCur);
  return rc;
}

int vkgs_capture_frame(vkgs_ctx* c)
{
  if(!c)
    return VKGS_ERR_INVALID_ARGUMENT;
  if(c->lastSlot < 0 || !c->slots[c->lastSlot].haveFrame)
  {
    c->lastError = "vkgs_capture_frame: no frame rendered yet";
    return VKGS_ERR_INVALID_ARGUMENT;
  }
  if(c->opt.target_format != VKGS_FORMAT_FLOAT32)
  {
    c->lastError = "vkgs_capture_frame needs the fp32 colour target";
    return VKGS_ERR_UNSUPPORTED;
  }
  CU_TRY(c, cudaSetDevice(c->device));
  FrameSlot&   s     = c->slots[c->lastSlot];
  const size_t bytes = sizeof(float4) * static_cast<size_t>(s.imgW) * s.imgH;
  CU_TRY(c, cudaStreamSynchronize(s.stream));
  if(c->captureW != s.imgW || c->captureH != s.imgH)
  {
    freeDev(c->dCapture);
    CU_TRY(c, cudaMalloc(&c->dCapture, bytes));
    c->captureW = s.imgW, c->captureH = s.imgH;
  }
  // on the frame's own stream: a device-to-device cudaMemcpy on the legacy stream is asynchronous to the host and the next
  // frame (non-blocking streams) could overwrite the image under it
  CU_TRY(c, cudaMemcpyAsync(c->dCapture, s.dImage, bytes, cudaMemcpyDeviceToDevice, s.stream));
  CU_TRY(c, cudaStreamSynchronize(s.stream));
  return VKGS_OK;
}

int vkgs_compare_with_capture(vkgs_ctx* c, uint32_t flip_mode, vkgs_image_metrics* out)
{
  if(!c || !out || flip_mode > VKGS_FLIP_REFERENCE)
    return VKGS_ERR_INVALID_ARGUMENT;
  if(!c->dCapture || c->lastSlot < 0)
  {
    c->lastError = "vkgs_compare_with_capture: no captured frame";
    return VKGS_ERR_INVALID_ARGUMENT;
  }
  FrameSlot& s = c->slots[c->lastSlot];
  if(c->opt.target_format != VKGS_FORMAT_FLOAT32 || s.imgW != c->captureW || s.imgH != c->captureH)
  {
    // (the reference resamples a differently sized current image with a bilinear sampler; not built)
    c->lastError = "vkgs_compare_with_capture needs the fp32 colour target and the capture's frame size";
    return VKGS_ERR_UNSUPPORTED;
  }
  CU_TRY(c, cudaSetDevice(c->device));
  CU_TRY(c, cudaStreamSynchronize(s.stream));
  return runMetrics(c, static_cast<const float4*>(c->dCapture), static_cast<const float4*>(s.dImage), s.imgW, s.imgH, flip_mode, out);
}

}  // extern "C"
