// Denoise-loop glue on the device: Euler flow-match update, classifier-free-guidance combine and
// the latent pack / unpack permutations, so that a multi-step sampling loop never leaves the GPU between steps
// (reference: src/UniGenPipeline.py:1050-1116 loop body, :405-412 CFG, diffusers FlowMatchEulerDiscreteScheduler.step,
// FluxPipeline._pack_latents / _unpack_latents).
#include "ug_host.h"
#include "ug_ptx.cuh"

namespace ug {

// x <- bf16( float(x) + (sigma_next - sigma) * float(v) )
__global__ void __launch_bounds__(256) euler_step_kernel(__nv_bfloat16* __restrict__ x, const __nv_bfloat16* __restrict__ v,
                                                         float dsigma, long long n) {
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x)
    x[i] = __float2bfloat16(__bfloat162float(x[i]) + dsigma * __bfloat162float(v[i]));
}
// same, sigma / sigma_next read from the device-resident schedule
__global__ void __launch_bounds__(256) euler_step_table_kernel(__nv_bfloat16* __restrict__ x, const __nv_bfloat16* __restrict__ v,
                                                               const float* __restrict__ sigmas, int step, long long n) {
  const float dsigma = sigmas[step + 1] - sigmas[step];
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x)
    x[i] = __float2bfloat16(__bfloat162float(x[i]) + dsigma * __bfloat162float(v[i]));
}
// out = uncond + g * (text - uncond)
__global__ void __launch_bounds__(256) cfg_combine_kernel(const __nv_bfloat16* __restrict__ uncond,
                                                          const __nv_bfloat16* __restrict__ text, float g,
                                                          __nv_bfloat16* __restrict__ out, long long n) {
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
    const float u = __bfloat162float(uncond[i]);
    out[i] = __float2bfloat16(u + g * (__bfloat162float(text[i]) - u));
  }
}
// pack: (B, C, H, W) -> (B, (H/2)(W/2), C*4) with token = (h/2, w/2), channel = c*4 + (h%2)*2 + (w%2); unpack = inverse
template <bool kPack>
__global__ void __launch_bounds__(256) pack_latents_kernel(const __nv_bfloat16* __restrict__ src, __nv_bfloat16* __restrict__ dst,
                                                           int B, int C, int H, int W) {
  const long long total = (long long)B * C * H * W;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    // i indexes the PACKED layout (b, token, ch)
    const int ch = (int)(i % (C * 4));
    const long long t = i / (C * 4);
    const int tok = (int)(t % ((H / 2) * (W / 2)));
    const int b = (int)(t / ((H / 2) * (W / 2)));
    const int c = ch >> 2, dy = (ch >> 1) & 1, dx = ch & 1;
    const int h = (tok / (W / 2)) * 2 + dy, w = (tok % (W / 2)) * 2 + dx;
    const long long j = (((long long)b * C + c) * H + h) * W + w;  // index in (B, C, H, W)
    if (kPack) dst[i] = src[j]; else dst[j] = src[i];
  }
}

// SD3 un-patchify (src/UniGenTransformer.py:693-704): tokens (B, h*w, p*p*C) with channel = (py*p + px)*C + c
// -> image (B, C, h*p, w*p).  i indexes the IMAGE layout so that the stores are coalesced.
__global__ void __launch_bounds__(256) unpatchify_kernel(const __nv_bfloat16* __restrict__ tok, __nv_bfloat16* __restrict__ img,
                                                         int B, int h, int w, int p, int C) {
  const int H = h * p, W = w * p;
  const long long total = (long long)B * C * H * W;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const int x = (int)(i % W);
    long long t = i / W;
    const int y = (int)(t % H);
    t /= H;
    const int c = (int)(t % C), b = (int)(t / C);
    const long long token = ((long long)b * h + y / p) * w + x / p;
    img[i] = tok[token * (p * p * C) + ((y % p) * p + (x % p)) * C + c];
  }
}

static inline int grid1d(long long n) {
  long long g = (n + 255) / 256;
  const long long cap = (long long)num_sms() * 16;
  return (int)(g < 1 ? 1 : (g > cap ? cap : g));
}
}  // namespace ug

using namespace ug;

extern "C" int ug_euler_step(void* latents, const void* velocity, float sigma, float sigma_next, int64_t n, void* stream) {
  UG_CHECK_ARG(latents && velocity && n >= 1, "euler_step: bad arguments");
  euler_step_kernel<<<grid1d(n), 256, 0, reinterpret_cast<cudaStream_t>(stream)>>>((__nv_bfloat16*)latents, (const __nv_bfloat16*)velocity,
                                                                                sigma_next - sigma, n);
  UG_CHECK_LAUNCH("euler_step");
  return UG_OK;
}
extern "C" int ug_euler_step_table(void* latents, const void* velocity, const float* sigmas_dev, int32_t step, int64_t n, void* stream) {
  UG_CHECK_ARG(latents && velocity && sigmas_dev && step >= 0 && n >= 1, "euler_step_table: bad arguments");
  euler_step_table_kernel<<<grid1d(n), 256, 0, reinterpret_cast<cudaStream_t>(stream)>>>((__nv_bfloat16*)latents, (const __nv_bfloat16*)velocity,
                                                                                      sigmas_dev, step, n);
  UG_CHECK_LAUNCH("euler_step_table");
  return UG_OK;
}
extern "C" int ug_cfg_combine(const void* uncond, const void* text, float guidance_scale, void* out, int64_t n, void* stream) {
  UG_CHECK_ARG(uncond && text && out && n >= 1, "cfg_combine: bad arguments");
  cfg_combine_kernel<<<grid1d(n), 256, 0, reinterpret_cast<cudaStream_t>(stream)>>>((const __nv_bfloat16*)uncond, (const __nv_bfloat16*)text,
                                                                                 guidance_scale, (__nv_bfloat16*)out, n);
  UG_CHECK_LAUNCH("cfg_combine");
  return UG_OK;
}
extern "C" int ug_pack_latents(const void* src, void* dst, int32_t batch, int32_t channels, int32_t height, int32_t width, int32_t unpack,
                               void* stream) {
  UG_CHECK_ARG(src && dst && batch >= 1 && channels >= 1 && height >= 2 && width >= 2 && height % 2 == 0 && width % 2 == 0,
               "pack_latents: bad arguments (height / width must be even)");
  const long long n = (long long)batch * channels * height * width;
  auto s = reinterpret_cast<cudaStream_t>(stream);
  if (unpack) pack_latents_kernel<false><<<grid1d(n), 256, 0, s>>>((const __nv_bfloat16*)src, (__nv_bfloat16*)dst, batch, channels, height, width);
  else pack_latents_kernel<true><<<grid1d(n), 256, 0, s>>>((const __nv_bfloat16*)src, (__nv_bfloat16*)dst, batch, channels, height, width);
  UG_CHECK_LAUNCH("pack_latents");
  return UG_OK;
}
extern "C" int ug_unpatchify(const void* tokens, void* image, int32_t batch, int32_t h, int32_t w, int32_t p, int32_t channels,
                             void* stream) {
  UG_CHECK_ARG(tokens && image && batch >= 1 && h >= 1 && w >= 1 && p >= 1 && channels >= 1, "unpatchify: bad arguments");
  const long long n = (long long)batch * channels * h * p * w * p;
  unpatchify_kernel<<<grid1d(n), 256, 0, reinterpret_cast<cudaStream_t>(stream)>>>((const __nv_bfloat16*)tokens, (__nv_bfloat16*)image,
                                                                                 batch, h, w, p, channels);
  UG_CHECK_LAUNCH("unpatchify");
  return UG_OK;
}
