// CoMoE pre-stage routing as sparse index work.
// DeepSpeed top1gating builds dense (S,E,C) one-hot tensors and the reference dispatches / combines with
// einsum("sec,sm->ecm") / ("sec,ecm->sm") (src/UniGenUtils.py:140,183).  Here:
//   route_logits_kernel : fp32 gate logits, softmax, argmax (one warp per token)
//   route_select_kernel : per expert, Random-Token-Selection top-C by radix select on the uniform draw,
//                         slot = rank of the token among the kept tokens in token order (the cumsum of top1gating)
//   route_finish_kernel : l_aux
//   gather_modulate / combine : row gathers keyed by the slot maps.
// Integer outputs (expert_idx, slot, slot_token, exp_counts) are bit-exact w.r.t. the oracle given the same uniform.
#include "ug_host.h"
#include "ug_ptx.cuh"

namespace ug {

constexpr int kMaxExperts = 32;

__device__ __forceinline__ float warp_sum_f(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

// gates[token, e] = softmax_e( x[token,:] . wg[e,:] ), expert_idx = argmax (first max), prob = gates[argmax]
__global__ void __launch_bounds__(256) route_logits_kernel(const __nv_bfloat16* __restrict__ x,
                                                           const float* __restrict__ wg, int tokens, int d, int experts,
                                                           float* __restrict__ gates, int* __restrict__ expert_idx,
                                                           float* __restrict__ prob) {
  const int token = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int lane = threadIdx.x & 31;
  if (token >= tokens) return;
  float acc[kMaxExperts];
#pragma unroll
  for (int e = 0; e < kMaxExperts; ++e) acc[e] = 0.f;
  const __nv_bfloat16* xr = x + (long long)token * d;
  for (int kk = lane * 8; kk < d; kk += 256) {
    const uint4 u = *reinterpret_cast<const uint4*>(xr + kk);
    float2 a = unpack_bf16x2(u.x), b = unpack_bf16x2(u.y), c = unpack_bf16x2(u.z), dd = unpack_bf16x2(u.w);
    const float f[8] = {a.x, a.y, b.x, b.y, c.x, c.y, dd.x, dd.y};
#pragma unroll
    for (int e = 0; e < kMaxExperts; ++e) {
      if (e < experts) {
        const float4 w0 = __ldg(reinterpret_cast<const float4*>(wg + (long long)e * d + kk));
        const float4 w1 = __ldg(reinterpret_cast<const float4*>(wg + (long long)e * d + kk + 4));
        acc[e] += f[0] * w0.x + f[1] * w0.y + f[2] * w0.z + f[3] * w0.w + f[4] * w1.x + f[5] * w1.y + f[6] * w1.z + f[7] * w1.w;
      }
    }
  }
  float mx = -INFINITY;
  int arg = 0;
#pragma unroll
  for (int e = 0; e < kMaxExperts; ++e) {
    if (e < experts) {
      acc[e] = warp_sum_f(acc[e]);
      if (acc[e] > mx) { mx = acc[e]; arg = e; }
    }
  }
  float sum = 0.f;
#pragma unroll
  for (int e = 0; e < kMaxExperts; ++e) {
    if (e < experts) { acc[e] = expf(acc[e] - mx); sum += acc[e]; }
  }
  const float inv = 1.f / sum;
  if (lane == 0) {
    float p = 0.f;
#pragma unroll
    for (int e = 0; e < kMaxExperts; ++e) {
      if (e < experts) {
        const float g = acc[e] * inv;
        gates[(long long)token * experts + e] = g;
        if (e == arg) p = g;
      }
    }
    expert_idx[token] = arg;
    prob[token] = p;
  }
}

// One CTA per expert. Keeps the (up to) `capacity` assigned tokens with the largest uniform value, assigns slots in
// token order, fills the inverse map and the pre-capacity count.
__global__ void __launch_bounds__(1024) route_select_kernel(const int* __restrict__ expert_idx,
                                                            const float* __restrict__ uniform,
                                                            const float* __restrict__ gates, int tokens, int experts,
                                                            int capacity, int* __restrict__ slot,
                                                            int* __restrict__ slot_token,
                                                            long long* __restrict__ exp_counts,
                                                            float* __restrict__ me_sum) {
  const int e = blockIdx.x;
  const int tid = threadIdx.x;
  const int nthr = blockDim.x;
  __shared__ unsigned int hist[256];
  __shared__ unsigned int s_prefix, s_remaining, s_count, s_base;
  __shared__ unsigned int warp_tot[32];
  __shared__ float fsum[32];

  for (int i = tid; i < capacity; i += nthr) slot_token[(long long)e * capacity + i] = -1;
  // count assigned tokens and sum gates[:, e] (deterministic tree order)
  unsigned int cnt = 0;
  float gsum = 0.f;
  for (int t = tid; t < tokens; t += nthr) {
    cnt += (expert_idx[t] == e);
    gsum += gates[(long long)t * experts + e];
  }
  if (tid == 0) s_count = 0;
  __syncthreads();
  cnt = __reduce_add_sync(0xffffffffu, cnt);
  gsum = warp_sum_f(gsum);
  if ((tid & 31) == 0) { atomicAdd(&s_count, cnt); fsum[tid >> 5] = gsum; }
  __syncthreads();
  if (tid == 0) {
    float s = 0.f;
    for (int w = 0; w < (nthr >> 5); ++w) s += fsum[w];
    me_sum[e] = s;
    exp_counts[e] = (long long)s_count;
  }
  const unsigned int n_assigned = s_count;

  // threshold = capacity-th largest key among assigned tokens (keys are the IEEE bits of u >= 0, order preserving)
  unsigned int thr = 0;          // keep key > thr ...
  unsigned int ties_to_keep = 0; // ... plus the first `ties_to_keep` tokens with key == thr (token order)
  bool keep_all = n_assigned <= (unsigned int)capacity;
  if (!keep_all) {
    if (tid == 0) { s_prefix = 0; s_remaining = (unsigned int)capacity; }
    for (int shift = 24; shift >= 0; shift -= 8) {
      for (int i = tid; i < 256; i += nthr) hist[i] = 0;
      __syncthreads();
      const unsigned int prefix = s_prefix;
      const unsigned int himask = shift == 24 ? 0u : (0xffffffffu << (shift + 8));
      for (int t = tid; t < tokens; t += nthr) {
        if (expert_idx[t] == e) {
          const unsigned int key = __float_as_uint(uniform[(long long)t * experts + e]);
          if ((key & himask) == (prefix & himask)) atomicAdd(&hist[(key >> shift) & 255u], 1u);
        }
      }
      __syncthreads();
      if (tid == 0) {
        unsigned int rem = s_remaining;  // how many still to take from the current prefix group, from the top
        int b = 255;
        for (; b > 0; --b) {
          if (hist[b] >= rem) break;
          rem -= hist[b];
        }
        s_prefix = prefix | ((unsigned int)b << shift);
        s_remaining = rem;
      }
      __syncthreads();
    }
    thr = s_prefix;
    ties_to_keep = s_remaining;  // number of tokens with key == thr that are kept
  }

  // slot assignment: exclusive running count of kept tokens in token order
  if (tid == 0) s_base = 0;
  unsigned int ties_seen_base = 0;  // uniform across the block (tracked redundantly per thread via shared scans)
  __shared__ unsigned int s_ties_base;
  if (tid == 0) s_ties_base = 0;
  __syncthreads();
  for (int t0 = 0; t0 < tokens; t0 += nthr) {
    const int t = t0 + tid;
    bool mine = false, gt = false, eq = false;
    if (t < tokens && expert_idx[t] == e) {
      mine = true;
      if (!keep_all) {
        const unsigned int key = __float_as_uint(uniform[(long long)t * experts + e]);
        gt = key > thr;
        eq = key == thr;
      }
    }
    // first scan: rank among ties (only needed when not keep_all)
    unsigned int tie_rank = 0;
    const int lane = tid & 31, wid = tid >> 5;
    if (!keep_all) {
      const unsigned int bal = __ballot_sync(0xffffffffu, eq);
      const unsigned int before = __popc(bal & ((1u << lane) - 1u));
      if (lane == 0) warp_tot[wid] = __popc(bal);
      __syncthreads();
      unsigned int wbase = 0;
      for (int w = 0; w < wid; ++w) wbase += warp_tot[w];
      unsigned int total = 0;
      for (int w = 0; w < (nthr >> 5); ++w) total += warp_tot[w];
      ties_seen_base = s_ties_base;
      tie_rank = ties_seen_base + wbase + before;
      __syncthreads();
      if (tid == 0) s_ties_base = ties_seen_base + total;
    }
    const bool kept = mine && (keep_all || gt || (eq && tie_rank < ties_to_keep));
    const unsigned int bal = __ballot_sync(0xffffffffu, kept);
    const unsigned int before = __popc(bal & ((1u << lane) - 1u));
    if (lane == 0) warp_tot[wid] = __popc(bal);
    __syncthreads();
    unsigned int wbase = 0;
    for (int w = 0; w < wid; ++w) wbase += warp_tot[w];
    unsigned int total = 0;
    for (int w = 0; w < (nthr >> 5); ++w) total += warp_tot[w];
    const unsigned int base = s_base;
    if (mine) {
      if (kept) {
        const int sl = (int)(base + wbase + before);
        slot[t] = sl;
        slot_token[(long long)e * capacity + sl] = t;
      } else {
        slot[t] = -1;
      }
    }
    __syncthreads();
    if (tid == 0) s_base = base + total;
    __syncthreads();
  }
}

// l_aux = E * sum_e mean_s(gates[:,e]) * mean_s(mask1[:,e])
__global__ void route_finish_kernel(const float* __restrict__ me_sum, const long long* __restrict__ exp_counts, int tokens,
                                    int experts, float* __restrict__ l_aux) {
  if (threadIdx.x == 0 && blockIdx.x == 0) {
    float s = 0.f;
    for (int e = 0; e < experts; ++e) s += (me_sum[e] / (float)tokens) * ((float)exp_counts[e] / (float)tokens);
    *l_aux = s * (float)experts;
  }
}

__global__ void __launch_bounds__(256) gather_modulate_kernel(const __nv_bfloat16* __restrict__ x,
                                                              const int* __restrict__ slot_token,
                                                              const float* __restrict__ mod, long long mod_es,
                                                              long long mod_bs, const __nv_bfloat16* __restrict__ addend,
                                                              __nv_bfloat16* __restrict__ out, int experts, int capacity,
                                                              int tokens_per_batch, int d) {
  const int nvec = d >> 3;
  const long long total = (long long)experts * capacity * nvec;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const int v = (int)(i % nvec);
    const long long row = i / nvec;
    const int e = (int)(row / capacity);
    const int token = slot_token[row];
    uint4 o = make_uint4(0, 0, 0, 0);
    if (token >= 0) {
      const int b = token / tokens_per_batch;
      const uint4 u = *reinterpret_cast<const uint4*>(x + (long long)token * d + 8 * v);
      float2 p0 = unpack_bf16x2(u.x), p1 = unpack_bf16x2(u.y), p2 = unpack_bf16x2(u.z), p3 = unpack_bf16x2(u.w);
      float f[8] = {p0.x, p0.y, p1.x, p1.y, p2.x, p2.y, p3.x, p3.y};
      if (addend) {
        const uint4 a = *reinterpret_cast<const uint4*>(addend + row * d + 8 * v);
        float2 q0 = unpack_bf16x2(a.x), q1 = unpack_bf16x2(a.y), q2 = unpack_bf16x2(a.z), q3 = unpack_bf16x2(a.w);
        f[0] += q0.x; f[1] += q0.y; f[2] += q1.x; f[3] += q1.y; f[4] += q2.x; f[5] += q2.y; f[6] += q3.x; f[7] += q3.y;
        // the reference forms (hidden + cond') in bf16 before modulating: round once here as well
#pragma unroll
        for (int j = 0; j < 8; ++j) f[j] = __bfloat162float(__float2bfloat16(f[j]));
      }
      if (mod) {
        const float* m = mod + (long long)e * mod_es + (long long)b * mod_bs + 8 * v;
        const float4 m0 = *reinterpret_cast<const float4*>(m), m1 = *reinterpret_cast<const float4*>(m + 4);
        f[0] *= m0.x; f[1] *= m0.y; f[2] *= m0.z; f[3] *= m0.w; f[4] *= m1.x; f[5] *= m1.y; f[6] *= m1.z; f[7] *= m1.w;
      }
      o.x = pack_bf16x2(f[0], f[1]); o.y = pack_bf16x2(f[2], f[3]);
      o.z = pack_bf16x2(f[4], f[5]); o.w = pack_bf16x2(f[6], f[7]);
    }
    *reinterpret_cast<uint4*>(out + row * d + 8 * v) = o;
  }
}

__global__ void __launch_bounds__(256) combine_kernel(const __nv_bfloat16* __restrict__ y, const int* __restrict__ expert_idx,
                                                      const int* __restrict__ slot, const float* __restrict__ prob,
                                                      __nv_bfloat16* __restrict__ out, int tokens, int capacity, int d) {
  const int nvec = d >> 3;
  const long long total = (long long)tokens * nvec;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const int v = (int)(i % nvec);
    const int t = (int)(i / nvec);
    const int sl = slot[t];
    uint4 o = make_uint4(0, 0, 0, 0);
    if (sl >= 0) {
      const float p = prob[t];
      const long long row = (long long)expert_idx[t] * capacity + sl;
      const uint4 u = *reinterpret_cast<const uint4*>(y + row * d + 8 * v);
      float2 p0 = unpack_bf16x2(u.x), p1 = unpack_bf16x2(u.y), p2 = unpack_bf16x2(u.z), p3 = unpack_bf16x2(u.w);
      o.x = pack_bf16x2(p * p0.x, p * p0.y); o.y = pack_bf16x2(p * p1.x, p * p1.y);
      o.z = pack_bf16x2(p * p2.x, p * p2.y); o.w = pack_bf16x2(p * p3.x, p * p3.y);
    }
    *reinterpret_cast<uint4*>(out + (long long)t * d + 8 * v) = o;
  }
}

static inline int grid_cap(long long threads, int block) {
  long long g = (threads + block - 1) / block;
  const long long cap = (long long)num_sms() * 16;
  return (int)(g < 1 ? 1 : (g > cap ? cap : g));
}

}  // namespace ug

using namespace ug;

extern "C" int ug_moe_route(const void* x, const float* wg, const float* rts_uniform, int32_t tokens, int32_t d,
                            int32_t experts, int32_t capacity, int32_t* expert_idx, int32_t* slot, float* prob,
                            int32_t* slot_token, int64_t* exp_counts, float* l_aux, float* workspace, void* stream) {
  UG_CHECK_ARG(x && wg && rts_uniform && expert_idx && slot && prob && slot_token && exp_counts && l_aux && workspace,
               "moe_route: null pointer");
  UG_CHECK_ARG(tokens >= 1 && d >= 8 && d % 8 == 0 && experts >= 1 && experts <= kMaxExperts && capacity >= 1,
               "moe_route: bad shape tokens %d d %d experts %d capacity %d", tokens, d, experts, capacity);
  UG_CHECK_ARG((reinterpret_cast<uintptr_t>(x) & 15) == 0 && (reinterpret_cast<uintptr_t>(wg) & 15) == 0, "moe_route: alignment");
  auto s = reinterpret_cast<cudaStream_t>(stream);
  // workspace layout: gates [tokens*experts] | me_sum [experts]
  float* gates = workspace;
  float* me_sum = workspace + (long long)tokens * experts;
  const int block = 256;
  const int grid = (int)(((long long)tokens * 32 + block - 1) / block);
  route_logits_kernel<<<grid, block, 0, s>>>((const __nv_bfloat16*)x, wg, tokens, d, experts, gates, expert_idx, prob);
  UG_CHECK_LAUNCH("moe_route/logits");
  route_select_kernel<<<experts, 1024, 0, s>>>(expert_idx, rts_uniform, gates, tokens, experts, capacity, slot, slot_token,
                                               (long long*)exp_counts, me_sum);
  UG_CHECK_LAUNCH("moe_route/select");
  route_finish_kernel<<<1, 32, 0, s>>>(me_sum, (const long long*)exp_counts, tokens, experts, l_aux);
  UG_CHECK_LAUNCH("moe_route/finish");
  return UG_OK;
}

extern "C" int ug_moe_gather_modulate(const void* x, const int32_t* slot_token, const float* mod, int64_t mod_es,
                                      int64_t mod_bs, const void* addend, void* out, int32_t experts, int32_t capacity,
                                      int32_t tokens_per_batch, int32_t d, void* stream) {
  UG_CHECK_ARG(x && slot_token && out, "moe_gather_modulate: null pointer");
  UG_CHECK_ARG(experts >= 1 && capacity >= 1 && tokens_per_batch >= 1 && d >= 8 && d % 8 == 0, "moe_gather_modulate: bad shape");
  UG_CHECK_ARG(mod_es % 4 == 0 && mod_bs % 4 == 0, "moe_gather_modulate: modulation strides must be multiples of 4");
  const long long total = (long long)experts * capacity * (d >> 3);
  gather_modulate_kernel<<<grid_cap(total, 256), 256, 0, reinterpret_cast<cudaStream_t>(stream)>>>(
      (const __nv_bfloat16*)x, slot_token, mod, mod_es, mod_bs, (const __nv_bfloat16*)addend, (__nv_bfloat16*)out, experts,
      capacity, tokens_per_batch, d);
  UG_CHECK_LAUNCH("moe_gather_modulate");
  return UG_OK;
}

extern "C" int ug_moe_combine(const void* y, const int32_t* expert_idx, const int32_t* slot, const float* prob, void* out,
                              int32_t tokens, int32_t capacity, int32_t d, void* stream) {
  UG_CHECK_ARG(y && expert_idx && slot && prob && out, "moe_combine: null pointer");
  UG_CHECK_ARG(tokens >= 1 && capacity >= 1 && d >= 8 && d % 8 == 0, "moe_combine: bad shape");
  const long long total = (long long)tokens * (d >> 3);
  combine_kernel<<<grid_cap(total, 256), 256, 0, reinterpret_cast<cudaStream_t>(stream)>>>(
      (const __nv_bfloat16*)y, expert_idx, slot, prob, (__nv_bfloat16*)out, tokens, capacity, d);
  UG_CHECK_LAUNCH("moe_combine");
  return UG_OK;
}
