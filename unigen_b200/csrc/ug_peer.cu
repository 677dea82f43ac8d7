// Peer-memory exchange for Ulysses sequence parallelism over NVLink 5 / NVSwitch (one process per GPU).
//
// The reference has no sequence parallelism; the north-star asks for Ulysses attention on one
// 8 x B200 box. Instead of staging copies + NCCL all-to-all around attention, the exchange is FUSED into the kernels
// that produce the data, over peer-mapped memory (CUDA IPC):
//   * ug_qkv_scatter      reads the local rows of the fused q|k|v projection once, applies the per-head RMSNorm + RoPE
//                         (the in-place pass of the single-GPU path) and stores every head straight into the receive
//                         buffer of the rank that owns that head  (seq-shard x all heads -> all tokens x head-shard).
//   * ug_attention_bf16_peer (ug_attention.cu) stores every output row straight into the attention-output buffer of the
//                         rank that owns that token row              (all tokens x head-shard -> seq-shard x all heads).
//   * ug_peer_bcast_rows  all-gather by peer stores (residual stream for the replicated CoMoE pre-stage, final velocity).
//   * ug_peer_barrier     flag barrier between those phases (st.release.sys / ld.acquire.sys on peer flags); no host sync,
//                         no NCCL, capturable in a CUDA graph (the epoch counter lives in device memory).
// Pool layout (every rank allocates the same size): bytes [0, UG_PEER_HEADER_BYTES) = control block (flags[src rank] u32 at
// 0.., epoch u32 at 256, error u32 at 260), payload behind it at offsets the host chooses identically on all ranks.
#include <cstdlib>

#include "ug_host.h"
#include "ug_ptx.cuh"

namespace ug {

struct PeerPtrs {
  int world, rank;
  uint8_t* base[UG_MAX_PEERS];
};

__device__ __forceinline__ void st_release_sys(unsigned int* p, unsigned int v) {
  asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
__device__ __forceinline__ unsigned int ld_acquire_sys(const unsigned int* p) {
  unsigned int v;
  asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}

// One block; thread r < world signals rank r and waits for rank r's signal of the same epoch. Every peer store of the kernels
// launched before it on this stream happens-before this kernel (stream order); the system-scope fence + release store make
// them visible to a peer that acquires the flag, so the producing kernels need no fence of their own. A rank that never arrives
// (a bug, a dead peer, or host-side skew longer than the timeout) must not hang the GPU: after `timeout_ns` of spinning the
// STICKY error word (ctrl[65]) is set and the kernel returns; once it is set, later barriers still signal their peers but do
// not wait, so a failure costs one timeout, not one per barrier. The host surfaces the word (ug_peer_error; the sequence-parallel
// modules raise on it) — results produced after it was set are invalid.
__global__ void peer_barrier_kernel(PeerPtrs t, unsigned long long timeout_ns) {
  __shared__ unsigned int epoch, failed;
  unsigned int* ctrl = reinterpret_cast<unsigned int*>(t.base[t.rank]);
  if (threadIdx.x == 0) {
    epoch = ctrl[64] + 1;
    ctrl[64] = epoch;
    failed = ctrl[65];
  }
  __syncthreads();
  if ((int)threadIdx.x < t.world) {
    __threadfence_system();
    st_release_sys(reinterpret_cast<unsigned int*>(t.base[threadIdx.x]) + t.rank, epoch);
    if (!failed) {
      const unsigned int* mine = ctrl + threadIdx.x;
      const uint64_t t0 = globaltimer_ns();
      unsigned int spins = 0;
      while ((int)(ld_acquire_sys(mine) - epoch) < 0) {
        if ((++spins & 0xffu) == 0 && globaltimer_ns() - t0 > timeout_ns) {
          atomicExch(ctrl + 65, 1u + threadIdx.x);  // 1 + the rank that did not arrive
          break;
        }
      }
    }
    __threadfence_system();
  }
}

static unsigned long long g_peer_timeout_ns = 0;  // 0 = not initialised yet
static unsigned long long peer_timeout_ns() {
  if (g_peer_timeout_ns == 0) {
    unsigned long long ms = 20000;  // default 20 s: covers graph instantiation / IO skew between ranks
    if (const char* env = getenv("UG_PEER_TIMEOUT_MS")) {
      const long long v = atoll(env);
      if (v > 0) ms = (unsigned long long)v;
    }
    g_peer_timeout_ns = ms * 1000000ull;
  }
  return g_peer_timeout_ns;
}

__device__ __forceinline__ void unpack8f(const uint4& u, float (&f)[8]) {
  float2 a = unpack_bf16x2(u.x), b = unpack_bf16x2(u.y), c = unpack_bf16x2(u.z), d = unpack_bf16x2(u.w);
  f[0] = a.x; f[1] = a.y; f[2] = b.x; f[3] = b.y; f[4] = c.x; f[5] = c.y; f[6] = d.x; f[7] = d.y;
}
__device__ __forceinline__ uint4 pack8f(const float (&f)[8]) {
  uint4 u;
  u.x = pack_bf16x2(f[0], f[1]); u.y = pack_bf16x2(f[2], f[3]);
  u.z = pack_bf16x2(f[4], f[5]); u.w = pack_bf16x2(f[6], f[7]);
  return u;
}

// One warp per local token row; 16 bytes per lane, a head spans LPH lanes, a warp covers HPW heads per pass.
// Column block `which` = h / heads (0 = q, 1 = k: RMSNorm(weight[which]) + RoPE; 2 = v: pass-through).
// Head hh = h % heads goes to rank hh / heads_per_rank at  recv[which][dst_row0 + r][(hh % heads_per_rank) * dh + e].
template <int kDh>
__global__ void __launch_bounds__(256) qkv_scatter_kernel(const __nv_bfloat16* __restrict__ qkv, long long rs, int rows, int heads,
                                                          const __nv_bfloat16* __restrict__ w, float eps,
                                                          const float* __restrict__ cos_sin, PeerPtrs t, long long dst_offset,
                                                          int seq_total, int dst_row0) {
  constexpr int LPH = kDh / 8, HPW = 32 / LPH, kUnroll = 4;
  // one warp per (row, chunk of HPW * kUnroll heads): a row is spread over several warps so that enough 16-byte stores are
  // in flight to cover the NVLink latency even when a rank owns only a few hundred rows
  const int chunks = (3 * heads + HPW * kUnroll - 1) / (HPW * kUnroll);
  const int wg = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int r = wg / chunks, chunk = wg % chunks;
  const int lane = threadIdx.x & 31;
  if (r >= rows) return;
  const __nv_bfloat16* row = qkv + (long long)r * rs;
  const int sub = lane / LPH, e0 = (lane % LPH) * 8;
  const int hpr = heads / t.world;       // heads per rank
  const long long hd = (long long)hpr * kDh;  // receive-buffer row length
  float cs[4], sn[4];
  if (cos_sin) {
    const float* c = cos_sin + (long long)r * kDh + e0;
    const float4 t0 = *reinterpret_cast<const float4*>(c), t1 = *reinterpret_cast<const float4*>(c + 4);
    cs[0] = t0.x; sn[0] = t0.y; cs[1] = t0.z; sn[1] = t0.w; cs[2] = t1.x; sn[2] = t1.y; cs[3] = t1.z; sn[3] = t1.w;
  }
  const int total_heads = 3 * heads;
  {
    const int h0 = chunk * HPW * kUnroll;
    uint4 u[kUnroll];
#pragma unroll
    for (int k = 0; k < kUnroll; ++k) {
      const int h = h0 + k * HPW + sub;
      if (h < total_heads) u[k] = *reinterpret_cast<const uint4*>(row + (long long)h * kDh + e0);
    }
#pragma unroll
    for (int k = 0; k < kUnroll; ++k) {
      const int h = h0 + k * HPW + sub;
      const bool on = h < total_heads;
      const int which = on ? h / heads : 2, hh = on ? h % heads : 0;
      float f[8];
      if (on) unpack8f(u[k], f);
      else {
#pragma unroll
        for (int i = 0; i < 8; ++i) f[i] = 0.f;
      }
      float ss = 0.f;
#pragma unroll
      for (int i = 0; i < 8; ++i) ss += f[i] * f[i];
#pragma unroll
      for (int o = LPH / 2; o > 0; o >>= 1) ss += __shfl_xor_sync(0xffffffffu, ss, o);
      if (!on) continue;
      uint4 outv = u[k];
      if (which < 2) {
        if (w) {
          const float rstd = rsqrtf(ss / (float)kDh + eps);
          float wf[8];
          unpack8f(*reinterpret_cast<const uint4*>(w + which * kDh + e0), wf);
#pragma unroll
          for (int i = 0; i < 8; ++i) f[i] = f[i] * rstd * wf[i];
        }
        if (cos_sin) {
#pragma unroll
          for (int q = 0; q < 4; ++q) {
            const float x0 = f[2 * q], x1 = f[2 * q + 1];
            f[2 * q] = x0 * cs[q] - x1 * sn[q];
            f[2 * q + 1] = x1 * cs[q] + x0 * sn[q];
          }
        }
        outv = pack8f(f);
      }
      const int dst_rank = hh / hpr;
      __nv_bfloat16* dst = reinterpret_cast<__nv_bfloat16*>(t.base[dst_rank] + dst_offset) +
                           ((long long)which * seq_total + dst_row0 + r) * hd + (long long)(hh % hpr) * kDh + e0;
      *reinterpret_cast<uint4*>(dst) = outv;
    }
  }
  // no per-thread system fence here (it cost ~45 % of this kernel's stall samples): the stores are ordered before the
  // following ug_peer_barrier by stream order, and that kernel's fence + release store publishes them to the peers
}

// dst pool rows [dst_row0, dst_row0 + rows) of EVERY rank <- src rows (16-byte vectors; blockIdx.y = destination rank)
__global__ void __launch_bounds__(256) peer_bcast_rows_kernel(const __nv_bfloat16* __restrict__ src, long long s_rs, int rows, int d,
                                                              PeerPtrs t, long long dst_offset, long long d_rs, int dst_row0) {
  const int nvec = d >> 3;
  const long long total = (long long)rows * nvec;
  __nv_bfloat16* dst = reinterpret_cast<__nv_bfloat16*>(t.base[blockIdx.y] + dst_offset) + (long long)dst_row0 * d_rs;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const int v = (int)(i % nvec);
    const long long r = i / nvec;
    *reinterpret_cast<uint4*>(dst + r * d_rs + 8 * v) = *reinterpret_cast<const uint4*>(src + r * s_rs + 8 * v);
  }
}

static int to_ptrs(const ug_peer_table* t, PeerPtrs* p, const char* name) {
  UG_CHECK_ARG(t != nullptr, "%s: null peer table", name);
  UG_CHECK_ARG(t->world >= 1 && t->world <= UG_MAX_PEERS && t->rank >= 0 && t->rank < t->world, "%s: bad world %d / rank %d", name,
               t->world, t->rank);
  p->world = t->world;
  p->rank = t->rank;
  for (int i = 0; i < UG_MAX_PEERS; ++i) {
    p->base[i] = i < t->world ? reinterpret_cast<uint8_t*>(t->base[i]) : nullptr;
    UG_CHECK_ARG(i >= t->world || (t->base[i] && (reinterpret_cast<uintptr_t>(t->base[i]) & 255) == 0), "%s: peer base %d null / unaligned",
                 name, i);
  }
  return UG_OK;
}

}  // namespace ug

using namespace ug;

#define UG_CUDA_CALL(expr, what)                                          \
  do {                                                                    \
    cudaError_t e__ = (expr);                                             \
    if (e__ != cudaSuccess) {                                             \
      (void)cudaGetLastError();                                           \
      set_error("%s: %s", what, cudaGetErrorString(e__));                 \
      return UG_ERR_CUDA;                                                 \
    }                                                                     \
  } while (0)

extern "C" int ug_peer_alloc(size_t bytes, void** dev_ptr) {
  UG_CHECK_ARG(dev_ptr && bytes >= UG_PEER_HEADER_BYTES, "peer_alloc: need at least the %d-byte control block", UG_PEER_HEADER_BYTES);
  void* p = nullptr;
  UG_CUDA_CALL(cudaMalloc(&p, bytes), "peer_alloc: cudaMalloc");
  UG_CUDA_CALL(cudaMemset(p, 0, bytes), "peer_alloc: cudaMemset");
  UG_CUDA_CALL(cudaDeviceSynchronize(), "peer_alloc: sync");
  *dev_ptr = p;
  return UG_OK;
}
extern "C" int ug_peer_free(void* dev_ptr) {
  UG_CHECK_ARG(dev_ptr, "peer_free: null pointer");
  UG_CUDA_CALL(cudaFree(dev_ptr), "peer_free");
  return UG_OK;
}
extern "C" int ug_peer_export(const void* dev_ptr, uint8_t handle[UG_PEER_HANDLE_BYTES]) {
  static_assert(sizeof(cudaIpcMemHandle_t) == UG_PEER_HANDLE_BYTES, "IPC handle size");
  UG_CHECK_ARG(dev_ptr && handle, "peer_export: null pointer");
  cudaIpcMemHandle_t h;
  UG_CUDA_CALL(cudaIpcGetMemHandle(&h, const_cast<void*>(dev_ptr)), "peer_export: cudaIpcGetMemHandle");
  memcpy(handle, &h, sizeof(h));
  return UG_OK;
}
extern "C" int ug_peer_open(const uint8_t handle[UG_PEER_HANDLE_BYTES], void** peer_ptr) {
  UG_CHECK_ARG(handle && peer_ptr, "peer_open: null pointer");
  cudaIpcMemHandle_t h;
  memcpy(&h, handle, sizeof(h));
  void* p = nullptr;
  UG_CUDA_CALL(cudaIpcOpenMemHandle(&p, h, cudaIpcMemLazyEnablePeerAccess), "peer_open: cudaIpcOpenMemHandle");
  *peer_ptr = p;
  return UG_OK;
}
extern "C" int ug_peer_close(void* peer_ptr) {
  UG_CHECK_ARG(peer_ptr, "peer_close: null pointer");
  UG_CUDA_CALL(cudaIpcCloseMemHandle(peer_ptr), "peer_close");
  return UG_OK;
}

extern "C" int ug_peer_barrier(const ug_peer_table* table, void* stream) {
  PeerPtrs t;
  int st = to_ptrs(table, &t, "peer_barrier");
  if (st != UG_OK) return st;
  peer_barrier_kernel<<<1, 32, 0, reinterpret_cast<cudaStream_t>(stream)>>>(t, peer_timeout_ns());
  UG_CHECK_LAUNCH("peer_barrier");
  return UG_OK;
}

extern "C" int ug_peer_set_timeout_ms(int64_t ms) {
  UG_CHECK_ARG(ms >= 1, "peer_set_timeout_ms: timeout must be positive");
  g_peer_timeout_ns = (unsigned long long)ms * 1000000ull;
  return UG_OK;
}

// Asynchronous read-out of the sticky error word into (pinned) host memory, stream-ordered and graph-capturable: the
// sequence-parallel modules enqueue it at the end of every forward and test the host word before the next one.
extern "C" int ug_peer_error_async(const ug_peer_table* table, int32_t* error_host_pinned, void* stream) {
  PeerPtrs t;
  int st = to_ptrs(table, &t, "peer_error_async");
  if (st != UG_OK) return st;
  UG_CHECK_ARG(error_host_pinned, "peer_error_async: null pointer");
  UG_CUDA_CALL(cudaMemcpyAsync(error_host_pinned, t.base[t.rank] + 260, 4, cudaMemcpyDeviceToHost, reinterpret_cast<cudaStream_t>(stream)),
               "peer_error_async: cudaMemcpyAsync");
  return UG_OK;
}

extern "C" int ug_peer_error(const ug_peer_table* table, int32_t* error_host) {
  PeerPtrs t;
  int st = to_ptrs(table, &t, "peer_error");
  if (st != UG_OK) return st;
  UG_CHECK_ARG(error_host, "peer_error: null pointer");
  unsigned int v = 0;
  UG_CUDA_CALL(cudaMemcpy(&v, t.base[t.rank] + 260, 4, cudaMemcpyDeviceToHost), "peer_error: cudaMemcpy");
  *error_host = (int32_t)v;
  return UG_OK;
}

extern "C" int ug_qkv_scatter(const ug_peer_table* table, const ug_qkv_scatter_args* a, void* stream) {
  PeerPtrs t;
  int st = to_ptrs(table, &t, "qkv_scatter");
  if (st != UG_OK) return st;
  UG_CHECK_ARG(a && a->qkv, "qkv_scatter: null pointer");
  UG_CHECK_ARG(a->rows >= 1 && a->heads >= 1 && a->heads % t.world == 0, "qkv_scatter: heads %d must be a multiple of world %d",
               a->heads, t.world);
  UG_CHECK_ARG(a->row_stride % 8 == 0 && (reinterpret_cast<uintptr_t>(a->qkv) & 15) == 0 && a->dst_offset % 16 == 0 &&
                   a->dst_offset >= UG_PEER_HEADER_BYTES && (!a->norm_weight || (reinterpret_cast<uintptr_t>(a->norm_weight) & 15) == 0) &&
                   (!a->cos_sin || (reinterpret_cast<uintptr_t>(a->cos_sin) & 15) == 0),
               "qkv_scatter: alignment");
  UG_CHECK_ARG(a->dst_row0 >= 0 && a->dst_row0 + a->rows <= a->seq_total, "qkv_scatter: rows [%d, %d) outside the %d-row sequence",
               a->dst_row0, a->dst_row0 + a->rows, a->seq_total);
  const int heads_per_warp = (32 / (a->head_dim / 8)) * 4;
  const long long warps = (long long)a->rows * ((3 * a->heads + heads_per_warp - 1) / heads_per_warp);
  const int grid = (int)((warps * 32 + 255) / 256);
  auto s = reinterpret_cast<cudaStream_t>(stream);
  auto q = (const __nv_bfloat16*)a->qkv;
  auto w = (const __nv_bfloat16*)a->norm_weight;
  if (a->head_dim == 128)
    qkv_scatter_kernel<128><<<grid, 256, 0, s>>>(q, a->row_stride, a->rows, a->heads, w, a->eps, a->cos_sin, t, a->dst_offset, a->seq_total, a->dst_row0);
  else if (a->head_dim == 64)
    qkv_scatter_kernel<64><<<grid, 256, 0, s>>>(q, a->row_stride, a->rows, a->heads, w, a->eps, a->cos_sin, t, a->dst_offset, a->seq_total, a->dst_row0);
  else {
    set_error("qkv_scatter: head_dim %d not supported (64 or 128)", a->head_dim);
    return UG_ERR_UNSUPPORTED;
  }
  UG_CHECK_LAUNCH("qkv_scatter");
  return UG_OK;
}

extern "C" int ug_peer_bcast_rows(const ug_peer_table* table, const void* src, int64_t src_row_stride, int32_t rows, int32_t d,
                                  int64_t dst_offset, int64_t dst_row_stride, int32_t dst_row0, void* stream) {
  PeerPtrs t;
  int st = to_ptrs(table, &t, "peer_bcast_rows");
  if (st != UG_OK) return st;
  UG_CHECK_ARG(src && rows >= 1 && d >= 8 && d % 8 == 0 && dst_row0 >= 0, "peer_bcast_rows: bad shape");
  UG_CHECK_ARG(src_row_stride % 8 == 0 && dst_row_stride % 8 == 0 && dst_offset % 16 == 0 && dst_offset >= UG_PEER_HEADER_BYTES &&
                   (reinterpret_cast<uintptr_t>(src) & 15) == 0, "peer_bcast_rows: alignment");
  const long long total = (long long)rows * (d >> 3);
  long long gx = (total + 255) / 256;
  const long long cap = (long long)num_sms() * 4;
  if (gx > cap) gx = cap;
  peer_bcast_rows_kernel<<<dim3((unsigned)gx, (unsigned)t.world), 256, 0, reinterpret_cast<cudaStream_t>(stream)>>>(
      (const __nv_bfloat16*)src, src_row_stride, rows, d, t, dst_offset, dst_row_stride, dst_row0);
  UG_CHECK_LAUNCH("peer_bcast_rows");
  return UG_OK;
}
