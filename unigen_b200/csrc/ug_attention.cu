// Joint attention for sm_100a: softmax(Q K^T * scale + segment mask) V, non-causal, one (128-query tile, head, batch)
// per CTA.  TMA feeds 128B-swizzled Q/K/V tiles; both GEMMs run on tcgen05 with accumulators in TMEM:
//   S = Q K^T  (SS MMA, K-major operands)             -> TMEM S[2] (double buffered, 128 fp32 columns each)
//   O += P V   (P from TMEM (variant 1) or smem (2); V is the MN-major B operand) -> TMEM O (head_dim columns)
// Warp roles (192 threads): warp 0 TMA producer, warp 1 MMA issuer / TMEM owner, warps 2..5 softmax + epilogue
// (one query row per thread; online softmax in fp32, exp2 domain; O is rescaled in TMEM only when a row max moves).
// The segment rule (condition tokens / image tokens visibility) is applied as whole-tile skipping plus a per-row
// column mask on tiles that straddle a segment boundary; it is the same rule ug_expand_segment_mask materialises.
//
// Replaces F.scaled_dot_product_attention in diffusers FluxAttnProcessor2_0.
#include <algorithm>

#include "ug_host.h"
#include "ug_ptx.cuh"

namespace ug {

constexpr int kBlockQ = 128;
constexpr int kBlockKV = 128;
constexpr int kMaxTiles = 512;  // seq <= 65536
constexpr float kRescaleLog2 = 8.0f;  // lazy-rescale threshold in the exp2 domain

struct AttnParams {
  __nv_bfloat16* o;
  long long o_rs, o_bs;
  int seq, heads, batch;
  float scale_log2;  // scale * log2(e)
  int n_seg;
  int bounds[UG_MAX_SEGMENTS + 1];
  unsigned int visible[UG_MAX_SEGMENTS];
  // Ulysses heads->sequence exchange fused into the epilogue (ug_attention_bf16_peer): query row q belongs to rank
  // q / o_rows_per_rank and is stored at o_peer[rank] (that rank's peer-mapped output buffer, already offset to this
  // rank's head columns) + (q % o_rows_per_rank) * o_rs.  o_rows_per_rank == 0: plain local output `o`.
  int o_rows_per_rank;
  __nv_bfloat16* o_peer[UG_MAX_PEERS];
  // CTA order: 0 = query tile fastest (grid = (q_tiles, heads, batch): concurrently running CTAs share a head's K/V in L2);
  // 1 = head fastest (grid = (heads, q_tiles, batch)): with few CTAs per SM and a segment mask, the query tiles with the
  // most visible keys (text / image rows come first) all start in the first wave instead of trailing in the last one
  int head_fastest;
  int q_tiles2, n_units;  // persistent two-tile kernel: 256-row query tiles and work units (= q_tiles2 * heads * batch)
};

__device__ __forceinline__ __nv_bfloat16* o_row_ptr(const AttnParams& p, int b, int q_row, int head, int dh) {
  if (p.o_rows_per_rank > 0) {
    const int r = q_row / p.o_rows_per_rank;
    return p.o_peer[r] + (long long)(q_row - r * p.o_rows_per_rank) * p.o_rs + head * dh;
  }
  if (p.o_rows_per_rank < 0) {
    // segment-sharded rows (world = -o_rows_per_rank): every segment [b_s, b_{s+1}) is split evenly over the ranks and a
    // rank keeps its shards in segment order, so its local segment s starts at b_s / world
    const int world = -p.o_rows_per_rank;
    int sq = p.n_seg - 1;
    for (int s = 0; s < p.n_seg; ++s)
      if (q_row >= p.bounds[s] && q_row < p.bounds[s + 1]) sq = s;
    const int per = (p.bounds[sq + 1] - p.bounds[sq]) / world, off = q_row - p.bounds[sq];
    const int r = off / per;
    return p.o_peer[r] + (long long)(p.bounds[sq] / world + off - r * per) * p.o_rs + head * dh;
  }
  return p.o + (long long)b * p.o_bs + (long long)q_row * p.o_rs + head * dh;
}

template <int kDh, bool kPInTmem>
struct AttnCfg {
  static constexpr int SLABS = kDh / 64;                 // 64-column (128-byte) slabs per row
  static constexpr int SLAB_BYTES = 128 * 128;           // 128 rows x 128 B
  static constexpr int TILE_BYTES = SLABS * SLAB_BYTES;  // one Q / K / V tile
  static constexpr int KV_STAGES = 2;
  static constexpr int P_BYTES = kPInTmem ? 0 : 2 * 2 * SLAB_BYTES;  // 2 buffers x (128 keys = 2 slabs)
  static constexpr int SMEM_TILES = TILE_BYTES * (1 + 2 * KV_STAGES) + P_BYTES;
  static constexpr int NUM_BARS = 1 + 4 * KV_STAGES + 6;
  static constexpr int SMEM_BYTES = SMEM_TILES + NUM_BARS * 8 + 16 + kMaxTiles * 2 + 1024;
  static constexpr int TMEM_COLS = 512;
  static constexpr int TMEM_S0 = 0, TMEM_S1 = 128, TMEM_O = 256;
  static_assert(SMEM_BYTES <= 227 * 1024, "shared memory budget");
};

// Is key tile `kt` needed by query tile `qt` (any visible (query segment, key segment) pair)?  flags bit0 = needed,
// bit1 = per-element masking required (some pair invisible, or the tile holds the sequence tail).
__device__ __forceinline__ int classify_tile(const AttnParams& p, int qt, int kt, int block_q = kBlockQ) {
  const int q_lo = qt * block_q, q_hi = min(q_lo + block_q, p.seq);
  const int k_lo = kt * kBlockKV, k_hi_full = k_lo + kBlockKV, k_hi = min(k_hi_full, p.seq);
  int needed = 0, partial = (k_hi_full > p.seq) ? 1 : 0;
  if (p.n_seg == 0) return 1 | (partial << 1);
  for (int sq = 0; sq < p.n_seg; ++sq) {
    if (p.bounds[sq + 1] <= q_lo || p.bounds[sq] >= q_hi) continue;
    for (int sk = 0; sk < p.n_seg; ++sk) {
      if (p.bounds[sk + 1] <= k_lo || p.bounds[sk] >= k_hi) continue;
      if ((p.visible[sq] >> sk) & 1u) needed = 1; else partial = 1;
    }
  }
  return needed | (partial << 1);
}

// Key tiles a query tile has to visit, built by all 32 lanes of one warp (lane l classifies tiles l, l + 32, ...; a ballot
// + prefix count compacts them in order). A single thread doing this serially cost tens of microseconds per CTA at 132 key
// tiles x 5 segments — a large fraction of a short (condition-row) CTA's lifetime.
__device__ __forceinline__ int build_tile_list(const AttnParams& p, int qt, int block_q, uint16_t* tile_list, int lane) {
  const int total = (p.seq + kBlockKV - 1) / kBlockKV;
  int n = 0;
  for (int k0 = 0; k0 < total; k0 += 32) {
    const int kt = k0 + lane;
    const int f = kt < total ? classify_tile(p, qt, kt, block_q) : 0;
    const unsigned int m = __ballot_sync(0xffffffffu, f & 1);
    if (f & 1) tile_list[n + __popc(m & ((1u << lane) - 1u))] = (uint16_t)(kt | ((f >> 1) << 15));
    n += __popc(m);
  }
  return n;
}

template <int kDh, bool kPInTmem>
__global__ void __launch_bounds__(192, 1)
attention_kernel(const __grid_constant__ CUtensorMap tma_q, const __grid_constant__ CUtensorMap tma_k,
                 const __grid_constant__ CUtensorMap tma_v, const AttnParams p) {
  using Cfg = AttnCfg<kDh, kPInTmem>;
  constexpr int KS = Cfg::KV_STAGES;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
  uint8_t* smem_q = smem;
  uint8_t* smem_k = smem_q + Cfg::TILE_BYTES;
  uint8_t* smem_v = smem_k + KS * Cfg::TILE_BYTES;
  uint8_t* smem_p = smem_v + KS * Cfg::TILE_BYTES;
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + Cfg::SMEM_TILES);
  uint64_t* q_full = bars;
  uint64_t* k_full = q_full + 1;
  uint64_t* k_empty = k_full + KS;
  uint64_t* v_full = k_empty + KS;
  uint64_t* v_empty = v_full + KS;
  uint64_t* s_full = v_empty + KS;   // [2]
  uint64_t* p_ready = s_full + 2;    // [2]
  uint64_t* pv_done = p_ready + 2;   // [2]
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(pv_done + 2);
  int* n_tiles_smem = reinterpret_cast<int*>(tmem_slot + 1);
  uint16_t* tile_list = reinterpret_cast<uint16_t*>(tmem_slot + 4);  // entry = tile | (needs_mask << 15)

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int qt = p.head_fastest ? blockIdx.y : blockIdx.x, head = p.head_fastest ? blockIdx.x : blockIdx.y, b = blockIdx.z;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tma_q);
    tma_prefetch_desc(&tma_k);
    tma_prefetch_desc(&tma_v);
    mbar_init(q_full, 1);
    for (int s = 0; s < KS; ++s) {
      mbar_init(&k_full[s], 1); mbar_init(&k_empty[s], 1);
      mbar_init(&v_full[s], 1); mbar_init(&v_empty[s], 1);
    }
    for (int s = 0; s < 2; ++s) {
      mbar_init(&s_full[s], 1);
      mbar_init(&p_ready[s], 128);
      mbar_init(&pv_done[s], 1);
    }
    fence_mbar_init();
  }
  if (warp == 0) {
    const int n = build_tile_list(p, qt, kBlockQ, tile_list, lane);
    if (lane == 0) *n_tiles_smem = n;
  }
  if (warp == 1) tmem_alloc<1>(tmem_slot, Cfg::TMEM_COLS);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  const int n_tiles = *n_tiles_smem;

  if (warp == 0) {
    // ------------------------------ TMA producer ------------------------------
    if (elect_one_sync()) {
      mbar_arrive_expect_tx(q_full, Cfg::TILE_BYTES);
#pragma unroll
      for (int s = 0; s < Cfg::SLABS; ++s)
        tma_load_4d(smem_q + s * Cfg::SLAB_BYTES, &tma_q, q_full, s * 64, head, qt * kBlockQ, b);
    }
    int stage = 0;
    uint32_t phase = 0;
    for (int i = 0; i < n_tiles; ++i) {
      const int kt = tile_list[i] & 0x7fff;
      mbar_wait(&k_empty[stage], phase ^ 1);
      if (elect_one_sync()) {
        mbar_arrive_expect_tx(&k_full[stage], Cfg::TILE_BYTES);
#pragma unroll
        for (int s = 0; s < Cfg::SLABS; ++s)
          tma_load_4d(smem_k + stage * Cfg::TILE_BYTES + s * Cfg::SLAB_BYTES, &tma_k, &k_full[stage], s * 64, head,
                      kt * kBlockKV, b);
      }
      mbar_wait(&v_empty[stage], phase ^ 1);
      if (elect_one_sync()) {
        mbar_arrive_expect_tx(&v_full[stage], Cfg::TILE_BYTES);
#pragma unroll
        for (int s = 0; s < Cfg::SLABS; ++s)
          tma_load_4d(smem_v + stage * Cfg::TILE_BYTES + s * Cfg::SLAB_BYTES, &tma_v, &v_full[stage], s * 64, head,
                      kt * kBlockKV, b);
      }
      __syncwarp();
      if (++stage == KS) { stage = 0; phase ^= 1; }
    }
  } else if (warp == 1) {
    // ------------------------------ MMA issuer ------------------------------
    constexpr uint32_t idesc_s = make_idesc_bf16(128, kBlockKV, false, false);  // S = Q K^T : 128 x 128, K = dh
    constexpr uint32_t idesc_o = make_idesc_bf16(128, kDh, false, true);        // O = P V   : 128 x dh, V is MN-major
    auto issue_s = [&](int stage, int buf) {
      const uint32_t d = tmem_base + (buf ? Cfg::TMEM_S1 : Cfg::TMEM_S0);
#pragma unroll
      for (int k = 0; k < kDh / 16; ++k) {
        const uint32_t off = (k >> 2) * Cfg::SLAB_BYTES + (k & 3) * 32;
        const uint64_t a_desc = make_sdesc_sw128(smem_u32(smem_q) + off, 16, 1024);
        const uint64_t b_desc = make_sdesc_sw128(smem_u32(smem_k + stage * Cfg::TILE_BYTES) + off, 16, 1024);
        umma_ss<1>(d, a_desc, b_desc, idesc_s, k != 0 ? 1u : 0u);
      }
    };
    auto issue_pv = [&](int stage, int buf, bool accumulate) {
      const uint32_t d = tmem_base + Cfg::TMEM_O;
#pragma unroll
      for (int k = 0; k < kBlockKV / 16; ++k) {
        // V tile: SLABS slabs of [128 keys][64 dh]; 16 keys = 2048 B down the slab; LBO = slab stride (next 64 dh)
        const uint64_t b_desc =
            make_sdesc_sw128(smem_u32(smem_v + stage * Cfg::TILE_BYTES) + k * 2048, Cfg::SLAB_BYTES, 1024);
        const uint32_t acc = (accumulate || k != 0) ? 1u : 0u;
        if constexpr (kPInTmem) {
          // P (bf16 pairs) aliases the first 64 columns of its S buffer; 16 keys = 8 columns per MMA
          const uint32_t a_tmem = tmem_base + (buf ? Cfg::TMEM_S1 : Cfg::TMEM_S0) + k * 8;
          umma_ts(d, a_tmem, b_desc, idesc_o, acc);
        } else {
          const uint32_t off = buf * 2 * Cfg::SLAB_BYTES + (k >> 2) * Cfg::SLAB_BYTES + (k & 3) * 32;
          const uint64_t a_desc = make_sdesc_sw128(smem_u32(smem_p) + off, 16, 1024);
          umma_ss<1>(d, a_desc, b_desc, idesc_o, acc);
        }
      }
    };
    if (n_tiles > 0) {
      mbar_wait(q_full, 0);
      mbar_wait(&k_full[0], 0);
      tc_fence_after();
      if (elect_one_sync()) {
        issue_s(0, 0);
        umma_commit(&k_empty[0]);
        umma_commit(&s_full[0]);
      }
      __syncwarp();
      for (int i = 0; i < n_tiles; ++i) {
        const int stage = i % KS;
        const uint32_t phase = (i / KS) & 1;
        if (i + 1 < n_tiles) {
          const int nstage = (i + 1) % KS;
          mbar_wait(&k_full[nstage], ((i + 1) / KS) & 1);
          tc_fence_after();
          if (elect_one_sync()) {
            issue_s(nstage, (i + 1) & 1);
            umma_commit(&k_empty[nstage]);
            umma_commit(&s_full[(i + 1) & 1]);
          }
          __syncwarp();
        }
        mbar_wait(&p_ready[i & 1], (i >> 1) & 1);
        mbar_wait(&v_full[stage], phase);
        tc_fence_after();
        if (elect_one_sync()) {
          issue_pv(stage, i & 1, i > 0);
          umma_commit(&v_empty[stage]);
          umma_commit(&pv_done[i & 1]);
        }
        __syncwarp();
      }
    }
  } else {
    // ------------------------------ softmax + epilogue (warps 2..5) ------------------------------
    const int qd = warp & 3;                 // TMEM lane quarter owned by this warp
    const int row_local = qd * 32 + lane;    // query row inside the tile
    const int q_row = qt * kBlockQ + row_local;
    const uint32_t lane_off = (uint32_t)(qd * 32) << 16;
    unsigned int vis = 0xffffffffu;
    if (p.n_seg > 0) {
      int sq = p.n_seg - 1;
      for (int s = 0; s < p.n_seg; ++s)
        if (q_row >= p.bounds[s] && q_row < p.bounds[s + 1]) sq = s;
      vis = p.visible[sq];
    }
    float m = -INFINITY, l = 0.f;
    for (int i = 0; i < n_tiles; ++i) {
      const int entry = tile_list[i];
      const int kt = entry & 0x7fff;
      const int buf = i & 1;
      const uint32_t s_addr = tmem_base + lane_off + (buf ? Cfg::TMEM_S1 : Cfg::TMEM_S0);
      mbar_wait(&s_full[buf], (i >> 1) & 1);
      tc_fence_after();
      uint32_t s[128];
#pragma unroll
      for (int c = 0; c < 4; ++c) tmem_ld_32x32(s_addr + 32 * c, *reinterpret_cast<uint32_t(*)[32]>(&s[32 * c]));
      tmem_ld_wait();
      if (entry & 0x8000) {
        // per-row column mask: invisible key segments and the sequence tail
        unsigned int mw[4] = {0u, 0u, 0u, 0u};
        const int k_lo = kt * kBlockKV;
        auto mask_range = [&](int lo, int hi) {  // columns [lo, hi) of this tile
          lo = max(lo, 0); hi = min(hi, kBlockKV);
          for (int w = 0; w < 4; ++w) {
            const int a = max(lo - 32 * w, 0), e = min(hi - 32 * w, 32);
            if (a < e) mw[w] |= (e - a == 32) ? 0xffffffffu : (((1u << (e - a)) - 1u) << a);
          }
        };
        if (p.seq < k_lo + kBlockKV) mask_range(p.seq - k_lo, kBlockKV);
        for (int sk = 0; sk < p.n_seg; ++sk)
          if (!((vis >> sk) & 1u)) mask_range(p.bounds[sk] - k_lo, p.bounds[sk + 1] - k_lo);
#pragma unroll
        for (int c = 0; c < 128; ++c)
          if ((mw[c >> 5] >> (c & 31)) & 1u) s[c] = 0xff800000u;  // -inf
      }
      // 8 independent max chains (3-input FMNMX3), then a short tree: the row reduction is latency-, not issue-bound
      float rm[8];
#pragma unroll
      for (int c = 0; c < 8; ++c) rm[c] = fmaxf(__uint_as_float(s[c]), __uint_as_float(s[c + 8]));
#pragma unroll
      for (int c = 16; c < 128; c += 16) {
#pragma unroll
        for (int j = 0; j < 8; ++j) rm[j] = fmaxf(rm[j], fmaxf(__uint_as_float(s[c + j]), __uint_as_float(s[c + j + 8])));
      }
      const float rmax = fmaxf(fmaxf(fmaxf(rm[0], rm[1]), fmaxf(rm[2], rm[3])), fmaxf(fmaxf(rm[4], rm[5]), fmaxf(rm[6], rm[7])));
      // Lazy rescaling: the running reference max m only moves when some row of this warp would otherwise see
      // exp2 arguments above kRescaleLog2 (P values up to 2^8 are harmless in bf16 / fp32); O and l stay consistent
      // with the reference max, so the result is exact. Most tiles then skip the TMEM round trip of O.
      const float m_cand = fmaxf(m, rmax);
      const bool grow = (m_cand - m) * p.scale_log2 > kRescaleLog2;  // also true when m == -inf and rmax is finite
      const bool rescale = __any_sync(0xffffffffu, grow);
      const float m_new = rescale ? m_cand : m;
      const float m_use = (m_new == -INFINITY) ? 0.f : m_new;
      const float alpha = rescale ? ex2_ftz((m - m_use) * p.scale_log2) : 1.f;  // m = -inf -> 0
      const float neg_ms = -m_use * p.scale_log2;
      // packed fp32x2 scale-subtract and row sums (FFMA2 / FADD2): see attention2_kernel
      float rs[8];
      uint64_t rs2[4] = {0ull, 0ull, 0ull, 0ull};
      const uint64_t sc2 = pack_f32x2(p.scale_log2, p.scale_log2), nm2 = pack_f32x2(neg_ms, neg_ms);
      uint32_t pk[64];
#pragma unroll
      for (int c = 0; c < 64; ++c) {
        float a0, a1;
        unpack_f32x2(fma_f32x2(pack_f32x2(__uint_as_float(s[2 * c]), __uint_as_float(s[2 * c + 1])), sc2, nm2), a0, a1);
        const float p0 = ex2_ftz(a0), p1 = ex2_ftz(a1);
        rs2[c & 3] = add_f32x2(rs2[c & 3], pack_f32x2(p0, p1));
        pk[c] = pack_bf16x2(p0, p1);
      }
#pragma unroll
      for (int j = 0; j < 4; ++j) unpack_f32x2(rs2[j], rs[2 * j], rs[2 * j + 1]);
      l = l * alpha + (((rs[0] + rs[1]) + (rs[2] + rs[3])) + ((rs[4] + rs[5]) + (rs[6] + rs[7])));
      if constexpr (kPInTmem) {
        tmem_st_32x32(s_addr, *reinterpret_cast<const uint32_t(*)[32]>(&pk[0]));
        tmem_st_32x32(s_addr + 32, *reinterpret_cast<const uint32_t(*)[32]>(&pk[32]));
      } else {
        // K-major, 128B-swizzled: row r, 16-byte chunk c of slab (keys 64*slab ..): chunk index XOR (r & 7)
        uint8_t* pb = smem_p + buf * 2 * Cfg::SLAB_BYTES + row_local * 128;
#pragma unroll
        for (int ch = 0; ch < 16; ++ch) {
          uint8_t* dst = pb + (ch >> 3) * Cfg::SLAB_BYTES + (((ch & 7) ^ (row_local & 7)) << 4);
          *reinterpret_cast<uint4*>(dst) = make_uint4(pk[4 * ch], pk[4 * ch + 1], pk[4 * ch + 2], pk[4 * ch + 3]);
        }
      }
      if (i > 0) {
        // O must be quiescent (PV of the previous tile retired) before it is rescaled
        mbar_wait(&pv_done[(i - 1) & 1], ((i - 1) >> 1) & 1);
        tc_fence_after();
        if (rescale) {
          const uint32_t o_addr = tmem_base + lane_off + Cfg::TMEM_O;
#pragma unroll
          for (int c = 0; c < kDh / 32; ++c) {
            uint32_t o[32];
            tmem_ld_32x32(o_addr + 32 * c, o);
            tmem_ld_wait();
#pragma unroll
            for (int j = 0; j < 32; ++j) o[j] = __float_as_uint(__uint_as_float(o[j]) * alpha);
            tmem_st_32x32(o_addr + 32 * c, o);
          }
        }
      }
      if constexpr (kPInTmem) {
        tmem_st_wait();
      } else {
        tmem_st_wait();
        fence_proxy_async_smem();
      }
      tc_fence_before();
      mbar_arrive(&p_ready[buf]);
      m = m_new;
    }
    // epilogue: O / l -> bf16 -> HBM
    if (n_tiles > 0) {
      mbar_wait(&pv_done[(n_tiles - 1) & 1], ((n_tiles - 1) >> 1) & 1);
      tc_fence_after();
    }
    const float inv_l = l > 0.f ? 1.f / l : 0.f;
    __nv_bfloat16* orow = q_row < p.seq ? o_row_ptr(p, b, q_row, head, kDh) : p.o;
    const uint32_t o_addr = tmem_base + lane_off + Cfg::TMEM_O;
#pragma unroll
    for (int c = 0; c < kDh / 32; ++c) {
      uint32_t o[32];
      if (n_tiles > 0) {
        tmem_ld_32x32(o_addr + 32 * c, o);
        tmem_ld_wait();
      } else {
#pragma unroll
        for (int j = 0; j < 32; ++j) o[j] = 0u;
      }
      if (q_row < p.seq) {
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          uint4 u;
          u.x = pack_bf16x2(__uint_as_float(o[8 * j]) * inv_l, __uint_as_float(o[8 * j + 1]) * inv_l);
          u.y = pack_bf16x2(__uint_as_float(o[8 * j + 2]) * inv_l, __uint_as_float(o[8 * j + 3]) * inv_l);
          u.z = pack_bf16x2(__uint_as_float(o[8 * j + 4]) * inv_l, __uint_as_float(o[8 * j + 5]) * inv_l);
          u.w = pack_bf16x2(__uint_as_float(o[8 * j + 6]) * inv_l, __uint_as_float(o[8 * j + 7]) * inv_l);
          *reinterpret_cast<uint4*>(orow + 32 * c + 8 * j) = u;
        }
      }
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc<1>(tmem_base, Cfg::TMEM_COLS);
  }
}


// =====================================================================================================================
// v2: 256 query rows per CTA = two 128-row tiles that share every K/V tile and ping-pong on the tensor pipe.
// Warps: 0 TMA, 1 MMA, 2-3 idle (register donors), 4-7 softmax group 0 (rows 0..127), 8-11 softmax group 1 (rows 128..255).
// Each group owns S_w (128 TMEM columns, P aliases its first 64) and O_w (head_dim columns). The MMA warp issues
//   PV0(j), S0(j+1), PV1(j), S1(j+1)
// so group 0's softmax overlaps group 1's MMAs and vice versa; K/V smem traffic per FLOP halves vs v1.
// s_full[w](j) is committed after S_w(j), i.e. after PV_w(j-1) in issue order, so it also tells group w that O_w is
// quiescent and may be rescaled.
// =====================================================================================================================
template <int kDh>
struct Attn2Cfg {
  static constexpr int SLABS = kDh / 64;
  static constexpr int SLAB_BYTES = 128 * 128;
  static constexpr int TILE_BYTES = SLABS * SLAB_BYTES;
  static constexpr int KV_STAGES = 2;
  static constexpr int SMEM_TILES = TILE_BYTES * (2 + 2 * KV_STAGES);
  static constexpr int NUM_BARS = 1 + 4 * KV_STAGES + 8 + 5;   // + q_empty, list_full[2], list_empty[2] (persistent mode)
  static constexpr int SMEM_BYTES = SMEM_TILES + NUM_BARS * 8 + 16 + 2 * kMaxTiles * 2 + 1024;
  static constexpr int TMEM_COLS = 512;
  static_assert(SMEM_BYTES <= 227 * 1024, "shared memory budget");
};

// kSplitP: the softmax groups publish P in two 64-key halves (p_ready / p_ready2) and the MMA warp issues the first four PV
// MMAs as soon as the first half is in TMEM, so half of the PV work overlaps the exponentials of the second half instead of
// waiting behind them (the per-group chain softmax -> PV -> S(next) -> softmax is what bounds this kernel: ncu shows tensor
// pipe and MUFU ~53 % active each, i.e. mostly taking turns).
// kPersist: grid = min(units, SMs); every CTA walks the work units u = blockIdx.x, + gridDim.x, ... (unit = one 256-row query
// tile of one head). The K / V ring, the MMA issue stream and the barrier phases simply continue across units: the producer
// prefetches the next unit's Q (as soon as the last S MMAs of the current one have retired) and its first K / V tiles, the
// MMA warp issues S(0) of the next unit behind the last PV of the current one, so the O epilogue of one unit overlaps the
// first MMAs of the next, and TMEM allocation / barrier init / CTA launch are paid once. Per-unit key-tile lists are built
// one unit ahead by warp 2 into a double buffer. (Cycle stamps of the one-unit-per-CTA form at S = 4608: ~9 of 83 us per
// CTA were prologue, epilogue, the tail where one group idles, and the relaunch gap.)
template <int kDh, int kPolyMod, bool kSplitP, bool kPersist>
__global__ void __launch_bounds__(384, 1)
attention2_kernel(const __grid_constant__ CUtensorMap tma_q, const __grid_constant__ CUtensorMap tma_k,
                  const __grid_constant__ CUtensorMap tma_v, const AttnParams p) {
  using Cfg = Attn2Cfg<kDh>;
  constexpr int KS = Cfg::KV_STAGES;
  constexpr int kBlockQ2 = 256;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
  uint8_t* smem_q = smem;                                // 2 query tiles
  uint8_t* smem_k = smem_q + 2 * Cfg::TILE_BYTES;
  uint8_t* smem_v = smem_k + KS * Cfg::TILE_BYTES;
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + Cfg::SMEM_TILES);
  uint64_t* q_full = bars;
  uint64_t* k_full = q_full + 1;
  uint64_t* k_empty = k_full + KS;
  uint64_t* v_full = k_empty + KS;
  uint64_t* v_empty = v_full + KS;
  uint64_t* s_full = v_empty + KS;   // [2] per softmax group
  uint64_t* p_ready = s_full + 2;    // [2]
  uint64_t* pv_done = p_ready + 2;   // [2]
  uint64_t* p_ready2 = pv_done + 2;  // [2] second half of P (kSplitP)
  uint64_t* q_empty = p_ready2 + 2;    // persistent mode: the unit's last S MMAs retired, Q may be overwritten
  uint64_t* list_full = q_empty + 1;   // [2] tile list of unit k is in buffer k & 1
  uint64_t* list_empty = list_full + 2;  // [2] every consumer is done with it (10 arrivals: producer, MMA, 8 softmax warps)
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(list_empty + 2);
  int* n_tiles_smem = reinterpret_cast<int*>(tmem_slot + 1);  // [2]
  uint16_t* tile_lists = reinterpret_cast<uint16_t*>(tmem_slot + 4);  // [2][kMaxTiles]

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int n_units = kPersist ? p.n_units : 1;
  const int u_first = kPersist ? (int)blockIdx.x : 0, u_step = kPersist ? (int)gridDim.x : 1;
  // work unit -> (256-row query tile, head, batch); same order as the one-unit-per-CTA grid
  auto decode = [&](int u, int& qt, int& head, int& b) {
    if constexpr (kPersist) {
      const int fast = p.head_fastest ? p.heads : p.q_tiles2, slow = p.head_fastest ? p.q_tiles2 : p.heads;
      const int x = u % fast, y = (u / fast) % slow;
      b = u / (fast * slow);
      qt = p.head_fastest ? y : x;
      head = p.head_fastest ? x : y;
    } else {
      qt = p.head_fastest ? blockIdx.y : blockIdx.x;
      head = p.head_fastest ? blockIdx.x : blockIdx.y;
      b = blockIdx.z;
    }
  };
  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tma_q);
    tma_prefetch_desc(&tma_k);
    tma_prefetch_desc(&tma_v);
    mbar_init(q_full, 1);
    for (int s = 0; s < KS; ++s) {
      mbar_init(&k_full[s], 1); mbar_init(&k_empty[s], 1);
      mbar_init(&v_full[s], 1); mbar_init(&v_empty[s], 1);
    }
    for (int s = 0; s < 2; ++s) {
      mbar_init(&s_full[s], 1);
      mbar_init(&p_ready[s], 128);
      mbar_init(&p_ready2[s], 128);
      mbar_init(&pv_done[s], 1);
      mbar_init(&list_full[s], 1);
      mbar_init(&list_empty[s], 10);
    }
    mbar_init(q_empty, 1);
    fence_mbar_init();
  }
  if constexpr (!kPersist) {
    if (warp == 0) {
      int qt, head, b;
      decode(0, qt, head, b);
      const int n = build_tile_list(p, qt, kBlockQ2, tile_lists, lane);
      if (lane == 0) n_tiles_smem[0] = n;
    }
  }
  if (warp == 1) tmem_alloc<1>(tmem_slot, Cfg::TMEM_COLS);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  // tile list of the k-th unit of this CTA (persistent: wait until warp 2 has built it)
  auto acquire_list = [&](int k, const uint16_t*& tile_list) -> int {
    if constexpr (kPersist) mbar_wait(&list_full[k & 1], (k >> 1) & 1);
    tile_list = tile_lists + (k & 1) * kMaxTiles;
    return n_tiles_smem[k & 1];
  };
  auto release_list = [&](int k) {  // by ONE lane per consumer warp
    if constexpr (kPersist) mbar_arrive(&list_empty[k & 1]);
  };

  if (warp < 4) {
setmaxnreg_dec<72>();  // 168*384 = 64512 registers per CTA: 72*128 + 208*256 = 62464 fits
    if (warp == 0) {
      // ------------------------------ TMA producer ------------------------------
      int stage = 0;
      uint32_t phase = 0;
      int k = 0;
      for (int u = u_first; u < n_units; u += u_step, ++k) {
        int qt, head, b;
        decode(u, qt, head, b);
        const uint16_t* tile_list;
        const int n_tiles = acquire_list(k, tile_list);
        if constexpr (kPersist) mbar_wait(q_empty, (k & 1) ^ 1);  // the previous unit's S MMAs have all read Q
        if (elect_one_sync()) {
          mbar_arrive_expect_tx(q_full, 2 * Cfg::TILE_BYTES);
#pragma unroll
          for (int t = 0; t < 2; ++t)
#pragma unroll
            for (int s = 0; s < Cfg::SLABS; ++s)
              tma_load_4d(smem_q + t * Cfg::TILE_BYTES + s * Cfg::SLAB_BYTES, &tma_q, q_full, s * 64, head,
                          qt * kBlockQ2 + t * 128, b);
        }
        for (int i = 0; i < n_tiles; ++i) {
          const int kt = tile_list[i] & 0x7fff;
          mbar_wait(&k_empty[stage], phase ^ 1);
          if (elect_one_sync()) {
            mbar_arrive_expect_tx(&k_full[stage], Cfg::TILE_BYTES);
#pragma unroll
            for (int s = 0; s < Cfg::SLABS; ++s)
              tma_load_4d(smem_k + stage * Cfg::TILE_BYTES + s * Cfg::SLAB_BYTES, &tma_k, &k_full[stage], s * 64, head,
                          kt * kBlockKV, b);
          }
          mbar_wait(&v_empty[stage], phase ^ 1);
          if (elect_one_sync()) {
            mbar_arrive_expect_tx(&v_full[stage], Cfg::TILE_BYTES);
#pragma unroll
            for (int s = 0; s < Cfg::SLABS; ++s)
              tma_load_4d(smem_v + stage * Cfg::TILE_BYTES + s * Cfg::SLAB_BYTES, &tma_v, &v_full[stage], s * 64, head,
                          kt * kBlockKV, b);
          }
          __syncwarp();
          if (++stage == KS) { stage = 0; phase ^= 1; }
        }
        if (elect_one_sync()) release_list(k);
        __syncwarp();
      }
    } else if (warp == 1) {
      // ------------------------------ MMA issuer ------------------------------
      constexpr uint32_t idesc_s = make_idesc_bf16(128, kBlockKV, false, false);
      constexpr uint32_t idesc_o = make_idesc_bf16(128, kDh, false, true);
      // Descriptors are built ONCE: per MMA only the 14-bit start-address field moves (+2 per 32 bytes along K inside a
      // 128-byte swizzle row, + a slab / 2 KB step otherwise), so each tcgen05.mma costs the issuing thread one add per
      // operand instead of re-encoding the descriptor (the issue thread, not the tensor pipe, was the bottleneck:
      // ~85 cycles of scalar work per 64-cycle 128x128x16 MMA).
      const uint32_t tb = __shfl_sync(0xffffffffu, tmem_base, 0);
      const uint64_t q_desc0 = make_sdesc_sw128(smem_u32(smem_q), 16, 1024);
      const uint64_t k_desc0 = make_sdesc_sw128(smem_u32(smem_k), 16, 1024);
      const uint64_t v_desc0 = make_sdesc_sw128(smem_u32(smem_v), Cfg::SLAB_BYTES, 1024);
      // (the empty asm keeps ptxas from hoisting the 8 + 8 per-MMA addresses of every call site out of the tile loop: it
      // would park ~50 loop-invariant values in registers, spill them, and reload them on the issue path)
      auto issue_s = [&](int w, int stage) {
        uint32_t tbl = tb;
        asm volatile("" : "+r"(tbl));
        const uint32_t d = tbl + w * 128;
        const uint64_t a0 = q_desc0 + (uint64_t)((w * Cfg::TILE_BYTES) >> 4);
        const uint64_t b0 = k_desc0 + (uint64_t)((stage * Cfg::TILE_BYTES) >> 4);
#pragma unroll
        for (int k = 0; k < kDh / 16; ++k) {
          const uint32_t off = ((k >> 2) * Cfg::SLAB_BYTES + (k & 3) * 32) >> 4;
          umma_ss<1>(d, a0 + off, b0 + off, idesc_s, k != 0 ? 1u : 0u);
        }
      };
      auto issue_pv = [&](int w, int stage, bool accumulate, int k0 = 0, int k1 = kBlockKV / 16) {
        uint32_t tbl = tb;
        asm volatile("" : "+r"(tbl));
        const uint32_t d = tbl + 256 + w * 128;
        const uint64_t b0 = v_desc0 + (uint64_t)((stage * Cfg::TILE_BYTES) >> 4);
#pragma unroll
        for (int k = 0; k < kBlockKV / 16; ++k) {
          if (k < k0 || k >= k1) continue;
          umma_ts(d, tbl + w * 128 + k * 8, b0 + (uint64_t)((k * 2048) >> 4), idesc_o, (accumulate || k != 0) ? 1u : 0u);
        }
      };
      // One continuous stream of tiles over all units of this CTA. g counts tiles (barrier parities), the K / V ring runs on.
      int stage = 0, nstage = 1 % KS;
      uint32_t phase = 0, nphase = (1 / KS) & 1;
      uint32_t g = 0;
      int k = 0;
      const uint16_t* unused_list;
      int nt = u_first < n_units ? __shfl_sync(0xffffffffu, acquire_list(0, unused_list), 0) : 0;
      // S of the first tile of the first unit
      if (nt > 0) {
        mbar_wait(q_full, 0);
        mbar_wait(&k_full[0], 0);
        tc_fence_after();
        if (elect_one_sync()) {
          issue_s(0, 0);
          umma_commit(&s_full[0]);
          issue_s(1, 0);
          umma_commit(&s_full[1]);
          umma_commit(&k_empty[0]);
          if (kPersist && nt == 1) umma_commit(q_empty);
        }
        __syncwarp();
      }
      for (int u = u_first; u < n_units; u += u_step, ++k) {
        const bool next_unit = kPersist && (u + u_step < n_units);
        int nt_next = 0;
        for (int i = 0; i < nt; ++i, ++g) {
          const bool last = i + 1 == nt;
          // `more`: there is a next tile whose S is issued behind this tile's PV (the first tile of the next unit included)
          const bool more = !last || next_unit;
          // K / V arrive long before P: poll them first so that nothing but the issue itself follows the wake-up on p_ready
          // (measured 225 cycles between "P seen" and the first MMA the other way round). At a unit boundary the next unit's
          // S(0) is issued AFTER both PVs of this unit's last tile (`cross`): its Q is still on its way (it could only be
          // requested once this unit's last S MMAs had retired), and this unit's epilogue must not wait for it.
          const bool cross = last && next_unit;
          const bool s_now = more && !cross;
          mbar_wait(&v_full[stage], phase);
          if (s_now) mbar_wait(&k_full[nstage], nphase);
          // the next tile is the last one of its unit: once its S MMAs have been issued, Q is no longer needed
          const bool q_done = kPersist && s_now && i + 2 == nt;
          // ---- group 0 ----
          mbar_wait(&p_ready[0], g & 1);
          tc_fence_after();
          if constexpr (kSplitP) {
            if (elect_one_sync()) issue_pv(0, stage, i > 0, 0, kBlockKV / 32);
            __syncwarp();
            mbar_wait(&p_ready2[0], g & 1);
            tc_fence_after();
          }
          if (elect_one_sync()) {
            if constexpr (kSplitP) issue_pv(0, stage, true, kBlockKV / 32, kBlockKV / 16);
            else issue_pv(0, stage, i > 0);
            umma_commit(&pv_done[0]);
            if (s_now) {
              issue_s(0, nstage);
              umma_commit(&s_full[0]);
            }
          }
          __syncwarp();
          // ---- group 1 ----
          mbar_wait(&p_ready[1], g & 1);
          tc_fence_after();
          if constexpr (kSplitP) {
            if (elect_one_sync()) issue_pv(1, stage, i > 0, 0, kBlockKV / 32);
            __syncwarp();
            mbar_wait(&p_ready2[1], g & 1);
            tc_fence_after();
          }
          if (elect_one_sync()) {
            if constexpr (kSplitP) issue_pv(1, stage, true, kBlockKV / 32, kBlockKV / 16);
            else issue_pv(1, stage, i > 0);
            umma_commit(&pv_done[1]);
            umma_commit(&v_empty[stage]);
            if (s_now) {
              issue_s(1, nstage);
              umma_commit(&s_full[1]);
              umma_commit(&k_empty[nstage]);
              if (q_done) umma_commit(q_empty);
            }
          }
          __syncwarp();
          if (cross) {
            nt_next = __shfl_sync(0xffffffffu, acquire_list(k + 1, unused_list), 0);
            mbar_wait(q_full, (k + 1) & 1);
            mbar_wait(&k_full[nstage], nphase);
            tc_fence_after();
            if (elect_one_sync()) {
              issue_s(0, nstage);
              umma_commit(&s_full[0]);
              issue_s(1, nstage);
              umma_commit(&s_full[1]);
              umma_commit(&k_empty[nstage]);
              if (nt_next == 1) umma_commit(q_empty);
            }
            __syncwarp();
          }
          stage = nstage; phase = nphase;
          if (++nstage == KS) { nstage = 0; nphase ^= 1; }
        }
        if (elect_one_sync()) release_list(k);
        __syncwarp();
        nt = nt_next;
      }
    } else if (warp == 2) {
      // ------------------------------ tile lists, one unit ahead (persistent mode) ------------------------------
      if constexpr (kPersist) {
        int k = 0;
        for (int u = u_first; u < n_units; u += u_step, ++k) {
          int qt, head, b;
          decode(u, qt, head, b);
          mbar_wait(&list_empty[k & 1], ((k >> 1) & 1) ^ 1);
          const int n = build_tile_list(p, qt, kBlockQ2, tile_lists + (k & 1) * kMaxTiles, lane);
          if (lane == 0) n_tiles_smem[k & 1] = n;
          __syncwarp();
          if (lane == 0) mbar_arrive(&list_full[k & 1]);
        }
      }
    }
  } else {
    // ------------------------------ softmax groups ------------------------------
    setmaxnreg_inc<208>();
    const int w = (warp - 4) >> 2;            // group
    const int qd = warp & 3;                  // TMEM lane quarter
    constexpr int NC = 128;                   // score columns per thread (one query row)
    constexpr int NP = NC / 2;                // column pairs = packed P columns
    const int row_local = qd * 32 + lane;
    const uint32_t lane_off = (uint32_t)(qd * 32) << 16;
    const uint32_t s_addr = tmem_base + lane_off + w * 128;
    const uint32_t o_addr = tmem_base + lane_off + 256 + w * 128;
    uint32_t g = 0;  // tiles processed so far (barrier parities run on across units)
    int k = 0;
    for (int u = u_first; u < n_units; u += u_step, ++k) {
    int qt, head, b;
    decode(u, qt, head, b);
    const uint16_t* tile_list;
    const int n_tiles = acquire_list(k, tile_list);
    const int q_row = qt * kBlockQ2 + w * 128 + row_local;
    unsigned int vis = 0xffffffffu;
    if (p.n_seg > 0) {
      int sq = p.n_seg - 1;
      for (int s = 0; s < p.n_seg; ++s)
        if (q_row >= p.bounds[s] && q_row < p.bounds[s + 1]) sq = s;
      vis = p.visible[sq];
    }
    float m = -INFINITY, l = 0.f;
    for (int i = 0; i < n_tiles; ++i, ++g) {
      const int entry = tile_list[i];
      const int kt = entry & 0x7fff;
      // the column mask of a partially visible tile does not depend on S: build it while the S MMA is still in flight and
      // register pressure is low (with the score registers live, ptxas spilled a quarter of them around this branch)
      unsigned int mw[NC / 32];
#pragma unroll
      for (int ww = 0; ww < NC / 32; ++ww) mw[ww] = 0u;
      if (entry & 0x8000) {
        const int k_lo = kt * kBlockKV;
        auto mask_range = [&](int lo, int hi) {
          lo = max(lo, 0); hi = min(hi, NC);
#pragma unroll
          for (int ww = 0; ww < NC / 32; ++ww) {
            const int a = max(lo - 32 * ww, 0), e = min(hi - 32 * ww, 32);
            mw[ww] |= a < e ? ((e - a == 32) ? 0xffffffffu : (((1u << (e - a)) - 1u) << a)) : 0u;
          }
        };
        if (p.seq < k_lo + NC) mask_range(p.seq - k_lo, NC);
        for (int sk = 0; sk < p.n_seg; ++sk)
          if (!((vis >> sk) & 1u)) mask_range(p.bounds[sk] - k_lo, p.bounds[sk + 1] - k_lo);
      }
      mbar_wait(&s_full[w], g & 1);
      tc_fence_after();
      uint32_t s[NC];
#pragma unroll
      for (int c = 0; c < NC / 32; ++c) tmem_ld_32x32(s_addr + 32 * c, *reinterpret_cast<uint32_t(*)[32]>(&s[32 * c]));
      tmem_ld_wait();
      if (entry & 0x8000) {
#pragma unroll
        for (int c = 0; c < NC; ++c)
          if ((mw[c >> 5] >> (c & 31)) & 1u) s[c] = 0xff800000u;
      }
      // 8 independent max chains (3-input FMNMX3), then a short tree: the row reduction is latency-, not issue-bound
      float rm[8];
#pragma unroll
      for (int c = 0; c < 8; ++c) rm[c] = fmaxf(__uint_as_float(s[c]), __uint_as_float(s[c + 8]));
#pragma unroll
      for (int c = 16; c < NC; c += 16) {
#pragma unroll
        for (int j = 0; j < 8; ++j) rm[j] = fmaxf(rm[j], fmaxf(__uint_as_float(s[c + j]), __uint_as_float(s[c + j + 8])));
      }
      const float rmax = fmaxf(fmaxf(fmaxf(rm[0], rm[1]), fmaxf(rm[2], rm[3])), fmaxf(fmaxf(rm[4], rm[5]), fmaxf(rm[6], rm[7])));
      const float m_cand = fmaxf(m, rmax);
      const bool grow = (m_cand - m) * p.scale_log2 > kRescaleLog2;  // lazy rescaling, see attention_kernel
      const bool rescale = __any_sync(0xffffffffu, grow);
      const float m_new = rescale ? m_cand : m;
      const float m_use = (m_new == -INFINITY) ? 0.f : m_new;
      const float alpha = rescale ? ex2_ftz((m - m_use) * p.scale_log2) : 1.f;
      const float neg_ms = -m_use * p.scale_log2;
      // O_w is quiescent here (s_full(i) is committed after PV_w(i-1)): rescale it when the reference max moved
      // (8-column chunks: with the score registers live there is no room for 32-column round trips; ptxas spilled a
      // quarter of the scores on the hot path to make room for this rare branch)
      if (i > 0 && rescale) {
#pragma unroll 1
        for (int c = 0; c < kDh / 8; ++c) {
          uint32_t o[8];
          tmem_ld_32x8(o_addr + 8 * c, o);
          tmem_ld_wait();
#pragma unroll
          for (int j = 0; j < 8; ++j) o[j] = __float_as_uint(__uint_as_float(o[j]) * alpha);
          tmem_st_32x8(o_addr + 8 * c, o);
        }
      }
      // The scale-subtract and the row sums run as packed fp32x2 instructions (FFMA2 / FADD2: the same IEEE operations, half
      // the issue slots).
      float rs[8];
      uint64_t rs2[4] = {0ull, 0ull, 0ull, 0ull};
      const uint64_t sc2 = pack_f32x2(p.scale_log2, p.scale_log2), nm2 = pack_f32x2(neg_ms, neg_ms);
      {
        uint32_t pk[NP];
#pragma unroll
        for (int c = 0; c < NP; ++c) {
          float p0, p1;
          const uint64_t a2 = fma_f32x2(pack_f32x2(__uint_as_float(s[2 * c]), __uint_as_float(s[2 * c + 1])), sc2, nm2);
          if (kPolyMod > 0 && (c % kPolyMod) == kPolyMod - 1) {
            // every kPolyMod-th PAIR takes the polynomial (FMA-pipe) exp2 on both lanes instead of two MUFU.EX2
            ex2_poly_x2(a2, p0, p1);
          } else {
            float a0, a1;
            unpack_f32x2(a2, a0, a1);
            p0 = ex2_ftz(a0);
            p1 = ex2_ftz(a1);
          }
          rs2[c & 3] = add_f32x2(rs2[c & 3], pack_f32x2(p0, p1));
          pk[c] = pack_bf16x2(p0, p1);
          if (c % 32 == 31) {
            tmem_st_32x32(s_addr + (c - 31), *reinterpret_cast<const uint32_t(*)[32]>(&pk[c - 31]));
            if constexpr (kSplitP) {  // publish the first 64 keys of P: the MMA warp starts PV on them meanwhile
              if (c == 31) {
                tmem_st_wait();
                tc_fence_before();
                mbar_arrive(&p_ready[w]);
              }
            }
          }
        }
      }
#pragma unroll
      for (int j = 0; j < 4; ++j) unpack_f32x2(rs2[j], rs[2 * j], rs[2 * j + 1]);
      l = l * alpha + (((rs[0] + rs[1]) + (rs[2] + rs[3])) + ((rs[4] + rs[5]) + (rs[6] + rs[7])));
      tmem_st_wait();
      tc_fence_before();
      mbar_arrive(kSplitP ? &p_ready2[w] : &p_ready[w]);
      m = m_new;
    }
    if (n_tiles > 0) {
      mbar_wait(&pv_done[w], (g - 1) & 1);
      tc_fence_after();
    }
    const float inv_l = l > 0.f ? 1.f / l : 0.f;
    __nv_bfloat16* orow = q_row < p.seq ? o_row_ptr(p, b, q_row, head, kDh) : p.o;
#pragma unroll
    for (int c = 0; c < kDh / 32; ++c) {
      uint32_t o[32];
      if (n_tiles > 0) {
        tmem_ld_32x32(o_addr + 32 * c, o);
        tmem_ld_wait();
      } else {
#pragma unroll
        for (int j = 0; j < 32; ++j) o[j] = 0u;
      }
      if (q_row < p.seq) {
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          uint4 u;
          u.x = pack_bf16x2(__uint_as_float(o[8 * j]) * inv_l, __uint_as_float(o[8 * j + 1]) * inv_l);
          u.y = pack_bf16x2(__uint_as_float(o[8 * j + 2]) * inv_l, __uint_as_float(o[8 * j + 3]) * inv_l);
          u.z = pack_bf16x2(__uint_as_float(o[8 * j + 4]) * inv_l, __uint_as_float(o[8 * j + 5]) * inv_l);
          u.w = pack_bf16x2(__uint_as_float(o[8 * j + 6]) * inv_l, __uint_as_float(o[8 * j + 7]) * inv_l);
          *reinterpret_cast<uint4*>(orow + 32 * c + 8 * j) = u;
        }
      }
    }
    __syncwarp();
    if (lane == 0) release_list(k);
    }  // units
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc<1>(tmem_base, Cfg::TMEM_COLS);
  }
}


__global__ void expand_mask_kernel(int seq, int n_seg, AttnParams p, uint8_t* __restrict__ mask) {
  const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= (long long)seq * seq) return;
  const int q = (int)(idx / seq), k = (int)(idx % seq);
  int sq = 0, sk = 0;
  for (int s = 0; s < n_seg; ++s) {
    if (q >= p.bounds[s] && q < p.bounds[s + 1]) sq = s;
    if (k >= p.bounds[s] && k < p.bounds[s + 1]) sk = s;
  }
  mask[idx] = n_seg == 0 ? 1 : (uint8_t)((p.visible[sq] >> sk) & 1u);
}

static int fill_segments(AttnParams& p, int seq, int n_seg, const int32_t* bounds, const uint32_t* visible) {
  UG_CHECK_ARG(n_seg >= 0 && n_seg <= UG_MAX_SEGMENTS, "attention: n_seg %d out of range [0, %d]", n_seg, UG_MAX_SEGMENTS);
  p.n_seg = n_seg;
  for (int i = 0; i <= UG_MAX_SEGMENTS; ++i) p.bounds[i] = seq;
  for (int i = 0; i < UG_MAX_SEGMENTS; ++i) p.visible[i] = 0;
  if (n_seg > 0) {
    UG_CHECK_ARG(bounds && visible, "attention: segment arrays are null");
    UG_CHECK_ARG(bounds[0] == 0 && bounds[n_seg] == seq, "attention: seg_bounds must start at 0 and end at seq");
    for (int i = 0; i < n_seg; ++i) {
      UG_CHECK_ARG(bounds[i + 1] >= bounds[i], "attention: seg_bounds must be non-decreasing");
      p.bounds[i] = bounds[i];
      p.visible[i] = visible[i];
    }
    p.bounds[n_seg] = seq;
  }
  return UG_OK;
}

struct PeerO {
  int rows_per_rank;
  __nv_bfloat16* base[UG_MAX_PEERS];
};
static void set_peer(AttnParams& p, const PeerO* peer) {
  p.o_rows_per_rank = peer ? peer->rows_per_rank : 0;
  for (int i = 0; i < UG_MAX_PEERS; ++i) p.o_peer[i] = peer ? peer->base[i] : nullptr;
}

template <int kDh, bool kPInTmem>
static int launch_attention(const ug_attn_args& a, const PeerO* peer, cudaStream_t stream) {
  using Cfg = AttnCfg<kDh, kPInTmem>;
  auto kern = attention_kernel<kDh, kPInTmem>;
  static bool attr_done[64] = {false};
  if (int st = ensure_dynamic_smem(kern, Cfg::SMEM_BYTES, attr_done, "attention"); st != UG_OK) return st;
  CUtensorMap maps[3];
  const void* ptrs[3] = {a.q, a.k, a.v};
  const int64_t rs[3] = {a.q_row_stride, a.k_row_stride, a.v_row_stride};
  const int64_t bs[3] = {a.q_batch_stride, a.k_batch_stride, a.v_batch_stride};
  for (int i = 0; i < 3; ++i) {
    uint64_t dims[4] = {(uint64_t)kDh, (uint64_t)a.heads, (uint64_t)a.seq, (uint64_t)a.batch};
    uint64_t bstride = a.batch > 1 ? (uint64_t)bs[i] : (uint64_t)a.seq * rs[i];
    uint64_t strides[3] = {(uint64_t)kDh * 2, (uint64_t)rs[i] * 2, bstride * 2};
    uint32_t box[4] = {64, 1, 128, 1};
    int st = encode_tmap_bf16(&maps[i], ptrs[i], 4, dims, strides, box);
    if (st != UG_OK) return st;
  }
  AttnParams p;
  p.o = (__nv_bfloat16*)a.o; p.o_rs = a.o_row_stride; p.o_bs = a.o_batch_stride;
  p.seq = a.seq; p.heads = a.heads; p.batch = a.batch;
  set_peer(p, peer);
  p.scale_log2 = a.scale * 1.4426950408889634f;
  int st = fill_segments(p, a.seq, a.n_seg, a.seg_bounds, a.seg_visible);
  if (st != UG_OK) return st;
  const int q_tiles = (a.seq + kBlockQ - 1) / kBlockQ;
  p.head_fastest = ((long long)q_tiles * a.heads * a.batch < 4LL * num_sms()) ? 1 : 0;
  dim3 grid(p.head_fastest ? a.heads : q_tiles, p.head_fastest ? q_tiles : a.heads, a.batch);
  kern<<<grid, 192, Cfg::SMEM_BYTES, stream>>>(maps[0], maps[1], maps[2], p);
  UG_CHECK_LAUNCH("attention");
  return UG_OK;
}


template <int kDh, int kPolyMod, bool kSplitP = false, bool kPersist = false>
static int launch_attention2(const ug_attn_args& a, const PeerO* peer, cudaStream_t stream) {
  using Cfg = Attn2Cfg<kDh>;
  auto kern = attention2_kernel<kDh, kPolyMod, kSplitP, kPersist>;
  static bool attr_done[64] = {false};
  if (int st = ensure_dynamic_smem(kern, Cfg::SMEM_BYTES, attr_done, "attention2"); st != UG_OK) return st;
  CUtensorMap maps[3];
  const void* ptrs[3] = {a.q, a.k, a.v};
  const int64_t rs[3] = {a.q_row_stride, a.k_row_stride, a.v_row_stride};
  const int64_t bs[3] = {a.q_batch_stride, a.k_batch_stride, a.v_batch_stride};
  for (int i = 0; i < 3; ++i) {
    uint64_t dims[4] = {(uint64_t)kDh, (uint64_t)a.heads, (uint64_t)a.seq, (uint64_t)a.batch};
    uint64_t bstride = a.batch > 1 ? (uint64_t)bs[i] : (uint64_t)a.seq * rs[i];
    uint64_t strides[3] = {(uint64_t)kDh * 2, (uint64_t)rs[i] * 2, bstride * 2};
    uint32_t box[4] = {64, 1, 128, 1};
    int st = encode_tmap_bf16(&maps[i], ptrs[i], 4, dims, strides, box);
    if (st != UG_OK) return st;
  }
  AttnParams p;
  p.o = (__nv_bfloat16*)a.o; p.o_rs = a.o_row_stride; p.o_bs = a.o_batch_stride;
  p.seq = a.seq; p.heads = a.heads; p.batch = a.batch;
  set_peer(p, peer);
  p.scale_log2 = a.scale * 1.4426950408889634f;
  int st = fill_segments(p, a.seq, a.n_seg, a.seg_bounds, a.seg_visible);
  if (st != UG_OK) return st;
  const int q_tiles = (a.seq + 255) / 256;
  p.head_fastest = ((long long)q_tiles * a.heads * a.batch < 4LL * num_sms()) ? 1 : 0;
  p.q_tiles2 = q_tiles;
  p.n_units = q_tiles * a.heads * a.batch;
  if constexpr (kPersist) {
    // the persistent stream issues S(0) of a unit behind the last PV of the previous one: every unit needs >= 1 key tile,
    // i.e. every non-empty segment must see a non-empty segment (always true for the reference's visibility rules)
    for (int sq = 0; sq < p.n_seg; ++sq) {
      if (p.bounds[sq + 1] <= p.bounds[sq]) continue;
      bool sees = false;
      for (int sk = 0; sk < p.n_seg; ++sk) sees |= ((p.visible[sq] >> sk) & 1u) && p.bounds[sk + 1] > p.bounds[sk];
      if (!sees) return launch_attention2<kDh, kPolyMod, kSplitP, false>(a, peer, stream);
    }
    kern<<<dim3(std::min(p.n_units, num_sms())), 384, Cfg::SMEM_BYTES, stream>>>(maps[0], maps[1], maps[2], p);
  } else {
    dim3 grid(p.head_fastest ? a.heads : q_tiles, p.head_fastest ? q_tiles : a.heads, a.batch);
    kern<<<grid, 384, Cfg::SMEM_BYTES, stream>>>(maps[0], maps[1], maps[2], p);
  }
  UG_CHECK_LAUNCH("attention2");
  return UG_OK;
}

}  // namespace ug

using namespace ug;

static int attention_dispatch(const ug_attn_args* args, const PeerO* peer, void* stream) {
  UG_CHECK_ARG(args != nullptr, "attention: null args");
  const ug_attn_args& a = *args;
  UG_CHECK_ARG(a.q && a.k && a.v && (a.o || peer), "attention: null operand pointer");
  UG_CHECK_ARG(a.batch >= 1 && a.heads >= 1 && a.seq >= 1, "attention: empty problem");
  UG_CHECK_ARG(a.seq <= kMaxTiles * kBlockKV, "attention: seq %d exceeds %d", a.seq, kMaxTiles * kBlockKV);
  UG_CHECK_ARG(a.q_row_stride % 8 == 0 && a.k_row_stride % 8 == 0 && a.v_row_stride % 8 == 0 && a.o_row_stride % 8 == 0,
               "attention: row strides must be multiples of 8 elements");
  UG_CHECK_ARG((reinterpret_cast<uintptr_t>(a.o) & 15) == 0, "attention: O not 16-byte aligned");
  UG_CHECK_ARG(a.batch == 1 || (a.q_batch_stride % 8 == 0 && a.k_batch_stride % 8 == 0 && a.v_batch_stride % 8 == 0 &&
                                a.o_batch_stride % 8 == 0),
               "attention: batch strides must be multiples of 8 elements");
  cudaStream_t s = reinterpret_cast<cudaStream_t>(stream);
  // auto: the two-tile ping-pong kernel once there are enough 256-row query tiles to fill the machine, with P published in two
  // halves (bit-identical to variant 3; -0.4 % per cfg3 step once the MMA issue stopped being the bottleneck)
  // (variants 9 / 10 = split-P + packed polynomial exp2 on every 3rd / 2nd pair: 9 is +5 % in isolation at the SD3.5 head_dim-64 shape
  // and +1 % at head_dim 128, but does not shorten the cfg5 step — 120.9 vs 119.3 ms, same box — so the MUFU path stays the default)
  int variant = a.variant;
  if (variant == 0) variant = ((long long)((a.seq + 255) / 256) * a.heads * a.batch >= num_sms()) ? 5 : 1;
  if (a.head_dim == 128) {
    if (variant == 1) return launch_attention<128, true>(a, peer, s);
    if (variant == 2) return launch_attention<128, false>(a, peer, s);
    if (variant == 3) return launch_attention2<128, 0>(a, peer, s);
    if (variant == 4) return launch_attention2<128, 4>(a, peer, s);
    if (variant == 5) return launch_attention2<128, 0, true>(a, peer, s);
    if (variant == 6) return launch_attention2<128, 3>(a, peer, s);
    if (variant == 7) return launch_attention2<128, 0, true, true>(a, peer, s);
    if (variant == 9) return launch_attention2<128, 3, true>(a, peer, s);
    if (variant == 10) return launch_attention2<128, 2, true>(a, peer, s);
  } else if (a.head_dim == 64) {
    if (variant == 1) return launch_attention<64, true>(a, peer, s);
    if (variant == 2) return launch_attention<64, false>(a, peer, s);
    if (variant == 3) return launch_attention2<64, 0>(a, peer, s);
    if (variant == 4) return launch_attention2<64, 4>(a, peer, s);
    if (variant == 5) return launch_attention2<64, 0, true>(a, peer, s);
    if (variant == 6) return launch_attention2<64, 3>(a, peer, s);
    if (variant == 7) return launch_attention2<64, 0, true, true>(a, peer, s);
    if (variant == 9) return launch_attention2<64, 3, true>(a, peer, s);
    if (variant == 10) return launch_attention2<64, 2, true>(a, peer, s);
  } else {
    set_error("attention: head_dim %d not supported (64 or 128)", a.head_dim);
    return UG_ERR_UNSUPPORTED;
  }
  set_error("attention: unknown variant %d", a.variant);
  return UG_ERR_INVALID;
}

extern "C" int ug_attention_bf16(const ug_attn_args* args, void* stream) { return attention_dispatch(args, nullptr, stream); }

extern "C" int ug_attention_bf16_peer(const ug_attn_args* args, const ug_peer_table* table, int64_t o_offset, int32_t rows_per_rank,
                                      void* stream) {
  UG_CHECK_ARG(args && table, "attention_peer: null args");
  UG_CHECK_ARG(table->world >= 1 && table->world <= UG_MAX_PEERS && table->rank >= 0 && table->rank < table->world,
               "attention_peer: bad world %d / rank %d", table->world, table->rank);
  UG_CHECK_ARG(args->batch == 1, "attention_peer: one sample is sharded across the ranks (batch must be 1)");
  if (rows_per_rank == 0) {  // segment-sharded rows
    UG_CHECK_ARG(args->n_seg >= 1 && args->seg_bounds, "attention_peer: rows_per_rank = 0 selects segment-sharded rows and needs segments");
    for (int i = 0; i <= args->n_seg; ++i)
      UG_CHECK_ARG(args->seg_bounds[i] % table->world == 0, "attention_peer: segment bound %d is not a multiple of the world size %d",
                   args->seg_bounds[i], table->world);
  } else {
    UG_CHECK_ARG(rows_per_rank >= 1 && (long long)rows_per_rank * table->world >= args->seq,
                 "attention_peer: %d rows per rank x %d ranks do not cover seq %d", rows_per_rank, table->world, args->seq);
  }
  UG_CHECK_ARG(o_offset >= UG_PEER_HEADER_BYTES && o_offset % 16 == 0, "attention_peer: bad output offset");
  PeerO peer;
  peer.rows_per_rank = rows_per_rank > 0 ? rows_per_rank : -table->world;
  // this rank's heads occupy columns [rank * heads * head_dim, (rank + 1) * heads * head_dim) of every rank's output rows
  const long long col0 = (long long)table->rank * args->heads * args->head_dim;
  for (int i = 0; i < UG_MAX_PEERS; ++i) {
    peer.base[i] = nullptr;
    if (i < table->world) {
      UG_CHECK_ARG(table->base[i], "attention_peer: null peer base %d", i);
      peer.base[i] = reinterpret_cast<__nv_bfloat16*>(reinterpret_cast<uint8_t*>(table->base[i]) + o_offset) + col0;
    }
  }
  return attention_dispatch(args, &peer, stream);
}

extern "C" int ug_expand_segment_mask(int32_t seq, int32_t n_seg, const int32_t* bounds, const uint32_t* visible,
                                      uint8_t* mask_dev, void* stream) {
  UG_CHECK_ARG(seq >= 1 && mask_dev, "expand_segment_mask: bad arguments");
  AttnParams p;
  int st = fill_segments(p, seq, n_seg, bounds, visible);
  if (st != UG_OK) return st;
  const long long total = (long long)seq * seq;
  expand_mask_kernel<<<(unsigned)((total + 255) / 256), 256, 0, reinterpret_cast<cudaStream_t>(stream)>>>(seq, n_seg, p, mask_dev);
  UG_CHECK_LAUNCH("expand_segment_mask");
  return UG_OK;
}
