// Whole-step handle of the C ABI: one denoise step of UniGenFlux (reference src/UniGenTransformer.py:1182-1271 with
// base_forward :1106-1180, control_forward :1070-1104, preprocess_moe_forward :1028-1068, moe_forward :969-1026) sequenced in
// C++ — a host that is not Python creates a handle, binds the weight pointers under the reference's state-dict names, sizes and
// supplies ONE workspace, and calls ug_flux_forward on its stream. Every arithmetic op is one of the granular entry points of
// this library (the same kernels, in the same order, as the Python mirror unigen_b200/model.py issues them: results are
// bit-identical, tests/test_chandle_gpu.py). No device memory is allocated after ug_flux_create; the only CUDA objects the
// handle owns are two side streams and a few events (fork / join of the AdaLN weight stream and the text-stream GEMMs).
#include <cuda_bf16.h>

#include <cmath>
#include <cstring>
#include <map>
#include <string>
#include <vector>

#include "ug_host.h"

namespace {

struct Tensor {
  const void* p = nullptr;
  int dtype = 0;  // 0 bf16, 1 f32
  int64_t shape[4] = {0, 0, 0, 0};
  int ndim = 0;
};

struct Lin { const void* w = nullptr; const void* b = nullptr; };

struct DoubleW {
  Lin norm1, norm1_ctx, qkv, add_qkv, to_out, to_add_out, ff1, ff2, ffc1, ffc2;
  const void* rms = nullptr;      // [norm_q; norm_k]
  const void* rms_ctx = nullptr;  // [norm_added_q; norm_added_k]
};
struct SingleW { Lin norm, qkv, mlp, out; const void* rms = nullptr; };
struct TimeTextW { Lin t1, t2, g1, g2, p1, p2; };

struct View {  // bf16 [batch, rows, cols] view (element strides)
  void* p; int64_t rs, bs; int rows;
};

constexpr size_t kAlign = 256;
inline size_t align_up(size_t x) { return (x + kAlign - 1) / kAlign * kAlign; }

}  // namespace

struct ug_flux {
  ug_flux_desc d;
  int D = 0;
  std::map<std::string, Tensor> weights;
  bool resolved = false;
  // resolved weight table
  Lin x_embedder, context_embedder, norm_out, proj_out, control_context_embedder, control_x_embedder;
  TimeTextW time_text, control_time_text, control_condition;
  std::vector<DoubleW> dbl, cdbl, shared;
  std::vector<SingleW> sgl, csgl;
  std::vector<Lin> add_dbl, add_sgl;
  const float* gate_wg = nullptr;
  const void *exp_w[2] = {nullptr, nullptr}, *exp_b[2] = {nullptr, nullptr}, *exp_mod_w[2] = {nullptr, nullptr}, *exp_mod_b[2] = {nullptr, nullptr};
  // streams / events
  cudaStream_t side = nullptr, text = nullptr;
  cudaEvent_t ev_fork = nullptr, ev_mod = nullptr, ev_t0 = nullptr, ev_t1 = nullptr;
  // job tables cached per (workspace, shape)
  void* plan_ws = nullptr;
  int plan_B = 0, plan_N = 0, plan_T = 0;
  int early_jobs = 0, early_groups = 0, late_jobs = 0, late_groups = 0;
};

using namespace ug;

#define UG_TRY(expr)            \
  do {                          \
    int st__ = (expr);          \
    if (st__ != UG_OK) return st__; \
  } while (0)

extern "C" int ug_flux_create(const ug_flux_desc* desc, ug_flux** out) {
  UG_CHECK_ARG(desc && out, "flux_create: null argument");
  UG_CHECK_ARG(desc->num_layers >= 1 && desc->num_single_layers >= 0 && desc->heads >= 1 && (desc->head_dim == 64 || desc->head_dim == 128),
               "flux_create: bad architecture (head_dim must be 64 or 128)");
  UG_CHECK_ARG(desc->axes_dims_rope[0] + desc->axes_dims_rope[1] + desc->axes_dims_rope[2] == desc->head_dim,
               "flux_create: axes_dims_rope must sum to head_dim");
  UG_CHECK_ARG(desc->n_ctrl_double >= 1 && desc->n_ctrl_double <= desc->num_layers && desc->n_ctrl_single >= 0 &&
                   desc->n_ctrl_single <= desc->num_single_layers, "flux_create: control block counts out of range");
  UG_CHECK_ARG(desc->condition_nums >= 1 && desc->condition_nums <= UG_FLUX_MAX_CONDITIONS && desc->experts >= 1 && desc->experts <= 32,
               "flux_create: 1..%d conditions and 1..32 experts", UG_FLUX_MAX_CONDITIONS);
  ug_flux* h = new ug_flux();
  h->d = *desc;
  h->D = desc->heads * desc->head_dim;
  cudaError_t e = cudaStreamCreateWithFlags(&h->side, cudaStreamNonBlocking);
  if (e == cudaSuccess) e = cudaStreamCreateWithFlags(&h->text, cudaStreamNonBlocking);
  for (cudaEvent_t* ev : {&h->ev_fork, &h->ev_mod, &h->ev_t0, &h->ev_t1})
    if (e == cudaSuccess) e = cudaEventCreateWithFlags(ev, cudaEventDisableTiming);
  if (e != cudaSuccess) {
    set_error("flux_create: %s", cudaGetErrorString(e));
    delete h;
    return UG_ERR_CUDA;
  }
  *out = h;
  return UG_OK;
}

extern "C" void ug_flux_destroy(ug_flux* h) {
  if (!h) return;
  if (h->side) cudaStreamDestroy(h->side);
  if (h->text) cudaStreamDestroy(h->text);
  for (cudaEvent_t ev : {h->ev_fork, h->ev_mod, h->ev_t0, h->ev_t1})
    if (ev) cudaEventDestroy(ev);
  delete h;
}

extern "C" int ug_flux_bind_weight(ug_flux* h, const char* name, const void* dev_ptr, int32_t dtype, const int64_t* shape, int32_t ndim) {
  UG_CHECK_ARG(h && name && dev_ptr && shape && ndim >= 1 && ndim <= 4 && (dtype == 0 || dtype == 1), "flux_bind_weight: bad argument");
  Tensor t;
  t.p = dev_ptr; t.dtype = dtype; t.ndim = ndim;
  for (int i = 0; i < ndim; ++i) t.shape[i] = shape[i];
  h->weights[name] = t;
  h->resolved = false;
  h->plan_ws = nullptr;
  return UG_OK;
}

namespace {

struct Resolver {
  ug_flux* h;
  int status = UG_OK;
  const Tensor* get(const std::string& name, int dtype, int64_t s0, int64_t s1 = -1) {
    auto it = h->weights.find(name);
    if (it == h->weights.end()) { fail("weight %s is not bound", name); return nullptr; }
    const Tensor& t = it->second;
    if (t.dtype != dtype) { fail("weight %s has the wrong dtype", name); return nullptr; }
    if (t.shape[0] != s0 || (s1 >= 0 && (t.ndim < 2 || t.shape[1] != s1)) || (s1 < 0 && t.ndim != 1)) {
      fail("weight %s has the wrong shape", name);
      return nullptr;
    }
    return &t;
  }
  void fail(const char* fmt, const std::string& name) {
    if (status == UG_OK) { set_error(fmt, name.c_str()); status = UG_ERR_INVALID; }
  }
  Lin lin(const std::string& p, int64_t out_f, int64_t in_f) {
    Lin l;
    const Tensor* w = get(p + ".weight", 0, out_f, in_f);
    const Tensor* b = get(p + ".bias", 0, out_f);
    if (w && b) { l.w = w->p; l.b = b->p; }
    return l;
  }
  // `names` are row blocks of ONE fused [len * out_each, in_f] matrix: they must be bound to contiguous memory
  Lin fused(const std::vector<std::string>& names, int64_t out_each, int64_t in_f) {
    Lin first = lin(names[0], out_each, in_f);
    for (size_t i = 1; i < names.size() && status == UG_OK; ++i) {
      Lin l = lin(names[i], out_each, in_f);
      if (status != UG_OK) break;
      const char* want_w = static_cast<const char*>(first.w) + i * out_each * in_f * 2;
      const char* want_b = static_cast<const char*>(first.b) + i * out_each * 2;
      if (l.w != want_w || l.b != want_b) fail("%s: q / k / v (stacked expert) weights must be bound as one contiguous block", names[i]);
    }
    return first;
  }
  const void* rms_pair(const std::string& a, const std::string& b, int64_t dh) {
    const Tensor* ta = get(a, 0, dh);
    const Tensor* tb = get(b, 0, dh);
    if (!ta || !tb) return nullptr;
    if (static_cast<const char*>(tb->p) != static_cast<const char*>(ta->p) + dh * 2) fail("%s: norm_q / norm_k weights must be contiguous", b);
    return ta->p;
  }
  DoubleW dbl(const std::string& p, int64_t D, int64_t dh) {
    DoubleW w;
    w.norm1 = lin(p + ".norm1.linear", 6 * D, D);
    w.norm1_ctx = lin(p + ".norm1_context.linear", 6 * D, D);
    w.qkv = fused({p + ".attn.to_q", p + ".attn.to_k", p + ".attn.to_v"}, D, D);
    w.add_qkv = fused({p + ".attn.add_q_proj", p + ".attn.add_k_proj", p + ".attn.add_v_proj"}, D, D);
    w.to_out = lin(p + ".attn.to_out.0", D, D);
    w.to_add_out = lin(p + ".attn.to_add_out", D, D);
    w.rms = rms_pair(p + ".attn.norm_q.weight", p + ".attn.norm_k.weight", dh);
    w.rms_ctx = rms_pair(p + ".attn.norm_added_q.weight", p + ".attn.norm_added_k.weight", dh);
    w.ff1 = lin(p + ".ff.net.0.proj", 4 * D, D);
    w.ff2 = lin(p + ".ff.net.2", D, 4 * D);
    w.ffc1 = lin(p + ".ff_context.net.0.proj", 4 * D, D);
    w.ffc2 = lin(p + ".ff_context.net.2", D, 4 * D);
    return w;
  }
  SingleW sgl(const std::string& p, int64_t D, int64_t dh) {
    SingleW w;
    w.norm = lin(p + ".norm.linear", 3 * D, D);
    w.qkv = fused({p + ".attn.to_q", p + ".attn.to_k", p + ".attn.to_v"}, D, D);
    w.mlp = lin(p + ".proj_mlp", 4 * D, D);
    w.out = lin(p + ".proj_out", D, 5 * D);
    w.rms = rms_pair(p + ".attn.norm_q.weight", p + ".attn.norm_k.weight", dh);
    return w;
  }
  TimeTextW time_text(const std::string& p, int64_t D, int64_t pooled, bool guidance) {
    TimeTextW w;
    w.t1 = lin(p + ".timestep_embedder.linear_1", D, 256);
    w.t2 = lin(p + ".timestep_embedder.linear_2", D, D);
    if (guidance) {
      w.g1 = lin(p + ".guidance_embedder.linear_1", D, 256);
      w.g2 = lin(p + ".guidance_embedder.linear_2", D, D);
    }
    w.p1 = lin(p + ".text_embedder.linear_1", D, pooled);
    w.p2 = lin(p + ".text_embedder.linear_2", D, D);
    return w;
  }
};

int resolve(ug_flux* h) {
  if (h->resolved) return UG_OK;
  const ug_flux_desc& d = h->d;
  const int64_t D = h->D, dh = d.head_dim, P = d.pooled_dim, E = d.experts;
  Resolver r{h};
  auto idx = [](const char* p, int i) { return std::string(p) + "." + std::to_string(i); };
  h->x_embedder = r.lin("x_embedder", D, d.in_channels);
  h->context_embedder = r.lin("context_embedder", D, d.joint_dim);
  h->time_text = r.time_text("time_text_embed", D, P, d.guidance_embeds != 0);
  h->dbl.clear(); h->sgl.clear(); h->cdbl.clear(); h->csgl.clear(); h->add_dbl.clear(); h->add_sgl.clear(); h->shared.clear();
  for (int i = 0; i < d.num_layers; ++i) h->dbl.push_back(r.dbl(idx("transformer_blocks", i), D, dh));
  for (int i = 0; i < d.num_single_layers; ++i) h->sgl.push_back(r.sgl(idx("single_transformer_blocks", i), D, dh));
  h->norm_out = r.lin("norm_out.linear", 2 * D, D);
  h->proj_out = r.lin("proj_out", d.in_channels, D);
  h->control_time_text = r.time_text("control_time_text_embed", D, P, d.guidance_embeds != 0);
  h->control_condition = r.time_text("control_condition_embed", D, P, d.guidance_embeds != 0);
  h->control_context_embedder = r.lin("control_context_embedder", D, D);
  h->control_x_embedder = r.lin("control_x_embedder", D, d.in_channels);
  for (int j = 0; j < d.n_ctrl_double; ++j) {
    h->cdbl.push_back(r.dbl(idx("control_joint_trans_blocks", j), D, dh));
    h->add_dbl.push_back(r.lin(idx("controlnet_add_joint_blocks", j), D, D));
  }
  for (int j = 0; j < d.n_ctrl_single; ++j) {
    h->csgl.push_back(r.sgl(idx("control_single_trans_blocks", j), D, dh));
    h->add_sgl.push_back(r.lin(idx("controlnet_add_single_blocks", j), D, D));
  }
  if (const Tensor* g = r.get("moe.moe_layer.gate.wg.weight", 1, E, D)) h->gate_wg = static_cast<const float*>(g->p);
  for (int br = 0; br < 2; ++br) {
    std::vector<std::string> lin0, lin1;
    for (int e = 0; e < E; ++e) {
      const std::string p = "moe.moe_layer.experts.deepspeed_experts." + std::to_string(e) + "." + std::to_string(br);
      lin0.push_back(p + ".0");
      lin1.push_back(p + ".1");
    }
    Lin a = r.fused(lin0, D, D), m = r.fused(lin1, D, P);
    h->exp_w[br] = a.w; h->exp_b[br] = a.b; h->exp_mod_w[br] = m.w; h->exp_mod_b[br] = m.b;
  }
  if (d.use_shared_expert)
    for (int s = 0; s < 2; ++s) h->shared.push_back(r.dbl(idx("shared_expert", s), D, dh));
  if (r.status != UG_OK) return r.status;
  h->resolved = true;
  return UG_OK;
}

// ---- workspace layout (bytes, every buffer 256-byte aligned) ----
struct Layout {
  size_t X, NX, QKV, AO, FF, CAT, CH, CS, CENC, COND, HC, G, A, YC, YH, EH, EC, CIN, MOD, TEMBS, STEMBS, tmp, t_emb, g_emb, MODC, MODH,
      rope, rope0, rope1, NO, route_idx, route_slot, route_prob, route_slot_token, route_ws, jobs, total;
  int n_mod, capacity;
};

Layout make_layout(const ug_flux* h, int B, int N, int T) {
  const ug_flux_desc& d = h->d;
  const size_t D = h->D, S = T + N, Smax = T + 2 * (size_t)N, E = d.experts;
  Layout L;
  size_t off = 0;
  auto take = [&](size_t bytes) { size_t o = off; off = align_up(off + bytes); return o; };
  L.capacity = (int)((B * (size_t)N + E - 1) / E);
  if (L.capacity < 4) L.capacity = 4;
  const size_t C = L.capacity;
  L.n_mod = 12 * (d.num_layers + d.n_ctrl_double + 1 + d.condition_nums) + 3 * (d.num_single_layers + d.n_ctrl_single) + 2;
  L.X = take(B * S * D * 2); L.NX = take(B * Smax * D * 2); L.QKV = take(B * Smax * 3 * D * 2); L.AO = take(B * Smax * D * 2);
  L.FF = take(B * Smax * 4 * D * 2); L.CAT = take(B * S * 5 * D * 2); L.CH = take(B * (size_t)N * D * 2); L.CS = take(B * S * D * 2);
  L.CENC = take(B * (size_t)T * D * 2); L.COND = take(B * (size_t)N * D * 2); L.HC = take(B * 2 * (size_t)N * D * 2);
  L.G = take(B * (size_t)N * D * 2); L.A = take(E * C * D * 2); L.YC = take(E * C * D * 2); L.YH = take(E * C * D * 2);
  L.EH = take(B * (size_t)N * D * 2); L.EC = take(B * (size_t)N * D * 2); L.CIN = take(B * (size_t)N * D * 2);
  L.MOD = take((size_t)B * L.n_mod * D * 4);
  L.TEMBS = take((3 + d.condition_nums) * (size_t)B * D * 4); L.STEMBS = take((3 + d.condition_nums) * (size_t)B * D * 4);
  L.tmp = take((size_t)B * D * 4); L.t_emb = take((size_t)B * 256 * 4); L.g_emb = take((size_t)B * 256 * 4);
  L.MODC = take((size_t)B * E * D * 4); L.MODH = take((size_t)B * E * D * 4);
  L.rope = take(S * d.head_dim * 4); L.rope0 = take(2 * (size_t)N * d.head_dim * 4); L.rope1 = take(Smax * d.head_dim * 4);
  L.NO = take(B * (size_t)N * D * 2);
  L.route_idx = take((size_t)B * N * 4); L.route_slot = take((size_t)B * N * 4); L.route_prob = take((size_t)B * N * 4);
  L.route_slot_token = take(E * C * 4); L.route_ws = take(((size_t)B * N * E + E) * 4);
  L.jobs = take(UG_MAX_GEMV_JOBS * sizeof(ug_gemv_job));
  L.total = off;
  return L;
}

}  // namespace

extern "C" size_t ug_flux_workspace_bytes(const ug_flux* h, int32_t batch, int32_t n_img, int32_t n_txt) {
  if (!h || batch < 1 || n_img < 1 || n_txt < 1) return 0;
  return make_layout(h, batch, n_img, n_txt).total;
}

namespace {

struct Step {
  ug_flux* h;
  const ug_flux_inputs* in;
  const ug_flux_outputs* out;
  char* ws;
  Layout L;
  cudaStream_t main;
  int B, N, T, S, D, H, dh, E;
  float cscale;
  // mod table: chunk c of slot s for batch row b lives at MOD[b * n_mod * D + (s + c) * D]
  float* mod(int slot) const { return reinterpret_cast<float*>(ws + L.MOD) + (size_t)slot * D; }
  int64_t mod_bs() const { return (int64_t)L.n_mod * D; }
  __nv_bfloat16* bf(size_t off) const { return reinterpret_cast<__nv_bfloat16*>(ws + off); }
  float* f32(size_t off) const { return reinterpret_cast<float*>(ws + off); }
  View view(size_t off, int row0, int rows, int64_t rs, int64_t total_rows) const {
    return View{bf(off) + (int64_t)row0 * rs, rs, total_rows * rs, rows};
  }

  int gemm(cudaStream_t s, View a, int k, Lin w, int n, View c, const float* gate = nullptr, const View* res = nullptr, int act = UG_ACT_NONE,
           float alpha = 1.0f, int batch = -1, int64_t w_batch_stride = 0, int64_t bias_batch_stride = 0, const void* qk_rms = nullptr,
           const float* qk_cos_sin = nullptr) const {
    ug_gemm_args g;
    memset(&g, 0, sizeof(g));
    g.a = a.p; g.a_row_stride = a.rs; g.a_batch_stride = a.bs;
    g.w = w.w; g.w_row_stride = k; g.w_batch_stride = w_batch_stride;
    g.c = c.p; g.c_row_stride = c.rs; g.c_batch_stride = c.bs;
    g.batch = batch < 0 ? B : batch; g.rows = a.rows; g.n = n; g.k = k;
    g.bias = w.b; g.bias_batch_stride = bias_batch_stride;
    g.gate = gate; g.gate_batch_stride = gate ? mod_bs() : 0;
    g.alpha = alpha; g.act = act;
    if (res) { g.residual = res->p; g.res_row_stride = res->rs; g.res_batch_stride = res->bs; }
    if (qk_rms) {  // q|k|v projection: per-head RMSNorm + RoPE in the GEMM epilogue (model.py: fuse_qk_norm, the default)
      g.qk_norm_weight = qk_rms; g.qk_cos_sin = qk_cos_sin; g.qk_head_dim = dh; g.qk_d = D; g.qk_eps = 1e-6f;
    }
    return ug_gemm_bf16(&g, s);
  }
  int ln(cudaStream_t s, View x, View o, const float* shift, const float* scale) const {
    return ug_ln_modulate(x.p, x.rs, x.bs, o.p, o.rs, o.bs, shift, scale, mod_bs(), B, x.rows, D, 1e-6f, s);
  }
  int gemv(const float* x, Lin w, float* o, int n, int k, int silu_out, int accumulate, int64_t x_stride, int64_t o_stride) const {
    return ug_gemv(x, x_stride, w.w, w.b, o, o_stride, B, n, k, 0, silu_out, accumulate, main);
  }
  int time_text(const TimeTextW& w, const float* pooled, int64_t pooled_stride, float* dst, bool accumulate) const {
    float* tmp = f32(L.tmp);
    UG_TRY(gemv(f32(L.t_emb), w.t1, tmp, D, 256, 1, 0, 256, D));
    UG_TRY(gemv(tmp, w.t2, dst, D, D, 0, accumulate ? 1 : 0, D, D));
    if (in->guidance && w.g1.w) {
      UG_TRY(gemv(f32(L.g_emb), w.g1, tmp, D, 256, 1, 0, 256, D));
      UG_TRY(gemv(tmp, w.g2, dst, D, D, 0, 1, D, D));
    }
    UG_TRY(gemv(pooled, w.p1, tmp, D, h->d.pooled_dim, 1, 0, pooled_stride, D));
    return gemv(tmp, w.p2, dst, D, D, 0, 1, D, D);
  }
  int rope(size_t off, std::initializer_list<std::pair<const float*, int>> tables) const {
    int r0 = 0;
    for (auto& t : tables) {
      UG_TRY(ug_rope_table(t.first, t.second, h->d.axes_dims_rope, h->d.theta, f32(off) + (size_t)r0 * dh, main));
      r0 += t.second;
    }
    return UG_OK;
  }
  int attention(int Sq, View o) const {
    ug_attn_args a;
    memset(&a, 0, sizeof(a));
    __nv_bfloat16* qkv = bf(L.QKV);
    const int64_t rs = 3 * (int64_t)D, bs = (int64_t)(T + 2 * N) * rs;
    a.q = qkv; a.k = qkv + D; a.v = qkv + 2 * D; a.o = o.p;
    a.q_row_stride = a.k_row_stride = a.v_row_stride = rs;
    a.q_batch_stride = a.k_batch_stride = a.v_batch_stride = bs;
    a.o_row_stride = o.rs; a.o_batch_stride = o.bs;
    a.batch = B; a.heads = H; a.seq = Sq; a.head_dim = dh;
    a.scale = (float)(1.0 / sqrt((double)dh));  // the double-precision 1 / sqrt(head_dim) of the reference, rounded once
    return ug_attention_bf16(&a, main);
  }
  // diffusers FluxTransformerBlock over [ctx (n_ctx rows) | smp (n_smp rows)] (model.py::_double_block)
  int double_block(const DoubleW& w, int slot_smp, int slot_ctx, View smp_in, View ctx_in, View smp_out, const View* ctx_out, size_t rope_off) const {
    const int n_ctx = ctx_in.rows, n_smp = smp_in.rows, Sq = n_ctx + n_smp;
    const int64_t Smax = T + 2 * (int64_t)N;
    View nx_c = view(L.NX, 0, n_ctx, D, Smax), nx_s = view(L.NX, n_ctx, n_smp, D, Smax);
    View qkv_c = view(L.QKV, 0, n_ctx, 3 * D, Smax), qkv_s = view(L.QKV, n_ctx, n_smp, 3 * D, Smax);
    View ff_c = view(L.FF, 0, n_ctx, 4 * D, Smax), ff_s = view(L.FF, n_ctx, n_smp, 4 * D, Smax);
    View ao_c = view(L.AO, 0, n_ctx, D, Smax), ao_s = view(L.AO, n_ctx, n_smp, D, Smax), ao = view(L.AO, 0, Sq, D, Smax);
    const float *sh_a = mod(slot_smp), *sc_a = mod(slot_smp + 1), *g_a = mod(slot_smp + 2), *sh_m = mod(slot_smp + 3), *sc_m = mod(slot_smp + 4),
                *g_m = mod(slot_smp + 5);
    const float *csh_a = mod(slot_ctx), *csc_a = mod(slot_ctx + 1), *cg_a = mod(slot_ctx + 2), *csh_m = mod(slot_ctx + 3),
                *csc_m = mod(slot_ctx + 4), *cg_m = mod(slot_ctx + 5);
    const bool fork = ctx_out != nullptr;  // text-stream GEMMs beside the image stream's (disjoint rows)
    cudaStream_t ts = fork ? h->text : main;
    if (fork) { cudaEventRecord(h->ev_t0, main); cudaStreamWaitEvent(ts, h->ev_t0, 0); }
    const float* rp = f32(rope_off);
    UG_TRY(ln(ts, ctx_in, nx_c, csh_a, csc_a));
    UG_TRY(gemm(ts, nx_c, D, w.add_qkv, 3 * D, qkv_c, nullptr, nullptr, UG_ACT_NONE, 1.0f, -1, 0, 0, w.rms_ctx, rp));
    UG_TRY(ln(main, smp_in, nx_s, sh_a, sc_a));
    UG_TRY(gemm(main, nx_s, D, w.qkv, 3 * D, qkv_s, nullptr, nullptr, UG_ACT_NONE, 1.0f, -1, 0, 0, w.rms, rp + (size_t)n_ctx * dh));
    if (fork) { cudaEventRecord(h->ev_t1, ts); cudaStreamWaitEvent(main, h->ev_t1, 0); }
    UG_TRY(attention(Sq, ao));
    if (fork) {
      cudaEventRecord(h->ev_t0, main); cudaStreamWaitEvent(ts, h->ev_t0, 0);
      UG_TRY(gemm(ts, ao_c, D, w.to_add_out, D, *ctx_out, cg_a, &ctx_in));
      UG_TRY(ln(ts, *ctx_out, nx_c, csh_m, csc_m));
      UG_TRY(gemm(ts, nx_c, D, w.ffc1, 4 * D, ff_c, nullptr, nullptr, UG_ACT_GELU_TANH));
      UG_TRY(gemm(ts, ff_c, 4 * D, w.ffc2, D, *ctx_out, cg_m, ctx_out));
    }
    UG_TRY(gemm(main, ao_s, D, w.to_out, D, smp_out, g_a, &smp_in));
    UG_TRY(ln(main, smp_out, nx_s, sh_m, sc_m));
    UG_TRY(gemm(main, nx_s, D, w.ff1, 4 * D, ff_s, nullptr, nullptr, UG_ACT_GELU_TANH));
    UG_TRY(gemm(main, ff_s, 4 * D, w.ff2, D, smp_out, g_m, &smp_out));
    if (fork) { cudaEventRecord(h->ev_t1, ts); cudaStreamWaitEvent(main, h->ev_t1, 0); }
    return UG_OK;
  }

  // diffusers FluxSingleTransformerBlock (model.py::_single_block)
  int single_block(const SingleW& w, int slot, View x_in, View x_out) const {
    const int64_t Smax = T + 2 * (int64_t)N;
    View nx = view(L.NX, 0, S, D, Smax), qkv = view(L.QKV, 0, S, 3 * D, Smax);
    View cat = view(L.CAT, 0, S, 5 * D, S);
    View cat_attn = cat, cat_mlp = cat;
    cat_mlp.p = bf(L.CAT) + D;
    UG_TRY(ln(main, x_in, nx, mod(slot), mod(slot + 1)));
    UG_TRY(gemm(main, nx, D, w.qkv, 3 * D, qkv, nullptr, nullptr, UG_ACT_NONE, 1.0f, -1, 0, 0, w.rms, f32(L.rope)));
    UG_TRY(gemm(main, nx, D, w.mlp, 4 * D, cat_mlp, nullptr, nullptr, UG_ACT_GELU_TANH));
    UG_TRY(attention(S, cat_attn));
    return gemm(main, cat, 5 * D, w.out, D, x_out, mod(slot + 2), &x_in);
  }

  int add(View a, View b, View o) const {
    return ug_add_bf16(a.p, a.rs, a.bs, b.p, b.rs, b.bs, o.p, o.rs, o.bs, B, a.rows, D, main);
  }
};

struct JobBuilder {
  std::vector<ug_gemv_job> early, late;
  int g_early = 0, g_late = 0;
  void push(bool is_early, Lin w, const float* x, float* out, int n, int k, int64_t out_stride) {
    ug_gemv_job j;
    memset(&j, 0, sizeof(j));
    j.w = w.w; j.bias = w.b; j.x = x; j.out = out; j.x_stride = k; j.out_stride = out_stride; j.n = n; j.k = k; j.flags = 0;
    int& g = is_early ? g_early : g_late;
    j.first_group = g;
    g += (n + 3) / 4;
    (is_early ? early : late).push_back(j);
  }
};

}  // namespace

extern "C" int ug_flux_forward(ug_flux* h, const ug_flux_inputs* in, const ug_flux_outputs* out, void* workspace, size_t workspace_bytes,
                               void* stream) {
  UG_CHECK_ARG(h && in && out && workspace, "flux_forward: null argument");
  UG_TRY(resolve(h));
  const ug_flux_desc& d = h->d;
  const int B = in->batch, N = in->n_img, T = in->n_txt, D = h->D, E = d.experts, n_cond = d.condition_nums;
  UG_CHECK_ARG(B >= 1 && B <= 8 && N >= 1 && T >= 1, "flux_forward: bad shape batch %d n_img %d n_txt %d (batch <= 8)", B, N, T);
  UG_CHECK_ARG(T != N, "flux_forward: T == N: the reference MOELayer would also dispatch the text tensor (unsupported)");
  UG_CHECK_ARG(in->hidden_states && in->encoder_hidden_states && in->pooled_projections && in->timestep && in->img_ids && in->txt_ids,
               "flux_forward: missing input");
  UG_CHECK_ARG(!in->guidance == !d.guidance_embeds || !in->guidance, "flux_forward: guidance given to a model without guidance embeddings");
  for (int c = 0; c < n_cond; ++c)
    UG_CHECK_ARG(in->condition_hidden_states[c] && in->condition_pooled_projections[c] && in->condition_ids[c] && in->rts_uniform[c],
                 "flux_forward: missing input of condition %d", c);
  UG_CHECK_ARG(out->velocity && out->expert_counts && out->l_aux, "flux_forward: missing output buffer");
  UG_CHECK_ARG((reinterpret_cast<uintptr_t>(workspace) & (kAlign - 1)) == 0, "flux_forward: workspace must be %d-byte aligned", (int)kAlign);
  Step s;
  s.h = h; s.in = in; s.out = out; s.ws = static_cast<char*>(workspace);
  s.L = make_layout(h, B, N, T);
  UG_CHECK_ARG(workspace_bytes >= s.L.total, "flux_forward: workspace of %zu bytes is smaller than ug_flux_workspace_bytes() = %zu",
               workspace_bytes, s.L.total);
  s.main = reinterpret_cast<cudaStream_t>(stream);
  s.B = B; s.N = N; s.T = T; s.S = T + N; s.D = D; s.H = d.heads; s.dh = d.head_dim; s.E = E;
  s.cscale = in->conditioning_scale;
  const Layout& L = s.L;
  const int S = s.S, C = L.capacity;
  const int64_t Smax = T + 2 * (int64_t)N;
  cudaStream_t main = s.main;

  // ---- slot table of the AdaLN vectors (same order as model.py::_mod_plans) + job tables, built once per (workspace, shape) ----
  const int nd = d.num_layers, ncd = d.n_ctrl_double, ns = d.num_single_layers, ncs = d.n_ctrl_single;
  std::vector<int> s_d(nd), s_dc(nd), s_cd(ncd), s_cdc(ncd), s_s(ns), s_cs(ncs), s_sh0(n_cond), s_sh0c(n_cond);
  int s_sh1 = 0, s_sh1c = 0, s_out = 0, slot = 0;
  float* stembs = s.f32(L.STEMBS);
  const float *s_temb = stembs, *s_ctemb = stembs + (size_t)B * D, *s_cdtemb = stembs + 2 * (size_t)B * D;
  JobBuilder jb;
  const int64_t mbs = s.mod_bs();
  auto job = [&](bool early, Lin w, const float* x, int chunks) {
    const int s0 = slot;
    slot += chunks;
    jb.push(early, w, x, s.mod(s0), chunks * D, D, mbs);
    return s0;
  };
  for (int i = 0; i < nd; ++i) { s_d[i] = job(i == 0, h->dbl[i].norm1, s_temb, 6); s_dc[i] = job(i == 0, h->dbl[i].norm1_ctx, s_temb, 6); }
  for (int j = 0; j < ncd; ++j) { s_cd[j] = job(j == 0, h->cdbl[j].norm1, s_cdtemb, 6); s_cdc[j] = job(j == 0, h->cdbl[j].norm1_ctx, s_cdtemb, 6); }
  for (int i = 0; i < ns; ++i) s_s[i] = job(false, h->sgl[i].norm, s_temb, 3);
  for (int j = 0; j < ncs; ++j) s_cs[j] = job(false, h->csgl[j].norm, s_cdtemb, 3);
  if (d.use_shared_expert) {
    for (int c = 0; c < n_cond; ++c) {
      const float* x = stembs + (size_t)(3 + c) * B * D;
      s_sh0[c] = job(true, h->shared[0].norm1, x, 6);
      s_sh0c[c] = job(true, h->shared[0].norm1_ctx, x, 6);
    }
    s_sh1 = job(true, h->shared[1].norm1, s_ctemb, 6);
    s_sh1c = job(true, h->shared[1].norm1_ctx, s_ctemb, 6);
  }
  s_out = job(false, h->norm_out, s_temb, 2);
  UG_CHECK_ARG(slot <= L.n_mod && jb.early.size() + jb.late.size() <= UG_MAX_GEMV_JOBS, "flux_forward: AdaLN job table overflow");
  ug_gemv_job* jobs_dev = reinterpret_cast<ug_gemv_job*>(s.ws + L.jobs);
  if (h->plan_ws != workspace || h->plan_B != B || h->plan_N != N || h->plan_T != T) {
    // (first call with this workspace / shape: must not run under stream capture — the table is copied from host memory)
    std::vector<ug_gemv_job> all(jb.early);
    all.insert(all.end(), jb.late.begin(), jb.late.end());
    cudaError_t e = cudaMemcpyAsync(jobs_dev, all.data(), all.size() * sizeof(ug_gemv_job), cudaMemcpyHostToDevice, main);
    if (e == cudaSuccess) e = cudaStreamSynchronize(main);
    if (e != cudaSuccess) {
      set_error("flux_forward: uploading the AdaLN job table failed: %s", cudaGetErrorString(e));
      return UG_ERR_CUDA;
    }
    h->plan_ws = workspace; h->plan_B = B; h->plan_N = N; h->plan_T = T;
  }

  // ---- embeddings (:1215-1239) ----
  View X = s.view(L.X, 0, S, D, S), x_txt = s.view(L.X, 0, T, D, S), x_img = s.view(L.X, T, N, D, S);
  View hs{const_cast<void*>(in->hidden_states), d.in_channels, (int64_t)N * d.in_channels, N};
  View es{const_cast<void*>(in->encoder_hidden_states), d.joint_dim, (int64_t)T * d.joint_dim, T};
  UG_TRY(s.gemm(main, hs, d.in_channels, h->x_embedder, D, x_img));
  UG_TRY(s.gemm(main, es, d.joint_dim, h->context_embedder, D, x_txt));
  UG_TRY(ug_timestep_embedding(in->timestep, in->timestep_stride, B, 256, 1000.0f, s.f32(L.t_emb), main));
  if (in->guidance) UG_TRY(ug_timestep_embedding(in->guidance, 1, B, 256, 1000.0f, s.f32(L.g_emb), main));
  float* tembs = s.f32(L.TEMBS);
  float *temb = tembs, *ctemb = tembs + (size_t)B * D, *cdtemb = tembs + 2 * (size_t)B * D;
  const int64_t P = d.pooled_dim;
  UG_TRY(s.time_text(h->time_text, in->pooled_projections, P, temb, false));
  if (d.use_pooled_prompt_embeds) {
    UG_TRY(s.time_text(h->control_time_text, in->pooled_projections, P, ctemb, false));
  } else {
    // control_pooled_projections = zeros_like(pooled) (:1040-1043): tmp2 is not available, reuse MODC as a zero [B, P] source
    cudaMemsetAsync(s.f32(L.MODC), 0, (size_t)B * P * 4, main);
    UG_TRY(s.time_text(h->control_time_text, s.f32(L.MODC), P, ctemb, false));
  }
  for (int c = 0; c < n_cond; ++c) {
    UG_TRY(s.time_text(h->control_condition, in->condition_pooled_projections[c], P, tembs + (size_t)(3 + c) * B * D, false));
    UG_TRY(s.time_text(h->control_condition, in->condition_pooled_projections[c], P, cdtemb, c > 0));
  }
  UG_TRY(ug_silu_f32(tembs, stembs, (int64_t)(3 + n_cond) * B * D, main));
  UG_TRY(s.rope(L.rope, {{in->txt_ids, T}, {in->img_ids, N}}));

  // ---- AdaLN vectors of every block: `early` on the main stream, `late` on the side stream under the blocks ----
  const int n_early = (int)jb.early.size(), n_late = (int)jb.late.size();
  UG_TRY(ug_gemv_grouped(jobs_dev, n_early, jb.g_early, B, 0, jb.g_early, nullptr, main));
  cudaEventRecord(h->ev_fork, main);
  cudaStreamWaitEvent(h->side, h->ev_fork, 0);
  UG_TRY(ug_gemv_grouped(jobs_dev + n_early, n_late, jb.g_late, B, 0, jb.g_late, nullptr, h->side));
  cudaEventRecord(h->ev_mod, h->side);
  bool mods_joined = false;

  // ---- 19 x [base double -> control double -> add] (:1124-1141) ----
  View CENC = s.view(L.CENC, 0, T, D, T), CH = s.view(L.CH, 0, N, D, N), CIN = s.view(L.CIN, 0, N, D, N), COND = s.view(L.COND, 0, N, D, N);
  bool routed = false;
  for (int i = 0; i < nd; ++i) {
    if (i == 1 && !mods_joined) { cudaStreamWaitEvent(main, h->ev_mod, 0); mods_joined = true; }
    UG_TRY(s.double_block(h->dbl[i], s_d[i], s_dc[i], x_img, x_txt, x_img, &x_txt, L.rope));
    const int j = (int)((double)i / ((double)nd / (double)ncd));  // int(i / (n_base / n_ctrl)) (:1126-1127)
    View ctrl_in = x_img;
    if (!routed) {
      // ---- CoMoE pre-stage at the first control call (:1084-1089, preprocess_moe_forward, moe_forward) ----
      UG_TRY(s.gemm(main, x_txt, D, h->control_context_embedder, D, CENC));
      for (int c = 0; c < n_cond; ++c) {
        View cs{const_cast<void*>(in->condition_hidden_states[c]), d.in_channels, (int64_t)N * d.in_channels, N};
        UG_TRY(s.gemm(main, cs, d.in_channels, h->control_x_embedder, D, COND));
        UG_TRY(s.rope(L.rope0, {{in->condition_ids[c], N}, {in->img_ids, N}}));
        UG_TRY(s.rope(L.rope1, {{in->txt_ids, T}, {in->img_ids, N}, {in->condition_ids[c], N}}));
        View G = s.view(L.G, 0, N, D, N);
        UG_TRY(s.add(x_img, COND, G));
        UG_TRY(ug_moe_route(s.bf(L.G), h->gate_wg, in->rts_uniform[c], B * N, D, E, C, reinterpret_cast<int32_t*>(s.ws + L.route_idx),
                            reinterpret_cast<int32_t*>(s.ws + L.route_slot), s.f32(L.route_prob),
                            reinterpret_cast<int32_t*>(s.ws + L.route_slot_token), out->expert_counts, out->l_aux, s.f32(L.route_ws), main));
        const int32_t* slot_token = reinterpret_cast<const int32_t*>(s.ws + L.route_slot_token);
        // experts: cond' = Wc (s_c . cond) + bc ; hid' = Wh (s_h . (hid + cond')) + bh  (expert_forward :925-967)
        Lin mc{h->exp_mod_w[0], h->exp_mod_b[0]}, mh{h->exp_mod_w[1], h->exp_mod_b[1]};
        UG_TRY(ug_gemv(in->condition_pooled_projections[c], P, mc.w, mc.b, s.f32(L.MODC), (int64_t)E * D, B, E * D, (int)P, 0, 0, 0, main));
        UG_TRY(ug_gemv(in->pooled_projections, P, mh.w, mh.b, s.f32(L.MODH), (int64_t)E * D, B, E * D, (int)P, 0, 0, 0, main));
        UG_TRY(ug_moe_gather_modulate(s.bf(L.COND), slot_token, s.f32(L.MODC), D, (int64_t)E * D, nullptr, s.bf(L.A), E, C, N, D, main));
        View A{s.bf(L.A), D, (int64_t)C * D, C}, YC{s.bf(L.YC), D, (int64_t)C * D, C}, YH{s.bf(L.YH), D, (int64_t)C * D, C};
        UG_TRY(s.gemm(main, A, D, Lin{h->exp_w[0], h->exp_b[0]}, D, YC, nullptr, nullptr, UG_ACT_NONE, 1.0f, E, (int64_t)D * D, D));
        View EH = s.view(L.EH, 0, N, D, N), EC = s.view(L.EC, 0, N, D, N);
        UG_TRY(ug_copy_bf16(x_img.p, x_img.rs, x_img.bs, EH.p, EH.rs, EH.bs, B, N, D, main));
        UG_TRY(ug_moe_gather_modulate(s.bf(L.EH), slot_token, s.f32(L.MODH), D, (int64_t)E * D, s.bf(L.YC), s.bf(L.A), E, C, N, D, main));
        UG_TRY(s.gemm(main, A, D, Lin{h->exp_w[1], h->exp_b[1]}, D, YH, nullptr, nullptr, UG_ACT_NONE, 1.0f, E, (int64_t)D * D, D));
        View HC = s.view(L.HC, 0, 2 * N, D, 2 * N), hc_h = s.view(L.HC, 0, N, D, 2 * N), hc_c = s.view(L.HC, N, N, D, 2 * N);
        if (d.use_shared_expert) {  // V2 shared experts (:1013-1022)
          UG_TRY(s.double_block(h->shared[0], s_sh0[c], s_sh0c[c], x_img, COND, hc_h, &hc_c, L.rope0));
          UG_TRY(s.double_block(h->shared[1], s_sh1, s_sh1c, HC, CENC, HC, nullptr, L.rope1));
        }
        const int32_t* ridx = reinterpret_cast<const int32_t*>(s.ws + L.route_idx);
        const int32_t* rslot = reinterpret_cast<const int32_t*>(s.ws + L.route_slot);
        UG_TRY(ug_moe_combine(s.bf(L.YH), ridx, rslot, s.f32(L.route_prob), s.bf(L.EH), B * N, C, D, main));
        UG_TRY(ug_moe_combine(s.bf(L.YC), ridx, rslot, s.f32(L.route_prob), s.bf(L.EC), B * N, C, D, main));
        if (!d.use_shared_expert) {
          if (c == 0) { UG_TRY(s.add(EH, EC, CIN)); }
          else { UG_TRY(s.add(CIN, EH, CIN)); UG_TRY(s.add(CIN, EC, CIN)); }
        } else {
          if (c == 0) { UG_TRY(s.add(hc_h, EH, CIN)); }
          else { UG_TRY(s.add(CIN, hc_h, CIN)); UG_TRY(s.add(CIN, EH, CIN)); }
          UG_TRY(s.add(CIN, hc_c, CIN));
          UG_TRY(s.add(CIN, EC, CIN));
        }
      }
      routed = true;
      ctrl_in = CIN;
    }
    UG_TRY(s.double_block(h->cdbl[j], s_cd[j], s_cdc[j], ctrl_in, CENC, CH, nullptr, L.rope));
    UG_TRY(s.gemm(main, CH, D, h->add_dbl[j], D, x_img, nullptr, &x_img, UG_ACT_NONE, s.cscale));
  }
  if (!mods_joined) cudaStreamWaitEvent(main, h->ev_mod, 0);

  // ---- 38 x [base single -> control single -> add] (:1146-1172) ----
  View CS = s.view(L.CS, 0, S, D, S);
  for (int i = 0; i < ns; ++i) {
    UG_TRY(s.single_block(h->sgl[i], s_s[i], X, X));
    if (ncs > 0) {
      const int j = (int)((double)i / ((double)ns / (double)ncs));
      UG_TRY(s.single_block(h->csgl[j], s_cs[j], X, CS));
      if (d.single_add) {
        View cs_img = s.view(L.CS, T, N, D, S);
        UG_TRY(s.gemm(main, cs_img, D, h->add_sgl[j], D, x_img, nullptr, &x_img, UG_ACT_NONE, s.cscale));
      } else {
        UG_TRY(s.gemm(main, CS, D, h->add_sgl[j], D, X, nullptr, &X, UG_ACT_NONE, s.cscale));
      }
    }
  }
  // ---- norm_out (AdaLayerNormContinuous: scale first, then shift) + proj_out (:1264-1265) ----
  View NO = s.view(L.NO, 0, N, D, N);
  UG_TRY(s.ln(main, x_img, NO, s.mod(s_out + 1), s.mod(s_out)));
  View vel{out->velocity, d.in_channels, (int64_t)N * d.in_channels, N};
  UG_TRY(s.gemm(main, NO, D, h->proj_out, d.in_channels, vel));
  (void)Smax;
  return UG_OK;
}
