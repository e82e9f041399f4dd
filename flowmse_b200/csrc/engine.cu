// libflowse engine: context, weight packing, the static per-(B,T) execution plan of one NCSN++ evaluation,
// the N-step sampler loop, and the C ABI declared in include/flowse.h.
//
// The reference walks a flat nn.ModuleList by index in Python for every NFE
// (/root/reference/flowmse/backbones/ncsnpp.py:247-404), ~1,860 ATen dispatches each.  Here the walk is done ONCE per
// (B,T): it resolves every buffer, tensor map and launch shape into a flat op list that is replayed (optionally as a
// CUDA graph) for each of the N (Euler) or 2N-1 (Heun) evaluations of the sampler loop
// (/root/reference/flowmse/sampling/__init__.py:48-57).
#include "../../include/flowse.h"
#include "flowse_internal.h"

#include <algorithm>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <functional>
#include <map>
#include <memory>
#include <string>
#include <unordered_map>
#include <vector>

using namespace flowse;

namespace {

std::string g_create_error;

#define CK(call)                                                                           \
  do {                                                                                     \
    cudaError_t _e = (call);                                                               \
    if (_e != cudaSuccess) {                                                               \
      ctx->err = std::string(#call) + ": " + cudaGetErrorString(_e);                       \
      return 1;                                                                            \
    }                                                                                      \
  } while (0)

constexpr int NF = 128;
constexpr int kNumLevels = 7;
constexpr int kChMult[kNumLevels] = {1, 1, 2, 2, 2, 2, 2};
constexpr int kNumResBlocks = 2;
constexpr int kImage = 256;
constexpr int kAttnRes = 16;
constexpr int kTemb = 512;
constexpr int kStatSlotDoubles = kStatReplicas * 64 * 2;   // per batch element: up to 256 channels
constexpr int kMaxStatSlots = 160;

struct ModSpec {
  enum Kind { Fourier, Linear, Conv3, RB, Attn, Combine, GN } kind;
  int cin = 0, cout = 0;
  bool up = false, down = false;
};

// Same derivation as flowmse/backbones/ncsnpp.py:99-245 (default config).
std::vector<ModSpec> build_modules() {
  std::vector<ModSpec> m;
  auto add = [&](ModSpec::Kind k, int ci, int co, bool up = false, bool down = false) {
    ModSpec s; s.kind = k; s.cin = ci; s.cout = co; s.up = up; s.down = down; m.push_back(s);
  };
  add(ModSpec::Fourier, 0, NF);
  add(ModSpec::Linear, 2 * NF, 4 * NF);
  add(ModSpec::Linear, 4 * NF, 4 * NF);
  add(ModSpec::Conv3, 4, NF);
  std::vector<int> hs_c = {NF};
  int in_ch = NF;
  for (int l = 0; l < kNumLevels; ++l) {
    for (int i = 0; i < kNumResBlocks; ++i) {
      const int out_ch = NF * kChMult[l];
      add(ModSpec::RB, in_ch, out_ch);
      in_ch = out_ch;
      if ((kImage >> l) == kAttnRes) add(ModSpec::Attn, in_ch, in_ch);
      hs_c.push_back(in_ch);
    }
    if (l != kNumLevels - 1) {
      add(ModSpec::RB, in_ch, in_ch, false, true);
      add(ModSpec::Combine, 4, in_ch);
      hs_c.push_back(in_ch);
    }
  }
  in_ch = hs_c.back();
  add(ModSpec::RB, in_ch, in_ch);
  add(ModSpec::Attn, in_ch, in_ch);
  add(ModSpec::RB, in_ch, in_ch);
  for (int l = kNumLevels - 1; l >= 0; --l) {
    for (int i = 0; i < kNumResBlocks + 1; ++i) {
      const int out_ch = NF * kChMult[l];
      add(ModSpec::RB, in_ch + hs_c.back(), out_ch);
      hs_c.pop_back();
      in_ch = out_ch;
    }
    if ((kImage >> l) == kAttnRes) add(ModSpec::Attn, in_ch, in_ch);
    add(ModSpec::GN, in_ch, in_ch);
    add(ModSpec::Conv3, in_ch, 4);
    if (l != 0) add(ModSpec::RB, in_ch, in_ch, true, false);
  }
  return m;
}

struct ConvW {
  __half* wp = nullptr;   // [2][Npad][K]
  int Npad = 0, K = 0;
  float wscale_inv = 1.f;
};

struct RBW {
  int cin = 0, cout = 0;
  bool up = false, down = false, has_sc = false;
  float *gn0_g = nullptr, *gn0_b = nullptr, *gn1_g = nullptr, *gn1_b = nullptr;
  ConvW conv0, conv1;      // conv1 includes the Conv_2 shortcut K blocks when has_sc
  float* bias1 = nullptr;  // Conv_1.bias (+ Conv_2.bias)
  int dense_off = 0;       // row offset into the Dense_0 table
};
struct AttnW {
  int c = 0;
  float *gn_g = nullptr, *gn_b = nullptr, *wqkv = nullptr, *bqkv = nullptr, *w3 = nullptr, *b3 = nullptr;
};
struct CombW { int c = 0; float *w = nullptr, *b = nullptr; };
struct HeadW { int c = 0; float *gn_g = nullptr, *gn_b = nullptr, *bias = nullptr, *wf = nullptr; };   // wf: [9][C][4] fp32

struct Act { float* p = nullptr; int C = 0, H = 0, W = 0; double* qs = nullptr; };   // qs: quad statistics [B][C/4][2]

// kind: 0 misc, 1 gn_stats, 2 gn_prep, 3 conv_gemm (per-tap kernel), 4 attention, 5 small 4-channel / input kernels,
//       6 time embedding, 7 conv_halo (halo kernel; chosen at plan time with the conv_impl option in force)
// branch: 0 = the main chain; 1 = output pyramid (the heads read h of the main chain and the previous head: nothing on the
// main chain needs them before the final kernel); 2 = input pyramid (the FIR-downsample chain depends only on the input).
// Under graph capture the side branches become parallel graph branches: they fill the SMs the short low-resolution
// kernels of the main chain leave idle.  needs_main: wait for the main chain's current position; wait_branch: a main op
// that consumes what the branch has produced so far.
struct Op { std::function<int(cudaStream_t)> fn; int nk; int kind; double flops; int info[4]; int branch = 0; bool needs_main = true;
            int wait_branch = 0; };

struct Plan;
void destroy_graphs(Plan* p);

struct Plan {
  int B = 0, T = 0;
  char* arena = nullptr;
  size_t arena_bytes = 0;
  std::vector<Op> ops;
  int kernels_per_forward = 0;
  std::map<int, Act> taps;
  // fixed I/O buffers
  float2 *x = nullptr, *y = nullptr, *z = nullptr, *vout = nullptr, *xa = nullptr, *va = nullptr, *vb = nullptr;
  float* t_dev = nullptr;       // [B]
  float* step_dev = nullptr;    // [1]
  float4* pyr_out = nullptr;    // final 4-plane pyramid [B,256,T,4]
  double* stats = nullptr; size_t stats_bytes = 0;
  double* gn_partials = nullptr; unsigned* gn_counters = nullptr;
  cudaGraphExec_t graph_fwd = nullptr;      // forward, final mode 0/1 chosen at launch via separate final op
  bool graph_ready = false;
  int eager_runs = 0;   // the first evaluation of a plan runs eagerly (sets kernel attributes, surfaces errors)
  // time embedding: a function of t only, so it is NOT part of the replayed evaluation.  flowse_sample evaluates it for
  // every time point of the schedule in one launch pair (rows = evaluations x B) and copies row block e into the
  // fixed bias table before evaluation e; flowse_ncsnpp_forward runs it for its B rows.
  std::function<int(const float* t_rows, int rows, float* table, cudaStream_t)> temb_fn;
  float* bias_table = nullptr;              // [B][R], read by the Conv_0 epilogues
  float* t_all = nullptr;                   // [kMaxEvals * B] evaluation times of a sampler call
  float* temb_all = nullptr;                // [kMaxEvals * B][512]
  float* bias_all = nullptr;                // [kMaxEvals][B][R]
  float2* yp = nullptr;                     // prior mean when it differs from y
  // whole-sampler graphs (prior + every evaluation + fused updates), keyed by schedule / solver / sigma
  struct SamplerGraph { std::vector<float> ts; int solver; float sigma; bool own_prior; cudaGraphExec_t exec; long long kernels; int seen; };
  std::vector<SamplerGraph> sampler_graphs;
};
constexpr int kMaxEvals = 64;               // evaluations whose time embeddings are computed by one launch pair

}  // namespace

struct flowse_ctx {
  int device = 0;
  std::string err;
  bool weights_loaded = false;
  std::vector<ModSpec> mods;
  // weights: every device buffer in upload order (the packed export / import walks this list), and the per-conv
  // power-of-two weight scales in upload order
  std::vector<std::pair<void*, size_t>> dev_allocs;
  std::vector<float> conv_scales;
  struct PackedSrc { const char* base; std::vector<unsigned long long> seg; std::vector<float> scales; size_t seg_i, scale_i, off; };
  PackedSrc* packed = nullptr;              // set while flowse_load_packed rebuilds the weights from a packed blob
  TembWeights temb{};
  float *conv_in_w = nullptr, *conv_in_b = nullptr, *out_w = nullptr, *out_b = nullptr;
  std::map<int, RBW> rbs;
  std::map<int, AttnW> attns;
  std::map<int, CombW> combs;
  std::map<int, HeadW> heads;   // keyed by the GN module index
  int dense_rows = 0;
  // options
  int conv_impl = 0;
  int use_graph = 1;
  long long graph_kernels = 0;      // kernels executed through graph replays (not seen by launch_counter)
  long long counter_base = 0;
  std::unique_ptr<Plan> plan;
  cudaStream_t cap_stream = nullptr;   // capture happens here: the caller's stream may be the legacy default stream
  cudaStream_t side[2] = {nullptr, nullptr};          // side branches of the captured graphs (output / input pyramid)
  cudaEvent_t ev_main = nullptr, ev_branch[3] = {nullptr, nullptr, nullptr};
  int fork_branches = 1;                              // option "fork": capture the pyramid branches as parallel graph branches
  // scratch for op-level entry points
  double* op_stats = nullptr; double* op_partials = nullptr; unsigned* op_counters = nullptr;
  float* op_scratch = nullptr; size_t op_scratch_bytes = 0;
  float* op_splitk = nullptr;
  // STFT / iSTFT (stft.cu): bases built on first use, scratch grown on demand
  // one grow-only workspace shared by successive plans (cudaMalloc / cudaFree of multi-GiB arenas on every (B,T) switch
  // cost up to hundreds of milliseconds in the bucketed evaluate driver)
  char* arena = nullptr; size_t arena_cap = 0;
  float* stft_basis = nullptr;
  int stft_window = 0, stft_window_built = -1;   // 0 hann, 1 sqrthann (data_module.py:13-19)
  int spec_transform = 0;                        // 0 exponent, 1 log, 2 none (data_module.py:149-175)
  char* stft_scratch = nullptr; size_t stft_scratch_bytes = 0;
  // sticky fp16 range flag: operand-producing kernels count the values whose magnitude exceeds the fp16 hi/lo range
  // (|v| > 65504 saturates silently otherwise); read back through flowse_fp16_overflow
  unsigned long long* overflow = nullptr;
  double* rk_acc = nullptr;                 // device accumulator of flowse_rk_lincomb's error norm
  int whole_graph = 1;                      // capture the whole sampler call as one CUDA graph (second call with the same schedule)
  int fuse_prep = 1;                        // the halo conv kernel prepares its own operands (no standalone prep pass for those layers)
};

namespace {

// ------------------------------------------------------------------------------------------------
// weights
// ------------------------------------------------------------------------------------------------
struct HostBlob {
  const float* base;
  std::unordered_map<std::string, std::pair<long long, long long>> idx;
  const float* zeros = nullptr;     // packed import: the fp32 tensors are not available (nor needed), every lookup gets zeros
  const float* get(const std::string& name, long long numel, std::string* err) const {
    if (zeros) return zeros;
    auto it = idx.find(name);
    if (it == idx.end()) { *err = "missing tensor '" + name + "'"; return nullptr; }
    if (it->second.second != numel) {
      *err = "tensor '" + name + "' has " + std::to_string(it->second.second) + " elements, expected " +
             std::to_string(numel);
      return nullptr;
    }
    return base + it->second.first;
  }
};

int dev_upload(flowse_ctx* ctx, const void* host, size_t bytes, void** out) {
  if (ctx->packed) {               // packed import: the next segment of the blob IS this buffer
    auto* pk = ctx->packed;
    if (pk->seg_i >= pk->seg.size() || pk->seg[pk->seg_i] != bytes) {
      ctx->err = "packed weights: segment " + std::to_string(pk->seg_i) + " does not match this library's layout";
      return 2;
    }
    host = pk->base + pk->off;
    pk->off += (bytes + 255) & ~static_cast<size_t>(255);
    ++pk->seg_i;
  }
  void* d = nullptr;
  CK(cudaMalloc(&d, bytes));
  ctx->dev_allocs.push_back({d, bytes});
  CK(cudaMemcpy(d, host, bytes, cudaMemcpyHostToDevice));
  *out = d;
  return 0;
}
int up_f32(flowse_ctx* ctx, const float* host, size_t n, float** out) {
  return dev_upload(ctx, host, n * sizeof(float), reinterpret_cast<void**>(out));
}

int upload_conv(flowse_ctx* ctx, const float* w_main, int Cout, int Cin, int ntaps, const float* w_sc, int Cin2,
                int Npad, ConvW* out) {
  const int K = ntaps * Cin + (w_sc ? Cin2 : 0);
  out->Npad = Npad; out->K = K;
  const size_t bytes = static_cast<size_t>(2) * Npad * K * sizeof(__half);
  if (ctx->packed) {               // already in the K-major fp16 hi/lo format: no host-side packing
    auto* pk = ctx->packed;
    if (pk->scale_i >= pk->scales.size()) { ctx->err = "packed weights: conv scale list too short"; return 2; }
    out->wscale_inv = pk->scales[pk->scale_i++];
    ctx->conv_scales.push_back(out->wscale_inv);
    return dev_upload(ctx, nullptr, bytes, reinterpret_cast<void**>(&out->wp));
  }
  std::vector<__half> buf(static_cast<size_t>(2) * Npad * K);
  const int e = pack_conv_weights_host(w_main, Cout, Cin, ntaps, w_sc, Cin2, Npad, buf.data(),
                                       buf.data() + static_cast<size_t>(Npad) * K);
  out->wscale_inv = std::ldexp(1.0f, -e);
  ctx->conv_scales.push_back(out->wscale_inv);
  return dev_upload(ctx, buf.data(), bytes, reinterpret_cast<void**>(&out->wp));
}

int load_weights_impl(flowse_ctx* ctx, const HostBlob& hb) {
  std::string e;
  auto P = [&](int i) { return "all_modules." + std::to_string(i) + "."; };
#define GET(var, name, numel)                                      \
  const float* var = hb.get(name, numel, &e);                      \
  if (!var) { ctx->err = e; return 2; }
  const auto& mods = ctx->mods;

  // time embedding
  GET(fw, P(0) + "W", NF);
  GET(l1w, P(1) + "weight", 512LL * 256); GET(l1b, P(1) + "bias", 512);
  GET(l2w, P(2) + "weight", 512LL * 512); GET(l2b, P(2) + "bias", 512);
  float *d_fw, *d_l1w, *d_l1b, *d_l2w, *d_l2b;
  if (up_f32(ctx, fw, NF, &d_fw) || up_f32(ctx, l1w, 512 * 256, &d_l1w) || up_f32(ctx, l1b, 512, &d_l1b) ||
      up_f32(ctx, l2w, 512 * 512, &d_l2w) || up_f32(ctx, l2b, 512, &d_l2b)) return 1;
  ctx->temb.fourier_W = d_fw; ctx->temb.l1_w = d_l1w; ctx->temb.l1_b = d_l1b; ctx->temb.l2_w = d_l2w; ctx->temb.l2_b = d_l2b;

  GET(ciw, P(3) + "weight", 128LL * 4 * 9); GET(cib, P(3) + "bias", 128);
  if (up_f32(ctx, ciw, 128 * 36, &ctx->conv_in_w) || up_f32(ctx, cib, 128, &ctx->conv_in_b)) return 1;
  GET(ow, "output_layer.weight", 8); GET(ob, "output_layer.bias", 2);
  if (up_f32(ctx, ow, 8, &ctx->out_w) || up_f32(ctx, ob, 2, &ctx->out_b)) return 1;

  int R = 0;
  for (const auto& m : mods) if (m.kind == ModSpec::RB) R += m.cout;
  ctx->dense_rows = R;
  std::vector<float> dense_w(static_cast<size_t>(R) * kTemb), dense_b(R);
  int row = 0;
  for (int i = 0; i < static_cast<int>(mods.size()); ++i) {
    const ModSpec& m = mods[i];
    if (m.kind == ModSpec::RB) {
      RBW r; r.cin = m.cin; r.cout = m.cout; r.up = m.up; r.down = m.down;
      r.has_sc = (m.cin != m.cout) || m.up || m.down;
      GET(g0, P(i) + "GroupNorm_0.weight", m.cin); GET(b0, P(i) + "GroupNorm_0.bias", m.cin);
      GET(c0w, P(i) + "Conv_0.weight", 9LL * m.cin * m.cout); GET(c0b, P(i) + "Conv_0.bias", m.cout);
      GET(dw, P(i) + "Dense_0.weight", static_cast<long long>(m.cout) * kTemb); GET(db, P(i) + "Dense_0.bias", m.cout);
      GET(g1, P(i) + "GroupNorm_1.weight", m.cout); GET(b1, P(i) + "GroupNorm_1.bias", m.cout);
      GET(c1w, P(i) + "Conv_1.weight", 9LL * m.cout * m.cout); GET(c1b, P(i) + "Conv_1.bias", m.cout);
      const float *c2w = nullptr, *c2b = nullptr;
      if (r.has_sc) {
        c2w = hb.get(P(i) + "Conv_2.weight", static_cast<long long>(m.cin) * m.cout, &e);
        c2b = hb.get(P(i) + "Conv_2.bias", m.cout, &e);
        if (!c2w || !c2b) { ctx->err = e; return 2; }
      }
      if (up_f32(ctx, g0, m.cin, &r.gn0_g) || up_f32(ctx, b0, m.cin, &r.gn0_b) || up_f32(ctx, g1, m.cout, &r.gn1_g) ||
          up_f32(ctx, b1, m.cout, &r.gn1_b)) return 1;
      if (upload_conv(ctx, c0w, m.cout, m.cin, 9, nullptr, 0, m.cout, &r.conv0)) return 1;
      if (upload_conv(ctx, c1w, m.cout, m.cout, 9, c2w, m.cin, m.cout, &r.conv1)) return 1;
      std::vector<float> bias1(m.cout);
      for (int c = 0; c < m.cout; ++c) bias1[c] = c1b[c] + (c2b ? c2b[c] : 0.f);
      if (up_f32(ctx, bias1.data(), m.cout, &r.bias1)) return 1;
      r.dense_off = row;
      std::memcpy(dense_w.data() + static_cast<size_t>(row) * kTemb, dw, sizeof(float) * m.cout * kTemb);
      for (int c = 0; c < m.cout; ++c) dense_b[row + c] = db[c] + c0b[c];
      row += m.cout;
      ctx->rbs[i] = r;
    } else if (m.kind == ModSpec::Attn) {
      const int c = m.cin;
      AttnW a; a.c = c;
      GET(g, P(i) + "GroupNorm_0.weight", c); GET(b, P(i) + "GroupNorm_0.bias", c);
      const float* W[4]; const float* bb[4];
      for (int k = 0; k < 4; ++k) {
        W[k] = hb.get(P(i) + "NIN_" + std::to_string(k) + ".W", static_cast<long long>(c) * c, &e);
        bb[k] = hb.get(P(i) + "NIN_" + std::to_string(k) + ".b", c, &e);
        if (!W[k] || !bb[k]) { ctx->err = e; return 2; }
      }
      std::vector<float> wqkv(static_cast<size_t>(c) * 3 * c), bqkv(3 * c);
      for (int k = 0; k < 3; ++k) {
        for (int r = 0; r < c; ++r)
          std::memcpy(wqkv.data() + (static_cast<size_t>(r) * 3 + k) * c, W[k] + static_cast<size_t>(r) * c, sizeof(float) * c);
        std::memcpy(bqkv.data() + k * c, bb[k], sizeof(float) * c);
      }
      if (up_f32(ctx, g, c, &a.gn_g) || up_f32(ctx, b, c, &a.gn_b) || up_f32(ctx, wqkv.data(), wqkv.size(), &a.wqkv) ||
          up_f32(ctx, bqkv.data(), bqkv.size(), &a.bqkv) || up_f32(ctx, W[3], static_cast<size_t>(c) * c, &a.w3) ||
          up_f32(ctx, bb[3], c, &a.b3)) return 1;
      ctx->attns[i] = a;
    } else if (m.kind == ModSpec::Combine) {
      CombW cw; cw.c = m.cout;
      GET(w, P(i) + "Conv_0.weight", 4LL * m.cout); GET(b, P(i) + "Conv_0.bias", m.cout);
      if (up_f32(ctx, w, 4 * m.cout, &cw.w) || up_f32(ctx, b, m.cout, &cw.b)) return 1;
      ctx->combs[i] = cw;
    } else if (m.kind == ModSpec::GN) {
      HeadW h; h.c = m.cin;
      GET(g, P(i) + "weight", m.cin); GET(b, P(i) + "bias", m.cin);
      GET(cw, P(i + 1) + "weight", 9LL * m.cin * 4); GET(cb, P(i + 1) + "bias", 4);
      if (up_f32(ctx, g, m.cin, &h.gn_g) || up_f32(ctx, b, m.cin, &h.gn_b) || up_f32(ctx, cb, 4, &h.bias)) return 1;
      // PyTorch [4][C][3][3] -> tap-major [9][C][4] (the 4 outputs of one input channel contiguous) for head_conv_kernel
      std::vector<float> wf(static_cast<size_t>(9) * m.cin * 4);
      for (int o = 0; o < 4; ++o)
        for (int c = 0; c < m.cin; ++c)
          for (int t = 0; t < 9; ++t) wf[(static_cast<size_t>(t) * m.cin + c) * 4 + o] = cw[(static_cast<size_t>(o) * m.cin + c) * 9 + t];
      if (up_f32(ctx, wf.data(), wf.size(), &h.wf)) return 1;
      ctx->heads[i] = h;
    }
  }
  float *d_dw, *d_db;
  if (up_f32(ctx, dense_w.data(), dense_w.size(), &d_dw) || up_f32(ctx, dense_b.data(), dense_b.size(), &d_db)) return 1;
  ctx->temb.dense_w = d_dw; ctx->temb.dense_b = d_db; ctx->temb.R = R;
#undef GET
  return 0;
}

// ------------------------------------------------------------------------------------------------
// plan
// ------------------------------------------------------------------------------------------------
struct Arena {
  char* base = nullptr;
  size_t off = 0;
  template <typename T>
  T* alloc(size_t n) {
    off = (off + 1023) & ~static_cast<size_t>(1023);
    T* p = base ? reinterpret_cast<T*>(base + off) : nullptr;
    off += n * sizeof(T);
    return p;
  }
};

// conv_impl: 0 = auto (halo kernel for the high-resolution layers, per-tap kernel otherwise), 1 = SIMT cross-check,
// 2 = per-tap tcgen05 kernel everywhere, 3 = auto with 3 rotating main accumulators in the halo kernel,
// 4 = auto with the CTA-pair (cta_group::2) halo kernel.
// the halo kernel takes the layers with at least ~2/3 of a wave of its 16 x 8 pixel x 128 channel tiles
bool conv_uses_halo(const flowse_ctx* ctx, const ConvGemmArgs& a) {
  return ctx->conv_impl != 1 && ctx->conv_impl != 2 && conv_halo_supported(a) &&
         static_cast<long long>(a.B) * (a.H / 16) * (a.W / 8) * ((a.Cout + 127) / 128) >= 100;
}

int run_conv(flowse_ctx* ctx, const ConvGemmArgs& a, cudaStream_t s) {
  std::string e;
  int rc;
  if (ctx->conv_impl == 1) rc = launch_conv_gemm_simt(a, s, &e);
  else if (conv_uses_halo(ctx, a))
    rc = launch_conv_halo(a, ctx->conv_impl == 3 ? 3 : (ctx->conv_impl == 4 ? 2 : 1), s, &e);
  else rc = launch_conv_gemm(a, s, &e);
  if (rc) ctx->err = e;
  return rc;
}

struct Builder {
  flowse_ctx* ctx;
  Plan* plan;
  Arena ar;
  bool dry;
  int B, T;
  int stat_slots = 0;
  // scratch
  __half *scrA = nullptr, *scrX = nullptr;
  float *scrH1 = nullptr, *scrF = nullptr, *scrQKV = nullptr, *scrS = nullptr, *scrO = nullptr;
  float *temb_act = nullptr, *bias_table = nullptr, *splitk = nullptr;

  void push(int nk, std::function<int(cudaStream_t)> fn, int kind = 0, double flops = 0.0, int i0 = 0, int i1 = 0,
            int i2 = 0, int i3 = 0) {
    if (!dry) plan->ops.push_back(Op{std::move(fn), nk, kind, flops, {i0, i1, i2, i3}});
  }
  void tag_last(int branch, bool needs_main, int wait_branch = 0) {
    if (!dry) { Op& o = plan->ops.back(); o.branch = branch; o.needs_main = needs_main; o.wait_branch = wait_branch; }
  }
  static double conv_flops(const ConvGemmArgs& a) {
    const double K = static_cast<double>(a.ntaps) * a.Cin + a.Cin2;      // Cin2 = 0 without a folded shortcut
    return 2.0 * a.B * a.H * a.W * a.Cout * K;
  }
  // one slot = quad statistics of a tensor with up to 256 channels: [B][64][2] doubles
  double* stat_slot() {
    double* p = plan->stats ? plan->stats + static_cast<size_t>(stat_slots) * B * kStatSlotDoubles : nullptr;
    ++stat_slots;
    return p;
  }
  Act new_act(int C, int H, int W) {
    Act a; a.C = C; a.H = H; a.W = W;
    a.p = ar.alloc<float>(static_cast<size_t>(B) * H * W * C);
    a.qs = stat_slot();
    return a;
  }


  Act resblock(int mi, const Act& in1, const Act* in2) {
    const RBW& r = ctx->rbs.at(mi);
    const int Cin = in1.C + (in2 ? in2->C : 0);
    const int H = in1.H, W = in1.W;
    const int Ho = r.down ? H / 2 : (r.up ? H * 2 : H), Wo = r.down ? W / 2 : (r.up ? W * 2 : W);
    Act out = new_act(r.cout, Ho, Wo);
    double* st1 = stat_slot();                    // statistics of the Conv_0 output, filled by its epilogue
    const float* s1 = in1.p; const int C1 = in1.C;
    const float* s2 = in2 ? in2->p : nullptr; const int C2 = in2 ? in2->C : 0;
    // conv arguments first: whether a layer runs on the halo kernel decides who prepares its operands
    ConvGemmArgs c0{};
    c0.A = scrA; c0.Cin = Cin; c0.ntaps = 9; c0.X = nullptr; c0.Cin2 = 0; c0.Wp = r.conv0.wp; c0.Npad = r.conv0.Npad;
    c0.wscale_inv = r.conv0.wscale_inv; c0.bias = bias_table + r.dense_off; c0.bias_bstride = ctx->dense_rows;
    c0.residual = nullptr; c0.div_sqrt2 = 0; c0.out = scrH1; c0.Cout = r.cout; c0.ldc = r.cout;
    c0.B = B; c0.H = Ho; c0.W = Wo;
    c0.splitk_scratch = splitk; c0.splitk_scratch_elems = kSplitKScratchElems; c0.qstats = st1;
    c0.overflow = ctx->overflow;
    float* h1 = scrH1; const int Co = r.cout;
    ConvGemmArgs c1{};
    c1.A = scrA; c1.Cin = Co; c1.ntaps = 9; c1.X = r.has_sc ? scrX : nullptr; c1.Cin2 = r.has_sc ? Cin : 0;
    c1.Wp = r.conv1.wp; c1.Npad = r.conv1.Npad; c1.wscale_inv = r.conv1.wscale_inv; c1.bias = r.bias1;
    c1.bias_bstride = 0; c1.residual = r.has_sc ? nullptr : s1; c1.div_sqrt2 = 1; c1.out = out.p; c1.Cout = Co;
    c1.ldc = Co; c1.B = B; c1.H = Ho; c1.W = Wo;
    c1.splitk_scratch = splitk; c1.splitk_scratch_elems = kSplitKScratchElems; c1.qstats = out.qs;
    c1.overflow = ctx->overflow;
    // Fused operand preparation (halo-kernel layers): GroupNorm + SiLU + fp16 split happen in the conv kernel's transform
    // warps, reading the fp32 activations directly.  A resampling block keeps its FIR prep for Conv_0 and for the shortcut
    // operand; its Conv_1 still normalises h1 itself.
    const bool resample = r.up || r.down;
    // Only the halo kernel fuses: the same transform in the per-tap kernel (low-resolution layers, 2-10 K blocks per CTA)
    // was measured 2x slower than prep + TMA - the operand of every K block costs ~3 us to prepare against 0.5 us to fetch
    // (profiles/README.md, round 2) - and was removed.
    auto can_fuse = [&](const ConvGemmArgs& c) { return ctx->fuse_prep >= 1 && conv_uses_halo(ctx, c); };
    const bool fuse0 = !resample && can_fuse(c0) && Cin <= 512;
    const bool fuse1 = can_fuse(c1) && Co <= 512;
    if (fuse0) {
      c0.A = nullptr;
      c0.fA.s1 = s1; c0.fA.C1 = C1; c0.fA.s2 = s2; c0.fA.C2 = C2; c0.fA.qs1 = in1.qs; c0.fA.qs2 = in2 ? in2->qs : nullptr;
      c0.fA.gamma = r.gn0_g; c0.fA.beta = r.gn0_b; c0.fA.silu = 1;
    }
    if (fuse1) {
      c1.A = nullptr;
      c1.fA.s1 = h1; c1.fA.C1 = Co; c1.fA.qs1 = st1; c1.fA.gamma = r.gn1_g; c1.fA.beta = r.gn1_b; c1.fA.silu = 1;
      if (r.has_sc && !resample) {          // raw block input as the 1x1 shortcut operand
        c1.X = nullptr;
        c1.fX.s1 = s1; c1.fX.C1 = C1; c1.fX.s2 = s2; c1.fX.C2 = C2;
      }
    }
    const bool need_x_operand = r.has_sc && !(fuse1 && !resample);
    if (!fuse0 || need_x_operand) {
      PrepArgs pa{};
      pa.src1 = s1; pa.C1 = C1; pa.src2 = s2; pa.C2 = C2; pa.qs1 = in1.qs; pa.qs2 = in2 ? in2->qs : nullptr;
      pa.gamma = r.gn0_g; pa.beta = r.gn0_b;
      pa.B = B; pa.H = H; pa.W = W; pa.mode = r.down ? kPrepDown : (r.up ? kPrepUp : kPrepPlain); pa.silu = 1;
      pa.outA = fuse0 ? nullptr : scrA; pa.outX = need_x_operand ? scrX : nullptr; pa.overflow = ctx->overflow;
      push(1, [=](cudaStream_t s) { launch_gn_prep(pa, s); return 0; }, 2);
    }
    { flowse_ctx* cx = ctx; push(1, [=](cudaStream_t s) { return run_conv(cx, c0, s); }, conv_uses_halo(cx, c0) ? 7 : 3, conv_flops(c0), c0.H, c0.W, c0.ntaps * c0.Cin, c0.Cout); }
    if (!fuse1) {
      PrepArgs pb{};
      pb.src1 = h1; pb.C1 = Co; pb.src2 = nullptr; pb.C2 = 0; pb.qs1 = st1; pb.gamma = r.gn1_g; pb.beta = r.gn1_b;
      pb.B = B; pb.H = Ho; pb.W = Wo; pb.mode = kPrepPlain; pb.silu = 1; pb.outA = scrA; pb.overflow = ctx->overflow;
      push(1, [=](cudaStream_t s) { launch_gn_prep(pb, s); return 0; }, 2);
    }
    { flowse_ctx* cx = ctx; push(1, [=](cudaStream_t s) { return run_conv(cx, c1, s); }, conv_uses_halo(cx, c1) ? 7 : 3, conv_flops(c1), c1.H, c1.W, c1.ntaps * c1.Cin + c1.Cin2, c1.Cout); }
    plan->taps[mi] = out;
    return out;
  }

  Act attention(int mi, const Act& in) {
    const AttnW& a = ctx->attns.at(mi);
    const int C = a.c, H = in.H, W = in.W, L = H * W;
    Act out = new_act(C, H, W);
    AttentionArgs aa{};
    aa.x = in.p; aa.qs = in.qs; aa.gamma = a.gn_g; aa.beta = a.gn_b; aa.wqkv = a.wqkv; aa.bqkv = a.bqkv; aa.w3 = a.w3;
    aa.b3 = a.b3; aa.out = out.p; aa.qstats = out.qs; aa.scratch = scrQKV; aa.B = B; aa.L = L; aa.C = C;
    flowse_ctx* cx = ctx;
    // GroupNorm + q,k,v = NIN_0..2(h); softmax(q k^T C^-1/2) v; (x + NIN_3(h)) / sqrt(2) + output statistics: 2 kernels
    push(2, [=](cudaStream_t s) { std::string e; const int rc = launch_attention(aa, s, &e); if (rc) cx->err = e; return rc; }, 4,
         2.0 * B * L * (4.0 * C * C + 2.0 * L * C));
    plan->taps[mi] = out;
    return out;
  }

  int build() {
    const auto& mods = ctx->mods;
    const int H0 = kImage, W0 = T;
    const size_t HW = static_cast<size_t>(H0) * W0;
    plan->x = ar.alloc<float2>(B * HW); plan->y = ar.alloc<float2>(B * HW); plan->z = ar.alloc<float2>(B * HW);
    plan->vout = ar.alloc<float2>(B * HW); plan->xa = ar.alloc<float2>(B * HW);
    plan->va = ar.alloc<float2>(B * HW); plan->vb = ar.alloc<float2>(B * HW);
    plan->yp = ar.alloc<float2>(B * HW);
    plan->t_dev = ar.alloc<float>(B); plan->step_dev = ar.alloc<float>(4);
    temb_act = ar.alloc<float>(static_cast<size_t>(B) * kTemb);
    bias_table = ar.alloc<float>(static_cast<size_t>(B) * ctx->dense_rows);
    plan->bias_table = bias_table;
    plan->t_all = ar.alloc<float>(static_cast<size_t>(kMaxEvals) * B);
    plan->temb_all = ar.alloc<float>(static_cast<size_t>(kMaxEvals) * B * kTemb);
    plan->bias_all = ar.alloc<float>(static_cast<size_t>(kMaxEvals) * B * ctx->dense_rows);
    // scratch sized for the largest user (level 0, 256 concatenated input channels)
    const size_t top = static_cast<size_t>(B) * HW;
    scrA = ar.alloc<__half>(2 * top * 256);
    scrX = ar.alloc<__half>(2 * top * 256);
    scrH1 = ar.alloc<float>(top * 128);
    const int La = kAttnRes * (T / 16);           // tokens of the 16 x T/16 attention blocks
    scrQKV = ar.alloc<float>(attention_scratch_floats(B, La));
    splitk = ar.alloc<float>(kSplitKScratchElems);
    plan->stats_bytes = static_cast<size_t>(kMaxStatSlots) * B * kStatSlotDoubles * sizeof(double);
    plan->stats = ar.alloc<double>(plan->stats_bytes / sizeof(double));
    plan->gn_partials = ar.alloc<double>(static_cast<size_t>(B) * gn_stats_max_blocks() * 256);
    plan->gn_counters = ar.alloc<unsigned>(B);      // arena is zero-initialised; the kernel restores zero

    // ---- the walk (ncsnpp.py:247-404) ----
    {
      // the time embedding is a prologue outside the replayed op list (see Plan::temb_fn)
      TembWeights tw = ctx->temb; float* ta = plan->temb_all;
      if (!dry) plan->temb_fn = [=](const float* t_rows, int rows, float* table, cudaStream_t s) {
        launch_temb(tw, t_rows, rows, ta, table, s); return 0; };
      // producers accumulate quad statistics with atomics: clear all slots once per evaluation
      double* st = plan->stats; const size_t sb = plan->stats_bytes;
      push(0, [=](cudaStream_t s) { return cudaMemsetAsync(st, 0, sb, s) == cudaSuccess ? 0 : 1; }, 0);
    }
    std::vector<float4*> pin(kNumLevels);
    for (int l = 0; l < kNumLevels; ++l) pin[l] = ar.alloc<float4>(static_cast<size_t>(B) * (H0 >> l) * (W0 >> l));
    int m = 3;
    Act h0 = new_act(NF, H0, W0);
    {
      const float2 *px = plan->x, *py = plan->y; const float *w = ctx->conv_in_w, *bb = ctx->conv_in_b;
      float* o = h0.p; float4* p0 = pin[0]; const int Bc = B; double* q0 = h0.qs;
      push(1, [=](cudaStream_t s) { launch_conv_in(px, py, w, bb, o, p0, q0, Bc, H0, W0, s); return 0; }, 5);
    }
    plan->taps[3] = h0;
    ++m;
    // input pyramid (ncsnpp.py:310): the six FIR downsamples depend on nothing but the input conv's 4-plane copy, so they
    // are queued here as a side branch instead of one by one in front of their Combine
    for (int l = 0; l + 1 < kNumLevels; ++l) {
      const float4* src = pin[l]; float4* dst = pin[l + 1]; const int Bc = B, Hh = H0 >> (l + 1), Ww = W0 >> (l + 1);
      push(1, [=](cudaStream_t s) { launch_fir_down4(src, dst, Bc, Hh, Ww, s); return 0; }, 5);
      tag_last(2, l == 0);
    }
    std::vector<Act> hs = {h0};
    Act h = h0;
    for (int l = 0; l < kNumLevels; ++l) {
      for (int i = 0; i < kNumResBlocks; ++i) {
        h = resblock(m, hs.back(), nullptr); ++m;
        if (h.H == kAttnRes) { h = attention(m, h); ++m; }
        hs.push_back(h);
      }
      if (l != kNumLevels - 1) {
        h = resblock(m, hs.back(), nullptr); ++m;
        const CombW& cw = ctx->combs.at(m);
        Act o = new_act(cw.c, h.H, h.W);
        {
          float4* dst = pin[l + 1]; const int Bc = B, Hh = h.H, Ww = h.W, C = cw.c;
          const float* hp = h.p; float* op = o.p; const float *w = cw.w, *bb = cw.b; double* oq = o.qs;
          push(1, [=](cudaStream_t s) { launch_combine(hp, dst, w, bb, op, oq, Bc, Hh, Ww, C, s); return 0; }, 5);
          tag_last(0, true, 2);
        }
        plan->taps[m] = o;
        h = o; ++m;
        hs.push_back(h);
      }
    }
    h = hs.back();
    h = resblock(m, h, nullptr); ++m;
    h = attention(m, h); ++m;
    h = resblock(m, h, nullptr); ++m;
    float4* pyr_prev = nullptr;
    for (int l = kNumLevels - 1; l >= 0; --l) {
      for (int i = 0; i < kNumResBlocks + 1; ++i) {
        Act skip = hs.back(); hs.pop_back();
        h = resblock(m, h, &skip); ++m;
      }
      if (h.H == kAttnRes) { h = attention(m, h); ++m; }
      {
        const HeadW& hw = ctx->heads.at(m);
        const float* hp = h.p; const int C = h.C, Hh = h.H, Ww = h.W, Bc = B;
        float4* pyr = ar.alloc<float4>(static_cast<size_t>(B) * Hh * Ww);
        const float4* prev = pyr_prev; const double* hq = h.qs;
        const float *gg = hw.gn_g, *gb = hw.gn_b, *wf = hw.wf, *hb = hw.bias;
        // GN + SiLU + conv3x3(C -> 4) + FIR-up(previous level) in one fp32 SIMT kernel
        push(1, [=](cudaStream_t s) { launch_head_conv(hp, hq, gg, gb, wf, hb, prev, pyr, Bc, Hh, Ww, C, s); return 0; },
             5, 2.0 * Bc * Hh * Ww * 4 * 9.0 * C, Hh, Ww, 9 * C, 4);
        if (l != 0) tag_last(1, true);              // the full-resolution head is the last op before the final kernel anyway
        Act tp; tp.p = reinterpret_cast<float*>(pyr); tp.C = 4; tp.H = Hh; tp.W = Ww;
        plan->taps[m + 1] = tp;
        pyr_prev = pyr;
        m += 2;
      }
      if (l != 0) { h = resblock(m, h, nullptr); ++m; }
    }
    if (!hs.empty() || m != static_cast<int>(mods.size())) { ctx->err = "internal: module walk mismatch"; return 3; }
    if (stat_slots > kMaxStatSlots) { ctx->err = "internal: too many GroupNorm stat slots"; return 3; }
    plan->pyr_out = pyr_prev;
    return 0;
  }
};

void destroy_graphs(Plan* p) {
  if (p->graph_fwd) { cudaGraphExecDestroy(p->graph_fwd); p->graph_fwd = nullptr; }
  p->graph_ready = false;
  for (auto& g : p->sampler_graphs) if (g.exec) cudaGraphExecDestroy(g.exec);
  p->sampler_graphs.clear();
}

// `s` is the stream the caller is about to enqueue work on: the plan's zero-initialisation is ordered on it (a
// synchronous cudaMemset runs on the legacy NULL stream, which non-blocking streams are not ordered against).
int ensure_plan(flowse_ctx* ctx, int B, int T, cudaStream_t s) {
  if (!ctx->weights_loaded) { ctx->err = "weights not loaded (call flowse_load_weights first)"; return 2; }
  if (B <= 0 || T <= 0 || T % 64 != 0) { ctx->err = "need B > 0 and T a positive multiple of 64 (pad_spec)"; return 2; }
  if (ctx->plan && ctx->plan->B == B && ctx->plan->T == T) return 0;
  CK(cudaSetDevice(ctx->device));
  if (ctx->plan) {
    CK(cudaDeviceSynchronize());
    destroy_graphs(ctx->plan.get());
    ctx->plan.reset();          // the workspace belongs to the context and is reused
  }
  std::unique_ptr<Plan> plan(new Plan());
  plan->B = B; plan->T = T;
  {
    Builder dry{ctx, plan.get(), Arena{}, true, B, T};
    if (int rc = dry.build()) return rc;
    plan->arena_bytes = dry.ar.off + 4096;
    plan->taps.clear();
  }
  if (ctx->arena_cap < plan->arena_bytes) {
    if (ctx->arena) { CK(cudaFree(ctx->arena)); ctx->arena = nullptr; ctx->arena_cap = 0; }
    const size_t cap = plan->arena_bytes + plan->arena_bytes / 8;      // headroom: slightly larger buckets reuse it
    CK(cudaMalloc(reinterpret_cast<void**>(&ctx->arena), cap));
    ctx->arena_cap = cap;
  }
  plan->arena = ctx->arena;
  {
    Builder real{ctx, plan.get(), Arena{plan->arena, 0}, false, B, T};
    if (int rc = real.build()) return rc;
  }
  // only the statistics slots and the last-block counters need to start from zero (every evaluation clears the slots
  // again; the counters are restored to zero by their kernel).  Ordered on the caller's stream, then waited for, so a
  // later call on another stream cannot overtake it.
  CK(cudaMemsetAsync(plan->stats, 0, plan->stats_bytes, s));
  CK(cudaMemsetAsync(plan->gn_counters, 0, sizeof(unsigned) * B, s));
  CK(cudaStreamSynchronize(s));
  plan->kernels_per_forward = 0;
  for (const auto& op : plan->ops) plan->kernels_per_forward += op.nk;
  ctx->plan = std::move(plan);
  return 0;
}

// Enqueue the backbone ops on s.  fork (only while s is being captured): ops tagged with a side branch go to side streams
// that fork from / join s through events, i.e. they become parallel branches of the captured graph.
int run_ops(flowse_ctx* ctx, cudaStream_t s, bool fork = false) {
  if (fork && !ctx->ev_main) {
    for (auto& st : ctx->side) CK(cudaStreamCreateWithFlags(&st, cudaStreamNonBlocking));
    CK(cudaEventCreateWithFlags(&ctx->ev_main, cudaEventDisableTiming));
    for (int b = 1; b < 3; ++b) CK(cudaEventCreateWithFlags(&ctx->ev_branch[b], cudaEventDisableTiming));
  }
  bool used[3] = {false, false, false};
  for (auto& op : ctx->plan->ops) {
    int rc;
    if (!fork || op.branch == 0) {
      if (fork && op.wait_branch && used[op.wait_branch]) CK(cudaStreamWaitEvent(s, ctx->ev_branch[op.wait_branch], 0));
      rc = op.fn(s);
    } else {
      cudaStream_t sb = ctx->side[op.branch - 1];
      if (op.needs_main || !used[op.branch]) {          // (the first op of a branch joins the capture through this edge)
        CK(cudaEventRecord(ctx->ev_main, s));
        CK(cudaStreamWaitEvent(sb, ctx->ev_main, 0));
      }
      rc = op.fn(sb);
      CK(cudaEventRecord(ctx->ev_branch[op.branch], sb));
      used[op.branch] = true;
    }
    if (rc) { if (ctx->err.empty()) ctx->err = "plan op failed"; return rc; }
  }
  for (int b = 1; b < 3; ++b)
    if (used[b]) CK(cudaStreamWaitEvent(s, ctx->ev_branch[b], 0));      // join: the final kernel reads the output pyramid
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) { ctx->err = std::string("kernel launch failed: ") + cudaGetErrorString(e); return 1; }
  return 0;
}

// One network evaluation on the plan's fixed buffers (x, y, t_dev, bias_table) -> pyr_out.  inline_ops: enqueue the ops
// themselves even when a per-evaluation graph exists (used while a whole sampler call is being captured).
int run_backbone(flowse_ctx* ctx, cudaStream_t s, bool inline_ops = false) {
  Plan* p = ctx->plan.get();
  if (inline_ops) return run_ops(ctx, s, ctx->fork_branches != 0);      // s is being captured by the caller
  if (ctx->use_graph && p->eager_runs >= 1) {
    if (!p->graph_ready) {
      cudaGraph_t g = nullptr;
      if (!ctx->cap_stream) CK(cudaStreamCreateWithFlags(&ctx->cap_stream, cudaStreamNonBlocking));
      CK(cudaStreamBeginCapture(ctx->cap_stream, cudaStreamCaptureModeThreadLocal));
      const long long c0 = launch_counter();
      const int rc = run_ops(ctx, ctx->cap_stream, ctx->fork_branches != 0);
      p->kernels_per_forward = static_cast<int>(launch_counter() - c0);
      ctx->counter_base += p->kernels_per_forward;    // captured, not executed
      cudaError_t e = cudaStreamEndCapture(ctx->cap_stream, &g);
      if (rc) { if (g) cudaGraphDestroy(g); return rc; }
      if (e != cudaSuccess) { ctx->err = std::string("graph capture: ") + cudaGetErrorString(e); return 1; }
      e = cudaGraphInstantiate(&p->graph_fwd, g, 0);
      cudaGraphDestroy(g);
      if (e != cudaSuccess) { ctx->err = std::string("graph instantiate: ") + cudaGetErrorString(e); return 1; }
      p->graph_ready = true;
    }
    CK(cudaGraphLaunch(p->graph_fwd, s));
    ctx->graph_kernels += p->kernels_per_forward;
  } else {
    if (int rc = run_ops(ctx, s)) return rc;
    ++p->eager_runs;
  }
  return 0;
}

int final_op(flowse_ctx* ctx, int mode, const float2* xin, float2* out, cudaStream_t s) {
  Plan* p = ctx->plan.get();
  launch_final(p->pyr_out, p->t_dev, ctx->out_w, ctx->out_b, xin, p->step_dev, out, mode, p->B, kImage * p->T, s);
  return 0;
}

int set_t(flowse_ctx* ctx, float t, float step, cudaStream_t s) {
  // t and the step size travel as kernel arguments: no host buffer, no stream synchronisation
  Plan* p = ctx->plan.get();
  launch_set_scalars(p->t_dev, p->B, t, p->step_dev, step, s);
  return 0;
}

}  // namespace

// ================================================================================================
// C ABI
// ================================================================================================
extern "C" {

int flowse_create(flowse_ctx** out, int device) {
  if (!out) return 2;
  *out = nullptr;
  int n = 0;
  cudaError_t e = cudaGetDeviceCount(&n);
  if (e != cudaSuccess || n <= 0) {
    g_create_error = std::string("no CUDA device available: ") + cudaGetErrorString(e) +
                     " (libflowse has no CPU fallback)";
    return 1;
  }
  if (device < 0 || device >= n) { g_create_error = "invalid device ordinal"; return 2; }
  cudaDeviceProp prop;
  cudaGetDeviceProperties(&prop, device);
  if (prop.major != 10) {
    g_create_error = std::string("device '") + prop.name + "' is sm_" + std::to_string(prop.major) +
                     std::to_string(prop.minor) + "; libflowse is built for sm_100a (B200) only";
    return 1;
  }
  flowse_ctx* ctx = new flowse_ctx();
  ctx->device = device;
  ctx->mods = build_modules();
  ctx->counter_base = launch_counter();
  // A/B switches for measurements (the options of flowse_set_option, preset from the environment)
  if (const char* e = getenv("FLOWSE_FUSE_PREP")) ctx->fuse_prep = atoi(e);
  if (const char* e = getenv("FLOWSE_WHOLE_GRAPH")) ctx->whole_graph = atoi(e);
  if (const char* e = getenv("FLOWSE_FORK")) ctx->fork_branches = atoi(e);
  cudaSetDevice(device);
  if (cudaMalloc(reinterpret_cast<void**>(&ctx->overflow), sizeof(unsigned long long)) != cudaSuccess ||
      cudaMemset(ctx->overflow, 0, sizeof(unsigned long long)) != cudaSuccess) {
    g_create_error = "cudaMalloc failed in flowse_create";
    delete ctx;
    return 1;
  }
  *out = ctx;
  return 0;
}

void flowse_destroy(flowse_ctx* ctx) {
  if (!ctx) return;
  cudaSetDevice(ctx->device);
  cudaDeviceSynchronize();
  if (ctx->plan) destroy_graphs(ctx->plan.get());
  if (ctx->overflow) cudaFree(ctx->overflow);
  if (ctx->rk_acc) cudaFree(ctx->rk_acc);
  if (ctx->arena) cudaFree(ctx->arena);
  for (auto& p : ctx->dev_allocs) cudaFree(p.first);
  if (ctx->cap_stream) cudaStreamDestroy(ctx->cap_stream);
  for (auto& st : ctx->side) if (st) cudaStreamDestroy(st);
  if (ctx->ev_main) cudaEventDestroy(ctx->ev_main);
  for (auto& ev : ctx->ev_branch) if (ev) cudaEventDestroy(ev);
  if (ctx->op_stats) cudaFree(ctx->op_stats);
  if (ctx->op_partials) cudaFree(ctx->op_partials);
  if (ctx->op_counters) cudaFree(ctx->op_counters);
  if (ctx->op_scratch) cudaFree(ctx->op_scratch);
  if (ctx->op_splitk) cudaFree(ctx->op_splitk);
  if (ctx->stft_basis) cudaFree(ctx->stft_basis);
  if (ctx->stft_scratch) cudaFree(ctx->stft_scratch);
  delete ctx;
}

const char* flowse_last_error(const flowse_ctx* ctx) { return ctx ? ctx->err.c_str() : g_create_error.c_str(); }

int flowse_load_weights(flowse_ctx* ctx, const float* host_blob, const flowse_tensor_desc* descs, int n) {
  if (!ctx) return 2;
  ctx->err.clear();
  if (!host_blob || !descs || n <= 0) { ctx->err = "load_weights: null arguments"; return 2; }
  if (ctx->weights_loaded) { ctx->err = "weights already loaded; create a new context"; return 2; }
  CK(cudaSetDevice(ctx->device));
  HostBlob hb; hb.base = host_blob;
  for (int i = 0; i < n; ++i) {
    char nm[97]; std::memcpy(nm, descs[i].name, 96); nm[96] = 0;
    hb.idx[std::string(nm)] = {descs[i].offset, descs[i].numel};
  }
  if (int rc = load_weights_impl(ctx, hb)) return rc;
  CK(cudaDeviceSynchronize());
  ctx->weights_loaded = true;
  return 0;
}

// ---- packed weights: the context's device buffers - conv weights already in the K-major fp16 hi/lo
// format, small tensors in fp32 - as one relocatable blob, so a deployment ships / loads them without the fp32 checkpoint
// and without the host-side packing pass.
namespace {
constexpr unsigned long long kPackedMagic = 0x31304b5045534c46ull;    // "FLSEPK01"
struct PackedHeader { unsigned long long magic; unsigned version, n_seg, n_conv, reserved; };
size_t packed_payload_offset(unsigned n_seg, unsigned n_conv) {
  const size_t h = sizeof(PackedHeader) + sizeof(unsigned long long) * n_seg + sizeof(float) * n_conv;
  return (h + 255) & ~static_cast<size_t>(255);
}
}  // namespace

size_t flowse_packed_bytes(flowse_ctx* ctx) {
  if (!ctx || !ctx->weights_loaded) return 0;
  size_t n = packed_payload_offset(static_cast<unsigned>(ctx->dev_allocs.size()), static_cast<unsigned>(ctx->conv_scales.size()));
  for (auto& a : ctx->dev_allocs) n += (a.second + 255) & ~static_cast<size_t>(255);
  return n;
}

int flowse_export_packed(flowse_ctx* ctx, void* host_out, size_t bytes) {
  if (!ctx) return 2;
  ctx->err.clear();
  if (!ctx->weights_loaded) { ctx->err = "export_packed: weights not loaded"; return 2; }
  if (!host_out || bytes < flowse_packed_bytes(ctx)) { ctx->err = "export_packed: buffer smaller than flowse_packed_bytes()"; return 2; }
  CK(cudaSetDevice(ctx->device));
  char* out = static_cast<char*>(host_out);
  std::memset(out, 0, flowse_packed_bytes(ctx));
  PackedHeader h{kPackedMagic, 1u, static_cast<unsigned>(ctx->dev_allocs.size()), static_cast<unsigned>(ctx->conv_scales.size()), 0u};
  std::memcpy(out, &h, sizeof h);
  auto* seg = reinterpret_cast<unsigned long long*>(out + sizeof h);
  for (size_t i = 0; i < ctx->dev_allocs.size(); ++i) seg[i] = ctx->dev_allocs[i].second;
  std::memcpy(out + sizeof h + sizeof(unsigned long long) * h.n_seg, ctx->conv_scales.data(), sizeof(float) * h.n_conv);
  size_t off = packed_payload_offset(h.n_seg, h.n_conv);
  for (auto& a : ctx->dev_allocs) {
    CK(cudaMemcpy(out + off, a.first, a.second, cudaMemcpyDeviceToHost));
    off += (a.second + 255) & ~static_cast<size_t>(255);
  }
  return 0;
}

int flowse_load_packed(flowse_ctx* ctx, const void* host_blob, size_t bytes) {
  if (!ctx) return 2;
  ctx->err.clear();
  if (ctx->weights_loaded) { ctx->err = "weights already loaded; create a new context"; return 2; }
  if (!host_blob || bytes < sizeof(PackedHeader)) { ctx->err = "load_packed: blob too small"; return 2; }
  PackedHeader h;
  std::memcpy(&h, host_blob, sizeof h);
  if (h.magic != kPackedMagic || h.version != 1u) { ctx->err = "load_packed: not a libflowse packed-weights blob (magic / version)"; return 2; }
  const size_t payload = packed_payload_offset(h.n_seg, h.n_conv);
  if (h.n_seg > 100000u || h.n_conv > 100000u || bytes < payload) { ctx->err = "load_packed: truncated header"; return 2; }
  flowse_ctx::PackedSrc pk;
  pk.base = static_cast<const char*>(host_blob);
  const auto* seg = reinterpret_cast<const unsigned long long*>(pk.base + sizeof h);
  pk.seg.assign(seg, seg + h.n_seg);
  const auto* sc = reinterpret_cast<const float*>(pk.base + sizeof h + sizeof(unsigned long long) * h.n_seg);
  pk.scales.assign(sc, sc + h.n_conv);
  pk.seg_i = 0; pk.scale_i = 0; pk.off = payload;
  size_t need = payload;
  unsigned long long biggest = 0;
  for (auto v : pk.seg) { need += (v + 255) & ~static_cast<size_t>(255); biggest = std::max(biggest, v); }
  if (bytes < need) { ctx->err = "load_packed: truncated payload"; return 2; }
  CK(cudaSetDevice(ctx->device));
  std::vector<float> zeros(static_cast<size_t>(9) * 512 * 512, 0.f);    // larger than any tensor of the backbone
  HostBlob hb; hb.base = nullptr; hb.zeros = zeros.data();
  ctx->packed = &pk;
  const int rc = load_weights_impl(ctx, hb);
  ctx->packed = nullptr;
  if (rc) return rc;
  if (pk.seg_i != pk.seg.size() || pk.scale_i != pk.scales.size()) { ctx->err = "load_packed: blob has more segments than this library's layout"; return 2; }
  CK(cudaDeviceSynchronize());
  ctx->weights_loaded = true;
  return 0;
}

size_t flowse_workspace_bytes(flowse_ctx* ctx, int B, int T) {
  if (!ctx) return 0;
  ctx->err.clear();
  if (ensure_plan(ctx, B, T, nullptr)) return 0;
  return ctx->plan->arena_bytes;
}

int flowse_prior_sample(flowse_ctx* ctx, const void* y, const void* z, float sigma, void* x, long long n, void* stream) {
  if (!ctx) return 2;
  ctx->err.clear();
  if (n <= 0) return 0;
  launch_prior(static_cast<const float2*>(y), static_cast<const float2*>(z), sigma, static_cast<float2*>(x),
               static_cast<size_t>(n), static_cast<cudaStream_t>(stream));
  CK(cudaGetLastError());
  return 0;
}

int flowse_euler_step(flowse_ctx* ctx, const void* x, const void* v, float stepsize, void* x_out, long long n,
                      void* stream) {
  if (!ctx) return 2;
  ctx->err.clear();
  if (n <= 0) return 0;
  launch_euler_update(static_cast<const float2*>(x), static_cast<const float2*>(v), -stepsize,
                      static_cast<float2*>(x_out), static_cast<size_t>(n), static_cast<cudaStream_t>(stream));
  CK(cudaGetLastError());
  return 0;
}

int flowse_ncsnpp_forward(flowse_ctx* ctx, const void* x, long long x_bstride, const void* y, long long y_bstride,
                          const float* t, void* out, int negate, int B, int T, void* stream) {
  if (!ctx) return 2;
  ctx->err.clear();
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  if (int rc = ensure_plan(ctx, B, T, s)) return rc;
  Plan* p = ctx->plan.get();
  const size_t row = static_cast<size_t>(kImage) * T * sizeof(float2);
  CK(cudaMemcpy2DAsync(p->x, row, x, x_bstride * sizeof(float2), row, B, cudaMemcpyDeviceToDevice, s));
  CK(cudaMemcpy2DAsync(p->y, row, y, y_bstride * sizeof(float2), row, B, cudaMemcpyDeviceToDevice, s));
  CK(cudaMemcpyAsync(p->t_dev, t, sizeof(float) * B, cudaMemcpyDeviceToDevice, s));
  if (int rc = p->temb_fn(p->t_dev, B, p->bias_table, s)) return rc;
  if (int rc = run_backbone(ctx, s)) return rc;
  final_op(ctx, negate ? 1 : 0, nullptr, static_cast<float2*>(out), s);
  CK(cudaGetLastError());
  return 0;
}

}  // extern "C"

namespace {

// Times at which a sampler call evaluates the network, in order (sampling/__init__.py:48-57 for Euler; SURVEY.md 8 A4 for
// Heun / midpoint with an Euler step on the last interval).  fp32 arithmetic, as the reference's tensors.
std::vector<float> evaluation_times(const float* ts, int N, int solver) {
  std::vector<float> ev;
  for (int i = 0; i < N; ++i) {
    const float t = ts[i];
    const float step = (i != N - 1) ? (t - ts[i + 1]) : ts[N - 1];
    ev.push_back(t);
    if (solver != FLOWSE_SOLVER_EULER && i != N - 1) {
      const float dt = -step;
      ev.push_back(solver == FLOWSE_SOLVER_HEUN ? t + dt : t + dt / 2);
    }
  }
  return ev;
}

// Everything of a sampler call between "y, z (and yp) are in the plan's buffers" and "the result is in p->x", enqueued
// on s.  inline_ops: see run_backbone.
int enqueue_sampler(flowse_ctx* ctx, const float* ts, int N, int solver, float sigma, bool own_prior, cudaStream_t s,
                    bool inline_ops) {
  Plan* p = ctx->plan.get();
  const int B = p->B;
  const size_t n = static_cast<size_t>(B) * kImage * p->T;
  const std::vector<float> ev = evaluation_times(ts, N, solver);
  const int E = static_cast<int>(ev.size());
  const size_t table_floats = static_cast<size_t>(B) * ctx->dense_rows;
  int e = 0;                                     // index of the next evaluation
  // time embeddings of up to kMaxEvals evaluations per launch pair, then one 4*B*R-byte copy per evaluation
  auto begin_eval = [&](float t, float step) -> int {
    if (e >= E || ev[e] != t) { ctx->err = "internal: evaluation schedule mismatch"; return 3; }
    if (e % kMaxEvals == 0) {
      const int cnt = std::min(kMaxEvals, E - e);
      launch_set_times(p->t_all, B, ev.data() + e, cnt, s);
      if (int rc = p->temb_fn(p->t_all, cnt * B, p->bias_all, s)) return rc;
    }
    CK(cudaMemcpyAsync(p->bias_table, p->bias_all + static_cast<size_t>(e % kMaxEvals) * table_floats,
                       table_floats * sizeof(float), cudaMemcpyDeviceToDevice, s));
    ++e;
    return set_t(ctx, t, step, s);
  };
  launch_prior(own_prior ? p->yp : p->y, p->z, sigma, p->x, n, s);         // x_T = y_prior + sigma z
  for (int i = 0; i < N; ++i) {
    const float t = ts[i];
    const float step = (i != N - 1) ? (t - ts[i + 1]) : ts[N - 1];        // fp32, as sampling/__init__.py:50-53
    const bool last = (i == N - 1);
    if (int rc = begin_eval(t, step)) return rc;
    if (solver == FLOWSE_SOLVER_EULER || last) {
      if (int rc = run_backbone(ctx, s, inline_ops)) return rc;
      final_op(ctx, 2, p->x, p->x, s);                                      // x += step * dnn(x, y, t)
    } else if (solver == FLOWSE_SOLVER_HEUN) {
      // v0 = VF(x,t); x_next = x + dt v0; x = x + dt/2 (v0 + VF(x_next, t+dt)),  dt = -step, VF = -dnn
      const float dt = -step;
      if (int rc = run_backbone(ctx, s, inline_ops)) return rc;
      final_op(ctx, 1, nullptr, p->va, s);                                  // va = v0
      CK(cudaMemcpyAsync(p->xa, p->x, n * sizeof(float2), cudaMemcpyDeviceToDevice, s));
      launch_axpy_c(p->xa, p->va, dt, p->x, n, s);                          // x := x_next (network input)
      if (int rc = begin_eval(t + dt, step)) return rc;
      if (int rc = run_backbone(ctx, s, inline_ops)) return rc;
      final_op(ctx, 1, nullptr, p->vb, s);                                  // vb = VF(x_next, t+dt)
      launch_heun_combine(p->xa, p->va, p->vb, dt / 2, p->x, n, s);
    } else {
      // x = x + dt VF(x + dt/2 VF(x,t), t + dt/2)
      const float dt = -step;
      if (int rc = run_backbone(ctx, s, inline_ops)) return rc;
      final_op(ctx, 1, nullptr, p->va, s);
      CK(cudaMemcpyAsync(p->xa, p->x, n * sizeof(float2), cudaMemcpyDeviceToDevice, s));
      launch_axpy_c(p->xa, p->va, dt / 2, p->x, n, s);
      if (int rc = begin_eval(t + dt / 2, step)) return rc;
      if (int rc = run_backbone(ctx, s, inline_ops)) return rc;
      final_op(ctx, 1, nullptr, p->vb, s);
      launch_axpy_c(p->xa, p->vb, dt, p->x, n, s);
    }
  }
  return 0;
}

constexpr int kWholeGraphMaxEvals = 12;      // ~240 kernel nodes per evaluation: keep instantiation in the low milliseconds
constexpr size_t kMaxSamplerGraphs = 4;

}  // namespace

extern "C" {

int flowse_sample(flowse_ctx* ctx, const void* y, const void* y_prior, const void* z, const float* ts, int N, int solver, float sigma,
                  void* x_out, int B, int T, void* stream) {
  if (!ctx) return 2;
  ctx->err.clear();
  if (N <= 0 || !ts) { ctx->err = "sample: need N >= 1 timesteps"; return 2; }
  if (solver < 0 || solver > 2) { ctx->err = "ODEsolver unknown (0 euler, 1 heun, 2 midpoint)"; return 2; }
  for (int i = 0; i < N; ++i)
    if (!(ts[i] > 0.f)) { ctx->err = "sample: timesteps must be > 0 (the backbone takes log t and divides by t)"; return 2; }
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  if (int rc = ensure_plan(ctx, B, T, s)) return rc;
  Plan* p = ctx->plan.get();
  const size_t n = static_cast<size_t>(B) * kImage * T;
  const bool own_prior = y_prior != nullptr && y_prior != y;
  CK(cudaMemcpyAsync(p->y, y, n * sizeof(float2), cudaMemcpyDeviceToDevice, s));
  CK(cudaMemcpyAsync(p->z, z, n * sizeof(float2), cudaMemcpyDeviceToDevice, s));
  if (own_prior) CK(cudaMemcpyAsync(p->yp, y_prior, n * sizeof(float2), cudaMemcpyDeviceToDevice, s));

  // Whole call as ONE graph launch: the second call with a given (schedule, solver, sigma) on a warm plan captures
  // prior + every evaluation + the fused updates; later calls replay it.
  const int E = (solver == FLOWSE_SOLVER_EULER) ? N : 2 * N - 1;
  bool done = false;
  if (ctx->use_graph && ctx->whole_graph && p->eager_runs >= 1 && E <= kWholeGraphMaxEvals) {
    Plan::SamplerGraph* g = nullptr;
    for (auto& c : p->sampler_graphs)
      if (c.solver == solver && c.sigma == sigma && c.own_prior == own_prior && static_cast<int>(c.ts.size()) == N &&
          std::memcmp(c.ts.data(), ts, sizeof(float) * N) == 0) { g = &c; break; }
    if (!g) {
      if (p->sampler_graphs.size() >= kMaxSamplerGraphs) {
        if (p->sampler_graphs.front().exec) cudaGraphExecDestroy(p->sampler_graphs.front().exec);
        p->sampler_graphs.erase(p->sampler_graphs.begin());
      }
      p->sampler_graphs.push_back(Plan::SamplerGraph{std::vector<float>(ts, ts + N), solver, sigma, own_prior, nullptr, 0, 1});
    } else {
      if (!g->exec) {
        cudaGraph_t cg = nullptr;
        if (!ctx->cap_stream) CK(cudaStreamCreateWithFlags(&ctx->cap_stream, cudaStreamNonBlocking));
        CK(cudaStreamBeginCapture(ctx->cap_stream, cudaStreamCaptureModeThreadLocal));
        const long long c0 = launch_counter();
        const int rc = enqueue_sampler(ctx, ts, N, solver, sigma, own_prior, ctx->cap_stream, true);
        g->kernels = launch_counter() - c0;
        ctx->counter_base += g->kernels;              // captured, not executed
        cudaError_t ce = cudaStreamEndCapture(ctx->cap_stream, &cg);
        if (rc) { if (cg) cudaGraphDestroy(cg); return rc; }
        if (ce != cudaSuccess) { ctx->err = std::string("sampler graph capture: ") + cudaGetErrorString(ce); return 1; }
        ce = cudaGraphInstantiate(&g->exec, cg, 0);
        cudaGraphDestroy(cg);
        if (ce != cudaSuccess) { g->exec = nullptr; ctx->err = std::string("sampler graph instantiate: ") + cudaGetErrorString(ce); return 1; }
      }
      CK(cudaGraphLaunch(g->exec, s));
      ctx->graph_kernels += g->kernels;
      done = true;
    }
  }
  if (!done)
    if (int rc = enqueue_sampler(ctx, ts, N, solver, sigma, own_prior, s, false)) return rc;
  CK(cudaMemcpyAsync(x_out, p->x, n * sizeof(float2), cudaMemcpyDeviceToDevice, s));
  CK(cudaGetLastError());
  return 0;
}

int flowse_rk_lincomb(flowse_ctx* ctx, const void* base64, const void* K32, long long k_stride, const double* coef_host, int S,
                      void* out64, void* out32, const void* ya64, const void* yb64, double rtol, double atol,
                      double* sumsq_host, long long n, void* stream) {
  if (!ctx) return 2;
  ctx->err.clear();
  if (n <= 0) return 0;
  if (S < 0 || S > 8 || (S > 0 && (!K32 || !coef_host))) { ctx->err = "rk_lincomb: 0 <= S <= 8 stages with K and coefficients"; return 2; }
  if (sumsq_host && (!ya64 || !yb64)) { ctx->err = "rk_lincomb: the scaled norm needs ya and yb"; return 2; }
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  CK(cudaSetDevice(ctx->device));
  if (sumsq_host) {
    if (!ctx->rk_acc) CK(cudaMalloc(reinterpret_cast<void**>(&ctx->rk_acc), sizeof(double)));
    CK(cudaMemsetAsync(ctx->rk_acc, 0, sizeof(double), s));
  }
  launch_rk_lincomb(static_cast<const double2*>(base64), static_cast<const float2*>(K32), k_stride, coef_host, S,
                    static_cast<double2*>(out64), static_cast<float2*>(out32), static_cast<const double2*>(ya64),
                    static_cast<const double2*>(yb64), rtol, atol, sumsq_host ? ctx->rk_acc : nullptr, static_cast<size_t>(n), s);
  CK(cudaGetLastError());
  if (sumsq_host) {       // the step-size controller needs this number on the host: the one synchronisation per RK step
    CK(cudaMemcpyAsync(sumsq_host, ctx->rk_acc, sizeof(double), cudaMemcpyDeviceToHost, s));
    CK(cudaStreamSynchronize(s));
  }
  return 0;
}

int flowse_fp16_overflow(flowse_ctx* ctx, long long* count, int reset) {
  if (!ctx) return 2;
  ctx->err.clear();
  if (!count) { ctx->err = "fp16_overflow: null count"; return 2; }
  *count = 0;
  if (!ctx->overflow) return 0;
  CK(cudaSetDevice(ctx->device));
  CK(cudaDeviceSynchronize());
  unsigned long long v = 0;
  CK(cudaMemcpy(&v, ctx->overflow, sizeof(v), cudaMemcpyDeviceToHost));
  *count = static_cast<long long>(v);
  if (reset) CK(cudaMemset(ctx->overflow, 0, sizeof(v)));
  return 0;
}

int flowse_profile_forward(flowse_ctx* ctx, int max_ops, int* kinds, float* ms, double* flops, int* info, int* n_ops) {
  if (!ctx) return 2;
  ctx->err.clear();
  if (!ctx->plan) { ctx->err = "no plan yet: run a forward or a sampler first"; return 2; }
  Plan* p = ctx->plan.get();
  const int n = static_cast<int>(p->ops.size());
  if (n > max_ops) { ctx->err = "profile_forward: buffer too small"; return 2; }
  CK(cudaSetDevice(ctx->device));
  cudaStream_t s = nullptr;
  CK(cudaStreamCreateWithFlags(&s, cudaStreamNonBlocking));
  std::vector<cudaEvent_t> ev(n + 1);
  for (auto& e : ev) CK(cudaEventCreate(&e));
  CK(cudaDeviceSynchronize());
  // the whole evaluation is queued behind an 8 ms spin so that the kernels run back to back and the events between them
  // measure device time, not the host's launch cadence (tensor-map encoding + launch is ~5 us per conv on the host)
  const long long c0 = launch_counter();
  launch_spin(8000000ull, s);
  CK(cudaEventRecord(ev[0], s));
  for (int i = 0; i < n; ++i) {
    if (int rc = p->ops[i].fn(s)) return rc;
    CK(cudaEventRecord(ev[i + 1], s));
  }
  CK(cudaStreamSynchronize(s));
  (void)c0;
  for (int i = 0; i < n; ++i) {
    CK(cudaEventElapsedTime(&ms[i], ev[i], ev[i + 1]));
    kinds[i] = p->ops[i].kind; flops[i] = p->ops[i].flops;
    for (int k = 0; k < 4; ++k) info[4 * i + k] = p->ops[i].info[k];
  }
  *n_ops = n;
  for (auto& e : ev) cudaEventDestroy(e);
  cudaStreamDestroy(s);
  return 0;
}

int flowse_set_option(flowse_ctx* ctx, const char* key, int value) {
  if (!ctx || !key) return 2;
  ctx->err.clear();
  const std::string k(key);
  bool replan = false;            // the op list depends on the option: rebuild the plan on the next call
  if (k == "conv_impl") { ctx->conv_impl = value; replan = true; }
  else if (k == "fuse_prep") { ctx->fuse_prep = value; replan = true; }
  else if (k == "graph") ctx->use_graph = value;
  else if (k == "pdl") pdl_mode() = static_cast<int>(value);
  else if (k == "whole_graph") ctx->whole_graph = value;
  else if (k == "fork") ctx->fork_branches = value;
  else if (k == "stft_window") {
    if (value < 0 || value > 1) { ctx->err = "stft_window: 0 (hann) or 1 (sqrthann)"; return 2; }
    ctx->stft_window = value;   // the bases are rebuilt on the next STFT call
    return 0;
  } else if (k == "spec_transform") {
    if (value < 0 || value > 2) { ctx->err = "spec_transform: 0 (exponent), 1 (log) or 2 (none)"; return 2; }
    ctx->spec_transform = value;
    return 0;
  }
  else { ctx->err = "unknown option '" + k + "'"; return 2; }
  if (ctx->plan) {   // captured graphs bake the old setting
    cudaSetDevice(ctx->device);
    cudaDeviceSynchronize();
    destroy_graphs(ctx->plan.get());
    if (replan) ctx->plan.reset();
  }
  return 0;
}

long long flowse_kernel_launches(const flowse_ctx* ctx) {
  return ctx ? (launch_counter() - ctx->counter_base + ctx->graph_kernels) : 0;
}

int flowse_debug_tap(flowse_ctx* ctx, int module_idx, const float** ptr, int* C, int* H, int* W) {
  if (!ctx) return 2;
  ctx->err.clear();
  if (!ctx->plan) { ctx->err = "no plan yet"; return 2; }
  auto it = ctx->plan->taps.find(module_idx);
  if (it == ctx->plan->taps.end()) { ctx->err = "no tap for module " + std::to_string(module_idx); return 2; }
  *ptr = it->second.p; *C = it->second.C; *H = it->second.H; *W = it->second.W;
  return 0;
}

static int ensure_op_stats(flowse_ctx* ctx) {
  if (ctx->op_stats) return 0;
  CK(cudaMalloc(reinterpret_cast<void**>(&ctx->op_stats), 2 * 64 * kStatSlotDoubles * sizeof(double)));   // two sources
  CK(cudaMemset(ctx->op_stats, 0, 2 * 64 * kStatSlotDoubles * sizeof(double)));   // quad_stats fills replica 0 only
  CK(cudaMalloc(reinterpret_cast<void**>(&ctx->op_partials), static_cast<size_t>(64) * gn_stats_max_blocks() * 256 * sizeof(double)));
  CK(cudaMalloc(reinterpret_cast<void**>(&ctx->op_counters), 64 * sizeof(unsigned)));
  CK(cudaMemset(ctx->op_counters, 0, 64 * sizeof(unsigned)));
  return 0;
}

int flowse_debug_copy(flowse_ctx* ctx, const void* src, void* dst, size_t bytes) {
  if (!ctx) return 2;
  ctx->err.clear();
  CK(cudaDeviceSynchronize());
  CK(cudaMemcpy(dst, src, bytes, cudaMemcpyDeviceToDevice));
  return 0;
}

int flowse_pack_conv_weights(const float* w_main_host, int Cout, int Cin, int ntaps, const float* w_sc_host, int Cin2,
                             int Npad, void* dev_out, int* wexp) {
  const int K = ntaps * Cin + (w_sc_host ? Cin2 : 0);
  std::vector<__half> buf(static_cast<size_t>(2) * Npad * K);
  const int e = pack_conv_weights_host(w_main_host, Cout, Cin, ntaps, w_sc_host, Cin2, Npad, buf.data(),
                                       buf.data() + static_cast<size_t>(Npad) * K);
  if (wexp) *wexp = e;
  return cudaMemcpy(dev_out, buf.data(), buf.size() * sizeof(__half), cudaMemcpyHostToDevice) == cudaSuccess ? 0 : 1;
}

int flowse_op_gn_prep(flowse_ctx* ctx, const float* src1, int C1, const float* src2, int C2, const float* gamma,
                      const float* beta, int B, int H, int W, int mode, int silu, void* outA, void* outX, float* outF,
                      float* outXF, void* stream) {
  if (!ctx) return 2;
  ctx->err.clear();
  const int C = C1 + (src2 ? C2 : 0);
  if (C % 128 != 0 || C > 1024 || C1 % 4 != 0) { ctx->err = "gn_prep: channel count must be a multiple of 128 (<= 1024)"; return 2; }
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  if (B > 64) { ctx->err = "gn_prep op: B <= 64"; return 2; }
  if (int rc = ensure_op_stats(ctx)) return rc;
  double* qs2 = ctx->op_stats + static_cast<size_t>(64) * kStatSlotDoubles;
  // the slot layout depends on the channel count: clear what an earlier call with another shape left in the replicas
  CK(cudaMemsetAsync(ctx->op_stats, 0, 2 * 64 * kStatSlotDoubles * sizeof(double), s));
  launch_quad_stats(src1, C1, B, H * W, ctx->op_stats, ctx->op_partials, ctx->op_counters, s);
  if (src2) launch_quad_stats(src2, C2, B, H * W, qs2, ctx->op_partials, ctx->op_counters, s);
  PrepArgs pa{};
  pa.src1 = src1; pa.C1 = C1; pa.src2 = src2; pa.C2 = C2; pa.qs1 = ctx->op_stats; pa.qs2 = qs2; pa.gamma = gamma; pa.beta = beta;
  pa.B = B; pa.H = H; pa.W = W; pa.mode = mode; pa.silu = silu;
  pa.outA = static_cast<__half*>(outA); pa.outX = static_cast<__half*>(outX); pa.outF = outF; pa.outXF = outXF;
  pa.overflow = ctx->overflow;
  launch_gn_prep(pa, s);
  CK(cudaGetLastError());
  return 0;
}

int flowse_op_conv_gemm(flowse_ctx* ctx, const void* A, int Cin, int ntaps, const void* X, int Cin2, const void* Wp,
                        int Npad, int wexp, const float* bias, int bias_bstride, const float* residual, int div_sqrt2,
                        float* out, int Cout, int ldc, int B, int H, int W, int impl, void* stream) {
  if (!ctx) return 2;
  ctx->err.clear();
  ConvGemmArgs a{};
  a.A = static_cast<const __half*>(A); a.Cin = Cin; a.ntaps = ntaps; a.X = static_cast<const __half*>(X); a.Cin2 = Cin2;
  a.Wp = static_cast<const __half*>(Wp); a.Npad = Npad; a.wscale_inv = std::ldexp(1.0f, -wexp); a.bias = bias;
  a.bias_bstride = bias_bstride; a.residual = residual; a.div_sqrt2 = div_sqrt2; a.out = out; a.Cout = Cout; a.ldc = ldc;
  a.B = B; a.H = H; a.W = W;
  if (!ctx->op_splitk) {
    CK(cudaMalloc(reinterpret_cast<void**>(&ctx->op_splitk), kSplitKScratchElems * sizeof(float)));
  }
  a.splitk_scratch = ctx->op_splitk; a.splitk_scratch_elems = kSplitKScratchElems;
  std::string e;
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  int rc;
  if (impl == 1) rc = launch_conv_gemm_simt(a, st, &e);
  else if (impl >= 2 && impl <= 4) rc = launch_conv_halo(a, impl == 3 ? 3 : (impl == 4 ? 2 : 1), st, &e);   // halo kernel
  else rc = launch_conv_gemm(a, st, &e);
  if (rc) ctx->err = e;
  return rc;
}

// ---- STFT / iSTFT ------------------------------------------------------------------------------
namespace {
struct StftScratch { int* lengths; unsigned* peak; float* xpad; long long xpad_stride; float2* S; float* frames; };

int stft_prepare(flowse_ctx* ctx, const int* lengths_host, int B, int min_frames, int* Lmax_out, StftScratch* sc, cudaStream_t s) {
  if (!lengths_host || B <= 0) { ctx->err = "stft: need B > 0 and a host lengths array"; return 2; }
  int Lmax = 0;
  for (int b = 0; b < B; ++b) {
    if (lengths_host[b] < 256) { ctx->err = "stft: every utterance needs >= 256 samples (reflect padding of n_fft/2 = 255)"; return 2; }
    Lmax = std::max(Lmax, lengths_host[b]);
  }
  CK(cudaSetDevice(ctx->device));
  if (!ctx->stft_basis) CK(cudaMalloc(reinterpret_cast<void**>(&ctx->stft_basis), stft_basis_floats() * sizeof(float)));
  if (ctx->stft_window_built != ctx->stft_window) {
    CK(cudaDeviceSynchronize());       // a basis of the other window may still be in use on another stream
    launch_stft_basis(ctx->stft_basis, ctx->stft_window, s);
    CK(cudaStreamSynchronize(s));      // built once per window: later calls may come on other streams
    ctx->stft_window_built = ctx->stft_window;
  }
  const int Tmax = std::max(stft_frames(Lmax), min_frames);
  const long long xs = ((static_cast<long long>(Lmax) + 510 + 128 + 127) / 128) * 128;
  auto up = [](size_t v) { return (v + 255) & ~static_cast<size_t>(255); };
  const size_t o_len = 0, o_peak = up(o_len + sizeof(int) * B), o_xpad = up(o_peak + sizeof(unsigned) * B);
  const size_t o_S = up(o_xpad + sizeof(float) * B * xs), o_fr = up(o_S + sizeof(float2) * B * Tmax * 256);
  const size_t need = up(o_fr + sizeof(float) * B * Tmax * 512);
  if (ctx->stft_scratch_bytes < need) {
    CK(cudaDeviceSynchronize());
    if (ctx->stft_scratch) cudaFree(ctx->stft_scratch);
    CK(cudaMalloc(reinterpret_cast<void**>(&ctx->stft_scratch), need));
    ctx->stft_scratch_bytes = need;
  }
  char* base = ctx->stft_scratch;
  sc->lengths = reinterpret_cast<int*>(base + o_len); sc->peak = reinterpret_cast<unsigned*>(base + o_peak);
  sc->xpad = reinterpret_cast<float*>(base + o_xpad); sc->xpad_stride = xs;
  sc->S = reinterpret_cast<float2*>(base + o_S); sc->frames = reinterpret_cast<float*>(base + o_fr);
  CK(cudaMemcpyAsync(sc->lengths, lengths_host, sizeof(int) * B, cudaMemcpyHostToDevice, s));
  *Lmax_out = Lmax;
  return 0;
}
}  // namespace

int flowse_stft_spec(flowse_ctx* ctx, const float* wav, long long wav_stride, const int* lengths_host, int B, int normalize,
                     float spec_factor, float abs_exponent, void* Y, int Tpad, float* peak_out, void* stream) {
  if (!ctx) return 2;
  ctx->err.clear();
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  int Lmax = 0; StftScratch sc{};
  if (int rc = stft_prepare(ctx, lengths_host, B, 0, &Lmax, &sc, s)) return rc;
  if (!wav || !Y || wav_stride < Lmax) { ctx->err = "stft_spec: null pointer or wav_stride < longest utterance"; return 2; }
  if (Tpad < stft_frames(Lmax)) { ctx->err = "stft_spec: Tpad is smaller than the frame count of the longest utterance"; return 2; }
  if (normalize && !peak_out) { ctx->err = "stft_spec: normalize needs peak_out"; return 2; }
  if (!(spec_factor > 0.f) || !(abs_exponent > 0.f)) { ctx->err = "stft_spec: spec_factor and abs_exponent must be > 0"; return 2; }
  // the peaks are accumulated as bit patterns directly in the caller's buffer
  launch_stft_spec(ctx->stft_basis, wav, wav_stride, sc.lengths, B, Lmax, normalize != 0, spec_factor, abs_exponent,
                   ctx->spec_transform, sc.xpad, sc.xpad_stride, sc.S, reinterpret_cast<unsigned*>(peak_out), static_cast<float2*>(Y), Tpad, s);
  CK(cudaGetLastError());
  return 0;
}

int flowse_spec_istft(flowse_ctx* ctx, const void* X, int Tpad, const int* lengths_host, int B, float spec_factor,
                      float abs_exponent, const float* peak, float* wav_out, long long wav_stride, void* stream) {
  if (!ctx) return 2;
  ctx->err.clear();
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  int Lmax = 0; StftScratch sc{};
  if (int rc = stft_prepare(ctx, lengths_host, B, Tpad, &Lmax, &sc, s)) return rc;
  if (!X || !wav_out || wav_stride < Lmax) { ctx->err = "spec_istft: null pointer or wav_stride < longest utterance"; return 2; }
  if (Tpad < stft_frames(Lmax)) { ctx->err = "spec_istft: Tpad is smaller than the frame count of the longest utterance"; return 2; }
  if (!(spec_factor > 0.f) || !(abs_exponent > 0.f)) { ctx->err = "spec_istft: spec_factor and abs_exponent must be > 0"; return 2; }
  launch_spec_istft(ctx->stft_basis, static_cast<const float2*>(X), Tpad, sc.lengths, B, Lmax, spec_factor, abs_exponent,
                    ctx->spec_transform, peak, sc.S, sc.frames, wav_out, wav_stride, s);
  CK(cudaGetLastError());
  return 0;
}

int flowse_op_head_conv(flowse_ctx* ctx, const float* h, const float* gamma, const float* beta, const float* wf,
                        const float* bias, const void* prev, void* out, int B, int H, int W, int C, void* stream) {
  if (!ctx) return 2;
  ctx->err.clear();
  if (C % 128 != 0 || C > 256 || B > 64) { ctx->err = "head_conv op: C in {128, 256}, B <= 64"; return 2; }
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  if (int rc = ensure_op_stats(ctx)) return rc;
  CK(cudaMemsetAsync(ctx->op_stats, 0, 2 * 64 * kStatSlotDoubles * sizeof(double), s));
  launch_quad_stats(h, C, B, H * W, ctx->op_stats, ctx->op_partials, ctx->op_counters, s);
  launch_head_conv(h, ctx->op_stats, gamma, beta, wf, bias, static_cast<const float4*>(prev), static_cast<float4*>(out), B, H,
                   W, C, s);
  CK(cudaGetLastError());
  return 0;
}

int flowse_op_attention(flowse_ctx* ctx, int module_idx, const float* x, float* out, int B, int H, int W, void* stream) {
  if (!ctx) return 2;
  ctx->err.clear();
  if (!ctx->weights_loaded) { ctx->err = "weights not loaded"; return 2; }
  auto it = ctx->attns.find(module_idx);
  if (it == ctx->attns.end()) { ctx->err = "module is not an attention block"; return 2; }
  const AttnW& a = it->second;
  const int C = a.c, L = H * W;
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  const size_t need = attention_scratch_floats(B, L) * sizeof(float);
  if (ctx->op_scratch_bytes < need) {
    CK(cudaDeviceSynchronize());
    if (ctx->op_scratch) cudaFree(ctx->op_scratch);
    CK(cudaMalloc(reinterpret_cast<void**>(&ctx->op_scratch), need));
    ctx->op_scratch_bytes = need;
  }
  if (B > 64) { ctx->err = "attention op: B <= 64"; return 2; }
  if (int rc = ensure_op_stats(ctx)) return rc;
  CK(cudaMemsetAsync(ctx->op_stats, 0, 2 * 64 * kStatSlotDoubles * sizeof(double), s));
  launch_quad_stats(x, C, B, L, ctx->op_stats, ctx->op_partials, ctx->op_counters, s);
  AttentionArgs aa{};
  aa.x = x; aa.qs = ctx->op_stats; aa.gamma = a.gn_g; aa.beta = a.gn_b; aa.wqkv = a.wqkv; aa.bqkv = a.bqkv; aa.w3 = a.w3;
  aa.b3 = a.b3; aa.out = out; aa.qstats = nullptr; aa.scratch = ctx->op_scratch; aa.B = B; aa.L = L; aa.C = C;
  std::string e;
  if (int rc = launch_attention(aa, s, &e)) { ctx->err = e; return rc; }
  CK(cudaGetLastError());
  return 0;
}

}  // extern "C"
