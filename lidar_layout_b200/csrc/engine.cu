// Host engine + C ABI (include/lidm_b200.h): weight store/packing, static activation planning, the U-Net and
// decoder launch plans, the on-device DDIM loop.  Mirrors, for the unconditional LiDM sampling path,
//   UNetModel.forward                 reference lidm/modules/diffusion/openaimodel.py:719-751
//   ResBlock / AttentionBlock         openaimodel.py:256-276 / 320-326
//   Decoder.forward, ResnetBlock, AttnBlock, Upsample   lidm/modules/diffusion/model_lidm.py:385-417,127-147,184-208,57-61
//   VQModelInterface.decode           lidm/models/ae/autoencoder.py:290-302
//   DDIMSampler.ddim_sampling         lidm/models/diffusion/ddim.py:115-165
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

#include "../../include/lidm_b200.h"
#include "common.h"

namespace lidm {

std::atomic<int64_t> g_launch_count{0};

namespace {

thread_local std::string tls_error;

struct Op {
  std::function<void(cudaStream_t)> fn;
  int cat;
  double flops, bytes;
  std::string label;
};

// ---- optional per-op CUDA-event profiler (bench.py roofline) ----
struct ProfState {
  bool on = false;
  std::vector<cudaEvent_t> ev;            // pairs
  std::vector<int> cat;
  std::vector<std::string> label;
  std::vector<double> op_flops, op_bytes;
  double flops[PROF_NCAT] = {0, 0, 0, 0}, bytes[PROF_NCAT] = {0, 0, 0, 0};
  int64_t launches[PROF_NCAT] = {0, 0, 0, 0};
} g_prof;

double gemm_flops(const View& a, int ntaps, int N) { return 2.0 * a.B * a.H * a.W * (double)N * ntaps * a.C; }

inline size_t align_up(size_t v, size_t a) { return (v + a - 1) / a * a; }

// ---------------------------------------------------------------------------------------------------------
// Plan-time arena allocator (first fit + coalescing).  Buffers are assigned once when a plan is built; stream
// order makes reuse after `release` safe.
class ArenaPlanner {
 public:
  size_t alloc(size_t bytes) {
    bytes = align_up(bytes, 1024);
    for (size_t i = 0; i < free_.size(); ++i) {
      if (free_[i].second >= bytes) {
        const size_t off = free_[i].first;
        if (free_[i].second == bytes) free_.erase(free_.begin() + i);
        else { free_[i].first += bytes; free_[i].second -= bytes; }
        return off;
      }
    }
    const size_t off = top_;
    top_ += bytes;
    high_ = std::max(high_, top_);
    return off;
  }
  void release(size_t off, size_t bytes) {
    bytes = align_up(bytes, 1024);
    free_.emplace_back(off, bytes);
    std::sort(free_.begin(), free_.end());
    for (size_t i = 0; i + 1 < free_.size();) {
      if (free_[i].first + free_[i].second == free_[i + 1].first) {
        free_[i].second += free_[i + 1].second;
        free_.erase(free_.begin() + i + 1);
      } else ++i;
    }
    if (!free_.empty() && free_.back().first + free_.back().second == top_) {
      top_ = free_.back().first;
      free_.pop_back();
    }
  }
  size_t high() const { return high_; }

 private:
  std::vector<std::pair<size_t, size_t>> free_;
  size_t top_ = 0, high_ = 0;
};

struct Buf { size_t off = 0, bytes = 0; };

struct DevTensor {
  float* p = nullptr;
  std::vector<int64_t> shape;
  int64_t numel = 0;
};

struct ConvW {
  bf16* w = nullptr;
  float* bias = nullptr;
  int cout = 0, cin = 0, kh = 1, kw = 1;
  int n_alloc = 0, k_alloc = 0;
  int nseg = 1;   // operand-split segments per tap (precise mode: 3 for hi/lo inputs, 2 for exact-bf16 inputs)
  bool f16 = false;   // packed as IEEE half instead of bf16
  // optional copy with an identity block appended along K ([W | I], row stride k_alloc_id): a 1x1 projection whose residual
  // is added by the GEMM itself, the residual tensor riding in as a second A operand (GemmEpilogue::a2 + a2_diag)
  bf16* w_id = nullptr;
  int k_alloc_id = 0;
};
struct NormW { float* gamma = nullptr; float* beta = nullptr; int C = 0; };
struct ResW {
  NormW n1, n2; ConvW c1, c2, skip; ConvW c2s; /* conv2 with the 1x1 skip appended along K */ bool has_skip = false; int emb_off = -1; int cin = 0, cout = 0;
  // layout U-Net (object_cross_unet.py:253-283): FiLM scale-shift conditioning (emb_layers gives 2*cout values) and
  // ResBlock up- (1) / down- (2) sampling of both h and x
  bool film = false; int updown = 0;
};
struct AttnW { NormW n; ConvW qkv, proj; int ch = 0, heads = 0; };
// BasicTransformerBlock / SpatialTransformer weights (reference lidm/modules/attention.py:196-261)
struct STBlockW {
  NormW n1, n2, n3;           // LayerNorms
  ConvW qkv1, out1;           // attn1: packed [q | k | v] rows (q, k pre-scaled by d^-1/4), to_out.0
  ConvW q2, out2;             // attn2: to_q (pre-scaled by d^-1/2), to_out.0; k / v come from the shared context GEMM
  ConvW ff0, ff2;             // GEGLU proj (C -> 8C), Linear (4C -> C)
  int kv_col = 0;             // column of this block's K rows inside the context K/V matrix (V at kv_col + C)
};
struct STW { NormW n; ConvW proj_in, proj_out; std::vector<STBlockW> blocks; int ch = 0, heads = 0; };
// ObjectAwareCrossAttention (object_cross_unet.py:380-565): qkv rows with 128^-1/4 folded into q and k; the positional /
// layout-content projectors stay fp32 (they run once per conditioning, lidm_layout_set_cond, into the buffers below)
struct OacaW {
  NormW n_qkv, n_img_pos, n_lay_pos, n_cls;
  ConvW qkv, proj;
  float *w_pos = nullptr, *b_pos = nullptr, *w_content = nullptr, *b_content = nullptr;
  int ch = 0, heads = 0, rows = 0;    // rows: feature-map rows of this block = key of image_patch_bbox_embedding_for_resolution{rows}
  bf16 *pos_img = nullptr, *klay = nullptr, *vlay = nullptr;    // conditioning buffers: (pos_batch,T,C), (B,16,2C), (B,16,C)
  int pos_batch = 0;
};
struct Layer {
  enum Kind { CONV, RES, ATTN, DOWN, UP, ST, OACA } kind;
  ResW r; AttnW a; ConvW c; STW st; OacaW oa;
  ConvW cpar[4];   // UP: the 3x3 conv after nearest x2, folded into one 2x2 conv per output parity (py * 2 + px)
  int cin = 0, cout = 0;
};
struct EncLevel { std::vector<ResW> blocks; bool has_down = false; int sh = 1, sw = 1, pl = 0, pt = 0; ConvW down; int ch = 0; };
struct DecLevel { std::vector<ResW> blocks; int kh = 1, kw = 4; bool has_up = false; int sh = 1, sw = 1; ConvW up; int ch = 0; };

ConvTaps taps_rect(int kh, int kw, int pl, int pt) {
  ConvTaps t;
  t.n = kh * kw;
  for (int ky = 0; ky < kh; ++ky)
    for (int kx = 0; kx < kw; ++kx) { t.dy[ky * kw + kx] = (int8_t)(ky - pt); t.dx[ky * kw + kx] = (int8_t)(kx - pl); }
  return t;
}
ConvTaps taps_1x1() { ConvTaps t; t.n = 1; return t; }

// log2 e rides on the softmax scale that is folded into the packed q / k weights (its square root on each of q and k, all of
// it on q where k comes from the shared context matrix): q . k is then the base-2 exponent the attention kernels feed to ex2
const float kLog2e = 1.4426950408889634f;
const float kSqrtLog2e = 1.2011224087864498f;

int round_n_alloc(int cout) {
  if (cout >= 128) return (cout + 127) / 128 * 128;
  if (cout > 16) return (cout + 63) / 64 * 64;
  return 16;
}

struct Plan {
  int B = 0;
  char* arena = nullptr;
  size_t arena_bytes = 0;
  std::vector<Op> ops;
  float* gn_partials = nullptr;
  // launch-time IO (closures read these when they run)
  const float* x = nullptr;         // U-Net latent x_t / decoder latent (fp32 NCHW)
  const float* xin = nullptr;       // U-Net conv-in source: x itself, or the assembled [x | c_concat] tensor
  const float* context = nullptr;   // cross-attention context (B, ctx_len, context_dim) fp32, or null
  int ctx_len = 0;
  float* out = nullptr;             // eps_out / img_out
  int32_t* idx_out = nullptr;
  int quantize = 1;
  int64_t cond_gen = 0;             // generation of the layout conditioning buffers the ops read (lidm_layout_set_cond)
  bool eff_input_ready = false;     // EfficientUNet: the constant coordinate-encoding channels of the input buffer are filled
  const float* rowadd_base = nullptr;
  int rowadd_ld = 0;
  const float* ddim_noise = nullptr;
  float* ddim_x_prev = nullptr;
  float* ddim_pred_x0 = nullptr;
  const float* ddim_coef = nullptr;
  // CUDA graphs of this plan (the DDIM loop replays one per step): valid while the launch-time IO pointers are unchanged
  struct GraphSlot {
    cudaGraphExec_t exec = nullptr;
    std::vector<const void*> key;
    int64_t kernels = 0;
  };
  GraphSlot gslot[3];
  bool warmed = false;              // ran eagerly at least once (lazy one-time kernel attribute setup happens there)
  std::vector<const void*> io_key() const {
    return {x, xin, context, out, idx_out, rowadd_base, reinterpret_cast<const void*>((intptr_t)rowadd_ld), ddim_noise,
            ddim_x_prev, ddim_pred_x0, ddim_coef, reinterpret_cast<const void*>((intptr_t)quantize),
            reinterpret_cast<const void*>((intptr_t)cond_gen)};
  }
  ~Plan() {
    for (auto& g : gslot) if (g.exec) cudaGraphExecDestroy(g.exec);
    if (arena) cudaFree(arena);
    if (gn_partials) cudaFree(gn_partials);
  }
};

}  // namespace
}  // namespace lidm

using namespace lidm;

struct lidm_handle {
  lidm_config cfg{};
  std::string error;
  bool finalized = false;
  int unet_prec = LIDM_PREC_BF16, ae_prec = LIDM_PREC_BF16;   // numeric mode of the U-Net / of the first stage
  // layout-conditioned U-Net: conditioning state written by lidm_layout_set_cond
  bool is_layout = false;
  int cond_B = 0, cond_n_layout = 0;
  int64_t cond_gen = 0;                                   // bumps whenever the conditioning buffers move (CUDA-graph key)
  float* cond_xf_proj = nullptr;                          // (cond_B, ted)
  std::vector<void*> cond_owned;
  // EfficientUNet (R2DM, unet_type 2)
  bool is_eff = false;
  struct EffBlock {
    bool has_down = false, has_up = false, has_attn = false;
    ConvW down, up;                                       // conv before the FIR down-sampler / after the FIR up-sampler
    std::vector<ResW> res;
    NormW attn_norm; ConvW attn_qkv, attn_proj; int attn_heads = 0, ch = 0, cin = 0;
  };
  EffBlock eff_d[4], eff_u[4];
  ConvW eff_in_conv;
  float* eff_cenc = nullptr; int eff_cenc_ch = 0;          // Fourier features of the polar coordinates (extra_ch, H, W) fp32
  float* eff_in_map = nullptr;                             // in_conv of the coordinate channels + bias: constant (H, W, C0) fp32 map
  float* eff_in_w = nullptr;                               // in_conv weights of the image channels [C0][Cx][9] fp32
  float *ones_c = nullptr, *zeros_c = nullptr;            // AdaGN = GroupNorm without affine parameters
  // LayoutTransformerEncoder (cond_stage_model.*), fp32; packed when the state-dict carries it
  bool has_layout_encoder = false;
  void* lenc_layers = nullptr;                            // device array of n_layers x 12 float pointers
  float *lenc_cls = nullptr, *lenc_be_w = nullptr, *lenc_be_b = nullptr, *lenc_bx_w = nullptr, *lenc_bx_b = nullptr,
        *lenc_fln_g = nullptr, *lenc_fln_b = nullptr, *lenc_tp_w = nullptr, *lenc_tp_b = nullptr;
  std::unordered_map<std::string, DevTensor> raw;     // fp32 state-dict tensors on the device
  std::vector<void*> owned;                           // packed weight allocations
  // U-Net
  std::vector<std::vector<Layer>> in_blocks, out_blocks;
  std::vector<Layer> mid_block;
  std::vector<int> in_chans;                          // channels of hs[k]
  NormW out_norm; ConvW out_conv;
  bool has_st = false;                                // SpatialTransformer U-Net (cross-attention conditioning)
  bf16* ctx_w = nullptr; int ctx_n = 0;               // every block's [to_k ; to_v] rows stacked: [ctx_n][context_dim]
  int latent_channels = 0;                            // channels of x / eps (in_channels - latent = concat channels)
  float* xcat = nullptr; size_t xcat_elems = 0;       // assembled [x | c_concat] (and the 2B classifier-free batch)
  float* ctx2 = nullptr; size_t ctx2_elems = 0;       // [uncond ; cond] context for classifier-free guidance
  float* eps2 = nullptr; size_t eps2_elems = 0;       // (2B) eps of the guidance batch
  float *te_w0 = nullptr, *te_b0 = nullptr, *te_w2 = nullptr, *te_b2 = nullptr;
  float *emb_w = nullptr, *emb_b = nullptr;           // concatenated emb_layers Linear weights [emb_total][ted]
  int emb_total = 0, ted = 0;
  // decoder
  float *codebook = nullptr, *cb_norm = nullptr, *pq_w = nullptr, *pq_b = nullptr;
  ConvW dec_conv_in, dec_conv_out;
  ResW dec_mid1, dec_mid2;
  AttnW dec_attn;
  std::vector<DecLevel> dec_levels;                   // indexed by i_level
  NormW dec_norm_out;
  int dec_top = 0, dec_last = 0, img_h = 0, img_w = 0;
  // encoder (optional: packed when the state-dict carries first_stage_model.encoder.*)
  bool has_encoder = false;
  ConvW enc_conv_in, enc_conv_out;                    // conv_out has quant_conv folded in
  std::vector<EncLevel> enc_levels;
  ResW enc_mid1, enc_mid2;
  AttnW enc_attn;
  NormW enc_norm_out;
  int enc_top = 0;
  // plans
  std::map<int64_t, std::unique_ptr<Plan>> unet_plans, dec_plans, enc_plans;   // key = B | ctx_len << 24
  // time-embedding scratch
  float *te_tmp = nullptr, *te_emb = nullptr, *emb_out = nullptr;
  int64_t* t_dev = nullptr;
  int te_rows = 0;
  float* coef_dev = nullptr; int coef_rows = 0;
  float *xa = nullptr, *xb = nullptr; size_t xbuf_elems = 0;
  // per-step staging the graphed DDIM loop reads through fixed addresses: this step's emb_layers row, coefficients,
  // noise slice; pred_x0 scratch
  float *emb_cur = nullptr, *coef_cur = nullptr, *noise_cur = nullptr, *pred_scratch = nullptr;
  size_t stage_elems = 0;
  cudaStream_t cap_stream = nullptr;

  ~lidm_handle() {
    for (auto& kv : raw) cudaFree(kv.second.p);
    for (void* p : owned) cudaFree(p);
    for (void* p : cond_owned) cudaFree(p);
    cudaFree(cond_xf_proj);
    cudaFree(te_tmp); cudaFree(te_emb); cudaFree(emb_out); cudaFree(t_dev); cudaFree(coef_dev);
    cudaFree(xa); cudaFree(xb); cudaFree(xcat); cudaFree(ctx2); cudaFree(eps2);
    cudaFree(emb_cur); cudaFree(coef_cur); cudaFree(noise_cur); cudaFree(pred_scratch);
    if (cap_stream) cudaStreamDestroy(cap_stream);
  }
};

namespace lidm {
namespace {

template <class T>
T* dev_alloc(lidm_handle* h, size_t n) {
  void* p = nullptr;
  LIDM_CUDA_CHECK(cudaMalloc(&p, std::max<size_t>(n, 1) * sizeof(T)));
  if (h) h->owned.push_back(p);
  return reinterpret_cast<T*>(p);
}

// ------------------------------------------------------------------------------------------- weight lookup
const DevTensor& find_raw(lidm_handle* h, const std::string& name, bool use_ema) {
  if (use_ema) {
    // LitEma buffer name: parameter name relative to DDPM.model with the dots removed (lidm/modules/ema.py:16-21)
    const std::string pre = "model.";
    if (name.compare(0, pre.size(), pre) == 0) {
      std::string s = name.substr(pre.size());
      s.erase(std::remove(s.begin(), s.end(), '.'), s.end());
      auto it = h->raw.find("model_ema." + s);
      if (it != h->raw.end()) return it->second;
    }
  }
  auto it = h->raw.find(name);
  if (it == h->raw.end()) throw Error(LIDM_ERR_STATE, "missing weight tensor '" + name + "'");
  return it->second;
}

struct Packer {
  lidm_handle* h;
  bool ema;
  cudaStream_t s = 0;
  bool precise = false;
  bool f16 = false;    // pack GEMM operands as IEEE half (never together with `precise`)

  float* f32(const std::string& name, int64_t expect_numel) {
    const DevTensor& t = find_raw(h, name, ema);
    if (t.numel != expect_numel)
      throw Error(LIDM_ERR_STATE, "weight '" + name + "' has " + std::to_string(t.numel) + " elements, expected " +
                                      std::to_string(expect_numel));
    float* p = dev_alloc<float>(h, (size_t)t.numel);
    LIDM_CUDA_CHECK(cudaMemcpyAsync(p, t.p, t.numel * sizeof(float), cudaMemcpyDeviceToDevice, s));
    return p;
  }
  NormW norm(const std::string& prefix, int C) {
    NormW n;
    n.C = C;
    n.gamma = f32(prefix + ".weight", C);
    n.beta = f32(prefix + ".bias", C);
    return n;
  }
  // k_alloc_override: K of the packed matrix (for im2col'd operands padded to a multiple of 64)
  // k_alloc_override: per-plane K of an im2col'd small-channel operand (padded to a multiple of 64)
  // nseg_p: operand-split segments used in precise mode (3: hi/lo activations, 2: exact bf16 activations)
  ConvW conv(const std::string& prefix, int cout, int cin, int kh, int kw, int k_alloc_override = 0, int nseg_p = 3,
             float scale = 1.f /* multiplies weights and bias (bf16 / fp16 modes only) */) {
    const DevTensor& t = find_raw(h, prefix + ".weight", ema);
    if (t.numel != (int64_t)cout * cin * kh * kw)
      throw Error(LIDM_ERR_STATE, "weight '" + prefix + ".weight' has unexpected size");
    ConvW c;
    c.cout = cout; c.cin = cin; c.kh = kh; c.kw = kw;
    c.n_alloc = round_n_alloc(cout);
    c.nseg = precise ? nseg_p : 1;
    c.f16 = f16;
    if (!precise) {
      c.k_alloc = k_alloc_override ? k_alloc_override : kh * kw * cin;
      c.w = dev_alloc<bf16>(h, (size_t)c.n_alloc * c.k_alloc);
      launch_pack_conv_weight(t.p, cout, cin, kh, kw, c.n_alloc, c.k_alloc, nullptr, nullptr, scale, scale != 1.f ? cout : 0,
                              c.w, s, f16);
    } else if (k_alloc_override) {
      // K layout [seg][kpad]: reorder to fp32 [cout][kpad] (k = tap*cin + c) and split-pack it as a 1x1 conv
      float* tmp = nullptr;
      LIDM_CUDA_CHECK(cudaMalloc(reinterpret_cast<void**>(&tmp), (size_t)cout * k_alloc_override * sizeof(float)));
      launch_reorder_weight_f32(t.p, cout, cin, kh * kw, k_alloc_override, tmp, s);
      c.k_alloc = c.nseg * k_alloc_override;
      c.w = dev_alloc<bf16>(h, (size_t)c.n_alloc * c.k_alloc);
      launch_pack_conv_weight_split(tmp, cout, k_alloc_override, 1, 1, c.n_alloc, c.nseg, nullptr, 1.f, 0, c.w, s);
      LIDM_CUDA_CHECK(cudaStreamSynchronize(s));
      cudaFree(tmp);
    } else {
      c.k_alloc = kh * kw * c.nseg * cin;
      c.w = dev_alloc<bf16>(h, (size_t)c.n_alloc * c.k_alloc);
      launch_pack_conv_weight_split(t.p, cout, cin, kh, kw, c.n_alloc, c.nseg, nullptr, 1.f, 0, c.w, s);
    }
    c.bias = f32(prefix + ".bias", cout);
    if (scale != 1.f) {
      if (precise) throw Error(LIDM_ERR_INVALID, "internal: scaled convolutions are not packed in the operand-split mode");
      std::vector<float> bh(cout);
      LIDM_CUDA_CHECK(cudaMemcpyAsync(bh.data(), c.bias, cout * sizeof(float), cudaMemcpyDeviceToHost, s));
      LIDM_CUDA_CHECK(cudaStreamSynchronize(s));
      for (float& v : bh) v *= scale;
      LIDM_CUDA_CHECK(cudaMemcpy(c.bias, bh.data(), cout * sizeof(float), cudaMemcpyHostToDevice));
    }
    return c;
  }
};

// nn.Linear / 1x1 weights stacked row-wise into one K-major bf16 matrix [n_alloc][cin]; part i may be scaled.
struct LinPart { std::string name; float scale; };
ConvW pack_stacked_linear(Packer& pk, const std::vector<LinPart>& parts, int cout_each, int cin, const std::string& bias_name) {
  lidm_handle* h = pk.h;
  if (pk.precise || pk.f16) throw Error(LIDM_ERR_INVALID, "SpatialTransformer U-Nets run in the bf16 mode only");
  ConvW c;
  c.cout = cout_each * (int)parts.size(); c.cin = cin; c.kh = c.kw = 1;
  c.n_alloc = round_n_alloc(c.cout);
  c.k_alloc = cin;
  c.w = dev_alloc<bf16>(h, (size_t)c.n_alloc * cin);
  LIDM_CUDA_CHECK(cudaMemsetAsync(c.w, 0, (size_t)c.n_alloc * cin * sizeof(bf16), pk.s));
  for (size_t i = 0; i < parts.size(); ++i) {
    const DevTensor& w = find_raw(h, parts[i].name, pk.ema);
    if (w.numel != (int64_t)cout_each * cin) throw Error(LIDM_ERR_STATE, "weight '" + parts[i].name + "' has unexpected size");
    launch_pack_conv_weight(w.p, cout_each, cin, 1, 1, cout_each, cin, nullptr, nullptr, parts[i].scale,
                            parts[i].scale != 1.f ? cout_each : 0, c.w + i * (size_t)cout_each * cin, pk.s);
  }
  c.bias = bias_name.empty() ? nullptr : pk.f32(bias_name, c.cout);
  return c;
}

// GEGLU projection (Linear dim -> 2 * inner, output = x * gelu(gate), lidm/modules/attention.py:36-44) packed for the fused
// epilogue: output rows interleaved in blocks of [16 value rows | 16 gate rows], bias permuted alike.
ConvW pack_geglu_linear(Packer& pk, const std::string& prefix, int cout, int cin) {
  lidm_handle* h = pk.h;
  const DevTensor& w = find_raw(h, prefix + ".weight", pk.ema);
  const DevTensor& bsrc = find_raw(h, prefix + ".bias", pk.ema);
  if (w.numel != (int64_t)cout * cin || bsrc.numel != cout || cout % 32 != 0)
    throw Error(LIDM_ERR_STATE, "GEGLU projection weight size: " + prefix);
  const int inner = cout / 2;
  std::vector<int> perm(cout);
  for (int r = 0; r < cout; ++r) {
    const int blk = r / 32, i = r % 32;
    perm[r] = i < 16 ? blk * 16 + i : inner + blk * 16 + (i - 16);
  }
  int* perm_dev = dev_alloc<int>(h, perm.size());
  LIDM_CUDA_CHECK(cudaMemcpyAsync(perm_dev, perm.data(), perm.size() * sizeof(int), cudaMemcpyHostToDevice, pk.s));
  ConvW c;
  c.cout = cout; c.cin = cin; c.kh = c.kw = 1;
  c.n_alloc = round_n_alloc(cout);
  c.k_alloc = cin;
  c.w = dev_alloc<bf16>(h, (size_t)c.n_alloc * cin);
  launch_pack_conv_weight(w.p, cout, cin, 1, 1, c.n_alloc, cin, perm_dev, nullptr, 1.f, 0, c.w, pk.s);
  std::vector<float> bh(cout), bp(cout);
  LIDM_CUDA_CHECK(cudaMemcpyAsync(bh.data(), bsrc.p, cout * sizeof(float), cudaMemcpyDeviceToHost, pk.s));
  LIDM_CUDA_CHECK(cudaStreamSynchronize(pk.s));
  for (int r = 0; r < cout; ++r) bp[r] = bh[perm[r]];
  c.bias = dev_alloc<float>(h, cout);
  LIDM_CUDA_CHECK(cudaMemcpy(c.bias, bp.data(), cout * sizeof(float), cudaMemcpyHostToDevice));
  return c;
}

// ------------------------------------------------------------------------------------------- plan builder
struct Builder {
  lidm_handle* h;
  Plan* P;
  ArenaPlanner ap;
  bool dry;
  bool precise = false;   // this plan runs the operand-split (fp32-class) kernels
  bool f16 = false;       // 2-byte activations of this plan are IEEE half instead of bf16
  int gn_groups = 32;     // GroupNorm group count (32 everywhere but the R2DM U-Net: 8)
  // GroupNorm statistics written by GEMM epilogues (View::gst): which 8-channel granules of each statistics buffer have
  // been produced so far in plan order; a GroupNorm whose whole input is covered skips its statistics pass
  std::map<const float*, std::vector<char>> gst_cover;
  bool gst_enabled = getenv("LIDM_NO_GN_FUSE") == nullptr;

  View act(int B, int H, int W, int C, int hl, int hr, Buf* buf) {
    View v;
    v.B = B; v.H = H; v.W = W; v.C = C; v.hl = hl; v.hr = hr; v.ld = C; v.f16 = f16;
    const size_t act_bytes = align_up((size_t)B * H * (W + hl + hr) * C * sizeof(bf16), 256);
    // halo-free tensors made of whole 128-pixel tiles get a statistics side buffer [B][H*W/128][C/8][2] floats
    const bool want_gst = gst_enabled && !precise && hl == 0 && hr == 0 && C % 8 == 0 && (H * W) % 128 == 0;
    const size_t gst_bytes = want_gst ? (size_t)B * (H * W / 128) * (C / 8) * 2 * sizeof(float) : 0;
    buf->bytes = act_bytes + gst_bytes;
    buf->off = ap.alloc(buf->bytes);
    v.p = reinterpret_cast<bf16*>(P->arena + buf->off);
    if (want_gst) {
      v.gst = reinterpret_cast<float*>(P->arena + buf->off + act_bytes);
      v.gst_ld = (C / 8) * 2;
      v.gst_slots = H * W / 128;
      gst_cover[v.gst] = std::vector<char>(C / 8, 0);
      v.gst_base = v.gst;
    }
    return v;
  }
  bool gst_covered(const View& x) const {
    if (x.gst == nullptr || x.gst_base == nullptr || x.C % 8 != 0) return false;
    auto it = gst_cover.find(x.gst_base);
    if (it == gst_cover.end()) return false;
    const int g0 = (int)((x.gst - x.gst_base) / 2);
    for (int g = g0; g < g0 + x.C / 8; ++g)
      if (g >= (int)it->second.size() || !it->second[g]) return false;
    return true;
  }
  void gst_mark(const View& o, int N, int slot0, int nslots) {
    // a buffer is covered granule by granule; parity-strided outputs (folded upsample) cover a granule only once all four
    // parities have been written, which up_folded does back to back, so marking on the last parity is enough
    if (o.gst == nullptr || o.gst_base == nullptr) return;
    if (slot0 + nslots < o.gst_slots) return;
    auto it = gst_cover.find(o.gst_base);
    if (it == gst_cover.end()) return;
    const int g0 = (int)((o.gst - o.gst_base) / 2);
    for (int g = g0; g < g0 + N / 8 && g < (int)it->second.size(); ++g) it->second[g] = 1;
  }
  template <class T>
  T* raw(size_t n, Buf* buf) {
    buf->bytes = n * sizeof(T);
    buf->off = ap.alloc(buf->bytes);
    return reinterpret_cast<T*>(P->arena + buf->off);
  }
  void release(const Buf& b) { ap.release(b.off, b.bytes); }
  void op(std::function<void(cudaStream_t)> f, int cat = PROF_OTHER, double flops = 0, double bytes = 0,
          std::string label = std::string()) {
    if (!dry) P->ops.push_back(Op{std::move(f), cat, flops, bytes, std::move(label)});
  }
  static std::string gemm_label(const View& a, int ntaps, int N) {
    return "gemm taps" + std::to_string(ntaps) + " " + std::to_string(a.C) + "->" + std::to_string(N) + " @" +
           std::to_string(a.H) + "x" + std::to_string(a.W);
  }

  static GemmB gb(const ConvW& w) {
    GemmB b; b.p = w.w; b.n_alloc = w.n_alloc; b.ld = w.k_alloc; b.nseg = w.nseg; b.f16 = w.f16;
    return b;
  }
  // a plain channels-last matrix view over a raw buffer (im2col output, attention probabilities ...)
  View mat(bf16* p, int B, int H, int W, int C, int ld) const {
    View v; v.p = p; v.B = B; v.H = H; v.W = W; v.C = C; v.ld = ld; v.f16 = f16;
    return v;
  }

  static View chan_slice(const View& v, int c0, int C) {
    View s = v;
    s.p = v.p + c0;
    s.C = C;
    if (v.gst != nullptr) s.gst = (c0 % 8 == 0) ? v.gst + (c0 / 8) * 2 : nullptr;
    return s;
  }

  // decide at plan time whether this GEMM's epilogue also writes the GroupNorm statistics of its output
  void prep_gst(GemmEpilogue& ep, int N, int n_alloc) {
    if (ep.out.gst == nullptr) return;
    if (N % 8 == 0 && conv_gemm_emits_gstats(ep, n_alloc)) gst_mark(ep.out, N, ep.out.gst_slot0, ep.out.H * ep.out.W / 128);
    else ep.out.gst = nullptr;
  }

  void gemm(const View& a, const ConvTaps& taps, const ConvW& w, const GemmEpilogue& ep_in) {
    const GemmB b = gb(w);
    const int N = w.cout;
    GemmEpilogue ep = ep_in;
    prep_gst(ep, N, w.n_alloc);
    op([=](cudaStream_t s) { launch_conv_gemm(a, taps, b, N, ep, s); }, PROF_GEMM,
       gemm_flops(a, taps.n, N) * w.nseg / (taps.sx * taps.sy), 0,
       gemm_label(a, taps.n * w.nseg, N) + (taps.sx * taps.sy > 1 ? " /" + std::to_string(taps.sy) + "x" + std::to_string(taps.sx) : ""));
  }

  // 1x1 GEMM + residual: through the identity-extended weights when they exist (the residual is a second A operand,
  // full 128-byte rows through the TMA ring), else as a per-thread residual read in the epilogue
  void gemm_res(const View& a, const ConvW& w, const GemmEpilogue& ep_in, const View& residual, float res_scale = 1.f) {
    static const bool no_idres = getenv("LIDM_NO_IDRES") != nullptr;   // A/B switch
    GemmEpilogue ep = ep_in;
    if (w.w_id != nullptr && !no_idres && res_scale == 1.f && residual.wpitch == 0 && residual.C == w.cout && residual.f16 == a.f16 &&
        residual.B == a.B && residual.H == a.H && residual.W == a.W && !ep.geglu) {
      ConvW wi = w;
      wi.w = w.w_id; wi.k_alloc = w.k_alloc_id;
      ep.a2 = residual;
      ep.a2_diag = true;
      gemm(a, taps_1x1(), wi, ep);
    } else {
      ep.residual = residual;
      ep.res_scale = res_scale;
      gemm(a, taps_1x1(), w, ep);
    }
  }

  // =============================================================================================== precise mode
  // fp32 residual stream (ViewF) between GEMMs; GEMM inputs are bf16 hi/lo planes (View with lo_off = C).
  ViewF actf(int B, int H, int W, int C, Buf* buf) {
    ViewF v;
    v.B = B; v.H = H; v.W = W; v.C = C; v.ld = C;
    buf->bytes = (size_t)B * H * W * C * sizeof(float);
    buf->off = ap.alloc(buf->bytes);
    v.p = reinterpret_cast<float*>(P->arena + buf->off);
    return v;
  }
  View act_hl(int B, int H, int W, int C, int hl, int hr, Buf* buf) {
    View v;
    v.B = B; v.H = H; v.W = W; v.C = C; v.hl = hl; v.hr = hr; v.ld = 2 * C; v.lo_off = C;
    buf->bytes = (size_t)B * H * (W + hl + hr) * 2 * C * sizeof(bf16);
    buf->off = ap.alloc(buf->bytes);
    v.p = reinterpret_cast<bf16*>(P->arena + buf->off);
    return v;
  }
  static ViewF chan_slice_f(const ViewF& v, int c0, int C) {
    ViewF s = v;
    s.p = v.p + c0;
    s.C = C;
    return s;
  }
  void groupnorm_f(const ViewF& x, const View& y, const NormW& n, float eps, bool silu) {
    Plan* P_ = P;
    op([=](cudaStream_t s) { launch_groupnorm_f32(x, y, n.gamma, n.beta, eps, 32, silu, P_->gn_partials, s); }, PROF_NORM,
       0, 8.0 * x.B * x.H * x.W * x.C, "gn(f32) C" + std::to_string(x.C) + " @" + std::to_string(x.H) + "x" + std::to_string(x.W));
  }
  static void set_out_f(GemmEpilogue& ep, const ViewF& dst) { ep.out_f32_nhwc = dst.p; ep.out_f32_ld = dst.ld; }
  static void set_res_f(GemmEpilogue& ep, const ViewF& r) { ep.res_f32 = r.p; ep.res_f32_ld = r.ld; }

  void res_block_p(const ResW& r, const ViewF& x, const ViewF& dst, int kh, int kw, int pl, int pr, int pt, float eps) {
    const int B = x.B, H = x.H, W = x.W;
    Buf bg1, bh, bg2, bsk, bxs;
    View g1 = act_hl(B, H, W, r.cin, pl, pr, &bg1);
    groupnorm_f(x, g1, r.n1, eps, true);
    ViewF hmid = actf(B, H, W, r.cout, &bh);
    {
      GemmEpilogue ep;
      ep.bias = r.c1.bias;
      set_out_f(ep, hmid);
      GemmB b; b.p = r.c1.w; b.n_alloc = r.c1.n_alloc; b.ld = r.c1.k_alloc; b.nseg = r.c1.nseg;
      const ConvTaps taps = taps_rect(kh, kw, pl, pt);
      const int N = r.cout, emb_off = r.emb_off;
      Plan* P_ = P;
      op([=](cudaStream_t s) {
        GemmEpilogue e = ep;
        if (emb_off >= 0) { e.rowadd = P_->rowadd_base + emb_off; e.rowadd_ld = P_->rowadd_ld; }
        launch_conv_gemm(g1, taps, b, N, e, s);
      }, PROF_GEMM, gemm_flops(g1, taps.n, N) * b.nseg, 0, gemm_label(g1, taps.n * b.nseg, N));
    }
    release(bg1);
    View g2 = act_hl(B, H, W, r.cout, pl, pr, &bg2);
    groupnorm_f(hmid, g2, r.n2, eps, true);
    release(bh);
    ViewF resid = x;
    if (r.has_skip) {
      View xs = act_hl(B, H, W, r.cin, 0, 0, &bxs);
      op([=](cudaStream_t s) { launch_split_f32(x, xs, s); });
      ViewF sk = actf(B, H, W, r.cout, &bsk);
      GemmEpilogue ep;
      ep.bias = r.skip.bias;
      set_out_f(ep, sk);
      gemm(xs, taps_1x1(), r.skip, ep);
      release(bxs);
      resid = sk;
    }
    {
      GemmEpilogue ep;
      ep.bias = r.c2.bias;
      set_res_f(ep, resid);
      set_out_f(ep, dst);
      gemm(g2, taps_rect(kh, kw, pl, pt), r.c2, ep);
    }
    release(bg2);
    if (r.has_skip) release(bsk);
  }

  void attn_block_p(const AttnW& a, const ViewF& x, const ViewF& dst) {
    const int B = x.B, H = x.H, W = x.W, C = a.ch, T = H * W;
    Buf bg, bqk, ba;
    View g = act_hl(B, H, W, C, 0, 0, &bg);
    groupnorm_f(x, g, a.n, 1e-5f, false);
    View qkv = act(B, H, W, 3 * C, 0, 0, &bqk);
    {
      GemmEpilogue ep;
      ep.bias = a.qkv.bias;
      ep.out = qkv;
      gemm(g, taps_1x1(), a.qkv, ep);
    }
    release(bg);
    View ao = act(B, H, W, C, 0, 0, &ba);
    const int heads = a.heads;
    op([=](cudaStream_t s) { launch_attention_d32_packed(qkv.p, ao, B, T, heads, s, true); }, PROF_ATTN,
       4.0 * B * heads * (double)T * T * 32, 0, "attn T" + std::to_string(T) + " heads" + std::to_string(heads));
    release(bqk);
    {
      GemmEpilogue ep;
      ep.bias = a.proj.bias;
      set_res_f(ep, x);
      set_out_f(ep, dst);
      gemm(ao, taps_1x1(), a.proj, ep);
    }
    release(ba);
  }

  void dec_attn_block_p(const AttnW& a, const ViewF& x, const ViewF& dst) {
    const int B = x.B, H = x.H, W = x.W, C = a.ch, T = H * W;
    Buf bg, bqk, bvt, bs, bp, ba;
    View g = act_hl(B, H, W, C, 0, 0, &bg);
    groupnorm_f(x, g, a.n, 1e-6f, false);
    View qk = act(B, H, W, 2 * C, 0, 0, &bqk);
    bf16* vt = raw<bf16>((size_t)B * C * T, &bvt);
    {
      GemmEpilogue ep;
      ep.bias = a.qkv.bias;
      ep.out = qk;
      ep.split_n = 2 * C;
      ep.out_t = vt;
      gemm(g, taps_1x1(), a.qkv, ep);
    }
    release(bg);
    float* S = raw<float>((size_t)B * T * T, &bs);
    {
      View q = chan_slice(qk, 0, C);
      GemmB kb; kb.p = qk.p + C; kb.n_alloc = T; kb.ld = 2 * C; kb.batch_stride = (int64_t)T * 2 * C;
      GemmEpilogue ep;
      ep.out_f32_nhwc = S;
      op([=](cudaStream_t s) { launch_conv_gemm(q, taps_1x1(), kb, T, ep, s); }, PROF_GEMM, gemm_flops(q, 1, T));
    }
    bf16* Pm = raw<bf16>((size_t)B * T * T, &bp);
    op([=](cudaStream_t s) { launch_softmax_rows(S, Pm, (int64_t)B * T, T, s); });
    release(bs);
    View ao = act(B, H, W, C, 0, 0, &ba);
    {
      View pv; pv.p = Pm; pv.B = B; pv.H = T / 128; pv.W = 128; pv.C = T; pv.ld = T;
      GemmB vb; vb.p = vt; vb.n_alloc = C; vb.ld = T; vb.batch_stride = (int64_t)C * T;
      GemmEpilogue ep;
      View o2 = ao; o2.H = T / 128; o2.W = 128;
      ep.out = o2;
      op([=](cudaStream_t s) { launch_conv_gemm(pv, taps_1x1(), vb, C, ep, s); }, PROF_GEMM, gemm_flops(pv, 1, C));
    }
    release(bp); release(bqk); release(bvt);
    {
      GemmEpilogue ep;
      ep.bias = a.proj.bias;
      set_res_f(ep, x);
      set_out_f(ep, dst);
      gemm(ao, taps_1x1(), a.proj, ep);
    }
    release(ba);
  }

  // strided conv on the fp32 stream: U-Net Downsample (3x3 stride 2, pad 1/1) and the encoder's Downsample
  // (model_lidm.py:68-83: 3x3, stride (sh, sw), pad left pl / top pt only)
  void down_p(const ConvW& c, const ViewF& x, const ViewF& dst, int sh = 2, int sw = 2, int pl = 1, int pt = 1) {
    const int B = x.B, Ho = dst.H, Wo = dst.W, C = x.C, nt = c.kh * c.kw;
    Buf bxs, bc;
    View xs = act_hl(B, x.H, x.W, C, 0, 0, &bxs);
    op([=](cudaStream_t s) { launch_split_f32(x, xs, s); });
    bf16* col = raw<bf16>((size_t)B * Ho * Wo * nt * 2 * C, &bc);
    View xs2 = xs; xs2.C = 2 * C; xs2.lo_off = 0;      // im2col moves both planes as one 2C-channel tensor
    const int kh = c.kh, kw = c.kw;
    op([=](cudaStream_t s) { launch_im2col_nhwc(xs2, kh, kw, sh, pl, pt, Ho, Wo, col, s, sw); });
    release(bxs);
    View a; a.p = col; a.B = B; a.H = Ho; a.W = Wo; a.C = C; a.ld = 2 * nt * C; a.lo_off = C; a.cphys = 2 * nt * C;
    ConvTaps taps; taps.n = nt; taps.cstep = 2 * C;    // row = [tap][hi C | lo C]
    GemmEpilogue ep;
    ep.bias = c.bias;
    set_out_f(ep, dst);
    gemm(a, taps, c, ep);
    release(bc);
  }

  void up_p(const ConvW& c, const ViewF& x, const ViewF& dst) {
    const int C = x.C;
    Buf bxs, bu;
    View xs = act_hl(x.B, x.H, x.W, C, 0, 0, &bxs);
    op([=](cudaStream_t s) { launch_split_f32(x, xs, s); });
    View u = act_hl(x.B, x.H * 2, x.W * 2, C, 1, 1, &bu);
    View xs2 = xs; xs2.C = 2 * C; xs2.lo_off = 0;
    View u2 = u; u2.C = 2 * C; u2.lo_off = 0;
    op([=](cudaStream_t s) { launch_upsample_nearest2x(xs2, u2, s); });
    release(bxs);
    GemmEpilogue ep;
    ep.bias = c.bias;
    set_out_f(ep, dst);
    gemm(u, taps_rect(3, 3, 1, 1), c, ep);
    release(bu);
  }
  // film_off >= 0: FiLM rows (scale | shift) of this block inside the per-sample emb_layers output (Plan::rowadd_base)
  void groupnorm(const View& x, const View& y, const NormW& n, float eps, bool silu, int film_off = -1) {
    Plan* P_ = P;
    const std::string label = "gn C" + std::to_string(x.C) + " @" + std::to_string(x.H) + "x" + std::to_string(x.W);
    const int G = gn_groups;
    if (x.C % G == 0 && (x.C / G) % 8 == 0 && x.hl == 0 && x.hr == 0 && gst_covered(x)) {
      // every producer of x left its granule statistics behind: one pass (read x, write y)
      op([=](cudaStream_t s) {
        launch_groupnorm_from_gstats(x, y, n.gamma, n.beta, eps, G, silu, s,
                                     film_off >= 0 ? P_->rowadd_base + film_off : nullptr, P_->rowadd_ld);
      }, PROF_NORM, 0, 4.0 * x.B * x.H * x.W * x.C, label);
      return;
    }
    op([=](cudaStream_t s) {
      launch_groupnorm(x, y, n.gamma, n.beta, eps, G, silu, P_->gn_partials, s,
                       film_off >= 0 ? P_->rowadd_base + film_off : nullptr, P_->rowadd_ld);
    }, PROF_NORM, 0, 4.0 * x.B * x.H * x.W * x.C, label);
  }

  static ConvTaps taps_zero3x3() { ConvTaps t = taps_rect(3, 3, 1, 1); t.zero_w = true; return t; }

  // ResBlock.forward of the layout U-Net (object_cross_unet.py:253-283): GN32 -> SiLU -> [nearest x2 | 2x2 average of h
  // AND x] -> zero-padded conv3x3 -> GN32 * (1 + scale) + shift -> SiLU -> zero-padded conv3x3 -> + skip(x)
  void res_block_film(const ResW& r, const View& x, const View& dst) {
    const int B = x.B, H = x.H, W = x.W;
    const int Ho = r.updown == 1 ? 2 * H : (r.updown == 2 ? H / 2 : H), Wo = r.updown == 1 ? 2 * W : (r.updown == 2 ? W / 2 : W);
    Buf bg1, bg1r, bxr, bh, bg2;
    View g1 = act(B, H, W, r.cin, 0, 0, &bg1);
    groupnorm(x, g1, r.n1, 1e-5f, true);
    View a1 = g1, xr = x;
    if (r.updown) {
      a1 = act(B, Ho, Wo, r.cin, 0, 0, &bg1r);
      xr = act(B, Ho, Wo, r.cin, 0, 0, &bxr);
      if (r.updown == 1) {
        op([=](cudaStream_t s) { launch_upsample_nearest2x(g1, a1, s); });
        op([=](cudaStream_t s) { launch_upsample_nearest2x(x, xr, s); });
      } else {
        op([=](cudaStream_t s) { launch_avgpool2(g1, a1, s); });
        op([=](cudaStream_t s) { launch_avgpool2(x, xr, s); });
      }
      release(bg1);
    }
    View hmid = act(B, Ho, Wo, r.cout, 0, 0, &bh);
    {
      GemmEpilogue ep;
      ep.bias = r.c1.bias;
      ep.out = hmid;
      gemm(a1, taps_zero3x3(), r.c1, ep);
    }
    release(r.updown ? bg1r : bg1);
    View g2 = act(B, Ho, Wo, r.cout, 0, 0, &bg2);
    groupnorm(hmid, g2, r.n2, 1e-5f, true, r.film ? r.emb_off : -1);
    release(bh);
    if (r.has_skip && r.c2s.w != nullptr && xr.wpitch == 0) {
      GemmEpilogue ep;                   // conv2(h) + skip(x) as one GEMM (x rides along as one more K segment)
      ep.bias = r.c2s.bias;
      ep.a2 = xr;
      ep.out = dst;
      prep_gst(ep, r.cout, r.c2s.n_alloc);
      const GemmB b = gb(r.c2s);
      const ConvTaps taps = taps_zero3x3();
      const int N = r.cout;
      op([=](cudaStream_t s) { launch_conv_gemm(g2, taps, b, N, ep, s); }, PROF_GEMM,
         gemm_flops(g2, taps.n, N) + gemm_flops(xr, 1, N), 0, gemm_label(g2, taps.n, N) + " +skip" + std::to_string(r.cin));
    } else {
      View resid = xr;
      Buf bsk;
      if (r.has_skip) {
        View sk = act(B, Ho, Wo, r.cout, 0, 0, &bsk);
        GemmEpilogue ep;
        ep.bias = r.skip.bias;
        ep.out = sk;
        gemm(xr, taps_1x1(), r.skip, ep);
        resid = sk;
      }
      GemmEpilogue ep;
      ep.bias = r.c2.bias;
      ep.residual = resid;
      ep.out = dst;
      gemm(g2, taps_zero3x3(), r.c2, ep);
      if (r.has_skip) release(bsk);
    }
    release(bg2);
    if (r.updown) release(bxr);
  }

  // ResidualBlock.forward of the R2DM U-Net (efficient_unet.py:55-110): GN -> SiLU -> ring conv -> AdaGN(temb) -> SiLU ->
  // ring conv; (skip(x) + h) / sqrt 2 with the scale folded into conv2 / skip (an identity skip enters as res_scale * x)
  void res_block_eff(const ResW& r, const View& x, const View& dst, float eps) {
    const int B = x.B, H = x.H, W = x.W;
    Buf bg1, bh, bg2;
    View g1 = act(B, H, W, r.cin, 1, 1, &bg1);
    groupnorm(x, g1, r.n1, eps, true);
    View hmid = act(B, H, W, r.cout, 0, 0, &bh);
    {
      GemmEpilogue ep;
      ep.bias = r.c1.bias;
      ep.out = hmid;
      gemm(g1, taps_rect(3, 3, 1, 1), r.c1, ep);
    }
    release(bg1);
    View g2 = act(B, H, W, r.cout, 1, 1, &bg2);
    groupnorm(hmid, g2, r.n2, eps, true, r.emb_off);
    release(bh);
    if (r.has_skip && r.c2s.w != nullptr && x.wpitch == 0) {
      GemmEpilogue ep;
      ep.bias = r.c2s.bias;
      ep.a2 = x;
      ep.out = dst;
      prep_gst(ep, r.cout, r.c2s.n_alloc);
      const GemmB b = gb(r.c2s);
      const ConvTaps taps = taps_rect(3, 3, 1, 1);
      const int N = r.cout;
      op([=](cudaStream_t s) { launch_conv_gemm(g2, taps, b, N, ep, s); }, PROF_GEMM,
         gemm_flops(g2, taps.n, N) + gemm_flops(x, 1, N), 0, gemm_label(g2, taps.n, N) + " +skip" + std::to_string(r.cin));
    } else {
      if (r.has_skip) throw Error(LIDM_ERR_INVALID, "internal: EfficientUNet skip convolutions are always folded");
      GemmEpilogue ep;
      ep.bias = r.c2.bias;
      ep.residual = x;
      ep.res_scale = 0.70710678118654752f;
      ep.out = dst;
      gemm(g2, taps_rect(3, 3, 1, 1), r.c2, ep);
    }
    release(bg2);
  }

  // SelfAttentionBlock.forward (efficient_unet.py:23-52): GN -> nn.MultiheadAttention -> (x + h) / sqrt 2.  Head width 32
  // runs the flash kernel; head width 64 runs per-head GEMMs (S = Q K^T, row softmax, P V) on the conv-GEMM kernel.
  void attn_block_eff(const lidm_handle::EffBlock& blk, const View& x, const View& dst, float eps) {
    const int B = x.B, H = x.H, W = x.W, C = blk.ch, T = H * W, heads = blk.attn_heads, d = C / heads;
    Buf bg, bqk, bvt, bs, bp, ba;
    View g = act(B, H, W, C, 0, 0, &bg);
    groupnorm(x, g, blk.attn_norm, eps, false);
    View ao;
    if (d == 32) {
      View qkv = act(B, H, W, 3 * C, 0, 0, &bqk);
      GemmEpilogue ep;
      ep.bias = blk.attn_qkv.bias;
      ep.out = qkv;
      ep.out.gst = nullptr;
      gemm(g, taps_1x1(), blk.attn_qkv, ep);
      release(bg);
      ao = act(B, H, W, C, 0, 0, &ba);
      op([=](cudaStream_t s) { launch_attention_d32_packed(qkv.p, ao, B, T, heads, s, true); }, PROF_ATTN,
         4.0 * B * heads * (double)T * T * 32, 0, "attn T" + std::to_string(T) + " heads" + std::to_string(heads));
      release(bqk);
    } else {
      if (T % 128 != 0 || T > 4096) throw Error(LIDM_ERR_INVALID, "EfficientUNet attention (head width 64): T must be a multiple of 128");
      View qk = act(B, H, W, 2 * C, 0, 0, &bqk);
      bf16* vt = raw<bf16>((size_t)B * C * T, &bvt);
      {
        GemmEpilogue ep;
        ep.bias = blk.attn_qkv.bias;
        ep.out = qk;
        ep.out.gst = nullptr;
        ep.split_n = 2 * C;
        ep.out_t = vt;
        gemm(g, taps_1x1(), blk.attn_qkv, ep);
      }
      release(bg);
      ao = act(B, H, W, C, 0, 0, &ba);
      float* S = raw<float>((size_t)B * T * T, &bs);
      bf16* Pm = raw<bf16>((size_t)B * T * T, &bp);
      const bool f16_ = f16;
      for (int hd = 0; hd < heads; ++hd) {
        View q = chan_slice(qk, hd * d, d);
        q.gst = nullptr;
        GemmB kb; kb.p = qk.p + C + hd * d; kb.n_alloc = T; kb.ld = 2 * C; kb.batch_stride = (int64_t)T * 2 * C; kb.f16 = f16;
        GemmEpilogue es;
        es.out_f32_nhwc = S;
        op([=](cudaStream_t s) { launch_conv_gemm(q, taps_1x1(), kb, T, es, s); }, PROF_ATTN, gemm_flops(q, 1, T));
        op([=](cudaStream_t s) { launch_softmax_rows(S, Pm, (int64_t)B * T, T, s, f16_); }, PROF_ATTN);
        View pv = mat(Pm, B, T / 128, 128, T, T);
        GemmB vb; vb.p = vt + (size_t)hd * d * T; vb.n_alloc = d; vb.ld = T; vb.batch_stride = (int64_t)C * T; vb.f16 = f16;
        GemmEpilogue eo;
        View o2 = chan_slice(ao, hd * d, d);
        o2.H = T / 128; o2.W = 128; o2.gst = nullptr;
        eo.out = o2;
        op([=](cudaStream_t s) { launch_conv_gemm(pv, taps_1x1(), vb, d, eo, s); }, PROF_ATTN, gemm_flops(pv, 1, d));
      }
      release(bs); release(bp); release(bqk); release(bvt);
    }
    {
      GemmEpilogue ep;
      ep.bias = blk.attn_proj.bias;
      ep.residual = x;
      ep.res_scale = 0.70710678118654752f;
      ep.out = dst;
      gemm(ao, taps_1x1(), blk.attn_proj, ep);
    }
    release(ba);
  }

  // ObjectAwareCrossAttention.forward (object_cross_unet.py:447-565): GN32 -> qkv 1x1 -> attention over image + layout
  // keys (conditioning-only operands come from the buffers lidm_layout_set_cond filled) -> proj_out + x
  void oaca_block(const OacaW* w, const View& x, const View& dst) {
    const int B = x.B, H = x.H, W = x.W, C = w->ch, T = H * W;
    Buf bg, bqkv, bao;
    View g = act(B, H, W, C, 0, 0, &bg);
    groupnorm(x, g, w->n_qkv, 1e-5f, false);
    View qkv = act(B, H, W, 3 * C, 0, 0, &bqkv);
    {
      GemmEpilogue ep;
      ep.bias = w->qkv.bias;
      ep.out = qkv;
      ep.out.gst = nullptr;
      gemm(g, taps_1x1(), w->qkv, ep);
    }
    release(bg);
    View ao = act(B, H, W, C, 0, 0, &bao);
    lidm_handle* h_ = h;
    op([=](cudaStream_t s) {
      if (w->pos_img == nullptr || h_->cond_B != B)
        throw Error(LIDM_ERR_STATE, "layout conditioning missing for this batch size: call lidm_layout_set_cond first");
      launch_oaca_attention(qkv.p, w->pos_img, w->pos_batch, w->klay, w->vlay, h_->cond_n_layout, ao, B, T, C, s);
    }, PROF_ATTN, 2.0 * B * (double)T * (T + 13) * (3 * C), 0, "oaca T" + std::to_string(T) + " C" + std::to_string(C));
    release(bqkv);
    {
      GemmEpilogue ep;
      ep.bias = w->proj.bias;
      ep.residual = x;
      ep.out = dst;
      gemm(ao, taps_1x1(), w->proj, ep);
    }
    release(bao);
  }

  // ResBlock._forward (openaimodel.py:256-276) / ResnetBlock.forward (model_lidm.py:127-147, temb None)
  void res_block(const ResW& r, const View& x, const View& dst, int kh, int kw, int pl, int pr, int pt, float eps) {
    const int B = x.B, H = x.H, W = x.W;
    Buf bg1, bh, bg2, bsk;
    View g1 = act(B, H, W, r.cin, pl, pr, &bg1);
    groupnorm(x, g1, r.n1, eps, true);
    View hmid = act(B, H, W, r.cout, 0, 0, &bh);
    {
      GemmEpilogue ep;
      ep.bias = r.c1.bias;
      ep.out = hmid;
      prep_gst(ep, r.cout, r.c1.n_alloc);
      const GemmB b = gb(r.c1);
      const ConvTaps taps = taps_rect(kh, kw, pl, pt);
      const int N = r.cout, emb_off = r.emb_off;
      Plan* P_ = P;
      op([=](cudaStream_t s) {
        GemmEpilogue e = ep;
        if (emb_off >= 0) { e.rowadd = P_->rowadd_base + emb_off; e.rowadd_ld = P_->rowadd_ld; }
        launch_conv_gemm(g1, taps, b, N, e, s);
      }, PROF_GEMM, gemm_flops(g1, taps.n, N), 0, gemm_label(g1, taps.n, N));
    }
    release(bg1);
    View g2 = act(B, H, W, r.cout, pl, pr, &bg2);
    groupnorm(hmid, g2, r.n2, eps, true);
    release(bh);
    static const bool no_fold = getenv("LIDM_NO_SKIP_FOLD") != nullptr;   // A/B switch
    if (r.has_skip && r.c2s.w != nullptr && !no_fold && x.wpitch == 0) {
      // out = conv2(h) + skip(x) as ONE GEMM: x rides along as one more K segment (GemmEpilogue::a2)
      GemmEpilogue ep;
      ep.bias = r.c2s.bias;
      ep.a2 = x;
      ep.out = dst;
      prep_gst(ep, r.cout, r.c2s.n_alloc);
      const GemmB b = gb(r.c2s);
      const ConvTaps taps = taps_rect(kh, kw, pl, pt);
      const int N = r.cout;
      op([=](cudaStream_t s) { launch_conv_gemm(g2, taps, b, N, ep, s); }, PROF_GEMM,
         gemm_flops(g2, taps.n, N) + gemm_flops(x, 1, N), 0, gemm_label(g2, taps.n, N) + " +skip" + std::to_string(r.cin));
      release(bg2);
      return;
    }
    View resid = x;
    if (r.has_skip) {
      View sk = act(B, H, W, r.cout, 0, 0, &bsk);
      GemmEpilogue ep;
      ep.bias = r.skip.bias;
      ep.out = sk;
      gemm(x, taps_1x1(), r.skip, ep);
      resid = sk;
    }
    {
      GemmEpilogue ep;
      ep.bias = r.c2.bias;
      ep.residual = resid;
      ep.out = dst;
      gemm(g2, taps_rect(kh, kw, pl, pt), r.c2, ep);
    }
    release(bg2);
    if (r.has_skip) release(bsk);
  }

  // AttentionBlock._forward + QKVAttentionLegacy (openaimodel.py:320-326, 358-374)
  void attn_block(const AttnW& a, const View& x, const View& dst) {
    const int B = x.B, H = x.H, W = x.W, C = a.ch, T = H * W;
    Buf bg, bqk, ba;
    View g = act(B, H, W, C, 0, 0, &bg);
    groupnorm(x, g, a.n, 1e-5f, false);
    const int heads = a.heads;
    // one packed (B,T,3C) = [q | k | v] tensor straight out of the qkv GEMM (TMA-store epilogue)
    View qkv = act(B, H, W, 3 * C, 0, 0, &bqk);
    {
      GemmEpilogue ep;
      ep.bias = a.qkv.bias;
      ep.out = qkv;
      ep.out.gst = nullptr;        // q | k | v feed the attention kernel, never a GroupNorm: no statistics pass
      gemm(g, taps_1x1(), a.qkv, ep);
    }
    release(bg);
    View ao = act(B, H, W, C, 0, 0, &ba);
    op([=](cudaStream_t s) { launch_attention_d32_packed(qkv.p, ao, B, T, heads, s, true); }, PROF_ATTN,
       4.0 * B * heads * (double)T * T * 32, 0, "attn T" + std::to_string(T) + " heads" + std::to_string(heads));
    release(bqk);
    {
      GemmEpilogue ep;
      ep.bias = a.proj.bias;
      ep.out = dst;
      gemm_res(ao, a.proj, ep, x);
    }
    release(ba);
  }

  // AttnBlock.forward (model_lidm.py:184-208): single head, d = C, scale C^-1/2 folded into the packed q rows
  void dec_attn_block(const AttnW& a, const View& x, const View& dst) {
    const int B = x.B, H = x.H, W = x.W, C = a.ch, T = H * W;
    Buf bg, bqk, bvt, bs, bp, ba;
    View g = act(B, H, W, C, 0, 0, &bg);
    groupnorm(x, g, a.n, 1e-6f, false);
    View qk = act(B, H, W, 2 * C, 0, 0, &bqk);
    bf16* vt = raw<bf16>((size_t)B * C * T, &bvt);
    {
      GemmEpilogue ep;
      ep.bias = a.qkv.bias;
      ep.out = qk;
      ep.split_n = 2 * C;
      ep.out_t = vt;
      gemm(g, taps_1x1(), a.qkv, ep);
    }
    release(bg);
    float* S = raw<float>((size_t)B * T * T, &bs);
    {
      View q = chan_slice(qk, 0, C);
      GemmB kb; kb.p = qk.p + C; kb.n_alloc = T; kb.ld = 2 * C; kb.batch_stride = (int64_t)T * 2 * C; kb.f16 = f16;
      GemmEpilogue ep;
      ep.out_f32_nhwc = S;
      op([=](cudaStream_t s) { launch_conv_gemm(q, taps_1x1(), kb, T, ep, s); }, PROF_GEMM, gemm_flops(q, 1, T));
    }
    bf16* Pm = raw<bf16>((size_t)B * T * T, &bp);
    const bool f16_ = f16;
    op([=](cudaStream_t s) { launch_softmax_rows(S, Pm, (int64_t)B * T, T, s, f16_); });
    release(bs);
    View ao = act(B, H, W, C, 0, 0, &ba);
    {
      View pv = mat(Pm, B, T / 128, 128, T, T);
      GemmB vb; vb.p = vt; vb.n_alloc = C; vb.ld = T; vb.batch_stride = (int64_t)C * T; vb.f16 = f16;
      GemmEpilogue ep;
      View o2 = ao; o2.H = T / 128; o2.W = 128; o2.gst = nullptr;
      ep.out = o2;
      op([=](cudaStream_t s) { launch_conv_gemm(pv, taps_1x1(), vb, C, ep, s); }, PROF_GEMM, gemm_flops(pv, 1, C));
    }
    release(bp); release(bqk); release(bvt);
    {
      GemmEpilogue ep;
      ep.bias = a.proj.bias;
      ep.residual = x;
      ep.out = dst;
      gemm(ao, taps_1x1(), a.proj, ep);
    }
    release(ba);
  }

  void layernorm(const View& x, const View& y, const NormW& n) {
    op([=](cudaStream_t s) { launch_layernorm(x, y, n.gamma, n.beta, 1e-5f, s); }, PROF_NORM, 0,
       4.0 * x.B * x.H * x.W * x.C, "ln C" + std::to_string(x.C) + " @" + std::to_string(x.H) + "x" + std::to_string(x.W));
  }
  // `stats`: the output feeds a GroupNorm (the producing GEMM emits its granule statistics); LayerNorm / attention consumers
  // take none, and the statistics pass is a fifth of a short-K epilogue
  void lin(const View& a, const ConvW& w, const View& out, const View* residual = nullptr, bool stats = false) {
    GemmEpilogue ep;
    ep.bias = w.bias;
    ep.out = out;
    if (!stats) ep.out.gst = nullptr;
    if (residual) gemm_res(a, w, ep, *residual);
    else gemm(a, taps_1x1(), w, ep);
  }

  // SpatialTransformer.forward / BasicTransformerBlock._forward / CrossAttention.forward / GEGLU
  // (lidm/modules/attention.py:250-261, 211-215, 170-193, 36-44).  ctx_kv: every block's K | V rows of the context,
  // (B * ctx_len padded to 128 rows, ctx_n) bf16, produced once per U-Net evaluation by one GEMM.
  void st_block(const STW& t, const View& x, const View& dst, const bf16* ctx_kv, int ctx_n, int ctx_len) {
    const int B = x.B, H = x.H, W = x.W, C = t.ch, T = H * W, heads = t.heads;
    Buf bg, bh;
    View g = act(B, H, W, C, 0, 0, &bg);
    groupnorm(x, g, t.n, 1e-6f, false);
    View hcur = act(B, H, W, C, 0, 0, &bh);
    lin(g, t.proj_in, hcur);
    release(bg);
    for (const STBlockW& k : t.blocks) {
      // x = attn1(norm1(x)) + x   (self-attention)
      Buf bn, bqkv, bao, bh1;
      View n1 = act(B, H, W, C, 0, 0, &bn);
      layernorm(hcur, n1, k.n1);
      View qkv = act(B, H, W, 3 * C, 0, 0, &bqkv);
      lin(n1, k.qkv1, qkv);
      release(bn);
      View ao = act(B, H, W, C, 0, 0, &bao);
      op([=](cudaStream_t s) { launch_attention_d32_packed(qkv.p, ao, B, T, heads, s, true); }, PROF_ATTN,
         4.0 * B * heads * (double)T * T * 32, 0, "attn T" + std::to_string(T) + " heads" + std::to_string(heads));
      release(bqkv);
      View h1 = act(B, H, W, C, 0, 0, &bh1);
      lin(ao, k.out1, h1, &hcur);
      release(bao);
      release(bh);
      // x = attn2(norm2(x), context) + x   (cross-attention)
      Buf bn2, bq2, bao2, bh2;
      View n2 = act(B, H, W, C, 0, 0, &bn2);
      layernorm(h1, n2, k.n2);
      View q2 = act(B, H, W, C, 0, 0, &bq2);
      lin(n2, k.q2, q2);
      release(bn2);
      View ao2 = act(B, H, W, C, 0, 0, &bao2);
      const int kv_col = k.kv_col;
      op([=](cudaStream_t s) {
        launch_cross_attention_d32(q2.p, q2.ld, ctx_kv, ctx_n, kv_col, kv_col + C, ctx_len, ao2, B, T, heads, s, true);
      }, PROF_ATTN, 4.0 * B * heads * (double)T * ctx_len * 32, 0,
         "xattn T" + std::to_string(T) + " L" + std::to_string(ctx_len) + " heads" + std::to_string(heads));
      release(bq2);
      View h2 = act(B, H, W, C, 0, 0, &bh2);
      lin(ao2, k.out2, h2, &h1);
      release(bao2);
      release(bh1);
      // x = ff(norm3(x)) + x   (GEGLU feed-forward)
      Buf bn3, bgl, bh3;
      View n3 = act(B, H, W, C, 0, 0, &bn3);
      layernorm(h2, n3, k.n3);
      View ffg = act(B, H, W, 4 * C, 0, 0, &bgl);
      {
        GemmEpilogue ep;                 // proj + GEGLU in one GEMM: the (B,T,8C) intermediate never reaches HBM
        ep.bias = k.ff0.bias;
        ep.out = ffg;
        ep.out.gst = nullptr;
        ep.geglu = true;
        gemm(n3, taps_1x1(), k.ff0, ep);
      }
      release(bn3);
      View h3 = act(B, H, W, C, 0, 0, &bh3);
      lin(ffg, k.ff2, h3, &h2);
      release(bgl);
      release(bh2);
      hcur = h3; bh = bh3;
    }
    lin(hcur, t.proj_out, dst, &x, true);
    release(bh);
  }

  // Strided circular convolution (kernel (kh,kw), stride (sh,sw), circular pad pl on the left of W, zero pad pt on top of H;
  // whatever the kernel reaches beyond the right / bottom edge is the circular wrap / zero as well).  Implicit GEMM: the
  // operand is the input tensor itself, read through TMA traversal strides (tap (ky,kx) of output pixel (h,w) = input
  // (h*sh + ky - pt, w*sw + kx - pl)); the circular wrap columns come from a halo, added by one copy when the producer wrote
  // none.  Shapes whose output does not tile into 128-pixel boxes fall back to channels-last im2col + GEMM.
  void conv_strided(const ConvW& c, const View& x, const View& dst, int sh, int sw, int pl, int pt) {
    static const bool no_implicit = getenv("LIDM_NO_STRIDED_TMA") != nullptr;     // A/B switch
    const int B = x.B, Ho = dst.H, Wo = dst.W, kh = c.kh, kw = c.kw;
    const int need_hr = std::max(0, (Wo - 1) * sw + kw - 1 - pl - (x.W - 1));
    const bool tiles = (Wo <= 128 && 128 % Wo == 0 && Ho % (128 / Wo) == 0) || Wo % 128 == 0;
    const bool implicit = !no_implicit && !precise && c.nseg == 1 && x.C % 64 == 0 && x.W == Wo * sw && x.H == Ho * sh && tiles && x.wpitch == 0 &&
                          std::min(Wo, 128) * sw <= 256 && (128 / std::min(Wo, 128)) * sh <= 256;
    GemmEpilogue ep;
    ep.bias = c.bias;
    ep.out = dst;
    if (implicit) {
      View xin = x;
      Buf bh;
      const bool need_copy = x.hl < pl || x.hr < need_hr;
      if (need_copy) {
        xin = act(B, x.H, x.W, x.C, pl, need_hr, &bh);
        const View xh = xin;
        op([=](cudaStream_t s) { launch_copy_with_halo(x, xh, s); });
      }
      ConvTaps taps = taps_rect(kh, kw, pl, pt);
      taps.sx = sw; taps.sy = sh;
      gemm(xin, taps, c, ep);
      if (need_copy) release(bh);
      return;
    }
    const int ntap = kh * kw;
    Buf bc;
    bf16* col = raw<bf16>((size_t)B * Ho * Wo * ntap * x.C, &bc);
    op([=](cudaStream_t s) { launch_im2col_nhwc(x, kh, kw, sh, pl, pt, Ho, Wo, col, s, sw); });
    View a = mat(col, B, Ho, Wo, ntap * x.C, ntap * x.C);
    gemm(a, taps_1x1(), c, ep);
    release(bc);
  }

  // Downsample.forward (openaimodel.py:159-161): circular 3x3 stride 2, pad 1
  void down(const ConvW& c, const View& x, const View& dst) { conv_strided(c, x, dst, 2, 2, 1, 1); }

  // Downsample.forward of the first-stage encoder (model_lidm.py:68-83): CircularConv2d kernel (kh,kw), stride (sh,sw),
  // circular pad (pl, .) on W / zero pad (pt, .) on H
  void down_strided(const ConvW& c, const View& x, const View& dst, int sh, int sw, int pl, int pt) {
    conv_strided(c, x, dst, sh, sw, pl, pt);
  }

  // Upsample.forward (openaimodel.py:108-118), nearest x2 followed by the circular 3x3 conv, without materialising
  // the 4x tensor: output pixel (2h+py, 2w+px) only ever sees input rows {h-1+py, h+py} and columns {w-1+px, w+px}, so each
  // of the four output parities is a 2x2 conv on the low-resolution input whose taps are sums of the 3x3 taps
  // (cpar[py*2+px], folded in fp32 at pack time): 16 C^2 instead of 36 C^2 MACs per input pixel.  Each GEMM writes its
  // parity straight into the 2x output through a strided view (ld and row pitch doubled).
  void up_folded(const ConvW* cpar, const ConvW& c, const View& x, const View& dst) {
    if (dst.hl + dst.hr != 0 || dst.wpitch != 0) throw Error(LIDM_ERR_INVALID, "internal: upsample output must be halo-free");
    Buf bx;
    View xh = act(x.B, x.H, x.W, x.C, 1, 1, &bx);
    op([=](cudaStream_t s) { launch_copy_with_halo(x, xh, s); });
    for (int py = 0; py < 2; ++py)
      for (int px = 0; px < 2; ++px) {
        ConvTaps t;
        t.n = 4;
        for (int a = 0; a < 2; ++a)
          for (int bb = 0; bb < 2; ++bb) { t.dy[a * 2 + bb] = (int8_t)(a - 1 + py); t.dx[a * 2 + bb] = (int8_t)(bb - 1 + px); }
        View o = dst;
        o.H = x.H; o.W = x.W;
        o.p = dst.p + (size_t)(py * dst.W + px) * dst.ld;
        o.ld = 2 * dst.ld;
        o.wpitch = dst.W;                 // one output row pair = 2 * dst.W pixels of dst.ld = dst.W pixels of 2 * dst.ld
        o.gst_slot0 = (py * 2 + px) * (x.H * x.W / 128);   // each parity fills its own quarter of the statistics slots
        GemmEpilogue ep;
        ep.bias = c.bias;
        ep.out = o;
        gemm(xh, t, cpar[py * 2 + px], ep);
      }
    release(bx);
  }
  void up(const ConvW& c, const View& x, const View& dst) {
    Buf bu;
    View u = act(x.B, x.H * 2, x.W * 2, x.C, 1, 1, &bu);
    op([=](cudaStream_t s) { launch_upsample_nearest2x(x, u, s); });
    GemmEpilogue ep;
    ep.bias = c.bias;
    ep.out = dst;
    gemm(u, taps_rect(3, 3, 1, 1), c, ep);
    release(bu);
  }
};

// ------------------------------------------------------------------------------------------- U-Net plan
void build_unet_plan_pass(lidm_handle* h, Plan* P, bool dry, size_t* high) {
  Builder b{h, P, ArenaPlanner(), dry};
  b.f16 = h->unet_prec == LIDM_PREC_FP16;
  const lidm_config& cfg = h->cfg;
  const int B = P->B;
  const int n_in = (int)h->in_blocks.size(), n_out = (int)h->out_blocks.size();
  // resolution of every input block's output
  std::vector<int> rh(n_in), rw(n_in);
  {
    int H = cfg.latent_h, W = cfg.latent_w;
    for (int k = 0; k < n_in; ++k) {
      const Layer& L0 = h->in_blocks[k][0];
      if (L0.kind == Layer::DOWN || (L0.kind == Layer::RES && L0.r.updown == 2)) { H /= 2; W /= 2; }
      rh[k] = H; rw[k] = W;
    }
  }
  const bool zw = h->is_layout;        // plain zero-padded convolutions (no circular halo)
  // concat buffers: output block i consumes cat[h_prev (Ca) | hs[n_in-1-i] (Cb)]
  std::vector<View> cat(n_out);
  std::vector<Buf> catbuf(n_out);
  std::vector<int> Ca(n_out);
  {
    int ch_prev = h->mid_block.back().cout;
    for (int i = 0; i < n_out; ++i) {
      const int k = n_in - 1 - i;
      Ca[i] = ch_prev;
      cat[i] = b.act(B, rh[k], rw[k], ch_prev + h->in_chans[k], 0, 0, &catbuf[i]);
      ch_prev = h->out_blocks[i][0].cout;
    }
  }
  // cross-attention context: fp32 (B, L, context_dim) -> bf16 rows (padded to a multiple of 128) -> one GEMM against the
  // stacked [to_k ; to_v] rows of every transformer block -> ctx_kv (rows, ctx_n)
  const bf16* ctx_kv = nullptr;
  Buf bctx_kv;
  const int ctx_len = P->ctx_len;
  if (h->has_st) {
    if (ctx_len <= 0) throw Error(LIDM_ERR_INVALID, "this U-Net needs a cross-attention context");
    const int D = cfg.context_dim;
    const int64_t rows = (int64_t)B * ctx_len, rows_pad = (rows + 127) / 128 * 128;
    Buf bc16;
    bf16* c16 = b.raw<bf16>((size_t)rows_pad * D, &bc16);
    b.op([=](cudaStream_t s) { launch_f32_rows_to_bf16(P->context, rows, rows_pad, D, c16, s); });
    bf16* kv = b.raw<bf16>((size_t)rows_pad * h->ctx_n, &bctx_kv);
    View a; a.p = c16; a.B = 1; a.H = 1; a.W = (int)rows_pad; a.C = D; a.ld = D;
    View o; o.p = kv; o.B = 1; o.H = 1; o.W = (int)rows_pad; o.C = h->ctx_n; o.ld = h->ctx_n;
    GemmB wb; wb.p = h->ctx_w; wb.n_alloc = h->ctx_n; wb.ld = D;
    GemmEpilogue ep;
    ep.out = o;
    const int N = h->ctx_n;
    b.op([=](cudaStream_t s) { launch_conv_gemm(a, taps_1x1(), wb, N, ep, s); }, PROF_GEMM, gemm_flops(a, 1, N), 0,
         "gemm context kv " + std::to_string(D) + "->" + std::to_string(N) + " rows" + std::to_string(rows_pad));
    b.release(bc16);
    ctx_kv = kv;
  }
  auto run_layers = [&](const std::vector<Layer>& layers, View x, const View& dst) {
    Buf prev_buf; bool have_prev = false;
    for (size_t j = 0; j < layers.size(); ++j) {
      const Layer& L = layers[j];
      const bool last = (j + 1 == layers.size());
      int Ho = x.H, Wo = x.W;
      if (L.kind == Layer::DOWN || (L.kind == Layer::RES && L.r.updown == 2)) { Ho /= 2; Wo /= 2; }
      if (L.kind == Layer::UP || (L.kind == Layer::RES && L.r.updown == 1)) { Ho *= 2; Wo *= 2; }
      Buf ob; View o;
      if (last) o = dst;
      else o = b.act(B, Ho, Wo, L.cout, 0, 0, &ob);
      switch (L.kind) {
        case Layer::RES:
          if (zw) b.res_block_film(L.r, x, o);
          else b.res_block(L.r, x, o, 3, 3, 1, 1, 1, 1e-5f);
          break;
        case Layer::OACA: b.oaca_block(&L.oa, x, o); break;
        case Layer::ATTN: b.attn_block(L.a, x, o); break;
        case Layer::ST: b.st_block(L.st, x, o, ctx_kv, h->ctx_n, ctx_len); break;
        case Layer::DOWN: b.down(L.c, x, o); break;
        case Layer::UP:
          if (L.cpar[0].w != nullptr) b.up_folded(L.cpar, L.c, x, o);
          else b.up(L.c, x, o);
          break;
        case Layer::CONV: throw Error(LIDM_ERR_INVALID, "unexpected conv layer");
      }
      if (have_prev) b.release(prev_buf);
      prev_buf = ob; have_prev = !last;
      x = o;
    }
  };

  // input_blocks[0]: conv 3x3 in_channels -> model_channels via fp32-NCHW im2col (K padded to 128)
  {
    const int H = cfg.latent_h, W = cfg.latent_w, kpad = h->in_blocks[0][0].c.k_alloc;
    Buf bc;
    bf16* col = b.raw<bf16>((size_t)B * H * W * kpad, &bc);
    const int Cin = cfg.in_channels;
    const bool f16 = b.f16;
    b.op([=](cudaStream_t s) { launch_im2col_nchw_f32(P->xin, B, Cin, H, W, 3, 3, 1, 1, col, kpad, s, f16, zw); });
    View a = b.mat(col, B, H, W, kpad, kpad);
    GemmEpilogue ep;
    ep.bias = h->in_blocks[0][0].c.bias;
    ep.out = Builder::chan_slice(cat[n_out - 1], Ca[n_out - 1], h->in_chans[0]);
    b.gemm(a, taps_1x1(), h->in_blocks[0][0].c, ep);
    b.release(bc);
  }
  for (int k = 1; k < n_in; ++k) {
    const int iprev = n_out - 1 - (k - 1), icur = n_out - 1 - k;
    View x = Builder::chan_slice(cat[iprev], Ca[iprev], h->in_chans[k - 1]);
    View dst = Builder::chan_slice(cat[icur], Ca[icur], h->in_chans[k]);
    run_layers(h->in_blocks[k], x, dst);
  }
  {
    View x = Builder::chan_slice(cat[0], Ca[0], h->in_chans[n_in - 1]);
    View dst = Builder::chan_slice(cat[0], 0, Ca[0]);
    run_layers(h->mid_block, x, dst);
  }
  Buf bfinal;
  View hfinal;
  for (int i = 0; i < n_out; ++i) {
    View dst;
    if (i + 1 < n_out) dst = Builder::chan_slice(cat[i + 1], 0, Ca[i + 1]);
    else { hfinal = b.act(B, cfg.latent_h, cfg.latent_w, h->out_blocks[i].back().cout, 0, 0, &bfinal); dst = hfinal; }
    run_layers(h->out_blocks[i], cat[i], dst);
    b.release(catbuf[i]);
  }
  // out: GroupNorm32 + SiLU + conv3x3 (zero-init in the reference) -> eps (fp32 NCHW) [+ fused DDIM update]
  {
    Buf bg;
    View g = b.act(B, hfinal.H, hfinal.W, hfinal.C, zw ? 0 : 1, zw ? 0 : 1, &bg);
    b.groupnorm(hfinal, g, h->out_norm, 1e-5f, true);
    const GemmB wb = Builder::gb(h->out_conv);
    const float* bias = h->out_conv.bias;
    const int N = h->out_conv.cout;
    const ConvTaps taps = zw ? Builder::taps_zero3x3() : taps_rect(3, 3, 1, 1);
    b.op([=](cudaStream_t s) {
      GemmEpilogue ep;
      ep.bias = bias;
      ep.out_f32_nchw = P->out;
      if (P->ddim_x_prev != nullptr) {
        ep.ddim_x = P->x; ep.ddim_noise = P->ddim_noise; ep.ddim_x_prev = P->ddim_x_prev;
        ep.ddim_pred_x0 = P->ddim_pred_x0; ep.ddim_coef = P->ddim_coef;
      }
      launch_conv_gemm(g, taps, wb, N, ep, s);
    }, PROF_GEMM, gemm_flops(g, taps.n, N));
    b.release(bg);
    b.release(bfinal);
  }
  if (h->has_st) b.release(bctx_kv);
  *high = b.ap.high();
}

// ------------------------------------------------------------------------------------------- EfficientUNet plan
// EfficientUNet.forward (efficient_unet.py:262-295): [x | Fourier features] -> in_conv -> 4 down blocks -> 4 up blocks (skip
// concatenations written in place) -> out_conv -> eps (fp32 NCHW) [+ fused DDIM update]
void build_eff_plan_pass(lidm_handle* h, Plan* P, bool dry, size_t* high) {
  Builder b{h, P, ArenaPlanner(), dry};
  b.f16 = h->unet_prec == LIDM_PREC_FP16;
  b.gn_groups = h->cfg.eff_gn_groups;
  const lidm_config& cfg = h->cfg;
  const float eps = cfg.eff_gn_eps;
  const int B = P->B, H = cfg.latent_h, W = cfg.latent_w, Cx = cfg.in_channels, Ce = h->eff_cenc_ch, Cin = Cx + Ce;
  const int C0 = cfg.model_channels;
  int C[5] = {C0, 0, 0, 0, 0};
  for (int i = 0; i < 4; ++i) C[i + 1] = C0 * cfg.channel_mult[i];
  // skip-concatenation buffers [up-path tensor | down-path tensor] at levels 1..3
  Buf bcat[4];
  View cat[4];
  for (int l = 1; l <= 3; ++l) cat[l] = b.act(B, H >> (l - 1), W >> (l - 1), 2 * C[l], 0, 0, &bcat[l]);
  Buf bh0;
  View h0 = b.act(B, H, W, C[0], 0, 0, &bh0);
  h0.gst = nullptr;     // written by the CUDA-core input convolution: the first GroupNorm takes its own statistics
  if (Cx <= 2 && C[0] % 8 == 0) {
    b.op([=](cudaStream_t s) { launch_eff_in_conv(P->x, h->eff_in_w, h->eff_in_map, Cx, h0, s); });
  } else {
    throw Error(LIDM_ERR_INVALID, "EfficientUNet input convolution: one or two image channels");
  }
  // one Block (efficient_unet.py:113-186)
  auto run_block = [&](const lidm_handle::EffBlock& blk, View x, Buf xbuf, bool own_x, const View& dst) {
    // [ring conv 3x3 + FIR x1/2]
    if (blk.has_down) {
      Buf bxh, by, bz;
      View xh = b.act(B, x.H, x.W, x.C, 1, 1, &bxh);
      b.op([=](cudaStream_t s) { launch_copy_with_halo(x, xh, s); });
      if (own_x) b.release(xbuf);
      View y = b.act(B, x.H, x.W, blk.ch, 0, 0, &by);
      GemmEpilogue ep;
      ep.bias = blk.down.bias;
      ep.out = y;
      ep.out.gst = nullptr;
      b.gemm(xh, taps_rect(3, 3, 1, 1), blk.down, ep);
      b.release(bxh);
      View z = b.act(B, x.H / 2, x.W / 2, blk.ch, 0, 0, &bz);
      z.gst = nullptr;                      // produced by the FIR kernel: the first GroupNorm takes its own statistics
      b.op([=](cudaStream_t s) { launch_fir_down2(y, z, s); });
      b.release(by);
      x = z; xbuf = bz; own_x = true;
    }
    const int n = (int)blk.res.size();
    const bool tail = blk.has_attn || blk.has_up;
    for (int i = 0; i < n; ++i) {
      const bool last = (i + 1 == n) && !tail;
      Buf bo; View o;
      if (last) o = dst;
      else o = b.act(B, x.H, x.W, blk.ch, 0, 0, &bo);
      b.res_block_eff(blk.res[i], x, o, eps);
      if (own_x) b.release(xbuf);
      x = o; xbuf = bo; own_x = !last;
    }
    if (blk.has_attn) {
      const bool last = !blk.has_up;
      Buf bo; View o;
      if (last) o = dst;
      else o = b.act(B, x.H, x.W, blk.ch, 0, 0, &bo);
      b.attn_block_eff(blk, x, o, eps);
      if (own_x) b.release(xbuf);
      x = o; xbuf = bo; own_x = !last;
    }
    if (blk.has_up) {
      Buf bu;
      View u = b.act(B, 2 * x.H, 2 * x.W, blk.ch, 1, 1, &bu);
      b.op([=](cudaStream_t s) { launch_fir_up2(x, u, s); });
      if (own_x) b.release(xbuf);
      GemmEpilogue ep;
      ep.bias = blk.up.bias;
      ep.out = dst;
      b.gemm(u, taps_rect(3, 3, 1, 1), blk.up, ep);
      b.release(bu);
    }
  };
  // down path: h1, h2, h3 land in the second halves of the concatenation buffers
  run_block(h->eff_d[0], h0, bh0, true, Builder::chan_slice(cat[1], C[1], C[1]));
  run_block(h->eff_d[1], Builder::chan_slice(cat[1], C[1], C[1]), Buf(), false, Builder::chan_slice(cat[2], C[2], C[2]));
  run_block(h->eff_d[2], Builder::chan_slice(cat[2], C[2], C[2]), Buf(), false, Builder::chan_slice(cat[3], C[3], C[3]));
  Buf bh4;
  View h4 = b.act(B, H >> 3, W >> 3, C[4], 0, 0, &bh4);
  run_block(h->eff_d[3], Builder::chan_slice(cat[3], C[3], C[3]), Buf(), false, h4);
  // up path
  run_block(h->eff_u[3], h4, bh4, true, Builder::chan_slice(cat[3], 0, C[3]));
  run_block(h->eff_u[2], cat[3], bcat[3], true, Builder::chan_slice(cat[2], 0, C[2]));
  run_block(h->eff_u[1], cat[2], bcat[2], true, Builder::chan_slice(cat[1], 0, C[1]));
  Buf bhf;
  View hf = b.act(B, H, W, C[0], 0, 0, &bhf);
  run_block(h->eff_u[0], cat[1], bcat[1], true, hf);
  // out_conv: ring conv 3x3 straight on h (no norm) -> eps [+ DDIM update]
  {
    Buf bg;
    View g = b.act(B, H, W, C[0], 1, 1, &bg);
    b.op([=](cudaStream_t s) { launch_copy_with_halo(hf, g, s); });
    b.release(bhf);
    const GemmB wb = Builder::gb(h->out_conv);
    const float* bias = h->out_conv.bias;
    const int N = h->out_conv.cout;
    const ConvTaps taps = taps_rect(3, 3, 1, 1);
    b.op([=](cudaStream_t s) {
      GemmEpilogue ep;
      ep.bias = bias;
      ep.out_f32_nchw = P->out;
      if (P->ddim_x_prev != nullptr) {
        ep.ddim_x = P->x; ep.ddim_noise = P->ddim_noise; ep.ddim_x_prev = P->ddim_x_prev;
        ep.ddim_pred_x0 = P->ddim_pred_x0; ep.ddim_coef = P->ddim_coef;
      }
      launch_conv_gemm(g, taps, wb, N, ep, s);
    }, PROF_GEMM, gemm_flops(g, taps.n, N));
    b.release(bg);
  }
  *high = b.ap.high();
}

// ------------------------------------------------------------------------------------------- decoder plan
void build_dec_plan_pass(lidm_handle* h, Plan* P, bool dry, size_t* high) {
  Builder b{h, P, ArenaPlanner(), dry};
  b.f16 = h->ae_prec == LIDM_PREC_FP16;
  const lidm_config& cfg = h->cfg;
  const int B = P->B, lh = cfg.latent_h, lw = cfg.latent_w, zc = cfg.z_channels;
  Buf bzq, bcol;
  float* zq = b.raw<float>((size_t)B * zc * lh * lw, &bzq);
  {
    const float inv_scale = 1.0f / cfg.scale_factor;
    const int n_embed = cfg.n_embed;
    b.op([=](cudaStream_t s) {
      launch_vq(P->x, B, zc, lh * lw, h->codebook, h->cb_norm, n_embed, P->quantize, h->pq_w, h->pq_b, inv_scale, zq,
                P->idx_out, s);
    });
  }
  const int kpad = h->dec_conv_in.k_alloc;
  bf16* col = b.raw<bf16>((size_t)B * lh * lw * kpad, &bcol);
  const bool f16 = b.f16;
  b.op([=](cudaStream_t s) { launch_im2col_nchw_f32(zq, B, zc, lh, lw, 3, 3, 1, 1, col, kpad, s, f16); });
  b.release(bzq);
  Buf bx;
  View x = b.act(B, lh, lw, h->dec_top, 0, 0, &bx);
  {
    View a = b.mat(col, B, lh, lw, kpad, kpad);
    GemmEpilogue ep;
    ep.bias = h->dec_conv_in.bias;
    ep.out = x;
    b.gemm(a, taps_1x1(), h->dec_conv_in, ep);
  }
  b.release(bcol);
  auto step = [&](std::function<void(const View&, const View&)> f, int Ho, int Wo, int C, int hl, int hr) {
    Buf bo;
    View o = b.act(B, Ho, Wo, C, hl, hr, &bo);
    f(x, o);
    b.release(bx);
    bx = bo; x = o;
  };
  step([&](const View& i, const View& o) { b.res_block(h->dec_mid1, i, o, 3, 3, 1, 1, 1, 1e-6f); }, lh, lw, h->dec_top, 0, 0);
  step([&](const View& i, const View& o) { b.dec_attn_block(h->dec_attn, i, o); }, lh, lw, h->dec_top, 0, 0);
  step([&](const View& i, const View& o) { b.res_block(h->dec_mid2, i, o, 3, 3, 1, 1, 1, 1e-6f); }, lh, lw, h->dec_top, 0, 0);
  int H = lh, W = lw;
  for (int lv = cfg.ae_n_ch_mult - 1; lv >= 0; --lv) {
    const DecLevel& L = h->dec_levels[lv];
    const int pl = L.kw == 3 ? 1 : 1, pr = L.kw == 3 ? 1 : 2, pt = L.kh == 3 ? 1 : 0;
    for (const ResW& r : L.blocks)
      step([&](const View& i, const View& o) { b.res_block(r, i, o, L.kh, L.kw, pl, pr, pt, 1e-6f); }, H, W, r.cout, 0, 0);
    if (L.has_up) {
      // Upsample.forward (model_lidm.py:57-61): bilinear align_corners=True, then circular conv
      const int Ho = H * L.sh, Wo = W * L.sw;
      const int ukh = L.up.kh, ukw = L.up.kw;
      const int uh = (ukw - 1) / 2, upt = (ukh - 1) / 2;
      step([&](const View& i, const View& o) { b.op([=](cudaStream_t s) { launch_upsample_bilinear(i, o, s); }); }, Ho, Wo,
           L.ch, uh, uh);
      H = Ho; W = Wo;
      step([&](const View& i, const View& o) {
        GemmEpilogue ep;
        ep.bias = L.up.bias;
        ep.out = o;
        b.gemm(i, taps_rect(ukh, ukw, uh, upt), L.up, ep);
      }, H, W, L.ch, 0, 0);
    }
  }
  // norm_out + swish + conv_out (1,4), pad (1,2)
  {
    Buf bg;
    View g = b.act(B, H, W, h->dec_last, 1, 2, &bg);
    b.groupnorm(x, g, h->dec_norm_out, 1e-6f, true);
    b.release(bx);
    const GemmB wb = Builder::gb(h->dec_conv_out);
    const float* bias = h->dec_conv_out.bias;
    const int N = h->dec_conv_out.cout;
    const ConvTaps taps = taps_rect(1, 4, 1, 0);
    if (cfg.ae_use_mask) {
      Buf bd;
      float* dec = b.raw<float>((size_t)B * N * H * W, &bd);
      b.op([=](cudaStream_t s) {
        GemmEpilogue ep; ep.bias = bias; ep.out_f32_nchw = dec;
        launch_conv_gemm(g, taps, wb, N, ep, s);
      }, PROF_GEMM, gemm_flops(g, taps.n, N));
      const int HW = H * W;
      b.op([=](cudaStream_t s) { launch_mask_select(dec, B, HW, P->out, s); });
      b.release(bd);
    } else {
      b.op([=](cudaStream_t s) {
        GemmEpilogue ep; ep.bias = bias; ep.out_f32_nchw = P->out;
        launch_conv_gemm(g, taps, wb, N, ep, s);
      }, PROF_GEMM, gemm_flops(g, taps.n, N));
    }
    b.release(bg);
  }
  *high = b.ap.high();
}

// ------------------------------------------------------------------------------------------- encoder plan
// Encoder.forward + quant_conv (model_lidm.py:284-312, autoencoder.py:285-288); image (B,Cin,H,W) fp32 -> (B,embed,h,w)
void build_enc_plan_pass(lidm_handle* h, Plan* P, bool dry, size_t* high) {
  Builder b{h, P, ArenaPlanner(), dry};
  b.f16 = h->ae_prec == LIDM_PREC_FP16;
  const lidm_config& cfg = h->cfg;
  const int B = P->B;
  int H = h->img_h, W = h->img_w;
  const int kpad = h->enc_conv_in.k_alloc;
  Buf bcol, bx;
  bf16* col = b.raw<bf16>((size_t)B * H * W * kpad, &bcol);
  const int Cin = cfg.ae_in_channels;
  const bool f16 = b.f16;
  b.op([=](cudaStream_t s) { launch_im2col_nchw_f32(P->x, B, Cin, H, W, 3, 3, 1, 1, col, kpad, s, f16); });
  View x = b.act(B, H, W, cfg.ae_ch, 0, 0, &bx);
  {
    View a = b.mat(col, B, H, W, kpad, kpad);
    GemmEpilogue ep;
    ep.bias = h->enc_conv_in.bias;
    ep.out = x;
    b.gemm(a, taps_1x1(), h->enc_conv_in, ep);
  }
  b.release(bcol);
  auto step = [&](std::function<void(const View&, const View&)> f, int Ho, int Wo, int C) {
    Buf bo;
    View o = b.act(B, Ho, Wo, C, 0, 0, &bo);
    f(x, o);
    b.release(bx);
    bx = bo; x = o;
  };
  for (const EncLevel& L : h->enc_levels) {
    for (const ResW& r : L.blocks)
      step([&](const View& i, const View& o) { b.res_block(r, i, o, 3, 3, 1, 1, 1, 1e-6f); }, H, W, r.cout);
    if (L.has_down) {
      const int Ho = H / L.sh, Wo = W / L.sw;
      step([&](const View& i, const View& o) { b.down_strided(L.down, i, o, L.sh, L.sw, L.pl, L.pt); }, Ho, Wo, L.ch);
      H = Ho; W = Wo;
    }
  }
  step([&](const View& i, const View& o) { b.res_block(h->enc_mid1, i, o, 3, 3, 1, 1, 1, 1e-6f); }, H, W, h->enc_top);
  step([&](const View& i, const View& o) { b.dec_attn_block(h->enc_attn, i, o); }, H, W, h->enc_top);
  step([&](const View& i, const View& o) { b.res_block(h->enc_mid2, i, o, 3, 3, 1, 1, 1, 1e-6f); }, H, W, h->enc_top);
  {
    Buf bg;
    View g = b.act(B, H, W, h->enc_top, 1, 1, &bg);
    b.groupnorm(x, g, h->enc_norm_out, 1e-6f, true);
    b.release(bx);
    const GemmB wb = Builder::gb(h->enc_conv_out);
    const float* bias = h->enc_conv_out.bias;
    const int N = h->enc_conv_out.cout;
    const ConvTaps taps = taps_rect(3, 3, 1, 1);
    b.op([=](cudaStream_t s) {
      GemmEpilogue ep; ep.bias = bias; ep.out_f32_nchw = P->out;
      launch_conv_gemm(g, taps, wb, N, ep, s);
    }, PROF_GEMM, gemm_flops(g, taps.n, N));
    b.release(bg);
  }
  *high = b.ap.high();
}

// ------------------------------------------------------------------------------------------- precise plans
void build_unet_plan_pass_p(lidm_handle* h, Plan* P, bool dry, size_t* high) {
  Builder b{h, P, ArenaPlanner(), dry};
  b.precise = true;
  const lidm_config& cfg = h->cfg;
  const int B = P->B;
  const int n_in = (int)h->in_blocks.size(), n_out = (int)h->out_blocks.size();
  std::vector<int> rh(n_in), rw(n_in);
  {
    int H = cfg.latent_h, W = cfg.latent_w;
    for (int k = 0; k < n_in; ++k) {
      if (h->in_blocks[k][0].kind == Layer::DOWN) { H /= 2; W /= 2; }
      rh[k] = H; rw[k] = W;
    }
  }
  std::vector<ViewF> cat(n_out);
  std::vector<Buf> catbuf(n_out);
  std::vector<int> Ca(n_out);
  {
    int ch_prev = h->mid_block.back().cout;
    for (int i = 0; i < n_out; ++i) {
      const int k = n_in - 1 - i;
      Ca[i] = ch_prev;
      cat[i] = b.actf(B, rh[k], rw[k], ch_prev + h->in_chans[k], &catbuf[i]);
      ch_prev = h->out_blocks[i][0].cout;
    }
  }
  auto run_layers = [&](const std::vector<Layer>& layers, ViewF x, const ViewF& dst) {
    Buf prev_buf; bool have_prev = false;
    for (size_t j = 0; j < layers.size(); ++j) {
      const Layer& L = layers[j];
      const bool last = (j + 1 == layers.size());
      int Ho = x.H, Wo = x.W;
      if (L.kind == Layer::DOWN) { Ho /= 2; Wo /= 2; }
      if (L.kind == Layer::UP) { Ho *= 2; Wo *= 2; }
      Buf ob; ViewF o;
      if (last) o = dst;
      else o = b.actf(B, Ho, Wo, L.cout, &ob);
      switch (L.kind) {
        case Layer::RES: b.res_block_p(L.r, x, o, 3, 3, 1, 1, 1, 1e-5f); break;
        case Layer::ATTN: b.attn_block_p(L.a, x, o); break;
        case Layer::ST: throw Error(LIDM_ERR_INVALID, "precise mode does not cover SpatialTransformer U-Nets");
        case Layer::DOWN: b.down_p(L.c, x, o); break;
        case Layer::UP: b.up_p(L.c, x, o); break;
        case Layer::CONV: throw Error(LIDM_ERR_INVALID, "unexpected conv layer");
      }
      if (have_prev) b.release(prev_buf);
      prev_buf = ob; have_prev = !last;
      x = o;
    }
  };
  {
    const int H = cfg.latent_h, W = cfg.latent_w;
    const ConvW& cw = h->in_blocks[0][0].c;
    const int kpad = cw.k_alloc / cw.nseg;
    Buf bc;
    bf16* col = b.raw<bf16>((size_t)B * H * W * 2 * kpad, &bc);
    const int Cin = cfg.in_channels;
    b.op([=](cudaStream_t s) { launch_im2col_nchw_f32_hl(P->xin, B, Cin, H, W, 3, 3, 1, 1, col, kpad, s); });
    View a; a.p = col; a.B = B; a.H = H; a.W = W; a.C = kpad; a.ld = 2 * kpad; a.lo_off = kpad;
    GemmEpilogue ep;
    ep.bias = cw.bias;
    Builder::set_out_f(ep, Builder::chan_slice_f(cat[n_out - 1], Ca[n_out - 1], h->in_chans[0]));
    b.gemm(a, taps_1x1(), cw, ep);
    b.release(bc);
  }
  for (int k = 1; k < n_in; ++k) {
    const int iprev = n_out - 1 - (k - 1), icur = n_out - 1 - k;
    run_layers(h->in_blocks[k], Builder::chan_slice_f(cat[iprev], Ca[iprev], h->in_chans[k - 1]),
               Builder::chan_slice_f(cat[icur], Ca[icur], h->in_chans[k]));
  }
  run_layers(h->mid_block, Builder::chan_slice_f(cat[0], Ca[0], h->in_chans[n_in - 1]),
             Builder::chan_slice_f(cat[0], 0, Ca[0]));
  Buf bfinal;
  ViewF hfinal;
  for (int i = 0; i < n_out; ++i) {
    ViewF dst;
    if (i + 1 < n_out) dst = Builder::chan_slice_f(cat[i + 1], 0, Ca[i + 1]);
    else { hfinal = b.actf(B, cfg.latent_h, cfg.latent_w, h->out_blocks[i].back().cout, &bfinal); dst = hfinal; }
    run_layers(h->out_blocks[i], cat[i], dst);
    b.release(catbuf[i]);
  }
  {
    Buf bg;
    View g = b.act_hl(B, hfinal.H, hfinal.W, hfinal.C, 1, 1, &bg);
    b.groupnorm_f(hfinal, g, h->out_norm, 1e-5f, true);
    GemmB wb; wb.p = h->out_conv.w; wb.n_alloc = h->out_conv.n_alloc; wb.ld = h->out_conv.k_alloc; wb.nseg = h->out_conv.nseg;
    const float* bias = h->out_conv.bias;
    const int N = h->out_conv.cout;
    const ConvTaps taps = taps_rect(3, 3, 1, 1);
    b.op([=](cudaStream_t s) {
      GemmEpilogue ep;
      ep.bias = bias;
      ep.out_f32_nchw = P->out;
      if (P->ddim_x_prev != nullptr) {
        ep.ddim_x = P->x; ep.ddim_noise = P->ddim_noise; ep.ddim_x_prev = P->ddim_x_prev;
        ep.ddim_pred_x0 = P->ddim_pred_x0; ep.ddim_coef = P->ddim_coef;
      }
      launch_conv_gemm(g, taps, wb, N, ep, s);
    }, PROF_GEMM, gemm_flops(g, taps.n, N) * wb.nseg);
    b.release(bg);
    b.release(bfinal);
  }
  *high = b.ap.high();
}

void build_dec_plan_pass_p(lidm_handle* h, Plan* P, bool dry, size_t* high) {
  Builder b{h, P, ArenaPlanner(), dry};
  b.precise = true;
  const lidm_config& cfg = h->cfg;
  const int B = P->B, lh = cfg.latent_h, lw = cfg.latent_w, zc = cfg.z_channels;
  Buf bzq, bcol;
  float* zq = b.raw<float>((size_t)B * zc * lh * lw, &bzq);
  {
    const float inv_scale = 1.0f / cfg.scale_factor;
    const int n_embed = cfg.n_embed;
    b.op([=](cudaStream_t s) {
      launch_vq(P->x, B, zc, lh * lw, h->codebook, h->cb_norm, n_embed, P->quantize, h->pq_w, h->pq_b, inv_scale, zq,
                P->idx_out, s);
    });
  }
  const int kpad = h->dec_conv_in.k_alloc / h->dec_conv_in.nseg;
  bf16* col = b.raw<bf16>((size_t)B * lh * lw * 2 * kpad, &bcol);
  b.op([=](cudaStream_t s) { launch_im2col_nchw_f32_hl(zq, B, zc, lh, lw, 3, 3, 1, 1, col, kpad, s); });
  b.release(bzq);
  Buf bx;
  ViewF x = b.actf(B, lh, lw, h->dec_top, &bx);
  {
    View a; a.p = col; a.B = B; a.H = lh; a.W = lw; a.C = kpad; a.ld = 2 * kpad; a.lo_off = kpad;
    GemmEpilogue ep;
    ep.bias = h->dec_conv_in.bias;
    Builder::set_out_f(ep, x);
    b.gemm(a, taps_1x1(), h->dec_conv_in, ep);
  }
  b.release(bcol);
  auto step = [&](std::function<void(const ViewF&, const ViewF&)> f, int Ho, int Wo, int C) {
    Buf bo;
    ViewF o = b.actf(B, Ho, Wo, C, &bo);
    f(x, o);
    b.release(bx);
    bx = bo; x = o;
  };
  step([&](const ViewF& i, const ViewF& o) { b.res_block_p(h->dec_mid1, i, o, 3, 3, 1, 1, 1, 1e-6f); }, lh, lw, h->dec_top);
  step([&](const ViewF& i, const ViewF& o) { b.dec_attn_block_p(h->dec_attn, i, o); }, lh, lw, h->dec_top);
  step([&](const ViewF& i, const ViewF& o) { b.res_block_p(h->dec_mid2, i, o, 3, 3, 1, 1, 1, 1e-6f); }, lh, lw, h->dec_top);
  int H = lh, W = lw;
  for (int lv = cfg.ae_n_ch_mult - 1; lv >= 0; --lv) {
    const DecLevel& L = h->dec_levels[lv];
    const int pl = 1, pr = L.kw == 3 ? 1 : 2, pt = L.kh == 3 ? 1 : 0;
    for (const ResW& r : L.blocks)
      step([&](const ViewF& i, const ViewF& o) { b.res_block_p(r, i, o, L.kh, L.kw, pl, pr, pt, 1e-6f); }, H, W, r.cout);
    if (L.has_up) {
      const int Ho = H * L.sh, Wo = W * L.sw;
      const int ukh = L.up.kh, ukw = L.up.kw;
      const int uh = (ukw - 1) / 2, upt = (ukh - 1) / 2;
      step([&](const ViewF& i, const ViewF& o) {
        Buf bu;
        View u = b.act_hl(B, Ho, Wo, L.ch, uh, uh, &bu);
        b.op([=](cudaStream_t s) { launch_upsample_bilinear_f32(i, u, s); });
        GemmEpilogue ep;
        ep.bias = L.up.bias;
        Builder::set_out_f(ep, o);
        b.gemm(u, taps_rect(ukh, ukw, uh, upt), L.up, ep);
        b.release(bu);
      }, Ho, Wo, L.ch);
      H = Ho; W = Wo;
    }
  }
  {
    Buf bg;
    View g = b.act_hl(B, H, W, h->dec_last, 1, 2, &bg);
    b.groupnorm_f(x, g, h->dec_norm_out, 1e-6f, true);
    b.release(bx);
    GemmB wb; wb.p = h->dec_conv_out.w; wb.n_alloc = h->dec_conv_out.n_alloc; wb.ld = h->dec_conv_out.k_alloc;
    wb.nseg = h->dec_conv_out.nseg;
    const float* bias = h->dec_conv_out.bias;
    const int N = h->dec_conv_out.cout;
    const ConvTaps taps = taps_rect(1, 4, 1, 0);
    if (cfg.ae_use_mask) {
      Buf bd;
      float* dec = b.raw<float>((size_t)B * N * H * W, &bd);
      b.op([=](cudaStream_t s) {
        GemmEpilogue ep; ep.bias = bias; ep.out_f32_nchw = dec;
        launch_conv_gemm(g, taps, wb, N, ep, s);
      }, PROF_GEMM, gemm_flops(g, taps.n, N) * wb.nseg);
      const int HW = H * W;
      b.op([=](cudaStream_t s) { launch_mask_select(dec, B, HW, P->out, s); });
      b.release(bd);
    } else {
      b.op([=](cudaStream_t s) {
        GemmEpilogue ep; ep.bias = bias; ep.out_f32_nchw = P->out;
        launch_conv_gemm(g, taps, wb, N, ep, s);
      }, PROF_GEMM, gemm_flops(g, taps.n, N) * wb.nseg);
    }
    b.release(bg);
  }
  *high = b.ap.high();
}

// Encoder.forward + quant_conv in the operand-split (fp32-class) mode: fp32 residual stream, hi/lo bf16 GEMM operands
void build_enc_plan_pass_p(lidm_handle* h, Plan* P, bool dry, size_t* high) {
  Builder b{h, P, ArenaPlanner(), dry};
  b.precise = true;
  const lidm_config& cfg = h->cfg;
  const int B = P->B;
  int H = h->img_h, W = h->img_w;
  const int kpad = h->enc_conv_in.k_alloc / h->enc_conv_in.nseg;
  Buf bcol, bx;
  bf16* col = b.raw<bf16>((size_t)B * H * W * 2 * kpad, &bcol);
  const int Cin = cfg.ae_in_channels;
  b.op([=](cudaStream_t s) { launch_im2col_nchw_f32_hl(P->x, B, Cin, H, W, 3, 3, 1, 1, col, kpad, s); });
  ViewF x = b.actf(B, H, W, cfg.ae_ch, &bx);
  {
    View a; a.p = col; a.B = B; a.H = H; a.W = W; a.C = kpad; a.ld = 2 * kpad; a.lo_off = kpad;
    GemmEpilogue ep;
    ep.bias = h->enc_conv_in.bias;
    Builder::set_out_f(ep, x);
    b.gemm(a, taps_1x1(), h->enc_conv_in, ep);
  }
  b.release(bcol);
  auto step = [&](std::function<void(const ViewF&, const ViewF&)> f, int Ho, int Wo, int C) {
    Buf bo;
    ViewF o = b.actf(B, Ho, Wo, C, &bo);
    f(x, o);
    b.release(bx);
    bx = bo; x = o;
  };
  for (const EncLevel& L : h->enc_levels) {
    for (const ResW& r : L.blocks)
      step([&](const ViewF& i, const ViewF& o) { b.res_block_p(r, i, o, 3, 3, 1, 1, 1, 1e-6f); }, H, W, r.cout);
    if (L.has_down) {
      const int Ho = H / L.sh, Wo = W / L.sw;
      step([&](const ViewF& i, const ViewF& o) { b.down_p(L.down, i, o, L.sh, L.sw, L.pl, L.pt); }, Ho, Wo, L.ch);
      H = Ho; W = Wo;
    }
  }
  step([&](const ViewF& i, const ViewF& o) { b.res_block_p(h->enc_mid1, i, o, 3, 3, 1, 1, 1, 1e-6f); }, H, W, h->enc_top);
  step([&](const ViewF& i, const ViewF& o) { b.dec_attn_block_p(h->enc_attn, i, o); }, H, W, h->enc_top);
  step([&](const ViewF& i, const ViewF& o) { b.res_block_p(h->enc_mid2, i, o, 3, 3, 1, 1, 1, 1e-6f); }, H, W, h->enc_top);
  {
    Buf bg;
    View g = b.act_hl(B, H, W, h->enc_top, 1, 1, &bg);
    b.groupnorm_f(x, g, h->enc_norm_out, 1e-6f, true);
    b.release(bx);
    const GemmB wb = Builder::gb(h->enc_conv_out);
    const float* bias = h->enc_conv_out.bias;
    const int N = h->enc_conv_out.cout;
    const ConvTaps taps = taps_rect(3, 3, 1, 1);
    b.op([=](cudaStream_t s) {
      GemmEpilogue ep; ep.bias = bias; ep.out_f32_nchw = P->out;
      launch_conv_gemm(g, taps, wb, N, ep, s);
    }, PROF_GEMM, gemm_flops(g, taps.n, N) * wb.nseg);
    b.release(bg);
  }
  *high = b.ap.high();
}

Plan* get_plan(lidm_handle* h, std::map<int64_t, std::unique_ptr<Plan>>& cache, int B,
               void (*pass)(lidm_handle*, Plan*, bool, size_t*), int ctx_len = 0) {
  const int64_t key = (int64_t)B | ((int64_t)ctx_len << 24);
  auto it = cache.find(key);
  if (it != cache.end()) return it->second.get();
  std::unique_ptr<Plan> P(new Plan());
  P->B = B;
  P->ctx_len = ctx_len;
  size_t high = 0;
  pass(h, P.get(), true, &high);
  P->arena_bytes = high;
  LIDM_CUDA_CHECK(cudaMalloc(reinterpret_cast<void**>(&P->arena), std::max<size_t>(high, 1024)));
  LIDM_CUDA_CHECK(cudaMalloc(reinterpret_cast<void**>(&P->gn_partials), (size_t)B * 32 * 2 * GN_MAX_CHUNKS * sizeof(float)));
  size_t high2 = 0;
  pass(h, P.get(), false, &high2);
  if (high2 != high) throw Error(LIDM_ERR_STATE, "internal: non-deterministic activation plan");
  // every plan owns its activation arena (GBs at large batches): keep a handful of shapes, drop the rest
  constexpr size_t kMaxPlans = 6;
  while (cache.size() >= kMaxPlans) {
    LIDM_CUDA_CHECK(cudaDeviceSynchronize());     // nothing may still be running out of the arena we are about to free
    cache.erase(cache.begin());
  }
  Plan* ret = P.get();
  cache[key] = std::move(P);
  return ret;
}

void run_plan(Plan* P, cudaStream_t s) {
  pdl_set_batch(P->B);
  if (!g_prof.on) {
    for (auto& o : P->ops) o.fn(s);
    return;
  }
  for (auto& o : P->ops) {
    cudaEvent_t a, b;
    LIDM_CUDA_CHECK(cudaEventCreate(&a));
    LIDM_CUDA_CHECK(cudaEventCreate(&b));
    LIDM_CUDA_CHECK(cudaEventRecord(a, s));
    const int64_t before = g_launch_count.load();
    o.fn(s);
    LIDM_CUDA_CHECK(cudaEventRecord(b, s));
    g_prof.ev.push_back(a); g_prof.ev.push_back(b); g_prof.cat.push_back(o.cat);
    g_prof.label.push_back(o.label); g_prof.op_flops.push_back(o.flops); g_prof.op_bytes.push_back(o.bytes);
    g_prof.flops[o.cat] += o.flops; g_prof.bytes[o.cat] += o.bytes;
    g_prof.launches[o.cat] += g_launch_count.load() - before;
  }
}

// Replays the plan as a CUDA graph (captured on a private stream the first time the same IO pointers come round): at
// small batches the ~160-400 launches of one U-Net evaluation cost more host time than device time.
void run_plan_graphed(lidm_handle* h, Plan* P, int slot, cudaStream_t s) {
  static const bool no_graph = getenv("LIDM_NO_GRAPH") != nullptr;
  if (g_prof.on || no_graph || !P->warmed) {
    run_plan(P, s);
    P->warmed = true;
    return;
  }
  Plan::GraphSlot& g = P->gslot[slot];
  const std::vector<const void*> key = P->io_key();
  if (g.exec == nullptr || g.key != key) {
    if (h->cap_stream == nullptr) LIDM_CUDA_CHECK(cudaStreamCreateWithFlags(&h->cap_stream, cudaStreamNonBlocking));
    LIDM_CUDA_CHECK(cudaStreamBeginCapture(h->cap_stream, cudaStreamCaptureModeThreadLocal));
    const int64_t before = g_launch_count.load();
    cudaGraph_t graph = nullptr;
    try {
      run_plan(P, h->cap_stream);
    } catch (...) {
      cudaStreamEndCapture(h->cap_stream, &graph);
      if (graph) cudaGraphDestroy(graph);
      throw;
    }
    LIDM_CUDA_CHECK(cudaStreamEndCapture(h->cap_stream, &graph));
    const int64_t kernels = g_launch_count.load() - before;
    g_launch_count.fetch_sub(kernels);          // captured, not launched
    cudaGraphExec_t exec = nullptr;
    cudaError_t e = cudaGraphInstantiate(&exec, graph, 0);
    cudaGraphDestroy(graph);
    LIDM_CUDA_CHECK(e);
    if (g.exec) cudaGraphExecDestroy(g.exec);
    g.exec = exec; g.key = key; g.kernels = kernels;
  }
  LIDM_CUDA_CHECK(cudaGraphLaunch(g.exec, s));
  g_launch_count.fetch_add(g.kernels);
}

// ------------------------------------------------------------------------------------------- finalize
// [W | I]: a 1x1 projection whose residual is added by the GEMM itself (GemmEpilogue::a2 + a2_diag): the residual tensor
// streams through the TMA ring as a second A operand against the identity block.
void add_identity(Packer& pk, ConvW& c) {
  if (pk.precise || c.w == nullptr || c.nseg != 1 || c.kh != 1 || c.kw != 1 || c.k_alloc != c.cin || c.cin % 64 != 0 ||
      c.cout % 256 != 0 || c.cout > 1024)
    return;
  lidm_handle* h = pk.h;
  c.k_alloc_id = c.cin + c.cout;
  c.w_id = dev_alloc<bf16>(h, (size_t)c.n_alloc * c.k_alloc_id);
  LIDM_CUDA_CHECK(cudaMemcpy2DAsync(c.w_id, (size_t)c.k_alloc_id * sizeof(bf16), c.w, (size_t)c.k_alloc * sizeof(bf16),
                                    (size_t)c.cin * sizeof(bf16), c.n_alloc, cudaMemcpyDeviceToDevice, pk.s));
  std::vector<uint16_t> eye((size_t)c.n_alloc * c.cout, 0);
  for (int i = 0; i < c.cout; ++i) eye[(size_t)i * c.cout + i] = pk.f16 ? 0x3c00 : 0x3f80;   // 1.0 as IEEE half / bf16
  LIDM_CUDA_CHECK(cudaMemcpy2DAsync(c.w_id + c.cin, (size_t)c.k_alloc_id * sizeof(bf16), eye.data(), (size_t)c.cout * sizeof(bf16),
                                    (size_t)c.cout * sizeof(bf16), c.n_alloc, cudaMemcpyHostToDevice, pk.s));
  LIDM_CUDA_CHECK(cudaStreamSynchronize(pk.s));
}

ResW pack_res(Packer& pk, const std::string& p, int cin, int cout, int kh, int kw, bool unet) {
  ResW r;
  r.cin = cin; r.cout = cout;
  if (unet) {
    r.n1 = pk.norm(p + ".in_layers.0", cin);
    r.c1 = pk.conv(p + ".in_layers.2", cout, cin, kh, kw);
    r.n2 = pk.norm(p + ".out_layers.0", cout);
    r.c2 = pk.conv(p + ".out_layers.3", cout, cout, kh, kw);
    if (cin != cout) { r.has_skip = true; r.skip = pk.conv(p + ".skip_connection", cout, cin, 1, 1); }
  } else {
    r.n1 = pk.norm(p + ".norm1", cin);
    r.c1 = pk.conv(p + ".conv1", cout, cin, kh, kw);
    r.n2 = pk.norm(p + ".norm2", cout);
    r.c2 = pk.conv(p + ".conv2", cout, cout, kh, kw);
    if (cin != cout) { r.has_skip = true; r.skip = pk.conv(p + ".nin_shortcut", cout, cin, 1, 1); }
  }
  if (r.has_skip && !pk.precise && cin % 64 == 0 && cout % 64 == 0) {
    // fold the 1x1 skip convolution into conv2: K = [kh*kw taps of h | x], bias = b2 + b_skip (ResBlock output =
    // skip(x) + conv2(h)); the skip tensor is never written or re-read
    lidm_handle* h = pk.h;
    const std::string c2name = p + (unet ? ".out_layers.3" : ".conv2");
    const DevTensor& t = find_raw(h, c2name + ".weight", pk.ema);
    ConvW f;
    f.cout = cout; f.cin = cout; f.kh = kh; f.kw = kw;
    f.n_alloc = r.c2.n_alloc; f.nseg = 1;
    f.k_alloc = kh * kw * cout + cin;
    f.w = dev_alloc<bf16>(h, (size_t)f.n_alloc * f.k_alloc);
    f.f16 = pk.f16;
    launch_pack_conv_weight(t.p, cout, cout, kh, kw, f.n_alloc, f.k_alloc, nullptr, nullptr, 1.f, 0, f.w, pk.s, pk.f16);
    LIDM_CUDA_CHECK(cudaMemcpy2DAsync(f.w + (size_t)kh * kw * cout, (size_t)f.k_alloc * sizeof(bf16), r.skip.w,
                                      (size_t)r.skip.k_alloc * sizeof(bf16), (size_t)cin * sizeof(bf16), r.skip.n_alloc,
                                      cudaMemcpyDeviceToDevice, pk.s));
    std::vector<float> b2(cout), bs(cout);
    LIDM_CUDA_CHECK(cudaMemcpyAsync(b2.data(), r.c2.bias, cout * sizeof(float), cudaMemcpyDeviceToHost, pk.s));
    LIDM_CUDA_CHECK(cudaMemcpyAsync(bs.data(), r.skip.bias, cout * sizeof(float), cudaMemcpyDeviceToHost, pk.s));
    LIDM_CUDA_CHECK(cudaStreamSynchronize(pk.s));
    for (int i = 0; i < cout; ++i) b2[i] += bs[i];
    f.bias = dev_alloc<float>(h, cout);
    LIDM_CUDA_CHECK(cudaMemcpy(f.bias, b2.data(), cout * sizeof(float), cudaMemcpyHostToDevice));
    r.c2s = f;
  }
  return r;
}

// qkv Conv1d of AttentionBlock: legacy channel order [head][q|k|v][ch] -> packed rows [q all heads | k all heads | v],
// q and k rows (and biases) pre-multiplied by ch^-1/4 (openaimodel.py:367-370).
AttnW pack_unet_attn(Packer& pk, const std::string& p, int ch, int heads) {
  lidm_handle* h = pk.h;
  AttnW a;
  a.ch = ch; a.heads = heads;
  const int d = ch / heads;
  if (d != 32) throw Error(LIDM_ERR_INVALID, "attention head dim must be 32 (num_head_channels)");
  a.n = pk.norm(p + ".norm", ch);
  const DevTensor& w = find_raw(h, p + ".qkv.weight", pk.ema);
  const DevTensor& bsrc = find_raw(h, p + ".qkv.bias", pk.ema);
  if (w.numel != (int64_t)3 * ch * ch || bsrc.numel != 3 * ch) throw Error(LIDM_ERR_STATE, "qkv weight size: " + p);
  std::vector<int> perm(3 * ch);
  for (int part = 0; part < 3; ++part)
    for (int hd = 0; hd < heads; ++hd)
      for (int c = 0; c < d; ++c) perm[part * ch + hd * d + c] = hd * 3 * d + part * d + c;
  int* perm_dev = dev_alloc<int>(h, perm.size());
  LIDM_CUDA_CHECK(cudaMemcpyAsync(perm_dev, perm.data(), perm.size() * sizeof(int), cudaMemcpyHostToDevice, pk.s));
  const float scale = kSqrtLog2e / std::sqrt(std::sqrt((float)d));
  ConvW c;
  c.cout = 3 * ch; c.cin = ch; c.kh = c.kw = 1;
  c.n_alloc = round_n_alloc(3 * ch);
  c.nseg = pk.precise ? 3 : 1;
  c.f16 = pk.f16;
  c.k_alloc = c.nseg * ch;
  c.w = dev_alloc<bf16>(h, (size_t)c.n_alloc * c.k_alloc);
  if (pk.precise)
    launch_pack_conv_weight_split(w.p, 3 * ch, ch, 1, 1, c.n_alloc, 3, perm_dev, scale, 2 * ch, c.w, pk.s);
  else
    launch_pack_conv_weight(w.p, 3 * ch, ch, 1, 1, c.n_alloc, c.k_alloc, perm_dev, nullptr, scale, 2 * ch, c.w, pk.s, pk.f16);
  std::vector<float> bh(3 * ch), bp(3 * ch);
  LIDM_CUDA_CHECK(cudaMemcpyAsync(bh.data(), bsrc.p, bh.size() * sizeof(float), cudaMemcpyDeviceToHost, pk.s));
  LIDM_CUDA_CHECK(cudaStreamSynchronize(pk.s));
  for (int i = 0; i < 3 * ch; ++i) bp[i] = bh[perm[i]] * (i < 2 * ch ? scale : 1.0f);
  c.bias = dev_alloc<float>(h, bp.size());
  LIDM_CUDA_CHECK(cudaMemcpy(c.bias, bp.data(), bp.size() * sizeof(float), cudaMemcpyHostToDevice));
  a.qkv = c;
  a.proj = pk.conv(p + ".proj_out", ch, ch, 1, 1, 0, 2);   // its input (attention output) is exact bf16
  add_identity(pk, a.proj);
  return a;
}

// Decoder AttnBlock: separate q/k/v 1x1 convs -> one packed [q * C^-1/2 | k | v] matrix
AttnW pack_dec_attn(Packer& pk, const std::string& p, int ch) {
  lidm_handle* h = pk.h;
  AttnW a;
  a.ch = ch; a.heads = 1;
  a.n = pk.norm(p + ".norm", ch);
  ConvW c;
  c.cout = 3 * ch; c.cin = ch; c.kh = c.kw = 1;
  c.n_alloc = round_n_alloc(3 * ch);
  c.nseg = pk.precise ? 3 : 1;
  c.f16 = pk.f16;
  c.k_alloc = c.nseg * ch;
  if (c.n_alloc != 3 * ch) throw Error(LIDM_ERR_INVALID, "decoder attention channels must be a multiple of 128");
  c.w = dev_alloc<bf16>(h, (size_t)c.n_alloc * c.k_alloc);
  c.bias = dev_alloc<float>(h, 3 * ch);
  const char* names[3] = {".q", ".k", ".v"};
  const float scale = 1.0f / std::sqrt((float)ch);
  std::vector<float> bh(ch);
  for (int i = 0; i < 3; ++i) {
    const DevTensor& w = find_raw(h, p + names[i] + ".weight", pk.ema);
    const DevTensor& bsrc = find_raw(h, p + names[i] + ".bias", pk.ema);
    if (w.numel != (int64_t)ch * ch || bsrc.numel != ch) throw Error(LIDM_ERR_STATE, "decoder attention weight size");
    if (pk.precise)
      launch_pack_conv_weight_split(w.p, ch, ch, 1, 1, ch, 3, nullptr, i == 0 ? scale : 1.f, i == 0 ? ch : 0,
                                    c.w + (size_t)i * ch * c.k_alloc, pk.s);
    else
      launch_pack_conv_weight(w.p, ch, ch, 1, 1, ch, ch, nullptr, nullptr, i == 0 ? scale : 1.f, i == 0 ? ch : 0,
                              c.w + (size_t)i * ch * ch, pk.s, pk.f16);
    LIDM_CUDA_CHECK(cudaMemcpy(bh.data(), bsrc.p, ch * sizeof(float), cudaMemcpyDeviceToHost));
    if (i == 0) for (float& v : bh) v *= scale;
    LIDM_CUDA_CHECK(cudaMemcpy(c.bias + (size_t)i * ch, bh.data(), ch * sizeof(float), cudaMemcpyHostToDevice));
  }
  a.qkv = c;
  a.proj = pk.conv(p + ".proj_out", ch, ch, 1, 1, 0, 2);
  return a;
}

// The 3x3 conv that follows a nearest x2 upsample, folded into four 2x2 convs (one per output parity): tap (a, b) of
// parity (py, px) is the sum of the 3x3 taps that land on the same low-resolution pixel.  Summed in fp32 on the host,
// then packed like any other conv weight.
void fold_upsample_conv(Packer& pk, const std::string& prefix, int ch, ConvW* out4) {
  lidm_handle* h = pk.h;
  const DevTensor& w = find_raw(h, prefix + ".weight", pk.ema);
  if (w.numel != (int64_t)ch * ch * 9) throw Error(LIDM_ERR_STATE, "weight '" + prefix + ".weight' has unexpected size");
  std::vector<float> W3((size_t)w.numel), W2((size_t)ch * ch * 4);
  LIDM_CUDA_CHECK(cudaMemcpy(W3.data(), w.p, W3.size() * sizeof(float), cudaMemcpyDeviceToHost));
  // rows of the 3x3 kernel feeding 2x2 row a:  parity 0: a=0 -> {0}, a=1 -> {1,2};  parity 1: a=0 -> {0,1}, a=1 -> {2}
  auto span = [](int parity, int a, int* lo, int* hi) {
    if (parity == 0) { *lo = a == 0 ? 0 : 1; *hi = a == 0 ? 0 : 2; }
    else { *lo = a == 0 ? 0 : 2; *hi = a == 0 ? 1 : 2; }
  };
  for (int py = 0; py < 2; ++py)
    for (int px = 0; px < 2; ++px) {
      for (size_t oc = 0; oc < (size_t)ch * ch; ++oc) {
        const float* s3 = &W3[oc * 9];
        for (int a = 0; a < 2; ++a)
          for (int b = 0; b < 2; ++b) {
            int y0, y1, x0, x1;
            span(py, a, &y0, &y1);
            span(px, b, &x0, &x1);
            float acc = 0.f;
            for (int ky = y0; ky <= y1; ++ky)
              for (int kx = x0; kx <= x1; ++kx) acc += s3[ky * 3 + kx];
            W2[oc * 4 + a * 2 + b] = acc;
          }
      }
      const std::string name = prefix + "_fold" + std::to_string(py * 2 + px);
      DevTensor t;
      t.numel = (int64_t)W2.size();
      t.shape = {ch, ch, 2, 2};
      LIDM_CUDA_CHECK(cudaMalloc(reinterpret_cast<void**>(&t.p), W2.size() * sizeof(float)));
      LIDM_CUDA_CHECK(cudaMemcpy(t.p, W2.data(), W2.size() * sizeof(float), cudaMemcpyHostToDevice));
      auto it = h->raw.find(name + ".weight");
      if (it != h->raw.end()) { cudaFree(it->second.p); h->raw.erase(it); }
      h->raw.emplace(name + ".weight", std::move(t));
      // bias: shared with the 3x3 conv (Packer::conv copies it per call)
      const DevTensor& b3 = find_raw(h, prefix + ".bias", pk.ema);
      DevTensor tb;
      tb.numel = b3.numel; tb.shape = b3.shape;
      LIDM_CUDA_CHECK(cudaMalloc(reinterpret_cast<void**>(&tb.p), b3.numel * sizeof(float)));
      LIDM_CUDA_CHECK(cudaMemcpy(tb.p, b3.p, b3.numel * sizeof(float), cudaMemcpyDeviceToDevice));
      auto itb = h->raw.find(name + ".bias");
      if (itb != h->raw.end()) { cudaFree(itb->second.p); h->raw.erase(itb); }
      h->raw.emplace(name + ".bias", std::move(tb));
      Packer pk2{h, false, pk.s};
      pk2.f16 = pk.f16;
      out4[py * 2 + px] = pk2.conv(name, ch, ch, 2, 2);
    }
}

bool in_list(const int32_t* v, int n, int x) {
  for (int i = 0; i < n; ++i) if (v[i] == x) return true;
  return false;
}

// ObjectAwareCrossAttention weights (object_cross_unet.py:430-446, shipped options: norm_first False, positional scale 1)
OacaW pack_oaca(Packer& pk, const std::string& p, int ch, int rows) {
  lidm_handle* h = pk.h;
  const int E = h->cfg.encoder_channels;
  OacaW a;
  a.ch = ch; a.heads = ch / 64; a.rows = rows;
  a.n_qkv = pk.norm(p + ".norm_for_qkv", ch);
  a.n_img_pos = pk.norm(p + ".norm_for_image_patch_positional_embedding", ch);
  a.n_lay_pos = pk.norm(p + ".norm_for_layout_positional_embedding", ch);
  a.n_cls = pk.norm(p + ".norm_for_obj_class_embedding", E);
  if (h->raw.count(p + ".norm_for_obj_embedding.weight"))
    throw Error(LIDM_ERR_INVALID, "ObjectAwareCrossAttention: norm_for_obj_embedding / norm_first are not supported");
  // q and k rows (and biases) carry the (2 * 64)^-1/4 scale of both score factors (object_cross_unet.py:533-536)
  const float scale = 1.0f / std::sqrt(std::sqrt(128.0f));
  const DevTensor& w = find_raw(h, p + ".qkv_projector.weight", pk.ema);
  const DevTensor& bsrc = find_raw(h, p + ".qkv_projector.bias", pk.ema);
  if (w.numel != (int64_t)3 * ch * ch || bsrc.numel != 3 * ch) throw Error(LIDM_ERR_STATE, "qkv_projector weight size: " + p);
  ConvW c;
  c.cout = 3 * ch; c.cin = ch; c.kh = c.kw = 1;
  c.n_alloc = round_n_alloc(3 * ch);
  c.f16 = pk.f16;
  c.k_alloc = ch;
  c.w = dev_alloc<bf16>(h, (size_t)c.n_alloc * c.k_alloc);
  launch_pack_conv_weight(w.p, 3 * ch, ch, 1, 1, c.n_alloc, c.k_alloc, nullptr, nullptr, scale, 2 * ch, c.w, pk.s, pk.f16);
  std::vector<float> bh(3 * ch);
  LIDM_CUDA_CHECK(cudaMemcpyAsync(bh.data(), bsrc.p, bh.size() * sizeof(float), cudaMemcpyDeviceToHost, pk.s));
  LIDM_CUDA_CHECK(cudaStreamSynchronize(pk.s));
  for (int i = 0; i < 2 * ch; ++i) bh[i] *= scale;
  c.bias = dev_alloc<float>(h, bh.size());
  LIDM_CUDA_CHECK(cudaMemcpy(c.bias, bh.data(), bh.size() * sizeof(float), cudaMemcpyHostToDevice));
  a.qkv = c;
  a.proj = pk.conv(p + ".proj_out", ch, ch, 1, 1);
  a.w_pos = pk.f32(p + ".layout_position_embedding_projector.weight", (int64_t)ch * E);
  a.b_pos = pk.f32(p + ".layout_position_embedding_projector.bias", ch);
  a.w_content = pk.f32(p + ".layout_content_embedding_projector.weight", (int64_t)2 * ch * E);
  a.b_content = pk.f32(p + ".layout_content_embedding_projector.bias", 2 * ch);
  return a;
}

// LayoutDiffusionUNetModel.__init__ (object_cross_unet.py:742-912, resblock_updown = True, use_scale_shift_norm = True)
void finalize_unet_layout(lidm_handle* h, Packer& pk) {
  const lidm_config& cfg = h->cfg;
  const std::string U = "model.diffusion_model.";
  const int mc = cfg.model_channels;
  const int nab = cfg.num_attention_blocks > 0 ? cfg.num_attention_blocks : 1;
  LIDM_REQUIRE(!pk.precise, "the layout U-Net runs in the bf16 / fp16 modes");
  LIDM_REQUIRE(cfg.num_head_channels == 64 && cfg.encoder_channels > 0 && cfg.encoder_channels % 32 == 0,
               "layout U-Net: num_head_channels must be 64 and encoder_channels a positive multiple of 32");
  {
    Layer L; L.kind = Layer::CONV; L.cin = cfg.in_channels; L.cout = mc;
    L.c = pk.conv(U + "input_blocks.0.0", mc, cfg.in_channels, 3, 3, (9 * cfg.in_channels + 63) / 64 * 64);
    h->in_blocks.push_back({L});
    h->in_chans.push_back(mc);
  }
  auto make_res = [&](const std::string& p, int cin, int cout, int updown) {
    Layer L; L.kind = Layer::RES; L.cin = cin; L.cout = cout;
    L.r = pack_res(pk, p, cin, cout, 3, 3, true);
    L.r.film = true; L.r.updown = updown;
    return L;
  };
  auto make_attn = [&](const std::string& p, int c, int ds) {
    Layer L; L.kind = Layer::OACA; L.cin = L.cout = c;
    LIDM_REQUIRE(cfg.latent_h % ds == 0, "attention_ds must divide the latent height");
    L.oa = pack_oaca(pk, p, c, cfg.latent_h / ds);
    return L;
  };
  int ch = mc, ds = 1;
  for (int level = 0; level < cfg.n_channel_mult; ++level) {
    for (int r = 0; r < cfg.num_res_blocks; ++r) {
      const std::string p = U + "input_blocks." + std::to_string(h->in_blocks.size());
      std::vector<Layer> layers;
      layers.push_back(make_res(p + ".0", ch, cfg.channel_mult[level] * mc, 0));
      ch = cfg.channel_mult[level] * mc;
      if (in_list(cfg.attention_resolutions, cfg.n_attention_resolutions, ds))
        for (int a = 0; a < nab; ++a) layers.push_back(make_attn(p + "." + std::to_string(1 + a), ch, ds));
      h->in_blocks.push_back(layers);
      h->in_chans.push_back(ch);
    }
    if (level != cfg.n_channel_mult - 1) {
      const std::string p = U + "input_blocks." + std::to_string(h->in_blocks.size());
      h->in_blocks.push_back({make_res(p + ".0", ch, ch, 2)});
      h->in_chans.push_back(ch);
      ds *= 2;
    }
  }
  h->mid_block.push_back(make_res(U + "middle_block.0", ch, ch, 0));
  h->mid_block.push_back(make_attn(U + "middle_block.1", ch, ds));
  h->mid_block.push_back(make_res(U + "middle_block.2", ch, ch, 0));
  std::vector<int> chans = h->in_chans;
  for (int level = cfg.n_channel_mult - 1; level >= 0; --level) {
    for (int i = 0; i <= cfg.num_res_blocks; ++i) {
      const int ich = chans.back();
      chans.pop_back();
      const std::string p = U + "output_blocks." + std::to_string(h->out_blocks.size());
      std::vector<Layer> layers;
      layers.push_back(make_res(p + ".0", ch + ich, mc * cfg.channel_mult[level], 0));
      ch = mc * cfg.channel_mult[level];
      int j = 1;
      if (in_list(cfg.attention_resolutions, cfg.n_attention_resolutions, ds))
        for (int a = 0; a < nab; ++a) layers.push_back(make_attn(p + "." + std::to_string(j++), ch, ds));
      if (level && i == cfg.num_res_blocks) {
        layers.push_back(make_res(p + "." + std::to_string(j++), ch, ch, 1));
        ds /= 2;
      }
      h->out_blocks.push_back(layers);
    }
  }
  h->out_norm = pk.norm(U + "out.0", ch);
  h->out_conv = pk.conv(U + "out.2", cfg.out_channels, mc, 3, 3);
  if (ch != mc) throw Error(LIDM_ERR_INVALID, "U-Net must end at model_channels");
}

// EfficientUNet.__init__ (lidm/modules/unets/efficient_unet.py:188-260).  Every residual path ends in "* 1/sqrt 2": the
// scale is folded into the last conv's (and the skip conv's) weights and bias; an identity skip enters the GEMM epilogue as
// res_scale * x.
void finalize_unet_efficient(lidm_handle* h, Packer& pk, std::vector<std::pair<std::string, ResW*>>& emb) {
  const lidm_config& cfg = h->cfg;
  const std::string U = "model.diffusion_model.";
  LIDM_REQUIRE(!pk.precise, "the R2DM U-Net runs in the bf16 / fp16 modes");
  LIDM_REQUIRE(cfg.n_channel_mult == 4, "EfficientUNet has four resolution levels");
  LIDM_REQUIRE(cfg.eff_gn_groups >= 1 && cfg.eff_attn_heads >= 1 && cfg.eff_gn_eps > 0.f, "EfficientUNet: GroupNorm / attention options");
  const int mc = cfg.model_channels, H = cfg.latent_h, W = cfg.latent_w;
  const float S = 0.70710678118654752f;
  int C[5] = {mc, 0, 0, 0, 0};
  for (int i = 0; i < 4; ++i) {
    C[i + 1] = mc * cfg.channel_mult[i];
    LIDM_REQUIRE(C[i + 1] % 64 == 0 && (C[i + 1] / cfg.eff_gn_groups) % 8 == 0 && cfg.eff_res_blocks[i] >= 1,
                 "EfficientUNet: channels must be multiples of 64 with >= 8 channels per GroupNorm group");
  }
  h->ones_c = dev_alloc<float>(h, 4096);
  h->zeros_c = dev_alloc<float>(h, 4096);
  {
    std::vector<float> o(4096, 1.f);
    LIDM_CUDA_CHECK(cudaMemcpy(h->ones_c, o.data(), o.size() * sizeof(float), cudaMemcpyHostToDevice));
    LIDM_CUDA_CHECK(cudaMemset(h->zeros_c, 0, 4096 * sizeof(float)));
  }
  // Fourier features of the polar coordinates (encoding.py:93-105, 133-163), float32 like the reference's buffers
  {
    const int Lh = (int)std::ceil(std::log2((double)H)), Lw = (int)std::ceil(std::log2((double)W)), nf = Lh + Lw;
    h->eff_cenc_ch = 2 * nf;
    std::vector<float> ce((size_t)2 * nf * H * W);
    const float d2r = (float)(M_PI / 180.0);
    for (int y = 0; y < H; ++y) {
      const float el = ((1.f - (float)y / (float)H) * 40.f + (-30.f)) * d2r;
      for (int x = 0; x < W; ++x) {
        const float az = ((1.f - (float)x / (float)W) * 360.f + (-180.f)) * d2r;
        for (int k = 0; k < nf; ++k) {
          const float a = k < Lh ? el * std::exp2((float)k) : az * std::exp2((float)(k - Lh));
          ce[((size_t)k * H + y) * W + x] = std::sin(a);
          ce[((size_t)(nf + k) * H + y) * W + x] = std::cos(a);
        }
      }
    }
    h->eff_cenc = dev_alloc<float>(h, ce.size());
    LIDM_CUDA_CHECK(cudaMemcpy(h->eff_cenc, ce.data(), ce.size() * sizeof(float), cudaMemcpyHostToDevice));
  }
  const int cin0 = cfg.in_channels + h->eff_cenc_ch;
  h->eff_in_conv = pk.conv(U + "in_conv", C[0], cin0, 3, 3, (9 * cin0 + 63) / 64 * 64);
  {
    // the coordinate channels are constants: their share of in_conv (+ bias) becomes a per-pixel map, the image channels keep
    // their fp32 weights for the per-step CUDA-core convolution (launch_eff_in_conv)
    const DevTensor& wraw = find_raw(h, U + "in_conv.weight", pk.ema);
    const DevTensor& braw = find_raw(h, U + "in_conv.bias", pk.ema);
    LIDM_REQUIRE(wraw.numel == (int64_t)C[0] * cin0 * 9 && braw.numel == C[0], "in_conv weight size");
    const int Cx = cfg.in_channels;
    h->eff_in_map = dev_alloc<float>(h, (size_t)H * W * C[0]);
    h->eff_in_w = dev_alloc<float>(h, (size_t)C[0] * Cx * 9);
    launch_eff_in_map(wraw.p, braw.p, h->eff_cenc, Cx, h->eff_cenc_ch, H, W, C[0], h->eff_in_map, pk.s);
    LIDM_CUDA_CHECK(cudaMemcpy2DAsync(h->eff_in_w, (size_t)Cx * 9 * sizeof(float), wraw.p, (size_t)cin0 * 9 * sizeof(float),
                                      (size_t)Cx * 9 * sizeof(float), C[0], cudaMemcpyDeviceToDevice, pk.s));
    LIDM_CUDA_CHECK(cudaStreamSynchronize(pk.s));
  }
  auto pack_res_eff = [&](const std::string& p, int cin, int cout, ResW& r) {
    r.cin = cin; r.cout = cout; r.film = true;
    r.n1 = pk.norm(p + ".norm1", cin);
    r.c1 = pk.conv(p + ".conv1", cout, cin, 3, 3);
    r.n2.C = cout; r.n2.gamma = h->ones_c; r.n2.beta = h->zeros_c;
    r.c2 = pk.conv(p + ".conv2", cout, cout, 3, 3, 0, 3, S);
    r.has_skip = cin != cout;
    if (r.has_skip) {
      r.skip = pk.conv(p + ".skip", cout, cin, 1, 1, 0, 3, S);
      // conv2(h) + skip(x) as one GEMM (both already carry the 1/sqrt 2)
      const DevTensor& t = find_raw(h, p + ".conv2.weight", pk.ema);
      ConvW f;
      f.cout = cout; f.cin = cout; f.kh = f.kw = 3; f.f16 = pk.f16;
      f.n_alloc = r.c2.n_alloc; f.k_alloc = 9 * cout + cin;
      f.w = dev_alloc<bf16>(h, (size_t)f.n_alloc * f.k_alloc);
      launch_pack_conv_weight(t.p, cout, cout, 3, 3, f.n_alloc, f.k_alloc, nullptr, nullptr, S, cout, f.w, pk.s, pk.f16);
      LIDM_CUDA_CHECK(cudaMemcpy2DAsync(f.w + (size_t)9 * cout, (size_t)f.k_alloc * sizeof(bf16), r.skip.w,
                                        (size_t)r.skip.k_alloc * sizeof(bf16), (size_t)cin * sizeof(bf16), r.skip.n_alloc,
                                        cudaMemcpyDeviceToDevice, pk.s));
      std::vector<float> b2(cout), bs(cout);
      LIDM_CUDA_CHECK(cudaMemcpyAsync(b2.data(), r.c2.bias, cout * sizeof(float), cudaMemcpyDeviceToHost, pk.s));
      LIDM_CUDA_CHECK(cudaMemcpyAsync(bs.data(), r.skip.bias, cout * sizeof(float), cudaMemcpyDeviceToHost, pk.s));
      LIDM_CUDA_CHECK(cudaStreamSynchronize(pk.s));
      for (int i = 0; i < cout; ++i) b2[i] += bs[i];
      f.bias = dev_alloc<float>(h, cout);
      LIDM_CUDA_CHECK(cudaMemcpy(f.bias, b2.data(), cout * sizeof(float), cudaMemcpyHostToDevice));
      r.c2s = f;
    }
  };
  auto pack_block = [&](lidm_handle::EffBlock& blk, const std::string& p, int cin, int cout, int nres, bool down, bool up, bool attn) {
    blk.cin = cin; blk.ch = cout; blk.has_down = down; blk.has_up = up; blk.has_attn = attn;
    if (down) blk.down = pk.conv(p + ".downsample.0", cout, cin, 3, 3);
    blk.res.resize(nres);
    for (int i = 0; i < nres; ++i) {
      const std::string rp = p + ".residual_blocks." + std::to_string(i);
      pack_res_eff(rp, (i != 0 || down) ? cout : cin, cout, blk.res[i]);
      emb.emplace_back(rp + ".norm2.proj.1", &blk.res[i]);
    }
    if (attn) {
      const std::string ap = p + ".self_attn_block";
      const int heads = cfg.eff_attn_heads, d = cout / heads;
      LIDM_REQUIRE(cout % heads == 0 && (d == 32 || d == 64), "EfficientUNet attention: head width 32 or 64");
      blk.attn_heads = heads;
      blk.attn_norm = pk.norm(ap + ".norm", cout);
      // nn.MultiheadAttention: in_proj rows [q | k | v], softmax(q k^T / sqrt d): d^-1/4 on q and k rows (head width 32, flash
      // kernel) or d^-1/2 on the q rows (head width 64, GEMM path)
      const DevTensor& w = find_raw(h, ap + ".attn.in_proj_weight", pk.ema);
      const DevTensor& bsrc = find_raw(h, ap + ".attn.in_proj_bias", pk.ema);
      if (w.numel != (int64_t)3 * cout * cout || bsrc.numel != 3 * cout) throw Error(LIDM_ERR_STATE, "in_proj weight size: " + ap);
      const float sc = d == 32 ? kSqrtLog2e / std::sqrt(std::sqrt((float)d)) : 1.0f / std::sqrt((float)d);
      const int nsc = d == 32 ? 2 * cout : cout;
      ConvW c;
      c.cout = 3 * cout; c.cin = cout; c.kh = c.kw = 1; c.f16 = pk.f16;
      c.n_alloc = round_n_alloc(3 * cout); c.k_alloc = cout;
      c.w = dev_alloc<bf16>(h, (size_t)c.n_alloc * c.k_alloc);
      launch_pack_conv_weight(w.p, 3 * cout, cout, 1, 1, c.n_alloc, c.k_alloc, nullptr, nullptr, sc, nsc, c.w, pk.s, pk.f16);
      std::vector<float> bh(3 * cout);
      LIDM_CUDA_CHECK(cudaMemcpyAsync(bh.data(), bsrc.p, bh.size() * sizeof(float), cudaMemcpyDeviceToHost, pk.s));
      LIDM_CUDA_CHECK(cudaStreamSynchronize(pk.s));
      for (int i = 0; i < nsc; ++i) bh[i] *= sc;
      c.bias = dev_alloc<float>(h, bh.size());
      LIDM_CUDA_CHECK(cudaMemcpy(c.bias, bh.data(), bh.size() * sizeof(float), cudaMemcpyHostToDevice));
      blk.attn_qkv = c;
      // out_proj is an nn.Linear (2-D weight): same memory layout as a 1x1 conv
      blk.attn_proj = pk.conv(ap + ".attn.out_proj", cout, cout, 1, 1, 0, 3, S);
    }
    if (up) blk.up = pk.conv(p + ".upsample.1", cout, cout, 3, 3);
  };
  const int32_t* N = cfg.eff_res_blocks;
  pack_block(h->eff_d[0], U + "d_block1", C[0], C[1], N[0], false, false, false);
  pack_block(h->eff_d[1], U + "d_block2", C[1], C[2], N[1], true, false, false);
  pack_block(h->eff_d[2], U + "d_block3", C[2], C[3], N[2], true, false, false);
  pack_block(h->eff_d[3], U + "d_block4", C[3], C[4], N[3], true, false, true);
  pack_block(h->eff_u[3], U + "u_block4", C[4], C[3], N[3], false, true, true);
  pack_block(h->eff_u[2], U + "u_block3", 2 * C[3], C[2], N[2], false, true, false);
  pack_block(h->eff_u[1], U + "u_block2", 2 * C[2], C[1], N[1], false, true, false);
  pack_block(h->eff_u[0], U + "u_block1", 2 * C[1], C[0], N[0], false, false, false);
  h->out_conv = pk.conv(U + "out_conv", cfg.out_channels, C[0], 3, 3);
}

void finalize(lidm_handle* h, bool use_ema) {
  const lidm_config& cfg = h->cfg;
  Packer pk{h, use_ema};
  pk.precise = h->unet_prec == LIDM_PREC_BF16X3;
  pk.f16 = h->unet_prec == LIDM_PREC_FP16;
  const std::string U = "model.diffusion_model.";
  const int mc = cfg.model_channels, ted = mc * 4;
  h->latent_channels = cfg.latent_channels > 0 ? cfg.latent_channels : cfg.in_channels;
  LIDM_REQUIRE(h->latent_channels <= cfg.in_channels && h->latent_channels == cfg.out_channels,
               "latent channels must equal out_channels and not exceed in_channels");
  h->ted = ted;
  const bool eff_names = cfg.unet_type == 2;      // EfficientUNet: time_embedding = [sinusoidal, Linear, SiLU, Linear]
  h->te_w0 = pk.f32(U + (eff_names ? "time_embedding.1.weight" : "time_embed.0.weight"), (int64_t)ted * mc);
  h->te_b0 = pk.f32(U + (eff_names ? "time_embedding.1.bias" : "time_embed.0.bias"), ted);
  h->te_w2 = pk.f32(U + (eff_names ? "time_embedding.3.weight" : "time_embed.2.weight"), (int64_t)ted * ted);
  h->te_b2 = pk.f32(U + (eff_names ? "time_embedding.3.bias" : "time_embed.2.bias"), ted);

  std::vector<std::pair<std::string, ResW*>> emb_layers;
  h->in_blocks.clear(); h->out_blocks.clear(); h->mid_block.clear(); h->in_chans.clear();
  h->is_layout = cfg.unet_type == 1;
  h->is_eff = cfg.unet_type == 2;
  h->has_st = false;
  h->ctx_n = 0;
  std::vector<std::pair<std::string, ResW*>> eff_emb;
  if (h->is_layout) finalize_unet_layout(h, pk);
  else if (h->is_eff) finalize_unet_efficient(h, pk, eff_emb);
  else {
  // ---- U-Net topology, exactly as UNetModel.__init__ walks it (openaimodel.py:516-687)
  {
    Layer L; L.kind = Layer::CONV; L.cin = cfg.in_channels; L.cout = mc;
    const int kpad = (9 * cfg.in_channels + 63) / 64 * 64;
    L.c = pk.conv(U + "input_blocks.0.0", mc, cfg.in_channels, 3, 3, kpad);
    h->in_blocks.push_back({L});
    h->in_chans.push_back(mc);
  }
  int ch = mc, ds = 1;
  auto make_res = [&](const std::string& p, int cin, int cout) {
    Layer L; L.kind = Layer::RES; L.cin = cin; L.cout = cout;
    L.r = pack_res(pk, p, cin, cout, 3, 3, true);
    return L;
  };
  std::vector<std::pair<std::string, int>> st_blocks;   // (transformer block prefix, channels) in packing order
  auto make_attn = [&](const std::string& p, int c) {
    Layer L; L.cin = L.cout = c;
    if (!cfg.use_spatial_transformer) {
      L.kind = Layer::ATTN;
      L.a = pack_unet_attn(pk, p, c, c / cfg.num_head_channels);
      return L;
    }
    // SpatialTransformer(ch, num_heads = ch / num_head_channels, dim_head = num_head_channels) (openaimodel.py:546-561)
    L.kind = Layer::ST;
    STW& t = L.st;
    t.ch = c; t.heads = c / cfg.num_head_channels;
    const float d = (float)cfg.num_head_channels;
    const float s4 = kSqrtLog2e / std::sqrt(std::sqrt(d)), s2 = kLog2e / std::sqrt(d);
    t.n = pk.norm(p + ".norm", c);
    t.proj_in = pk.conv(p + ".proj_in", c, c, 1, 1);
    t.proj_out = pk.conv(p + ".proj_out", c, c, 1, 1);
    add_identity(pk, t.proj_out);
    for (int k = 0; k < cfg.transformer_depth; ++k) {
      const std::string bp = p + ".transformer_blocks." + std::to_string(k);
      STBlockW w;
      w.n1 = pk.norm(bp + ".norm1", c); w.n2 = pk.norm(bp + ".norm2", c); w.n3 = pk.norm(bp + ".norm3", c);
      // softmax(q k^T * d^-1/2) (attention.py:158,183): d^-1/4 folded into the q and k rows of the self-attention,
      // d^-1/2 into the q rows of the cross-attention (its k rows live in the shared context matrix)
      w.qkv1 = pack_stacked_linear(pk, {{bp + ".attn1.to_q.weight", s4}, {bp + ".attn1.to_k.weight", s4},
                                        {bp + ".attn1.to_v.weight", 1.f}}, c, c, "");
      w.out1 = pack_stacked_linear(pk, {{bp + ".attn1.to_out.0.weight", 1.f}}, c, c, bp + ".attn1.to_out.0.bias");
      w.q2 = pack_stacked_linear(pk, {{bp + ".attn2.to_q.weight", s2}}, c, c, "");
      w.out2 = pack_stacked_linear(pk, {{bp + ".attn2.to_out.0.weight", 1.f}}, c, c, bp + ".attn2.to_out.0.bias");
      w.ff0 = pack_geglu_linear(pk, bp + ".ff.net.0.proj", 8 * c, c);
      w.ff2 = pack_stacked_linear(pk, {{bp + ".ff.net.2.weight", 1.f}}, c, 4 * c, bp + ".ff.net.2.bias");
      add_identity(pk, w.out1); add_identity(pk, w.out2);   // (ff2: K = 4C is long enough to hide the residual read)
      w.kv_col = h->ctx_n;
      h->ctx_n += 2 * c;
      st_blocks.emplace_back(bp, c);
      t.blocks.push_back(w);
    }
    return L;
  };
  h->has_st = cfg.use_spatial_transformer != 0;
  h->ctx_n = 0;
  if (h->has_st) {
    LIDM_REQUIRE(cfg.context_dim > 0 && cfg.context_dim % 64 == 0, "context_dim must be a positive multiple of 64");
    LIDM_REQUIRE(cfg.transformer_depth >= 1, "transformer_depth");
    LIDM_REQUIRE(!pk.precise && !pk.f16, "SpatialTransformer U-Nets run in the bf16 mode only");
  }
  for (int level = 0; level < cfg.n_channel_mult; ++level) {
    const int mult = cfg.channel_mult[level];
    for (int r = 0; r < cfg.num_res_blocks; ++r) {
      const std::string p = U + "input_blocks." + std::to_string(h->in_blocks.size());
      std::vector<Layer> layers;
      layers.push_back(make_res(p + ".0", ch, mult * mc));
      ch = mult * mc;
      if (in_list(cfg.attention_resolutions, cfg.n_attention_resolutions, ds)) layers.push_back(make_attn(p + ".1", ch));
      h->in_blocks.push_back(layers);
      h->in_chans.push_back(ch);
    }
    if (level != cfg.n_channel_mult - 1) {
      const std::string p = U + "input_blocks." + std::to_string(h->in_blocks.size());
      Layer L; L.kind = Layer::DOWN; L.cin = L.cout = ch;
      L.c = pk.conv(p + ".0.op", ch, ch, 3, 3);
      h->in_blocks.push_back({L});
      h->in_chans.push_back(ch);
      ds *= 2;
    }
  }
  h->mid_block.push_back(make_res(U + "middle_block.0", ch, ch));
  h->mid_block.push_back(make_attn(U + "middle_block.1", ch));
  h->mid_block.push_back(make_res(U + "middle_block.2", ch, ch));
  {
    std::vector<int> chans = h->in_chans;
    for (int level = cfg.n_channel_mult - 1; level >= 0; --level) {
      const int mult = cfg.channel_mult[level];
      for (int i = 0; i <= cfg.num_res_blocks; ++i) {
        const int ich = chans.back();
        chans.pop_back();
        const std::string p = U + "output_blocks." + std::to_string(h->out_blocks.size());
        std::vector<Layer> layers;
        layers.push_back(make_res(p + ".0", ch + ich, mc * mult));
        ch = mc * mult;
        int j = 1;
        if (in_list(cfg.attention_resolutions, cfg.n_attention_resolutions, ds))
          layers.push_back(make_attn(p + "." + std::to_string(j++), ch));
        if (level && i == cfg.num_res_blocks) {
          Layer L; L.kind = Layer::UP; L.cin = L.cout = ch;
          const std::string cp = p + "." + std::to_string(j++) + ".conv";
          L.c = pk.conv(cp, ch, ch, 3, 3);
          static const bool no_fold = getenv("LIDM_NO_UPSAMPLE_FOLD") != nullptr;
          if (!pk.precise && !no_fold) fold_upsample_conv(pk, cp, ch, L.cpar);
          layers.push_back(L);
          ds /= 2;
        }
        h->out_blocks.push_back(layers);
      }
    }
  }
  if (h->has_st) {
    // one K-major matrix holding every transformer block's [to_k ; to_v] rows: the context is projected once per U-Net
    // evaluation by a single GEMM
    const int D = cfg.context_dim;
    h->ctx_w = dev_alloc<bf16>(h, (size_t)h->ctx_n * D);
    size_t row = 0;
    for (auto& sb : st_blocks) {
      for (const char* nm : {".attn2.to_k.weight", ".attn2.to_v.weight"}) {
        const DevTensor& w = find_raw(h, sb.first + nm, use_ema);
        if (w.numel != (int64_t)sb.second * D) throw Error(LIDM_ERR_STATE, "weight '" + sb.first + nm + "' has unexpected size");
        launch_pack_conv_weight(w.p, sb.second, D, 1, 1, sb.second, D, nullptr, nullptr, 1.f, 0, h->ctx_w + row * D, pk.s);
        row += sb.second;
      }
    }
  }
  h->out_norm = pk.norm(U + "out.0", ch);
  h->out_conv = pk.conv(U + "out.2", cfg.out_channels, mc, 3, 3);
  if (ch != mc) throw Error(LIDM_ERR_INVALID, "U-Net must end at model_channels");
  }   // openaimodel.UNetModel

  // ---- emb_layers: one concatenated [emb_total][ted] fp32 matrix, evaluated once per step for all ResBlocks
  {
    std::vector<std::pair<std::string, ResW*>> all;
    auto collect = [&](std::vector<Layer>& layers, const std::string& p) {
      for (size_t j = 0; j < layers.size(); ++j)
        if (layers[j].kind == Layer::RES) all.emplace_back(p + "." + std::to_string(j) + ".emb_layers.1", &layers[j].r);
    };
    for (size_t i = 0; i < h->in_blocks.size(); ++i) collect(h->in_blocks[i], U + "input_blocks." + std::to_string(i));
    collect(h->mid_block, U + "middle_block");
    for (size_t i = 0; i < h->out_blocks.size(); ++i) collect(h->out_blocks[i], U + "output_blocks." + std::to_string(i));
    for (auto& e : eff_emb) all.push_back(e);
    int total = 0;
    for (auto& e : all) { e.second->emb_off = total; total += (e.second->film ? 2 : 1) * e.second->cout; }
    h->emb_total = total;
    h->emb_w = dev_alloc<float>(h, (size_t)total * ted);
    h->emb_b = dev_alloc<float>(h, total);
    for (auto& e : all) {
      const DevTensor& w = find_raw(h, e.first + ".weight", use_ema);
      const DevTensor& bb = find_raw(h, e.first + ".bias", use_ema);
      const int erows = (e.second->film ? 2 : 1) * e.second->cout;
      if (w.numel != (int64_t)erows * ted || bb.numel != erows)
        throw Error(LIDM_ERR_STATE, "emb_layers size mismatch at " + e.first);
      LIDM_CUDA_CHECK(cudaMemcpy(h->emb_w + (size_t)e.second->emb_off * ted, w.p, w.numel * sizeof(float), cudaMemcpyDeviceToDevice));
      LIDM_CUDA_CHECK(cudaMemcpy(h->emb_b + e.second->emb_off, bb.p, bb.numel * sizeof(float), cudaMemcpyDeviceToDevice));
    }
  }

  // ---- layout encoder (the cond stage of layout_crossattn models), fp32
  h->has_layout_encoder = h->is_layout && h->raw.count("cond_stage_model.transformer_proj.weight") != 0;
  if (h->has_layout_encoder) {
    const std::string Cn = "cond_stage_model.";
    const int Hd = cfg.encoder_channels;
    LIDM_REQUIRE(cfg.enc_layers >= 0 && cfg.enc_heads >= 1 && cfg.enc_out_dim == ted && cfg.enc_num_classes >= 1,
                 "layout encoder: enc_layers / enc_heads / enc_num_classes, and enc_out_dim must equal 4 * model_channels");
    std::vector<const float*> ptrs;
    for (int i = 0; i < cfg.enc_layers; ++i) {
      const std::string p = Cn + "transform.resblocks." + std::to_string(i);
      ptrs.push_back(pk.f32(p + ".ln_1.weight", Hd)); ptrs.push_back(pk.f32(p + ".ln_1.bias", Hd));
      ptrs.push_back(pk.f32(p + ".attn.c_qkv.weight", (int64_t)3 * Hd * Hd)); ptrs.push_back(pk.f32(p + ".attn.c_qkv.bias", 3 * Hd));
      ptrs.push_back(pk.f32(p + ".attn.c_proj.weight", (int64_t)Hd * Hd)); ptrs.push_back(pk.f32(p + ".attn.c_proj.bias", Hd));
      ptrs.push_back(pk.f32(p + ".ln_2.weight", Hd)); ptrs.push_back(pk.f32(p + ".ln_2.bias", Hd));
      ptrs.push_back(pk.f32(p + ".mlp.c_fc.weight", (int64_t)4 * Hd * Hd)); ptrs.push_back(pk.f32(p + ".mlp.c_fc.bias", 4 * Hd));
      ptrs.push_back(pk.f32(p + ".mlp.c_proj.weight", (int64_t)4 * Hd * Hd)); ptrs.push_back(pk.f32(p + ".mlp.c_proj.bias", Hd));
    }
    h->lenc_layers = dev_alloc<const float*>(h, ptrs.size());
    LIDM_CUDA_CHECK(cudaMemcpy(h->lenc_layers, ptrs.data(), ptrs.size() * sizeof(const float*), cudaMemcpyHostToDevice));
    h->lenc_cls = pk.f32(Cn + "obj_class_embedding.weight", (int64_t)cfg.enc_num_classes * Hd);
    h->lenc_be_w = pk.f32(Cn + "obj_bbox_embedding.weight", (int64_t)Hd * 4);
    h->lenc_be_b = pk.f32(Cn + "obj_bbox_embedding.bias", Hd);
    h->lenc_bx_w = pk.f32(Cn + "obj_bbox_encoding.weight", (int64_t)Hd * 8);
    h->lenc_bx_b = pk.f32(Cn + "obj_bbox_encoding.bias", Hd);
    if (h->raw.count(Cn + "final_ln.weight")) {
      h->lenc_fln_g = pk.f32(Cn + "final_ln.weight", Hd);
      h->lenc_fln_b = pk.f32(Cn + "final_ln.bias", Hd);
    }
    h->lenc_tp_w = pk.f32(Cn + "transformer_proj.weight", (int64_t)ted * Hd);
    h->lenc_tp_b = pk.f32(Cn + "transformer_proj.bias", ted);
  }

  if (h->is_eff) {                 // pixel-space model: no first stage
    LIDM_CUDA_CHECK(cudaDeviceSynchronize());
    for (auto& kv : h->raw) cudaFree(kv.second.p);
    h->raw.clear();
    h->finalized = true;
    return;
  }
  // ---- first stage (decode side): its own numeric mode
  pk.precise = h->ae_prec == LIDM_PREC_BF16X3;
  pk.f16 = h->ae_prec == LIDM_PREC_FP16;
  const std::string A = "first_stage_model.";
  if (cfg.embed_dim != 8 || cfg.z_channels != 8) throw Error(LIDM_ERR_INVALID, "embed_dim and z_channels must be 8");
  h->codebook = pk.f32(A + "quantize.embedding.weight", (int64_t)cfg.n_embed * cfg.embed_dim);
  h->cb_norm = dev_alloc<float>(h, cfg.n_embed);
  launch_codebook_norm(h->codebook, cfg.n_embed, cfg.embed_dim, h->cb_norm, 0);
  h->pq_w = pk.f32(A + "post_quant_conv.weight", (int64_t)cfg.z_channels * cfg.embed_dim);
  h->pq_b = pk.f32(A + "post_quant_conv.bias", cfg.z_channels);
  const std::string D = A + "decoder.";
  const int nres = cfg.ae_n_ch_mult;
  int block_in = cfg.ae_ch * cfg.ae_ch_mult[nres - 1];
  h->dec_top = block_in;
  h->dec_conv_in = pk.conv(D + "conv_in", block_in, cfg.z_channels, 3, 3, (9 * cfg.z_channels + 63) / 64 * 64);
  h->dec_mid1 = pack_res(pk, D + "mid.block_1", block_in, block_in, 3, 3, false);
  h->dec_attn = pack_dec_attn(pk, D + "mid.attn_1", block_in);
  h->dec_mid2 = pack_res(pk, D + "mid.block_2", block_in, block_in, 3, 3, false);
  h->dec_levels.assign(nres, DecLevel());
  int H = cfg.latent_h, W = cfg.latent_w;
  for (int lv = nres - 1; lv >= 0; --lv) {
    DecLevel& L = h->dec_levels[lv];
    if (lv > 0) {
      L.sh = cfg.ae_strides[lv - 1][0]; L.sw = cfg.ae_strides[lv - 1][1];
      if (L.sh == 2 && L.sw == 2) { L.kh = 3; L.kw = 3; }
      else if (L.sh == 1 && L.sw == 2) { L.kh = 1; L.kw = 4; }
      else throw Error(LIDM_ERR_INVALID, "unsupported decoder stride");
      L.has_up = true;
    } else { L.kh = 1; L.kw = 4; }
    const int block_out = cfg.ae_ch * cfg.ae_ch_mult[lv];
    for (int i = 0; i <= cfg.ae_num_res_blocks; ++i) {
      L.blocks.push_back(pack_res(pk, D + "up." + std::to_string(lv) + ".block." + std::to_string(i), block_in, block_out, L.kh, L.kw, false));
      block_in = block_out;
    }
    L.ch = block_in;
    if (L.has_up) {
      const int ukh = (L.sh == 2) ? 3 : 1, ukw = (L.sh == 2) ? 3 : 5;   // UPSAMPLE_STRIDE2KERNEL_DICT (model_lidm.py:44)
      L.up = pk.conv(D + "up." + std::to_string(lv) + ".upsample.conv", block_in, block_in, ukh, ukw);
      H *= L.sh; W *= L.sw;
    }
  }
  h->dec_last = block_in;
  h->img_h = H; h->img_w = W;
  h->dec_norm_out = pk.norm(D + "norm_out", block_in);
  h->dec_conv_out = pk.conv(D + "conv_out", cfg.ae_out_ch, block_in, 1, 4);
  if (cfg.ae_use_mask && cfg.ae_out_ch != 2) throw Error(LIDM_ERR_INVALID, "use_mask needs out_ch == 2");
  // ---- first stage (encode side), only when the state-dict carries it
  h->has_encoder = h->raw.count(A + "encoder.conv_in.weight") != 0;
  if (h->has_encoder) {
    const std::string E = A + "encoder.";
    h->enc_conv_in = pk.conv(E + "conv_in", cfg.ae_ch, cfg.ae_in_channels, 3, 3, (9 * cfg.ae_in_channels + 63) / 64 * 64);
    h->enc_levels.assign(nres, EncLevel());
    int bin = cfg.ae_ch;
    for (int lv = 0; lv < nres; ++lv) {
      EncLevel& L = h->enc_levels[lv];
      bin = cfg.ae_ch * (lv == 0 ? 1 : cfg.ae_ch_mult[lv - 1]);
      const int bout = cfg.ae_ch * cfg.ae_ch_mult[lv];
      for (int i = 0; i < cfg.ae_num_res_blocks; ++i) {
        L.blocks.push_back(pack_res(pk, E + "down." + std::to_string(lv) + ".block." + std::to_string(i), bin, bout, 3, 3, false));
        bin = bout;
      }
      L.ch = bin;
      if (lv != nres - 1) {
        L.has_down = true;
        L.sh = cfg.ae_strides[lv][0]; L.sw = cfg.ae_strides[lv][1];
        // DOWNSAMPLE_STRIDE2{KERNEL,PAD}_DICT (model_lidm.py:64-65): (1,2) -> 3x3, pad (0,1,1,1); (2,2) -> 3x3, pad (0,1,0,1)
        if (L.sh == 1 && L.sw == 2) { L.pl = 0; L.pt = 1; }
        else if (L.sh == 2 && L.sw == 2) { L.pl = 0; L.pt = 0; }
        else throw Error(LIDM_ERR_INVALID, "unsupported encoder stride");
        L.down = pk.conv(E + "down." + std::to_string(lv) + ".downsample.conv", bin, bin, 3, 3);
      }
    }
    h->enc_top = bin;
    h->enc_mid1 = pack_res(pk, E + "mid.block_1", bin, bin, 3, 3, false);
    h->enc_attn = pack_dec_attn(pk, E + "mid.attn_1", bin);
    h->enc_mid2 = pack_res(pk, E + "mid.block_2", bin, bin, 3, 3, false);
    h->enc_norm_out = pk.norm(E + "norm_out", bin);
    // quant_conv (1x1, z_channels -> embed_dim) folded into conv_out: W' = Wq Wo, b' = Wq bo + bq (fp32 on the host)
    {
      const DevTensor& wo = find_raw(h, E + "conv_out.weight", use_ema);
      const DevTensor& bo = find_raw(h, E + "conv_out.bias", use_ema);
      const DevTensor& wq = find_raw(h, A + "quant_conv.weight", use_ema);
      const DevTensor& bq = find_raw(h, A + "quant_conv.bias", use_ema);
      const int zc = cfg.z_channels, ed = cfg.embed_dim;
      const int64_t kk = (int64_t)bin * 9;
      if (wo.numel != zc * kk || bo.numel != zc || wq.numel != (int64_t)ed * zc || bq.numel != ed)
        throw Error(LIDM_ERR_STATE, "encoder conv_out / quant_conv weight size");
      std::vector<float> Wo(wo.numel), Bo(zc), Wq(wq.numel), Bq(ed), Wf((size_t)ed * kk), Bf(ed);
      LIDM_CUDA_CHECK(cudaMemcpy(Wo.data(), wo.p, Wo.size() * 4, cudaMemcpyDeviceToHost));
      LIDM_CUDA_CHECK(cudaMemcpy(Bo.data(), bo.p, Bo.size() * 4, cudaMemcpyDeviceToHost));
      LIDM_CUDA_CHECK(cudaMemcpy(Wq.data(), wq.p, Wq.size() * 4, cudaMemcpyDeviceToHost));
      LIDM_CUDA_CHECK(cudaMemcpy(Bq.data(), bq.p, Bq.size() * 4, cudaMemcpyDeviceToHost));
      for (int o = 0; o < ed; ++o) {
        double bacc = Bq[o];
        for (int m = 0; m < zc; ++m) bacc += (double)Wq[o * zc + m] * Bo[m];
        Bf[o] = (float)bacc;
        for (int64_t k = 0; k < kk; ++k) {
          double acc = 0;
          for (int m = 0; m < zc; ++m) acc += (double)Wq[o * zc + m] * Wo[m * kk + k];
          Wf[o * kk + k] = (float)acc;
        }
      }
      int64_t wshape[4] = {ed, bin, 3, 3}, bshape[1] = {ed};
      auto put = [&](const std::string& name, const std::vector<float>& v, int nd, const int64_t* shp) {
        DevTensor t; t.numel = (int64_t)v.size();
        for (int i = 0; i < nd; ++i) t.shape.push_back(shp[i]);
        LIDM_CUDA_CHECK(cudaMalloc(reinterpret_cast<void**>(&t.p), v.size() * 4));
        LIDM_CUDA_CHECK(cudaMemcpy(t.p, v.data(), v.size() * 4, cudaMemcpyHostToDevice));
        auto it = h->raw.find(name);
        if (it != h->raw.end()) { cudaFree(it->second.p); h->raw.erase(it); }
        h->raw.emplace(name, std::move(t));
      };
      put(E + "conv_out_q.weight", Wf, 4, wshape);
      put(E + "conv_out_q.bias", Bf, 1, bshape);
      Packer pk2{h, false};
      pk2.precise = pk.precise; pk2.f16 = pk.f16;
      h->enc_conv_out = pk2.conv(E + "conv_out_q", ed, bin, 3, 3);
    }
  }
  LIDM_CUDA_CHECK(cudaDeviceSynchronize());
  // raw fp32 copies are no longer needed
  for (auto& kv : h->raw) cudaFree(kv.second.p);
  h->raw.clear();
  h->finalized = true;
}

void ensure_time_buffers(lidm_handle* h, int rows) {
  if (rows <= h->te_rows) return;
  cudaFree(h->te_tmp); cudaFree(h->te_emb); cudaFree(h->emb_out); cudaFree(h->t_dev);
  h->te_tmp = h->te_emb = h->emb_out = nullptr; h->t_dev = nullptr; h->te_rows = 0;
  LIDM_CUDA_CHECK(cudaMalloc(reinterpret_cast<void**>(&h->te_tmp), (size_t)rows * h->ted * sizeof(float)));
  LIDM_CUDA_CHECK(cudaMalloc(reinterpret_cast<void**>(&h->te_emb), (size_t)rows * h->ted * sizeof(float)));
  LIDM_CUDA_CHECK(cudaMalloc(reinterpret_cast<void**>(&h->emb_out), (size_t)rows * h->emb_total * sizeof(float)));
  LIDM_CUDA_CHECK(cudaMalloc(reinterpret_cast<void**>(&h->t_dev), (size_t)rows * sizeof(int64_t)));
  h->te_rows = rows;
}

// timestep_embedding -> time_embed -> all emb_layers (basic.py:278-296, openaimodel.py:732-733, 262)
void run_time_embed(lidm_handle* h, const int64_t* t_dev, int rows, cudaStream_t s, int t_stride = 1,
                    const float* rowbias = nullptr) {
  launch_time_embed(t_dev, rows, h->cfg.model_channels, h->te_w0, h->te_b0, h->te_w2, h->te_b2, h->ted, h->te_tmp,
                    h->te_emb, s, t_stride, rowbias, h->is_eff ? 1 : 0);
  launch_linear_rows(h->te_emb, rows, h->ted, h->emb_w, h->emb_b, h->emb_total, h->emb_out, s);
}

typedef void (*PlanPass)(lidm_handle*, Plan*, bool, size_t*);
PlanPass unet_pass(lidm_handle* h) {
  if (h->is_eff) return build_eff_plan_pass;
  return h->unet_prec == LIDM_PREC_BF16X3 ? build_unet_plan_pass_p : build_unet_plan_pass;
}

void require_layout_cond(lidm_handle* h, int B) {
  if (h->cond_B != B || h->cond_xf_proj == nullptr)
    throw Error(LIDM_ERR_STATE, "layout U-Net: lidm_layout_set_cond has not been called for batch size " + std::to_string(B));
}

void require_ready(lidm_handle* h, int B) {
  if (h == nullptr) throw Error(LIDM_ERR_INVALID, "null handle");
  if (!h->finalized) throw Error(LIDM_ERR_STATE, "lidm_finalize_weights has not been called");
  if (B <= 0) throw Error(LIDM_ERR_INVALID, "batch size must be positive");
}

template <class F>
int guarded(lidm_handle* h, F&& f) {
  try {
    f();
    return LIDM_OK;
  } catch (const Error& e) {
    (h ? h->error : tls_error) = e.what();
    return e.code;
  } catch (const std::exception& e) {
    (h ? h->error : tls_error) = e.what();
    return LIDM_ERR_INVALID;
  }
}

__global__ void qkv_legacy_to_internal_kernel(const float* __restrict__ qkv, int B, int heads, int T, float scale,
                                              bf16* __restrict__ out) {
  // qkv: (B, heads*96, T) with channel = head*96 + part*32 + c  ->  packed (B, T, 3C) bf16 = [q | k | v], q and k scaled
  const int C = heads * 32;
  const int64_t total = (int64_t)B * heads * 96 * T;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    const int t = (int)(i % T);
    int64_t r = i / T;
    const int chn = (int)(r % (heads * 96));
    const int b = (int)(r / (heads * 96));
    const int hd = chn / 96, part = (chn % 96) / 32, c = chn % 32;
    const float v = qkv[i];
    out[((size_t)b * T + t) * (3 * C) + part * C + hd * 32 + c] = __float2bfloat16(part < 2 ? v * scale : v);
  }
}

// Which conditioning tensors this model takes (DiffusionWrapper.forward, ddpm.py:2313-2339).
void check_conditioning(lidm_handle* h, const float* c_concat, const float* context, int ctx_len) {
  if (h->is_layout && (c_concat != nullptr || context != nullptr))
    throw Error(LIDM_ERR_INVALID, "the layout U-Net takes its conditioning through lidm_layout_set_cond");
  const bool wants_concat = h->cfg.in_channels > h->latent_channels;
  if (wants_concat != (c_concat != nullptr))
    throw Error(LIDM_ERR_INVALID, wants_concat ? "this model is concat-conditioned: c_concat is required"
                                               : "this model takes no concat conditioning");
  if (h->has_st != (context != nullptr && ctx_len > 0))
    throw Error(LIDM_ERR_INVALID, h->has_st ? "this model is cross-attention conditioned: a context (B, L, context_dim) is required"
                                            : "this model takes no cross-attention context");
}

// DiffusionWrapper 'concat': xc = torch.cat([x] + c_concat, dim=1).  Writes rows [row0, row0 + B) of the assembled
// (rows_total, in_channels, H, W) buffer and returns the buffer (or x itself when there is nothing to concatenate).
const float* assemble_input(lidm_handle* h, const float* x, const float* c_concat, int B, int row0, int rows_total,
                            cudaStream_t s, bool force_copy = false) {
  const lidm_config& cfg = h->cfg;
  const size_t HW = (size_t)cfg.latent_h * cfg.latent_w;
  const int Cl = h->latent_channels, Cin = cfg.in_channels;
  if (Cin == Cl && !force_copy) return x;
  const size_t need = (size_t)rows_total * Cin * HW;
  if (need > h->xcat_elems) {
    LIDM_CUDA_CHECK(cudaStreamSynchronize(s));
    cudaFree(h->xcat); h->xcat = nullptr; h->xcat_elems = 0;
    LIDM_CUDA_CHECK(cudaMalloc(reinterpret_cast<void**>(&h->xcat), need * sizeof(float)));
    h->xcat_elems = need;
  }
  float* dst = h->xcat + (size_t)row0 * Cin * HW;
  LIDM_CUDA_CHECK(cudaMemcpy2DAsync(dst, Cin * HW * sizeof(float), x, Cl * HW * sizeof(float), Cl * HW * sizeof(float), B,
                                    cudaMemcpyDeviceToDevice, s));
  if (Cin > Cl)
    LIDM_CUDA_CHECK(cudaMemcpy2DAsync(dst + Cl * HW, Cin * HW * sizeof(float), c_concat, (Cin - Cl) * HW * sizeof(float),
                                      (Cin - Cl) * HW * sizeof(float), B, cudaMemcpyDeviceToDevice, s));
  return h->xcat;
}

struct TmpBufs {
  std::vector<void*> p;
  template <class T> T* get(size_t n) {
    void* q = nullptr;
    LIDM_CUDA_CHECK(cudaMalloc(&q, std::max<size_t>(n, 1) * sizeof(T)));
    p.push_back(q);
    return reinterpret_cast<T*>(q);
  }
  ~TmpBufs() { for (void* q : p) cudaFree(q); }
};

}  // namespace
}  // namespace lidm

// =========================================================================================================
// C ABI
// =========================================================================================================
extern "C" {

const char* lidm_last_error(const lidm_handle* h) { return h ? h->error.c_str() : tls_error.c_str(); }

int64_t lidm_launch_count(void) { return g_launch_count.load(); }

int lidm_profile_begin(void) {
  return guarded(nullptr, [&] {
    for (cudaEvent_t e : g_prof.ev) cudaEventDestroy(e);
    g_prof = ProfState();
    g_prof.on = true;
  });
}

int lidm_profile_end(double* ms, double* flops, double* bytes, int64_t* launches) {
  return guarded(nullptr, [&] {
    LIDM_REQUIRE(ms && flops && bytes && launches, "null output");
    g_prof.on = false;
    LIDM_CUDA_CHECK(cudaDeviceSynchronize());
    for (int c = 0; c < PROF_NCAT; ++c) { ms[c] = 0; flops[c] = g_prof.flops[c]; bytes[c] = g_prof.bytes[c]; launches[c] = g_prof.launches[c]; }
    FILE* dump = nullptr;
    if (const char* path = getenv("LIDM_PROFILE_DUMP")) dump = fopen(path, "a");
    for (size_t i = 0; i < g_prof.cat.size(); ++i) {
      float t = 0;
      LIDM_CUDA_CHECK(cudaEventElapsedTime(&t, g_prof.ev[2 * i], g_prof.ev[2 * i + 1]));
      ms[g_prof.cat[i]] += t;
      if (dump) fprintf(dump, "%d,%.4f,%.6g,%.6g,%s\n", g_prof.cat[i], t, g_prof.op_flops[i], g_prof.op_bytes[i],
                        g_prof.label[i].c_str());
    }
    if (dump) fclose(dump);
    for (cudaEvent_t e : g_prof.ev) cudaEventDestroy(e);
    g_prof.ev.clear(); g_prof.cat.clear(); g_prof.label.clear(); g_prof.op_flops.clear(); g_prof.op_bytes.clear();
  });
}

int lidm_create(const lidm_config* cfg, lidm_handle** out) {
  return guarded(nullptr, [&] {
    LIDM_REQUIRE(cfg != nullptr && out != nullptr, "null argument");
    int ndev = 0;
    cudaError_t e = cudaGetDeviceCount(&ndev);
    if (e != cudaSuccess || ndev == 0)
      throw Error(LIDM_ERR_CUDA, "no CUDA device available: this library has no CPU fallback");
    cudaDeviceProp prop;
    int dev = 0;
    LIDM_CUDA_CHECK(cudaGetDevice(&dev));
    LIDM_CUDA_CHECK(cudaGetDeviceProperties(&prop, dev));
    if (prop.major != 10) throw Error(LIDM_ERR_CUDA, std::string("sm_100a (B200) required, found ") + prop.name);
    LIDM_REQUIRE(cfg->n_channel_mult >= 1 && cfg->n_channel_mult <= LIDM_MAX_LEVELS, "n_channel_mult");
    LIDM_REQUIRE(cfg->unet_type == 2 || (cfg->ae_n_ch_mult >= 1 && cfg->ae_n_ch_mult <= LIDM_MAX_LEVELS), "ae_n_ch_mult");
    LIDM_REQUIRE(cfg->model_channels % 64 == 0, "model_channels must be a multiple of 64");
    LIDM_REQUIRE(cfg->unet_type >= 0 && cfg->unet_type <= 2, "unet_type");
    LIDM_REQUIRE(cfg->unet_type == 2 || cfg->num_head_channels == (cfg->unet_type == 1 ? 64 : 32),
                 "num_head_channels must be 32 (openaimodel.UNetModel) / 64 (LayoutDiffusionUNetModel)");
    LIDM_REQUIRE(cfg->ae_ch % 64 == 0, "ae ch must be a multiple of 64");
    LIDM_REQUIRE(cfg->latent_h > 0 && cfg->latent_w > 0 && cfg->scale_factor != 0.f, "latent shape / scale_factor");
    const int down = 1 << (cfg->n_channel_mult - 1);
    LIDM_REQUIRE(cfg->latent_h % down == 0 && cfg->latent_w % down == 0, "latent not divisible by U-Net downsampling");
    const int lw = cfg->latent_w / down, lh = cfg->latent_h / down;
    LIDM_REQUIRE((lw >= 128 ? lw % 128 == 0 : (128 % lw == 0 && (lh % (128 / lw) == 0 || lh * lw == 64 || lh * lw == 32))),
                 "coarsest U-Net level must tile into 128-pixel patches (or hold 32 / 64 pixels)");
    LIDM_REQUIRE(cfg->precision >= LIDM_PREC_BF16 && cfg->precision <= LIDM_PREC_FP16, "precision");
    LIDM_REQUIRE(cfg->ae_precision >= 0 && cfg->ae_precision <= LIDM_PREC_FP16 + 1, "ae_precision");
    lidm_handle* h = new lidm_handle();
    h->cfg = *cfg;
    h->unet_prec = cfg->precision;
    h->ae_prec = cfg->ae_precision == 0 ? cfg->precision : cfg->ae_precision - 1;
    if (h->cfg.ae_in_channels <= 0) h->cfg.ae_in_channels = 1;
    if (h->cfg.transformer_depth <= 0) h->cfg.transformer_depth = 1;
    *out = h;
  });
}

void lidm_destroy(lidm_handle* h) { delete h; }

int lidm_load_weight(lidm_handle* h, const char* name, const float* data, int32_t ndim, const int64_t* shape) {
  return guarded(h, [&] {
    LIDM_REQUIRE(h != nullptr && name != nullptr && data != nullptr && ndim >= 0 && ndim <= 8, "bad argument");
    if (h->finalized) throw Error(LIDM_ERR_STATE, "weights already finalized");
    DevTensor t;
    t.numel = 1;
    for (int i = 0; i < ndim; ++i) { t.shape.push_back(shape[i]); t.numel *= shape[i]; }
    LIDM_CUDA_CHECK(cudaMalloc(reinterpret_cast<void**>(&t.p), std::max<int64_t>(t.numel, 1) * sizeof(float)));
    cudaError_t e = cudaMemcpy(t.p, data, t.numel * sizeof(float), cudaMemcpyDefault);
    if (e != cudaSuccess) { cudaFree(t.p); LIDM_CUDA_CHECK(e); }
    auto it = h->raw.find(name);
    if (it != h->raw.end()) { cudaFree(it->second.p); h->raw.erase(it); }
    h->raw.emplace(name, std::move(t));
  });
}

int lidm_finalize_weights(lidm_handle* h, int32_t use_ema) {
  return guarded(h, [&] {
    LIDM_REQUIRE(h != nullptr, "null handle");
    if (h->finalized) throw Error(LIDM_ERR_STATE, "weights already finalized");
    finalize(h, use_ema != 0);
  });
}

int lidm_vq_quantize(lidm_handle* h, const float* z, float* zq_out, int32_t* idx_out, int32_t B, void* stream) {
  return guarded(h, [&] {
    require_ready(h, B);
    LIDM_REQUIRE(z != nullptr && zq_out != nullptr, "null tensor");
    const lidm_config& cfg = h->cfg;
    launch_vq(z, B, cfg.z_channels, cfg.latent_h * cfg.latent_w, h->codebook, h->cb_norm, cfg.n_embed, 1, nullptr, nullptr,
              1.0f, zq_out, idx_out, reinterpret_cast<cudaStream_t>(stream));
  });
}

int lidm_vq_encode(lidm_handle* h, const float* img, float* z_out, int32_t B, void* stream) {
  return guarded(h, [&] {
    require_ready(h, B);
    LIDM_REQUIRE(img != nullptr && z_out != nullptr, "null tensor");
    if (h->is_eff) throw Error(LIDM_ERR_INVALID, "the R2DM pixel-space model has no first stage");
    if (!h->has_encoder)
      throw Error(LIDM_ERR_STATE, "no encoder weights were loaded (first_stage_model.encoder.* / quant_conv.* missing from the state-dict)");
    cudaStream_t s = reinterpret_cast<cudaStream_t>(stream);
    Plan* P = get_plan(h, h->enc_plans, B, h->ae_prec == LIDM_PREC_BF16X3 ? build_enc_plan_pass_p : build_enc_plan_pass);
    P->x = img; P->out = z_out;
    run_plan(P, s);
  });
}

int lidm_image_shape(const lidm_handle* h, int32_t* c, int32_t* hh, int32_t* ww) {
  if (h == nullptr || !h->finalized) return LIDM_ERR_STATE;
  if (c) *c = h->cfg.ae_use_mask ? 1 : h->cfg.ae_out_ch;
  if (hh) *hh = h->img_h;
  if (ww) *ww = h->img_w;
  return LIDM_OK;
}

int lidm_unet_forward_cond(lidm_handle* h, const float* x, const int64_t* t, const float* c_concat, const float* context,
                           int32_t ctx_len, float* eps_out, int32_t B, void* stream) {
  return guarded(h, [&] {
    require_ready(h, B);
    LIDM_REQUIRE(x != nullptr && t != nullptr && eps_out != nullptr, "null tensor");
    cudaStream_t s = reinterpret_cast<cudaStream_t>(stream);
    check_conditioning(h, c_concat, context, ctx_len);
    Plan* P = get_plan(h, h->unet_plans, B, unet_pass(h),
                       h->has_st ? ctx_len : 0);
    ensure_time_buffers(h, B);
    if (h->is_layout) {
      require_layout_cond(h, B);
      run_time_embed(h, t, B, s, 1, h->cond_xf_proj);     // emb = time_embed(t) + xf_proj (object_cross_unet.py:934-936)
    } else run_time_embed(h, t, B, s);
    P->cond_gen = h->cond_gen;
    P->x = x; P->out = eps_out;
    P->xin = assemble_input(h, x, c_concat, B, 0, B, s);
    P->context = context;
    P->rowadd_base = h->emb_out; P->rowadd_ld = h->emb_total;
    P->ddim_x_prev = nullptr; P->ddim_noise = nullptr; P->ddim_pred_x0 = nullptr; P->ddim_coef = nullptr;
    run_plan(P, s);
  });
}

int lidm_layout_set_cond(lidm_handle* h, int32_t B, int32_t n_layout, const float* xf_proj, const float* xf_out,
                         const float* obj_class_embedding, const float* obj_bbox_embedding, int32_t n_res,
                         const int32_t* res_rows, const float* const* patch_emb, const int32_t* patch_batch, void* stream) {
  return guarded(h, [&] {
    require_ready(h, B);
    if (!h->is_layout) throw Error(LIDM_ERR_INVALID, "this model is not a layout U-Net");
    LIDM_REQUIRE(xf_proj && xf_out && obj_class_embedding && obj_bbox_embedding && n_layout >= 1 && n_layout <= 16,
                 "layout conditioning tensors (1..16 layout tokens)");
    LIDM_REQUIRE(n_res >= 0 && (n_res == 0 || (res_rows && patch_emb && patch_batch)), "patch embedding tables");
    cudaStream_t s = reinterpret_cast<cudaStream_t>(stream);
    const lidm_config& cfg = h->cfg;
    const int E = cfg.encoder_channels;
    const bool f16 = h->unet_prec == LIDM_PREC_FP16;
    const float scale = 1.0f / std::sqrt(std::sqrt(128.0f));
    // (re)allocate when the batch size or a broadcast flag changes; the buffers are read by plans / graphs of this B
    std::vector<OacaW*> blocks;
    auto collect = [&](std::vector<Layer>& ls) { for (Layer& L : ls) if (L.kind == Layer::OACA) blocks.push_back(&L.oa); };
    for (auto& ls : h->in_blocks) collect(ls);
    collect(h->mid_block);
    for (auto& ls : h->out_blocks) collect(ls);
    auto find_res = [&](int rows) {
      for (int i = 0; i < n_res; ++i) if (res_rows[i] == rows) return i;
      throw Error(LIDM_ERR_INVALID, "image_patch_bbox_embedding_for_resolution" + std::to_string(rows) + " missing");
    };
    bool realloc = h->cond_B != B;
    for (OacaW* w : blocks) {
      const int pb = patch_batch[find_res(w->rows)];
      LIDM_REQUIRE(pb == 1 || pb == B, "patch embedding batch must be 1 or B");
      if (w->pos_batch != pb) realloc = true;
    }
    if (realloc) {
      LIDM_CUDA_CHECK(cudaDeviceSynchronize());
      for (void* p : h->cond_owned) cudaFree(p);
      h->cond_owned.clear();
      cudaFree(h->cond_xf_proj); h->cond_xf_proj = nullptr;
      LIDM_CUDA_CHECK(cudaMalloc(reinterpret_cast<void**>(&h->cond_xf_proj), (size_t)B * h->ted * sizeof(float)));
      auto alloc = [&](size_t elems) {
        void* p = nullptr;
        LIDM_CUDA_CHECK(cudaMalloc(&p, elems * sizeof(bf16)));
        LIDM_CUDA_CHECK(cudaMemset(p, 0, elems * sizeof(bf16)));
        h->cond_owned.push_back(p);
        return reinterpret_cast<bf16*>(p);
      };
      for (OacaW* w : blocks) {
        const int T = w->rows * (cfg.latent_w * w->rows / cfg.latent_h);
        w->pos_batch = patch_batch[find_res(w->rows)];
        w->pos_img = alloc((size_t)w->pos_batch * T * w->ch);
        w->klay = alloc((size_t)B * 16 * 2 * w->ch);
        w->vlay = alloc((size_t)B * 16 * w->ch);
      }
      h->cond_B = B;
      ++h->cond_gen;
    }
    h->cond_n_layout = n_layout;
    LIDM_CUDA_CHECK(cudaMemcpyAsync(h->cond_xf_proj, xf_proj, (size_t)B * h->ted * sizeof(float), cudaMemcpyDeviceToDevice, s));
    for (OacaW* w : blocks) {
      const int T = w->rows * (cfg.latent_w * w->rows / cfg.latent_h), C = w->ch;
      // positional half of the queries / image keys: GN32(conv1(image_patch_bbox_embedding)) (object_cross_unet.py:468-476)
      launch_oaca_pos(patch_emb[find_res(w->rows)], w->pos_batch, E, T, w->w_pos, w->b_pos, w->n_img_pos.gamma,
                      w->n_img_pos.beta, C, scale, w->pos_img, T, C, 0, f16, s);
      // positional half of the layout keys: GN32(conv1(obj_bbox_embedding)) (:493-503) -> columns [C, 2C) of klay
      launch_oaca_pos(obj_bbox_embedding, B, E, n_layout, w->w_pos, w->b_pos, w->n_lay_pos.gamma, w->n_lay_pos.beta, C, scale,
                      w->klay, 16, 2 * C, C, f16, s);
      // content keys / values of the layout tokens (:505-520)
      launch_oaca_layout_kv(xf_out, obj_class_embedding, B, E, n_layout, w->n_cls.gamma, w->n_cls.beta, w->w_content,
                            w->b_content, C, scale, w->klay, w->vlay, f16, s);
    }
  });
}

int lidm_layout_encode(lidm_handle* h, const float* layout, int32_t B, int32_t n_layout, float* xf_proj, float* xf_out,
                       float* obj_class_embedding, float* obj_bbox_embedding, int32_t n_res, const int32_t* res_rows,
                       float* const* patch_emb, void* stream) {
  return guarded(h, [&] {
    require_ready(h, B);
    if (!h->has_layout_encoder)
      throw Error(LIDM_ERR_STATE, "no layout encoder weights were loaded (cond_stage_model.* missing from the state-dict)");
    LIDM_REQUIRE(layout && xf_proj && xf_out && obj_class_embedding && obj_bbox_embedding && n_layout >= 1 && n_layout <= 16,
                 "layout encoder tensors (1..16 layout tokens)");
    cudaStream_t s = reinterpret_cast<cudaStream_t>(stream);
    const lidm_config& cfg = h->cfg;
    launch_layout_encoder(layout, B, n_layout, cfg.encoder_channels, cfg.enc_heads, cfg.enc_layers, h->lenc_layers, h->lenc_cls,
                          cfg.enc_num_classes, h->lenc_be_w, h->lenc_be_b, h->lenc_bx_w, h->lenc_bx_b, h->lenc_fln_g, h->lenc_fln_b,
                          h->lenc_tp_w, h->lenc_tp_b, h->ted, xf_proj, xf_out, obj_class_embedding, obj_bbox_embedding, s);
    for (int i = 0; i < n_res; ++i) {
      LIDM_REQUIRE(res_rows[i] >= 1 && cfg.latent_h % res_rows[i] == 0 && patch_emb[i] != nullptr, "patch table resolution");
      launch_patch_table(h->lenc_be_w, h->lenc_be_b, cfg.encoder_channels, res_rows[i], cfg.latent_w * res_rows[i] / cfg.latent_h,
                         patch_emb[i], s);
    }
  });
}

int lidm_unet_forward(lidm_handle* h, const float* x, const int64_t* t, float* eps_out, int32_t B, void* stream) {
  return lidm_unet_forward_cond(h, x, t, nullptr, nullptr, 0, eps_out, B, stream);
}

int lidm_ddim_step(const float* x, const float* eps, const float* noise, float a_t, float a_prev, float sigma_t,
                   float sqrt_one_minus_at, float temperature, float* x_prev, float* pred_x0, int64_t n, void* stream) {
  return guarded(nullptr, [&] {
    LIDM_REQUIRE(x != nullptr && eps != nullptr && x_prev != nullptr && n >= 0, "null tensor");
    if (n == 0) return;
    cudaStream_t s = reinterpret_cast<cudaStream_t>(stream);
    static thread_local float* coef_dev = nullptr;
    if (coef_dev == nullptr) LIDM_CUDA_CHECK(cudaMalloc(reinterpret_cast<void**>(&coef_dev), 5 * sizeof(float) * 64));
    static thread_local int slot = 0;
    slot = (slot + 1) % 64;
    const float c[5] = {a_t, a_prev, sigma_t, sqrt_one_minus_at, temperature};
    LIDM_CUDA_CHECK(cudaMemcpyAsync(coef_dev + slot * 5, c, sizeof(c), cudaMemcpyHostToDevice, s));
    launch_ddim_step(x, eps, noise, coef_dev + slot * 5, x_prev, pred_x0, n, s);
  });
}

int lidm_ddpm_step(const float* x, const float* eps, const float* noise, const float* coef, int32_t B, int64_t n_per_sample,
                   int32_t clip_denoised, float* x_prev, float* x_recon, void* stream) {
  return guarded(nullptr, [&] {
    LIDM_REQUIRE(x != nullptr && eps != nullptr && noise != nullptr && coef != nullptr && x_prev != nullptr, "null tensor");
    LIDM_REQUIRE(B > 0 && n_per_sample > 0, "ddpm_step: empty batch");
    launch_ddpm_step(x, eps, noise, coef, B, n_per_sample, clip_denoised, x_prev, x_recon, reinterpret_cast<cudaStream_t>(stream));
  });
}

int lidm_ddim_sample_cond(lidm_handle* h, float* x_inout, const int64_t* timesteps, const float* sched, int32_t n_steps,
                          const float* noise, float temperature, float* pred_x0_out, int32_t B, const float* c_concat,
                          const float* context, int32_t ctx_len, const float* uncond_concat, const float* uncond_context,
                          float guidance_scale, void* stream) {
  return guarded(h, [&] {
    require_ready(h, B);
    LIDM_REQUIRE(x_inout != nullptr && timesteps != nullptr && sched != nullptr && n_steps > 0, "null argument");
    cudaStream_t s = reinterpret_cast<cudaStream_t>(stream);
    const lidm_config& cfg = h->cfg;
    for (int i = 0; i < n_steps; ++i)
      LIDM_REQUIRE(timesteps[i] >= 0, "negative timestep");
    check_conditioning(h, c_concat, context, ctx_len);
    const bool wants_concat = cfg.in_channels > h->latent_channels;
    // classifier-free guidance (ddim.py:173-180) needs the unconditional twin of every conditioning tensor
    const bool guided = (uncond_concat != nullptr || uncond_context != nullptr) && guidance_scale != 1.0f;
    if (guided) {
      LIDM_REQUIRE(!wants_concat || uncond_concat != nullptr, "guidance: unconditional concat tensor missing");
      LIDM_REQUIRE(!h->has_st || uncond_context != nullptr, "guidance: unconditional context missing");
    }
    const int Bp = guided ? 2 * B : B;       // batch the U-Net plan runs at
    const int L = h->has_st ? ctx_len : 0;
    Plan* P = get_plan(h, h->unet_plans, Bp, unet_pass(h), L);
    ensure_time_buffers(h, std::max(n_steps, B));
    if (h->is_layout) {
      require_layout_cond(h, B);
      LIDM_REQUIRE(!guided, "classifier-free guidance is not wired for the layout U-Net");
    }
    P->cond_gen = h->cond_gen;
    const size_t HW = (size_t)cfg.latent_h * cfg.latent_w;
    const size_t elems = (size_t)B * h->latent_channels * HW;
    if (elems > h->xbuf_elems) {
      cudaFree(h->xa); cudaFree(h->xb); h->xa = h->xb = nullptr; h->xbuf_elems = 0;
      LIDM_CUDA_CHECK(cudaMalloc(reinterpret_cast<void**>(&h->xa), elems * sizeof(float)));
      LIDM_CUDA_CHECK(cudaMalloc(reinterpret_cast<void**>(&h->xb), elems * sizeof(float)));
      h->xbuf_elems = elems;
    }
    if (n_steps > h->coef_rows) {
      cudaFree(h->coef_dev); h->coef_dev = nullptr; h->coef_rows = 0;
      LIDM_CUDA_CHECK(cudaMalloc(reinterpret_cast<void**>(&h->coef_dev), (size_t)n_steps * 5 * sizeof(float)));
      h->coef_rows = n_steps;
    }
    if (h->emb_cur == nullptr) {
      LIDM_CUDA_CHECK(cudaMalloc(reinterpret_cast<void**>(&h->emb_cur), (size_t)h->emb_total * sizeof(float)));
      LIDM_CUDA_CHECK(cudaMalloc(reinterpret_cast<void**>(&h->coef_cur), 8 * sizeof(float)));
    }
    if (elems > h->stage_elems) {
      cudaFree(h->noise_cur); cudaFree(h->pred_scratch); h->noise_cur = h->pred_scratch = nullptr; h->stage_elems = 0;
      LIDM_CUDA_CHECK(cudaMalloc(reinterpret_cast<void**>(&h->noise_cur), elems * sizeof(float)));
      LIDM_CUDA_CHECK(cudaMalloc(reinterpret_cast<void**>(&h->pred_scratch), elems * sizeof(float)));
      h->stage_elems = elems;
    }
    const float* ctx_run = context;
    if (guided) {
      if (2 * elems > h->eps2_elems) {
        cudaFree(h->eps2); h->eps2 = nullptr; h->eps2_elems = 0;
        LIDM_CUDA_CHECK(cudaMalloc(reinterpret_cast<void**>(&h->eps2), 2 * elems * sizeof(float)));
        h->eps2_elems = 2 * elems;
      }
      if (h->has_st) {   // c_in = torch.cat([unconditional_conditioning, c])
        const size_t ce = (size_t)B * ctx_len * cfg.context_dim;
        if (2 * ce > h->ctx2_elems) {
          cudaFree(h->ctx2); h->ctx2 = nullptr; h->ctx2_elems = 0;
          LIDM_CUDA_CHECK(cudaMalloc(reinterpret_cast<void**>(&h->ctx2), 2 * ce * sizeof(float)));
          h->ctx2_elems = 2 * ce;
        }
        LIDM_CUDA_CHECK(cudaMemcpyAsync(h->ctx2, uncond_context, ce * sizeof(float), cudaMemcpyDeviceToDevice, s));
        LIDM_CUDA_CHECK(cudaMemcpyAsync(h->ctx2 + ce, context, ce * sizeof(float), cudaMemcpyDeviceToDevice, s));
        ctx_run = h->ctx2;
      }
    }
    // loop order: i-th iteration uses index = n_steps-1-i (np.flip(ddim_timesteps), ddim.py:136-143)
    std::vector<int64_t> t_loop(n_steps);
    std::vector<float> coef((size_t)n_steps * 5);
    for (int i = 0; i < n_steps; ++i) {
      const int index = n_steps - 1 - i;
      t_loop[i] = timesteps[index];
      for (int k = 0; k < 4; ++k) coef[(size_t)i * 5 + k] = sched[(size_t)index * 4 + k];
      coef[(size_t)i * 5 + 4] = temperature;
    }
    // pageable -> device copies are staged by the runtime before returning, so the host vectors may go out of scope
    LIDM_CUDA_CHECK(cudaMemcpyAsync(h->t_dev, t_loop.data(), n_steps * sizeof(int64_t), cudaMemcpyHostToDevice, s));
    LIDM_CUDA_CHECK(cudaMemcpyAsync(h->coef_dev, coef.data(), coef.size() * sizeof(float), cudaMemcpyHostToDevice, s));
    if (!h->is_layout) run_time_embed(h, h->t_dev, n_steps, s);   // all steps' embeddings at once (t is uniform over the batch)
    LIDM_CUDA_CHECK(cudaMemcpyAsync(h->xa, x_inout, elems * sizeof(float), cudaMemcpyDeviceToDevice, s));
    float* cur = h->xa;
    float* nxt = h->xb;
    for (int i = 0; i < n_steps; ++i) {
      // this step's timestep-embedding row, coefficients and noise slice go to fixed staging addresses, so the same two
      // CUDA graphs (x ping-pong parity) replay the whole loop
      if (h->is_layout)     // per-sample rows: emb = time_embed(t_i) + xf_proj[b] goes through a SiLU before emb_layers
        run_time_embed(h, h->t_dev + i, B, s, 0, h->cond_xf_proj);
      else
        LIDM_CUDA_CHECK(cudaMemcpyAsync(h->emb_cur, h->emb_out + (size_t)i * h->emb_total, (size_t)h->emb_total * sizeof(float),
                                        cudaMemcpyDeviceToDevice, s));
      LIDM_CUDA_CHECK(cudaMemcpyAsync(h->coef_cur, h->coef_dev + (size_t)i * 5, 5 * sizeof(float), cudaMemcpyDeviceToDevice, s));
      if (noise) LIDM_CUDA_CHECK(cudaMemcpyAsync(h->noise_cur, noise + (size_t)i * elems, elems * sizeof(float),
                                                 cudaMemcpyDeviceToDevice, s));
      P->context = ctx_run;
      if (h->is_layout) { P->rowadd_base = h->emb_out; P->rowadd_ld = h->emb_total; }
      else { P->rowadd_base = h->emb_cur; P->rowadd_ld = 0; }
      const float* nz = noise ? h->noise_cur : nullptr;
      if (!guided) {
        P->x = cur;
        P->xin = assemble_input(h, cur, c_concat, B, 0, B, s);
        P->out = nullptr;
        P->ddim_x_prev = nxt; P->ddim_noise = nz; P->ddim_pred_x0 = h->pred_scratch;
        P->ddim_coef = h->coef_cur;
        run_plan_graphed(h, P, (cur == h->xa) ? 0 : 1, s);
      } else {
        // x_in = torch.cat([x] * 2); c_in = torch.cat([uncond, cond]); one 2B evaluation, then the guided update
        assemble_input(h, cur, uncond_concat, B, 0, 2 * B, s, true);
        P->xin = assemble_input(h, cur, c_concat, B, B, 2 * B, s, true);
        P->x = P->xin;
        P->out = h->eps2;
        P->ddim_x_prev = nullptr; P->ddim_noise = nullptr; P->ddim_pred_x0 = nullptr; P->ddim_coef = nullptr;
        run_plan_graphed(h, P, 2, s);
        launch_cfg_ddim_step(cur, h->eps2, guidance_scale, nz, h->coef_cur, nxt, h->pred_scratch, nullptr, (int64_t)elems, s);
      }
      std::swap(cur, nxt);
    }
    if (pred_x0_out != nullptr)
      LIDM_CUDA_CHECK(cudaMemcpyAsync(pred_x0_out, h->pred_scratch, elems * sizeof(float), cudaMemcpyDeviceToDevice, s));
    LIDM_CUDA_CHECK(cudaMemcpyAsync(x_inout, cur, elems * sizeof(float), cudaMemcpyDeviceToDevice, s));
  });
}

int lidm_ddim_sample(lidm_handle* h, float* x_inout, const int64_t* timesteps, const float* sched, int32_t n_steps,
                     const float* noise, float temperature, float* pred_x0_out, int32_t B, void* stream) {
  return lidm_ddim_sample_cond(h, x_inout, timesteps, sched, n_steps, noise, temperature, pred_x0_out, B, nullptr, nullptr,
                               0, nullptr, nullptr, 1.0f, stream);
}

int lidm_cfg_combine(const float* eps2, float guidance_scale, float* eps_out, int64_t n, void* stream) {
  return guarded(nullptr, [&] {
    LIDM_REQUIRE(eps2 != nullptr && eps_out != nullptr && n > 0, "null tensor");
    launch_cfg_ddim_step(nullptr, eps2, guidance_scale, nullptr, nullptr, nullptr, nullptr, eps_out, n,
                         reinterpret_cast<cudaStream_t>(stream));
  });
}

int lidm_vq_decode(lidm_handle* h, const float* z, int32_t force_not_quantize, float* img_out, int32_t* idx_out,
                   int32_t B, void* stream) {
  return guarded(h, [&] {
    require_ready(h, B);
    LIDM_REQUIRE(z != nullptr && img_out != nullptr, "null tensor");
    cudaStream_t s = reinterpret_cast<cudaStream_t>(stream);
    if (h->is_eff) throw Error(LIDM_ERR_INVALID, "the R2DM pixel-space model has no first stage");
    Plan* P = get_plan(h, h->dec_plans, B, h->ae_prec == LIDM_PREC_BF16X3 ? build_dec_plan_pass_p : build_dec_plan_pass);
    P->x = z; P->out = img_out; P->idx_out = idx_out; P->quantize = force_not_quantize ? 0 : 1;
    run_plan(P, s);
  });
}

int lidm_backproject(const float* img, int32_t B, int32_t H, int32_t W, float fov_up_deg, float fov_down_deg,
                     float depth_min, float depth_max, float depth_scale, int32_t log_scale, int32_t input_is_unit,
                     float* xyz_out, uint8_t* mask_out, void* stream) {
  return guarded(nullptr, [&] {
    LIDM_REQUIRE(img != nullptr && xyz_out != nullptr, "null tensor");
    launch_backproject(img, B, H, W, fov_up_deg, fov_down_deg, depth_min, depth_max, depth_scale, log_scale, input_is_unit,
                       xyz_out, mask_out, reinterpret_cast<cudaStream_t>(stream));
  });
}

int lidm_to_uint8_image(const float* x, uint8_t* out, int64_t n, void* stream) {
  return guarded(nullptr, [&] {
    LIDM_REQUIRE(x != nullptr && out != nullptr, "null tensor");
    launch_to_uint8_image(x, out, n, reinterpret_cast<cudaStream_t>(stream));
  });
}

int lidm_compact_points(const float* xyz, const uint8_t* mask, int32_t B, int32_t HW, float* points, int32_t* counts,
                        void* stream) {
  return guarded(nullptr, [&] {
    LIDM_REQUIRE(xyz != nullptr && mask != nullptr && points != nullptr && counts != nullptr, "null tensor");
    launch_compact_points(xyz, mask, B, HW, points, counts, reinterpret_cast<cudaStream_t>(stream));
  });
}

int lidm_chamfer_nn_ex(const float* xyz1, const float* xyz2, int32_t B, int32_t N, int32_t M, int32_t dim, float* dist1,
                       int32_t* idx1, float* dist2, int32_t* idx2, int32_t contract_fma, void* stream) {
  return guarded(nullptr, [&] {
    LIDM_REQUIRE(xyz1 && xyz2 && dist1 && idx1 && dist2 && idx2, "null tensor");
    LIDM_REQUIRE(B > 0 && N > 0 && M > 0 && (dim == 2 || dim == 3), "chamfer: B, N, M > 0 and dim 2 or 3");
    cudaStream_t s = reinterpret_cast<cudaStream_t>(stream);
    launch_nn_dist(xyz1, N, xyz2, M, B, dim, dist1, idx1, s, contract_fma != 0);
    launch_nn_dist(xyz2, M, xyz1, N, B, dim, dist2, idx2, s, contract_fma != 0);
  });
}

int lidm_chamfer_nn(const float* xyz1, const float* xyz2, int32_t B, int32_t N, int32_t M, int32_t dim, float* dist1,
                    int32_t* idx1, float* dist2, int32_t* idx2, void* stream) {
  return lidm_chamfer_nn_ex(xyz1, xyz2, B, N, M, dim, dist1, idx1, dist2, idx2, 1, stream);
}

int lidm_chamfer_backward(const float* xyz1, const float* xyz2, int32_t B, int32_t N, int32_t M, int32_t dim, const float* graddist1,
                          const float* graddist2, const int32_t* idx1, const int32_t* idx2, float* gradxyz1, float* gradxyz2,
                          void* stream) {
  return guarded(nullptr, [&] {
    LIDM_REQUIRE(xyz1 && xyz2 && graddist1 && graddist2 && idx1 && idx2 && gradxyz1 && gradxyz2, "null tensor");
    LIDM_REQUIRE(B > 0 && N > 0 && M > 0 && (dim == 2 || dim == 3), "chamfer: B, N, M > 0 and dim 2 or 3");
    cudaStream_t s = reinterpret_cast<cudaStream_t>(stream);
    LIDM_CUDA_CHECK(cudaMemsetAsync(gradxyz1, 0, (size_t)B * N * dim * sizeof(float), s));
    LIDM_CUDA_CHECK(cudaMemsetAsync(gradxyz2, 0, (size_t)B * M * dim * sizeof(float), s));
    launch_chamfer_grad(xyz1, N, xyz2, M, B, dim, graddist1, idx1, gradxyz1, gradxyz2, s);
    launch_chamfer_grad(xyz2, M, xyz1, N, B, dim, graddist2, idx2, gradxyz2, gradxyz1, s);
  });
}

int lidm_emd_forward(const float* xyz1, const float* xyz2, int32_t B, int32_t n, float eps, int32_t iters, float* dist,
                     int32_t* assignment, void* stream) {
  return guarded(nullptr, [&] {
    LIDM_REQUIRE(xyz1 && xyz2 && dist && assignment, "null tensor");
    // the limits of the reference extension (emd_cuda.cu:232-245)
    LIDM_REQUIRE(B > 0 && B <= 512, "emd: the batch size should be no greater than 512");
    LIDM_REQUIRE(n > 0 && n % 1024 == 0, "emd: the size of the point clouds should be a multiple of 1024");
    LIDM_REQUIRE(iters >= 1, "emd: iters");
    cudaStream_t s = reinterpret_cast<cudaStream_t>(stream);
    void* ws = nullptr;
    LIDM_CUDA_CHECK(cudaMallocAsync(&ws, ((size_t)7 * B * n + B) * 4, s));
    try {
      launch_emd_forward(xyz1, xyz2, B, n, eps, iters, dist, assignment, ws, s);
    } catch (...) {
      cudaFreeAsync(ws, s);
      throw;
    }
    LIDM_CUDA_CHECK(cudaFreeAsync(ws, s));
  });
}

int lidm_emd_backward(const float* xyz1, const float* xyz2, const float* graddist, const int32_t* assignment, int32_t B, int32_t n,
                      float* gradxyz1, void* stream) {
  return guarded(nullptr, [&] {
    LIDM_REQUIRE(xyz1 && xyz2 && graddist && assignment && gradxyz1 && B > 0 && n > 0, "emd backward arguments");
    cudaStream_t s = reinterpret_cast<cudaStream_t>(stream);
    LIDM_CUDA_CHECK(cudaMemsetAsync(gradxyz1, 0, (size_t)B * n * 3 * sizeof(float), s));
    launch_emd_backward(xyz1, xyz2, graddist, assignment, B, n, gradxyz1, s);
  });
}

int lidm_op_circular_conv2d(const float* x, int32_t B, int32_t Cin, int32_t H, int32_t W, const float* weight,
                            const float* bias, int32_t Cout, int32_t kh, int32_t kw, int32_t pad_l, int32_t pad_r,
                            int32_t pad_t, int32_t pad_b, int32_t stride, const float* residual, float* out,
                            void* stream) {
  return guarded(nullptr, [&] {
    LIDM_REQUIRE(x && weight && out && B > 0 && Cin > 0 && Cout > 0 && stride >= 1, "bad argument");
    cudaStream_t s = reinterpret_cast<cudaStream_t>(stream);
    const int Ho = (H + pad_t + pad_b - kh) / stride + 1, Wo = (W + pad_l + pad_r - kw) / stride + 1;
    TmpBufs tmp;
    const int n_alloc = round_n_alloc(Cout);
    const bool strided_ok = stride > 1 && H % stride == 0 && W % stride == 0 && Ho == H / stride && Wo == W / stride &&
                            ((Wo <= 128 && 128 % Wo == 0 && Ho % (128 / Wo) == 0) || Wo % 128 == 0) && std::min(Wo, 128) * stride <= 256;
    const bool implicit = Cin % 64 == 0 && ((stride == 1 && Ho == H && Wo == W) || strided_ok);
    const int K = kh * kw * Cin;
    const int k_alloc = implicit ? K : (K + 63) / 64 * 64;
    bf16* wp = tmp.get<bf16>((size_t)n_alloc * k_alloc);
    launch_pack_conv_weight(weight, Cout, Cin, kh, kw, n_alloc, k_alloc, nullptr, nullptr, 1.f, 0, wp, s);
    GemmB wb; wb.p = wp; wb.n_alloc = n_alloc; wb.ld = k_alloc;
    GemmEpilogue ep;
    ep.bias = bias;
    ep.out_f32_nchw = out;
    View rv;
    if (residual != nullptr) {
      rv.B = B; rv.H = Ho; rv.W = Wo; rv.C = Cout; rv.ld = (Cout + 7) / 8 * 8;
      rv.p = tmp.get<bf16>((size_t)B * Ho * Wo * rv.ld);
      launch_f32_to_nhwc_bf16(residual, B, Cout, Ho * Wo, rv, s);
      ep.residual = rv;
    }
    if (implicit) {
      View a; a.B = B; a.H = H; a.W = W; a.C = Cin; a.ld = Cin; a.hl = pad_l; a.hr = pad_r;
      a.p = tmp.get<bf16>((size_t)B * H * a.Wp() * Cin);
      launch_f32_to_nhwc_bf16(x, B, Cin, H * W, a, s);
      ConvTaps taps = taps_rect(kh, kw, pad_l, pad_t);
      taps.sx = taps.sy = stride;
      launch_conv_gemm(a, taps, wb, Cout, ep, s);
    } else {
      bf16* col = tmp.get<bf16>((size_t)B * Ho * Wo * k_alloc);
      if (Cin % 8 == 0 && k_alloc == K) {
        View a; a.B = B; a.H = H; a.W = W; a.C = Cin; a.ld = Cin;
        a.p = tmp.get<bf16>((size_t)B * H * W * Cin);
        launch_f32_to_nhwc_bf16(x, B, Cin, H * W, a, s);
        launch_im2col_nhwc(a, kh, kw, stride, pad_l, pad_t, Ho, Wo, col, s);
      } else {
        LIDM_REQUIRE(stride == 1 && Ho == H && Wo == W, "small-channel im2col path supports stride 1 'same' convs only");
        launch_im2col_nchw_f32(x, B, Cin, H, W, kh, kw, pad_l, pad_t, col, k_alloc, s);
      }
      View a; a.p = col; a.B = B; a.H = Ho; a.W = Wo; a.C = k_alloc; a.ld = k_alloc;
      launch_conv_gemm(a, taps_1x1(), wb, Cout, ep, s);
    }
    LIDM_CUDA_CHECK(cudaStreamSynchronize(s));
  });
}

int lidm_op_conv2d_stored(const float* x, int32_t B, int32_t Cin, int32_t H, int32_t W, const float* weight, const float* bias,
                          int32_t Cout, int32_t kh, int32_t kw, int32_t pad_l, int32_t pad_r, int32_t pad_t, const float* residual,
                          float res_scale, int32_t halo_kernel, float* out, float* gst_out, void* stream) {
  return guarded(nullptr, [&] {
    LIDM_REQUIRE(x && weight && out && B > 0 && Cin % 64 == 0 && Cout % 64 == 0 && (H * W) % 128 == 0, "bad argument");
    cudaStream_t s = reinterpret_cast<cudaStream_t>(stream);
    TmpBufs tmp;
    const int n_alloc = round_n_alloc(Cout), K = kh * kw * Cin;
    bf16* wp = tmp.get<bf16>((size_t)n_alloc * K);
    launch_pack_conv_weight(weight, Cout, Cin, kh, kw, n_alloc, K, nullptr, nullptr, 1.f, 0, wp, s);
    GemmB wb; wb.p = wp; wb.n_alloc = n_alloc; wb.ld = K;
    GemmEpilogue ep;
    ep.bias = bias;
    View o; o.B = B; o.H = H; o.W = W; o.C = Cout; o.ld = Cout;
    o.p = tmp.get<bf16>((size_t)B * H * W * Cout);
    o.gst_ld = (Cout / 8) * 2; o.gst_slots = H * W / 128;
    o.gst = o.gst_base = tmp.get<float>((size_t)B * o.gst_slots * o.gst_ld);
    ep.out = o;
    if (residual != nullptr) {
      View rv = o; rv.gst = rv.gst_base = nullptr;
      rv.p = tmp.get<bf16>((size_t)B * H * W * Cout);
      launch_f32_to_nhwc_bf16(residual, B, Cout, H * W, rv, s);
      ep.residual = rv; ep.res_scale = res_scale;
    }
    View a; a.B = B; a.H = H; a.W = W; a.C = Cin; a.ld = Cin; a.hl = pad_l; a.hr = pad_r;
    a.p = tmp.get<bf16>((size_t)B * H * a.Wp() * Cin);
    launch_f32_to_nhwc_bf16(x, B, Cin, H * W, a, s);
    const ConvTaps taps = taps_rect(kh, kw, pad_l, pad_t);
    conv_halo64_override(halo_kernel ? 1 : 0);
    const bool ok = !halo_kernel || conv_halo64_applicable(a, taps, wb, Cout, ep);
    if (ok) launch_conv_gemm(a, taps, wb, Cout, ep, s);
    conv_halo64_override(-1);
    LIDM_REQUIRE(ok, "halo-tile kernel requested for a shape it does not take");
    launch_nhwc_bf16_to_f32_nchw(o, out, s);
    if (gst_out != nullptr)
      LIDM_CUDA_CHECK(cudaMemcpyAsync(gst_out, o.gst, (size_t)B * o.gst_slots * o.gst_ld * sizeof(float), cudaMemcpyDeviceToDevice, s));
    LIDM_CUDA_CHECK(cudaStreamSynchronize(s));
  });
}

int lidm_op_groupnorm(const float* x, int32_t B, int32_t C, int32_t H, int32_t W, const float* gamma,
                      const float* beta, float eps, int32_t groups, int32_t silu, float* out, void* stream) {
  return guarded(nullptr, [&] {
    LIDM_REQUIRE(x && gamma && beta && out && B > 0, "bad argument");
    cudaStream_t s = reinterpret_cast<cudaStream_t>(stream);
    TmpBufs tmp;
    View a; a.B = B; a.H = H; a.W = W; a.C = C; a.ld = C;
    a.p = tmp.get<bf16>((size_t)B * H * W * C);
    View y = a; y.hl = 1; y.hr = 2;
    y.p = tmp.get<bf16>((size_t)B * H * y.Wp() * C);
    float* partials = tmp.get<float>((size_t)B * groups * 2 * GN_MAX_CHUNKS);
    launch_f32_to_nhwc_bf16(x, B, C, H * W, a, s);
    launch_groupnorm(a, y, gamma, beta, eps, groups, silu != 0, partials, s);
    launch_nhwc_bf16_to_f32_nchw(y, out, s);
    LIDM_CUDA_CHECK(cudaStreamSynchronize(s));
  });
}

int lidm_op_qkv_attention_legacy(const float* qkv, int32_t B, int32_t heads, int32_t T, float* out, void* stream) {
  return guarded(nullptr, [&] {
    LIDM_REQUIRE(qkv && out && B > 0 && heads > 0 && T > 0 && T % 128 == 0, "bad argument (T must be a multiple of 128)");
    cudaStream_t s = reinterpret_cast<cudaStream_t>(stream);
    const int C = heads * 32;
    TmpBufs tmp;
    bf16* qk = tmp.get<bf16>((size_t)B * T * 3 * C);
    const float scale = kSqrtLog2e / std::sqrt(std::sqrt(32.0f));
    qkv_legacy_to_internal_kernel<<<148 * 8, 256, 0, s>>>(qkv, B, heads, T, scale, qk);
    LIDM_CUDA_CHECK(cudaGetLastError());
    View o; o.B = B; o.H = T / 128; o.W = 128; o.C = C; o.ld = C;
    o.p = tmp.get<bf16>((size_t)B * T * C);
    launch_attention_d32_packed(qk, o, B, T, heads, s, true);
    launch_nhwc_bf16_to_f32_nchw(o, out, s);
    LIDM_CUDA_CHECK(cudaStreamSynchronize(s));
  });
}

}  // extern "C"
