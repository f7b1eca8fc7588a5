// Host-side shared declarations: tensor views, error plumbing, TMA descriptor creation, kernel launchers.
#pragma once
#include <cuda.h>
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <stdint.h>

#include <atomic>
#include <stdexcept>
#include <string>

namespace lidm {

typedef __nv_bfloat16 bf16;

extern std::atomic<int64_t> g_launch_count;
#define LIDM_COUNT_LAUNCH(n) (::lidm::g_launch_count.fetch_add((n), std::memory_order_relaxed))

// Optional per-launch CUDA-event profiler (bench.py's roofline numbers): categories 0 conv-gemm, 1 groupnorm,
// 2 attention, 3 other.  Off by default; zero cost when off.
enum ProfCat { PROF_GEMM = 0, PROF_NORM = 1, PROF_ATTN = 2, PROF_OTHER = 3, PROF_NCAT = 4 };

struct Error : std::runtime_error {
  int code;
  Error(int c, const std::string& m) : std::runtime_error(m), code(c) {}
};

#define LIDM_CUDA_CHECK(expr)                                                                              \
  do {                                                                                                     \
    cudaError_t _e = (expr);                                                                               \
    if (_e != cudaSuccess)                                                                                 \
      throw ::lidm::Error(-2, std::string(#expr) + " failed: " + cudaGetErrorString(_e) + " at " + __FILE__ + \
                                  ":" + std::to_string(__LINE__));                                         \
  } while (0)

#define LIDM_REQUIRE(cond, msg)                                                                            \
  do {                                                                                                     \
    if (!(cond)) throw ::lidm::Error(-1, std::string("invalid argument: ") + (msg) + " [" #cond "]");      \
  } while (0)

// Channels-last bf16 activation view with materialised circular halo columns:
// physical shape (B, H, W + hl + hr, ld) where column hl+w holds logical pixel w, the hl leftmost columns hold
// logical pixels W-hl..W-1 and the hr rightmost hold logical pixels 0..hr-1.  `p` already includes the channel
// offset of the view inside a wider (concatenated) buffer; `ld` is the channel stride of one pixel (elements).
struct View {
  bf16* p = nullptr;
  int B = 0, H = 0, W = 0, C = 0;
  int hl = 0, hr = 0;
  int ld = 0;
  // precise ("fp32-class") mode: the value is carried as two bf16 planes x = hi + lo; hi occupies channels
  // [0, C) of the view, lo occupies [lo_off, lo_off + C) (lo_off == 0 => ordinary single-plane tensor)
  int lo_off = 0;
  int cphys = 0;   // physical channel extent when it is not derivable (im2col'd hi/lo operands); 0 => derived
  // row pitch in pixels of `ld` elements when it differs from W + hl + hr (GEMM outputs only): lets a GEMM write every
  // second pixel of every second row of a 2x larger tensor (ld = 2 * ld_big, wpitch = 2 * W_big / 2 ... see Builder::up)
  int wpitch = 0;
  // GroupNorm statistics produced by the GEMM that wrote this tensor: per (sample, 128-pixel tile, 8-channel granule)
  // partial (sum, sum of squares) of the stored bf16 values, layout gst[(b * gst_slots + slot) * gst_ld + granule * 2 + {0,1}]
  // with `gst` already offset to the first granule of this (channel-sliced) view; null => none
  float* gst = nullptr;
  float* gst_base = nullptr;   // start of the statistics buffer (granule 0 of the underlying tensor): plan-time bookkeeping
  int gst_ld = 0;      // floats per (sample, slot) row = 2 * granules of the underlying buffer
  int gst_slots = 0;   // slots per sample (= H * W / 128 of the buffer)
  int gst_slot0 = 0;   // first slot this GEMM writes (output parity of a folded upsample conv)
  // element format of this 2-byte tensor: false = bf16, true = IEEE half (same layouts, TMA maps and tensor rate; used by
  // the first stage, whose bf16 rounding alone exceeds the 1e-2 final-image budget)
  bool f16 = false;
  int Cphys() const { return cphys ? cphys : (lo_off ? lo_off + C : C); }
  int Wp() const { return W + hl + hr; }
  int pitch() const { return wpitch ? wpitch : Wp(); }
  size_t pix_index(int b, int h, int w) const { return ((size_t)(b * H + h) * Wp() + (w + hl)); }
};

// Kernel launch with the programmatic-dependent-launch attribute (the kernel must call pdl_wait() before touching anything
// an earlier kernel wrote).  LIDM_NO_PDL turns the attribute off (A/B runs): the same kernels then serialise normally.
bool pdl_enabled();
void pdl_set_batch(int batch);   // the engine tells the launchers which batch size the plan it is about to run has
template <class... KArgs, class... Args>
inline void launch_pdl(void (*kernel)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t s, Args&&... args) {
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = grid; cfg.blockDim = block; cfg.dynamicSmemBytes = smem; cfg.stream = s;
  cudaLaunchAttribute attr[1];
  if (pdl_enabled()) {
    attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = attr; cfg.numAttrs = 1;
  }
  LIDM_CUDA_CHECK(cudaLaunchKernelEx(&cfg, kernel, std::forward<Args>(args)...));
}

// ---- TMA maps (cuTensorMapEncodeTiled through the runtime's driver entry point; no -lcuda link) -------
CUtensorMap make_tma_act(const View& v, int box_c, int box_w, int box_h, int swizzle_bytes, int box_b = 1, int sx = 1, int sy = 1);
CUtensorMap make_tma_3d(const void* base, uint64_t d0, uint64_t d1, uint64_t d2, uint64_t stride1_bytes,
                        uint64_t stride2_bytes, uint32_t b0, uint32_t b1, int swizzle_bytes);

// ---- implicit-GEMM convolution (gemm_conv.cu) ------------------------------------------------------------
struct ConvTaps {
  int n = 1;
  int8_t dx[9] = {0};
  int8_t dy[9] = {0};
  int cstep = 0;   // channels between consecutive taps inside an im2col'd operand (0 => taps are spatial shifts)
  // zero padding on W as well as H (plain nn.Conv2d(padding=1) of the layout U-Net, object_cross_unet.py): the operand is
  // a halo-free tensor and every out-of-range column comes from TMA's out-of-bounds zero fill
  bool zero_w = false;
  // strided convolution: output pixel (h, w) reads input (h * sy + dy, w * sx + dx); the operand view is the INPUT tensor
  // (a.H = H_out * sy, a.W = W_out * sx) and arrives through TMA traversal strides - no im2col
  int sx = 1, sy = 1;
};

// fp32 channels-last tensor (no halo): the residual stream of the precise mode
struct ViewF {
  float* p = nullptr;
  int B = 0, H = 0, W = 0, C = 0;
  int ld = 0;
};

struct GemmEpilogue {
  const float* bias = nullptr;     // [N] fp32 or null
  const float* rowadd = nullptr;   // per-sample additive term (timestep embedding): [B][rowadd_ld] or null
  int rowadd_ld = 0;               // 0 => same row for every sample
  View residual;                   // optional bf16 residual (p == null => none), same logical shape as out
  float res_scale = 1.f;           // the residual enters as res_scale * residual
  // optional second A operand appended along K as one more (unshifted) tap: out += a2 * B[:, Kmain : Kmain + a2.C]
  // (a ResBlock's 1x1 skip convolution folded into its second conv: one GEMM, no skip tensor written or re-read)
  View a2;
  // a2 as a RESIDUAL (out += a2, a2.C == N): the appended B block is the identity, so an output-channel tile only needs the
  // a2 channels it covers - the kernel walks those K chunks alone.  The residual then arrives through the TMA ring as full
  // 128-byte rows instead of per-thread 16-byte loads strided by the row pitch (ncu / A-B: 28 of the 65 us of the
  // 256 -> 256 attention output projection at 16x128, B = 64), at the price of BN more K columns on an idle tensor pipe.
  bool a2_diag = false;
  View out;                        // bf16 NHWC output (p == null => none); halos are written when hl/hr > 0
  // columns >= split_n go, transposed, to out_t[(b * (N - split_n) + (n - split_n)) * HW + pixel]  (V^T for attention)
  int split_n = 1 << 30;
  bf16* out_t = nullptr;
  // GEGLU fused into the epilogue: the B rows are packed as blocks of [16 value rows | 16 gate rows]; `out` has N/2
  // channels and receives value * gelu(gate) (exact erf GELU)
  bool geglu = false;
  float* out_f32_nchw = nullptr;   // optional fp32 NCHW output (B, N, H, W)
  float* out_f32_nhwc = nullptr;   // optional fp32 channels-last output (B, H, W, N) (no halo)
  int out_f32_ld = 0;              // channel stride of out_f32_nhwc (0 => N)
  const float* res_f32 = nullptr;  // optional fp32 channels-last residual, channel stride res_f32_ld
  int res_f32_ld = 0;
  // fused DDIM update (only with out_f32_nchw semantics; eps itself is still written to out_f32_nchw if non-null)
  const float* ddim_x = nullptr;   // x_t fp32 NCHW
  const float* ddim_noise = nullptr;
  float* ddim_x_prev = nullptr;
  float* ddim_pred_x0 = nullptr;
  const float* ddim_coef = nullptr;  // device pointer to 5 floats: a_t, a_prev, sigma_t, sqrt(1-a_t), temperature
};

// B operand: K-major bf16 rows [n_alloc][ld] (ld >= ntaps*Cin elements); batched => one matrix per sample.
struct GemmB {
  const bf16* p = nullptr;
  int n_alloc = 0;          // rows available (multiple of the N tile; padding rows are zero)
  int64_t ld = 0;           // row stride in elements (0 => ntaps*nseg*Cin)
  int64_t batch_stride = 0; // elements between per-sample matrices (0 => shared weights)
  // operand-split segments per tap (precise mode, x*w = xh*wh + xl*wh + xh*wl): packed K order [tap][seg][Cin]
  //   1: [w]                      A single plane
  //   2: [w_hi, w_lo]             A single plane (exact bf16 values, e.g. attention output)
  //   3: [w_hi, w_hi, w_lo]       A = hi/lo planes: segments read (hi, lo, hi)
  int nseg = 1;
  bool f16 = false;         // element format (must match the A operand's View::f16)
};

// A: activation view (taps applied on the halo'd view).  N = logical output channels.
void launch_conv_gemm(const View& a, const ConvTaps& taps, const GemmB& wt, int N, const GemmEpilogue& ep,
                      cudaStream_t stream);

// True when launch_conv_gemm will take the TMA-store epilogue for this output, the one that can also emit the
// GroupNorm granule statistics of View::gst.
// halo-tile variant for 64 -> 64 channel convolutions (gemm_halo.cu); launch_conv_gemm dispatches to it
bool conv_halo64_applicable(const View& a, const ConvTaps& taps, const GemmB& wtb, int N, const GemmEpilogue& ep);
void launch_conv_halo64(const View& a, const ConvTaps& taps, const GemmB& wtb, int N, const GemmEpilogue& ep, cudaStream_t stream);
void conv_halo64_override(int on);   // -1: LIDM_GEMM_HALO decides (default); 0 / 1: off / on (parity tests of the two paths)
bool conv_gemm_emits_gstats(const GemmEpilogue& ep, int n_alloc);

// ---- normalisation (norm.cu) -----------------------------------------------------------------------------
// film (optional): per-sample FiLM rows [B][film_ld] = [scale (C) | shift (C)]: y = norm(x) * (1 + scale) + shift, then SiLU
void launch_groupnorm(const View& x, const View& y, const float* gamma, const float* beta, float eps, int groups,
                      bool silu, float* partials /* workspace >= B*groups*2*GN_MAX_CHUNKS floats */, cudaStream_t s,
                      const float* film = nullptr, int film_ld = 0);
constexpr int GN_MAX_CHUNKS = 64;
// same, with the statistics taken from x.gst (written by the producing GEMMs' epilogues): one pass over x
void launch_groupnorm_from_gstats(const View& x, const View& y, const float* gamma, const float* beta, float eps, int groups,
                                  bool silu, cudaStream_t s, const float* film = nullptr, int film_ld = 0);

// ---- attention (attention.cu) ----------------------------------------------------------------------------
// qkv: (B, T, 3C) bf16 = [q | k | v] (plain qkv GEMM output); V consumed as an MN-major tcgen05 operand
// log2_scores: the q / k weights carry log2 e on top of the softmax scale (q . k is a base-2 exponent): the kernel then runs
// its exponentials on the raw scores wherever the row maximum allows reference 0
void launch_attention_d32_packed(const bf16* qkv, const View& out, int B, int T, int heads, cudaStream_t s, bool log2_scores = false);

// q (B,T,q_ld) bf16 (heads at columns head*32, softmax scale already folded in); k / v rows of the context
// (B, L, kv_ld) at columns k_col / v_col; out (B,T,C) view.  Any L >= 1 (ragged last tile is masked).
void launch_cross_attention_d32(const bf16* q, int q_ld, const bf16* kv, int kv_ld, int k_col, int v_col, int L,
                                const View& out, int B, int T, int heads, cudaStream_t s, bool log2_scores = false);

// ---- layout-conditioned denoiser (layout.cu) --------------------------------------------------------------------
// ObjectAwareCrossAttention core (object_cross_unet.py:447-565): qkv (B,T,3C) = [q | k | v] with the 128^-1/4 scale folded
// into q and k; pos (pos_batch,T,C) positional half of queries / image keys (pre-scaled); klay (B,16,2C) layout keys
// [content | positional]; vlay (B,16,C); out (B,T,C).  Head width 64 (+64 positional).  T in {64, 128, 256}.
void launch_oaca_attention(const bf16* qkv, const bf16* pos, int pos_batch, const bf16* klay, const bf16* vlay, int n_layout,
                           const View& out, int B, int T, int C, cudaStream_t s);
// dst[(b*rows + l)*dst_ld + dst_col + c] = scale * GroupNorm32(bias[c] + W[c,:] . src[b,:,l])   (src fp32 (Bsrc,E,Lt))
void launch_oaca_pos(const float* src, int Bsrc, int E, int Lt, const float* W, const float* bias, const float* gamma,
                     const float* beta, int C, float scale, bf16* dst, int rows, int dst_ld, int dst_col, bool f16,
                     cudaStream_t s);
// layout tokens' keys (scaled) and values from xf_out and the class embedding (both fp32 (B,E,Lt)), 16 slots per sample
void launch_oaca_layout_kv(const float* xf_out, const float* cls, int B, int E, int Lt, const float* gamma, const float* beta,
                           const float* Wc, const float* bc, int C, float scale, bf16* klay, bf16* vlay, bool f16,
                           cudaStream_t s);
void launch_avgpool2(const View& x, const View& y, cudaStream_t s);
// ops.Resample (lidm/modules/unets/ops.py:52-143) with the [1,3,3,1] window, ring = True: x2 up / down sampling of a
// halo-free channels-last tensor (circular on W, zeros on H)
// R2DM input convolution: constant coordinate-channel map (once) + per-step convolution of the image channels (layout.cu)
void launch_eff_in_map(const float* w, const float* bias, const float* cenc, int Cx, int Ce, int H, int W, int C0, float* map,
                        cudaStream_t s);
void launch_eff_in_conv(const float* x, const float* w, const float* map, int Cx, const View& out, cudaStream_t s);
void launch_fir_down2(const View& x, const View& y, cudaStream_t s);
void launch_fir_up2(const View& x, const View& y, cudaStream_t s);
// LayoutTransformerEncoder.forward (layout_encoder.py:222-281), fp32, one CTA per sample.  layers_dev: device array of
// n_layers x 12 float pointers {ln_1 g/b, c_qkv w/b, c_proj w/b, ln_2 g/b, c_fc w/b, mlp.c_proj w/b}.  Outputs (NCL fp32):
// xf_proj (B,out_dim), xf_out / obj_class_embedding / obj_bbox_embedding (B,H,L).
void launch_layout_encoder(const float* layout, int B, int L, int H, int heads, int n_layers, const void* layers_dev,
                           const float* cls_emb, int n_classes, const float* be_w, const float* be_b, const float* bx_w,
                           const float* bx_b, const float* fln_g, const float* fln_b, const float* tp_w, const float* tp_b,
                           int out_dim, float* xf_proj, float* xf_out, float* cls_out, float* bbox_out, cudaStream_t s);
void launch_patch_table(const float* be_w, const float* be_b, int H, int rows, int cols, float* out, cudaStream_t s);

// ---- transformer pieces (transformer.cu) -------------------------------------------------------------------
// nn.LayerNorm over the channel dimension of every pixel/token (eps 1e-5), bf16 in / bf16 out, fp32 statistics
void launch_layernorm(const View& x, const View& y, const float* gamma, const float* beta, float eps, cudaStream_t s);
// fp32 (rows, cols) -> bf16 (rows_pad, cols) with zero rows appended
void launch_f32_rows_to_bf16(const float* x, int64_t rows, int64_t rows_pad, int cols, bf16* y, cudaStream_t s);
// classifier-free guidance + DDIM update (ddim.py:173-206): e = e_u + scale (e_c - e_u), then ddim_update(x, e, ...)
// eps2: (2, n) = [uncond | cond]; eps_out (optional) receives the guided eps
void launch_cfg_ddim_step(const float* x, const float* eps2, float scale, const float* noise, const float* coef_dev,
                          float* x_prev, float* pred_x0, float* eps_out, int64_t n, cudaStream_t s);

// ---- sample post-processing (postprocess.cu) ----------------------------------------------------------------
void launch_to_uint8_image(const float* x, uint8_t* y, int64_t n, cudaStream_t s);
// xyz (B,3,HW) + mask (B,HW) -> points (B, HW, 3) with the valid points of sample b packed, in pixel order, at the
// front of block b; counts (B) = number of valid points per sample
// eval_kernels.cu: nearest point of set b (B,m,dim) for every point of set a (B,n,dim): squared distance + index
void launch_nn_dist(const float* a, int n, const float* b, int m, int B, int dim, float* dist, int32_t* idx, cudaStream_t s,
                    bool fma = true);
// Chamfer backward (chamfer3D.cu:155-185): grad_a / grad_b accumulate (callers zero them), one call per direction
void launch_chamfer_grad(const float* a, int n, const float* b, int m, int B, int dim, const float* grad_dist, const int32_t* idx,
                         float* grad_a, float* grad_b, cudaStream_t s);
// EMD auction (emd_cuda.cu:226-284): workspace >= (7 * B * n + B) * 4 bytes
void launch_emd_forward(const float* xyz1, const float* xyz2, int B, int n, float eps, int iters, float* dist, int32_t* assignment,
                        void* workspace, cudaStream_t s, bool fma = true);
void launch_emd_backward(const float* xyz1, const float* xyz2, const float* grad_dist, const int32_t* assignment, int B, int n,
                         float* grad_xyz1, cudaStream_t s);
void launch_compact_points(const float* xyz, const uint8_t* mask, int B, int HW, float* points, int32_t* counts,
                           cudaStream_t s);

// ---- elementwise (elementwise.cu) ------------------------------------------------------------------------
// ancestral DDPM update (ddpm.py:1090-1119) with per-sample coefficients coef[B][5] on the device
void launch_ddpm_step(const float* x, const float* eps, const float* noise, const float* coef, int B, int64_t n_per_sample, int clip,
                      float* x_prev, float* x_recon, cudaStream_t s);
void launch_ddim_step(const float* x, const float* eps, const float* noise, const float* coef_dev, float* x_prev,
                      float* pred_x0, int64_t n, cudaStream_t s);
void launch_backproject(const float* img, int B, int H, int W, float fov_up_deg, float fov_down_deg, float dmin,
                        float dmax, float depth_scale, int log_scale, int input_is_unit, float* xyz, uint8_t* mask,
                        cudaStream_t s);
void launch_im2col_nchw_f32(const float* x, int B, int C, int H, int W, int kh, int kw, int pl, int pt, bf16* out,
                            int kpad, cudaStream_t s, bool f16 = false, bool zero_w = false);
// stride = vertical stride; stride_w = horizontal stride (0 => same as stride)
void launch_im2col_nhwc(const View& x, int kh, int kw, int stride, int pl, int pt, int Ho, int Wo, bf16* out,
                        cudaStream_t s, int stride_w = 0);
void launch_upsample_nearest2x(const View& x, const View& y, cudaStream_t s);
void launch_upsample_bilinear(const View& x, const View& y, cudaStream_t s);
void launch_copy_with_halo(const View& x, const View& y, cudaStream_t s);
void launch_softmax_rows(const float* s, bf16* p, int64_t rows, int cols, cudaStream_t st, bool f16 = false);
void launch_vq(const float* z, int B, int C, int HW, const float* codebook, const float* cb_norm, int n_embed,
               int quantize, const float* pq_w, const float* pq_b, float scale, float* out, int32_t* idx,
               cudaStream_t s);
void launch_codebook_norm(const float* codebook, int n_embed, int dim, float* out, cudaStream_t s);
void launch_time_embed(const int64_t* t_dev, int nt, int model_ch, const float* w0, const float* b0, const float* w2,
                       const float* b2, int ted, float* tmp /* nt*ted */, float* emb_silu /* nt*ted */,
                       cudaStream_t s, int t_stride = 1 /* 0: one timestep for every row */,
                       const float* rowbias = nullptr /* [nt][ted] added to emb before the SiLU (layout encoder xf_proj) */,
                       int style = 0 /* 0: [cos | sin], f_i = P^(-i/half) (basic.py:278-296); 1: [sin | cos], f_i = P^(-i/(half-1))
                                        (R2DM, unets/ops.py:14-27) */);
void launch_linear_rows(const float* x, int nt, int K, const float* w, const float* b, int N, float* out,
                        cudaStream_t s);
void launch_pack_conv_weight(const float* w, int cout, int cin, int kh, int kw, int n_alloc, int k_alloc,
                             const int* row_perm, const float* row_scale_rows, float row_scale, int n_scaled_rows,
                             bf16* out, cudaStream_t s, bool f16 = false);
void launch_mask_select(const float* dec, int B, int HW, float* out, cudaStream_t s);

// ---- precise ("fp32-class") mode helpers (precise.cu): fp32 stream tensors <-> bf16 hi/lo planes -------------
// y.lo_off > 0 => write hi and lo planes (x = hi + lo, both bf16), else a single rounded plane.  Halos are written.
void launch_split_f32(const ViewF& x, const View& y, cudaStream_t s);
void launch_groupnorm_f32(const ViewF& x, const View& y, const float* gamma, const float* beta, float eps, int groups,
                          bool silu, float* partials, cudaStream_t s);
void launch_upsample_bilinear_f32(const ViewF& x, const View& y, cudaStream_t s);
void launch_im2col_nchw_f32_hl(const float* x, int B, int C, int H, int W, int kh, int kw, int pl, int pt, bf16* out,
                               int kpad, cudaStream_t s);
void launch_reorder_weight_f32(const float* w, int cout, int cin, int taps, int kpad, float* out, cudaStream_t s);
// packed K order [tap][seg][cin]; seg values per GemmB::nseg
void launch_pack_conv_weight_split(const float* w, int cout, int cin, int kh, int kw, int n_alloc, int nseg,
                                   const int* row_perm, float row_scale, int n_scaled_rows, bf16* out, cudaStream_t s);
void launch_f32_to_nhwc_bf16(const float* x, int B, int C, int HW, const View& y, cudaStream_t s);
void launch_nhwc_bf16_to_f32_nchw(const View& x, float* y, cudaStream_t s);

}  // namespace lidm
