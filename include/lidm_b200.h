/* lidm_b200.h — C ABI of the B200-native LiDM sampling path (liblidm_b200.so).
 *
 * The reference (AlanLiangC/LiDAR-Layout) has no FFI layer: the path sits behind a plain Python object API.
 * Each entry point below names the reference interface (file:line under the reference tree) whose GPU work it
 * replaces; the Python mirror in lidar_layout_b200/ keeps the reference signatures and calls these through ctypes
 * (see INTEGRATION.md for the binding a reference maintainer would add).
 *
 * Conventions: extern "C"; returns 0 on success, a negative code on failure (never throws); message via
 * lidm_last_error(); plain pointers and sizes only.  Unless stated otherwise pointers are DEVICE pointers to
 * contiguous fp32 NCHW tensors exactly as PyTorch lays them out, borrowed only for the duration of the call; all
 * work is enqueued on the caller's `stream` (a cudaStream_t passed as void*) and is asynchronous to the host.
 * A handle owns packed weights and workspaces for one device and is not thread-safe.  There is no CPU fallback.
 */
#ifndef LIDM_B200_H_
#define LIDM_B200_H_

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define LIDM_OK 0
#define LIDM_ERR_INVALID (-1)   /* bad argument / unsupported configuration */
#define LIDM_ERR_CUDA (-2)      /* CUDA runtime / driver failure */
#define LIDM_ERR_STATE (-3)     /* call order violated (e.g. forward before finalize) or missing weight */

#define LIDM_MAX_LEVELS 8

/* Numeric modes of the tensor-core paths (lidm_config.precision / ae_precision).  All accumulate in fp32 in TMEM.
 *   BF16   : bf16 operands and bf16 activations (north_star bf16 budget: eps within 2e-2)
 *   BF16X3 : "precise", every GEMM as a 3-way bf16 operand split x*w = xh*wh + xl*wh + xh*wl with an fp32 residual
 *            stream (fp32-class: eps within 1e-3, image within 1e-2), about 3x the GEMM work
 *   FP16   : IEEE half operands and activations at the same tensor rate as bf16 (3 more mantissa bits: the first
 *            stage's image error drops from 2.4e-2 to 3e-3); conversions saturate at +-65504 */
#define LIDM_PREC_BF16 0
#define LIDM_PREC_BF16X3 1
#define LIDM_PREC_FP16 2

typedef struct lidm_handle lidm_handle;

/* Mirrors the YAML blocks the reference instantiates from (models/lidm/kitti/uncond/config.yaml):
 * model.params.unet_config.params (UNetModel.__init__, lidm/modules/diffusion/openaimodel.py:445-470) and
 * model.params.first_stage_config.params (+ddconfig) (VQModel.__init__, lidm/models/ae/autoencoder.py:15-50;
 * Decoder.__init__, lidm/modules/diffusion/model_lidm.py:315-383). */
typedef struct lidm_config {
  /* U-Net */
  int32_t in_channels, out_channels, model_channels, num_res_blocks, num_head_channels;
  int32_t n_channel_mult;
  int32_t channel_mult[LIDM_MAX_LEVELS];
  int32_t n_attention_resolutions;
  int32_t attention_resolutions[LIDM_MAX_LEVELS];
  int32_t latent_h, latent_w;
  /* first stage (decode side) */
  int32_t embed_dim, n_embed, z_channels, ae_ch, ae_out_ch, ae_num_res_blocks, ae_use_mask;
  int32_t ae_n_ch_mult;
  int32_t ae_ch_mult[LIDM_MAX_LEVELS];
  int32_t ae_strides[LIDM_MAX_LEVELS][2]; /* ae_n_ch_mult-1 entries (h, w) */
  float scale_factor;                     /* LatentDiffusion.scale_factor (ddpm.py:438,725) */
  /* numeric mode of the U-Net (LIDM_PREC_*), and of the first stage unless ae_precision says otherwise */
  int32_t precision;
  /* conditioning (DiffusionWrapper.forward, ddpm.py:2313-2339):
   * latent_channels: channels of x_t / eps; in_channels - latent_channels > 0 channels come from a 'concat' conditioning
   *   tensor (0 means latent_channels == in_channels).
   * use_spatial_transformer / context_dim / transformer_depth: UNetModel's cross-attention options
   *   (openaimodel.py:465-467): SpatialTransformer blocks replace the AttentionBlocks and take a (B, L, context_dim)
   *   context. */
  int32_t latent_channels;
  int32_t use_spatial_transformer, context_dim, transformer_depth;
  int32_t ae_in_channels;                  /* first-stage ddconfig.in_channels (encoder input), 0 => 1 */
  /* numeric mode of the first stage (decoder, encoder): 0 = same as `precision`, otherwise LIDM_PREC_* + 1.  The
   * benchmarked mix is precision = BF16, ae_precision = FP16 + 1: the decoder's bf16 rounding alone exceeds the 1e-2
   * final-image budget, IEEE half meets it at the same speed. */
  int32_t ae_precision;
  /* denoiser family: 0 = openaimodel.UNetModel (the fields above); 1 = LayoutDiffusionUNetModel
   * (lidm/modules/unets/object_cross_unet.py:632-952, conditioning_key 'layout_crossattn'): FiLM scale-shift ResBlocks,
   * ResBlock up/down-sampling, zero-padded 3x3 convs, ObjectAwareCrossAttention at the `attention_resolutions`
   * (= attention_ds) downsampling rates with num_head_channels = 64.  encoder_channels = hidden width of the layout
   * encoder (channels of xf_out / obj_*_embedding); conditioning arrives through lidm_layout_set_cond. */
  int32_t unet_type;
  int32_t encoder_channels;
  int32_t num_attention_blocks;             /* 0 => 1 */
  /* LayoutTransformerEncoder (cond_stage_config of layout2lidar; lidm/modules/encoders/layout_encoder.py:140-220):
   * hidden_dim = encoder_channels, enc_out_dim = output_dim (= 4 * model_channels), transformer depth / heads, classes */
  int32_t enc_layers, enc_heads, enc_out_dim, enc_num_classes;
  /* unet_type 2 = EfficientUNet, the pixel-space R2DM denoiser (lidm/modules/unets/efficient_unet.py:188-295; no first
   * stage: the ae_* fields are ignored): in_channels (2) image channels on a latent_h x latent_w = resolution grid,
   * model_channels = base_channels, channel_mult[4] = channel_multiplier, eff_res_blocks[4] = num_residual_blocks,
   * GroupNorm with eff_gn_groups groups and eps eff_gn_eps, eff_attn_heads attention heads, Fourier-feature coordinate
   * encoding, ring (circular) padding. */
  int32_t eff_res_blocks[LIDM_MAX_LEVELS];
  int32_t eff_gn_groups, eff_attn_heads;
  float eff_gn_eps;
} lidm_config;

/* Last error message for `h` (or, with h == NULL, for the calling thread's last failed lidm_create / stateless call). */
const char* lidm_last_error(const lidm_handle* h);

/* instantiate_from_config(config.model) (lidm/utils/misc_utils.py:118) for the supported topology. */
int lidm_create(const lidm_config* cfg, lidm_handle** out);
void lidm_destroy(lidm_handle* h);

/* model.load_state_dict(sd, strict=False) (scripts/sample.py:268-273): one call per state-dict tensor, keys in the
 * reference scheme (`model.diffusion_model.*`, `model_ema.*`, `first_stage_model.*`).  `data` is fp32, HOST or
 * DEVICE memory; it is copied before the call returns.  Unknown keys are ignored (strict=False). */
int lidm_load_weight(lidm_handle* h, const char* name, const float* data, int32_t ndim, const int64_t* shape);

/* Packs weights into kernel layouts (bf16, [Cout][tap][Cin]); use_ema != 0 selects the LitEma shadow
 * (`model_ema.<name without dots>`, lidm/modules/ema.py:16-21) where present — the one-time equivalent of
 * DDPM.ema_scope (ddpm.py:174-187).  Fails with LIDM_ERR_STATE and names the first missing tensor. */
int lidm_finalize_weights(lidm_handle* h, int32_t use_ema);

/* LatentDiffusion.apply_model -> DiffusionWrapper.forward -> UNetModel.forward
 * (ddpm.py:900, 2313; openaimodel.py:719-751), unconditional.  x: (B,C,H,W) fp32, t: (B,) int64, eps_out: (B,C,H,W). */
int lidm_unet_forward(lidm_handle* h, const float* x, const int64_t* t, float* eps_out, int32_t B, void* stream);

/* The same hook for conditioned models (DiffusionWrapper.forward, ddpm.py:2313-2339):
 *   'concat'    : c_concat (B, in_channels - latent_channels, H, W) is concatenated to x on the channel axis;
 *   'crossattn' : context (B, ctx_len, context_dim) feeds every SpatialTransformer's cross-attention
 *                 (lidm/modules/attention.py:170-261); ctx_len is free (camera: 4 tokens, text: 77).
 * Pass NULL / 0 for the conditioning the model does not take; a mismatch is LIDM_ERR_INVALID. */
int lidm_unet_forward_cond(lidm_handle* h, const float* x, const int64_t* t, const float* c_concat, const float* context,
                           int32_t ctx_len, float* eps_out, int32_t B, void* stream);

/* Conditioning of the layout U-Net: the output dict of LayoutTransformerEncoder.forward
 * (lidm/modules/encoders/layout_encoder.py:222-281) exactly as LatentDiffusion.apply_model hands it to
 * DiffusionWrapper.forward (ddpm.py:2334-2335, `layout_outputs=kwargs`), all DEVICE fp32:
 *   xf_proj (B, 4*model_channels); xf_out, obj_class_embedding, obj_bbox_embedding (B, encoder_channels, n_layout);
 *   for each attention resolution r = res_rows[i]: image_patch_bbox_embedding_for_resolution{r} = patch_emb[i],
 *   (patch_batch[i], encoder_channels, L1_r) with patch_batch[i] = B, or 1 when every sample holds the same rows (the
 *   reference encoder repeat_interleaves one tensor, layout_encoder.py:251-257).
 * Everything that depends on the conditioning only (positional projections + GroupNorm32 of image patches and layout
 * boxes, the layout tokens' keys and values; object_cross_unet.py:462-520) is computed here once and reused by every
 * lidm_unet_forward / lidm_ddim_sample call at this batch size until the next lidm_layout_set_cond. */
int lidm_layout_set_cond(lidm_handle* h, int32_t B, int32_t n_layout, const float* xf_proj, const float* xf_out,
                         const float* obj_class_embedding, const float* obj_bbox_embedding, int32_t n_res,
                         const int32_t* res_rows, const float* const* patch_emb, const int32_t* patch_batch, void* stream);

/* LatentDiffusion.get_learned_conditioning(layout) -> LayoutTransformerEncoder.forward
 * (ddpm.py:558-569, lidm/modules/encoders/layout_encoder.py:222-281) for the shipped condition types (obj_class, obj_bbox,
 * is_valid_obj), fp32.  layout: DEVICE (B, n_layout, 13) = [bbox 8 | bbox_2d 4 | class 1].  Outputs, DEVICE fp32:
 * xf_proj (B, 4*model_channels); xf_out / obj_class_embedding / obj_bbox_embedding (B, encoder_channels, n_layout);
 * patch_emb[i] (1, encoder_channels, L1) = image_patch_bbox_embedding_for_resolution{res_rows[i]} (one row: the reference
 * repeats it over the batch).  Needs cond_stage_model.* in the loaded state-dict. */
int lidm_layout_encode(lidm_handle* h, const float* layout, int32_t B, int32_t n_layout, float* xf_proj, float* xf_out,
                       float* obj_class_embedding, float* obj_bbox_embedding, int32_t n_res, const int32_t* res_rows,
                       float* const* patch_emb, void* stream);

/* DDIMSampler.p_sample_ddim update arithmetic (lidm/models/diffusion/ddim.py:191-206), stateless:
 * pred_x0 = (x - sqrt(1-a_t) eps)/sqrt(a_t); x_prev = sqrt(a_prev) pred_x0 + sqrt(1-a_prev-sigma^2) eps + sigma noise T.
 * noise and pred_x0 may be NULL. n = element count. Bit-exact with the reference's fp32 op order. */
int lidm_ddim_step(const float* x, const float* eps, const float* noise, float a_t, float a_prev, float sigma_t,
                   float sqrt_one_minus_at, float temperature, float* x_prev, float* pred_x0, int64_t n, void* stream);

/* The ancestral DDPM update (LatentDiffusion.p_sample after the U-Net: predict_start_from_noise + q_posterior + noise,
 * reference lidm/models/diffusion/ddpm.py:219-232, 1090-1119), fused.  coef: device [B][5] fp32 per-sample coefficients
 * (sqrt_recip_alphas_cumprod[t], sqrt_recipm1_alphas_cumprod[t], posterior_mean_coef1[t], posterior_mean_coef2[t],
 * (t != 0) * exp(0.5 * posterior_log_variance_clipped[t])); noise already holds temperature (and dropout).  x_recon may be
 * NULL.  Bit-identical to the reference's eager expression (every product and sum rounded on its own, same order). */
int lidm_ddpm_step(const float* x, const float* eps, const float* noise, const float* coef, int32_t B, int64_t n_per_sample,
                   int32_t clip_denoised, float* x_prev, float* x_recon, void* stream);

/* DDIMSampler.ddim_sampling loop (ddim.py:115-165), unconditional, whole loop on the device with the DDIM update
 * fused into the U-Net's last conv epilogue.  x_inout: x_T in, x_0 estimate out, (B,C,H,W).
 * timesteps: HOST int64[n_steps] ascending (ddim_timesteps); sched: HOST float[n_steps*4] rows
 * {a_t, a_prev, sigma_t, sqrt_one_minus_at} indexed like the reference's `index` (the loop walks them backwards).
 * noise: NULL (eta = 0) or DEVICE fp32 (n_steps,B,C,H,W) consumed in loop order (first row = first iteration).
 * pred_x0_out: NULL or DEVICE (B,C,H,W) receiving the last step's pred_x0. */
int lidm_ddim_sample(lidm_handle* h, float* x_inout, const int64_t* timesteps, const float* sched, int32_t n_steps,
                     const float* noise, float temperature, float* pred_x0_out, int32_t B, void* stream);

/* The same loop for conditioned models, with optional classifier-free guidance (ddim.py:173-180): when an
 * unconditional twin of the conditioning is given and guidance_scale != 1, every step evaluates the U-Net once on the
 * 2B batch [uncond | cond] and steps with e_u + guidance_scale * (e_c - e_u). */
int lidm_ddim_sample_cond(lidm_handle* h, float* x_inout, const int64_t* timesteps, const float* sched, int32_t n_steps,
                          const float* noise, float temperature, float* pred_x0_out, int32_t B, const float* c_concat,
                          const float* context, int32_t ctx_len, const float* uncond_concat, const float* uncond_context,
                          float guidance_scale, void* stream);

/* e_t = e_t_uncond + scale * (e_t - e_t_uncond) (ddim.py:180), stateless: eps2 = (2, n) [uncond | cond] -> eps_out (n). */
int lidm_cfg_combine(const float* eps2, float guidance_scale, float* eps_out, int64_t n, void* stream);

/* LatentDiffusion.decode_first_stage -> VQModelInterface.decode (ddpm.py:717-775, autoencoder.py:290-302):
 * z/scale_factor -> [quantize] -> post_quant_conv -> Decoder -> [use_mask].  z: (B,C,h,w); img_out: (B,1 or out_ch,H,W)
 * fp32; idx_out: NULL or int32 (B*h*w) codebook indices (-1 when force_not_quantize). */
int lidm_vq_decode(lidm_handle* h, const float* z, int32_t force_not_quantize, float* img_out, int32_t* idx_out,
                   int32_t B, void* stream);

/* first_stage_model.quantize(z) (taming VectorQuantizer2.forward, arithmetic as lidm/models/ae/vq.py:66-79; used by the
 * samplers' quantize_x0 / quantize_denoised options, ddim.py:199, ddpm.py:1081): nearest codebook entry per latent pixel.
 * z: (B, embed_dim, h, w); zq_out: same shape (z + (e_idx - z), the straight-through value); idx_out: NULL or int32 (B*h*w). */
int lidm_vq_quantize(lidm_handle* h, const float* z, float* zq_out, int32_t* idx_out, int32_t B, void* stream);

/* LatentDiffusion.encode_first_stage -> VQModelInterface.encode (ddpm.py:837, autoencoder.py:285-288): Encoder
 * (model_lidm.py:284-312) + quant_conv, NOT quantised (the quantiser sits in decode) and not yet multiplied by
 * scale_factor (get_first_stage_encoding, ddpm.py:546-556).  img: (B, in_channels, H, W) fp32; z_out: (B, embed_dim, h, w).
 * Needs first_stage_model.encoder.* / quant_conv.* in the loaded state-dict (LIDM_ERR_STATE otherwise). */
int lidm_vq_encode(lidm_handle* h, const float* img, float* z_out, int32_t B, void* stream);

/* Output geometry of lidm_vq_decode for this config: channels, height, width of img_out. */
int lidm_image_shape(const lidm_handle* h, int32_t* c, int32_t* hh, int32_t* ww);

/* custom_to_pcd + range2xyz / range2pcd geometry (scripts/sample.py:29-35, lidm/utils/lidar_utils.py:134-204),
 * stateless.  img: (B,H,W) fp32 in [-1,1] (clipped and mapped to [0,1] like custom_to_pcd) or, with input_is_unit != 0,
 * already in [0,1] (the argument range2pcd/range2xyz themselves take); xyz_out: (B,3,H,W) fp32 with -1 where masked (range2xyz);
 * mask_out: NULL or uint8 (B,H,W), 1 where depth_min < depth < depth_max (the points range2pcd keeps). */
int lidm_backproject(const float* img, int32_t B, int32_t H, int32_t W, float fov_up_deg, float fov_down_deg,
                     float depth_min, float depth_max, float depth_scale, int32_t log_scale, int32_t input_is_unit,
                     float* xyz_out, uint8_t* mask_out, void* stream);

/* custom_to_pil (scripts/sample.py:38-45), stateless: out[i] = uint8(255 * (clip(x[i], -1, 1) + 1) / 2), the exact fp32
 * op order and truncation of the reference.  x: n fp32 values (16-byte aligned), out: n bytes. */
int lidm_to_uint8_image(const float* x, uint8_t* out, int64_t n, void* stream);

/* The `pcd[mask, :]` gather of range2pcd / custom_to_pcd (lidar_utils.py:169-171, scripts/sample.py:29-35) for a whole
 * batch: xyz (B,3,HW) fp32 and mask (B,HW) uint8 as lidm_backproject wrote them -> points (B,HW,3) fp32 where block b
 * starts with its counts[b] valid points in row-major pixel order (numpy boolean-index order); counts: int32 (B). */
int lidm_compact_points(const float* xyz, const uint8_t* mask, int32_t B, int32_t HW, float* points, int32_t* counts,
                        void* stream);

/* Forward of the reference's Chamfer-distance extension (lidm/eval/modules/chamfer3D/chamfer_cuda.cpp:13-15 ->
 * chamfer3D.cu:12-155; chamfer2D/chamfer2D.cu:12-145; called from lidm/eval/metric_utils.py:414-440), stateless:
 * xyz1 (B,N,dim), xyz2 (B,M,dim) fp32 device pointers, dim 2 or 3.  dist1 (B,N) / idx1 (B,N) int32: squared distance to
 * and index of the nearest point of xyz2 for every point of xyz1; dist2 / idx2 (B,M) the same the other way round.
 * Ties go to the lowest index; d = (dx*dx + dy*dy) + dz*dz with every operation rounded separately. */
int lidm_chamfer_nn(const float* xyz1, const float* xyz2, int32_t B, int32_t N, int32_t M, int32_t dim, float* dist1,
                    int32_t* idx1, float* dist2, int32_t* idx2, void* stream);

/* The same with the rounding of the distance chosen: contract_fma != 0 evaluates fma(dz, dz, fma(dx, dx, dy*dy)), which is
 * how nvcc compiles the reference's `dx*dx + dy*dy + dz*dz` (bit-identical to the reference extension built for sm_100,
 * tests/test_gpu_eval_ref.py) and what lidm_chamfer_nn uses; 0 rounds every operation separately (the numpy form). */
int lidm_chamfer_nn_ex(const float* xyz1, const float* xyz2, int32_t B, int32_t N, int32_t M, int32_t dim, float* dist1,
                       int32_t* idx1, float* dist2, int32_t* idx2, int32_t contract_fma, void* stream);

/* Backward of the Chamfer extension (chamfer_cuda.cpp:22-26 -> chamfer3D.cu:155-185, NmDistanceGradKernel): gradxyz1 (B,N,dim)
 * and gradxyz2 (B,M,dim) are zeroed and receive 2 g (p - nn(p)) from both directions (atomics, like the reference). */
int lidm_chamfer_backward(const float* xyz1, const float* xyz2, int32_t B, int32_t N, int32_t M, int32_t dim, const float* graddist1,
                          const float* graddist2, const int32_t* idx1, const int32_t* idx2, float* gradxyz1, float* gradxyz2,
                          void* stream);

/* The reference's EMD extension (lidm/eval/modules/emd/emd.cpp:17-25 -> emd_cuda.cu:226-284, auction algorithm; called from
 * lidm/eval/metric_utils.py:447-458 with eps 0.005, 50 iterations): xyz1 / xyz2 (B,n,3) fp32 in [0,1], n a multiple of 1024,
 * B <= 512 (the reference's own limits).  dist (B,n) squared distance of every xyz1 point to its assigned xyz2 point;
 * assignment (B,n) int32.  Workspaces are allocated from the stream's memory pool. */
int lidm_emd_forward(const float* xyz1, const float* xyz2, int32_t B, int32_t n, float eps, int32_t iters, float* dist,
                     int32_t* assignment, void* stream);
/* emd_cuda.cu:286-316: gradxyz1 (B,n,3) = 2 g (p - assigned(p)); only xyz1 receives a gradient, as in the reference. */
int lidm_emd_backward(const float* xyz1, const float* xyz2, const float* graddist, const int32_t* assignment, int32_t B, int32_t n,
                      float* gradxyz1, void* stream);

/* ---- operator-level entry points (the same kernels the model uses; exposed for parity tests and reuse) ---------
 * All tensors fp32 NCHW device pointers; conversions to the internal channels-last bf16 layout happen inside. */

/* CircularConv2d.forward (lidm/modules/basic.py:52-59): circular pad (pad_l,pad_r) on W, zero pad (pad_t,pad_b) on H,
 * conv stride `stride`.  weight: (Cout,Cin,kh,kw), bias: (Cout) or NULL, residual: NULL or (B,Cout,Ho,Wo) added. */
int lidm_op_circular_conv2d(const float* x, int32_t B, int32_t Cin, int32_t H, int32_t W, const float* weight,
                            const float* bias, int32_t Cout, int32_t kh, int32_t kw, int32_t pad_l, int32_t pad_r,
                            int32_t pad_t, int32_t pad_b, int32_t stride, const float* residual, float* out,
                            void* stream);

/* The same convolution as the MODEL runs it (stride 1, Cin and Cout multiples of 64, H*W a multiple of 128): the output is
 * stored channels-last bf16 and the epilogue also writes the GroupNorm granule statistics of the stored values, both of which
 * this entry point hands back (out: the stored values widened to fp32 NCHW; gst_out: (B, H*W/128, Cout/8, 2) partial sums and
 * sums of squares, or NULL).  residual enters as res_scale * residual.  halo_kernel = 1 forces the halo-tile kernel
 * (gemm_halo.cu; error if the shape is not one it takes), 0 the streamed implicit-GEMM kernel: the two must agree bit for bit. */
int lidm_op_conv2d_stored(const float* x, int32_t B, int32_t Cin, int32_t H, int32_t W, const float* weight, const float* bias,
                          int32_t Cout, int32_t kh, int32_t kw, int32_t pad_l, int32_t pad_r, int32_t pad_t, const float* residual,
                          float res_scale, int32_t halo_kernel, float* out, float* gst_out, void* stream);

/* GroupNorm32 (+SiLU) (lidm/modules/basic.py:339-341, openaimodel.py:205-207). */
int lidm_op_groupnorm(const float* x, int32_t B, int32_t C, int32_t H, int32_t W, const float* gamma,
                      const float* beta, float eps, int32_t groups, int32_t silu, float* out, void* stream);

/* QKVAttentionLegacy.forward (openaimodel.py:358-374): qkv (B, heads*3*32, T) -> out (B, heads*32, T). */
int lidm_op_qkv_attention_legacy(const float* qkv, int32_t B, int32_t heads, int32_t T, float* out, void* stream);

/* Number of kernels this library has launched so far in the process (bench.py's gpu_launches). */
int64_t lidm_launch_count(void);

/* Per-op CUDA-event profiler for bench.py's roofline line.  Between begin and end every op of a model-level call
 * is bracketed by events on the launching stream.  Arrays of 4: {conv-gemm, groupnorm, attention, other};
 * ms = summed device time, flops / bytes = summed ALGORITHMIC work, launches = kernels launched. */
int lidm_profile_begin(void);
int lidm_profile_end(double* ms, double* flops, double* bytes, int64_t* launches);

#ifdef __cplusplus
}
#endif
#endif /* LIDM_B200_H_ */
