/*
 * b200lidar.h -- C-ABI of libb200lidar.so: the sm_100a kernels behind the LiDARCrafter denoiser hot path.
 *
 * Conventions: plain pointers + sizes, no torch types; every pointer is a
 * DEVICE pointer owned by the caller; no allocation inside the library; every call is asynchronous on
 * `stream` (a cudaStream_t passed as void*) and CUDA-graph capturable; return value 0 = ok, negative =
 * B200_E_* (never exit()).  Activations are NHWC ("pixels x channels") fp32 unless noted; the sampler
 * state x_t / eps prediction stay in the reference's NCHW [B,2,H,W] layout.
 *
 * Each entry point names the reference code it replaces (paths relative to /root/reference/lidargen).
 */
#ifndef B200LIDAR_H_
#define B200LIDAR_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define B200_OK 0
#define B200_E_ARG (-1)     /* invalid argument / unsupported shape */
#define B200_E_CUDA (-2)    /* CUDA runtime error at launch */
#define B200_E_ARCH (-3)    /* device is not sm_100 */

/* library / device info ------------------------------------------------------------------------- */
int b200_version(void);
/* returns SM count (>0) if device `dev` is compute capability 10.x, else B200_E_ARCH */
int b200_device_check(int dev);
const char* b200_last_error(void);

/* ---- K1: ring conv as implicit GEMM on tcgen05 tensor cores ------------------------------------
 * replaces: F.pad(circular W / zero H) + nn.Conv2d  (models/unets/ops.py:32-49,149-173) together with
 *           the bias add, the residual add and the 1/sqrt(2) scale of ResidualBlock.forward
 *           (models/unets/efficient_unet.py:112-115), and the GroupNorm statistics of the NEXT norm.
 *   a        : "conv operand layout" (tile-major slabs): [planes][B][H][W/128][Cin/8][130][8] fp16-sized elements.
 *              For one image row, one 128-pixel tile and one 8-channel group: 130 pixels at a 16-byte pitch = the ring
 *              neighbour (w0-1) mod W, the tile's 128 pixels, the ring neighbour (w0+128) mod W.  That IS the tcgen05
 *              no-swizzle K-major shared-memory image of the tile with its 3x3 halo, and the slabs of consecutive
 *              channel groups are contiguous, so the TMA engine stages one (row, plane, K chunk) with ONE bulk copy.
 *              Written (halo duplicates included) by b200_gn_act_f16 / the attention kernels;
 *              parts = 1: plain fp16;  parts = 2: a = a[0] (hi) + a[1] (lo), the error-compensated split
 *              (3 tensor-core MMAs per product, ~fp32 accuracy: ~2e-6 through the UNet);
 *              parts = 3 ("fp16 + fp8 correction", ~5e-5 through the UNet, inside the 1e-3 tolerance): plane 0 =
 *              fp16 hi as above; plane 1 (same byte size) = e4m3 pairs [B][H][W/128][Cin/16][2][130][16 bytes]: for
 *              every 16-channel chunk and pixel one 16-byte unit L8 = e4m3((a - hi) * 2^11) and one unit A8 = e4m3(a).
 *              The kernel evaluates  hi x w16  with kind::f16 and BOTH cross terms a_lo*w_hi + a_hi*w_lo with ONE
 *              kind::f8f6f4 MMA (K = 32 = [L8 | A8] x [e4m3(w 2^-11) | e4m3(w - w16)]): 2 MMA time units instead
 *              of 3.  wscale must put max|w * wscale| in [2^14, 2^15) (both e4m3 weight planes in normal range).
 *              parts = 4: as 3 with the correction MMAs accumulating into their own TMEM columns (added in the
 *              epilogue) -- diagnostic variant, needs 2*rows*bn <= 256.
 *   wpacked  : fp16 image from b200_pack_conv_weight (same bn, same parts), pre-scaled by wscale = 1/w_inv
 *   out      : fp32 [B,H,W,Cout] = (conv(a)*w_inv + bias + res) * out_scale   (res may be NULL / == out)
 *   stats    : fp64 [B,Cout,2] += {sum, sum of squares} of `out` over H*W  (may be NULL)
 *   taps     : 9 (3x3, padding 1) or 1 (1x1);  ring: 1 = circular in W, 0 = zero pad
 *   bn       : output-channel tile (64 or 128, Cout % bn == 0);  rows: image rows per tile (1,2,4; H % rows == 0,
 *              rows*bn <= 256: two TMEM accumulator sets so the epilogue overlaps the next tile's MMAs)
 * constraints: W % 128 == 0, Cin % 32 == 0.  Persistent kernel: grid = min(#tiles, #SMs).          */
int b200_conv_tc(const void* a, const void* wpacked, const float* bias, const float* res,
                 float out_scale, float w_inv, float* out, double* stats, int B, int H, int W, int Cin,
                 int Cout, int taps, int ring, int bn, int rows, int parts, void* stream);
/* b200_conv_tc with the reduction split over `splits` (2, 4 or 8; (Cin / 16) % splits == 0, Cin / 32 for parts 1) CTAs per
 * output tile: for the layers with few tiles and a long K (deep levels at small batch -- 4x128 C512 at B = 1 is 32 CTAs that
 * each walk K = 4608 alone).  workspace: fp32 [splits][B, H*W, Cout] scratch (raw partial sums of the channel slices); a
 * second kernel adds the slices in index order and applies the epilogue of b200_conv_tc.  Same arguments / result otherwise
 * (up to the fp32 summation order).                                                                    */
int b200_conv_tc_splitk(const void* a, const void* wpacked, const float* bias, const float* res,
                        float out_scale, float w_inv, float* out, double* stats, float* workspace, int splits,
                        int B, int H, int W, int Cin, int Cout, int taps, int ring, int bn, int rows, int parts,
                        void* stream);
/* GroupNorm(+AdaGN)-apply + SiLU + operand split fused IN FRONT of b200_conv_tc: replaces  conv(silu(norm(x)))  of
 * ResidualBlock.forward (models/unets/efficient_unet.py:104-115, ops.py:176-200; layout_unet_v1.py:229-249), the GroupNorm
 * -> QKV projection of the attention blocks (efficient_unet.py:46-49) and the plain fp32 -> operand casts in ONE launch.
 * The A operand never exists in HBM: four transform warps of the conv kernel read the fp32 NHWC activation
 *   x = x0 [B,H*W,C0] (| x1 [B,H*W,C1], channel concatenation; x1 NULL <=> C1 == 0; C0 % 16 == C1 % 16 == 0),
 * apply  y = act(x * a_c + b_c)  with the per-(sample, channel) coefficients of GroupNorm(groups, eps)[*gamma + beta]
 * [*(1 + ada[b, c]) + ada[b, Cin + c]] computed in the kernel from the COMPLETE per-channel statistics stats0 / stats1
 * (fp64 {sum, sum of squares} [B,C,2], as accumulated by the producing kernel; stats0 NULL: no normalisation, a = 1,
 * b = 0), act = SiLU if silu else identity, split y into the operand planes of `parts` (2: fp16 hi + lo; 3: fp16 hi +
 * e4m3 pair) and write them straight into the K-major shared-memory slab (rows above / below the image and non-ring edges
 * are exact zeros) that the tcgen05 issuers read.  Everything after the operand is b200_conv_tc: same wpacked / bias /
 * res / out / stats / tiles, Cin = C0 + C1 <= 1024, groups <= 32.  Result == b200_gn_act_f16 followed by b200_conv_tc
 * (bit-identical operands).  640 threads per CTA, register file re-balanced with setmaxnreg.
 * rows = 0 selects the COLUMN WALK (csrc/conv_col.cuh) for the 64 -> 64 channel 3x3 layers (parts == 3, taps == 9, bn == 64,
 * C0 == 64, C1 == 0, Cout == 64; wpacked from b200_pack_conv_weight with rows = 0): a CTA owns a run of image rows of one
 * 128-pixel column, converts every input row ONCE into shared memory, keeps all weights resident and adds a row's contribution
 * to its three output rows with one N = 192 MMA per (chunk, dx, operand plane); same result up to the fp32 summation order.  */
int b200_conv_gn_tc(const float* x0, int C0, const float* x1, int C1, const double* stats0,
                    const double* stats1, const float* gamma, const float* beta, const float* ada,
                    int ada_stride, int groups, float eps, int silu, const void* wpacked, const float* bias,
                    const float* res, float out_scale, float w_inv, float* out, double* stats, int B, int H,
                    int W, int Cout, int taps, int ring, int bn, int rows, int parts, void* stream);
/* profiling aid: device buffer of [#CTAs][8] uint64 cycle counters filled by b200_conv_tc (NULL disables; see
 * conv_tc.cu g_conv_dbg for the slot meaning).  Not used on the product path.                          */
int b200_conv_set_debug(void* dbg_u64);
/* diagnostic builds only (-DB200_CONV_ABLATE): timing ablations of the conv pipeline (see conv_tc.cu); mask 0 = off */
int b200_conv_set_ablate(int mask);
/* number of fp16-sized elements of the packed weight image (== planes*taps*Cout*Cin, planes = 1 for parts 1 else 2) */
size_t b200_packed_weight_elems(int Cout, int Cin, int taps, int parts);
/* w: fp32 OIHW [Cout,Cin,k,k] (k*k == taps), multiplied by wscale (a power of two), ->
 * packed fp16 tiles [Cout/bn][Cin/KC][taps][parts][KC/8][bn][8] (merged mode: [..][KC/8][parts][bn][8]),
 * KC = 32 (parts 1) or 16 (parts 2);  parts 3/4: per (n-tile, 16-channel chunk, tap) [2][bn][8] fp16(w*wscale)
 * followed by [2][bn][16] e4m3 {w*wscale*2^-11, w*wscale - fp16(w*wscale)};
 * rows == 0 (column walk of b200_conv_gn_tc; parts 3, taps 9, Cin = Cout = bn = 64): per (16-channel chunk, dx) two planes of
 * [2 k-groups][192 = dy*64 + cout][16 B]: fp16(w*wscale) | e4m3 {w*wscale*2^-11, w*wscale - fp16(w*wscale)}            */
int b200_pack_conv_weight(const float* w, void* wpacked, int Cout, int Cin, int taps, int bn, int rows,
                          int parts, float wscale, void* stream);
/* 1 if b200_conv_tc runs (bn, rows, parts) in merged mode (hi/lo weight rows adjacent: the packed image and the
 * tile must be created with the same (bn, rows, parts))                                              */
int b200_conv_merged(int bn, int rows, int parts);
/* plain fp16 copy w16[parts][tap][Cout][Cin] for the CUDA-core checking kernel (parts 3: planes 1, 2 hold the
 * fp16 images of the two e4m3 weight planes) */
int b200_pack_conv_weight_plain(const float* w, void* w16, int Cout, int Cin, int taps, int parts,
                                float wscale, void* stream);
/* CUDA-core (FFMA) implementation of exactly the same contract as b200_conv_tc, weights from
 * b200_pack_conv_weight_plain.  Debug / cross-check path (tests), any W, Cin % 8 == 0.            */
int b200_conv_ffma(const void* a, const void* w16, const float* bias, const float* res, float out_scale,
                   float w_inv, float* out, double* stats, int B, int H, int W, int Cin, int Cout, int taps,
                   int ring, int parts, void* stream);

/* ---- GroupNorm apply (+AdaGN) + SiLU + fp16 cast (+ channel concat) ------------------------------
 * replaces: nn.GroupNorm -> SiLU  (efficient_unet.py:77-79,106-110), AdaGN (ops.py:176-200),
 *           torch.cat of skip features (efficient_unet.py:295-297).
 *   y[b,p,c] = act( ((x[b,p,c]-mean)*rstd*gamma[c]+beta[c]) * (1+scale[b,c]) + shift[b,c] )  as fp16
 *   x0/x1    : fp32 [B,HW,C0] / [B,HW,C1] sources concatenated along C (x1 may be NULL, C1 = 0)
 *   stats0/1 : fp64 [B,C,2] per-channel {sum,sumsq} of the sources; NULL => no normalisation (cast only)
 *   gamma/beta: [C0+C1] or NULL;  ada: fp32, scale at ada[b*ada_stride + c], shift at
 *               ada[b*ada_stride + (C0+C1) + c], NULL => none;  silu: 1 = apply x*sigmoid(x)
 *   y        : conv operand layout [planes][B][H][W/128][(C0+C1)/8][130][8] (see b200_conv_tc; W % 128 == 0);
 *              parts = 2 also writes the residual lo = fp16(v - fp32(hi)); parts = 3 writes the e4m3 pair plane
 *              (see b200_conv_tc; needs (C0+C1) % 32 == 0)
 *   y_raw    : optional second output (same layout): the UN-normalised concatenated input as a conv operand
 *              (input of the block's 1x1 skip conv, efficient_unet.py:92-96) -- one read of x feeds both     */
int b200_gn_act_f16(const float* x0, int C0, const float* x1, int C1, const double* stats0,
                    const double* stats1, const float* gamma, const float* beta, const float* ada,
                    int ada_stride, int groups, float eps, int silu, void* y, void* y_raw, int parts, int B,
                    int H, int W, void* stream);
/* fp32 NHWC variant y = act(GroupNorm(x)) (no concat / AdaGN): feeds the FIR resampler of the up/down
 * ResBlocks (layout_unet_v1.py:229-235) and the final `out` head (layout_unet_v1.py:899-900)        */
int b200_gn_act_f32(const float* x, const double* stats, const float* gamma, const float* beta, int groups,
                    float eps, int silu, float* y, int B, int HW, int C, void* stream);
/* per-(b,c) {sum,sumsq} of an fp32 NHWC tensor, accumulated (+=) into stats fp64 [B,C,2] */
int b200_channel_stats(const float* x, double* stats, int B, int HW, int C, void* stream);

/* ---- K3: FIR resampling ------------------------------------------------------------------------
 * replaces: ops.Resample (ops.py:52-146), window [1,3,3,1], circular in W / zero in H.
 *   up=0: [B,H,W,C] -> [B,H/2,W/2,C];  up=1: [B,H,W,C] -> [B,2H,2W,C];  stats (optional) as above     */
int b200_fir_resample(const float* x, float* y, double* stats, int B, int H, int W, int C, int up,
                      int ring, void* stream);
/* FIR x2 upsample (as b200_fir_resample with up = 1) written directly as the conv operand (layout of b200_conv_tc,
 * `parts` planes) of the ring conv that follows it in Block.forward (efficient_unet.py:176-190): x fp32 [B,H,W,C] ->
 * y operand of the image [B,2H,2W,C].  2W % 128 == 0, C % 8 == 0 (parts 3: C % 32 == 0).                     */
int b200_fir_up_operand(const float* x, void* y, int parts, int B, int H, int W, int C, int ring, void* stream);

/* ---- K4: time embedding + all per-block (scale,shift) projections --------------------------------
 * replaces: SinusoidalPositionalEmbedding + 2 Linear (efficient_unet.py:237-242, ops.py:14-26) and the
 *           24 AdaGN proj Linear layers (ops.py:190-200), one launch per denoiser step.
 *   t [B] (log-SNR) -> temb [B,E] = W2 silu(W1 sincos(t) + b1) + b2 (+ temb_add[B,E] if not NULL)
 *   ada [B,P] = Wp silu(temb) + bp   where Wp [P,E] stacks every block's projection                  */
int b200_time_embed(const float* t, const float* w1, const float* b1, const float* w2, const float* b2,
                    const float* temb_add, const float* wp, const float* bp, float* temb, float* ada,
                    int B, int Cs, int E, int P, void* stream);

/* ---- first / last convs on CUDA cores ------------------------------------------------------------
 * in_conv: out[b,h,w,:] = cst[(b),h,w,:] + ring_conv3x3(x[b,0:Cx]) ; x NCHW fp32 [B,Cx,H,W] (Cx <= 4),
 *          w fp32 [Cout,Cx,3,3], cst fp32 [Bc,H,W,Cout] (Bc = 1 broadcast or B) already holds bias +
 *          the conv of the step-invariant channels (Fourier features / layout condition).
 *          replaces in_conv + FourierFeatures + cat (efficient_unet.py:283-289, encoding.py:141-146)  */
int b200_in_conv(const float* x, const float* w, const float* cst, int cst_batched, float* out,
                 double* stats, int B, int H, int W, int Cx, int Cout, int ring, void* stream);
/* generic small fp32 direct conv (one-time constant folding): x NHWC fp32 [B,H,W,Cin], w OIHW fp32     */
int b200_conv_direct_f32(const float* x, const float* w, const float* bias, float* out, int B, int H,
                         int W, int Cin, int Cout, int k, int ring, void* stream);
/* out_conv: pred NCHW fp32 [B,Cout<=4,H,W] = ring_conv3x3(a) + bias ; a fp32 NHWC (a_is_f16=0) or fp16 */
int b200_out_conv(const void* a, int a_is_f16, const float* w, const float* bias, float* pred, int B,
                  int H, int W, int Cin, int Cout, int ring, void* stream);

/* ---- K2: attention ---------------------------------------------------------------------------------
 * softmax(q k^T * scale) v per (batch, head), flash style (online softmax; scores never leave the SM).
 *   self-attention (replaces the nn.MultiheadAttention core, efficient_unet.py:39-53):
 *     qkv fp32 [B,T,3E] (q | k | v of the fused in-projection, head-major), d = E / heads = 32 or 64
 *   object-aware cross attention (replaces ObjectAwareCrossAttention / QKVAttentionLegacy, layout_unet_v1.py:416-532,
 *   norm_first=False, scale 1.0):
 *     qkv fp32 [B,T,3C] (q|k|v of qkv_projector, head-major), pos_p fp32 [B,T,C] (normalised image-patch positional
 *     embedding), kl / pos_l / vl fp32 [B,L2,C] (layout key content, positional, value); per head (d = C/heads = 32):
 *     q = [q_c;pos_p], k = [[k_c;pos_p] | [k_l;pos_l]], v = [v_c | v_l]; scale2 = 1/sqrt(2d) (the reference multiplies q and
 *     k each by (2d)^-1/4)
 *   out: conv operand layout (parts 1 | 2 | 3) for an image [B][T/out_w][out_w][E] (token t = row t/out_w, col t%out_w;
 *   out_w % 128 == 0) = the A operand of the out-projection conv                                                      */

/* optional in-kernel profile of the tcgen05 attention kernel: 16 x u64 cycle counters per CTA (device buffer of
 * grid-size x 16 u64, or NULL to switch off); see csrc/attention.cu                                                 */
int b200_attn_set_debug(void* dbg_u64);

/* Product path: tcgen05.mma with the S and O accumulators in TMEM (csrc/attention.cu flash_attn_tc_kernel): a pack pass
 * writes the fp16 hi | lo Q / K / V^T tile images of every (batch, head) into `workspace` in the shared-memory layout of the
 * MMA operands, the main kernel stages them with one bulk copy per tile.
 *   workspace: b200_flash_attention_workspace(B, heads, T, Tx, dq, dv) bytes of device memory (Tx = extra layout keys, dq =
 *   query / key width per head (2 x 32 for the object-aware variant), dv = value width per head); contents are scratch.  */
size_t b200_flash_attention_workspace(int B, int heads, int T, int Tx, int dq, int dv);
int b200_flash_attention(const float* qkv, int E, void* out, int out_w, int parts, int B, int heads, int T,
                         float scale, void* workspace, void* stream);
int b200_flash_attention_oa(const float* qkv, const float* pos_p, const float* kl, const float* pos_l,
                            const float* vl, void* out, int out_w, int parts, int B, int C, int heads, int T,
                            int L2, float scale2, void* workspace, void* stream);

/* ---- K5: sampler update --------------------------------------------------------------------------
 * replaces p_step's ~25 elementwise ops (diffusion/continuous_time.py:205-231).
 *   coef fp32 [B,8]: {alpha_t, sigma_t, alpha_s, sigma_s, c1, c2, ddpm_c, 0}; mode 0 = ddim, 1 = ddpm;
 *   objective 0 = eps, 1 = v, 2 = x_0; clip <= 0 disables clamping; noise may be NULL when c1 == 0.  */
int b200_sampler_update(const float* x_t, const float* pred, const float* noise, const float* coef,
                        float* x_s, int B, int n_per_sample, int mode, int objective, float clip,
                        void* stream);

/* log-SNR schedule + step coefficients of a batch in one launch (continuous_time.py:14-63,200-231):
 *   step_t, step_s fp32 [B] in [0,1];  schedule 0 = linear, 1 = cosine, 2 = cosine_shifted, 3 = cosine_interpolated;
 *   t_min = atan(exp(-logsnr_max/2)), t_span = atan(exp(-logsnr_min/2)) - t_min, shift_* = 2 log(noise_d/image_d);
 *   log_snr_t fp32 [B] (the network's time condition), coef fp32 [B,8] as b200_sampler_update expects.      */
int b200_sampler_coefficients(const float* step_t, const float* step_s, int schedule, float t_min, float t_span,
                              float shift_lo, float shift_hi, float ddim_eta, float* log_snr_t, float* coef,
                              int B, void* stream);

/* ---- K6: point cloud -> range image ----------------------------------------------------------------
 * replaces load_points_as_images (dataset/transforms_3d/common.py:26-91, scan_unfolding=False).
 *   points fp32 [F,M,4] (x,y,z,intensity), npts int32 [F] valid points per frame
 *   out    fp32 [F,H,W,6] (x,y,z,i,depth,mask), nearest point per pixel wins (ties: highest index)
 *   grid   int32 [F,M,2] (grid_h, grid_w) for parity checks (may be NULL)
 *   zbuf   uint64 scratch [F,H,W]                                                                    */
int b200_range_project(const float* points, const int* npts, float* out, int* grid, void* zbuf, int F,
                       int M, int H, int W, float min_depth, float max_depth, float fov_up_deg,
                       float fov_down_deg, void* stream);

/* the same projection of a FLOAT64 point array (the temporal glue re-projects the ego-motion-warped background, a float64
 * array: tools/vis_tools/utils/pipe_related.py:244-255,271-280 -> custom_dataset.py:59-62; every intermediate of
 * common.py:38-91 is float64 there, only the final image is cast to float32).
 *   points fp64 [F,M,4], npts int32 [F] (may be NULL), out fp32 [F,H,W,6], zbuf uint64 scratch [F,H,W],
 *   winner int32 scratch [F,H,W] (index of the surviving point per pixel, -1: none)                                   */
int b200_range_project_f64(const double* points, const int* npts, float* out, void* zbuf, int* winner, int F, int M,
                           int H, int W, float min_depth, float max_depth, float fov_up_deg, float fov_down_deg,
                           void* stream);

/* ---- 3-D boxes -> 2-D boxes, condition mask, loss-weight map --------------------------------------------
 * replaces convert_boxes_to_2d + convert_points_to_2d (dataset/transforms_3d/common.py:99-216), called per frame by
 * NuscDataset.pre_process (dataset/nuscenes_dataset.py:389-398).
 *   boxes     [F,N,8] (x, y, z, l, w, h, yaw, class) fp32 (boxes_f64 = 0) or fp64 (1): the reference's dtype flow
 *             (corner offsets / cos, sin / centre depth in the array's dtype, projection in fp64) follows the input
 *   boxes_2d  fp64 [F,N,4] normalised (x1, y1, x2, y2);  mask fp32 [F,2,H,W] (class id, centre depth; later boxes
 *             overwrite earlier ones; a rectangle wider than 0.6 W wraps around the azimuth seam);
 *   weight    fp32 [F,H,W] = exp(sum_i inside_i (3 - area_i / max area)) or NULL
 *   workspace b200_boxes_to_mask_workspace(F, N) bytes of device memory                                            */
size_t b200_boxes_to_mask_workspace(int F, int N);
int b200_boxes_to_mask(const void* boxes, int boxes_f64, int F, int N, int H, int W, float fov_up_deg,
                       float fov_down_deg, double* boxes_2d, float* mask, float* weight, void* workspace,
                       void* stream);

/* ---- K7: points in boxes / voxel index ----------------------------------------------------------------
 * points_in_boxes_cpu semantics (ops/roiaware_pool3d/src/roiaware_pool3d.cpp:121-168, MARGIN 1e-2):
 *   pts [M,3], boxes [N,7] -> out int32 [N,M] in {0,1}                                                */
int b200_points_in_boxes(const float* pts, const float* boxes, int* out, int N, int M, void* stream);
/* points_in_boxes_gpu semantics (roiaware_pool3d_kernel.cu:313-336, MARGIN 1e-5): first box or -1     */
int b200_points_in_boxes_first(const float* pts, const float* boxes, int* out, int B, int N, int M,
                               void* stream);
/* generate_pts_mask_for_box3d (roiaware_pool3d_kernel.cu:39-75): -1 or x<<16|y<<8|z voxel code          */
int b200_voxel_index(const float* pts, const float* rois, int* out, int N, int M, int out_x, int out_y,
                     int out_z, void* stream);

/* ---- LiDARUtility (utils/lidar.py:34-132) --------------------------------------------------------- */
/* normalized depth [B,1,H,W] in [-1,1] (sampler output) -> metric depth [B,H,W] and xyz [B,3,H,W]     */
int b200_depth_to_xyz(const float* x_norm, const float* ray_angles, float* depth, float* xyz, int B,
                      int H, int W, float min_depth, float max_depth, void* stream);

/* ---- K8: evaluation-side projection + integer voxel quantiser (metrics/metric_utils.py) --------------
 * pcd2range (metric_utils.py:65-121): F clouds, pcd fp32 [F,M,3], npts int32 [F] (NULL: all M valid), optional per-point
 *   feature fp32 [F,M] (remission: feature_fill = -1; labels: 0).  Points with depth_min < ||xyz|| < depth_max (strict)
 *   are binned (fp32 arithmetic of NumPy >= 2 on float32 input; no +1e-6, no modulo, clamp to the image); the nearest
 *   point per pixel wins (equal depth: lowest index).  proj_range fp32 [F,H,W] (-1 where empty), proj_feature fp32
 *   [F,H,W] or NULL;  zbuf: uint64 scratch [F,H,W].                                                          */
int b200_pcd2range(const float* pcd, const int* npts, const float* feature, float* proj_range,
                   float* proj_feature, void* zbuf, int F, int M, int H, int W, float fov_up_deg,
                   float fov_down_deg, float depth_min, float depth_max, float feature_fill, void* stream);
/* range2xyz (metric_utils.py:124-154): range_img fp32 [F,H,W] -> xyz fp64 [F,3,H,W] (-1 outside the depth range);
 *   log_scale: depth = exp2(range * depth_scale) - 1 (fp32), else depth = range.                              */
int b200_range2xyz(const float* range_img, double* xyz, int F, int H, int W, float fov_up_deg,
                   float fov_down_deg, float depth_min, float depth_max, float depth_scale, int log_scale,
                   void* stream);
/* np.floor(coords / voxel_size).astype(np.int32) (sparse_quantize :51; pcd2bev_* / pcd2voxel_full :189,249):
 *   coords [M,stride] fp32 (coords_f64 = 0) or fp64 (1), first D (2|3) columns used;  div_f32 = 1: fp32 division by
 *   fp32(v) (python-float voxel size), 0: fp64 division (np.array voxel size).  voxel int32 [M,D];
 *   minmax int32 [6] (device): per-axis min in [0:3], max in [3:6] -- the caller reads it back to size the workspace.  */
int b200_quantize_coords(const void* coords, int coords_f64, int M, int D, int stride, double v0, double v1,
                         double v2, int div_f32, int32_t* voxel, int32_t* minmax, void* stream);
/* bytes of workspace b200_sparse_quantize needs for this bounding grid (minmax on the HOST); 0 = grid too large */
size_t b200_sparse_quantize_workspace(const int32_t* minmax_host, int D, int M);
/* ravel_hash (metric_utils.py:28-41): uint64 key of every voxel [M]                                          */
int b200_ravel_hash(const int32_t* voxel, int M, int D, const int32_t* minmax_host, uint64_t* out, void* stream);
/* np.unique(ravel_hash(voxel), return_index=True, return_inverse=True) (metric_utils.py:53-62), sort-free: bitmap of
 *   the bounding grid -> popcount scan -> rank.  uniq_coords int32 [M,D] (first n_unique rows valid, hash order),
 *   indices int64 [M] (first occurrence), inverse int64 [M], n_unique int32 [1] (device); any of the first three may
 *   be NULL.                                                                                                  */
int b200_sparse_quantize(const int32_t* voxel, int M, int D, const int32_t* minmax_host, void* workspace,
                         size_t workspace_bytes, int32_t* uniq_coords, int64_t* indices, int64_t* inverse,
                         int32_t* n_unique, void* stream);
/* pcd2bev_sum (metric_utils.py:231-256): clouds concatenated in pcd fp32 [total,stride] with offsets int32
 *   [n_clouds+1] (device); volume_sum fp32 [X,Y] += 1 per (cloud, occupied cell); bitmap_ws: n_clouds *
 *   ceil(X*Y/32) uint32 scratch.  Strict range test, cell = floor(v / fp32(voxel)) - min_bound.               */
int b200_bev_occupancy_sum(const float* pcd, const int* offsets, int n_clouds, int max_cloud_pts, int stride,
                           float x_lo, float x_hi, float y_lo, float y_hi, float voxel, int min_bx, int min_by,
                           int X, int Y, void* bitmap_ws, float* volume_sum, void* stream);
/* pcd2voxel_full (metric_utils.py:170-199): one cloud -> vol fp32 [X,Y,Z] of 0/1.  range_lo_hi_host = {x0,x1,y0,y1,
 *   z0,z1}, min_bound_host / dims_host int32 [3] (HOST pointers).                                            */
int b200_voxel_occupancy(const float* pcd, int M, int stride, const float* range_lo_hi_host, float voxel,
                         const int32_t* min_bound_host, const int32_t* dims_host, float* vol, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* B200LIDAR_H_ */
