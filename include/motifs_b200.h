/* motifs_b200.h — C ABI of libmotifs_b200.so (sm_100a kernels for the neural-motifs hot path).
 *
 * Every entry point takes plain pointers, sizes and a cudaStream_t; no torch types.
 * Device pointers are marked DEV, host pointers HOST.  Unless noted, the caller owns and
 * allocates every buffer (inputs, outputs and scratch), exactly as the reference's callers
 * do (SURVEY.md §8b "Ownership").
 *
 * Return convention (superset of the reference's): 1 = ok (what the reference launchers
 * return), 0 = bad argument (roi_align_cuda.c:19-22), negative = CUDA failure or
 * unsupported shape; mb200_last_error() returns the message.  Nothing calls exit() or
 * prints, unlike roi_align_kernel.cu:94-98 / highway_lstm_kernel.cu:17-29.
 *
 * PART 1 keeps the reference's own launcher names and argument order (drop-in symbols).
 * PART 2 are new entry points (prefix mb200_) the B200 host code uses.
 */
#ifndef MOTIFS_B200_H_
#define MOTIFS_B200_H_

#include <cuda_runtime_api.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

/* ------------------------------------------------------------------ PART 1: drop-in symbols */

/* replaces lib/fpn/roi_align/src/cuda/roi_align_kernel.h:11-15 (roi_align_kernel.cu:82-101).
 * image DEV [batch,depth,H,W] fp32; boxes DEV [num_boxes,5] = (batch idx, x1,y1,x2,y2) with the
 * corners already normalised by (W-1)/scale, (H-1)/scale (functions/roi_align.py:25-31);
 * crops DEV [num_boxes,depth,crop_h,crop_w]. Every element of crops is written. */
int ROIAlignForwardLaucher(const float* image_ptr, const float* boxes_ptr, int num_boxes, int batch,
                           int image_height, int image_width, int crop_height, int crop_width,
                           int depth, float extrapolation_value, float* crops_ptr, cudaStream_t stream);

/* replaces roi_align_kernel.h:21-23 (roi_align_kernel.cu:172-191). grads_image DEV
 * [batch,depth,H,W] must be zeroed by the caller; gradients are accumulated atomically. */
int ROIAlignBackwardLaucher(const float* grads_ptr, const float* boxes_ptr, int num_boxes, int batch,
                            int image_height, int image_width, int crop_height, int crop_width,
                            int depth, float* grads_image_ptr, cudaStream_t stream);

/* replaces lib/fpn/nms/src/cuda/nms_kernel.h:1-2 (nms_kernel.cu:88-131). keep_out HOST
 * [boxes_num] int32; boxes_dev DEV [boxes_num,4] sorted by descending score. Returns the
 * number of kept boxes (>= 0) or a negative error. Synchronous, like the reference. */
int ApplyNMSGPU(int* keep_out, const float* boxes_dev, const int boxes_num, float nms_overlap_thresh,
                int device_id);

/* replaces lib/lstm/highway_lstm_cuda/src/highway_lstm_kernel.h:7 (highway_lstm_kernel.cu:377-496).
 * Same argument order. lengths HOST [miniBatch] int32, sorted descending. `handle` is accepted
 * for ABI compatibility and ignored (no cuBLAS is used). Buffers as in
 * alternating_highway_lstm.py:71-107: x [T,B,In]; h_data,c_data [L,T+1,B,H] zero-initialised;
 * tmp_i [B,6H], tmp_h [B,5H] (unused scratch, may be NULL); T flat weights; bias [L,5H];
 * dropout [L,B,H]; gates [L,T,B,6H] or NULL when !is_training. Asynchronous on `stream`. */
void highway_lstm_forward_ongpu(int inputSize, int hiddenSize, int miniBatch, int numLayers, int seqLength,
                                float* x, int* lengths, float* h_data, float* c_data, float* tmp_i,
                                float* tmp_h, float* T, float* bias, float* dropout, float* gates,
                                int is_training, cudaStream_t stream, void* handle);

/* replaces highway_lstm_kernel.h:9 (highway_lstm_kernel.cu:162-375). Same argument order.
 * T_grad and bias_grad are ACCUMULATED into (caller pre-zeroes), as the reference's beta=1 GEMMs. */
void highway_lstm_backward_ongpu(int inputSize, int hiddenSize, int miniBatch, int numLayers, int seqLength,
                                 float* out_grad, int* lengths, float* h_data_grad, float* c_data_grad,
                                 float* x, float* h_data, float* c_data, float* T, float* gates_out,
                                 float* dropout_in, float* h_gates_grad, float* i_gates_grad,
                                 float* h_out_grad, float* x_grad, float* T_grad, float* bias_grad,
                                 int isTraining, int do_weight_grad, cudaStream_t stream, void* handle);

/* ------------------------------------------------------------------ PART 2: mb200_ entry points */

const char* mb200_last_error(void);
int mb200_abi_version(void);
int mb200_compiled_arch(void);
int mb200_device_ok(void);

/* RoIAlign, NHWC feature map in -> [num_boxes, crop_h*crop_w, depth] out (bin-major,
 * channel-minor). Same sampling as ROIAlignForwardLaucher. */
int mb200_roi_align_forward_nhwc(const float* image_nhwc, const float* boxes_ptr, int num_boxes, int batch,
                                 int image_height, int image_width, int crop_height, int crop_width,
                                 int depth, float extrapolation_value, float* crops_nhwc, cudaStream_t stream);

/* Same, but the output is the reference's [num_boxes, depth, crop_h, crop_w] layout (depth % 4 == 0). */
int mb200_roi_align_forward_nhwc_to_nchw(const float* image_nhwc, const float* boxes_ptr, int num_boxes, int batch,
                                         int image_height, int image_width, int crop_height, int crop_width,
                                         int depth, float extrapolation_value, float* crops_nchw, cudaStream_t stream);

/* Segmented greedy NMS entirely on the device (no D2H). See csrc/nms.cu. */
long long mb200_nms_mask_words(const int* seg_sizes_host, int num_segments);
int mb200_nms_segmented(const float* boxes_dev, const int* seg_off_dev, const long long* mask_off_dev,
                        int num_segments, int max_seg, float thresh, int max_keep,
                        unsigned long long* mask_dev, int* keep_dev, int* num_keep_dev, cudaStream_t stream);

/* replaces lib/fpn/box_utils.py:109-131 (fp32 IoU) on the device. out DEV [A,B]. */
int mb200_bbox_overlaps_f32(const float* boxes_a, int A, const float* boxes_b, int B, float* out,
                            cudaStream_t stream);
/* replaces lib/fpn/box_intersections_cpu/bbox.pyx:15-62 (mode 0) and :64-107 (mode 1), float64. */
int mb200_bbox_overlaps_f64(const double* boxes, int N, const double* query, int K, int mode, double* out,
                            cudaStream_t stream);
/* replaces lib/fpn/anchor_targets.py:50-67: float64 IoU of N anchors x G GT boxes, per-anchor max / first arg-max,
 * labels {-1, 0, 1} before subsampling. One warp per anchor, the IoU matrix is never materialised. All pointers DEV;
 * gt_max_ws: G x 8 bytes of scratch (zeroed inside). */
int mb200_anchor_targets(const double* anchors, int N, const double* gt_boxes, int G, double neg_thr, double pos_thr,
                         unsigned long long* gt_max_ws, double* max_overlaps, int* argmax, long long* labels,
                         cudaStream_t stream);
/* replaces the host loop of lib/lstm/decoder_rnn.py:230-247 (overlap-aware greedy label commitment in SGDet eval):
 * boxes DEV [N,C,4] (class-specific boxes of the detections), probs DEV [N,C] (softmax), thresh = nms_thresh (0.3);
 * commit DEV [N] int64. Returns MB200_ERR_UNSUPPORTED when N*C*4 bytes exceed 200 KB of shared memory. */
int mb200_decoder_commit(const float* boxes, const float* probs, int N, int C, float thresh, long long* commit,
                         cudaStream_t stream);
/* replaces lib/get_union_boxes.py:82-87 (union roi) and the pair gather of :47. */
int mb200_union_rois(const float* rois, const long long* pairs, int num_pairs, float* union_rois,
                     float* pair_boxes, cudaStream_t stream);
/* replaces lib/draw_rectangles/draw_rectangles.pyx:12-67; out DEV [N,2,P,P] = mask - offset. */
int mb200_draw_union_boxes(const float* pair_boxes, int num_pairs, int pooling_size, float offset, float* out,
                           cudaStream_t stream);
/* replaces lib/fpn/box_utils.py:28-48 (+ the clamps of object_detector.py:383-387). */
int mb200_bbox_preds(const float* boxes, const float* deltas, long long num_rows, int rows_per_box,
                     const float* im_hw, const int* im_idx, float* out, cudaStream_t stream);

/* Highway LSTM with DEVICE lengths and caller-provided scratch (what the torch host code calls;
 * the drop-in launchers above wrap these). See csrc/lstm.cu. proj_scratch / dG_scratch are
 * mb200_highway_lstm_scratch_floats() floats. */
size_t mb200_highway_lstm_scratch_floats(int hiddenSize, int miniBatch, int seqLength);
int mb200_highway_lstm_forward(int inputSize, int hiddenSize, int miniBatch, int numLayers, int seqLength,
                               const float* x, const int* lengths_dev, float* h_data, float* c_data,
                               const float* T, const float* bias, const float* dropout, float* gates,
                               float* proj_scratch, cudaStream_t stream);
int mb200_highway_lstm_backward(int inputSize, int hiddenSize, int miniBatch, int numLayers, int seqLength,
                                const float* out_grad, const int* lengths_dev, float* h_data_grad,
                                float* c_data_grad, const float* x, const float* h_data, const float* c_data,
                                const float* T, const float* gates_out, const float* dropout_in,
                                float* h_out_grad, float* x_grad, float* T_grad, float* bias_grad,
                                int do_weight_grad, float* dG_scratch, cudaStream_t stream);

/* One layer of the recurrence with the input projection P [T,B,6H] supplied by the caller (hoisted
 * tensor-core GEMM); gates may alias P. Backward fills dG [T,B,6H]; dX/dW/db are the caller's GEMMs. */
int mb200_highway_lstm_layer_forward(int hiddenSize, int miniBatch, int seqLength, int dir, const float* P,
                                     const float* Wh, const float* bias, const float* dropout, float* h, float* c,
                                     float* gates, const int* lengths_dev, cudaStream_t stream);
int mb200_highway_lstm_layer_backward(int hiddenSize, int miniBatch, int seqLength, int dir, const float* out_grad,
                                      const float* Wh, const float* h, const float* c, const float* gates,
                                      const float* dropout, float* h_grad, float* c_grad, float* dG,
                                      const int* lengths_dev, cudaStream_t stream);

/* Exact-fp32 SIMT GEMM, row-major, C = alpha*op(A)*op(B) + beta*C (replaces the cublasSgemm calls
 * of highway_lstm_kernel.cu:441-465 for small shapes; cross-check for the tcgen05 path). */
int mb200_sgemm(int transA, int transB, int M, int N, int K, float alpha, const float* A, int lda,
                const float* B, int ldb, float beta, float* C, int ldc, cudaStream_t stream);

/* ---- bf16x3 tensor-core path (csrc/gemm_tc.cu, csrc/convert.cu): tcgen05.mma + TMA.
 * Operands are (hi, lo) bf16 pairs, K contiguous, row pitch Kp % 64 == 0, zero padded. */

/* replaces cublasSgemm under nn.Linear (lib/object_detector.py:102-103, lib/rel_model.py:360-390):
 * C[M,N] = A[M,K] * B[N,K]^T (+ bias[N]) (ReLU). Outputs C fp32 (ldc) and/or (Chi,Clo) (ldsplit). */
long long mb200_gemm_workspace_floats(int M, int N, int Kp);
int mb200_gemm_bf16x3(const void* Ahi, const void* Alo, const void* Bhi, const void* Blo, int M, int N, int Kp,
                      const float* bias, int relu, float* C, long long ldc, void* Chi, void* Clo,
                      long long ldsplit, float* workspace, cudaStream_t stream);
/* replaces cuDNN under the 3x3/pad-1 VGG convolutions (lib/object_detector.py:110-127):
 * x NHWC pair [B,H,W,Cin], w [Cout, 9*Cin] pair in (kh,kw,cin) order, y NHWC fp32 and/or pair. */
int mb200_conv3x3_bf16x3(const void* xhi, const void* xlo, const void* whi, const void* wlo, int B, int H, int W,
                         int Cin, int Cout, const float* bias, int relu, float* y, void* yhi, void* ylo,
                         cudaStream_t stream);
int mb200_split_bf16(const float* src, long long rows, int cols, long long ld, int Kp, void* hi, void* lo,
                     cudaStream_t stream);
int mb200_split_transpose_bf16(const float* src, int rows, int cols, long long ld, int Rp, void* hi, void* lo,
                               cudaStream_t stream);
int mb200_conv_weight_split(const float* w_oihw, int O, int I, int Ip, void* hi, void* lo, cudaStream_t stream);
int mb200_im2col3_split(const float* x_nchw, int B, int H, int W, void* hi, void* lo, cudaStream_t stream);
int mb200_stem_weight_split(const float* w_oihw, int O, void* hi, void* lo, cudaStream_t stream);
/* VGG stem conv1_1 (3->64, 3x3, pad 1) + bias (+ReLU): fp32 NCHW image -> NHWC bf16 pair, exact fp32
 * direct convolution (replaces the first nn.Conv2d of lib/object_detector.py:110-127). Cout must be 64. */
int mb200_conv3x3_stem_split(const float* x_nchw, const float* w_oihw, const float* bias, int B, int H, int W,
                             int Cout, int relu, void* yhi, void* ylo, cudaStream_t stream);
int mb200_maxpool2_nhwc_split(const void* xhi, const void* xlo, int B, int H, int W, int C, void* yhi, void* ylo,
                              cudaStream_t stream);

/* 3x3 / stride 2 / pad 1 max-pool on [planes,H,W] fp32 (nn.MaxPool2d(3,2,1) of lib/get_union_boxes.py:34):
 * forward stores the arg-max (0..8) per output, backward is a deterministic gather. */
int mb200_maxpool3s2_forward(const float* x, long long planes, int H, int W, float* y, unsigned char* argmax,
                             cudaStream_t stream);
int mb200_maxpool3s2_backward(const float* grad_y, const unsigned char* argmax, long long planes, int H, int W,
                              float* grad_x, cudaStream_t stream);

/* ---- union-box mask branch (lib/get_union_boxes.py:28-37: conv7x7/s2 + ReLU + BN + maxpool3x3/s2 + conv3x3 +
 * ReLU + BN on the [R,2,27,27] masks of draw_union_boxes), NHWC streaming kernels around mb200_gemm_bf16x3. ---- */
/* im2col of the 7x7/s2/p3 stem over masks [R,2,S,S]: k = (ky*7+kx)*2 + c, 98 of 128 columns used.
 * transposed == 0: (hi, lo) [R*Ho*Ho, 128]; transposed == 1: [128, Pp] with Pp >= R*Ho*Ho, Pp % 2 == 0. */
int mb200_im2col7s2_split(const float* masks, int R, int S, int transposed, long long Pp, void* hi, void* lo,
                          cudaStream_t stream);
/* im2col of a 3x3/s1/p1 convolution over NHWC fp32 x [R,H,W,C]: k = (ky*3+kx)*C + c (the K order of
 * mb200_conv_weight_split). transposed == 0: [R*H*W, 9C]; 1: [9C, Pp], Pp % 64 == 0, C % 32 == 0. */
int mb200_im2col3_nhwc_split(const float* x, int R, int H, int W, int C, int transposed, long long Pp, void* hi,
                             void* lo, cudaStream_t stream);
/* nn.BatchNorm2d training statistics of x [P,C] (rows = N*H*W): two-pass mean / biased variance with double
 * accumulation -> mean, invstd = rsqrt(var + eps); running_* (nullable) updated with `momentum` and the
 * unbiased variance. sums: scratch of 4*C doubles. */
int mb200_bn_stats(const float* x, long long P, int C, float eps, float momentum, double* sums, float* mean,
                   float* invstd, float* running_mean, float* running_var, cudaStream_t stream);
/* y = maxpool3x3/s2/p1(BN(x)) on NHWC [R,H,W,C] -> [R,Ho,Wo,C], arg-max code (0..8, first maximum) per output. */
int mb200_bn_pool3s2_nhwc(const float* x, const float* mean, const float* invstd, const float* gamma, const float* beta,
                          int R, int H, int W, int C, float* y, unsigned char* argmax, cudaStream_t stream);
int mb200_unpool3s2_nhwc(const float* grad_y, const unsigned char* argmax, int R, int H, int W, int C, float* grad_x,
                         cudaStream_t stream);
/* out[R,C,HW] = BN(x[R,HW,C]) (+ addend[R,C,HW], nullable); HW <= 64. */
int mb200_bn_nhwc_to_nchw(const float* x, const float* mean, const float* invstd, const float* gamma, const float* beta,
                          const float* addend, int R, int HW, int C, float* out, cudaStream_t stream);
int mb200_nchw_to_nhwc(const float* x, int R, int C, int HW, float* out, cudaStream_t stream);
/* Backward of y = BN_train(x), x = ReLU(z): g = dL/dy [P,C]. sums[0:C] = sum g (dbeta), sums[C:2C] = sum g*xhat
 * (dgamma), dz = dL/dz [P,C], dbias[C] = column sums of dz (doubles; all three overwritten). */
int mb200_bn_relu_backward(const float* g, const float* x, const float* mean, const float* invstd, const float* gamma,
                           long long P, int C, double* sums, float* dz, double* dbias, cudaStream_t stream);
/* Fused form for the weight-gradient GEMMs: dz leaves as (hi, lo) bf16 pairs, transposed [C, Pp] (Pp % 64 == 0,
 * zero padded; C % 32 == 0) and, when p_hi != NULL, plain [P, C]. argmax != NULL: g is the pooled gradient
 * [R,Ho,Wo,C], routed through the 3x3/s2/p1 max-pool on the fly (P = R*H*W, x = the pool's input [R,H,W,C]). */
int mb200_bn_relu_backward_split(const float* g, const unsigned char* argmax, const float* x, const float* mean,
                                 const float* invstd, const float* gamma, long long P, long long Pp, int C, int H, int W,
                                 double* sums, void* t_hi, void* t_lo, void* p_hi, void* p_lo, double* dbias,
                                 cudaStream_t stream);
/* dx[R,H,W,C] from dcol [R*H*W, 9C] (adjoint of mb200_im2col3_nhwc_split, transposed == 0). */
int mb200_col2im3_nhwc(const float* dcol, int R, int H, int W, int C, float* dx, cudaStream_t stream);

/* EXPERIMENTAL (csrc/gemm_mn.cu; not yet run on a B200): the same bf16x3 product for operands whose reduction
 * dimension is the ROW index — C[M,N] (fp32, pitch ldc) = A^T B with A stored [K, M] (pitch lda) and B stored [K, N]
 * (pitch ldb), (hi, lo) bf16 pairs, M / N contiguous ("MN-major" UMMA operands): the weight-gradient products
 * dW = dY^T X of every nn.Linear on the path without the transposed operand copies. lda, ldb % 8 == 0. */
long long mb200_gemm_mn_workspace_floats(int M, int N, int K);
int mb200_gemm_bf16x3_mn(const void* Ahi, const void* Alo, long long lda, const void* Bhi, const void* Blo, long long ldb,
                         int M, int N, int K, float* C, long long ldc, float* workspace, cudaStream_t stream);

/* tcgen05 GEMM / conv kernel selection: 0 = 1-CTA kernels only, 1 = per-shape choice (default), 2 = the CTA-pair
 * (cta_group::2, 256-row tiles) kernel whenever the shape allows. Returns the previous mode. Tests and A/B runs. */
int mb200_gemm_set_pair_mode(int mode);
/* SMs the persistent tcgen05 kernels may occupy (clamped to [16, 148], even). Lowered by the host while a gradient
 * all-reduce is in flight (NCCL's channel CTAs hold SMs: a full-width persistent grid would run its last CTAs in a second
 * wave). Returns the previous value. */
int mb200_set_sm_budget(int n);
/* cudaMemsetAsync(ptr, 0, bytes, stream): zero-fill without occupying an SM. */
int mb200_zero_async(void* ptr, long long bytes, cudaStream_t stream);
/* 3x3 convolution kernel: 0 = tap-by-tap shifted TMA boxes, 1 = per layer (default), 2 = shared-memory halo staging always. */
int mb200_conv_set_halo_mode(int mode);

/* Highway-LSTM recurrence of ONE layer on the tensor cores (csrc/lstm_tc.cu; large batches — BASELINE configs[4]): same
 * contract as mb200_highway_lstm_layer_forward. Wt_hi / Wt_lo: W_h as bf16 pairs [H/16 * 80, H], K contiguous, row
 * (s*80 + g*16 + u) = column (g*H + 16 s + u) of W_h [H,5H]; hb_hi / hb_lo: [T+1,B,H] bf16 scratch, ZERO on entry (the
 * bf16 pair of the hidden state). mb200_highway_lstm_tc_supported: H % 64 == 0, weight slice fits shared memory. */
int mb200_highway_lstm_tc_supported(int hiddenSize, int miniBatch);
int mb200_highway_lstm_layer_forward_tc(int hiddenSize, int miniBatch, int seqLength, int dir, const float* P,
                                        const void* Wt_hi, const void* Wt_lo, const float* bias, const float* dropout,
                                        float* h, float* c, void* hb_hi, void* hb_lo, float* gates,
                                        const int* lengths_dev, cudaStream_t stream);

/* *acc += sum_i x[i]^2 (double accumulator on the device, caller zeroes it): the global gradient norm of
 * clip_grad_norm (lib/pytorch_misc.py:416-459) as one pass per flat buffer. x 16-byte aligned. */
int mb200_sumsq_accum(const float* x, long long n, double* acc, cudaStream_t stream);

/* Fused clip + weight-decay + momentum SGD over a flat fp32 buffer (replaces the caller-side
 * clip_grad_norm + optim.SGD.step of models/train_rels.py:145-150). total_norm_dev: device scalar with
 * the global gradient norm, or NULL for no clipping. Pointers 16-byte aligned. */
int mb200_sgd_momentum_clip(float* params, float* grads, float* momentum_buf, long long n, float lr, float momentum,
                            float weight_decay, const float* total_norm_dev, float max_norm, int first_step,
                            int zero_grad, cudaStream_t stream);
/* Same, with `grads` holding grad_scale^-1 times the gradient (data parallel: the all-reduced SUM and
 * grad_scale = 1/world, which saves the separate averaging pass over the 1.1 GB buffer); *total_norm_dev is the
 * norm of the SCALED gradient. */
int mb200_sgd_momentum_clip_scaled(float* params, float* grads, float* momentum_buf, long long n, float lr, float momentum,
                                   float weight_decay, const float* total_norm_dev, float max_norm, float grad_scale,
                                   int first_step, int zero_grad, cudaStream_t stream);
/* Same, and the bf16 (hi, lo) pair of every UPDATED parameter is written to hi[i], lo[i] (same flat index; 8-byte aligned):
 * for weight matrices whose row length is a multiple of 64 these ARE the K-major GEMM operands of mb200_gemm_bf16x3. */
int mb200_sgd_momentum_clip_split(float* params, float* grads, float* momentum_buf, void* hi, void* lo, long long n, float lr,
                                  float momentum, float weight_decay, const float* total_norm_dev, float max_norm,
                                  float grad_scale, int first_step, int zero_grad, cudaStream_t stream);

/* ---- Data-parallel sharded update over NVSwitch multicast (lib/fused_optim.py, comm="nvls"; replaces the reference's
 * replicate / parallel_apply / Gather of lib/rel_model.py:549-560 together with the gradient exchange). `*_mc` arguments are
 * MULTICAST virtual addresses of symmetric buffers (one physical copy per rank, same offset everywhere), obtained by the host
 * from torch's symmetric-memory rendezvous; everything else is an ordinary local device pointer.
 *   pass 1  mb200_dp_reduce_shard_sumsq: out[i] = SUM over ranks of the gradient at x_mc[i] (multimem.ld_reduce: the sum is
 *           formed inside the switch), *acc += sum of out[i]^2                                   -> reduce-scatter + norm
 *           mb200_dp_bcast_slot: *v -> slots[idx] on every rank                                  -> the ranks' partial norms
 *   pass 2  mb200_sgd_momentum_clip_mc: the fused clip + SGD update of this rank's shard, the new parameters (and their bf16
 *           operand pairs when hi_mc / lo_mc are non-null) stored with multimem.st into EVERY rank's copy   -> all-gather
 * n % 4 == 0, 16-byte aligned (8 for hi / lo). The caller orders the passes with cross-rank barriers.
 * mb200_optim_set_background(1): sumsq / sgd / the two passes launch as ONE 128-thread CTA per SM (<= 80 registers), which
 * fits beside a resident tcgen05 GEMM CTA — for updates deferred underneath the next step's backbone. */
/* comm="ce": the gradient shards travel by COPY ENGINE over NVLink (peer-mapped symmetric memory, no SM involved);
 * the only kernels are local and shard-sized: g[i] += sum of the nslots staged copies (stage + k * stride), *acc += |g|^2,
 * then the ordinary mb200_sgd_momentum_clip_split on the shard; the updated shard is copied back out by the copy engines. */
int mb200_dp_reduce_staged_sumsq(float* g, const float* stage, long long stride, int nslots, long long n, double* acc,
                                 cudaStream_t stream);
int mb200_optim_set_background(int on);
int mb200_dp_reduce_shard_sumsq(const float* x_mc, float* out, long long n, double* acc, cudaStream_t stream);
int mb200_dp_bcast_slot(const double* v, double* slots_mc, int idx, cudaStream_t stream);
int mb200_sgd_momentum_clip_mc(float* params, float* grads, float* momentum_buf, float* p_mc, void* hi_mc, void* lo_mc,
                               long long n, float lr, float momentum, float weight_decay, const float* total_norm_dev,
                               float max_norm, float grad_scale, int first_step, int zero_grad, cudaStream_t stream);

#ifdef __cplusplus
}
#endif
#endif /* MOTIFS_B200_H_ */
