/* shf_b200 -- C ABI of the B200-native smallhardface inference hot path (libshf_b200.so).
 *
 * Every entry point takes plain pointers and sizes.  Device pointers are marked [dev]; `stream` is a
 * cudaStream_t passed as void* (NULL = default stream).  All functions except `_nms` return 0 on
 * success, a negative code otherwise, with a human-readable message from shf_last_error().
 * Launches are asynchronous on `stream`; nothing synchronises unless stated.
 *
 * Activation tensors are NHWC, 4 bytes per element in two equally sized planes; every function that touches
 * one takes the format of what it reads / writes:
 *   SHF_FMT_H2  (0, "h2", precise): planes [2][N][H][W][C] of __half with x = hi + lo (hi = rn(x), lo = rn(x - hi));
 *               C a multiple of 8 (16-byte rows for TMA).  The conv runs 3 fp16 MMAs per 16 input channels.
 *   SHF_FMT_HF8 (1, "hf8", fast): plane 0 = hi as above; plane 1 = per pixel and 64-channel block, 64 bytes
 *               e4m3((x - hi) * 2^6) then 64 bytes e4m3(hi * 2^-5), both saturating; C (and channel windows) multiples
 *               of 64.  The conv runs 1 fp16 + 1 fp8 (K = 32) MMA per 16 input channels; x = hi + al8 * 2^-6 is carried
 *               to ~2^-16 relative for 2 <~ |x| < 16384 (fewer residual bits below, saturation above).
 * Each declaration cites the reference interface it stands in for (paths under the reference repo).
 */
#ifndef SHF_B200_H
#define SHF_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define SHF_FMT_H2 0
#define SHF_FMT_HF8 1

const char* shf_last_error(void);
int shf_abi_version(void);
int shf_device_info(int device, int* sm_count, int* cc_major, int* cc_minor, long long* total_mem);

/* ---- Net layers (caffe/src/caffe/layers/*) -------------------------------------------------- */

/* ConvolutionLayer<float>::Forward (conv_layer.cpp:30-46 -> base_conv_layer.cpp:255-279) followed by the
 * in-place ReLULayer (relu_layer.cpp:9-19), for 3x3 (pad == dilation, stride 1) and 1x1 convolutions with
 * Cin, Cout multiples of 64.  tcgen05 implicit GEMM.  w_h2 [dev]: weights pre-packed as
 * [2 planes][taps][Cout][Cin] fp16 of (w * 2^k) (in_format h2: hi / lo planes; in_format hf8: plane 1 holds, per
 * (tap, Cout, 64-channel block), 64 bytes e4m3(hi * 2^-6) then 64 bytes e4m3(lo * 2^5)); out_scale = 2^-k.
 * in_format describes in_h2 AND w_h2, out_format what is written.  The result lands in channels
 * [out_channel_offset, +cout) of an h2 tensor with out_channels_total channels (ConcatLayer by construction,
 * concat_layer.cpp:47-74).
 * range_guard [dev, may be NULL]: one 32-bit slot that receives (atomic max) the float bits of max |x| over the values
 * this launch writes.  The fixed exponent windows of the hf8 format and the fp16 range of either format make that the
 * quantity a caller must watch on weights it has not seen: >= 65504 overflows hi, >= 14336 saturates the hf8 ah8 bytes,
 * a small maximum (< ~32) leaves the hf8 residual bytes fewer than their 4 bits.  shf_conv1_tc and
 * shf_deconv_depthwise take the same argument. */
int shf_conv_igemm(const void* in_h2, const void* w_h2, const float* bias, void* out_h2, int batch, int H, int W,
                   int cin, int cout, int ksize, int dilation, int out_channels_total, int out_channel_offset,
                   float out_scale, int relu, int in_format, int out_format, unsigned int* range_guard, void* stream);

/* shf_conv_igemm + the PoolingLayer MAX 2x2/2 that follows it (pooling_layer.cpp:140-187) in one launch: the pooled
 * map goes to pool_out_h2 (N, H/2, W/2, pool_channels_total) at pool_channel_offset; out_h2 may be NULL when the
 * un-pooled blob has no other consumer.  H and W must be even. */
int shf_conv_igemm_pool(const void* in_h2, const void* w_h2, const float* bias, void* out_h2, void* pool_out_h2, int batch,
                        int H, int W, int cin, int cout, int ksize, int dilation, int out_channels_total,
                        int out_channel_offset, int pool_channels_total, int pool_channel_offset, float out_scale,
                        int relu, int in_format, int out_format, unsigned int* range_guard, void* stream);

/* shf_conv_igemm for a 1x1 kernel with spatial stride (pad 0): the projection shortcuts and downsampling 1x1 convolutions
 * of a ResNet bottleneck (conv_layer.cpp:8-28 output size (H - 1) / stride + 1).  H x W are the INPUT dims.  The same
 * tcgen05 kernel reads the activations through a strided TMA view (every stride-th pixel), no gather kernel. */
int shf_conv_igemm_strided(const void* in_h2, const void* w_h2, const float* bias, void* out_h2, int batch, int H, int W,
                           int stride, int cin, int cout, int out_channels_total, int out_channel_offset, float out_scale,
                           int relu, int in_format, int out_format, unsigned int* range_guard, void* stream);

/* shf_conv_igemm with the residual add of a ResNet block fused into the epilogue (EltwiseLayer SUM with unit coefficients,
 * eltwise_layer.cpp:52-57, then the block's in-place ReLU): out = [relu](conv(in) * out_scale + bias + residual).  residual
 * [dev]: an activation tensor of the same N x H x W with res_channels_total channels, read at res_channel_offset, in
 * res_format.  The branch output is never written to HBM. */
int shf_conv_igemm_res(const void* in_h2, const void* w_h2, const float* bias, const void* residual, void* out_h2, int batch,
                       int H, int W, int cin, int cout, int ksize, int dilation, int out_channels_total, int out_channel_offset,
                       int res_channels_total, int res_channel_offset, float out_scale, int relu, int in_format, int res_format,
                       int out_format, unsigned int* range_guard, void* stream);

/* 3x3 convolution with stride 2 and pad 1 (stage transitions of torchvision-style ResNets and similar backbones): H x W are
 * the INPUT dims (both >= 2), the output is ((H - 1) / 2 + 1) x ((W - 1) / 2 + 1).  w_h2 as for a 3x3 shf_conv_igemm.  The
 * tcgen05 kernel walks (tap, 64-channel chunk) pairs and reads every tap from the parity view of the input it lives in
 * (four strided TMA maps; TMA zero fill outside a view is the zero padding): no gather pass, no MMAs on discarded pixels. */
int shf_conv3x3_s2(const void* in_h2, const void* w_h2, const float* bias, void* out_h2, int batch, int H, int W, int cin,
                   int cout, int out_channels_total, int out_channel_offset, float out_scale, int relu, int in_format,
                   int out_format, unsigned int* range_guard, void* stream);

/* ---- layers a ResNet-style backbone adds (BASELINE north_star names "the ResNet/VGG backbone"; csrc/resnet_kernels.cu) -- */
/* EltwiseLayer SUM (eltwise_layer.cpp:37-77): out = sum_t coeffs[t] * ins[t] over n_in (<= 4) activation tensors of
 * `pixels` x C [dev; ins is a HOST array of device pointers], coeffs [host, may be NULL = all 1], optionally followed by the
 * in-place ReLU of a residual block; written at a channel offset like shf_conv_igemm. */
int shf_eltwise_sum(const void* const* ins, const float* coeffs, int n_in, void* out, long long pixels, int C,
                    int out_channels_total, int out_channel_offset, int relu, int in_format, int out_format,
                    unsigned int* range_guard, void* stream);

/* PoolingLayer MAX with any kernel / stride / pad (pooling_layer.cpp:79-123,140-187: ceil-mode output size, windows clipped
 * to the image).  shf_pool_out_size is the layer's output-size rule for one dimension (any_pad = pad_h || pad_w). */
int shf_pool_out_size(int size, int k, int stride, int pad, int any_pad);
int shf_maxpool(const void* in_h2, void* out_h2, int batch, int H, int W, int C, int kh, int kw, int sh, int sw, int ph, int pw,
                int format, void* stream);

/* First convolution of a net whose input has 3 channels, square kernel <= 11, stride <= 4 (ResNet conv1: 7x7 stride 2 pad 3):
 * fp32 NCHW (N,3,H,W) [dev] -> activation tensor (N,HO,WO,64); weights OIHW fp32 [dev] (64,3,k,k) (BatchNorm / Scale already
 * folded in by the caller), fp32 FMAs.  The VGG16 conv1_1 (3x3 stride 1) stays on shf_conv1_tc. */
int shf_conv_first(const float* in_nchw, const float* w_oihw, const float* bias, void* out_act, int batch, int H, int W,
                   int cout, int ksize, int stride, int pad, int relu, int out_format, unsigned int* range_guard, void* stream);

/* First convolutions (3 -> 64 channels) on the tensor cores: ksize 7 / stride 2 (ResNet conv1) and ksize 3 / stride 1 (the
 * VGG16 conv1_1; shf_conv1_tc is its one-thread-per-pixel predecessor and regression twin).  The 3 k^2 taps of an output pixel
 * form one split-fp16 K-major operand row [hi(Kp) | lo(Kp)], Kp = 3 k^2 rounded up to 16, in 128-byte-swizzled column blocks;
 * 3 Kp / 16 tcgen05 MMAs per 128-pixel tile; two threads per pixel.  w_packed [dev]: fp16 [2][64][64 * WB] of (w * 2^e):
 * [0] = hi(k), [1] = lo(k), k = (c*ks + r)*ks + s zero-padded to WB whole 64-half blocks (WB = 3 / 1); out_scale = 2^-e.
 * Output ((H + 2 pad - ks) / stride + 1) x ((W + 2 pad - ks) / stride + 1).  shf_conv_first is the fp32 SIMT twin. */
int shf_conv_first_tc(const float* in_nchw, const void* w_packed, const float* bias, void* out_act, int batch, int H, int W,
                      int cout, int ksize, int stride, int pad, float out_scale, int relu, int out_format,
                      unsigned int* range_guard, void* stream);

/* Test hook: 8 = the conv kernel on CTA pairs (tcgen05 cta_group::2, the product path and the default), 7 = the same
 * persistent streaming-drain kernel with one CTA per tile (regression twin of the pair protocol; same results). */
int shf_set_conv_impl(int impl);

/* conv1_1: fp32 NCHW (N,3,H,W) [dev] -> h2 (N,H,W,64); weights OIHW fp32 [dev] (64,3,3,3), pad 1. */
int shf_conv1_c3(const float* in_nchw, const float* w_oihw, const float* bias, void* out_h2, int batch, int H, int W,
                 int cout, int relu, int out_format, void* stream);

/* conv1_1 on the tensor cores (the product path; shf_conv1_c3 is the fp32 SIMT twin kept for validation): the 27 taps of
 * a pixel form one split-fp16 K-major operand row built in shared memory, six tcgen05 MMAs per 128-pixel tile.
 * w_packed [dev]: fp16 [2][64][64] of (w * 2^k): [0] = rows [hi(32) | hi(32)], [1] = rows [lo(32) | 0], K index
 * c*9 + r*3 + s padded to 32; out_scale = 2^-k. */
int shf_conv1_tc(const float* in_nchw, const void* w_packed, const float* bias, void* out_act, int batch, int H, int W,
                 int cout, float out_scale, int relu, int out_format, unsigned int* range_guard, void* stream);

/* PoolingLayer MAX 2x2 stride 2 (pooling_layer.cpp:79-123,140-187), ceil-mode output (H+1)/2 x (W+1)/2. */
int shf_maxpool2x2(const void* in_h2, void* out_h2, int batch, int H, int W, int C, int format, void* stream);

/* DeconvolutionLayer with group == channels (deconv_layer.cpp:8-46; the net's conv5_256_up k4 s2 p1).
 * w [dev]: fp32 (C,1,k,k). Output size stride*(H-1)+k-2*pad, written at a channel offset like shf_conv_igemm. */
int shf_deconv_depthwise(const void* in_h2, const float* w, void* out_h2, int batch, int H, int W, int C, int ksize,
                         int stride, int pad, int out_channels_total, int out_channel_offset, int in_format,
                         int out_format, unsigned int* range_guard, void* stream);

/* Blob boundary (caffe/python/caffe/_caffe.cpp:205-242 exposes blobs as fp32 NCHW arrays). */
int shf_h2_to_nchw(const void* in_h2, float* out_nchw, int batch, int H, int W, int channels_total, int channel_offset,
                   int channels, int format, void* stream);
int shf_nchw_to_h2(const float* in_nchw, void* out_h2, int batch, int channels, int H, int W, int channels_total,
                   int channel_offset, int format, void* stream);

/* ---- pre-processing (lib/utils/test_utils.py:29-46, lib/utils/blob.py:16-32, lib/test.py:30-38,147-155) -- */
/* uint8 HWC BGR image [dev] -> one pyramid level: mean-subtract, bilinear resize by `scale` (cv2 fx=fy
 * semantics, output out_h x out_w = rint(h*scale) x rint(w*scale)), optional mirror, zero pad to
 * padded_h x padded_w, fp32 CHW [dev].  means: 3 host doubles (cfg.PIXEL_MEANS). */
int shf_preprocess_level(const uint8_t* img_hwc, int h, int w, float* out_chw, int out_h, int out_w, int padded_h,
                         int padded_w, double scale, int flip, const double* means, void* stream);

/* The same for num_images stacked same-sized images (N, h, w, 3) and passes in {1, 2} (plain [+ mirrored]) in ONE launch:
 * out_nchw is the (num_images * passes, 3, padded_h, padded_w) level batch, slot j * passes + f. */
int shf_preprocess_level_batched(const uint8_t* imgs_nhwc, int num_images, int h, int w, float* out_nchw, int out_h,
                                 int out_w, int padded_h, int padded_w, double scale, int passes, const double* means,
                                 void* stream);

/* ---- detection tail -------------------------------------------------------------------------- */
/* cls_score*/bbox_pred* 1x1 convs + Concat/Reshape + SoftmaxLayer (softmax_layer.cpp:27-60) + the decode half of
 * ProposalLayer.forward (lib/layers/proposal_layer.py:96-173, lib/utils/bbox_transform.py:33-93).
 * feat_h2: host array of num_anchors [dev] pointers to ONE image's hi-plane rows (H,W,C) inside an h2-FORMAT tensor
 * (the head convs write h2 whatever format they read);
 * feat_plane_stride = elements between its hi and lo planes (N*H*W*C for a batch of N; 0 means H*W*C).
 * w_cls [A][2][C], b_cls [A][2], w_box [A][4][C], b_box [A][4], all fp32 [dev]; base_anchors: host [A][4].
 * Outputs [dev]: prob (2A,H,W) fp32 = cls_prob_reshape_output; delta (4A,H,W) = bbox_pred_output;
 * boxes (H*W*A,4) decoded+clipped, rows ordered (h,w,a); keys (H*W*A) u64 sort keys
 * (desc(score)<<32 | row with desc() = complement of the order-preserving integer image of a float, so ascending keys =
 * descending scores for any float; ~0 for rows below score_thresh / min_size); count = rows >= score_thresh;
 * best_key = smallest key over rows passing min_size. */
int shf_head_decode(const void* const* feat_h2, long long feat_plane_stride, int num_anchors, const float* w_cls,
                    const float* b_cls, const float* w_box, const float* b_box, const float* base_anchors, int H, int W,
                    int C, int feat_stride, float im_h, float im_w, float min_size, float score_thresh, float* prob,
                    float* delta, float* boxes, unsigned long long* keys, int* count, unsigned long long* best_key,
                    void* stream);

/* The same for all images of one pyramid-level batch in ONE launch (the pyramid driver's form).  feat_h2 point at
 * image 0; image i starts feat_image_stride elements later.  Outputs are [image][...] slices of the single-image
 * layout; keys are tagged [image:5][desc(score):32][row:27]; shf_sort_keys with one segment per image slot orders them. */
int shf_head_decode_batched(const void* const* feat_h2, long long feat_plane_stride, long long feat_image_stride,
                            int num_images, int num_anchors, const float* w_cls, const float* b_cls, const float* w_box,
                            const float* b_box, const float* base_anchors, int H, int W, int C, int feat_stride,
                            float im_h, float im_w, float min_size, float score_thresh, float* prob, float* delta,
                            float* boxes, unsigned long long* keys, int* count, unsigned long long* best_key,
                            void* stream);

/* lib/test.py:52-66,141-167 for one level of a batch: per image, the plain and the mirrored pass (slots 2j, 2j+1 of the
 * batched decode) are un-mirrored, divided by im_scale, thresholded (fg > det_thresh) and appended in that order to
 * dets[image_base + j] at pass_offsets[image][pass_base], updating pass_offsets[image][pass_base + 1 (+2)]. */
int shf_gather_dets_batched(const unsigned long long* sorted_keys, const int* count, const unsigned long long* best_key,
                            const float* prob, const float* boxes, int num_anchors, int hw, int topn, int num_images,
                            int passes_per_image, float* dets, int* pass_offsets, int image_base, int passes_total,
                            int pass_base, int det_cap, float level_w, float im_scale, float det_thresh, void* stream);

/* `max_score.argsort()[::-1]` (proposal_layer.py:181; also lib/test.py:182, cpu_nms.pyx:25) as an ascending, STABLE,
 * segmented radix sort of the keys above, hand-written (one CTA per segment, one launch): segment s occupies
 * keys[s * segment_stride, + len), len = min(segment_len[s], fixed_len) (segment_len [dev] may be NULL: fixed_len).
 * Keys whose 32-bit score field (bits [begin_bit, begin_bit + 32)) is all ones are sentinels ("not a candidate") and
 * are dropped; the survivors are ordered by that field alone and ties keep their arrival order -- the keys arrive in
 * ascending row order, so ties resolve to the lower row, the order the oracle defines.  begin_bit = 32 for
 * shf_head_decode keys, 27 for the image-tagged batched keys (segment = image slot).  keys_out[s * stride, + survivors)
 * holds the result (the rest of the segment is ~0); keys_in is left untouched.  workspace: one more key buffer. */
long long shf_sort_keys_workspace(int total_keys);
int shf_sort_keys(const unsigned long long* keys_in, unsigned long long* keys_out, int num_segments, int segment_stride,
                  const int* segment_len, int fixed_len, int begin_bit, void* workspace, long long workspace_bytes,
                  void* stream);

/* proposal_layer.py:183-220: R = min(count, topn) rows (1 if nothing cleared the threshold) ->
 * out_boxes (topn,5) [0,x1,y1,x2,y2], out_probs (topn,2) [bg,fg], *out_rows = R  [all dev].
 * If dets != NULL also performs lib/test.py:52-66,163-167 for this pass: un-mirror when flip
 * (x1' = level_w - x2), divide by im_scale, keep fg > det_thresh, append to dets (det_cap,5) at
 * pass_offsets[pass] and write pass_offsets[pass+1]. */
int shf_proposal_gather(const unsigned long long* sorted_keys, const int* count, const unsigned long long* best_key,
                        const float* prob, const float* boxes, int num_anchors, int hw, int topn, float* out_boxes,
                        float* out_probs, int* out_rows, float* dets, int* pass_offsets, int pass, int det_cap,
                        int flip, float level_w, float im_scale, float det_thresh, void* stream);

/* Batched per-image post-processing over dets (rows [seg_begin[i], seg_end[i]) per image, all [dev]):
 *   method 0: greedy NMS -> out_idx[i][out_cap] kept row indices relative to seg_begin[i], descending score
 *             (lib/nms/cpu_nms.pyx:17-68 for mode 0 `(double)ovr >= thresh`; lib/nms/nms_kernel.cu:45-155 for
 *              mode 1 `ovr > (float)thresh`; mode 2 `ovr >= (float)thresh`)
 *   method 1: bbox_vote (lib/test.py:181-217) -> out_dets[i][out_cap][5], singleton clusters dropped.
 * out_count[i] = rows PRODUCED, which may exceed out_cap: only the first out_cap are stored and the caller must treat
 * out_count[i] > out_cap as an error (box voting cannot produce more than cap_per_image / 2 + 1 rows, NMS no more than
 * cap_per_image).  cap_per_image bounds rows considered per image.  Images with up to 16384 rows use IoU bit masks
 * (parallel 64 x 64 tiles + a 64-rows-per-step sweep), larger ones a serial one-CTA sweep; same results. */
long long shf_postprocess_workspace(int num_images, int cap_per_image);
int shf_postprocess(const float* dets, const int* seg_begin, const int* seg_end, int num_images, int cap_per_image,
                    double thresh, int method, int mode, int* out_idx, float* out_dets, int* out_count, int out_cap,
                    void* workspace, long long workspace_bytes, void* stream);

/* lib/nms/gpu_nms.hpp:1-2 -- the reference's own symbol: HOST pointers, boxes pre-sorted by descending score,
 * keep_out caller-allocated (boxes_num ints), synchronous, selects device_id.  `IoU > thresh` (nms_kernel.cu:82).
 * On a CUDA error *num_out = -1 (the reference prints and continues). */
void _nms(int* keep_out, int* num_out, const float* boxes_host, int boxes_num, int boxes_dim, float nms_overlap_thresh,
          int device_id);
int shf_nms_host(int* keep_out, int* num_out, const float* boxes_host, int boxes_num, int boxes_dim, double thresh,
                 int mode, int device_id);
/* lib/test.py:181-217 `bbox_vote(det)` with the same host-buffer contract: dets_host (num_dets x 5 float32, any order) ->
 * out_dets (caller-allocated, >= num_dets / 2 + 1 rows of 5 floats), *num_out rows; synchronous. */
int shf_bbox_vote_host(float* out_dets, int* num_out, const float* dets_host, int num_dets, double thresh, int device_id);

/* lib/utils/bbox.pyx: kind 0 bbox_overlaps (:14-54), 1 bbox_overlaps_IoA (:56-102), 2 bbox_overlaps_itself
 * (:106-142).  boxes (n,4), query (k,4), out (n,k): float64 [dev]. */
int shf_bbox_overlaps(const double* boxes, const double* query, int n, int k, int kind, double* out, void* stream);

/* Validation only (not on the product path): direct fp32 convolution on h2 tensors, OIHW fp32 weights. */
int shf_debug_conv_direct(const void* in_h2, const float* w_oihw, const float* bias, void* out_h2, int batch, int H,
                          int W, int cin, int cout, int ksize, int dilation, int pad, int out_channels_total,
                          int out_channel_offset, int relu, int in_format, int out_format, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* SHF_B200_H */
