/*
 * monocon_b200.h -- C ABI of the B200-native MonoCon forward + decode engine.
 *
 * The reference (2gunsu/monocon-pytorch) has no native / FFI layer: its boundary for this
 * path is the Python nn.Module surface of model/detector/monocon_detector.py.  This library
 * is what a thin Python module binds (ctypes) to replace the bodies of
 *
 *   MonoConDetector._extract_feat_from_data_dict   monocon_detector.py:85-87   (backbone + neck)
 *   MonoConDenseHeads._get_predictions             monocon_heads.py:165-200    (heads)
 *   MonoConDenseHeads.decode_heatmap/_get_bboxes   monocon_heads.py:313-329,399-482 (decode)
 *
 * Conventions: every entry point returns an int status (0 = ok, non-zero = error, message via
 * mc_last_error); nothing throws across the ABI; all device work is enqueued on the stream that
 * is passed in (a cudaStream_t passed as void*), with no hidden synchronisation unless stated;
 * outputs are written into caller-provided buffers; the library owns only its packed weights
 * and its activation arena.  A handle is NOT thread-safe: one handle per (device, stream).
 * All file:line citations are relative to the reference repository root.
 */
#ifndef MONOCON_B200_H_
#define MONOCON_B200_H_

#include <stddef.h>
#include <stdint.h>

#if defined(__GNUC__)
#define MC_API __attribute__((visibility("default")))
#else
#define MC_API
#endif

#ifdef __cplusplus
extern "C" {
#endif

typedef struct mc_handle mc_handle;

/* precision_mode of mc_create */
#define MC_PREC_BF16 0 /* bf16 storage, tcgen05 tensor-core convolutions, fp32 accumulate (throughput mode)   */
#define MC_PREC_FP32 1 /* fp32 storage and fp32 FFMA convolutions (the reference's TF32-off arithmetic,        */
                       /* test.py:30-33); the training engine, and the slow twin of MC_PREC_FP32_TC            */
#define MC_PREC_FP32_TC 2 /* fp32-ACCURATE results on the tensor cores (inference; the 1e-3 / identical-top-k   */
                       /* parity gate runs in this mode): every activation and weight is held as two fp16      */
                       /* pieces hi + lo (~22 significant bits, per-tensor / per-filter power-of-two scales),   */
                       /* every K-block is three tcgen05 MMAs (hi*w_lo, lo*w_hi, hi*w_hi) into one fp32 TMEM    */
                       /* accumulator, the stems / AttnBN / 1x1 heads / decode stay fp32.  mc_calibrate_scales  */
                       /* fits the per-tensor scales to a sample batch; mc_scale_status reports range use.     */

/* conv implementation override of mc_set_option("conv_impl", ...) */
#define MC_CONV_AUTO 0 /* tcgen05 in MC_PREC_BF16 / MC_PREC_FP32_TC, FFMA in MC_PREC_FP32 */
#define MC_CONV_SIMT 1 /* force the FFMA kernels (MC_PREC_BF16 / MC_PREC_FP32 storage) */

/* Order of the ten prediction maps in pred_out[] / pred[] -- the keys of the dict returned by
 * MonoConDenseHeads._get_predictions (monocon_heads.py:190-200), NCHW fp32, channels:
 *   0 center_heatmap_pred 3 | 1 kpt_heatmap_pred 9 | 2 wh_pred 2 | 3 offset_pred 2 |
 *   4 kpt_heatmap_offset_pred 2 | 5 center2kpt_offset_pred 18 | 6 dim_pred 3 | 7 depth_pred 2 |
 *   8 alpha_cls_pred 12 | 9 alpha_offset_pred 12                                              */
#define MC_NUM_PRED 10

/* neck_variant of mc_create_ex */
#define MC_NECK_CONV 0 /* the reference's neck: plain 3x3 Conv2dBlocks in IDAUp (model/backbone/dla_neck.py:11-38,56-57)   */
#define MC_NECK_DCN 1  /* the DCN variant BASELINE.json's north_star names: the proj_j / node_j 3x3 convolutions are       */
                       /* modulated deformable convolutions (DCNv2 pack: conv.weight (Cout,Cin,3,3) without bias,          */
                       /* conv.conv_offset.weight (27,Cin,3,3) + .bias (27) -> 18 offsets ((dy,dx) per tap) + 9 mask       */
                       /* logits), bn1 + ReLU unchanged.  Operator semantics = torchvision.ops.deform_conv2d; the          */
                       /* reference repository itself has no deformable layer.  Inference only.      */

/* Build the engine for DLA-34 + DLAUp + MonoCon heads (MonoConDetector.__init__,
 * monocon_detector.py:29-50) at a fixed input geometry (H, W multiples of 32, the reference pads
 * to /32: transforms/default_transforms.py:421-426) and a maximum batch. */
MC_API int mc_create(mc_handle** out, int device, int max_batch, int H, int W, int precision_mode);
/* The same with the neck variant chosen (mc_create = MC_NECK_CONV). */
MC_API int mc_create_ex(mc_handle** out, int device, int max_batch, int H, int W, int precision_mode, int neck_variant);

/* Hand one entry of the reference state_dict to the engine (nn.Module.load_state_dict,
 * engine/base_engine.py:208; monocon_detector.py:80-82).  `key` is the reference's key
 * (e.g. "backbone.level2.tree1.conv1.weight"); `data` is fp32 (int64 num_batches_tracked entries
 * are not needed), host or device memory; copied before returning (synchronous). */
MC_API int mc_set_param(mc_handle* h, const char* key, const float* data, const int64_t* shape, int ndim);

/* Fold eval-mode BatchNorm into per-channel scale/shift and repack the weights for the kernels
 * (OIHW fp32 -> tap-major K-blocked bf16 / fp32).  training = 1 keeps the BatchNorm parameters separate for mc_forward_train;
 * training = 2 additionally keeps what the backward pass needs (raw convolution outputs, batch statistics) and allocates the
 * gradient buffers (mc_backward_train below).  Two training engines, same entry points:
 *   MC_PREC_FP32  fp32 storage, FFMA kernels everywhere -- the strict twin (gradients pinned to the reference's own step);
 *   MC_PREC_BF16  the tensor-core step (BASELINE.json configs[2] "bf16"): bf16 activations / raw outputs / activation gradients,
 *                 tcgen05 forward, dgrad (the forward kernels on flipped weights) and wgrad (csrc/wgrad_tc.cu), fp32 master weights,
 *                 statistics, parameter gradients and optimiser state (csrc/train_engine_tc.cu).  ~30x the fp32 engine's speed.
 * MC_PREC_FP32_TC handles are inference-only.  Eval entry points must not be used on a training handle.  Synchronous. */
MC_API int mc_finalize_params(mc_handle* h, int training);
/* New weights into a finalized engine: stage EVERY tensor again with mc_set_param, then mc_refresh_params folds / packs them into
 * the same device buffers (same plan, same mode).  Pointers handed out earlier (mc_train_tensor, tensor maps inside captured CUDA
 * graphs, optimiser handles) stay valid; gradients are invalidated.  Synchronises the device.  This is what a training loop that
 * steps torch-side parameters calls once per iteration instead of rebuilding the handle. */
MC_API int mc_refresh_params(mc_handle* h);

/* MonoConDetector.forward in eval mode (monocon_detector.py:53-65): img (B,3,H,W) fp32 NCHW on
 * the device -> the ten prediction maps, NCHW fp32 on the device, (B,C_i,H/4,W/4). */
MC_API int mc_forward(mc_handle* h, const float* img_nchw, int B, float* const pred_out[MC_NUM_PRED], void* stream);

/* MonoConDetector.forward in train() mode up to the ten prediction maps (monocon_detector.py:53-61, first half of
 * SURVEY.md 8(f) row 1): every BatchNorm normalises with the statistics of this batch and updates its running statistics
 * (momentum 0.1; AttnBatchNorm2d: base BN momentum 0.03 / eps 1e-3, and the 10-channel BatchNorm of the attention branch
 * over the batch, so 2 <= B).  Needs mc_finalize_params(h, 1 or 2) on an MC_PREC_FP32 or MC_PREC_BF16 handle; the backward pass is mc_backward_train.
 * mc_get_buffer copies the current running_mean / running_var of a BatchNorm of the plan (reference state_dict key, e.g.
 * "backbone.level2.tree1.bn1.running_var") to the host; the two unused outer `project` BatchNorms of level3 / level4
 *  are not part of the plan and are not updated. */
MC_API int mc_forward_train(mc_handle* h, const float* img_nchw, int B, float* const pred_out[MC_NUM_PRED], void* stream);
/* Number of mc_forward_train calls so far.  The raw outputs / batch statistics mc_backward_train differentiates are those of the
 * LAST train-mode forward: a caller that holds an older graph (gradient accumulation, two losses) compares the generation it
 * recorded with this before calling mc_backward_train. */
MC_API long long mc_train_generation(const mc_handle* h);
MC_API int mc_get_buffer(mc_handle* h, const char* key, float* out_host, int n);

/* decode_heatmap + the origin shift of _get_bboxes (monocon_heads.py:399-482, 313-329;
 * utils/tensor_ops.py:17-59) with fixed-shape outputs (the reference's ragged per-image lists are
 * rows with valid != 0, in order):
 *   P2      (B,3,4) fp32 device  -- KITTICalibration.P2 per image (monocon_heads.py:501,543)
 *   invP    (B,4,4) fp32 device  -- inverse of the 4x4-padded P2, computed by the caller on the host
 *                                   exactly as the reference does (monocon_heads.py:544-546)
 *   box2d   (B,K,5) fp32   x1,y1,x2,y2,score*sigma
 *   box3d   (B,K,7) fp32   x,y(bottom-centre),z,dim0,dim1,dim2,rot_y
 *   labels  (B,K)   int64  class id        inds (B,K) int64  flat y*W+x index on the feature map
 *   valid   (B,K)   uint8  score*sigma > thres
 * Top-k ties are broken by the lowest flat index (class-major), deterministic. */
MC_API int mc_decode(mc_handle* h, const float* const pred[MC_NUM_PRED], int B, const float* P2, const float* invP,
              int img_h, int img_w, int topk, float thres, float* box2d, float* box3d, int64_t* labels,
              int64_t* inds, uint8_t* valid, void* stream);

/* Host-buffer end-to-end call (what MonoConDetector.batch_eval does per batch,
 * monocon_detector.py:68-77, including the H2D of engine_utils.move_data_device and the .cpu() of
 * bbox_*_to_result, monocon_heads.py:561-586): pinned or pageable host img / P2 / invP in, host
 * decode outputs out.  Copies, forward, decode and the read-back run on `stream`; the call returns
 * after the stream is synchronised. */
MC_API int mc_infer_host(mc_handle* h, const float* img_nchw_host, int B, const float* P2_host, const float* invP_host,
                  int topk, float thres, float* box2d_host, float* box3d_host, int64_t* labels_host,
                  int64_t* inds_host, uint8_t* valid_host, void* stream);

/* Pipelined form of mc_infer_host for streaming inference (the data-loader loop of engine/monocon_engine.py:134-139):
 * two slots; mc_infer_host_submit enqueues H2D (copy stream) -> forward + decode (compute stream) -> D2H (copy
 * stream) and returns immediately, mc_infer_host_wait blocks until that slot's host outputs are complete.  Submitting
 * batch i+1 before waiting for batch i overlaps its H2D copy with the compute of batch i.  The host buffers must stay
 * valid (and should be pinned) until the wait returns.  Uses the engine's own streams. */
MC_API int mc_infer_host_submit(mc_handle* h, int slot, const float* img_nchw_host, int B, const float* P2_host,
                                const float* invP_host, int topk, float thres, float* box2d_host, float* box3d_host,
                                int64_t* labels_host, int64_t* inds_host, uint8_t* valid_host);
MC_API int mc_infer_host_wait(mc_handle* h, int slot);
/* The same pipeline fed with what the reference's data loader actually holds before its transforms
 * (dataset/monocon_dataset.py:38-42, transforms/default_transforms.py:376-431): pinned host uint8 HWC frames
 * (B, H0, W0, 3) + per-frame valid sizes hw[B][2]; Normalize + Pad + ToTensor run inside the input-packing kernel
 * (mc_infer_device_u8).  A quarter of the H2D bytes of the fp32 NCHW form: 23.6 MB instead of 94.4 MB per batch of 16. */
MC_API int mc_infer_host_u8_submit(mc_handle* h, int slot, const uint8_t* img_hwc_host, const int32_t* hw_host, int B, int H0,
                                   int W0, const float* P2_host, const float* invP_host, int topk, float thres,
                                   float* box2d_host, float* box3d_host, int64_t* labels_host, int64_t* inds_host,
                                   uint8_t* valid_host);

/* Same as mc_forward + mc_decode with device inputs and device outputs, keeping the ten maps
 * inside the engine (used by the throughput bench and the multi-GPU shard path). */
MC_API int mc_infer_device(mc_handle* h, const float* img_nchw, int B, const float* P2, const float* invP, int topk,
                    float thres, float* box2d, float* box3d, int64_t* labels, int64_t* inds, uint8_t* valid,
                    void* stream);

/* Post-decode KITTI conversion of the decoded 3D boxes on the device: get_valid_bboxes_3d +
 * convert_to_kitti_3d (utils/kitti_convert_utils.py:16-171; corners / projection of utils/geometry_ops.py:7-163), one
 * thread per detection, the reference's float32 / float64 split.  box3d (B,K,7), valid (B,K) from mc_decode; P2 (B,3,4);
 * img_hw (B,2) int32 = img_metas['ori_shape'].  bbox_out (B,K,4) float64: projected 2D box clipped to the image; alpha_out
 * (B,K) float32 = -atan2(x, z) + rot_y; keep_out (B,K) uint8 = valid and the box touches the image.  Stateless (errors via
 * mc_last_error(NULL)). */
MC_API int mc_kitti_boxes(int device, const float* box3d, const uint8_t* valid, const float* P2, const int32_t* img_hw, int B,
                          int K, double* bbox_out, float* alpha_out, uint8_t* keep_out, void* stream);

/* Input pipeline fused into the engine: Normalize(mean, std) + Pad(32) + ToTensor of the
 * reference's test pipeline (transforms/default_transforms.py:376-431, dataset/monocon_dataset.py:38-42) happen inside
 * the input-packing kernel.  img_hwc: (B, H0, W0, 3) uint8 on the device, dense, the frames top-left aligned;
 * hw: (B, 2) int32 on the device, the valid (height, width) of every frame (<= H0, W0; H0 <= H, W0 <= W of mc_create).
 * Values are float((u - mean) / std) computed in double like numpy does, pixels outside a frame are zero (Pad's canvas):
 * the result is bit-identical to feeding the reference-transformed fp32 tensor to mc_forward.  The default table is the
 * reference's mean = [123.675, 116.28, 103.53], std = [58.395, 57.12, 57.375]. */
MC_API int mc_set_normalization(mc_handle* h, const double mean[3], const double std_dev[3]);
MC_API int mc_forward_u8(mc_handle* h, const uint8_t* img_hwc, const int32_t* hw, int B, int H0, int W0,
                         float* const pred_out[MC_NUM_PRED], void* stream);
MC_API int mc_infer_device_u8(mc_handle* h, const uint8_t* img_hwc, const int32_t* hw, int B, int H0, int W0, const float* P2,
                              const float* invP, int topk, float thres, float* box2d, float* box3d, int64_t* labels,
                              int64_t* inds, uint8_t* valid, void* stream);

/* Multi-GPU inference (one process per GPU, the batch sharded across ranks): all-gather of the decode
 * outputs over peer memory, fused into the decode kernel.  Every rank owns a gather block of 2 buffers x world slots;
 * a slot holds one rank's (max_batch, topk) decode outputs packed as box2d | box3d | labels | inds | valid (each field
 * 16-byte aligned, the layout of monocon_pytorch_b200/dist.py).  The decode kernel stores every detection row into its
 * own slot locally AND into the same slot of every peer's block (plain stores over NVLink, no NCCL kernel, no extra
 * pass over the data); its last CTA publishes a generation number to every peer.
 *   mc_gather_create   allocate the local block, return its 64-byte cudaIpcMemHandle_t (exchange them out of band,
 *                      e.g. torch.distributed.all_gather, and pass all `world` handles, rank order, to _connect)
 *   mc_infer_device_gather   mc_infer_device with the outputs going to buffer `buf` (0 / 1) of every rank; first tells
 *                      the peers that this rank is done reading the previous contents of `buf`
 *   mc_gather_wait     enqueue, on `stream`, a wait until the slots of all ranks for the last generation issued on
 *                      `buf` have arrived; later work on the stream may read mc_gather_buffer(buf)
 * Alternate buf = 0, 1 and wait one batch late to overlap the exchange with the next forward.  B must be max_batch. */
MC_API int mc_gather_create(mc_handle* h, int world, int rank, int topk, void* ipc_handle_out_64B);
MC_API int mc_gather_connect(mc_handle* h, const void* all_handles_world_x_64B);
MC_API size_t mc_gather_slot_bytes(const mc_handle* h);
MC_API int mc_gather_buffer(mc_handle* h, int buf, void** device_ptr);
MC_API int mc_infer_device_gather(mc_handle* h, const float* img_nchw, int B, const float* P2, const float* invP,
                                  float thres, int buf, void* stream);
MC_API int mc_gather_wait(mc_handle* h, int buf, void* stream);

/* Engine-owned copies of the ten maps of the last mc_infer_* call (device pointers, NCHW fp32). */
MC_API int mc_get_pred_ptrs(mc_handle* h, float* out_ptrs[MC_NUM_PRED]);
/* Copy those maps (first B images) into caller-provided NCHW fp32 device buffers, on `stream`. */
MC_API int mc_copy_pred(mc_handle* h, int B, float* const dst[MC_NUM_PRED], void* stream);

/* Options: "conv_impl" (MC_CONV_*), "use_graph" (0/1: replay the forward as a CUDA graph); bf16 training engines: "train_debug"
 * (1: also keep the fp32 gradient of the head stems for mc_debug_train_dump), "head_backward" (0: the fp32 twin's heads kernels). */
MC_API int mc_set_option(mc_handle* h, const char* name, int value);

/* MC_PREC_FP32_TC only.  mc_calibrate_scales: run the forward on a sample batch (device fp32 NCHW, as mc_forward) and fit the
 * per-tensor power-of-two scales of the fp16 hi / lo planes to it (each tensor's largest value lands in [2^11, 2^12): 16-32x
 * headroom, ~22 significant bits for everything within 2^-14 of the maximum); synchronises.  The scales live in a device table
 * the kernels read at run time, so captured CUDA graphs follow a recalibration.  An engine that was never calibrated uses
 * scale 1 for every tensor -- exact for activations of order 0.1 ... 10^4, as DLA-34 behind its BatchNorms produces.
 * mc_scale_status: the largest |stored value| any tensor has seen since the previous call, as a fraction of the fp16 limit,
 * and the number of tensors that hit the limit (their values were clamped: recalibrate and repeat the batch); synchronises. */
MC_API int mc_calibrate_scales(mc_handle* h, const float* img_nchw, int B, void* stream);
MC_API int mc_scale_status(mc_handle* h, float* max_fraction, int* n_saturated);

/* Introspection. */
MC_API size_t mc_workspace_bytes(const mc_handle* h);           /* activation arena + packed weights            */
MC_API int mc_num_kernel_launches(const mc_handle* h);          /* kernels enqueued by one mc_infer_device call */
MC_API double mc_flops_per_image(const mc_handle* h);           /* 2*MAC of every convolution in the plan       */
MC_API double mc_bytes_per_image(const mc_handle* h);           /* layer-wise activation bytes (read+write)     */
MC_API const char* mc_last_error(const mc_handle* h);           /* h may be NULL: last error of mc_create       */
MC_API void mc_destroy(mc_handle* h);

/* Per-stage device timing (CUDA events on `stream`, eager launches, mean over `iters` passes after one
 * warm-up pass).  Stages: pack_input, every op of the plan (convolutions, pools, up-samplings, the
 * AttnBN + 1x1 head stage) and decode; mc_num_stages entries are written to ms_out. */
MC_API int mc_num_stages(mc_handle* h);
MC_API int mc_stage_info(mc_handle* h, int stage, char* name, int name_len, double* flops_per_image,
                         double* bytes_per_image, int* is_tensor_core);
MC_API int mc_profile_stages(mc_handle* h, const float* img_nchw, int B, const float* P2, const float* invP, int iters,
                             float* ms_out, void* stream);

/* Debug / per-layer parity: copy a named intermediate activation of the last forward (e.g.
 * "backbone.level2", "neck.feat", "head.stems") into an NCHW fp32 device buffer. */
MC_API int mc_debug_tensor_shape(mc_handle* h, const char* name, int* C, int* H, int* W);
MC_API int mc_debug_tensor(mc_handle* h, const char* name, int B, float* out_nchw, void* stream);

/* Stand-alone operator entry for kernel-level parity tests (the reference op is
 * torch.nn.functional.conv2d + eval BatchNorm + residual + ReLU as composed in BasicBlock.forward,
 * dla.py:34-51, Root.forward :124-132 and Conv2dBlock.forward, dla_neck.py:34-38):
 *   x      (B,Cin,H,W) fp32 NCHW device        w (Cout,Cin,k,k) fp32 device
 *   scale, shift (Cout) fp32 device (folded BN; bias-only: scale = 1)
 *   residual (B,Cout,Ho,Wo) fp32 NCHW device or NULL       y (B,Cout,Ho,Wo) fp32 NCHW device
 *   split: number of equal channel groups the input is presented as (tests the multi-source
 *          "concat-free" K-split used for Root / node convolutions); precision/impl as above. */
MC_API int mc_conv2d(int device, int precision_mode, int conv_impl, const float* x, int B, int Cin, int H, int W,
              const float* w, int Cout, int k, int stride, int pad, const float* scale, const float* shift,
              const float* residual, int relu, int split, float* y, void* stream, char* err, int err_len);

/* Stand-alone operator entry for the modulated deformable convolution of the MC_NECK_DCN plan -- the fused tcgen05 kernel of
 * csrc/dcn_tc.cu where it applies (tensor-core storage, channel groups that are multiples of 64, Cout <= 256, MC_DCN_FUSE != 0),
 * else csrc/dcn.cu: deformable columns, then a 1x1 convolution over 9 Cin column channels; reference op: torchvision.ops.deform_conv2d
 * (3x3, stride 1, padding 1, dilation 1, one offset group), torchvision/ops/deform_conv.py:14-96:
 *   x (B,Cin,H,W), offset (B,18,H,W) ((dy,dx) per tap), mask (B,9,H,W) (the modulation itself, already in (0,1)),
 *   w (Cout,Cin,3,3), bias (Cout) or NULL, y (B,Cout,H,W): fp32 NCHW device buffers.  split: the input presented as one or
 *   two equal channel groups (the concat-free sources of a node block). */
MC_API int mc_deform_conv2d(int device, int precision_mode, const float* x, int B, int Cin, int H, int W, const float* offset,
                            const float* mask, const float* w, const float* bias, int Cout, int split, float* y, void* stream,
                            char* err, int err_len);

/* Stand-alone operator entry for the tensor-core weight gradient of the bf16 training step (csrc/wgrad_tc.cu; reference op:
 * the weight gradient of torch.nn.functional.conv2d, k x k / stride 1 / pad (k - 1) / 2, as autograd computes it for
 * BasicBlock / Root / Conv2dBlock, dla.py:22-51,117-132, dla_neck.py:24-38):
 *   x  (B,Cin,H,W) fp32 NCHW device, presented as `split` equal channel groups (concat-free sources)
 *   dy (B,Cout,H,W) fp32 NCHW device -- both are rounded to bf16 NHWC, what the training engine stores
 *   dw [k*k][Cin][Cout] fp32 device, overwritten (the engine's master-weight layout). */
MC_API int mc_conv2d_wgrad_tc(int device, const float* x, int B, int Cin, int H, int W, const float* dy, int Cout, int k,
                              int split, float* dw, void* stream, char* err, int err_len);

/* ------------------------------------------------------------------------------------------------------------------
 * Training-side rows of the hot path.  Stateless entry points on caller-owned device memory
 * (fp32 unless stated); constants are the reference defaults num_classes 3, num_kpts 9, num_alpha_bins 12
 * (monocon_detector.py:12-17).  The network's own backward pass is the experimental block further down; these replace the Python
 * target generator, the loss block with its gradients w.r.t. the ten prediction maps, and clip + AdamW.
 * Errors: non-zero return, message via mc_train_last_error() (thread-local).
 * ------------------------------------------------------------------------------------------------------------------ */

/* label tensors of data_dict['label'] after collation (dataset/monocon_dataset.py:160-171, 197-205) */
typedef struct mc_labels {
    const float* gt_bboxes;             /* (B,M,4)  x1,y1,x2,y2 in padded-image pixels */
    const uint8_t* gt_labels;           /* (B,M)                                       */
    const float* gt_bboxes_3d;          /* (B,M,7)  loc(3) dim(3) ry ([6] = alpha source, target_generator.py:82) */
    const float* depths;                /* (B,M)                                       */
    const float* gt_kpts_2d;            /* (B,M,18) 9 x (x,y)                          */
    const uint8_t* gt_kpts_valid_mask;  /* (B,M,9)  visibility, used if >= 1           */
    const uint8_t* mask;                /* (B,M)    valid rows (not necessarily contiguous) */
} mc_labels;

/* the 15 tensors of TargetGenerator._create_empty_target (utils/target_generator.py:152-177) */
typedef struct mc_targets {
    float* center_heatmap;              /* (B,3,h,w)  */
    float* kpt_heatmap;                 /* (B,9,h,w)  */
    float* wh;                          /* (B,M,2)    */
    float* offset;                      /* (B,M,2)    */
    float* dim;                         /* (B,M,3)    */
    float* alpha_cls;                   /* (B,M,1)    */
    float* alpha_offset;                /* (B,M,1)    */
    float* depth;                       /* (B,M,1)    */
    float* center2kpt_offset;           /* (B,M,18)   */
    float* kpt_heatmap_offset;          /* (B,M,18)   */
    int64_t* indices;                   /* (B,M)      */
    int64_t* indices_kpt;               /* (B,M*9)    */
    uint8_t* mask_target;               /* (B,M) bool */
    float* mask_center2kpt_offset;      /* (B,M,18)   */
    float* mask_kpt_heatmap_offset;     /* (B,M,18)   */
} mc_targets;

/* TargetGenerator.__call__ (utils/target_generator.py:30-138; gaussian_radius / generate_gaussian_target,
 * utils/tensor_ops.py:62-125): zero-fills the targets, then one CTA per image compacts the valid rows, writes the
 * per-object / per-key-point targets and max-splats the Gaussians (integer atomicMax on non-negative floats).
 * Integer outputs are bit-identical to the reference, heat-map values within 1e-6 (expf). */
MC_API int mc_generate_targets(int device, int B, int max_objs, int feat_h, int feat_w, int pad_h, int pad_w,
                               const mc_labels* labels, const mc_targets* targets, void* stream);

/* MonoConDenseHeads._get_losses (monocon_heads.py:203-310) with losses/{focal,l1,dim,depth,cross_entropy}_loss.py.
 *   pred[]     the ten prediction maps (order of MC_NUM_PRED above), NCHW fp32
 *   losses_out 10 floats (device) in the reference's dict order: center_heatmap, wh, offset, dim, center2kpt_offset,
 *              kpt_heatmap, kpt_heatmap_offset, alpha_cls, alpha_reg, depth (loss weights applied; their plain sum is
 *              the training loss, utils/engine_utils.py:79-80)
 *   grad[]     NULL, or ten NCHW fp32 buffers receiving d(sum of the ten losses)/d(pred[i]) (what loss.backward() hands
 *              to the head convolutions)
 *   workspace  mc_losses_workspace_bytes() bytes of device memory; after the call ((double*)workspace)[6] != 0 means
 *              the batch had no valid object, where the reference asserts (losses/l1_loss.py:15). */
MC_API size_t mc_losses_workspace_bytes(void);
MC_API int mc_losses(int device, int B, int max_objs, int feat_h, int feat_w, const float* const pred[MC_NUM_PRED],
                     const mc_targets* targets, float* losses_out, float* const grad[MC_NUM_PRED], void* workspace,
                     void* stream);

/* clip_grad_norm_(parameters, max_norm, 2) + torch.optim.AdamW.step as called by engine/monocon_engine.py:94-100 on the
 * solver of :39-53, fused over all parameter tensors: one reduction launch + one update launch.
 *   create: host arrays of n device pointers (parameters, exp_avg, exp_avg_sq; the state is caller-owned so that it
 *           can live in the optimizer's state_dict) and element counts
 *   step:   host array of n gradient device pointers (NULL = parameter without gradient: skipped like torch does, e.g.
 *           the dead backbone.level{3,4}.project.* tensors); `step` is the 1-based count including this update; lr /
 *           beta1 change every iteration under the cyclic scheduler (solver/cyclic_scheduler.py:36-71);
 *           total_norm_out: optional device float receiving the pre-clip gradient norm. */
typedef struct mc_optimizer mc_optimizer;
MC_API int mc_optimizer_create(mc_optimizer** out, int device, int n_tensors, float* const* params, float* const* exp_avg,
                               float* const* exp_avg_sq, const int64_t* numel);
MC_API int mc_optimizer_step(mc_optimizer* o, float* const* grads, int step, double lr, double beta1, double beta2,
                             double eps, double weight_decay, double max_norm, float* total_norm_out, void* stream);
MC_API void mc_optimizer_destroy(mc_optimizer* o);
MC_API const char* mc_train_last_error(void);

/* ------------------------------------------------------------------------------------------------------------------
 * KITTI evaluation overlaps: the reference's numba.cuda rotated-IoU kernel and the numba CPU pass
 * behind it, as one CUDA kernel each (one thread per (box, query) pair).  Device pointers, stateless; errors via
 * mc_eval_last_error() (thread-local).
 *   mc_rotate_iou     rotate_iou_gpu_eval (engine/kitti_eval/rotate_iou.py:337-379): BEV boxes (N,5) / (K,5) float32
 *                     [cx, cy, dx, dy, angle]; criterion -1: IoU, 0: / query area, 1: / box area, 2: intersection area
 *                     (the reference evaluates devRotateIoUEval(query, box), :330-333); out (N,K) float32
 *   mc_box3d_overlap  d3_box_overlap (engine/kitti_eval/eval.py:128-164): camera boxes (N,7) / (K,7) float64
 *                     [x, y, z, l, h, w, ry]; BEV intersection x height overlap; criterion -1 / 0 / 1; out (N,K) float32
 * The reference's results are reproduced including its quirks (two identical boxes give 1/3: duplicate polygon vertices).
 * ------------------------------------------------------------------------------------------------------------------ */
MC_API int mc_rotate_iou(int device, const float* boxes, const float* qboxes, int N, int K, int criterion, float* out,
                         void* stream);
MC_API int mc_box3d_overlap(int device, const double* boxes, const double* qboxes, int N, int K, int criterion, float* out,
                            void* stream);
MC_API const char* mc_eval_last_error(void);

/* ------------------------------------------------------------------------------------------------------------------
 * Training step, backward kernels (second half) -- EXPERIMENTAL: one entry point per backward kernel
 * family of csrc/train_backward.cu, on plain fp32 NHWC device pointers.  Pinned on the CPU (the same kernel bodies run under
 * tests/host_shim against oracle/backward_oracle.py, which is pinned to the reference's own gradients) and on the B200
 * (tests/test_gpu_zz_train_backward.py); the engine's stage list drives them through mc_backward_train.  "+=": accumulates into a caller-zeroed
 * buffer; "=": overwrites.  Errors via mc_bw_last_error().
 *   mc_bw_conv        y = conv2d(cat(src...), w) (dla.py:22-31,117-121,228-236; dla_neck.py:24-31; monocon_heads.py:114-131):
 *                     dw[k*k][Cin][Cout] += wgrad (NULL = skip); dsrc[s] (dense NHWC, NULL = not needed) += dgrad.
 *                     srcWp / srcXoff: physical row pitch / first column of each source (NULL = dense); wT_scratch: k*k*Cin*Cout
 *                     floats for a [k*k][Cout][Cin] copy of w that makes the dgrad weight reads coalesced (NULL = strided reads)
 *   mc_bw_batchnorm   y = relu?(BN_train(raw) + res) (nn.BatchNorm2d in train(), dla.py:24,30,119,...): draw =, dres +=,
 *                     dgamma =, dbeta =; mean / inv: batch mean and rsqrt(biased var + eps) of raw; sums: 2*C doubles
 *   mc_bw_colsum      out[c] = sum_P x[P][C] (bias gradients); sums: C doubles
 *   mc_bw_maxpool2    MaxPool2d(2,2) (dla.py:176-177,193): dx += dy at the first maximum of each window
 *   mc_bw_upsample2   depthwise ConvTranspose2d k=4 s=2 p=1 (dla_neck.py:58-65): dx +=, dw[C][16] +=
 *   mc_bw_heads       dL/dpred (NCHW, mc_losses) -> output transforms -> ten 1x1 convolutions -> ReLU -> nine
 *                     AttnBatchNorm2d (monocon_heads.py:165-200; attentive_norm.py:79-91,154-164): dstems [B][HW][576] =,
 *                     dw [65][64] =, dbias [65] =, the AttnBN parameter gradients =.  sums / coefA / coefB: the forward's
 *                     per-sample statistics and affine of this batch
 * ------------------------------------------------------------------------------------------------------------------ */
MC_API int mc_bw_conv(int nsrc, const float* const* src, float* const* dsrc, const int* srcC, const int* srcWp, const int* srcXoff,
                      int B, int Hin, int Win, int Hout, int Wout, int Cout, int k, int stride, int pad, const float* w,
                      const float* dy, float* dw, float* wT_scratch, void* stream);
MC_API int mc_bw_batchnorm(const float* dy, const float* y, const float* raw, const float* mean, const float* inv, const float* gamma,
                           long long P, int C, int relu, double* sums, float* draw, float* dres, float* dgamma, float* dbeta,
                           void* stream);
MC_API int mc_bw_colsum(const float* x, long long P, int C, double* sums, float* out, void* stream);
MC_API int mc_bw_maxpool2(const float* x, const float* dy, float* dx, int B, int C, int Hin, int Win, void* stream);
MC_API int mc_bw_upsample2(const float* x, const float* w, const float* dy, float* dx, float* dw, int B, int C, int Hin, int Win,
                           void* stream);
MC_API long long mc_bw_heads_scratch_bytes(int B, int HW);
MC_API int mc_bw_heads(const float* const* pred, const float* const* dpred, const float* stems, const double* sums,
                       const float* coefA, const float* coefB, const float* att_w, const float* att_gamma, const float* att_beta,
                       const float* bank_w, const float* bank_b, const float* w, int B, int HW, void* scratch, float* dstems,
                       float* dw, float* dbias, float* datt_w, float* datt_gamma, float* datt_beta, float* dbank_w, float* dbank_b,
                       void* stream);
MC_API const char* mc_bw_last_error(void);

/* The backward pass over a whole stage list: the engine's forward ops (fused conv + BatchNorm + residual + ReLU over
 * channel-concatenated sources, 2x2 max-pool, depthwise 2x upsampling, the heads) described as plain records, walked in
 * reverse.  Every tensor with g != NULL has its gradient buffer zeroed first ([B][H][W][C] floats, dense), then each op adds
 * its contributions in stream order; parameter-gradient buffers (dw of convolutions and upsamplings) are zeroed as well.
 * Same experimental status as the kernels above; pinned on the CPU over the full DLA-34 + DLAUp + heads graph against the
 * reference-pinned oracle (tests/test_backward_graph_host.py). */
typedef struct mc_bw_tensor {
    const float* x;        /* forward value (post-activation output of its producer), NHWC, row pitch Wp, first column xoff */
    float* g;              /* gradient, dense NHWC; NULL = not needed (the input image) */
    int C, H, W, Wp, xoff;
} mc_bw_tensor;
typedef struct mc_bw_heads_args {
    const float* pred[MC_NUM_PRED];
    const float* dpred[MC_NUM_PRED];
    const double* sums;
    const float *coefA, *coefB, *att_w, *att_gamma, *att_beta, *bank_w, *bank_b, *w;
    void* scratch;         /* mc_bw_heads_scratch_bytes(B, HW) */
    float *dw, *dbias, *datt_w, *datt_gamma, *datt_beta, *dbank_w, *dbank_b;
} mc_bw_heads_args;
enum { MC_BW_CONV = 0, MC_BW_POOL = 1, MC_BW_UP = 2, MC_BW_HEADS = 3 };
typedef struct mc_bw_op {
    int type;
    int nsrc, src[4], dst;     /* tensor indices; POOL / UP / HEADS use src[0] (HEADS: the pre-norm stems, no dst) */
    int residual, relu;        /* CONV: residual tensor index or -1; ReLU after the add */
    int k, stride, pad, cout;  /* CONV */
    const float* w;            /* CONV: [k*k][Cin][Cout]; UP: [C][16] */
    float* dw;                 /* gradient of w (zeroed, then accumulated); NULL = skip */
    float* wT;                 /* CONV: scratch, k*k*Cin*Cout floats, for the transposed weights of the dgrad; NULL = strided reads */
    float* dbias;              /* CONV without BatchNorm: [Cout] = column sums of the output gradient; NULL = none */
    int has_bn;                /* CONV followed by a train-mode BatchNorm: */
    const float *raw, *mean, *inv, *gamma;     /* raw convolution output, batch mean, rsqrt(var + eps), weight (NULL = 1) */
    float *dgamma, *dbeta;
    float* draw;               /* scratch: B*H*W*Cout floats */
    double* sums;              /* scratch: 2*Cout doubles */
    const mc_bw_heads_args* heads;   /* HEADS */
} mc_bw_op;
/* The same pass driven by the engine: mc_finalize_params(h, 2) makes mc_forward_train keep what the backward needs (raw
 * convolution outputs, batch statistics); mc_backward_train(pred, dpred) -- the maps mc_forward_train wrote and dL/dpred from
 * mc_losses -- runs mc_bw_run_graph over the engine's own stage list; mc_get_grad copies one parameter gradient to the host
 * in the reference's state_dict layout (OIHW weights), i.e. what `loss.backward()` leaves in `param.grad`
 * (engine/monocon_engine.py:88-91).  The six `backbone.level{3,4}.project.*` tensors receive no gradient in the reference
 *  and are an error here. */
MC_API int mc_backward_train(mc_handle* h, const float* const pred[MC_NUM_PRED], const float* const dpred[MC_NUM_PRED], int B,
                             void* stream);
MC_API int mc_get_grad(mc_handle* h, const char* key, float* out_host, int64_t n);
/* The same pass in segments (see mc_bw_run_graph_range): stages [op_first, op_last) of the engine's stage list, the first call
 * of a pass has op_last == mc_num_backward_stages(h).  mc_train_tensor reports for every trainable buffer the stage whose
 * backward finishes its gradient. */
MC_API int mc_num_backward_stages(mc_handle* h);              /* -1: not a backward-enabled engine */
MC_API int mc_backward_train_segment(mc_handle* h, const float* const pred[MC_NUM_PRED], const float* const dpred[MC_NUM_PRED], int B,
                                     int op_first, int op_last, void* stream);
/* Engine-resident training: every trainable buffer of the plan IN THE ENGINE'S LAYOUT (packed [k*k][Cin][Cout] convolution
 * weights, BatchNorm weight / bias, biases, upsampling taps, the head matrices) with its gradient buffer.  clip_grad_norm_ and
 * AdamW (engine/monocon_engine.py:94-100) are element-wise, so mc_optimizer_create / _step over these pointers updates the
 * weights in place and the whole iteration -- forward, targets, losses, backward, optimiser -- stays on the device with no
 * unpacking; mc_get_param returns the current value of one parameter in state_dict layout (checkpointing, eval engines). */
/* Debug / test: the engine's own backward records (host arrays owned by the handle, device pointers inside) -- lets a test
 * replay the identical pass elsewhere (tests/test_gpu_zz_train_backward.py replays it on the CPU host shim). */
MC_API int mc_debug_bw_graph(mc_handle* h, const mc_bw_tensor** tensors, int* n_tensors, const mc_bw_op** ops, int* n_ops);
/* Debug / test: one buffer of the last training pass as NCHW fp32 on the device.  kind 0: forward tensor `index` (indices of
 * mc_debug_bw_graph); 1: its gradient; 2 / 3: raw output / gradient of the raw output of the convolution at STAGE `index` (3: bf16
 * tensor-core engines only; the gradient of a stride-2 convolution comes zero-inserted at input resolution).  In an MC_PREC_BF16
 * training engine the records of mc_debug_bw_graph describe the structure only: activation pointers are bf16, g / raw / draw are null. */
MC_API int mc_debug_train_dump(mc_handle* h, int kind, int index, int B, float* out_nchw, void* stream);
MC_API int mc_num_train_tensors(mc_handle* h);                /* -1: not a backward-enabled engine */
MC_API int mc_train_tensor(mc_handle* h, int i, float** param, float** grad, int64_t* numel, int* stage, char* key, int key_cap);
MC_API int mc_get_param(mc_handle* h, const char* key, float* out_host, int64_t n);
MC_API int mc_bw_run_graph(const mc_bw_tensor* tensors, int n_tensors, const mc_bw_op* ops, int n_ops, int B, void* stream);
/* One segment of the same pass: ops[op_last - 1] down to ops[op_first]; zero != 0 (the first segment of a pass, op_last ==
 * n_ops) zeroes every gradient buffer first.  Walking n_ops .. 0 in several calls equals one mc_bw_run_graph; between two calls
 * the gradients of the finished stages are final, which is where data-parallel training launches their all-reduce. */
MC_API int mc_bw_run_graph_range(const mc_bw_tensor* tensors, int n_tensors, const mc_bw_op* ops, int n_ops, int B, int op_first,
                                 int op_last, int zero, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* MONOCON_B200_H_ */
