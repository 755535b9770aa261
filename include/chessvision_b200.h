/* chessvision_b200.h — C ABI of the B200-native image->FEN path.
 *
 * The reference (gudbrandtandberg/ChessVision-3LC) has no FFI layer: its boundary is the Python class
 * `chessvision.ChessVision` (chessvision/core.py:22-567).  These entry points are what a ctypes binding inside that
 * class calls instead of PyTorch-eager + OpenCV; `chessvision-3lc_b200/chessvision/_native.py` is that binding and
 * INTEGRATION.md shows the stub a maintainer of the reference would add.
 *
 * Conventions: plain pointers and sizes only; every `*_dev` style pointer is a device pointer in caller-allocated
 * memory unless the function name ends in `_host`; `stream` is a `cudaStream_t` passed as `void*` (NULL = default
 * stream); return value 0 = ok, negative = error (text via cvb_last_error); no function throws; nothing allocates
 * device memory after cvb_create / cvb_load_* / cvb_train_create, except the lazily sized staging of cvb_decode_jpeg and of
 * the *_hw entry points (first use / change of image size).
 */
#ifndef CHESSVISION_B200_H
#define CHESSVISION_B200_H

#include <stdint.h>

#if defined(__GNUC__)
#define CVB_API __attribute__((visibility("default")))
#else
#define CVB_API
#endif

#ifdef __cplusplus
extern "C" {
#endif

typedef struct cvb_ctx cvb_ctx;

/* One entry of a PyTorch state_dict, host memory, float32, C-contiguous. */
typedef struct cvb_tensor {
    const char* name;
    const float* data;
    int32_t ndim;
    int64_t shape[4];
} cvb_tensor;

/* Per-board results; any pointer may be NULL (that output is then skipped). */
typedef struct cvb_outputs {
    float* logits;         /* [N,256,256]  UNet logits (BoardExtractionResult.probabilities, core.py:287,306) */
    uint8_t* mask;         /* [N,256,256]  {0,255}     (binary_mask, utils.py:101-112)                        */
    int32_t* quad;         /* [N,4,2]      corners (x,y) in the 256x256 mask frame, after _rotate_quadrangle   */
    uint8_t* found;        /* [N]          1 if a quadrangle was found (core.py:281)                           */
    int32_t* status;       /* [N]          0 none, 1 found, 2 capacity overflow in the contour kernel          */
    uint8_t* board;        /* [N,512,512]  warped, gray, flipped board (core.py:298-300); zero where !found     */
    float* probs;          /* [N,64,13]    softmax probabilities (PositionResult.model_probabilities)          */
    uint8_t* labels;       /* [N,64]       argmax class per square, order a8..h8,a7..h1 (core.py:326)          */
    uint8_t* labels_valid; /* [N,64]       after rule 1 "no_pawns_on_ends" (core.py:451-469)                   */
    char* fen;             /* [N,2,72]     [0] original_fen, [1] fen; NUL padded (core.py:336,350)             */
    uint8_t* squares;      /* [N,64,64,64] the board again as 64 crops in extract_squares order (core.py:420-439:
                              PositionResult.squares); written by the warp kernel, zero where !found               */
} cvb_outputs;

CVB_API int cvb_version(void);

/* ChessVision.__init__ (core.py:25-64): context bound to one GPU; workspaces sized for `max_batch` boards per chunk. */
CVB_API cvb_ctx* cvb_create(int device, int max_batch);
CVB_API void cvb_destroy(cvb_ctx* ctx);
CVB_API const char* cvb_last_error(const cvb_ctx* ctx);
CVB_API int cvb_max_batch(const cvb_ctx* ctx);

/* _initialize_board_extractor / _initialize_classifier (core.py:84-150) after utils.load_model_checkpoint
 * (utils.py:42-86): fold BatchNorm (eps 1e-5), pack fp16 K-major weight matrices, upload. */
CVB_API int cvb_load_unet(cvb_ctx* ctx, const cvb_tensor* state_dict, int n_tensors);
CVB_API int cvb_load_resnet18(cvb_ctx* ctx, const cvb_tensor* state_dict, int n_tensors);

/* cv2.resize(img, (256,256), INTER_AREA) for 512x512 inputs (core.py:212): u8[N,2h,2w,3] -> u8[N,h,w,3]. */
CVB_API int cvb_resize_area_half(cvb_ctx* ctx, const uint8_t* img, int N, int h, int w, uint8_t* out, void* stream);

/* extract_board up to the logits (core.py:212-220) + sigmoid/threshold (core.py:273-276).
 * img u8[N,512,512,3] BGR.  logits and/or mask may be NULL. */
CVB_API int cvb_unet_forward(cvb_ctx* ctx, const uint8_t* img, int N, float threshold, float* logits, uint8_t* mask, void* stream);

/* create_binary_mask(sigmoid(logits), thr) for externally produced logits (core.py:273-276). */
CVB_API int cvb_mask_from_logits(cvb_ctx* ctx, const float* logits, int N, float threshold, uint8_t* mask, void* stream);

/* ChessVision._find_quadrangle (core.py:358-379) on N masks u8[N,256,256]. */
CVB_API int cvb_mask_to_quad(cvb_ctx* ctx, const uint8_t* mask, int N, int32_t* quad, uint8_t* found, int32_t* status, void* stream);

/* _scale_quadrangle + utils.extract_perspective + BGR2GRAY + flip (core.py:291-300): img u8[N,H,W,3] -> board. */
CVB_API int cvb_warp_squares(cvb_ctx* ctx, const uint8_t* img, const int32_t* quad, const uint8_t* found, int N, int H, int W,
                     uint8_t* board, void* stream);

/* utils.extract_perspective (utils.py:115-132) for one image and caller-supplied corners: cv2.getPerspectiveTransform(corners,
 * ((0,0),(w,0),(w,h),(0,h))) + cv2.warpPerspective(img, M, (w,h)) (INTER_LINEAR, BORDER_CONSTANT 0), bit-identical to OpenCV.
 * img u8[H,W,C] (C = 1 or 3), corners f32[4][2] (x,y), out u8[out_h,out_w,C]; all device pointers. */
CVB_API int cvb_warp_perspective(cvb_ctx* ctx, const uint8_t* img, int H, int W, int C, const float* corners, int out_w, int out_h,
                                 uint8_t* out, void* stream);

/* classify_position + process_position_probabilities (core.py:225-249, 310-355) on boards u8[N,512,512]. */
CVB_API int cvb_classify(cvb_ctx* ctx, const uint8_t* board, int N, int flip, float* probs, uint8_t* labels,
                 uint8_t* labels_valid, char* fen, void* stream);

/* process_image for a batch (core.py:152-195); all pointers on the device. */
CVB_API int cvb_image_to_fen(cvb_ctx* ctx, const uint8_t* img, int N, float threshold, int flip, const cvb_outputs* out,
                     void* stream);

/* Inputs of any size H x W >= 256 x 256 (process_image accepts every u8[H,W,3], core.py:168-170,212): img u8[N,H,W,3], all
 * images of a call share one size.  512 x 512 takes the fused path above; other sizes go through the general INTER_AREA
 * reduction (cvb_resize_area), the quadrangle is scaled by H/256 on both axes (core.py:414-417) and the board is warped
 * from the original image.  The cell tables and two staging images are (re)built when the size changes. */
CVB_API int cvb_image_to_fen_hw(cvb_ctx* ctx, const uint8_t* img, int N, int H, int W, float threshold, int flip, const cvb_outputs* out,
                                void* stream);
CVB_API int cvb_unet_forward_hw(cvb_ctx* ctx, const uint8_t* img, int N, int H, int W, float threshold, float* logits, uint8_t* mask,
                                void* stream);
/* cv2.resize(img, (256,256), interpolation=INTER_AREA) (core.py:212) for any H, W >= 256, bit-identical to OpenCV (integer
 * cell averages for integer scale factors, its float32 accumulation order otherwise): u8[N,H,W,3] -> u8[N,256,256,3]. */
CVB_API int cvb_resize_area(cvb_ctx* ctx, const uint8_t* img, int N, int H, int W, uint8_t* out, void* stream);

/* Same, from HOST buffers to HOST buffers (pinned memory recommended): chunks of max_batch boards are copied in,
 * processed and copied out on three streams so that PCIe transfers overlap compute.  Synchronous on return. */
CVB_API int cvb_image_to_fen_host(cvb_ctx* ctx, const uint8_t* img_host, int N, float threshold, int flip,
                          const cvb_outputs* out_host);
/* Same; *boards_done (host memory, may be NULL) is advanced to the number of leading boards whose results have landed in
 * out_host, so that another host thread can consume finished groups while the rest of the batch is still in flight
 * (ChessVision.process_images builds its result objects that way). */
CVB_API int cvb_image_to_fen_host_progress(cvb_ctx* ctx, const uint8_t* img_host, int N, float threshold, int flip,
                                           const cvb_outputs* out_host, volatile int32_t* boards_done);

/* Building block exposed for parity tests: one convolution layer on fp16 NHWC device tensors through the tcgen05
 * kernel.  ksize in {1,3} (pad = ksize/2), stride in {1,2}; w_packed fp16 [Cout][ksize*ksize*Cin] with
 * k = (r*ksize+s)*Cin + ci; bias fp32 [Cout]; residual (optional) fp16 [N,Ho,Wo,Cout].  Cin % 64 == 0. */
CVB_API int cvb_conv2d_f16(cvb_ctx* ctx, const void* in, int N, int H, int W, int Cin, const void* w_packed, const float* bias,
                   int Cout, int ksize, int stride, int relu, const void* residual, void* out, void* stream);
/* ConvTranspose2d(k=2, s=2): w_packed fp16 [4*Cout][Cin] with row = (dy*2+dx)*Cout + co; bias fp32 [4*Cout]
 * (the per-channel bias repeated for the four taps); out fp16 [N,2H,2W,out_c_stride] at channel out_c_off. */
CVB_API int cvb_convt2x2_f16(cvb_ctx* ctx, const void* in, int N, int H, int W, int Cin, const void* w_packed, const float* bias,
                     int Cout, void* out, int out_c_stride, int out_c_off, void* stream);

/* conv3x3(Cin -> 128, pad 1) + bias + ReLU followed by ConvTranspose2d(128 -> 64, k2, s2) + bias in ONE kernel (the UNet's
 * up3.conv.double_conv.3 + up4.up, unet_parts.py:16-21,53): the 128-channel intermediate stays on chip.  Operand layouts as in
 * cvb_conv2d_f16 / cvb_convt2x2_f16 (bias2: fp32 [Cout2], Cout2 = 64); bit-identical to calling those two. */
CVB_API int cvb_conv3x3_convt2x2_f16(cvb_ctx* ctx, const void* in, int N, int H, int W, int Cin, const void* w_packed, const float* bias,
                                     const void* w2_packed, const float* bias2, int Cout2, void* out, int out_c_stride, int out_c_off,
                                     void* stream);

/* First layers alone (parity tests): fused preprocessing + first convolution of each network on tcgen05.
 * cvb_unet_stem:   img u8[N,512,512,3] -> fp16 NHWC [N,256,256,64]  = ReLU(BN(conv3x3(resize_area(img)/255)))  (core.py:212-216,
 *                  unet_parts.py:16-18);  cvb_resnet_stem: board u8[N,512,512] -> fp16 NHWC [N*64,16,16,64] = maxpool3x3s2(ReLU(BN(
 *                  conv7x7s2(square/255)))) per square (core.py:232-237, timm resnet18 conv1..maxpool).  Need loaded weights. */
CVB_API int cvb_unet_stem(cvb_ctx* ctx, const uint8_t* img, int N, void* out, void* stream);
CVB_API int cvb_resnet_stem(cvb_ctx* ctx, const uint8_t* board, int N, void* out, void* stream);

/* ---------------------------------------------------------------------------------------------------------------------
 * UNet board-extractor training step (BASELINE.json configs[4]; reference scripts/train/train_unet.py:236-245,293-323,
 * chessvision/pytorch_unet/utils/dice_score.py:5-30).  fp16 tensor-core operands, fp32 master weights / accumulation /
 * optimizer state (what the reference's `--amp` autocast path computes), BatchNorm in training mode with local batch
 * statistics.  Data-parallel use: forward_backward on every rank, all-reduce (sum) the flat gradient buffer over NCCL,
 * then optimizer_step with grad_scale = 1/world_size.
 * ------------------------------------------------------------------------------------------------------------------- */
typedef struct cvb_train_config {
    int32_t batch;        /* images per step on this GPU (fixed; scripts/bin/train_board_extractor.sh uses 2)           */
    float loss_scale;     /* static factor applied to the fp16 activation gradients (GradScaler's role), default 4096   */
    float momentum;       /* RMSprop momentum 0.999 (train_unet.py:240)                                                  */
    float alpha;          /* RMSprop alpha 0.99, eps 1e-8 (torch defaults)                                               */
    float eps;
    float weight_decay;   /* 1e-8 (train_unet.py:239)                                                                    */
    float max_grad_norm;  /* clip_grad_norm_ threshold 1.0 (train_unet.py:321)                                           */
    float bn_momentum;    /* nn.BatchNorm2d defaults 0.1 / 1e-5                                                          */
    float bn_eps;
} cvb_train_config;

CVB_API int cvb_train_default_config(cvb_train_config* cfg);
/* UNet(3,1) in model.train() state from a state_dict (same tensors cvb_load_unet takes); cfg NULL = reference defaults. */
CVB_API int cvb_train_create(cvb_ctx* ctx, const cvb_tensor* state_dict, int n_tensors, const cvb_train_config* cfg);
/* masks_pred = model(images); loss = BCEWithLogits + dice_loss; loss.backward()  (train_unet.py:309-319).
 * img fp32 [B,3,256,256] NCHW, mask fp32 [B,1,256,256] in {0,1}, loss: 1 float; all device pointers. */
CVB_API int cvb_train_forward_backward(cvb_ctx* ctx, const float* img, const float* mask, float* loss, void* stream);
/* The flat fp32 gradient buffer written by forward_backward (packed layout; only sums/norms are layout independent). */
CVB_API int cvb_train_grads(cvb_ctx* ctx, float** grads, int64_t* count);
/* Gradient buckets for overlapping the data-parallel all-reduce with the backward pass: bucket b is the range [lo[b], hi[b])
 * of the flat gradient buffer, numbered in the order forward_backward completes them (layer-reverse: the decoder's gradients
 * first).  Returns the number of buckets (lo / hi may be NULL).  cvb_train_bucket_wait makes `stream` (the communication
 * stream) wait until the last kernel of the most recent forward_backward that writes into bucket b has finished, so that
 * its all-reduce can start while the compute stream is still working on the earlier layers. */
CVB_API int cvb_train_buckets(cvb_ctx* ctx, int64_t* lo, int64_t* hi, int capacity);
CVB_API int cvb_train_bucket_wait(cvb_ctx* ctx, int bucket, void* stream);
/* clip_grad_norm_(params, max_grad_norm); optimizer.step()  (train_unet.py:321-322) on grads * grad_scale. */
CVB_API int cvb_train_optimizer_step(cvb_ctx* ctx, float lr, float grad_scale, void* stream);
/* Single-GPU convenience: forward_backward + optimizer_step(grad_scale 1). */
CVB_API int cvb_train_step(cvb_ctx* ctx, const float* img, const float* mask, float lr, float* loss, void* stream);
/* Copy parameters + BatchNorm running statistics (what = 0) or gradients (what = 1) back in torch state_dict layout:
 * out[i].name / shape select the tensor, out[i].data must point to WRITABLE host memory. */
CVB_API int cvb_train_export(cvb_ctx* ctx, int what, const cvb_tensor* out, int n_tensors);
/* Building block exposed for parity tests: weight gradient of Conv2d(3x3, pad 1) on the tcgen05 MN-major kernel.
 * dz fp16 [N,H,H,Cout], x fp16 [N,H,H,Cin] (dense NHWC) -> dw fp32 [Cout][9][Cin] = scale * sum_p dz[p] (x) x[p+tap]. */
CVB_API int cvb_wgrad3x3_f16(cvb_ctx* ctx, const void* dz, const void* x, int N, int H, int W, int Cout, int Cin, float scale,
                             float* dw, void* stream);

/* ---------------------------------------------------------------------------------------------------------------------
 * Piece-classifier training step (reference scripts/train/train_classifier.py:63-126,218-221):
 * timm resnet18(num_classes=13, in_chans=1) on 64x64 squares, CrossEntropyLoss, torch.optim.Adam; fp32 throughout, BatchNorm
 * with batch statistics in training mode and running statistics in eval mode (the validation loop, :91-113).
 * ------------------------------------------------------------------------------------------------------------------- */
typedef struct cvb_cls_train_config {
    int32_t batch;        /* squares per step on this GPU (fixed at creation)                    */
    float beta1, beta2;   /* Adam (0.9, 0.999), eps 1e-8, weight_decay 0: torch.optim.Adam's defaults, train_classifier.py:219 */
    float eps;
    float weight_decay;
    float bn_momentum;    /* nn.BatchNorm2d defaults 0.1 / 1e-5                                   */
    float bn_eps;
} cvb_cls_train_config;

CVB_API int cvb_cls_train_default_config(cvb_cls_train_config* cfg);
/* Model + optimizer state from a state_dict (the tensors cvb_load_resnet18 takes); cfg NULL = defaults with batch 64. */
CVB_API int cvb_cls_train_create(cvb_ctx* ctx, const cvb_tensor* state_dict, int n_tensors, const cvb_cls_train_config* cfg);
/* output = model(data) [+ loss, number of correct predictions]: data fp32 [B,1,64,64], target i32 [B] (may be NULL: no loss),
 * training != 0 = model.train() (batch statistics, running statistics updated), 0 = model.eval().  loss (1 float), correct
 * (1 int32), logits ([B,13]) are device pointers and may be NULL. */
CVB_API int cvb_cls_train_forward(cvb_ctx* ctx, const float* data, const int32_t* target, int training, float* loss, int32_t* correct,
                                  float* logits, void* stream);
/* forward in training mode + loss.backward() (train_classifier.py:77-80): the gradients land in the flat buffer below. */
CVB_API int cvb_cls_train_forward_backward(cvb_ctx* ctx, const float* data, const int32_t* target, float* loss, int32_t* correct, void* stream);
/* The flat fp32 gradient buffer (for a data-parallel all-reduce; conv weights are stored [co][r][q][ci]). */
CVB_API int cvb_cls_train_grads(cvb_ctx* ctx, float** grads, int64_t* count);
/* optimizer.step() (Adam) on grads * grad_scale with learning rate lr (the caller applies StepLR, train_classifier.py:220). */
CVB_API int cvb_cls_train_optimizer_step(cvb_ctx* ctx, float lr, float grad_scale, void* stream);
/* forward_backward + optimizer_step(grad_scale 1): the body of the reference's loop for one batch. */
CVB_API int cvb_cls_train_step(cvb_ctx* ctx, const float* data, const int32_t* target, float lr, float* loss, int32_t* correct, void* stream);
/* what = 0: parameters + BatchNorm running statistics, 1: gradients, 2 / 3: Adam exp_avg / exp_avg_sq -- in torch state_dict
 * layout; out[i].name / shape select the tensor, out[i].data must point to WRITABLE host memory. */
CVB_API int cvb_cls_train_export(cvb_ctx* ctx, int what, const cvb_tensor* out, int n_tensors);
CVB_API int64_t cvb_cls_train_steps(const cvb_ctx* ctx);

/* ---------------------------------------------------------------------------------------------------------------------
 * Consumers of the per-board outputs, computed on the device buffers the pipeline wrote.
 * ------------------------------------------------------------------------------------------------------------------- */
/* scripts/eval/evaluate.py:37-52,109-140 for N boards.  probs f32 [N,64,13]; labels / labels_valid u8 [N,64] (either may
 * be NULL); true_labels u8 [N,64] = class index of the ground-truth FEN in FEN order a8..h1 (evaluate.py:61-86, 12 = empty);
 * topk_hits i32 [N,k]: squares whose true class is among the i+1 most probable (compute_model_topk_accuracy * 64, ties
 * ranked like np.argsort read from the end); correct i32 [N,2] (may be NULL): squares on which original_fen / fen agree
 * with the truth (compute_position_accuracy.num_correct); flip = the orientation the labels were produced with. */
CVB_API int cvb_eval_metrics(cvb_ctx* ctx, const float* probs, const uint8_t* labels, const uint8_t* labels_valid,
                             const uint8_t* true_labels, int N, int flip, int k, int32_t* topk_hits, int32_t* correct, void* stream);
/* scripts/process_new_raw/process_pipeline.py:357-467 for N boards.  values f32 [N,L] (the reference passes
 * BoardExtractionResult.probabilities, L = 65536); quad f32 [N,4,2] (may be NULL) with found u8 [N] (may be NULL);
 * scores f64 [N,4] = {quadrangle_regularity, mask_completeness (process_pipeline.py:380-414; defined for L = 256*256,
 * NaN otherwise), probability_distribution, probability_confidence}. */
CVB_API int cvb_quality_scores(cvb_ctx* ctx, const float* values, const float* quad, const uint8_t* found, int N, int L,
                               double* scores, void* stream);

/* ---------------------------------------------------------------------------------------------------------------------
 * JPEG decode front-end: cv2.imread / cv2.imdecode(buf, IMREAD_COLOR) in front of process_image
 * (scripts/eval/evaluate.py:147, app/computeroot/cv_endpoint.py:151-153), bit-identical to OpenCV's libjpeg-turbo
 * defaults (islow IDCT, fancy upsampling, BGR).  Huffman decoding runs on the host (one image per thread), the inverse
 * DCT, chroma upsampling and colour conversion on the device.  Supported: baseline, 8-bit, 4:2:0, dimensions multiples
 * of 16, restart intervals, EXIF orientation absent or 1; anything else returns -6 with the reason in cvb_last_error.
 * ------------------------------------------------------------------------------------------------------------------- */
/* Host only: dimensions of one JPEG stream (0 ok, -6 unsupported / corrupt). */
CVB_API int cvb_jpeg_info(const uint8_t* data, int64_t nbytes, int32_t* h, int32_t* w);
/* Host only: the entropy-decoding half alone (tests): quantised coefficients int16 [Y blocks (H/8 x W/8, row-major) | Cb
 * blocks | Cr blocks][64] in natural order (H*W*3/2 values) and the three quantisation tables u16 [3][64] (may be NULL). */
CVB_API int cvb_jpeg_coefficients(const uint8_t* data, int64_t nbytes, int16_t* coef, uint16_t* qt);
/* data[i] / nbytes[i]: N JPEG streams in HOST memory, all H x W; img: DEVICE u8 [N,H,W,3] BGR.  The host buffers may be
 * released on return; the device work is enqueued on `stream`.  Staging buffers are sized on first use. */
CVB_API int cvb_decode_jpeg(cvb_ctx* ctx, const uint8_t* const* data, const int64_t* nbytes, int N, int H, int W, uint8_t* img,
                            void* stream);

/* Number of kernels launched by this context so far (bench.py reports it as gpu_launches). */
CVB_API int64_t cvb_launch_count(const cvb_ctx* ctx);
/* Pipeline passes of at most one chunk (max_batch boards) on a capturable stream are captured once into a CUDA graph and
 * replayed (ChessVision.process_image on one board is ~50 launches); this counts the replays.  CVB_NO_GRAPH=1 disables it. */
CVB_API int64_t cvb_graph_replays(const cvb_ctx* ctx);

/* Time (ms, CUDA events on the launching stream) accumulated per stage since the last reset; stage ids:
 * 0 unet convs (tcgen05), 1 unet aux (stem, pools), 2 mask->quad, 3 homography+warp, 4 resnet stem,
 * 5 resnet convs (tcgen05), 6 head.  Enabled with cvb_profile(ctx, 1); adds event records, so off by default. */
CVB_API int cvb_profile(cvb_ctx* ctx, int enable);
CVB_API int cvb_profile_read(cvb_ctx* ctx, float* ms_out, int n_stages);

#ifdef __cplusplus
}
#endif
#endif /* CHESSVISION_B200_H */
