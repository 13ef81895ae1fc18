/* yoloret_b200 C-ABI  -  the drop-in boundary of the B200 YOLO-ReT hot path.
 *
 * The reference (prakharg24/yoloret) is pure Python on TensorFlow and has no
 * FFI of its own; its hot path runs inside TF ops.  This header
 * is the seam the replacement creates: the Python host that mirrors the
 * reference call surface (yoloret_b200/yolo.py, yoloret_b200/yolo3/model.py)
 * binds these entry points with ctypes.  Each entry point cites the reference
 * code whose TF ops it replaces.
 *
 * Conventions
 *  - plain C: raw DEVICE pointers + sizes, no torch types; the caller owns all
 *    memory including workspaces; nothing here allocates or synchronises.
 *  - every call is stream-ordered on `stream` (a cudaStream_t passed as void*)
 *    and may be captured into a CUDA graph.
 *  - returns 0 on success, a negative yr_status otherwise; yr_last_error()
 *    returns a thread-local message for the last failure.
 *  - activations are NHWC fp32.  Every channel dimension is padded to a
 *    multiple of 8 ("Cp") and the pad channels hold exact zeros; `ld` is the
 *    element stride between consecutive pixels of a buffer (>= Cp), which lets
 *    a producer write straight into a channel slice of a concat buffer
 *    (Concatenate layers of reference code/yolo3/model.py:164-166,255,275,308,321
 *    are never materialised by a copy).
 */
#ifndef YOLORET_B200_H_
#define YOLORET_B200_H_

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef enum yr_status {
    YR_OK = 0,
    YR_ERR_INVALID = -1,   /* bad argument (shape, alignment, null pointer) */
    YR_ERR_UNSUPPORTED = -2,
    YR_ERR_CUDA = -3,      /* a CUDA runtime call / launch failed */
    YR_ERR_WORKSPACE = -4  /* caller-provided workspace too small */
} yr_status;

typedef enum yr_act { YR_ACT_NONE = 0, YR_ACT_RELU6 = 1, YR_ACT_SWISH = 2 } yr_act;

typedef enum yr_op_kind {
    YR_OP_STEM = 0,      /* dense 3x3 stride-2 conv, Cin=3 (+folded BN +act) */
    YR_OP_PW = 1,        /* 1x1 conv = GEMM (+folded BN/bias +act +residual +SE scale) */
    YR_OP_DW = 2,        /* depthwise kxk conv, k in {3,5}, stride in {1,2} (+folded BN +act) */
    YR_OP_RESAMPLE = 3,  /* nearest x2 up-sampling / 2x2 / 4x4 max-pooling into a channel slice */
    YR_OP_RFCR = 4,      /* fused RFCR fusion: 4x 1x1 conv + resize + weighted sum */
    YR_OP_SE = 5,        /* squeeze-excite gate: global mean -> FC -> swish -> FC -> sigmoid */
    YR_OP_SE_FC = 6,     /* the same gate from the channel sums a DW op left in `aux` (fused squeeze) */
    /* 7: retired (the round-1 expand+depthwise+project kernel, superseded by DWPW) */
    YR_OP_DWPW = 8       /* fused 3x3 depthwise (+BN +act) -> 1x1 conv (+BN +act +residual): the depthwise output never
                            exists in HBM */
} yr_op_kind;

typedef enum yr_resample_mode { YR_UP2 = 0, YR_POOL2 = 1, YR_POOL4 = 2 } yr_resample_mode;

/* One network layer.  A plan is an array of these executed in order by
 * yr_run_ops().  Fields not used by a kind are ignored.
 *
 *  STEM     in  [B,H,W,3] (u8 if in_is_u8 else f32; u8 is scaled by 1/255 as
 *               tf.io.decode_image(dtype=float32) does, reference code/yolo.py:106)
 *           w   [3*3*3][Cp_out] (kh,kw,cin major; BN scale folded), bias [Cp_out]
 *           out [B,Ho,Wo,ld_out]; TF 'SAME' padding (pad_t/pad_l = leading pad).
 *           replaces Conv1+bn_Conv1+ReLU6 of tf.keras.applications.MobileNetV2
 *           (reference code/yolo3/override.py:339) and the EfficientNet stem
 *           (code/yolo3/efficientnet.py:636-645).
 *  PW       in  [B*H*W rows][ld_in], C = K (multiple of 8)
 *           w   [K][N] row-major (N multiple of 8, BN scale folded), bias [N]
 *           res optional [rows][ld_res] added AFTER the activation-less BN
 *               (MobileNetV2 block_N_add / MBConv id_skip, efficientnet.py:527-533)
 *           scale optional [B][K]: SE gate multiplied into the A operand
 *               (Multiply layer of SEBlock, efficientnet.py:435)
 *           out [rows][ld_out].  Replaces every Conv2D(kernel_size=1) of
 *           reference code/yolo3/model.py:98-114,152-155,243-247,263-267,299-318
 *           and efficientnet.py:485-491,517-522.
 *  DW       in [B,H,W,ld_in] C channels; w [k*k][C] (BN folded), bias [C];
 *           out [B,Ho,Wo,ld_out].  DepthwiseConv2D of MobileNetV2 blocks,
 *           efficientnet.py:501-506 and model.py:20-23.
 *  RESAMPLE in [B,H,W,ld_in] C channels -> out [B,Ho,Wo,ld_out] (mode):
 *           UpSampling2D / MaxPooling2D of model.py:139-144,157,164-166,254,274,307,320.
 *  RFCR     in=b1 [B,H/2,W/2,ld_in] C=K1, in2=b2 [B,H,W,ld_in2] K2,
 *           in3=b3 [B,2H,2W,ld_in3] K3, in4=b4 [B,4H,4W,ld_in4] K4 (un-pooled tap);
 *           w = [K1+K2+K3+K4][N] stacked 1x1 kernels (N=48), bias = alpha[4];
 *           out[b,h,w,:] = a0*W1.b1[h/2,w/2] + a1*W2.b2[h,w]
 *                        + a2*max_{2x2}(W3.b3) + a3*W4.max_{4x4}(b4)
 *           (H,W = output grid).  rfcr_module + WeightedSum, model.py:117-157,190.
 *           aux optional [B][yr_dw_se_slots(op)][C]: per-CTA sums of the op's OUTPUT over its pixels
 *               (the squeeze of a following SEBlock, fused; deterministic, no atomics).
 *  SE_FC    in = the aux buffer of the producing DW op [B][K2 slots][C=F]; H,W = spatial size of the
 *           squeezed tensor; w = [w1^T R x F | w2 R x F] (first FC TRANSPOSED), bias = b1[R] then b2[F];
 *           N = R; out = gate [B][F].
 *  DWPW     in [B,H,W,ld_in] C channels (multiple of 8); k=3, stride 1|2, pad_t/pad_l = leading TF-SAME pads;
 *           mode = yr_act of the depthwise conv; N = output channels of the 1x1 conv (<= 192), act = its activation;
 *           bias = the 1x1 conv's bias [N]; res optional [B,Ho,Wo,ld_res]; w_tc = yr_dwpw_pack image (both layers'
 *           folded weights); out [B,Ho,Wo,ld_out].  Bit-identical to a DW op followed by a PW op (variant 2/3).
 *           One kernel for block_N_depthwise .. block_N_project(+add) of tf.keras.applications.MobileNetV2
 *           (reference code/yolo3/override.py:339-341) and the SE-less MBConv blocks of efficientnet.py:501-533.
 *  SE       in [B,H,W,ld_in] C=F channels; w = [F][R] then [R][F] (w2 = w + F*R),
 *           bias = b1[R] then b2[F]; N = R; out = gate [B][F].
 *           SEBlock, efficientnet.py:406-438.
 */
typedef struct yr_op {
    int32_t kind;      /* yr_op_kind */
    int32_t act;       /* yr_act */
    int32_t mode;      /* yr_resample_mode (RESAMPLE); yr_act of the depthwise conv (DWPW) */
    int32_t in_is_u8;  /* STEM */
    int32_t B, H, W, C;      /* input batch / height / width / channels (K) */
    int32_t Ho, Wo, N;       /* output height / width / channels */
    int32_t k, stride, pad_t, pad_l;
    int32_t ld_in, ld_in2, ld_in3, ld_in4, ld_out, ld_res;
    int32_t K2, K3, K4;      /* RFCR: channels of in2..in4.
                                PW (variant 3 / 4), stacked outputs: two 1x1 convs that read the SAME tensor (reference
                                code/yolo3/model.py:296-305: the y conv and the next bottom-up conv of a head stage) run as one
                                GEMM over [W1 | W2]: K2 > 0 = the first conv's (padded) column count - columns [0, K2) go to
                                `out` (row stride ld_out), columns [K2, N) to `aux` (row stride ld_in2), K2 % 4 == 0; K3 != 0 =
                                the first conv is linear (no activation) while `act` applies to the second */
    int32_t variant;         /* PW kernel choice: 0 = auto (tcgen05 when w_tc is given), 1 = SIMT fp32,
                                2 = tcgen05 3xTF32, A and B operands in shared memory (w_tc = yr_pw_tc_pack image),
                                3 = tcgen05 3xTF32, A operand in tensor memory (w_tc = yr_pw_ts_pack image),
                                4 = the same as a CTA PAIR (cluster of 2, tcgen05.mma.cta_group::2, M = 256: each CTA holds half
                                    of every weight tile; same yr_pw_ts_pack image; see yr_pw_ts2_supported) */
    const void* in;
    const void* in2;
    const void* in3;
    const void* in4;
    void* out;
    const float* w;
    const float* bias;
    const float* res;
    const float* scale;
    const float* w_tc;       /* PW: weight image made by yr_pw_tc_pack (variant 0/2) or yr_pw_ts_pack (variant 3), or NULL */
    float* aux;              /* DW: squeeze-excite partial sums (see above), or NULL; PW with K2 > 0: the second output */
} yr_op;

/* Library identity / errors. */
int yr_version(void);
const char* yr_last_error(void);
int yr_sizeof_op(void); /* sizeof(yr_op), for binding self-checks */
int yr_dw_se_slots(const yr_op* op); /* slots per image of a DW op's aux buffer (negative = invalid op) */

/* Executes ops[0..n_ops) in order on `stream`.  The network forward
 * (reference yolov3_body, code/yolo3/model.py:170-342, called from
 * YoloModel.call code/yolo.py:152) is one call of this. */
int yr_run_ops(const yr_op* ops, int n_ops, void* stream);

/* Tensor-core pointwise variant: the 1x1 kernel W [K][N] (row-major, BN scale folded, the same
 * array yr_op.w points to) is split into TF32 (hi, lo) halves and stored in the swizzled
 * shared-memory image the tcgen05 kernel bulk-copies.  Done once per layer at weight-load time
 * (replaces model.load_weights for these layers, reference code/yolo.py:87).
 *   yr_pw_tc_packed_floats  number of floats `packed` must hold (0 = shape unsupported)
 *   yr_pw_tc_pack           packed must be 128-byte aligned */
int64_t yr_pw_tc_packed_floats(int K, int N);
int yr_pw_tc_pack(const float* w, int K, int N, float* packed, void* stream);
/* Same for the variant-3 kernel (its n tiles are at most 192 columns wide, so the image differs). */
int64_t yr_pw_ts_packed_floats(int K, int N);
int yr_pw_ts_pack(const float* w, int K, int N, float* packed, void* stream);
int yr_pw_ts2_supported(int K, int N); /* 1 when variant 4 (CTA pair) has a tiling for this layer */

/* Fused depthwise -> pointwise pair (YR_OP_DWPW): w_pw [K][N] (the PW op's `w`), w_dw [9][K] + b_dw [K] (the DW op's
 * `w` / `bias`; K = the depthwise channel count = the 1x1 conv's input channels).  yr_dwpw_packed_floats returns 0
 * when N needs more than one n tile (> 192); yr_dwpw_supported says whether a given layer geometry has a fused
 * tiling (the caller otherwise runs the two ops separately). */
int64_t yr_dwpw_packed_floats(int K, int N);
int yr_dwpw_pack(const float* w_pw, int K, int N, const float* w_dw, const float* b_dw, float* packed, void* stream);
int yr_dwpw_supported(int C, int N, int stride, int Ho, int Wo);
/* Host-only query of the plan the fused kernel would run (tests, diagnostics): plan[16] = {tile rows, tile cols, box
 * rows, box cols, tiles per image (rows), (cols), converter groups, epilogue groups, box ring, TMEM stage ring, weight
 * ring, accumulators, weights resident, n tile width, k-blocks, dynamic shared memory bytes}; returns 1 / 0 as
 * yr_dwpw_supported. */
int yr_dwpw_plan(int C, int N, int stride, int Ho, int Wo, int32_t* plan);

/* ---- post-process: yolo_eval (reference code/yolo3/model.py:431-491) --------- */

typedef struct yr_decode_params {
    int32_t B;            /* images */
    int32_t num_classes;  /* C */
    int32_t num_scales;   /* 1..3, scales given coarse -> fine (s32, s16, s8) */
    int32_t grid_h[3], grid_w[3];
    int32_t ld[3];        /* element stride between grid cells of feats[s] (>= 3*(C+5)) */
    float anchors[3][3][2]; /* per scale, per anchor: (w, h) in pixels - already masked
                               with [[6,7,8],[3,4,5],[0,1,2]][-num_scales:] (model.py:444-445) */
    int32_t input_h, input_w; /* network input = grid(s32) * 32 (model.py:449) */
    float score_threshold;
    int32_t cand_cap;     /* capacity of each (image,class) candidate list */
} yr_decode_params;

/* yolo_head + yolo_correct_boxes + yolo_boxes_and_scores (model.py:344-428) for a
 * batch of independent images, fused with the score-threshold filter that
 * tf.image.non_max_suppression applies internally.
 *   feats[s]      [B,gh,gw,ld[s]] raw logits, cell layout (anchor, 5+C)
 *   image_shapes  [B][2] float (h, w) of the ORIGINAL images (model.py:380-385)
 *   boxes         [B][total_boxes][4] (ymin,xmin,ymax,xmax) image pixels
 *   cand_score / cand_index [B][C][cand_cap]; cand_count [B][C] (zeroed by this call)
 * Candidate (b,c) lists hold every box with conf*cls > score_threshold (strict),
 * in unspecified order; counts saturate at cand_cap and the overflow is
 * reported by yr_nms_classwise (never silently truncated). */
int yr_decode_filter(const float* const feats[3], const float* image_shapes, const yr_decode_params* p,
                     float* boxes, float* cand_score, int32_t* cand_index, int32_t* cand_count, void* stream);

/* yolo_head (model.py:344-371) as a standalone op of the public call surface.
 *   feats [B,gh,gw,ld] raw logits (cell layout (anchor, 5+C)), anchors_dev [A][2] (w,h px) on the device
 *   box_xy / box_wh [B,gh,gw,A,2], box_confidence [B,gh,gw,A,1], box_class_probs [B,gh,gw,A,C] (NULL when the
 *   caller wants calc_loss=True outputs), grid [gh,gw,1,2] (x,y) or NULL. */
int yr_yolo_head(const float* feats, int ld, int B, int gh, int gw, int A, int C, const float* anchors_dev,
                 int input_h, int input_w, float* box_xy, float* box_wh, float* box_confidence,
                 float* box_class_probs, float* grid, void* stream);

/* Class-wise greedy NMS == the `for c in range(num_classes):
 * tf.image.non_max_suppression(...)` loop of model.py:474-486, bit-exact
 * selection (score desc, ties -> lower index; IoU > thr strict; TF IoU formula).
 *   det      [B][C][max_boxes][6] (ymin,xmin,ymax,xmax,score,box_index as float bits)
 *   det_count[B][C]
 *   status   [1] int32, set to 1 if any candidate list had overflowed cand_cap.
 * cand_score is consumed (overwritten). */
int yr_nms_classwise(const float* boxes, int total_boxes, float* cand_score, const int32_t* cand_index,
                     const int32_t* cand_count, int B, int num_classes, int cand_cap, int max_boxes,
                     float iou_threshold, float* det, int32_t* det_count, int32_t* status, void* stream);

/* Concatenates per-class results class-major like model.py:487-490.
 *   out_boxes_f [B][C*max_boxes][4] float, out_boxes_i same as int32 (truncated, model.py:490),
 *   out_scores [B][C*max_boxes], out_classes [B][C*max_boxes] int32, out_count [B]. */
int yr_pack_detections(const float* det, const int32_t* det_count, int B, int num_classes, int max_boxes,
                       float* out_boxes_f, int32_t* out_boxes_i, float* out_scores, int32_t* out_classes,
                       int32_t* out_count, void* stream);

/* ---- pre-process: letterbox_image (reference code/yolo3/utils.py:67-83) ---------
 * src u8 [ih,iw,3] -> dst f32 [h,w,3]: (1/255) scaling, bilinear half-pixel
 * resize to (nh,nw), zero pad at (dy,dx).  nh/nw/dy/dx are computed by the
 * host exactly as utils.py:76-79 does (float64). */
int yr_letterbox_u8(const uint8_t* src, int ih, int iw, float* dst, int h, int w, int nh, int nw, int dy,
                    int dx, void* stream);

/* ---- training input: preprocess_true_boxes (reference code/yolo3/utils.py:298-376) ---
 * boxes [B][T][5] device f32 (xmin, ymin, xmax, ymax, class) in input pixels, zero-width rows = padding;
 * anchors: HOST pointer to 9 (w,h) pairs; y_true: HOST array of num_scales DEVICE pointers, each
 * [B][gh][gw][3][5+num_classes] (gh = round(input_h / {32,16,8}[l])); the tensors are cleared here.
 * Same slot-collision and row-indexing behaviour as the reference's sequential loop. */
int yr_encode_true_boxes(const float* boxes, int B, int T, const float* anchors, int input_h, int input_w,
                         int num_classes, int num_scales, float* const* y_true, void* stream);

/* ---- training: YoloLoss (reference code/yolo3/model.py:585-671) ------------------
 * One scale.  logits / y_true [B,gh,gw,A,5+C] with cell stride ld_logits / ld_true.
 * true_boxes [n_true][4] = tf.boolean_mask(true_box, object_mask) over the WHOLE
 * batch (model.py:643), produced by yr_yolo_loss_gather_true.
 *   loss_parts [4]: giou, confidence, class sums (already / B) and sum(ignore_mask)
 *   dlogits    same layout as logits (may be NULL for forward only):
 *              d(loss)/d(logits) of giou+conf+class (ignore mask is a constant,
 *              as in TF where tf.cast(bool) blocks the gradient).
 * workspace: partials, >= yr_yolo_loss_workspace(...) bytes. */
typedef struct yr_loss_params {
    int32_t B, gh, gw, A, C;
    int32_t ld_logits, ld_true; /* per-cell stride in elements, >= A*(5+C) */
    float anchors[3][2];
    int32_t input_h, input_w;
    float ignore_thresh;
    int32_t max_true;           /* capacity of true_boxes */
} yr_loss_params;

int64_t yr_yolo_loss_workspace(const yr_loss_params* p);
int yr_yolo_loss_gather_true(const float* y_true, const yr_loss_params* p, float* true_boxes,
                             int32_t* n_true, void* stream);
int yr_yolo_loss(const float* logits, const float* y_true, const float* true_boxes, const int32_t* n_true,
                 const yr_loss_params* p, float* loss_parts, float* dlogits, void* workspace,
                 int64_t workspace_bytes, void* stream);

/* ---- training, fast path: sparse y_true + the loss of ALL scales in one launch --------------------------
 * Replaces the same reference code as yr_encode_true_boxes + num_scales x yr_yolo_loss (preprocess_true_boxes,
 * code/yolo3/utils.py:298-376; YoloLoss.call, code/yolo3/model.py:607-671; the sum over scales, code/yolo3/train.py:11-16)
 * without materialising the dense y_true tensors (> 99.9 % zeros) - SURVEY.md section 8f-2.
 *
 * Sparse y_true of a batch:
 *   slot_maps[l]  int32 [B][gh_l][gw_l][3]: index of the slot's record in scale l's record list, or -1
 *   records       float [3][max_records][8]: (x, y, w, h) normalised centre form = y_true[..., 0:4], then 4 words of
 *                 class bits (bit c of word c / 32 = y_true[..., 5 + c]); num_classes <= 128
 *   counts        int32 [3]: records per scale (max_records >= B * T can never overflow)
 * yr_encode_true_boxes_sparse takes the same inputs as yr_encode_true_boxes and fills the three (it clears the
 * maps and counts itself).  slot_maps is a HOST array of DEVICE pointers; anchors a HOST pointer to 9 (w, h) pairs.
 *
 * yr_yolo_loss3: logits[l] / dlogits[l] [B][gh_l][gw_l][ld_logits[l]] (3 anchors x (5 + C) values per cell; HOST
 * arrays of DEVICE pointers, dlogits may be NULL for forward only).
 *   loss_parts [3][4]: per scale giou, confidence, class sums (already / B) and sum(ignore_mask); the reference's loss
 *                      is the sum of the first three over the scales.
 *   workspace  >= yr_yolo_loss3_workspace(p) bytes, ZEROED once before the first call (it holds the arrival counter
 *              the kernel re-arms itself); one launch, no host synchronisation, CUDA-graph capturable. */
typedef struct yr_loss3_params {
    int32_t B, C, num_scales;
    int32_t gh[3], gw[3], ld_logits[3];
    float anchors[3][3][2]; /* per scale, the 3 anchors (w, h) in input pixels: anchors[anchor_mask[l][k]] */
    int32_t input_h, input_w;
    float ignore_thresh;
    int32_t max_records;
} yr_loss3_params;

int yr_encode_true_boxes_sparse(const float* boxes, int B, int T, const float* anchors, int input_h, int input_w,
                                int num_classes, int num_scales, int32_t* const* slot_maps, float* records,
                                int32_t* counts, int max_records, void* stream);
int64_t yr_yolo_loss3_workspace(const yr_loss3_params* p);
int yr_yolo_loss3(const float* const* logits, const int32_t* const* slot_maps, const float* records,
                  const int32_t* counts, const yr_loss3_params* p, float* loss_parts, float* const* dlogits,
                  void* workspace, int64_t workspace_bytes, void* stream);

/* ---- training: optimizer step (reference code/train.py:158-160,195-197: tf.keras.optimizers.Adam(lr, epsilon=1e-8)) ---
 * In-place Adam update of a flat fp32 parameter vector (or a rank's shard of it) with TF's ApplyAdam arithmetic:
 *   alpha = lr * sqrt(1 - beta2^step) / (1 - beta1^step);  m += (g - m)(1 - beta1);  v += (g*g - v)(1 - beta2);
 *   param -= (m * alpha) / (sqrt(v) + epsilon).     step is 1-based.  lr comes from the host (epoch-wise CosineDecay,
 * code/train.py:92-100: yoloret_b200.train.cosine_decay). */
int yr_adam_step(float* param, const float* grad, float* m, float* v, int64_t n, float lr, float beta1, float beta2,
                 float epsilon, int64_t step, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* YOLORET_B200_H_ */
