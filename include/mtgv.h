/*
 * mtgv.h - C ABI of libmtgv.so, the B200 (sm_100a) synthetic-sample generator that
 * replaces the cv2/numpy hot path of nmichlo/mtg-vision.
 *
 * The reference has no FFI of its own: its seam is three Python surfaces
 * (SyntheticBgFgMtgImages statics, RanMtgEncDecDataset, Gen - SURVEY.md section 8b).
 * The Python drop-ins in mtgvision_b200/ keep those surfaces and bind the entry points
 * below with ctypes.  Each entry point cites the reference code it replaces
 * (paths relative to the reference root, file:line).
 *
 * Conventions
 *   - every function returns 0 on success or a negative mtgv_status; the message for
 *     the last failure on a context is returned by mtgv_last_error();
 *   - all image/label/tape/param pointers are DEVICE pointers owned by the caller
 *     (PyTorch tensors) unless the parameter name ends in `_host`;
 *   - work is enqueued on the caller's CUDA stream (`stream` is a cudaStream_t passed
 *     as void*); the library never synchronises except in mtgv_create/destroy and the
 *     pool setters;
 *   - one mtgv_ctx per device; calls on one ctx are not re-entrant; different ctxs are
 *     independent (one process per GPU, no collectives: samples are independent).
 *   - sizes are (H, W) like the reference (mtgvision/util/image.py:295).
 */
#ifndef MTGV_H_
#define MTGV_H_

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define MTGV_ABI_VERSION 1

typedef struct mtgv_ctx mtgv_ctx;

typedef enum {
  MTGV_OK = 0,
  MTGV_ERR_INVALID = -1,   /* bad argument / unsupported configuration */
  MTGV_ERR_CUDA = -2,      /* CUDA runtime error (message has the cudaError string) */
  MTGV_ERR_STATE = -3,     /* pools / config not set */
  MTGV_ERR_LIMIT = -4      /* a documented size limit was exceeded */
} mtgv_status;

typedef enum { MTGV_OUT_F16 = 0, MTGV_OUT_U8 = 1, MTGV_OUT_F32 = 2 } mtgv_out_dtype;

/* ------------------------------------------------------------------------------------ */
/* Tape: the sampled choices and magnitudes of ONE augmented sample, exactly the values  */
/* the reference draws from `random` / `np.random`.  In parity     */
/* tests the tape is recorded from the reference/oracle run; in production it is written  */
/* by mtgv_sample_encoder_tape() from a Philox counter stream.                            */
/* ------------------------------------------------------------------------------------ */

typedef enum {
  MTGV_OP_NONE = 0,
  MTGV_OP_DOWNUP = 1,       /* Mutate.downscale_upscale  encoder_datasets.py:142-163
                               i[0]=n, i[1]=interp_down, i[2]=interp_up (cv2 codes 0/1/2)   */
  MTGV_OP_WARP = 2,         /* Mutate.warp               :94-111   d[0..7]=np.random.rand(4,2) */
  MTGV_OP_AFFINE = 3,       /* Mutate.affine_transform   :353-375  d[0]=angle d[1]=tx d[2]=ty
                               d[3]=scale d[4]=shear; i[0]=1 -> d[5],d[6] hold the host-libm
                               alpha=cos*scale, beta=sin*scale of getRotationMatrix2D         */
  MTGV_OP_PERSPECTIVE = 4,  /* Mutate.perspective_transform :377-403  d[0..7]=uniform(-.1,.1) */
  MTGV_OP_TINT = 5,         /* Mutate.tint               :165-171  d[0..2]=np.random.random()  */
  MTGV_OP_FADE_BLACK = 6,   /* Mutate.fade_black         :180-185  d[0]=np.random.random()     */
  MTGV_OP_FADE_WHITE = 7,   /* Mutate.fade_white         :173-178  d[0]=np.random.random()     */
  MTGV_OP_BC = 8,           /* Mutate.brightness_contrast :187-193 d[0]=contrast draw, d[1]=brightness draw */
  MTGV_OP_FLIP = 9,         /* Mutate.flip               :73-80    i[0]=horr, i[1]=vert         */
  MTGV_OP_ROTATE = 10,      /* Mutate.rotate_bounded     :82-87 + util/image.py:380-398
                               d[0]=np.random.random() (deg=360*d[0]); i[0]=1 -> d[1],d[2]=alpha,beta */
  MTGV_OP_WARP_INV = 11,    /* Mutate.warp_inv           :113-116  d[0..7]=np.random.rand(4,2) */
  MTGV_OP_BLUR = 12,        /* Mutate.blur               :136-140  i[0]=ksize (1 or 3)          */
  MTGV_OP_SHARPEN = 13,     /* Mutate.sharpen            :242-247                               */
  MTGV_OP_NOISE = 14,       /* Mutate.noise              :118-134  i[0]=kind 0 speckle 1 gaussian
                               2 salt&pepper 3 poisson; d[0]=ratio draw; field (+field2)       */
  MTGV_OP_GAUSS_NOISE = 15, /* Mutate.gaussian_noise     :222-226  field                        */
  MTGV_OP_SALT_PEPPER = 16, /* Mutate.salt_pepper_noise  :228-240  field=salt, field2=pepper    */
  MTGV_OP_ERASE = 17,       /* Mutate.random_erasing     :273-351  d[0]=scale d[1]=aspect d[2]=flip draw,
                               i[0]=stage (0 returned before centre draw, 1 empty rect, 2 filled),
                               i[1]=cx i[2]=cy i[3]=colour mode (0 random 1 uniform_random 2 zeros
                               3 ones 4 mean), d[3..5]=uniform_random colour; field=random block;
                               i[4]=1 -> i[5],i[6]=host-computed block_w, block_h              */
  MTGV_OP_CUTOUT = 18       /* Mutate.cutout             :259-271  i[2k],i[2k+1]=(y,x) of hole k */
} mtgv_tape_opcode;

#define MTGV_FIELD_PHILOX (-1) /* `field` value: generate the random field on device */

typedef struct {
  int32_t code;
  int32_t n_field;  /* salt count (SALT_PEPPER / NOISE kind 2) */
  int32_t n_field2; /* pepper count */
  int32_t _pad;
  int32_t i[16];
  double d[8];
  int64_t field;  /* offset (in 4-byte words) into the `fields` buffer, or MTGV_FIELD_PHILOX */
  int64_t field2;
} mtgv_tape_op; /* 160 bytes */

#define MTGV_TAPE_MAX_OPS 18

typedef enum {
  MTGV_KIND_VIRTUAL = 0, /* make_virtual  (encoder_datasets.py:786-813) */
  MTGV_KIND_CROPPED = 1, /* make_cropped  (:733-753)                    */
  MTGV_KIND_BG_ONLY = 2  /* make_bg       (:774-784): only the bg ops of the tape are used */
} mtgv_sample_kind;

typedef struct {
  int32_t kind;        /* mtgv_sample_kind: make_virtual (encoder_datasets.py:786) or
                          make_cropped (:733) when the target-is-input draw fired
                          (encoder_train.py:178)                                        */
  int32_t card;        /* base card pool index (ran_card, encoder_datasets.py:662)      */
  int32_t swap_choice; /* hard negative (encoder_train.py:217-221): index drawn by
                          random.choice over the same-name group minus self, -1 = none  */
  int32_t bg;          /* background pool index                                         */
  int32_t upsidedown;  /* ApplyChoice(upsidedown, None) result (encoder_datasets.py:801) */
  int32_t n_fg, n_bg, n_vrtl; /* ops[] holds fg ops, then bg ops, then vrtl ops, each in
                                 the order the reference executed them                   */
  uint64_t seed;       /* Philox key for device-generated fields                         */
  mtgv_tape_op ops[MTGV_TAPE_MAX_OPS];
} mtgv_enc_tape;

/* ------------------------------------------------------------------------------------ */
/* Params: the tape expanded into kernel-ready form (inverse homographies, fixed-point    */
/* affine, integer geometry, float32 coefficients).  Opaque to callers except for tests.  */
/* ------------------------------------------------------------------------------------ */

typedef enum {
  MTGV_X_NONE = 0,
  MTGV_X_ELEM = 1,        /* y = a*x + b (separate roundings), optional clip; f[0..3]=a, f[4..7]=b per
                             channel RGBA, i[0]=channel mask, i[1]=clip                            */
  MTGV_X_DOWNUP = 2,      /* i[0]=n i[1]=down i[2]=up                                              */
  MTGV_X_WARP_PERSP = 3,  /* d[0..8] = cv::invert(M) as used by cv::warpPerspective                */
  MTGV_X_WARP_AFFINE = 4, /* d[0..5] = inverted 2x3 as used by cv::warpAffine                      */
  MTGV_X_BLUR3 = 5,
  MTGV_X_SHARPEN = 6,
  MTGV_X_NOISE = 7,       /* i[0]=kind, f[0]=ratio f[1]=1-ratio                                    */
  MTGV_X_GAUSS_NOISE = 8,
  MTGV_X_SALT_PEPPER = 9,
  MTGV_X_ERASE = 10,      /* i[0..3]=y0,y1,x0,x1 i[4]=mode f[0..2]=colour                          */
  MTGV_X_CUTOUT = 11      /* i[0..15]                                                              */
} mtgv_x_opcode;

typedef struct {
  int32_t code;
  int32_t n_field, n_field2;
  int32_t _pad;
  int32_t i[16];
  float f[8];
  double d[9];
  int64_t field, field2;
} mtgv_x_op; /* 200 bytes */

#define MTGV_X_MAX_OPS 16

typedef struct {
  int32_t kind, card, bg, upsidedown;
  int32_t out_h, out_w;
  int32_t card_h, card_w;
  /* foreground / cropped geometry: crop_to_size(pad=True) (util/image.py:349-377) or
     remove_border_resized (:337-346) */
  int32_t src_y0, src_x0, src_h, src_w; /* source window of the card that is area-resized */
  int32_t fg_rh, fg_rw, fg_y0, fg_x0;   /* resized size and paste offset inside (out_h,out_w) */
  /* background chain: flip -> rotate_bounded -> warp_inv -> crop_to_size */
  int32_t bg_h, bg_w, flip_h, flip_v;
  int32_t rot_nh, rot_nw;
  int32_t bg_rh, bg_rw, bg_y0, bg_x0;
  int32_t n_fg, n_pre, n_post, n_vrtl; /* ops[]: fg | bg elementwise before G0 | after G0 | vrtl */
  double rot_inv[6];
  double winv[9];
  uint64_t seed;
  int32_t status; /* 0 ok, else mtgv_status raised by the expansion of this sample */
  int32_t _pad;
  mtgv_x_op ops[MTGV_X_MAX_OPS];
} mtgv_enc_params;

typedef struct {
  int32_t out_h, out_w;          /* x_size_hw (encoder_train.py:98)                     */
  int32_t y_h, y_w;              /* y_size_hw (:99)                                      */
  double target_is_input_prob;   /* :101 */
  double similar_neg_prob;       /* :102 */
  int32_t half_upsidedown;       /* :100 */
  int32_t paired;                /* :96  */
  int32_t targets;               /* :97  */
  int32_t _pad;
} mtgv_enc_config;

/* ------------------------------------------------------------------------------------ */
/* Context and pools                                                                     */
/* ------------------------------------------------------------------------------------ */

int mtgv_abi_version(void);
int mtgv_sizeof(int which); /* sizeof of 0 mtgv_tape_op, 1 mtgv_enc_tape, 2 mtgv_x_op, 3 mtgv_enc_params, 4 mtgv_enc_config */
mtgv_ctx* mtgv_create(int device);
void mtgv_destroy(mtgv_ctx* ctx);
const char* mtgv_last_error(const mtgv_ctx* ctx);

/* Card pool: replaces SyntheticBgFgMtgImages._load_card_image / _init_ label tables
 * (encoder_datasets.py:557-590, 633-634).  cards: [N,H,W,3] uint8 RGB (device).  The
 * library keeps its own planar copy.  labels3: [N,3] int32 (id,name,set ranks, :586-590);
 * grp_off [N+1], grp_mem: CSR of same-name groups in insertion order (:570, :619-630).
 * Also builds the static rounded-rect alpha (make_masked :755-772, round_rect_mask
 * util/image.py:406-425) for (H,W). */
int mtgv_set_card_pool(mtgv_ctx* ctx, const uint8_t* cards, int n, int h, int w, const int32_t* labels3,
                       const int32_t* grp_off, const int32_t* grp_mem, int n_mem);

/* Background pool: replaces IlsvrcImages._load_image (encoder_datasets.py:457-474).
 * bgs: concatenated HWC uint8 images (device); offsets[n] byte offsets; hw[n][2].  The library keeps
 * its own copy as RGBX words (one 4-byte load per bilinear tap). */
int mtgv_set_bg_pool(mtgv_ctx* ctx, const uint8_t* bgs, const int64_t* offsets_host, const int32_t* hw_host, int n);

int mtgv_set_encoder_config(mtgv_ctx* ctx, const mtgv_enc_config* cfg_host);

/* ------------------------------------------------------------------------------------ */
/* Encoder path                                                                          */
/* ------------------------------------------------------------------------------------ */

/* Production sampler: writes the tapes of `n_pairs` x-samples followed by `n_pairs`
 * x2-samples (when paired) for global sample indices [first_index, first_index+n_pairs),
 * drawing what RanMtgEncDecDataset._random_image_batch/_make_image_batch draw
 * (encoder_train.py:149-156,189-230) from Philox4x32-10 keyed by (seed, index). */
int mtgv_sample_encoder_tape(mtgv_ctx* ctx, uint64_t seed, int64_t first_index, int n_pairs, mtgv_enc_tape* tape,
                             void* stream);

/* Same, for explicit cards (RanMtgEncDecDataset.image_batch_by_ids, encoder_train.py:122-139):
 * cards [n_pairs] device int32 pool indices (NULL = draw them); the two probabilities
 * override the configured ones when >= 0 (force_target_input / force_similar_neg). */
int mtgv_sample_encoder_tape_ex(mtgv_ctx* ctx, uint64_t seed, int64_t first_index, int n_pairs, const int32_t* cards,
                                const int32_t* bgs, double target_is_input_prob, double similar_neg_prob,
                                mtgv_enc_tape* tape, void* stream);

/* Streaming ingest (the per-batch `_load_card_image` / `ilsvrc.ran()` of _make_image_batch,
 * encoder_train.py:149-156,196): overwrite pool entries [first, first+n) in place with new
 * HWC uint8 images (device) of the entries' existing size; asynchronous on `stream`.
 * With explicit `bgs` in mtgv_sample_encoder_tape_ex, x uses bgs[i] and x2 uses bgs[slot]
 * for a uniformly drawn slot (bg1 = random.choice(bg_imgs), :224). */
int mtgv_update_card_images(mtgv_ctx* ctx, const uint8_t* cards, int first, int n, void* stream);
int mtgv_update_bg_images(mtgv_ctx* ctx, const uint8_t* bgs, int first, int n, void* stream);

/* Static rounded-rectangle mask of the card pool (round_rect_mask, util/image.py:406-425):
 * which = 0 encoder (radius_ratio 0.05, encoder_datasets.py:763), 1 detection (0.046,
 * od_datasets.py:223).  out: [card_h, card_w] float32 (device). */
int mtgv_get_mask(mtgv_ctx* ctx, int which, float* out, void* stream);

/* Subsystem (1): tape -> params.  Replaces cv2.getPerspectiveTransform / getRotationMatrix2D /
 * the matrix inversions inside cv2.warp* / crop_to_size geometry for every sample
 * (encoder_datasets.py:94-111,353-403; util/image.py:349-398).  Also resolves the
 * hard-negative card (encoder_datasets.py:619-630) and writes labels [n,3] int64
 * (card_get_labels :586-590) when `labels` is not NULL. */
int mtgv_expand_params(mtgv_ctx* ctx, const mtgv_enc_tape* tape, int n, mtgv_enc_params* params, int64_t* labels,
                       void* stream);

/* Subsystems (2)(3)(5): runs n samples; out is [n,3,out_h,out_w] NCHW of out_dtype.
 * Replaces make_virtual / make_cropped + np.stack + image_to_tensor
 * (encoder_datasets.py:733-813, encoder_train.py:143,228).  `fields` may be NULL when
 * every field offset is MTGV_FIELD_PHILOX. */
int mtgv_encoder_batch(mtgv_ctx* ctx, const mtgv_enc_params* params, int n, void* out, int out_dtype,
                       const void* fields, void* stream);

/* make_cropped targets y (encoder_train.py:171-173): out [n,3,y_h,y_w] for cards[n] (device int32). */
int mtgv_encoder_targets(mtgv_ctx* ctx, const int32_t* cards, int n, void* out, int out_dtype, void* stream);

/* Parity entry: cv2.warpPerspective(src f32 HWC, M, (dw,dh), INTER_LINEAR, BORDER_CONSTANT 0)
 * for n images (encoder_datasets.py:111,403; od_datasets.py:82).  src [n,sh,sw,c] f32,
 * M [n,9] f64 (forward matrix, inverted on device like cv2), dst [n,dh,dw,c] f32. */
int mtgv_warp_perspective(mtgv_ctx* ctx, const float* src, int n, int sh, int sw, int c, const double* M, float* dst,
                          int dh, int dw, void* stream);

/* Parity/debug entry: run a list of expanded ops on float32 HWC images with the same
 * plane interpreter the batch kernel uses.  img [n,h,w,c] f32 in/out (c = 3 or 4). */
int mtgv_run_plane_ops(mtgv_ctx* ctx, float* img, int n, int h, int w, int c, const mtgv_x_op* ops, int n_ops,
                       const void* fields, uint64_t seed, void* stream);

/* ------------------------------------------------------------------------------------ */
/* Detection path (mtgvision/od_datasets.py)                                             */
/* ------------------------------------------------------------------------------------ */

#define MTGV_DET_MAX_CARDS 32    /* np.random.randint(num_cards_min, num_cards_max) - 1 <= 32 (BASELINE config 4) */
#define MTGV_DET_MAX_ATTEMPTS 10 /* card_max_place_attempts (od_datasets.py:633)                                 */
#define MTGV_DET_MAX_KP 8        /* points per keypoint polygon: 4 (obb) or 8 (seg)                              */
#define MTGV_DET_MAX_KPOLY 3     /* polygons per card: 3 (obb: card, top, bottom) or 1 (seg)                     */
#define MTGV_DET_MAX_PRE 4
#define MTGV_DET_MAX_POST 8
#define MTGV_DET_MAX_CARD_OPS 3

/* Photometric ops: the subset of the albumentations graphs at od_datasets.py:420-512 named by the
 * north star.  Parameters are the values after sampling (alpha/beta, shifts, sigma, rectangle). */
typedef enum {
  MTGV_PH_NONE = 0,
  MTGV_PH_RBC = 1,         /* RandomBrightnessContrast: clip(x*d[0] + d[1])                          */
  MTGV_PH_HSV = 2,         /* HueSaturationValue: d[0] hue shift (deg), d[1] sat shift, d[2] val shift (in 1/255) */
  MTGV_PH_GAUSS_NOISE = 3, /* GaussNoise: clip(x + d[0]*N(0,1)); field = unit normals [H,W,3] f32 or PHILOX */
  MTGV_PH_GAUSS_BLUR = 4,  /* GaussianBlur: d[0] sigma; ksize = max(3, int(6 sigma + 1) | 1), REFLECT_101    */
  MTGV_PH_ERASE = 5,       /* Erasing: i[0]=top i[1]=left i[2]=h i[3]=w i[4]=fill (0 random,1 random_uniform,
                              2 ones,3 zeros), d[0..2]=uniform colour; field = random block [h,w,3] f32    */
  /* the remaining members of the noise / blur families of get_bg_transform (od_datasets.py:443-457), SURVEY 8f.3 */
  MTGV_PH_ISO_NOISE = 6,   /* ISONoise: d[0] color_shift, d[1] intensity.  In cv2's float HLS space: hue += N(0,1) * d[0]*360*d[1]
                              (mod 360), L += Poisson(std(L)*d[1]*255)/255 * (1 - L); std(L) = cv2.meanStdDev of the image the
                              op receives.  field = [H,W,2] f32 (Poisson counts, unit normals) or PHILOX               */
  MTGV_PH_SHOT_NOISE = 7,  /* ShotNoise: d[0] scale.  clip(Poisson((x^2.2 + d[0]*1e-6)/d[0]) * d[0])^(1/2.2);
                              field = Poisson counts [H,W,3] f32 or PHILOX                                            */
  MTGV_PH_MEDIAN_BLUR = 8, /* MedianBlur: i[0] = ksize in {3,5,7}; cv2.medianBlur of rint(255 x) as uint8 (edge replicated), / 255 */
  MTGV_PH_MOTION_BLUR = 9, /* MotionBlur: i[0] = ksize (odd, <= 11), i[1..4] = bit mask of the line kernel (bit y*ksize+x);
                              cv2.filter2D with mask / popcount(mask), BORDER_REFLECT_101                              */
  MTGV_PH_GLASS_BLUR = 10  /* GlassBlur (mode "fast"): d[0] sigma, i[0] max_delta (<= 8), i[1] iterations (<= 2): GaussianBlur(sigma),
                              per iteration every interior pixel swaps with the neighbour at its (dy,dx) in [-max_delta, max_delta)
                              (numpy's simultaneous fancy assignment, later index wins), GaussianBlur(sigma).
                              field = int32 [(H-2md)*(W-2md)][iterations][2] in albumentations' pixel order, or PHILOX     */
} mtgv_photo_code;

typedef struct {
  int32_t code;
  int32_t i[5];
  double d[3];
  int64_t field; /* word offset into `fields` or MTGV_FIELD_PHILOX */
} mtgv_photo_op; /* 56 bytes */

typedef struct {
  int32_t cx, cy;     /* random.randint centre (od_datasets.py:322-323)                              */
  int32_t dst_given;  /* 1: dst[] holds the float32 corner targets computed by the host (numpy
                         norm/arctan2/cos/sin + getRotationMatrix2D; the parity boundary)  */
  int32_t _pad;
  double deg;         /* np.random.uniform(0, 360)                        (:325)                    */
  double area;        /* exp(np.random.uniform(log a_min, log a_max))     (:332)                    */
  double jitter[4];   /* np.random.uniform(1-j, 1+j, size=4)              (:37)                     */
  float dst[8];       /* dst_pts.astype(float32)                          (:348)                    */
} mtgv_det_attempt;   /* 96 bytes */

typedef struct {
  int32_t card;       /* card pool index (mtg_ds.ran_path, :560) */
  int32_t n_attempts; /* attempts available in att[] */
  int32_t n_photo;
  int32_t _pad;
  mtgv_photo_op photo[MTGV_DET_MAX_CARD_OPS]; /* pre_transform_card, applied only if the card is placed (:581) */
  mtgv_det_attempt att[MTGV_DET_MAX_ATTEMPTS];
} mtgv_det_card;

typedef struct {
  int32_t bg_only; /* Gen.random took the ratio_bg branch (:686-687) */
  int32_t bg;      /* background pool index */
  int32_t bg_deg;  /* np.random.randint(0, 360) of make_background (:200) */
  int32_t n_cards; /* np.random.randint(num_cards_min, num_cards_max) (:558) */
  int32_t n_pre, n_post;
  uint64_t seed;
  int32_t bg_ab_given; /* 1: bg_ab holds the host-libm cos/sin of bg_deg (getRotationMatrix2D, :106) */
  int32_t _pad;
  double bg_ab[2];
  mtgv_photo_op pre[MTGV_DET_MAX_PRE];    /* pre_transform_bg  (get_bg_transform_light) */
  mtgv_photo_op post[MTGV_DET_MAX_POST];  /* post_transform_bg (get_bg_transform)       */
  mtgv_det_card cards[MTGV_DET_MAX_CARDS];
} mtgv_det_tape;

typedef struct {
  int32_t size_h, size_w;                 /* bg_size_hw (od_datasets.py:623)             */
  int32_t num_cards_min, num_cards_max;   /* :624-625 (max exclusive, like np.random.randint) */
  double min_visible;                     /* card_min_visible_ratio          :626 */
  double min_visible_edges;               /* card_min_visible_ratio_edges    :627; < 0 = None */
  double jitter_ratio;                    /* :628 */
  double min_area_ratio, max_area_ratio;  /* :629-630 */
  double ratio_bg;                        /* :634; 0 = None */
  int32_t no_contains;                    /* :632 */
  int32_t max_attempts;                   /* :633 */
  int32_t kind;                           /* 0 obb, 1 seg  (:638) */
  int32_t photometrics;                   /* 0 disables the albumentations stages (parity tests) */
  int32_t size_sample_mode;               /* card_size_sample_mode :631, :327-332: 0 log_uniform, 1 uniform */
  int32_t n_bgs_first;                    /* the background pool = bg_ds (slots [0, n_bgs_first)) followed by bg2_ds; 0 = one dataset */
  double bg_first_prob;                   /* ilsvrc_vs_coco_sample_weights normalised (:656-672): P(dataset = bg_ds); the image is then
                                             drawn uniformly inside the dataset.  Ignored when n_bgs_first is 0 or the whole pool. */
} mtgv_det_config;

int mtgv_set_det_config(mtgv_ctx* ctx, const mtgv_det_config* cfg_host);

/* Production sampler: draws what Gen.random / generate_synthetic_image draw for scenes
 * [first_index, first_index+n) from Philox (od_datasets.py:520-611, 685-704). */
int mtgv_sample_det_tape(mtgv_ctx* ctx, uint64_t seed, int64_t first_index, int n, mtgv_det_tape* tape, void* stream);

/* Subsystem (4) + (1): placement rejection sampling and label warping on device.
 * Replaces place_card_on_background_get_transform (:287-377, shapely tests restated with convex
 * clipping), cv2.getPerspectiveTransform per attempt (:347), apply_transform_2d on keypoints
 * (:352,597) and get_rotate_over_output_transform (:85-118).
 *   params    [n] opaque scene programs for mtgv_det_batch (mtgv_det_params_size() bytes each)
 *   accepted  [n, MAX_CARDS] int32: index of the accepted attempt per card, -1 = not placed
 *   keypoints [n, MAX_CARDS*MAX_KPOLY, MAX_KP, 2] float64, in the reference's output order
 *             (cards in reverse placement order, :594-601), pixel units
 *   labels    [n, MAX_CARDS*MAX_KPOLY] int32 (keypoints_labels), -1 padding
 *   counts    [n] int32 number of keypoint polygons */
int mtgv_det_place(mtgv_ctx* ctx, const mtgv_det_tape* tape, int n, void* params, int32_t* accepted, double* keypoints,
                   int32_t* labels, int32_t* counts, void* stream);
int mtgv_det_params_size(void);

/* Subsystems (2)(3)(5): background cover-warp, pre-augments, per-card warp of image + mask with
 * alpha composite in reverse placement order, post-augments, cast.  Replaces make_background
 * (:195-203), apply_transform_2d_img (:73-82), the composite loop (:594-601) and the
 * albumentations calls (:552,581,604).  images: [n,3,S,S] NCHW of out_dtype. */
int mtgv_det_batch(mtgv_ctx* ctx, const void* params, int n, void* images, int out_dtype, const void* fields, void* stream);

/* ------------------------------------------------------------------------------------ */
/* Serving-side dewarp (mtgvision/od_export.py), SURVEY 8f.4                              */
/* ------------------------------------------------------------------------------------ */

/* InstanceSeg.extract_dewarped (od_export.py:95-111) for n detected cards of one frame:
 * M_k = cv2.getPerspectiveTransform(quads[k] float32 (4 corners x,y), dst_rect float32 (4 corners of
 * the expanded output rectangle, computed by the caller like :102-104)), then
 * cv2.warpPerspective(frame uint8 HWC with c = 1..4 channels, M_k, (ow, oh)) - INTER_LINEAR, constant 0 border,
 * uint8 fixed-point blend.  out: [n, oh, ow, c] uint8.  Needs no pool or config.  Bit-exact. */
int mtgv_extract_dewarped(mtgv_ctx* ctx, const uint8_t* frame, int frame_h, int frame_w, int channels, const float* quads,
                          int n, const float* dst_rect, uint8_t* out, int out_h, int out_w, void* stream);

/* ------------------------------------------------------------------------------------ */
/* Image decode into the pools (mtgvision/util/image.py:107-114), SURVEY 8f.1             */
/* ------------------------------------------------------------------------------------ */

/* Frame size of one JPEG file (host memory): hw[0] = height, hw[1] = width.  Host-only marker walk; fails with
 * MTGV_ERR_INVALID and a message for files the decoder does not support (arithmetic coding, CMYK, 12-bit, chroma
 * subsampled by more than 2).  Baseline, extended sequential and progressive Huffman files are supported. */
int mtgv_jpeg_info(mtgv_ctx* ctx, const uint8_t* file, int64_t len, int32_t* hw);

/* mtgv_jpeg_info for n files in ONE call: file i = files[file_off[i] .. file_off[i+1]) (host memory, file_off has n+1
 * entries), hw: host [n][2].  The marker walks run on a few host threads; what a loader does before it sizes the
 * output of mtgv_decode_jpeg_batch (the reference opens one file at a time: IlsvrcImages._load_image,
 * encoder_datasets.py:457-474).  The first unsupported or damaged file fails the call, its index is in the message. */
int mtgv_jpeg_info_batch(mtgv_ctx* ctx, const uint8_t* files, const int64_t* file_off, int n, int32_t* hw);

/* n separate host buffers (the `bytes` a loader got from n open().read() calls: the reference reads one file per drawn
 * image, IlsvrcImages._load_image encoder_datasets.py:457-474, util/image.py:107-114) copied back to back into `dst`
 * (host memory, pinned for the upload that follows) by a few host threads; file_off (host, n+1 entries) receives the
 * offsets mtgv_decode_jpeg_batch / mtgv_decode_jpeg_to_pools take.  dst_cap: bytes available at dst.  Host-only: ctx may
 * be NULL (it only receives the error message). */
int mtgv_gather_files(mtgv_ctx* ctx, const uint8_t* const* srcs, const int64_t* lens, int n, uint8_t* dst, int64_t dst_cap, int64_t* file_off);

/* imread_float's cv2.imread(path, IMREAD_COLOR_RGB) (util/image.py:107-114; IlsvrcImages._load_image
 * encoder_datasets.py:457-474) for n baseline JPEG files at once, on the device.
 * files: HOST bytes, file i = files[file_off[i] .. file_off[i+1]) (file_off: host, n+1 entries).
 * out: DEVICE uint8; image i is written as [h,w,3] RGB at out + out_off[i] (out_off: host, n entries) - with
 *      out_off = cumulative 3*h*w this is exactly the buffer mtgv_set_bg_pool takes, with equal-size files and
 *      out_off = i*H*W*3 the tensor mtgv_set_card_pool / mtgv_update_*_images take.
 * hw: host [n][2] as returned by mtgv_jpeg_info (what the caller sized `out` with); a file whose frame header
 *     disagrees fails the call.  Bit-exact with cv2.imdecode (libjpeg-turbo: ISLOW IDCT, fancy upsampling).
 *     Progressive (SOF2) files take the same call: their scans are entropy-decoded on the host threads that walk the markers
 *     (refinement scans are bit-serial per block), the coefficients are uploaded, and dequantisation / IDCT / upsampling /
 *     colour conversion run in the same device kernels as for baseline files.
 * Work is queued on `stream`.  file_off, out_off and hw may be reused when the call returns; `files` is read by an
 * asynchronous copy when it is page-locked memory and must then stay valid until that work has run (pageable memory is
 * staged before the call returns). */
int mtgv_decode_jpeg_batch(mtgv_ctx* ctx, const uint8_t* files, const int64_t* file_off, int n, uint8_t* out,
                           const int64_t* out_off, const int32_t* hw, void* stream);

/* The loader step of one batch in one call: what `_random_image_batch` does with `ran_card` -> dl_and_open_im_resized and
 * `IlsvrcImages.ran` -> imread_float (encoder_train.py:149-156, encoder_datasets.py:470-474, 634), for files instead of
 * arrays.  Files [0, n_cards) are decoded STRAIGHT INTO card pool slots [first_card, first_card + n_cards) (planar layout),
 * files [n_cards, n_cards + n_bgs) into background slots [first_bg, ...) (RGBX layout) - mtgv_decode_jpeg_batch followed
 * by mtgv_update_card_images / mtgv_update_bg_images without the intermediate HWC image.  Every file's frame size must
 * equal its slot's image size (the pools keep their geometry), else MTGV_ERR_INVALID.  Same pixel values as
 * mtgv_decode_jpeg_batch (bit-exact with cv2.imdecode); `files` lifetime as there.  Work is queued on `stream`. */
int mtgv_decode_jpeg_to_pools(mtgv_ctx* ctx, const uint8_t* files, const int64_t* file_off, int n_cards, int first_card,
                              int n_bgs, int first_bg, void* stream);

/* Device time of the three kernels (entropy decode, inverse DCT, upsample + colour) of the last
 * mtgv_decode_jpeg_batch call, measured with CUDA events on its stream; waits for that batch (bench bookkeeping). */
int mtgv_jpeg_last_kernel_ms(mtgv_ctx* ctx, float* ms3);

/* ------------------------------------------------------------------------------------ */
/* Image encode for the on-disk dataset writer (mtgvision/od_datasets.py:794-832), 8f.2   */
/* ------------------------------------------------------------------------------------ */

#define MTGV_LAYOUT_NHWC 0 /* [n,h,w,3] */
#define MTGV_LAYOUT_NCHW 1 /* [n,3,h,w] (mtgv_det_batch's output) */

/* save_sample's imwrite (od_datasets.py:829-831; util/image.py:95-104: cv2.imwrite of the uint8 RGB image) for n
 * images at once, on the device: baseline JPEG, 4:2:0, standard Huffman tables, JFIF header - the FILE BYTES of
 * cv2.imencode(".jpg", bgr, [IMWRITE_JPEG_QUALITY, quality]) (cv2.imwrite's default quality is 95).
 * images: device uint8 RGB in `layout`, any size up to 16384 per side (sizes that are not whole 16x16 MCUs follow libjpeg's
 * edge-replication and dummy-block rules; save_sample itself asserts 640x640).
 * out: device, image i's file at out + i*cap (cap: bytes per image, multiple of 4); out_len: device int32 [n], the
 * file length, or -1 when the file does not fit in cap (nothing usable is written for that image). */
int mtgv_encode_jpeg_batch(mtgv_ctx* ctx, const uint8_t* images, int n, int h, int w, int layout, int quality, uint8_t* out,
                           int64_t cap, int32_t* out_len, void* stream);

/* Packs the files of mtgv_encode_jpeg_batch back to back so that one transfer brings them to the host: offsets (device
 * int64 [n+1]) receives the exclusive prefix sum of the file lengths (files that did not fit count as empty), compact
 * (device, at least the sum of the lengths; n*cap always suffices) file i at compact + offsets[i]. */
int mtgv_compact_jpeg_files(mtgv_ctx* ctx, const uint8_t* slots, int64_t cap, const int32_t* out_len, int n, uint8_t* compact,
                            int64_t* offsets, void* stream);

/* Device time of the two kernels (colour + DCT + quantisation, Huffman coding + stuffing) of the last
 * mtgv_encode_jpeg_batch call, CUDA events on its stream; waits for that batch (bench bookkeeping). */
int mtgv_jpeg_encode_last_kernel_ms(mtgv_ctx* ctx, float* ms2);

/* Number of kernels launched by this context since creation (bench bookkeeping). */
int64_t mtgv_launch_count(const mtgv_ctx* ctx);

#ifdef __cplusplus
}
#endif
#endif /* MTGV_H_ */
