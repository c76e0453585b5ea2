// mtgv_expand.cuh - subsystem (1): tape -> kernel-ready parameters for one encoder sample.
// Host/device so the same code is unit-tested on the CPU (tests/host_harness.cpp) against
// the matrices the reference's cv2 calls produced, and runs in k_expand on the GPU.
#pragma once

#include "../../include/mtgv.h"
#include "mtgv_geom.cuh"

namespace mtgv {

constexpr double kPiOver180 = 3.141592653589793238462643383279502884 / 180.0;  // CV_PI/180

struct PoolMeta {
  int card_h, card_w, n_cards, n_bgs;
  const int32_t* labels3;  // [n_cards,3]
  const int32_t* grp_off;  // [n_cards+1]
  const int32_t* grp_mem;
  const int32_t* bg_hw;    // [n_bgs,2]
};

MTGV_HD void x_op_clear(mtgv_x_op* o) {
  o->code = MTGV_X_NONE;
  o->n_field = o->n_field2 = 0;
  o->_pad = 0;
  for (int k = 0; k < 16; k++) o->i[k] = 0;
  for (int k = 0; k < 8; k++) o->f[k] = 0.f;
  for (int k = 0; k < 9; k++) o->d[k] = 0.0;
  o->field = o->field2 = MTGV_FIELD_PHILOX;
}

// y = a*x + b elementwise ops.  `nchan_mask` is the set of channels the reference touches:
// tint / fades write img[:, :, :3]; brightness_contrast rewrites the whole array.
MTGV_HD bool expand_elementwise(const mtgv_tape_op* t, mtgv_x_op* o) {
  x_op_clear(o);
  o->code = MTGV_X_ELEM;
  for (int c = 0; c < 4; c++) {
    o->f[c] = 1.f;
    o->f[4 + c] = 0.f;
  }
  switch (t->code) {
    case MTGV_OP_TINT:  // r = 1 + 0.15*(2u-1); img[:,:,i] = clip(r*img[:,:,i])
      for (int c = 0; c < 3; c++)
        o->f[c] = (float)MTGV_DADD(1.0, MTGV_DMUL(0.15, MTGV_DSUB(MTGV_DMUL(2.0, t->d[c]), 1.0)));
      o->i[0] = 7;
      o->i[1] = 1;
      return true;
    case MTGV_OP_FADE_BLACK: {  // ratio*0 + (1-ratio)*img
      double ratio = MTGV_DMUL(t->d[0], 0.5);
      for (int c = 0; c < 3; c++) o->f[c] = (float)MTGV_DSUB(1.0, ratio);
      o->i[0] = 7;
      o->i[1] = 0;
      return true;
    }
    case MTGV_OP_FADE_WHITE: {  // ratio*1 + (1-ratio)*img
      double ratio = MTGV_DMUL(t->d[0], 0.33);
      for (int c = 0; c < 3; c++) {
        o->f[c] = (float)MTGV_DSUB(1.0, ratio);
        o->f[4 + c] = (float)ratio;
      }
      o->i[0] = 7;
      o->i[1] = 0;
      return true;
    }
    case MTGV_OP_BC: {  // alpha*img + beta, clip; alpha = 1.0 + U(-c, c)
      double alpha = MTGV_DADD(1.0, t->d[0]);
      for (int c = 0; c < 4; c++) {
        o->f[c] = (float)alpha;
        o->f[4 + c] = (float)t->d[1];
      }
      o->i[0] = 15;
      o->i[1] = 1;
      return true;
    }
    default:
      return false;
  }
}

MTGV_HD bool is_elementwise(int code) {
  return code == MTGV_OP_TINT || code == MTGV_OP_FADE_BLACK || code == MTGV_OP_FADE_WHITE || code == MTGV_OP_BC;
}

// Ops that run on an (H, W) plane set: foreground RGBA and the composited RGB image.
// Returns 0 (nothing emitted: identity), 1 (one op emitted) or a negative status.
MTGV_HD int expand_plane_op(const mtgv_tape_op* t, int H, int W, mtgv_x_op* o) {
  if (is_elementwise(t->code)) return expand_elementwise(t, o) ? 1 : MTGV_ERR_INVALID;
  x_op_clear(o);
  o->field = t->field;
  o->field2 = t->field2;
  o->n_field = t->n_field;
  o->n_field2 = t->n_field2;
  switch (t->code) {
    case MTGV_OP_NONE:
      return 0;
    case MTGV_OP_DOWNUP:
      if (t->i[0] == 0) return 0;  // same-size cv2.resize is a copy
      if (t->i[0] < 0 || t->i[0] > 2) return MTGV_ERR_INVALID;
      o->code = MTGV_X_DOWNUP;
      o->i[0] = t->i[0];
      o->i[1] = t->i[1];
      o->i[2] = t->i[2];
      return 1;
    case MTGV_OP_WARP:
    case MTGV_OP_WARP_INV:
    case MTGV_OP_PERSPECTIVE: {
      float src[8], dst[8];
      double M[9];
      if (t->code == MTGV_OP_WARP)
        mutate_warp_points(H, W, t->d, 0.3, -0.25, src, dst);
      else if (t->code == MTGV_OP_WARP_INV)
        mutate_warp_points(H, W, t->d, -0.5, -0.25, src, dst);
      else
        mutate_perspective_points(H, W, t->d, src, dst);
      get_perspective_transform(src, dst, M);
      invert3x3(M, o->d);
      o->code = MTGV_X_WARP_PERSP;
      return 1;
    }
    case MTGV_OP_AFFINE: {
      double alpha, beta, M[6];
      if (t->i[0]) {
        alpha = t->d[5];
        beta = t->d[6];
      } else {
        double a = MTGV_DMUL(t->d[0], kPiOver180);
        alpha = MTGV_DMUL(cos(a), t->d[3]);
        beta = MTGV_DMUL(sin(a), t->d[3]);
      }
      mutate_affine_matrix(H, W, alpha, beta, t->d[1], t->d[2], t->d[4], M);
      invert_affine(M, o->d);
      o->code = MTGV_X_WARP_AFFINE;
      return 1;
    }
    case MTGV_OP_BLUR:
      if (t->i[0] == 1) return 0;
      if (t->i[0] != 3) return MTGV_ERR_INVALID;
      o->code = MTGV_X_BLUR3;
      return 1;
    case MTGV_OP_SHARPEN:
      o->code = MTGV_X_SHARPEN;
      return 1;
    case MTGV_OP_NOISE: {
      double ratio = MTGV_DMUL(t->d[0], 0.5);
      o->code = MTGV_X_NOISE;
      o->i[0] = t->i[0];
      o->f[0] = (float)ratio;
      o->f[1] = (float)MTGV_DSUB(1.0, ratio);
      if (t->i[0] == 2 && t->field == MTGV_FIELD_PHILOX) {
        // noise_salt_pepper(strength=0.1, svp=0.5): ceil(0.1 * img.size * 0.5) points each
        double sz = (double)(H * W * 3);
        o->n_field = (int)ceil(MTGV_DMUL(MTGV_DMUL(0.1, sz), 0.5));
        o->n_field2 = (int)ceil(MTGV_DMUL(MTGV_DMUL(0.1, sz), MTGV_DSUB(1.0, 0.5)));
      }
      return 1;
    }
    case MTGV_OP_GAUSS_NOISE:
      o->code = MTGV_X_GAUSS_NOISE;
      return 1;
    case MTGV_OP_SALT_PEPPER:
      o->code = MTGV_X_SALT_PEPPER;
      if (t->field == MTGV_FIELD_PHILOX) {
        double sz = (double)(H * W * 3);
        o->n_field = (int)ceil(MTGV_DMUL(0.01, sz));
        o->n_field2 = (int)ceil(MTGV_DMUL(0.01, sz));
      }
      return 1;
    case MTGV_OP_ERASE: {
      if (t->i[0] < 2) return 0;  // returned before filling anything
      int bw, bh;
      if (t->i[4]) {
        bw = t->i[5];
        bh = t->i[6];
      } else {
        double area = MTGV_DMUL(t->d[0], (double)(H * W));
        double aspect = t->d[1];
        if (t->d[2] < 0.5) aspect = MTGV_DDIV(1.0, aspect);
        bw = (int)sqrt(MTGV_DDIV(area, aspect));
        bh = (int)sqrt(MTGV_DMUL(area, aspect));
      }
      int cx = t->i[1], cy = t->i[2];
      int x0 = cx - bw / 2, y0 = cy - bh / 2, x1 = cx + bw / 2, y1 = cy + bh / 2;
      x0 = x0 < 0 ? 0 : x0;
      y0 = y0 < 0 ? 0 : y0;
      x1 = x1 > W ? W : x1;
      y1 = y1 > H ? H : y1;
      if (y1 <= y0 || x1 <= x0) return 0;
      o->code = MTGV_X_ERASE;
      o->i[0] = y0;
      o->i[1] = y1;
      o->i[2] = x0;
      o->i[3] = x1;
      int mode = t->i[3];
      o->i[4] = mode;
      float c = mode == 3 ? 1.f : 0.f;
      for (int k = 0; k < 3; k++) o->f[k] = mode == 1 ? (float)t->d[3 + k] : c;
      return 1;
    }
    case MTGV_OP_CUTOUT:
      o->code = MTGV_X_CUTOUT;
      for (int k = 0; k < 16; k++) o->i[k] = t->i[k];
      return 1;
    default:
      return MTGV_ERR_INVALID;
  }
}

// Hard negative: the swap_choice-th member of the same-name group with `card` removed
// (get_similar_card, mtgvision/encoder_datasets.py:619-630).
MTGV_HD int resolve_card(const PoolMeta& pm, int card, int swap_choice) {
  if (swap_choice < 0) return card;
  int lo = pm.grp_off[card], hi = pm.grp_off[card + 1], k = 0;
  for (int j = lo; j < hi; j++) {
    int m = pm.grp_mem[j];
    if (m == card) continue;
    if (k == swap_choice) return m;
    k++;
  }
  return -1;
}

MTGV_HD int expand_encoder_sample(const mtgv_enc_tape* t, const mtgv_enc_config* cfg, const PoolMeta& pm,
                                  mtgv_enc_params* p) {
  const int OH = cfg->out_h, OW = cfg->out_w;
  p->kind = t->kind;
  p->upsidedown = t->upsidedown;
  p->out_h = OH;
  p->out_w = OW;
  p->card_h = pm.card_h;
  p->card_w = pm.card_w;
  p->seed = t->seed;
  p->n_fg = p->n_pre = p->n_post = p->n_vrtl = 0;
  p->flip_h = p->flip_v = 0;
  p->bg_h = p->bg_w = p->rot_nh = p->rot_nw = p->bg_rh = p->bg_rw = p->bg_y0 = p->bg_x0 = 0;
  p->status = 0;
  p->_pad = 0;
  p->src_y0 = p->src_x0 = p->src_h = p->src_w = 0;
  p->fg_rh = p->fg_rw = p->fg_y0 = p->fg_x0 = 0;
  p->card = p->bg = 0;
  for (int k = 0; k < MTGV_X_MAX_OPS; k++) x_op_clear(&p->ops[k]);
  for (int k = 0; k < 6; k++) p->rot_inv[k] = 0.0;
  for (int k = 0; k < 9; k++) p->winv[k] = 0.0;
  if (t->card < 0 || t->card >= pm.n_cards) return p->status = MTGV_ERR_INVALID;
  int card = resolve_card(pm, t->card, t->swap_choice);
  if (card < 0) return p->status = MTGV_ERR_INVALID;
  p->card = card;
  p->bg = t->bg;

  if (t->kind == MTGV_KIND_CROPPED) {
    // make_cropped: border = ceil(max(0.02 H, 0.02 W)); plain resize to (OH, OW)
    int border = (int)ceil(fmax(MTGV_DMUL(0.02, (double)pm.card_h), MTGV_DMUL(0.02, (double)pm.card_w)));
    p->src_y0 = border;
    p->src_x0 = border;
    p->src_h = pm.card_h - 2 * border;
    p->src_w = pm.card_w - 2 * border;
    p->fg_rh = OH;
    p->fg_rw = OW;
    p->fg_y0 = p->fg_x0 = 0;
    if (p->src_h < OH || p->src_w < OW) return p->status = MTGV_ERR_LIMIT;
    if (p->src_h > 6 * OH || p->src_w > 6 * OW) return p->status = MTGV_ERR_LIMIT;
    return 0;
  }
  if (t->kind != MTGV_KIND_VIRTUAL && t->kind != MTGV_KIND_BG_ONLY) return p->status = MTGV_ERR_INVALID;
  const bool bg_only = t->kind == MTGV_KIND_BG_ONLY;
  if (t->bg < 0 || t->bg >= pm.n_bgs) return p->status = MTGV_ERR_INVALID;
  if (t->n_fg + t->n_bg + t->n_vrtl > MTGV_TAPE_MAX_OPS) return p->status = MTGV_ERR_INVALID;

  // foreground: whole card, crop_to_size(pad=True)
  p->src_y0 = p->src_x0 = 0;
  p->src_h = pm.card_h;
  p->src_w = pm.card_w;
  if (pm.card_h == OH && pm.card_w == OW) {
    p->fg_rh = OH;
    p->fg_rw = OW;
    p->fg_y0 = p->fg_x0 = 0;
  } else {
    crop_geometry(pm.card_h, pm.card_w, OH, OW, true, &p->fg_rh, &p->fg_rw, &p->fg_y0, &p->fg_x0);
  }
  if (p->fg_rh < 1 || p->fg_rw < 1 || p->fg_rh > pm.card_h || p->fg_rw > pm.card_w) return p->status = MTGV_ERR_LIMIT;
  if (pm.card_h > 6 * p->fg_rh || pm.card_w > 6 * p->fg_rw) return p->status = MTGV_ERR_LIMIT;

  int n = 0;
  const mtgv_tape_op* ops = t->ops;
  for (int k = 0; k < (bg_only ? 0 : t->n_fg); k++) {
    if (n >= MTGV_X_MAX_OPS) return p->status = MTGV_ERR_LIMIT;
    int r = expand_plane_op(&ops[k], OH, OW, &p->ops[n]);
    if (r < 0) return p->status = r;
    n += r;
  }
  p->n_fg = n;

  // background: elementwise ops before / after the geometric group flip->rotate->warp_inv
  const int bh = pm.bg_hw[2 * t->bg], bw = pm.bg_hw[2 * t->bg + 1];
  p->bg_h = bh;
  p->bg_w = bw;
  bool seen_geo = false;
  int n_pre = 0, n_post = 0;
  for (int k = t->n_fg; k < t->n_fg + t->n_bg; k++) {
    const mtgv_tape_op* o = &ops[k];
    if (o->code == MTGV_OP_FLIP) {
      p->flip_h = o->i[0];
      p->flip_v = o->i[1];
    } else if (o->code == MTGV_OP_ROTATE) {
      double alpha, beta, M[6];
      if (o->i[0]) {
        alpha = o->d[1];
        beta = o->d[2];
      } else {
        double deg = MTGV_DADD(0.0, MTGV_DMUL(o->d[0], 360.0));
        double a = MTGV_DMUL(deg, kPiOver180);
        alpha = cos(a);
        beta = sin(a);
      }
      rotate_bounded_matrix(bh, bw, alpha, beta, M, &p->rot_nh, &p->rot_nw);
      invert_affine(M, p->rot_inv);
    } else if (o->code == MTGV_OP_WARP_INV) {
      if (p->rot_nh < 2 || p->rot_nw < 2) return p->status = MTGV_ERR_INVALID;
      float src[8], dst[8];
      double M[9];
      mutate_warp_points(p->rot_nh, p->rot_nw, o->d, -0.5, -0.25, src, dst);
      get_perspective_transform(src, dst, M);
      invert3x3(M, p->winv);
      seen_geo = true;
    } else if (is_elementwise(o->code)) {
      if (n >= MTGV_X_MAX_OPS) return p->status = MTGV_ERR_LIMIT;
      expand_elementwise(o, &p->ops[n]);
      p->ops[n].i[0] &= 7;  // backgrounds are RGB
      n++;
      if (seen_geo) n_post++; else n_pre++;
    } else {
      return p->status = MTGV_ERR_INVALID;
    }
  }
  if (!seen_geo) return p->status = MTGV_ERR_INVALID;
  p->n_pre = n_pre;
  p->n_post = n_post;
  if (p->rot_nh == OH && p->rot_nw == OW) {
    p->bg_rh = OH;
    p->bg_rw = OW;
    p->bg_y0 = p->bg_x0 = 0;
  } else {
    crop_geometry(p->rot_nh, p->rot_nw, OH, OW, false, &p->bg_rh, &p->bg_rw, &p->bg_y0, &p->bg_x0);
  }
  // INTER_AREA: down-scale by at most kBgMaxAreaScale (photographs up to ~4600 px diagonal at 192x128; the pools refuse
  // larger images when they are filled, so this never trips at run time); when the canvas is smaller than the output cv2
  // enlarges with its 2-tap "area-linear" kernel on both axes (area_linear_compact)
  if (p->bg_rh < OH || p->bg_rw < OW) return p->status = MTGV_ERR_LIMIT;
  if (p->rot_nh > kBgMaxAreaScale * p->bg_rh || p->rot_nw > kBgMaxAreaScale * p->bg_rw) return p->status = MTGV_ERR_LIMIT;

  for (int k = t->n_fg + t->n_bg; k < t->n_fg + t->n_bg + (bg_only ? 0 : t->n_vrtl); k++) {
    if (n >= MTGV_X_MAX_OPS) return p->status = MTGV_ERR_LIMIT;
    int r = expand_plane_op(&ops[k], OH, OW, &p->ops[n]);
    if (r < 0) return p->status = r;
    if (r) p->ops[n].i[0] &= (p->ops[n].code == MTGV_X_ELEM ? 7 : -1);
    p->n_vrtl += r;
    n += r;
  }
  return 0;
}

}  // namespace mtgv
