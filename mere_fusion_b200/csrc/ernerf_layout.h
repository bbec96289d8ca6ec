// ernerf_layout.h -- layout of the packed ErNeRF checkpoint blob.  Single source of truth:
// the Python packer queries these numbers through mf_ernerf_blob_layout() instead of
// duplicating them.
#pragma once
#include <stdint.h>

// blob entry ids
enum {
    ER_ID_HEAD_PLANES = 1,   // f32 [3][rows]   encoder_xy, encoder_yz, encoder_xz embeddings
    ER_ID_BITFIELD = 2,      // u8  [cascade * G^3 / 8]
    ER_ID_TORSO_TABLE = 3,   // f16 [rows][2]   torso_encoder embeddings, halved once at pack time
    ER_ID_TORSO_DENSITY = 4, // f32 [G*G]       density_grid_torso
    ER_ID_HEAD_MLP = 5,      // shared-memory image of the head MLPs (layout below)
    ER_ID_TORSO_MLP = 6,     // shared-memory image of the torso MLPs
    ER_ID_AUDIO = 7,         // f16 audio_net + audio_att_net weights/biases, in module order
    ER_ID_MISC = 8,          // f32: anchor_points[12], individual_codes[0][4], individual_codes_torso[0][8]
    ER_ID_TORSO_CONST = 9,   // f16 [2][32][50]: columns of torso_deform_net.0 / torso_net.0 that
                             //     multiply the per-frame constants (enc_anchor 42 + ind code 8)
};

// All MLP matrices are stored as W[n][k] fp16 rows (n = output unit) with a row stride of
// K + 8 halfs, so that both ldmatrix row fetches and 32-bit fragment loads are bank-conflict
// free.  Offsets in halfs.
//
// head image.  K layouts:
//   enc_x block  : cols 0..35 = [xy L0..11 | yz L0..11 | xz L0..11], col 36 = e (eye * eye_att),
//                  cols 37..47 = 0
//   sigma_net.0  : K = 80 = enc_x block (48) | enc_w (32)            (orig cols 0..35,68 | 36..67)
//   sigma_net.2  : N = 72 = geo_feat 64 (orig rows 1..64) | sigma logit (orig row 0) | 7 zero rows
//   color_net.0  : K = 80 = geo_feat 64 (orig cols 16..79) | SH 16 (orig cols 0..15);
//                  the individual-code columns 80..83 are folded into COLBIAS
//   color_net.1  : N = 8 = rgb 3 | 5 zero rows
#define ER_H_AUD1 0                          /* [64][56] */
#define ER_H_AUD2 (ER_H_AUD1 + 64 * 56)      /* [32][72] */
#define ER_H_EYE1 (ER_H_AUD2 + 32 * 72)      /* [16][56] */
#define ER_H_SIG1 (ER_H_EYE1 + 16 * 56)      /* [64][88] */
#define ER_H_SIG2 (ER_H_SIG1 + 64 * 88)      /* [64][72] */
#define ER_H_SIG3 (ER_H_SIG2 + 64 * 72)      /* [72][72] */
#define ER_H_COL1 (ER_H_SIG3 + 72 * 72)      /* [64][88] */
#define ER_H_COL2 (ER_H_COL1 + 64 * 88)      /* [8][72] */
#define ER_H_EYE2 (ER_H_COL2 + 8 * 72)       /* [16] */
#define ER_H_HALFS (ER_H_EYE2 + 16)
#define ER_H_COLBIAS_BYTES (ER_H_HALFS * 2)  /* f32 [64] */
#define ER_H_BYTES (ER_H_COLBIAS_BYTES + 64 * 4)

// torso image.
//   torso_deform_net.0 : K = 48 = freq(x, deg 8) 34 | 14 zero  (constants folded into a bias)
//   torso_net.0        : K = 80 = grid feat 32 | freq(x) 34 | 14 zero
//   last layers        : N = 8 (2 resp. 4 rows used)
#define ER_T_DEF1 0                          /* [32][56] */
#define ER_T_DEF2 (ER_T_DEF1 + 32 * 56)      /* [32][40] */
#define ER_T_DEF3 (ER_T_DEF2 + 32 * 40)      /* [8][40] */
#define ER_T_TOR1 (ER_T_DEF3 + 8 * 40)       /* [32][88] */
#define ER_T_TOR2 (ER_T_TOR1 + 32 * 88)      /* [32][40] */
#define ER_T_TOR3 (ER_T_TOR2 + 32 * 40)      /* [8][40] */
#define ER_T_HALFS (ER_T_TOR3 + 8 * 40)
#define ER_T_BYTES (ER_T_HALFS * 2)
