/*
 * mf_b200.h -- C ABI of libmf_b200.so: the B200-native (sm_100a) audio -> face-frame hot path
 * that sits behind Caxson/mere-fusion's BaseReal / BaseASR plugin surface.
 *
 * Conventions:
 *   - extern "C", plain pointers and sizes, no C++ / torch types;
 *   - every function returns 0 on success or a negative MF_E_* code, never throws;
 *     mf_last_error() gives the message for the last failure on that context;
 *   - unless stated otherwise every data pointer is a DEVICE pointer owned by the caller
 *     (e.g. torch tensor .data_ptr()); pointers documented "host" are host memory;
 *   - `stream` is a cudaStream_t passed as void*; work is enqueued asynchronously on it, with
 *     no hidden synchronisation (mf_*_load are the exception: they synchronise once);
 *   - the library owns only its context: weights, workspaces and per-session state (the ErNeRF
 *     audio-feature EMA); one mf_ctx per (GPU, session); contexts are independent, a single
 *     context is not thread-safe.
 *
 * Each entry point cites the reference interface it replaces (paths relative to the
 * reference repository root).
 */
#ifndef MF_B200_H
#define MF_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define MF_ABI_VERSION 1

#define MF_OK 0
#define MF_E_INVALID -1     /* bad argument / shape */
#define MF_E_CUDA -2        /* a CUDA runtime call failed */
#define MF_E_STATE -3       /* weights for this head not loaded */
#define MF_E_UNSUPPORTED -4 /* configuration outside what the kernels implement */

typedef struct mf_ctx mf_ctx;

int mf_version(void);
/* one context per (GPU, session).  Fails with MF_E_CUDA when no sm_100 device is present:
 * there is no CPU fallback. */
int mf_create(int device, mf_ctx **out);
void mf_destroy(mf_ctx *ctx);
const char *mf_last_error(const mf_ctx *ctx);

/* ------------------------------------------------------------------------------------------
 * ErNeRF, kernel level: drop-in equivalents of the reference's pybind functions, same argument
 * order and meaning, at::Tensor replaced by device pointers.  Outputs are pre-allocated by the
 * caller exactly as the reference's Python wrappers do.
 * ------------------------------------------------------------------------------------------ */

/* ernerf/raymarching/src/raymarching.h:7  near_far_from_aabb (kernel raymarching.cu:91-145) */
int mf_near_far_from_aabb(mf_ctx *ctx, const float *rays_o, const float *rays_d, const float *aabb,
                          uint32_t N, float min_near, float *nears, float *fars, void *stream);

/* ernerf/raymarching/src/raymarching.h:19  march_rays (kernel raymarching.cu:827-929).
 * xyzs/dirs/deltas must be zero-filled by the caller (raymarching/raymarching.py:383-385).
 * noises may be NULL (= zeros, perturb=False). */
int mf_march_rays(mf_ctx *ctx, uint32_t n_alive, uint32_t n_step, const int32_t *rays_alive,
                  const float *rays_t, const float *rays_o, const float *rays_d, float bound,
                  float dt_gamma, uint32_t max_steps, uint32_t C, uint32_t H, const uint8_t *grid,
                  const float *nears, const float *fars, float *xyzs, float *dirs, float *deltas,
                  const float *noises, void *stream);

/* ernerf/raymarching/src/raymarching.h:37  composite_rays_triplane (kernel raymarching.cu:2141-2249).
 * ambs_aud / ambs_eye / uncertainties and their *_sum outputs may be NULL (unused by the live
 * path). */
int mf_composite_rays_triplane(mf_ctx *ctx, uint32_t n_alive, uint32_t n_step, float T_thresh,
                               int32_t *rays_alive, float *rays_t, const float *sigmas,
                               const float *rgbs, const float *deltas, const float *ambs_aud,
                               const float *ambs_eye, const float *uncertainties,
                               float *weights_sum, float *depth, float *image, float *amb_aud_sum,
                               float *amb_eye_sum, float *uncertainty_sum, void *stream);

/* ernerf/gridencoder/src/gridencoder.h:12  grid_encode_forward, D = 2 (kernel gridencoder.cu:75-175).
 * embeddings_is_half: 0 = fp32 table + fp32 outputs, 1 = fp16 table + fp16 outputs (the torso
 * encoder under autocast).  outputs is [L, B, C] like the reference kernel's. */
int mf_grid_encode_forward(mf_ctx *ctx, const float *inputs, const void *embeddings,
                           const int32_t *offsets, void *outputs, uint32_t B, uint32_t D,
                           uint32_t C, uint32_t L, float S, uint32_t H, uint32_t gridtype,
                           int align_corners, int embeddings_is_half, void *stream);

/* The per-level scale `exp2f(level * S) * H - 1.0f` (gridencoder.cu:124) as the device evaluates it
 * (CUDA exp2f is approximate): scales_host is a HOST array of L floats.  Lets a CPU checker use
 * the very constants the kernels use.  Synchronises. */
int mf_grid_level_scales(mf_ctx *ctx, float S, uint32_t H, uint32_t L, float *scales_host);

/* ernerf/shencoder/src/shencoder.h:9  sh_encode_forward, degree C = 4 (kernel shencoder.cu:27-68) */
int mf_sh_encode_forward(mf_ctx *ctx, const float *inputs, float *outputs, uint32_t B, uint32_t D,
                         uint32_t C, void *stream);

/* ernerf/freqencoder/src/freqencoder.h:7  freq_encode_forward (kernel freqencoder.cu:30-58) */
int mf_freq_encode_forward(mf_ctx *ctx, const float *inputs, uint32_t B, uint32_t D, uint32_t deg,
                           uint32_t C, float *outputs, void *stream);

/* ------------------------------------------------------------------------------------------
 * ErNeRF, frame level: the fused render that replaces
 *   NeRFReal.test_step -> Trainer.test_gui_with_data -> NeRFRenderer.run_cuda + run_torso
 *   (nerfreal.py:70-127, ernerf/nerf_triplane/utils.py:1191-1223, renderer.py:158-352).
 * ------------------------------------------------------------------------------------------ */

#define MF_ERNERF_HEAD_LEVELS 12
#define MF_ERNERF_TORSO_LEVELS 16

/* host struct.  Mirrors what reaches run_cuda through **vars(opt) (utils.py:949-950). */
typedef struct mf_ernerf_cfg {
    float bound;                /* opt.bound (1) */
    float min_near;             /* opt.min_near (0.05) */
    float dt_gamma;             /* opt.dt_gamma (1/256) */
    float T_thresh;             /* run_cuda default 1e-4 */
    float density_thresh_torso; /* min(opt.density_thresh_torso, mean_density_torso), renderer.py:325 */
    float torso_shrink;         /* opt.torso_shrink (0.8) */
    uint32_t max_steps;         /* opt.max_steps (16) */
    uint32_t cascade;           /* 1 + ceil(log2(bound)); only 1 is supported */
    uint32_t grid_size;         /* 128 */
    uint32_t smooth_lips;       /* opt.smooth_lips: EMA of the audio feature, renderer.py:190-194 */
    float head_log2_scale;      /* S = log2(per_level_scale) of encoder_xy/yz/xz (grid.py:31) */
    uint32_t head_base;         /* 64 */
    int32_t head_offsets[MF_ERNERF_HEAD_LEVELS + 1];
    float torso_log2_scale;
    uint32_t torso_base;        /* 16 */
    int32_t torso_offsets[MF_ERNERF_TORSO_LEVELS + 1];
    uint32_t audio_in_dim;      /* 44 for the esperanto wav2vec2 head (network.py:104-111) */
} mf_ernerf_cfg;

/* `blob` is the DEVICE-resident packed checkpoint produced by
 * mere_fusion_b200.ernerf_pack.pack_ernerf (so a torch.distributed-broadcast tensor can be
 * handed over directly).  The blob must stay alive until mf_destroy. */
int mf_ernerf_load(mf_ctx *ctx, const void *blob, size_t nbytes, const mf_ernerf_cfg *cfg /*host*/);

typedef struct mf_ernerf_frame {
    const float *pose;       /* host, [4,4] row-major cam2world after nerf_matrix_to_ngp (provider.py:19-26) */
    float fx, fy, cx, cy;    /* intrinsics (provider.py:263-270) */
    int32_t H, W;            /* render size: rays are the H*W pixel centres (utils.py:255-341) */
    const float *auds;       /* device, [8, audio_in_dim, 16] attention window (nerfasr.py:75-103) */
    const float *enc_a;      /* device or NULL: [32] pre-encoded audio feature; skips encode_audio and
                                the EMA state so frames can be rendered out of order */
    float eye;               /* eye area (provider.py:240-253) */
    const void *bg_color;    /* device or NULL: fp16 [H*W,3]; NULL = white (opt.bg_img default) */
    const float *rays_o;     /* device or NULL: explicit rays [N,3] (then N = n_rays, and */
    const float *rays_d;     /*   bg_coords [N,2] must be given too) instead of the pixel grid */
    const float *bg_coords;
    int32_t n_rays;
    int32_t outH, outW;      /* output size: bilinear resize, utils.py:1212; ignored for explicit rays */
    float *out_image_f32;    /* device or NULL: [outH,outW,3] fp32 in [0,1] (what test_gui_with_data returns) */
} mf_ernerf_frame;

/* optional observability for the parity tests (all device pointers, any may be NULL) */
typedef struct mf_ernerf_debug {
    float *nears, *fars;     /* [N] */
    int32_t *round_info;     /* [17*4] per round: n_alive, tiles handed out, samples emitted, n_step */
    float *weights_sum;      /* [N] */
    float *image_head;       /* [N,3] composited head before background */
    float *enc_a;            /* [32] the (smoothed) audio feature used */
    uint8_t *torso_mask;     /* [N] */
} mf_ernerf_debug;

/* out_rgb: device u8 [outH,outW,3] RGB = (image*255) truncated (nerfreal.py:110). */
int mf_ernerf_render(mf_ctx *ctx, const mf_ernerf_frame *frame /*host*/, uint8_t *out_rgb,
                     const mf_ernerf_debug *dbg /*host, nullable*/, void *stream);
/* The same render for n frames of n DIFFERENT sessions in one pass (n <= 4): ctxs[i] renders frames[i] into outs[i] (device u8
 * [outH,outW,3]) with its own per-session state, exactly as n mf_ernerf_render calls would (bit-identical images), but with ONE
 * launch of the fused march / encode / MLP / composite kernel for the whole batch -- the reference renders one session per
 * process (app.py:331-392) and has no equivalent.  All contexts must live on the same GPU and have been loaded from the same
 * device blob with the same configuration (one avatar model, many sessions).  Errors are reported on ctxs[0]. */
int mf_ernerf_render_batch(mf_ctx *const *ctxs /*host*/, const mf_ernerf_frame *frames /*host, [n]*/,
                           uint8_t *const *outs /*host array of device pointers*/, int n, void *stream);
/* The audio half of a frame on its own: encode_audio (network.py:222-237) + the EMA of renderer.py:190-194 on one attention window,
 * advancing the session's audio state exactly as mf_ernerf_render would, and writing the smoothed feature to enc_a_out (device fp32
 * [32]).  SURVEY 8(e): when ONE session's frames are sharded round-robin over several GPUs, every rank follows the session's audio
 * state with this call (8 small CTAs) and renders only its own frames with mf_ernerf_frame.enc_a set -- frames can then be rendered
 * out of order and are bit-identical to the in-order stream. */
int mf_ernerf_encode_audio(mf_ctx *ctx, const float *auds /*device [8, audio_in_dim, 16]*/, float *enc_a_out, void *stream);
/* forget the audio-feature EMA (a new session on a reused context) */
int mf_ernerf_reset_state(mf_ctx *ctx);
/* offsets of the packed-blob MLP images (csrc/ernerf_layout.h) for the Python packer; returns the
 * number of values defined and writes min(n, that) of them. */
int mf_ernerf_blob_layout(int32_t *out, int n);
/* measurement hooks (bench.py roofline): with profiling enabled every mf_ernerf_render records CUDA
 * events around the dominant kernel (k_head) on the launching stream; mf_ernerf_last_head_ms waits
 * for the last one and returns its duration and the number of samples it marched+shaded. */
int mf_ernerf_profile(mf_ctx *ctx, int enable);
int mf_ernerf_last_head_ms(mf_ctx *ctx, float *ms /*host*/, int64_t *samples /*host*/);
/* kernels launched by the last mf_ernerf_render on this context */
int mf_ernerf_last_launches(const mf_ctx *ctx);

/* ------------------------------------------------------------------------------------------
 * Wav2Lip: replaces the model(mel_batch, img_batch) call and the batch build / x255 around it in
 * the reference's inference() loop (lipreal.py:108-126; network wav2lip/models/wav2lip.py:87-125).
 * ------------------------------------------------------------------------------------------ */

/* `blob` = DEVICE-resident conv-net program + weights from mere_fusion_b200.wav2lip_pack.pack_wav2lip
 * (BatchNorm folded to per-channel scale/shift, weights bf16).  Activation buffers are allocated
 * for max_batch frames.  The blob must stay alive until mf_destroy.  Synchronises. */
int mf_wav2lip_load(mf_ctx *ctx, const void *blob, size_t nbytes, int max_batch);

/* mel    : device fp32 [B,1,80,16]   (LipASR.run_step chunks, lipasr.py:29-35)
 * faces  : device u8  [B,S,S,3] BGR  (the avatar face crops, lipreal.py:109-111); the lower-half mask,
 *          the 6-channel concat and /255 happen inside
 * out_u8 : device u8  [B,S,S,3] = (sigmoid * 255) truncated  (lipreal.py:126,209), nullable
 * out_f32: device fp32 [B,S,S,3] in (0,1) (what `pred.cpu().numpy().transpose(0,2,3,1)` holds), nullable */
int mf_wav2lip_forward(mf_ctx *ctx, const float *mel, const uint8_t *faces, uint8_t *out_u8, float *out_f32,
                       int B, void *stream);
int mf_wav2lip_last_launches(const mf_ctx *ctx);
/* measurement hooks: record CUDA events around conv op `op_index` (-1 = off) on the launching stream */
int mf_wav2lip_profile(mf_ctx *ctx, int op_index);
int mf_wav2lip_last_op_ms(mf_ctx *ctx, float *ms /*host*/);
/* unit-test entry: run the loaded program on an fp32 NHWC tensor placed in buffer in_buf and read
 * buffer out_buf back as fp32 NHWC (programs without an output head only) */
int mf_convnet_debug_run(mf_ctx *ctx, int in_buf, const float *in_f32, int out_buf, float *out_f32, int B,
                         void *stream);

/* ------------------------------------------------------------------------------------------
 * MuseTalk: replaces pe(whisper) -> unet.model(latents, t=0, encoder_hidden_states).sample ->
 * vae.decode_latents(pred) of the reference's inference() loop (musereal.py:99-108,
 * musetalk/models/unet.py:12-27, musetalk/models/vae.py:96-108).  The program blob comes from
 * mere_fusion_b200.musetalk_pack.pack_musetalk and is loaded with mf_wav2lip_load (same executor).
 *   latents_f16 : device fp16 [B,8,32,32] NCHW  (masked || reference latents, vae.py:110-122)
 *   whisper_f16 : device fp16 [B,50,384]        (MuseASR chunks, BEFORE the positional encoding)
 *   out_u8      : device u8  [B,256,256,3] BGR  = ((x/2+0.5).clamp(0,1)*255).round(), nullable
 *   out_f32     : device fp32 [B,256,256,3] RGB in [0,1], nullable
 * ------------------------------------------------------------------------------------------ */
int mf_musetalk_forward(mf_ctx *ctx, const void *latents_f16, const void *whisper_f16, uint8_t *out_u8,
                        float *out_f32, int B, void *stream);
/* unit-test helper: write an fp32 NHWC tensor into program buffer `buf` (second input of debug programs) */
int mf_convnet_debug_set(mf_ctx *ctx, int buf, const float *in_f32, int B, void *stream);

/* ------------------------------------------------------------------------------------------
 * Whisper audio features for MuseTalk: replaces Audio2Feature.audio2feat -> Whisper.transcribe ->
 * log_mel_spectrogram + AudioEncoder.forward(include_embeddings=True)
 * (musetalk/whisper/audio2feature.py:99-112, whisper/transcribe.py:85-128, whisper/audio.py:92-125,
 * whisper/model.py:143-171).  The program blob comes from mere_fusion_b200.whisper_pack.pack_whisper and
 * is loaded with mf_wav2lip_load(max_batch = 1) into its own context.
 *   audio   : device fp32 [n_samples], 16 kHz mono, one segment (n_samples / 160 <= 3000 frames)
 *   out_f32 : device fp32 [T, n_layer + 1, n_state] = embeddings.transpose(0,2,1,3)[0][:T]
 *             (T = int(n_frames / 2) is what audio2feat keeps)
 * ------------------------------------------------------------------------------------------ */
int mf_whisper_features(mf_ctx *ctx, const float *audio, int n_samples, float *out_f32, int T, void *stream);

/* ------------------------------------------------------------------------------------------
 * wav2vec2 CTC logits for ErNeRF: replaces `processor(frame) -> model(input_values).logits` of
 * NerfASR.__frame_to_text (nerfasr.py:128-143; third-party HF Wav2Vec2ForCTC, XLSR-53 large).  The program blob comes
 * from mere_fusion_b200.wav2vec2_pack.pack_wav2vec2 (built for one window length) and is loaded with
 * mf_wav2lip_load(max_batch = 1) into its own context.
 *   audio   : device fp32 [n_samples] (the (l + m + r) x 20 ms window: 8960 samples for the live defaults)
 *   out_f32 : device fp32 [n_frames, vocab] (27 x 44); NerfASR keeps rows [l : T - r + 1]
 * ------------------------------------------------------------------------------------------ */
int mf_wav2vec2_logits(mf_ctx *ctx, const float *audio, int n_samples, float *out_f32, void *stream);
/* the same for B windows at once (B <= the max_batch given to mf_wav2lip_load): the windows of B different ErNeRF sessions of one GPU in one
 * pass over the 630 MB of weights (the model is weight-streaming / launch-latency bound at one window).
 *   audio   : device fp32 [B, n_samples]          out_f32 : device fp32 [B, n_frames, vocab] */
int mf_wav2vec2_logits_batch(mf_ctx *ctx, const float *audio, int n_samples, int B, float *out_f32, void *stream);
/* debug tap (no reference counterpart): %globaltimer stamps in ns of one CTA at the 11 phase boundaries of the second transformer layer of
 * the last fused-stack launch (csrc/w2v_stack.cuh); synchronises the device.  scripts/time_w2v.py prints them. */
int mf_debug_w2v_phase_ns(mf_ctx *ctx, unsigned long long *out, int n);

/* ------------------------------------------------------------------------------------------
 * Paste-back (lipreal.py:207-214): out[i] = frames[idx_i] with faces[i] resized (cv2.resize, u8,
 * INTER_LINEAR, bit-exact) into the box (y1:y2, x1:x2).  coords order as wav2lip/genavatar.py:96.
 *   frames : device u8 [n_frames,H,W,3] the avatar's full frames, resident
 *   faces  : device u8 [B,S,S,3]        the generated crops (mf_wav2lip_forward out_u8)
 *   idx_bbox_host : HOST int32 [B,5] = (frame index, y1, y2, x1, x2) per item, B <= 64
 *   out    : device u8 [B,H,W,3]
 * ------------------------------------------------------------------------------------------ */
int mf_paste_resize_u8(mf_ctx *ctx, const uint8_t *frames, int n_frames, int H, int W, const uint8_t *faces,
                       int S, int B, const int32_t *idx_bbox_host, uint8_t *out, void *stream);

/* ------------------------------------------------------------------------------------------
 * Wav2Lip mel front-end: replaces audio.melspectrogram (wav2lip/audio.py:45-51: preemphasis, STFT n_fft 800 /
 * hop 200 / Hann, 80-band mel 55-7600 Hz, 20 log10 - 20, clip(8 (S + 100) / 100 - 4, +-4)) and the chunk slicing
 * of LipASR.run_step (lipasr.py:24-35) for one window.
 *   audio          : device fp32 [n_samples] (the l + 2B + r chunk window, 16 kHz)
 *   filters        : device fp32 [80][401], the mel basis (mere_fusion_b200.audio_mel.mel_filterbank())
 *   start_idx_host : HOST int32 [B] first mel column of each chunk (lipasr.py:29-33, tail clamp included)
 *   out            : device fp32 [B,1,80,16], what mf_wav2lip_forward takes as `mel`
 * ------------------------------------------------------------------------------------------ */
int mf_wav2lip_mel_chunks(mf_ctx *ctx, const float *audio, int n_samples, const float *filters,
                          const int32_t *start_idx_host, int B, float *out, void *stream);

/* ------------------------------------------------------------------------------------------
 * MuseTalk paste-back (musereal.py:229-250 -> musetalk/utils/blending.py:103-125 get_image_blending):
 * faces[i] is resized (cv2.resize, bit-exact) into its bbox inside a copy of the body crop, then
 *   body[ys:ye, xs:xe] = cv2.blendLinear(face_large, body_crop, m, 1 - m),  m = gray(mask) / 255
 * bit-exact against OpenCV 4.x (cvtColor BGR2GRAY 15-bit fixed point, blendLinear fp32 without FMA).
 *   rows_host     : HOST int32 [B,9] = (frame index, y1, y2, x1, x2 of the face bbox, ys, ye, xs, xe of
 *                   the mask crop box); note musereal.py keeps bboxes as (x1, y1, x2, y2)
 *   masks         : device u8, the avatar's masks packed back to back, each BGR [ye-ys, xe-xs, 3]
 *                   (mask/*.png as cv2.imread returns them, musereal.py:176-179)
 *   mask_off_host : HOST int64 [B] byte offset of each row's mask inside `masks`
 * ------------------------------------------------------------------------------------------ */
int mf_paste_blend_u8(mf_ctx *ctx, const uint8_t *frames, int n_frames, int H, int W, const uint8_t *faces,
                      int S, int B, const int32_t *rows_host, const uint8_t *masks, size_t masks_nbytes,
                      const int64_t *mask_off_host, uint8_t *out, void *stream);

#ifdef __cplusplus
}
#endif
#endif /* MF_B200_H */
