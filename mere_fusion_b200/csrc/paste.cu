// paste.cu -- paste-back of the generated face crop into the avatar's full frame, on the GPU.
//
// Replaces, per output frame (lipreal.py:207-214):
//     combine_frame = copy.deepcopy(frame_list_cycle[idx])
//     res_frame = cv2.resize(res_frame.astype(np.uint8), (x2 - x1, y2 - y1))
//     combine_frame[y1:y2, x1:x2] = res_frame
// cv2.resize(u8, INTER_LINEAR) is restated bit-exactly: OpenCV's fixed-point bilinear
// (imgproc/src/resize.cpp: coordinates (dx + 0.5) * scale - 0.5 evaluated in double and narrowed to
// float, 11-bit coefficients saturate_cast<short>(c * 2048), horizontal pass in int, vertical pass
// ((b0 * (r0 >> 4)) >> 16) + ((b1 * (r1 >> 4)) >> 16) + 2) >> 2), including its switch to the 2x2
// area average when both scale factors are exactly 2 (INTER_LINEAR -> INTER_AREA fast path).
// HBM-bound byte work: one thread per output pixel, coalesced 3-byte pixels, no staging needed
// (every source byte is read at most ~4 times and stays in L1/L2).
#include "mf_common.cuh"

#define PASTE_MAX_BATCH 64

struct PasteParams {
    const uint8_t *frames;   // [n_frames, H, W, 3]
    const uint8_t *faces;    // [B, S, S, 3]
    uint8_t *out;            // [B, H, W, 3]
    int H, W, S, B;
    int idx[PASTE_MAX_BATCH];
    int y1[PASTE_MAX_BATCH], y2[PASTE_MAX_BATCH], x1[PASTE_MAX_BATCH], x2[PASTE_MAX_BATCH];
};

// OpenCV resize.cpp, INTER_LINEAR coefficient set-up for one axis
// x axis: the fraction is zeroed when a tap falls off either end; y axis (VERTICAL): OpenCV keeps the
// fraction and clamps the two row indices instead, which rounds differently on the border rows.
template <bool VERTICAL>
__device__ __forceinline__ void cv_linear_coef(int d, int ssize, double scale, int &s0, int &s1, int &a0, int &a1) {
    float f = (float)((d + 0.5) * scale - 0.5);
    int s = (int)floorf(f);
    f -= s;
    if (!VERTICAL) {
        if (s < 0) { f = 0.f; s = 0; }
        if (s >= ssize - 1) { f = 0.f; s = ssize - 1; }
    }
    s0 = max(0, min(s, ssize - 1));
    s1 = max(0, min(s + 1, ssize - 1));
    a0 = max(-32768, min(32767, __float2int_rn((1.f - f) * 2048.f)));
    a1 = max(-32768, min(32767, __float2int_rn(f * 2048.f)));
}

// one pixel (dx, dy) of cv2.resize(face[S, S, 3], (dw, dh)), INTER_LINEAR
__device__ __forceinline__ void cv_resize_px(const uint8_t *face, int S, int dw, int dh, int dx, int dy, uint8_t (&o)[3]) {
    if (S == 2 * dw && S == 2 * dh) {  // INTER_AREA fast path (exact 2x decimation)
#pragma unroll
        for (int c = 0; c < 3; c++) {
            const int s = face[((2 * dy) * S + 2 * dx) * 3 + c] + face[((2 * dy) * S + 2 * dx + 1) * 3 + c] +
                          face[((2 * dy + 1) * S + 2 * dx) * 3 + c] + face[((2 * dy + 1) * S + 2 * dx + 1) * 3 + c];
            o[c] = (uint8_t)((s + 2) >> 2);
        }
        return;
    }
    int sx0, sx1, ax0, ax1, sy0, sy1, by0, by1;
    cv_linear_coef<false>(dx, S, (double)S / dw, sx0, sx1, ax0, ax1);
    cv_linear_coef<true>(dy, S, (double)S / dh, sy0, sy1, by0, by1);
#pragma unroll
    for (int c = 0; c < 3; c++) {
        const int r0 = face[(sy0 * S + sx0) * 3 + c] * ax0 + face[(sy0 * S + sx1) * 3 + c] * ax1;
        const int r1 = face[(sy1 * S + sx0) * 3 + c] * ax0 + face[(sy1 * S + sx1) * 3 + c] * ax1;
        const int v = (((by0 * (r0 >> 4)) >> 16) + ((by1 * (r1 >> 4)) >> 16) + 2) >> 2;
        o[c] = (uint8_t)max(0, min(255, v));
    }
}

__global__ void __launch_bounds__(256) k_paste_resize(const __grid_constant__ PasteParams p) {
    const int b = blockIdx.y;
    const int pix = blockIdx.x * blockDim.x + threadIdx.x;
    if (pix >= p.H * p.W) return;
    const int y = pix / p.W, x = pix - y * p.W;
    const uint8_t *src = p.frames + ((size_t)p.idx[b] * p.H * p.W + pix) * 3;
    uint8_t *dst = p.out + ((size_t)b * p.H * p.W + pix) * 3;
    const int y1 = p.y1[b], y2 = p.y2[b], x1 = p.x1[b], x2 = p.x2[b];
    if (y < y1 || y >= y2 || x < x1 || x >= x2) {
        dst[0] = src[0]; dst[1] = src[1]; dst[2] = src[2];
        return;
    }
    uint8_t o[3];
    cv_resize_px(p.faces + (size_t)b * p.S * p.S * 3, p.S, x2 - x1, y2 - y1, x - x1, y - y1, o);
    dst[0] = o[0]; dst[1] = o[1]; dst[2] = o[2];
}

// ---- MuseTalk paste-back (musereal.py:229-250 + musetalk/utils/blending.py:103-125) -------------------------
//     res = cv2.resize(res_frame.astype(np.uint8), (x2 - x1, y2 - y1))
//     face_large = body[ys:ye, xs:xe].copy();  face_large[y1-ys:y2-ys, x1-xs:x2-xs] = res
//     m = (cv2.cvtColor(mask, COLOR_BGR2GRAY) / 255).astype(np.float32)
//     body[ys:ye, xs:xe] = cv2.blendLinear(face_large, body[ys:ye, xs:xe], m, 1 - m)
// cvtColor(u8) is OpenCV's 15-bit fixed point (B 3735, G 19235, R 9798, + 16384 >> 15; checked exhaustively
// against cv2 4.13 over all 2^24 colours); `/ 255` is a float64 division
// narrowed to fp32; blendLinear is (s1 * w1 + s2 * w2) / (w1 + w2 + 1e-5f) in fp32 (products rounded separately, no
// FMA contraction) rounded to nearest-even and saturated.
struct BlendParams {
    const uint8_t *frames;   // [n_frames, H, W, 3]
    const uint8_t *faces;    // [B, S, S, 3]
    const uint8_t *masks;    // packed per-avatar-frame masks, BGR u8 [ye - ys, xe - xs, 3] at byte offset moff[b]
    uint8_t *out;            // [B, H, W, 3]
    int H, W, S, B;
    int idx[PASTE_MAX_BATCH];
    int y1[PASTE_MAX_BATCH], y2[PASTE_MAX_BATCH], x1[PASTE_MAX_BATCH], x2[PASTE_MAX_BATCH];
    int ys[PASTE_MAX_BATCH], ye[PASTE_MAX_BATCH], xs[PASTE_MAX_BATCH], xe[PASTE_MAX_BATCH];
    long long moff[PASTE_MAX_BATCH];
};

__global__ void __launch_bounds__(256) k_paste_blend(const __grid_constant__ BlendParams p) {
    const int b = blockIdx.y;
    const int pix = blockIdx.x * blockDim.x + threadIdx.x;
    if (pix >= p.H * p.W) return;
    const int y = pix / p.W, x = pix - y * p.W;
    const uint8_t *src = p.frames + ((size_t)p.idx[b] * p.H * p.W + pix) * 3;
    uint8_t *dst = p.out + ((size_t)b * p.H * p.W + pix) * 3;
    const uint8_t body[3] = {src[0], src[1], src[2]};
    const int ys = p.ys[b], ye = p.ye[b], xs = p.xs[b], xe = p.xe[b];
    if (y < ys || y >= ye || x < xs || x >= xe) {
        dst[0] = body[0]; dst[1] = body[1]; dst[2] = body[2];
        return;
    }
    const int y1 = p.y1[b], y2 = p.y2[b], x1 = p.x1[b], x2 = p.x2[b];
    uint8_t fl[3] = {body[0], body[1], body[2]};
    if (y >= y1 && y < y2 && x >= x1 && x < x2)
        cv_resize_px(p.faces + (size_t)b * p.S * p.S * 3, p.S, x2 - x1, y2 - y1, x - x1, y - y1, fl);
    const uint8_t *m = p.masks + p.moff[b] + ((size_t)(y - ys) * (xe - xs) + (x - xs)) * 3;
    const int gray = (m[0] * 3735 + m[1] * 19235 + m[2] * 9798 + (1 << 14)) >> 15;
    const float w1 = (float)((double)gray / 255.0);
    const float w2 = 1.0f - w1;
    const float den = __fadd_rn(__fadd_rn(w1, w2), 1e-5f);
#pragma unroll
    for (int c = 0; c < 3; c++) {
        const float num = __fadd_rn(__fmul_rn((float)fl[c], w1), __fmul_rn((float)body[c], w2));
        const int v = __float2int_rn(__fdiv_rn(num, den));
        dst[c] = (uint8_t)max(0, min(255, v));
    }
}

extern "C" int mf_paste_resize_u8(mf_ctx *ctx, const uint8_t *frames, int n_frames, int H, int W, const uint8_t *faces,
                                  int S, int B, const int32_t *idx_bbox_host, uint8_t *out, void *stream) {
    if (!ctx) return MF_E_INVALID;
    MF_REQUIRE(ctx, frames && faces && out && idx_bbox_host, "mf_paste_resize_u8: null pointer");
    MF_REQUIRE(ctx, B >= 1 && B <= PASTE_MAX_BATCH && H > 0 && W > 0 && S > 1 && n_frames > 0, "mf_paste_resize_u8: bad sizes");
    PasteParams p;
    p.frames = frames; p.faces = faces; p.out = out; p.H = H; p.W = W; p.S = S; p.B = B;
    for (int i = 0; i < B; i++) {
        const int32_t *r = idx_bbox_host + i * 5;
        MF_REQUIRE(ctx, r[0] >= 0 && r[0] < n_frames, "mf_paste_resize_u8: frame index %d out of range", r[0]);
        MF_REQUIRE(ctx, 0 <= r[1] && r[1] < r[2] && r[2] <= H && 0 <= r[3] && r[3] < r[4] && r[4] <= W,
                   "mf_paste_resize_u8: bbox (y1,y2,x1,x2)=(%d,%d,%d,%d) outside the %dx%d frame", r[1], r[2], r[3], r[4], H, W);
        p.idx[i] = r[0]; p.y1[i] = r[1]; p.y2[i] = r[2]; p.x1[i] = r[3]; p.x2[i] = r[4];
    }
    MF_CUDA(ctx, cudaSetDevice(ctx->device));   // the caller's thread may sit on another GPU (in-process multi-GPU scheduler)
    dim3 grid((H * W + 255) / 256, B);
    k_paste_resize<<<grid, 256, 0, (cudaStream_t)stream>>>(p);
    MF_CUDA(ctx, cudaGetLastError());
    return MF_OK;
}

// rows_host: B x 9 int32 (frame index, face bbox y1, y2, x1, x2, mask crop box ys, ye, xs, xe); mask_off_host: B int64
// byte offsets into `masks` (mask of row i is BGR u8 [ye - ys, xe - xs, 3])
extern "C" int mf_paste_blend_u8(mf_ctx *ctx, const uint8_t *frames, int n_frames, int H, int W, const uint8_t *faces,
                                 int S, int B, const int32_t *rows_host, const uint8_t *masks, size_t masks_nbytes,
                                 const int64_t *mask_off_host, uint8_t *out, void *stream) {
    if (!ctx) return MF_E_INVALID;
    MF_REQUIRE(ctx, frames && faces && out && rows_host && masks && mask_off_host, "mf_paste_blend_u8: null pointer");
    MF_REQUIRE(ctx, B >= 1 && B <= PASTE_MAX_BATCH && H > 0 && W > 0 && S > 1 && n_frames > 0, "mf_paste_blend_u8: bad sizes");
    BlendParams p;
    p.frames = frames; p.faces = faces; p.masks = masks; p.out = out; p.H = H; p.W = W; p.S = S; p.B = B;
    for (int i = 0; i < B; i++) {
        const int32_t *r = rows_host + i * 9;
        MF_REQUIRE(ctx, r[0] >= 0 && r[0] < n_frames, "mf_paste_blend_u8: frame index %d out of range", r[0]);
        MF_REQUIRE(ctx, 0 <= r[5] && r[5] < r[6] && r[6] <= H && 0 <= r[7] && r[7] < r[8] && r[8] <= W,
                   "mf_paste_blend_u8: crop box (ys,ye,xs,xe)=(%d,%d,%d,%d) outside the %dx%d frame", r[5], r[6], r[7], r[8], H, W);
        MF_REQUIRE(ctx, r[5] <= r[1] && r[1] < r[2] && r[2] <= r[6] && r[7] <= r[3] && r[3] < r[4] && r[4] <= r[8],
                   "mf_paste_blend_u8: face bbox (y1,y2,x1,x2)=(%d,%d,%d,%d) outside its crop box", r[1], r[2], r[3], r[4]);
        const long long need = (long long)(r[6] - r[5]) * (r[8] - r[7]) * 3;
        MF_REQUIRE(ctx, mask_off_host[i] >= 0 && (size_t)(mask_off_host[i] + need) <= masks_nbytes,
                   "mf_paste_blend_u8: mask %d exceeds the mask buffer", i);
        p.idx[i] = r[0]; p.y1[i] = r[1]; p.y2[i] = r[2]; p.x1[i] = r[3]; p.x2[i] = r[4];
        p.ys[i] = r[5]; p.ye[i] = r[6]; p.xs[i] = r[7]; p.xe[i] = r[8];
        p.moff[i] = mask_off_host[i];
    }
    MF_CUDA(ctx, cudaSetDevice(ctx->device));
    dim3 grid((H * W + 255) / 256, B);
    k_paste_blend<<<grid, 256, 0, (cudaStream_t)stream>>>(p);
    MF_CUDA(ctx, cudaGetLastError());
    return MF_OK;
}

// =====================================================================================================
// Wav2Lip mel front-end on the GPU: replaces audio.melspectrogram (wav2lip/audio.py:45-51 with :20-23 preemphasis,
// :57-61 librosa.stft n_fft 800 / hop 200 / Hann, :92-122 mel basis 80 x 401 (55-7600 Hz), amp_to_db, normalize) and the chunk
// slicing of LipASR.run_step (lipasr.py:24-35) for one l + 2B + r window: float audio in, [B,1,80,16] mel chunks out, in
// the layout mf_wav2lip_forward consumes.  Direct 800-point DFT per frame (84 frames per window: 27 MMAC).
// librosa >= 0.10 semantics (centre padding with zeros), like mere_fusion_b200/audio_mel.py: PARITY UNPINNED.
// =====================================================================================================
#define WM_NFFT 800
#define WM_HOP 200
#define WM_BINS 401
#define WM_MELS 80

struct W2lMelParams {
    const float *audio;      // device fp32 [n_samples]
    const float *filters;    // device fp32 [80][401]
    float *mel;              // scratch [n_frames][80]
    float *out;              // [B][1][80][16]
    int n_samples, n_frames, B;
    int start[PASTE_MAX_BATCH];
};

__global__ void __launch_bounds__(256) k_w2l_mel_frames(const __grid_constant__ W2lMelParams p) {
    __shared__ float xs[WM_NFFT], ct[WM_NFFT], st[WM_NFFT], mag[WM_BINS + 3];
    const int t = blockIdx.x;
    for (int i = threadIdx.x; i < WM_NFFT; i += blockDim.x) {
        float sn, cs;
        sincospif((float)i / 400.0f, &sn, &cs);   // 2 pi i / 800
        ct[i] = cs; st[i] = sn;
        const int src = t * WM_HOP + i - WM_NFFT / 2;   // centred frame, zero padding
        float v = 0.f;
        if (src >= 0 && src < p.n_samples) v = p.audio[src] - (src > 0 ? 0.97f * p.audio[src - 1] : 0.f);   // lfilter([1, -0.97], [1], wav)
        xs[i] = v * (0.5f - 0.5f * cs);           // periodic Hann(800)
    }
    __syncthreads();
    for (int k = threadIdx.x; k < WM_BINS; k += blockDim.x) {
        float re = 0.f, im = 0.f;
        int idx = 0;
        for (int n = 0; n < WM_NFFT; n++) {
            re = fmaf(xs[n], ct[idx], re);
            im = fmaf(xs[n], st[idx], im);
            idx += k;
            if (idx >= WM_NFFT) idx -= WM_NFFT;
        }
        mag[k] = sqrtf(re * re + im * im);
    }
    __syncthreads();
    if (threadIdx.x < WM_MELS) {
        const float *f = p.filters + threadIdx.x * WM_BINS;
        float acc = 0.f;
        for (int k = 0; k < WM_BINS; k++) acc = fmaf(__ldg(f + k), mag[k], acc);
        const float db = 20.0f * log10f(fmaxf(1e-5f, acc)) - 20.0f;                       // _amp_to_db - ref_level_db
        p.mel[t * WM_MELS + threadIdx.x] = fminf(fmaxf(8.0f * ((db + 100.0f) / 100.0f) - 4.0f, -4.0f), 4.0f);   // _normalize
    }
}
__global__ void __launch_bounds__(256) k_w2l_mel_chunks(const __grid_constant__ W2lMelParams p) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= p.B * WM_MELS * 16) return;
    const int c = i & 15, m = (i >> 4) % WM_MELS, b = i / (16 * WM_MELS);
    p.out[i] = p.mel[(p.start[b] + c) * WM_MELS + m];
}

extern "C" int mf_wav2lip_mel_chunks(mf_ctx *ctx, const float *audio, int n_samples, const float *filters, const int32_t *start_idx_host,
                                     int B, float *out, void *stream) {
    if (!ctx) return MF_E_INVALID;
    MF_REQUIRE(ctx, audio && filters && start_idx_host && out, "mf_wav2lip_mel_chunks: null pointer");
    MF_REQUIRE(ctx, n_samples >= 1 && n_samples <= 16000 * 60 && B >= 1 && B <= PASTE_MAX_BATCH, "mf_wav2lip_mel_chunks: bad sizes");
    W2lMelParams p;
    p.audio = audio; p.filters = filters; p.out = out; p.n_samples = n_samples; p.B = B;
    p.n_frames = n_samples / WM_HOP + 1;
    for (int i = 0; i < B; i++) {
        MF_REQUIRE(ctx, start_idx_host[i] >= 0 && start_idx_host[i] + 16 <= p.n_frames, "mf_wav2lip_mel_chunks: chunk %d starts at column %d of %d",
                   i, start_idx_host[i], p.n_frames);
        p.start[i] = start_idx_host[i];
    }
    MF_CUDA(ctx, cudaSetDevice(ctx->device));
    if (p.n_frames > ctx->mel_frames_cap) {
        MF_CUDA(ctx, cudaDeviceSynchronize());
        cudaFree(ctx->mel_scratch);
        ctx->mel_scratch = nullptr;
        ctx->mel_frames_cap = 0;
        MF_CUDA(ctx, cudaMalloc(&ctx->mel_scratch, (size_t)p.n_frames * WM_MELS * sizeof(float)));
        ctx->mel_frames_cap = p.n_frames;
    }
    p.mel = ctx->mel_scratch;
    k_w2l_mel_frames<<<p.n_frames, 256, 0, (cudaStream_t)stream>>>(p);
    k_w2l_mel_chunks<<<(B * WM_MELS * 16 + 255) / 256, 256, 0, (cudaStream_t)stream>>>(p);
    MF_CUDA(ctx, cudaGetLastError());
    return MF_OK;
}
