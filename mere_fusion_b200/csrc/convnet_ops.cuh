// convnet_ops.cuh -- the non-conv ops of the conv-net executor (csrc/wav2lip.cu), needed by the MuseTalk
// UNet / VAE decoder: GroupNorm(+SiLU), LayerNorm, batched attention GEMMs (warp-level mma.sync on bf16),
// row softmax, GEGLU, and the input-preparation kernels.  All tensors are NHWC bf16 ("tokens x channels").
// Every kernel takes exactly one by-value parameter struct so that the host can treat launches uniformly
// (direct launch or CUDA-graph node).
#pragma once
#include "mf_common.cuh"

// ---------------------------------------------------------------------------------------------------
// input preparation
// ---------------------------------------------------------------------------------------------------
struct PrepParams {
    const void *src;
    void *dst;
    int a, b, c, d;
};

// wav2lip faces u8 [B,S,S,3] BGR -> bf16 [B,S,S,8]: ch 0-2 = face with rows >= S/2 zeroed, ch 3-5 = face, /255
// (lipreal.py:108-122).  a = B, b = S
__global__ void k_prep_face(const PrepParams p) {
    pdl_launch();
    pdl_wait();
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    const int S = p.b;
    if (i >= p.a * S * S) return;
    const uint8_t *faces = reinterpret_cast<const uint8_t *>(p.src);
    const int row = (i / S) % S;
    float v[8];
#pragma unroll
    for (int c = 0; c < 3; c++) {
        const float f = (float)faces[(size_t)i * 3 + c] / 255.f;
        v[c] = row >= S / 2 ? 0.f : f;
        v[3 + c] = f;
    }
    v[6] = v[7] = 0.f;
    uint32_t o[4];
#pragma unroll
    for (int j = 0; j < 4; j++) {
        __nv_bfloat162 h = __floats2bfloat162_rn(v[2 * j], v[2 * j + 1]);
        o[j] = *reinterpret_cast<uint32_t *>(&h);
    }
    reinterpret_cast<uint4 *>(p.dst)[i] = make_uint4(o[0], o[1], o[2], o[3]);
}
// wav2lip mel fp32 [B,1,80,16] -> bf16 [B,80,16,8] (channel 0).  a = element count
__global__ void k_prep_mel(const PrepParams p) {
    pdl_launch();
    pdl_wait();
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= p.a) return;
    __nv_bfloat162 h = __floats2bfloat162_rn(reinterpret_cast<const float *>(p.src)[i], 0.f);
    reinterpret_cast<uint4 *>(p.dst)[i] = make_uint4(*reinterpret_cast<uint32_t *>(&h), 0u, 0u, 0u);
}
// musetalk latents fp16 NCHW [B,C,H,W] -> bf16 NHWC [B,H,W,Cpad].  a = B, b = C, c = H*W, d = Cpad
__global__ void k_prep_latents(const PrepParams p) {
    pdl_launch();
    pdl_wait();
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= p.a * p.c * p.d) return;
    const int ch = i % p.d, pix = (i / p.d) % p.c, b = i / (p.d * p.c);
    const __half *src = reinterpret_cast<const __half *>(p.src);
    const float v = ch < p.b ? __half2float(src[((size_t)b * p.b + ch) * p.c + pix]) : 0.f;
    reinterpret_cast<__nv_bfloat16 *>(p.dst)[i] = __float2bfloat16_rn(v);
}
// musetalk audio context: whisper fp16 [B,T,D] + sinusoidal PE in fp16 (musetalk/models/unet.py:12-27 with
// pe.half(), musereal.py:60,102) -> bf16 [B,T,1,D].  a = B, b = T, c = D
__global__ void k_prep_ctx(const PrepParams p) {
    pdl_launch();
    pdl_wait();
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= p.a * p.b * p.c) return;
    const int dch = i % p.c, t = (i / p.c) % p.b;
    const float div = expf((float)(dch & ~1) * (-logf(10000.0f) / (float)p.c));
    const float ang = (float)t * div;
    const float pe = (dch & 1) ? cosf(ang) : sinf(ang);
    const __half x = reinterpret_cast<const __half *>(p.src)[i];
    const __half s = __hadd(x, __float2half_rn(pe));
    reinterpret_cast<__nv_bfloat16 *>(p.dst)[i] = __float2bfloat16_rn(__half2float(s));
}
// ---------------------------------------------------------------------------------------------------
// Whisper front-end (musetalk/whisper/whisper/audio.py:92-125 log_mel_spectrogram + transcribe.py:99 pad_or_trim):
// centred STFT (n_fft 400, hop 160, periodic Hann, reflect padding) as a direct 400-point DFT per frame (104 frames
// for the 52-chunk live window: 33 MFLOP, not worth an FFT), power, 80-band mel filterbank, log10, and in a second
// pass max(., global max - 8), (x + 4) / 4 and zero padding to the 3000-frame context.
// ---------------------------------------------------------------------------------------------------
struct WhisperPrep {
    const float *audio;       // device fp32 [n_samples]
    float *logspec;           // scratch fp32 [n_ctx_frames][80]
    const float *filters;     // fp32 [80][201]
    int *maxslot;             // float bits of (max log10 + 16); reset to 0 by k_whisper_gather at the end of the program
    __nv_bfloat16 *melbuf;    // program input buffer [n_ctx_frames][1][80]
    int n_samples, n_frames, n_ctx_frames;
};
#define WH_NFFT 400
#define WH_HOP 160
#define WH_BINS 201
#define WH_MELS 80

__global__ void __launch_bounds__(256) k_logmel_frames(const WhisperPrep p) {
    pdl_launch();
    pdl_wait();
    __shared__ float xs[WH_NFFT], ct[WH_NFFT], st[WH_NFFT], pw[WH_BINS + 7];
    __shared__ float red[8];
    const int t = blockIdx.x;
    for (int i = threadIdx.x; i < WH_NFFT; i += blockDim.x) {
        float sn, cs;
        sincospif((float)i / 200.0f, &sn, &cs);   // 2 pi i / 400
        ct[i] = cs; st[i] = sn;
        int src = t * WH_HOP + i - WH_NFFT / 2;   // torch.stft(center=True, pad_mode="reflect")
        if (src < 0) src = -src;
        if (src >= p.n_samples) src = 2 * (p.n_samples - 1) - src;
        xs[i] = p.audio[src] * (0.5f - 0.5f * cs);  // torch.hann_window(400) (periodic)
    }
    __syncthreads();
    if (threadIdx.x < WH_BINS) {
        const int k = threadIdx.x;
        float re = 0.f, im = 0.f;
        int idx = 0;
        for (int n = 0; n < WH_NFFT; n++) {
            re = fmaf(xs[n], ct[idx], re);
            im = fmaf(xs[n], st[idx], im);
            idx += k;
            if (idx >= WH_NFFT) idx -= WH_NFFT;
        }
        pw[k] = re * re + im * im;
    }
    __syncthreads();
    float v = -1e30f;
    if (threadIdx.x < WH_MELS) {
        const float *f = p.filters + threadIdx.x * WH_BINS;
        float acc = 0.f;
        for (int k = 0; k < WH_BINS; k++) acc = fmaf(__ldg(f + k), pw[k], acc);
        v = log10f(fmaxf(acc, 1e-10f));
        p.logspec[t * WH_MELS + threadIdx.x] = v;
    }
#pragma unroll
    for (int o = 16; o; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = v;
    __syncthreads();
    if (threadIdx.x == 0) {
        for (int w = 1; w < 8; w++) v = fmaxf(v, red[w]);
        atomicMax(p.maxslot, __float_as_int(v + 16.0f));   // v >= -10: positive floats order like ints
    }
}
__global__ void __launch_bounds__(256) k_logmel_finish(const WhisperPrep p) {
    pdl_launch();
    pdl_wait();
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= p.n_ctx_frames * WH_MELS) return;
    const int t = i / WH_MELS;
    float v = 0.f;                                         // pad_or_trim pads the finished log-mel with zeros
    if (t < p.n_frames) {
        const float mx = __int_as_float(*p.maxslot) - 16.0f;
        v = (fmaxf(p.logspec[i], mx - 8.0f) + 4.0f) / 4.0f;
    }
    p.melbuf[i] = __float2bfloat16_rn(v);
}
// embeddings = stack([x0, block outputs...], axis=1) -> transpose(0,2,1,3) -> first T rows (model.py:155-167,
// audio2feature.py:103-110): out fp32 [T][n_src][C]
struct GatherParams {
    const __nv_bfloat16 *src[8];
    float *out;
    int *maxslot;
    int n_src, T, C;
};
__global__ void __launch_bounds__(256) k_whisper_gather(const GatherParams p) {
    pdl_launch();
    pdl_wait();
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i == 0) *p.maxslot = 0;
    if (i >= p.T * p.n_src * p.C) return;
    const int c = i % p.C, j = (i / p.C) % p.n_src, t = i / (p.C * p.n_src);
    p.out[i] = __bfloat162float(p.src[j][(size_t)t * p.C + c]);
}

// ---------------------------------------------------------------------------------------------------
// wav2vec2 front-end (HF Wav2Vec2FeatureExtractor do_normalize + the first conv layer of Wav2Vec2FeatureEncoder): the waveform is
// normalised to zero mean / unit variance ((x - mean) / sqrt(var + 1e-7)) and convolved (Cin = 1, k = 10, stride 5) in fp32 --
// the raw audio never gets rounded to bf16; the 512-channel output is the program's first bf16 buffer.
// ---------------------------------------------------------------------------------------------------
struct W2vPrep {
    const float *audio;     // device fp32 [B][n_samples]: one window per batch item (blockIdx.y)
    const float *conv0;     // fp32 [C0][k0] then [C0] bias
    float *stats;           // [B] x (mean, rstd)
    __nv_bfloat16 *out;     // [B][n_frames][1][C0]
    int n_samples, n_frames, C0, k0, s0;
};
__global__ void __launch_bounds__(1024) k_w2v_stats(const W2vPrep p) {
    pdl_launch();
    pdl_wait();
    __shared__ double red[32];
    double s = 0.0, q = 0.0;
    const float *audio = p.audio + (size_t)blockIdx.y * p.n_samples;
    for (int i = threadIdx.x; i < p.n_samples; i += blockDim.x) { const double v = audio[i]; s += v; q += v * v; }
    for (int o = 16; o; o >>= 1) { s += __shfl_xor_sync(0xffffffffu, s, o); q += __shfl_xor_sync(0xffffffffu, q, o); }
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = s;
    __syncthreads();
    double S = 0.0;
    if (threadIdx.x < 32) { S = red[threadIdx.x]; for (int o = 16; o; o >>= 1) S += __shfl_xor_sync(0xffffffffu, S, o); }
    __syncthreads();
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = q;
    __syncthreads();
    if (threadIdx.x < 32) {
        double Q = red[threadIdx.x];
        for (int o = 16; o; o >>= 1) Q += __shfl_xor_sync(0xffffffffu, Q, o);
        if (threadIdx.x == 0) {
            const double mean = S / p.n_samples, var = Q / p.n_samples - mean * mean;   // numpy var (population)
            p.stats[2 * blockIdx.y] = (float)mean;
            p.stats[2 * blockIdx.y + 1] = (float)(1.0 / sqrt(var + 1e-7));
        }
    }
}
// 8 output frames per CTA, one thread per (frame, channel pair ...): weights in shared memory
__global__ void __launch_bounds__(256) k_w2v_conv0(const W2vPrep p) {
    pdl_launch();
    pdl_wait();
    __shared__ float xs[7 * 16 + 16];   // s0, k0 <= 16 (checked at load)
    const int t0 = blockIdx.x * 8;
    const float mean = p.stats[2 * blockIdx.y], rstd = p.stats[2 * blockIdx.y + 1];
    const float *audio = p.audio + (size_t)blockIdx.y * p.n_samples;
    const int span = 7 * p.s0 + p.k0;   // samples touched by the CTA's 8 frames
    for (int i = threadIdx.x; i < span; i += blockDim.x) {
        const int src = t0 * p.s0 + i;
        xs[i] = src < p.n_samples ? (audio[src] - mean) * rstd : 0.f;
    }
    __syncthreads();
    for (int o = threadIdx.x; o < 8 * p.C0; o += blockDim.x) {
        const int c = o % p.C0, f = o / p.C0;
        if (t0 + f >= p.n_frames) continue;
        const float *w = p.conv0 + c * p.k0;
        float acc = __ldg(p.conv0 + p.C0 * p.k0 + c);
        for (int k = 0; k < p.k0; k++) acc = fmaf(__ldg(w + k), xs[f * p.s0 + k], acc);
        p.out[((size_t)blockIdx.y * p.n_frames + t0 + f) * p.C0 + c] = __float2bfloat16_rn(acc);
    }
}

__global__ void k_f32_to_bf16(const PrepParams p) {
    pdl_launch();
    pdl_wait();
    const size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x;
    const size_t n = ((size_t)(uint32_t)p.a) | ((size_t)(uint32_t)p.b << 32);
    if (i < n) reinterpret_cast<__nv_bfloat16 *>(p.dst)[i] = __float2bfloat16_rn(reinterpret_cast<const float *>(p.src)[i]);
}
__global__ void k_bf16_to_f32(const PrepParams p) {
    pdl_launch();
    pdl_wait();
    const size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x;
    const size_t n = ((size_t)(uint32_t)p.a) | ((size_t)(uint32_t)p.b << 32);
    if (i < n) reinterpret_cast<float *>(p.dst)[i] = __bfloat162float(reinterpret_cast<const __nv_bfloat16 *>(p.src)[i]);
}

// ---------------------------------------------------------------------------------------------------
// SiLU y * sigmoid(y) = 0.5 y (1 + tanh(0.5 y)) with ONE MUFU (tanh.approx.f32, relative error 2^-11: below the bf16 rounding of the
// result) and three FP32 ops.  The IEEE division of y / (1 + exp(-y)) made the apply pass instruction-bound (~20 instructions per
// element: 134 M elements of a 128-channel 256x256 x 16 tensor cost as much issue time as their 537 MB cost HBM time).
__device__ __forceinline__ float silu_fast(float y) {
    const float h = 0.5f * y;
    float t;
    asm("tanh.approx.f32 %0, %1;" : "=f"(t) : "f"(h));
    return fmaf(h, t, h);
}

// GroupNorm (+ SiLU).  Deterministic: every CTA of k_gn_stats reduces its pixel range in a fixed order and
// writes one partial (sum, sum of squares) per group; the CTA that arrives last at a per-batch-item counter
// adds the partials in index order and turns them into per-channel coefficients (a, b) with
// y = x * a + b  (a = rstd * gamma, b = beta - mean * a).  k_gn_apply is then one fused multiply-add per
// element.  Run-to-run results are bit-identical (no floating-point atomics anywhere).
// ---------------------------------------------------------------------------------------------------
#define GN_MAX_C 2560
struct NormParams {
    const __nv_bfloat16 *in;
    __nv_bfloat16 *out;
    const float *gamma, *beta;
    float *coef;          // [B][C][2] per-channel (a, b); scratch shared by all GroupNorm ops (ops run in order)
    float *partial;       // [B][gridDim.x][G][2] scratch
    unsigned *counter;    // [B], zero between launches (the finalising CTA resets it)
    int npix, C, G, silu;
    float eps;
    int pix_per_cta;
    int in_stride, in_coff;  // GroupNorm input may be a channel range of a wider (concat) buffer; output is dense
};

// grid (ceil(npix / pix_per_cta), B).  A thread owns fixed 8-channel chunk columns (so the sums of its channels
// stay in registers) and walks the CTA's pixel range with a stride.
__global__ void __launch_bounds__(256) k_gn_stats(const NormParams p) {
    pdl_launch();
    pdl_wait();
    __shared__ float part[2][GN_MAX_C];
    __shared__ float red[8][128];
    __shared__ float mr[64][2];
    __shared__ bool is_last;
    const int chunks = p.C >> 3;
    const int cpp = min(chunks, (int)blockDim.x);          // chunk columns handled per pass
    const int lanes = ((int)blockDim.x / cpp) * cpp;
    const int pstep = lanes / cpp;
    const int cpg = p.C / p.G;
    const int b = blockIdx.y;
    if ((int)threadIdx.x < lanes) {
        const int prow = threadIdx.x / cpp;
        const int p0 = blockIdx.x * p.pix_per_cta, p1 = min(p.npix, p0 + p.pix_per_cta);
        for (int chunk = threadIdx.x % cpp; chunk < chunks; chunk += cpp) {
            const __nv_bfloat16 *base = p.in + (size_t)b * p.npix * p.in_stride + p.in_coff + chunk * 8;
            float s[8], q[8];
#pragma unroll
            for (int j = 0; j < 8; j++) s[j] = q[j] = 0.f;
            auto acc8 = [&](const uint4 &v) {
                const uint32_t w[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
                for (int j = 0; j < 4; j++) {
                    const __nv_bfloat162 h = *reinterpret_cast<const __nv_bfloat162 *>(&w[j]);
                    const float a = __bfloat162float(h.x), c = __bfloat162float(h.y);
                    s[2 * j] += a; q[2 * j] += a * a;
                    s[2 * j + 1] += c; q[2 * j + 1] += c * c;
                }
            };
            int px = p0 + prow;
            for (; px + 3 * pstep < p1; px += 4 * pstep) {   // 4 independent 16-B loads in flight per thread
                const uint4 v0 = __ldg(reinterpret_cast<const uint4 *>(base + (size_t)px * p.in_stride));
                const uint4 v1 = __ldg(reinterpret_cast<const uint4 *>(base + (size_t)(px + pstep) * p.in_stride));
                const uint4 v2 = __ldg(reinterpret_cast<const uint4 *>(base + (size_t)(px + 2 * pstep) * p.in_stride));
                const uint4 v3 = __ldg(reinterpret_cast<const uint4 *>(base + (size_t)(px + 3 * pstep) * p.in_stride));
                acc8(v0); acc8(v1); acc8(v2); acc8(v3);
            }
            for (; px < p1; px += pstep) acc8(__ldg(reinterpret_cast<const uint4 *>(base + (size_t)px * p.in_stride)));
            // pstep > 1 only when all chunks fit one pass (C = cpp * 8), so prow * C + c < 2048
#pragma unroll
            for (int j = 0; j < 8; j++) {
                part[0][prow * p.C + chunk * 8 + j] = s[j];
                part[1][prow * p.C + chunk * 8 + j] = q[j];
            }
        }
    }
    __syncthreads();
    float *mine = p.partial + ((size_t)b * gridDim.x + blockIdx.x) * p.G * 2;
    if ((int)threadIdx.x < 2 * p.G) {
        const int g = threadIdx.x >> 1, st = threadIdx.x & 1;
        float t = 0.f;
        for (int r = 0; r < pstep; r++)
            for (int c = 0; c < cpg; c++) t += part[st][r * p.C + g * cpg + c];
        mine[threadIdx.x] = t;
    }
    __threadfence();
    __syncthreads();
    if (threadIdx.x == 0) is_last = atomicAdd(p.counter + b, 1u) == gridDim.x - 1;
    __syncthreads();
    if (!is_last) return;
    __threadfence();
    // ---- finalise (one CTA per batch item): partials in index order -> mean / rstd -> per-channel coefficients
    const int items = 2 * p.G, nl = min(8, (int)blockDim.x / items);   // G <= 64: nl >= 2
    const int item = threadIdx.x % items, ln = threadIdx.x / items;
    if (ln < nl) {
        const volatile float *pp = p.partial + (size_t)b * gridDim.x * items;
        float t = 0.f;
        for (int c = ln; c < (int)gridDim.x; c += nl) t += pp[(size_t)c * items + item];
        red[ln][item] = t;
    }
    __syncthreads();
    if ((int)threadIdx.x < p.G) {
        float s = 0.f, q = 0.f;
        for (int l = 0; l < nl; l++) { s += red[l][2 * threadIdx.x]; q += red[l][2 * threadIdx.x + 1]; }
        const float inv_n = 1.0f / ((float)p.npix * (float)cpg);
        const float mean = s * inv_n;
        const float var = fmaxf(q * inv_n - mean * mean, 0.f);
        mr[threadIdx.x][0] = mean;
        mr[threadIdx.x][1] = rsqrtf(var + p.eps);
    }
    __syncthreads();
    for (int c = threadIdx.x; c < p.C; c += blockDim.x) {
        const int g = c / cpg;
        const float a = mr[g][1] * __ldg(p.gamma + c);
        reinterpret_cast<float2 *>(p.coef)[(size_t)b * p.C + c] = make_float2(a, __ldg(p.beta + c) - mr[g][0] * a);
    }
    if (threadIdx.x == 0) p.counter[b] = 0u;
}

// GroupNorm statistics fused into the producing conv's epilogue (k_conv_tma writes one (sum, sum of squares) per (image, tile,
// 32-row quarter, group) into fixed slots): this kernel adds the slots of a batch item in index order and writes the per-channel
// coefficients, replacing the k_gn_stats pass (one full read of the tensor).  grid B, block GN_FIN_THREADS.
#define GN_FIN_THREADS 1024
struct GnFinalParams {
    const float *partial;   // [B][slots][G][2]
    const float *gamma, *beta;
    float *coef;            // [B][C][2]
    int slots, C, G, npix;
    float eps;
};
__global__ void __launch_bounds__(GN_FIN_THREADS) k_gn_finalize(const GnFinalParams p) {
    pdl_launch();
    pdl_wait();
    __shared__ float red[16][128];
    __shared__ float mr[64][2];
    const int b = blockIdx.x, items = 2 * p.G;
    const int nl = min(16, (int)blockDim.x / items);
    const int item = threadIdx.x % items, ln = threadIdx.x / items;
    if (ln < nl) {
        const float *pp = p.partial + (size_t)b * p.slots * items;
        float t = 0.f;
        for (int c = ln; c < p.slots; c += nl) t += __ldcg(pp + (size_t)c * items + item);
        red[ln][item] = t;
    }
    __syncthreads();
    const int cpg = p.C / p.G;
    if ((int)threadIdx.x < p.G) {
        float s = 0.f, q = 0.f;
        for (int l = 0; l < nl; l++) { s += red[l][2 * threadIdx.x]; q += red[l][2 * threadIdx.x + 1]; }
        const float inv_n = 1.0f / ((float)p.npix * (float)cpg);
        const float mean = s * inv_n;
        mr[threadIdx.x][0] = mean;
        mr[threadIdx.x][1] = rsqrtf(fmaxf(q * inv_n - mean * mean, 0.f) + p.eps);
    }
    __syncthreads();
    for (int c = threadIdx.x; c < p.C; c += blockDim.x) {
        const int g = c / cpg;
        const float a = mr[g][1] * __ldg(p.gamma + c);
        reinterpret_cast<float2 *>(p.coef)[(size_t)b * p.C + c] = make_float2(a, __ldg(p.beta + c) - mr[g][0] * a);
    }
}

// grid (ceil(npix / pix_per_cta), B): same thread -> channel-chunk mapping as the statistics pass, so the 16
// coefficients of a thread's chunk are loaded once and the pixel loop is pure streaming (16-B loads / stores)
__global__ void __launch_bounds__(256) k_gn_apply(const NormParams p) {
    pdl_launch();
    pdl_wait();
    const int chunks = p.C >> 3;
    const int cpp = min(chunks, (int)blockDim.x);
    const int lanes = ((int)blockDim.x / cpp) * cpp;
    if ((int)threadIdx.x >= lanes) return;
    const int pstep = lanes / cpp, prow = threadIdx.x / cpp;
    const int b = blockIdx.y;
    const int p0 = blockIdx.x * p.pix_per_cta, p1 = min(p.npix, p0 + p.pix_per_cta);
    for (int chunk = threadIdx.x % cpp; chunk < chunks; chunk += cpp) {
        float ca[8], cb[8];
        const float4 *cf = reinterpret_cast<const float4 *>(p.coef + ((size_t)b * p.C + chunk * 8) * 2);
#pragma unroll
        for (int j = 0; j < 4; j++) {
            const float4 t = cf[j];
            ca[2 * j] = t.x; cb[2 * j] = t.y; ca[2 * j + 1] = t.z; cb[2 * j + 1] = t.w;
        }
        const __nv_bfloat16 *ib = p.in + (size_t)b * p.npix * p.in_stride + p.in_coff + chunk * 8;
        __nv_bfloat16 *ob = p.out + (size_t)b * p.npix * p.C + chunk * 8;
        auto apply8 = [&](const uint4 &v, int px) {
            const uint32_t w[4] = {v.x, v.y, v.z, v.w};
            uint32_t o[4];
#pragma unroll
            for (int j = 0; j < 4; j++) {
                const __nv_bfloat162 h = *reinterpret_cast<const __nv_bfloat162 *>(&w[j]);
                float y0 = fmaf(__bfloat162float(h.x), ca[2 * j], cb[2 * j]);
                float y1 = fmaf(__bfloat162float(h.y), ca[2 * j + 1], cb[2 * j + 1]);
                if (p.silu) { y0 = silu_fast(y0); y1 = silu_fast(y1); }
                __nv_bfloat162 r = __floats2bfloat162_rn(y0, y1);
                o[j] = *reinterpret_cast<uint32_t *>(&r);
            }
            *reinterpret_cast<uint4 *>(ob + (size_t)px * p.C) = make_uint4(o[0], o[1], o[2], o[3]);
        };
        int px = p0 + prow;
        for (; px + 3 * pstep < p1; px += 4 * pstep) {
            const uint4 v0 = __ldg(reinterpret_cast<const uint4 *>(ib + (size_t)px * p.in_stride));
            const uint4 v1 = __ldg(reinterpret_cast<const uint4 *>(ib + (size_t)(px + pstep) * p.in_stride));
            const uint4 v2 = __ldg(reinterpret_cast<const uint4 *>(ib + (size_t)(px + 2 * pstep) * p.in_stride));
            const uint4 v3 = __ldg(reinterpret_cast<const uint4 *>(ib + (size_t)(px + 3 * pstep) * p.in_stride));
            apply8(v0, px); apply8(v1, px + pstep); apply8(v2, px + 2 * pstep); apply8(v3, px + 3 * pstep);
        }
        for (; px < p1; px += pstep) apply8(__ldg(reinterpret_cast<const uint4 *>(ib + (size_t)px * p.in_stride)), px);
    }
}

// small tensors (npix * C * 2 bytes <= ~200 KB per batch item, channels-per-group a multiple of 8): ONE kernel, one CTA per
// batch item: the slab is staged in shared memory, statistics are reduced in a fixed order, and the normalised values are
// written from shared memory -- one global read, one global write, one launch instead of two latency-bound ones.
#define GN_SMALL_THREADS 1024
__global__ void __launch_bounds__(GN_SMALL_THREADS) k_gn_small(const NormParams p) {
    pdl_launch();
    pdl_wait();
    extern __shared__ __align__(16) unsigned char gsm[];
    uint4 *slab = reinterpret_cast<uint4 *>(gsm);                                     // [npix][C / 8]
    float2 *tsum = reinterpret_cast<float2 *>(gsm + (size_t)p.npix * p.C * 2);        // [threads]
    float2 *mr = tsum + GN_SMALL_THREADS;                                             // [G] (mean, rstd)
    const int chunks = p.C >> 3, cpg = p.C / p.G, cpgc = cpg >> 3;
    const int cpp = min(chunks, (int)blockDim.x);
    const int lanes = ((int)blockDim.x / cpp) * cpp, pstep = lanes / cpp;
    const int b = blockIdx.x;
    const int chunk0 = threadIdx.x % cpp, prow = threadIdx.x / cpp;
    // chunks > blockDim never happens here (C <= 2560 -> 320 chunks <= 1024)
    float s = 0.f, q = 0.f;
    if ((int)threadIdx.x < lanes) {
        const __nv_bfloat16 *base = p.in + (size_t)b * p.npix * p.in_stride + p.in_coff + chunk0 * 8;
        for (int px = prow; px < p.npix; px += pstep) {
            const uint4 v = __ldg(reinterpret_cast<const uint4 *>(base + (size_t)px * p.in_stride));
            slab[px * chunks + chunk0] = v;
            const uint32_t w[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
            for (int j = 0; j < 4; j++) {
                const __nv_bfloat162 h = *reinterpret_cast<const __nv_bfloat162 *>(&w[j]);
                const float a = __bfloat162float(h.x), c = __bfloat162float(h.y);
                s += a + c; q += a * a + c * c;
            }
        }
    }
    tsum[threadIdx.x] = make_float2(s, q);
    __syncthreads();
    if ((int)threadIdx.x < p.G) {   // group g = chunk columns [g * cpgc, (g + 1) * cpgc) of every pixel row, fixed order
        float S = 0.f, Q = 0.f;
        for (int r = 0; r < pstep; r++)
            for (int c = 0; c < cpgc; c++) {
                const float2 t = tsum[r * cpp + threadIdx.x * cpgc + c];
                S += t.x; Q += t.y;
            }
        const float inv_n = 1.0f / ((float)p.npix * (float)cpg);
        const float mean = S * inv_n;
        mr[threadIdx.x] = make_float2(mean, rsqrtf(fmaxf(Q * inv_n - mean * mean, 0.f) + p.eps));
    }
    __syncthreads();
    if ((int)threadIdx.x >= lanes) return;
    const float2 m = mr[chunk0 / cpgc];
    float ca[8], cb[8];
#pragma unroll
    for (int j = 0; j < 8; j++) {
        ca[j] = m.y * __ldg(p.gamma + chunk0 * 8 + j);
        cb[j] = __ldg(p.beta + chunk0 * 8 + j) - m.x * ca[j];
    }
    __nv_bfloat16 *ob = p.out + (size_t)b * p.npix * p.C + chunk0 * 8;
    for (int px = prow; px < p.npix; px += pstep) {
        const uint4 v = slab[px * chunks + chunk0];
        const uint32_t w[4] = {v.x, v.y, v.z, v.w};
        uint32_t o[4];
#pragma unroll
        for (int j = 0; j < 4; j++) {
            const __nv_bfloat162 h = *reinterpret_cast<const __nv_bfloat162 *>(&w[j]);
            float y0 = fmaf(__bfloat162float(h.x), ca[2 * j], cb[2 * j]);
            float y1 = fmaf(__bfloat162float(h.y), ca[2 * j + 1], cb[2 * j + 1]);
            if (p.silu) { y0 = silu_fast(y0); y1 = silu_fast(y1); }
            __nv_bfloat162 r = __floats2bfloat162_rn(y0, y1);
            o[j] = *reinterpret_cast<uint32_t *>(&r);
        }
        *reinterpret_cast<uint4 *>(ob + (size_t)px * p.C) = make_uint4(o[0], o[1], o[2], o[3]);
    }
}

// LayerNorm over C per token: one warp per token (npix = total tokens over the batch), C <= 2048, C % 8 == 0
__global__ void __launch_bounds__(256) k_layernorm(const NormParams p) {
    pdl_launch();
    pdl_wait();
    const int tok = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5), lane = threadIdx.x & 31;
    if (tok >= p.npix) return;
    const __nv_bfloat16 *in = p.in + (size_t)tok * p.C;
    float x[64];
    int n = 0;
    float s = 0.f;
    for (int c = lane * 8; c < p.C; c += 256, n += 8) {
        const uint4 v = __ldg(reinterpret_cast<const uint4 *>(in + c));
        const uint32_t w[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
        for (int j = 0; j < 4; j++) {
            const __nv_bfloat162 h = *reinterpret_cast<const __nv_bfloat162 *>(&w[j]);
            x[n + 2 * j] = __bfloat162float(h.x);
            x[n + 2 * j + 1] = __bfloat162float(h.y);
            s += x[n + 2 * j] + x[n + 2 * j + 1];
        }
    }
#pragma unroll
    for (int o = 16; o; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
    const float mean = s / (float)p.C;
    float q = 0.f;
    for (int j = 0; j < n; j++) { const float d = x[j] - mean; q += d * d; }
#pragma unroll
    for (int o = 16; o; o >>= 1) q += __shfl_xor_sync(0xffffffffu, q, o);
    const float rstd = rsqrtf(q / (float)p.C + p.eps);
    __nv_bfloat16 *out = p.out + (size_t)tok * p.C;
    int k = 0;
    for (int c = lane * 8; c < p.C; c += 256, k += 8) {
        uint32_t o[4];
#pragma unroll
        for (int j = 0; j < 4; j++) {
            float a = (x[k + 2 * j] - mean) * rstd * __ldg(p.gamma + c + 2 * j) + __ldg(p.beta + c + 2 * j);
            float b = (x[k + 2 * j + 1] - mean) * rstd * __ldg(p.gamma + c + 2 * j + 1) + __ldg(p.beta + c + 2 * j + 1);
            if (p.silu == 2) {   // LayerNorm -> exact GELU (wav2vec2 feature encoder)
                a = 0.5f * a * (1.0f + erff(a * 0.70710678118654752f));
                b = 0.5f * b * (1.0f + erff(b * 0.70710678118654752f));
            }
            __nv_bfloat162 h = __floats2bfloat162_rn(a, b);
            o[j] = *reinterpret_cast<uint32_t *>(&h);
        }
        *reinterpret_cast<uint4 *>(out + c) = make_uint4(o[0], o[1], o[2], o[3]);
    }
}

// GEGLU (diffusers GEGLU): in [tokens][2*Hd] -> out [tokens][Hd] = in[:, :Hd] * gelu(in[:, Hd:]) (exact erf gelu)
struct GegluParams {
    const __nv_bfloat16 *in;
    __nv_bfloat16 *out;
    size_t tokens;
    int Hd;
};
__global__ void __launch_bounds__(256) k_geglu(const GegluParams p) {
    pdl_launch();
    pdl_wait();
    const size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x;
    const int chunks = p.Hd >> 3;
    if (i >= p.tokens * chunks) return;
    const size_t tok = i / chunks;
    const int c = (int)(i % chunks) * 8;
    const uint4 a = __ldg(reinterpret_cast<const uint4 *>(p.in + tok * 2 * p.Hd + c));
    const uint4 g = __ldg(reinterpret_cast<const uint4 *>(p.in + tok * 2 * p.Hd + p.Hd + c));
    const uint32_t aw[4] = {a.x, a.y, a.z, a.w}, gw[4] = {g.x, g.y, g.z, g.w};
    uint32_t o[4];
#pragma unroll
    for (int j = 0; j < 4; j++) {
        const __nv_bfloat162 ha = *reinterpret_cast<const __nv_bfloat162 *>(&aw[j]);
        const __nv_bfloat162 hg = *reinterpret_cast<const __nv_bfloat162 *>(&gw[j]);
        const float g0 = __bfloat162float(hg.x), g1 = __bfloat162float(hg.y);
        const float r0 = __bfloat162float(ha.x) * 0.5f * g0 * (1.0f + erff(g0 * 0.70710678118654752f));
        const float r1 = __bfloat162float(ha.y) * 0.5f * g1 * (1.0f + erff(g1 * 0.70710678118654752f));
        __nv_bfloat162 h = __floats2bfloat162_rn(r0, r1);
        o[j] = *reinterpret_cast<uint32_t *>(&h);
    }
    *reinterpret_cast<uint4 *>(p.out + tok * p.Hd + c) = make_uint4(o[0], o[1], o[2], o[3]);
}

// ---------------------------------------------------------------------------------------------------
// batched GEMM on warp-level tensor-core tiles (mma.sync m16n8k16 bf16): the attention products.
//   TRANSB = false : C[M,N] = A[M,K] * B[N,K]^T   (scores = Q K^T; fp32 out)
//   TRANSB = true  : C[M,N] = A[M,K] * B[K,N]     (out = P V; bf16 out)
// batch index z = b * heads + h; pointers advance by (b * *_bs + h * *_hs) elements.
// The attention GEMMs are ~1 % of the MuseTalk FLOPs with per-(batch, head) operands that change every
// call, i.e. no TMA-descriptor-friendly static operand: warp-level MMA is the right size for them.
// ---------------------------------------------------------------------------------------------------
struct GemmParams {
    const __nv_bfloat16 *A, *B;
    void *C;
    int M, N, K, heads;
    int lda, ldb, ldc;
    long long a_bs, a_hs, b_bs, b_hs, c_bs, c_hs;
};

__device__ __forceinline__ void mma_bf16(float (&c)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1) {
    asm volatile(
        "mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};\n"
        : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3])
        : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}
__device__ __forceinline__ void ldmatrix_x2_trans(uint32_t &r0, uint32_t &r1, uint32_t saddr) {
    asm volatile("ldmatrix.sync.aligned.m8n8.x2.trans.shared.b16 {%0,%1}, [%2];\n" : "=r"(r0), "=r"(r1) : "r"(saddr));
}

// CTA tile 64 x 64, K step 32, 4 warps (2 x 2), each warp 32 x 32
template <bool TRANSB, bool OUT_BF16>
__global__ void __launch_bounds__(128) k_bgemm(const GemmParams p) {
    pdl_launch();
    pdl_wait();
    __shared__ __align__(16) __nv_bfloat16 sA[64][40];
    __shared__ __align__(16) __nv_bfloat16 sB[TRANSB ? 32 : 64][TRANSB ? 72 : 40];
    const int z = blockIdx.z, bb = z / p.heads, hh = z % p.heads;
    const __nv_bfloat16 *A = p.A + bb * p.a_bs + hh * p.a_hs;
    const __nv_bfloat16 *B = p.B + bb * p.b_bs + hh * p.b_hs;
    const int m0 = blockIdx.y * 64, n0 = blockIdx.x * 64;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int wm = (warp >> 1) * 32, wn = (warp & 1) * 32;
    float acc[2][4][4];
#pragma unroll
    for (int i = 0; i < 2; i++)
#pragma unroll
        for (int j = 0; j < 4; j++) acc[i][j][0] = acc[i][j][1] = acc[i][j][2] = acc[i][j][3] = 0.f;

    for (int k0 = 0; k0 < p.K; k0 += 32) {
        // A tile 64 x 32: 256 chunks of 8 elements, 2 per thread
        for (int c = threadIdx.x; c < 256; c += 128) {
            const int r = c >> 2, kc = (c & 3) * 8;
            uint4 v = make_uint4(0u, 0u, 0u, 0u);
            if (m0 + r < p.M && k0 + kc < p.K) v = __ldg(reinterpret_cast<const uint4 *>(A + (size_t)(m0 + r) * p.lda + k0 + kc));
            *reinterpret_cast<uint4 *>(&sA[r][kc]) = v;
        }
        if (!TRANSB) {  // B tile 64 (n) x 32 (k)
            for (int c = threadIdx.x; c < 256; c += 128) {
                const int r = c >> 2, kc = (c & 3) * 8;
                uint4 v = make_uint4(0u, 0u, 0u, 0u);
                if (n0 + r < p.N && k0 + kc < p.K) v = __ldg(reinterpret_cast<const uint4 *>(B + (size_t)(n0 + r) * p.ldb + k0 + kc));
                *reinterpret_cast<uint4 *>(&sB[r][kc]) = v;
            }
        } else {  // B tile 32 (k) x 64 (n)
            for (int c = threadIdx.x; c < 256; c += 128) {
                const int r = c >> 3, nc = (c & 7) * 8;
                uint4 v = make_uint4(0u, 0u, 0u, 0u);
                if (k0 + r < p.K && n0 + nc < p.N) v = __ldg(reinterpret_cast<const uint4 *>(B + (size_t)(k0 + r) * p.ldb + n0 + nc));
                *reinterpret_cast<uint4 *>(&sB[r][nc]) = v;
            }
        }
        __syncthreads();
#pragma unroll
        for (int ks = 0; ks < 2; ks++) {
            uint32_t a[2][4];
#pragma unroll
            for (int mt = 0; mt < 2; mt++) {
                const int q = lane >> 3;
                const int row = wm + mt * 16 + (lane & 7) + 8 * (q & 1), col = ks * 16 + 8 * (q >> 1);
                ldmatrix_x4(a[mt], smem_u32(&sA[row][col]));
            }
#pragma unroll
            for (int nt = 0; nt < 4; nt++) {
                uint32_t b0, b1;
                if (!TRANSB) {
                    const int l = lane & 15;
                    ldmatrix_x2(b0, b1, smem_u32(&sB[wn + nt * 8 + (l & 7)][ks * 16 + 8 * (l >> 3)]));
                } else {
                    const int l = lane & 15;
                    ldmatrix_x2_trans(b0, b1, smem_u32(&sB[ks * 16 + l][wn + nt * 8]));
                }
                mma_bf16(acc[0][nt], a[0], b0, b1);
                mma_bf16(acc[1][nt], a[1], b0, b1);
            }
        }
        __syncthreads();
    }
    const int g = lane >> 2, t = lane & 3;
#pragma unroll
    for (int mt = 0; mt < 2; mt++)
#pragma unroll
        for (int nt = 0; nt < 4; nt++)
#pragma unroll
            for (int h = 0; h < 2; h++) {
                const int row = m0 + wm + mt * 16 + g + 8 * h, col = n0 + wn + nt * 8 + 2 * t;
                if (row >= p.M || col >= p.N) continue;
                const size_t off = (size_t)(bb * p.c_bs + hh * p.c_hs) + (size_t)row * p.ldc + col;
                const float v0 = acc[mt][nt][2 * h], v1 = acc[mt][nt][2 * h + 1];
                if (OUT_BF16) {
                    __nv_bfloat16 *C = reinterpret_cast<__nv_bfloat16 *>(p.C) + off;
                    C[0] = __float2bfloat16_rn(v0);
                    if (col + 1 < p.N) C[1] = __float2bfloat16_rn(v1);
                } else {
                    float *C = reinterpret_cast<float *>(p.C) + off;
                    C[0] = v0;
                    if (col + 1 < p.N) C[1] = v1;
                }
            }
}

// ---------------------------------------------------------------------------------------------------
// fused attention (flash-attention style, warp-level mma.sync bf16): O = softmax(Q K^T * scale) V per (batch, head) without
// materialising the score matrix (the unfused path writes nq x nk fp32 scores + bf16 probabilities to HBM: 0.8 GB per
// 32x32 self-attention at batch 16).  One CTA = 64 query rows (4 warps x 16), K / V streamed in 64-key tiles through
// shared memory; online softmax in fp32; P is rounded to bf16 before P V exactly like the unfused path.
// DP = head dim rounded up to 16 (zero padded in shared memory): 48 (dh 40), 64, 80, 160.
// ---------------------------------------------------------------------------------------------------
struct FlashParams {
    const __nv_bfloat16 *Q, *K, *V;
    __nv_bfloat16 *O;
    int nq, nk, dh, heads;
    int ldq, ldk, ldv, ldo;             // token strides (elements)
    long long q_bs, k_bs, v_bs, o_bs;   // batch strides (elements); head h starts at column h * dh
    float scale_log2;                   // scale * log2(e)
};

template <int DP>
__global__ void __launch_bounds__(128) k_flash(const FlashParams p) {
    pdl_launch();
    pdl_wait();
    constexpr int LD = DP + 8;          // +16 B per row: conflict-free ldmatrix
    constexpr int KS = DP / 16;         // k-steps of Q K^T
    constexpr int NT = DP / 8;          // n-tiles of P V
    __shared__ __align__(16) __nv_bfloat16 sK[64][LD];
    __shared__ __align__(16) __nv_bfloat16 sV[64][LD];
    const int b = blockIdx.z, h = blockIdx.y, q0 = blockIdx.x * 64;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int g = lane >> 2, t = lane & 3;
    const __nv_bfloat16 *Q = p.Q + b * p.q_bs + h * p.dh;
    const __nv_bfloat16 *K = p.K + b * p.k_bs + h * p.dh;
    const __nv_bfloat16 *V = p.V + b * p.v_bs + h * p.dh;
    const int dchunks = DP / 8;
    const int dvalid = p.dh / 8;        // dh % 8 == 0

    // ---- Q fragments (A operand, 16 rows x DP) via shared memory (reuse sK as staging)
    for (int c = threadIdx.x; c < 64 * dchunks; c += 128) {
        const int r = c / dchunks, dc = c % dchunks;
        uint4 v = make_uint4(0u, 0u, 0u, 0u);
        if (q0 + r < p.nq && dc < dvalid) v = __ldg(reinterpret_cast<const uint4 *>(Q + (size_t)(q0 + r) * p.ldq + dc * 8));
        *reinterpret_cast<uint4 *>(&sK[r][dc * 8]) = v;
    }
    __syncthreads();
    uint32_t qf[KS][4];
#pragma unroll
    for (int ks = 0; ks < KS; ks++) {
        const int qd = lane >> 3;
        ldmatrix_x4(qf[ks], smem_u32(&sK[warp * 16 + (lane & 7) + 8 * (qd & 1)][ks * 16 + 8 * (qd >> 1)]));
    }
    __syncthreads();

    float o[NT][4];
#pragma unroll
    for (int i = 0; i < NT; i++) o[i][0] = o[i][1] = o[i][2] = o[i][3] = 0.f;
    float m0 = -1e30f, m1 = -1e30f, l0 = 0.f, l1 = 0.f;   // rows g and g + 8 of this warp's 16

    for (int k0 = 0; k0 < p.nk; k0 += 64) {
        for (int c = threadIdx.x; c < 64 * dchunks; c += 128) {
            const int r = c / dchunks, dc = c % dchunks;
            uint4 kv = make_uint4(0u, 0u, 0u, 0u), vv = kv;
            if (k0 + r < p.nk && dc < dvalid) {
                kv = __ldg(reinterpret_cast<const uint4 *>(K + (size_t)(k0 + r) * p.ldk + dc * 8));
                vv = __ldg(reinterpret_cast<const uint4 *>(V + (size_t)(k0 + r) * p.ldv + dc * 8));
            }
            *reinterpret_cast<uint4 *>(&sK[r][dc * 8]) = kv;
            *reinterpret_cast<uint4 *>(&sV[r][dc * 8]) = vv;
        }
        __syncthreads();
        // ---- S = Q K^T : 16 x 64 per warp
        float sc[8][4];
#pragma unroll
        for (int nt = 0; nt < 8; nt++) sc[nt][0] = sc[nt][1] = sc[nt][2] = sc[nt][3] = 0.f;
#pragma unroll
        for (int ks = 0; ks < KS; ks++)
#pragma unroll
            for (int nt = 0; nt < 8; nt++) {
                uint32_t b0, b1;
                const int l = lane & 15;
                ldmatrix_x2(b0, b1, smem_u32(&sK[nt * 8 + (l & 7)][ks * 16 + 8 * (l >> 3)]));
                mma_bf16(sc[nt], qf[ks], b0, b1);
            }
        // ---- online softmax (exp2 domain); keys beyond nk are masked
        float mx0 = m0, mx1 = m1;
#pragma unroll
        for (int nt = 0; nt < 8; nt++) {
            const int key = k0 + nt * 8 + 2 * t;
#pragma unroll
            for (int e = 0; e < 4; e++) {
                const bool ok = key + (e & 1) < p.nk;
                sc[nt][e] = ok ? sc[nt][e] * p.scale_log2 : -1e30f;
            }
            mx0 = fmaxf(mx0, fmaxf(sc[nt][0], sc[nt][1]));
            mx1 = fmaxf(mx1, fmaxf(sc[nt][2], sc[nt][3]));
        }
        mx0 = fmaxf(mx0, __shfl_xor_sync(0xffffffffu, mx0, 1)); mx0 = fmaxf(mx0, __shfl_xor_sync(0xffffffffu, mx0, 2));
        mx1 = fmaxf(mx1, __shfl_xor_sync(0xffffffffu, mx1, 1)); mx1 = fmaxf(mx1, __shfl_xor_sync(0xffffffffu, mx1, 2));
        const float c0 = exp2f(m0 - mx0), c1 = exp2f(m1 - mx1);
        m0 = mx0; m1 = mx1;
        l0 *= c0; l1 *= c1;
#pragma unroll
        for (int i = 0; i < NT; i++) { o[i][0] *= c0; o[i][1] *= c0; o[i][2] *= c1; o[i][3] *= c1; }
        uint32_t pf[4][4];   // P as A fragments: k-step j = keys [16 j, 16 j + 16)
#pragma unroll
        for (int nt = 0; nt < 8; nt++) {
            const float e0 = exp2f(sc[nt][0] - m0), e1 = exp2f(sc[nt][1] - m0), e2 = exp2f(sc[nt][2] - m1), e3 = exp2f(sc[nt][3] - m1);
            __nv_bfloat162 h01 = __floats2bfloat162_rn(e0, e1), h23 = __floats2bfloat162_rn(e2, e3);
            // the row sums use the ROUNDED probabilities, so that sum(P) and P V see the same numbers
            l0 += __bfloat162float(h01.x) + __bfloat162float(h01.y);
            l1 += __bfloat162float(h23.x) + __bfloat162float(h23.y);
            pf[nt >> 1][(nt & 1) * 2] = *reinterpret_cast<uint32_t *>(&h01);
            pf[nt >> 1][(nt & 1) * 2 + 1] = *reinterpret_cast<uint32_t *>(&h23);
        }
        // ---- O += P V : V^T fragments via ldmatrix.trans
#pragma unroll
        for (int j = 0; j < 4; j++)
#pragma unroll
            for (int nt = 0; nt < NT; nt++) {
                uint32_t b0, b1;
                ldmatrix_x2_trans(b0, b1, smem_u32(&sV[j * 16 + (lane & 15)][nt * 8]));
                mma_bf16(o[nt], pf[j], b0, b1);
            }
        __syncthreads();
    }
    l0 += __shfl_xor_sync(0xffffffffu, l0, 1); l0 += __shfl_xor_sync(0xffffffffu, l0, 2);
    l1 += __shfl_xor_sync(0xffffffffu, l1, 1); l1 += __shfl_xor_sync(0xffffffffu, l1, 2);
    const float i0 = 1.0f / l0, i1 = 1.0f / l1;
    __nv_bfloat16 *O = p.O + b * p.o_bs + h * p.dh;
    const int r0 = q0 + warp * 16 + g, r1 = r0 + 8;
#pragma unroll
    for (int nt = 0; nt < NT; nt++) {
        const int d = nt * 8 + 2 * t;
        if (d >= p.dh) continue;
        if (r0 < p.nq) *reinterpret_cast<__nv_bfloat162 *>(O + (size_t)r0 * p.ldo + d) = __floats2bfloat162_rn(o[nt][0] * i0, o[nt][1] * i0);
        if (r1 < p.nq) *reinterpret_cast<__nv_bfloat162 *>(O + (size_t)r1 * p.ldo + d) = __floats2bfloat162_rn(o[nt][2] * i1, o[nt][3] * i1);
    }
}

// row softmax: S fp32 [rows][ld] (valid cols n) * scale -> P bf16 [rows][ld], padding columns zeroed.  One warp per row.
struct SoftmaxParams {
    const float *S;
    __nv_bfloat16 *P;
    size_t rows;
    int n, ld;
    float scale;
};
__global__ void __launch_bounds__(256) k_softmax(const SoftmaxParams p) {
    pdl_launch();
    pdl_wait();
    const size_t row = blockIdx.x * (size_t)(blockDim.x >> 5) + (threadIdx.x >> 5);
    const int lane = threadIdx.x & 31;
    if (row >= p.rows) return;
    const float *s = p.S + row * p.ld;
    float mx = -3.0e38f;
    for (int c = lane; c < p.n; c += 32) mx = fmaxf(mx, s[c] * p.scale);
#pragma unroll
    for (int o = 16; o; o >>= 1) mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, o));
    float sum = 0.f;
    for (int c = lane; c < p.n; c += 32) sum += __expf(s[c] * p.scale - mx);
#pragma unroll
    for (int o = 16; o; o >>= 1) sum += __shfl_xor_sync(0xffffffffu, sum, o);
    const float inv = 1.0f / sum;
    __nv_bfloat16 *out = p.P + row * p.ld;
    for (int c = lane; c < p.ld; c += 32) out[c] = __float2bfloat16_rn(c < p.n ? __expf(s[c] * p.scale - mx) * inv : 0.f);
}
