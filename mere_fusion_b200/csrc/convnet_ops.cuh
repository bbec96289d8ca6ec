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
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= p.a) return;
    __nv_bfloat162 h = __floats2bfloat162_rn(reinterpret_cast<const float *>(p.src)[i], 0.f);
    reinterpret_cast<uint4 *>(p.dst)[i] = make_uint4(*reinterpret_cast<uint32_t *>(&h), 0u, 0u, 0u);
}
// musetalk latents fp16 NCHW [B,C,H,W] -> bf16 NHWC [B,H,W,Cpad].  a = B, b = C, c = H*W, d = Cpad
__global__ void k_prep_latents(const PrepParams p) {
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
__global__ void k_f32_to_bf16(const PrepParams p) {
    const size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x;
    const size_t n = ((size_t)(uint32_t)p.a) | ((size_t)(uint32_t)p.b << 32);
    if (i < n) reinterpret_cast<__nv_bfloat16 *>(p.dst)[i] = __float2bfloat16_rn(reinterpret_cast<const float *>(p.src)[i]);
}
__global__ void k_bf16_to_f32(const PrepParams p) {
    const size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x;
    const size_t n = ((size_t)(uint32_t)p.a) | ((size_t)(uint32_t)p.b << 32);
    if (i < n) reinterpret_cast<float *>(p.dst)[i] = __bfloat162float(reinterpret_cast<const __nv_bfloat16 *>(p.src)[i]);
}

// ---------------------------------------------------------------------------------------------------
// GroupNorm (+ SiLU): statistics by atomics into a per-op fp32 slot [B][G][2], then an element-wise pass
// ---------------------------------------------------------------------------------------------------
struct NormParams {
    const __nv_bfloat16 *in;
    __nv_bfloat16 *out;
    const float *gamma, *beta;
    float *stats;  // [B][G][2] (sum, sum of squares), zeroed once per forward
    int npix, C, G, silu;
    float eps;
    int pix_per_cta;
    int in_stride, in_coff;  // GroupNorm input may be a channel range of a wider (concat) buffer; output is dense
};

// grid (ceil(npix / pix_per_cta), B).  A thread owns fixed 8-channel chunk columns (so the group of each of its 8
// channels is fixed and sums stay in registers) and walks the CTA's pixel range with a stride.
__global__ void __launch_bounds__(256) k_gn_stats(const NormParams p) {
    __shared__ float acc[64][2];
    const int chunks = p.C >> 3;
    const int cpp = min(chunks, (int)blockDim.x);          // chunk columns handled per pass
    const int lanes = ((int)blockDim.x / cpp) * cpp;
    for (int i = threadIdx.x; i < p.G * 2; i += blockDim.x) (&acc[0][0])[i] = 0.f;
    __syncthreads();
    if ((int)threadIdx.x < lanes) {
        const int prow = threadIdx.x / cpp, pstep = lanes / cpp;
        const int cpg = p.C / p.G;
        const int p0 = blockIdx.x * p.pix_per_cta, p1 = min(p.npix, p0 + p.pix_per_cta);
        for (int chunk = threadIdx.x % cpp; chunk < chunks; chunk += cpp) {
            const __nv_bfloat16 *base = p.in + (size_t)blockIdx.y * p.npix * p.in_stride + p.in_coff + chunk * 8;
            float s[8], q[8];
#pragma unroll
            for (int j = 0; j < 8; j++) s[j] = q[j] = 0.f;
            for (int px = p0 + prow; px < p1; px += pstep) {
                const uint4 v = __ldg(reinterpret_cast<const uint4 *>(base + (size_t)px * p.in_stride));
                const uint32_t w[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
                for (int j = 0; j < 4; j++) {
                    const __nv_bfloat162 h = *reinterpret_cast<const __nv_bfloat162 *>(&w[j]);
                    const float a = __bfloat162float(h.x), b = __bfloat162float(h.y);
                    s[2 * j] += a; q[2 * j] += a * a;
                    s[2 * j + 1] += b; q[2 * j + 1] += b * b;
                }
            }
#pragma unroll
            for (int j = 0; j < 8; j++) {
                const int g = (chunk * 8 + j) / cpg;
                atomicAdd(&acc[g][0], s[j]);
                atomicAdd(&acc[g][1], q[j]);
            }
        }
    }
    __syncthreads();
    for (int i = threadIdx.x; i < p.G * 2; i += blockDim.x)
        atomicAdd(p.stats + (size_t)blockIdx.y * p.G * 2 + i, (&acc[0][0])[i]);
}

// one thread per 8-channel chunk
__global__ void __launch_bounds__(256) k_gn_apply(const NormParams p) {
    const size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x;
    const int chunks = p.C >> 3;
    const size_t total = (size_t)p.npix * chunks;  // per batch item
    if (i >= total) return;
    const int b = blockIdx.y;
    const int chunk = (int)(i % chunks);
    const size_t off = ((size_t)b * p.npix + i / chunks) * p.C + chunk * 8;
    const size_t ioff = ((size_t)b * p.npix + i / chunks) * p.in_stride + p.in_coff + chunk * 8;
    const int cpg = p.C / p.G;
    const float inv_n = 1.0f / ((float)p.npix * (float)cpg);
    const uint4 v = __ldg(reinterpret_cast<const uint4 *>(p.in + ioff));
    const uint32_t w[4] = {v.x, v.y, v.z, v.w};
    float x[8];
#pragma unroll
    for (int j = 0; j < 4; j++) {
        const __nv_bfloat162 h = *reinterpret_cast<const __nv_bfloat162 *>(&w[j]);
        x[2 * j] = __bfloat162float(h.x);
        x[2 * j + 1] = __bfloat162float(h.y);
    }
    uint32_t o[4];
    float y[8];
#pragma unroll
    for (int j = 0; j < 8; j++) {
        const int c = chunk * 8 + j, g = c / cpg;
        const float sum = p.stats[((size_t)b * p.G + g) * 2], sq = p.stats[((size_t)b * p.G + g) * 2 + 1];
        const float mean = sum * inv_n;
        const float var = fmaxf(sq * inv_n - mean * mean, 0.f);
        float t = (x[j] - mean) * rsqrtf(var + p.eps) * __ldg(p.gamma + c) + __ldg(p.beta + c);
        if (p.silu) t = t / (1.0f + __expf(-t));
        y[j] = t;
    }
#pragma unroll
    for (int j = 0; j < 4; j++) {
        __nv_bfloat162 h = __floats2bfloat162_rn(y[2 * j], y[2 * j + 1]);
        o[j] = *reinterpret_cast<uint32_t *>(&h);
    }
    *reinterpret_cast<uint4 *>(p.out + off) = make_uint4(o[0], o[1], o[2], o[3]);
}

// LayerNorm over C per token: one warp per token (npix = total tokens over the batch), C <= 2048, C % 8 == 0
__global__ void __launch_bounds__(256) k_layernorm(const NormParams p) {
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
            const float a = (x[k + 2 * j] - mean) * rstd * __ldg(p.gamma + c + 2 * j) + __ldg(p.beta + c + 2 * j);
            const float b = (x[k + 2 * j + 1] - mean) * rstd * __ldg(p.gamma + c + 2 * j + 1) + __ldg(p.beta + c + 2 * j + 1);
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

// row softmax: S fp32 [rows][ld] (valid cols n) * scale -> P bf16 [rows][ld], padding columns zeroed.  One warp per row.
struct SoftmaxParams {
    const float *S;
    __nv_bfloat16 *P;
    size_t rows;
    int n, ld;
    float scale;
};
__global__ void __launch_bounds__(256) k_softmax(const SoftmaxParams p) {
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
