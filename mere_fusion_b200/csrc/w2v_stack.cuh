// w2v_stack.cuh -- the 24 pre-LN transformer layers of the wav2vec2 CTC acoustic model (NerfASR.__frame_to_text, nerfasr.py:128-143;
// transformers modeling_wav2vec2.py: Wav2Vec2EncoderLayerStableLayerNorm) as ONE persistent kernel.
//
// Round 1 ran the stack as 24 x 7 launches of the conv executor (LayerNorm, QKV GEMM, attention, out-proj, LayerNorm, FFN1, FFN2):
// M = 27 tokens per window, so every GEMM is a weight-streaming GEMV-like problem (25 MB of bf16 weights per layer, 604 MB in
// all = ~0.1 ms at HBM speed) and the 1.6 ms it took were launch latency and under-filled grids (profiles: 203 launches).
// Here: G = 128 CTAs (one per SM, all resident, clusters of 4), each owns a fixed slice of the OUTPUT columns of every matrix, so every
// weight is read from HBM exactly once per forward and results need no cross-CTA reduction (deterministic); a phase is
//   [stream this CTA's rows of W by cp.async.bulk (TMA 1-D, one copy per row) into padded shared-memory rows, double buffered,
//    the NEXT phase's slice issued before the grid barrier]  x  [A operand: LayerNorm of the fp32 residual stream -- each CTA of a
//    cluster normalises a quarter of the rows and stores them into all four operands through distributed shared memory -- or the
//    bf16 activations of the previous phase by cp.async]  ->  mma.sync.m16n8k16 bf16  ->  epilogue
// and phases are separated by a software grid barrier (5 per layer):
//   P1  qkv  = LN1(x) Wqkv^T + b          P2  attention per (window, head, row group) item      P3  x += ao Wo^T + b
//   P4  hid  = gelu(LN2(x) W1^T + b)      P5  x += hid W2^T + b  (K = 4096 split over the cluster's ranks, partial sums through DSMEM)
// The LayerNorm affines are folded into Wqkv / W1 and their biases at pack time.  The residual stream stays fp32 in a scratch buffer
// (the executor's buffers are bf16).  One window per launch (M <= 32 rows): an engine that batches the windows of several sessions keeps
// the op-by-op program.
// Measured (B200, XLSR-53 shape, scripts/time_w2v.py: %globaltimer stamps of CTA 0 in layer 1): 1.16 ms per window against 1.97 ms for
// the 203-launch program of round 1; per layer 38.5 us = LN1+QKV 6.6 (LN 3.5: x loaded + first reduction 1.3, normalise + remote stores
// 1.5, cluster sync 0.6; MMA 1.4; epilogue 0.7 with the bias loaded before the GEMM), attention 4.0 (7.0 before the K rows were padded:
// 32-way bank conflict in the score loop), out-proj 4.9, LN2+FFN1 8.4, FFN2 7.5 (11.4 as a 4-stage ring over the full K, 13.7 with
// 256-column chunks), five barriers ~1.5 each (release / acquire instead of full fences: 2.0 -> 1.5; the rest is arrival skew).
// Every phase is a few dependent L2 round trips at ~32 B/clk of L1 ingest per SM; the weight stream (25 MB per layer, 4 us at HBM speed)
// is nowhere near the limit.  Tried and dropped: the FFN2 ring fed by cp.async.bulk row copies (216 copies of 1 KB per phase: 17.6 us),
// bulk row copies for the whole-A operands of out-proj / FFN2 (no change), clusters of 8 (16 of them are not co-resident on this part).
#pragma once

#define WS_G 128          /* CTAs; output-column slices are multiples of 8 */
#define WS_THREADS_ 256
#define WS_KC 512         /* K chunk (elements) of the A ring used when K > WS_KA (256: 16 sync points per FFN2 phase, 13.7 us) */
#define WS_KS (WS_KC + 8) /* padded chunk row stride in smem (bank-conflict-free ldmatrix) */
#define WS_KA 1024        /* largest K whose A operand is resident as a whole */
#define WS_STAGES 4       /* A ring depth (chunks of WS_KC): 2 stages in sm.A, 2 in the weight buffer that is idle during the phase */
#define WS_RING_PAD 256   /* halfs added to sm.A and each weight buffer so that two ring stages (32 x (WS_KC + 8)) fit */
#define WS_MAX_NC 32      /* output columns per CTA and phase */
#define WS_MAX_MT 2       /* 16-row tiles of tokens: one window (<= 32 frames); batched windows use the op-by-op program */
#define WS_WBUF_HALFS (WS_MAX_NC * (WS_KA + 8) + WS_RING_PAD)   /* one weight slice: nc rows x (K + 8), nc * K <= 32 * 1024 */

struct W2vStackParams {
    const unsigned char *image;   // packed weights (layout below)
    const __nv_bfloat16 *x_in;    // [M][D] bf16 (after the positional conv)
    __nv_bfloat16 *x_out;         // [M][D] bf16
    float *xres;                  // scratch [M][D] fp32 residual stream
    __nv_bfloat16 *qkv, *ao, *hid;  // scratch [M][3D], [M][D], [M][I]
    unsigned *barrier;            // zeroed before the launch; 16 counters, then 24 x u64 phase time stamps of CTA 0 in layer 1 (debug tap)
    int M, T, B, D, I, heads, layers;
    float eps, scale_log2;        // softmax scale * log2(e)
};

// image layout per layer (bytes): vectors fp32 [ln1_g D | ln1_b D | bqkv 3D | bo D | ln2_g D | ln2_b D | b1 I | b2 D] (the LayerNorm
// affines are folded into Wqkv / bqkv and W1 / b1 by the packer; their slots hold 1 / 0 and are not read), then matrices bf16
// row-major [Wqkv 3D x D | Wo D x D | W1 I x D | W2 D x I]
struct W2vLayerPtrs {
    const float *ln1_g, *ln1_b, *bqkv, *bo, *ln2_g, *ln2_b, *b1, *b2;
    const __nv_bfloat16 *Wqkv, *Wo, *W1, *W2;
};
__host__ __device__ inline size_t w2v_layer_bytes(int D, int I) { return (size_t)(9 * D + I) * 4 + ((size_t)4 * D * D + (size_t)2 * D * I) * 2; }
__device__ __forceinline__ W2vLayerPtrs w2v_layer(const unsigned char *image, int l, int D, int I) {
    const unsigned char *p = image + (size_t)l * w2v_layer_bytes(D, I);
    W2vLayerPtrs w;
    const float *f = reinterpret_cast<const float *>(p);
    w.ln1_g = f; f += D; w.ln1_b = f; f += D; w.bqkv = f; f += 3 * D; w.bo = f; f += D;
    w.ln2_g = f; f += D; w.ln2_b = f; f += D; w.b1 = f; f += I; w.b2 = f; f += D;
    const __nv_bfloat16 *h = reinterpret_cast<const __nv_bfloat16 *>(f);
    w.Wqkv = h; h += (size_t)3 * D * D; w.Wo = h; h += (size_t)D * D; w.W1 = h; h += (size_t)I * D; w.W2 = h;
    return w;
}

struct W2vSmem {
    alignas(16) __nv_bfloat16 W[2][WS_WBUF_HALFS];                // this phase's weight slice (whole K) / the next phase's, in flight
    alignas(16) __nv_bfloat16 A[WS_MAX_MT * 16 * (WS_KA + 8) + WS_RING_PAD];    // A operand: whole (K <= WS_KA) or a WS_STAGES ring of WS_KC chunks
    float red[8][32][8];                                          // split-K partials: [warp][lane][c]
    float mean[WS_MAX_MT * 16], rstd[WS_MAX_MT * 16];
    alignas(8) uint64_t wbar[2];
};

// Software grid barrier.  Arrivals are spread over 16 counters (CTA b -> counter b & 15), warp 0 polls: lane i watches counter i.  The
// arrival is a red.release.gpu, the poll an ld.acquire.gpu: with __threadfence() on both sides a barrier cost ~2.0 us, so ~1.5 (the
// full fences drained every outstanding access of the thread; the multi-counter split alone changed nothing: the rest is arrival skew).
#define WS_BAR_CTRS 16
__device__ __forceinline__ void w2v_grid_barrier(unsigned *ctr, unsigned &gen) {
    __syncthreads();
    gen++;
    if (threadIdx.x < 32) {
        const int lane = threadIdx.x;
        if (lane == 0)   // release: the CTA's writes (ordered before this by the __syncthreads above) are visible to whoever acquires the counter
            asm volatile("red.release.gpu.global.add.u32 [%0], 1;\n" ::"l"(ctr + (blockIdx.x & (WS_BAR_CTRS - 1))) : "memory");
        // CTAs that share counter i: b = i, i + 16, ... < gridDim.x
        const unsigned per = lane < WS_BAR_CTRS ? (gridDim.x - lane + WS_BAR_CTRS - 1) / WS_BAR_CTRS : 0u;
        const unsigned target = gen * per;
        bool ok;
        do {
            unsigned v = target;
            if (lane < WS_BAR_CTRS && per != 0u) asm volatile("ld.acquire.gpu.global.u32 %0, [%1];\n" : "=r"(v) : "l"(ctr + lane) : "memory");
            ok = v >= target;
        } while (!__all_sync(0xffffffffu, ok));
    }
    __syncthreads();
}

__device__ __forceinline__ void w2v_stamp(const W2vStackParams &p, int l, int i) {   // %globaltimer at a phase boundary (mf_debug_w2v_phase_ns)
    if (l == 1 && blockIdx.x == 0 && threadIdx.x == 0) {
        unsigned long long t;
        asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t));
        reinterpret_cast<unsigned long long *>(p.barrier + 16)[i] = t;
    }
}
__device__ __forceinline__ void mma_bf16_16816(float (&c)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1) {
    asm volatile(
        "mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};\n"
        : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3])
        : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}

// issue the bulk copies (TMA 1-D, one per row) of this CTA's weight rows [n0, n0 + nc) of W [N][K] into buffer `slot` (one thread)
__device__ __forceinline__ void w2v_issue_w(W2vSmem &sm, const __nv_bfloat16 *W, int K, int n0, int nc, int slot) {
    mbar_expect_tx(&sm.wbar[slot], (uint32_t)(nc * K * 2));
    for (int r = 0; r < nc; r++) bulk_g2s(sm.W[slot] + (size_t)r * (K + 8), W + (size_t)(n0 + r) * K, (uint32_t)(K * 2), &sm.wbar[slot]);
}
// ... columns [k0, k0 + Ks) of those rows only (row stride ldK in HBM, Ks + 8 in shared memory): the K-split form of FFN2
__device__ __forceinline__ void w2v_issue_w_sub(W2vSmem &sm, const __nv_bfloat16 *W, int ldK, int k0, int Ks, int n0, int nc, int slot) {
    mbar_expect_tx(&sm.wbar[slot], (uint32_t)(nc * Ks * 2));
    for (int r = 0; r < nc; r++) bulk_g2s(sm.W[slot] + (size_t)r * (Ks + 8), W + (size_t)(n0 + r) * ldK + k0, (uint32_t)(Ks * 2), &sm.wbar[slot]);
}
__device__ __forceinline__ float w2v_ld_cluster_f32(uint32_t local_saddr, uint32_t rank) {
    uint32_t remote;
    float v;
    asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(remote) : "r"(local_saddr), "r"(rank));
    asm volatile("ld.shared::cluster.f32 %0, [%1];" : "=f"(v) : "r"(remote) : "memory");
    return v;
}
__device__ __forceinline__ void w2v_cp16(void *dst, const void *src) {
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;\n" ::"r"(smem_u32(dst)), "l"(src) : "memory");
}
__device__ __forceinline__ void w2v_cp_commit() { asm volatile("cp.async.commit_group;\n" ::: "memory"); }
template <int N>
__device__ __forceinline__ void w2v_cp_wait() { asm volatile("cp.async.wait_group %0;\n" ::"n"(N) : "memory"); }

// A operand builders: rows 0..Mpad-1 of sm.A with row stride `ks` halfs ------------------------------------------------------
// LayerNorm of the fp32 residual rows -> bf16 A operand (row stride D + 8).  Every CTA needs ALL rows; computing them in every CTA was
// 6.0 us of the 9.3 us LN1+QKV phase (128-fold redundant: 110 KB of fp32 rows through each SM's L1, ~270 instructions per row and lane).
// The CTAs of a thread-block cluster share the work instead: CTA rank r normalises rows r, r + csize, ... (one warp per row, the row in
// registers, two-pass statistics like torch.nn.functional.layer_norm) and stores the bf16 row into the shared memory of EVERY CTA of the
// cluster (st.shared::cluster through mapa), then the cluster synchronises.  csize = 1 degenerates to every CTA doing every row.
// Safe against the neighbours' use of sm.A: every CTA of the grid has passed the grid barrier that precedes an LN phase, i.e. has finished
// the previous phase's reads of its sm.A, before any CTA gets here; nobody writes sm.A remotely outside an LN phase.  D <= 1024, M <= 32.
__device__ __forceinline__ uint32_t w2v_cluster_rank() { uint32_t r; asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r)); return r; }
__device__ __forceinline__ uint32_t w2v_cluster_size() { uint32_t r; asm volatile("mov.u32 %0, %%cluster_nctarank;" : "=r"(r)); return r; }
__device__ __forceinline__ void w2v_cluster_sync() {
    asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;\n" ::: "memory");
}
__device__ __forceinline__ void w2v_st_cluster_v4(uint32_t local_saddr, uint32_t rank, const uint32_t (&w)[4]) {
    uint32_t remote;
    asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(remote) : "r"(local_saddr), "r"(rank));
    asm volatile("st.shared::cluster.v4.u32 [%0], {%1, %2, %3, %4};" ::"r"(remote), "r"(w[0]), "r"(w[1]), "r"(w[2]), "r"(w[3]) : "memory");
}
// The affine part (gamma, beta) is folded into the following linear layer at pack time (wav2vec2_pack.py): the phase starts with the
// loads of x instead of a round trip to HBM for two vectors.
__device__ __forceinline__ void w2v_ln_rows(W2vSmem &sm, const float *xres, int M, int Mpad, int D, float eps, const W2vStackParams &dbg_p, int dbg_l = -1) {
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, ks = D + 8;
    constexpr int NV = WS_KA / 128;                                              // float4 per lane and row
    const int crank = (int)w2v_cluster_rank(), csize = (int)w2v_cluster_size();
    // the zero rows M .. Mpad-1 of this CTA's own operand
    for (int i = threadIdx.x; i < (Mpad - M) * (D / 4); i += WS_THREADS_) {
        const int m = M + i / (D / 4), c = (i - (i / (D / 4)) * (D / 4)) * 4;
        *reinterpret_cast<uint2 *>(sm.A + (size_t)m * ks + c) = make_uint2(0u, 0u);
    }
    // a lane owns 8 consecutive columns per 256-column group: two float4 loads, one 16-byte remote store per destination CTA
    for (int m = crank + csize * warp; m < M; m += csize * (WS_THREADS_ / 32)) {
        float4 v[NV];
#pragma unroll
        for (int i = 0; i < NV; i++) {
            const int c = lane * 8 + (i >> 1) * 256 + (i & 1) * 4;
            v[i] = c < D ? __ldcg(reinterpret_cast<const float4 *>(xres + (size_t)m * D + c)) : make_float4(0.f, 0.f, 0.f, 0.f);   // written by other CTAs: not through L1
        }
        float s = 0.f;
#pragma unroll
        for (int i = 0; i < NV; i++) s += (v[i].x + v[i].y) + (v[i].z + v[i].w);
#pragma unroll
        for (int o = 16; o; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
        if (dbg_l >= 0) w2v_stamp(dbg_p, dbg_l, 16);
        const float mean = s / (float)D;
        float q = 0.f;
#pragma unroll
        for (int i = 0; i < NV; i++)
            if (lane * 8 + (i >> 1) * 256 + (i & 1) * 4 < D) {
                const float a0 = v[i].x - mean, a1 = v[i].y - mean, a2 = v[i].z - mean, a3 = v[i].w - mean;
                q += (a0 * a0 + a1 * a1) + (a2 * a2 + a3 * a3);
            }
#pragma unroll
        for (int o = 16; o; o >>= 1) q += __shfl_xor_sync(0xffffffffu, q, o);
        const float rs = rsqrtf(q / (float)D + eps);
        const uint32_t row_saddr = smem_u32(sm.A + (size_t)m * ks);
#pragma unroll
        for (int i = 0; i < NV; i += 2) {
            const int c = lane * 8 + (i >> 1) * 256;
            if (c < D) {
                uint32_t w[4];
#pragma unroll
                for (int h = 0; h < 2; h++) {
                    __nv_bfloat162 h0 = __floats2bfloat162_rn((v[i + h].x - mean) * rs, (v[i + h].y - mean) * rs);
                    __nv_bfloat162 h1 = __floats2bfloat162_rn((v[i + h].z - mean) * rs, (v[i + h].w - mean) * rs);
                    w[2 * h] = *reinterpret_cast<uint32_t *>(&h0); w[2 * h + 1] = *reinterpret_cast<uint32_t *>(&h1);
                }
                for (int d = 0; d < csize; d++) w2v_st_cluster_v4(row_saddr + (uint32_t)c * 2u, (uint32_t)d, w);
            }
        }
    }
    if (dbg_l >= 0) w2v_stamp(dbg_p, dbg_l, 17);
    w2v_cluster_sync();                                    // every CTA's rows have landed in every CTA's operand
}
// columns [k0, k0 + kn) of bf16 activations [M][ld] -> rows of `dst` (stride ks) by cp.async (global -> shared, no registers)
__device__ __forceinline__ void w2v_copy_a(__nv_bfloat16 *dst, int ks, const __nv_bfloat16 *src, int M, int Mpad, int ld, int k0, int kn) {
    for (int i = threadIdx.x; i < Mpad * (kn / 8); i += WS_THREADS_) {
        const int m = i / (kn / 8), j = (i - m * (kn / 8)) * 8;
        if (m < M) w2v_cp16(dst + (size_t)m * ks + j, src + (size_t)m * ld + k0 + j);
        else *reinterpret_cast<uint4 *>(dst + (size_t)m * ks + j) = make_uint4(0u, 0u, 0u, 0u);
    }
}
// One phase GEMM for this CTA: C[M x nc] = A[M x K] W[n0 .. n0+nc)[K]^T.  The weight slice is whole in sm.W[wslot] (in flight since
// before the preceding grid barrier); the A operand is whole in sm.A when K <= WS_KA (built by the caller before the call), else
// streamed from `a_src` (bf16 [M][K]) through a WS_STAGES-deep cp.async ring of WS_KC-column chunks.  epi(m, n, value) is called for
// every valid (m < M, n < nc) exactly once; the order of summation is fixed.
// Work split: (m-tile, n-tile) pairs over the 8 warps; with fewer pairs than warps the k-steps are split over the spare warps and
// the partials are added in warp order through shared memory.
// pre(m, n) -> float2 is called for the thread's (up to four) outputs BEFORE the MMA loop and handed to epi(m, n, value, pre): the bias
// (cold in HBM: part of the weight image) and the residual element are loaded while the GEMM runs instead of one L2 / HBM round trip
// after it (~1 us per phase).
template <class Pre, class Epi>
__device__ __forceinline__ void w2v_phase_gemm(W2vSmem &sm, int K, int nc, int M, const __nv_bfloat16 *a_src, Pre pre, Epi epi, uint32_t (&wphase)[2],
                                               int &wslot, const W2vStackParams &dbg_p, int dbg_l = -1, int dbg_i = 0) {
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int mt = (M + 15) / 16, Mpad = mt * 16, nt = nc / 8, pairs = mt * nt;
    const int ksplit = pairs >= 8 ? 1 : (8 / pairs >= 4 ? 4 : (8 / pairs >= 2 ? 2 : 1));
    const int groups = 8 / ksplit;                 // warps working on distinct pairs at a time
    const int kpart = warp / groups, pg = warp % groups;
    const int pr = pg;                             // at most one pair per warp: mt <= 2, nt <= 4
    const bool active = pr < pairs;
    const int mi = active ? pr / nt : 0, ni = active ? pr - mi * nt : 0;
    const int wks = K + 8;
    float acc[4] = {0.f, 0.f, 0.f, 0.f};
    float2 pv[4] = {make_float2(0.f, 0.f), make_float2(0.f, 0.f), make_float2(0.f, 0.f), make_float2(0.f, 0.f)};
    if (active && kpart == 0) {
        const int r0 = mi * 16 + (lane >> 2), cc = ni * 8 + 2 * (lane & 3);
        if (r0 < M) { pv[0] = pre(r0, cc); pv[1] = pre(r0, cc + 1); }
        if (r0 + 8 < M) { pv[2] = pre(r0 + 8, cc); pv[3] = pre(r0 + 8, cc + 1); }
    }
    const __nv_bfloat16 *Wsm = sm.W[wslot];
    const int q = lane >> 3;
    if (K <= WS_KA) {
        __syncthreads();                           // the caller's A is complete
        mbar_wait(&sm.wbar[wslot], wphase[wslot]);
        if (dbg_l >= 0) w2v_stamp(dbg_p, dbg_l, dbg_i);
        const int ksteps = K / 16, kper = ksteps / ksplit, aks = K + 8;
        if (active)
#pragma unroll 4
            for (int ks = kpart * kper; ks < (kpart + 1) * kper; ks++) {
                uint32_t a[4], b0, b1;
                ldmatrix_x4(a, smem_u32(sm.A + (size_t)(mi * 16 + (lane & 7) + 8 * (q & 1)) * aks + ks * 16 + 8 * (q >> 1)));
                ldmatrix_x2(b0, b1, smem_u32(Wsm + (size_t)(ni * 8 + (lane & 7)) * wks + ks * 16 + 8 * (q & 1)));
                mma_bf16_16816(acc, a, b0, b1);
            }
    } else {
        const int nchunks = K / WS_KC;
        __nv_bfloat16 *idle = sm.W[wslot ^ 1];     // the next phase's weights are issued only after this call returns
        auto stage = [&](int c) { const int st = c % WS_STAGES; return st < WS_STAGES / 2 ? sm.A + (size_t)st * Mpad * WS_KS : idle + (size_t)(st - WS_STAGES / 2) * Mpad * WS_KS; };
        static_assert((WS_STAGES / 2) * WS_MAX_MT * 16 * WS_KS <= WS_MAX_MT * 16 * (WS_KA + 8) + WS_RING_PAD && (WS_STAGES / 2) * WS_MAX_MT * 16 * WS_KS <= WS_WBUF_HALFS, "ring stages do not fit");
        __syncthreads();                           // everyone is done with sm.A of the phase before
        for (int c = 0; c < WS_STAGES - 1 && c < nchunks; c++) {
            w2v_copy_a(stage(c), WS_KS, a_src, M, Mpad, K, c * WS_KC, WS_KC);
            w2v_cp_commit();
        }
        mbar_wait(&sm.wbar[wslot], wphase[wslot]);
        if (dbg_l >= 0) w2v_stamp(dbg_p, dbg_l, dbg_i);
        for (int kc = 0; kc < nchunks; kc++) {
            if (kc + WS_STAGES - 1 < nchunks) w2v_cp_wait<WS_STAGES - 2>(); else w2v_cp_wait<0>();
            __syncthreads();                       // chunk kc has landed for everyone; the stage refilled below was consumed in kc - 1
            if (kc + WS_STAGES - 1 < nchunks) {
                const int c = kc + WS_STAGES - 1;
                w2v_copy_a(stage(c), WS_KS, a_src, M, Mpad, K, c * WS_KC, WS_KC);
                w2v_cp_commit();
            }
            const __nv_bfloat16 *As = stage(kc);
            const int kper = (WS_KC / 16) / ksplit;
            if (active)
                for (int ks = kpart * kper; ks < (kpart + 1) * kper; ks++) {
                    uint32_t a[4], b0, b1;
                    ldmatrix_x4(a, smem_u32(As + (size_t)(mi * 16 + (lane & 7) + 8 * (q & 1)) * WS_KS + ks * 16 + 8 * (q >> 1)));
                    ldmatrix_x2(b0, b1, smem_u32(Wsm + (size_t)(ni * 8 + (lane & 7)) * wks + kc * WS_KC + ks * 16 + 8 * (q & 1)));
                    mma_bf16_16816(acc, a, b0, b1);
                }
        }
    }
    wphase[wslot] ^= 1u;
    wslot ^= 1;
    if (dbg_l >= 0) { __syncthreads(); w2v_stamp(dbg_p, dbg_l, dbg_i + 1); }
    // split-K reduction in warp order, then the epilogue by the kpart == 0 warps
    if (ksplit > 1) {
        if (active && kpart > 0) {
#pragma unroll
            for (int c = 0; c < 4; c++) sm.red[warp][lane][c] = acc[c];
        }
        __syncthreads();
        if (active && kpart == 0)
            for (int kp = 1; kp < ksplit; kp++)
#pragma unroll
                for (int c = 0; c < 4; c++) acc[c] += sm.red[kp * groups + pg][lane][c];
    }
    if (active && kpart == 0) {
        const int r0 = mi * 16 + (lane >> 2), cc = ni * 8 + 2 * (lane & 3);
        if (r0 < M) { epi(r0, cc, acc[0], pv[0]); epi(r0, cc + 1, acc[1], pv[1]); }
        if (r0 + 8 < M) { epi(r0 + 8, cc, acc[2], pv[2]); epi(r0 + 8, cc + 1, acc[3], pv[3]); }
    }
    __syncthreads();
}

__device__ __forceinline__ float w2v_gelu(float x) { return 0.5f * x * (1.0f + erff(x * 0.70710678118654752f)); }

// this CTA's column slice of an N-wide output
__device__ __forceinline__ void w2v_slice(int N, int &n0, int &nc) {
    int per = (N + WS_G - 1) / WS_G;
    per = (per + 7) / 8 * 8;
    n0 = blockIdx.x * per;
    nc = n0 >= N ? 0 : min(per, N - n0);
}

struct ZeroParams { unsigned *p; int n; };
__global__ void k_zero_u32(const ZeroParams z) {
    pdl_launch();
    pdl_wait();
    if ((int)threadIdx.x < z.n) z.p[threadIdx.x] = 0u;
}

__global__ void __launch_bounds__(WS_THREADS_, 1) k_w2v_stack(const __grid_constant__ W2vStackParams p) {
    pdl_launch();
    pdl_wait();
    extern __shared__ __align__(128) unsigned char w2v_smem_raw[];
    W2vSmem &sm = *reinterpret_cast<W2vSmem *>(w2v_smem_raw);
    const int D = p.D, I = p.I, M = p.M, Mpad = (M + 15) / 16 * 16;
    unsigned gen = 0;
    uint32_t wphase[2] = {0u, 0u};
    int wslot = 0;                                 // sm.W[wslot] holds (or is receiving) the NEXT phase's weight slice
    if (threadIdx.x == 0) {
        mbar_init(&sm.wbar[0], 1);
        mbar_init(&sm.wbar[1], 1);
        fence_barrier_init();
    }
    // residual stream in fp32
    for (size_t i = (size_t)blockIdx.x * WS_THREADS_ + threadIdx.x; i < (size_t)M * D; i += (size_t)gridDim.x * WS_THREADS_)
        p.xres[i] = __bfloat162float(p.x_in[i]);
    __syncthreads();
    int n0q, ncq, n0d, ncd, n0i, nci;
    w2v_slice(3 * D, n0q, ncq);
    w2v_slice(D, n0d, ncd);
    w2v_slice(I, n0i, nci);
    // FFN2 (K = I = 4096) split over the K dimension inside the cluster: rank j multiplies columns [j * I / csize, ..) of hid with the
    // matching weight columns for ALL the cluster's output columns (A and W slices whole in shared memory: no ring, a quarter of the A
    // traffic), the partial sums meet through distributed shared memory in rank order (deterministic)
    const int crank = (int)w2v_cluster_rank(), csize = (int)w2v_cluster_size();
    const int Ks = I / csize;
    const bool p5_split = csize > 1 && I > WS_KA && I % csize == 0 && Ks <= WS_KA && Ks % 64 == 0 && ncd * csize <= WS_MAX_NC &&
                          D % ((int)gridDim.x * 8) == 0 && D / (int)gridDim.x == ncd;
    const int n0c = ((int)blockIdx.x - crank) * ncd;   // first output column of the cluster
    // first weight chunk of layer 0 / P1 in flight before the first barrier
    {
        const W2vLayerPtrs w0 = w2v_layer(p.image, 0, D, I);
        if (threadIdx.x == 0 && ncq) w2v_issue_w(sm, w0.Wqkv, D, n0q, ncq, wslot);
    }
    w2v_grid_barrier(p.barrier, gen);

    for (int l = 0; l < p.layers; l++) {
        const W2vLayerPtrs w = w2v_layer(p.image, l, D, I);
        w2v_stamp(p, l, 0);
        // ---- P1: qkv = LN1(x) Wqkv^T + b
        w2v_ln_rows(sm, p.xres, M, Mpad, D, p.eps, p, l);   // every CTA: the rows are shared inside the cluster
        w2v_stamp(p, l, 15);
        if (ncq) {
            w2v_phase_gemm(sm, D, ncq, M, nullptr,
                           [&](int, int n) { return make_float2(__ldg(w.bqkv + n0q + n), 0.f); },
                           [&](int m, int n, float v, float2 pr) { p.qkv[(size_t)m * 3 * D + n0q + n] = __float2bfloat16_rn(v + pr.x); },
                           wphase, wslot, p, l, 11);
        }
        if (threadIdx.x == 0 && ncd) w2v_issue_w(sm, w.Wo, D, n0d, ncd, wslot);         // P3's weights stream during P2
        w2v_stamp(p, l, 1);
        w2v_grid_barrier(p.barrier, gen);
        w2v_stamp(p, l, 2);
        // ---- P2: attention; item = (window, head, group of query rows) so that all CTAs take part (one CTA per head took 30 us)
        {
            const int dh = D / p.heads, T = p.T;
            const int RG = max(1, min(T, (int)gridDim.x / (p.B * p.heads)));      // query-row groups per (window, head)
            const int rows_per = (T + RG - 1) / RG;
            float *sq = reinterpret_cast<float *>(sm.A);             // q rows of the group | k [T][dh + 1] | v [T][dh], scores [rows_per][T]
            const int dk = dh + 1;                                   // k rows padded: the score loop reads k[tk][d] with tk across the lanes
            float *sk = sq + rows_per * dh, *sv = sk + T * dk, *sc = sv + T * dh;
            for (int item = blockIdx.x; item < p.B * p.heads * RG; item += gridDim.x) {
                const int bh = item / RG, rg = item - bh * RG, b = bh / p.heads, h = bh - b * p.heads;
                const int t0 = rg * rows_per, nr = min(rows_per, T - t0);
                if (nr <= 0) continue;
                const int nvec = (nr + 2 * T) * (dh / 8);
                for (int i = threadIdx.x; i < nvec; i += WS_THREADS_) {                 // 16-byte loads, all in flight together
                    const int r = i / (dh / 8), d = (i - r * (dh / 8)) * 8;
                    const int which = r < nr ? 0 : (r < nr + T ? 1 : 2), t = which == 0 ? t0 + r : (which == 1 ? r - nr : r - nr - T);
                    const uint4 v = __ldcg(reinterpret_cast<const uint4 *>(p.qkv + (size_t)(b * T + t) * 3 * D + which * D + h * dh + d));   // not through L1
                    const uint32_t wv[4] = {v.x, v.y, v.z, v.w};
                    float *dstf = (which == 0 ? sq + r * dh : (which == 1 ? sk + t * dk : sv + t * dh)) + d;
#pragma unroll
                    for (int j = 0; j < 4; j++) {
                        dstf[2 * j] = __uint_as_float(wv[j] << 16);
                        dstf[2 * j + 1] = __uint_as_float(wv[j] & 0xffff0000u);
                    }
                }
                __syncthreads();
                for (int i = threadIdx.x; i < nr * T; i += WS_THREADS_) {
                    const int tq = i / T, tk = i - tq * T;
                    float s0 = 0.f, s1 = 0.f;
                    for (int d = 0; d < dh; d += 2) {
                        s0 = fmaf(sq[tq * dh + d], sk[tk * dk + d], s0);
                        s1 = fmaf(sq[tq * dh + d + 1], sk[tk * dk + d + 1], s1);
                    }
                    sc[i] = (s0 + s1) * p.scale_log2;
                }
                __syncthreads();
                for (int tq = threadIdx.x >> 5; tq < nr; tq += WS_THREADS_ / 32) {      // softmax per row, one warp per row
                    const int lane = threadIdx.x & 31;
                    float mx = -3.0e38f;
                    for (int tk = lane; tk < T; tk += 32) mx = fmaxf(mx, sc[tq * T + tk]);
#pragma unroll
                    for (int o = 16; o; o >>= 1) mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, o));
                    float sum = 0.f;
                    for (int tk = lane; tk < T; tk += 32) { const float e = exp2f(sc[tq * T + tk] - mx); sc[tq * T + tk] = e; sum += e; }
#pragma unroll
                    for (int o = 16; o; o >>= 1) sum += __shfl_xor_sync(0xffffffffu, sum, o);
                    const float inv = 1.0f / sum;
                    for (int tk = lane; tk < T; tk += 32) sc[tq * T + tk] = __bfloat162float(__float2bfloat16_rn(sc[tq * T + tk] * inv));
                }
                __syncthreads();
                for (int i = threadIdx.x; i < nr * dh; i += WS_THREADS_) {
                    const int t = i / dh, d = i - t * dh;
                    float o0 = 0.f, o1 = 0.f;
                    int tk = 0;
                    for (; tk + 1 < T; tk += 2) {
                        o0 = fmaf(sc[t * T + tk], sv[tk * dh + d], o0);
                        o1 = fmaf(sc[t * T + tk + 1], sv[(tk + 1) * dh + d], o1);
                    }
                    if (tk < T) o0 = fmaf(sc[t * T + tk], sv[tk * dh + d], o0);
                    p.ao[(size_t)(b * T + t0 + t) * D + h * dh + d] = __float2bfloat16_rn(o0 + o1);
                }
                __syncthreads();
            }
        }
        w2v_stamp(p, l, 3);
        w2v_grid_barrier(p.barrier, gen);
        w2v_stamp(p, l, 4);
        // ---- P3: x += ao Wo^T + b
        if (ncd) {
            w2v_copy_a(sm.A, D + 8, p.ao, M, Mpad, D, 0, D);
            w2v_cp_commit();
            w2v_cp_wait<0>();
            w2v_phase_gemm(sm, D, ncd, M, nullptr,
                           [&](int m, int n) { return make_float2(__ldg(w.bo + n0d + n), __ldcg(p.xres + (size_t)m * D + n0d + n)); },
                           [&](int m, int n, float v, float2 pr) { p.xres[(size_t)m * D + n0d + n] = pr.y + (v + pr.x); },
                           wphase, wslot, p);
        }
        if (threadIdx.x == 0 && nci) w2v_issue_w(sm, w.W1, D, n0i, nci, wslot);
        w2v_stamp(p, l, 5);
        w2v_grid_barrier(p.barrier, gen);
        w2v_stamp(p, l, 6);
        // ---- P4: hid = gelu(LN2(x) W1^T + b)
        w2v_ln_rows(sm, p.xres, M, Mpad, D, p.eps, p);
        if (nci) {
            w2v_phase_gemm(sm, D, nci, M, nullptr,
                           [&](int, int n) { return make_float2(__ldg(w.b1 + n0i + n), 0.f); },
                           [&](int m, int n, float v, float2 pr) { p.hid[(size_t)m * I + n0i + n] = __float2bfloat16_rn(w2v_gelu(v + pr.x)); },
                           wphase, wslot, p);
        }
        if (threadIdx.x == 0 && ncd) {
            if (p5_split) w2v_issue_w_sub(sm, w.W2, I, crank * Ks, Ks, n0c, ncd * csize, wslot);
            else w2v_issue_w(sm, w.W2, I, n0d, ncd, wslot);
        }
        w2v_stamp(p, l, 7);
        w2v_grid_barrier(p.barrier, gen);
        w2v_stamp(p, l, 8);
        // ---- P5: x += hid W2^T + b
        if (p5_split) {
            float *part = reinterpret_cast<float *>(sm.red);            // [M][ncd * csize] partial sums of this rank's K slice
            const int ncc = ncd * csize;
            w2v_copy_a(sm.A, Ks + 8, p.hid, M, Mpad, I, crank * Ks, Ks);
            w2v_cp_commit();
            w2v_cp_wait<0>();
            // this thread's output element of the final reduction (M * ncd <= 256): residual and bias are loaded before the GEMM
            const int om = (int)threadIdx.x / ncd, on = (int)threadIdx.x - om * ncd;
            const bool mine = (int)threadIdx.x < M * ncd;
            float xr = 0.f, br = 0.f;
            if (mine) { xr = __ldcg(p.xres + (size_t)om * D + n0d + on); br = __ldg(w.b2 + n0d + on); }
            w2v_phase_gemm(sm, Ks, ncc, M, nullptr, [](int, int) { return make_float2(0.f, 0.f); },
                           [&](int m, int n, float v, float2) { part[m * ncc + n] = v; }, wphase, wslot, p, l, 13);
            w2v_cluster_sync();
            if (mine) {
                const uint32_t a = smem_u32(part + om * ncc + crank * ncd + on);
                float v = 0.f;
                for (int r = 0; r < csize; r++) v += w2v_ld_cluster_f32(a, (uint32_t)r);
                p.xres[(size_t)om * D + n0d + on] = xr + (v + br);
            }
            for (int i = threadIdx.x + WS_THREADS_; i < M * ncd; i += WS_THREADS_) {   // (not reached for M <= 32, ncd = 8)
                const int m = i / ncd, n = i - m * ncd;
                const uint32_t a = smem_u32(part + m * ncc + crank * ncd + n);
                float v = 0.f;
                for (int r = 0; r < csize; r++) v += w2v_ld_cluster_f32(a, (uint32_t)r);
                float *x = p.xres + (size_t)m * D + n0d + n;
                *x = __ldcg(x) + (v + __ldg(w.b2 + n0d + n));
            }
        } else if (ncd) {
            if (I <= WS_KA) {
                w2v_copy_a(sm.A, I + 8, p.hid, M, Mpad, I, 0, I);
                w2v_cp_commit();
                w2v_cp_wait<0>();
            }
            w2v_phase_gemm(sm, I, ncd, M, p.hid,
                           [&](int m, int n) { return make_float2(__ldg(w.b2 + n0d + n), __ldcg(p.xres + (size_t)m * D + n0d + n)); },
                           [&](int m, int n, float v, float2 pr) { p.xres[(size_t)m * D + n0d + n] = pr.y + (v + pr.x); },
                           wphase, wslot, p, l, 13);
        }
        if (l + 1 < p.layers) {
            const W2vLayerPtrs wn = w2v_layer(p.image, l + 1, D, I);
            if (threadIdx.x == 0 && ncq) w2v_issue_w(sm, wn.Wqkv, D, n0q, ncq, wslot);
        }
        w2v_stamp(p, l, 9);
        w2v_grid_barrier(p.barrier, gen);
        w2v_stamp(p, l, 10);
    }
    for (size_t i = (size_t)blockIdx.x * WS_THREADS_ + threadIdx.x; i < (size_t)M * D; i += (size_t)gridDim.x * WS_THREADS_)
        p.x_out[i] = __float2bfloat16_rn(__ldcg(p.xres + i));
}
