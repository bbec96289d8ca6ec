// conv_tma.cuh -- k_conv_tma: the stride-1 implicit-GEMM convolution (and every Linear / 1x1 conv) with BOTH operands
// staged by TMA, persistent over a static tile schedule, accumulators double-buffered in TMEM.
//
//   D[128 px, BN ch] (fp32, TMEM) += A[128 px, 64 ch of one tap] (bf16) x B[BN, 64]^T (bf16),  UMMA 128 x BN x 16
//
//   * A: the activation tensor is NHWC bf16; one output tile is a TW x TH x TB box of output pixels (TW*TH*TB = 128, all
//     powers of two, chosen per layer to minimise the tile count).  The A tile of filter tap (dy, dx) and channel block c
//     is ONE 4-D TMA box {64 ch, TW, TH, TB} at (c, x0 + dx, y0 + dy, b0): its 128 rows of 128 bytes land in exactly the
//     K-major SWIZZLE_128B layout tcgen05 reads; zero padding and ragged edges are TMA out-of-bounds zero fill.  No thread
//     touches the operand bytes (the cp.async im2col producer of k_conv measured 2x slower than its own MMA pipe:
//     profiles/r01_conv_bound_experiment.md).
//   * A row-halo reuse: when the tile is one 128-pixel row segment and the taps come in runs of consecutive dx (3x3: three runs of
//     3), ONE box of 128 + ndx - 1 pixels is loaded per (dy, channel block) and the ndx shifted A operands are descriptors into it
//     (start address + j * 128 B): A ingest drops 3x for 3x3 convs.
//   * B: weights, K = tap * Cin + channel, stored re-tiled [Cout_pad / 16][K / 64][16][64] (2 KB contiguous pieces), 4-D TMA
//     box {64, 16, 1, BN / 16}: lands as BN rows of 128 bytes, the same K-major SWIZZLE_128B layout.
//   * roles (320 threads): warp 5 lane 0 = TMA producer, warp 4 lane 0 = MMA issuer (also owns the TMEM allocation),
//     warps 0-3 and 6-9 = epilogue (TMEM lane == tile row == output pixel; the two groups split the tile's columns).  Rings: smem full/empty (TMA <-> MMA) and TMEM
//     full/empty (MMA <-> epilogue, 2 accumulators): the epilogue of tile i overlaps the main loop of tile i + 1.
//   * BN is a runtime value (any multiple of 16 up to 256): idesc, stage size and ring depth are computed per launch, so a
//     320-channel layer runs as 2 x 160 instead of 5 x 64.
//   * split-K for small-M layers (the 8x8 / 16x16 UNet levels): every split writes its fp32 partial tile to a workspace; the
//     split that arrives last at the tile's counter adds all partials in split order (deterministic) and runs the epilogue.
//
// Eligibility (host, add_conv_tma): kind 0, input stride 1, no folded upsampling, Cin % 64 == 0.
#pragma once

#define CT_THREADS 320   // warps 0-3 and 6-9: epilogue (two column halves), warp 4: MMA, warp 5: TMA producer
#define CT_MAX_STAGES 8
#define CT_SMEM_LIMIT (227 * 1024)
#define CT_ACC_STRIDE 256  // TMEM columns between the two accumulators

struct ConvTmaParams {
    alignas(64) CUtensorMap amap;
    alignas(64) CUtensorMap wmap;
    void *out;
    const __nv_bfloat16 *res;
    const float *scale, *shift;
    float *ws;            // split-K partial tiles [m_tile][n_tile][split][128][BN] fp32
    unsigned *counters;   // [m_tile * n_tiles], zero between launches
    int out_stride, out_coff, Hout, Wout;
    int res_stride, res_coff;
    int Mh, Mw, B, oy0, ox0, osy, osx;
    int in_coff, ntaps, cblocks, nkb, Cout, relu, flags;
    int BN, n_tiles, splits, stages;
    int lTW, lTH, tiles_x, tiles_y, total_items;
    float *gn_partial;   // fused GroupNorm statistics of the OUTPUT tensor (nullable): [B][tiles per image * 4][G][2]
    int gn_cpg, gn_G;    // channels per group (4 / 8 / 16) and number of groups
    int ndx, n_groups, a_stage_bytes;   // row-halo A reuse: a k-step is (tap group, channel block): one A load, ndx B loads, 4 ndx UMMAs
    int dbg;
    uint32_t mg_tx, mg_txy;   // ceil(2^32 / tiles_x), ceil(2^32 / (tiles_x * tiles_y)) for ct_fastdiv
    int mode;        // 0 bf16 NHWC; 1 wav2lip head (sigmoid, x255 truncated, u8 + fp32); 2 VAE head ((x/2+.5).clamp, round, BGR u8 + RGB fp32);
                     // 3 fp32 tokens x Cout to out_f32 (wav2vec2 logits)
    float *out_f32;  // modes 1 / 2
    int8_t tap_dy[CONV_MAX_TAPS], tap_dx[CONV_MAX_TAPS];   // per tap GROUP: dy, first dx
    int8_t grp_tap0[CONV_MAX_TAPS];                         // per tap group: index of its first tap in the weight K order
};

__device__ __forceinline__ void tma_prefetch_desc(const CUtensorMap *map) {
    asm volatile("prefetch.tensormap [%0];\n" ::"l"(map) : "memory");
}
__device__ __forceinline__ void tmem_ld16_nowait(uint32_t taddr, uint32_t (&v)[16]) {
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];\n"
        : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]),
          "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15])
        : "r"(taddr)
        : "memory");
}
__device__ __forceinline__ void tmem_ld_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;\n" ::: "memory"); }
__device__ __forceinline__ void epi_bar_sync() { asm volatile("bar.sync 1, 256;\n" ::: "memory"); }

// ---- cta_group::2 (CTA pair) helpers: PTX forms as in cute/arch/copy_sm100_tma.hpp, mma_sm100_umma.hpp, cutlass/arch/barrier.h ----
__device__ __forceinline__ uint32_t cluster_ctarank() { uint32_t r; asm volatile("mov.u32 %0, %%cluster_ctarank;\n" : "=r"(r)); return r; }
__device__ __forceinline__ void cluster_sync_all() {
    asm volatile("barrier.cluster.arrive.release.aligned;\n" ::: "memory");
    asm volatile("barrier.cluster.wait.acquire.aligned;\n" ::: "memory");
}
// the pair leader's copy of a barrier: shared::cluster address of CTA rank 0 = own shared::cta address with the peer bit cleared
#define CT_PEER_BIT_MASK 0xFEFFFFFFu
__device__ __forceinline__ void tma_load_4d_2sm(void *smem_dst, const CUtensorMap *map, int c0, int c1, int c2, int c3, uint64_t *bar) {
    asm volatile(
        "cp.async.bulk.tensor.4d.cta_group::2.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6}], [%2];\n" ::
            "r"(smem_u32(smem_dst)),
        "l"(map), "r"(smem_u32(bar) & CT_PEER_BIT_MASK), "r"(c0), "r"(c1), "r"(c2), "r"(c3)
        : "memory");
}
__device__ __forceinline__ void tmem_alloc_2sm(uint32_t *slot, uint32_t ncols) {
    asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;\n" ::"r"(smem_u32(slot)), "r"(ncols) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;\n" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc_2sm(uint32_t addr, uint32_t ncols) {
    asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;\n" ::"r"(addr), "r"(ncols) : "memory");
}
__device__ __forceinline__ void umma_f16_2sm(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t acc) {
    asm volatile(
        "{\n.reg .pred p;\nsetp.ne.b32 p, %4, 0;\n"
        "tcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p;\n}\n" ::"r"(tmem_d),
        "l"(adesc), "l"(bdesc), "r"(idesc), "r"(acc)
        : "memory");
}
// arrives on the barrier at this shared-memory offset in BOTH CTAs of the pair once the MMAs issued so far have completed
__device__ __forceinline__ void umma_commit_2sm(uint64_t *bar) {
    asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;\n" ::"r"(smem_u32(bar)),
                 "h"((uint16_t)3)
                 : "memory");
}
__device__ __forceinline__ void mbar_arrive_leader(uint64_t *bar) {   // arrive on CTA rank 0's copy of `bar`
    uint32_t remote;
    asm volatile("mapa.shared::cluster.u32 %0, %1, %2;\n" : "=r"(remote) : "r"(smem_u32(bar)), "r"(0));
    // default semantics (.release at CTA scope), the form cutlass::arch::ClusterBarrier::arrive(cta_id) uses.  The barrier only
    // says "this CTA's tcgen05.ld of the accumulator have completed" (tcgen05.wait::ld + fence::before_thread_sync + bar.sync come
    // first), so nothing has to be made visible cluster-wide: `.release.cluster` compiled to MEMBAR.ALL.GPU + ERRBAR, i.e. thread 0
    // waited for its output stores to drain once per tile and everyone waited for thread 0 at the next epilogue barrier
    asm volatile("mbarrier.arrive.shared::cluster.b64 _, [%0];\n" ::"r"(remote) : "memory");
}

// scale/shift (staged in shared memory per tile: cs = scale[256] | shift[256], indexed by the column inside the tile) ->
// (+residual, already loaded: r0 | r1 = 16 bf16) -> activation -> bf16, 16 channels of one pixel
__device__ __forceinline__ float4 lds_f4(uint32_t saddr) {   // explicit ld.shared (a generic pointer compiles to LD.E: long scoreboard)
    float4 v;
    asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];\n" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "r"(saddr));
    return v;
}

// n / d with m = ceil(2^32 / d) from the host: exact while n * d < 2^32 (checked when the launch is planned)
__device__ __forceinline__ uint32_t ct_fastdiv(uint32_t n, uint32_t m, uint32_t d) { return d == 1u ? n : __umulhi(n, m); }

template <int ACT>   // 0 none, 1 ReLU, 2 exact-erf GELU: compile-time, so that the epilogue loop stays a few hundred instructions
__device__ __forceinline__ void ct_finish16(const ConvTmaParams &p, float (&f)[16], uint32_t cs, int c, int n0, size_t opix,
                                            const uint4 &r0, const uint4 &r1) {
    {   // cs = shared-memory address of scale[256] | shift[256]
        const uint32_t sh = cs + (256 + c) * 4, sc = cs + c * 4;
        if (p.flags & 2) {   // every scale is 1 (no folded BatchNorm): bias only
#pragma unroll
            for (int j = 0; j < 4; j++) {
                const float4 t = lds_f4(sh + 16 * j);
                f[4 * j] += t.x; f[4 * j + 1] += t.y; f[4 * j + 2] += t.z; f[4 * j + 3] += t.w;
            }
        } else {
#pragma unroll
            for (int j = 0; j < 4; j++) {
                const float4 t = lds_f4(sh + 16 * j), u = lds_f4(sc + 16 * j);
                f[4 * j] = fmaf(f[4 * j], u.x, t.x); f[4 * j + 1] = fmaf(f[4 * j + 1], u.y, t.y);
                f[4 * j + 2] = fmaf(f[4 * j + 2], u.z, t.z); f[4 * j + 3] = fmaf(f[4 * j + 3], u.w, t.w);
            }
        }
    }
    const bool res_late = p.flags & 1;
    if (ACT != 0 && res_late) {
#pragma unroll
        for (int j = 0; j < 16; j++) f[j] = ACT == 1 ? fmaxf(f[j], 0.f) : gelu_erf(f[j]);
    }
    if (p.res) {
        const uint32_t rw[8] = {r0.x, r0.y, r0.z, r0.w, r1.x, r1.y, r1.z, r1.w};
#pragma unroll
        for (int j = 0; j < 8; j++) {
            const __nv_bfloat162 h = *reinterpret_cast<const __nv_bfloat162 *>(&rw[j]);
            f[2 * j] += __bfloat162float(h.x);
            f[2 * j + 1] += __bfloat162float(h.y);
        }
    }
    uint32_t o[8];
#pragma unroll
    for (int j = 0; j < 8; j++) {
        float a = f[2 * j], c = f[2 * j + 1];
        if (ACT != 0 && !res_late) {
            if (ACT == 1) { a = fmaxf(a, 0.f); c = fmaxf(c, 0.f); }
            else { a = gelu_erf(a); c = gelu_erf(c); }
        }
        f[2 * j] = a; f[2 * j + 1] = c;
        __nv_bfloat162 h = __floats2bfloat162_rn(a, c);
        o[j] = *reinterpret_cast<uint32_t *>(&h);
    }
    if (p.mode == 3) {   // fp32 output head: the unrounded values, only the real columns
        float *dst = p.out_f32 + opix * p.Cout + n0;
#pragma unroll
        for (int j = 0; j < 8; j++) {
            if (n0 + 2 * j < p.Cout) dst[2 * j] = f[2 * j];
            if (n0 + 2 * j + 1 < p.Cout) dst[2 * j + 1] = f[2 * j + 1];
        }
        return;
    }
    uint4 *op = reinterpret_cast<uint4 *>(reinterpret_cast<__nv_bfloat16 *>(p.out) + opix * p.out_stride + p.out_coff + n0);
    op[0] = make_uint4(o[0], o[1], o[2], o[3]);
    op[1] = make_uint4(o[4], o[5], o[6], o[7]);
    if (p.gn_partial) {   // the fused GroupNorm statistics are taken from the values as stored (only the layers that feed a GroupNorm)
#pragma unroll
        for (int j = 0; j < 8; j++) {
            const __nv_bfloat162 h = *reinterpret_cast<const __nv_bfloat162 *>(&o[j]);
            f[2 * j] = __bfloat162float(h.x); f[2 * j + 1] = __bfloat162float(h.y);
        }
    }
}

// CG = 1: one CTA per 128 x BN tile.  CG = 2: a CTA pair (cluster of 2, tcgen05 cta_group::2) per 256 x BN tile: each CTA stages
// its own 128-row A tile and HALF of the B tile, the leader issues UMMA 256 x BN x 16 for both, each CTA keeps its 128 rows of the
// accumulator in its own TMEM and runs its own epilogue.  Per CTA and k-block that is 16 KB + BN * 64 B of TMA ingest and shared-
// memory operand reads instead of 16 KB + BN * 128 B -- the measured bound of the CG = 1 kernel.
// fused GroupNorm statistics: per-group (sum, sum of squares) of one 16-channel chunk over the 32 rows of this warp, reduced with
// shuffles and written by lane 0 to a FIXED slot (tile, 32-row quarter, group): no atomics, the finalise kernel adds the slots in
// index order.  NG = groups per chunk = 16 / channels-per-group.  Rows that are not stored contribute zeros.
template <int NG>
__device__ __forceinline__ void ct_gn_chunk(const float (&f)[16], bool valid, float *dst, int lane) {
    constexpr int CPG = 16 / NG;
    float gs[NG], gq[NG];
#pragma unroll
    for (int g = 0; g < NG; g++) {
        gs[g] = 0.f; gq[g] = 0.f;
#pragma unroll
        for (int j = 0; j < CPG; j++) {
            const float v = valid ? f[g * CPG + j] : 0.f;
            gs[g] += v; gq[g] = fmaf(v, v, gq[g]);
        }
    }
#pragma unroll
    for (int o = 16; o; o >>= 1)
#pragma unroll
        for (int g = 0; g < NG; g++) {
            gs[g] += __shfl_xor_sync(0xffffffffu, gs[g], o);
            gq[g] += __shfl_xor_sync(0xffffffffu, gq[g], o);
        }
    if (lane == 0) {
#pragma unroll
        for (int g = 0; g < NG; g++) reinterpret_cast<float2 *>(dst)[g] = make_float2(gs[g], gq[g]);
    }
}

// the epilogue is done with accumulator `acc`: CG 1 every thread arrives (count 256); CG 2 one elected thread per CTA arrives on
// the leader's barrier (count 2), after all 256 epilogue threads of this CTA have finished their TMEM loads
#define CT_RELEASE_ACC()                                                  \
    do {                                                                  \
        tc_fence_before();                                                \
        if (CG == 2) {                                                    \
            epi_bar_sync();                                               \
            if (threadIdx.x == 0) mbar_arrive_leader(&tempty[acc]);       \
        } else {                                                          \
            mbar_arrive(&tempty[acc]);                                    \
        }                                                                 \
    } while (0)

template <int CG, int ACT>
__global__ void __launch_bounds__(CT_THREADS, 1) k_conv_tma(const __grid_constant__ ConvTmaParams p) {   // 10 warps = 3 on one SM sub-partition (16 K registers): 168 registers per thread is the hardware cap
    extern __shared__ unsigned char smem_raw[];
    unsigned char *smem = reinterpret_cast<unsigned char *>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    const int ST = p.stages;
    const uint32_t B_STAGE = (uint32_t)p.BN * (128u / CG);   // this CTA's share of the B tile
    const uint32_t cta_rank = CG == 2 ? cluster_ctarank() : 0u;
    unsigned char *sA = smem;
    const uint32_t A_STAGE = (uint32_t)p.a_stage_bytes;              // 16 KB, or 17 KB with the row halo
    const uint32_t A_TX = (uint32_t)(CONV_BM + p.ndx - 1) * 128u;       // bytes one A box delivers
    const int NDX = p.ndx;
    unsigned char *sB = smem + ST * A_STAGE;
    uint64_t *bars = reinterpret_cast<uint64_t *>(sB + (size_t)ST * NDX * B_STAGE);
    uint64_t *full = bars, *empty = bars + CT_MAX_STAGES, *tfull = bars + 2 * CT_MAX_STAGES, *tempty = tfull + 2;
    uint32_t *tmem_slot = reinterpret_cast<uint32_t *>(tempty + 2);
    volatile int *last_flag = reinterpret_cast<volatile int *>(tmem_slot + 1);
    float *coef = reinterpret_cast<float *>(reinterpret_cast<unsigned char *>(bars) + 256);   // [2 (item parity)][scale 256 | shift 256]
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

    if (threadIdx.x == 0) {
        for (int i = 0; i < ST; i++) {
            mbar_init(&full[i], 1);
            mbar_init(&empty[i], 1);
        }
        for (int i = 0; i < 2; i++) {
            mbar_init(&tfull[i], 1);
            mbar_init(&tempty[i], CG == 2 ? 2 : 256);   // CG 2: one elected arrival per CTA of the pair, on the leader's copy
        }
        fence_barrier_init();
    }
    if (warp == 4) {
        if (CG == 2) tmem_alloc_2sm(tmem_slot, 512);
        else tmem_alloc(tmem_slot, 512);
    }
    tc_fence_before();
    if (CG == 2) cluster_sync_all();   // the peer's barriers are initialised before anything can signal them
    else __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;
    if (warp == 5 && lane == 0) {   // the descriptors are kernel arguments: fetched while the predecessor drains
        tma_prefetch_desc(&p.amap);
        tma_prefetch_desc(&p.wmap);
    }
    pdl_launch();
    pdl_wait();   // barrier init, TMEM allocation and descriptor fetch above overlap the predecessor's tail
    const int total = p.total_items;            // work items of a CTA (CG 1) or of a CTA pair (CG 2)
    const int per_tile = p.n_tiles * p.splits;
    const int item0 = blockIdx.x / CG, item_step = gridDim.x / CG;

    if (warp == 5) {
        // =========================== TMA producer ==============================================
        if (lane == 0) {
            int s = 0;
            uint32_t ph = 1;   // parity to wait for on empty[s]: a fresh barrier passes a wait on parity 1 (first lap = free)
            for (int item = item0; item < total; item += item_step) {
                const int mp = item / per_tile, rem = item - mp * per_tile;
                const int mt = mp * CG + (int)cta_rank;   // an m-tile past the end is all out-of-bounds: TMA zero-fills it
                const int nt = rem / p.splits, sp = rem - nt * p.splits;
                const int txi = mt % p.tiles_x, tyi = (mt / p.tiles_x) % p.tiles_y, tbi = mt / (p.tiles_x * p.tiles_y);
                const int x0 = txi << p.lTW, y0 = tyi << p.lTH, b0 = tbi << (7 - p.lTW - p.lTH);
                const int kb0 = (int)((long long)sp * p.nkb / p.splits), kb1 = (int)((long long)(sp + 1) * p.nkb / p.splits);
                int grp = kb0 / p.cblocks, cb = kb0 - grp * p.cblocks;
                for (int kb = kb0; kb < kb1; kb++) {
                    mbar_wait(&empty[s], ph);
                    unsigned char *a_dst = sA + (size_t)s * A_STAGE, *b_dst = sB + (size_t)s * NDX * B_STAGE;
                    const int ax = x0 + p.tap_dx[grp], ay = y0 + p.tap_dy[grp], ac = p.in_coff + cb * CONV_BK;
                    const int tap0 = p.grp_tap0[grp];
                    if (CG == 2) {
                        // both CTAs' loads complete on the LEADER's full barrier, which expects the bytes of both
                        if (cta_rank == 0) mbar_expect_tx(&full[s], 2 * (A_TX + NDX * B_STAGE));
                        tma_load_4d_2sm(a_dst, &p.amap, ac, ax, ay, b0, &full[s]);
                        for (int j = 0; j < NDX; j++)
                            tma_load_4d_2sm(b_dst + (size_t)j * B_STAGE, &p.wmap, 0, 0, (tap0 + j) * p.cblocks + cb,
                                            (nt * p.BN + (int)cta_rank * (p.BN >> 1)) >> 4, &full[s]);
                    } else {
                        mbar_expect_tx(&full[s], A_TX + NDX * B_STAGE);
                        if (!(p.dbg & 1)) tma_load_4d(a_dst, &p.amap, ac, ax, ay, b0, &full[s]);
                        else asm volatile("mbarrier.complete_tx.shared::cta.b64 [%0], %1;\n" ::"r"(smem_u32(&full[s])), "r"(A_TX) : "memory");
                        for (int j = 0; j < NDX; j++) {
                            if (!(p.dbg & 2)) tma_load_4d(b_dst + (size_t)j * B_STAGE, &p.wmap, 0, 0, (tap0 + j) * p.cblocks + cb, (nt * p.BN) >> 4, &full[s]);
                            else asm volatile("mbarrier.complete_tx.shared::cta.b64 [%0], %1;\n" ::"r"(smem_u32(&full[s])), "r"(B_STAGE) : "memory");
                        }
                    }
                    if (++cb == p.cblocks) { cb = 0; grp++; }
                    if (++s == ST) { s = 0; ph ^= 1; }
                }
            }
        }
    } else if (warp == 4) {
        // =========================== MMA issuer ================================================
        if (lane == 0 && cta_rank == 0) {
            // the issue loop runs on ONE thread: every instruction in it delays the next UMMA (at BN = 128 a k-block of four
            // 128x128x16 UMMAs is only 256 tensor cycles), so no division / modulo / descriptor rebuild per k-block
            const uint32_t idesc = make_idesc(CONV_BM * CG, p.BN);
            const uint64_t adesc0 = make_sdesc(smem_u32(sA)), bdesc0 = make_sdesc(smem_u32(sB));
            const uint32_t a_step = A_STAGE >> 4, b_step = (NDX * B_STAGE) >> 4;   // descriptor start-address units of 16 B
            const uint32_t b_tap = B_STAGE >> 4;
            // shifted A operand j of the row halo: start address + j * 128 B, nothing else: the 128B swizzle is a function of the
            // absolute shared-memory address bits (TMA wrote it that way and UMMA reads it that way), so the descriptor's matrix
            // base offset stays 0 (measured: setting it to j gives wrong results; tests/test_convnet_gpu.py row-halo cases)
            const uint64_t a_shift = 8ull;
            int s = 0;
            uint32_t ph = 0, n = 0;
            uint64_t adesc = adesc0, bdesc = bdesc0;
            for (int item = item0; item < total; item += item_step, n++) {
                const int rem = item % per_tile;
                const int sp = rem % p.splits;
                const int kb0 = (int)((long long)sp * p.nkb / p.splits), kb1 = (int)((long long)(sp + 1) * p.nkb / p.splits);
                const uint32_t acc = n & 1;
                if (n >= 2) mbar_wait(&tempty[acc], ((n >> 1) - 1) & 1);
                tc_fence_after();
                const uint32_t d_tmem = tmem_base + acc * CT_ACC_STRIDE;
                uint32_t accum = 0;
                for (int kb = kb0; kb < kb1; kb++) {
                    mbar_wait(&full[s], ph);
                    tc_fence_after();
                    uint64_t ad = adesc, bd = bdesc;
                    for (int j = 0; j < NDX; j++) {
                        if (CG == 2) {
                            umma_f16_2sm(d_tmem, ad, bd, idesc, accum);
                            umma_f16_2sm(d_tmem, ad + 2, bd + 2, idesc, 1);
                            umma_f16_2sm(d_tmem, ad + 4, bd + 4, idesc, 1);
                            umma_f16_2sm(d_tmem, ad + 6, bd + 6, idesc, 1);
                        } else if (!(p.dbg & 4)) {
                            umma_f16(d_tmem, ad, bd, idesc, accum);
                            umma_f16(d_tmem, ad + 2, bd + 2, idesc, 1);
                            umma_f16(d_tmem, ad + 4, bd + 4, idesc, 1);
                            umma_f16(d_tmem, ad + 6, bd + 6, idesc, 1);
                        }
                        accum = 1;
                        ad += a_shift; bd += b_tap;
                    }
                    if (CG == 2) umma_commit_2sm(&empty[s]); else umma_commit(&empty[s]);
                    adesc += a_step; bdesc += b_step;
                    if (++s == ST) { s = 0; ph ^= 1; adesc = adesc0; bdesc = bdesc0; }
                }
                if (CG == 2) umma_commit_2sm(&tfull[acc]); else umma_commit(&tfull[acc]);
            }
        }
    } else {
        // =========================== epilogue: TMEM lane == tile row == output pixel ===========
        // a warp may only touch TMEM lanes [32 (warp % 4), +32): warps 0-3 take the first half of the tile's columns,
        // warps 6-9 the second half
        const int r = (warp & 3) * 32 + lane;                 // tile row 0..127
        const int half = warp < 4 ? 0 : 1;
        const int c_half = ((p.BN + 31) / 32) * 16;           // columns [0, c_half) and [c_half, BN)
        const int cb = half ? c_half : 0, ce = half ? p.BN : min(c_half, p.BN);
        const bool leader = threadIdx.x == 0;
        const int TWm = (1 << p.lTW) - 1, THm = (1 << p.lTH) - 1;
        uint32_t n = 0;
        // the epilogue is instruction-issue bound on the short-K layers (8 warps, ~700 instructions per warp and tile for 32
        // columns, profiles/r01_conv64_ncu_summary.md): the tile coordinates come from two multiply-high divisions by constants
        // prepared on the host instead of five integer divisions per tile (nothing is carried across tiles: the kernel sits at
        // its register cap)
        const bool simple = per_tile == 1;   // one n-tile, no split-K: an item IS an m-tile
        for (int item = item0; item < total; item += item_step, n++) {
            int mt, nt, sp;
            if (simple) {
                mt = item * CG + (int)cta_rank; nt = 0; sp = 0;
            } else {
                const int mp = item / per_tile, rem = item - mp * per_tile;
                mt = mp * CG + (int)cta_rank;
                nt = rem / p.splits; sp = rem - nt * p.splits;
            }
            const int q1 = (int)ct_fastdiv((uint32_t)mt, p.mg_tx, (uint32_t)p.tiles_x);
            const int tbi = (int)ct_fastdiv((uint32_t)mt, p.mg_txy, (uint32_t)(p.tiles_x * p.tiles_y));
            const int txi = mt - q1 * p.tiles_x, tyi = q1 - tbi * p.tiles_y;
            const int mx = (txi << p.lTW) + (r & TWm), my = (tyi << p.lTH) + ((r >> p.lTW) & THm);
            const int b = (tbi << (7 - p.lTW - p.lTH)) + (r >> (p.lTW + p.lTH));
            const bool row_ok = mx < p.Mw && my < p.Mh && b < p.B;
            const size_t opix = ((size_t)b * p.Hout + (p.oy0 + p.osy * my)) * p.Wout + (p.ox0 + p.osx * mx);
            const uint32_t acc = n & 1;
            const int n_base = nt * p.BN;
            // this tile's scale / shift go to shared memory once (the loads are in flight while the accumulator is awaited):
            // per-chunk __ldg of the bias was the largest stall of the epilogue (profiles/r01_conv_bound_experiment.md)
            // (a layer with ONE n-tile has the same coefficients for every item: staged once, no per-tile load / store / barrier)
            const int et = half * 128 + r;
            const bool stage_coef = p.n_tiles > 1 || n == 0;
            float my_sc = 1.f, my_sh = 0.f;
            if (stage_coef && et < p.BN && n_base + et < p.Cout) { my_sc = __ldg(p.scale + n_base + et); my_sh = __ldg(p.shift + n_base + et); }
            mbar_wait(&tfull[acc], (n >> 1) & 1);
            tc_fence_after();
            float *csp = coef + (p.n_tiles > 1 ? (n & 1) * 512 : 0);
            const uint32_t cs = smem_u32(csp);
            if (stage_coef) {
                csp[et] = my_sc; csp[256 + et] = my_sh;
                epi_bar_sync();
            }
            const uint32_t taddr = tmem_base + acc * CT_ACC_STRIDE + ((uint32_t)((warp & 3) * 32) << 16);
            if (p.mode == 1 || p.mode == 2) {
                // image output heads (Cout <= 16, one 16-column chunk; splits == 1)
                uint32_t v[16];
                tmem_ld16_nowait(taddr, v);
                tmem_ld_wait();
                CT_RELEASE_ACC();
                if (row_ok && half == 0) {
                    for (int j = 0; j < p.Cout; j++) {
                        const float a = fmaf(__uint_as_float(v[j]), __ldg(p.scale + j), __ldg(p.shift + j));
                        if (p.mode == 2) {   // musetalk/models/vae.py:102-107
                            const float im = fminf(fmaxf(a * 0.5f + 0.5f, 0.f), 1.f);
                            if (p.out_f32) p.out_f32[opix * p.Cout + j] = im;
                            if (p.out) reinterpret_cast<uint8_t *>(p.out)[opix * p.Cout + (p.Cout - 1 - j)] = (uint8_t)rintf(im * 255.f);
                        } else {             // wav2lip.py:83-85 sigmoid; lipreal.py:126,209 x255 truncated
                            const float sg = 1.0f / (1.0f + __expf(-a));
                            if (p.out_f32) p.out_f32[opix * p.Cout + j] = sg;
                            if (p.out) reinterpret_cast<uint8_t *>(p.out)[opix * p.Cout + j] = (uint8_t)(sg * 255.f);
                        }
                    }
                }
            } else if (p.splits == 1) {
#pragma unroll 1
                for (int c0 = cb; c0 < ce; c0 += 64) {
                    // one group = up to four 16-column chunks: their TMEM loads and residual loads are all issued before the
                    // single wait (the per-chunk load -> wait -> use chain was the epilogue's largest stall)
                    bool doc[4];
                    uint4 ra[4], rb[4];
                    uint32_t v[4][16];
#pragma unroll
                    for (int u = 0; u < 4; u++) {
                        const int cc = c0 + 16 * u;
                        doc[u] = row_ok && cc < ce && n_base + cc < p.Cout;
                        ra[u] = make_uint4(0u, 0u, 0u, 0u); rb[u] = ra[u];
                        if (p.res && doc[u]) {
                            const uint4 *rp = reinterpret_cast<const uint4 *>(p.res + opix * p.res_stride + p.res_coff + n_base + cc);
                            ra[u] = __ldg(rp); rb[u] = __ldg(rp + 1);
                        }
                    }
#pragma unroll
                    for (int u = 0; u < 4; u++)
                        if (c0 + 16 * u < ce) tmem_ld16_nowait(taddr + c0 + 16 * u, v[u]);
                    tmem_ld_wait();
#pragma unroll
                    for (int u = 0; u < 4; u++) {
                        float f[16];
#pragma unroll
                        for (int j = 0; j < 16; j++) f[j] = __uint_as_float(v[u][j]);
                        if (doc[u]) ct_finish16<ACT>(p, f, cs, c0 + 16 * u, n_base + c0 + 16 * u, opix, ra[u], rb[u]);
                        const int cc = c0 + 16 * u;
                        if (p.gn_partial && cc < ce && n_base + cc < p.Cout && tbi < p.B) {   // warp-uniform condition
                            // tiles are single-image here (TB = 1): slot = (image, tile in image, 32-row quarter)
                            const int tiles_img = p.tiles_x * p.tiles_y;
                            float *dst = p.gn_partial + ((((size_t)tbi * tiles_img + (mt - tbi * tiles_img)) * 4 + (warp & 3)) * p.gn_G +
                                                         (n_base + cc) / p.gn_cpg) * 2;
                            if (p.gn_cpg == 4) ct_gn_chunk<4>(f, doc[u], dst, lane);
                            else if (p.gn_cpg == 8) ct_gn_chunk<2>(f, doc[u], dst, lane);
                            else ct_gn_chunk<1>(f, doc[u], dst, lane);
                        }
                    }
                }
                CT_RELEASE_ACC();
            } else {
                // split-K: dump the fp32 partial, then the last-arriving split reduces in split order
                float *wtile = p.ws + ((size_t)(mt * p.n_tiles + nt) * p.splits) * (128 * (size_t)p.BN);
                float *mine = wtile + ((size_t)sp * 128 + r) * p.BN;
#pragma unroll 1
                for (int c0 = cb; c0 < ce; c0 += 16) {
                    uint32_t v[16];
                    tmem_ld16_nowait(taddr + c0, v);
                    tmem_ld_wait();
                    float4 *d = reinterpret_cast<float4 *>(mine + c0);
#pragma unroll
                    for (int j = 0; j < 4; j++)
                        d[j] = make_float4(__uint_as_float(v[4 * j]), __uint_as_float(v[4 * j + 1]), __uint_as_float(v[4 * j + 2]), __uint_as_float(v[4 * j + 3]));
                }
                CT_RELEASE_ACC();
                __threadfence();
                epi_bar_sync();
                if (leader) *last_flag = atomicAdd(p.counters + mt * p.n_tiles + nt, 1u) == (unsigned)(p.splits - 1);
                epi_bar_sync();
                const bool last = *last_flag != 0;
                epi_bar_sync();  // everyone has read the flag before the next item may overwrite it
                if (last) {
                    __threadfence();
                    if (row_ok) {
                        // all partials of a 16-column chunk are requested before the first add (up to 4 splits x 4 float4 in
                        // flight per thread; a dependent load-add chain here cost ~0.6 us of L2 latency per split per chunk);
                        // the adds run in split order, so the result does not depend on arrival order
#pragma unroll 1
                        for (int c0 = cb; c0 < ce && n_base + c0 < p.Cout; c0 += 16) {
                            float f[16];
#pragma unroll
                            for (int j = 0; j < 16; j++) f[j] = 0.f;
                            for (int s0 = 0; s0 < p.splits; s0 += 4) {
                                float4 t[4][4];
#pragma unroll
                                for (int u = 0; u < 4; u++) {
                                    const int s2 = min(s0 + u, p.splits - 1);
                                    const float4 *q = reinterpret_cast<const float4 *>(wtile + ((size_t)s2 * 128 + r) * p.BN + c0);
#pragma unroll
                                    for (int j = 0; j < 4; j++) t[u][j] = __ldcg(q + j);
                                }
#pragma unroll
                                for (int u = 0; u < 4; u++) {
                                    if (s0 + u < p.splits) {
#pragma unroll
                                        for (int j = 0; j < 4; j++) {
                                            f[4 * j] += t[u][j].x; f[4 * j + 1] += t[u][j].y; f[4 * j + 2] += t[u][j].z; f[4 * j + 3] += t[u][j].w;
                                        }
                                    }
                                }
                            }
                            uint4 q0 = make_uint4(0u, 0u, 0u, 0u), q1 = q0;
                            if (p.res) {
                                const uint4 *rp = reinterpret_cast<const uint4 *>(p.res + opix * p.res_stride + p.res_coff + n_base + c0);
                                q0 = __ldg(rp); q1 = __ldg(rp + 1);
                            }
                            ct_finish16<ACT>(p, f, cs, c0, n_base + c0, opix, q0, q1);
                        }
                    }
                    if (leader) p.counters[mt * p.n_tiles + nt] = 0u;
                }
            }
        }
    }
    __syncwarp();
    tc_fence_before();
    if (CG == 2) cluster_sync_all();   // the peer may still signal this CTA's barriers / read its operands until it is done too
    else __syncthreads();
    if (warp == 4) {
        if (CG == 2) tmem_dealloc_2sm(tmem_base, 512);
        else tmem_dealloc(tmem_base, 512);
    }
}
