// wav2lip.cu -- sm_100a conv-net executor for the Wav2Lip head: every Conv2d / ConvTranspose2d
// (+ folded BatchNorm, residual add, ReLU / sigmoid) of wav2lip/models/wav2lip.py:12-125 runs as
// an implicit GEMM on the 5th-generation tensor cores:
//
//   D[128 output pixels, BN channels] (fp32, TMEM) += A[128, 64] (bf16, smem) x B[BN, 64]^T (bf16, smem)
//
//   * one CTA = one 128-pixel x BN-channel output tile, 5 warps:
//       warps 0-3  A producers: one output pixel per thread, im2col gather of 16-byte channel
//                  chunks with cp.async (zero-fill for padding) straight into the 128B-swizzled
//                  K-major layout tcgen05 expects; afterwards the same warps run the epilogue
//                  (tcgen05.ld TMEM -> registers -> scale/shift (BN) -> +residual -> ReLU -> bf16
//                  NHWC store, optionally at a channel offset of a concat buffer, so
//                  torch.cat((x, feats[-1]), dim=1) never copies)
//       warp 4     allocates TMEM; lane 0 issues tcgen05.mma (UMMA 128 x BN x 16) per 64-wide
//                  k-block and tcgen05.commit to recycle the smem stage
//       thread 0   also issues the TMA (cp.async.bulk.tensor.2d, SWIZZLE_128B) load of the
//                  weight tile B for each k-block
//   * 4-stage mbarrier ring (full: 128 producer arrivals + TMA tx bytes; empty: tcgen05.commit)
//   * K = taps x Cin (k index = tap * Cin + channel).  A stride-2 ConvTranspose2d is executed
//     as its 4 output-parity classes, each an ordinary gather with only the taps that hit real
//     inputs (no zero-insertion waste); a stride-1 one as a gather with negative tap offsets.
//
// The layer list itself is data: the Python packer (mere_fusion_b200/wav2lip_pack.py) emits a
// "program" of conv ops + buffer table; this file only executes it.
#include <cuda.h>

#include <algorithm>
#include <cstdlib>
#include <vector>

#include "convnet_ops.cuh"

#define CONV_BM 128
#define CONV_BK 64
#define CONV_THREADS 160
#define CONV_MAX_TAPS 128
#define A_STAGE_BYTES (CONV_BM * CONV_BK * 2)
#define GN_MAX_CTAS (148 * 8)

// blob entry ids (kind = 2)
enum { W2L_ID_PROGRAM = 1, W2L_ID_AUX = 2, W2L_ID_FIRST_TENSOR = 16 };

// ---- program records written by the packer (all int32, little endian) ----------------------
struct W2LHeader {
    int32_t n_buffers, n_ops, in_face_buf, in_mel_buf, face_hw, mel_h, mel_w, out_hw;
};
struct W2LBuffer {
    int32_t H, W, C, init_entry;  // init_entry > 0: constant buffer, bf16 [H,W,C] blob entry copied into every batch slot at load
};
// kind 0 = conv (fields as named).  Other kinds reuse the integer fields (see the packer, convnet_pack.py):
//   1 GroupNorm(+SiLU): in_buf -> out_buf, Cin = C, ntaps = groups, relu = silu, Kpad = eps (float bits),
//                       scale_entry / shift_entry = gamma / beta (fp32), Mh = stats slot
//   2 LayerNorm       : in_buf -> out_buf, Cin = C, Kpad = eps bits, gamma / beta as above
//   3 attention       : q = (in_buf, in_coff), k = (res_buf, res_coff), v = (Mh, Mw), out = (out_buf, out_coff),
//                       ntaps = heads, Cin = dim_head, Kpad = scale (float bits)
//   4 GEGLU           : in_buf [.., 2 * Cin] -> out_buf [.., Cin]
//   5 transformer stack (wav2vec2, w2v_stack.cuh): in_buf [T,1,D] -> out_buf [T,1,D], Cin = D, Mw = intermediate size, ntaps = heads,
//                       Mh = layers, Kpad = LayerNorm eps bits, w_entry = packed weights of all layers
struct W2LOp {
    int32_t in_buf, in_coff, out_buf, out_coff, res_buf, res_coff;
    int32_t Mh, Mw, oy0, ox0, osy, osx, isy, isx;
    int32_t ntaps, Cin, Kpad, Cout, Cout_pad, BN, relu, mode;
    int32_t w_entry, scale_entry, shift_entry, kind, ups, flags;
    int8_t tap_dy[CONV_MAX_TAPS], tap_dx[CONV_MAX_TAPS];
};

struct ConvParams {
    alignas(64) CUtensorMap wmap;
    const __nv_bfloat16 *in;
    void *out;
    const __nv_bfloat16 *res;
    const float *scale, *shift;
    float *out_f32;
    int in_stride, in_coff, Hin, Win;
    int out_stride, out_coff, Hout, Wout;
    int res_stride, res_coff;
    int Mh, Mw, oy0, ox0, osy, osx, isy, isx;
    int ntaps, Cin, nkb, Cout, M, relu, mode;
    int ups;  // input is a nearest-neighbour 2^ups upsampling of the stored tensor (folded into the gather)
    int flags;  // bit 0: residual is added AFTER the activation (Whisper: gelu(conv2(x)) + positional embedding)
    int dbg;  // timing experiments only (MF_CONV_DBG): 1 skip A loads, 2 skip B TMA, 4 skip MMA
    int8_t tap_dy[CONV_MAX_TAPS], tap_dx[CONV_MAX_TAPS];
};

// ---- PTX wrappers ---------------------------------------------------------------------------
__device__ __forceinline__ void mbar_arrive(uint64_t *bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];\n" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void cp_async16(uint32_t dst, const void *src, uint32_t src_bytes) {
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;\n" ::"r"(dst), "l"(src), "r"(src_bytes) : "memory");
}
__device__ __forceinline__ void cp_async_mbar_arrive_noinc(uint64_t *bar) {
    asm volatile("cp.async.mbarrier.arrive.noinc.shared::cta.b64 [%0];\n" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;\n" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_async_wait() {
    asm volatile("cp.async.wait_group %0;\n" ::"n"(N) : "memory");
}
__device__ __forceinline__ void tma_load_4d(void *smem_dst, const CUtensorMap *map, int c0, int c1, int c2, int c3, uint64_t *bar) {
    asm volatile(
        "cp.async.bulk.tensor.4d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6}], [%2];\n" ::
            "r"(smem_u32(smem_dst)),
        "l"(map), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "r"(c2), "r"(c3)
        : "memory");
}
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;\n" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;\n" ::: "memory"); }
__device__ __forceinline__ void tmem_alloc(uint32_t *slot, uint32_t ncols) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;\n" ::"r"(smem_u32(slot)), "r"(ncols)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;\n" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc(uint32_t addr, uint32_t ncols) {
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;\n" ::"r"(addr), "r"(ncols) : "memory");
}
__device__ __forceinline__ void umma_f16(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t acc) {
    asm volatile(
        "{\n.reg .pred p;\nsetp.ne.b32 p, %4, 0;\n"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n}\n" ::"r"(tmem_d),
        "l"(adesc), "l"(bdesc), "r"(idesc), "r"(acc)
        : "memory");
}
__device__ __forceinline__ void umma_commit(uint64_t *bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];\n" ::"r"(smem_u32(bar))
                 : "memory");
}
__device__ __forceinline__ void tmem_ld16(uint32_t taddr, uint32_t (&v)[16]) {
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];\n"
        : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]),
          "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15])
        : "r"(taddr)
        : "memory");
    asm volatile("tcgen05.wait::ld.sync.aligned;\n" ::: "memory");
}

__device__ __forceinline__ float gelu_erf(float x) { return 0.5f * x * (1.0f + erff(x * 0.70710678118654752f)); }

// K-major, SWIZZLE_128B shared-memory matrix descriptor (cute/arch/mma_sm100_desc.hpp SmemDescriptor):
// start address >> 4 [0,14), LBO >> 4 [16,30) (unused for swizzled K-major, 1), SBO >> 4 [32,46) = 1024 B between
// 8-row groups, version 1 [46,48), layout type SWIZZLE_128B = 2 [61,64)
__device__ __forceinline__ uint64_t make_sdesc(uint32_t saddr) {
    return (uint64_t)((saddr & 0x3FFFFu) >> 4) | (1ull << 16) | (64ull << 32) | (1ull << 46) | (2ull << 61);
}
// instruction descriptor, kind::f16 (InstrDescriptor): D fp32 (1 << 4), A/B bf16 (1 << 7, 1 << 10), both K-major,
// N >> 3 at [17,23), M >> 4 at [24,29)
__host__ __device__ constexpr uint32_t make_idesc(int M, int N) {
    return (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
}

// CONV_STAGES smem stages = k-blocks of loads in flight per CTA.  Two depths per
// tile width: 4 stages (2 CTAs / SM: one CTA's epilogue overlaps the other's main loop) for layers with
// many tiles, and a deep ring (1 CTA / SM) for the few-tile, long-K, load-latency-bound layers.
template <int BN, int CONV_STAGES>
struct ConvSmem {
    static constexpr int B_STAGE = BN * CONV_BK * 2;
    static constexpr int BAR_OFF = CONV_STAGES * (A_STAGE_BYTES + B_STAGE);
    static constexpr int TOTAL = BAR_OFF + 256 + 1024;  // + barriers/slot + alignment slack
};

template <int BN, int CONV_STAGES>
__global__ void __launch_bounds__(CONV_THREADS) k_conv(const __grid_constant__ ConvParams p) {
    extern __shared__ unsigned char smem_raw[];
    unsigned char *smem = reinterpret_cast<unsigned char *>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    unsigned char *sA = smem;
    unsigned char *sB = smem + CONV_STAGES * A_STAGE_BYTES;
    uint64_t *full = reinterpret_cast<uint64_t *>(smem + ConvSmem<BN, CONV_STAGES>::BAR_OFF);
    uint64_t *empty = full + CONV_STAGES;
    uint64_t *accum = empty + CONV_STAGES;
    uint32_t *tmem_slot = reinterpret_cast<uint32_t *>(accum + 1);
    constexpr uint32_t TMEM_COLS = BN < 32 ? 32 : BN;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int nkb = p.nkb;

    if (threadIdx.x == 0) {
        for (int i = 0; i < CONV_STAGES; i++) {
            mbar_init(&full[i], CONV_BM + 1);
            mbar_init(&empty[i], 1);
        }
        mbar_init(accum, 1);
        fence_barrier_init();
    }
    if (warp == 4) tmem_alloc(tmem_slot, TMEM_COLS);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;
    pdl_launch();
    pdl_wait();   // everything above is CTA-local set-up and overlaps the predecessor's tail

    if (warp < 4) {
        // =========================== A producer: one output pixel per thread ===================
        const int r = threadIdx.x;
        const int m = blockIdx.x * CONV_BM + r;
        const bool row_ok = m < p.M;
        const int mx = m % p.Mw, my = (m / p.Mw) % p.Mh, b = m / (p.Mw * p.Mh);
        const int iy0 = my * p.isy, ix0 = mx * p.isx;
        const __nv_bfloat16 *in_b = p.in + (size_t)b * p.Hin * p.Win * p.in_stride + p.in_coff;
        const uint32_t a_row = smem_u32(sA) + r * 128;
        const uint32_t sw = r & 7;
        int tap = 0, ch = 0;
        for (int kb = 0; kb < nkb; kb++) {
            const int s = kb % CONV_STAGES;
            if (kb >= CONV_STAGES) mbar_wait(&empty[s], ((kb / CONV_STAGES) - 1) & 1);
            if (threadIdx.x == 0) {
                if (p.dbg & 2) mbar_arrive(&full[s]);
                else {
                mbar_expect_tx(&full[s], ConvSmem<BN, CONV_STAGES>::B_STAGE);
                tma_load_4d(sB + s * ConvSmem<BN, CONV_STAGES>::B_STAGE, &p.wmap, 0, 0, kb, (blockIdx.y * BN) >> 4, &full[s]);
                }
            }
            if (p.dbg & 1) { mbar_arrive(&full[s]); continue; }
            // the 8 16-byte chunks of this k-block, walked as runs that stay inside one filter tap: the
            // bounds test and the source pointer are computed once per run, not once per chunk
            {
                const uint32_t dst0 = a_row + s * A_STAGE_BYTES;
                int c = 0;
                while (c < 8) {
                    const int n = min(8 - c, (p.Cin - ch) >> 3);
                    bool ok = row_ok && tap < p.ntaps;
                    const __nv_bfloat16 *src = p.in;
                    if (ok) {
                        const int iy = iy0 + p.tap_dy[tap], ix = ix0 + p.tap_dx[tap];
                        ok = iy >= 0 && iy < (p.Hin << p.ups) && ix >= 0 && ix < (p.Win << p.ups);
                        if (ok) src = in_b + ((iy >> p.ups) * p.Win + (ix >> p.ups)) * p.in_stride + ch;
                    }
                    const uint32_t nbytes = ok ? 16u : 0u;
                    const int step = ok ? 8 : 0;
                    for (int j = 0; j < n; j++) cp_async16(dst0 + (((c + j) ^ sw) << 4), src + j * step, nbytes);
                    c += n;
                    ch += n << 3;
                    if (ch >= p.Cin) { ch = 0; tap++; }
                }
            }
            // this thread's arrival on full[s] fires when its cp.asyncs above have landed: no wait_group,
            // no per-k-block proxy fence (a fence.proxy.async here drains every copy in flight and
            // serialises the ring: measured 1.25 us per k-block regardless of depth)
            cp_async_mbar_arrive_noinc(&full[s]);
        }

        // =========================== epilogue: TMEM lane == output pixel =======================
        mbar_wait(accum, 0);
        tc_fence_after();
        const int oy = p.oy0 + p.osy * my, ox = p.ox0 + p.osx * mx;
        const size_t opix = ((size_t)b * p.Hout + oy) * p.Wout + ox;
        const uint32_t taddr = tmem_base + ((uint32_t)(warp * 32) << 16);
        const int n_base = blockIdx.y * BN;
#pragma unroll 1
        for (int c0 = 0; c0 < BN; c0 += 16) {
            uint32_t v[16];
            tmem_ld16(taddr + c0, v);
            if (!row_ok) continue;
            const int n0 = n_base + c0;
            if (p.mode == 0) {
                if (n0 >= p.Cout) continue;
                float f[16];
#pragma unroll
                for (int j = 0; j < 16; j++) f[j] = fmaf(__uint_as_float(v[j]), __ldg(p.scale + n0 + j), __ldg(p.shift + n0 + j));
                const bool res_late = p.flags & 1;
                if (res_late) {  // activation first (relu field: 0 none, 1 ReLU, 2 exact-erf GELU)
#pragma unroll
                    for (int j = 0; j < 16; j++) f[j] = p.relu == 1 ? fmaxf(f[j], 0.f) : (p.relu == 2 ? gelu_erf(f[j]) : f[j]);
                }
                if (p.res) {
                    const uint4 *rp = reinterpret_cast<const uint4 *>(p.res + opix * p.res_stride + p.res_coff + n0);
                    const uint4 r0 = __ldg(rp), r1 = __ldg(rp + 1);
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
                    if (!res_late) {
                        if (p.relu == 1) { a = fmaxf(a, 0.f); c = fmaxf(c, 0.f); }
                        else if (p.relu == 2) { a = gelu_erf(a); c = gelu_erf(c); }
                    }
                    __nv_bfloat162 h = __floats2bfloat162_rn(a, c);
                    o[j] = *reinterpret_cast<uint32_t *>(&h);
                }
                uint4 *op = reinterpret_cast<uint4 *>(reinterpret_cast<__nv_bfloat16 *>(p.out) + opix * p.out_stride + p.out_coff + n0);
                op[0] = make_uint4(o[0], o[1], o[2], o[3]);
                op[1] = make_uint4(o[4], o[5], o[6], o[7]);
            } else if (p.mode == 2) {
                // VAE output head (musetalk/models/vae.py:102-107): (x / 2 + 0.5).clamp(0, 1) -> * 255 -> round -> u8,
                // channels reversed RGB -> BGR; optional fp32 copy in [0,1], RGB
                if (c0 != 0) continue;
                for (int j = 0; j < p.Cout; j++) {
                    const float a = fmaf(__uint_as_float(v[j]), __ldg(p.scale + j), __ldg(p.shift + j));
                    const float im = fminf(fmaxf(a * 0.5f + 0.5f, 0.f), 1.f);
                    if (p.out_f32) p.out_f32[opix * p.Cout + j] = im;
                    if (p.out) reinterpret_cast<uint8_t *>(p.out)[opix * p.Cout + (p.Cout - 1 - j)] = (uint8_t)rintf(im * 255.f);
                }
            } else {
                // output head: bare conv + bias -> sigmoid (wav2lip.py:83-85) -> x255 truncated to u8
                // (lipreal.py:126 `* 255.`, :209 `astype(np.uint8)`), NHWC
                if (c0 != 0) continue;
                for (int j = 0; j < p.Cout; j++) {
                    const float a = fmaf(__uint_as_float(v[j]), __ldg(p.scale + j), __ldg(p.shift + j));
                    const float sg = 1.0f / (1.0f + __expf(-a));
                    if (p.out_f32) p.out_f32[opix * p.Cout + j] = sg;
                    if (p.out) reinterpret_cast<uint8_t *>(p.out)[opix * p.Cout + j] = (uint8_t)(sg * 255.f);
                }
            }
        }
    } else if (lane == 0) {
        // =========================== MMA issuer ================================================
        constexpr uint32_t idesc = make_idesc(CONV_BM, BN);
        for (int kb = 0; kb < nkb; kb++) {
            const int s = kb % CONV_STAGES;
            mbar_wait(&full[s], (kb / CONV_STAGES) & 1);
            tc_fence_after();
            const uint64_t adesc = make_sdesc(smem_u32(sA + s * A_STAGE_BYTES));
            const uint64_t bdesc = make_sdesc(smem_u32(sB + s * ConvSmem<BN, CONV_STAGES>::B_STAGE));
#pragma unroll
            for (int k = 0; k < CONV_BK / 16; k++)  // +32 B per UMMA_K inside the 128B swizzle atom
                if (!(p.dbg & 4)) umma_f16(tmem_base, adesc + 2 * k, bdesc + 2 * k, idesc, (kb | k) != 0);
            umma_commit(&empty[s]);
        }
        umma_commit(accum);
    }
    __syncwarp();  // re-converge warp 4 (lane 0 ran the issue loop alone) before the CTA barrier
    tc_fence_before();
    __syncthreads();
    if (warp == 4) tmem_dealloc(tmem_base, TMEM_COLS);
}

#include "conv_tma.cuh"
#include "w2v_stack.cuh"

// =====================================================================================================
// host: program loader, launch list, direct / CUDA-graph execution
// =====================================================================================================
struct LaunchDesc {
    void *func;
    dim3 grid;
    int smem;
};

template <int BN, int STAGES>
static cudaError_t conv_desc_s(dim3 grid, LaunchDesc *d) {
    static mf_per_device_flag attr_set;
    int dev = 0;
    cudaGetDevice(&dev);
    if (!attr_set.test_and_set(dev)) {
        cudaError_t e = cudaFuncSetAttribute(k_conv<BN, STAGES>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                             ConvSmem<BN, STAGES>::TOTAL);
        if (e != cudaSuccess) return e;
    }
    d->func = (void *)k_conv<BN, STAGES>;
    d->grid = grid;
    d->smem = ConvSmem<BN, STAGES>::TOTAL;
    return cudaSuccess;
}

template <int BN>
static cudaError_t conv_desc(const ConvParams &p, LaunchDesc *d) {
    dim3 grid((p.M + CONV_BM - 1) / CONV_BM, (p.Cout + BN - 1) / BN);
    constexpr int DEEP = BN <= 32 ? 10 : (BN == 64 ? 8 : 6);  // 200 / 192 / 192 KB
    const bool deep = grid.x * grid.y <= 148 && p.nkb > 4;
    return deep ? conv_desc_s<BN, DEEP>(grid, d) : conv_desc_s<BN, 4>(grid, d);
}

static cudaError_t conv_desc_any(int BN, const ConvParams &p, LaunchDesc *d) {
    switch (BN) {
        case 16: return conv_desc<16>(p, d);
        case 32: return conv_desc<32>(p, d);
        case 64: return conv_desc<64>(p, d);
        default: return conv_desc<128>(p, d);
    }
}

// one kernel launch with its by-value parameter struct; io != 0 marks the launches that see caller pointers
enum { IO_NONE = 0, IO_IN0 = 1, IO_IN1 = 2, IO_OUT = 3, IO_WH_FRAMES = 4, IO_WH_FINISH = 5, IO_WH_OUT = 6, IO_W2V_IN = 7 };
struct Launch {
    void *func = nullptr;
    dim3 grid, block;
    int smem = 0;
    int io = IO_NONE;
    int cluster = 1;  // thread-block cluster size (k_conv_tma<2>: CTA pairs)
    int op = -1;  // program op index (conv ops: profiling hook), -1 for helper kernels
    size_t ws_bytes = 0;  // k_conv_tma split-K: workspace / counter requirement (patched in by finish_launch_list)
    int ws_counters = 0;
    int par = 0;          // != 0: branch id inside a parallel region (W2LOp.flags bits 8..15): a maximal run of tagged launches is a region,
                          // equal ids run in order, different ids may run side by side
    std::vector<unsigned char> params;
    template <class T>
    void set(const T &p) { params.assign((const unsigned char *)&p, (const unsigned char *)&p + sizeof(T)); }
    template <class T>
    T &as() { return *reinterpret_cast<T *>(params.data()); }
};

typedef CUresult (*PFN_encodeTiled)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *,
                                    const cuuint64_t *, const cuuint32_t *, const cuuint32_t *, CUtensorMapInterleave,
                                    CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static PFN_encodeTiled get_encode() {
    static PFN_encodeTiled fn = nullptr;
    if (!fn) {
        void *p = nullptr;
        cudaDriverEntryPointQueryResult q;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess &&
            q == cudaDriverEntryPointSuccess)
            fn = (PFN_encodeTiled)p;
    }
    return fn;
}

struct Wav2LipState {
    W2LHeader hdr;
    std::vector<W2LBuffer> bufs;
    std::vector<W2LOp> ops;
    std::vector<ConvParams> params;  // conv ops: batch-independent fields filled at load
    std::vector<__nv_bfloat16 *> dbuf;
    const unsigned char *blob = nullptr;
    std::vector<unsigned char> *entry_table = nullptr;  // host copy of the blob's entry table
    std::vector<const mf_blob_entry *> ent_of_op_scale, ent_of_op_shift;
    float *gn_coef = nullptr;       // GroupNorm scratch shared by all GN ops (ops run in order): [max_batch][GN_MAX_C][2]
    float *gn_partial = nullptr;    // per-CTA partial sums [<= GN_MAX_CTAS + max_batch][64][2]
    unsigned *gn_counter = nullptr; // [max_batch], zero between launches
    bool has_gn = false;
    float *scores = nullptr;  // attention scratch: fp32 scores
    __nv_bfloat16 *probs = nullptr;
    size_t score_elems = 0;
    int max_batch = 0;
    int last_launches = 0;
    bool profile = false;
    cudaEvent_t ev[2] = {nullptr, nullptr};
    int profile_op = -1;
    // cached launch list + instantiated CUDA graph per batch size: a forward is ONE cudaGraphLaunch; only the nodes
    // that see caller pointers are re-parameterised when those change
    struct Plan {
        int B = 0;
        int n_samples = 0, T = 0;  // whisper plans
        float *ws = nullptr;       // split-K workspace + tile counters of this plan's k_conv_tma launches
        unsigned *counters = nullptr;
        std::vector<Launch> launches;
        cudaGraph_t graph = nullptr;
        cudaGraphExec_t exec = nullptr;
        std::vector<cudaGraphNode_t> nodes;
        const void *in0 = nullptr, *in1 = nullptr;
        void *out_u8 = nullptr;
        float *out_f32 = nullptr;
    };
    std::vector<Plan *> plans;
    bool use_graph = true;
    bool use_pdl = true;
    cudaStream_t capture_stream = nullptr;
    std::vector<cudaStream_t> par_streams;   // fork-join branches of independent launches inside the captured graph
    std::vector<cudaEvent_t> par_events;     // [0] fork, [1 + j] end of branch j
    // Whisper program (hdr.mel_w == -2): log-mel scratch, embedding buffers to gather, filterbank
    float *wh_logspec = nullptr;
    int *wh_maxslot = nullptr;
    const float *wh_filters = nullptr;
    std::vector<int> wh_embed_bufs;
    // wav2vec2 program (hdr.mel_w == -3): fixed window length, first conv layer in fp32, waveform statistics scratch
    int w2v_samples = 0, w2v_frames = 0, w2v_vocab = 0, w2v_c0 = 0, w2v_k0 = 0, w2v_s0 = 0;
    const float *w2v_conv0 = nullptr;   // [C0][k0] weights then [C0] bias, fp32
    float *w2v_stats = nullptr;         // (mean, rstd) of the window
    float *w2v_xres = nullptr;          // transformer stack (op kind 5) scratch: fp32 residual stream, qkv / attention out / FFN hidden, barrier
    __nv_bfloat16 *w2v_qkv = nullptr, *w2v_ao = nullptr, *w2v_hid = nullptr;
    unsigned *w2v_barrier = nullptr;
    int w2v_cluster = 1;                // cluster size of k_w2v_stack: the largest of 8 / 4 / 2 / 1 with all WS_G CTAs co-resident
    const unsigned char *w2v_image = nullptr;
    // GroupNorm statistics fused into the producing conv (k_conv_tma epilogue): per conv op its consumer GN op (or -1), per GN op
    // its producer conv (or -1), the per-conv slot buffers, and the slot count chosen while the current launch list is built
    std::vector<int> gn_consumer, gn_producer, gn_fused_slots;
    std::vector<float *> gn_fused_buf;
    std::vector<size_t> gn_fused_cap;
    float *dbg_ws = nullptr;        // workspace of the ad-hoc launch lists of mf_convnet_debug_run
    unsigned *dbg_counters = nullptr;
    size_t dbg_ws_bytes = 0;
    int dbg_n_counters = 0;
    PFN_encodeTiled encode = nullptr;
    int sm_count = 148;
};

// program kinds (hdr.mel_w): >= 0 wav2lip, -1 musetalk, -2 whisper encoder, -3 wav2vec2 CTC
static bool is_musetalk(const Wav2LipState *s) { return s->hdr.mel_w == -1; }
static bool is_whisper(const Wav2LipState *s) { return s->hdr.mel_w == -2; }
static bool is_wav2vec2(const Wav2LipState *s) { return s->hdr.mel_w == -3; }

void wav2lip_destroy(mf_ctx *ctx) {
    Wav2LipState *s = ctx->wav2lip;
    if (!s) return;
    for (auto p : s->dbuf) cudaFree(p);
    cudaFree(s->w2v_xres); cudaFree(s->w2v_qkv); cudaFree(s->w2v_ao); cudaFree(s->w2v_hid); cudaFree(s->w2v_barrier);
    for (auto pl : s->plans) {
        if (pl->exec) cudaGraphExecDestroy(pl->exec);
        if (pl->graph) cudaGraphDestroy(pl->graph);
        cudaFree(pl->ws);
        cudaFree(pl->counters);
        delete pl;
    }
    delete s->entry_table;
    if (s->capture_stream) cudaStreamDestroy(s->capture_stream);
    for (auto st : s->par_streams) cudaStreamDestroy(st);
    for (auto ev : s->par_events) cudaEventDestroy(ev);
    for (auto b : s->gn_fused_buf) cudaFree(b);
    cudaFree(s->dbg_ws);
    cudaFree(s->dbg_counters);
    cudaFree(s->w2v_stats);
    cudaFree(s->wh_logspec);
    cudaFree(s->wh_maxslot);
    cudaFree(s->gn_coef);
    cudaFree(s->gn_partial);
    cudaFree(s->gn_counter);
    cudaFree(s->scores);
    cudaFree(s->probs);
    if (s->ev[0]) { cudaEventDestroy(s->ev[0]); cudaEventDestroy(s->ev[1]); }
    delete s;
    ctx->wav2lip = nullptr;
}

static inline float bits_to_float(int32_t v) { float f; memcpy(&f, &v, 4); return f; }

static bool conv_tma_eligible(const W2LOp &o);

extern "C" int mf_wav2lip_load(mf_ctx *ctx, const void *blob, size_t nbytes, int max_batch) {
    if (!ctx) return MF_E_INVALID;
    MF_REQUIRE(ctx, blob && max_batch >= 1 && max_batch <= 256, "mf_wav2lip_load: bad arguments");
    MF_CUDA(ctx, cudaSetDevice(ctx->device));
    wav2lip_destroy(ctx);
    PFN_encodeTiled encode = get_encode();
    if (!encode) return mf_fail(ctx, MF_E_CUDA, "cuTensorMapEncodeTiled entry point not found");
    std::vector<unsigned char> *head = new std::vector<unsigned char>(sizeof(mf_blob_header) + 4096 * sizeof(mf_blob_entry));
    const size_t hbytes = std::min(head->size(), nbytes);
    MF_CUDA(ctx, cudaMemcpy(head->data(), blob, hbytes, cudaMemcpyDeviceToHost));
    const mf_blob_header *h = reinterpret_cast<const mf_blob_header *>(head->data());
    MF_REQUIRE(ctx, hbytes >= sizeof(mf_blob_header) && h->magic == MF_BLOB_MAGIC && h->kind == 2,
               "mf_wav2lip_load: not a conv-net blob");
    MF_REQUIRE(ctx, h->n_entries <= 4096 && sizeof(mf_blob_header) + h->n_entries * sizeof(mf_blob_entry) <= hbytes,
               "mf_wav2lip_load: too many entries");
    const mf_blob_entry *ent = reinterpret_cast<const mf_blob_entry *>(head->data() + sizeof(mf_blob_header));
    const unsigned char *base = reinterpret_cast<const unsigned char *>(blob);
    std::vector<const mf_blob_entry *> by_id;
    for (uint32_t i = 0; i < h->n_entries; i++) {
        if (ent[i].offset + ent[i].nbytes > nbytes) continue;
        if (ent[i].id >= by_id.size()) by_id.resize(ent[i].id + 1, nullptr);
        by_id[ent[i].id] = &ent[i];
    }
    auto find = [&](int32_t id) -> const mf_blob_entry * { return (id >= 0 && (size_t)id < by_id.size()) ? by_id[id] : nullptr; };
    const mf_blob_entry *pe = find(W2L_ID_PROGRAM);
    MF_REQUIRE(ctx, pe && pe->nbytes >= sizeof(W2LHeader), "mf_wav2lip_load: program entry missing");
    std::vector<unsigned char> prog(pe->nbytes);
    MF_CUDA(ctx, cudaMemcpy(prog.data(), base + pe->offset, pe->nbytes, cudaMemcpyDeviceToHost));
    Wav2LipState *s = new (std::nothrow) Wav2LipState();
    MF_REQUIRE(ctx, s, "out of host memory");
    ctx->wav2lip = s;
    s->blob = base;
    s->entry_table = head;
    s->hdr = *reinterpret_cast<const W2LHeader *>(prog.data());
    const size_t need = sizeof(W2LHeader) + (size_t)s->hdr.n_buffers * sizeof(W2LBuffer) + (size_t)s->hdr.n_ops * sizeof(W2LOp);
    MF_REQUIRE(ctx, s->hdr.n_buffers > 0 && s->hdr.n_ops > 0 && need == pe->nbytes, "mf_wav2lip_load: program size mismatch");
    const W2LBuffer *pb = reinterpret_cast<const W2LBuffer *>(prog.data() + sizeof(W2LHeader));
    const W2LOp *po = reinterpret_cast<const W2LOp *>(pb + s->hdr.n_buffers);
    s->bufs.assign(pb, pb + s->hdr.n_buffers);
    s->ops.assign(po, po + s->hdr.n_ops);
    s->max_batch = max_batch;
    s->encode = encode;
    s->sm_count = ctx->sm_count > 0 ? ctx->sm_count : 148;
    { const char *e = getenv("MF_NO_GRAPH"); s->use_graph = !(e && atoi(e)); }
    { const char *e = getenv("MF_PDL"); s->use_pdl = !(e && !atoi(e)); }
    s->dbuf.assign(s->hdr.n_buffers, nullptr);
    for (int i = 0; i < s->hdr.n_buffers; i++) {
        const size_t bytes = (size_t)max_batch * s->bufs[i].H * s->bufs[i].W * s->bufs[i].C * 2;
        MF_CUDA(ctx, cudaMalloc(&s->dbuf[i], bytes));
        MF_CUDA(ctx, cudaMemset(s->dbuf[i], 0, bytes));
        if (s->bufs[i].init_entry > 0) {
            const mf_blob_entry *ie = find(s->bufs[i].init_entry);
            MF_REQUIRE(ctx, ie && ie->nbytes == bytes / max_batch, "buffer %d: init tensor does not match the buffer shape", i);
            for (int b = 0; b < max_batch; b++)
                MF_CUDA(ctx, cudaMemcpy(reinterpret_cast<unsigned char *>(s->dbuf[i]) + (size_t)b * ie->nbytes, base + ie->offset,
                                        ie->nbytes, cudaMemcpyDeviceToDevice));
        }
    }
    s->params.resize(s->hdr.n_ops);
    s->ent_of_op_scale.assign(s->hdr.n_ops, nullptr);
    s->ent_of_op_shift.assign(s->hdr.n_ops, nullptr);
    auto okbuf = [&](int b) { return b >= 0 && b < s->hdr.n_buffers; };
    for (int i = 0; i < s->hdr.n_ops; i++) {
        const W2LOp &o = s->ops[i];
        if (o.kind != 0) {
            MF_REQUIRE(ctx, o.kind >= 1 && o.kind <= 5, "op %d: unknown kind %d", i, o.kind);
            MF_REQUIRE(ctx, ((o.flags >> 8) & 0xff) == 0, "op %d: only conv ops can be part of a parallel region", i);
            MF_REQUIRE(ctx, okbuf(o.in_buf) && okbuf(o.out_buf), "op %d: bad buffer id", i);
            if (o.kind == 5) {
                const int D = o.Cin, I = o.Mw, heads = o.ntaps, layers = o.Mh;
                const W2LBuffer &ib = s->bufs[o.in_buf], &ob = s->bufs[o.out_buf];
                const int T = ib.H * ib.W;
                const mf_blob_entry *we = find(o.w_entry);
                auto slice = [](int N) { int per = (N + WS_G - 1) / WS_G; return (per + 7) / 8 * 8; };
                MF_REQUIRE(ctx, D > 0 && I > 0 && heads > 0 && layers > 0 && ib.C == D && ob.C == D && ob.H * ob.W == T && D % 64 == 0 &&
                                    (I <= WS_KA ? I % 64 == 0 : I % WS_KC == 0) && D <= WS_KA && D % heads == 0 && slice(3 * D) <= WS_MAX_NC && slice(I) <= WS_MAX_NC &&
                                    (size_t)slice(3 * D) * (D + 8) <= WS_WBUF_HALFS && (size_t)slice(I) * (D + 8) <= WS_WBUF_HALFS &&
                                    (size_t)slice(D) * (I + 8) <= WS_WBUF_HALFS && (size_t)max_batch * T <= WS_MAX_MT * 16 &&
                                    (size_t)(3 * T * (D / heads + 1) + T * T) * 4 <= sizeof(((W2vSmem *)nullptr)->A),
                           "op %d: transformer stack outside what k_w2v_stack implements (one window, D <= 1024)", i);
                MF_REQUIRE(ctx, we && we->nbytes == (size_t)layers * w2v_layer_bytes(D, I), "op %d: transformer weight image does not match the geometry", i);
                const size_t M = (size_t)max_batch * T;
                MF_REQUIRE(ctx, !s->w2v_xres, "only one transformer stack per program");
                MF_CUDA(ctx, cudaMalloc(&s->w2v_xres, M * D * 4));
                MF_CUDA(ctx, cudaMalloc(&s->w2v_qkv, M * 3 * D * 2));
                MF_CUDA(ctx, cudaMalloc(&s->w2v_ao, M * D * 2));
                MF_CUDA(ctx, cudaMalloc(&s->w2v_hid, M * I * 2));
                MF_CUDA(ctx, cudaMalloc(&s->w2v_barrier, 64 + 24 * 8));
                MF_CUDA(ctx, cudaMemset(s->w2v_barrier, 0, 64 + 24 * 8));
                s->w2v_image = base + we->offset;
                MF_CUDA(ctx, cudaFuncSetAttribute(k_w2v_stack, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(W2vSmem)));
                {   // the software grid barrier needs every CTA resident: pick the largest cluster size the device can co-schedule WS_G / size times
                    s->w2v_cluster = 1;
                    const char *force = getenv("MF_W2V_CLUSTER");
                    for (int cs = 8; cs >= 2; cs >>= 1) {
                        if (force && atoi(force) != cs) continue;
                        cudaLaunchConfig_t cfg;
                        memset(&cfg, 0, sizeof(cfg));
                        cfg.gridDim = dim3(WS_G); cfg.blockDim = dim3(WS_THREADS_); cfg.dynamicSmemBytes = sizeof(W2vSmem);
                        cudaLaunchAttribute at[1];
                        memset(at, 0, sizeof(at));
                        at[0].id = cudaLaunchAttributeClusterDimension;
                        at[0].val.clusterDim.x = (unsigned)cs; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
                        cfg.attrs = at; cfg.numAttrs = 1;
                        int n_clusters = 0;
                        if (cudaOccupancyMaxActiveClusters(&n_clusters, k_w2v_stack, &cfg) == cudaSuccess && n_clusters * cs >= WS_G) { s->w2v_cluster = cs; break; }
                        cudaGetLastError();
                    }
                }
                continue;
            }
            if (o.kind == 1 || o.kind == 2) {
                const mf_blob_entry *se = find(o.scale_entry), *he = find(o.shift_entry);
                MF_REQUIRE(ctx, se && he && se->nbytes == (size_t)o.Cin * 4 && he->nbytes == (size_t)o.Cin * 4 &&
                                    o.in_coff % 8 == 0 && o.in_coff + o.Cin <= s->bufs[o.in_buf].C &&
                                    (o.kind == 1 || s->bufs[o.in_buf].C == o.Cin) && s->bufs[o.out_buf].C == o.Cin && o.Cin % 8 == 0,
                           "op %d: norm parameters do not match the buffers", i);
                s->ent_of_op_scale[i] = se;
                s->ent_of_op_shift[i] = he;
                if (o.kind == 1) {
                    MF_REQUIRE(ctx, o.ntaps >= 1 && o.ntaps <= 64 && o.Cin % o.ntaps == 0 && o.Cin <= GN_MAX_C, "op %d: bad group count", i);
                    s->has_gn = true;
                }
            } else if (o.kind == 3) {
                MF_REQUIRE(ctx, okbuf(o.res_buf) && okbuf(o.Mh) && o.Cin % 8 == 0 && o.ntaps >= 1, "op %d: bad attention op", i);
                const size_t nq = (size_t)s->bufs[o.in_buf].H * s->bufs[o.in_buf].W;
                const size_t nk = (size_t)s->bufs[o.res_buf].H * s->bufs[o.res_buf].W;
                s->score_elems = std::max(s->score_elems, (size_t)max_batch * o.ntaps * nq * ((nk + 63) / 64 * 64));  // (unfused path / MF_FLASH=0)
            } else {
                MF_REQUIRE(ctx, s->bufs[o.in_buf].C == 2 * o.Cin && s->bufs[o.out_buf].C == o.Cin && o.Cin % 8 == 0, "op %d: bad GEGLU op", i);
            }
            continue;
        }
        ConvParams &p = s->params[i];
        memset(&p, 0, sizeof(p));
        MF_REQUIRE(ctx, okbuf(o.in_buf) && (o.mode != 0 || okbuf(o.out_buf)) && (o.res_buf < 0 || okbuf(o.res_buf)),
                   "op %d: bad buffer id", i);
        MF_REQUIRE(ctx, o.BN == 16 || o.BN == 32 || o.BN == 64 || o.BN == 128, "op %d: BN %d unsupported", i, o.BN);
        MF_REQUIRE(ctx, o.ntaps >= 1 && o.ntaps <= CONV_MAX_TAPS && o.Cin % 8 == 0 && o.Kpad % CONV_BK == 0 &&
                            o.Kpad >= o.ntaps * o.Cin && o.Cout_pad % o.BN == 0 && o.Cout <= o.Cout_pad && o.ups >= 0 && o.ups <= 2,
                   "op %d: bad geometry", i);
        if (const int par = (o.flags >> 8) & 0xff) {
            // the packer's promise for a parallel region, checked: this op's output channels are touched by no op of ANOTHER branch of
            // the region, and it touches no output of theirs
            MF_REQUIRE(ctx, o.mode == 0, "op %d: an output-head op cannot be part of a parallel region", i);
            for (int j = i - 1; j >= 0 && s->ops[j].kind == 0 && ((s->ops[j].flags >> 8) & 0xff) != 0; j--) {
                const W2LOp &q = s->ops[j];
                if (((q.flags >> 8) & 0xff) == par) continue;
                auto overlap = [](int b0, int c0, int n0, int b1, int c1, int n1) { return b0 == b1 && b0 >= 0 && c0 < c1 + n1 && c1 < c0 + n0; };
                const bool bad = overlap(o.out_buf, o.out_coff, o.Cout, q.out_buf, q.out_coff, q.Cout) ||
                                 overlap(o.out_buf, o.out_coff, o.Cout, q.in_buf, q.in_coff, q.Cin) ||
                                 overlap(o.out_buf, o.out_coff, o.Cout, q.res_buf, q.res_coff, q.Cout) ||
                                 overlap(q.out_buf, q.out_coff, q.Cout, o.in_buf, o.in_coff, o.Cin) ||
                                 overlap(q.out_buf, q.out_coff, q.Cout, o.res_buf, o.res_coff, o.Cout);
                MF_REQUIRE(ctx, !bad, "ops %d and %d are in different branches of a parallel region but touch each other's output channels", j, i);
            }
        }
        const mf_blob_entry *we = find(o.w_entry), *se = find(o.scale_entry), *he = find(o.shift_entry);
        MF_REQUIRE(ctx, we && se && he && we->nbytes == (size_t)o.Cout_pad * o.Kpad * 2 &&
                            se->nbytes == (size_t)o.Cout_pad * 4 && he->nbytes == (size_t)o.Cout_pad * 4,
                   "op %d: weight/scale/shift entries do not match the geometry (strict loader)", i);
        const W2LBuffer &ib = s->bufs[o.in_buf];
        MF_REQUIRE(ctx, o.in_coff % 8 == 0 && ib.C % 8 == 0 && o.in_coff + o.Cin <= ib.C, "op %d: input channels out of range", i);
        p.in = s->dbuf[o.in_buf];
        p.in_stride = ib.C; p.in_coff = o.in_coff; p.Hin = ib.H; p.Win = ib.W; p.ups = o.ups;
        if (o.mode == 0) {
            const W2LBuffer &ob = s->bufs[o.out_buf];
            MF_REQUIRE(ctx, o.out_coff % 8 == 0 && ob.C % 8 == 0 && o.out_coff + o.Cout <= ob.C && o.Cout % 16 == 0,
                       "op %d: output channels out of range", i);
            p.out = s->dbuf[o.out_buf];
            p.out_stride = ob.C; p.out_coff = o.out_coff; p.Hout = ob.H; p.Wout = ob.W;
        } else {
            if (o.mode == 3) {   // fp32 token output (wav2vec2 logits): [Mh * Mw tokens][Cout]
                MF_REQUIRE(ctx, o.Cout <= 256 && conv_tma_eligible(o), "op %d: fp32 output head must be a TMA-eligible Linear with Cout <= 256", i);
                p.Hout = o.Mh; p.Wout = o.Mw;
            } else {
                MF_REQUIRE(ctx, o.Cout <= 16 && o.BN == 16 && (o.mode == 1 || o.mode == 2), "op %d: output head must have Cout <= 16", i);
                p.Hout = s->hdr.out_hw; p.Wout = s->hdr.out_hw;
            }
        }
        MF_REQUIRE(ctx, o.oy0 + o.osy * (o.Mh - 1) < p.Hout && o.ox0 + o.osx * (o.Mw - 1) < p.Wout, "op %d: output grid out of range", i);
        if (o.res_buf >= 0) {
            const W2LBuffer &rb = s->bufs[o.res_buf];
            MF_REQUIRE(ctx, rb.H == p.Hout && rb.W == p.Wout && o.res_coff % 8 == 0 && o.res_coff + o.Cout <= rb.C,
                       "op %d: residual shape mismatch", i);
            p.res = s->dbuf[o.res_buf];
            p.res_stride = rb.C; p.res_coff = o.res_coff;
        }
        p.scale = reinterpret_cast<const float *>(base + se->offset);
        p.shift = reinterpret_cast<const float *>(base + he->offset);
        p.Mh = o.Mh; p.Mw = o.Mw; p.oy0 = o.oy0; p.ox0 = o.ox0; p.osy = o.osy; p.osx = o.osx; p.isy = o.isy; p.isx = o.isx;
        p.ntaps = o.ntaps; p.Cin = o.Cin; p.nkb = o.Kpad / CONV_BK; p.Cout = o.Cout; p.relu = o.relu; p.mode = o.mode; p.flags = o.flags;
        memcpy(p.tap_dy, o.tap_dy, CONV_MAX_TAPS);
        memcpy(p.tap_dx, o.tap_dx, CONV_MAX_TAPS);
        // weights re-tiled [Cout_pad / 16][Kpad / 64][16][64] bf16: TMA box = {64 k, 16 rows, 1 k-block, BN / 16 row groups}
        cuuint64_t dims[4] = {CONV_BK, 16, (cuuint64_t)(o.Kpad / CONV_BK), (cuuint64_t)(o.Cout_pad / 16)};
        cuuint64_t strides[3] = {CONV_BK * 2, 16 * CONV_BK * 2, (cuuint64_t)(o.Kpad / CONV_BK) * 16 * CONV_BK * 2};
        cuuint32_t box[4] = {CONV_BK, 16, 1, (cuuint32_t)(o.BN / 16)};
        cuuint32_t estr[4] = {1, 1, 1, 1};
        CUresult cr = encode(&p.wmap, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 4, (void *)(base + we->offset), dims, strides, box,
                             estr, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B,
                             CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
        if (cr != CUDA_SUCCESS) return mf_fail(ctx, MF_E_CUDA, "op %d: cuTensorMapEncodeTiled failed (%d)", i, (int)cr);
    }
    // ---- GroupNorm <- producing conv pairs (statistics fused into the conv epilogue)
    s->gn_consumer.assign(s->hdr.n_ops, -1);
    s->gn_producer.assign(s->hdr.n_ops, -1);
    s->gn_fused_slots.assign(s->hdr.n_ops, 0);
    s->gn_fused_buf.assign(s->hdr.n_ops, nullptr);
    s->gn_fused_cap.assign(s->hdr.n_ops, 0);
    {
        static int fuse = -1;
        if (fuse < 0) { const char *e = getenv("MF_GN_FUSE"); fuse = e ? atoi(e) : 1; }
        for (int i = 0; fuse && i < s->hdr.n_ops; i++) {
            const W2LOp &g = s->ops[i];
            if (g.kind != 1 || g.in_coff != 0 || s->bufs[g.in_buf].C != g.Cin) continue;
            const int cpg = g.Cin / g.ntaps;
            if (cpg != 4 && cpg != 8 && cpg != 16) continue;
            for (int j = i - 1; j >= 0; j--) {
                const W2LOp &c = s->ops[j];
                const bool writes = (c.kind == 0 && c.mode == 0 && c.out_buf == g.in_buf) || (c.kind != 0 && c.out_buf == g.in_buf);
                if (!writes) continue;
                const W2LBuffer &ob = s->bufs[g.in_buf];
                if (c.kind == 0 && conv_tma_eligible(c) && c.out_coff == 0 && c.Cout == g.Cin && c.Mh == ob.H && c.Mw == ob.W && c.osy == 1 &&
                    c.osx == 1 && c.oy0 == 0 && c.ox0 == 0 && (s->gn_consumer[j] < 0)) {
                    s->gn_consumer[j] = i;
                    s->gn_producer[i] = j;
                }
                break;   // the last writer decides
            }
        }
    }
    if (is_whisper(s)) {
        const mf_blob_entry *ae = find(W2L_ID_AUX);
        MF_REQUIRE(ctx, ae && ae->nbytes >= 12 && ae->nbytes % 4 == 0, "whisper program: aux entry missing");
        std::vector<int32_t> aux(ae->nbytes / 4);
        MF_CUDA(ctx, cudaMemcpy(aux.data(), base + ae->offset, ae->nbytes, cudaMemcpyDeviceToHost));
        const int n = aux[0];
        MF_REQUIRE(ctx, n >= 1 && n <= 8 && (size_t)n + 2 <= aux.size(), "whisper program: bad aux entry");
        const W2LBuffer &mb = s->bufs[s->hdr.in_face_buf];
        MF_REQUIRE(ctx, mb.C == WH_MELS && mb.W == 1, "whisper program: input buffer must be [frames,1,80]");
        for (int i = 0; i < n; i++) {
            MF_REQUIRE(ctx, okbuf(aux[1 + i]) && s->bufs[aux[1 + i]].W == 1 && s->bufs[aux[1 + i]].C == s->bufs[aux[1]].C,
                       "whisper program: bad embedding buffer");
            s->wh_embed_bufs.push_back(aux[1 + i]);
        }
        const mf_blob_entry *fe = find(aux[1 + n]);
        MF_REQUIRE(ctx, fe && fe->nbytes == (size_t)WH_MELS * WH_BINS * 4, "whisper program: mel filterbank entry missing");
        s->wh_filters = reinterpret_cast<const float *>(base + fe->offset);
        MF_CUDA(ctx, cudaMalloc(&s->wh_logspec, (size_t)mb.H * WH_MELS * sizeof(float)));
        MF_CUDA(ctx, cudaMalloc(&s->wh_maxslot, sizeof(int)));
        MF_CUDA(ctx, cudaMemset(s->wh_maxslot, 0, sizeof(int)));
    }
    if (is_wav2vec2(s)) {
        const mf_blob_entry *ae = find(W2L_ID_AUX);
        MF_REQUIRE(ctx, ae && ae->nbytes == 7 * 4, "wav2vec2 program: aux entry missing");
        int32_t aux[7];
        MF_CUDA(ctx, cudaMemcpy(aux, base + ae->offset, sizeof(aux), cudaMemcpyDeviceToHost));
        s->w2v_samples = aux[0]; s->w2v_frames = aux[1]; s->w2v_vocab = aux[2]; s->w2v_c0 = aux[4]; s->w2v_k0 = aux[5]; s->w2v_s0 = aux[6];
        const mf_blob_entry *ce = find(aux[3]);
        const W2LBuffer &b0 = s->bufs[s->hdr.in_face_buf];
        MF_REQUIRE(ctx, ce && ce->nbytes == (size_t)(s->w2v_c0 * s->w2v_k0 + s->w2v_c0) * 4 && b0.C == s->w2v_c0 && b0.W == 1 &&
                            b0.H == (s->w2v_samples - s->w2v_k0) / s->w2v_s0 + 1 && s->w2v_c0 <= 1024 && s->w2v_k0 >= 1 && s->w2v_k0 <= 16 && s->w2v_s0 >= 1 && s->w2v_s0 <= 16 &&
                            s->ops.back().kind == 0 && s->ops.back().mode == 3 && s->ops.back().Cout == s->w2v_vocab,
                   "wav2vec2 program: first-layer entry / buffers do not match the aux record");
        s->w2v_conv0 = reinterpret_cast<const float *>(base + ce->offset);
        MF_CUDA(ctx, cudaMalloc(&s->w2v_stats, (size_t)max_batch * 2 * sizeof(float)));
    }
    if (s->has_gn) {
        MF_CUDA(ctx, cudaMalloc(&s->gn_coef, (size_t)max_batch * GN_MAX_C * 2 * sizeof(float)));
        MF_CUDA(ctx, cudaMalloc(&s->gn_partial, (size_t)(GN_MAX_CTAS + 2 * max_batch) * 128 * sizeof(float)));
        MF_CUDA(ctx, cudaMalloc(&s->gn_counter, (size_t)max_batch * sizeof(unsigned)));
        MF_CUDA(ctx, cudaMemset(s->gn_counter, 0, (size_t)max_batch * sizeof(unsigned)));
    }
    if (s->score_elems) {
        MF_CUDA(ctx, cudaMalloc(&s->scores, s->score_elems * sizeof(float)));
        MF_CUDA(ctx, cudaMalloc(&s->probs, s->score_elems * sizeof(__nv_bfloat16)));
    }
    MF_CUDA(ctx, cudaDeviceSynchronize());
    return MF_OK;
}


// ---- k_conv_tma planning: tile box, BN, split-K, ring depth -------------------------------------------
static void *conv_tma_func(int cg, int act) {
    static void *const tab[2][3] = {{(void *)k_conv_tma<1, 0>, (void *)k_conv_tma<1, 1>, (void *)k_conv_tma<1, 2>},
                                    {(void *)k_conv_tma<2, 0>, (void *)k_conv_tma<2, 1>, (void *)k_conv_tma<2, 2>}};
    return tab[cg - 1][act];
}
static bool is_conv_tma(const void *f) {
    for (int c = 1; c <= 2; c++)
        for (int a = 0; a < 3; a++)
            if (f == conv_tma_func(c, a)) return true;
    return false;
}

static bool conv_tma_eligible(const W2LOp &o) {
    static int on = -1;
    if (on < 0) { const char *e = getenv("MF_CONV_TMA"); on = e ? atoi(e) : 1; }
    return on && o.kind == 0 && o.isy == 1 && o.isx == 1 && o.ups == 0 && o.Cin % CONV_BK == 0 &&
           o.Kpad == o.ntaps * o.Cin;
}

static int add_conv_tma(mf_ctx *ctx, Wav2LipState *s, int i, int B, std::vector<Launch> &L) {
    const W2LOp &o = s->ops[i];
    const ConvParams &cp = s->params[i];
    ConvTmaParams p;
    memset(&p, 0, sizeof(p));
    // tile box TW x TH x TB = 128 output pixels: fewest tiles, ties to the widest rows
    int best_tiles = 1 << 30, lTW = 0, lTH = 0;
    for (int a = 7; a >= 0; a--)
        for (int b = 7 - a; b >= 0; b--) {
            const int TW = 1 << a, TH = 1 << b, TB = 1 << (7 - a - b);
            if (TW > 256 || TH > 256 || TB > 256) continue;
            const int t = ((o.Mw + TW - 1) / TW) * ((o.Mh + TH - 1) / TH) * ((B + TB - 1) / TB);
            if (t < best_tiles) { best_tiles = t; lTW = a; lTH = b; }
        }
    const int TW = 1 << lTW, TH = 1 << lTH, TB = 1 << (7 - lTW - lTH);
    p.lTW = lTW; p.lTH = lTH;
    p.tiles_x = (o.Mw + TW - 1) / TW; p.tiles_y = (o.Mh + TH - 1) / TH;
    const int m_tiles = best_tiles;
    // tap groups for the row-halo A reuse: runs of taps with the same dy and consecutive dx (3x3: three runs of 3; upsample parity
    // convs: two runs of 2), only for single-row tiles (TW = 128); otherwise every tap is its own group (ndx = 1)
    int ndx = 1, n_groups = o.ntaps;
    {
        static int halo = -1;
        if (halo < 0) { const char *e = getenv("MF_CONV_HALO"); halo = e ? atoi(e) : 1; }
        int run = 1;
        while (run < o.ntaps && o.tap_dy[run] == o.tap_dy[0] && o.tap_dx[run] == o.tap_dx[run - 1] + 1) run++;
        bool ok = halo && lTW == 7 && run >= 2 && run <= 3 && o.ntaps % run == 0;
        for (int t = 0; ok && t < o.ntaps; t++)
            if (t % run != 0 && (o.tap_dy[t] != o.tap_dy[t - 1] || o.tap_dx[t] != o.tap_dx[t - 1] + 1)) ok = false;
        if (ok) { ndx = run; n_groups = o.ntaps / run; }
    }
    const int cblocks = o.Cin / CONV_BK;
    const int nkb = n_groups * cblocks;   // k-steps: (tap group, channel block), each 4 * ndx UMMAs
    // BN / split-K: minimise a simple cost model (cycles): per k-block max(MMA, smem traffic), per item a pipeline fill, per split
    // a reduction term.  Sweeps over forced (BN, S) (profiles/r01_conv_planner_sweep.log) show it within ~20 % of the best
    // measured choice on the small-M layers; a fitted per-SM ingest / split-latency model did worse and was dropped.
    const int sms = s->sm_count;
    double best = 1e30;
    int BN = 0, S = 1;
    for (int nt = (o.Cout + 255) / 256; nt <= std::max(1, o.Cout / 32); nt++) {
        const int bn = (((o.Cout + nt - 1) / nt) + 15) / 16 * 16;
        if (bn > 256) continue;
        {   // the choice must leave room for the ring: with the row-halo groups a stage holds ndx B tiles (a 256-channel layer
            // with fewer than one tile per SM -- no CTA pair to halve the B tile -- would need 113 KB per stage: take a narrower BN)
            const int cg_c = (m_tiles >= sms && bn % 32 == 0 && o.mode == 0) ? 2 : 1;
            const int stage_c = (ndx > 1 ? 17 * 1024 : A_STAGE_BYTES) + ndx * (bn * 128 / cg_c);
            if ((CT_SMEM_LIMIT - 1024 - 256 - 4096) / stage_c < (ndx > 1 ? 3 : 2)) continue;
        }
        const double kb_cost = ndx * std::max(2.0 * bn, 1.5 * (128 + bn));
        static const int splits[] = {1, 2, 3, 4, 6, 8};
        for (int sp : splits) {
            if (sp > 1 && (nkb / sp < 6 || o.mode != 0)) break;
            const long items = (long)m_tiles * nt * sp;
            const long waves = (items + sms - 1) / sms;
            const double per_item = (double)((nkb + sp - 1) / sp) * kb_cost + 2500.0 + (sp > 1 ? 1500.0 + 12.0 * bn * sp : 0.0);
            const double cost = (double)waves * per_item;
            if (cost < best * 0.999) { best = cost; BN = bn; S = sp; }
        }
    }
    MF_REQUIRE(ctx, BN >= 16, "op %d: no BN for the TMA conv", i);
    {   // experiments (scripts/bench_conv.py): MF_CONV_FORCE="BN,S" overrides the cost model, MF_CONV_VERBOSE=1 prints the choice
        const char *e = getenv("MF_CONV_FORCE");
        int fbn = 0, fs = 0;
        if (e && sscanf(e, "%d,%d", &fbn, &fs) == 2 && fbn >= 16 && fbn <= 256 && fbn % 16 == 0 && fs >= 1 && fs <= 8 && nkb / fs >= 1 &&
            (o.mode == 0 || fs == 1)) { BN = fbn; S = fs; }
        const char *v = getenv("MF_CONV_VERBOSE");
        if (v && atoi(v))
            fprintf(stderr, "[conv_tma] op %d Cin %d Cout %d taps %d M %dx%dx%d: tile %dx%dx%d, m_tiles %d, BN %d, splits %d, k-steps %d x %d taps\n", i, o.Cin,
                    o.Cout, o.ntaps, B, o.Mh, o.Mw, TW, TH, TB, m_tiles, BN, S, nkb, ndx);
    }
    // BN of a CTA pair must split into two halves of whole 16-row weight pieces
    // CTA pairs (cta_group::2) for the layers with enough 128-row tiles to fill the SM pairs: halves the per-CTA B traffic
    int CG = (m_tiles >= sms && BN % 32 == 0 && o.mode == 0) ? 2 : 1;
    { const char *e = getenv("MF_CONV_CG"); if (e && atoi(e) == 1) CG = 1; if (e && atoi(e) == 2 && BN % 32 == 0 && m_tiles >= 2) CG = 2; }
    p.BN = BN; p.n_tiles = (o.Cout + BN - 1) / BN; p.splits = S;
    p.ndx = ndx; p.n_groups = n_groups;
    p.a_stage_bytes = ndx > 1 ? 17 * 1024 : A_STAGE_BYTES;   // 128 + ndx - 1 pixel rows of 128 B, rounded up to the 1024-B swizzle period
    const int stage = p.a_stage_bytes + ndx * (BN * 128 / CG);
    p.stages = std::min(CT_MAX_STAGES, (CT_SMEM_LIMIT - 1024 - 256 - 4096) / stage);
    MF_REQUIRE(ctx, p.stages >= 2, "op %d: a pipeline stage of %d bytes does not fit twice", i, stage);
    {   // multiply-high constants for the epilogue's tile-coordinate divisions (conv_tma.cuh ct_fastdiv)
        const uint64_t dx = (uint64_t)p.tiles_x, dxy = (uint64_t)p.tiles_x * p.tiles_y;
        MF_REQUIRE(ctx, ((uint64_t)m_tiles + 2) * dxy < (1ull << 32), "op %d: too many tiles for the 32-bit tile index arithmetic", i);
        p.mg_tx = dx > 1 ? (uint32_t)(((1ull << 32) + dx - 1) / dx) : 0u;
        p.mg_txy = dxy > 1 ? (uint32_t)(((1ull << 32) + dxy - 1) / dxy) : 0u;
    }
    const int m_groups = (m_tiles + CG - 1) / CG;          // work items are 128-row tiles (CG 1) or 256-row tile pairs (CG 2)
    p.total_items = m_groups * p.n_tiles * S;
    p.out = cp.out; p.res = cp.res; p.scale = cp.scale; p.shift = cp.shift;
    p.out_stride = cp.out_stride; p.out_coff = cp.out_coff; p.Hout = cp.Hout; p.Wout = cp.Wout;
    p.res_stride = cp.res_stride; p.res_coff = cp.res_coff;
    p.Mh = o.Mh; p.Mw = o.Mw; p.B = B; p.oy0 = o.oy0; p.ox0 = o.ox0; p.osy = o.osy; p.osx = o.osx;
    p.in_coff = o.in_coff; p.ntaps = o.ntaps; p.cblocks = cblocks; p.nkb = nkb; p.Cout = o.Cout; p.relu = o.relu; p.flags = o.flags;
    for (int g = 0; g < n_groups; g++) { p.tap_dy[g] = o.tap_dy[g * ndx]; p.tap_dx[g] = o.tap_dx[g * ndx]; p.grp_tap0[g] = (int8_t)(g * ndx); }
    { const char *e = getenv("MF_CONV_DBG"); p.dbg = e ? atoi(e) : 0; }
    // activation map: NHWC bf16 buffer [max_batch][H][W][C]
    const W2LBuffer &ib = s->bufs[o.in_buf];
    {
        cuuint64_t dims[4] = {(cuuint64_t)ib.C, (cuuint64_t)ib.W, (cuuint64_t)ib.H, (cuuint64_t)s->max_batch};
        cuuint64_t strides[3] = {(cuuint64_t)ib.C * 2, (cuuint64_t)ib.W * ib.C * 2, (cuuint64_t)ib.H * ib.W * ib.C * 2};
        cuuint32_t box[4] = {CONV_BK, (cuuint32_t)(TW + ndx - 1), (cuuint32_t)TH, (cuuint32_t)TB};
        cuuint32_t estr[4] = {1, 1, 1, 1};
        CUresult cr = s->encode(&p.amap, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 4, (void *)s->dbuf[o.in_buf], dims, strides, box, estr,
                                CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                                CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
        if (cr != CUDA_SUCCESS) return mf_fail(ctx, MF_E_CUDA, "op %d: activation cuTensorMapEncodeTiled failed (%d)", i, (int)cr);
    }
    {
        const mf_blob_entry *we = nullptr;
        const mf_blob_entry *ent = reinterpret_cast<const mf_blob_entry *>(s->entry_table->data() + sizeof(mf_blob_header));
        const mf_blob_header *h = reinterpret_cast<const mf_blob_header *>(s->entry_table->data());
        for (uint32_t e = 0; e < h->n_entries; e++) if ((int32_t)ent[e].id == o.w_entry) we = &ent[e];
        MF_REQUIRE(ctx, we, "op %d: weight entry missing", i);
        cuuint64_t dims[4] = {CONV_BK, 16, (cuuint64_t)(o.Kpad / CONV_BK), (cuuint64_t)(o.Cout_pad / 16)};
        cuuint64_t strides[3] = {CONV_BK * 2, 16 * CONV_BK * 2, (cuuint64_t)(o.Kpad / CONV_BK) * 16 * CONV_BK * 2};
        cuuint32_t box[4] = {CONV_BK, 16, 1, (cuuint32_t)(BN / 16 / CG)};
        cuuint32_t estr[4] = {1, 1, 1, 1};
        CUresult cr = s->encode(&p.wmap, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 4, (void *)(s->blob + we->offset), dims, strides, box, estr,
                                CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                                CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
        if (cr != CUDA_SUCCESS) return mf_fail(ctx, MF_E_CUDA, "op %d: weight cuTensorMapEncodeTiled failed (%d)", i, (int)cr);
    }
    // fused GroupNorm statistics for the consumer GN op (single-image tiles, no split-K)
    s->gn_fused_slots[i] = 0;
    if (s->gn_consumer[i] >= 0 && S == 1 && TB == 1 && o.mode == 0) {
        const W2LOp &g = s->ops[s->gn_consumer[i]];
        const int tiles_img = p.tiles_x * p.tiles_y, slots = tiles_img * 4;
        const size_t need = (size_t)s->max_batch * slots * g.ntaps * 2 * sizeof(float);
        if (!s->gn_fused_buf[i]) {
            // capacity for any tile decomposition of this layer (ragged shapes can need up to ~2x the minimal tile count)
            const size_t cap = (size_t)s->max_batch * (2 * ((size_t)o.Mh * o.Mw + 127) / 128 + 4) * 4 * g.ntaps * 2 * sizeof(float);
            MF_CUDA(ctx, cudaMalloc(&s->gn_fused_buf[i], cap));
            s->gn_fused_cap[i] = cap;
        }
        if (need <= s->gn_fused_cap[i]) {
            p.gn_partial = s->gn_fused_buf[i]; p.gn_cpg = g.Cin / g.ntaps; p.gn_G = g.ntaps;
            s->gn_fused_slots[i] = slots;
        }
    }
    static mf_per_device_flag attr_set;
    if (!attr_set.test_and_set(ctx->device)) {
        for (int f = 0; f < 6; f++)
            MF_CUDA(ctx, cudaFuncSetAttribute(conv_tma_func(f / 3 + 1, f % 3), cudaFuncAttributeMaxDynamicSharedMemorySize, CT_SMEM_LIMIT));
    }
    MF_REQUIRE(ctx, o.relu >= 0 && o.relu <= 2, "op %d: unknown activation %d", i, o.relu);
    Launch l;
    l.func = conv_tma_func(CG, o.relu);
    l.cluster = CG;
    l.grid = dim3(CG * std::min(p.total_items, sms / CG)); l.block = dim3(CT_THREADS);
    p.mode = o.mode;
    l.smem = p.stages * stage + 256 + 4096 + 1024; l.op = i; l.io = o.mode != 0 ? IO_OUT : IO_NONE;
    if (S > 1) {
        l.ws_bytes = (size_t)m_groups * CG * p.n_tiles * S * 128 * BN * sizeof(float);
        l.ws_counters = m_groups * CG * p.n_tiles;
    }
    l.set(p);
    L.push_back(std::move(l));
    return MF_OK;
}

// allocate (or grow) the split-K workspace for a finished launch list and patch it into the k_conv_tma launches
static int finish_launch_list(mf_ctx *ctx, std::vector<Launch> &L, float **ws, unsigned **counters, size_t *ws_bytes, int *n_counters) {
    // launches run one after the other and share the workspace from offset 0 -- except inside a parallel region (Launch::par), whose
    // branches may execute concurrently: there every BRANCH gets its own region (its launches run in order and share it)
    size_t need = 0;
    int nc = 0;
    std::vector<size_t> ws_off(L.size(), 0);
    std::vector<int> c_off(L.size(), 0);
    for (size_t a = 0; a < L.size();) {
        size_t b = a + 1;
        if (L[a].par != 0) while (b < L.size() && L[b].par != 0) b++;
        size_t w = 0;
        int c = 0;
        if (L[a].par == 0) {
            w = L[a].ws_bytes; c = L[a].ws_counters;
        } else {
            std::vector<int> tags;
            for (size_t j = a; j < b; j++) if (std::find(tags.begin(), tags.end(), L[j].par) == tags.end()) tags.push_back(L[j].par);
            for (int t : tags) {
                size_t bw = 0;
                int bc = 0;
                for (size_t j = a; j < b; j++) if (L[j].par == t) { bw = std::max(bw, L[j].ws_bytes); bc = std::max(bc, L[j].ws_counters); }
                for (size_t j = a; j < b; j++) if (L[j].par == t) { ws_off[j] = w; c_off[j] = c; }
                w += (bw + 255) / 256 * 256; c += bc;
            }
        }
        need = std::max(need, w); nc = std::max(nc, c);
        a = b;
    }
    if (need > *ws_bytes) {
        MF_CUDA(ctx, cudaDeviceSynchronize());
        cudaFree(*ws);
        *ws = nullptr;
        MF_CUDA(ctx, cudaMalloc(ws, need));
        *ws_bytes = need;
    }
    if (nc > *n_counters) {
        MF_CUDA(ctx, cudaDeviceSynchronize());
        cudaFree(*counters);
        *counters = nullptr;
        MF_CUDA(ctx, cudaMalloc(counters, (size_t)nc * sizeof(unsigned)));
        MF_CUDA(ctx, cudaMemset(*counters, 0, (size_t)nc * sizeof(unsigned)));
        *n_counters = nc;
    }
    for (size_t j = 0; j < L.size(); j++)
        if (is_conv_tma(L[j].func)) {
            L[j].as<ConvTmaParams>().ws = reinterpret_cast<float *>(reinterpret_cast<unsigned char *>(*ws) + ws_off[j]);
            L[j].as<ConvTmaParams>().counters = *counters + c_off[j];
        }
    return MF_OK;
}

// ---- launch list -----------------------------------------------------------------------------------
static int add_op_launches_(mf_ctx *ctx, Wav2LipState *s, int i, int B, std::vector<Launch> &L);
static int add_op_launches(mf_ctx *ctx, Wav2LipState *s, int i, int B, std::vector<Launch> &L) {
    const size_t n0 = L.size();
    const int rc = add_op_launches_(ctx, s, i, B, L);
    const int par = (s->ops[i].flags >> 8) & 0xff;
    if (rc == MF_OK && par) for (size_t j = n0; j < L.size(); j++) L[j].par = par;
    return rc;
}
static int add_op_launches_(mf_ctx *ctx, Wav2LipState *s, int i, int B, std::vector<Launch> &L) {
    const W2LOp &o = s->ops[i];
    if (conv_tma_eligible(o)) return add_conv_tma(ctx, s, i, B, L);
    if (o.kind == 0) {
        ConvParams p = s->params[i];
        p.M = B * p.Mh * p.Mw;
        {
            const char *e = getenv("MF_CONV_DBG");  // timing experiments only (scripts/bench_conv.py)
            p.dbg = e ? atoi(e) : 0;
        }
        LaunchDesc d;
        MF_CUDA(ctx, conv_desc_any(o.BN, p, &d));
        Launch l;
        l.func = d.func; l.grid = d.grid; l.block = dim3(CONV_THREADS); l.smem = d.smem; l.op = i;
        l.io = p.mode != 0 ? IO_OUT : IO_NONE;
        l.set(p);
        L.push_back(std::move(l));
        return MF_OK;
    }
    const W2LBuffer &ib = s->bufs[o.in_buf];
    if (o.kind == 1 || o.kind == 2) {
        NormParams n;
        n.in = s->dbuf[o.in_buf]; n.out = s->dbuf[o.out_buf];
        n.gamma = reinterpret_cast<const float *>(s->blob + s->ent_of_op_scale[i]->offset);
        n.beta = reinterpret_cast<const float *>(s->blob + s->ent_of_op_shift[i]->offset);
        n.C = o.Cin; n.G = o.ntaps; n.silu = o.relu; n.eps = bits_to_float(o.Kpad);
        n.coef = s->gn_coef; n.partial = s->gn_partial; n.counter = s->gn_counter;
        n.pix_per_cta = 0; n.in_stride = ib.C; n.in_coff = o.in_coff;
        if (o.kind == 1) {
            n.npix = ib.H * ib.W;
            const size_t slab = (size_t)n.npix * n.C * 2;
            const int cpg = n.C / n.G;
            if (slab <= 200 * 1024 && cpg % 8 == 0 && n.C / 8 <= GN_SMALL_THREADS) {
                static mf_per_device_flag attr_set;
                if (!attr_set.test_and_set(ctx->device))
                    MF_CUDA(ctx, cudaFuncSetAttribute(k_gn_small, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
                Launch a;
                a.func = (void *)k_gn_small; a.grid = dim3(B); a.block = dim3(GN_SMALL_THREADS); a.op = i;
                a.smem = (int)(slab + GN_SMALL_THREADS * 8 + 64 * 8);
                a.set(n);
                L.push_back(std::move(a));
                return MF_OK;
            }
            const int target = std::max(1, GN_MAX_CTAS / B);
            n.pix_per_cta = std::max(8, (n.npix + target - 1) / target);
            const int prod = s->gn_producer[i];
            if (prod >= 0 && s->gn_fused_slots[prod] > 0) {
                // the producing conv already left per-tile statistics: add them up instead of re-reading the tensor
                GnFinalParams f;
                f.partial = s->gn_fused_buf[prod]; f.gamma = n.gamma; f.beta = n.beta; f.coef = n.coef;
                f.slots = s->gn_fused_slots[prod]; f.C = n.C; f.G = n.G; f.npix = n.npix; f.eps = n.eps;
                Launch a;
                a.func = (void *)k_gn_finalize; a.grid = dim3(B); a.block = dim3(GN_FIN_THREADS); a.op = i;
                a.set(f);
                L.push_back(std::move(a));
            } else {
                Launch a;
                a.func = (void *)k_gn_stats; a.grid = dim3((n.npix + n.pix_per_cta - 1) / n.pix_per_cta, B); a.block = dim3(256); a.op = i;
                a.set(n);
                L.push_back(std::move(a));
            }
            Launch b;
            b.func = (void *)k_gn_apply; b.grid = dim3((n.npix + n.pix_per_cta - 1) / n.pix_per_cta, B); b.block = dim3(256); b.op = i;
            b.set(n);
            L.push_back(std::move(b));
        } else {
            n.npix = B * ib.H * ib.W;
            Launch a;
            a.func = (void *)k_layernorm; a.grid = dim3((n.npix + 7) / 8); a.block = dim3(256); a.op = i;
            a.set(n);
            L.push_back(std::move(a));
        }
        return MF_OK;
    }
    if (o.kind == 5) {
        Launch z;
        ZeroParams zp;
        zp.p = s->w2v_barrier; zp.n = 16;
        z.func = (void *)k_zero_u32; z.grid = dim3(1); z.block = dim3(32); z.op = i;
        z.set(zp);
        L.push_back(std::move(z));
        W2vStackParams w;
        w.image = s->w2v_image;
        w.x_in = s->dbuf[o.in_buf]; w.x_out = s->dbuf[o.out_buf];
        w.xres = s->w2v_xres; w.qkv = s->w2v_qkv; w.ao = s->w2v_ao; w.hid = s->w2v_hid; w.barrier = s->w2v_barrier;
        w.T = ib.H * ib.W; w.B = B; w.M = B * w.T; w.D = o.Cin; w.I = o.Mw; w.heads = o.ntaps; w.layers = o.Mh;
        w.eps = bits_to_float(o.Kpad);
        w.scale_log2 = 1.4426950408889634f / sqrtf((float)(o.Cin / o.ntaps));
        Launch a;
        a.func = (void *)k_w2v_stack; a.grid = dim3(WS_G); a.block = dim3(WS_THREADS_); a.smem = (int)sizeof(W2vSmem); a.op = i;
        a.cluster = s->w2v_cluster;   // LayerNorm rows are shared inside a cluster (w2v_ln_rows)
        a.set(w);
        L.push_back(std::move(a));
        return MF_OK;
    }
    if (o.kind == 4) {
        GegluParams g;
        g.in = s->dbuf[o.in_buf]; g.out = s->dbuf[o.out_buf]; g.tokens = (size_t)B * ib.H * ib.W; g.Hd = o.Cin;
        Launch a;
        a.func = (void *)k_geglu; a.grid = dim3((unsigned)((g.tokens * (g.Hd / 8) + 255) / 256)); a.block = dim3(256); a.op = i;
        a.set(g);
        L.push_back(std::move(a));
        return MF_OK;
    }
    // attention: fused (k_flash) for head dims up to 160; otherwise scores = Q K^T -> softmax -> P V through HBM
    {
        const W2LBuffer &kb0 = s->bufs[o.res_buf], &vb0 = s->bufs[o.Mh], &ob0 = s->bufs[o.out_buf];
        const int dh0 = o.Cin;
        static int fused = -1;
        if (fused < 0) { const char *e = getenv("MF_FLASH"); fused = e ? atoi(e) : 1; }
        if (fused && dh0 % 8 == 0 && dh0 <= 160) {
            FlashParams f;
            f.Q = s->dbuf[o.in_buf] + o.in_coff; f.K = s->dbuf[o.res_buf] + o.res_coff; f.V = s->dbuf[o.Mh] + o.Mw;
            f.O = s->dbuf[o.out_buf] + o.out_coff;
            f.nq = ib.H * ib.W; f.nk = kb0.H * kb0.W; f.dh = dh0; f.heads = o.ntaps;
            f.ldq = ib.C; f.ldk = kb0.C; f.ldv = vb0.C; f.ldo = ob0.C;
            f.q_bs = (long long)f.nq * ib.C; f.k_bs = (long long)f.nk * kb0.C; f.v_bs = (long long)f.nk * vb0.C; f.o_bs = (long long)f.nq * ob0.C;
            f.scale_log2 = bits_to_float(o.Kpad) * 1.4426950408889634f;
            Launch a;
            a.func = dh0 <= 48 ? (void *)k_flash<48> : dh0 <= 64 ? (void *)k_flash<64> : dh0 <= 80 ? (void *)k_flash<80> : (void *)k_flash<160>;
            a.grid = dim3((f.nq + 63) / 64, o.ntaps, B); a.block = dim3(128); a.op = i;
            a.set(f);
            L.push_back(std::move(a));
            return MF_OK;
        }
    }
    const W2LBuffer &kb = s->bufs[o.res_buf], &vb = s->bufs[o.Mh], &ob = s->bufs[o.out_buf];
    const int nq = ib.H * ib.W, nk = kb.H * kb.W, heads = o.ntaps, dh = o.Cin, ld = (nk + 63) / 64 * 64;
    GemmParams g1;
    g1.A = s->dbuf[o.in_buf] + o.in_coff; g1.lda = ib.C; g1.a_bs = (long long)nq * ib.C; g1.a_hs = dh;
    g1.B = s->dbuf[o.res_buf] + o.res_coff; g1.ldb = kb.C; g1.b_bs = (long long)nk * kb.C; g1.b_hs = dh;
    g1.C = s->scores; g1.ldc = ld; g1.c_bs = (long long)heads * nq * ld; g1.c_hs = (long long)nq * ld;
    g1.M = nq; g1.N = nk; g1.K = dh; g1.heads = heads;
    Launch a;
    a.func = (void *)k_bgemm<false, false>; a.grid = dim3((nk + 63) / 64, (nq + 63) / 64, B * heads); a.block = dim3(128); a.op = i;
    a.set(g1);
    L.push_back(std::move(a));
    SoftmaxParams sp;
    sp.S = s->scores; sp.P = s->probs; sp.rows = (size_t)B * heads * nq; sp.n = nk; sp.ld = ld; sp.scale = bits_to_float(o.Kpad);
    Launch b;
    b.func = (void *)k_softmax; b.grid = dim3((unsigned)((sp.rows + 7) / 8)); b.block = dim3(256); b.op = i;
    b.set(sp);
    L.push_back(std::move(b));
    GemmParams g2;
    g2.A = s->probs; g2.lda = ld; g2.a_bs = (long long)heads * nq * ld; g2.a_hs = (long long)nq * ld;
    g2.B = s->dbuf[o.Mh] + o.Mw; g2.ldb = vb.C; g2.b_bs = (long long)nk * vb.C; g2.b_hs = dh;
    g2.C = s->dbuf[o.out_buf] + o.out_coff; g2.ldc = ob.C; g2.c_bs = (long long)nq * ob.C; g2.c_hs = dh;
    g2.M = nq; g2.N = dh; g2.K = nk; g2.heads = heads;
    Launch c;
    c.func = (void *)k_bgemm<true, true>; c.grid = dim3((dh + 63) / 64, (nq + 63) / 64, B * heads); c.block = dim3(128); c.op = i;
    c.set(g2);
    L.push_back(std::move(c));
    return MF_OK;
}

// program inputs: wav2lip (in_face_buf >= 0): in0 = faces u8, in1 = mel fp32;  musetalk (mel_h == 0 marks it):
// in0 = latents fp16 NCHW [B,C,H,W], in1 = whisper fp16 [B,T,D]

static int build_plan(mf_ctx *ctx, Wav2LipState *s, Wav2LipState::Plan *pl, int B) {
    std::vector<Launch> &L = pl->launches;
    pl->B = B;
    if (is_whisper(s)) {
        const W2LBuffer &mb = s->bufs[s->hdr.in_face_buf];
        WhisperPrep w;
        w.audio = nullptr; w.logspec = s->wh_logspec; w.filters = s->wh_filters; w.maxslot = s->wh_maxslot;
        w.melbuf = s->dbuf[s->hdr.in_face_buf]; w.n_samples = 0; w.n_frames = 0; w.n_ctx_frames = mb.H;
        Launch l0, l1;
        l0.func = (void *)k_logmel_frames; l0.grid = dim3(1); l0.block = dim3(256); l0.io = IO_WH_FRAMES; l0.set(w);
        l1.func = (void *)k_logmel_finish; l1.grid = dim3((mb.H * WH_MELS + 255) / 256); l1.block = dim3(256); l1.io = IO_WH_FINISH; l1.set(w);
        L.push_back(std::move(l0));
        L.push_back(std::move(l1));
        for (int i = 0; i < s->hdr.n_ops; i++) {
            int rc = add_op_launches(ctx, s, i, B, L);
            if (rc) return rc;
        }
        GatherParams g;
        memset(&g, 0, sizeof(g));
        g.n_src = (int)s->wh_embed_bufs.size();
        for (int i = 0; i < g.n_src; i++) g.src[i] = s->dbuf[s->wh_embed_bufs[i]];
        g.C = s->bufs[s->wh_embed_bufs[0]].C; g.maxslot = s->wh_maxslot;
        Launch lg;
        lg.func = (void *)k_whisper_gather; lg.grid = dim3(1); lg.block = dim3(256); lg.io = IO_WH_OUT; lg.set(g);
        L.push_back(std::move(lg));
        return MF_OK;
    }
    if (is_wav2vec2(s)) {
        W2vPrep w;
        w.audio = nullptr; w.conv0 = s->w2v_conv0; w.stats = s->w2v_stats; w.out = s->dbuf[s->hdr.in_face_buf];
        w.n_samples = s->w2v_samples; w.n_frames = s->bufs[s->hdr.in_face_buf].H; w.C0 = s->w2v_c0; w.k0 = s->w2v_k0; w.s0 = s->w2v_s0;
        Launch l0, l1;
        l0.func = (void *)k_w2v_stats; l0.grid = dim3(1, B); l0.block = dim3(1024); l0.io = IO_W2V_IN; l0.set(w);
        l1.func = (void *)k_w2v_conv0; l1.grid = dim3((w.n_frames + 7) / 8, B); l1.block = dim3(256); l1.io = IO_W2V_IN; l1.set(w);
        L.push_back(std::move(l0));
        L.push_back(std::move(l1));
        for (int i = 0; i < s->hdr.n_ops; i++) {
            int rc = add_op_launches(ctx, s, i, B, L);
            if (rc) return rc;
        }
        return MF_OK;
    }
    const W2LBuffer &b0 = s->bufs[s->hdr.in_face_buf], &b1 = s->bufs[s->hdr.in_mel_buf];
    {
        PrepParams p0, p1;
        Launch l0, l1;
        if (!is_musetalk(s)) {
            p0 = {nullptr, s->dbuf[s->hdr.in_face_buf], B, s->hdr.face_hw, 0, 0};
            l0.func = (void *)k_prep_face; l0.grid = dim3((B * b0.H * b0.W + 255) / 256);
            p1 = {nullptr, s->dbuf[s->hdr.in_mel_buf], B * b1.H * b1.W, 0, 0, 0};
            l1.func = (void *)k_prep_mel; l1.grid = dim3((B * b1.H * b1.W + 255) / 256);
        } else {
            p0 = {nullptr, s->dbuf[s->hdr.in_face_buf], B, s->hdr.face_hw /* real latent channels */, b0.H * b0.W, b0.C};
            l0.func = (void *)k_prep_latents; l0.grid = dim3((B * b0.H * b0.W * b0.C + 255) / 256);
            p1 = {nullptr, s->dbuf[s->hdr.in_mel_buf], B, b1.H * b1.W, b1.C, 0};
            l1.func = (void *)k_prep_ctx; l1.grid = dim3((B * b1.H * b1.W * b1.C + 255) / 256);
        }
        l0.block = l1.block = dim3(256);
        l0.io = IO_IN0; l1.io = IO_IN1;
        l0.set(p0); l1.set(p1);
        L.push_back(std::move(l0));
        L.push_back(std::move(l1));
    }
    for (int i = 0; i < s->hdr.n_ops; i++) {
        int rc = add_op_launches(ctx, s, i, B, L);
        if (rc) return rc;
    }
    return MF_OK;
}

static void patch_io(Wav2LipState::Plan *pl, const void *in0, const void *in1, void *out_u8, float *out_f32) {
    for (auto &l : pl->launches) {
        if (l.io == IO_IN0) l.as<PrepParams>().src = in0;
        else if (l.io == IO_IN1) l.as<PrepParams>().src = in1;
        else if (l.io == IO_OUT && is_conv_tma(l.func)) { l.as<ConvTmaParams>().out = out_u8; l.as<ConvTmaParams>().out_f32 = out_f32; }
        else if (l.io == IO_OUT) { l.as<ConvParams>().out = out_u8; l.as<ConvParams>().out_f32 = out_f32; }
        else if (l.io == IO_W2V_IN) l.as<W2vPrep>().audio = reinterpret_cast<const float *>(in0);
        else if (l.io == IO_WH_FRAMES || l.io == IO_WH_FINISH) {
            WhisperPrep &w = l.as<WhisperPrep>();
            w.audio = reinterpret_cast<const float *>(in0); w.n_samples = pl->n_samples; w.n_frames = pl->n_samples / WH_HOP;
            if (l.io == IO_WH_FRAMES) l.grid = dim3(w.n_frames);
        } else if (l.io == IO_WH_OUT) {
            GatherParams &g = l.as<GatherParams>();
            g.out = out_f32; g.T = pl->T;
            l.grid = dim3((g.T * g.n_src * g.C + 255) / 256);
        }
    }
    pl->in0 = in0; pl->in1 = in1; pl->out_u8 = out_u8; pl->out_f32 = out_f32;
}

#define MF_PAR_BRANCHES 16
static int launch_one(mf_ctx *ctx, Launch &l, cudaStream_t st, bool pdl_attr) {
    void *args[] = {l.params.data()};
    if (pdl_attr || l.cluster > 1) {
        // programmatic dependent launch: this kernel's CTA-local prologue may overlap the predecessor's tail; every executor
        // kernel calls griddepcontrol.wait before its first global access.  CTA pairs are launched as clusters of 2.
        cudaLaunchConfig_t cfg;
        memset(&cfg, 0, sizeof(cfg));
        cfg.gridDim = l.grid; cfg.blockDim = l.block; cfg.dynamicSmemBytes = (size_t)l.smem; cfg.stream = st;
        cudaLaunchAttribute at[2];
        memset(at, 0, sizeof(at));
        int na = 0;
        if (pdl_attr) {
            at[na].id = cudaLaunchAttributeProgrammaticStreamSerialization;
            at[na].val.programmaticStreamSerializationAllowed = 1;
            na++;
        }
        if (l.cluster > 1) {
            at[na].id = cudaLaunchAttributeClusterDimension;
            at[na].val.clusterDim.x = (unsigned)l.cluster; at[na].val.clusterDim.y = 1; at[na].val.clusterDim.z = 1;
            na++;
        }
        cfg.attrs = at; cfg.numAttrs = (unsigned)na;
        MF_CUDA(ctx, cudaLaunchKernelExC(&cfg, l.func, args));
    } else {
        MF_CUDA(ctx, cudaLaunchKernel(l.func, l.grid, l.block, args, l.smem, st));
    }
    return MF_OK;
}

// `capturing`: st is a capturing stream (the graph path).  A parallel region (maximal run of launches with Launch::par != 0) then becomes a
// fork-join with one branch per tag: the 16 groups of wav2vec2's positional conv (16 launches of 16 CTAs, 22 us apiece one after the
// other), Wav2Lip's audio encoder beside its face encoder.  Launched directly (profiling, MF_NO_GRAPH) the region simply runs in order.
static int launch_direct(mf_ctx *ctx, Wav2LipState *s, std::vector<Launch> &L, cudaStream_t st, bool pdl = false, bool capturing = false) {
    bool started = false, first = true;
    for (size_t a = 0; a < L.size();) {
        size_t b = a + 1;
        if (L[a].par != 0) while (b < L.size() && L[b].par != 0) b++;
        std::vector<int> tags;
        if (L[a].par != 0)
            for (size_t j = a; j < b; j++) if (std::find(tags.begin(), tags.end(), L[j].par) == tags.end()) tags.push_back(L[j].par);
        if (capturing && tags.size() >= 2 && tags.size() <= MF_PAR_BRANCHES) {
            MF_REQUIRE(ctx, !s->par_streams.empty(), "fork-join streams were not created before the capture");
            const int nb = (int)tags.size();
            MF_CUDA(ctx, cudaEventRecord(s->par_events[0], st));
            for (int j = 0; j < nb; j++) MF_CUDA(ctx, cudaStreamWaitEvent(s->par_streams[j], s->par_events[0], 0));
            std::vector<char> first_in_branch(nb, 1);
            for (size_t j = a; j < b; j++) {
                const int br = (int)(std::find(tags.begin(), tags.end(), L[j].par) - tags.begin());
                const int rc = launch_one(ctx, L[j], s->par_streams[br], pdl && !first_in_branch[br]);
                if (rc) return rc;
                first_in_branch[br] = 0;
            }
            for (int j = 0; j < nb; j++) {
                MF_CUDA(ctx, cudaEventRecord(s->par_events[1 + j], s->par_streams[j]));
                MF_CUDA(ctx, cudaStreamWaitEvent(st, s->par_events[1 + j], 0));
            }
            first = true;   // the launch after the join has several predecessors: a plain (fully serialised) launch
            a = b;
            continue;
        }
        for (size_t j = a; j < b; j++) {
            Launch &l = L[j];
            // profiling: events around ALL launches of the profiled op (an op may expand to several kernels)
            const bool prof = s->profile && l.op >= 0 && l.op == s->profile_op;
            if (prof && !started) { cudaEventRecord(s->ev[0], st); started = true; }
            const int rc = launch_one(ctx, l, st, pdl && !first);
            if (rc) return rc;
            first = false;
            if (prof) cudaEventRecord(s->ev[1], st);
        }
        a = b;
    }
    return MF_OK;
}

static int forward_common(mf_ctx *ctx, Wav2LipState *s, const void *in0, const void *in1, void *out_u8, float *out_f32,
                          int B, cudaStream_t st, int n_samples = 0, int T = 0) {
    Wav2LipState::Plan *pl = nullptr;
    for (auto q : s->plans) if (q->B == B) pl = q;
    if (!pl) {
        pl = new Wav2LipState::Plan();
        s->plans.push_back(pl);
        int rc = build_plan(ctx, s, pl, B);
        if (rc) return rc;
        size_t wsb = 0;
        int ncnt = 0;
        rc = finish_launch_list(ctx, pl->launches, &pl->ws, &pl->counters, &wsb, &ncnt);
        if (rc) return rc;
    }
    const bool io_changed = pl->in0 != in0 || pl->in1 != in1 || pl->out_u8 != out_u8 || pl->out_f32 != out_f32 ||
                            pl->n_samples != n_samples || pl->T != T;
    pl->n_samples = n_samples; pl->T = T;
    if (io_changed) patch_io(pl, in0, in1, out_u8, out_f32);
    s->last_launches = (int)pl->launches.size();
    if (!s->use_graph || s->profile) return launch_direct(ctx, s, pl->launches, st);
    if (!pl->exec || io_changed) {
        // (re)build the graph by capturing the launch list on a private stream: capture keeps the programmatic-dependency
        // edges of the PDL launches.  Caller pointers change rarely (the plugins reuse their staging buffers).
        if (pl->exec) { cudaGraphExecDestroy(pl->exec); pl->exec = nullptr; }
        if (pl->graph) { cudaGraphDestroy(pl->graph); pl->graph = nullptr; }
        if (!s->capture_stream) MF_CUDA(ctx, cudaStreamCreateWithFlags(&s->capture_stream, cudaStreamNonBlocking));
        if (s->par_streams.empty()) {   // created outside the capture
            bool any = false;
            for (auto &l : pl->launches) any = any || l.par != 0;
            if (any) {
                s->par_streams.resize(MF_PAR_BRANCHES);
                s->par_events.resize(MF_PAR_BRANCHES + 1);
                for (auto &x : s->par_streams) MF_CUDA(ctx, cudaStreamCreateWithFlags(&x, cudaStreamNonBlocking));
                for (auto &e : s->par_events) MF_CUDA(ctx, cudaEventCreateWithFlags(&e, cudaEventDisableTiming));
            }
        }
        MF_CUDA(ctx, cudaStreamBeginCapture(s->capture_stream, cudaStreamCaptureModeThreadLocal));
        int rc = launch_direct(ctx, s, pl->launches, s->capture_stream, s->use_pdl, true);
        cudaError_t ce = cudaStreamEndCapture(s->capture_stream, &pl->graph);
        if (rc) { if (pl->graph) { cudaGraphDestroy(pl->graph); pl->graph = nullptr; } return rc; }
        MF_CUDA(ctx, ce);
        MF_CUDA(ctx, cudaGraphInstantiate(&pl->exec, pl->graph, 0));
    }
    MF_CUDA(ctx, cudaGraphLaunch(pl->exec, st));
    return MF_OK;
}

extern "C" int mf_wav2lip_forward(mf_ctx *ctx, const float *mel, const uint8_t *faces, uint8_t *out_u8, float *out_f32,
                                  int B, void *stream) {
    if (!ctx) return MF_E_INVALID;
    Wav2LipState *s = ctx->wav2lip;
    if (!s) return mf_fail(ctx, MF_E_STATE, "mf_wav2lip_forward: weights not loaded");
    MF_REQUIRE(ctx, mel && faces && (out_u8 || out_f32), "mf_wav2lip_forward: null pointer");
    MF_REQUIRE(ctx, B >= 1 && B <= s->max_batch, "mf_wav2lip_forward: batch %d outside [1, %d]", B, s->max_batch);
    MF_REQUIRE(ctx, s->hdr.in_face_buf >= 0 && s->hdr.in_mel_buf >= 0 && !is_musetalk(s) && s->ops.back().mode == 1,
               "the loaded program is not a wav2lip program");
    MF_CUDA(ctx, cudaSetDevice(ctx->device));
    return forward_common(ctx, s, faces, mel, out_u8, out_f32, B, (cudaStream_t)stream);
}

extern "C" int mf_musetalk_forward(mf_ctx *ctx, const void *latents_f16, const void *whisper_f16, uint8_t *out_u8,
                                   float *out_f32, int B, void *stream) {
    if (!ctx) return MF_E_INVALID;
    Wav2LipState *s = ctx->wav2lip;
    if (!s) return mf_fail(ctx, MF_E_STATE, "mf_musetalk_forward: weights not loaded");
    MF_REQUIRE(ctx, latents_f16 && whisper_f16 && (out_u8 || out_f32), "mf_musetalk_forward: null pointer");
    MF_REQUIRE(ctx, B >= 1 && B <= s->max_batch, "mf_musetalk_forward: batch %d outside [1, %d]", B, s->max_batch);
    MF_REQUIRE(ctx, s->hdr.in_face_buf >= 0 && s->hdr.in_mel_buf >= 0 && is_musetalk(s) && s->ops.back().mode == 2,
               "the loaded program is not a musetalk program");
    MF_CUDA(ctx, cudaSetDevice(ctx->device));
    return forward_common(ctx, s, latents_f16, whisper_f16, out_u8, out_f32, B, (cudaStream_t)stream);
}

extern "C" int mf_whisper_features(mf_ctx *ctx, const float *audio, int n_samples, float *out_f32, int T, void *stream) {
    if (!ctx) return MF_E_INVALID;
    Wav2LipState *s = ctx->wav2lip;
    if (!s) return mf_fail(ctx, MF_E_STATE, "mf_whisper_features: weights not loaded");
    MF_REQUIRE(ctx, is_whisper(s), "the loaded program is not a whisper program");
    MF_REQUIRE(ctx, audio && out_f32, "mf_whisper_features: null pointer");
    const int n_ctx_frames = s->bufs[s->hdr.in_face_buf].H;
    MF_REQUIRE(ctx, n_samples > WH_NFFT / 2 && n_samples / WH_HOP <= n_ctx_frames,
               "mf_whisper_features: %d samples outside (200, %d] (one 30 s segment per call)", n_samples, n_ctx_frames * WH_HOP);
    MF_REQUIRE(ctx, T >= 1 && T <= s->bufs[s->wh_embed_bufs[0]].H, "mf_whisper_features: T = %d rows out of range", T);
    MF_CUDA(ctx, cudaSetDevice(ctx->device));
    return forward_common(ctx, s, audio, nullptr, nullptr, out_f32, 1, (cudaStream_t)stream, n_samples, T);
}

extern "C" int mf_wav2vec2_logits_batch(mf_ctx *ctx, const float *audio, int n_samples, int B, float *out_f32, void *stream) {
    if (!ctx) return MF_E_INVALID;
    Wav2LipState *s = ctx->wav2lip;
    if (!s) return mf_fail(ctx, MF_E_STATE, "mf_wav2vec2_logits: weights not loaded");
    MF_REQUIRE(ctx, is_wav2vec2(s), "the loaded program is not a wav2vec2 program");
    MF_REQUIRE(ctx, audio && out_f32, "mf_wav2vec2_logits: null pointer");
    MF_REQUIRE(ctx, B >= 1 && B <= s->max_batch, "mf_wav2vec2_logits: batch %d outside [1, %d]", B, s->max_batch);
    MF_REQUIRE(ctx, n_samples == s->w2v_samples, "mf_wav2vec2_logits: the program was packed for windows of %d samples, got %d", s->w2v_samples,
               n_samples);
    MF_CUDA(ctx, cudaSetDevice(ctx->device));
    return forward_common(ctx, s, audio, nullptr, nullptr, out_f32, B, (cudaStream_t)stream);
}

// debug tap: %globaltimer stamps (ns) of CTA 0 at the phase boundaries of layer 1 of the last k_w2v_stack launch
// [P1 start, P1 end, P2 start, P2 end, P3 start, P3 end, P4 start, P4 end, P5 start, P5 end, next layer's start]; synchronises the device
extern "C" int mf_debug_w2v_phase_ns(mf_ctx *ctx, unsigned long long *out, int n) {
    if (!ctx) return MF_E_INVALID;
    Wav2LipState *s = ctx->wav2lip;
    if (!s || !s->w2v_barrier) return mf_fail(ctx, MF_E_STATE, "mf_debug_w2v_phase_ns: no fused transformer stack loaded");
    MF_REQUIRE(ctx, out && n >= 1 && n <= 24, "mf_debug_w2v_phase_ns: bad arguments");
    MF_CUDA(ctx, cudaSetDevice(ctx->device));
    MF_CUDA(ctx, cudaDeviceSynchronize());
    MF_CUDA(ctx, cudaMemcpy(out, s->w2v_barrier + 16, (size_t)n * 8, cudaMemcpyDeviceToHost));
    return MF_OK;
}

extern "C" int mf_wav2vec2_logits(mf_ctx *ctx, const float *audio, int n_samples, float *out_f32, void *stream) {
    return mf_wav2vec2_logits_batch(ctx, audio, n_samples, 1, out_f32, stream);
}

// unit-test entry: run the loaded program on an fp32 NHWC tensor written into buffer `in_buf` (all of its channels)
// and read buffer `out_buf` back as fp32 NHWC (programs without an output head).
extern "C" int mf_convnet_debug_run(mf_ctx *ctx, int in_buf, const float *in_f32, int out_buf, float *out_f32, int B,
                                    void *stream) {
    if (!ctx) return MF_E_INVALID;
    Wav2LipState *s = ctx->wav2lip;
    if (!s) return mf_fail(ctx, MF_E_STATE, "mf_convnet_debug_run: program not loaded");
    MF_REQUIRE(ctx, in_f32 && out_f32 && B >= 1 && B <= s->max_batch && in_buf >= 0 && in_buf < s->hdr.n_buffers &&
                        out_buf >= 0 && out_buf < s->hdr.n_buffers,
               "mf_convnet_debug_run: bad arguments");
    cudaStream_t st = (cudaStream_t)stream;
    const size_t n_in = (size_t)B * s->bufs[in_buf].H * s->bufs[in_buf].W * s->bufs[in_buf].C;
    const size_t n_out = (size_t)B * s->bufs[out_buf].H * s->bufs[out_buf].W * s->bufs[out_buf].C;
    std::vector<Launch> L;
    {
        PrepParams p = {in_f32, s->dbuf[in_buf], (int)(n_in & 0xffffffffu), (int)(n_in >> 32), 0, 0};
        Launch l;
        l.func = (void *)k_f32_to_bf16; l.grid = dim3((unsigned)((n_in + 255) / 256)); l.block = dim3(256);
        l.set(p);
        L.push_back(std::move(l));
    }
    for (int i = 0; i < s->hdr.n_ops; i++) {
        MF_REQUIRE(ctx, s->ops[i].kind != 0 || s->ops[i].mode == 0, "debug run supports programs without an output head");
        int rc = add_op_launches(ctx, s, i, B, L);
        if (rc) return rc;
    }
    {
        PrepParams p = {s->dbuf[out_buf], out_f32, (int)(n_out & 0xffffffffu), (int)(n_out >> 32), 0, 0};
        Launch l;
        l.func = (void *)k_bf16_to_f32; l.grid = dim3((unsigned)((n_out + 255) / 256)); l.block = dim3(256);
        l.set(p);
        L.push_back(std::move(l));
    }
    int rc = finish_launch_list(ctx, L, &s->dbg_ws, &s->dbg_counters, &s->dbg_ws_bytes, &s->dbg_n_counters);
    if (rc) return rc;
    rc = launch_direct(ctx, s, L, st);
    if (rc) return rc;
    s->last_launches = (int)L.size();
    return MF_OK;
}

// second input of a two-input debug program (e.g. the attention context): fp32 NHWC -> buffer, synchronous helper
extern "C" int mf_convnet_debug_set(mf_ctx *ctx, int buf, const float *in_f32, int B, void *stream) {
    if (!ctx) return MF_E_INVALID;
    Wav2LipState *s = ctx->wav2lip;
    if (!s) return mf_fail(ctx, MF_E_STATE, "mf_convnet_debug_set: program not loaded");
    MF_REQUIRE(ctx, in_f32 && B >= 1 && B <= s->max_batch && buf >= 0 && buf < s->hdr.n_buffers, "mf_convnet_debug_set: bad arguments");
    const size_t n = (size_t)B * s->bufs[buf].H * s->bufs[buf].W * s->bufs[buf].C;
    PrepParams p = {in_f32, s->dbuf[buf], (int)(n & 0xffffffffu), (int)(n >> 32), 0, 0};
    k_f32_to_bf16<<<(unsigned)((n + 255) / 256), 256, 0, (cudaStream_t)stream>>>(p);
    MF_CUDA(ctx, cudaGetLastError());
    return MF_OK;
}

extern "C" int mf_wav2lip_last_launches(const mf_ctx *ctx) { return (ctx && ctx->wav2lip) ? ctx->wav2lip->last_launches : 0; }

extern "C" int mf_wav2lip_profile(mf_ctx *ctx, int op_index) {
    if (!ctx) return MF_E_INVALID;
    Wav2LipState *s = ctx->wav2lip;
    if (!s) return mf_fail(ctx, MF_E_STATE, "weights not loaded");
    if (op_index >= 0 && !s->ev[0]) {
        MF_CUDA(ctx, cudaEventCreate(&s->ev[0]));
        MF_CUDA(ctx, cudaEventCreate(&s->ev[1]));
    }
    s->profile = op_index >= 0;
    s->profile_op = op_index;
    return MF_OK;
}

extern "C" int mf_wav2lip_last_op_ms(mf_ctx *ctx, float *ms) {
    if (!ctx) return MF_E_INVALID;
    Wav2LipState *s = ctx->wav2lip;
    if (!s || !s->ev[0] || !ms) return mf_fail(ctx, MF_E_STATE, "profiling not enabled");
    MF_CUDA(ctx, cudaEventSynchronize(s->ev[1]));
    MF_CUDA(ctx, cudaEventElapsedTime(ms, s->ev[0], s->ev[1]));
    return MF_OK;
}
