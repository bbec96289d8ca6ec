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

#include "mf_common.cuh"

#define CONV_BM 128
#define CONV_BK 64
#define CONV_THREADS 160
#define CONV_MAX_TAPS 52
#define A_STAGE_BYTES (CONV_BM * CONV_BK * 2)

// blob entry ids (kind = 2)
enum { W2L_ID_PROGRAM = 1, W2L_ID_FIRST_TENSOR = 16 };

// ---- program records written by the packer (all int32, little endian) ----------------------
struct W2LHeader {
    int32_t n_buffers, n_ops, in_face_buf, in_mel_buf, face_hw, mel_h, mel_w, out_hw;
};
struct W2LBuffer {
    int32_t H, W, C, reserved;
};
struct W2LOp {
    int32_t in_buf, in_coff, out_buf, out_coff, res_buf, res_coff;
    int32_t Mh, Mw, oy0, ox0, osy, osx, isy, isx;
    int32_t ntaps, Cin, Kpad, Cout, Cout_pad, BN, relu, mode;
    int32_t w_entry, scale_entry, shift_entry, reserved;
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
__device__ __forceinline__ void tma_load_2d(void *smem_dst, const CUtensorMap *map, int c0, int c1, uint64_t *bar) {
    asm volatile(
        "cp.async.bulk.tensor.2d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];\n" ::
            "r"(smem_u32(smem_dst)),
        "l"(map), "r"(smem_u32(bar)), "r"(c0), "r"(c1)
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
                tma_load_2d(sB + s * ConvSmem<BN, CONV_STAGES>::B_STAGE, &p.wmap, kb * CONV_BK, blockIdx.y * BN, &full[s]);
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
                        ok = iy >= 0 && iy < p.Hin && ix >= 0 && ix < p.Win;
                        if (ok) src = in_b + (iy * p.Win + ix) * p.in_stride + ch;
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
                    if (p.relu) { a = fmaxf(a, 0.f); c = fmaxf(c, 0.f); }
                    __nv_bfloat162 h = __floats2bfloat162_rn(a, c);
                    o[j] = *reinterpret_cast<uint32_t *>(&h);
                }
                uint4 *op = reinterpret_cast<uint4 *>(reinterpret_cast<__nv_bfloat16 *>(p.out) + opix * p.out_stride + p.out_coff + n0);
                op[0] = make_uint4(o[0], o[1], o[2], o[3]);
                op[1] = make_uint4(o[4], o[5], o[6], o[7]);
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

// ---- input preparation (lipreal.py:108-122) ---------------------------------------------------
// faces u8 [B,S,S,3] BGR -> bf16 [B,S,S,8]: ch 0-2 = face with rows >= S/2 zeroed, ch 3-5 = face, /255
__global__ void k_w2l_prep_face(const uint8_t *__restrict__ faces, __nv_bfloat16 *__restrict__ out, int B, int S) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= B * S * S) return;
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
    reinterpret_cast<uint4 *>(out)[i] = make_uint4(o[0], o[1], o[2], o[3]);
}
// mel fp32 [B,1,80,16] -> bf16 [B,80,16,8] (channel 0)
__global__ void k_w2l_prep_mel(const float *__restrict__ mel, __nv_bfloat16 *__restrict__ out, int n) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    __nv_bfloat162 h = __floats2bfloat162_rn(mel[i], 0.f);
    reinterpret_cast<uint4 *>(out)[i] = make_uint4(*reinterpret_cast<uint32_t *>(&h), 0u, 0u, 0u);
}
// generic bf16 <-> fp32 NHWC converters for the unit-test entry point
__global__ void k_f32_to_bf16(const float *__restrict__ in, __nv_bfloat16 *__restrict__ out, size_t n) {
    const size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x;
    if (i < n) out[i] = __float2bfloat16_rn(in[i]);
}
__global__ void k_bf16_to_f32(const __nv_bfloat16 *__restrict__ in, float *__restrict__ out, size_t n) {
    const size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x;
    if (i < n) out[i] = __bfloat162float(in[i]);
}

// ---- host --------------------------------------------------------------------------------------
struct LaunchDesc {
    void *func;
    dim3 grid;
    int smem;
};

struct Wav2LipState {
    W2LHeader hdr;
    std::vector<W2LBuffer> bufs;
    std::vector<W2LOp> ops;
    std::vector<ConvParams> params;  // per op, batch-independent fields filled at load
    std::vector<__nv_bfloat16 *> dbuf;
    int max_batch = 0;
    int last_launches = 0;
    bool profile = false;
    cudaEvent_t ev[2] = {nullptr, nullptr};
    int profile_op = -1;
    // one instantiated CUDA graph per batch size: the ~70 launches of a forward are replayed with a single
    // cudaGraphLaunch; only the three nodes that see caller pointers are re-parameterised per call
    struct GraphEntry {
        int B = 0;
        cudaGraph_t graph = nullptr;
        cudaGraphExec_t exec = nullptr;
        cudaGraphNode_t n_face = nullptr, n_mel = nullptr, n_out = nullptr;
        const float *mel = nullptr;
        const uint8_t *faces = nullptr;
        uint8_t *out_u8 = nullptr;
        float *out_f32 = nullptr;
        ConvParams out_params;
        LaunchDesc out_desc;
    };
    std::vector<GraphEntry> graphs;
    bool use_graph = true;
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

void wav2lip_destroy(mf_ctx *ctx) {
    Wav2LipState *s = ctx->wav2lip;
    if (!s) return;
    for (auto p : s->dbuf) cudaFree(p);
    for (auto &g : s->graphs) {
        if (g.exec) cudaGraphExecDestroy(g.exec);
        if (g.graph) cudaGraphDestroy(g.graph);
    }
    if (s->ev[0]) { cudaEventDestroy(s->ev[0]); cudaEventDestroy(s->ev[1]); }
    delete s;
    ctx->wav2lip = nullptr;
}

template <int BN, int STAGES>
static cudaError_t conv_desc_s(dim3 grid, LaunchDesc *d) {
    static bool attr_set = false;
    if (!attr_set) {
        cudaError_t e = cudaFuncSetAttribute(k_conv<BN, STAGES>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                             ConvSmem<BN, STAGES>::TOTAL);
        if (e != cudaSuccess) return e;
        attr_set = true;
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

extern "C" int mf_wav2lip_load(mf_ctx *ctx, const void *blob, size_t nbytes, int max_batch) {
    if (!ctx) return MF_E_INVALID;
    MF_REQUIRE(ctx, blob && max_batch >= 1 && max_batch <= 256, "mf_wav2lip_load: bad arguments");
    MF_CUDA(ctx, cudaSetDevice(ctx->device));
    wav2lip_destroy(ctx);
    PFN_encodeTiled encode = get_encode();
    if (!encode) return mf_fail(ctx, MF_E_CUDA, "cuTensorMapEncodeTiled entry point not found");
    std::vector<unsigned char> head(sizeof(mf_blob_header) + MF_BLOB_MAX_ENTRIES * sizeof(mf_blob_entry));
    const size_t hbytes = std::min(head.size(), nbytes);
    MF_CUDA(ctx, cudaMemcpy(head.data(), blob, hbytes, cudaMemcpyDeviceToHost));
    const mf_blob_header *h = reinterpret_cast<const mf_blob_header *>(head.data());
    MF_REQUIRE(ctx, hbytes >= sizeof(mf_blob_header) && h->magic == MF_BLOB_MAGIC && h->kind == 2,
               "mf_wav2lip_load: not a conv-net blob");
    MF_REQUIRE(ctx, h->n_entries <= MF_BLOB_MAX_ENTRIES, "mf_wav2lip_load: too many entries");
    const mf_blob_entry *ent = reinterpret_cast<const mf_blob_entry *>(head.data() + sizeof(mf_blob_header));
    const unsigned char *base = reinterpret_cast<const unsigned char *>(blob);
    auto find = [&](uint32_t id) -> const mf_blob_entry * {
        for (uint32_t i = 0; i < h->n_entries; i++)
            if (ent[i].id == id && ent[i].offset + ent[i].nbytes <= nbytes) return &ent[i];
        return nullptr;
    };
    const mf_blob_entry *pe = find(W2L_ID_PROGRAM);
    MF_REQUIRE(ctx, pe && pe->nbytes >= sizeof(W2LHeader), "mf_wav2lip_load: program entry missing");
    std::vector<unsigned char> prog(pe->nbytes);
    MF_CUDA(ctx, cudaMemcpy(prog.data(), base + pe->offset, pe->nbytes, cudaMemcpyDeviceToHost));
    Wav2LipState *s = new (std::nothrow) Wav2LipState();
    MF_REQUIRE(ctx, s, "out of host memory");
    ctx->wav2lip = s;
    s->hdr = *reinterpret_cast<const W2LHeader *>(prog.data());
    const size_t need = sizeof(W2LHeader) + (size_t)s->hdr.n_buffers * sizeof(W2LBuffer) + (size_t)s->hdr.n_ops * sizeof(W2LOp);
    MF_REQUIRE(ctx, s->hdr.n_buffers > 0 && s->hdr.n_ops > 0 && need == pe->nbytes, "mf_wav2lip_load: program size mismatch");
    const W2LBuffer *pb = reinterpret_cast<const W2LBuffer *>(prog.data() + sizeof(W2LHeader));
    const W2LOp *po = reinterpret_cast<const W2LOp *>(pb + s->hdr.n_buffers);
    s->bufs.assign(pb, pb + s->hdr.n_buffers);
    s->ops.assign(po, po + s->hdr.n_ops);
    s->max_batch = max_batch;
    { const char *e = getenv("MF_NO_GRAPH"); s->use_graph = !(e && atoi(e)); }
    s->dbuf.assign(s->hdr.n_buffers, nullptr);
    for (int i = 0; i < s->hdr.n_buffers; i++) {
        const size_t bytes = (size_t)max_batch * s->bufs[i].H * s->bufs[i].W * s->bufs[i].C * 2;
        MF_CUDA(ctx, cudaMalloc(&s->dbuf[i], bytes));
        MF_CUDA(ctx, cudaMemset(s->dbuf[i], 0, bytes));
    }
    s->params.resize(s->hdr.n_ops);
    for (int i = 0; i < s->hdr.n_ops; i++) {
        const W2LOp &o = s->ops[i];
        ConvParams &p = s->params[i];
        memset(&p, 0, sizeof(p));
        auto okbuf = [&](int b) { return b >= 0 && b < s->hdr.n_buffers; };
        MF_REQUIRE(ctx, okbuf(o.in_buf) && (o.mode == 1 || okbuf(o.out_buf)) && (o.res_buf < 0 || okbuf(o.res_buf)),
                   "op %d: bad buffer id", i);
        MF_REQUIRE(ctx, o.BN == 16 || o.BN == 32 || o.BN == 64 || o.BN == 128, "op %d: BN %d unsupported", i, o.BN);
        MF_REQUIRE(ctx, o.ntaps >= 1 && o.ntaps <= CONV_MAX_TAPS && o.Cin % 8 == 0 && o.Kpad % CONV_BK == 0 &&
                            o.Kpad >= o.ntaps * o.Cin && o.Cout_pad % o.BN == 0 && o.Cout <= o.Cout_pad,
                   "op %d: bad geometry", i);
        const mf_blob_entry *we = find(o.w_entry), *se = find(o.scale_entry), *he = find(o.shift_entry);
        MF_REQUIRE(ctx, we && se && he && we->nbytes == (size_t)o.Cout_pad * o.Kpad * 2 &&
                            se->nbytes == (size_t)o.Cout_pad * 4 && he->nbytes == (size_t)o.Cout_pad * 4,
                   "op %d: weight/scale/shift entries do not match the geometry (strict loader)", i);
        const W2LBuffer &ib = s->bufs[o.in_buf];
        MF_REQUIRE(ctx, o.in_coff % 8 == 0 && ib.C % 8 == 0 && o.in_coff + o.Cin <= ib.C, "op %d: input channels out of range", i);
        p.in = s->dbuf[o.in_buf];
        p.in_stride = ib.C; p.in_coff = o.in_coff; p.Hin = ib.H; p.Win = ib.W;
        if (o.mode == 0) {
            const W2LBuffer &ob = s->bufs[o.out_buf];
            MF_REQUIRE(ctx, o.out_coff % 8 == 0 && ob.C % 8 == 0 && o.out_coff + o.Cout <= ob.C && o.Cout % 16 == 0,
                       "op %d: output channels out of range", i);
            p.out = s->dbuf[o.out_buf];
            p.out_stride = ob.C; p.out_coff = o.out_coff; p.Hout = ob.H; p.Wout = ob.W;
        } else {
            MF_REQUIRE(ctx, o.Cout <= 16 && o.BN == 16, "op %d: output head must have Cout <= 16", i);
            p.Hout = s->hdr.out_hw; p.Wout = s->hdr.out_hw;
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
        p.ntaps = o.ntaps; p.Cin = o.Cin; p.nkb = o.Kpad / CONV_BK; p.Cout = o.Cout; p.relu = o.relu; p.mode = o.mode;
        memcpy(p.tap_dy, o.tap_dy, CONV_MAX_TAPS);
        memcpy(p.tap_dx, o.tap_dx, CONV_MAX_TAPS);
        // weights [Cout_pad][Kpad] bf16, K contiguous: TMA box = 64 (K) x BN rows, 128B swizzle
        cuuint64_t dims[2] = {(cuuint64_t)o.Kpad, (cuuint64_t)o.Cout_pad};
        cuuint64_t strides[1] = {(cuuint64_t)o.Kpad * 2};
        cuuint32_t box[2] = {CONV_BK, (cuuint32_t)o.BN};
        cuuint32_t estr[2] = {1, 1};
        CUresult cr = encode(&p.wmap, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, (void *)(base + we->offset), dims, strides, box,
                             estr, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B,
                             CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
        if (cr != CUDA_SUCCESS) return mf_fail(ctx, MF_E_CUDA, "op %d: cuTensorMapEncodeTiled failed (%d)", i, (int)cr);
    }
    MF_CUDA(ctx, cudaDeviceSynchronize());
    return MF_OK;
}

static int run_ops(mf_ctx *ctx, Wav2LipState *s, int B, uint8_t *out_u8, float *out_f32, cudaStream_t st, int *launches) {
    for (int i = 0; i < s->hdr.n_ops; i++) {
        ConvParams p = s->params[i];
        p.M = B * p.Mh * p.Mw;
        {
            static int dbg = -1;
            if (dbg < 0) { const char *e = getenv("MF_CONV_DBG"); dbg = e ? atoi(e) : 0; }
            p.dbg = dbg;
        }
        if (p.mode == 1) {
            p.out = out_u8;
            p.out_f32 = out_f32;
        }
        if (s->profile && i == s->profile_op) cudaEventRecord(s->ev[0], st);
        LaunchDesc d;
        cudaError_t e = conv_desc_any(s->ops[i].BN, p, &d);
        if (e == cudaSuccess) {
            void *args[] = {&p};
            e = cudaLaunchKernel(d.func, d.grid, dim3(CONV_THREADS), args, d.smem, st);
        }
        if (s->profile && i == s->profile_op) cudaEventRecord(s->ev[1], st);
        if (e != cudaSuccess) return mf_fail(ctx, MF_E_CUDA, "conv op %d launch: %s", i, cudaGetErrorString(e));
        (*launches)++;
    }
    return MF_OK;
}

static int add_kernel_node(mf_ctx *ctx, cudaGraph_t g, cudaGraphNode_t *prev, cudaGraphNode_t *out, void *func, dim3 grid,
                           dim3 block, int smem, void **args) {
    cudaKernelNodeParams kp;
    memset(&kp, 0, sizeof(kp));
    kp.func = func; kp.gridDim = grid; kp.blockDim = block; kp.sharedMemBytes = (unsigned)smem; kp.kernelParams = args;
    MF_CUDA(ctx, cudaGraphAddKernelNode(out, g, *prev ? prev : nullptr, *prev ? 1 : 0, &kp));
    *prev = *out;
    return MF_OK;
}

static int set_kernel_node(mf_ctx *ctx, cudaGraphExec_t ex, cudaGraphNode_t node, void *func, dim3 grid, dim3 block, int smem,
                           void **args) {
    cudaKernelNodeParams kp;
    memset(&kp, 0, sizeof(kp));
    kp.func = func; kp.gridDim = grid; kp.blockDim = block; kp.sharedMemBytes = (unsigned)smem; kp.kernelParams = args;
    MF_CUDA(ctx, cudaGraphExecKernelNodeSetParams(ex, node, &kp));
    return MF_OK;
}

// the whole forward as one graph launch; returns MF_E_UNSUPPORTED when the graph path must be skipped
static int forward_graph(mf_ctx *ctx, Wav2LipState *s, const float *mel, const uint8_t *faces, uint8_t *out_u8,
                         float *out_f32, int B, cudaStream_t st) {
    const int S = s->hdr.face_hw;
    int nf = B * S * S, nm = B * s->hdr.mel_h * s->hdr.mel_w, Bv = B, Sv = S;
    __nv_bfloat16 *fbuf = s->dbuf[s->hdr.in_face_buf], *mbuf = s->dbuf[s->hdr.in_mel_buf];
    Wav2LipState::GraphEntry *ge = nullptr;
    for (auto &g : s->graphs) if (g.B == B) ge = &g;
    if (!ge) {
        if (s->ops.empty() || s->ops.back().mode != 1) return MF_E_UNSUPPORTED;
        s->graphs.emplace_back();
        ge = &s->graphs.back();
        ge->B = B;
        MF_CUDA(ctx, cudaGraphCreate(&ge->graph, 0));
        cudaGraphNode_t prev = nullptr, node = nullptr;
        {
            void *a0[] = {(void *)&faces, (void *)&fbuf, &Bv, &Sv};
            int rc = add_kernel_node(ctx, ge->graph, &prev, &ge->n_face, (void *)k_w2l_prep_face, dim3((nf + 255) / 256), dim3(256), 0, a0);
            if (rc) return rc;
            void *a1[] = {(void *)&mel, (void *)&mbuf, &nm};
            rc = add_kernel_node(ctx, ge->graph, &prev, &ge->n_mel, (void *)k_w2l_prep_mel, dim3((nm + 255) / 256), dim3(256), 0, a1);
            if (rc) return rc;
        }
        for (int i = 0; i < s->hdr.n_ops; i++) {
            ConvParams p = s->params[i];
            p.M = B * p.Mh * p.Mw;
            p.dbg = 0;
            if (p.mode == 1) { p.out = out_u8; p.out_f32 = out_f32; }
            LaunchDesc d;
            MF_CUDA(ctx, conv_desc_any(s->ops[i].BN, p, &d));
            void *a[] = {&p};
            int rc = add_kernel_node(ctx, ge->graph, &prev, &node, d.func, d.grid, dim3(CONV_THREADS), d.smem, a);
            if (rc) return rc;
            if (p.mode == 1) { ge->n_out = node; ge->out_params = p; ge->out_desc = d; }
        }
        MF_CUDA(ctx, cudaGraphInstantiate(&ge->exec, ge->graph, 0));
        ge->mel = mel; ge->faces = faces; ge->out_u8 = out_u8; ge->out_f32 = out_f32;
    }
    if (ge->faces != faces) {
        void *a0[] = {(void *)&faces, (void *)&fbuf, &Bv, &Sv};
        int rc = set_kernel_node(ctx, ge->exec, ge->n_face, (void *)k_w2l_prep_face, dim3((nf + 255) / 256), dim3(256), 0, a0);
        if (rc) return rc;
        ge->faces = faces;
    }
    if (ge->mel != mel) {
        void *a1[] = {(void *)&mel, (void *)&mbuf, &nm};
        int rc = set_kernel_node(ctx, ge->exec, ge->n_mel, (void *)k_w2l_prep_mel, dim3((nm + 255) / 256), dim3(256), 0, a1);
        if (rc) return rc;
        ge->mel = mel;
    }
    if (ge->out_u8 != out_u8 || ge->out_f32 != out_f32) {
        ge->out_params.out = out_u8;
        ge->out_params.out_f32 = out_f32;
        void *a[] = {&ge->out_params};
        int rc = set_kernel_node(ctx, ge->exec, ge->n_out, ge->out_desc.func, ge->out_desc.grid, dim3(CONV_THREADS), ge->out_desc.smem, a);
        if (rc) return rc;
        ge->out_u8 = out_u8; ge->out_f32 = out_f32;
    }
    MF_CUDA(ctx, cudaGraphLaunch(ge->exec, st));
    s->last_launches = 2 + s->hdr.n_ops;
    return MF_OK;
}

extern "C" int mf_wav2lip_forward(mf_ctx *ctx, const float *mel, const uint8_t *faces, uint8_t *out_u8, float *out_f32,
                                  int B, void *stream) {
    if (!ctx) return MF_E_INVALID;
    Wav2LipState *s = ctx->wav2lip;
    if (!s) return mf_fail(ctx, MF_E_STATE, "mf_wav2lip_forward: weights not loaded");
    MF_REQUIRE(ctx, mel && faces && (out_u8 || out_f32), "mf_wav2lip_forward: null pointer");
    MF_REQUIRE(ctx, B >= 1 && B <= s->max_batch, "mf_wav2lip_forward: batch %d outside [1, %d]", B, s->max_batch);
    MF_REQUIRE(ctx, s->hdr.in_face_buf >= 0 && s->hdr.in_mel_buf >= 0, "program has no wav2lip inputs");
    cudaStream_t st = (cudaStream_t)stream;
    MF_CUDA(ctx, cudaSetDevice(ctx->device));
    int launches = 0;
    const int S = s->hdr.face_hw;
    const int nf = B * S * S, nm = B * s->hdr.mel_h * s->hdr.mel_w;
    if (s->use_graph && !s->profile) {
        int rc = forward_graph(ctx, s, mel, faces, out_u8, out_f32, B, st);
        if (rc != MF_E_UNSUPPORTED) return rc;
    }
    k_w2l_prep_face<<<(nf + 255) / 256, 256, 0, st>>>(faces, s->dbuf[s->hdr.in_face_buf], B, S);
    k_w2l_prep_mel<<<(nm + 255) / 256, 256, 0, st>>>(mel, s->dbuf[s->hdr.in_mel_buf], nm);
    launches += 2;
    int rc = run_ops(ctx, s, B, out_u8, out_f32, st, &launches);
    if (rc) return rc;
    s->last_launches = launches;
    return MF_OK;
}

// unit-test entry: run the loaded program on an fp32 NHWC tensor written into buffer `in_buf`
// (all of its channels) and read buffer `out_buf` back as fp32 NHWC.
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
    k_f32_to_bf16<<<(unsigned)((n_in + 255) / 256), 256, 0, st>>>(in_f32, s->dbuf[in_buf], n_in);
    int launches = 1;
    for (int i = 0; i < s->hdr.n_ops; i++) MF_REQUIRE(ctx, s->ops[i].mode == 0, "debug run supports mode-0 ops only");
    int rc = run_ops(ctx, s, B, nullptr, nullptr, st, &launches);
    if (rc) return rc;
    k_bf16_to_f32<<<(unsigned)((n_out + 255) / 256), 256, 0, st>>>(s->dbuf[out_buf], out_f32, n_out);
    MF_CUDA(ctx, cudaGetLastError());
    s->last_launches = launches + 1;
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
