// mf_common.cuh -- context, error handling and small PTX helpers shared by all heads.
#pragma once
#include <cuda_fp16.h>
#include <cuda_bf16.h>
#include <cuda_runtime.h>

#include <cstdarg>
#include <cstdint>
#include <cstdio>
#include <cstring>
#include <new>

#include "../../include/mf_b200.h"

struct ErnerfState;
struct Wav2LipState;

struct mf_ctx {
    int device = 0;
    int sm_count = 0;
    char err[512] = {0};
    ErnerfState *ernerf = nullptr;
    Wav2LipState *wav2lip = nullptr;
    float *mel_scratch = nullptr;   // Wav2Lip mel front-end: [frames][80] fp32
    int mel_frames_cap = 0;
};

static inline int mf_fail(mf_ctx *ctx, int code, const char *fmt, ...) {
    if (ctx) {
        va_list ap;
        va_start(ap, fmt);
        vsnprintf(ctx->err, sizeof(ctx->err), fmt, ap);
        va_end(ap);
    }
    return code;
}

#define MF_CUDA(ctx, call)                                                                        \
    do {                                                                                          \
        cudaError_t e__ = (call);                                                                 \
        if (e__ != cudaSuccess)                                                                   \
            return mf_fail((ctx), MF_E_CUDA, "%s:%d %s -> %s", __FILE__, __LINE__, #call,         \
                           cudaGetErrorString(e__));                                              \
    } while (0)

// cudaFuncSetAttribute is per DEVICE: remember per (call site, device) whether it has been done
struct mf_per_device_flag {
    bool done[64] = {false};
    bool test_and_set(int device) {
        const int d = device & 63;
        const bool was = done[d];
        done[d] = true;
        return was;
    }
};

#define MF_REQUIRE(ctx, cond, ...)                                                                \
    do {                                                                                          \
        if (!(cond)) return mf_fail((ctx), MF_E_INVALID, __VA_ARGS__);                            \
    } while (0)

// ------------------------------------------------------------------------------------------
// packed-weight blob header (written by mere_fusion_b200/*_pack.py)
// ------------------------------------------------------------------------------------------
#define MF_BLOB_MAGIC 0x3242464Du /* 'MFB2' */
struct mf_blob_entry {
    uint32_t id;
    uint32_t reserved;
    uint64_t offset;
    uint64_t nbytes;
};
struct mf_blob_header {
    uint32_t magic, kind, version, n_entries;
};
#define MF_BLOB_MAX_ENTRIES 256

// ------------------------------------------------------------------------------------------
// device helpers
// ------------------------------------------------------------------------------------------
#ifdef __CUDACC__
// programmatic dependent launch: a kernel launched with cudaLaunchAttributeProgrammaticStreamSerialization may start while its
// predecessor drains; pdl_wait() blocks until the predecessor grid has completed and its writes are visible (no-op otherwise)
__device__ __forceinline__ void pdl_launch() { asm volatile("griddepcontrol.launch_dependents;\n" ::: "memory"); }
__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;\n" ::: "memory"); }

__device__ __forceinline__ uint32_t smem_u32(const void *p) {
    return (uint32_t)__cvta_generic_to_shared(p);
}
__device__ __forceinline__ uint32_t pack_half2(float lo, float hi) {
    __half2 h = __floats2half2_rn(lo, hi);
    return *reinterpret_cast<uint32_t *>(&h);
}
__device__ __forceinline__ float round_half(float v) { return __half2float(__float2half_rn(v)); }

// mma.sync m16n8k16 fp16 x fp16 -> fp32 (the ErNeRF per-sample MLPs are K,N <= 80: warp-level
// tiles chained in registers; the dense conv stacks use tcgen05 instead, see wav2lip_conv.cu)
__device__ __forceinline__ void mma16816(float (&c)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1) {
    asm volatile(
        "mma.sync.aligned.m16n8k16.row.col.f32.f16.f16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, "
        "{%0,%1,%2,%3};\n"
        : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3])
        : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}
__device__ __forceinline__ void ldmatrix_x4(uint32_t (&r)[4], uint32_t saddr) {
    asm volatile("ldmatrix.sync.aligned.m8n8.x4.shared.b16 {%0,%1,%2,%3}, [%4];\n"
                 : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3])
                 : "r"(saddr));
}
__device__ __forceinline__ void ldmatrix_x2(uint32_t &r0, uint32_t &r1, uint32_t saddr) {
    asm volatile("ldmatrix.sync.aligned.m8n8.x2.shared.b16 {%0,%1}, [%2];\n"
                 : "=r"(r0), "=r"(r1)
                 : "r"(saddr));
}

// mbarrier + 1-D bulk TMA (cp.async.bulk): stage a contiguous weight image into shared memory
__device__ __forceinline__ void mbar_init(uint64_t *bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;\n" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t *bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;\n" ::"r"(smem_u32(bar)), "r"(bytes)
                 : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t *bar, uint32_t phase) {
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "WAIT_%=:\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
        "@p bra DONE_%=;\n"
        "bra WAIT_%=;\n"
        "DONE_%=:\n"
        "}\n" ::"r"(smem_u32(bar)),
        "r"(phase)
        : "memory");
}
__device__ __forceinline__ void fence_barrier_init() {
    asm volatile("fence.mbarrier_init.release.cluster;\n" ::: "memory");
}
__device__ __forceinline__ void fence_proxy_async() {
    asm volatile("fence.proxy.async.shared::cta;\n" ::: "memory");
}
__device__ __forceinline__ void bulk_g2s(void *smem_dst, const void *gmem_src, uint32_t bytes, uint64_t *bar) {
    asm volatile(
        "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];\n" ::"r"(
            smem_u32(smem_dst)),
        "l"(gmem_src), "r"(bytes), "r"(smem_u32(bar))
        : "memory");
}
#endif
