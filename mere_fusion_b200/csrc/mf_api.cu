// mf_api.cu -- context lifetime of libmf_b200.so (include/mf_b200.h).
#include "mf_common.cuh"

void ernerf_destroy(mf_ctx *ctx);
void wav2lip_destroy(mf_ctx *ctx);

extern "C" int mf_version(void) { return MF_ABI_VERSION; }

extern "C" int mf_create(int device, mf_ctx **out) {
    if (!out) return MF_E_INVALID;
    *out = nullptr;
    int n = 0;
    if (cudaGetDeviceCount(&n) != cudaSuccess || device < 0 || device >= n) return MF_E_CUDA;
    cudaDeviceProp prop;
    if (cudaGetDeviceProperties(&prop, device) != cudaSuccess) return MF_E_CUDA;
    // sm_100a SASS only: no PTX fallback, no other architecture, no CPU path
    if (prop.major != 10) return MF_E_UNSUPPORTED;
    if (cudaSetDevice(device) != cudaSuccess) return MF_E_CUDA;
    mf_ctx *ctx = new (std::nothrow) mf_ctx();
    if (!ctx) return MF_E_INVALID;
    ctx->device = device;
    ctx->sm_count = prop.multiProcessorCount;
    *out = ctx;
    return MF_OK;
}

extern "C" void mf_destroy(mf_ctx *ctx) {
    if (!ctx) return;
    cudaSetDevice(ctx->device);
    ernerf_destroy(ctx);
    wav2lip_destroy(ctx);
    cudaFree(ctx->mel_scratch);
    delete ctx;
}

extern "C" const char *mf_last_error(const mf_ctx *ctx) { return ctx ? ctx->err : "null context"; }
