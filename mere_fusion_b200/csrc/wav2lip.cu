#include "mf_common.cuh"
void wav2lip_destroy(mf_ctx *ctx) { (void)ctx; }
