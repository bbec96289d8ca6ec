// ernerf.cu -- sm_100a ErNeRF frame renderer behind the C ABI of include/mf_b200.h.
//
// Replaces, per frame, the ~400 launches + <=16 host syncs of the reference's
// NeRFRenderer.run_cuda / run_torso (ernerf/nerf_triplane/renderer.py:158-352) by
//   k_setup         CTAs < 8 n_frames: AudioNet on one attention window each, the last one to finish adds AudioAttNet +
//                                      EMA (network.py:9-66,222-237) of its frame
//                   the other CTAs   : the RAY PASS -- 32-ray tiles (8 x 4 pixel blocks) drawn by ticket: ray generation
//                                      (utils.py:255-341) -> near/far (raymarching.cu:91-145) -> march to the first sample
//                                      (raymarching.cu:827-929); rays that have one are appended to the frame's hit list
//                                      (this is the reference's round 0 minus its shading: n_step = N / N = 1)
//   k_head          persistent, NO grid barrier: each warp owns 32 / CH rays at a time, marches CH samples of each
//                   side by side -> tri-plane hash-grid gather (gridencoder.cu:75-175) -> SH (shencoder.cu:27-68) ->
//                   aud_ch_att / eye_att / sigma / color MLPs (network.py:249-308) on warp-level tensor-core tiles ->
//                   composite (raymarching.cu:2141-2249) in sample order, and refills finished rays from the hit list
//   k_torso_compose torso occupancy + deformation + tiled grid + MLPs (network.py:166-201, renderer.py:294-352) over the
//                   background for all its tiles first, THEN it waits for k_head, resolves the loop control from the life
//                   histogram and composes: head + (1 - weights_sum) * that, clamp, fp32 / u8 output
//   k_resize_u8     bilinear resize (utils.py:1212) + u8 (nerfreal.py:110) when sizes differ
// The three kernels of a frame are chained by programmatic dependent launch (griddepcontrol): k_head stages its MLP image
// while k_setup drains, k_torso_compose's torso pass runs on the SMs k_head's CTAs have already left.
//
// Round semantics WITHOUT rounds.  The reference marches every alive ray n_step = clamp(N / n_alive, 1, 8) samples per
// round and stops when the summed n_step reaches max_steps (renderer.py:246-270); n_alive is a global quantity, which is
// why round 1 of this kernel had one grid barrier per round (~45 % of its warp-time waited there, profiles/r01_k_head_v4).
// But the samples of a ray do not depend on how they are cut into rounds: march_rays continues from rays_t with a
// sigma-independent step (raymarching.cu:872-928), the compositor is a recurrence over the ray's samples in order with two
// exits -- the ray runs out of samples (deltas[0] == 0) or T < T_thresh after accumulating a sample (:2189-2218).  So a
// ray's result is the composite of its first min(C, own end) samples, where C = sum of n_step over the rounds is the ONLY
// thing the round structure decides, and C follows from A(c) = number of rays still alive after c samples:
//     c = 0; while c < max_steps: n_alive = A(c); n_step = clamp(N / n_alive, 1, 8); c += n_step;   C = c  (<= max_steps + 7)
// A ray is alive after c samples iff life >= c, life = min(samples available, index of the T-exit sample - 1).  k_head
// therefore shades every ray at its own pace (CH samples per pass, any order), counts rays by life in a histogram, and the
// few rays that are still alive after max_steps samples go on speculatively to max_steps + 7 while recording their state
// after each of the samples max_steps .. max_steps + 7; k_torso_compose derives the rounds (n_alive, n_step) and C from the
// histogram and picks the snapshot C - max_steps for those rays.  Same arithmetic per ray, same samples, same exits: the
// image is the round-structured one bit for bit (tests/test_ernerf_reference_render_gpu.py pins it on the reference render).

#include <algorithm>
#include <cmath>
#include <cstring>
#include <new>
#include <vector>

#include "ernerf_device.cuh"
#include "ernerf_layout.h"

using namespace ernerf;

#define ER_MAX_ROUNDS 16
#define ER_MAX_STEPS 32  /* cfg.max_steps supported by the life histogram */
// per-frame counters (ints); two copies, alternating per frame of a context: k_setup of frame n zeroes the copy frame n+1 uses
#define CT_NHIT 0     /* rays with a first sample = length of the hit list */
#define CT_TICKET 1   /* next unclaimed hit-list slot (k_head refill) */
#define CT_SAMPLES 2  /* samples shaded by k_head */
#define CT_NSURV 3    /* rays alive after max_steps samples = snapshot slots in use */
#define CT_PASSES 5   /* k_head telemetry: warp passes (32 sample lanes each) of the launch (frame 0's counters) */
#define CT_TILES 6    /* k_head telemetry: 16-row MLP tiles evaluated */
#define CT_RTICKET 7  /* next unclaimed 32-ray tile of the ray pass */
#define CT_HIST 8     /* [ER_MAX_STEPS + 1] rays by life; bin max_steps = alive after max_steps samples */
#define CT_ROUNDS 48  /* [ER_MAX_ROUNDS + 1][4] derived by k_torso_compose: n_alive, k_head telemetry (row 0: warp passes, row 1: MLP tiles), samples emitted (round 0; -1 = not tracked), n_step */
#define CT_INTS 128
#define ER_SNAPS 8    /* states recorded per surviving ray: after sample max_steps + j, j = 0..7 */

// levels [0, n_dense) are IDX_DENSE, the rest IDX_HASH2 (head) / IDX_TILE2 (torso); n_dense < 0: generic
struct HeadLevels { GridLevel lv[MF_ERNERF_HEAD_LEVELS]; int n_dense; };
struct TorsoLevels { GridLevel lv[MF_ERNERF_TORSO_LEVELS]; int n_dense; };

static int classify_levels(const GridLevel *lv, int L, int tail_class) {
    int n = 0;
    while (n < L && grid_level_class(lv[n]) == IDX_DENSE) n++;
    for (int l = n; l < L; l++)
        if (grid_level_class(lv[l]) != tail_class) return -1;
    return n;
}

struct FrameGeom {
    int N, H, W;
    float R[9], T[3];
    float inv_fx, inv_fy, cx, cy;
    const float *rays_o, *rays_d, *bg_coords;  // nullable: explicit rays
    float inv_Hm1, inv_Wm1;
};

struct ErnerfState {
    mf_ernerf_cfg cfg;
    // blob views
    const float *planes = nullptr;
    uint32_t plane_rows = 0;
    const uint8_t *bitfield = nullptr;
    uint8_t *bitfield_linear = nullptr;   // owned: (z * H + y) * H + x copy for the fused kernels (cascade 1, H a multiple of 32)
    const __half2 *torso_table = nullptr;
    const float *torso_density = nullptr;
    const __half *head_mlp = nullptr, *torso_mlp = nullptr, *audio = nullptr, *torso_const = nullptr;
    const float *misc = nullptr;
    size_t audio_halfs = 0;
    HeadLevels hl;
    TorsoLevels tl;
    // persistent device state
    float *state = nullptr;  // [ST_FLOATS]: [0..31] enc_a (smoothed), [32] has_prev flag, [33] audio CTAs done, [128..383] encoded windows
    int *counters = nullptr; // [2][CT_INTS]
    unsigned frame_no = 0;   // parity selects the counter copy
    // per-N workspace
    int capN = 0;
    int *hits = nullptr;     // [N] ray ids with a first sample
    float4 *snap = nullptr;  // [N][ER_SNAPS] (ws, r, g, b) of rays alive after max_steps samples
    float *rays_t = nullptr, *fars = nullptr, *weights_sum = nullptr, *image = nullptr;
    float *final_f32 = nullptr;  // [N,3] when a resize follows
    float *torso_rgb = nullptr;  // [3][N] torso over background, written and read back by k_torso_compose (across its wait for k_head)
    int head_grid = 0;
    bool chunk_forced = false;
    int setup_per_sm = 1;    // k_setup CTAs per SM for a BATCHED pass (2 measured 1.5 % faster per frame at 4 sessions; for one frame the more
                             // interleaved hit list costs k_head 2 %, so a single frame keeps one CTA per SM)
    int chunk = 2;           // CH: samples of a ray shaded side by side (MF_HEAD_CHUNK = 1 | 2 | 4 | 8; 2 measured best)
    int last_launches = 0;
    float misc_host[24] = {0};
    bool profile = false;  // CUDA events around k_head on the launching stream (bench roofline)
    cudaEvent_t ev_head[2] = {nullptr, nullptr};  // anchor_points[12], individual_codes[0], individual_codes_torso[0]
};

// =========================================================================================
// load-time: level scales on the device (same exp2f as the reference kernel)
// =========================================================================================
__global__ void k_level_scales(float S, uint32_t H, int L, float *out) {
    const uint32_t level = threadIdx.x;
    if ((int)level < L) out[level] = exp2f(level * S) * H - 1.0f;
}

// load-time: the density bitfield (Morton order, raymarching.cu:56-71,894-895) re-indexed as (z * H + y) * H + x, same bits
__global__ void k_linear_bitfield(const uint8_t *__restrict__ morton, uint32_t H, uint32_t *linear_words) {
    const uint32_t w = blockIdx.x * blockDim.x + threadIdx.x;   // one 32-voxel word along x
    if (w >= H * H * H / 32) return;
    const uint32_t x0 = (w * 32) % H, y = (w * 32 / H) % H, z = w * 32 / (H * H);
    uint32_t bits = 0;
    for (uint32_t i = 0; i < 32; i++) {
        const uint32_t m = morton3D(x0 + i, y, z);
        if (morton[m / 8] & (1u << (m % 8))) bits |= 1u << i;
    }
    linear_words[w] = bits;
}

static void fill_levels(GridLevel *lv, int L, const float *scales, const int32_t *offsets, uint32_t gridtype) {
    for (int l = 0; l < L; l++) {
        GridLevel g;
        g.scale = scales[l];
        const uint32_t resolution = (uint32_t)std::ceil(scales[l]) + 1;
        g.stride1 = resolution + 1;
        g.offset = (uint32_t)offsets[l];
        g.size = (uint32_t)(offsets[l + 1] - offsets[l]);
        // get_grid_index (gridencoder.cu:54-72): d = 0 always contributes; d = 1 iff stride <= size
        uint32_t stride = g.stride1;
        uint32_t mode = 0;
        if (stride <= g.size) {
            mode |= 1u;
            stride *= g.stride1;
        }
        if (gridtype == 0 && stride > g.size) mode |= 2u;
        if ((g.size & (g.size - 1)) == 0) mode |= 4u;
        g.mode = mode;
        lv[l] = g;
    }
}

// =========================================================================================
// k_setup: audio encoder + per-frame constants
// =========================================================================================
struct SetupParams {
    const float *auds;        // [8, A, 16] or null
    const float *enc_a_in;    // [32] or null
    const __half *audio;      // weights image
    float *state;
    int *counters_next;       // the counter copy of this context's NEXT frame: zeroed here
    int A;
    int N;
    int smooth;
    int audio_halfs;          // size of the audio weight image (fp16 elements)
    float *dbg_enc_a;
};

__device__ __forceinline__ float lrelu16(float v16) { return v16 > 0.f ? v16 : round_half(v16 * 0.02f); }

#define SETUP_THREADS 1024
// ray pass (the CTAs of k_setup beyond the first n): one lane per ray
struct RayPassFrame {
    FrameGeom g;
    int *hits, *counters;
    float *rays_t, *fars, *nears, *weights_sum, *image;
};
struct SetupBatch {   // one audio CTA per frame of a batched render (HEAD_MAX_FRAMES) + the ray-pass CTAs
    int n;
    SetupParams f[4];
    RayPassFrame r[4];
    float bound, min_near, dt_gamma;
    uint32_t max_steps, cascade, grid_size;
    const uint8_t *bitfield;
    const uint8_t *bitfield_linear;   // nullable
    MarchParams mp;                   // make_march_params of the fields above (host), grid = the linear bitfield when there is one
    float aabb[6];
};
__device__ __forceinline__ void audio_cta(const SetupParams &p, int b);
__device__ __forceinline__ void gen_ray(const FrameGeom &g, int idx, Ray &r);

// The reference's round 0 (n_alive = N, n_step = 1) without its shading: which rays have a first sample, and where.
// rays_t[ray] = t AT the first sample (march_next from there finds it again at once), fars[ray]; rays without a sample get
// their final head result here (weights_sum = 0, image = 0).
__device__ __forceinline__ void ray_pass(const SetupBatch &b, int cta, int n_ctas) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, warps = blockDim.x >> 5;
    const MarchParams &mp = b.mp;
    for (int f = 0; f < b.n; f++) {
        const RayPassFrame &fr = b.r[f];
        const int n_tiles = (fr.g.N + 31) / 32;
        // tiles are handed out by ticket, in image order: a tile costs between nothing (rays that miss the box) and ~150 voxel tests
        // per ray, and a static split left the launch waiting for the warps that drew two expensive ones (in image order the hit
        // list also stays nearly sorted, which keeps the rays of a k_head warp alike: starting from the middle of the image
        // measured 2 % slower per frame)
        while (true) {
            int tile = 0;
            if (lane == 0) tile = atomicAdd(&fr.counters[CT_RTICKET], 1);
            tile = __shfl_sync(0xffffffffu, tile, 0);
            if (tile >= n_tiles) break;
            // a tile is an 8 x 4 pixel block when the image allows it (a warp's rays then stay together in k_head: nearer samples, more
            // shared table texels, more alike lives than a 32 x 1 row segment), else 32 consecutive rays
            int ray = tile * 32 + lane;
            if (!fr.g.rays_o && (fr.g.W & 7) == 0 && (fr.g.H & 3) == 0 && fr.g.N == fr.g.H * fr.g.W) {
                const int bw = fr.g.W >> 3, ty = tile / bw, tx = tile - ty * bw;
                ray = (ty * 4 + (lane >> 3)) * fr.g.W + tx * 8 + (lane & 7);
            }
            const bool valid = ray < fr.g.N;
            bool has = false;
            float t = 0.f;
            if (valid) {
                Ray ry;
                gen_ray(fr.g, ray, ry);
                float near, far, x, y, z, dt;
                uint32_t vox;
                near_far_aabb(ry.ox, ry.oy, ry.oz, ry.dx, ry.dy, ry.dz, b.aabb, b.min_near, near, far);
                t = near;
                has = march_find(mp, ry, t, far, x, y, z, dt, vox);
                fr.fars[ray] = far;
                if (fr.nears) fr.nears[ray] = near;
                if (has) {
                    fr.rays_t[ray] = t;
                } else {
                    fr.weights_sum[ray] = 0.f;
                    fr.image[ray * 3] = 0.f; fr.image[ray * 3 + 1] = 0.f; fr.image[ray * 3 + 2] = 0.f;
                }
            }
            const uint32_t hm = __ballot_sync(0xffffffffu, has);
            if (hm) {
                int base = 0;
                if (lane == 0) base = atomicAdd(&fr.counters[CT_NHIT], __popc(hm));
                base = __shfl_sync(0xffffffffu, base, 0);
                if (has) fr.hits[base + __popc(hm & ((1u << lane) - 1u))] = ray;
            }
        }
    }
}

// ---- audio encoder (network.py:9-66,222-237; smoothing renderer.py:190-194) -----------------------------------------------
// Round 1 ran it in ONE CTA (35-59 us of dependent layers, the 66 KB weight image staged in shared memory -- which also took
// the L1 away from the ray-pass CTAs of the same launch).  Now: one CTA per attention window (8 per frame), every output's dot
// product split over G lanes + a shuffle reduction, weights read straight from L2 with coalesced loads, and the CTA that
// finishes last runs the attention net + EMA on the 8 encoded windows.
// y[o] = act(fp16(bias[o] + sum_j W[o][j] * x[idx(o, j)])), o < O: G (power of two <= 32) lanes per output
template <class IdxF>
__device__ __forceinline__ void dense_split(const __half *__restrict__ W, const __half *__restrict__ bias, int O, int K, int G,
                                            const float *x, float *y, IdxF idx, bool lrelu) {
    for (int base = 0; base < O * G; base += blockDim.x) {   // uniform trip count: the shuffles below need whole warps
        const int gi = base + threadIdx.x;
        const int o = gi / G, sub = gi - o * G;
        const bool ok = o < O;
        float acc = 0.f;
        if (ok)
            for (int j = sub; j < K; j += G) {
                const int xi = idx(o, j);
                if (xi >= 0) acc = fmaf(__half2float(__ldg(W + (size_t)o * K + j)), x[xi], acc);
            }
        for (int m = G >> 1; m > 0; m >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, m);
        if (ok && sub == 0) {
            const float v = round_half(acc + __half2float(__ldg(bias + o)));
            y[o] = lrelu ? lrelu16(v) : v;
        }
    }
    __syncthreads();
}

// nn.Conv1d(k = 3, padding 1) on x[Cin][T] -> y[Cout][To] (values held as fp16-rounded floats): the weight row of output
// o = (co, t) is W[co]
template <class IdxF>
__device__ __forceinline__ void dense_rows(const __half *__restrict__ W, const __half *__restrict__ bias, int O, int K, int G, int To,
                                           const float *x, float *y, IdxF idx, bool lrelu) {
    for (int base = 0; base < O * G; base += blockDim.x) {
        const int gi = base + threadIdx.x;
        const int o = gi / G, sub = gi - o * G;
        const bool ok = o < O;
        const int co = o / To;
        float acc = 0.f;
        if (ok)
            for (int j = sub; j < K; j += G) {
                const int xi = idx(o, j);
                if (xi >= 0) acc = fmaf(__half2float(__ldg(W + (size_t)co * K + j)), x[xi], acc);
            }
        for (int m = G >> 1; m > 0; m >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, m);
        if (ok && sub == 0) {
            const float v = round_half(acc + __half2float(__ldg(bias + co)));
            y[o] = lrelu ? lrelu16(v) : v;
        }
    }
    __syncthreads();
}
__device__ __forceinline__ void conv1d_rows(const __half *W, const __half *b, int Cin, int T, int Cout, int stride, int G,
                                            const float *x, float *y) {
    const int To = (T + 2 - 3) / stride + 1;
    dense_rows(W, b, Cout * To, Cin * 3, G, To, x, y,
               [=](int o, int j) {
                   const int t = o % To, c = j / 3, k = j - 3 * c, ti = t * stride + k - 1;
                   return (ti >= 0 && ti < T) ? c * T + ti : -1;
               }, true);
}

// state layout (floats): [0..31] smoothed enc_a, [32] has-previous flag, [33] (int) audio CTAs done, [128..383] the 8 encoded windows
#define ST_DONE 33
#define ST_ENC8 128
#define ST_FLOATS 512

__device__ __forceinline__ void audio_cta(const SetupParams &p, int b) {
    __shared__ float bufA[64 * 16], bufB[64 * 16], att[8];
    __shared__ int s_last;
    const int tid = threadIdx.x;
    if (b == 0 && p.counters_next)   // the counters of this context's next frame (this frame's copy is being written by the ray-pass CTAs right now)
        for (int i = tid; i < CT_INTS; i += blockDim.x) p.counters_next[i] = 0;
    if (p.enc_a_in) {  // pre-encoded feature: k_head reads it where the caller put it; no EMA, the session's audio state is untouched
        if (b == 0 && tid < 32 && p.dbg_enc_a) p.dbg_enc_a[tid] = p.enc_a_in[tid];
        return;
    }
    // ---- AudioNet (network.py:40-66) on window b: x[:, 0:16] -> 4x conv(k3,s2,p1)+LeakyReLU -> fc
    const int A = p.A;
    const __half *w = p.audio;
    for (int i = tid; i < A * 16; i += blockDim.x) bufA[i] = round_half(p.auds[(size_t)b * A * 16 + i]);
    __syncthreads();
    conv1d_rows(w, w + 32 * A * 3, A, 16, 32, 2, 4, bufA, bufB);    w += 32 * A * 3 + 32;     // -> [32][8]
    conv1d_rows(w, w + 32 * 32 * 3, 32, 8, 32, 2, 8, bufB, bufA);   w += 32 * 32 * 3 + 32;    // -> [32][4]
    conv1d_rows(w, w + 64 * 32 * 3, 32, 4, 64, 2, 8, bufA, bufB);   w += 64 * 32 * 3 + 64;    // -> [64][2]
    conv1d_rows(w, w + 64 * 64 * 3, 64, 2, 64, 2, 16, bufB, bufA);  w += 64 * 64 * 3 + 64;    // -> [64][1]
    dense_split(w, w + 64 * 64, 64, 64, 16, bufA, bufB, [](int, int j) { return j; }, true);   w += 64 * 64 + 64;   // fc1.0 + LeakyReLU
    dense_split(w, w + 32 * 64, 32, 64, 32, bufB, bufA, [](int, int j) { return j; }, false);  w += 32 * 64 + 32;   // fc1.2
    if (tid < 32) p.state[ST_ENC8 + b * 32 + tid] = bufA[tid];
    // ---- the CTA that arrives last has all 8 windows: attention net + smoothing
    __syncthreads();
    if (tid == 0) {
        __threadfence();
        s_last = atomicAdd(reinterpret_cast<int *>(p.state) + ST_DONE, 1) == 7;
    }
    __syncthreads();
    if (!s_last) return;
    __threadfence();
    float *enc = bufB + 512;   // [8][32]
    for (int i = tid; i < 8 * 32; i += blockDim.x) enc[i] = __ldcg(p.state + ST_ENC8 + i);
    __syncthreads();
    // AudioAttNet (network.py:9-36): y = enc^T [32][8] -> 5x conv(k3,s1,p1)+LeakyReLU -> Linear(8,8) -> softmax
    for (int i = tid; i < 32 * 8; i += blockDim.x) bufA[i] = enc[(i % 8) * 32 + i / 8];
    __syncthreads();
    conv1d_rows(w, w + 16 * 32 * 3, 32, 8, 16, 1, 8, bufA, bufB);  w += 16 * 32 * 3 + 16;
    conv1d_rows(w, w + 8 * 16 * 3, 16, 8, 8, 1, 16, bufB, bufA);   w += 8 * 16 * 3 + 8;
    conv1d_rows(w, w + 4 * 8 * 3, 8, 8, 4, 1, 8, bufA, bufB);      w += 4 * 8 * 3 + 4;
    conv1d_rows(w, w + 2 * 4 * 3, 4, 8, 2, 1, 4, bufB, bufA);      w += 2 * 4 * 3 + 2;
    conv1d_rows(w, w + 1 * 2 * 3, 2, 8, 1, 1, 2, bufA, bufB);      w += 1 * 2 * 3 + 1;
    if (tid < 8) {
        float acc = 0.f;
        for (int k = 0; k < 8; k++) acc = fmaf(__half2float(__ldg(w + tid * 8 + k)), bufB[k], acc);
        att[tid] = round_half(acc + __half2float(__ldg(w + 64 + tid)));
    }
    __syncthreads();
    if (tid < 32) {
        float mx = att[0];
        for (int k = 1; k < 8; k++) mx = fmaxf(mx, att[k]);
        float e[8], sum = 0.f;
        for (int k = 0; k < 8; k++) { e[k] = expf(att[k] - mx); sum += e[k]; }
        float out = 0.f;
        for (int k = 0; k < 8; k++) out += (e[k] / sum) * enc[k * 32 + tid];
        // renderer.py:190-194
        if (p.smooth) {
            if (p.state[32] != 0.f) out = 0.35f * p.state[tid] + (1 - 0.35f) * out;
        }
        __syncwarp();
        p.state[tid] = out;
        if (tid == 0 && p.smooth) p.state[32] = 1.f;
        if (tid == 0) reinterpret_cast<int *>(p.state)[ST_DONE] = 0;
        if (p.dbg_enc_a) p.dbg_enc_a[tid] = out;
    }
}

// =========================================================================================
// warp-level MLP tiles (mma.sync m16n8k16, activations chained in registers)
// =========================================================================================
// one layer: c[NT] += a[KS] x W^T, W = smem [>=NT*8][WS] halfs.  B fragments via ldmatrix.
template <int KS, int NT, int WS>
__device__ __forceinline__ void mlp_layer(float (&c)[NT][4], const uint32_t (&a)[KS][4], const __half *W, int lane) {
    // ldmatrix.x4 row address for this lane: row (lane & 7) of matrix q = lane >> 3, k offset 8q
    const uint32_t base = smem_u32(W) + ((lane & 7) * WS + (lane >> 3) * 8) * 2;
#pragma unroll
    for (int nt = 0; nt < NT; nt++) {
#pragma unroll
        for (int kk = 0; kk + 1 < KS; kk += 2) {
            uint32_t b[4];
            ldmatrix_x4(b, base + (nt * 8 * WS + kk * 16) * 2);
            mma16816(c[nt], a[kk], b[0], b[1]);
            mma16816(c[nt], a[kk + 1], b[2], b[3]);
        }
        if (KS & 1) {
            uint32_t b0, b1;
            ldmatrix_x2(b0, b1, base + (nt * 8 * WS + (KS - 1) * 16) * 2);
            mma16816(c[nt], a[KS - 1], b0, b1);
        }
    }
}

template <int NT>
__device__ __forceinline__ void zero_acc(float (&c)[NT][4]) {
#pragma unroll
    for (int i = 0; i < NT; i++) c[i][0] = c[i][1] = c[i][2] = c[i][3] = 0.f;
}

__device__ __forceinline__ uint32_t relu_half2(uint32_t v) {
    const __half2 r = __hmax2(*reinterpret_cast<const __half2 *>(&v), __float2half2_rn(0.f));
    return *reinterpret_cast<const uint32_t *>(&r);
}
// accumulators of NT n-tiles -> A fragments of NT/2 k-steps for the next layer (ReLU optional)
template <int NT, bool RELU>
__device__ __forceinline__ void acc_to_a(const float (&c)[NT][4], uint32_t (&a)[NT / 2][4]) {
#pragma unroll
    for (int j = 0; j < NT / 2; j++) {
#pragma unroll
        for (int h = 0; h < 2; h++) {
            // ReLU on the packed pair (one HMNMX2 for two values): rounding to fp16 is monotone and keeps the sign, so
            // max(rn(x), 0) == rn(max(x, 0)) up to the sign of a zero, which a dot product cannot see
            const uint32_t lo = pack_half2(c[2 * j + h][0], c[2 * j + h][1]), hi = pack_half2(c[2 * j + h][2], c[2 * j + h][3]);
            a[j][2 * h + 0] = RELU ? relu_half2(lo) : lo;
            a[j][2 * h + 1] = RELU ? relu_half2(hi) : hi;
        }
    }
}

// load the A fragments of k-step `kk` of a 16-row tile stored row-major in shared memory
__device__ __forceinline__ void load_a(uint32_t (&a)[4], const __half *tile, int row_stride, int kk, int lane) {
    const int q = lane >> 3;
    const int row = (lane & 7) + 8 * (q & 1), col = kk * 16 + 8 * (q >> 1);
    ldmatrix_x4(a, smem_u32(tile + row * row_stride + col));
}

// =========================================================================================
// torso tiles (run by k_head's warps once they are out of rays) and k_compose
// =========================================================================================
#define TX_STRIDE 88 /* torso_net input tile: [feat 32 | freq 34 | pad] = 80 + 8 */
#ifndef TORSO_GL
#define TORSO_GL 4   /* torso grid levels gathered together */
#endif
#ifndef TORSO_MINB
#define TORSO_MINB 2 /* k_torso_compose CTAs per SM: 2 (105 registers, no spills) measured 4 % faster per frame than 3 (80 registers, 100 B of spills) */
#endif

struct TorsoModel {   // shared by the frames of a batch
    TorsoLevels tl;
    const __half2 *table;
    const float *density;    // [G*G]
    const __half *mlp_image, *torso_const;
    const float *misc;
    float thresh, shrink;
    int G;
};
struct TorsoFrame {
    const __half *bg_color;  // [N,3] or null (white)
    uint8_t *dbg_mask;
    float wa[6];             // wrapped anchor (fp16-rounded), network.py:175-176
};

// F.grid_sample(bilinear, zeros padding, align_corners=True) on a [G, G] image (renderer.py:326)
__device__ __forceinline__ float grid_sample_ac(const float *img, int G, float x, float y) {
    const float ix = ((x + 1.f) / 2) * (G - 1), iy = ((y + 1.f) / 2) * (G - 1);
    const float fx = floorf(ix), fy = floorf(iy);
    const int x0 = (int)fx, y0 = (int)fy, x1 = x0 + 1, y1 = y0 + 1;
    const float nw = ((fx + 1) - ix) * ((fy + 1) - iy), ne = (ix - fx) * ((fy + 1) - iy);
    const float sw = ((fx + 1) - ix) * (iy - fy), se = (ix - fx) * (iy - fy);
    auto at = [&](int yy, int xx) { return (xx >= 0 && xx < G && yy >= 0 && yy < G) ? __ldg(img + yy * G + xx) : 0.f; };
    float out = 0.f;
    out = fmaf(at(y0, x0), nw, out);
    out = fmaf(at(y0, x1), ne, out);
    out = fmaf(at(y1, x0), sw, out);
    out = fmaf(at(y1, x1), se, out);
    return out;
}

// the 16 levels of the torso grid (fp16 table, C = 2), four levels at a time: corner indices first, then their 16 loads back to
// back, then the per-level sums -- same arithmetic per level as grid_level_f16x2 (bit-identical features), but 4 exposed L2
// latencies per pixel instead of 16.  ND = number of leading dense levels (9 for the shipped 16 -> 2048 grid); ND < 0 = generic.
template <int ND>
__device__ __forceinline__ void torso_gather_t(const TorsoModel &tm, float u, float v, __half *row) {
    if (ND < 0) {
#pragma unroll
        for (int l = 0; l < MF_ERNERF_TORSO_LEVELS; l++) {
            float o0, o1;
            grid_level_f16x2<IDX_ANY>(tm.table, tm.tl.lv[l], u, v, o0, o1);
            *reinterpret_cast<uint32_t *>(row + 2 * l) = pack_half2(o0, o1);
        }
        return;
    }
    if (u < 0.f || u > 1.f || v < 0.f || v > 1.f) {   // grid_level_f16x2's bounds test
#pragma unroll
        for (int l = 0; l < MF_ERNERF_TORSO_LEVELS; l++) *reinterpret_cast<uint32_t *>(row + 2 * l) = 0u;
        return;
    }
    constexpr int GL = TORSO_GL;
#pragma unroll
    for (int l0 = 0; l0 < MF_ERNERF_TORSO_LEVELS; l0 += GL) {
        uint32_t idx[GL][4];
        float pu[GL], pv[GL];
        __half2 val[GL][4];
#pragma unroll
        for (int j = 0; j < GL; j++) {
            if (l0 + j < ND) grid_level_prep<IDX_DENSE>(tm.tl.lv[l0 + j], tm.tl.lv[l0 + j].offset, u, v, idx[j], pu[j], pv[j]);
            else grid_level_prep<IDX_TILE2>(tm.tl.lv[l0 + j], tm.tl.lv[l0 + j].offset, u, v, idx[j], pu[j], pv[j]);
        }
#pragma unroll
        for (int j = 0; j < GL; j++)
#pragma unroll
            for (int c = 0; c < 4; c++) val[j][c] = __ldg(tm.table + idx[j][c]);
#pragma unroll
        for (int j = 0; j < GL; j++) {
            const float w[4] = {(1 - pu[j]) * (1 - pv[j]), pu[j] * (1 - pv[j]), (1 - pu[j]) * pv[j], pu[j] * pv[j]};
            // Half += float (grid_level_f16x2): the product rounded to half, then a half + half sum rounded to half.  __hadd2 IS that sum:
            // the fp32 sum of two halves is either exact or (exponents > 13 bits apart) strictly inside the larger one's rounding
            // interval, so rounding it to half equals rounding the exact sum once.  4 instead of 12 instructions per corner.
            __half2 o = __float2half2_rn(0.f);
#pragma unroll
            for (int c = 0; c < 4; c++) {
                const float2 f = __half22float2(val[j][c]);
                o = __hadd2(o, __floats2half2_rn(w[c] * f.x, w[c] * f.y));
            }
            *reinterpret_cast<__half2 *>(row + 2 * (l0 + j)) = o;
        }
    }
}
__device__ __noinline__ void torso_gather_generic(const TorsoModel &tm, float u, float v, __half *row) { torso_gather_t<-1>(tm, u, v, row); }

// torso occupancy + deformation + tiled grid + MLPs (network.py:166-201, renderer.py:294-344) for one tile of 32 pixels;
// returns this lane's pixel's bgc = torso_color * torso_alpha + bg * (1 - torso_alpha).  mlp / bias: shared-memory torso MLP
// image and the frame's first-layer constants; xt: the warp's [32][TX_STRIDE] activation tile.
__device__ __forceinline__ void torso_tile(const TorsoModel &tm, const FrameGeom &g, const TorsoFrame &tf, int tile, int lane,
                                           __half *xt, const __half *mlp, const float *bias, float (&bgc)[3]) {
    const int N = g.N;
    const int t = lane & 3;
const int pix = tile * 32 + lane;
const bool valid = pix < N;
float c0 = 0.f, c1 = 0.f;  // bg_coords: c0 runs over image rows (utils.py:246-251)
if (valid) {
    if (g.bg_coords) {
        c0 = g.bg_coords[pix * 2];
        c1 = g.bg_coords[pix * 2 + 1];
    } else {
        const int row = pix / g.W, col = pix - row * g.W;
        c0 = ((float)row * g.inv_Hm1) * 2 - 1;
        c1 = ((float)col * g.inv_Wm1) * 2 - 1;
    }
}
const bool masked = valid && (grid_sample_ac(tm.density, tm.G, c0, c1) > tm.thresh);
const uint32_t mmask = __ballot_sync(0xffffffffu, masked);
float alpha = 0.f, tc0 = 0.f, tc1 = 0.f, tc2 = 0.f;

if (mmask) {
    // x * torso_shrink -> freq(deg 8) 34 values -> cols 32..65 of the tile (cols 66..79 zero)
    const float xin[2] = {c0 * tm.shrink, c1 * tm.shrink};
    __half *row = xt + lane * TX_STRIDE;
#pragma unroll
    for (int c = 0; c < 34; c += 2)
        *reinterpret_cast<uint32_t *>(row + 32 + c) =
            masked ? pack_half2(freq_elem(xin, 2, c), freq_elem(xin, 2, c + 1)) : 0u;
#pragma unroll
    for (int c = 66; c < 80; c += 2) *reinterpret_cast<uint32_t *>(row + c) = 0u;
    __syncwarp();

    float dxy[2] = {0.f, 0.f};
#pragma unroll 1
    for (int m = 0; m < 2; m++) {
        if (((mmask >> (16 * m)) & 0xffffu) == 0u) continue;
        // torso_deform_net: [freq 34 (+ constants as bias)] -> 32 -> 32 -> 2
        uint32_t a3[3][4];
#pragma unroll
        for (int kk = 0; kk < 3; kk++) load_a(a3[kk], xt + m * 16 * TX_STRIDE + 32, TX_STRIDE, kk, lane);
        float c[4][4];
#pragma unroll
        for (int nt = 0; nt < 4; nt++) {
            c[nt][0] = c[nt][2] = bias[nt * 8 + 2 * t];
            c[nt][1] = c[nt][3] = bias[nt * 8 + 2 * t + 1];
        }
        mlp_layer<3, 4, 56>(c, a3, mlp + ER_T_DEF1, lane);
        uint32_t ah[2][4];
        acc_to_a<4, true>(c, ah);
        zero_acc(c);
        mlp_layer<2, 4, 40>(c, ah, mlp + ER_T_DEF2, lane);
        acc_to_a<4, true>(c, ah);
        float cd[1][4];
        zero_acc(cd);
        mlp_layer<2, 1, 40>(cd, ah, mlp + ER_T_DEF3, lane);
        // dx (cols 0, 1) lives in lanes t == 0
        const int src0 = 4 * (lane & 7);
        const float d0lo = __shfl_sync(0xffffffffu, cd[0][0], src0), d1lo = __shfl_sync(0xffffffffu, cd[0][1], src0);
        const float d0hi = __shfl_sync(0xffffffffu, cd[0][2], src0), d1hi = __shfl_sync(0xffffffffu, cd[0][3], src0);
        if ((lane >> 4) == m) {
            dxy[0] = round_half((lane & 8) ? d0hi : d0lo);
            dxy[1] = round_half((lane & 8) ? d1hi : d1lo);
        }
    }
    // x = clamp(x + dx, -1, 1); tiled grid on (x + 1) / 2 (network.py:188-190), fp16 table
    if (masked) {
        const float u = (clampf_(xin[0] + dxy[0], -1.f, 1.f) + 1.f) * 0.5f;
        const float v = (clampf_(xin[1] + dxy[1], -1.f, 1.f) + 1.f) * 0.5f;
        if (tm.tl.n_dense == 9) torso_gather_t<9>(tm, u, v, row);
        else torso_gather_generic(tm, u, v, row);
    } else {
#pragma unroll
        for (int l = 0; l < 16; l++) *reinterpret_cast<uint32_t *>(row + 2 * l) = 0u;
    }
    __syncwarp();
#pragma unroll 1
    for (int m = 0; m < 2; m++) {
        if (((mmask >> (16 * m)) & 0xffffu) == 0u) continue;
        // torso_net: [feat 32 | freq 34 (+ constants as bias)] -> 32 -> 32 -> 4
        uint32_t a5[5][4];
#pragma unroll
        for (int kk = 0; kk < 5; kk++) load_a(a5[kk], xt + m * 16 * TX_STRIDE, TX_STRIDE, kk, lane);
        float c[4][4];
#pragma unroll
        for (int nt = 0; nt < 4; nt++) {
            c[nt][0] = c[nt][2] = bias[32 + nt * 8 + 2 * t];
            c[nt][1] = c[nt][3] = bias[32 + nt * 8 + 2 * t + 1];
        }
        mlp_layer<5, 4, 88>(c, a5, mlp + ER_T_TOR1, lane);
        uint32_t ah[2][4];
        acc_to_a<4, true>(c, ah);
        zero_acc(c);
        mlp_layer<2, 4, 40>(c, ah, mlp + ER_T_TOR2, lane);
        acc_to_a<4, true>(c, ah);
        float co[1][4];
        zero_acc(co);
        mlp_layer<2, 1, 40>(co, ah, mlp + ER_T_TOR3, lane);
        float o[4];
#pragma unroll
        for (int i = 0; i < 4; i++) o[i] = affine16(sigmoid16(round_half(co[0][i])));
        // cols: 0 alpha, 1..3 color -> lanes t == 0 hold (alpha, r), t == 1 hold (g, b)
        const int src0 = 4 * (lane & 7), src1 = src0 + 1;
        const float a_lo = __shfl_sync(0xffffffffu, o[0], src0), a_hi = __shfl_sync(0xffffffffu, o[2], src0);
        const float r_lo = __shfl_sync(0xffffffffu, o[1], src0), r_hi = __shfl_sync(0xffffffffu, o[3], src0);
        const float g_lo = __shfl_sync(0xffffffffu, o[0], src1), g_hi = __shfl_sync(0xffffffffu, o[2], src1);
        const float b_lo = __shfl_sync(0xffffffffu, o[1], src1), b_hi = __shfl_sync(0xffffffffu, o[3], src1);
        if ((lane >> 4) == m && masked) {
            const bool hi = lane & 8;
            alpha = hi ? a_hi : a_lo;
            tc0 = hi ? r_hi : r_lo;
            tc1 = hi ? g_hi : g_lo;
            tc2 = hi ? b_hi : b_lo;
        }
    }
    __syncwarp();
}

    if (valid) {
        // renderer.py:344 bg = torso_color * torso_alpha + bg * (1 - torso_alpha)
        float bg[3] = {1.f, 1.f, 1.f};
        if (tf.bg_color) {
            bg[0] = __half2float(tf.bg_color[pix * 3]);
            bg[1] = __half2float(tf.bg_color[pix * 3 + 1]);
            bg[2] = __half2float(tf.bg_color[pix * 3 + 2]);
        }
        const float tc[3] = {tc0, tc1, tc2};
#pragma unroll
        for (int k = 0; k < 3; k++) bgc[k] = tc[k] * alpha + bg[k] * (1 - alpha);
        if (tf.dbg_mask) tf.dbg_mask[pix] = masked ? 1 : 0;
    }
}

// per-frame torso constants: enc_anchor = freq(wrapped_anchor, deg 3) (network.py:177), then the contribution of
// [enc_anchor(42) | ind_code_torso(8)] to the first layer of both torso MLPs; called by threads 0..63 after `anchor` is filled
__device__ __forceinline__ float torso_bias_elem(const __half *tconst_sm, const float *anchor_sm, const float *misc, int n) {
    const __half *Wc = tconst_sm + n * 50;
    float acc = 0.f;
    for (int k = 0; k < 42; k++) acc = fmaf(__half2float(Wc[k]), round_half(anchor_sm[k]), acc);
    for (int k = 0; k < 8; k++) acc = fmaf(__half2float(Wc[42 + k]), round_half(__ldg(misc + 16 + k)), acc);
    return acc;
}

// =========================================================================================
// k_head
// =========================================================================================
#ifndef HEAD_THREADS
#define HEAD_THREADS 512
#endif
#define HEAD_WARPS (HEAD_THREADS / 32)
#define XS_STRIDE 56 /* enc_x tile row stride in halfs (48 + 8) */
#define SH_STRIDE 24 /* SH tile row stride in halfs (16 + 8) */

// per-frame part of the k_head arguments: one launch renders up to HEAD_MAX_FRAMES frames of DIFFERENT sessions (contexts loaded
// from the same blob); their hit lists form one queue, a warp's rays may belong to different frames
#define HEAD_MAX_FRAMES 4
struct HeadFrame {
    FrameGeom g;
    const float *state;  // enc_a at [0..31]
    float eye;
    const int *hits;
    int *counters;
    const float *rays_t, *fars;
    float *weights_sum, *image;
    float4 *snap;
};

struct HeadParams {
    HeadLevels hl;
    const float *planes;
    uint32_t plane_rows;
    const uint8_t *bitfield;
    const uint8_t *bitfield_linear;   // nullable
    const __half *mlp_image;
    float bound, min_near, dt_gamma, T_thresh;
    uint32_t max_steps, cascade, grid_size;
    int n_frames;
    MarchParams mp;
    HeadFrame f[HEAD_MAX_FRAMES];
};

struct HeadWarpSmem {   // one per warp
    alignas(16) __half xs[32 * XS_STRIDE];
    alignas(16) __half sh[32 * SH_STRIDE];
    float ray[9][32];                  // per lane: origin, direction, 1 / direction (kept across the passes of a ray)
};

struct HeadSmem {
    alignas(16) unsigned char mlp[ER_H_BYTES];
    alignas(16) HeadWarpSmem w[HEAD_WARPS];
    float enc_a[HEAD_MAX_FRAMES][32];
    float eye[HEAD_MAX_FRAMES];
    int end[HEAD_MAX_FRAMES];          // cumulative hit-list lengths
    int hist[HEAD_MAX_FRAMES][ER_MAX_STEPS + 1];
    int samples[HEAD_MAX_FRAMES];
    int passes, tiles;                 // telemetry: warp passes / MLP tiles of this CTA
    alignas(8) uint64_t bar;
};

__device__ __forceinline__ void gen_ray(const FrameGeom &g, int idx, Ray &r) {
    if (g.rays_o) {
        r.ox = g.rays_o[idx * 3]; r.oy = g.rays_o[idx * 3 + 1]; r.oz = g.rays_o[idx * 3 + 2];
        r.dx = g.rays_d[idx * 3]; r.dy = g.rays_d[idx * 3 + 1]; r.dz = g.rays_d[idx * 3 + 2];
    } else {
        // utils.py:274-277,318-328: i = col + 0.5, j = row + 0.5; (i - cx) / fx on CUDA is a multiply
        // by the fp32 reciprocal; directions normalised, then @ R^T.  Evaluation order pinned bit for bit on the
        // reference's own get_rays on the B200 (tests/test_ernerf_reference_render_gpu.py, scripts/diag_rays.py):
        // torch.norm over the 3-vector reduces as (x*x + z*z) + y*y with separately rounded squares (two threads
        // split the reduction: {x, z} and {y}); the [N,3] @ [3,3] sgemm accumulates k = 0, 1, 2 with FMAs.
        const int row = idx / g.W, col = idx - row * g.W;
        const float xs = ((float)col + 0.5f - g.cx) * g.inv_fx;
        const float ys = ((float)row + 0.5f - g.cy) * g.inv_fy;
        const float nrm = sqrtf(__fadd_rn(__fadd_rn(__fmul_rn(xs, xs), 1.0f), __fmul_rn(ys, ys)));
        const float d0 = xs / nrm, d1 = ys / nrm, d2 = 1.0f / nrm;
        r.dx = fmaf(d2, g.R[2], fmaf(d1, g.R[1], d0 * g.R[0]));
        r.dy = fmaf(d2, g.R[5], fmaf(d1, g.R[4], d0 * g.R[3]));
        r.dz = fmaf(d2, g.R[8], fmaf(d1, g.R[7], d0 * g.R[6]));
        r.ox = g.T[0]; r.oy = g.T[1]; r.oz = g.T[2];
    }
    r.rdx = 1 / r.dx; r.rdy = 1 / r.dy; r.rdz = 1 / r.dz;
}

// density + color for one 16-row tile (rows m*16..m*16+15 of the warp's 32-sample tile).
// Returns in lane (t == 0): sigma logit of rows g / g+8; rgb LOGITS (fp16-rounded) in lanes t == 0 (r, g) and t == 1 (b).
// ea_lo / eye_lo belong to the frame of row g, ea_hi / eye_hi to that of row g + 8 (a tile may mix sessions).
__device__ __forceinline__ void head_mlp_tile(const unsigned char *mlp, const float *ea_lo, const float *ea_hi, __half *xs, const __half *sh,
                                              int m, int lane, float eye_lo, float eye_hi, float &sig_lo, float &sig_hi, float (&rgb)[4]) {
    const __half *W = reinterpret_cast<const __half *>(mlp);
    const float *colbias = reinterpret_cast<const float *>(mlp + ER_H_COLBIAS_BYTES);
    const int g = lane >> 2, t = lane & 3;
    __half *xt = xs + m * 16 * XS_STRIDE;

    uint32_t ax[3][4];
#pragma unroll
    for (int kk = 0; kk < 3; kk++) load_a(ax[kk], xt, XS_STRIDE, kk, lane);

    uint32_t aw[2][4];  // enc_w = enc_a * aud_ch_att  -> sigma_net K cols 48..79
    {
        float c1[8][4];
        zero_acc(c1);
        mlp_layer<3, 8, 56>(c1, ax, W + ER_H_AUD1, lane);
        uint32_t ah[4][4];
        acc_to_a<8, true>(c1, ah);
        float c2[4][4];
        zero_acc(c2);
        mlp_layer<4, 4, 72>(c2, ah, W + ER_H_AUD2, lane);
#pragma unroll
        for (int j = 0; j < 2; j++)
#pragma unroll
            for (int h = 0; h < 2; h++) {
                const int col = (2 * j + h) * 8 + 2 * t;
                aw[j][2 * h + 0] = pack_half2(ea_lo[col] * round_half(c2[2 * j + h][0]), ea_lo[col + 1] * round_half(c2[2 * j + h][1]));
                aw[j][2 * h + 1] = pack_half2(ea_hi[col] * round_half(c2[2 * j + h][2]), ea_hi[col + 1] * round_half(c2[2 * j + h][3]));
            }
    }
    {   // eye_att = sigmoid(MLP(36 -> 16 -> 1)); e = eye * eye_att -> column 36 of the enc_x block
        float ce[2][4];
        zero_acc(ce);
        mlp_layer<3, 2, 56>(ce, ax, W + ER_H_EYE1, lane);
        const __half *w2 = W + ER_H_EYE2;
        float lo = 0.f, hi = 0.f;
#pragma unroll
        for (int nt = 0; nt < 2; nt++) {
            const float w0 = __half2float(w2[nt * 8 + 2 * t]), w1 = __half2float(w2[nt * 8 + 2 * t + 1]);
            lo = fmaf(round_half(fmaxf(ce[nt][0], 0.f)), w0, lo);
            lo = fmaf(round_half(fmaxf(ce[nt][1], 0.f)), w1, lo);
            hi = fmaf(round_half(fmaxf(ce[nt][2], 0.f)), w0, hi);
            hi = fmaf(round_half(fmaxf(ce[nt][3], 0.f)), w1, hi);
        }
        lo += __shfl_xor_sync(0xffffffffu, lo, 1);
        lo += __shfl_xor_sync(0xffffffffu, lo, 2);
        hi += __shfl_xor_sync(0xffffffffu, hi, 1);
        hi += __shfl_xor_sync(0xffffffffu, hi, 2);
        if (t == 0) {
            xt[g * XS_STRIDE + 36] = __float2half_rn(eye_lo * sigmoid16(round_half(lo)));
            xt[(g + 8) * XS_STRIDE + 36] = __float2half_rn(eye_hi * sigmoid16(round_half(hi)));
        }
        __syncwarp();
        load_a(ax[2], xt, XS_STRIDE, 2, lane);
    }
    uint32_t ag[4][4];  // geo_feat (fp16)
    {
        uint32_t a5[5][4];
#pragma unroll
        for (int i = 0; i < 4; i++) {
            a5[0][i] = ax[0][i]; a5[1][i] = ax[1][i]; a5[2][i] = ax[2][i];
            a5[3][i] = aw[0][i]; a5[4][i] = aw[1][i];
        }
        float c[8][4];
        zero_acc(c);
        mlp_layer<5, 8, 88>(c, a5, W + ER_H_SIG1, lane);
        uint32_t ah[4][4];
        acc_to_a<8, true>(c, ah);
        zero_acc(c);
        mlp_layer<4, 8, 72>(c, ah, W + ER_H_SIG2, lane);
        acc_to_a<8, true>(c, ah);
        zero_acc(c);
        mlp_layer<4, 8, 72>(c, ah, W + ER_H_SIG3, lane);
        acc_to_a<8, false>(c, ag);
        float cs[1][4];
        zero_acc(cs);
        mlp_layer<4, 1, 72>(cs, ah, W + ER_H_SIG3 + 64 * 72, lane);
        sig_lo = round_half(cs[0][0]);  // valid in lanes t == 0 (column 0 of the 9th n-tile)
        sig_hi = round_half(cs[0][2]);
    }
    {   // color_net: [geo 64 | SH 16] -> 64 -> 3, individual code folded into the accumulator init
        uint32_t a5[5][4];
#pragma unroll
        for (int i = 0; i < 4; i++) { a5[0][i] = ag[0][i]; a5[1][i] = ag[1][i]; a5[2][i] = ag[2][i]; a5[3][i] = ag[3][i]; }
        load_a(a5[4], sh + m * 16 * SH_STRIDE, SH_STRIDE, 0, lane);
        float c[8][4];
#pragma unroll
        for (int nt = 0; nt < 8; nt++) {
            c[nt][0] = c[nt][2] = colbias[nt * 8 + 2 * t];
            c[nt][1] = c[nt][3] = colbias[nt * 8 + 2 * t + 1];
        }
        mlp_layer<5, 8, 88>(c, a5, W + ER_H_COL1, lane);
        uint32_t ah[4][4];
        acc_to_a<8, true>(c, ah);
        float cc[1][4];
        zero_acc(cc);
        mlp_layer<4, 1, 72>(cc, ah, W + ER_H_COL2, lane);
        // the colour LOGITS: sigmoid + the affine map are applied by the caller once each value sits in its sample's lane (3 per lane
        // and pass instead of 4 per lane and tile here, most of them on padding columns)
#pragma unroll
        for (int i = 0; i < 4; i++) rgb[i] = round_half(cc[0][i]);
    }
}

// enc_x of one sample -> fp16 row of the warp's tile.  ND = number of leading dense levels (the shipped
// architecture has 4: base 64, 512 desired, 2^14 hash rows); ND < 0 = generic indexing.

// enc_x of one sample -> fp16 row of the warp's tile.  ND = number of leading dense levels (the shipped
// architecture has 4: base 64, 512 desired, 2^14 hash rows); ND < 0 = generic indexing.
// The 12 levels of a plane are gathered GL at a time: indices of GL levels first, then their 4 * GL loads back to back, then the
// bilinear sums -- with one level at a time only 4 loads were in flight per thread and every level exposed a full L1 / L2 latency
// (36 of them per sample; long scoreboard is the kernel's top stall once the barrier wait is gone).  Same arithmetic per level as
// grid_level_f32 (gridencoder.cu:75-175): bit-identical features.
#ifndef ER_GATHER_LEVELS
#define ER_GATHER_LEVELS 6   /* levels whose loads are written back to back in the source (2, 3, 4, 6 measured within 1 %: ptxas schedules the loads itself); must divide 12 */
#endif
template <int ND>
__device__ __forceinline__ void gather_planes_t(const HeadParams &p, float x, float y, float z, __half *row) {
    constexpr int GL = ER_GATHER_LEVELS;
    const float rb = 1.0f / (2.0f * p.bound);
    const float u0 = (x + p.bound) * rb, u1 = (y + p.bound) * rb, u2 = (z + p.bound) * rb;
    // the plane loop is NOT unrolled: fully unrolled, k_head was 168 KB of SASS and its top stall was instruction fetch
    // ("no_instruction" 2.0 per issue, profiles/r02_k_head_v5_ncu_summary.md)
#pragma unroll 1
    for (int pl = 0; pl < 3; pl++) {
        const float a = (pl == 1) ? u1 : u0;
        const float b = (pl == 0) ? u1 : u2;
        const float *tab = p.planes + (size_t)pl * p.plane_rows;   // generic path
        const uint32_t pbase = (uint32_t)pl * p.plane_rows;         // 32-bit element offset of the plane (3 * plane_rows < 2^32: checked at load)
        float f[12];
        if (ND < 0) {
#pragma unroll
            for (int l = 0; l < 12; l++) f[l] = grid_level_f32<IDX_ANY>(tab, p.hl.lv[l], a, b);
        } else if (a < 0.f || a > 1.f || b < 0.f || b > 1.f) {   // grid_level_f32's bounds test, once per plane
#pragma unroll
            for (int l = 0; l < 12; l++) f[l] = 0.f;
        } else {
#pragma unroll
            for (int l0 = 0; l0 < 12; l0 += GL) {
                uint32_t idx[GL][4];
                float pu[GL], pv[GL], v[GL][4];
#pragma unroll
                for (int j = 0; j < GL; j++) {
                    if (l0 + j < ND) grid_level_prep<IDX_DENSE>(p.hl.lv[l0 + j], pbase + p.hl.lv[l0 + j].offset, a, b, idx[j], pu[j], pv[j]);
                    else grid_level_prep<IDX_HASH2>(p.hl.lv[l0 + j], pbase + p.hl.lv[l0 + j].offset, a, b, idx[j], pu[j], pv[j]);
                }
#pragma unroll
                for (int j = 0; j < GL; j++)
#pragma unroll
                    for (int c = 0; c < 4; c++) v[j][c] = __ldg(p.planes + idx[j][c]);
#pragma unroll
                for (int j = 0; j < GL; j++) {
                    float r = 0.f;
                    r = fmaf((1 - pu[j]) * (1 - pv[j]), v[j][0], r);
                    r = fmaf(pu[j] * (1 - pv[j]), v[j][1], r);
                    r = fmaf((1 - pu[j]) * pv[j], v[j][2], r);
                    r = fmaf(pu[j] * pv[j], v[j][3], r);
                    f[l0 + j] = r;
                }
            }
        }
#pragma unroll
        for (int l = 0; l < 6; l++)
            *reinterpret_cast<uint32_t *>(row + pl * 12 + 2 * l) = pack_half2(f[2 * l], f[2 * l + 1]);
    }
#pragma unroll
    for (int i = 18; i < 24; i++) *reinterpret_cast<uint32_t *>(row + 2 * i) = 0u;
}

// any other level structure than the shipped one (4 dense + 8 hashed levels): kept out of line, off the hot code path
__device__ __noinline__ void gather_planes_generic(const HeadParams &p, float x, float y, float z, __half *row) {
    gather_planes_t<-1>(p, x, y, z, row);
}

// CH = samples of a ray shaded side by side in one pass: lane = q * CH + k holds sample k of the warp's ray slot q (32 / CH slots).
// The CH samples of a ray are independent until the compositor (the march never looks at sigma), exactly like the n_step samples
// of a reference round (renderer.py:258-264); samples shaded beyond a ray's exit are discarded by the compositing recurrence.
template <int CH>
__global__ void __launch_bounds__(HEAD_THREADS, 1) k_head(const __grid_constant__ HeadParams p) {
    extern __shared__ __align__(128) unsigned char smem_raw[];
    HeadSmem &sm = *reinterpret_cast<HeadSmem *>(smem_raw);
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int F = p.n_frames;
    // programmatic dependent launch: k_torso_compose may be scheduled onto the SMs that CTAs of this grid leave (its torso pass does not
    // depend on the head; it waits for this grid before compositing), and this grid itself may have started while k_setup drains
    pdl_launch();

    // stage the MLP image with one bulk TMA copy
    if (threadIdx.x == 0) {
        mbar_init(&sm.bar, 1);
        fence_barrier_init();
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        mbar_expect_tx(&sm.bar, ER_H_BYTES);
        bulk_g2s(sm.mlp, p.mlp_image, ER_H_BYTES, &sm.bar);
    }
    for (int i = threadIdx.x; i < HEAD_MAX_FRAMES * (ER_MAX_STEPS + 1); i += HEAD_THREADS) (&sm.hist[0][0])[i] = 0;
    pdl_wait();                    // k_setup has completed: hit lists, counters and the audio feature are visible
    if (threadIdx.x == 0) {
        int end = 0;
        for (int f = 0; f < HEAD_MAX_FRAMES; f++) {
            if (f < F) { end += p.f[f].counters[CT_NHIT]; sm.eye[f] = p.f[f].eye; }
            sm.end[f] = end;
            sm.samples[f] = 0;
        }
        sm.passes = sm.tiles = 0;
    }
    if (threadIdx.x < 32 * F) sm.enc_a[threadIdx.x >> 5][threadIdx.x & 31] = p.f[threadIdx.x >> 5].state[threadIdx.x & 31];
    mbar_wait(&sm.bar, 0);
    __syncthreads();

    const MarchParams &mp = p.mp;                // built on the host: read from the constant bank, not carried in registers
    __half *xs = sm.w[warp].xs;
    __half *sh = sm.w[warp].sh;
    float (*rs)[32] = sm.w[warp].ray;
    const int total = sm.end[HEAD_MAX_FRAMES - 1];
    int *ticket = &p.f[0].counters[CT_TICKET];   // the batch draws its rays from ONE queue (frame 0's ticket)
    const int max_steps = (int)p.max_steps, cap = max_steps + (ER_SNAPS - 1);

    const int k = lane % CH, lead = lane - k;    // sample inside the pass, first lane of the ray slot
    // per-lane ray state (identical in the CH lanes of a slot)
    int ray = -1, fi = 0, cnt = 0, sidx = -1;    // cnt = samples accumulated so far, sidx = snapshot slot once alive after max_steps
    float t = 0.f, far = 0.f, ws = 0.f, cr = 0.f, cg_ = 0.f, cb = 0.f;
    bool exhausted = total == 0;

    while (true) {
        // ---- refill the empty ray slots from the hit list
        const uint32_t empty = __ballot_sync(0xffffffffu, ray < 0 && k == 0);
        if (empty && !exhausted) {
            const int need = __popc(empty);
            int base = 0;
            if (lane == 0) base = atomicAdd(ticket, need);
            base = __shfl_sync(0xffffffffu, base, 0);
            if ((empty >> lead) & 1u) {
                const int idx = base + __popc(empty & ((1u << lead) - 1u));
                if (idx < total) {
                    int f = 0;
                    while (idx >= sm.end[f]) f++;
                    const HeadFrame &fr = p.f[f];
                    ray = fr.hits[idx - (f ? sm.end[f - 1] : 0)];
                    fi = f;
                    t = fr.rays_t[ray];
                    far = fr.fars[ray];
                    ws = cr = cg_ = cb = 0.f;
                    cnt = 0;
                    sidx = -1;
                    Ray ry;
                    gen_ray(fr.g, ray, ry);
                    rs[0][lane] = ry.ox; rs[1][lane] = ry.oy; rs[2][lane] = ry.oz;
                    rs[3][lane] = ry.dx; rs[4][lane] = ry.dy; rs[5][lane] = ry.dz;
                    rs[6][lane] = ry.rdx; rs[7][lane] = ry.rdy; rs[8][lane] = ry.rdz;
                    float shv[16];
                    sh4(ry.dx, ry.dy, ry.dz, shv);
#pragma unroll
                    for (int i = 0; i < 8; i++)
                        *reinterpret_cast<uint32_t *>(sh + lane * SH_STRIDE + 2 * i) = pack_half2(shv[2 * i], shv[2 * i + 1]);
                }
            }
            if (base + need >= total) exhausted = true;
        }
        const bool valid = ray >= 0;
        if (!__any_sync(0xffffffffu, valid)) break;
        const HeadFrame &fr = p.f[fi];

        // ---- march to this lane's sample: k + 1 steps from the ray's t (the same sequence the reference marches)
        float x = 0.f, y = 0.f, z = 0.f, dt = 0.f, tt = t;
        bool has = valid && cnt + k < cap;
        if (has) {
            Ray ry;
            ry.ox = rs[0][lane]; ry.oy = rs[1][lane]; ry.oz = rs[2][lane];
            ry.dx = rs[3][lane]; ry.dy = rs[4][lane]; ry.dz = rs[5][lane];
            ry.rdx = rs[6][lane]; ry.rdy = rs[7][lane]; ry.rdz = rs[8][lane];
            uint32_t vox;
            for (int j = 0; j <= k && has; j++) has = march_next(mp, ry, tt, far, x, y, z, dt, vox);
        }
        const uint32_t hasmask = __ballot_sync(0xffffffffu, has);
        float sigma_logit = 0.f, c0 = 0.f, c1 = 0.f, c2 = 0.f;
        if (hasmask != 0u) {
            // ---- tri-plane gather: 3 planes x 12 levels x 4 corners, fp32 (gridencoder.cu:75-175)
            __half *row = xs + lane * XS_STRIDE;
            if (has) {
                if (p.hl.n_dense == 4) gather_planes_t<4>(p, x, y, z, row);
                else gather_planes_generic(p, x, y, z, row);
            } else {
#pragma unroll
                for (int i = 0; i < 24; i++) *reinterpret_cast<uint32_t *>(row + 2 * i) = 0u;
            }
            __syncwarp();

            // ---- MLPs on two 16-row tensor-core tiles; results routed back to lane == sample
#pragma unroll 1
            for (int m = 0; m < 2; m++) {
                if (((hasmask >> (16 * m)) & 0xffffu) == 0u) continue;
                const int g = lane >> 2;
                const int f_lo = __shfl_sync(0xffffffffu, fi, m * 16 + g), f_hi = __shfl_sync(0xffffffffu, fi, m * 16 + g + 8);
                float slo, shi, rgb[4];
                head_mlp_tile(sm.mlp, sm.enc_a[f_lo], sm.enc_a[f_hi], xs, sh, m, lane, sm.eye[f_lo], sm.eye[f_hi], slo, shi, rgb);
                const int src0 = 4 * (lane & 7), src1 = src0 + 1;
                const float v_slo = __shfl_sync(0xffffffffu, slo, src0), v_shi = __shfl_sync(0xffffffffu, shi, src0);
                const float r_lo = __shfl_sync(0xffffffffu, rgb[0], src0), r_hi = __shfl_sync(0xffffffffu, rgb[2], src0);
                const float g_lo = __shfl_sync(0xffffffffu, rgb[1], src0), g_hi = __shfl_sync(0xffffffffu, rgb[3], src0);
                const float b_lo = __shfl_sync(0xffffffffu, rgb[0], src1), b_hi = __shfl_sync(0xffffffffu, rgb[2], src1);
                if ((lane >> 4) == m) {
                    const bool hi = lane & 8;
                    sigma_logit = hi ? v_shi : v_slo;
                    c0 = hi ? r_hi : r_lo;
                    c1 = hi ? g_hi : g_lo;
                    c2 = hi ? b_hi : b_lo;
                }
            }
            c0 = affine16(sigmoid16(c0)); c1 = affine16(sigmoid16(c1)); c2 = affine16(sigmoid16(c2));   // network.py:296-297 (fp16 ops)
            __syncwarp();
            for (int f = 0; f < F; f++) {
                const int c = __popc(__ballot_sync(0xffffffffu, has && fi == f));
                if (lane == 0 && c) atomicAdd(&sm.samples[f], c);
            }
            if (lane == 0) {
                atomicAdd(&sm.passes, 1);
                atomicAdd(&sm.tiles, ((hasmask & 0xffffu) != 0u) + ((hasmask >> 16) != 0u));
            }
        }

        // ---- composite (raymarching.cu:2189-2218): the recurrence over the ray's samples, in order; every lane of a slot
        // replays it from the shuffled per-sample values (identical arithmetic), the slot's first lane writes
        const float my_alpha = has ? 1.0f - __expf(-expf(sigma_logit) * dt) : 0.f;   // torch.exp runs in fp32 under autocast
        bool done = !valid;
        int life = max_steps;
#pragma unroll
        for (int j = 0; j < CH; j++) {
            const int src = lead + j;
            const bool h_j = (hasmask >> src) & 1u;
            const float a_j = __shfl_sync(0xffffffffu, my_alpha, src);
            const float r_j = __shfl_sync(0xffffffffu, c0, src), g_j = __shfl_sync(0xffffffffu, c1, src);
            const float b_j = __shfl_sync(0xffffffffu, c2, src);
            if (!done) {
                if (cnt >= cap) {
                    done = true;             // every sample a reference round structure could reach has been accumulated
                } else if (!h_j) {
                    done = true;             // deltas[0] == 0 in the reference compositor: the ray ran out of samples
                    life = min(cnt, max_steps);
                } else {
                    const float T = 1 - ws;
                    const float weight = a_j * T;
                    ws += weight;
                    cr = fmaf(weight, r_j, cr);
                    cg_ = fmaf(weight, g_j, cg_);
                    cb = fmaf(weight, b_j, cb);
                    cnt++;
                    if (T < p.T_thresh) {
                        done = true;         // the ray is dropped AFTER this sample was accumulated
                        life = min(cnt - 1, max_steps);
                    }
                    if (cnt >= max_steps && k == 0) {   // the states a later cut-off C = max_steps + j may select
                        if (cnt == max_steps && !done) sidx = atomicAdd(&fr.counters[CT_NSURV], 1);
                        if (sidx >= 0) fr.snap[(size_t)sidx * ER_SNAPS + (cnt - max_steps)] = make_float4(ws, cr, cg_, cb);
                    }
                }
            }
        }
        if (!done && cnt >= cap) done = true;
        // t after the ray's last sample of this pass (lane lead + CH - 1 marched all of them)
        const float t_end = __shfl_sync(0xffffffffu, tt, lead + CH - 1);
        if (valid) {
            if (done) {
                if (k == 0) {
                    if (sidx >= 0) {
                        for (int j = cnt - max_steps + 1; j < ER_SNAPS; j++)
                            fr.snap[(size_t)sidx * ER_SNAPS + j] = make_float4(ws, cr, cg_, cb);
                        fr.weights_sum[ray] = -(float)(sidx + 1);   // resolved by k_torso_compose once C is known
                    } else {
                        fr.weights_sum[ray] = ws;
                        fr.image[ray * 3] = cr; fr.image[ray * 3 + 1] = cg_; fr.image[ray * 3 + 2] = cb;
                    }
                    atomicAdd(&sm.hist[fi][life], 1);
                }
                ray = -1;
            } else {
                t = t_end;
            }
        }
        __syncwarp();
    }
    __syncthreads();
    for (int i = threadIdx.x; i < F * (ER_MAX_STEPS + 1); i += HEAD_THREADS) {
        const int f = i / (ER_MAX_STEPS + 1), b = i % (ER_MAX_STEPS + 1);
        if (sm.hist[f][b]) atomicAdd(&p.f[f].counters[CT_HIST + b], sm.hist[f][b]);
    }
    if (threadIdx.x < F && sm.samples[threadIdx.x]) atomicAdd(&p.f[threadIdx.x].counters[CT_SAMPLES], sm.samples[threadIdx.x]);
    if (threadIdx.x == 32 && sm.passes) { atomicAdd(&p.f[0].counters[CT_PASSES], sm.passes); atomicAdd(&p.f[0].counters[CT_TILES], sm.tiles); }
}

__global__ void __launch_bounds__(SETUP_THREADS) k_setup(const __grid_constant__ SetupBatch b) {
    pdl_launch();                  // k_head's CTAs may take the SMs this grid's CTAs leave (they wait for the grid's completion themselves)
    const int n_audio = 8 * b.n;   // one CTA per attention window
    if ((int)blockIdx.x < n_audio) audio_cta(b.f[blockIdx.x >> 3], blockIdx.x & 7);
    else ray_pass(b, (int)blockIdx.x - n_audio, (int)gridDim.x - n_audio);
}

// =========================================================================================
// torso + final compose (renderer.py:275-277,294-352): resolve the loop control, torso over background,
// image + (1 - weights_sum) * bg, clamp, fp32 / u8
#define TORSO_THREADS 256
#define TORSO_WARPS (TORSO_THREADS / 32)
struct ComposeParams {
    FrameGeom g;
    TorsoModel tm;
    TorsoFrame tf;
    float *weights_sum, *image;   // head result per ray; rays alive after max_steps samples hold -(snapshot slot + 1) in weights_sum
    const float4 *snap;
    int *counters;                // this frame's counters: life histogram in, derived rounds out
    int max_steps;
    float *torso_rgb;             // scratch [3][N]
    float *out_f32;               // [N,3] or null
    uint8_t *out_u8;              // [N,3] or null
    float *dbg_image_head;
};
struct TorsoSmem {
    alignas(16) __half mlp[ER_T_HALFS];
    alignas(16) __half tconst[2 * 32 * 50];
    alignas(16) __half xt[TORSO_WARPS][32 * TX_STRIDE];
    float bias[64];
    float anchor[48];
    int hist[ER_MAX_STEPS + 2];   // the frame's life histogram + the hit count, staged for the loop-control replay
    int snap;
};

// The grid is ONE wave (TORSO_MINB CTAs per SM): under programmatic dependent launch its CTAs take the SMs that k_head's CTAs have
// already left and run the torso pass of ALL their tiles -- independent of the head -- while the last rays of the frame are still being
// shaded; the torso-over-background colour waits in an L2-resident scratch (each thread reads back what it wrote).  Then the grid
// waits for k_head, replays the loop control and composes.
__global__ void __launch_bounds__(TORSO_THREADS, TORSO_MINB) k_torso_compose(const __grid_constant__ ComposeParams p) {
    extern __shared__ __align__(128) unsigned char smem_raw[];
    TorsoSmem &sm = *reinterpret_cast<TorsoSmem *>(smem_raw);
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    for (int i = threadIdx.x; i < ER_T_HALFS / 8; i += blockDim.x)
        reinterpret_cast<uint4 *>(sm.mlp)[i] = __ldg(reinterpret_cast<const uint4 *>(p.tm.mlp_image) + i);
    for (int i = threadIdx.x; i < 2 * 32 * 50 / 8; i += blockDim.x)
        reinterpret_cast<uint4 *>(sm.tconst)[i] = __ldg(reinterpret_cast<const uint4 *>(p.tm.torso_const) + i);
    if (threadIdx.x < 42) sm.anchor[threadIdx.x] = freq_elem(p.tf.wa, 6, threadIdx.x);
    __syncthreads();
    if (threadIdx.x < 64) sm.bias[threadIdx.x] = torso_bias_elem(sm.tconst, sm.anchor, p.tm.misc, threadIdx.x);
    __syncthreads();

    const int N = p.g.N;
    const int n_tiles = (N + 31) / 32;
    const int stride = gridDim.x * TORSO_WARPS;
    // ---- torso pass
#pragma unroll 1
    for (int tile = blockIdx.x * TORSO_WARPS + warp; tile < n_tiles; tile += stride) {
        float b[3] = {0.f, 0.f, 0.f};
        torso_tile(p.tm, p.g, p.tf, tile, lane, sm.xt[warp], sm.mlp, sm.bias, b);
        const int pix = tile * 32 + lane;
        if (pix < N) { p.torso_rgb[pix] = b[0]; p.torso_rgb[N + pix] = b[1]; p.torso_rgb[2 * N + pix] = b[2]; }
        __syncwarp();
    }
    // ---- k_head has completed.  The reference's loop control (renderer.py:246-256) replayed on the life histogram it filled:
    // A(0) = N, A(c) = hits - #{rays with life < c}; round r starts at c_r with n_alive = A(c_r), n_step = clamp(N / n_alive, 1, 8);
    // the summed n_step C selects the snapshot of the rays that were still alive after max_steps samples
    pdl_wait();
    if (threadIdx.x <= ER_MAX_STEPS) sm.hist[threadIdx.x] = p.counters[CT_HIST + threadIdx.x];     // one round trip for the histogram
    if (threadIdx.x == ER_MAX_STEPS + 1) sm.hist[ER_MAX_STEPS + 1] = p.counters[CT_NHIT];
    __syncthreads();
    if (threadIdx.x == 0) {
        const int nhit = sm.hist[ER_MAX_STEPS + 1];
        int c = 0, r = 0, gone = 0, upto = 1;
        while (c < p.max_steps && r < ER_MAX_ROUNDS) {
            for (; upto < c; upto++) gone += sm.hist[upto];
            const int n_alive = c == 0 ? N : nhit - gone;
            if (n_alive <= 0) break;
            const int n_step = max(min(N / n_alive, 8), 1);
            if (blockIdx.x == 0) {
                int *row = p.counters + CT_ROUNDS + 4 * r;
                // column 1 carries k_head's telemetry: warp passes (row 0), 16-row MLP tiles (row 1)
                row[0] = n_alive; row[1] = r == 0 ? p.counters[CT_PASSES] : (r == 1 ? p.counters[CT_TILES] : 0);
                row[2] = r == 0 ? nhit : -1; row[3] = n_step;
            }
            c += n_step;
            r++;
        }
        sm.snap = min(max(c - p.max_steps, 0), ER_SNAPS - 1);
    }
    __syncthreads();
    const int s_snap = sm.snap;
    // ---- resolve + compose (renderer.py:275-277)
#pragma unroll 1
    for (int tile = blockIdx.x * TORSO_WARPS + warp; tile < n_tiles; tile += stride) {
        const int pix = tile * 32 + lane;
        if (pix >= N) continue;
        float ws = p.weights_sum[pix];
        const float b[3] = {p.torso_rgb[pix], p.torso_rgb[N + pix], p.torso_rgb[2 * N + pix]};
        float hd[3];
        if (ws < 0.f) {   // alive after max_steps samples: the state after C samples
            const float4 v = p.snap[(size_t)((int)(-ws) - 1) * ER_SNAPS + s_snap];
            ws = v.x; hd[0] = v.y; hd[1] = v.z; hd[2] = v.w;
            p.weights_sum[pix] = ws;
            p.image[pix * 3] = hd[0]; p.image[pix * 3 + 1] = hd[1]; p.image[pix * 3 + 2] = hd[2];
        } else {
            hd[0] = p.image[pix * 3]; hd[1] = p.image[pix * 3 + 1]; hd[2] = p.image[pix * 3 + 2];
        }
        float out[3];
#pragma unroll
        for (int k = 0; k < 3; k++) {
            out[k] = fminf(fmaxf(hd[k] + (1 - ws) * b[k], 0.f), 1.f);
            if (p.dbg_image_head) p.dbg_image_head[pix * 3 + k] = hd[k];
        }
        if (p.out_f32) { p.out_f32[pix * 3] = out[0]; p.out_f32[pix * 3 + 1] = out[1]; p.out_f32[pix * 3 + 2] = out[2]; }
        if (p.out_u8) {
            p.out_u8[pix * 3] = (uint8_t)(out[0] * 255.f);
            p.out_u8[pix * 3 + 1] = (uint8_t)(out[1] * 255.f);
            p.out_u8[pix * 3 + 2] = (uint8_t)(out[2] * 255.f);
        }
    }
}

// F.interpolate(mode='bilinear', align_corners=False) (utils.py:1212) + (x*255).astype(uint8)
__global__ void k_resize_u8(const float *__restrict__ in, int H, int W, int outH, int outW, float *out_f32,
                            uint8_t *out_u8) {
    const int idx = blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= outH * outW) return;
    const int oy = idx / outW, ox = idx - oy * outW;
    const float sy = fmaxf(((float)H / outH) * (oy + 0.5f) - 0.5f, 0.f);
    const float sx = fmaxf(((float)W / outW) * (ox + 0.5f) - 0.5f, 0.f);
    const int y0 = min((int)sy, H - 1), x0 = min((int)sx, W - 1);
    const int y1 = min(y0 + 1, H - 1), x1 = min(x0 + 1, W - 1);
    const float ly = sy - y0, lx = sx - x0;
#pragma unroll
    for (int k = 0; k < 3; k++) {
        const float v00 = in[(y0 * W + x0) * 3 + k], v01 = in[(y0 * W + x1) * 3 + k];
        const float v10 = in[(y1 * W + x0) * 3 + k], v11 = in[(y1 * W + x1) * 3 + k];
        const float v = (1 - ly) * ((1 - lx) * v00 + lx * v01) + ly * ((1 - lx) * v10 + lx * v11);
        if (out_f32) out_f32[idx * 3 + k] = v;
        if (out_u8) out_u8[idx * 3 + k] = (uint8_t)(v * 255.f);
    }
}

// =========================================================================================
// kernel-level drop-ins (same device functions as the fused path)
// =========================================================================================
__global__ void k_near_far(const float *__restrict__ rays_o, const float *__restrict__ rays_d,
                           const float *__restrict__ aabb, uint32_t N, float min_near, float *nears, float *fars) {
    const uint32_t n = threadIdx.x + blockIdx.x * blockDim.x;
    if (n >= N) return;
    float ab[6];
#pragma unroll
    for (int i = 0; i < 6; i++) ab[i] = aabb[i];
    near_far_aabb(rays_o[n * 3], rays_o[n * 3 + 1], rays_o[n * 3 + 2], rays_d[n * 3], rays_d[n * 3 + 1],
                  rays_d[n * 3 + 2], ab, min_near, nears[n], fars[n]);
}

__global__ void k_march_rays(uint32_t n_alive, uint32_t n_step, const int *__restrict__ rays_alive,
                             const float *__restrict__ rays_t, const float *__restrict__ rays_o,
                             const float *__restrict__ rays_d, float bound, float dt_gamma, uint32_t max_steps,
                             uint32_t C, uint32_t H, const uint8_t *__restrict__ grid,
                             const float *__restrict__ nears, const float *__restrict__ fars, float *xyzs,
                             float *dirs, float *deltas, const float *__restrict__ noises) {
    const uint32_t n = threadIdx.x + blockIdx.x * blockDim.x;
    if (n >= n_alive) return;
    const int index = rays_alive[n];
    const float noise = noises ? noises[n] : 0.f;
    Ray r;
    r.ox = rays_o[index * 3]; r.oy = rays_o[index * 3 + 1]; r.oz = rays_o[index * 3 + 2];
    r.dx = rays_d[index * 3]; r.dy = rays_d[index * 3 + 1]; r.dz = rays_d[index * 3 + 2];
    r.rdx = 1 / r.dx; r.rdy = 1 / r.dy; r.rdz = 1 / r.dz;
    const MarchParams mp = make_march_params(bound, dt_gamma, max_steps, C, H, grid);
    xyzs += (size_t)n * n_step * 3;
    dirs += (size_t)n * n_step * 3;
    deltas += (size_t)n * n_step * 2;
    float t = rays_t[index];
    const float far = fars[index];
    (void)nears;
    t += clampf_(t * dt_gamma, mp.dt_min, mp.dt_max) * noise;
    for (uint32_t step = 0; step < n_step; step++) {
        float x, y, z, dt;
        uint32_t vox;
        if (!march_next(mp, r, t, far, x, y, z, dt, vox)) break;
        xyzs[0] = x; xyzs[1] = y; xyzs[2] = z;
        dirs[0] = r.dx; dirs[1] = r.dy; dirs[2] = r.dz;
        deltas[0] = dt; deltas[1] = t;
        xyzs += 3; dirs += 3; deltas += 2;
    }
}

__global__ void k_composite_rays_triplane(uint32_t n_alive, uint32_t n_step, float T_thresh, int *rays_alive,
                                          float *rays_t, const float *__restrict__ sigmas,
                                          const float *__restrict__ rgbs, const float *__restrict__ deltas,
                                          const float *__restrict__ ambs_aud, const float *__restrict__ ambs_eye,
                                          const float *__restrict__ uncertainties, float *weights_sum, float *depth,
                                          float *image, float *amb_aud_sum, float *amb_eye_sum,
                                          float *uncertainty_sum) {
    const uint32_t n = threadIdx.x + blockIdx.x * blockDim.x;
    if (n >= n_alive) return;
    const int index = rays_alive[n];
    sigmas += (size_t)n * n_step; rgbs += (size_t)n * n_step * 3; deltas += (size_t)n * n_step * 2;
    if (ambs_aud) ambs_aud += (size_t)n * n_step;
    if (ambs_eye) ambs_eye += (size_t)n * n_step;
    if (uncertainties) uncertainties += (size_t)n * n_step;
    float t = rays_t[index];
    float weight_sum = weights_sum[index], d = depth ? depth[index] : 0.f;
    float r = image[index * 3], g = image[index * 3 + 1], b = image[index * 3 + 2];
    float a_aud = amb_aud_sum ? amb_aud_sum[index] : 0.f, a_eye = amb_eye_sum ? amb_eye_sum[index] : 0.f;
    float u = uncertainty_sum ? uncertainty_sum[index] : 0.f;
    uint32_t step = 0;
    while (step < n_step) {
        if (deltas[0] == 0) break;
        const float alpha = 1.0f - __expf(-sigmas[0] * deltas[0]);
        const float T = 1 - weight_sum;
        const float weight = alpha * T;
        weight_sum += weight;
        t = deltas[1];
        d += weight * t;
        r += weight * rgbs[0];
        g += weight * rgbs[1];
        b += weight * rgbs[2];
        if (ambs_aud) a_aud += ambs_aud[0];
        if (ambs_eye) a_eye += ambs_eye[0];
        if (uncertainties) u += weight * uncertainties[0];
        if (T < T_thresh) break;
        sigmas++; rgbs += 3; deltas += 2; step++;
        if (ambs_aud) ambs_aud++;
        if (ambs_eye) ambs_eye++;
        if (uncertainties) uncertainties++;
    }
    if (step < n_step) rays_alive[n] = -1;
    else rays_t[index] = t;
    weights_sum[index] = weight_sum;
    if (depth) depth[index] = d;
    image[index * 3] = r; image[index * 3 + 1] = g; image[index * 3 + 2] = b;
    if (amb_aud_sum) amb_aud_sum[index] = a_aud;
    if (amb_eye_sum) amb_eye_sum[index] = a_eye;
    if (uncertainty_sum) uncertainty_sum[index] = u;
}

struct GridLevelsAny { GridLevel lv[16]; };

__global__ void k_grid_encode(const float *__restrict__ inputs, const void *__restrict__ table, GridLevelsAny lv,
                              void *outputs, uint32_t B, uint32_t C, int is_half) {
    const uint32_t b = blockIdx.x * blockDim.x + threadIdx.x;
    if (b >= B) return;
    const uint32_t level = blockIdx.y;
    const float u = inputs[b * 2], v = inputs[b * 2 + 1];
    if (!is_half) {  // C == 1
        reinterpret_cast<float *>(outputs)[(size_t)level * B + b] =
            grid_level_f32<IDX_ANY>(reinterpret_cast<const float *>(table), lv.lv[level], u, v);
    } else {  // C == 2
        float o0, o1;
        grid_level_f16x2<IDX_ANY>(reinterpret_cast<const __half2 *>(table), lv.lv[level], u, v, o0, o1);
        reinterpret_cast<__half2 *>(outputs)[(size_t)level * B + b] = __floats2half2_rn(o0, o1);
    }
}

__global__ void k_sh4(const float *__restrict__ inputs, float *outputs, uint32_t B) {
    const uint32_t b = threadIdx.x + blockIdx.x * blockDim.x;
    if (b >= B) return;
    float o[16];
    sh4(inputs[b * 3], inputs[b * 3 + 1], inputs[b * 3 + 2], o);
#pragma unroll
    for (int i = 0; i < 16; i++) outputs[(size_t)b * 16 + i] = o[i];
}

__global__ void k_freq(const float *__restrict__ inputs, uint32_t B, uint32_t D, uint32_t C, float *outputs) {
    const uint32_t t = threadIdx.x + blockIdx.x * blockDim.x;
    if (t >= B * C) return;
    const uint32_t b = t / C, c = t - b * C;
    outputs[t] = freq_elem(inputs + (size_t)b * D, D, c);
}

// =========================================================================================
// host side
// =========================================================================================
static int compute_scales(mf_ctx *ctx, float S, uint32_t H, int L, float *host_out) {
    float *d = nullptr;
    MF_CUDA(ctx, cudaMalloc(&d, 32 * sizeof(float)));
    k_level_scales<<<1, 32>>>(S, H, L, d);
    cudaError_t e = cudaMemcpy(host_out, d, L * sizeof(float), cudaMemcpyDeviceToHost);
    cudaFree(d);
    if (e != cudaSuccess) return mf_fail(ctx, MF_E_CUDA, "level scales: %s", cudaGetErrorString(e));
    return MF_OK;
}

extern "C" int mf_ernerf_blob_layout(int32_t *out, int n) {
    const int32_t v[] = {ER_H_AUD1, ER_H_AUD2, ER_H_EYE1, ER_H_SIG1, ER_H_SIG2, ER_H_SIG3, ER_H_COL1, ER_H_COL2,
                         ER_H_EYE2, ER_H_HALFS, ER_H_COLBIAS_BYTES, ER_H_BYTES,
                         ER_T_DEF1, ER_T_DEF2, ER_T_DEF3, ER_T_TOR1, ER_T_TOR2, ER_T_TOR3, ER_T_HALFS, ER_T_BYTES};
    const int cnt = (int)(sizeof(v) / sizeof(v[0]));
    for (int i = 0; i < cnt && i < n; i++) out[i] = v[i];
    return cnt;
}

static void ernerf_free_ws(ErnerfState *s) {
    cudaFree(s->hits); cudaFree(s->snap); cudaFree(s->rays_t); cudaFree(s->fars);
    cudaFree(s->weights_sum); cudaFree(s->image); cudaFree(s->final_f32); cudaFree(s->torso_rgb);
    s->hits = nullptr;
    s->snap = nullptr;
    s->rays_t = s->fars = s->weights_sum = s->image = s->final_f32 = s->torso_rgb = nullptr;
    s->capN = 0;
}

void ernerf_destroy(mf_ctx *ctx) {
    ErnerfState *s = ctx->ernerf;
    if (!s) return;
    ernerf_free_ws(s);
    cudaFree(s->state);
    cudaFree(s->counters);
    cudaFree(s->bitfield_linear);
    if (s->ev_head[0]) { cudaEventDestroy(s->ev_head[0]); cudaEventDestroy(s->ev_head[1]); }
    delete s;
    ctx->ernerf = nullptr;
}

extern "C" int mf_ernerf_load(mf_ctx *ctx, const void *blob, size_t nbytes, const mf_ernerf_cfg *cfg) {
    if (!ctx) return MF_E_INVALID;
    MF_REQUIRE(ctx, blob && cfg, "mf_ernerf_load: null blob/cfg");
    MF_REQUIRE(ctx, cfg->cascade == 1, "mf_ernerf_load: only cascade == 1 (bound <= 1) is implemented");
    MF_REQUIRE(ctx, cfg->max_steps >= 1 && cfg->max_steps <= ER_MAX_STEPS, "mf_ernerf_load: max_steps %u outside 1..%d", cfg->max_steps,
               ER_MAX_STEPS);
    MF_REQUIRE(ctx, cfg->audio_in_dim >= 1 && cfg->audio_in_dim <= 64, "mf_ernerf_load: audio_in_dim %u unsupported",
               cfg->audio_in_dim);
    MF_CUDA(ctx, cudaSetDevice(ctx->device));
    ernerf_destroy(ctx);
    // header
    std::vector<unsigned char> head(sizeof(mf_blob_header) + MF_BLOB_MAX_ENTRIES * sizeof(mf_blob_entry));
    const size_t hbytes = std::min(head.size(), nbytes);
    MF_CUDA(ctx, cudaMemcpy(head.data(), blob, hbytes, cudaMemcpyDeviceToHost));
    const mf_blob_header *h = reinterpret_cast<const mf_blob_header *>(head.data());
    MF_REQUIRE(ctx, hbytes >= sizeof(mf_blob_header) && h->magic == MF_BLOB_MAGIC && h->kind == 1,
               "mf_ernerf_load: not an ErNeRF blob");
    MF_REQUIRE(ctx, h->n_entries <= MF_BLOB_MAX_ENTRIES, "mf_ernerf_load: too many entries");
    const mf_blob_entry *ent = reinterpret_cast<const mf_blob_entry *>(head.data() + sizeof(mf_blob_header));
    ErnerfState *s = new (std::nothrow) ErnerfState();
    MF_REQUIRE(ctx, s, "mf_ernerf_load: out of host memory");
    s->cfg = *cfg;
    const unsigned char *base = reinterpret_cast<const unsigned char *>(blob);
    size_t sz[16] = {0};
    const void *ptr[16] = {nullptr};
    for (uint32_t i = 0; i < h->n_entries; i++) {
        if (ent[i].id < 16) {
            if (ent[i].offset + ent[i].nbytes > nbytes) {
                delete s;
                return mf_fail(ctx, MF_E_INVALID, "mf_ernerf_load: entry %u out of range", ent[i].id);
            }
            ptr[ent[i].id] = base + ent[i].offset;
            sz[ent[i].id] = ent[i].nbytes;
        }
    }
    const uint32_t head_rows = (uint32_t)cfg->head_offsets[MF_ERNERF_HEAD_LEVELS];
    const uint32_t torso_rows = (uint32_t)cfg->torso_offsets[MF_ERNERF_TORSO_LEVELS];
    const size_t G = cfg->grid_size;
    const size_t audio_halfs = (size_t)32 * cfg->audio_in_dim * 3 + 32 + 32 * 32 * 3 + 32 + 64 * 32 * 3 + 64 +
                               64 * 64 * 3 + 64 + 64 * 64 + 64 + 32 * 64 + 32 + 16 * 32 * 3 + 16 + 8 * 16 * 3 + 8 +
                               4 * 8 * 3 + 4 + 2 * 4 * 3 + 2 + 1 * 2 * 3 + 1 + 64 + 8;
    bool ok = (uint64_t)3 * head_rows < (1ull << 32) &&   // k_head indexes the three planes with 32-bit element offsets
              sz[ER_ID_HEAD_PLANES] == (size_t)3 * head_rows * 4 && sz[ER_ID_BITFIELD] == G * G * G / 8 &&
              sz[ER_ID_TORSO_TABLE] == (size_t)torso_rows * 4 && sz[ER_ID_TORSO_DENSITY] == G * G * 4 &&
              sz[ER_ID_HEAD_MLP] == ER_H_BYTES && sz[ER_ID_TORSO_MLP] == ER_T_BYTES &&
              sz[ER_ID_AUDIO] == audio_halfs * 2 && sz[ER_ID_MISC] == 24 * 4 && sz[ER_ID_TORSO_CONST] == 2 * 32 * 50 * 2;
    if (!ok) {
        delete s;
        return mf_fail(ctx, MF_E_INVALID, "mf_ernerf_load: blob entry sizes do not match cfg (strict loader)");
    }
    s->audio_halfs = audio_halfs;
    s->planes = (const float *)ptr[ER_ID_HEAD_PLANES];
    s->plane_rows = head_rows;
    s->bitfield = (const uint8_t *)ptr[ER_ID_BITFIELD];
    s->torso_table = (const __half2 *)ptr[ER_ID_TORSO_TABLE];
    s->torso_density = (const float *)ptr[ER_ID_TORSO_DENSITY];
    s->head_mlp = (const __half *)ptr[ER_ID_HEAD_MLP];
    s->torso_mlp = (const __half *)ptr[ER_ID_TORSO_MLP];
    s->audio = (const __half *)ptr[ER_ID_AUDIO];
    s->misc = (const float *)ptr[ER_ID_MISC];
    s->torso_const = (const __half *)ptr[ER_ID_TORSO_CONST];
    ctx->ernerf = s;
    MF_CUDA(ctx, cudaMemcpy(s->misc_host, s->misc, sizeof(s->misc_host), cudaMemcpyDeviceToHost));

    float sc[16];
    int rc = compute_scales(ctx, cfg->head_log2_scale, cfg->head_base, MF_ERNERF_HEAD_LEVELS, sc);
    if (rc) return rc;
    fill_levels(s->hl.lv, MF_ERNERF_HEAD_LEVELS, sc, cfg->head_offsets, 0);
    s->hl.n_dense = classify_levels(s->hl.lv, MF_ERNERF_HEAD_LEVELS, IDX_HASH2);
    rc = compute_scales(ctx, cfg->torso_log2_scale, cfg->torso_base, MF_ERNERF_TORSO_LEVELS, sc);
    if (rc) return rc;
    fill_levels(s->tl.lv, MF_ERNERF_TORSO_LEVELS, sc, cfg->torso_offsets, 1);
    s->tl.n_dense = classify_levels(s->tl.lv, MF_ERNERF_TORSO_LEVELS, IDX_TILE2);

    MF_CUDA(ctx, cudaMalloc(&s->state, ST_FLOATS * sizeof(float)));
    MF_CUDA(ctx, cudaMemset(s->state, 0, ST_FLOATS * sizeof(float)));
    if (cfg->cascade == 1 && (G & (G - 1)) == 0 && G >= 32) {
        MF_CUDA(ctx, cudaMalloc(&s->bitfield_linear, G * G * G / 8));
        k_linear_bitfield<<<(unsigned)((G * G * G / 32 + 255) / 256), 256>>>(s->bitfield, (uint32_t)G, reinterpret_cast<uint32_t *>(s->bitfield_linear));
    }
    MF_CUDA(ctx, cudaMalloc(&s->counters, 2 * CT_INTS * sizeof(int)));
    MF_CUDA(ctx, cudaMemset(s->counters, 0, 2 * CT_INTS * sizeof(int)));

    MF_CUDA(ctx, cudaFuncSetAttribute(k_torso_compose, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(TorsoSmem)));
    MF_CUDA(ctx, cudaFuncSetAttribute(k_head<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(HeadSmem)));
    MF_CUDA(ctx, cudaFuncSetAttribute(k_head<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(HeadSmem)));
    MF_CUDA(ctx, cudaFuncSetAttribute(k_head<4>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(HeadSmem)));
    MF_CUDA(ctx, cudaFuncSetAttribute(k_head<8>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(HeadSmem)));
    int per_sm = 0;
    MF_CUDA(ctx, cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, k_head<2>, HEAD_THREADS, sizeof(HeadSmem)));
    MF_REQUIRE(ctx, per_sm >= 1, "k_head does not fit on an SM");
    MF_CUDA(ctx, cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, k_setup, SETUP_THREADS, 0));
    s->setup_per_sm = std::max(1, std::min(per_sm, 2));
    if (const char *e = getenv("MF_SETUP_PER_SM")) s->setup_per_sm = std::max(1, std::min(atoi(e), 2));
    {   // experiment hook: MF_HEAD_CHUNK = samples of a ray shaded side by side per pass (2, 4 or 8)
        const char *e = getenv("MF_HEAD_CHUNK");
        if (e && (atoi(e) == 1 || atoi(e) == 2 || atoi(e) == 4 || atoi(e) == 8)) { s->chunk = atoi(e); s->chunk_forced = true; }
    }
    s->head_grid = ctx->sm_count;
    MF_CUDA(ctx, cudaDeviceSynchronize());
    return MF_OK;
}

extern "C" int mf_ernerf_reset_state(mf_ctx *ctx) {
    if (!ctx) return MF_E_INVALID;
    if (!ctx->ernerf) return mf_fail(ctx, MF_E_STATE, "ErNeRF weights not loaded");
    MF_CUDA(ctx, cudaMemset(ctx->ernerf->state, 0, 64 * sizeof(float)));
    return MF_OK;
}

extern "C" int mf_ernerf_profile(mf_ctx *ctx, int enable) {
    if (!ctx) return MF_E_INVALID;
    ErnerfState *s = ctx->ernerf;
    if (!s) return mf_fail(ctx, MF_E_STATE, "ErNeRF weights not loaded");
    if (enable && !s->ev_head[0]) {
        MF_CUDA(ctx, cudaEventCreate(&s->ev_head[0]));
        MF_CUDA(ctx, cudaEventCreate(&s->ev_head[1]));
    }
    s->profile = enable != 0;
    return MF_OK;
}

extern "C" int mf_ernerf_last_head_ms(mf_ctx *ctx, float *ms, int64_t *samples) {
    if (!ctx) return MF_E_INVALID;
    ErnerfState *s = ctx->ernerf;
    if (!s || !s->ev_head[0]) return mf_fail(ctx, MF_E_STATE, "profiling not enabled");
    MF_CUDA(ctx, cudaEventSynchronize(s->ev_head[1]));
    if (ms) MF_CUDA(ctx, cudaEventElapsedTime(ms, s->ev_head[0], s->ev_head[1]));
    if (samples) {   // samples shaded by the last k_head launch for this context's frame
        int v = 0;
        MF_CUDA(ctx, cudaMemcpy(&v, s->counters + ((s->frame_no + 1) & 1u) * CT_INTS + CT_SAMPLES, sizeof(int), cudaMemcpyDeviceToHost));
        *samples = v;
    }
    return MF_OK;
}

extern "C" int mf_ernerf_last_launches(const mf_ctx *ctx) { return (ctx && ctx->ernerf) ? ctx->ernerf->last_launches : 0; }

static int ensure_ws(mf_ctx *ctx, ErnerfState *s, int N) {
    if (N <= s->capN) return MF_OK;
    ernerf_free_ws(s);
    MF_CUDA(ctx, cudaMalloc(&s->hits, (size_t)N * 4));
    MF_CUDA(ctx, cudaMalloc(&s->snap, (size_t)N * ER_SNAPS * sizeof(float4)));
    MF_CUDA(ctx, cudaMalloc(&s->rays_t, (size_t)N * 4));
    MF_CUDA(ctx, cudaMalloc(&s->fars, (size_t)N * 4));
    MF_CUDA(ctx, cudaMalloc(&s->weights_sum, (size_t)N * 4));
    MF_CUDA(ctx, cudaMalloc(&s->image, (size_t)N * 12));
    MF_CUDA(ctx, cudaMalloc(&s->final_f32, (size_t)N * 12));
    MF_CUDA(ctx, cudaMalloc(&s->torso_rgb, (size_t)N * 12));
    s->capN = N;
    return MF_OK;
}

// inverse of a general 4x4 (Gauss-Jordan, partial pivoting) in double
static bool inv4(const double *m, double *out) {
    double a[4][8];
    for (int i = 0; i < 4; i++)
        for (int j = 0; j < 4; j++) { a[i][j] = m[i * 4 + j]; a[i][4 + j] = (i == j); }
    for (int c = 0; c < 4; c++) {
        int piv = c;
        for (int r = c + 1; r < 4; r++) if (std::fabs(a[r][c]) > std::fabs(a[piv][c])) piv = r;
        if (std::fabs(a[piv][c]) < 1e-300) return false;
        if (piv != c) for (int j = 0; j < 8; j++) std::swap(a[piv][j], a[c][j]);
        const double d = a[c][c];
        for (int j = 0; j < 8; j++) a[c][j] /= d;
        for (int r = 0; r < 4; r++) if (r != c) {
            const double f = a[r][c];
            for (int j = 0; j < 8; j++) a[r][j] -= f * a[c][j];
        }
    }
    for (int i = 0; i < 4; i++) for (int j = 0; j < 4; j++) out[i * 4 + j] = a[i][4 + j];
    return true;
}

static inline float h16(float v) { return __half2float(__float2half_rn(v)); }

// launch configuration with programmatic stream serialization (MF_ERNERF_PDL=0 turns it off): the kernel may start while its
// predecessor in the stream drains; it calls griddepcontrol.wait before touching anything the predecessor writes
static void pdl_config(cudaLaunchConfig_t &cfg, cudaLaunchAttribute *at, dim3 grid, dim3 block, size_t smem, cudaStream_t stream) {
    static int on = -1;
    if (on < 0) { const char *e = getenv("MF_ERNERF_PDL"); on = e ? atoi(e) != 0 : 1; }
    memset(&cfg, 0, sizeof(cfg));
    cfg.gridDim = grid; cfg.blockDim = block; cfg.dynamicSmemBytes = smem; cfg.stream = stream;
    at[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    at[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = at; cfg.numAttrs = on ? 1 : 0;
}

// one frame of a (possibly batched) render: argument checks, workspace, geometry and the k_setup arguments
struct PreparedFrame {
    ErnerfState *s;
    FrameGeom g;
    SetupParams sp;
    float wa[6];   // wrapped anchor (fp16-rounded), network.py:175-176
    int N, outH, outW;
    bool resize;
};

static int prepare_frame(mf_ctx *ctx, const mf_ernerf_frame *f, const mf_ernerf_debug *dbg, PreparedFrame &pf) {
    ErnerfState *s = ctx->ernerf;
    if (!s) return mf_fail(ctx, MF_E_STATE, "mf_ernerf_render: ErNeRF weights not loaded");
    MF_REQUIRE(ctx, f && f->pose, "mf_ernerf_render: null frame/pose");
    MF_REQUIRE(ctx, f->auds || f->enc_a, "mf_ernerf_render: need auds or enc_a");
    const bool explicit_rays = f->rays_o != nullptr;
    if (explicit_rays)
        MF_REQUIRE(ctx, f->rays_d && f->bg_coords && f->n_rays > 0, "explicit rays need rays_d, bg_coords, n_rays");
    else
        MF_REQUIRE(ctx, f->H > 1 && f->W > 1, "mf_ernerf_render: bad H/W");
    const int N = explicit_rays ? f->n_rays : f->H * f->W;
    pf.s = s; pf.N = N;
    pf.outH = explicit_rays ? 1 : (f->outH > 0 ? f->outH : f->H);
    pf.outW = explicit_rays ? N : (f->outW > 0 ? f->outW : f->W);
    pf.resize = !explicit_rays && (pf.outH != f->H || pf.outW != f->W);
    int rc = ensure_ws(ctx, s, N);
    if (rc) return rc;

    FrameGeom &g = pf.g;
    g.N = N; g.H = explicit_rays ? 1 : f->H; g.W = explicit_rays ? N : f->W;
    for (int i = 0; i < 3; i++) {
        for (int j = 0; j < 3; j++) g.R[i * 3 + j] = f->pose[i * 4 + j];
        g.T[i] = f->pose[i * 4 + 3];
    }
    g.inv_fx = 1.0f / f->fx; g.inv_fy = 1.0f / f->fy; g.cx = f->cx; g.cy = f->cy;
    g.rays_o = f->rays_o; g.rays_d = f->rays_d; g.bg_coords = f->bg_coords;
    g.inv_Hm1 = explicit_rays ? 0.f : 1.0f / (float)(f->H - 1);
    g.inv_Wm1 = explicit_rays ? 0.f : 1.0f / (float)(f->W - 1);

    // wrapped anchor (network.py:175-176): anchor_points @ inverse(pose^T) runs as an fp16 matmul
    // under autocast, the two divisions in fp16
    SetupParams &sp = pf.sp;
    {
        double pt[16], inv[16];
        for (int i = 0; i < 4; i++) for (int j = 0; j < 4; j++) pt[i * 4 + j] = f->pose[j * 4 + i];
        MF_REQUIRE(ctx, inv4(pt, inv), "mf_ernerf_render: singular pose");
        const float *ap = s->misc_host;  // anchor_points, cached at load time
        for (int a = 0; a < 3; a++) {
            float w[4];
            for (int j = 0; j < 4; j++) {
                float acc = 0.f;
                for (int k = 0; k < 4; k++) acc += h16(ap[a * 4 + k]) * h16((float)inv[k * 4 + j]);
                w[j] = h16(acc);
            }
            pf.wa[a * 2 + 0] = h16(h16(w[0] / w[3]) / w[2]);
            pf.wa[a * 2 + 1] = h16(h16(w[1] / w[3]) / w[2]);
        }
    }
    sp.auds = f->auds; sp.enc_a_in = f->enc_a; sp.audio = s->audio;
    sp.state = s->state; sp.counters_next = s->counters + ((s->frame_no + 1) & 1u) * CT_INTS;
    sp.A = (int)s->cfg.audio_in_dim;
    sp.N = N; sp.smooth = (int)s->cfg.smooth_lips; sp.dbg_enc_a = dbg ? dbg->enc_a : nullptr;
    sp.audio_halfs = (int)s->audio_halfs;
    return MF_OK;
}

// n frames of n contexts (n == 1: the plain render).  k_setup: one CTA per frame; k_head: ONE cooperative launch for the batch;
// k_torso_compose (+ resize): per frame.
static int render_frames(mf_ctx *const *ctxs, const mf_ernerf_frame *frames, uint8_t *const *outs, const mf_ernerf_debug *dbg,
                         int n, cudaStream_t stream) {
    mf_ctx *ctx = ctxs[0];
    MF_CUDA(ctx, cudaSetDevice(ctx->device));
    PreparedFrame pf[HEAD_MAX_FRAMES];
    for (int i = 0; i < n; i++) {
        int rc = prepare_frame(ctxs[i], &frames[i], i == 0 ? dbg : nullptr, pf[i]);
        if (rc) {
            if (i > 0) mf_fail(ctx, rc, "mf_ernerf_render_batch: frame %d: %s", i, mf_last_error(ctxs[i]));
            return rc;
        }
    }
    ErnerfState *s0 = pf[0].s;
    const float bnd = s0->cfg.bound;  // renderer.py:86
    const float aabb[6] = {-bnd, -bnd / 2, -bnd, bnd, bnd / 2, bnd};
    int *ctr[HEAD_MAX_FRAMES];   // this frame's counter copy per context (zeroed by the context's previous frame)
    for (int i = 0; i < n; i++) ctr[i] = pf[i].s->counters + (pf[i].s->frame_no & 1u) * CT_INTS;
    SetupBatch sb;
    sb.n = n;
    long total_tiles = 0;
    for (int i = 0; i < n; i++) {
        ErnerfState *s = pf[i].s;
        sb.f[i] = pf[i].sp;
        RayPassFrame &rp = sb.r[i];
        rp.g = pf[i].g; rp.hits = s->hits; rp.counters = ctr[i]; rp.rays_t = s->rays_t; rp.fars = s->fars;
        rp.nears = (i == 0 && dbg && dbg->nears) ? dbg->nears : nullptr;
        rp.weights_sum = s->weights_sum; rp.image = s->image;
        total_tiles += (pf[i].N + 31) / 32;
    }
    sb.bound = s0->cfg.bound; sb.min_near = s0->cfg.min_near; sb.dt_gamma = s0->cfg.dt_gamma;
    sb.max_steps = s0->cfg.max_steps; sb.cascade = s0->cfg.cascade; sb.grid_size = s0->cfg.grid_size;
    sb.bitfield = s0->bitfield; sb.bitfield_linear = s0->bitfield_linear;
    sb.mp = make_march_params(sb.bound, sb.dt_gamma, sb.max_steps, sb.cascade, sb.grid_size, sb.bitfield);
    if (sb.bitfield_linear) { sb.mp.grid = sb.bitfield_linear; sb.mp.linear = true; }
    for (int i = 0; i < 6; i++) sb.aabb[i] = aabb[i];
    {   // 8 audio CTAs per frame first, then the ray pass: one CTA (32 warps, one ray per lane) per SM that is left
        const int n_audio = 8 * n;
        const int ray_ctas = (int)std::max<long>(1, std::min<long>((n > 1 ? s0->setup_per_sm : 1) * ctx->sm_count - n_audio, (total_tiles + SETUP_THREADS / 32 - 1) / (SETUP_THREADS / 32)));
        k_setup<<<n_audio + ray_ctas, SETUP_THREADS, 0, stream>>>(sb);
    }
    int launches = 1;

    HeadParams hp;
    hp.hl = s0->hl; hp.planes = s0->planes; hp.plane_rows = s0->plane_rows; hp.bitfield = s0->bitfield;
    hp.bitfield_linear = s0->bitfield_linear;
    hp.mlp_image = s0->head_mlp;
    hp.bound = s0->cfg.bound; hp.min_near = s0->cfg.min_near; hp.dt_gamma = s0->cfg.dt_gamma; hp.T_thresh = s0->cfg.T_thresh;
    hp.max_steps = s0->cfg.max_steps; hp.cascade = s0->cfg.cascade; hp.grid_size = s0->cfg.grid_size;
    hp.n_frames = n;
    hp.mp = sb.mp;
    for (int i = 0; i < n; i++) {
        ErnerfState *s = pf[i].s;
        HeadFrame &hf = hp.f[i];
        hf.g = pf[i].g; hf.state = frames[i].enc_a ? frames[i].enc_a : s->state; hf.eye = frames[i].eye;
        hf.hits = s->hits; hf.counters = ctr[i];
        hf.rays_t = s->rays_t; hf.fars = s->fars; hf.weights_sum = s->weights_sum; hf.image = s->image; hf.snap = s->snap;
    }
    {
        // Samples of a ray shaded side by side per pass: 2 for full frames (least speculation, measured best); a launch with few
        // rays cannot fill the ray slots anyway and is bound by the number of sequential passes per ray, so it takes 4 or 8
        // (the literal "2048 rays per frame" reading of BASELINE configs[3]).  MF_HEAD_CHUNK overrides.
        int chunk = s0->chunk;
        if (!s0->chunk_forced) chunk = total_tiles * 32 <= 16384 ? 8 : (total_tiles * 32 <= 65536 ? 4 : 2);
        // one CTA per SM; small ray counts get fewer CTAs (every CTA stages the 57 KB MLP image): 16 warps x 32 / chunk ray slots each
        const long slots_per_cta = HEAD_WARPS * (32 / chunk);
        const int grid = (int)std::min<long>(s0->head_grid, std::max<long>(1, (total_tiles * 32 + slots_per_cta - 1) / slots_per_cta));
        if (s0->profile) MF_CUDA(ctx, cudaEventRecord(s0->ev_head[0], stream));
        cudaLaunchConfig_t cfg;
        cudaLaunchAttribute at[1];
        pdl_config(cfg, at, dim3(grid), dim3(HEAD_THREADS), sizeof(HeadSmem), stream);
        if (chunk == 1) MF_CUDA(ctx, cudaLaunchKernelEx(&cfg, k_head<1>, hp));
        else if (chunk == 4) MF_CUDA(ctx, cudaLaunchKernelEx(&cfg, k_head<4>, hp));
        else if (chunk == 8) MF_CUDA(ctx, cudaLaunchKernelEx(&cfg, k_head<8>, hp));
        else MF_CUDA(ctx, cudaLaunchKernelEx(&cfg, k_head<2>, hp));
        if (s0->profile) MF_CUDA(ctx, cudaEventRecord(s0->ev_head[1], stream));
        launches++;
    }

    for (int i = 0; i < n; i++) {
        ErnerfState *s = pf[i].s;
        const mf_ernerf_frame *f = &frames[i];
        const mf_ernerf_debug *d = i == 0 ? dbg : nullptr;
        const int N = pf[i].N;
        ComposeParams cp;
        cp.g = pf[i].g;
        cp.tm.tl = s->tl; cp.tm.table = s->torso_table; cp.tm.density = s->torso_density; cp.tm.mlp_image = s->torso_mlp;
        cp.tm.torso_const = s->torso_const; cp.tm.misc = s->misc; cp.tm.thresh = s->cfg.density_thresh_torso;
        cp.tm.shrink = s->cfg.torso_shrink; cp.tm.G = (int)s->cfg.grid_size;
        cp.tf.bg_color = (const __half *)f->bg_color; cp.tf.dbg_mask = d ? d->torso_mask : nullptr;
        for (int k = 0; k < 6; k++) cp.tf.wa[k] = pf[i].wa[k];
        cp.weights_sum = s->weights_sum; cp.image = s->image; cp.snap = s->snap; cp.counters = ctr[i]; cp.torso_rgb = s->torso_rgb;
        cp.max_steps = (int)s->cfg.max_steps;
        cp.out_f32 = pf[i].resize ? s->final_f32 : f->out_image_f32;
        cp.out_u8 = pf[i].resize ? nullptr : outs[i];
        cp.dbg_image_head = d ? d->image_head : nullptr;
        {
            const int n_tiles = (N + 31) / 32;
            const int grid = std::max(1, std::min((n_tiles + TORSO_WARPS - 1) / TORSO_WARPS, ctx->sm_count * TORSO_MINB));   // one wave
            cudaLaunchConfig_t cfg;
            cudaLaunchAttribute at[1];
            pdl_config(cfg, at, dim3(grid), dim3(TORSO_THREADS), sizeof(TorsoSmem), stream);
            // n > 1 (batched sessions): the second frame's kernel must not pass the first one's -- only the first is a programmatic dependent
            if (i > 0) cfg.numAttrs = 0;
            MF_CUDA(ctx, cudaLaunchKernelEx(&cfg, k_torso_compose, cp));
            launches++;
        }
        if (pf[i].resize) {
            const int tot = pf[i].outH * pf[i].outW;
            k_resize_u8<<<(tot + 255) / 256, 256, 0, stream>>>(s->final_f32, f->H, f->W, pf[i].outH, pf[i].outW, f->out_image_f32, outs[i]);
            launches++;
        }
        if (d) {
            if (d->fars) MF_CUDA(ctx, cudaMemcpyAsync(d->fars, s->fars, (size_t)N * 4, cudaMemcpyDeviceToDevice, stream));
            if (d->weights_sum)
                MF_CUDA(ctx, cudaMemcpyAsync(d->weights_sum, s->weights_sum, (size_t)N * 4, cudaMemcpyDeviceToDevice, stream));
            if (d->round_info)
                MF_CUDA(ctx, cudaMemcpyAsync(d->round_info, ctr[i] + CT_ROUNDS, (ER_MAX_ROUNDS + 1) * 4 * sizeof(int),
                                             cudaMemcpyDeviceToDevice, stream));
        }
    }
    MF_CUDA(ctx, cudaGetLastError());
    for (int i = 0; i < n; i++) {
        pf[i].s->last_launches = launches;
        pf[i].s->frame_no++;
    }
    return MF_OK;
}

// encode_audio + EMA of one attention window, exactly the audio CTAs of a render (same kernel, same state update), without a frame
extern "C" int mf_ernerf_encode_audio(mf_ctx *ctx, const float *auds, float *enc_a_out, void *stream_) {
    if (!ctx) return MF_E_INVALID;
    ErnerfState *s = ctx->ernerf;
    if (!s) return mf_fail(ctx, MF_E_STATE, "mf_ernerf_encode_audio: ErNeRF weights not loaded");
    MF_REQUIRE(ctx, auds && enc_a_out, "mf_ernerf_encode_audio: null pointer");
    MF_CUDA(ctx, cudaSetDevice(ctx->device));
    SetupBatch sb;
    memset(&sb, 0, sizeof(sb));
    sb.n = 1;
    SetupParams &sp = sb.f[0];
    sp.auds = auds; sp.enc_a_in = nullptr; sp.audio = s->audio; sp.state = s->state; sp.counters_next = nullptr;
    sp.A = (int)s->cfg.audio_in_dim; sp.N = 0; sp.smooth = (int)s->cfg.smooth_lips; sp.audio_halfs = (int)s->audio_halfs;
    sp.dbg_enc_a = enc_a_out;
    k_setup<<<8, SETUP_THREADS, 0, (cudaStream_t)stream_>>>(sb);
    MF_CUDA(ctx, cudaGetLastError());
    return MF_OK;
}

extern "C" int mf_ernerf_render(mf_ctx *ctx, const mf_ernerf_frame *f, uint8_t *out_rgb, const mf_ernerf_debug *dbg,
                                void *stream_) {
    if (!ctx) return MF_E_INVALID;
    if (!ctx->ernerf) return mf_fail(ctx, MF_E_STATE, "mf_ernerf_render: ErNeRF weights not loaded");
    MF_REQUIRE(ctx, f, "mf_ernerf_render: null frame/pose");
    return render_frames(&ctx, f, &out_rgb, dbg, 1, (cudaStream_t)stream_);
}

extern "C" int mf_ernerf_render_batch(mf_ctx *const *ctxs, const mf_ernerf_frame *frames, uint8_t *const *outs, int n, void *stream_) {
    if (!ctxs || n < 1 || !ctxs[0]) return MF_E_INVALID;
    mf_ctx *ctx = ctxs[0];
    MF_REQUIRE(ctx, frames && outs, "mf_ernerf_render_batch: null pointer");
    MF_REQUIRE(ctx, n <= HEAD_MAX_FRAMES, "mf_ernerf_render_batch: %d frames, at most %d per call", n, HEAD_MAX_FRAMES);
    for (int i = 0; i < n; i++) {
        MF_REQUIRE(ctx, ctxs[i] && ctxs[i]->ernerf, "mf_ernerf_render_batch: context %d has no ErNeRF weights", i);
        MF_REQUIRE(ctx, ctxs[i]->device == ctx->device, "mf_ernerf_render_batch: contexts on different devices");
        for (int j = 0; j < i; j++) MF_REQUIRE(ctx, ctxs[j] != ctxs[i], "mf_ernerf_render_batch: a context appears twice (one frame per session)");
        const ErnerfState *a = ctx->ernerf, *b = ctxs[i]->ernerf;
        // one MLP image / one set of tables is staged per CTA: the sessions of a batch must render the same avatar model
        MF_REQUIRE(ctx, a->head_mlp == b->head_mlp && a->planes == b->planes && a->bitfield == b->bitfield &&
                            memcmp(&a->cfg, &b->cfg, sizeof(a->cfg)) == 0,
                   "mf_ernerf_render_batch: context %d was loaded from a different blob / configuration", i);
    }
    return render_frames(ctxs, frames, outs, nullptr, n, (cudaStream_t)stream_);
}

// ---- kernel-level entry points ---------------------------------------------------------
extern "C" int mf_near_far_from_aabb(mf_ctx *ctx, const float *rays_o, const float *rays_d, const float *aabb,
                                     uint32_t N, float min_near, float *nears, float *fars, void *stream) {
    if (!ctx) return MF_E_INVALID;
    MF_REQUIRE(ctx, rays_o && rays_d && aabb && nears && fars, "mf_near_far_from_aabb: null pointer");
    if (N == 0) return MF_OK;
    k_near_far<<<(N + 127) / 128, 128, 0, (cudaStream_t)stream>>>(rays_o, rays_d, aabb, N, min_near, nears, fars);
    MF_CUDA(ctx, cudaGetLastError());
    return MF_OK;
}

extern "C" int mf_march_rays(mf_ctx *ctx, uint32_t n_alive, uint32_t n_step, const int32_t *rays_alive,
                             const float *rays_t, const float *rays_o, const float *rays_d, float bound,
                             float dt_gamma, uint32_t max_steps, uint32_t C, uint32_t H, const uint8_t *grid,
                             const float *nears, const float *fars, float *xyzs, float *dirs, float *deltas,
                             const float *noises, void *stream) {
    if (!ctx) return MF_E_INVALID;
    MF_REQUIRE(ctx, rays_alive && rays_t && rays_o && rays_d && grid && nears && fars && xyzs && dirs && deltas,
               "mf_march_rays: null pointer");
    MF_REQUIRE(ctx, C >= 1 && H >= 1 && max_steps >= 1, "mf_march_rays: bad C/H/max_steps");
    if (n_alive == 0 || n_step == 0) return MF_OK;
    k_march_rays<<<(n_alive + 127) / 128, 128, 0, (cudaStream_t)stream>>>(n_alive, n_step, rays_alive, rays_t, rays_o,
                                                                            rays_d, bound, dt_gamma, max_steps, C, H,
                                                                            grid, nears, fars, xyzs, dirs, deltas, noises);
    MF_CUDA(ctx, cudaGetLastError());
    return MF_OK;
}

extern "C" int mf_composite_rays_triplane(mf_ctx *ctx, uint32_t n_alive, uint32_t n_step, float T_thresh,
                                          int32_t *rays_alive, float *rays_t, const float *sigmas, const float *rgbs,
                                          const float *deltas, const float *ambs_aud, const float *ambs_eye,
                                          const float *uncertainties, float *weights_sum, float *depth, float *image,
                                          float *amb_aud_sum, float *amb_eye_sum, float *uncertainty_sum,
                                          void *stream) {
    if (!ctx) return MF_E_INVALID;
    MF_REQUIRE(ctx, rays_alive && rays_t && sigmas && rgbs && deltas && weights_sum && image,
               "mf_composite_rays_triplane: null pointer");
    if (n_alive == 0) return MF_OK;
    k_composite_rays_triplane<<<(n_alive + 127) / 128, 128, 0, (cudaStream_t)stream>>>(
        n_alive, n_step, T_thresh, rays_alive, rays_t, sigmas, rgbs, deltas, ambs_aud, ambs_eye, uncertainties,
        weights_sum, depth, image, amb_aud_sum, amb_eye_sum, uncertainty_sum);
    MF_CUDA(ctx, cudaGetLastError());
    return MF_OK;
}

extern "C" int mf_grid_encode_forward(mf_ctx *ctx, const float *inputs, const void *embeddings,
                                      const int32_t *offsets, void *outputs, uint32_t B, uint32_t D, uint32_t C,
                                      uint32_t L, float S, uint32_t H, uint32_t gridtype, int align_corners,
                                      int embeddings_is_half, void *stream_) {
    if (!ctx) return MF_E_INVALID;
    MF_REQUIRE(ctx, inputs && embeddings && offsets && outputs, "mf_grid_encode_forward: null pointer");
    if (D != 2 || align_corners || L > 16 || L < 1 || !((C == 1 && !embeddings_is_half) || (C == 2 && embeddings_is_half)))
        return mf_fail(ctx, MF_E_UNSUPPORTED,
                       "mf_grid_encode_forward: only D=2, align_corners=False, (C=1,fp32) or (C=2,fp16) are on the hot path");
    if (B == 0) return MF_OK;
    cudaStream_t stream = (cudaStream_t)stream_;
    int32_t off[17];
    MF_CUDA(ctx, cudaMemcpyAsync(off, offsets, (L + 1) * sizeof(int32_t), cudaMemcpyDeviceToHost, stream));
    MF_CUDA(ctx, cudaStreamSynchronize(stream));
    float sc[16];
    int rc = compute_scales(ctx, S, H, (int)L, sc);
    if (rc) return rc;
    GridLevelsAny lv;
    fill_levels(lv.lv, (int)L, sc, off, gridtype);
    k_grid_encode<<<dim3((B + 255) / 256, L), 256, 0, stream>>>(inputs, embeddings, lv, outputs, B, C, embeddings_is_half);
    MF_CUDA(ctx, cudaGetLastError());
    return MF_OK;
}

extern "C" int mf_grid_level_scales(mf_ctx *ctx, float S, uint32_t H, uint32_t L, float *scales_host) {
    if (!ctx) return MF_E_INVALID;
    MF_REQUIRE(ctx, scales_host && L >= 1 && L <= 16, "mf_grid_level_scales: bad arguments");
    return compute_scales(ctx, S, H, (int)L, scales_host);
}

extern "C" int mf_sh_encode_forward(mf_ctx *ctx, const float *inputs, float *outputs, uint32_t B, uint32_t D,
                                    uint32_t C, void *stream) {
    if (!ctx) return MF_E_INVALID;
    MF_REQUIRE(ctx, inputs && outputs, "mf_sh_encode_forward: null pointer");
    if (D != 3 || C != 4) return mf_fail(ctx, MF_E_UNSUPPORTED, "mf_sh_encode_forward: only D=3, degree 4 is on the hot path");
    if (B == 0) return MF_OK;
    k_sh4<<<(B + 255) / 256, 256, 0, (cudaStream_t)stream>>>(inputs, outputs, B);
    MF_CUDA(ctx, cudaGetLastError());
    return MF_OK;
}

extern "C" int mf_freq_encode_forward(mf_ctx *ctx, const float *inputs, uint32_t B, uint32_t D, uint32_t deg,
                                      uint32_t C, float *outputs, void *stream) {
    if (!ctx) return MF_E_INVALID;
    MF_REQUIRE(ctx, inputs && outputs, "mf_freq_encode_forward: null pointer");
    MF_REQUIRE(ctx, C == D + D * deg * 2, "mf_freq_encode_forward: C != D + 2*D*deg");
    if (B == 0) return MF_OK;
    const uint32_t tot = B * C;
    k_freq<<<(tot + 127) / 128, 128, 0, (cudaStream_t)stream>>>(inputs, B, D, C, outputs);
    MF_CUDA(ctx, cudaGetLastError());
    return MF_OK;
}
