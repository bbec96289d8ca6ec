/*
 * oracle/ernerf_kernels.c -- CPU restatement of the reference's ErNeRF inference kernels.
 *
 * TEST INFRASTRUCTURE ONLY: this file is the checker for the sm_100a kernels in
 * mere_fusion_b200/csrc.  It is never linked into, imported by, or called from the product
 * path; only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference
 * legs may use it.
 *
 * Every function follows one reference CUDA kernel expression by expression (paths relative
 * to /root/reference/ernerf):
 *   orc_near_far_from_aabb        raymarching/src/raymarching.cu:91-145
 *   orc_march_rays                raymarching/src/raymarching.cu:827-929 (+ helpers :19-71)
 *   orc_composite_rays_triplane   raymarching/src/raymarching.cu:2141-2249
 *   orc_grid_encode_f32 / _f16    gridencoder/src/gridencoder.cu:35-72,75-175
 *   orc_sh_encode4                shencoder/src/shencoder.cu:27-68 (degree 4)
 *   orc_freq_encode               freqencoder/src/freqencoder.cu:30-58
 *
 * Where nvcc's default -fmad=true would contract a*b+c the restatement uses fmaf explicitly
 * (SURVEY.md note N3); where the CUDA source mixes double literals into float expressions
 * (note N2) the same promotions are kept.  __expf/__sinf are restated with expf/sinf: those
 * two are approximate on the GPU, so float outputs that pass through them are compared with
 * a tolerance, integer outputs (alive sets, sample counts, voxel indices) exactly.
 *
 * Parity pin: tests/test_ernerf_ref_gpu.py runs the reference's own compiled kernels
 * (oracle/_ref, built by oracle/build_ref.py) on the GPU box against these functions on the
 * shipped checkpoint's data; golden vectors produced there are committed under tests/golden/.
 */
#include <math.h>
#include <stdint.h>
#include <string.h>
#include <float.h>

typedef _Float16 half_t;

static inline float signf_(const float x) { return copysignf(1.0f, x); }
static inline float clampf_(const float x, const float lo, const float hi) { return fminf(hi, fmaxf(lo, x)); }

/* raymarching.cu:42-47 */
static inline int mip_from_pos(const float x, const float y, const float z, const float max_cascade) {
    const float mx = fmaxf(fabsf(x), fmaxf(fabsf(y), fabsf(z)));
    int exponent;
    frexpf(mx, &exponent);
    return (int)fminf(max_cascade - 1, fmaxf(0, (float)exponent));
}

/* raymarching.cu:49-54: dt * H * 0.5 -- the 0.5 literal is double, narrowed into a float */
static inline int mip_from_dt(const float dt, const float H, const float max_cascade) {
    const float mx = (float)((double)(dt * H) * 0.5);
    int exponent;
    frexpf(mx, &exponent);
    return (int)fminf(max_cascade - 1, fmaxf(0, (float)exponent));
}

/* raymarching.cu:56-71 */
static inline uint32_t expand_bits(uint32_t v) {
    v = (v * 0x00010001u) & 0xFF0000FFu;
    v = (v * 0x00000101u) & 0x0F00F00Fu;
    v = (v * 0x00000011u) & 0xC30C30C3u;
    v = (v * 0x00000005u) & 0x49249249u;
    return v;
}
static inline uint32_t morton3D(uint32_t x, uint32_t y, uint32_t z) {
    return expand_bits(x) | (expand_bits(y) << 1) | (expand_bits(z) << 2);
}

/* raymarching.cu:91-145 */
void orc_near_far_from_aabb(const float *rays_o, const float *rays_d, const float *aabb,
                            uint32_t N, float min_near, float *nears, float *fars) {
#pragma omp parallel for schedule(static)
    for (uint32_t n = 0; n < N; n++) {
        const float ox = rays_o[n * 3 + 0], oy = rays_o[n * 3 + 1], oz = rays_o[n * 3 + 2];
        const float dx = rays_d[n * 3 + 0], dy = rays_d[n * 3 + 1], dz = rays_d[n * 3 + 2];
        const float rdx = 1 / dx, rdy = 1 / dy, rdz = 1 / dz;
        float near = (aabb[0] - ox) * rdx, far = (aabb[3] - ox) * rdx, tmp;
        if (near > far) { tmp = near; near = far; far = tmp; }
        float near_y = (aabb[1] - oy) * rdy, far_y = (aabb[4] - oy) * rdy;
        if (near_y > far_y) { tmp = near_y; near_y = far_y; far_y = tmp; }
        if (near > far_y || near_y > far) { nears[n] = fars[n] = FLT_MAX; continue; }
        if (near_y > near) near = near_y;
        if (far_y < far) far = far_y;
        float near_z = (aabb[2] - oz) * rdz, far_z = (aabb[5] - oz) * rdz;
        if (near_z > far_z) { tmp = near_z; near_z = far_z; far_z = tmp; }
        if (near > far_z || near_z > far) { nears[n] = fars[n] = FLT_MAX; continue; }
        if (near_z > near) near = near_z;
        if (far_z < far) far = far_z;
        if (near < min_near) near = min_near;
        nears[n] = near;
        fars[n] = far;
    }
}

/* raymarching.cu:827-929.  xyzs/dirs/deltas must be zero-filled by the caller, as the
 * reference wrapper does (raymarching/raymarching.py:383-385).  vox (nullable) receives the
 * bitfield index tested for every emitted sample: it is not a reference output, it makes the
 * "voxel indices bit-exact" check of SURVEY.md N4 observable. */
void orc_march_rays(uint32_t n_alive, uint32_t n_step, const int *rays_alive, const float *rays_t,
                    const float *rays_o, const float *rays_d, float bound, float dt_gamma,
                    uint32_t max_steps, uint32_t C, uint32_t H, const uint8_t *grid,
                    const float *nears, const float *fars, float *xyzs, float *dirs, float *deltas,
                    const float *noises, int32_t *vox) {
#pragma omp parallel for schedule(dynamic, 256)
    for (uint32_t n = 0; n < n_alive; n++) {
        const int index = rays_alive[n];
        const float noise = noises ? noises[n] : 0.0f;
        const float *ro = rays_o + (size_t)index * 3, *rd = rays_d + (size_t)index * 3;
        float *px = xyzs + (size_t)n * n_step * 3, *pd = dirs + (size_t)n * n_step * 3;
        float *pdl = deltas + (size_t)n * n_step * 2;
        int32_t *pv = vox ? vox + (size_t)n * n_step : 0;
        const float ox = ro[0], oy = ro[1], oz = ro[2];
        const float dx = rd[0], dy = rd[1], dz = rd[2];
        const float rdx = 1 / dx, rdy = 1 / dy, rdz = 1 / dz;
        const float rH = 1 / (float)H;
        const float H3 = (float)(H * H * H);
        float t = rays_t[index];
        const float far = fars[index];
        /* 2 * SQRT3() * (1 << (C - 1)) / H : float * int -> float, / uint32 -> float */
        const float dt_max = 2 * 1.7320508075688772f * (float)(1 << (C - 1)) / (float)H;
        const float dt_min = fminf(dt_max, 2 * 1.7320508075688772f / (float)max_steps);
        uint32_t step = 0;
        t += clampf_(t * dt_gamma, dt_min, dt_max) * noise;
        while (t < far && step < n_step) {
            const float x = clampf_(fmaf(t, dx, ox), -bound, bound);
            const float y = clampf_(fmaf(t, dy, oy), -bound, bound);
            const float z = clampf_(fmaf(t, dz, oz), -bound, bound);
            const float dt = clampf_(t * dt_gamma, dt_min, dt_max);
            const int a = mip_from_pos(x, y, z, (float)C), b = mip_from_dt(dt, (float)H, (float)C);
            const int level = a > b ? a : b;
            const float mip_bound = fminf(scalbnf(1, level), bound);
            const float mip_rbound = 1 / mip_bound;
            /* 0.5 * (x * mip_rbound + 1) * H evaluated in double, narrowed by clamp(const float ...) */
            const int nx = (int)clampf_((float)(0.5 * (double)fmaf(x, mip_rbound, 1.0f) * (double)H), 0.0f, (float)(H - 1));
            const int ny = (int)clampf_((float)(0.5 * (double)fmaf(y, mip_rbound, 1.0f) * (double)H), 0.0f, (float)(H - 1));
            const int nz = (int)clampf_((float)(0.5 * (double)fmaf(z, mip_rbound, 1.0f) * (double)H), 0.0f, (float)(H - 1));
            /* level * H3 + morton: uint32 -> float arithmetic, back to uint32 (exact below 2^24 .. level 0) */
            const uint32_t gi = (uint32_t)((float)level * H3 + (float)morton3D(nx, ny, nz));
            const int occ = grid[gi / 8] & (1 << (gi % 8));
            if (occ) {
                px[0] = x; px[1] = y; px[2] = z;
                pd[0] = dx; pd[1] = dy; pd[2] = dz;
                t += dt;
                pdl[0] = dt; pdl[1] = t;
                if (pv) { *pv++ = (int32_t)gi; }
                px += 3; pd += 3; pdl += 2;
                step++;
            } else {
                const float tx = (((nx + 0.5f + 0.5f * signf_(dx)) * rH * 2 - 1) * mip_bound - x) * rdx;
                const float ty = (((ny + 0.5f + 0.5f * signf_(dy)) * rH * 2 - 1) * mip_bound - y) * rdy;
                const float tz = (((nz + 0.5f + 0.5f * signf_(dz)) * rH * 2 - 1) * mip_bound - z) * rdz;
                const float tt = t + fmaxf(0.0f, fminf(tx, fminf(ty, tz)));
                do {
                    t += clampf_(t * dt_gamma, dt_min, dt_max);
                } while (t < tt);
            }
        }
    }
}

/* raymarching.cu:2141-2249 */
void orc_composite_rays_triplane(uint32_t n_alive, uint32_t n_step, float T_thresh, int *rays_alive,
                                 float *rays_t, const float *sigmas, const float *rgbs,
                                 const float *deltas, const float *ambs_aud, const float *ambs_eye,
                                 const float *uncertainties, float *weights_sum, float *depth,
                                 float *image, float *amb_aud_sum, float *amb_eye_sum,
                                 float *uncertainty_sum) {
#pragma omp parallel for schedule(static)
    for (uint32_t n = 0; n < n_alive; n++) {
        const int index = rays_alive[n];
        const float *sg = sigmas + (size_t)n * n_step, *rg = rgbs + (size_t)n * n_step * 3;
        const float *dl = deltas + (size_t)n * n_step * 2;
        const float *aa = ambs_aud ? ambs_aud + (size_t)n * n_step : 0;
        const float *ae = ambs_eye ? ambs_eye + (size_t)n * n_step : 0;
        const float *un = uncertainties ? uncertainties + (size_t)n * n_step : 0;
        float t = rays_t[index];
        float weight_sum = weights_sum[index], d = depth[index];
        float r = image[index * 3], g = image[index * 3 + 1], b = image[index * 3 + 2];
        float a_aud = amb_aud_sum ? amb_aud_sum[index] : 0, a_eye = amb_eye_sum ? amb_eye_sum[index] : 0;
        float u = uncertainty_sum ? uncertainty_sum[index] : 0;
        uint32_t step = 0;
        while (step < n_step) {
            if (dl[0] == 0) break;
            const float alpha = 1.0f - expf(-sg[0] * dl[0]);
            const float T = 1 - weight_sum;
            const float weight = alpha * T;
            weight_sum += weight;
            t = dl[1];
            d = fmaf(weight, t, d);
            r = fmaf(weight, rg[0], r);
            g = fmaf(weight, rg[1], g);
            b = fmaf(weight, rg[2], b);
            if (aa) a_aud += aa[0];
            if (ae) a_eye += ae[0];
            if (un) u = fmaf(weight, un[0], u);
            if (T < T_thresh) break;
            sg++; rg += 3; dl += 2; step++;
            if (aa) aa++;
            if (ae) ae++;
            if (un) un++;
        }
        if (step < n_step) rays_alive[n] = -1;
        else rays_t[index] = t;
        weights_sum[index] = weight_sum;
        depth[index] = d;
        image[index * 3] = r; image[index * 3 + 1] = g; image[index * 3 + 2] = b;
        if (amb_aud_sum) amb_aud_sum[index] = a_aud;
        if (amb_eye_sum) amb_eye_sum[index] = a_eye;
        if (uncertainty_sum) uncertainty_sum[index] = u;
    }
}

/* gridencoder.cu:35-51 (D == 2 -> primes 1, 2654435761) and :54-72 */
static inline uint32_t grid_index2(uint32_t gridtype, int align_corners, uint32_t hashmap_size,
                                   uint32_t resolution, const uint32_t pg[2]) {
    uint32_t stride = 1, index = 0;
    for (uint32_t d = 0; d < 2 && stride <= hashmap_size; d++) {
        index += pg[d] * stride;
        stride *= align_corners ? resolution : (resolution + 1);
    }
    if (gridtype == 0 && stride > hashmap_size) index = (pg[0] * 1u) ^ (pg[1] * 2654435761u);
    return index % hashmap_size;
}

/* gridencoder.cu:75-175, D = 2.  outputs are [L, B, C] like the kernel's; the permute to
 * [B, L*C] (gridencoder/grid.py:52) is done by the caller.  S = log2(per_level_scale) as float.
 * `scales` (nullable) overrides the per-level scale: CUDA's exp2f is a 2-ulp approximation
 * (MUFU.EX2), so glibc's exp2f differs from it in the last bit on some levels; feeding the
 * device-evaluated scales (tests/golden/ernerf_level_scales.json, or mf_grid_level_scales on the
 * GPU box) makes this restatement bit-exact against the reference kernel. */
#define GRID_BODY(SCALAR, ACCUM)                                                                   \
    for (uint32_t level = 0; level < L; level++) {                                                 \
        const SCALAR *g = grid + (size_t)(uint32_t)offsets[level] * C;                             \
        const uint32_t hashmap_size = offsets[level + 1] - offsets[level];                         \
        const float scale = scales ? scales[level] : exp2f(level * S) * H - 1.0f;                  \
        const uint32_t resolution = (uint32_t)ceilf(scale) + 1;                                    \
        _Pragma("omp parallel for schedule(static)")                                               \
        for (uint32_t b = 0; b < B; b++) {                                                         \
            const float *in = inputs + (size_t)b * 2;                                              \
            SCALAR *out = outputs + ((size_t)level * B + b) * C;                                   \
            if (in[0] < 0 || in[0] > 1 || in[1] < 0 || in[1] > 1) {                                \
                for (uint32_t ch = 0; ch < C; ch++) out[ch] = 0;                                   \
                continue;                                                                          \
            }                                                                                      \
            float pos[2]; uint32_t pg[2];                                                          \
            for (int d = 0; d < 2; d++) {                                                          \
                pos[d] = fmaf(in[d], scale, align_corners ? 0.0f : 0.5f);                          \
                pg[d] = (uint32_t)floorf(pos[d]);                                                  \
                pos[d] -= (float)pg[d];                                                            \
            }                                                                                      \
            SCALAR res[8] = {0};                                                                   \
            for (uint32_t idx = 0; idx < 4; idx++) {                                               \
                float w = 1; uint32_t pl[2];                                                       \
                for (int d = 0; d < 2; d++) {                                                      \
                    if ((idx & (1u << d)) == 0) { w *= 1 - pos[d]; pl[d] = pg[d]; }                \
                    else { w *= pos[d]; pl[d] = pg[d] + 1; }                                       \
                }                                                                                  \
                const uint32_t gi = grid_index2(gridtype, align_corners, hashmap_size, resolution, pl) * C; \
                for (uint32_t ch = 0; ch < C; ch++) { ACCUM; }                                     \
            }                                                                                      \
            for (uint32_t ch = 0; ch < C; ch++) out[ch] = res[ch];                                 \
        }                                                                                          \
    }

/* fp32 tables (head planes, C = 1: never halved, gridencoder/grid.py:37-39): results += w * grid,
 * contracted to an fma by nvcc */
void orc_grid_encode_f32(const float *inputs, const float *grid, const int *offsets, float *outputs,
                         uint32_t B, uint32_t C, uint32_t L, float S, uint32_t H, uint32_t gridtype,
                         int align_corners, const float *scales) {
    GRID_BODY(float, res[ch] = fmaf(w, g[gi + ch], res[ch]))
}

/* fp16 tables (torso, C = 2 under autocast): scalar_t = at::Half, so `results[ch] += w * grid[..]`
 * is Half += float, i.e. the float product is rounded to half, added in float, rounded to half
 * (c10/util/Half-inl.h operator+= / operator+) */
void orc_grid_encode_f16(const float *inputs, const half_t *grid, const int *offsets, half_t *outputs,
                         uint32_t B, uint32_t C, uint32_t L, float S, uint32_t H, uint32_t gridtype,
                         int align_corners, const float *scales) {
    GRID_BODY(half_t, res[ch] = (half_t)((float)res[ch] + (float)(half_t)(w * (float)g[gi + ch])))
}

/* shencoder.cu:27-68, degree 4 (C = 4 -> 16 outputs), fp32 (sphere_harmonics.py:16) */
void orc_sh_encode4(const float *inputs, float *outputs, uint32_t B) {
#pragma omp parallel for schedule(static)
    for (uint32_t b = 0; b < B; b++) {
        const float x = inputs[b * 3], y = inputs[b * 3 + 1], z = inputs[b * 3 + 2];
        float *o = outputs + (size_t)b * 16;
        const float xy = x * y, xz = x * z, yz = y * z, x2 = x * x, y2 = y * y, z2 = z * z;
        o[0] = 0.28209479177387814f;
        o[1] = -0.48860251190291987f * y;
        o[2] = 0.48860251190291987f * z;
        o[3] = -0.48860251190291987f * x;
        o[4] = 1.0925484305920792f * xy;
        o[5] = -1.0925484305920792f * yz;
        o[6] = fmaf(0.94617469575755997f, z2, -0.31539156525251999f);
        o[7] = -1.0925484305920792f * xz;
        o[8] = fmaf(0.54627421529603959f, x2, -(0.54627421529603959f * y2));
        o[9] = 0.59004358992664352f * y * fmaf(-3.0f, x2, y2);
        o[10] = 2.8906114426405538f * xy * z;
        o[11] = 0.45704579946446572f * y * fmaf(-5.0f, z2, 1.0f);
        o[12] = 0.3731763325901154f * z * fmaf(5.0f, z2, -3.0f);
        o[13] = 0.45704579946446572f * x * fmaf(-5.0f, z2, 1.0f);
        o[14] = 1.4453057213202769f * z * (x2 - y2);
        o[15] = 0.59004358992664352f * x * fmaf(3.0f, y2, -x2);
    }
}

/* freqencoder.cu:30-58; __sinf restated with sinf (freqencoder is the one -use_fast_math build) */
void orc_freq_encode(const float *inputs, uint32_t B, uint32_t D, uint32_t deg, uint32_t C, float *outputs) {
    (void)deg;
#pragma omp parallel for schedule(static)
    for (uint32_t t = 0; t < B * C; t++) {
        const uint32_t b = t / C, c = t - b * C;
        const float *in = inputs + (size_t)b * D;
        if (c < D) { outputs[t] = in[c]; continue; }
        const uint32_t col = c / D - 1, d = c % D, freq = col / 2;
        const float phase_shift = (col % 2) * (3.141592653589793f / 2);
        outputs[t] = sinf(scalbnf(in[d], (int)freq) + phase_shift);
    }
}

int orc_abi_version(void) { return 1; }
