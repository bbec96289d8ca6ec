// ernerf_device.cuh -- device functions of the ErNeRF path, shared by the kernel-level
// drop-ins (mf_march_rays, mf_grid_encode_forward, ...) and the fused frame kernels, so both
// produce the same bits.
//
// Arithmetic follows the reference kernels expression by expression (see the citations);
// a*b+c contractions are written as explicit fmaf where nvcc -fmad=true contracts the
// reference source, and the mixed float/double sub-expressions of march_rays are kept.
#pragma once
#include "mf_common.cuh"

namespace ernerf {

// ---------------------------------------------------------------------------------------
// multiresolution grid levels (gridencoder.cu:124-126 + get_grid_index :54-72), precomputed
// ---------------------------------------------------------------------------------------
struct GridLevel {
    float scale;       // exp2f(level * S) * H - 1.0f, evaluated on the device at load time
    uint32_t stride1;  // resolution + 1  (align_corners = false)
    uint32_t offset;   // offsets[level] (rows)
    uint32_t size;     // hashmap_size = offsets[level+1] - offsets[level]
    uint32_t mode;     // bit0: y term used (stride1 <= size); bit1: hashed; bit2: size is 2^n
};

__device__ __forceinline__ uint32_t grid_index2(const GridLevel &lv, uint32_t x, uint32_t y) {
    uint32_t index = x;
    if (lv.mode & 1u) index += y * lv.stride1;
    if (lv.mode & 2u) index = x ^ (y * 2654435761u);  // fast_hash<2>, primes {1, 2654435761}
    if (lv.mode & 4u) return index & (lv.size - 1u);
    return index % lv.size;
}

// Index classes.  get_grid_index is a run-time function of (gridtype, resolution, hashmap_size) in the
// reference; per level it always reduces to one of three closed forms, selected here at compile time so
// the 144 corner indices of a sample cost 2-3 instructions each instead of a generic `%`:
//   IDX_DENSE : x + y * stride1            ((res+1)^2 <= size: no wrap possible)
//   IDX_HASH2 : (x ^ y * 2654435761) & (size - 1)   (hashed level, size = 2^n)
//   IDX_TILE2 : (x + y * stride1) & (size - 1)      (tiled level that wraps, size = 2^n)
//   IDX_ANY   : the generic form
enum { IDX_ANY = 0, IDX_DENSE = 1, IDX_HASH2 = 2, IDX_TILE2 = 3 };

__host__ __device__ __forceinline__ int grid_level_class(const GridLevel &lv) {
    const bool yterm = lv.mode & 1u, hashed = lv.mode & 2u, pow2 = lv.mode & 4u;
    if (hashed) return pow2 ? IDX_HASH2 : IDX_ANY;
    if (yterm && (uint64_t)lv.stride1 * lv.stride1 <= lv.size) return IDX_DENSE;
    if (yterm && pow2) return IDX_TILE2;
    return IDX_ANY;
}

template <int CLS>
__device__ __forceinline__ uint32_t grid_index2c(const GridLevel &lv, uint32_t x, uint32_t y) {
    if (CLS == IDX_DENSE) return x + y * lv.stride1;
    if (CLS == IDX_HASH2) return (x ^ (y * 2654435761u)) & (lv.size - 1u);
    if (CLS == IDX_TILE2) return (x + y * lv.stride1) & (lv.size - 1u);
    return grid_index2(lv, x, y);
}

// gridencoder.cu:75-175 for D = 2, C = 1, fp32 table: one level of one plane
template <int CLS>
__device__ __forceinline__ float grid_level_f32(const float *__restrict__ table, const GridLevel &lv, float u,
                                                float v) {
    if (u < 0.f || u > 1.f || v < 0.f || v > 1.f) return 0.f;
    const float *g = table + lv.offset;
    float pu = fmaf(u, lv.scale, 0.5f), pv = fmaf(v, lv.scale, 0.5f);
    const float flu = floorf(pu), flv = floorf(pv);
    const uint32_t gx = (uint32_t)flu, gy = (uint32_t)flv;
    pu -= (float)gx;
    pv -= (float)gy;
    const float v00 = __ldg(g + grid_index2c<CLS>(lv, gx, gy));
    const float v10 = __ldg(g + grid_index2c<CLS>(lv, gx + 1, gy));
    const float v01 = __ldg(g + grid_index2c<CLS>(lv, gx, gy + 1));
    const float v11 = __ldg(g + grid_index2c<CLS>(lv, gx + 1, gy + 1));
    float r = 0.f;
    r = fmaf((1 - pu) * (1 - pv), v00, r);
    r = fmaf(pu * (1 - pv), v10, r);
    r = fmaf((1 - pu) * pv, v01, r);
    r = fmaf(pu * pv, v11, r);
    return r;
}

// corner indices and fractions of one level (the first half of grid_level_f32 / grid_level_f16x2): lets a caller
// issue the loads of several levels back to back.  `base` (32-bit element offset of the level's table inside the caller's
// allocation, lv.offset + whatever precedes it) is folded into the indices HERE, in 32-bit arithmetic: added to the pointer
// by the caller it became a 64-bit multiply-add per corner (5 instructions per load instead of 2, ~10 % of k_head's
// instruction count, profiles/r02_k_head_v8 source page).  Same indices as grid_index2c corner by corner:
// (gx + 1) + gy * s == (gx + gy * s) + 1 and (gy + 1) * P == gy * P + P in uint32 arithmetic.
template <int CLS>
__device__ __forceinline__ void grid_level_prep(const GridLevel &lv, uint32_t base, float u, float v, uint32_t (&idx)[4], float &pu, float &pv) {
    pu = fmaf(u, lv.scale, 0.5f); pv = fmaf(v, lv.scale, 0.5f);
    const float flu = floorf(pu), flv = floorf(pv);
    const uint32_t gx = (uint32_t)flu, gy = (uint32_t)flv;
    pu -= (float)gx;
    pv -= (float)gy;
    if (CLS == IDX_DENSE) {
        const uint32_t i0 = base + gx + gy * lv.stride1, i2 = i0 + lv.stride1;
        idx[0] = i0; idx[1] = i0 + 1u; idx[2] = i2; idx[3] = i2 + 1u;
    } else if (CLS == IDX_HASH2) {
        const uint32_t m = lv.size - 1u, h0 = gy * 2654435761u, h1 = h0 + 2654435761u;
        idx[0] = base + ((gx ^ h0) & m);
        idx[1] = base + (((gx + 1u) ^ h0) & m);
        idx[2] = base + ((gx ^ h1) & m);
        idx[3] = base + (((gx + 1u) ^ h1) & m);
    } else {
        idx[0] = base + grid_index2c<CLS>(lv, gx, gy);
        idx[1] = base + grid_index2c<CLS>(lv, gx + 1, gy);
        idx[2] = base + grid_index2c<CLS>(lv, gx, gy + 1);
        idx[3] = base + grid_index2c<CLS>(lv, gx + 1, gy + 1);
    }
}

// same kernel instantiated for scalar_t = at::Half, C = 2 (torso encoder under autocast):
// `results[ch] += w * grid[..]` is Half += float: product rounded to half, sum rounded to half
template <int CLS>
__device__ __forceinline__ void grid_level_f16x2(const __half2 *__restrict__ table, const GridLevel &lv, float u,
                                                 float v, float &o0, float &o1) {
    o0 = 0.f;
    o1 = 0.f;
    if (u < 0.f || u > 1.f || v < 0.f || v > 1.f) return;
    const __half2 *g = table + lv.offset;
    float pu = fmaf(u, lv.scale, 0.5f), pv = fmaf(v, lv.scale, 0.5f);
    const uint32_t gx = (uint32_t)floorf(pu), gy = (uint32_t)floorf(pv);
    pu -= (float)gx;
    pv -= (float)gy;
    const float w[4] = {(1 - pu) * (1 - pv), pu * (1 - pv), (1 - pu) * pv, pu * pv};
    const uint32_t idx[4] = {grid_index2c<CLS>(lv, gx, gy), grid_index2c<CLS>(lv, gx + 1, gy),
                             grid_index2c<CLS>(lv, gx, gy + 1), grid_index2c<CLS>(lv, gx + 1, gy + 1)};
    __half2 val[4];
#pragma unroll
    for (int i = 0; i < 4; i++) val[i] = __ldg(g + idx[i]);
#pragma unroll
    for (int i = 0; i < 4; i++) {
        const float2 f = __half22float2(val[i]);
        o0 = round_half(o0 + round_half(w[i] * f.x));
        o1 = round_half(o1 + round_half(w[i] * f.y));
    }
}

// ---------------------------------------------------------------------------------------
// raymarching.cu:19-71 helpers
// ---------------------------------------------------------------------------------------
__device__ __forceinline__ float signf_(const float x) { return copysignf(1.0f, x); }
__device__ __forceinline__ float clampf_(const float x, const float lo, const float hi) {
    return fminf(hi, fmaxf(lo, x));
}
__device__ __forceinline__ int mip_from_pos(const float x, const float y, const float z, const float max_cascade) {
    const float mx = fmaxf(fabsf(x), fmaxf(fabsf(y), fabsf(z)));
    int exponent;
    frexpf(mx, &exponent);
    return fminf(max_cascade - 1, fmaxf(0, exponent));
}
__device__ __forceinline__ int mip_from_dt(const float dt, const float H, const float max_cascade) {
    const float mx = dt * H * 0.5;  // double literal, as in the reference
    int exponent;
    frexpf(mx, &exponent);
    return fminf(max_cascade - 1, fmaxf(0, exponent));
}
__device__ __forceinline__ uint32_t expand_bits(uint32_t v) {
    v = (v * 0x00010001u) & 0xFF0000FFu;
    v = (v * 0x00000101u) & 0x0F00F00Fu;
    v = (v * 0x00000011u) & 0xC30C30C3u;
    v = (v * 0x00000005u) & 0x49249249u;
    return v;
}
__device__ __forceinline__ uint32_t morton3D(uint32_t x, uint32_t y, uint32_t z) {
    return expand_bits(x) | (expand_bits(y) << 1) | (expand_bits(z) << 2);
}

// raymarching.cu:91-145
__device__ __forceinline__ void near_far_aabb(float ox, float oy, float oz, float dx, float dy, float dz,
                                              const float *aabb, float min_near, float &near_out, float &far_out) {
    const float rdx = 1 / dx, rdy = 1 / dy, rdz = 1 / dz;
    float near = (aabb[0] - ox) * rdx, far = (aabb[3] - ox) * rdx, tmp;
    if (near > far) { tmp = near; near = far; far = tmp; }
    float near_y = (aabb[1] - oy) * rdy, far_y = (aabb[4] - oy) * rdy;
    if (near_y > far_y) { tmp = near_y; near_y = far_y; far_y = tmp; }
    if (near > far_y || near_y > far) { near_out = far_out = 3.402823466e+38f; return; }
    if (near_y > near) near = near_y;
    if (far_y < far) far = far_y;
    float near_z = (aabb[2] - oz) * rdz, far_z = (aabb[5] - oz) * rdz;
    if (near_z > far_z) { tmp = near_z; near_z = far_z; far_z = tmp; }
    if (near > far_z || near_z > far) { near_out = far_out = 3.402823466e+38f; return; }
    if (near_z > near) near = near_z;
    if (far_z < far) far = far_z;
    if (near < min_near) near = min_near;
    near_out = near;
    far_out = far;
}

struct MarchParams {
    float bound, dt_gamma, dt_min, dt_max, rH, H3, Hf, Cf;
    float mip_bound0, mip_rbound0, halfH, rH2;  // level-0 constants of the fast path
    uint32_t H;
    bool fast;  // cascade == 1, H a power of two and dt_min == dt_max: level is always 0, the double sub-expression
                // 0.5 * (x * rbound + 1) * H is an exact power-of-two scaling (fp32 gives the same bits), and the step
                // clamp(t * dt_gamma, dt_min, dt_max) is the constant dt_max (max_steps <= H, every shipped configuration)
    bool linear;  // `grid` is a copy of the bitfield re-indexed as (z * H + y) * H + x (same bits, made at load time): the fused
                  // frame kernels test ~10^7 voxels per frame and the Morton expansion was a quarter of that loop's instructions
    const uint8_t *grid;
};

// __host__ too: the fused frame kernels take the finished struct as a kernel argument (constant bank, no registers); every field is a
// correctly rounded IEEE fp32 operation on both sides
__host__ __device__ __forceinline__ MarchParams make_march_params(float bound, float dt_gamma, uint32_t max_steps, uint32_t C,
                                                         uint32_t H, const uint8_t *grid) {
    MarchParams p;
    p.bound = bound;
    p.dt_gamma = dt_gamma;
    p.dt_max = 2 * 1.7320508075688772f * (1 << (C - 1)) / H;
    p.dt_min = fminf(p.dt_max, 2 * 1.7320508075688772f / max_steps);
    p.rH = 1 / (float)H;
    p.H3 = H * H * H;
    p.Hf = (float)H;
    p.Cf = (float)C;
    p.H = H;
    p.grid = grid;
    p.mip_bound0 = fminf(1.0f, bound);
    p.mip_rbound0 = 1 / p.mip_bound0;
    p.halfH = 0.5f * (float)H;
    p.rH2 = 2.0f * p.rH;
    p.fast = (C == 1) && ((H & (H - 1)) == 0) && p.dt_min == p.dt_max;
    p.linear = false;
    return p;
}

struct Ray {
    float ox, oy, oz, dx, dy, dz, rdx, rdy, rdz;
};

// body of the while-loop of kernel_march_rays (raymarching.cu:872-928): advance t until the
// next occupied voxel; returns true with the sample (x, y, z, dt) and t AT the sample (not yet advanced), or false
// once t >= far.  march_next = march_find + `t += dt`, the reference's own order of operations.
template <bool FAST>
__device__ __forceinline__ bool march_find_t(const MarchParams &p, const Ray &r, float &t, const float far, float &x,
                                             float &y, float &z, float &dt_out, uint32_t &vox) {
    if (FAST) {
        // Same values as the general loop below, fewer instructions per voxel (the ray pass of k_setup tests ~10^7 voxels per frame,
        // profiles/r02_k_setup_torso_v10.md):  dt is the constant dt_max;  (int)clamp(v, 0, hi) == (int)min(v, hi) for v > -1
        // (conversion truncates toward zero; v >= -H * 2^-24 because x is clamped to the bound);  nx + 0.5 + 0.5 * sign(d) and
        // (..) * rH * 2 - 1 are exact in fp32 (small integers, power-of-two scale), so they may be regrouped.
        const float dt = p.dt_max, hi = (float)(p.H - 1);
        // 0.5 + 0.5 * copysign(1, d) is 1 or 0 by the sign BIT of d: added to the voxel coordinate as an integer
        auto pos = [](float d) { return (int)((~__float_as_uint(d)) >> 31); };
        while (t < far) {
            x = clampf_(fmaf(t, r.dx, r.ox), -p.bound, p.bound);
            y = clampf_(fmaf(t, r.dy, r.oy), -p.bound, p.bound);
            z = clampf_(fmaf(t, r.dz, r.oz), -p.bound, p.bound);
            const int nx = fminf(fmaf(x, p.mip_rbound0, 1.0f) * p.halfH, hi);
            const int ny = fminf(fmaf(y, p.mip_rbound0, 1.0f) * p.halfH, hi);
            const int nz = fminf(fmaf(z, p.mip_rbound0, 1.0f) * p.halfH, hi);
            const uint32_t index = p.linear ? ((uint32_t)nz * p.H + (uint32_t)ny) * p.H + (uint32_t)nx : morton3D(nx, ny, nz);
            const bool occ = __ldg(p.grid + index / 8) & (1 << (index % 8));
            if (occ) {
                dt_out = dt;
                vox = index;
                return true;
            }
            const float tx = (fmaf((float)(nx + pos(r.dx)), p.rH2, -1.0f) * p.mip_bound0 - x) * r.rdx;
            const float ty = (fmaf((float)(ny + pos(r.dy)), p.rH2, -1.0f) * p.mip_bound0 - y) * r.rdy;
            const float tz = (fmaf((float)(nz + pos(r.dz)), p.rH2, -1.0f) * p.mip_bound0 - z) * r.rdz;
            const float tt = t + fmaxf(0.0f, fminf(tx, fminf(ty, tz)));
            do {
                t += dt;
            } while (t < tt);
        }
        return false;
    }
    while (t < far) {
        x = clampf_(fmaf(t, r.dx, r.ox), -p.bound, p.bound);
        y = clampf_(fmaf(t, r.dy, r.oy), -p.bound, p.bound);
        z = clampf_(fmaf(t, r.dz, r.oz), -p.bound, p.bound);
        const float dt = clampf_(t * p.dt_gamma, p.dt_min, p.dt_max);
        const int level = max(mip_from_pos(x, y, z, p.Cf), mip_from_dt(dt, p.Hf, p.Cf));
        const float mip_bound = fminf(scalbnf(1, level), p.bound);
        const float mip_rbound = 1 / mip_bound;
        const int nx = clampf_(0.5 * fmaf(x, mip_rbound, 1.0f) * p.H, 0.0f, (float)(p.H - 1));
        const int ny = clampf_(0.5 * fmaf(y, mip_rbound, 1.0f) * p.H, 0.0f, (float)(p.H - 1));
        const int nz = clampf_(0.5 * fmaf(z, mip_rbound, 1.0f) * p.H, 0.0f, (float)(p.H - 1));
        const uint32_t index = level * p.H3 + morton3D(nx, ny, nz);
        const bool occ = __ldg(p.grid + index / 8) & (1 << (index % 8));
        if (occ) {
            dt_out = dt;
            vox = index;
            return true;
        }
        const float tx = (((nx + 0.5f + 0.5f * signf_(r.dx)) * p.rH * 2 - 1) * mip_bound - x) * r.rdx;
        const float ty = (((ny + 0.5f + 0.5f * signf_(r.dy)) * p.rH * 2 - 1) * mip_bound - y) * r.rdy;
        const float tz = (((nz + 0.5f + 0.5f * signf_(r.dz)) * p.rH * 2 - 1) * mip_bound - z) * r.rdz;
        const float tt = t + fmaxf(0.0f, fminf(tx, fminf(ty, tz)));
        do {
            t += clampf_(t * p.dt_gamma, p.dt_min, p.dt_max);
        } while (t < tt);
    }
    return false;
}

template <bool FAST>
__device__ __forceinline__ bool march_next_t(const MarchParams &p, const Ray &r, float &t, const float far, float &x,
                                             float &y, float &z, float &dt_out, uint32_t &vox) {
    if (!march_find_t<FAST>(p, r, t, far, x, y, z, dt_out, vox)) return false;
    t += dt_out;
    return true;
}

// the general form (cascades, non power-of-two grids) stays out of line: it is not on the shipped model's path and would only
// dilute the instruction cache of the fused kernels
__device__ __noinline__ bool march_find_general(const MarchParams &p, const Ray &r, float &t, const float far, float &x,
                                                float &y, float &z, float &dt_out, uint32_t &vox) {
    return march_find_t<false>(p, r, t, far, x, y, z, dt_out, vox);
}
__device__ __forceinline__ bool march_find(const MarchParams &p, const Ray &r, float &t, const float far, float &x,
                                           float &y, float &z, float &dt_out, uint32_t &vox) {
    return p.fast ? march_find_t<true>(p, r, t, far, x, y, z, dt_out, vox) : march_find_general(p, r, t, far, x, y, z, dt_out, vox);
}
__device__ __forceinline__ bool march_next(const MarchParams &p, const Ray &r, float &t, const float far, float &x,
                                           float &y, float &z, float &dt_out, uint32_t &vox) {
    if (!march_find(p, r, t, far, x, y, z, dt_out, vox)) return false;
    t += dt_out;
    return true;
}

// shencoder.cu:43-68, degree 4
__device__ __forceinline__ void sh4(float x, float y, float z, float *o) {
    const float xy = x * y, xz = x * z, yz = y * z, x2 = x * x, y2 = y * y, z2 = z * z;
    o[0] = 0.28209479177387814f;
    o[1] = -0.48860251190291987f * y;
    o[2] = 0.48860251190291987f * z;
    o[3] = -0.48860251190291987f * x;
    o[4] = 1.0925484305920792f * xy;
    o[5] = -1.0925484305920792f * yz;
    o[6] = 0.94617469575755997f * z2 - 0.31539156525251999f;
    o[7] = -1.0925484305920792f * xz;
    o[8] = 0.54627421529603959f * x2 - 0.54627421529603959f * y2;
    o[9] = 0.59004358992664352f * y * (-3.0f * x2 + y2);
    o[10] = 2.8906114426405538f * xy * z;
    o[11] = 0.45704579946446572f * y * (1.0f - 5.0f * z2);
    o[12] = 0.3731763325901154f * z * (5.0f * z2 - 3.0f);
    o[13] = 0.45704579946446572f * x * (1.0f - 5.0f * z2);
    o[14] = 1.4453057213202769f * z * (x2 - y2);
    o[15] = 0.59004358992664352f * x * (-x2 + 3.0f * y2);
}

// freqencoder.cu:45-57, element c of the encoding of in[0..D)
__device__ __forceinline__ float freq_elem(const float *in, uint32_t D, uint32_t c) {
    if (c < D) return in[c];
    const uint32_t col = c / D - 1, d = c % D, freq = col / 2;
    const float phase_shift = (col % 2) * (3.141592653589793f / 2);
    return __sinf(in[d] * (float)(1u << freq) + phase_shift);   // scalbnf(in[d], freq): the same exact power-of-two scaling
}

// autocast(fp16) elementwise helpers
__device__ __forceinline__ float sigmoid16(float h16) {  // torch.sigmoid on an fp16 tensor
    return round_half(1.0f / (1.0f + expf(-h16)));
}
__device__ __forceinline__ float affine16(float s16) {  // s*(1+2*0.001) - 0.001, op by op in fp16
    return round_half(round_half(s16 * 1.002f) - 0.001f);
}

}  // namespace ernerf
