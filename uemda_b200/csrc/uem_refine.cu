// a6: Aligner.label_refine as ONE full-resolution kernel, plus the fused refine->select chain.
// Reference: uemda/gast/alignment.py:194-293 (~60 eager kernels, ~58 full-res map passes):
//   prototype view  :215-223  1/pearson (low-res) -> bilinear up -> softmax(T=1) -> /max
//   prediction view :225-236  logits (low-res) -> bilinear up -> softmax(/temp) [mean of 2 heads] -> /max
//   superpixel view :238-258  region max of soft (torch_scatter 'max') -> gather -> softmax(/temp) -> /max,
//                             applied multiplicatively outside the batch-global "ignored" id (:241-243,:255)
//   final           :291-292  soft' = w*soft / (sum_c w*soft + 1e-7)
// Compulsory traffic per pixel: read soft (4c) + sup (8), write refined (4c); everything else (three
// low-res maps, the per-region weight table) is L2/L1 resident.
//
// Kernel shape (round-1 ncu: the first version was issue-bound at 752 instr/pixel and ran 1.7 waves, so this one
// is built to minimise instructions and to have no tail):
//   * persistent CTAs of 128 threads, one contiguous range of image rows each (grid = SMs x resident CTAs);
//   * a thread = 4 consecutive pixels; the soft tile and the superpixel ids go global->shared with cp.async
//     (no registers held while the weights are computed) and are read back by the same thread;
//   * per row the low-res maps are interpolated vertically once into shared memory, class-interleaved
//     ([x'][map][class padded to 4]) so a tap is two LDS.128;
//   * softmax followed by "/ (max + 1e-7)" collapses algebraically: max_c softmax = 1/S, hence
//     w_c = e_c / (1 + 1e-7*S) = e_c * (1 - 1e-7*S) to O(1e-13): no division; exp is one MUFU (ex2.approx.ftz);
//   * the default configuration (all views, two heads, >=3x up-sampling) runs a branch-free 3-tap form with
//     packed fp32x2 math (FFMA2/FMUL2/FADD2, new on sm_100): class pairs share an instruction;
//   * the superpixel view depends only on (image, region): softmax(/temp)/max of the region maxima is
//     precomputed per region by a tiny kernel, the pixel kernel just gathers 2 x LDG.128;
//   * per-(image,class) maxima for pseudo_selection (pseudo_generation.py:76) fall out as a (b, c+2) stats
//     table updated with one atomicMax per class per CTA.
#include "uem_common.cuh"

int uem_region_table_f32(const float* src, int64_t sb, int64_t sn, int64_t sc, const int64_t* index, int b, int64_t N,
                         int c, int64_t R, int op, const int64_t* hot_ptr, int64_t hot_val, int skip_hot, unsigned* table,
                         unsigned* cnt, int* status, cudaStream_t st);
int uem_i64_max_accumulate(const int64_t* x, int64_t n, int64_t* out_max, cudaStream_t st);

namespace {

constexpr int kRefineThreads = 128;

template <int C> struct Lay {
    static constexpr int CP = (C + 3) & ~3;   // class slots per map, padded for LDS.128 / LDG.128
    static constexpr int STRIDE = 3 * CP;     // floats per low-res column: [simi | pred1 | pred2]
    static constexpr int PC = (C + 1) / 2;    // class pairs for packed fp32x2 math
};

__device__ __forceinline__ float rcp_approx(float x) {
    float r;
    asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x));
    return r;
}
__device__ __forceinline__ float ex2_approx(float x) {
    float r;
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x));
    return r;
}
__device__ __forceinline__ void cp_async_4(void* smem, const void* gmem) {
    unsigned s = (unsigned)__cvta_generic_to_shared(smem);
    asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" ::"r"(s), "l"(gmem) : "memory");
}
__device__ __forceinline__ void cp_async_8(void* smem, const void* gmem) {
    unsigned s = (unsigned)__cvta_generic_to_shared(smem);
    asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" ::"r"(s), "l"(gmem) : "memory");
}

struct RefineParams {
    int views;
    const float* maps[3];   // slot 0 simi, 1 pred1, 2 pred2 (low-res (b,C,h,w)); null = unused
    float map_scale[3];     // folded into the staged rows (1/temp for the logits when temp is a power of two)
    int n_pred;             // prediction heads (0..2)
    int b, h, w, H, W;
    float sy, sx;
    float temp;
    int div_temp;           // temp is not a power of two: the staged logits are divided (one rounding) instead
    const int64_t* sup;
    const float* sw;        // (b,R,CP) per-region superpixel-view weights
    int64_t R;
    const int64_t* ignored_id;
    const float* soft;
    float* out;
    unsigned* stats;        // (b, C+2) ordered-u32: per-class max of `out`, -(min of out), bad flag; atomically raised
};

// e_c = exp(z_c - max z) (one MUFU each; results below 2^-126 flush to 0, irrelevant for a softmax numerator),
// returns S = sum e_c
template <int C> __device__ __forceinline__ float exp_shifted(float (&z)[C]) {
    float mx = z[0];
#pragma unroll
    for (int i = 1; i < C; ++i) mx = fmaxf(mx, z[i]);
    float s = 0.f;
#pragma unroll
    for (int i = 0; i < C; ++i) {
        z[i] = ex2_approx((z[i] - mx) * 1.4426950408889634f);
        s += z[i];
    }
    return s;
}
// packed variant over PC class pairs
template <int PC> __device__ __forceinline__ float exp_shifted2(float2 (&z)[PC]) {
    float mx = fmaxf(z[0].x, z[0].y);
#pragma unroll
    for (int j = 1; j < PC; ++j) mx = fmaxf(mx, fmaxf(z[j].x, z[j].y));
    const float2 nmx = make_float2(-mx, -mx), l2e = make_float2(1.4426950408889634f, 1.4426950408889634f);
    float2 acc = make_float2(0.f, 0.f);
#pragma unroll
    for (int j = 0; j < PC; ++j) {
        const float2 d = __fmul2_rn(__fadd2_rn(z[j], nmx), l2e);
        z[j] = make_float2(ex2_approx(d.x), ex2_approx(d.y));
        acc = __fadd2_rn(acc, z[j]);
    }
    return acc.x + acc.y;
}

template <int C> __device__ __forceinline__ void load_tap(const float* col, float (&t)[Lay<C>::CP]) {
#pragma unroll
    for (int q = 0; q < Lay<C>::CP / 4; ++q) {
        const float4 v = *reinterpret_cast<const float4*>(col + 4 * q);
        t[4 * q] = v.x; t[4 * q + 1] = v.y; t[4 * q + 2] = v.z; t[4 * q + 3] = v.w;
    }
}
template <int C>
__device__ __forceinline__ void load_tap3p(const float* col, int stride, float2 (&t)[3][Lay<C>::PC]) {
#pragma unroll
    for (int k3 = 0; k3 < 3; ++k3)
#pragma unroll
        for (int q = 0; q < Lay<C>::CP / 4; ++q) {
            const float4 v = *reinterpret_cast<const float4*>(col + k3 * stride + 4 * q);
            t[k3][2 * q] = make_float2(v.x, v.y);
            if (2 * q + 1 < Lay<C>::PC) t[k3][2 * q + 1] = make_float2(v.z, v.w);
        }
}

// vertical lerp of the low-res rows feeding output row y into row[(w+2)][STRIDE]:
// warp -> (map, class slot), lane -> low-res column: coalesced loads, no integer division by runtime values.
// Two replicated columns at the right edge (column a+2 of the 3-tap form always exists); padded class slots hold
// -1e30 so they exponentiate to exactly 0 and never win a max.
template <int C>
__device__ __forceinline__ void stage_row(float* row, const RefineParams& p, int bi, int y) {
    constexpr int CP = Lay<C>::CP, STRIDE = Lay<C>::STRIDE;
    const Lerp ly = make_lerp(y, p.h, p.sy);
    const int nwarps = blockDim.x >> 5, wid = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int hw_low = p.h * p.w;
    for (int mc = wid; mc < 3 * CP; mc += nwarps) {
        const int m = mc / CP, ci = mc - m * CP;
        if (p.maps[m] == nullptr) continue;
        if (ci >= C) {
            for (int x = lane; x < p.w + 2; x += 32) row[x * STRIDE + m * CP + ci] = -1e30f;
            continue;
        }
        const float* plane = p.maps[m] + ((int64_t)bi * C + ci) * hw_low;
        const float* r0 = plane + ly.i0 * p.w;
        const float* r1 = plane + ly.i1 * p.w;
        const float sc = p.map_scale[m];
        const bool dv = p.div_temp && m > 0;
        for (int x = lane; x < p.w + 2; x += 32) {
            const int xs = min(x, p.w - 1);
            const float v = ly.l0 * __ldg(r0 + xs) + ly.l1 * __ldg(r1 + xs);
            row[x * STRIDE + m * CP + ci] = dv ? __fdiv_rn(v, p.temp) : v * sc;
        }
    }
}

// block-reduce the running statistics of image bi and raise its stats row (one atomic per class per CTA)
template <int C>
__device__ __forceinline__ void flush_stats(unsigned* stats, int bi, float (&cmax)[C], float& cmin, bool& bad, float (*red)[C + 2]) {
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
#pragma unroll
    for (int ci = 0; ci < C; ++ci) {
        const float v = warp_max(cmax[ci]);
        if (lane == 0) red[warp][ci] = v;
    }
    const float mn = warp_min(cmin);
    const int anybad = __any_sync(0xffffffffu, bad);
    if (lane == 0) { red[warp][C] = mn; red[warp][C + 1] = anybad ? 1.f : 0.f; }
    __syncthreads();
    if (threadIdx.x <= C + 1) {
        const int nw = blockDim.x >> 5;
        float v = red[0][threadIdx.x];
        for (int i = 1; i < nw; ++i) v = (threadIdx.x == C) ? fminf(v, red[i][threadIdx.x]) : fmaxf(v, red[i][threadIdx.x]);
        unsigned* s = stats + (int64_t)bi * (C + 2);
        if (threadIdx.x < C) atomicMax(s + threadIdx.x, f32_to_ordered(v));
        else if (threadIdx.x == C) atomicMax(s + C, f32_to_ordered(-v));
        else if (v != 0.f) atomicOr(s + C + 1, 1u);
    }
    __syncthreads();
#pragma unroll
    for (int ci = 0; ci < C; ++ci) cmax[ci] = -INFINITY;
    cmin = INFINITY;
    bad = false;
}

// FAST: all three views, two heads, 128-bit path, >=3x up-sampling (branch-free packed 3-tap form).
// Otherwise: any view subset / one head / scalar path / any scale (2-tap scalar form, runtime view flags).
template <int C, int VEC, bool FAST>
__global__ void __launch_bounds__(kRefineThreads) refine_kernel(const RefineParams p) {
    constexpr int CP = Lay<C>::CP, STRIDE = Lay<C>::STRIDE, PC = Lay<C>::PC;
    extern __shared__ __align__(16) unsigned char smem_raw[];
    float* row = reinterpret_cast<float*>(smem_raw);                                        // [w+2][STRIDE]
    float* soft_s = row + (size_t)(p.w + 2) * STRIDE;                                       // [C][threads][VEC]
    int64_t* ids_s = reinterpret_cast<int64_t*>(soft_s + (size_t)C * kRefineThreads * VEC); // [threads][VEC]
    __shared__ float red[kRefineThreads / 32][C + 2];

    const int64_t HW = (int64_t)p.H * p.W;
    const bool vP = p.views & UEM_VIEW_PROTO, vL = p.views & UEM_VIEW_PRED, vS = p.views & UEM_VIEW_SUP;
    const int64_t ignored_id = vS ? *p.ignored_id : -1;
    const bool have_maps = vP || vL;

    float cmax[C];
#pragma unroll
    for (int ci = 0; ci < C; ++ci) cmax[ci] = -INFINITY;
    float cmin = INFINITY;
    bool bad = false;

    // persistent: CTA i owns the contiguous rows [R0, R1) of the flattened (image, row) space
    const int64_t total_rows = (int64_t)p.b * p.H;
    const int64_t R0 = total_rows * blockIdx.x / gridDim.x, R1 = total_rows * (blockIdx.x + 1) / gridDim.x;
    const int groups = p.W / VEC;
    int cur_b = -1;
    for (int64_t rr = R0; rr < R1; ++rr) {
        const int bi = (int)(rr / p.H), y = (int)(rr - (int64_t)bi * p.H);
        if (bi != cur_b) {
            if (cur_b >= 0 && p.stats) flush_stats<C>(p.stats, cur_b, cmax, cmin, bad, red);
            cur_b = bi;
        }
        const float* softb = p.soft + (int64_t)bi * C * HW;
        float* outb = p.out + (int64_t)bi * C * HW;
        const float* swb = p.sw + (int64_t)bi * p.R * CP;
        for (int g0 = 0; g0 < groups; g0 += kRefineThreads) {
            const int g = g0 + threadIdx.x;
            const bool active = g < groups;
            const int x0 = g * VEC;
            const int64_t px = (int64_t)y * p.W + x0;
            // ---- async: soft tile (+ ids) global -> shared, consumed by this same thread at the end
            if (active) {
#pragma unroll
                for (int ci = 0; ci < C; ++ci) {
                    float* dst = soft_s + ((size_t)ci * kRefineThreads + threadIdx.x) * VEC;
                    if constexpr (VEC == 4) cp_async_16(dst, softb + (int64_t)ci * HW + px);
                    else cp_async_4(dst, softb + (int64_t)ci * HW + px);
                }
                if (vS) {
                    int64_t* dst = ids_s + (size_t)threadIdx.x * VEC;
                    const int64_t* src = p.sup + (int64_t)bi * HW + px;
                    if constexpr (VEC == 4) { cp_async_16(dst, src); cp_async_16(dst + 2, src + 2); }
                    else cp_async_8(dst, src);
                }
            }
            cp_async_commit_group();
            if (have_maps && g0 == 0) {
                __syncthreads();  // previous row's readers are done
                stage_row<C>(row, p, bi, y);
                __syncthreads();
            }
            if (!active) { cp_async_wait_group<0>(); continue; }

            float wgt[C][VEC];
            if constexpr (FAST) {
                // horizontal taps: column a = i0 of the first pixel; pixel i uses columns a+d_i, a+d_i+1 with the
                // 3-tap weights (l0,l1,0) or (0,l0,l1): same products, same rounding as the 2-tap form
                float wa[VEC], wb[VEC], wc[VEC];
                const int a = make_lerp(x0, p.w, p.sx).i0;
#pragma unroll
                for (int i = 0; i < VEC; ++i) {
                    const Lerp lx = make_lerp(x0 + i, p.w, p.sx);
                    const bool d = lx.i0 != a;
                    wa[i] = d ? 0.f : lx.l0;
                    wb[i] = d ? lx.l0 : lx.l1;
                    wc[i] = d ? lx.l1 : 0.f;
                }
                const float* col = row + a * STRIDE;
                float2 wgt2[PC][VEC];
                {   // prototype view
                    float2 t[3][PC];
                    load_tap3p<C>(col, STRIDE, t);
#pragma unroll
                    for (int i = 0; i < VEC; ++i) {
                        const float2 a2 = make_float2(wa[i], wa[i]), b2 = make_float2(wb[i], wb[i]), c2 = make_float2(wc[i], wc[i]);
                        float2 z[PC];
#pragma unroll
                        for (int j = 0; j < PC; ++j) z[j] = __ffma2_rn(c2, t[2][j], __ffma2_rn(b2, t[1][j], __fmul2_rn(a2, t[0][j])));
                        const float S = exp_shifted2<PC>(z);
                        const float rs = fmaf(-1e-7f, S, 1.0f);
                        const float2 rs2 = make_float2(rs, rs);
#pragma unroll
                        for (int j = 0; j < PC; ++j) wgt2[j][i] = __fmul2_rn(z[j], rs2);
                    }
                }
                {   // prediction view, two heads
                    float2 t[3][PC], u[3][PC];
                    load_tap3p<C>(col + CP, STRIDE, t);
                    load_tap3p<C>(col + 2 * CP, STRIDE, u);
#pragma unroll
                    for (int i = 0; i < VEC; ++i) {
                        const float2 a2 = make_float2(wa[i], wa[i]), b2 = make_float2(wb[i], wb[i]), c2 = make_float2(wc[i], wc[i]);
                        float2 z[PC], z2[PC];
#pragma unroll
                        for (int j = 0; j < PC; ++j) {
                            z[j] = __ffma2_rn(c2, t[2][j], __ffma2_rn(b2, t[1][j], __fmul2_rn(a2, t[0][j])));
                            z2[j] = __ffma2_rn(c2, u[2][j], __ffma2_rn(b2, u[1][j], __fmul2_rn(a2, u[0][j])));
                        }
                        const float S1 = exp_shifted2<PC>(z);
                        const float S2 = exp_shifted2<PC>(z2);
                        const float h1 = 0.5f * rcp_approx(S1), h2 = 0.5f * rcp_approx(S2);
                        const float2 h1v = make_float2(h1, h1), h2v = make_float2(h2, h2);
                        float mx = 0.f;
#pragma unroll
                        for (int j = 0; j < PC; ++j) {
                            z[j] = __ffma2_rn(z[j], h1v, __fmul2_rn(z2[j], h2v));
                            mx = fmaxf(mx, fmaxf(z[j].x, z[j].y));
                        }
                        const float inv = rcp_approx(mx + 1e-7f);
                        const float2 inv2 = make_float2(inv, inv);
#pragma unroll
                        for (int j = 0; j < PC; ++j) wgt2[j][i] = __ffma2_rn(z[j], inv2, wgt2[j][i]);
                    }
                }
                cp_async_wait_group<0>();
                // superpixel view: multiplicative outside the ignored id
#pragma unroll
                for (int i = 0; i < VEC; ++i) {
                    const int64_t rid = ids_s[(size_t)threadIdx.x * VEC + i];
                    if (rid != ignored_id && (uint64_t)rid < (uint64_t)p.R) {
#pragma unroll
                        for (int q = 0; q < CP / 4; ++q) {
                            const float4 v = __ldg(reinterpret_cast<const float4*>(swb + rid * CP) + q);
                            wgt2[2 * q][i] = __fmul2_rn(wgt2[2 * q][i], make_float2(v.x, v.y));
                            if (2 * q + 1 < PC) wgt2[2 * q + 1][i] = __fmul2_rn(wgt2[2 * q + 1][i], make_float2(v.z, v.w));
                        }
                    }
                }
#pragma unroll
                for (int ci = 0; ci < C; ++ci)
#pragma unroll
                    for (int i = 0; i < VEC; ++i) wgt[ci][i] = (ci & 1) ? wgt2[ci >> 1][i].y : wgt2[ci >> 1][i].x;
            } else {
                bool have = false;
                if (vP) {  // prototype view: softmax(T=1) of the up-sampled 1/distance, peak-normalised
#pragma unroll
                    for (int i = 0; i < VEC; ++i) {
                        const Lerp lx = make_lerp(x0 + i, p.w, p.sx);
                        float t0[CP], t1[CP], z[C];
                        load_tap<C>(row + lx.i0 * STRIDE, t0);
                        load_tap<C>(row + lx.i1 * STRIDE, t1);
#pragma unroll
                        for (int ci = 0; ci < C; ++ci) z[ci] = lx.l0 * t0[ci] + lx.l1 * t1[ci];
                        const float S = exp_shifted<C>(z);
                        const float rs = fmaf(-1e-7f, S, 1.0f);
#pragma unroll
                        for (int ci = 0; ci < C; ++ci) wgt[ci][i] = z[ci] * rs;
                    }
                    have = true;
                }
                if (vL) {  // prediction view: softmax(logits/temp), mean of the heads, peak-normalised
#pragma unroll
                    for (int i = 0; i < VEC; ++i) {
                        const Lerp lx = make_lerp(x0 + i, p.w, p.sx);
                        float t0[CP], t1[CP], z[C], acc[C];
                        load_tap<C>(row + lx.i0 * STRIDE + CP, t0);
                        load_tap<C>(row + lx.i1 * STRIDE + CP, t1);
#pragma unroll
                        for (int ci = 0; ci < C; ++ci) z[ci] = lx.l0 * t0[ci] + lx.l1 * t1[ci];
                        const float S1 = exp_shifted<C>(z);
                        if (p.n_pred == 2) {
                            float z2[C];
                            load_tap<C>(row + lx.i0 * STRIDE + 2 * CP, t0);
                            load_tap<C>(row + lx.i1 * STRIDE + 2 * CP, t1);
#pragma unroll
                            for (int ci = 0; ci < C; ++ci) z2[ci] = lx.l0 * t0[ci] + lx.l1 * t1[ci];
                            const float S2 = exp_shifted<C>(z2);
                            const float h1 = 0.5f * rcp_approx(S1), h2 = 0.5f * rcp_approx(S2);
                            float mx = 0.f;
#pragma unroll
                            for (int ci = 0; ci < C; ++ci) { acc[ci] = fmaf(z[ci], h1, z2[ci] * h2); mx = fmaxf(mx, acc[ci]); }
                            const float inv = rcp_approx(mx + 1e-7f);
#pragma unroll
                            for (int ci = 0; ci < C; ++ci) acc[ci] *= inv;
                        } else {
                            const float rs = fmaf(-1e-7f, S1, 1.0f);
#pragma unroll
                            for (int ci = 0; ci < C; ++ci) acc[ci] = z[ci] * rs;
                        }
#pragma unroll
                        for (int ci = 0; ci < C; ++ci) wgt[ci][i] = have ? wgt[ci][i] + acc[ci] : acc[ci];
                    }
                    have = true;
                }
                cp_async_wait_group<0>();
                if (vS) {  // superpixel view: per-region weights, multiplicative outside the ignored id
#pragma unroll
                    for (int i = 0; i < VEC; ++i) {
                        const int64_t rid = ids_s[(size_t)threadIdx.x * VEC + i];
                        if (rid != ignored_id && (uint64_t)rid < (uint64_t)p.R) {
                            float sw[CP];
#pragma unroll
                            for (int q = 0; q < CP / 4; ++q) {
                                const float4 v = __ldg(reinterpret_cast<const float4*>(swb + rid * CP) + q);
                                sw[4 * q] = v.x; sw[4 * q + 1] = v.y; sw[4 * q + 2] = v.z; sw[4 * q + 3] = v.w;
                            }
#pragma unroll
                            for (int ci = 0; ci < C; ++ci) wgt[ci][i] = have ? wgt[ci][i] * sw[ci] : sw[ci];
                        } else if (!have) {
#pragma unroll
                            for (int ci = 0; ci < C; ++ci) wgt[ci][i] = 1.0f;
                        }
                    }
                }
            }
            // ---- soft' = w*soft / (sum + 1e-7)   (alignment.py:291-292, :324-325)
#pragma unroll
            for (int ci = 0; ci < C; ++ci) {
                const float* sp = soft_s + ((size_t)ci * kRefineThreads + threadIdx.x) * VEC;
                if constexpr (VEC == 4) {
                    const float4 v = *reinterpret_cast<const float4*>(sp);
                    wgt[ci][0] *= v.x; wgt[ci][1] *= v.y; wgt[ci][2] *= v.z; wgt[ci][3] *= v.w;
                } else {
                    wgt[ci][0] *= sp[0];
                }
            }
#pragma unroll
            for (int i = 0; i < VEC; ++i) {
                float s = 0.f;
#pragma unroll
                for (int ci = 0; ci < C; ++ci) s += wgt[ci][i];
                const float inv = rcp_approx(s + 1e-7f);
                bad |= !(s < INFINITY);
                float lo = INFINITY;
#pragma unroll
                for (int ci = 0; ci < C; ++ci) {
                    const float o = wgt[ci][i] * inv;
                    wgt[ci][i] = o;
                    cmax[ci] = fmaxf(cmax[ci], o);
                    lo = fminf(lo, o);
                }
                cmin = fminf(cmin, lo);
            }
#pragma unroll
            for (int ci = 0; ci < C; ++ci) {
                PixVec<VEC> t;
#pragma unroll
                for (int i = 0; i < VEC; ++i) t.v[i] = wgt[ci][i];
                t.store(outb + (int64_t)ci * HW + px);
            }
        }
    }
    if (cur_b >= 0 && p.stats) flush_stats<C>(p.stats, cur_b, cmax, cmin, bad, red);
}

// superpixel-view weight of every (image, region): softmax(region_max/temp) / (max + 1e-7)  (alignment.py:252-253)
// table: (b,R,C) ordered-u32 (encoded != 0) or fp32 region maxima -> sw (b,R,CP)
template <int C>
__global__ void __launch_bounds__(256) region_weight_kernel(const void* __restrict__ table, int encoded, int64_t regions, float temp,
                                                            float inv_temp, int div_temp, float* __restrict__ sw) {
    constexpr int CP = Lay<C>::CP;
    for (int64_t r = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; r < regions; r += (int64_t)gridDim.x * blockDim.x) {
        float z[C];
#pragma unroll
        for (int ci = 0; ci < C; ++ci) {
            float v;
            if (encoded) { const unsigned e = reinterpret_cast<const unsigned*>(table)[r * C + ci]; v = e ? ordered_to_f32(e) : 0.f; }
            else v = reinterpret_cast<const float*>(table)[r * C + ci];
            z[ci] = div_temp ? __fdiv_rn(v, temp) : v * inv_temp;
        }
        const float S = exp_shifted<C>(z);
        const float rs = fmaf(-1e-7f, S, 1.0f);
#pragma unroll
        for (int ci = 0; ci < CP; ++ci) sw[r * CP + ci] = ci < C ? z[ci < C ? ci : 0] * rs : 0.f;
    }
}

// a13: weight of each pixel's own hard-label class under the prototype view (alignment.py:295-309)
template <int C>
__global__ void __launch_bounds__(256) proto_weight_4pixel_kernel(const float* __restrict__ simi, int h, int w, int H, int W, float sy,
                                                                  float sx, const int64_t* __restrict__ hard, int64_t ignore_label,
                                                                  float eps, float* __restrict__ out) {
    extern __shared__ float prow[];
    const int y = blockIdx.x, bi = blockIdx.y;
    const Lerp ly = make_lerp(y, h, sy);
    for (int i = threadIdx.x; i < C * w; i += blockDim.x) {
        const int x = i % w, ci = i / w;
        const float* plane = simi + ((int64_t)bi * C + ci) * h * w;
        prow[i] = ly.l0 * __ldg(plane + (int64_t)ly.i0 * w + x) + ly.l1 * __ldg(plane + (int64_t)ly.i1 * w + x);
    }
    __syncthreads();
    for (int x = threadIdx.x; x < W; x += blockDim.x) {
        const Lerp lx = make_lerp(x, w, sx);
        float z[C];
#pragma unroll
        for (int ci = 0; ci < C; ++ci) z[ci] = lx.l0 * prow[ci * w + lx.i0] + lx.l1 * prow[ci * w + lx.i1];
        softmax_regs<C>(z);
        float mx = z[0];
#pragma unroll
        for (int ci = 1; ci < C; ++ci) mx = fmaxf(mx, z[ci]);
        const float inv = 1.0f / (mx + eps);
        const int64_t l = hard[((int64_t)bi * H + y) * W + x];
        float o = 0.f;
#pragma unroll
        for (int ci = 0; ci < C; ++ci) o = (l == ci && l != ignore_label) ? z[ci] * inv : o;
        out[((int64_t)bi * H + y) * W + x] = o;
    }
}

static inline int64_t align16(int64_t x) { return (x + 15) & ~(int64_t)15; }
static inline int cp_of(int c) { return (c + 3) & ~3; }

static int sm_count() {
    static int n = 0;
    if (n == 0) {
        int dev = 0;
        if (cudaGetDevice(&dev) != cudaSuccess || cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || n <= 0)
            n = UEM_SMS;
    }
    return n;
}

template <typename K>
static int launch_persistent(K kernel, const RefineParams& p, size_t smem, cudaStream_t st) {
    if (smem > 48 * 1024) UEM_CUDA(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    int per_sm = 0;
    UEM_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kernel, kRefineThreads, smem));
    if (per_sm < 1) per_sm = 1;
    const int64_t total_rows = (int64_t)p.b * p.H;
    const int grid = (int)min(total_rows, (int64_t)sm_count() * per_sm);
    kernel<<<grid, kRefineThreads, smem, st>>>(p);
    return 0;
}

// sw_ws: b*R*CP floats of scratch for the per-region weights (superpixel view only)
static int launch_refine(int views, const float* simi, const float* pred1, const float* pred2, int h, int w, const int64_t* sup,
                         const void* table, int table_encoded, int64_t R, const int64_t* ignored_id, const float* soft, int b,
                         int c, int H, int W, float temp, float* out, unsigned* stats, float* sw_ws, cudaStream_t st) {
    UEM_REQUIRE(soft && out && b > 0 && H > 0 && W > 0, "uem_label_refine_f32: bad arguments");
    UEM_REQUIRE(views > 0 && views < 8, "uem_label_refine_f32: views must be a non-empty mask of UEM_VIEW_*");
    UEM_REQUIRE(temp > 0.f, "uem_label_refine_f32: temp must be > 0");  // alignment.py:313
    RefineParams p{};
    p.views = views;
    int ex;
    const bool pow2 = (frexpf(temp, &ex) == 0.5f);
    for (int m = 0; m < 3; ++m) p.map_scale[m] = 1.0f;
    if (views & UEM_VIEW_PROTO) { UEM_REQUIRE(simi, "uem_label_refine_f32: prototype view needs simi"); p.maps[0] = simi; }
    if (views & UEM_VIEW_PRED) {
        UEM_REQUIRE(pred1, "uem_label_refine_f32: prediction view needs pred1");
        p.maps[1] = pred1;
        p.n_pred = 1;
        if (pred2) { p.maps[2] = pred2; p.n_pred = 2; }
        if (pow2) p.map_scale[1] = p.map_scale[2] = 1.0f / temp;  // exact: scaling by 2^k commutes with rounding
    }
    const bool maps = views & (UEM_VIEW_PROTO | UEM_VIEW_PRED);
    if (views & UEM_VIEW_SUP)
        UEM_REQUIRE(sup && table && ignored_id && R > 0 && sw_ws, "uem_label_refine_f32: superpixel view needs sup, region table, ignored id, ws");
    if (maps) UEM_REQUIRE(h > 0 && w > 0, "uem_label_refine_f32: bad low-res size");
    if (!maps) { w = 0; h = 0; }
    p.b = b; p.h = h; p.w = w; p.H = H; p.W = W;
    p.sy = uem_align_corners_scale(h, H);
    p.sx = uem_align_corners_scale(w, W);
    p.temp = temp;
    p.div_temp = pow2 ? 0 : 1;
    p.sup = sup; p.sw = sw_ws; p.R = R; p.ignored_id = ignored_id;
    p.soft = soft; p.out = out; p.stats = stats;
    const bool vec = (W % 4 == 0) && uem_aligned16(soft) && uem_aligned16(out) && (!sup || uem_aligned16(sup));
    const bool fast = vec && views == (UEM_VIEW_PROTO | UEM_VIEW_PRED | UEM_VIEW_SUP) && p.n_pred == 2 && 3.0f * p.sx <= 0.999f;
    void *ev0 = nullptr, *ev1 = nullptr;
    int launched = 1, rc = 0;
    UEM_DISPATCH_C(c, {
        if (views & UEM_VIEW_SUP) {
            const int64_t regions = (int64_t)b * R;
            region_weight_kernel<C><<<(int)min((int64_t)UEM_SMS * 4, (regions + 255) / 256), 256, 0, st>>>(
                table, table_encoded, regions, temp, 1.0f / temp, p.div_temp, sw_ws);
            launched = 2;
        }
        uem_take_profile_events(&ev0, &ev1);
        if (ev0) UEM_CUDA(cudaEventRecord((cudaEvent_t)ev0, st));
        const int vecw = vec ? 4 : 1;
        const size_t smem = (size_t)(w + 2) * Lay<C>::STRIDE * 4 + (size_t)C * kRefineThreads * vecw * 4 + (size_t)kRefineThreads * vecw * 8;
        UEM_REQUIRE(smem <= 200 * 1024, "uem_label_refine_f32: low-res width %d too large", w);
        if (fast) rc = launch_persistent(refine_kernel<C, 4, true>, p, smem, st);
        else if (vec) rc = launch_persistent(refine_kernel<C, 4, false>, p, smem, st);
        else rc = launch_persistent(refine_kernel<C, 1, false>, p, smem, st);
    });
    if (rc) return rc;
    if (ev1) UEM_CUDA(cudaEventRecord((cudaEvent_t)ev1, st));
    UEM_CHECK_LAUNCH_N(launched);
    return 0;
}

}  // namespace

extern "C" int64_t uem_class_stats_bytes(int b, int c) { return align16((int64_t)b * (c + 2) * 4); }

extern "C" int64_t uem_label_refine_ws_bytes(int b, int c, int64_t R) { return align16((int64_t)b * (R > 0 ? R : 1) * cp_of(c) * 4); }

extern "C" int uem_label_refine_f32(int views, const float* simi, const float* pred1, const float* pred2, int h, int w,
                                    const int64_t* sup, const float* region_max, int64_t R, const int64_t* ignored_id,
                                    const float* soft, int b, int c, int H, int W, float temp, float* out,
                                    uint32_t* class_stats, void* ws, void* stream) {
    return launch_refine(views, simi, pred1, pred2, h, w, sup, region_max, 0, R, ignored_id, soft, b, c, H, W, temp, out,
                         class_stats, (float*)ws, (cudaStream_t)stream);
}

extern "C" int uem_proto_weight_4pixel_f32(const float* simi, int h, int w, const int64_t* hard, int b, int c, int H, int W,
                                           int64_t ignore_label, float eps, float* out, void* stream) {
    UEM_REQUIRE(simi && hard && out && b > 0 && h > 0 && w > 0 && H > 0 && W > 0, "uem_proto_weight_4pixel_f32: bad arguments");
    const float sy = uem_align_corners_scale(h, H), sx = uem_align_corners_scale(w, W);
    UEM_DISPATCH_C(c, {
        size_t smem = (size_t)C * w * sizeof(float);
        UEM_REQUIRE(smem <= 48 * 1024, "uem_proto_weight_4pixel_f32: low-res width %d too large", w);
        dim3 grid(H, b);
        proto_weight_4pixel_kernel<C><<<grid, 256, smem, (cudaStream_t)stream>>>(simi, h, w, H, W, sy, sx, hard, ignore_label, eps, out);
    });
    UEM_CHECK_LAUNCH();
    return 0;
}

// ------------------------------------------------------------------------------------------------
// Fused chain: label_refine (all requested views) -> pseudo_selection (+ entropy / UVEM weight of the
// refined map), one C call, no host sync.  tools/train_ssl_uem.py:209-214, balance.py:372-396.
// ws layout (all 16-byte aligned):
//   [0]      status   int32[4]   bit1: label out of range, bit2: superpixel id outside [0,R)
//   [16]     simi     f32[b*c*h*w]
//            pearson ws (uem_pearson_ws_bytes)
//            sw       f32[b*R*CP]  per-region superpixel-view weights
//            --- zeroed by ONE memset per call ---
//            maxid    int64[2]     [0] = batch max superpixel id (alignment.py:241), ids are >= 0
//            stats    u32[b*(c+2)] class stats of the refined map (ordered-encoded, atomically raised)
//            region   u32[b*R*c]   ordered-encoded region maxima
// ------------------------------------------------------------------------------------------------
struct MineLayout {
    int64_t simi, pearson, sw, zero_begin, maxid, stats, region, end;
};
static MineLayout mine_layout(int b, int c, int h, int w, int k, int64_t R) {
    MineLayout L;
    int64_t n = 16;
    L.simi = n; n += align16((int64_t)b * c * h * w * 4);
    L.pearson = n; n += align16(uem_pearson_ws_bytes(c, k));
    L.sw = n; n += uem_label_refine_ws_bytes(b, c, R);
    L.zero_begin = n;
    L.maxid = n; n += 16;
    L.stats = n; n += uem_class_stats_bytes(b, c);
    L.region = n; n += align16((int64_t)b * R * c * 4);
    L.end = n;
    return L;
}

extern "C" int64_t uem_mine_ws_bytes(int b, int c, int H, int W, int h, int w, int k, int64_t R) {
    (void)H; (void)W;
    return mine_layout(b, c, h, w, k, R).end;
}
extern "C" int64_t uem_mine_ws_stats_offset(int b, int c, int H, int W, int h, int w, int k, int64_t R) {
    (void)H; (void)W;
    return mine_layout(b, c, h, w, k, R).stats;
}

extern "C" int uem_mine_refine_select_f32(int views, const float* feat, int k, const float* protos, const float* pred1,
                                          const float* pred2, int h, int w, const int64_t* sup, int64_t R,
                                          const int64_t* ignored_id, const float* soft, int b, int c, int H, int W, float temp,
                                          float eps, float cutoff_top, float cutoff_low, int64_t ignore_label, float* refined,
                                          int64_t* hard, const float* uvem /* NULL or host {m,t,1/gamma,coef_left,coef_right} */,
                                          float* entropy, float* weight, void* ws, void* stream) {
    UEM_REQUIRE(ws && soft && refined, "uem_mine_refine_select_f32: bad arguments");
    cudaStream_t st = (cudaStream_t)stream;
    char* base = (char*)ws;
    const MineLayout L = mine_layout(b, c, h, w, k, R);
    int* status = (int*)base;
    float* simi = (float*)(base + L.simi);
    void* pws = base + L.pearson;
    float* sw = (float*)(base + L.sw);
    int64_t* maxid = (int64_t*)(base + L.maxid);
    unsigned* stats = (unsigned*)(base + L.stats);
    unsigned* table = (unsigned*)(base + L.region);
    const int64_t HW = (int64_t)H * W;
    int rc;
    UEM_CUDA(cudaMemsetAsync(base + L.zero_begin, 0, (size_t)(L.end - L.zero_begin), st));
    // order matters for L2: the feature pass runs first so that soft/sup, read by the region-max pass, are
    // still L2 resident when the refine kernel reads them again
    if (views & UEM_VIEW_PROTO) {
        UEM_REQUIRE(feat && protos, "uem_mine_refine_select_f32: prototype view needs feat and prototypes");
        if ((rc = uem_pearson_dist_nchw_f32(feat, b, k, (int64_t)h * w, protos, c, eps, 1, simi, pws, stream))) return rc;
    }
    if (views & UEM_VIEW_SUP) {
        UEM_REQUIRE(sup && R > 0, "uem_mine_refine_select_f32: superpixel view needs sup and a region capacity R");
        if (!ignored_id) {
            if ((rc = uem_i64_max_accumulate(sup, (int64_t)b * HW, maxid, st))) return rc;
            ignored_id = maxid;
        }
        // region maxima of soft, NCHW viewed as (b,N,c): strides {c*N, 1, N}; the ignored id is never gathered
        // (alignment.py:255), so its pixels are skipped
        if ((rc = uem_region_table_f32(soft, (int64_t)c * HW, 1, HW, sup, b, HW, c, R, UEM_REDUCE_MAX, ignored_id, -1, 1, table,
                                       nullptr, status, st)))
            return rc;
    }
    if ((rc = launch_refine(views, simi, pred1, pred2, h, w, sup, table, 1, R, ignored_id, soft, b, c, H, W, temp, refined,
                            stats, sw, st)))
        return rc;
    if (hard || entropy || weight) {
        UEM_REQUIRE(hard, "uem_mine_refine_select_f32: entropy/weight outputs come with the selection (hard must be given)");
        UEM_REQUIRE(!(weight && !uvem), "uem_mine_refine_select_f32: weight output needs the uvem parameter block");
        if ((rc = uem_select_entropy_stats_f32(refined, stats, b, c, HW, cutoff_top, cutoff_low, ignore_label, hard, uvem, entropy,
                                               weight, stream)))
            return rc;
    }
    return 0;
}
