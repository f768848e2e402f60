// a5: class-wise confidence thresholds and hard pseudo-labels with ignore mask.
// Reference: uemda/gast/pseudo_generation.py:59-93 (pseudo_selection) and :24-56 (pseudo_selection1).
//
// HBM-bound integer/compare work: one pass for the per-(image,class) max (4c B/px read), one pass for
// the selection (4c B/px read + 8 B/px int64 write).  128-bit loads along the pixel dimension of each
// class plane, class vector in registers, thresholds in shared memory.
#include "uem_common.cuh"

int uem_select_entropy_stats_impl(const float* mask, const uint32_t* class_stats, int b, int c, int64_t hw, float cutoff_top,
                                  float cutoff_low, int64_t ignore_label, int64_t* out, const float* uvem, float* entropy,
                                  float* weight, int64_t* zero_after, int pdl, cudaStream_t st);

namespace {

constexpr int kThreads = 256;

__host__ __device__ inline int class_max_splits(int b, int c, int64_t hw) {
    // enough CTAs for >= 4 waves of 148 SMs, but at least 4096 pixels per CTA
    int64_t planes = (int64_t)b * c;
    int64_t want = (4 * UEM_SMS + planes - 1) / planes;
    int64_t cap = (hw + 4095) / 4096;
    int64_t s = want < cap ? want : cap;
    return (int)(s < 1 ? 1 : s);
}

// stage 1: partial max/min of one chunk of one (b,c) plane
template <int VEC>
__global__ void __launch_bounds__(kThreads) class_max_partial_kernel(const float* __restrict__ mask, int64_t hw,
                                                                     int splits, float* __restrict__ part) {
    const int plane = blockIdx.y, split = blockIdx.x;
    const float* p = mask + (int64_t)plane * hw;
    const int64_t groups = hw / VEC;
    const int64_t per = (groups + splits - 1) / splits;
    const int64_t g0 = split * per, g1 = min(g0 + per, groups);
    float mx = -INFINITY, mn = INFINITY;
    int nan = 0;
    for (int64_t g = g0 + threadIdx.x; g < g1; g += kThreads) {
        PixVec<VEC> v;
        v.load(p + g * VEC);
#pragma unroll
        for (int i = 0; i < VEC; ++i) {
            mx = fmaxf(mx, v.v[i]);
            mn = fminf(mn, v.v[i]);
            nan |= (v.v[i] != v.v[i]);
        }
    }
    mx = warp_max(mx);
    mn = warp_min(mn);
    nan = __any_sync(0xffffffffu, nan);
    __shared__ float smx[kThreads / 32], smn[kThreads / 32];
    __shared__ int snan[kThreads / 32];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    if (lane == 0) { smx[warp] = mx; smn[warp] = mn; snan[warp] = nan; }
    __syncthreads();
    if (threadIdx.x == 0) {
        for (int i = 1; i < kThreads / 32; ++i) { mx = fmaxf(mx, smx[i]); mn = fminf(mn, smn[i]); nan |= snan[i]; }
        float* o = part + ((int64_t)plane * splits + split) * 3;
        o[0] = mx; o[1] = mn; o[2] = nan ? 1.f : 0.f;
    }
}

// stage 2: fold the partials; torch.max/min propagate NaN
__global__ void class_max_final_kernel(const float* __restrict__ part, int planes, int splits,
                                       float* __restrict__ cmax, float* __restrict__ cmin, int32_t* __restrict__ has_nan) {
    int plane = blockIdx.x * blockDim.x + threadIdx.x;
    if (plane >= planes) return;
    float mx = -INFINITY, mn = INFINITY;
    bool nan = false;
    for (int s = 0; s < splits; ++s) {
        const float* o = part + ((int64_t)plane * splits + s) * 3;
        mx = fmaxf(mx, o[0]); mn = fminf(mn, o[1]); nan |= (o[2] != 0.f);
    }
    if (nan) { mx = NAN; mn = NAN; if (has_nan) atomicOr(has_nan, 1); }
    if (cmax) cmax[plane] = mx;
    if (cmin) cmin[plane] = mn;
}

// thr = max(fp32 max * fp32 cutoff_top, fp32 cutoff_low): one rounded multiply, no FMA
// (pseudo_generation.py:76-81; torch.maximum propagates NaN)
__device__ __forceinline__ float class_threshold(float cmax, float top, float low) {
    float t = __fmul_rn(cmax, top);
    return (t != t) ? t : fmaxf(t, low);
}

template <int C, int VEC>
__device__ __forceinline__ void select_body(const float* __restrict__ mask, const float* thr, int64_t hw, int64_t base,
                                            int64_t ignore_label, int variant, int64_t* __restrict__ out) {
    float p[C][VEC];
#pragma unroll
    for (int ci = 0; ci < C; ++ci) {
        PixVec<VEC> v;
        v.load(mask + (int64_t)ci * hw + base);
#pragma unroll
        for (int i = 0; i < VEC; ++i) p[ci][i] = v.v[i];
    }
    int64_t lab[VEC];
#pragma unroll
    for (int i = 0; i < VEC; ++i) {
        if (variant == 0) {
            // exactly one class strictly above its threshold (pseudo_generation.py:83-88)
            int wins = 0, first = 0;
#pragma unroll
            for (int ci = C - 1; ci >= 0; --ci) {
                bool w = p[ci][i] > thr[ci];
                wins += w;
                first = w ? ci : first;
            }
            lab[i] = (wins == 1) ? (int64_t)first : ignore_label;
        } else {
            // argmax (first index), dropped if below its class threshold (pseudo_generation.py:47-51)
            float best = p[0][i];
            int arg = 0;
#pragma unroll
            for (int ci = 1; ci < C; ++ci) {
                bool g = p[ci][i] > best;
                best = g ? p[ci][i] : best;
                arg = g ? ci : arg;
            }
            float t = thr[0];
#pragma unroll
            for (int ci = 1; ci < C; ++ci) t = (arg == ci) ? thr[ci] : t;
            lab[i] = (best < t) ? ignore_label : (int64_t)arg;
        }
    }
    store_ids<VEC>(out + base, lab);
}

template <int C, int VEC>
__global__ void __launch_bounds__(kThreads) select_kernel(const float* __restrict__ mask, const float* __restrict__ cmax,
                                                          int64_t hw, float top, float low, int64_t ignore_label,
                                                          int variant, int64_t* __restrict__ out) {
    __shared__ float thr[C];
    const int bi = blockIdx.y;
    if (threadIdx.x < C) thr[threadIdx.x] = class_threshold(cmax[bi * C + threadIdx.x], top, low);
    __syncthreads();
    const int64_t g = (int64_t)blockIdx.x * kThreads + threadIdx.x;
    if (g * VEC >= hw) return;
    select_body<C, VEC>(mask + (int64_t)bi * C * hw, thr, hw, g * VEC, ignore_label, variant,
                        out + (int64_t)bi * hw);
}

// thresholds of image bi from the (b, C+2) stats table the refine kernel raised atomically:
// [C ordered-encoded class maxima | ordered-encoded -(min) | bad flag]; an untouched slot (0) decodes to -inf and
// the bad flag (a NaN/inf row sum was produced) poisons every threshold like torch.max does
template <int C>
__device__ __forceinline__ void thresholds_from_stats(const unsigned* __restrict__ stats, int bi, float top, float low, float* thr) {
    if (threadIdx.x < C) {
        const unsigned* s = stats + (int64_t)bi * (C + 2);
        const unsigned e = s[threadIdx.x];
        float v = e ? ordered_to_f32(e) : -INFINITY;
        if (s[C + 1]) v = NAN;
        thr[threadIdx.x] = class_threshold(v, top, low);
    }
    __syncthreads();
}

// selection (+ optional entropy + UVEM weight) of the same map in one pass (the map is read once)
template <int C, int VEC, bool EXTRA>
__global__ void __launch_bounds__(kThreads) select_stats_kernel(const float* __restrict__ mask, const unsigned* __restrict__ stats,
                                                                int64_t hw, float top, float low, int64_t ignore_label,
                                                                int64_t* __restrict__ out, int has_uvem, float um, float ut,
                                                                float uig, float ucl, float ucr, float* __restrict__ entropy,
                                                                float* __restrict__ weight, int64_t* __restrict__ zero_after, int l2) {
    __shared__ float thr[C];
    const int bi = blockIdx.y;
    // programmatic dependent launch: the CTAs may already be resident while the refine kernel drains
    asm volatile("griddepcontrol.wait;" ::: "memory");
    if (zero_after && blockIdx.x == 0 && bi == 0 && threadIdx.x == 0) zero_after[0] = 0;  // max-id slot of the fused chain: clean for the next call
    thresholds_from_stats<C>(stats, bi, top, low, thr);
    const int64_t g = (int64_t)blockIdx.x * kThreads + threadIdx.x;
    if (g * VEC >= hw) return;
    const int64_t base = g * VEC;
    const float* mk = mask + (int64_t)bi * C * hw;
    if constexpr (!EXTRA) {
        select_body<C, VEC>(mk, thr, hw, base, ignore_label, 0, out + (int64_t)bi * hw);
    } else {
        // last reader of the refined map, only writer of the three outputs: everything here is touched once
        const uint64_t pol = l2_policy(l2);
        float p[C][VEC];
#pragma unroll
        for (int ci = 0; ci < C; ++ci) {
            PixVec<VEC> v;
            v.load(mk + (int64_t)ci * hw + base, pol);
#pragma unroll
            for (int i = 0; i < VEC; ++i) p[ci][i] = v.v[i];
        }
        int64_t lab[VEC];
        PixVec<VEC> ev, wv;
#pragma unroll
        for (int i = 0; i < VEC; ++i) {
            int wins = 0, first = 0;
            float px[C];
#pragma unroll
            for (int ci = C - 1; ci >= 0; --ci) {
                px[ci] = p[ci][i];
                const bool w = px[ci] > thr[ci];
                wins += w;
                first = w ? ci : first;
            }
            lab[i] = (wins == 1) ? (int64_t)first : ignore_label;
            const float u = entropy_px<C>(px);
            ev.v[i] = u;
            wv.v[i] = has_uvem ? uvem_weight_dev(u, um, ut, uig, ucl, ucr) : 1.0f;
        }
        store_ids<VEC>(out + (int64_t)bi * hw + base, lab, pol);
        if (entropy) ev.store(entropy + (int64_t)bi * hw + base, pol);
        if (weight) wv.store(weight + (int64_t)bi * hw + base, pol);
    }
}

// decode the stats table into plain floats: cmax/cmin (b*c... cmin is per image) for the host-side range assert
__global__ void class_stats_decode_kernel(const unsigned* __restrict__ stats, int b, int c, float* __restrict__ cmax,
                                          float* __restrict__ imin) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= b * (c + 1)) return;
    const int bi = i / (c + 1), j = i - bi * (c + 1);
    const unsigned* s = stats + (int64_t)bi * (c + 2);
    const bool bad = s[c + 1] != 0;
    const unsigned e = s[j];
    float v = e ? ordered_to_f32(e) : -INFINITY;
    if (j == c) v = -v;
    if (bad) v = NAN;
    if (j < c) cmax[bi * c + j] = v; else imin[bi] = v;
}

}  // namespace

extern "C" int uem_select_entropy_stats_f32(const float* mask, const uint32_t* class_stats, int b, int c, int64_t hw,
                                            float cutoff_top, float cutoff_low, int64_t ignore_label, int64_t* out,
                                            const float* uvem, float* entropy, float* weight, void* stream) {
    return uem_select_entropy_stats_impl(mask, class_stats, b, c, hw, cutoff_top, cutoff_low, ignore_label, out, uvem, entropy, weight,
                                         nullptr, 0, (cudaStream_t)stream);
}

namespace {
template <typename K, typename... Args>
int launch_maybe_pdl(K kernel, dim3 grid, int threads, cudaStream_t st, int pdl, Args... args) {
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = grid;
    cfg.blockDim = dim3(threads, 1, 1);
    cfg.stream = st;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = attr;
    cfg.numAttrs = pdl ? 1 : 0;
    UEM_CUDA(cudaLaunchKernelEx(&cfg, kernel, args...));
    return 0;
}
}  // namespace

// zero_after: optional int64 slot cleared by the kernel (fused chain housekeeping); pdl: launch with programmatic
// stream serialization (the kernel waits for the preceding kernel itself, its launch latency overlaps that kernel)
int uem_select_entropy_stats_impl(const float* mask, const uint32_t* class_stats, int b, int c, int64_t hw, float cutoff_top,
                                  float cutoff_low, int64_t ignore_label, int64_t* out, const float* uvem, float* entropy,
                                  float* weight, int64_t* zero_after, int pdl, cudaStream_t st) {
    UEM_REQUIRE(mask && class_stats && out && b > 0 && hw > 0, "uem_select_entropy_stats_f32: bad arguments");
    UEM_REQUIRE(!(weight && !uvem), "uem_select_entropy_stats_f32: weight output needs the uvem parameter block");
    const bool vec = uem_aligned16(mask) && uem_aligned16(out) && (hw % 4 == 0) && (!entropy || uem_aligned16(entropy)) &&
                     (!weight || uem_aligned16(weight));
    const int hu = uvem ? 1 : 0;
    const float um = hu ? uvem[0] : 0.f, ut = hu ? uvem[1] : 1.f, uig = hu ? uvem[2] : 1.f, ucl = hu ? uvem[3] : 0.f,
                ucr = hu ? uvem[4] : 0.f;
    const bool extra = entropy || weight;
    const unsigned* stats = class_stats;
    const int l2 = g_uem_l2_stream ? 1 : 0;
    int rc = 0;
    UEM_DISPATCH_C(c, {
        if (vec) {
            dim3 grid(uem_div_up(hw / 4, kThreads), b);
            if (extra)
                rc = launch_maybe_pdl(select_stats_kernel<C, 4, true>, grid, kThreads, st, pdl, mask, stats, hw, cutoff_top, cutoff_low,
                                      ignore_label, out, hu, um, ut, uig, ucl, ucr, entropy, weight, zero_after, l2);
            else
                rc = launch_maybe_pdl(select_stats_kernel<C, 4, false>, grid, kThreads, st, pdl, mask, stats, hw, cutoff_top, cutoff_low,
                                      ignore_label, out, hu, um, ut, uig, ucl, ucr, entropy, weight, zero_after, l2);
        } else {
            dim3 grid(uem_div_up(hw, kThreads), b);
            if (extra)
                rc = launch_maybe_pdl(select_stats_kernel<C, 1, true>, grid, kThreads, st, pdl, mask, stats, hw, cutoff_top, cutoff_low,
                                      ignore_label, out, hu, um, ut, uig, ucl, ucr, entropy, weight, zero_after, l2);
            else
                rc = launch_maybe_pdl(select_stats_kernel<C, 1, false>, grid, kThreads, st, pdl, mask, stats, hw, cutoff_top, cutoff_low,
                                      ignore_label, out, hu, um, ut, uig, ucl, ucr, entropy, weight, zero_after, l2);
        }
    });
    if (rc) return rc;
    UEM_CHECK_LAUNCH();
    return 0;
}

extern "C" int uem_class_stats_decode_f32(const uint32_t* class_stats, int b, int c, float* cmax, float* image_min, void* stream) {
    UEM_REQUIRE(class_stats && cmax && image_min && b > 0 && c > 0, "uem_class_stats_decode_f32: bad arguments");
    class_stats_decode_kernel<<<uem_div_up((int64_t)b * (c + 1), 128), 128, 0, (cudaStream_t)stream>>>(class_stats, b, c, cmax, image_min);
    UEM_CHECK_LAUNCH();
    return 0;
}

extern "C" int64_t uem_class_max_ws_bytes(int b, int c, int64_t hw) {
    return (int64_t)b * c * class_max_splits(b, c, hw) * 3 * sizeof(float);
}

extern "C" int uem_class_max_f32(const float* mask, int b, int c, int64_t hw, float* cmax, float* cmin,
                                 int32_t* has_nan, void* ws, void* stream) {
    UEM_REQUIRE(mask && ws && b > 0 && c > 0 && hw > 0, "uem_class_max_f32: bad arguments");
    cudaStream_t st = (cudaStream_t)stream;
    const int planes = b * c, splits = class_max_splits(b, c, hw);
    const bool vec = uem_aligned16(mask) && (hw % 4 == 0);
    dim3 grid(splits, planes);
    if (vec) class_max_partial_kernel<4><<<grid, kThreads, 0, st>>>(mask, hw, splits, (float*)ws);
    else class_max_partial_kernel<1><<<grid, kThreads, 0, st>>>(mask, hw, splits, (float*)ws);
    UEM_CHECK_LAUNCH();
    class_max_final_kernel<<<uem_div_up(planes, 128), 128, 0, st>>>((const float*)ws, planes, splits, cmax, cmin, has_nan);
    UEM_CHECK_LAUNCH();
    return 0;
}

extern "C" int uem_pseudo_select_f32(const float* mask, const float* cmax, int b, int c, int64_t hw, float cutoff_top,
                                     float cutoff_low, int64_t ignore_label, int variant, int64_t* out, void* stream) {
    UEM_REQUIRE(mask && cmax && out && b > 0 && hw > 0, "uem_pseudo_select_f32: bad arguments");
    UEM_REQUIRE(variant == 0 || variant == 1, "uem_pseudo_select_f32: variant must be 0 or 1");
    cudaStream_t st = (cudaStream_t)stream;
    const bool vec = uem_aligned16(mask) && uem_aligned16(out) && (hw % 4 == 0);
    UEM_DISPATCH_C(c, {
        if (vec) {
            dim3 grid(uem_div_up(hw / 4, kThreads), b);
            select_kernel<C, 4><<<grid, kThreads, 0, st>>>(mask, cmax, hw, cutoff_top, cutoff_low, ignore_label, variant, out);
        } else {
            dim3 grid(uem_div_up(hw, kThreads), b);
            select_kernel<C, 1><<<grid, kThreads, 0, st>>>(mask, cmax, hw, cutoff_top, cutoff_low, ignore_label, variant, out);
        }
    });
    UEM_CHECK_LAUNCH();
    return 0;
}
