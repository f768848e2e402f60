// f4: IAST class-wise percentile thresholds from per-class confidence histograms, and the sliding-window
// accumulation that precedes pseudo-label generation.  Reference: uemda/utils/tools.py
//   :323-333  ias_thresh        per class: np.percentile(confidences, 100 * (1 - alpha * w_c ** gamma)), linear interpolation
//   :347-371  generate_pseudo   confidences = max prob of the pixels whose argmax is the class, CAST TO float16, with the
//                               previous threshold as one extra sample; thresholds move by an EMA; label = argmax + 1,
//                               0 where the winning confidence is below its class threshold
//   :61-97    pre_slide         half-overlapping windows summed into a full map, divided by the cover count
//   :132-152  tta_predict       mean of the de-augmented views
// Because the reference casts the confidences to float16, an EXACT histogram over the 65536 half bit patterns (ordered
// like the numbers) holds the whole sample: the percentile's two order statistics are read off its prefix sum and
// interpolated with numpy's own fp64 formula (_lerp), so the thresholds come out bit for bit -- no sort, no host lists.
// Histogram: max / argmax per pixel from 128-bit planar loads, keys = (class, half bits); the bins of [0.125, 1] (3073 per
// class: every confidence a softmax over <= 8 classes can produce) live in a shared-memory table per CTA, raised with
// warp-aggregated atomics (__match_any_sync: the confident pixels of a warp share one key), everything else goes to the
// global table directly; CTAs merge their non-zero bins at the end.
#include "uem_common.cuh"
#include <cuda_fp16.h>

namespace {

constexpr int kBins = 65536;
constexpr int kWinLo = 0x8000 + 0x3000;   // ordered index of half(0.125)
constexpr int kWinN = 0x3C00 - 0x3000 + 1;   // .. half(1.0) inclusive: 3073 bins

// half bits -> index that sorts like the value (negatives descending magnitude, then positives); NaN last
__device__ __forceinline__ unsigned ordered_half(unsigned bits) {
    const unsigned mag = bits & 0x7FFFu;
    if (mag > 0x7C00u) return 0xFFFFu;
    return (bits & 0x8000u) ? (0x7FFFu - mag) : (0x8000u + mag);
}
__device__ __forceinline__ double ordered_half_value(unsigned idx) {
    const unsigned bits = idx >= 0x8000u ? (idx - 0x8000u) : (0x8000u | (0x7FFFu - idx));
    return (double)__half2float(__ushort_as_half((unsigned short)bits));
}

template <int C, int VEC>
__global__ void __launch_bounds__(512) iast_hist_kernel(const float* __restrict__ probs, int64_t hw, int64_t total_groups,
                                                        unsigned* __restrict__ hist) {
    extern __shared__ unsigned win[];   // [C][kWinN]
    for (int i = threadIdx.x; i < C * kWinN; i += blockDim.x) win[i] = 0u;
    __syncthreads();
    const int64_t groups_per_img = hw / VEC;
    for (int64_t g = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; g < total_groups + ((-total_groups) & 31);
         g += (int64_t)gridDim.x * blockDim.x) {
        const bool live = g < total_groups;   // whole warps iterate together (the match below is warp-wide)
        float best[VEC];
        int arg[VEC];
        if (live) {
            const int64_t bi = g / groups_per_img, px = (g - bi * groups_per_img) * VEC;
            const float* base = probs + bi * C * hw + px;
#pragma unroll
            for (int ci = 0; ci < C; ++ci) {
                PixVec<VEC> v;
                v.load(base + (int64_t)ci * hw);
#pragma unroll
                for (int i = 0; i < VEC; ++i)   // torch.max(dim=1): the first maximum wins, a NaN is a maximum
                    if (ci == 0 || v.v[i] > best[i] || (v.v[i] != v.v[i] && best[i] == best[i])) { best[i] = v.v[i]; arg[i] = ci; }
            }
        }
#pragma unroll
        for (int i = 0; i < VEC; ++i) {
            unsigned key = 0xFFFFFFFFu;
            if (live) key = ((unsigned)arg[i] << 16) | ordered_half((unsigned)__half_as_ushort(__float2half_rn(best[i])));
            const unsigned peers = __match_any_sync(0xffffffffu, key);
            if (live && (int)(__ffs(peers) - 1) == (int)(threadIdx.x & 31)) {
                const unsigned n = __popc(peers), idx = key & 0xFFFFu, ci = key >> 16;
                if (idx >= (unsigned)kWinLo && idx < (unsigned)(kWinLo + kWinN)) atomicAdd(&win[ci * kWinN + (idx - kWinLo)], n);
                else atomicAdd(&hist[(size_t)ci * kBins + idx], n);
            }
        }
    }
    __syncthreads();
    for (int i = threadIdx.x; i < C * kWinN; i += blockDim.x) {
        const unsigned n = win[i];
        if (n) {
            const int ci = i / kWinN, j = i - ci * kWinN;
            atomicAdd(&hist[(size_t)ci * kBins + kWinLo + j], n);
        }
    }
}

struct IastQ {
    double qfrac[UEM_MAX_C];   // q / 100 per class, computed on the host with Python / numpy arithmetic
};

// one CTA per class: prefix sum over the 65536 ordered bins, the extra sample (the previous threshold) merged in, numpy's
// linear-interpolation percentile in fp64, then the EMA of generate_pseudo (tools.py:357-361)
__global__ void __launch_bounds__(1024) iast_threshold_kernel(const unsigned* __restrict__ hist, const IastQ q, double beta,
                                                              float one_minus_beta, double* __restrict__ cls_thresh,
                                                              float* __restrict__ tmp_out, long long* __restrict__ count_out) {
    const int ci = blockIdx.x;
    const unsigned* h = hist + (size_t)ci * kBins;
    const double extra = cls_thresh[ci];
    constexpr int PER = kBins / 1024;   // 64 consecutive bins per thread
    const int b0 = threadIdx.x * PER;
    unsigned long long local = 0, below = 0;
    for (int j = 0; j < PER; ++j) {
        const unsigned n = h[b0 + j];
        local += n;
        if (b0 + j != 0xFFFF && ordered_half_value(b0 + j) < extra) below += n;
    }
    // block exclusive scan of `local`, block sum of `below`
    __shared__ unsigned long long wsum[32], wbelow[32];
    __shared__ unsigned long long s_total, s_below;
    __shared__ double s_val[2];
    __shared__ int s_nan;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    unsigned long long incl = local, bl = below;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        const unsigned long long t = __shfl_up_sync(0xffffffffu, incl, o);
        if (lane >= o) incl += t;
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) bl += __shfl_xor_sync(0xffffffffu, bl, o);
    if (lane == 31) wsum[warp] = incl;
    if (lane == 0) wbelow[warp] = bl;
    if (threadIdx.x == 0) s_nan = 0;
    __syncthreads();
    if (warp == 0) {
        unsigned long long v = wsum[lane], w2 = wbelow[lane];
        unsigned long long inc2 = v;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const unsigned long long t = __shfl_up_sync(0xffffffffu, inc2, o);
            if (lane >= o) inc2 += t;
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) w2 += __shfl_xor_sync(0xffffffffu, w2, o);
        wsum[lane] = inc2 - v;   // exclusive prefix of the warps
        if (lane == 31) s_total = inc2;
        if (lane == 0) s_below = w2;
    }
    __syncthreads();
    const unsigned long long excl = wsum[warp] + (incl - local);
    const unsigned long long n_hist = s_total, rank_e = s_below;
    const long long n = (long long)n_hist + 1;   // + the previous threshold
    // numpy: virtual = (n - 1) * q; previous = floor(virtual), next = previous + 1, both -> n - 1 when virtual >= n - 1
    const double virt = __dmul_rn((double)(n - 1), q.qfrac[ci]);
    long long prev = (long long)floor(virt), next = prev + 1;
    const bool above = virt >= (double)(n - 1), below0 = virt < 0.0;
    if (above) prev = next = n - 1;
    if (below0) prev = next = 0;
    // merged order statistic k: the histogram's sorted samples with `extra` inserted at position rank_e
    const long long want[2] = {prev, next};
#pragma unroll
    for (int t = 0; t < 2; ++t) {
        const long long k = want[t];
        if (k == (long long)rank_e) { if (threadIdx.x == 0) s_val[t] = extra; continue; }
        const unsigned long long j = (unsigned long long)(k < (long long)rank_e ? k : k - 1);   // index into the histogram's samples
        if (j >= excl && j < excl + local) {
            unsigned long long run = excl;
            for (int u = 0; u < PER; ++u) {
                const unsigned nb = h[b0 + u];
                if (j < run + nb) { s_val[t] = ordered_half_value(b0 + u); break; }
                run += nb;
            }
        }
    }
    if (threadIdx.x == 1023 && h[0xFFFF]) s_nan = 1;   // NaN confidences sort last and poison the percentile (numpy)
    __syncthreads();
    if (threadIdx.x == 0) {
        const double a = s_val[0], bq = s_val[1];
        const double gamma = above ? __dsub_rn(virt, -1.0) : __dsub_rn(virt, (double)prev);   // numpy subtracts the CLIPPED index (-1)
        const double diff = __dsub_rn(bq, a);
        double r = __dadd_rn(a, __dmul_rn(diff, gamma));
        if (gamma >= 0.5) r = __dsub_rn(bq, __dmul_rn(diff, __dsub_rn(1.0, gamma)));   // numpy _lerp
        if (s_nan) r = nan("");
        const float tmp = (float)r;                                   // cls_thresh is a float32 array in ias_thresh (:327)
        if (tmp_out) tmp_out[ci] = tmp;
        if (count_out) count_out[ci] = (long long)n_hist;
        // tools.py:360-361: beta * cls_thresh (float64) + (1 - beta) * tmp (float32 product), then the clamp
        double t = __dadd_rn(__dmul_rn(beta, extra), (double)__fmul_rn(one_minus_beta, tmp));
        if (t >= 1.0) t = 0.999;
        cls_thresh[ci] = t;
    }
}

template <int C, int VEC>
__global__ void __launch_bounds__(256) iast_label_kernel(const float* __restrict__ probs, int64_t hw, int64_t total_groups,
                                                         const double* __restrict__ cls_thresh, uint8_t* __restrict__ out) {
    __shared__ double thr[C];
    if (threadIdx.x < C) thr[threadIdx.x] = cls_thresh[threadIdx.x];
    __syncthreads();
    const int64_t groups_per_img = hw / VEC;
    for (int64_t g = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; g < total_groups; g += (int64_t)gridDim.x * blockDim.x) {
        const int64_t bi = g / groups_per_img, px = (g - bi * groups_per_img) * VEC;
        const float* base = probs + bi * C * hw + px;
        float best[VEC];
        int arg[VEC];
#pragma unroll
        for (int ci = 0; ci < C; ++ci) {
            PixVec<VEC> v;
            v.load(base + (int64_t)ci * hw);
#pragma unroll
            for (int i = 0; i < VEC; ++i)   // np.argmax / np.amax: the first maximum wins, a NaN is a maximum
                if (ci == 0 || v.v[i] > best[i] || (v.v[i] != v.v[i] && best[i] == best[i])) { best[i] = v.v[i]; arg[i] = ci; }
        }
        uint8_t o[VEC];
#pragma unroll
        for (int i = 0; i < VEC; ++i) o[i] = ((double)best[i] < thr[arg[i]]) ? (uint8_t)0 : (uint8_t)(arg[i] + 1);   // :369-372
        uint8_t* dst = out + bi * hw + px;
        if constexpr (VEC == 4) *reinterpret_cast<uchar4*>(dst) = make_uchar4(o[0], o[1], o[2], o[3]);
        else dst[0] = o[0];
    }
}

// full[:, :, y1:y2, x1:x2] += tile[:, :, :y2-y1, :x2-x1]; count[:, :, y1:y2, x1:x2] += 1   (tools.py:94-95)
__global__ void __launch_bounds__(256) window_accumulate_kernel(float* __restrict__ full, float* __restrict__ count,
                                                                const float* __restrict__ tile, int b, int c, int H, int W, int th,
                                                                int tw, int y1, int x1, int hh, int ww) {
    const int64_t total = (int64_t)b * (c + 1) * hh * ww;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
        const int x = (int)(i % ww);
        const int y = (int)((i / ww) % hh);
        const int ch = (int)((i / ((int64_t)ww * hh)) % (c + 1));
        const int bi = (int)(i / ((int64_t)ww * hh * (c + 1)));
        if (ch == c) {
            if (count) count[((int64_t)bi * H + y1 + y) * W + x1 + x] += 1.0f;
        } else {
            full[(((int64_t)bi * c + ch) * H + y1 + y) * W + x1 + x] += tile[(((int64_t)bi * c + ch) * th + y) * tw + x];
        }
    }
}
// full /= count (broadcast over classes)   (tools.py:97)
__global__ void __launch_bounds__(256) window_average_kernel(float* __restrict__ full, const float* __restrict__ count, int b, int c,
                                                             int64_t hw) {
    const int64_t total = (int64_t)b * c * hw;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
        const int64_t px = i % hw, bi = i / (c * hw);
        full[i] = __fdiv_rn(full[i], count[bi * hw + px]);
    }
}
// mean of n stacked views (n, numel): sum in view order, then / n   (tools.py:149-150)
__global__ void __launch_bounds__(256) views_mean_kernel(const float* __restrict__ views, int n, int64_t numel, float* __restrict__ out) {
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < numel; i += (int64_t)gridDim.x * blockDim.x) {
        float s = views[i];
        for (int v = 1; v < n; ++v) s += views[(int64_t)v * numel + i];
        out[i] = s / (float)n;
    }
}

}  // namespace

extern "C" int64_t uem_iast_hist_bytes(int c) { return (int64_t)c * kBins * 4; }

extern "C" int uem_iast_conf_hist_f32(const float* probs, int b, int c, int64_t hw, uint32_t* hist, void* stream) {
    UEM_REQUIRE(probs && hist && b > 0 && hw > 0, "uem_iast_conf_hist_f32: bad arguments");
    cudaStream_t st = (cudaStream_t)stream;
    UEM_CUDA(cudaMemsetAsync(hist, 0, (size_t)uem_iast_hist_bytes(c), st));
    const bool vec = (hw % 4 == 0) && uem_aligned16(probs);
    UEM_DISPATCH_C(c, {
        const size_t smem = (size_t)C * kWinN * 4;
        if (vec) {
            const int64_t groups = (int64_t)b * (hw / 4);
            UEM_CUDA(cudaFuncSetAttribute(iast_hist_kernel<C, 4>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
            iast_hist_kernel<C, 4><<<(int)min((int64_t)UEM_SMS * 2, (groups + 511) / 512), 512, smem, st>>>(probs, hw, groups, hist);
        } else {
            const int64_t groups = (int64_t)b * hw;
            UEM_CUDA(cudaFuncSetAttribute(iast_hist_kernel<C, 1>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
            iast_hist_kernel<C, 1><<<(int)min((int64_t)UEM_SMS * 2, (groups + 511) / 512), 512, smem, st>>>(probs, hw, groups, hist);
        }
    });
    UEM_CHECK_LAUNCH();
    return 0;
}

extern "C" int uem_iast_thresholds_f64(const uint32_t* hist, int c, const double* qfrac_host, double beta, float one_minus_beta,
                                       double* cls_thresh, float* tmp_out, int64_t* count_out, void* stream) {
    UEM_REQUIRE(hist && qfrac_host && cls_thresh && c > 0 && c <= UEM_MAX_C, "uem_iast_thresholds_f64: bad arguments");
    IastQ q;
    for (int i = 0; i < UEM_MAX_C; ++i) q.qfrac[i] = i < c ? qfrac_host[i] : 0.0;
    for (int i = 0; i < c; ++i)
        UEM_REQUIRE(q.qfrac[i] >= 0.0 && q.qfrac[i] <= 1.0, "Percentiles must be in the range [0, 100]");   // numpy's ValueError
    iast_threshold_kernel<<<c, 1024, 0, (cudaStream_t)stream>>>(hist, q, beta, one_minus_beta, cls_thresh, tmp_out,
                                                                (long long*)count_out);
    UEM_CHECK_LAUNCH();
    return 0;
}

extern "C" int uem_iast_labels_u8(const float* probs, int b, int c, int64_t hw, const double* cls_thresh, uint8_t* out, void* stream) {
    UEM_REQUIRE(probs && cls_thresh && out && b > 0 && hw > 0, "uem_iast_labels_u8: bad arguments");
    cudaStream_t st = (cudaStream_t)stream;
    const bool vec = (hw % 4 == 0) && uem_aligned16(probs) && ((reinterpret_cast<uintptr_t>(out) & 3u) == 0);
    UEM_DISPATCH_C(c, {
        if (vec) {
            const int64_t groups = (int64_t)b * (hw / 4);
            iast_label_kernel<C, 4><<<(int)min((int64_t)UEM_SMS * 8, (groups + 255) / 256), 256, 0, st>>>(probs, hw, groups, cls_thresh, out);
        } else {
            const int64_t groups = (int64_t)b * hw;
            iast_label_kernel<C, 1><<<(int)min((int64_t)UEM_SMS * 8, (groups + 255) / 256), 256, 0, st>>>(probs, hw, groups, cls_thresh, out);
        }
    });
    UEM_CHECK_LAUNCH();
    return 0;
}

extern "C" int uem_window_accumulate_f32(float* full, float* count, const float* tile, int b, int c, int H, int W, int th, int tw,
                                         int y1, int x1, int y2, int x2, void* stream) {
    UEM_REQUIRE(full && tile && b > 0 && c > 0 && y1 >= 0 && x1 >= 0 && y2 <= H && x2 <= W && y2 > y1 && x2 > x1 && y2 - y1 <= th &&
                    x2 - x1 <= tw,
                "uem_window_accumulate_f32: bad window");
    const int hh = y2 - y1, ww = x2 - x1;
    const int64_t total = (int64_t)b * (c + 1) * hh * ww;
    window_accumulate_kernel<<<(int)min((int64_t)UEM_SMS * 8, (total + 255) / 256), 256, 0, (cudaStream_t)stream>>>(full, count, tile, b, c, H,
                                                                                                                  W, th, tw, y1, x1, hh, ww);
    UEM_CHECK_LAUNCH();
    return 0;
}

extern "C" int uem_window_average_f32(float* full, const float* count, int b, int c, int64_t hw, void* stream) {
    UEM_REQUIRE(full && count && b > 0 && c > 0 && hw > 0, "uem_window_average_f32: bad arguments");
    const int64_t total = (int64_t)b * c * hw;
    window_average_kernel<<<(int)min((int64_t)UEM_SMS * 8, (total + 255) / 256), 256, 0, (cudaStream_t)stream>>>(full, count, b, c, hw);
    UEM_CHECK_LAUNCH();
    return 0;
}

extern "C" int uem_views_mean_f32(const float* views, int n, int64_t numel, float* out, void* stream) {
    UEM_REQUIRE(views && out && n > 0 && numel > 0, "uem_views_mean_f32: bad arguments");
    views_mean_kernel<<<(int)min((int64_t)UEM_SMS * 8, (numel + 255) / 256), 256, 0, (cudaStream_t)stream>>>(views, n, numel, out);
    UEM_CHECK_LAUNCH();
    return 0;
}
