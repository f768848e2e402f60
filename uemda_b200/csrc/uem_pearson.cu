// a9: Pearson distance between every feature vector and every class prototype.
// Reference: Aligner._pearson_dist, uemda/gast/alignment.py:424-451:
//   d[i,j] = ( -cov(f_i,p_j) / ((k-1+eps) * (std(f_i)*std(p_j) + eps)) + 1 ) * 0.5, unbiased std.
// The reference transposes NCHW features to (n,k) (alignment.py:213) and materialises an (n,c,k)
// product (:438-440, 403 MB at config 2).  Here the NCHW map is read exactly once, in place:
// threads run along the pixel dimension (128-bit loads, 128 B per class-of-8-lanes), the k dimension is
// split over 32 slices per CTA, and each thread keeps 2+m running sums per pixel:
//   S1 = sum g, S2 = sum g^2, Cj = sum g * pc_j      with g = f - pivot (pivot = f at channel 0)
// where pc_j is the centred prototype (sum_k pc_j ~ 0, so centring f is unnecessary for the covariance;
// the residual mean(g)*sum(pc_j) is subtracted anyway).  HBM-bound on feat: 4k B per feature pixel.
#include "uem_common.cuh"

namespace {

constexpr int kPearsonThreads = 256;
constexpr int kPxLanes = 4;    // lanes along pixels; each handles VEC pixels (64 B of a channel row per warp quarter)
constexpr int kSlices = kPearsonThreads / kPxLanes;  // k-slices per CTA (64): 512 CTAs of 16 pixels at config 2
constexpr int kDepth = 8;      // cp.async ring depth per thread
constexpr int kPcStride = 8;   // transposed centred prototypes: pcT[kk][8] -> two LDG.128 fetch all classes of a channel

// centre the prototypes once: pc (m,k), stats[j] = {std_j (unbiased), sum_k pc_j}
__global__ void __launch_bounds__(256) proto_center_kernel(const float* __restrict__ protos, int k, int transposed,
                                                           float* __restrict__ pc, float* __restrict__ stats) {
    const int j = blockIdx.x;
    const float* p = protos + (int64_t)j * k;
    __shared__ float red[8];
    __shared__ float bc;
    float s = 0.f;
    for (int i = threadIdx.x; i < k; i += 256) s += p[i];
    s = warp_sum(s);
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = s;
    __syncthreads();
    if (threadIdx.x == 0) { float t = 0.f; for (int i = 0; i < 8; ++i) t += red[i]; bc = t / (float)k; }
    __syncthreads();
    const float mean = bc;
    float s2 = 0.f, s1 = 0.f;
    for (int i = threadIdx.x; i < k; i += 256) {
        float d = p[i] - mean;
        if (transposed) pc[(int64_t)i * kPcStride + j] = d;
        else pc[(int64_t)j * k + i] = d;
        s2 += d * d;
        s1 += d;
    }
    s2 = warp_sum(s2);
    s1 = warp_sum(s1);
    __syncthreads();
    __shared__ float red1[8];
    if ((threadIdx.x & 31) == 0) { red[threadIdx.x >> 5] = s2; red1[threadIdx.x >> 5] = s1; }
    __syncthreads();
    if (threadIdx.x == 0) {
        float t2 = 0.f, t1 = 0.f;
        for (int i = 0; i < 8; ++i) { t2 += red[i]; t1 += red1[i]; }
        stats[2 * j] = sqrtf(t2 / (float)(k - 1));
        stats[2 * j + 1] = t1;
    }
}

__device__ __forceinline__ float pearson_finish(float S1, float S2, float Cj, float std_p, float sum_pc, int k, float eps,
                                                int reciprocal) {
    const float kf = (float)k;
    const float mean_g = S1 / kf;
    float var = (S2 - S1 * mean_g) / (float)(k - 1);
    var = fmaxf(var, 0.f);
    const float std_f = sqrtf(var);
    const float cov = (Cj - mean_g * sum_pc) / ((float)(k - 1) + eps);
    const float den = std_f * std_p + eps;
    const float d = (-1.0f * cov / den + 1.0f) * 0.5f;
    return reciprocal ? 1.0f / d : d;
}

// feat (b,k,hw) planar -> out (b,M,hw) planar
template <int M, int VEC>
__global__ void __launch_bounds__(kPearsonThreads, 4) pearson_nchw_kernel(const float* __restrict__ feat, int k, int64_t hw,
                                                                       const float* __restrict__ pc, const float* __restrict__ stats,
                                                                       float eps, int reciprocal, float* __restrict__ out) {
    constexpr int NA = 2 + M;
    constexpr int TILE = kPxLanes * VEC;
    const int bi = blockIdx.y;
    const int pl = threadIdx.x % kPxLanes, ks = threadIdx.x / kPxLanes;
    const int64_t px0 = (int64_t)blockIdx.x * TILE + pl * VEC;
    const bool active = px0 < hw;  // hw % VEC == 0 on the vector path
    const float* f = feat + (int64_t)bi * k * hw + px0;

    float acc[NA][VEC];
#pragma unroll
    for (int a = 0; a < NA; ++a)
#pragma unroll
        for (int i = 0; i < VEC; ++i) acc[a][i] = 0.f;

    if (active) {
        PixVec<VEC> piv;
        piv.load(f);  // channel 0 of these pixels: the shift that keeps the one-pass variance stable
        auto consume = [&](const PixVec<VEC>& v, int kk) {
            float pcv[kPcStride];
            const float4 a = __ldg(reinterpret_cast<const float4*>(pc + (int64_t)kk * kPcStride));
            pcv[0] = a.x; pcv[1] = a.y; pcv[2] = a.z; pcv[3] = a.w;
            if (M > 4) {
                const float4 c4 = __ldg(reinterpret_cast<const float4*>(pc + (int64_t)kk * kPcStride) + 1);
                pcv[4] = c4.x; pcv[5] = c4.y; pcv[6] = c4.z; pcv[7] = c4.w;
            }
#pragma unroll
            for (int i = 0; i < VEC; ++i) {
                const float g = v.v[i] - piv.v[i];
                acc[0][i] += g;
                acc[1][i] = fmaf(g, g, acc[1][i]);
#pragma unroll
                for (int j = 0; j < M; ++j) acc[2 + j][i] = fmaf(g, pcv[j], acc[2 + j][i]);
            }
        };
        if constexpr (VEC == 4) {
            // register-free software pipeline: every thread keeps kDepth 16-byte cp.async copies in flight in its own
            // shared-memory ring slots (128 B per thread, ~16 MB chip-wide: enough to cover HBM latency)
            __shared__ float4 ring[kDepth][kPearsonThreads];
            const int n_it = (k - ks + kSlices - 1) / kSlices;
#pragma unroll
            for (int d = 0; d < kDepth; ++d) {
                if (d < n_it) cp_async_16(&ring[d][threadIdx.x], f + (int64_t)(ks + d * kSlices) * hw);
                cp_async_commit_group();
            }
            for (int it = 0; it < n_it; ++it) {
                cp_async_wait_group<kDepth - 1>();
                const float4 q = ring[it % kDepth][threadIdx.x];
                if (it + kDepth < n_it) cp_async_16(&ring[it % kDepth][threadIdx.x], f + (int64_t)(ks + (it + kDepth) * kSlices) * hw);
                cp_async_commit_group();
                PixVec<VEC> v;
                v.v[0] = q.x; v.v[1] = q.y; v.v[2] = q.z; v.v[3] = q.w;
                consume(v, ks + it * kSlices);
            }
            cp_async_wait_group<0>();
        } else {
#pragma unroll 4
            for (int kk = ks; kk < k; kk += kSlices) {
                PixVec<VEC> v;
                v.load(f + (int64_t)kk * hw);
                consume(v, kk);
            }
        }
    }
    // fold the 8 slices that share a warp (lanes l, l+4, ..., l+28 hold the same pixels)
#pragma unroll
    for (int a = 0; a < NA; ++a)
#pragma unroll
        for (int i = 0; i < VEC; ++i) {
            float x = acc[a][i];
            x += __shfl_xor_sync(0xffffffffu, x, 4);
            x += __shfl_xor_sync(0xffffffffu, x, 8);
            x += __shfl_xor_sync(0xffffffffu, x, 16);
            acc[a][i] = x;
        }
    __shared__ float red[kPearsonThreads / 32][TILE][NA + 1];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    if (lane < kPxLanes) {
#pragma unroll
        for (int i = 0; i < VEC; ++i)
#pragma unroll
            for (int a = 0; a < NA; ++a) red[warp][lane * VEC + i][a] = acc[a][i];
    }
    __syncthreads();
    if (threadIdx.x < TILE) {
        const int64_t px = (int64_t)blockIdx.x * TILE + threadIdx.x;
        if (px < hw) {
            float s[NA];
#pragma unroll
            for (int a = 0; a < NA; ++a) {
                float t = 0.f;
#pragma unroll
                for (int wv = 0; wv < kPearsonThreads / 32; ++wv) t += red[wv][threadIdx.x][a];
                s[a] = t;
            }
#pragma unroll
            for (int j = 0; j < M; ++j)
                out[((int64_t)bi * M + j) * hw + px] =
                    pearson_finish(s[0], s[1], s[2 + j], stats[2 * j], stats[2 * j + 1], k, eps, reciprocal);
        }
    }
}

// generic row-major (n,k) x (m,k): one warp per feat1 row, classes in chunks of 8
__global__ void __launch_bounds__(256) pearson_rows_kernel(const float* __restrict__ f1, int64_t n, int k, const float* __restrict__ pc,
                                                           const float* __restrict__ stats, int m, float eps, float* __restrict__ out) {
    const int lane = threadIdx.x & 31;
    const int64_t row = (int64_t)blockIdx.x * 8 + (threadIdx.x >> 5);
    if (row >= n) return;
    const float* f = f1 + row * k;
    const float piv = f[0];
    for (int j0 = 0; j0 < m; j0 += 8) {
        const int mj = min(8, m - j0);
        float S1 = 0.f, S2 = 0.f, Cj[8];
#pragma unroll
        for (int j = 0; j < 8; ++j) Cj[j] = 0.f;
        for (int kk = lane; kk < k; kk += 32) {
            const float g = f[kk] - piv;
            S1 += g;
            S2 = fmaf(g, g, S2);
#pragma unroll
            for (int j = 0; j < 8; ++j)
                if (j < mj) Cj[j] = fmaf(g, __ldg(pc + (int64_t)(j0 + j) * k + kk), Cj[j]);
        }
        S1 = warp_sum(S1);
        S2 = warp_sum(S2);
#pragma unroll
        for (int j = 0; j < 8; ++j) Cj[j] = warp_sum(Cj[j]);
        if (lane == 0) {
#pragma unroll
            for (int j = 0; j < 8; ++j)
                if (j < mj)
                    out[row * m + j0 + j] = pearson_finish(S1, S2, Cj[j], stats[2 * (j0 + j)], stats[2 * (j0 + j) + 1], k, eps, 0);
        }
    }
}

}  // namespace

// ws layout: [pc max(m,8)*k f32 (row-major (m,k), or transposed (k,8) for the NCHW kernel)][stats 2*m f32]
extern "C" int64_t uem_pearson_ws_bytes(int m, int k) {
    const int64_t rows = m > kPcStride ? m : kPcStride;
    return (rows * k + 2 * (int64_t)m + 4) * sizeof(float);
}

extern "C" int uem_pearson_dist_nchw_f32(const float* feat, int b, int k, int64_t hw, const float* protos, int m, float eps,
                                         int reciprocal, float* out, void* ws, void* stream) {
    UEM_REQUIRE(feat && protos && out && ws && b > 0 && k > 0 && hw > 0, "uem_pearson_dist_nchw_f32: bad arguments");
    cudaStream_t st = (cudaStream_t)stream;
    float* pc = (float*)ws;
    float* stats = pc + (int64_t)kPcStride * k;
    const bool vec = (hw % 4 == 0) && uem_aligned16(feat);
    UEM_CUDA(cudaMemsetAsync(pc, 0, (size_t)kPcStride * k * sizeof(float), st));  // unused class slots stay 0
    UEM_DISPATCH_C(m, {
        proto_center_kernel<<<C, 256, 0, st>>>(protos, k, 1, pc, stats);
        if (vec) {
            dim3 grid(uem_div_up(hw, kPxLanes * 4), b);
            pearson_nchw_kernel<C, 4><<<grid, kPearsonThreads, 0, st>>>(feat, k, hw, pc, stats, eps, reciprocal, out);
        } else {
            dim3 grid(uem_div_up(hw, kPxLanes), b);
            pearson_nchw_kernel<C, 1><<<grid, kPearsonThreads, 0, st>>>(feat, k, hw, pc, stats, eps, reciprocal, out);
        }
    });
    UEM_CHECK_LAUNCH_N(2);
    return 0;
}

extern "C" int uem_pearson_dist_rows_f32(const float* feat1, int64_t n, int k, const float* feat2, int m, float eps, float* out,
                                         void* ws, void* stream) {
    UEM_REQUIRE(feat1 && feat2 && out && ws && n > 0 && k > 0 && m > 0, "uem_pearson_dist_rows_f32: bad arguments");
    cudaStream_t st = (cudaStream_t)stream;
    float* pc = (float*)ws;
    float* stats = pc + (int64_t)(m > kPcStride ? m : kPcStride) * k;
    proto_center_kernel<<<m, 256, 0, st>>>(feat2, k, 0, pc, stats);
    pearson_rows_kernel<<<uem_div_up(n, 8), 256, 0, st>>>(feat1, n, k, pc, stats, m, eps, out);
    UEM_CHECK_LAUNCH_N(2);
    return 0;
}
