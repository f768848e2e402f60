// a6: Aligner.label_refine as ONE full-resolution kernel, plus the fused refine->select chain.
// Reference: uemda/gast/alignment.py:194-293 (~60 eager kernels, ~58 full-res map passes):
//   prototype view  :215-223  1/pearson (low-res) -> bilinear up -> softmax(T=1) -> /max
//   prediction view :225-236  logits (low-res) -> bilinear up -> softmax(/temp) [mean of 2 heads] -> /max
//   superpixel view :238-258  region max of soft (torch_scatter 'max') -> gather -> softmax(/temp) -> /max,
//                             applied multiplicatively outside the batch-global "ignored" id (:241-243,:255)
//   final           :291-292  soft' = w*soft / (sum_c w*soft + 1e-7)
// Compulsory traffic per pixel: read soft (4c) + sup (8), write refined (4c).  Everything else (three
// low-res maps, the region table) is L2/L1 resident.  One CTA owns 4 output rows of one image: per
// row it interpolates the low-res maps vertically into shared memory once, then each thread handles
// 4 consecutive pixels with the class vectors in registers.  The per-(image,class) maxima that
// pseudo_selection needs (pseudo_generation.py:76) fall out of the same pass as per-CTA partials.
#include "uem_common.cuh"

int uem_region_table_f32(const float* src, int64_t sb, int64_t sn, int64_t sc, const int64_t* index, int b, int64_t N,
                         int c, int64_t R, int op, const int64_t* hot_ptr, int64_t hot_val, int skip_hot, unsigned* table,
                         unsigned* cnt, int* status, cudaStream_t st);

namespace {

constexpr int kRows = 4;  // output rows per CTA

struct RefineParams {
    int views;
    const float* maps[3];   // simi, pred1, pred2 (low-res (b,C,h,w)); unused entries null
    int nmaps, n_pred;      // staged maps; number of prediction heads (0..2)
    int h, w, H, W;
    float sy, sx;
    float temp, inv_temp;
    int temp_pow2;
    const int64_t* sup;
    const void* table;      // (b,R,C): fp32 (decoded) or ordered-u32 (encoded)
    int table_encoded;
    int64_t R;
    const int64_t* ignored_id;
    const float* soft;
    float* out;
    float* partial;         // (b, gridDim.x, C+1): per-class max + overall min of `out`
};

template <int C>
__device__ __forceinline__ void hlerp(const float* rowbase, int w, const Lerp& lx, float (&z)[C]) {
#pragma unroll
    for (int ci = 0; ci < C; ++ci) {
        const float* r = rowbase + ci * w;
        z[ci] = lx.l0 * r[lx.i0] + lx.l1 * r[lx.i1];
    }
}

template <int C, int VEC>
__global__ void __launch_bounds__(256) refine_kernel(const RefineParams p) {
    extern __shared__ float row[];  // [nmaps][C][w] vertically interpolated low-res rows
    const int bi = blockIdx.y;
    const int64_t HW = (int64_t)p.H * p.W;
    const bool vP = p.views & UEM_VIEW_PROTO, vL = p.views & UEM_VIEW_PRED, vS = p.views & UEM_VIEW_SUP;
    const int64_t ignored_id = vS ? *p.ignored_id : -1;
    const float* softb = p.soft + (int64_t)bi * C * HW;
    float* outb = p.out + (int64_t)bi * C * HW;

    float cmax[C];
#pragma unroll
    for (int ci = 0; ci < C; ++ci) cmax[ci] = -INFINITY;
    float cmin = INFINITY;
    bool bad = false;

    for (int r = 0; r < kRows; ++r) {
        const int y = blockIdx.x * kRows + r;
        if (y >= p.H) break;
        if (p.nmaps) {
            const Lerp ly = make_lerp(y, p.h, p.sy);
            __syncthreads();
            const int total = p.nmaps * C * p.w;
            for (int i = threadIdx.x; i < total; i += blockDim.x) {
                const int x = i % p.w, mc = i / p.w;
                const int m = mc / C, ci = mc - m * C;
                const float* plane = p.maps[m] + ((int64_t)bi * C + ci) * p.h * p.w;
                row[i] = ly.l0 * __ldg(plane + (int64_t)ly.i0 * p.w + x) + ly.l1 * __ldg(plane + (int64_t)ly.i1 * p.w + x);
            }
            __syncthreads();
        }
        const int groups = p.W / VEC;
        for (int g = threadIdx.x; g < groups; g += blockDim.x) {
            const int x0 = g * VEC;
            const int64_t px = (int64_t)y * p.W + x0;
            float sv[C][VEC];
#pragma unroll
            for (int ci = 0; ci < C; ++ci) {
                PixVec<VEC> t;
                t.load(softb + (int64_t)ci * HW + px);
#pragma unroll
                for (int i = 0; i < VEC; ++i) sv[ci][i] = t.v[i];
            }
            int64_t id[VEC];
            if (vS) load_ids<VEC>(p.sup + (int64_t)bi * HW + px, id);
#pragma unroll
            for (int i = 0; i < VEC; ++i) {
                const Lerp lx = make_lerp(x0 + i, p.w, p.sx);
                float wgt[C];
                bool have = false;
                int mi = 0;
                if (vP) {  // prototype view: softmax(T=1) of the up-sampled 1/distance, peak-normalised
                    hlerp<C>(row, p.w, lx, wgt);
                    softmax_regs<C>(wgt);
                    peak_norm_regs<C>(wgt);
                    have = true;
                    mi = 1;
                }
                if (vL) {  // prediction view
                    float acc[C];
#pragma unroll
                    for (int ci = 0; ci < C; ++ci) acc[ci] = 0.f;
                    for (int hd = 0; hd < p.n_pred; ++hd) {
                        float z[C];
                        hlerp<C>(row + (mi + hd) * C * p.w, p.w, lx, z);
#pragma unroll
                        for (int ci = 0; ci < C; ++ci) z[ci] = p.temp_pow2 ? z[ci] * p.inv_temp : __fdiv_rn(z[ci], p.temp);
                        softmax_regs<C>(z);
#pragma unroll
                        for (int ci = 0; ci < C; ++ci) acc[ci] += z[ci];
                    }
                    if (p.n_pred == 2) {
#pragma unroll
                        for (int ci = 0; ci < C; ++ci) acc[ci] *= 0.5f;
                    }
                    peak_norm_regs<C>(acc);
#pragma unroll
                    for (int ci = 0; ci < C; ++ci) wgt[ci] = have ? wgt[ci] + acc[ci] : acc[ci];
                    have = true;
                }
                if (vS) {  // superpixel view, multiplicative outside the ignored id
                    const int64_t rid = id[i];
                    if (rid != ignored_id && rid >= 0 && rid < p.R) {
                        float z[C];
                        if (p.table_encoded) {
                            const unsigned* t = (const unsigned*)p.table + ((int64_t)bi * p.R + rid) * C;
#pragma unroll
                            for (int ci = 0; ci < C; ++ci) { unsigned e = __ldg(t + ci); z[ci] = e ? ordered_to_f32(e) : 0.f; }
                        } else {
                            const float* t = (const float*)p.table + ((int64_t)bi * p.R + rid) * C;
#pragma unroll
                            for (int ci = 0; ci < C; ++ci) z[ci] = __ldg(t + ci);
                        }
#pragma unroll
                        for (int ci = 0; ci < C; ++ci) z[ci] = p.temp_pow2 ? z[ci] * p.inv_temp : __fdiv_rn(z[ci], p.temp);
                        softmax_regs<C>(z);
                        peak_norm_regs<C>(z);
#pragma unroll
                        for (int ci = 0; ci < C; ++ci) wgt[ci] = have ? wgt[ci] * z[ci] : z[ci];
                    } else if (!have) {
#pragma unroll
                        for (int ci = 0; ci < C; ++ci) wgt[ci] = 1.0f;
                    }
                }
                // soft' = w*soft / (sum + 1e-7)   (alignment.py:291-292, :324-325)
                float s = 0.f;
#pragma unroll
                for (int ci = 0; ci < C; ++ci) {
                    sv[ci][i] = wgt[ci] * sv[ci][i];
                    s += sv[ci][i];
                }
                const float inv = 1.0f / (s + 1e-7f);
#pragma unroll
                for (int ci = 0; ci < C; ++ci) {
                    const float o = sv[ci][i] * inv;
                    sv[ci][i] = o;
                    cmax[ci] = fmaxf(cmax[ci], o);
                    cmin = fminf(cmin, o);
                    bad |= (o != o);
                }
            }
#pragma unroll
            for (int ci = 0; ci < C; ++ci) {
                PixVec<VEC> t;
#pragma unroll
                for (int i = 0; i < VEC; ++i) t.v[i] = sv[ci][i];
                t.store(outb + (int64_t)ci * HW + px);
            }
        }
    }
    if (p.partial) {
        __shared__ float red[8][C + 2];
        const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
#pragma unroll
        for (int ci = 0; ci < C; ++ci) {
            float v = warp_max(cmax[ci]);
            if (lane == 0) red[warp][ci] = v;
        }
        float mn = warp_min(cmin);
        int anybad = __any_sync(0xffffffffu, bad);
        if (lane == 0) { red[warp][C] = mn; red[warp][C + 1] = anybad ? 1.f : 0.f; }
        __syncthreads();
        if (threadIdx.x <= C) {
            const int nw = (blockDim.x + 31) >> 5;
            float v = red[0][threadIdx.x];
            float nanf_ = 0.f;
            for (int i = 0; i < nw; ++i) {
                v = (threadIdx.x < C) ? fmaxf(v, red[i][threadIdx.x]) : fminf(v, red[i][threadIdx.x]);
                nanf_ += red[i][C + 1];
            }
            if (nanf_ != 0.f) v = NAN;
            p.partial[((int64_t)bi * gridDim.x + blockIdx.x) * (C + 1) + threadIdx.x] = v;
        }
    }
}

// a13: weight of each pixel's own hard-label class under the prototype view (alignment.py:295-309)
template <int C>
__global__ void __launch_bounds__(256) proto_weight_4pixel_kernel(const float* __restrict__ simi, int h, int w, int H, int W, float sy,
                                                                  float sx, const int64_t* __restrict__ hard, int64_t ignore_label,
                                                                  float eps, float* __restrict__ out) {
    extern __shared__ float row[];
    const int y = blockIdx.x, bi = blockIdx.y;
    const Lerp ly = make_lerp(y, h, sy);
    for (int i = threadIdx.x; i < C * w; i += blockDim.x) {
        const int x = i % w, ci = i / w;
        const float* plane = simi + ((int64_t)bi * C + ci) * h * w;
        row[i] = ly.l0 * __ldg(plane + (int64_t)ly.i0 * w + x) + ly.l1 * __ldg(plane + (int64_t)ly.i1 * w + x);
    }
    __syncthreads();
    for (int x = threadIdx.x; x < W; x += blockDim.x) {
        const Lerp lx = make_lerp(x, w, sx);
        float z[C];
        hlerp<C>(row, w, lx, z);
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

static int launch_refine(int views, const float* simi, const float* pred1, const float* pred2, int h, int w, const int64_t* sup,
                         const void* table, int table_encoded, int64_t R, const int64_t* ignored_id, const float* soft, int b,
                         int c, int H, int W, float temp, float* out, float* partial, cudaStream_t st) {
    UEM_REQUIRE(soft && out && b > 0 && H > 0 && W > 0, "uem_label_refine_f32: bad arguments");
    UEM_REQUIRE(views > 0 && views < 8, "uem_label_refine_f32: views must be a non-empty mask of UEM_VIEW_*");
    UEM_REQUIRE(temp > 0.f, "uem_label_refine_f32: temp must be > 0");  // alignment.py:313
    RefineParams p{};
    p.views = views;
    int nm = 0;
    if (views & UEM_VIEW_PROTO) { UEM_REQUIRE(simi, "uem_label_refine_f32: prototype view needs simi"); p.maps[nm++] = simi; }
    if (views & UEM_VIEW_PRED) {
        UEM_REQUIRE(pred1, "uem_label_refine_f32: prediction view needs pred1");
        p.maps[nm++] = pred1;
        p.n_pred = 1;
        if (pred2) { p.maps[nm++] = pred2; p.n_pred = 2; }
    }
    if (views & UEM_VIEW_SUP) UEM_REQUIRE(sup && table && ignored_id && R > 0, "uem_label_refine_f32: superpixel view needs sup, region table, ignored id");
    if (nm) UEM_REQUIRE(h > 0 && w > 0, "uem_label_refine_f32: bad low-res size");
    p.nmaps = nm;
    p.h = h; p.w = w; p.H = H; p.W = W;
    p.sy = uem_align_corners_scale(h, H);
    p.sx = uem_align_corners_scale(w, W);
    p.temp = temp;
    p.inv_temp = 1.0f / temp;
    int ex;
    p.temp_pow2 = (frexpf(temp, &ex) == 0.5f);
    p.sup = sup; p.table = table; p.table_encoded = table_encoded; p.R = R; p.ignored_id = ignored_id;
    p.soft = soft; p.out = out; p.partial = partial;
    const bool vec = (W % 4 == 0) && uem_aligned16(soft) && uem_aligned16(out) && (!sup || uem_aligned16(sup));
    void *ev0, *ev1;
    uem_take_profile_events(&ev0, &ev1);
    if (ev0) UEM_CUDA(cudaEventRecord((cudaEvent_t)ev0, st));
    UEM_DISPATCH_C(c, {
        size_t smem = (size_t)nm * C * w * sizeof(float);
        UEM_REQUIRE(smem <= 227 * 1024, "uem_label_refine_f32: low-res width %d too large", w);
        dim3 grid(uem_div_up(H, kRows), b);
        if (vec) {
            int threads = min(256, max(32, ((W / 4 + 31) / 32) * 32));
            if (smem > 48 * 1024) UEM_CUDA(cudaFuncSetAttribute(refine_kernel<C, 4>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
            refine_kernel<C, 4><<<grid, threads, smem, st>>>(p);
        } else {
            int threads = min(256, max(32, ((W + 31) / 32) * 32));
            if (smem > 48 * 1024) UEM_CUDA(cudaFuncSetAttribute(refine_kernel<C, 1>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
            refine_kernel<C, 1><<<grid, threads, smem, st>>>(p);
        }
    });
    if (ev1) UEM_CUDA(cudaEventRecord((cudaEvent_t)ev1, st));
    UEM_CHECK_LAUNCH();
    return 0;
}

}  // namespace

extern "C" int uem_label_refine_partials(int H) { return uem_div_up(H, kRows); }

extern "C" int uem_label_refine_f32(int views, const float* simi, const float* pred1, const float* pred2, int h, int w,
                                    const int64_t* sup, const float* region_max, int64_t R, const int64_t* ignored_id,
                                    const float* soft, int b, int c, int H, int W, float temp, float* out,
                                    float* class_max_partial, void* stream) {
    return launch_refine(views, simi, pred1, pred2, h, w, sup, region_max, 0, R, ignored_id, soft, b, c, H, W, temp, out,
                         class_max_partial, (cudaStream_t)stream);
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
// Fused chain: label_refine (all requested views) -> pseudo_selection, one C call, no host sync.
// tools/train_ssl_uem.py:209-214.  ws layout (all 16-byte aligned):
//   [0]      status   int32[4]   bit1: label out of range, bit2: superpixel id outside [0,R)
//   [16]     minmax   int64[2]   min / max superpixel id of the batch (alignment.py:241)
//   [32]     simi     f32[b*c*h*w]
//            pearson ws (uem_pearson_ws_bytes)
//            region   u32[b*R*c]  ordered-encoded region maxima
//            partial  f32[b*nparts*(c+1)]
// ------------------------------------------------------------------------------------------------
static inline int64_t align16(int64_t x) { return (x + 15) & ~(int64_t)15; }

extern "C" int64_t uem_mine_ws_bytes(int b, int c, int H, int W, int h, int w, int k, int64_t R) {
    (void)W;
    int64_t n = 32;
    n += align16((int64_t)b * c * h * w * 4);
    n += align16(uem_pearson_ws_bytes(c, k));
    n += align16((int64_t)b * R * c * 4);
    n += align16((int64_t)b * uem_label_refine_partials(H) * (c + 1) * 4);
    return n;
}

extern "C" int uem_mine_refine_select_f32(int views, const float* feat, int k, const float* protos, const float* pred1,
                                          const float* pred2, int h, int w, const int64_t* sup, int64_t R,
                                          const int64_t* ignored_id, const float* soft, int b, int c, int H, int W, float temp,
                                          float eps, float cutoff_top, float cutoff_low, int64_t ignore_label, float* refined,
                                          int64_t* hard, void* ws, void* stream) {
    UEM_REQUIRE(ws && soft && refined, "uem_mine_refine_select_f32: bad arguments");
    cudaStream_t st = (cudaStream_t)stream;
    char* base = (char*)ws;
    int* status = (int*)base;
    int64_t* minmax = (int64_t*)(base + 16);
    float* simi = (float*)(base + 32);
    char* q = base + 32 + align16((int64_t)b * c * h * w * 4);
    void* pws = q;
    q += align16(uem_pearson_ws_bytes(c, k));
    unsigned* table = (unsigned*)q;
    q += align16((int64_t)b * R * c * 4);
    float* partial = (float*)q;
    const int64_t HW = (int64_t)H * W;
    int rc;
    if (views & UEM_VIEW_SUP) {
        UEM_REQUIRE(sup && R > 0, "uem_mine_refine_select_f32: superpixel view needs sup and a region capacity R");
        if (!ignored_id) {
            if ((rc = uem_i64_minmax(sup, (int64_t)b * HW, minmax, stream))) return rc;
            ignored_id = minmax + 1;
        }
        UEM_CUDA(cudaMemsetAsync(table, 0, (size_t)b * R * c * 4, st));
        // region maxima of soft, NCHW viewed as (b,N,c): strides {c*N, 1, N}; the ignored id is never gathered
        // (alignment.py:255), so its pixels are skipped
        if ((rc = uem_region_table_f32(soft, (int64_t)c * HW, 1, HW, sup, b, HW, c, R, UEM_REDUCE_MAX, ignored_id, -1, 1, table,
                                       nullptr, status, st)))
            return rc;
    }
    if (views & UEM_VIEW_PROTO) {
        UEM_REQUIRE(feat && protos, "uem_mine_refine_select_f32: prototype view needs feat and prototypes");
        if ((rc = uem_pearson_dist_nchw_f32(feat, b, k, (int64_t)h * w, protos, c, eps, 1, simi, pws, stream))) return rc;
    }
    if ((rc = launch_refine(views, simi, pred1, pred2, h, w, sup, table, 1, R, ignored_id, soft, b, c, H, W, temp, refined,
                            partial, st)))
        return rc;
    if (hard) {
        if ((rc = uem_pseudo_select_partials_f32(refined, partial, uem_label_refine_partials(H), b, c, HW, cutoff_top, cutoff_low,
                                                 ignore_label, hard, stream)))
            return rc;
    }
    return 0;
}
