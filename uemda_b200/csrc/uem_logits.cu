// a1-a4: one fused pass over the target logits -> soft labels, confidence, entropy, argmax;
// a2/a3: entropy of a probability map + UVEM uncertainty weight.
// Reference: tools/train_align_uem.py:158-160 (bilinear up, softmax per head, mean),
// uemda/models/Encoder.py:153-155, uemda/gast/alignment.py:311-314 (temperature),
// uemda/gast/balance.py:368-373 (entropy, gate), :396-423 (get_weight), :331-342 (UPS),
// uemda/gast/pseudo_generation.py:47,148 (max / argmax).
//
// HBM-bound: the low-resolution logits are L2 resident (b*c*h*w*4 B, a few hundred KB); compulsory
// traffic is the full-resolution outputs.  One CTA owns one output row: it first interpolates the
// low-res rows vertically into shared memory (c values per low-res column per head), then every
// thread interpolates horizontally for its 4 consecutive pixels, keeps the class vector in
// registers, and writes 128-bit rows per class plane.
#include "uem_common.cuh"

namespace {

// vertical lerp of `nmaps` low-res maps into smem rows: row[(map*C+ci)*w + x]
template <int C>
__device__ __forceinline__ void stage_rows(float* row, const float* const* maps, int nmaps, int bi, int h, int w,
                                           const Lerp& ly) {
    const int total = nmaps * C * w;
    for (int i = threadIdx.x; i < total; i += blockDim.x) {
        int x = i % w, mc = i / w;
        int m = mc / C, ci = mc % C;
        const float* plane = maps[m] + ((int64_t)bi * C + ci) * h * w;
        float a = __ldg(plane + (int64_t)ly.i0 * w + x);
        float b2 = __ldg(plane + (int64_t)ly.i1 * w + x);
        row[i] = ly.l0 * a + ly.l1 * b2;
    }
}

template <int C, int VEC>
__global__ void __launch_bounds__(256) logits_pass_kernel(const float* __restrict__ x1, const float* __restrict__ x2,
                                                          int h, int w, int H, int W, float sy, float sx, float inv_temp,
                                                          int divide, float temp, float* __restrict__ soft,
                                                          float* __restrict__ conf, float* __restrict__ entropy,
                                                          int64_t* __restrict__ argmax) {
    extern __shared__ float row[];
    const int y = blockIdx.x, bi = blockIdx.y;
    const int nmaps = x2 ? 2 : 1;
    const float* maps[2] = {x1, x2};
    const Lerp ly = make_lerp(y, h, sy);
    stage_rows<C>(row, maps, nmaps, bi, h, w, ly);
    __syncthreads();
    const int64_t HW = (int64_t)H * W;
    const int groups = W / VEC;
    for (int g = threadIdx.x; g < groups; g += blockDim.x) {
        const int x0 = g * VEC;
        float p[C][VEC];
        float cf[VEC], en[VEC];
        int64_t am[VEC];
#pragma unroll
        for (int i = 0; i < VEC; ++i) {
            const Lerp lx = make_lerp(x0 + i, w, sx);
            float acc[C];
#pragma unroll
            for (int ci = 0; ci < C; ++ci) acc[ci] = 0.f;
            for (int m = 0; m < nmaps; ++m) {
                float z[C];
#pragma unroll
                for (int ci = 0; ci < C; ++ci) {
                    const float* r = row + (m * C + ci) * w;
                    float v = lx.l0 * r[lx.i0] + lx.l1 * r[lx.i1];
                    z[ci] = divide ? v / temp : v;  // alignment.py:314: division first
                }
                softmax_regs<C>(z);
#pragma unroll
                for (int ci = 0; ci < C; ++ci) acc[ci] += z[ci];
            }
            float best = -INFINITY;
            int arg = 0;
            float pp[C];
#pragma unroll
            for (int ci = 0; ci < C; ++ci) {
                float v = (nmaps == 2) ? acc[ci] * 0.5f : acc[ci];
                p[ci][i] = v;
                pp[ci] = v;
                bool gt = v > best;
                best = gt ? v : best;
                arg = gt ? ci : arg;
            }
            cf[i] = best; en[i] = entropy_px<C>(pp); am[i] = arg;  // balance.py:372; p==0 -> NaN like the reference
        }
        const int64_t px = (int64_t)y * W + x0;
        if (soft) {
#pragma unroll
            for (int ci = 0; ci < C; ++ci) {
                PixVec<VEC> o;
#pragma unroll
                for (int i = 0; i < VEC; ++i) o.v[i] = p[ci][i];
                o.store(soft + ((int64_t)bi * C + ci) * HW + px);
            }
        }
        if (conf) { PixVec<VEC> o; for (int i = 0; i < VEC; ++i) o.v[i] = cf[i]; o.store(conf + (int64_t)bi * HW + px); }
        if (entropy) { PixVec<VEC> o; for (int i = 0; i < VEC; ++i) o.v[i] = en[i]; o.store(entropy + (int64_t)bi * HW + px); }
        if (argmax) store_ids<VEC>(argmax + (int64_t)bi * HW + px, am);
    }
    (void)inv_temp;
}

// entropy (+ optional UVEM weight / gate / valid count) of a probability map (b,c,HW)
template <int C, int VEC, bool TERMS>
__global__ void __launch_bounds__(256) entropy_weight_kernel(const float* __restrict__ soft, const int64_t* __restrict__ target,
                                                             int64_t hw, float m, float t, float inv_gamma, float cl, float cr,
                                                             int use_weight, int64_t ignore_label, float* __restrict__ entropy,
                                                             float* __restrict__ weight, uint8_t* __restrict__ gate,
                                                             unsigned long long* __restrict__ valid_cnt) {
    const int bi = blockIdx.y;
    const int64_t g = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    int valid = 0;
    if (g * VEC < hw) {
        const int64_t px = g * VEC;
        const float* base = soft + (int64_t)bi * C * hw + px;
        float pv[VEC][C];
#pragma unroll
        for (int ci = 0; ci < C; ++ci) {
            PixVec<VEC> v;
            v.load(base + (int64_t)ci * hw);
#pragma unroll
            for (int i = 0; i < VEC; ++i) pv[i][ci] = v.v[i];
        }
        float u[VEC];
#pragma unroll
        for (int i = 0; i < VEC; ++i) u[i] = entropy_px<C>(pv[i]);
        const int64_t o = (int64_t)bi * hw + px;
        if (entropy) { PixVec<VEC> e; for (int i = 0; i < VEC; ++i) e.v[i] = u[i]; e.store(entropy + o); }
        if (weight) {
            PixVec<VEC> wv;
#pragma unroll
            for (int i = 0; i < VEC; ++i) wv.v[i] = use_weight ? uvem_weight_dev(u[i], m, t, inv_gamma, cl, cr) : 1.0f;
            wv.store(weight + o);
        }
        if (TERMS) {
            int64_t tg[VEC];
            load_ids<VEC>(target + o, tg);
#pragma unroll
            for (int i = 0; i < VEC; ++i) {
                if (gate) gate[o + i] = (u[i] > t) ? 1 : 0;            // balance.py:373 (NaN -> not gated)
                valid += (u[i] <= t) && (tg[i] != ignore_label);        // balance.py:382
            }
        }
    }
    if (TERMS) {
        valid = __reduce_add_sync(0xffffffffu, valid);
        __shared__ int sv[8];
        if ((threadIdx.x & 31) == 0) sv[threadIdx.x >> 5] = valid;
        __syncthreads();
        if (threadIdx.x == 0) {
            int tot = 0;
            for (int i = 0; i < (int)(blockDim.x >> 5); ++i) tot += sv[i];
            if (tot) atomicAdd(valid_cnt, (unsigned long long)tot);
        }
    }
}

__global__ void __launch_bounds__(256) uvem_weight_kernel(const float* __restrict__ u, int64_t n, float m, float t,
                                                          float inv_gamma, float cl, float cr, float* __restrict__ w) {
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x)
        w[i] = uvem_weight_dev(u[i], m, t, inv_gamma, cl, cr);
}

}  // namespace

extern "C" int uem_softmax_conf_entropy_argmax_f32(const float* x1, const float* x2, int b, int c, int h, int w, int H,
                                                   int W, float temp, float* soft, float* conf, float* entropy,
                                                   int64_t* argmax, void* stream) {
    UEM_REQUIRE(x1 && b > 0 && h > 0 && w > 0 && H > 0 && W > 0, "uem_softmax_conf_entropy_argmax_f32: bad arguments");
    UEM_REQUIRE(temp > 0.f, "uem_softmax_conf_entropy_argmax_f32: temp must be > 0");  // alignment.py:313
    cudaStream_t st = (cudaStream_t)stream;
    const bool vec = (W % 4 == 0) && (!soft || uem_aligned16(soft)) && (!conf || uem_aligned16(conf)) &&
                     (!entropy || uem_aligned16(entropy)) && (!argmax || uem_aligned16(argmax));
    const float sy = uem_align_corners_scale(h, H), sx = uem_align_corners_scale(w, W);
    const int nmaps = x2 ? 2 : 1;
    UEM_DISPATCH_C(c, {
        size_t smem = (size_t)nmaps * C * w * sizeof(float);
        UEM_REQUIRE(smem <= 227 * 1024, "uem_softmax_conf_entropy_argmax_f32: low-res width %d too large", w);
        dim3 grid(H, b);
        if (vec) {
            int threads = min(256, max(32, ((W / 4 + 31) / 32) * 32));
            if (smem > 48 * 1024) UEM_CUDA(cudaFuncSetAttribute(logits_pass_kernel<C, 4>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
            logits_pass_kernel<C, 4><<<grid, threads, smem, st>>>(x1, x2, h, w, H, W, sy, sx, 1.0f / temp, temp != 1.0f, temp,
                                                                 soft, conf, entropy, argmax);
        } else {
            int threads = min(256, max(32, ((W + 31) / 32) * 32));
            if (smem > 48 * 1024) UEM_CUDA(cudaFuncSetAttribute(logits_pass_kernel<C, 1>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
            logits_pass_kernel<C, 1><<<grid, threads, smem, st>>>(x1, x2, h, w, H, W, sy, sx, 1.0f / temp, temp != 1.0f, temp,
                                                                 soft, conf, entropy, argmax);
        }
    });
    UEM_CHECK_LAUNCH();
    return 0;
}

static int launch_entropy(const float* soft, const int64_t* target, int b, int c, int64_t hw, float m, float t,
                          float inv_gamma, float cl, float cr, int use_weight, int64_t ignore_label, float* entropy,
                          float* weight, uint8_t* gate, int64_t* valid_cnt, bool terms, cudaStream_t st) {
    const bool vec = (hw % 4 == 0) && uem_aligned16(soft) && (!entropy || uem_aligned16(entropy)) &&
                     (!weight || uem_aligned16(weight)) && (!target || uem_aligned16(target));
    UEM_DISPATCH_C(c, {
        if (vec) {
            dim3 grid(uem_div_up(hw / 4, 256), b);
            if (terms) entropy_weight_kernel<C, 4, true><<<grid, 256, 0, st>>>(soft, target, hw, m, t, inv_gamma, cl, cr, use_weight, ignore_label, entropy, weight, gate, (unsigned long long*)valid_cnt);
            else entropy_weight_kernel<C, 4, false><<<grid, 256, 0, st>>>(soft, target, hw, m, t, inv_gamma, cl, cr, use_weight, ignore_label, entropy, weight, gate, nullptr);
        } else {
            dim3 grid(uem_div_up(hw, 256), b);
            if (terms) entropy_weight_kernel<C, 1, true><<<grid, 256, 0, st>>>(soft, target, hw, m, t, inv_gamma, cl, cr, use_weight, ignore_label, entropy, weight, gate, (unsigned long long*)valid_cnt);
            else entropy_weight_kernel<C, 1, false><<<grid, 256, 0, st>>>(soft, target, hw, m, t, inv_gamma, cl, cr, use_weight, ignore_label, entropy, weight, gate, nullptr);
        }
    });
    UEM_CHECK_LAUNCH();
    return 0;
}

extern "C" int uem_entropy_uvem_weight_f32(const float* soft, int b, int c, int64_t hw, float m, float t, float inv_gamma,
                                           float coef_left, float coef_right, float* entropy, float* weight, void* stream) {
    UEM_REQUIRE(soft && b > 0 && hw > 0 && (entropy || weight), "uem_entropy_uvem_weight_f32: bad arguments");
    return launch_entropy(soft, nullptr, b, c, hw, m, t, inv_gamma, coef_left, coef_right, 1, -1, entropy, weight, nullptr,
                          nullptr, false, (cudaStream_t)stream);
}

extern "C" int uem_uvem_terms_f32(const float* soft, const int64_t* target, int b, int c, int64_t hw, float m, float t,
                                  float inv_gamma, float coef_left, float coef_right, int use_weight, int64_t ignore_label,
                                  float* weight, uint8_t* gate, int64_t* valid_cnt, void* stream) {
    UEM_REQUIRE(soft && target && valid_cnt && b > 0 && hw > 0, "uem_uvem_terms_f32: bad arguments");
    return launch_entropy(soft, target, b, c, hw, m, t, inv_gamma, coef_left, coef_right, use_weight, ignore_label, nullptr,
                          weight, gate, valid_cnt, true, (cudaStream_t)stream);
}

extern "C" int uem_uvem_weight_f32(const float* u, int64_t n, float m, float t, float inv_gamma, float coef_left,
                                   float coef_right, float* weight, void* stream) {
    UEM_REQUIRE(u && weight && n >= 0, "uem_uvem_weight_f32: bad arguments");
    if (n == 0) return 0;
    int grid = (int)min((int64_t)UEM_SMS * 8, (n + 255) / 256);
    uvem_weight_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(u, n, m, t, inv_gamma, coef_left, coef_right, weight);
    UEM_CHECK_LAUNCH();
    return 0;
}
