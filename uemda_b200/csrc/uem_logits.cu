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

// ------------------------------------------------------------------------------------------------
// Column-walk form of the fused logits pass (same decomposition as refine_col_kernel, uem_refine.cu): a thread owns
// ONE image column of a 128-column strip and walks down a contiguous range of rows.  The horizontal half of the
// bilinear interpolation is done once per low-res row pair (A = lerp_x(row i0), B = lerp_x(row i1), kept in registers
// pre-scaled by log2 e / temp as A and D = B - A); per output row a logit is one packed FFMA (z = A + t_y * D) and the
// softmax exponent a bare EX2.  There is no full-resolution input at all: the kernel is bound by its stores
// (4c + 16 bytes per pixel), which are 32-bit per lane and contiguous over the warp (full 128-byte lines).
// ------------------------------------------------------------------------------------------------
// VX = image columns per lane (2 when the row width is even: 64-bit stores, shared per-row bookkeeping, two interleaved
// dependent chains; 1 otherwise)
template <int C, int NM, int VX>
__global__ void __launch_bounds__(128, VX == 2 ? 4 : 5) logits_col_kernel(const float* __restrict__ x1, const float* __restrict__ x2, int b,
                                                                          int h, int w, int H, int W, float sy, float sx, float scale,
                                                                          float* __restrict__ soft, float* __restrict__ conf,
                                                                          float* __restrict__ entropy, int64_t* __restrict__ argmax,
                                                                          const int ncols_max) {
    constexpr int NT = 128, CP = (C + 3) & ~3, PC = (C + 1) / 2, TS = NM * CP, WC = 32 * VX;
    extern __shared__ __align__(16) float taps_all[];
    const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
    float* const taps = taps_all + (size_t)wid * 2 * ncols_max * TS;
    const int64_t HW = (int64_t)H * W;
    const int hw_low = h * w;
    const int nstrips = (W + NT * VX - 1) / (NT * VX);
    const int64_t total = (int64_t)b * nstrips * H;
    const int64_t U0 = total * blockIdx.x / gridDim.x, U1 = total * (blockIdx.x + 1) / gridDim.x;
    const int n = (int)(U1 - U0);
    if (n <= 0) return;
    int bs = (int)(U0 / H), y = (int)(U0 - (int64_t)bs * H);
    const float* maps[2] = {x1, x2};

    float2 A[VX][NM][PC], D[VX][NM][PC];
    int cur_bs = -1, cur_i0 = -1, bi = 0;
    int a0[VX], a1[VX], abase = 0, ncols = 1;
    float l0x[VX], l1x[VX];
#pragma unroll
    for (int v = 0; v < VX; ++v) { a0[v] = a1[v] = 0; l0x[v] = l1x[v] = 0.f; }
    bool active = false;
    uint32_t x = 0;
    for (int it = 0; it < n; ++it) {
        if (bs != cur_bs) {        // new (image, strip): column geometry
            bi = bs / nstrips;
            const int xw = (bs - bi * nstrips) * NT * VX + wid * WC;
            x = (uint32_t)(xw + lane * VX);
            active = (int)x < W;   // VX == 2 only with an even W: both columns inside or both outside
#pragma unroll
            for (int v = 0; v < VX; ++v) {
                const Lerp lx = make_lerp(active ? (int)x + v : W - 1, w, sx);
                a0[v] = lx.i0; a1[v] = lx.i1; l0x[v] = lx.l0; l1x[v] = lx.l1;
            }
            abase = make_lerp(min(xw, W - 1), w, sx).i0;
            ncols = make_lerp(min(xw + WC - 1, W - 1), w, sx).i1 - abase + 1;
            cur_bs = bs;
            cur_i0 = -1;
        }
        const Lerp ly = make_lerp(y, h, sy);
        if (ly.i0 != cur_i0) {     // new low-res row pair: fetched once per warp, interpolated horizontally once per lane
            cur_i0 = ly.i0;
            __syncwarp();
            if (lane < NM * C) {
                const int m = lane / C, ci = lane - m * C;
                const float* plane = maps[m] + ((int64_t)bi * C + ci) * hw_low + abase;
                float* dst = taps + m * CP + ci;
                const float* r0 = plane + ly.i0 * w;
                const float* r1 = plane + ly.i1 * w;
                for (int j0 = 0; j0 < ncols; j0 += 4) {   // 8 loads in flight: one L2 round trip per 4 columns
                    float u0[4], u1[4];
#pragma unroll
                    for (int u = 0; u < 4; ++u) {
                        const int j = min(j0 + u, ncols - 1);
                        u0[u] = __ldg(r0 + j);
                        u1[u] = __ldg(r1 + j);
                    }
#pragma unroll
                    for (int u = 0; u < 4; ++u)
                        if (j0 + u < ncols) {
                            dst[(j0 + u) * TS] = u0[u] * scale;
                            dst[(ncols_max + j0 + u) * TS] = u1[u] * scale;
                        }
                }
            }
            __syncwarp();
#pragma unroll
            for (int v = 0; v < VX; ++v) {
                const float* t00 = taps + (a0[v] - abase) * TS;
                const float* t01 = taps + (a1[v] - abase) * TS;
                const float2 l0 = make_float2(l0x[v], l0x[v]), l1 = make_float2(l1x[v], l1x[v]);
#pragma unroll
                for (int m = 0; m < NM; ++m) {
                    float v00[CP], v01[CP], v10[CP], v11[CP];
#pragma unroll
                    for (int q = 0; q < CP / 4; ++q) {
                        const float4 q00 = *reinterpret_cast<const float4*>(t00 + m * CP + 4 * q);
                        const float4 q01 = *reinterpret_cast<const float4*>(t01 + m * CP + 4 * q);
                        const float4 q10 = *reinterpret_cast<const float4*>(t00 + ncols_max * TS + m * CP + 4 * q);
                        const float4 q11 = *reinterpret_cast<const float4*>(t01 + ncols_max * TS + m * CP + 4 * q);
                        v00[4 * q] = q00.x; v00[4 * q + 1] = q00.y; v00[4 * q + 2] = q00.z; v00[4 * q + 3] = q00.w;
                        v01[4 * q] = q01.x; v01[4 * q + 1] = q01.y; v01[4 * q + 2] = q01.z; v01[4 * q + 3] = q01.w;
                        v10[4 * q] = q10.x; v10[4 * q + 1] = q10.y; v10[4 * q + 2] = q10.z; v10[4 * q + 3] = q10.w;
                        v11[4 * q] = q11.x; v11[4 * q + 1] = q11.y; v11[4 * q + 2] = q11.z; v11[4 * q + 3] = q11.w;
                    }
#pragma unroll
                    for (int j = 0; j < PC; ++j) {
                        const bool pad = 2 * j + 1 >= C;
                        const float2 p00 = make_float2(v00[2 * j], pad ? 0.f : v00[2 * j + 1]);
                        const float2 p01 = make_float2(v01[2 * j], pad ? 0.f : v01[2 * j + 1]);
                        const float2 p10 = make_float2(v10[2 * j], pad ? 0.f : v10[2 * j + 1]);
                        const float2 p11 = make_float2(v11[2 * j], pad ? 0.f : v11[2 * j + 1]);
                        float2 ta = __ffma2_rn(l1, p01, __fmul2_rn(l0, p00));
                        const float2 tb = __ffma2_rn(l1, p11, __fmul2_rn(l0, p10));
                        float2 td = __fadd2_rn(tb, make_float2(-ta.x, -ta.y));
                        if (pad) { ta.y = -1e30f; td.y = 0.f; }   // EX2 gives exactly 0, never wins a max
                        A[v][m][j] = ta;
                        D[v][m][j] = td;
                    }
                }
            }
        }
        if (active) {
            const float2 t2 = make_float2(ly.l1, ly.l1);
            float pr[VX][C], best[VX], ent[VX];
            int arg[VX];
#pragma unroll
            for (int v = 0; v < VX; ++v) {
                float2 e[NM][PC];
                float rs[NM];
#pragma unroll
                for (int m = 0; m < NM; ++m) {
                    float mx = -INFINITY;
#pragma unroll
                    for (int j = 0; j < PC; ++j) {
                        e[m][j] = __ffma2_rn(t2, D[v][m][j], A[v][m][j]);
                        mx = fmaxf(mx, fmaxf(e[m][j].x, e[m][j].y));
                    }
                    float2 acc;
#pragma unroll
                    for (int j = 0; j < PC; ++j) {
                        const float2 d = __fadd2_rn(e[m][j], make_float2(-mx, -mx));
                        e[m][j] = make_float2(ex2_approx(d.x), ex2_approx(d.y));
                        acc = j ? __fadd2_rn(acc, e[m][j]) : e[m][j];
                    }
                    // 1/S, Newton-refined (rcp.approx alone is 1 ulp; the soft labels are this kernel's product)
                    const float S = acc.x + acc.y;
                    float r = rcp_approx(S);
                    r = fmaf(r, fmaf(-S, r, 1.0f), r);
                    rs[m] = (NM == 2) ? 0.5f * r : r;
                }
#pragma unroll
                for (int j = 0; j < PC; ++j) {
                    float2 q = __fmul2_rn(e[0][j], make_float2(rs[0], rs[0]));
                    if (NM == 2) q = __ffma2_rn(e[1][j], make_float2(rs[1], rs[1]), q);
                    pr[v][2 * j] = q.x;
                    if (2 * j + 1 < C) pr[v][2 * j + 1] = q.y;
                }
                best[v] = -INFINITY;
                arg[v] = 0;
#pragma unroll
                for (int ci = 0; ci < C; ++ci) {   // first (lowest) index wins ties: torch.max / argmax
                    const bool gt = pr[v][ci] > best[v];
                    best[v] = gt ? pr[v][ci] : best[v];
                    arg[v] = gt ? ci : arg[v];
                }
                if (entropy) ent[v] = entropy_px<C>(pr[v]);   // balance.py:372; p==0 -> NaN like the reference
            }
            const uint32_t idx = (uint32_t)y * (uint32_t)W + x;
            if constexpr (VX == 2) {   // idx is even and every plane base is 8-byte aligned (checked by the launcher)
                if (soft) {
                    float* sb = soft + (int64_t)bi * C * HW;
#pragma unroll
                    for (int ci = 0; ci < C; ++ci) *reinterpret_cast<float2*>(sb + (int64_t)ci * HW + idx) = make_float2(pr[0][ci], pr[1][ci]);
                }
                if (conf) *reinterpret_cast<float2*>(conf + (int64_t)bi * HW + idx) = make_float2(best[0], best[1]);
                if (entropy) *reinterpret_cast<float2*>(entropy + (int64_t)bi * HW + idx) = make_float2(ent[0], ent[1]);
                if (argmax) stg_i64x2(argmax + (int64_t)bi * HW + idx, arg[0], arg[1]);
            } else {
                if (soft) {
                    float* sb = soft + (int64_t)bi * C * HW;
#pragma unroll
                    for (int ci = 0; ci < C; ++ci) (sb + (int64_t)ci * HW)[idx] = pr[0][ci];
                }
                if (conf) (conf + (int64_t)bi * HW)[idx] = best[0];
                if (entropy) (entropy + (int64_t)bi * HW)[idx] = ent[0];
                if (argmax) (argmax + (int64_t)bi * HW)[idx] = arg[0];
            }
        }
        if (++y == H) { y = 0; ++bs; }
    }
}

template <int C, int NM, int VX>
static int launch_logits_col_vx(const float* x1, const float* x2, int b, int h, int w, int H, int W, float sy, float sx, float temp,
                                float* soft, float* conf, float* entropy, int64_t* argmax, cudaStream_t st, bool* done) {
    *done = false;
    constexpr int CP = (C + 3) & ~3;
    int ncols_max = (int)((32 * VX - 1) * sx) + 3;
    if (ncols_max > w) ncols_max = w;
    const size_t smem = (size_t)4 * 2 * ncols_max * NM * CP * 4;
    if (smem > 64 * 1024 || (int64_t)H * W >= ((int64_t)1 << 31)) return 0;
    auto kernel = logits_col_kernel<C, NM, VX>;
    if (smem > 48 * 1024) UEM_CUDA(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    int per_sm = 0;
    UEM_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kernel, 128, smem));
    if (per_sm < 1) per_sm = 1;
    static int sms = 0;
    if (sms == 0) {
        int dev = 0;
        if (cudaGetDevice(&dev) != cudaSuccess || cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || sms <= 0)
            sms = UEM_SMS;
    }
    const int64_t total = (int64_t)b * ((W + 128 * VX - 1) / (128 * VX)) * H;
    const int grid = (int)min(total, (int64_t)sms * per_sm);
    kernel<<<grid, 128, smem, st>>>(x1, x2, b, h, w, H, W, sy, sx, (float)(1.4426950408889634 / (double)temp), soft, conf, entropy,
                                    argmax, ncols_max);
    *done = true;
    return 0;
}

template <int C, int NM>
static int launch_logits_col(const float* x1, const float* x2, int b, int h, int w, int H, int W, float sy, float sx, float temp,
                             float* soft, float* conf, float* entropy, int64_t* argmax, cudaStream_t st, bool* done) {
    // two columns per lane need 64-bit aligned rows: even W (then H*W and every plane offset are even) and aligned bases
    auto al8 = [](const void* p) { return (reinterpret_cast<uintptr_t>(p) & 7u) == 0; };
    const bool pair = (W % 2 == 0) && al8(soft) && al8(conf) && al8(entropy) && (!argmax || uem_aligned16(argmax));
    if (pair) return launch_logits_col_vx<C, NM, 2>(x1, x2, b, h, w, H, W, sy, sx, temp, soft, conf, entropy, argmax, st, done);
    return launch_logits_col_vx<C, NM, 1>(x1, x2, b, h, w, H, W, sy, sx, temp, soft, conf, entropy, argmax, st, done);
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
    {   // column-walk kernel (any width: its stores are scalar per lane)
        bool done = false;
        int rc = 0;
        UEM_DISPATCH_C(c, {
            if (nmaps == 2) rc = launch_logits_col<C, 2>(x1, x2, b, h, w, H, W, sy, sx, temp, soft, conf, entropy, argmax, st, &done);
            else rc = launch_logits_col<C, 1>(x1, x2, b, h, w, H, W, sy, sx, temp, soft, conf, entropy, argmax, st, &done);
        });
        if (rc) return rc;
        if (done) {
            UEM_CHECK_LAUNCH();
            return 0;
        }
    }
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
