// Grouped D x D projection with per-group weights (forward + backward).
//
//   Y[g][r][:] = act( X[g][r][:] . W[g] + bias[g] (+ Res[g][r][:]) )          g < G groups, r < R rows
//
// Element (g, r, j) of X / Res / Y lives at  base + g*group_stride + r*row_stride + j  (j < D contiguous).
// Two uses on the GPT-ST path, both over the (B,T,N,D) activation:
//   * time-adaptive projection  (group = (b,t): R = N rows, row_stride = D, group_stride = N*D)
//       hyperTem  GPTST.py:160-163,  MLP_RL  GPTST.py:29-32
//   * node-adaptive projection  (group = n:     R = B*T rows, row_stride = N*D, group_stride = D)
//       cap       GPTST.py:137-141,  MLP_RL  GPTST.py:24-27
// The contraction runs on tensor cores (mma.sync m16n8k8 tf32, optional 3xTF32 split, fp32 accumulate).
//
// Backward (same addressing):   dy = dY * act'(Y)
//   dX[g][r][:]  = dy . W[g]^T        dW[g] = sum_r X^T dy        dbias[g] = sum_r dy        dRes = dy
#include <cstdlib>

#include "common.cuh"

namespace gptst {

template <int D>
struct GProjCfg {
    static constexpr int BM = (D <= 64) ? 128 : 64;     // rows per tile
    static constexpr int WM = BM / 16;                  // warps along rows
    static constexpr int WN = 8 / WM;                   // warps along columns
    static constexpr int NT = D / 8 / WN;               // n-tiles per warp
    static constexpr int LDW_F = D + 8;                 // fwd:  B(k,n) = W[k][n]  (k-major)
    static constexpr int LDX_F = D + 4;                 // fwd:  A row-major
    static constexpr int LDW_B = D + 4;                 // bwd:  B(k=o,n=i) = W[i][o] (n-major)
    static constexpr int LDX_B = D + 8;                 // bwd:  X read transposed for dW
    static constexpr int LDY_B = D + 4;                 // bwd:  dy as A row-major (and k-major B for dW)
};

template <int D, int PREC>
__global__ void __launch_bounds__(256) gproj_fwd_kernel(const float* __restrict__ X, const float* __restrict__ W,
                                                        const float* __restrict__ bias, const float* __restrict__ Res,
                                                        float* __restrict__ Y, int G, int R, long group_stride,
                                                        long row_stride, int act) {
    using C = GProjCfg<D>;
    extern __shared__ __align__(16) unsigned char smem_raw[];
    uint32_t* Wh = reinterpret_cast<uint32_t*>(smem_raw);
    uint32_t* Wl = Wh + D * C::LDW_F;
    float* Xs = reinterpret_cast<float*>(Wl + (PREC == PREC_3XTF32 ? D * C::LDW_F : 0));
    float* bs = Xs + C::BM * C::LDX_F;

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int wm = warp % C::WM, wn = warp / C::WM;
    const int g = blockIdx.x;
    const float* Wg = W + (size_t)g * D * D;
    for (int i = tid; i < D * D / 4; i += 256) {
        int k = (i * 4) / D, n = (i * 4) % D;
        float4 w = *reinterpret_cast<const float4*>(Wg + (size_t)i * 4);
        float wv[4] = {w.x, w.y, w.z, w.w};
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            uint32_t hi, lo;
            split_tf32<PREC>(wv[j], hi, lo);
            Wh[k * C::LDW_F + n + j] = hi;
            if (PREC == PREC_3XTF32) Wl[k * C::LDW_F + n + j] = lo;
        }
    }
    for (int i = tid; i < D; i += 256) bs[i] = bias ? bias[(size_t)g * D + i] : 0.f;

    const float* Xg = X + (size_t)g * group_stride;
    const float* Rg = Res ? Res + (size_t)g * group_stride : nullptr;
    float* Yg = Y + (size_t)g * group_stride;
    const int ntiles = (R + C::BM - 1) / C::BM;
    for (int tile = blockIdx.y; tile < ntiles; tile += gridDim.y) {
        const int r0 = tile * C::BM;
        __syncthreads();  // previous tile's readers done (and W/bias visible on first pass)
        for (int i = tid; i < C::BM * D / 4; i += 256) {
            int r = (i * 4) / D, c = (i * 4) % D;
            float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
            if (r0 + r < R) v = *reinterpret_cast<const float4*>(Xg + (size_t)(r0 + r) * row_stride + c);
            *reinterpret_cast<float4*>(Xs + r * C::LDX_F + c) = v;
        }
        __syncthreads();
        float acc[C::NT][4];
#pragma unroll
        for (int nt = 0; nt < C::NT; ++nt) acc[nt][0] = acc[nt][1] = acc[nt][2] = acc[nt][3] = 0.f;
        const int ncol0 = wn * C::NT * 8;
        warp_gemm_presplit<D, C::NT, PREC, true>(acc, Xs + wm * 16 * C::LDX_F, C::LDX_F, Wh + ncol0, Wl + ncol0,
                                                 C::LDW_F, lane);
        const int gq = lane >> 2, tq = lane & 3;
#pragma unroll
        for (int half = 0; half < 2; ++half) {
            const int r = r0 + wm * 16 + gq + half * 8;
            if (r < R) {
#pragma unroll
                for (int nt = 0; nt < C::NT; ++nt) {
                    const int c = ncol0 + nt * 8 + 2 * tq;
                    float y0 = acc[nt][half * 2 + 0] + bs[c], y1 = acc[nt][half * 2 + 1] + bs[c + 1];
                    const size_t off = (size_t)r * row_stride + c;
                    if (Rg) {
                        float2 rr = *reinterpret_cast<const float2*>(Rg + off);
                        y0 += rr.x;
                        y1 += rr.y;
                    }
                    if (act) {
                        y0 = lrelu(y0);
                        y1 = lrelu(y1);
                    }
                    *reinterpret_cast<float2*>(Yg + off) = make_float2(y0, y1);
                }
            }
        }
    }
}

// dW / dbias partial layout: [gridDim.y][G][D*D] and [gridDim.y][G][D]; the host sums over the leading dim.
// 64-row tiles and a single fp32 copy of W[g] (operands are split on the fly: 2 instructions) keep the CTA at ~54 KB
// of shared memory for D = 64, i.e. four CTAs (32 warps) per SM to hide the global-load latency of the tile loop.
template <int D>
struct GProjBwdCfg {
    static constexpr int BM = 64;
    static constexpr int WM = BM / 16;                  // 4 warps along rows
    static constexpr int WN = 8 / WM;                   // 2 warps along columns
    static constexpr int NT = D / 8 / WN;               // n-tiles per warp (dX)
    static constexpr int LDW = D + 4;                   // B(k=o,n=i) = W[i][o]  (n-major)
    static constexpr int LDX = D + 8;                   // X read transposed for dW
    static constexpr int LDY = D + 4;                   // dy as A row-major (and k-major B for dW)
};

template <int D, int PREC>
__global__ void __launch_bounds__(256) gproj_bwd_kernel(const float* __restrict__ dY, const float* __restrict__ Y,
                                                        const float* __restrict__ X, const float* __restrict__ W,
                                                        float* __restrict__ dX, float* __restrict__ dWp,
                                                        float* __restrict__ dbp, float* __restrict__ dRes, int G, int R,
                                                        long group_stride, long row_stride, int act) {
    using C = GProjBwdCfg<D>;
    extern __shared__ __align__(16) unsigned char smem_raw[];
    float* Ws = reinterpret_cast<float*>(smem_raw);     // [D][LDW]  W[i][o]
    float* Xs = Ws + D * C::LDW;                        // [BM][LDX]
    float* Ys = Xs + C::BM * C::LDX;                    // [BM][LDY] dy
    float* red = Ys + C::BM * C::LDY;                   // [256/D][D] column partial sums

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int wm = warp % C::WM, wn = warp / C::WM;
    const int g = blockIdx.x;
    const float* Wg = W + (size_t)g * D * D;
    for (int i = tid; i < D * D / 4; i += 256) {
        int k = (i * 4) / D, n = (i * 4) % D;
        *reinterpret_cast<float4*>(Ws + k * C::LDW + n) = *reinterpret_cast<const float4*>(Wg + (size_t)i * 4);
    }
    // dW accumulators: output D x D, tiles of 16 x 8; warp owns one m-tile and NT_W n-tiles
    constexpr int MT = D / 16, NTT = D / 8;
    constexpr int WMG = (MT >= 8) ? 8 : MT;
    constexpr int WNG = 8 / WMG;
    constexpr int MT_W = MT / WMG;
    constexpr int NT_W = NTT / WNG;
    static_assert(MT_W == 1, "dW tiling");
    const int gm = warp % WMG, gn = warp / WMG;
    float gacc[NT_W][4];
#pragma unroll
    for (int nt = 0; nt < NT_W; ++nt) gacc[nt][0] = gacc[nt][1] = gacc[nt][2] = gacc[nt][3] = 0.f;
    float sigma = 0.f;  // thread tid < D owns column tid of dbias

    const float* Xg = X + (size_t)g * group_stride;
    const float* Yg = Y ? Y + (size_t)g * group_stride : nullptr;
    const float* dYg = dY + (size_t)g * group_stride;
    float* dXg = dX + (size_t)g * group_stride;
    float* dRg = dRes ? dRes + (size_t)g * group_stride : nullptr;
    const int ntiles = (R + C::BM - 1) / C::BM;
    for (int tile = blockIdx.y; tile < ntiles; tile += gridDim.y) {
        const int r0 = tile * C::BM;
        __syncthreads();
        for (int i = tid; i < C::BM * D / 4; i += 256) {
            int r = (i * 4) / D, c = (i * 4) % D;
            float4 x = make_float4(0.f, 0.f, 0.f, 0.f), d = x;
            if (r0 + r < R) {
                const size_t off = (size_t)(r0 + r) * row_stride + c;
                x = *reinterpret_cast<const float4*>(Xg + off);
                d = *reinterpret_cast<const float4*>(dYg + off);
                if (act) {
                    float4 y = *reinterpret_cast<const float4*>(Yg + off);
                    d.x = lrelu_grad(y.x, d.x);
                    d.y = lrelu_grad(y.y, d.y);
                    d.z = lrelu_grad(y.z, d.z);
                    d.w = lrelu_grad(y.w, d.w);
                }
                if (dRg) *reinterpret_cast<float4*>(dRg + off) = d;
            }
            *reinterpret_cast<float4*>(Xs + r * C::LDX + c) = x;
            *reinterpret_cast<float4*>(Ys + r * C::LDY + c) = d;
        }
        __syncthreads();
        // ---- dbias: column sums of dy
        {
            constexpr int PARTS = 256 / D;
            const int c = tid % D, part = tid / D;
            float s = 0.f;
            if (part < PARTS)
                for (int r = part; r < C::BM; r += PARTS) s += Ys[r * C::LDY + c];
            if (part < PARTS) red[part * D + c] = s;
        }
        // ---- dX = dy . W^T       B(k=o, n=i) = Ws[n][k]
        {
            float acc[C::NT][4];
#pragma unroll
            for (int nt = 0; nt < C::NT; ++nt) acc[nt][0] = acc[nt][1] = acc[nt][2] = acc[nt][3] = 0.f;
            const int ncol0 = wn * C::NT * 8;
            warp_gemm<D, C::NT, PREC, false, false>(acc, Ys + wm * 16 * C::LDY, C::LDY, Ws + ncol0 * C::LDW, C::LDW, lane);
            const int gq = lane >> 2, tq = lane & 3;
#pragma unroll
            for (int half = 0; half < 2; ++half) {
                const int r = r0 + wm * 16 + gq + half * 8;
                if (r < R) {
#pragma unroll
                    for (int nt = 0; nt < C::NT; ++nt) {
                        const int c = ncol0 + nt * 8 + 2 * tq;
                        *reinterpret_cast<float2*>(dXg + (size_t)r * row_stride + c) =
                            make_float2(acc[nt][half * 2 + 0], acc[nt][half * 2 + 1]);
                    }
                }
            }
        }
        // ---- dW += X^T . dy      (M = i, N = o, K = rows of this tile; zero rows contribute nothing)
        warp_gemm<C::BM, NT_W, PREC, true, true>(gacc, Xs + gm * 16, C::LDX, Ys + gn * NT_W * 8, C::LDY, lane);
        __syncthreads();
        if (tid < D) {
            constexpr int PARTS = 256 / D;
#pragma unroll
            for (int p = 0; p < PARTS; ++p) sigma += red[p * D + tid];
        }
    }
    // ---- write partials
    float* dWo = dWp + ((size_t)blockIdx.y * G + g) * D * D;
    const int gq = lane >> 2, tq = lane & 3;
#pragma unroll
    for (int nt = 0; nt < NT_W; ++nt) {
        const int c = (gn * NT_W + nt) * 8 + 2 * tq;
        const int r = gm * 16 + gq;
        *reinterpret_cast<float2*>(dWo + (size_t)r * D + c) = make_float2(gacc[nt][0], gacc[nt][1]);
        *reinterpret_cast<float2*>(dWo + (size_t)(r + 8) * D + c) = make_float2(gacc[nt][2], gacc[nt][3]);
    }
    if (tid < D) dbp[((size_t)blockIdx.y * G + g) * D + tid] = sigma;
}

template <int D>
static size_t gproj_fwd_smem(int prec) {
    using C = GProjCfg<D>;
    return (size_t)(prec == PREC_3XTF32 ? 2 : 1) * D * C::LDW_F * 4 + (size_t)C::BM * C::LDX_F * 4 + D * 4;
}
template <int D>
static size_t gproj_bwd_smem(int) {
    using C = GProjBwdCfg<D>;
    return (size_t)D * C::LDW * 4 + (size_t)C::BM * (C::LDX + C::LDY) * 4 + 256 * 4;
}

template <int D, int PREC>
static cudaError_t launch_fwd(const float* X, const float* W, const float* bias, const float* Res, float* Y, int G, int R,
                              long gs, long rs, int act, int splits, cudaStream_t st) {
    size_t smem = gproj_fwd_smem<D>(PREC);
    cudaError_t e = cudaFuncSetAttribute(gproj_fwd_kernel<D, PREC>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return e;
    gproj_fwd_kernel<D, PREC><<<dim3(G, splits), 256, smem, st>>>(X, W, bias, Res, Y, G, R, gs, rs, act);
    return cudaGetLastError();
}
template <int D, int PREC>
static cudaError_t launch_bwd(const float* dY, const float* Y, const float* X, const float* W, float* dX, float* dWp,
                              float* dbp, float* dRes, int G, int R, long gs, long rs, int act, int splits,
                              cudaStream_t st) {
    size_t smem = gproj_bwd_smem<D>(PREC);
    cudaError_t e = cudaFuncSetAttribute(gproj_bwd_kernel<D, PREC>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return e;
    gproj_bwd_kernel<D, PREC><<<dim3(G, splits), 256, smem, st>>>(dY, Y, X, W, dX, dWp, dbp, dRes, G, R, gs, rs, act);
    return cudaGetLastError();
}

}  // namespace gptst

namespace gptst {
// tcgen05 / TMEM implementation for D = 64 (gproj_umma.cu)
cudaError_t gproj_fwd_umma(const float* X, const float* W, const float* bias, const float* Res, float* Y, int G, int R,
                           long gs, long rs, int act, int prec, cudaStream_t st);
cudaError_t gproj_bwd_umma(const float* dY, const float* Y, const float* X, const float* W, float* dX, float* dWp, float* dbp,
                           float* dRes, int G, int R, long gs, long rs, int act, int prec, int splits, cudaStream_t st);
// second generation for D = 64 (gproj2.cu): fp16-split mma.sync m16n8k16, forward and backward
int gproj2_splits(int G, int R);
cudaError_t gproj2_fwd(const float* X, const float* W, const float* bias, const float* Res, float* Y, int G, int R, long gs,
                       long rs, int act, int prec, cudaStream_t st, const float* Gate = nullptr, float* Zout = nullptr);
cudaError_t gproj2_bwd(const float* dY, const float* Y, const float* X, const float* W, float* dX, float* dWp, float* dbp,
                       float* dRes, int G, int R, long gs, long rs, int act, int prec, int splits, int flags, cudaStream_t st);
// GPTST_B200_GPROJ = "umma" (tcgen05 forward) | "mma" (first-generation tf32 mma.sync) select the older kernels (A/B testing)
static int gproj_impl() {
    static int v = -1;
    if (v < 0) {
        const char* e = getenv("GPTST_B200_GPROJ");
        v = (e && e[0] == 'm') ? 0 : (e && e[0] == 'u') ? 1 : 2;
    }
    return v;
}
static bool use_umma() { return gproj_impl() == 1; }
static bool use_gp2(int D, int prec) { return gproj_impl() == 2 && D == 64 && (prec == 1 || prec == 3); }
// The tcgen05(dX)+mma.sync(dW) backward is correct but (one CTA per SM, serial phases) still a little slower than the
// two-CTA mma.sync kernel on PEMS08 shapes; opt-in with GPTST_B200_GPROJ_BWD=umma until it is pipelined.
static bool use_umma_bwd() {
    static int v = -1;
    if (v < 0) {
        const char* e = getenv("GPTST_B200_GPROJ_BWD");
        v = (e && e[0] == 'u') ? 1 : 0;
    }
    return v == 1;
}
}  // namespace gptst

using namespace gptst;

#define DISPATCH_D_PREC(D_, P_, CALL)                                          \
    do {                                                                       \
        if ((D_) == 32 && (P_) == 1) { CALL(32, 1); }                          \
        else if ((D_) == 32 && (P_) == 3) { CALL(32, 3); }                     \
        else if ((D_) == 64 && (P_) == 1) { CALL(64, 1); }                     \
        else if ((D_) == 64 && (P_) == 3) { CALL(64, 3); }                     \
        else if ((D_) == 128 && (P_) == 1) { CALL(128, 1); }                   \
        else if ((D_) == 128 && (P_) == 3) { CALL(128, 3); }                   \
        else return -2;                                                        \
    } while (0)

extern "C" int gptst_gproj_splits(int G, int R, int D) {
    if (use_gp2(D, 3)) return gproj2_splits(G, R);
    // enough CTAs to fill 148 SMs ~twice, never more than the number of row tiles
    int bm = (D <= 64) ? 128 : 64;
    int ntiles = (R + bm - 1) / bm;
    int want = (296 + G - 1) / G;
    int s = want < ntiles ? want : ntiles;
    return s < 1 ? 1 : s;
}

extern "C" int gptst_gproj_fwd(const float* X, const float* W, const float* bias, const float* Res, float* Y, int G,
                               int R, long group_stride, long row_stride, int D, int act, int prec, void* stream) {
    if (!X || !W || !Y || G <= 0 || R <= 0) return -1;
    cudaStream_t st = (cudaStream_t)stream;
    if (use_gp2(D, prec)) return (int)gproj2_fwd(X, W, bias, Res, Y, G, R, group_stride, row_stride, act, prec, st);
    if (D == 64 && (prec == 1 || prec == 3) && use_umma())
        return (int)gproj_fwd_umma(X, W, bias, Res, Y, G, R, group_stride, row_stride, act, prec, st);
    int splits = gptst_gproj_splits(G, R, D);
#define CALL(DD, PP) return (int)launch_fwd<DD, PP>(X, W, bias, Res, Y, G, R, group_stride, row_stride, act, splits, st)
    DISPATCH_D_PREC(D, prec, CALL);
#undef CALL
    return -2;
}

// Fusion gate of the eval path (reference model/Model.py:12-17) as the epilogue of its second product, D = 64:
//   z = sigmoid(time W + bias + xs),  h = z * flow + (1 - z) * time      (xs = HS_fc(flow), computed by a plain gptst_gproj_fwd)
// W is [in][out] (one group); z (rows, D) is kept for gptst_gate_bwd when non-null.  Other widths: -2 (use gptst_gate_blend).
extern "C" int gptst_gate_fwd(const float* flow, const float* time, const float* xs, const float* W, const float* bias,
                              float* h, float* z, long rows, int D, int prec, void* stream) {
    if (!flow || !time || !xs || !W || !h || rows <= 0) return -1;
    if (rows > 0x7fffffffL || !use_gp2(D, prec)) return -2;
    return (int)gproj2_fwd(time, W, bias, xs, h, 1, (int)rows, 0, D, 2, prec, (cudaStream_t)stream, flow, z);
}

extern "C" int gptst_gproj_bwd(const float* dY, const float* Y, const float* X, const float* W, float* dX, float* dW_part,
                               float* dbias_part, float* dRes, int G, int R, long group_stride, long row_stride, int D,
                               int act, int prec, int splits, void* stream) {
    if (!dY || !X || !W || !dX || !dW_part || !dbias_part || G <= 0 || R <= 0 || splits <= 0) return -1;
    if (act && !Y) return -1;
    cudaStream_t st = (cudaStream_t)stream;
    if (use_gp2(D, prec))
        return (int)gproj2_bwd(dY, Y, X, W, dX, dW_part, dbias_part, dRes, G, R, group_stride, row_stride, act, prec, splits, 0, st);
    if (D == 64 && (prec == 1 || prec == 3) && use_umma_bwd())
        return (int)gproj_bwd_umma(dY, Y, X, W, dX, dW_part, dbias_part, dRes, G, R, group_stride, row_stride, act, prec, splits, st);
#define CALL(DD, PP) \
    return (int)launch_bwd<DD, PP>(dY, Y, X, W, dX, dW_part, dbias_part, dRes, G, R, group_stride, row_stride, act, splits, st)
    DISPATCH_D_PREC(D, prec, CALL);
#undef CALL
    return -2;
}
