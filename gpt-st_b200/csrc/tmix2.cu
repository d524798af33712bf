// Backward of the per-node T x T temporal mix of hyperTem (reference GPTST.py:156-158), fused:
//     dx[b,s,n,:] += sum_t M[n][t][s] dy[b,t,n,:]            (the transposed mix, accumulated into the residual gradient)
//     dM[n][t][s]  = sum_{b,j} dy[b,t,n,j] x[b,s,n,j]        (gradient of the mix matrix)
// One warp per (node, batch range); per (b, n) the warp owns a 12 x 64 tile of dy and of x.  Both products run on the
// tensor cores (mma.sync m16n8k16, fp16-split operands, see mma_f16.cuh) with T = 12 padded to 16:
//     dM tile   : M-dim = t, N-dim = s, K = 64 columns  -> A and B fragments are float2 loads straight from global
//                 memory (rows g / g+8, columns 2t..2t+1), no shared memory at all;
//     mix tile  : M-dim = s, N-dim = 8 columns, K = t   -> A = M_n^T (8 registers per node, loaded once per warp),
//                 B = dy[t = 2t', 2t'+1][column g]: a second, transposed-friendly read of the same tile (L1 hit).
// dy is a gradient: each tile gets one power-of-two scale from its max |.| (exact, undone in fp32).  dM accumulates in
// fp32 registers over the warp's batches; partials per batch split are summed by the caller (deterministic).
// Replaces tmix(transpose, accumulate) + tmix_dM (24 + 50 us at B=64, N=170) with one pass over dy, x and dx.
#include "common.cuh"
#include "mma_f16.cuh"

namespace gptst {
namespace tm2 {

using namespace hf;
constexpr int T = 12, D = 64;

template <int PREC>
__global__ void __launch_bounds__(256, 3)
tmix_bwd_kernel(const float* __restrict__ dy, const float* __restrict__ x, const float* __restrict__ M,
                float* __restrict__ dx_io, float* __restrict__ dM_part, int B, int N, int bps) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int g = lane >> 2, t = lane & 3;
    const int n = blockIdx.x * 8 + warp;
    if (n >= N) return;
    const int b0 = blockIdx.y * bps;
    const int b1 = (b0 + bps < B) ? b0 + bps : B;
    const size_t slab = (size_t)N * D;
    const bool r1ok = g + 8 < T;                 // second fragment row (t or s = g + 8) exists

    // ---- A fragments of the mix: A[m = s][k = tt] = M[n][tt][s], one power-of-two scale per node
    uint32_t mh[4], ml[4];
    float m_inv;
    {
        const float* Mn = M + (size_t)n * T * T;
        float a[8];
#pragma unroll
        for (int i = 0; i < 8; ++i) {
            // i: 0,1 -> (s = g, tt = 2t, 2t+1)  2,3 -> (s = g+8, ..)  4,5 -> (s = g, tt = 2t+8, 2t+9)  6,7 -> (s = g+8, ..)
            const int tt = 2 * t + (i & 1) + ((i & 4) ? 8 : 0);
            const int s = g + ((i & 2) ? 8 : 0);
            a[i] = (tt < T && s < T) ? Mn[tt * T + s] : 0.f;
        }
        float mx = 0.f;
#pragma unroll
        for (int i = 0; i < 8; ++i) mx = fmaxf(mx, fabsf(a[i]));
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, o));
        const float2 sc = pow2_scale_for_fp16(mx);
        m_inv = sc.y;
        split_h2<PREC>(a[0] * sc.x, a[1] * sc.x, mh[0], ml[0]);
        split_h2<PREC>(a[2] * sc.x, a[3] * sc.x, mh[1], ml[1]);
        split_h2<PREC>(a[4] * sc.x, a[5] * sc.x, mh[2], ml[2]);
        split_h2<PREC>(a[6] * sc.x, a[7] * sc.x, mh[3], ml[3]);
    }

    float dm[2][4];                               // dM tile: [s tile][C fragment]
#pragma unroll
    for (int i = 0; i < 4; ++i) dm[0][i] = dm[1][i] = 0.f;

    for (int b = b0; b < b1; ++b) {
        const float* dyp = dy + (size_t)b * T * slab + (size_t)n * D;
        const float* xp = x + (size_t)b * T * slab + (size_t)n * D;
        float* dxp = dx_io + (size_t)b * T * slab + (size_t)n * D;
        // ---- dy tile in the A layout of the dM product: rows g, g+8; columns 16k + 2t (+8)
        float2 ya[4][4];
        float mx = 0.f;
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            const int c = 16 * k + 2 * t;
            ya[k][0] = *reinterpret_cast<const float2*>(dyp + (size_t)g * slab + c);
            ya[k][2] = *reinterpret_cast<const float2*>(dyp + (size_t)g * slab + c + 8);
            ya[k][1] = r1ok ? *reinterpret_cast<const float2*>(dyp + (size_t)(g + 8) * slab + c) : make_float2(0.f, 0.f);
            ya[k][3] = r1ok ? *reinterpret_cast<const float2*>(dyp + (size_t)(g + 8) * slab + c + 8) : make_float2(0.f, 0.f);
#pragma unroll
            for (int i = 0; i < 4; ++i) mx = fmaxf(mx, fmaxf(fabsf(ya[k][i].x), fabsf(ya[k][i].y)));
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, o));
        const float2 sc = pow2_scale_for_fp16(mx);
        // ---- dM tile += dy x^T
        {
            float th[2][4], tl[2][4];
#pragma unroll
            for (int i = 0; i < 4; ++i) th[0][i] = th[1][i] = tl[0][i] = tl[1][i] = 0.f;
#pragma unroll
            for (int k = 0; k < 4; ++k) {
                const int c = 16 * k + 2 * t;
                uint32_t ah[4], al[4];
#pragma unroll
                for (int i = 0; i < 4; ++i) split_h2<PREC>(ya[k][i].x * sc.x, ya[k][i].y * sc.x, ah[i], al[i]);
                // B fragments: x rows s = g (tile 0) and s = g + 8 (tile 1), same columns
                const float2 x00 = *reinterpret_cast<const float2*>(xp + (size_t)g * slab + c);
                const float2 x01 = *reinterpret_cast<const float2*>(xp + (size_t)g * slab + c + 8);
                const float2 x10 = r1ok ? *reinterpret_cast<const float2*>(xp + (size_t)(g + 8) * slab + c) : make_float2(0.f, 0.f);
                const float2 x11 = r1ok ? *reinterpret_cast<const float2*>(xp + (size_t)(g + 8) * slab + c + 8) : make_float2(0.f, 0.f);
                uint32_t bh[4], bl[4];
                split_h2<PREC>(x00.x, x00.y, bh[0], bl[0]);
                split_h2<PREC>(x01.x, x01.y, bh[1], bl[1]);
                split_h2<PREC>(x10.x, x10.y, bh[2], bl[2]);
                split_h2<PREC>(x11.x, x11.y, bh[3], bl[3]);
                if (PREC == PREC_3XTF32) {
                    mma_f16(tl[0], al, bh[0], bh[1]);
                    mma_f16(tl[1], al, bh[2], bh[3]);
                    mma_f16(tl[0], ah, bl[0], bl[1]);
                    mma_f16(tl[1], ah, bl[2], bl[3]);
                }
                mma_f16(th[0], ah, bh[0], bh[1]);
                mma_f16(th[1], ah, bh[2], bh[3]);
            }
#pragma unroll
            for (int i = 0; i < 4; ++i) {
                dm[0][i] = fmaf(th[0][i] + tl[0][i], sc.y, dm[0][i]);
                dm[1][i] = fmaf(th[1][i] + tl[1][i], sc.y, dm[1][i]);
            }
        }
        // ---- dx tile += M^T dy : B[k = tt][n = column] = dy[tt][8j + g]
        const float un = sc.y * m_inv;
#pragma unroll
        for (int j = 0; j < 8; ++j) {
            const int c = 8 * j + g;
            const float v0 = dyp[(size_t)(2 * t) * slab + c], v1 = dyp[(size_t)(2 * t + 1) * slab + c];
            float v2 = 0.f, v3 = 0.f;
            if (2 * t + 8 < T) { v2 = dyp[(size_t)(2 * t + 8) * slab + c]; v3 = dyp[(size_t)(2 * t + 9) * slab + c]; }
            uint32_t bh0, bl0, bh1, bl1;
            split_h2<PREC>(v0 * sc.x, v1 * sc.x, bh0, bl0);
            split_h2<PREC>(v2 * sc.x, v3 * sc.x, bh1, bl1);
            float acc[4] = {0.f, 0.f, 0.f, 0.f};
            mma3<PREC>(acc, mh, ml, bh0, bh1, bl0, bl1);
            float2* p0 = reinterpret_cast<float2*>(dxp + (size_t)g * slab + 8 * j + 2 * t);
            float2 o = *p0;
            o.x = fmaf(acc[0], un, o.x); o.y = fmaf(acc[1], un, o.y);
            *p0 = o;
            if (r1ok) {
                float2* p1 = reinterpret_cast<float2*>(dxp + (size_t)(g + 8) * slab + 8 * j + 2 * t);
                float2 o1 = *p1;
                o1.x = fmaf(acc[2], un, o1.x); o1.y = fmaf(acc[3], un, o1.y);
                *p1 = o1;
            }
        }
    }
    // ---- dM partial of this batch range: C fragment (t = g / g+8 ; s = 8*tile + 2t, 2t+1)
    float* out = dM_part + ((size_t)blockIdx.y * N + n) * T * T;
#pragma unroll
    for (int tile = 0; tile < 2; ++tile) {
        const int s = 8 * tile + 2 * t;
        if (s < T) {
            out[g * T + s] = dm[tile][0];
            out[g * T + s + 1] = dm[tile][1];
            if (r1ok) {
                out[(g + 8) * T + s] = dm[tile][2];
                out[(g + 8) * T + s + 1] = dm[tile][3];
            }
        }
    }
}

}  // namespace tm2
}  // namespace gptst

using namespace gptst;

extern "C" int gptst_tmix_bwd_splits(int B, int N) {
    const int groups = (N + 7) / 8;
    int s = (3 * 148 + groups - 1) / groups;     // ~3 CTAs of 8 warps per SM
    if (s > B) s = B;
    if (s < 1) s = 1;
    const int bps = (B + s - 1) / s;
    return (B + bps - 1) / bps;
}

extern "C" int gptst_tmix_bwd(const float* dy, const float* x, const float* M, float* dx_io, float* dM_part, int B, int T,
                              int N, int D, int prec, int splits, void* stream) {
    if (!dy || !x || !M || !dx_io || !dM_part || B <= 0 || N <= 0 || splits <= 0) return -1;
    if (T != tm2::T || D != tm2::D || (prec != 1 && prec != 3)) return -2;
    const int bps = (B + splits - 1) / splits;
    dim3 grid((N + 7) / 8, splits);
    cudaStream_t st = (cudaStream_t)stream;
    if (prec == 3) tm2::tmix_bwd_kernel<PREC_3XTF32><<<grid, 256, 0, st>>>(dy, x, M, dx_io, dM_part, B, N, bps);
    else tm2::tmix_bwd_kernel<PREC_TF32><<<grid, 256, 0, st>>>(dy, x, M, dx_io, dM_part, B, N, bps);
    return (int)cudaGetLastError();
}
