// Backward of the per-node T x T temporal mix of hyperTem (reference GPTST.py:156-158), fused:
//     dx[b,s,n,:] += sum_t M[n][t][s] dy[b,t,n,:]            (the transposed mix, accumulated into the residual gradient)
//     dM[n][t][s]  = sum_{b,j} dy[b,t,n,j] x[b,s,n,j]        (gradient of the mix matrix)
// One warp per (node, batch range); per (b, n) the warp owns a 12 x 64 tile of dy and of x.  Both products run on the
// tensor cores (mma.sync m16n8k16, fp16-split operands, see mma_f16.cuh) with T = 12 padded to 16:
//     dM tile   : M-dim = t, N-dim = s, K = 64 columns  -> A and B fragments by ldmatrix from the fp32 tiles the warp staged
//                 in shared memory (same k permutation on both operands, see mma_f16.cuh);
//     mix tile  : M-dim = s, N-dim = 8 columns, K = t   -> A = M_n^T (8 registers per node, loaded once per warp),
//                 B = dy[t = 2t', 2t'+1][column g] read from the same shared-memory tile; the update goes back through
//                 shared memory so that global memory only ever sees 16-byte chunks of full rows.
// dy is a gradient: each tile gets one power-of-two scale from its max |.| (exact, undone in fp32).  dM accumulates in
// fp32 registers over the warp's batches; partials per batch split are summed by the caller (deterministic).
// Replaces tmix(transpose, accumulate) + tmix_dM (24 + 50 us at B=64, N=170) with one pass over dy, x and dx.
#include "common.cuh"
#include "mma_f16.cuh"

namespace gptst {
namespace tm2 {

using namespace hf;
constexpr int T = 12, D = 64;
#ifndef TMIX_MINB
#define TMIX_MINB 3
#endif

// Shared-memory layout per warp: two 16-row tiles (dy, x) of 272-byte slots; rows 12..15 stay zero.  Rows are staged with
// 16-byte cp.async chunks (two full 256-byte rows per warp request: 4 cache lines per request instead of the 8 that a
// fragment-shaped global access touches -- the L1 tag stage, not DRAM, limited the first version of this kernel).
constexpr int TROWS = 16;
constexpr int WARP_BYTES = 2 * TROWS * ROWB;

__device__ __forceinline__ void stage_tile(unsigned char* tile, const float* base, size_t slab, int lane) {
#pragma unroll
    for (int k = 0; k < 6; ++k) {                  // 12 rows x 16 chunks = 192 chunks
        const int id = lane + 32 * k, r = id >> 4, ch = id & 15;
        cp_async16(tile + (size_t)r * ROWB + ch * 16, base + (size_t)r * slab + ch * 4);
    }
}

// DM_ONLY: only the dM tile (the side-stream half of the fused hyperTem backward, csrc/htem_fused.cu); dx_io / M unused
template <int PREC, bool DM_ONLY>
__global__ void __launch_bounds__(256, TMIX_MINB)
tmix_bwd_kernel(const float* __restrict__ dy, const float* __restrict__ x, const float* __restrict__ M,
                float* __restrict__ dx_io, float* __restrict__ dM_part, int B, int N, int bps) {
    extern __shared__ __align__(128) unsigned char smraw[];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int g = lane >> 2, t = lane & 3;
    unsigned char* Ty = smraw + (size_t)warp * WARP_BYTES;      // dy tile (fp32 rows)
    unsigned char* Tx = Ty + TROWS * ROWB;                      // x tile, later the staging area of the dx update
    // zero the padding rows 12..15 of both tiles once
    for (int i = lane; i < 2 * 4 * (ROWB / 16); i += 32) {
        const int tile = i / (4 * (ROWB / 16)), rem = i % (4 * (ROWB / 16));
        *reinterpret_cast<float4*>((tile ? Tx : Ty) + (size_t)(T + rem / (ROWB / 16)) * ROWB + (rem % (ROWB / 16)) * 16) =
            make_float4(0.f, 0.f, 0.f, 0.f);
    }
    const int n = blockIdx.x * 8 + warp;
    if (n >= N) return;
    const int b0 = blockIdx.y * bps;
    const int b1 = (b0 + bps < B) ? b0 + bps : B;
    const size_t slab = (size_t)N * D;
    const bool r1ok = g + 8 < T;

    // ---- A fragments of the mix: A[m = s][k = tt] = M[n][tt][s], one power-of-two scale per node
    uint32_t mh[4] = {0u, 0u, 0u, 0u}, ml[4] = {0u, 0u, 0u, 0u};
    float m_inv = 1.f;
    if (!DM_ONLY) {
        const float* Mn = M + (size_t)n * T * T;
        float a[8];
#pragma unroll
        for (int i = 0; i < 8; ++i) {
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

    float dm[2][4];
#pragma unroll
    for (int i = 0; i < 4; ++i) dm[0][i] = dm[1][i] = 0.f;

    for (int b = b0; b < b1; ++b) {
        const float* dyp = dy + (size_t)b * T * slab + (size_t)n * D;
        const float* xp = x + (size_t)b * T * slab + (size_t)n * D;
        float* dxp = dx_io + (size_t)b * T * slab + (size_t)n * D;
        __syncwarp();                               // the previous iteration's readers of Ty / Tx are done
        stage_tile(Ty, dyp, slab, lane);
        stage_tile(Tx, xp, slab, lane);
        cp_async_commit();
        // the dx tile this warp will update: same chunk mapping, straight to registers (overlaps the staging)
        float4 dxv[6];
        if (!DM_ONLY) {
#pragma unroll
            for (int k = 0; k < 6; ++k) {
                const int id = lane + 32 * k, r = id >> 4, ch = id & 15;
                dxv[k] = *reinterpret_cast<const float4*>(dxp + (size_t)r * slab + ch * 4);
            }
        }
        cp_async_wait_group<0>();
        __syncwarp();
        // ---- tile max of dy -> power-of-two scale
        float mx = 0.f;
#pragma unroll
        for (int k = 0; k < 6; ++k) {
            const int id = lane + 32 * k, r = id >> 4, ch = id & 15;
            const float4 v = *reinterpret_cast<const float4*>(Ty + (size_t)r * ROWB + ch * 16);
            mx = fmaxf(mx, fmaxf(fmaxf(fabsf(v.x), fabsf(v.y)), fmaxf(fabsf(v.z), fabsf(v.w))));
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, o));
        const float2 sc = pow2_scale_for_fp16(mx);
        // x is an activation of any magnitude: its tile gets a power-of-two scale of its own (exact, undone below)
        float mxx = 0.f;
#pragma unroll
        for (int k = 0; k < 6; ++k) {
            const int id = lane + 32 * k, r = id >> 4, ch = id & 15;
            const float4 v = *reinterpret_cast<const float4*>(Tx + (size_t)r * ROWB + ch * 16);
            mxx = fmaxf(mxx, fmaxf(fmaxf(fabsf(v.x), fabsf(v.y)), fmaxf(fabsf(v.z), fabsf(v.w))));
        }
        mxx = __uint_as_float(__reduce_max_sync(0xffffffffu, __float_as_uint(mxx)));
        const float2 sx = pow2_scale_for_fp16(mxx);
        // ---- dM tile += dy x^T   (both operands by ldmatrix from the fp32 tiles: same k permutation on both sides)
        {
            float th[2][4], tl[2][4];
#pragma unroll
            for (int i = 0; i < 4; ++i) th[0][i] = th[1][i] = tl[0][i] = tl[1][i] = 0.f;
#pragma unroll
            for (int k = 0; k < 4; ++k) {
                uint32_t f[4], ah[4], al[4], bh[4], bl[4];
                const uint32_t aaddr = smem_u32(Ty + (size_t)(lane & 7) * ROWB + (16 * k + 4 * (lane >> 3)) * 4);
                const uint32_t baddr = smem_u32(Tx + (size_t)(lane & 7) * ROWB + (16 * k + 4 * (lane >> 3)) * 4);
                ldsm_x4(f, aaddr);
                split_h2<PREC>(__uint_as_float(f[0]) * sc.x, __uint_as_float(f[1]) * sc.x, ah[0], al[0]);
                split_h2<PREC>(__uint_as_float(f[2]) * sc.x, __uint_as_float(f[3]) * sc.x, ah[2], al[2]);
                ldsm_x4(f, aaddr + 8 * ROWB);
                split_h2<PREC>(__uint_as_float(f[0]) * sc.x, __uint_as_float(f[1]) * sc.x, ah[1], al[1]);
                split_h2<PREC>(__uint_as_float(f[2]) * sc.x, __uint_as_float(f[3]) * sc.x, ah[3], al[3]);
                ldsm_x4(f, baddr);                  // x rows s = 0..7  -> B fragments of tile 0
                split_h2<PREC>(__uint_as_float(f[0]) * sx.x, __uint_as_float(f[1]) * sx.x, bh[0], bl[0]);
                split_h2<PREC>(__uint_as_float(f[2]) * sx.x, __uint_as_float(f[3]) * sx.x, bh[1], bl[1]);
                ldsm_x4(f, baddr + 8 * ROWB);       // x rows s = 8..15 -> tile 1
                split_h2<PREC>(__uint_as_float(f[0]) * sx.x, __uint_as_float(f[1]) * sx.x, bh[2], bl[2]);
                split_h2<PREC>(__uint_as_float(f[2]) * sx.x, __uint_as_float(f[3]) * sx.x, bh[3], bl[3]);
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
                dm[0][i] = fmaf(th[0][i] + tl[0][i], sc.y * sx.y, dm[0][i]);
                dm[1][i] = fmaf(th[1][i] + tl[1][i], sc.y * sx.y, dm[1][i]);
            }
        }
        __syncwarp();                               // x tile consumed: its slots become the staging area of the update
        if (DM_ONLY) continue;
        // ---- update tile = M^T dy : B[k = tt][n = column] = dy[tt][8j + g]  (conflict-free scalar reads of the fp32 tile)
        const float un = sc.y * m_inv;
#pragma unroll
        for (int j = 0; j < 8; ++j) {
            const int c = 8 * j + g;
            const float v0 = *reinterpret_cast<const float*>(Ty + (size_t)(2 * t) * ROWB + c * 4);
            const float v1 = *reinterpret_cast<const float*>(Ty + (size_t)(2 * t + 1) * ROWB + c * 4);
            const float v2 = *reinterpret_cast<const float*>(Ty + (size_t)(2 * t + 8) * ROWB + c * 4);   // rows >= 12 are zero
            const float v3 = *reinterpret_cast<const float*>(Ty + (size_t)(2 * t + 9) * ROWB + c * 4);
            uint32_t bh0, bl0, bh1, bl1;
            split_h2<PREC>(v0 * sc.x, v1 * sc.x, bh0, bl0);
            split_h2<PREC>(v2 * sc.x, v3 * sc.x, bh1, bl1);
            float acc[4] = {0.f, 0.f, 0.f, 0.f};
            mma3<PREC>(acc, mh, ml, bh0, bh1, bl0, bl1);
            *reinterpret_cast<float2*>(Tx + (size_t)g * ROWB + (8 * j + 2 * t) * 4) = make_float2(acc[0] * un, acc[1] * un);
            if (r1ok) *reinterpret_cast<float2*>(Tx + (size_t)(g + 8) * ROWB + (8 * j + 2 * t) * 4) = make_float2(acc[2] * un, acc[3] * un);
        }
        __syncwarp();
#pragma unroll
        for (int k = 0; k < 6; ++k) {
            const int id = lane + 32 * k, r = id >> 4, ch = id & 15;
            const float4 u = *reinterpret_cast<const float4*>(Tx + (size_t)r * ROWB + ch * 16);
            float4 o = dxv[k];
            o.x += u.x; o.y += u.y; o.z += u.z; o.w += u.w;
            *reinterpret_cast<float4*>(dxp + (size_t)r * slab + ch * 4) = o;
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
    const size_t smem = (size_t)8 * tm2::WARP_BYTES;
    cudaError_t e;
    if (prec == 3) {
        e = cudaFuncSetAttribute(tm2::tmix_bwd_kernel<PREC_3XTF32, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) return (int)e;
        tm2::tmix_bwd_kernel<PREC_3XTF32, false><<<grid, 256, smem, st>>>(dy, x, M, dx_io, dM_part, B, N, bps);
    } else {
        e = cudaFuncSetAttribute(tm2::tmix_bwd_kernel<PREC_TF32, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) return (int)e;
        tm2::tmix_bwd_kernel<PREC_TF32, false><<<grid, 256, smem, st>>>(dy, x, M, dx_io, dM_part, B, N, bps);
    }
    return (int)cudaGetLastError();
}

// dM only (side-stream half of the fused hyperTem backward): dM_part[split][n][t][s] = sum over the split's batches and the
// 64 columns of dret[b,t,n,:] x[b,s,n,:].  splits = gptst_tmix_bwd_splits(B, N).
extern "C" int gptst_tmix_dM2(const float* dret, const float* x, float* dM_part, int B, int T, int N, int D, int splits, void* stream) {
    if (!dret || !x || !dM_part || B <= 0 || N <= 0 || splits <= 0) return -1;
    if (T != tm2::T || D != tm2::D) return -2;
    const int bps = (B + splits - 1) / splits;
    dim3 grid((N + 7) / 8, splits);
    const size_t smem = (size_t)8 * tm2::WARP_BYTES;
    cudaError_t e = cudaFuncSetAttribute(tm2::tmix_bwd_kernel<PREC_3XTF32, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return (int)e;
    tm2::tmix_bwd_kernel<PREC_3XTF32, true><<<grid, 256, smem, (cudaStream_t)stream>>>(dret, x, nullptr, nullptr, dM_part, B, N, bps);
    return (int)cudaGetLastError();
}
