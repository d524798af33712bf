// Helpers shared by the cap kernels (see cap_route_fwd.cu / cap_route_bwd.cu / cap_small.cu).
#pragma once
#include <cooperative_groups.h>

#include "common.cuh"

namespace cg = cooperative_groups;

namespace gptst {

constexpr int kWarps = 8;
constexpr int kLPN = 4;               // lanes per node in the dot-product phases
constexpr int kNPB = 32 / kLPN;       // nodes per warp batch
constexpr int kCW = 20;               // padded row length of the per-warp c staging (>= H+1, multiple of 4)

template <int K, int NT, int PREC, bool A_TRANS, bool B_KMAJOR>
__device__ __forceinline__ void warp_gemm_rt(float (&acc)[NT][4], const float* __restrict__ As, int lda,
                                             const float* __restrict__ Bs, int ldb, int lane, int Krt) {
    const int g = lane >> 2, t = lane & 3;
    for (int k0 = 0; k0 < Krt; k0 += 8) {
        float af[4];
        if (!A_TRANS) {
            af[0] = As[g * lda + k0 + t];
            af[1] = As[(g + 8) * lda + k0 + t];
            af[2] = As[g * lda + k0 + t + 4];
            af[3] = As[(g + 8) * lda + k0 + t + 4];
        } else {
            af[0] = As[(k0 + t) * lda + g];
            af[1] = As[(k0 + t) * lda + g + 8];
            af[2] = As[(k0 + t + 4) * lda + g];
            af[3] = As[(k0 + t + 4) * lda + g + 8];
        }
        uint32_t ah[4], al[4];
#pragma unroll
        for (int i = 0; i < 4; ++i) split_tf32<PREC>(af[i], ah[i], al[i]);
#pragma unroll
        for (int nt = 0; nt < NT; ++nt) {
            float bf0, bf1;
            if (!B_KMAJOR) {
                bf0 = Bs[(nt * 8 + g) * ldb + k0 + t];
                bf1 = Bs[(nt * 8 + g) * ldb + k0 + t + 4];
            } else {
                bf0 = Bs[(k0 + t) * ldb + nt * 8 + g];
                bf1 = Bs[(k0 + t + 4) * ldb + nt * 8 + g];
            }
            uint32_t bh[2], bl[2];
            split_tf32<PREC>(bf0, bh[0], bl[0]);
            split_tf32<PREC>(bf1, bh[1], bl[1]);
            mma_split<PREC>(acc[nt], ah, al, bh, bl);
        }
    }
}

// Z = x Wp^T + bp for the 16-row tile `mt` (in place in Xs when Zs == Xs); returns per-row |Z|^2 for this
// lane's two rows (g and g+8), already reduced over the quad.
template <int D, int PREC>
__device__ __forceinline__ void ztile(const float* __restrict__ Xs, int ldx, const float* __restrict__ Wps, int ldw,
                                      const float* __restrict__ bps, int mt, int lane, float (&acc)[D / 8][4],
                                      float& q0, float& q1) {
#pragma unroll
    for (int nt = 0; nt < D / 8; ++nt) acc[nt][0] = acc[nt][1] = acc[nt][2] = acc[nt][3] = 0.f;
    warp_gemm<D, D / 8, PREC, false, false>(acc, Xs + mt * 16 * ldx, ldx, Wps, ldw, lane);
    const int tq = lane & 3;
    q0 = q1 = 0.f;
#pragma unroll
    for (int nt = 0; nt < D / 8; ++nt) {
        const float b0 = bps[nt * 8 + 2 * tq], b1 = bps[nt * 8 + 2 * tq + 1];
        acc[nt][0] += b0; acc[nt][1] += b1; acc[nt][2] += b0; acc[nt][3] += b1;
        q0 += acc[nt][0] * acc[nt][0] + acc[nt][1] * acc[nt][1];
        q1 += acc[nt][2] * acc[nt][2] + acc[nt][3] * acc[nt][3];
    }
    q0 += __shfl_xor_sync(0xffffffffu, q0, 1); q0 += __shfl_xor_sync(0xffffffffu, q0, 2);
    q1 += __shfl_xor_sync(0xffffffffu, q1, 1); q1 += __shfl_xor_sync(0xffffffffu, q1, 2);
}

// dots[h] = sum_d V[h][d] * Row[d] for one node handled by kLPN lanes (lane part `q` takes float4 index j = q mod kLPN)
template <int D, int HP>
__device__ __forceinline__ void node_dots(const float* __restrict__ row, const float* __restrict__ V, int H, int q,
                                          float (&dots)[HP]) {
#pragma unroll
    for (int h = 0; h < HP; ++h) dots[h] = 0.f;
#pragma unroll
    for (int jj = 0; jj < D / 4 / kLPN; ++jj) {
        const int j = jj * kLPN + q;
        const float4 p = *reinterpret_cast<const float4*>(row + 4 * j);
#pragma unroll
        for (int h = 0; h < HP; ++h) {
            if (h < H) {
                const float4 v = *reinterpret_cast<const float4*>(V + h * D + 4 * j);
                dots[h] = fmaf(p.x, v.x, fmaf(p.y, v.y, fmaf(p.z, v.z, fmaf(p.w, v.w, dots[h]))));
            }
        }
    }
#pragma unroll
    for (int h = 0; h < HP; ++h) {
        dots[h] += __shfl_xor_sync(0xffffffffu, dots[h], 1);
        dots[h] += __shfl_xor_sync(0xffffffffu, dots[h], 2);
    }
}

template <int HP>
__device__ __forceinline__ void softmax_h(float (&z)[HP], int H) {
    float m = -INFINITY;
#pragma unroll
    for (int h = 0; h < HP; ++h) if (h < H) m = fmaxf(m, z[h]);
    float s = 0.f;
#pragma unroll
    for (int h = 0; h < HP; ++h) {
        if (h < H) { z[h] = expf(z[h] - m); s += z[h]; } else z[h] = 0.f;
    }
    const float inv = 1.f / s;
#pragma unroll
    for (int h = 0; h < HP; ++h) z[h] *= inv;
}

// row-wise squash of an (rows x D) smem matrix in place, one warp per row
template <int D>
__device__ __forceinline__ void squash_rows(float* M, int rows, int warp, int lane) {
    for (int r = warp; r < rows; r += kWarps) {
        float q = 0.f;
        for (int d = lane; d < D; d += 32) { float v = M[r * D + d]; q += v * v; }
        q = warp_sum(q);
        const float f = squash_f(q);
        for (int d = lane; d < D; d += 32) M[r * D + d] *= f;
    }
}

constexpr size_t kSmemMax = 227 * 1024;

}  // namespace gptst
