// fp16-split tensor-core building blocks (mma.sync m16n8k16 + ldmatrix) shared by the second-generation cap kernels.
//
// A value a is carried as a_hi + a_lo (two fp16 numbers, 22 significant bits; fp16 keeps subnormals, so for |a| <= 1 the
// absolute error is <= 2^-25).  Products use three terms with fp32 accumulation: a.b ~= a_hi.b_hi + a_hi.b_lo + a_lo.b_hi.
// Operand rows live in shared memory as  [0,128) hi plane (64 halves) | [128,256) lo plane | 16 B pad  = 272 B per row:
// eight consecutive rows hit eight distinct 16-byte bank groups, so every ldmatrix is conflict-free.
#pragma once
#include <cuda_fp16.h>

#include "common.cuh"

namespace gptst {
namespace hf {

constexpr int ROWB = 272;     // bytes per shared-memory operand row
constexpr int LO = 128;       // byte offset of the lo plane inside a row

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void ldsm_x4(uint32_t (&r)[4], uint32_t addr) {
    asm volatile("ldmatrix.sync.aligned.m8n8.x4.shared.b16 {%0,%1,%2,%3}, [%4];"
                 : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]) : "r"(addr) : "memory");
}
__device__ __forceinline__ void ldsm_x4_t(uint32_t (&r)[4], uint32_t addr) {
    asm volatile("ldmatrix.sync.aligned.m8n8.x4.trans.shared.b16 {%0,%1,%2,%3}, [%4];"
                 : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]) : "r"(addr) : "memory");
}
// D(16x8, f32) += A(16x16, f16, row) * B(16x8, f16, col)
//   A regs: 0=(row g, k 2t..2t+1) 1=(row g+8, same k) 2=(row g, k 2t+8..2t+9) 3=(row g+8, k 2t+8..)     g = lane>>2, t = lane&3
//   B regs: b0=(k 2t..2t+1, n g)  b1=(k 2t+8..2t+9, n g)          C: 0=(g,2t) 1=(g,2t+1) 2=(g+8,2t) 3=(g+8,2t+1)
__device__ __forceinline__ void mma_f16(float (&c)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1) {
    asm volatile(
        "mma.sync.aligned.m16n8k16.row.col.f32.f16.f16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
        : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3])
        : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}
template <int PREC>
__device__ __forceinline__ void mma3(float (&c)[4], const uint32_t (&ah)[4], const uint32_t (&al)[4], uint32_t bh0,
                                     uint32_t bh1, uint32_t bl0, uint32_t bl1) {
    if (PREC == PREC_3XTF32) {
        mma_f16(c, al, bh0, bh1);
        mma_f16(c, ah, bl0, bl1);
    }
    mma_f16(c, ah, bh0, bh1);
}
// (a, b) -> packed fp16 pair hi and the packed residual lo
template <int PREC>
__device__ __forceinline__ void split_h2(float a, float b, uint32_t& hi, uint32_t& lo) {
    const __half2 h = __floats2half2_rn(a, b);
    hi = *reinterpret_cast<const uint32_t*>(&h);
    if (PREC == PREC_3XTF32) {
        const float2 f = __half22float2(h);
        const __half2 l = __floats2half2_rn(a - f.x, b - f.y);
        lo = *reinterpret_cast<const uint32_t*>(&l);
    } else {
        lo = 0u;
    }
}
__device__ __forceinline__ void cp_async16(void* dst, const void* src) {
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(smem_u32(dst)), "l"(src) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_async_wait_group() { asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory"); }
__device__ __forceinline__ void cp_async_wait_all() {
    asm volatile("cp.async.commit_group;\n\tcp.async.wait_group 0;" ::: "memory");
}



constexpr int REDLD = 72;     // floats per row of a cross-warp partial buffer (conflict-free float2 stores)

// Logit product of one warp:  z[(h, node)] += scale * sum_d V[h][d] * G[node][d]  for the warp's 16 rows n0..n0+15 of G.
//   V : 16 operand rows (hi|lo planes, rows >= H zero)  -> A fragments (M = 16 hyperedges)
//   G : operand rows of the nodes                       -> B fragments (N = 2 x 8 nodes), K = D = 64
// Result layout (= C fragment of the two node tiles = A fragment of warp_aggregate):
//   z[0]=(h0,na) z[1]=(h0,na+1) z[2]=(h1,na) z[3]=(h1,na+1) z[4]=(h0,nb) z[5]=(h0,nb+1) z[6]=(h1,nb) z[7]=(h1,nb+1)
//   h0 = lane>>2, h1 = h0+8, na = n0 + 2*(lane&3), nb = na+8.   Two accumulators per tile (hi.hi and the two small terms).
template <int PREC>
__device__ __forceinline__ void warp_logits(float (&z)[8], const unsigned char* vpl, const unsigned char* rows, int n0, int lane,
                                            float scale) {
    float zh[2][4], zl[2][4];
#pragma unroll
    for (int i = 0; i < 4; ++i) zh[0][i] = zh[1][i] = zl[0][i] = zl[1][i] = 0.f;
#pragma unroll
    for (int b = 0; b < 4; ++b) {
        uint32_t vh[4], vl[4] = {0u, 0u, 0u, 0u}, ph[4], pl[4] = {0u, 0u, 0u, 0u};
        const uint32_t aaddr = smem_u32(vpl + (size_t)(8 * ((lane >> 3) & 1) + (lane & 7)) * ROWB + (16 * b + 8 * (lane >> 4)) * 2);
        const uint32_t baddr =
            smem_u32(rows + (size_t)(n0 + 8 * (lane >> 4) + (lane & 7)) * ROWB + (16 * b + 8 * ((lane >> 3) & 1)) * 2);
        ldsm_x4(vh, aaddr);
        ldsm_x4(ph, baddr);
        if (PREC == PREC_3XTF32) {
            ldsm_x4(vl, aaddr + LO);
            ldsm_x4(pl, baddr + LO);
            mma_f16(zl[0], vl, ph[0], ph[1]);
            mma_f16(zl[1], vl, ph[2], ph[3]);
            mma_f16(zl[0], vh, pl[0], pl[1]);
            mma_f16(zl[1], vh, pl[2], pl[3]);
        }
        mma_f16(zh[0], vh, ph[0], ph[1]);
        mma_f16(zh[1], vh, ph[2], ph[3]);
    }
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        z[i] += (zh[0][i] + zl[0][i]) * scale;
        z[4 + i] += (zh[1][i] + zl[1][i]) * scale;
    }
}

// Aggregation of one warp:  red_w[h][:] = scale * sum over the warp's 16 nodes of c[h][n] G[n][:]   (rows h < HA).
// c comes in the layout warp_logits produces; G rows are fetched transposed (ldmatrix.trans), K = 16 nodes.
template <int PREC>
__device__ __forceinline__ void warp_aggregate(const float (&c)[8], int HA, const unsigned char* rows, int n0, float* red_w,
                                               int lane, float scale) {
    const int g = lane >> 2, t = lane & 3;
    const int h0 = g, h1 = g + 8;
    uint32_t ah[4], al[4];
    split_h2<PREC>(c[0], c[1], ah[0], al[0]);
    split_h2<PREC>(c[2], c[3], ah[1], al[1]);
    split_h2<PREC>(c[4], c[5], ah[2], al[2]);
    split_h2<PREC>(c[6], c[7], ah[3], al[3]);
    float* r0 = red_w + (size_t)h0 * REDLD + 2 * t;
    float* r1 = red_w + (size_t)h1 * REDLD + 2 * t;
#pragma unroll
    for (int jp = 0; jp < 4; ++jp) {
        uint32_t bh[4], bq[4] = {0u, 0u, 0u, 0u};
        const uint32_t baddr =
            smem_u32(rows + (size_t)(n0 + 8 * ((lane >> 3) & 1) + (lane & 7)) * ROWB + (16 * jp + 8 * (lane >> 4)) * 2);
        ldsm_x4_t(bh, baddr);
        if (PREC == PREC_3XTF32) ldsm_x4_t(bq, baddr + LO);
        float a0[4] = {0.f, 0.f, 0.f, 0.f}, a1[4] = {0.f, 0.f, 0.f, 0.f};
        mma3<PREC>(a0, ah, al, bh[0], bh[1], bq[0], bq[1]);
        mma3<PREC>(a1, ah, al, bh[2], bh[3], bq[2], bq[3]);
        if (h0 < HA) {
            *reinterpret_cast<float2*>(r0 + 16 * jp) = make_float2(a0[0] * scale, a0[1] * scale);
            *reinterpret_cast<float2*>(r0 + 16 * jp + 8) = make_float2(a1[0] * scale, a1[1] * scale);
        }
        if (h1 < HA) {
            *reinterpret_cast<float2*>(r1 + 16 * jp) = make_float2(a0[2] * scale, a0[3] * scale);
            *reinterpret_cast<float2*>(r1 + 16 * jp + 8) = make_float2(a1[2] * scale, a1[3] * scale);
        }
    }
}

// power-of-two scale that brings max|.| = m into [2^13, 2^14) (fp16-safe); returns (scale, 1/scale), (1,1) for m == 0
__device__ __forceinline__ float2 pow2_scale_for_fp16(float m) {
    if (!(m > 0.f)) return make_float2(1.f, 1.f);
    int sh = 13 - (((__float_as_int(m) >> 23) & 0xff) - 127);
    sh = sh > 100 ? 100 : (sh < -100 ? -100 : sh);
    return make_float2(__int_as_float((127 + sh) << 23), __int_as_float((127 - sh) << 23));
}


// ---- x W^T for a 16-row tile with the A operand fetched by ldmatrix straight from fp32 rows ------------------------
// Matrix i of an x4 load is the 8 x 4-float block at columns 16b+4i, so lane (g,t) receives x[g][16b+4i+t]: the MMA's
// logical k index is a permutation of the physical one inside every 16-block,
//     logical {2t, 2t+1, 2t+8, 2t+9}  <->  physical {t, 4+t, 8+t, 12+t},
// and the weight is staged with the same permutation, so the contraction is unchanged.
// stage_w_perm: W (64 x 64 fp32, [out][in] as nn.Linear stores it) * scale -> Wt rows = out, hi|lo planes along permuted in.
template <int PREC, int NT>
__device__ __forceinline__ void stage_w_perm(unsigned char* Wt, const float* __restrict__ W, float scale, int tid) {
    constexpr int WI = (64 * 16 + NT - 1) / NT;
    float4 wv[WI];
#pragma unroll
    for (int u = 0; u < WI; ++u) {               // all loads first: a rolled loop would serialise the global latencies
        const int i = tid + u * NT;
        wv[u] = (i < 64 * 16) ? *reinterpret_cast<const float4*>(W + (size_t)(i >> 4) * 64 + (i & 15) * 4) : make_float4(0.f, 0.f, 0.f, 0.f);
    }
#pragma unroll
    for (int u = 0; u < WI; ++u) {
        const int i = tid + u * NT;
        if (i < 64 * 16) {
            const int o = i >> 4, q4 = i & 15;
            const int q = q4 & 3;
            const int base = 16 * (q4 >> 2) + ((q >= 2) ? 8 : 0) + (q & 1);
            const float wq[4] = {wv[u].x * scale, wv[u].y * scale, wv[u].z * scale, wv[u].w * scale};
            __half* hrow = reinterpret_cast<__half*>(Wt + (size_t)o * ROWB);
            __half* lrow = reinterpret_cast<__half*>(Wt + (size_t)o * ROWB + LO);
#pragma unroll
            for (int e = 0; e < 4; ++e) {
                const __half h = __float2half_rn(wq[e]);
                hrow[base + 2 * e] = h;
                lrow[base + 2 * e] = (PREC == PREC_3XTF32) ? __float2half_rn(wq[e] - __half2float(h)) : __float2half_rn(0.f);
            }
        }
    }
}
// physical column of the permuted position p (inverse of the staging permutation)
__device__ __forceinline__ int perm_pos_to_col(int p) {
    const int b = p & ~15, r = p & 15;
    return b + ((r & 8) ? 8 : 0) + ((r & 1) ? 4 : 0) + ((r & 7) >> 1);
}
// acc[j] (j = 8 column tiles of 8) = rows[n0..n0+15] (fp32) * Wt^T   (scaled by the staging scale)
template <int PREC>
__device__ __forceinline__ void warp_xw_tile(float (&acc)[8][4], const unsigned char* rows, const unsigned char* Wt, int n0, int lane) {
#pragma unroll
    for (int j = 0; j < 8; ++j) acc[j][0] = acc[j][1] = acc[j][2] = acc[j][3] = 0.f;
#pragma unroll
    for (int b = 0; b < 4; ++b) {
        uint32_t f[4], ah[4], al[4];
        const uint32_t aaddr = smem_u32(rows + (size_t)(n0 + (lane & 7)) * ROWB + (16 * b + 4 * (lane >> 3)) * 4);
        ldsm_x4(f, aaddr);                       // rows n0..n0+7
        split_h2<PREC>(__uint_as_float(f[0]), __uint_as_float(f[1]), ah[0], al[0]);
        split_h2<PREC>(__uint_as_float(f[2]), __uint_as_float(f[3]), ah[2], al[2]);
        ldsm_x4(f, aaddr + 8 * ROWB);            // rows n0+8..n0+15
        split_h2<PREC>(__uint_as_float(f[0]), __uint_as_float(f[1]), ah[1], al[1]);
        split_h2<PREC>(__uint_as_float(f[2]), __uint_as_float(f[3]), ah[3], al[3]);
#pragma unroll
        for (int jp = 0; jp < 4; ++jp) {
            uint32_t bh[4], bl[4] = {0u, 0u, 0u, 0u};
            const uint32_t baddr =
                smem_u32(Wt + (size_t)(16 * jp + 8 * (lane >> 4) + (lane & 7)) * ROWB + (16 * b + 8 * ((lane >> 3) & 1)) * 2);
            ldsm_x4(bh, baddr);
            if (PREC == PREC_3XTF32) ldsm_x4(bl, baddr + LO);
            mma3<PREC>(acc[2 * jp], ah, al, bh[0], bh[1], bl[0], bl[1]);
            mma3<PREC>(acc[2 * jp + 1], ah, al, bh[2], bh[3], bl[2], bl[3]);
        }
    }
}

}  // namespace hf
}  // namespace gptst
