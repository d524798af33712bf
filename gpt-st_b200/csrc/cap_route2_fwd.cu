// cap: intra-cluster routing forward (reference GPTST.py:102-123), second generation, D = 64 and N <= 256.
//
// One CTA per (b,t) slab, one warp per 16 nodes; a warp only ever touches ITS OWN 16 rows of P, so the only
// cross-warp traffic is the (H+1) x D partial of each aggregation.  Everything that is a contraction runs on the
// tensor cores as a three-term fp16 split with fp32 accumulation (mma.sync m16n8k16, a = a_hi + a_lo):
//     a.b ~= a_hi.b_hi + a_hi.b_lo + a_lo.b_hi        (fp16 keeps 11 significant bits, the pair 22)
// P and c are bounded by 1 (squash / softmax outputs) and fp16 keeps subnormals, so hi+lo carries an absolute
// error <= 2^-25 there; x is split as it is (|x| must stay below 65504 -- an overflow turns the row into NaN, it
// does not go unnoticed) and Wp is pre-scaled by 64 so its low part stays in fp16's normal range.
// Compared with the tf32 generation (cap_route_fwd.cu, still used for D = 128 / N > 256) this halves the number
// of tensor instructions, needs no per-operand split on the hot loops (P, v, Wp are stored pre-split, operands are
// fetched with ldmatrix) and halves the shared-memory footprint of P (2 x 2 B per element).
//
//   stage   : x rows -> smem (cp.async, fp32, in the slot that later holds the row's P_hi|P_lo), Wp -> fp16 hi/lo
//   Z       : warp tile 16 nodes x 64: Z = x Wp^T + bp ; P = squash(Z) -> fp16 hi/lo planes (in place)
//   pass A  : c0 = softmax_H(dadj) ; [c0 ; 1] P  -> u = squash(c0 P), sumP          (first routing iteration: c = 1/H)
//   pass k  : b += v P^T ; c = softmax_H(b) ; v' = squash(u * (c P))                (R-1 times)
//   final   : b += v P^T ; c = softmax_H(b + dadj) -> c_out ; s = c P -> s_out
// The logit MMA (M = 16 hyperedges, N = 8 nodes, K = D) leaves its result in exactly the A-fragment layout of the
// aggregation MMA (M = 16 hyperedges, N = 8 columns of D, K = 16 nodes), so softmax outputs never leave registers.
#include <cstdlib>

#include "cap_common.cuh"
#include "mma_f16.cuh"

namespace gptst {
namespace r2 {

using namespace hf;
constexpr int D = 64;
constexpr float WSCALE = 64.f;
constexpr float LOG2E = 1.4426950408889634f;

// A warp owns TPW consecutive 16-node tiles (TPW = 1 is what runs).  ncu (round 2, profiles/ncu_route_fwd_r02.md): the kernel
// is bound by SHARED-MEMORY WAVEFRONTS (6.66 M per launch = 788 per tile; the LSU data pipe is 67 % busy while an SM is active)
// together with a 51 % issue rate at 31 % active warps: every MMA operand is an ldmatrix (W_p 128 wavefronts per tile, V and P
// 96 per routing sweep), plus the cross-warp partials (52).  SMs are active only 75 % of the launch (2.59 waves of slab CTAs).
__host__ __device__ inline size_t wred_bytes(int NW, int H) {
    const size_t w = (size_t)D * ROWB, r = (size_t)NW * (H + 1) * REDLD * 4;
    return r > w ? r : w;
}
__host__ __device__ inline size_t smem_bytes(int NW, int ntiles, int H) {
    return (size_t)ntiles * 16 * ROWB + wred_bytes(NW, H) + 16 * ROWB + (size_t)H * D * 4 + D * 4;
}

__device__ __forceinline__ float ex2_approx(float x) {
    float y;
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
    return y;
}

// softmax over the hyperedge index of the 4 node columns this lane holds.
//   z[0]=(h0,na) z[1]=(h0,na+1) z[2]=(h1,na) z[3]=(h1,na+1) z[4]=(h0,nb) z[5]=(h0,nb+1) z[6]=(h1,nb) z[7]=(h1,nb+1)
// h0 = g, h1 = g+8: the 16 logits of one node live in 2 registers of the 8 lanes that share t.
// BOUNDED = the logits are sums of R products of two squashed vectors (|z| <= R): no max subtraction needed.
template <bool BOUNDED>
__device__ __forceinline__ void softmax_cols(float (&z)[8], bool vh0, bool vh1, const bool (&vn)[4]) {
    float m[4], e0[4], e1[4], s[4];
#pragma unroll
    for (int col = 0; col < 4; ++col) {
        const int i0 = (col & 1) + 4 * (col >> 1), i1 = i0 + 2;
        m[col] = 0.f;
        if (!BOUNDED) m[col] = fmaxf(vh0 ? z[i0] : -INFINITY, vh1 ? z[i1] : -INFINITY);
    }
    if (!BOUNDED) {
#pragma unroll
        for (int o = 4; o < 32; o <<= 1)
#pragma unroll
            for (int col = 0; col < 4; ++col) m[col] = fmaxf(m[col], __shfl_xor_sync(0xffffffffu, m[col], o));
    }
#pragma unroll
    for (int col = 0; col < 4; ++col) {
        const int i0 = (col & 1) + 4 * (col >> 1), i1 = i0 + 2;
        e0[col] = vh0 ? ex2_approx((z[i0] - m[col]) * LOG2E) : 0.f;
        e1[col] = vh1 ? ex2_approx((z[i1] - m[col]) * LOG2E) : 0.f;
        s[col] = e0[col] + e1[col];
    }
#pragma unroll
    for (int o = 4; o < 32; o <<= 1)
#pragma unroll
        for (int col = 0; col < 4; ++col) s[col] += __shfl_xor_sync(0xffffffffu, s[col], o);
#pragma unroll
    for (int col = 0; col < 4; ++col) {
        const int i0 = (col & 1) + 4 * (col >> 1), i1 = i0 + 2;
        const float inv = vn[col] ? __frcp_rn(s[col]) : 0.f;
        z[i0] = e0[col] * inv;
        z[i1] = e1[col] * inv;
    }
}

// logits of the warp's tiles: z[k] += V P_k^T; the V fragments of a k-block are fetched once for all tiles
template <int PREC, int TPW>
__device__ __forceinline__ void warp_logits_tiles(float (&z)[TPW][8], int ntv, const unsigned char* vpl, const unsigned char* rows,
                                                  int n0, int lane) {
    float zh[TPW][2][4], zl[TPW][2][4];
#pragma unroll
    for (int k = 0; k < TPW; ++k)
#pragma unroll
        for (int i = 0; i < 4; ++i) zh[k][0][i] = zh[k][1][i] = zl[k][0][i] = zl[k][1][i] = 0.f;
#pragma unroll
    for (int b = 0; b < 4; ++b) {
        uint32_t vh[4], vl[4] = {0u, 0u, 0u, 0u};
        const uint32_t aaddr = smem_u32(vpl + (size_t)(8 * ((lane >> 3) & 1) + (lane & 7)) * ROWB + (16 * b + 8 * (lane >> 4)) * 2);
        ldsm_x4(vh, aaddr);
        if (PREC == PREC_3XTF32) ldsm_x4(vl, aaddr + LO);
#pragma unroll
        for (int k = 0; k < TPW; ++k) {
            if (k < ntv) {
                uint32_t ph[4], pl[4] = {0u, 0u, 0u, 0u};
                const uint32_t baddr =
                    smem_u32(rows + (size_t)(n0 + 16 * k + 8 * (lane >> 4) + (lane & 7)) * ROWB + (16 * b + 8 * ((lane >> 3) & 1)) * 2);
                ldsm_x4(ph, baddr);
                if (PREC == PREC_3XTF32) {
                    ldsm_x4(pl, baddr + LO);
                    mma_f16(zl[k][0], vl, ph[0], ph[1]);
                    mma_f16(zl[k][1], vl, ph[2], ph[3]);
                    mma_f16(zl[k][0], vh, pl[0], pl[1]);
                    mma_f16(zl[k][1], vh, pl[2], pl[3]);
                }
                mma_f16(zh[k][0], vh, ph[0], ph[1]);
                mma_f16(zh[k][1], vh, ph[2], ph[3]);
            }
        }
    }
#pragma unroll
    for (int k = 0; k < TPW; ++k)
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            z[k][i] += zh[k][0][i] + zl[k][0][i];
            z[k][4 + i] += zh[k][1][i] + zl[k][1][i];
        }
}

// aggregation over the warp's tiles: red_w[h][:] = sum over its 16 * ntv nodes of c[h][n] P[n][:]  (rows h < HA); the tiles
// accumulate in the MMA accumulators, one store per warp
template <int PREC, int TPW>
__device__ __forceinline__ void warp_aggregate_tiles(const float (&c)[TPW][8], int ntv, int HA, const unsigned char* rows, int n0,
                                                     float* red_w, int lane) {
    const int g = lane >> 2, t = lane & 3;
    const int h0 = g, h1 = g + 8;
    uint32_t ah[TPW][4], al[TPW][4];
#pragma unroll
    for (int k = 0; k < TPW; ++k) {
        split_h2<PREC>(c[k][0], c[k][1], ah[k][0], al[k][0]);
        split_h2<PREC>(c[k][2], c[k][3], ah[k][1], al[k][1]);
        split_h2<PREC>(c[k][4], c[k][5], ah[k][2], al[k][2]);
        split_h2<PREC>(c[k][6], c[k][7], ah[k][3], al[k][3]);
    }
    float* r0 = red_w + (size_t)h0 * REDLD + 2 * t;
    float* r1 = red_w + (size_t)h1 * REDLD + 2 * t;
#pragma unroll
    for (int jp = 0; jp < 4; ++jp) {
        float a0[4] = {0.f, 0.f, 0.f, 0.f}, a1[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
        for (int k = 0; k < TPW; ++k) {
            if (k < ntv) {
                uint32_t bh[4], bq[4] = {0u, 0u, 0u, 0u};
                const uint32_t baddr =
                    smem_u32(rows + (size_t)(n0 + 16 * k + 8 * ((lane >> 3) & 1) + (lane & 7)) * ROWB + (16 * jp + 8 * (lane >> 4)) * 2);
                ldsm_x4_t(bh, baddr);
                if (PREC == PREC_3XTF32) ldsm_x4_t(bq, baddr + LO);
                mma3<PREC>(a0, ah[k], al[k], bh[0], bh[1], bq[0], bq[1]);
                mma3<PREC>(a1, ah[k], al[k], bh[2], bh[3], bq[2], bq[3]);
            }
        }
        if (h0 < HA) {
            *reinterpret_cast<float2*>(r0 + 16 * jp) = make_float2(a0[0], a0[1]);
            *reinterpret_cast<float2*>(r0 + 16 * jp + 8) = make_float2(a1[0], a1[1]);
        }
        if (h1 < HA) {
            *reinterpret_cast<float2*>(r1 + 16 * jp) = make_float2(a0[2], a0[3]);
            *reinterpret_cast<float2*>(r1 + 16 * jp + 8) = make_float2(a1[2], a1[3]);
        }
    }
}

template <int NW, int TPW, int MINB, int PREC>
__global__ void __launch_bounds__(NW * 32, MINB)
cap_route2_fwd_kernel(const float* __restrict__ x, const float* __restrict__ Wp, const float* __restrict__ bp,
                      const float* __restrict__ dadj, float* __restrict__ c_out, float* __restrict__ s_out,
                      float* __restrict__ z_out, int N, int H, int R, int ntiles) {
    pdl_enter();
    extern __shared__ __align__(128) unsigned char smraw[];
    unsigned char* Prow = smraw;                                  // [ntiles*16][ROWB]
    unsigned char* Wt = Prow + (size_t)ntiles * 16 * ROWB;        // [64][ROWB]  (out o, permuted k), dead after Z
    float* red = reinterpret_cast<float*>(Wt);                    // [NW][H+1][REDLD]
    unsigned char* vpl = Wt + wred_bytes(NW, H);                  // [16][ROWB]  v hi|lo planes, rows >= H stay zero
    float* us = reinterpret_cast<float*>(vpl + 16 * ROWB);        // [H][64]  u = squash(softmax(dadj) P)
    float* bps = us + (size_t)H * D;                              // [64]

    constexpr int NT = NW * 32;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int g = lane >> 2, t = lane & 3;
    const int slab = blockIdx.x;
    const float* xs = x + (size_t)slab * N * D;

    // ---- stage x (fp32 rows into the P slots), Wp (fp16 hi/lo, scaled, k permuted), bias; zero the v planes
    // (staging Wp in front of the PDL wait was measured: the x copies then start later and the chain loses 4 us)
    for (int i = tid; i < ntiles * 16 * 16; i += NT) {
        const int r = i >> 4, ch = i & 15;
        unsigned char* dst = Prow + (size_t)r * ROWB + ch * 16;
        if (r < N) cp_async16(dst, xs + (size_t)r * D + ch * 4);
        else *reinterpret_cast<float4*>(dst) = make_float4(0.f, 0.f, 0.f, 0.f);
    }
    stage_w_perm<PREC, NT>(Wt, Wp, WSCALE, tid);
    for (int i = tid; i < D; i += NT) bps[i] = bp[i];
    for (int i = tid; i < 16 * ROWB / 16; i += NT) reinterpret_cast<float4*>(vpl)[i] = make_float4(0.f, 0.f, 0.f, 0.f);

    // this warp's tiles and the incidence logits they need (issued early: consumed after the Z product)
    const int n0 = warp * TPW * 16;
    int ntv = ntiles - warp * TPW;                                // valid tiles of this warp (warp-uniform)
    ntv = ntv < 0 ? 0 : (ntv > TPW ? TPW : ntv);
    const int h0 = g, h1 = g + 8;
    const bool vh0 = h0 < H, vh1 = h1 < H;
    bool vn[TPW][4];
    float dz[TPW][8];
    {
        const float* dj = dadj + (size_t)slab * H * N;
#pragma unroll
        for (int k = 0; k < TPW; ++k) {
            const int na = n0 + 16 * k + 2 * t, nb = na + 8;
            vn[k][0] = na < N; vn[k][1] = na + 1 < N; vn[k][2] = nb < N; vn[k][3] = nb + 1 < N;
#pragma unroll
            for (int i = 0; i < 8; ++i) {
                const int h = (i & 2) ? h1 : h0;
                const int n = ((i & 4) ? nb : na) + (i & 1);
                dz[k][i] = (h < H && n < N) ? dj[(size_t)h * N + n] : 0.f;
            }
        }
    }
    cp_async_wait_all();
    __syncthreads();

    // ---- Z = x Wp^T + bp ; P = squash(Z) -> hi/lo planes, in place over the warp's own x rows
#pragma unroll
    for (int k = 0; k < TPW; ++k) {
        if (k < ntv) {
            float acc[8][4];
            const int nk = n0 + 16 * k;
            warp_xw_tile<PREC>(acc, Prow, Wt, nk, lane);
            const int ra = nk + g, rb = ra + 8;
            float q0 = 0.f, q1 = 0.f;
            constexpr float inv_scale = 1.f / WSCALE;
#pragma unroll
            for (int j = 0; j < 8; ++j) {
                const float b0 = bps[8 * j + 2 * t], b1 = bps[8 * j + 2 * t + 1];
                acc[j][0] = fmaf(acc[j][0], inv_scale, b0); acc[j][1] = fmaf(acc[j][1], inv_scale, b1);
                acc[j][2] = fmaf(acc[j][2], inv_scale, b0); acc[j][3] = fmaf(acc[j][3], inv_scale, b1);
                q0 += acc[j][0] * acc[j][0] + acc[j][1] * acc[j][1];
                q1 += acc[j][2] * acc[j][2] + acc[j][3] * acc[j][3];
            }
            q0 += __shfl_xor_sync(0xffffffffu, q0, 1); q0 += __shfl_xor_sync(0xffffffffu, q0, 2);
            q1 += __shfl_xor_sync(0xffffffffu, q1, 1); q1 += __shfl_xor_sync(0xffffffffu, q1, 2);
            const float f0 = (ra < N) ? squash_f(q0) : 0.f, f1 = (rb < N) ? squash_f(q1) : 0.f;
            if (z_out) {                                        // training: the backward reads Z instead of recomputing it
                float* zo = z_out + (size_t)slab * N * D;
#pragma unroll
                for (int j = 0; j < 8; ++j) {
                    if (ra < N) *reinterpret_cast<float2*>(zo + (size_t)ra * D + 8 * j + 2 * t) = make_float2(acc[j][0], acc[j][1]);
                    if (rb < N) *reinterpret_cast<float2*>(zo + (size_t)rb * D + 8 * j + 2 * t) = make_float2(acc[j][2], acc[j][3]);
                }
            }
            __syncwarp();
#pragma unroll
            for (int j = 0; j < 8; ++j) {
                uint32_t hi, lo;
                split_h2<PREC>(acc[j][0] * f0, acc[j][1] * f0, hi, lo);
                *reinterpret_cast<uint32_t*>(Prow + (size_t)ra * ROWB + (8 * j + 2 * t) * 2) = hi;
                *reinterpret_cast<uint32_t*>(Prow + (size_t)ra * ROWB + LO + (8 * j + 2 * t) * 2) = lo;
                split_h2<PREC>(acc[j][2] * f1, acc[j][3] * f1, hi, lo);
                *reinterpret_cast<uint32_t*>(Prow + (size_t)rb * ROWB + (8 * j + 2 * t) * 2) = hi;
                *reinterpret_cast<uint32_t*>(Prow + (size_t)rb * ROWB + LO + (8 * j + 2 * t) * 2) = lo;
            }
        }
    }
    __syncthreads();   // every warp is done with the Wp tile: its storage becomes `red`; P rows are warp-private

    float bl[TPW][8];  // routing logits b of this lane's (h, node) cells
#pragma unroll
    for (int k = 0; k < TPW; ++k)
#pragma unroll
        for (int i = 0; i < 8; ++i) bl[k][i] = 0.f;

    float* red_w = red + (size_t)warp * (H + 1) * REDLD;
    // cross-warp sum of one row of the partials (deterministic order); lane owns columns 2*lane, 2*lane+1
    auto row_total = [&](int h) {
        float2 s = make_float2(0.f, 0.f);
#pragma unroll
        for (int w = 0; w < NW; ++w) {
            const float2 p = *reinterpret_cast<const float2*>(red + ((size_t)w * (H + 1) + h) * REDLD + 2 * lane);
            s.x += p.x; s.y += p.y;
        }
        return s;
    };
    // v row h <- squash(vt) as fp16 hi/lo planes
    auto store_v = [&](int h, float2 vt) {
        const float q = warp_sum(vt.x * vt.x + vt.y * vt.y);
        const float f = squash_f(q);
        uint32_t hi, lo;
        split_h2<PREC>(vt.x * f, vt.y * f, hi, lo);
        *reinterpret_cast<uint32_t*>(vpl + (size_t)h * ROWB + lane * 4) = hi;
        *reinterpret_cast<uint32_t*>(vpl + (size_t)h * ROWB + LO + lane * 4) = lo;
    };

    // ---- pass A: c0 = softmax_H(dadj), plus an all-ones row H that yields sumP
    {
        float c[TPW][8];
#pragma unroll
        for (int k = 0; k < TPW; ++k) {
#pragma unroll
            for (int i = 0; i < 8; ++i) c[k][i] = dz[k][i];
            softmax_cols<false>(c[k], vh0, vh1, vn[k]);
            if (h0 == H) { c[k][0] = vn[k][0] ? 1.f : 0.f; c[k][1] = vn[k][1] ? 1.f : 0.f; c[k][4] = vn[k][2] ? 1.f : 0.f; c[k][5] = vn[k][3] ? 1.f : 0.f; }
            if (h1 == H) { c[k][2] = vn[k][0] ? 1.f : 0.f; c[k][3] = vn[k][1] ? 1.f : 0.f; c[k][6] = vn[k][2] ? 1.f : 0.f; c[k][7] = vn[k][3] ? 1.f : 0.f; }
        }
        warp_aggregate_tiles<PREC, TPW>(c, ntv, H + 1, Prow, n0, red_w, lane);
    }
    __syncthreads();
    {
        const float invH = 1.f / (float)H;
        float2 sp = make_float2(0.f, 0.f);
        if (R >= 1 && warp < H) sp = row_total(H);
        for (int h = warp; h < H; h += NW) {
            const float2 tt = row_total(h);
            const float q = warp_sum(tt.x * tt.x + tt.y * tt.y);
            const float f = squash_f(q);
            const float2 u = make_float2(tt.x * f, tt.y * f);
            *reinterpret_cast<float2*>(us + (size_t)h * D + 2 * lane) = u;
            if (R >= 1) store_v(h, make_float2(u.x * sp.x * invH, u.y * sp.y * invH));
        }
    }
    __syncthreads();
    // ---- routing iterations 2..R  (|b| <= R: the softmax needs no max subtraction)
    for (int it = 2; it <= R; ++it) {
        warp_logits_tiles<PREC, TPW>(bl, ntv, vpl, Prow, n0, lane);
        float c[TPW][8];
#pragma unroll
        for (int k = 0; k < TPW; ++k) {
#pragma unroll
            for (int i = 0; i < 8; ++i) c[k][i] = bl[k][i];
            softmax_cols<true>(c[k], vh0, vh1, vn[k]);
        }
        warp_aggregate_tiles<PREC, TPW>(c, ntv, H, Prow, n0, red_w, lane);
        __syncthreads();
        for (int h = warp; h < H; h += NW) {
            const float2 tt = row_total(h);
            const float2 u = *reinterpret_cast<const float2*>(us + (size_t)h * D + 2 * lane);
            store_v(h, make_float2(u.x * tt.x, u.y * tt.y));
        }
        __syncthreads();
    }
    // ---- final assignment c = softmax_H(b + dadj), s = c P
    {
        if (R >= 1) warp_logits_tiles<PREC, TPW>(bl, ntv, vpl, Prow, n0, lane);
        float c[TPW][8];
        float* co = c_out + (size_t)slab * H * N;
#pragma unroll
        for (int k = 0; k < TPW; ++k) {
#pragma unroll
            for (int i = 0; i < 8; ++i) c[k][i] = bl[k][i] + dz[k][i];
            softmax_cols<false>(c[k], vh0, vh1, vn[k]);
            const int na = n0 + 16 * k + 2 * t, nb = na + 8;
#pragma unroll
            for (int i = 0; i < 8; ++i) {
                const int h = (i & 2) ? h1 : h0;
                const int n = ((i & 4) ? nb : na) + (i & 1);
                if (h < H && n < N) co[(size_t)h * N + n] = c[k][i];
            }
        }
        warp_aggregate_tiles<PREC, TPW>(c, ntv, H, Prow, n0, red_w, lane);
    }
    __syncthreads();
    for (int h = warp; h < H; h += NW) {
        const float2 tt = row_total(h);
        *reinterpret_cast<float2*>(s_out + ((size_t)slab * H + h) * D + 2 * lane) = tt;
    }
}

template <int NW, int TPW, int MINB, int PREC>
static cudaError_t launch(const float* x, const float* Wp, const float* bp, const float* dadj, float* c, float* s, float* z,
                          int BT, int N, int H, int R, cudaStream_t st) {
    const int ntiles = (N + 15) / 16;
    const size_t smem = smem_bytes(NW, ntiles, H);
    auto kern = cap_route2_fwd_kernel<NW, TPW, MINB, PREC>;
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return e;
    launch_pdl(kern, dim3(BT), dim3(NW * 32), smem, st, x, Wp, bp, dadj, c, s, z, N, H, R, ntiles);
    return cudaGetLastError();
}

template <int PREC>
static cudaError_t dispatch(const float* x, const float* Wp, const float* bp, const float* dadj, float* c, float* s, float* z,
                            int BT, int N, int H, int R, cudaStream_t st) {
    // One tile per warp.  Two tiles per warp (6 warps, 74 KB -> three slabs per SM) were measured on the B200 and lost:
    // 13 % fewer shared-memory wavefronts but 25 % instead of 31 % active warps -> 54.9 vs 50.4 us
    // (profiles/ncu_route_fwd_r02.md); the kernel template keeps TPW as a parameter, only TPW = 1 is instantiated.
    if (N <= 64) return launch<4, 1, 4, PREC>(x, Wp, bp, dadj, c, s, z, BT, N, H, R, st);
    if (N <= 128) return launch<8, 1, 3, PREC>(x, Wp, bp, dadj, c, s, z, BT, N, H, R, st);
    if (N <= 176) return launch<11, 1, 2, PREC>(x, Wp, bp, dadj, c, s, z, BT, N, H, R, st);
    if (N <= 208) return launch<13, 1, 2, PREC>(x, Wp, bp, dadj, c, s, z, BT, N, H, R, st);
    return launch<16, 1, 1, PREC>(x, Wp, bp, dadj, c, s, z, BT, N, H, R, st);
}

}  // namespace r2

// entry used by gptst_cap_route_fwd (cap_route_fwd.cu); returns cudaErrorNotSupported-free: the caller checks the shape
bool route2_supported(int N, int D, int H) {
    static int legacy = -1;
    if (legacy < 0) {
        const char* e = getenv("GPTST_B200_ROUTE");   // "legacy" forces the tf32 generation (A/B measurements)
        legacy = (e && e[0] == 'l') ? 1 : 0;
    }
    return !legacy && D == 64 && N <= 256 && H >= 1 && H <= 15;
}

cudaError_t route2_fwd(const float* x, const float* Wp, const float* bp, const float* dadj, float* c, float* s, float* z, int BT,
                       int N, int H, int R, int prec, cudaStream_t st) {
    if (prec == PREC_3XTF32) return r2::dispatch<PREC_3XTF32>(x, Wp, bp, dadj, c, s, z, BT, N, H, R, st);
    return r2::dispatch<PREC_TF32>(x, Wp, bp, dadj, c, s, z, BT, N, H, R, st);
}

}  // namespace gptst
