// cap backward through the routing block, second generation for D = 64, N <= 256.
//
// The first generation did everything in one kernel (recompute Z/P, dc, dL, dP, dZ, dx += dZ Wp, dWp += dZ^T x).  Here the
// two contractions with the SHARED weight Wp are handed to the grouped-projection backward (gproj2.cu, one group, dX
// accumulated in place), and this kernel only produces dZ and ddadj -- per (b,t) slab, one warp per 16 nodes, fp16-split
// mma.sync exactly like the forward (cap_route2_fwd.cu):
//     Z = x Wp^T + bp ; P = squash(Z)                       (recomputed; Z stays in registers)
//     dc = dcr + ds P^T                                     logit-type MMA, B operand = P straight from the Z fragments
//     dL = c * (dc - sum_h c dc)  -> ddadj                  (softmax backward; the routing logits are constants)
//     dP = c^T ds                                           M = 16 nodes, K = 16 hyperedges (one k-step)
//     dZ = squash'(Z, dP) = f dP + 2 Z f'(q) <Z, dP>        -> dZ (B,T,N,D)
// ds is a gradient of arbitrary magnitude: it is brought into fp16 range with one power-of-two scale per slab.
#include "cap_common.cuh"
#include "mma_f16.cuh"

namespace gptst {
namespace r2 {

using namespace hf;

// FROMZ: the forward kernel stored Z (gptst_cap_route_fwd_z), `x` then points at Z and nothing is recomputed -- no x / Wp
// staging, no Z product (96 of the 168 three-term MMAs per tile), 4.4 KB of shared memory instead of 70 KB.  These kernels are
// bound by issue slots and shared-memory operand traffic, not by HBM (profiles/ncu_route_fwd_r02.md), so reading one more
// activation is cheaper than recomputing it.
template <int NW, int MINB, int PREC, bool FROMZ>
__global__ void __launch_bounds__(NW * 32, MINB)
cap_route2_bwd_dz_kernel(const float* __restrict__ x, const float* __restrict__ Wp, const float* __restrict__ bp,
                         const float* __restrict__ c, const float* __restrict__ ds, const float* __restrict__ dcr,
                         float* __restrict__ dZ, float* __restrict__ ddadj, int N, int H) {
    constexpr int D = 64;
    constexpr float WSCALE = 64.f;
    extern __shared__ __align__(128) unsigned char smraw[];
    unsigned char* Xs = smraw;                                     // [NW*16][ROWB] x rows (fp32)          (absent with FROMZ)
    unsigned char* Wt = Xs + (FROMZ ? 0 : (size_t)NW * 16 * ROWB); // [64][ROWB]                           (absent with FROMZ)
    unsigned char* dsp = Wt + (FROMZ ? 0 : (size_t)D * ROWB);      // [16][ROWB] ds hi|lo planes (scaled), rows >= H zero
    float* bps = reinterpret_cast<float*>(dsp + 16 * ROWB);        // [64]
    float* wmax = bps + D;                                         // [NW]

    constexpr int NT = NW * 32;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int g = lane >> 2, t = lane & 3;
    const int slab = blockIdx.x;
    const float* xs = x + (size_t)slab * N * D;
    float z[8][4];
    if (FROMZ) {
        // Z rows straight into the accumulator-fragment layout: 8-byte loads, 8 rows x 32 B per warp request (full sectors)
        const int ra_ = warp * 16 + g, rb_ = ra_ + 8;
#pragma unroll
        for (int j = 0; j < 8; ++j) {
            const float2 a = (ra_ < N) ? *reinterpret_cast<const float2*>(xs + (size_t)ra_ * D + 8 * j + 2 * t) : make_float2(0.f, 0.f);
            const float2 b = (rb_ < N) ? *reinterpret_cast<const float2*>(xs + (size_t)rb_ * D + 8 * j + 2 * t) : make_float2(0.f, 0.f);
            z[j][0] = a.x; z[j][1] = a.y; z[j][2] = b.x; z[j][3] = b.y;
        }
    } else {
        for (int i = tid; i < NW * 16 * 16; i += NT) {
            const int r = i >> 4, ch = i & 15;
            unsigned char* dst = Xs + (size_t)r * ROWB + ch * 16;
            if (r < N) cp_async16(dst, xs + (size_t)r * D + ch * 4);
            else *reinterpret_cast<float4*>(dst) = make_float4(0.f, 0.f, 0.f, 0.f);
        }
        stage_w_perm<PREC, NT>(Wt, Wp, WSCALE, tid);
        for (int i = tid; i < D; i += NT) bps[i] = bp[i];
    }
    // ds: each thread keeps (at most two) float2 of the H x 64 block, the slab max goes through wmax
    constexpr int DSP = (16 * 32 + NT - 1) / NT;      // float2 items per thread (rows padded to 16)
    float2 dsv[DSP];
    float m = 0.f;
#pragma unroll
    for (int k = 0; k < DSP; ++k) {
        const int i = tid + k * NT, h = i >> 5, p = i & 31;
        dsv[k] = make_float2(0.f, 0.f);
        if (i < 16 * 32 && h < H) dsv[k] = *reinterpret_cast<const float2*>(ds + ((size_t)slab * H + h) * D + 2 * p);
        m = fmaxf(m, fmaxf(fabsf(dsv[k].x), fabsf(dsv[k].y)));
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, o));
    if (lane == 0) wmax[warp] = m;

    const int n0 = warp * 16;
    const int h0 = g, h1 = g + 8;
    const int na = n0 + 2 * t, nb = na + 8;
    // c and dcr in the logit-fragment layout: [0]=(h0,na) [1]=(h0,na+1) [2]=(h1,na) [3]=(h1,na+1) [4..7] same for nb
    float cf[8], dc[8];
    {
        const float* cs = c + (size_t)slab * H * N;
        const float* dr = dcr + (size_t)slab * H * N;
#pragma unroll
        for (int i = 0; i < 8; ++i) {
            const int h = (i & 2) ? h1 : h0;
            const int n = ((i & 4) ? nb : na) + (i & 1);
            const bool ok = h < H && n < N;
            cf[i] = ok ? cs[(size_t)h * N + n] : 0.f;
            dc[i] = ok ? dr[(size_t)h * N + n] : 0.f;
        }
    }
    cp_async_wait_all();
    __syncthreads();
    float2 sc;
    {
        float mm = 0.f;
#pragma unroll
        for (int w = 0; w < NW; ++w) mm = fmaxf(mm, wmax[w]);
        sc = pow2_scale_for_fp16(mm);
    }
#pragma unroll
    for (int k = 0; k < DSP; ++k) {
        const int i = tid + k * NT, h = i >> 5, p = i & 31;
        if (i < 16 * 32) {
            uint32_t hi, lo;
            split_h2<PREC>(dsv[k].x * sc.x, dsv[k].y * sc.x, hi, lo);
            *reinterpret_cast<uint32_t*>(dsp + (size_t)h * ROWB + p * 4) = hi;
            *reinterpret_cast<uint32_t*>(dsp + (size_t)h * ROWB + LO + p * 4) = lo;
        }
    }
    // ---- Z (registers), row norms
    if (!FROMZ) warp_xw_tile<PREC>(z, Xs, Wt, n0, lane);
    const int ra = n0 + g, rb = ra + 8;
    float q0 = 0.f, q1 = 0.f;
    {
        constexpr float inv_scale = 1.f / WSCALE;
#pragma unroll
        for (int j = 0; j < 8; ++j) {
            if (!FROMZ) {
                const float b0 = bps[8 * j + 2 * t], b1 = bps[8 * j + 2 * t + 1];
                z[j][0] = fmaf(z[j][0], inv_scale, b0); z[j][1] = fmaf(z[j][1], inv_scale, b1);
                z[j][2] = fmaf(z[j][2], inv_scale, b0); z[j][3] = fmaf(z[j][3], inv_scale, b1);
            }
            q0 += z[j][0] * z[j][0] + z[j][1] * z[j][1];
            q1 += z[j][2] * z[j][2] + z[j][3] * z[j][3];
        }
        q0 += __shfl_xor_sync(0xffffffffu, q0, 1); q0 += __shfl_xor_sync(0xffffffffu, q0, 2);
        q1 += __shfl_xor_sync(0xffffffffu, q1, 1); q1 += __shfl_xor_sync(0xffffffffu, q1, 2);
    }
    const bool va = ra < N, vb = rb < N;
    const float f0 = va ? squash_f(q0) : 0.f, f1 = vb ? squash_f(q1) : 0.f;
    const float fp0 = va ? squash_df(q0) : 0.f, fp1 = vb ? squash_df(q1) : 0.f;
    __syncthreads();   // ds planes complete
    // ---- dc += ds P^T : A = ds planes, B = P from the Z fragments (node tile 0 = rows g, tile 1 = rows g+8)
    {
        float zh[2][4], zl[2][4];
#pragma unroll
        for (int i = 0; i < 4; ++i) zh[0][i] = zh[1][i] = zl[0][i] = zl[1][i] = 0.f;
#pragma unroll
        for (int b = 0; b < 4; ++b) {
            uint32_t vh[4], vl[4] = {0u, 0u, 0u, 0u};
            const uint32_t aaddr = smem_u32(dsp + (size_t)(8 * ((lane >> 3) & 1) + (lane & 7)) * ROWB + (16 * b + 8 * (lane >> 4)) * 2);
            ldsm_x4(vh, aaddr);
            if (PREC == PREC_3XTF32) ldsm_x4(vl, aaddr + LO);
            uint32_t p0h[2], p0l[2], p1h[2], p1l[2];
            split_h2<PREC>(z[2 * b][0] * f0, z[2 * b][1] * f0, p0h[0], p0l[0]);
            split_h2<PREC>(z[2 * b + 1][0] * f0, z[2 * b + 1][1] * f0, p0h[1], p0l[1]);
            split_h2<PREC>(z[2 * b][2] * f1, z[2 * b][3] * f1, p1h[0], p1l[0]);
            split_h2<PREC>(z[2 * b + 1][2] * f1, z[2 * b + 1][3] * f1, p1h[1], p1l[1]);
            if (PREC == PREC_3XTF32) {
                mma_f16(zl[0], vl, p0h[0], p0h[1]);
                mma_f16(zl[1], vl, p1h[0], p1h[1]);
                mma_f16(zl[0], vh, p0l[0], p0l[1]);
                mma_f16(zl[1], vh, p1l[0], p1l[1]);
            }
            mma_f16(zh[0], vh, p0h[0], p0h[1]);
            mma_f16(zh[1], vh, p1h[0], p1h[1]);
        }
        // The B operand's n index is the node (tile 0 = nodes n0..n0+7 = rows g of the Z fragments), so the C fragment
        // (h0 / h1, nodes n0+2t, n0+2t+1 | +8 for tile 1) has exactly the layout of cf / dc.
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            dc[i] += (zh[0][i] + zl[0][i]) * sc.y;
            dc[4 + i] += (zh[1][i] + zl[1][i]) * sc.y;
        }
    }
    // ---- dL = c (dc - sum_h c dc)
    {
        float* dj = ddadj + (size_t)slab * H * N;
#pragma unroll
        for (int col = 0; col < 4; ++col) {
            const int i0 = (col & 1) + 4 * (col >> 1), i1 = i0 + 2;
            float s = cf[i0] * dc[i0] + cf[i1] * dc[i1];
#pragma unroll
            for (int o = 4; o < 32; o <<= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
            const int n = ((col >> 1) ? nb : na) + (col & 1);
            if (n < N) {
                if (h0 < H) dj[(size_t)h0 * N + n] = cf[i0] * (dc[i0] - s);
                if (h1 < H) dj[(size_t)h1 * N + n] = cf[i1] * (dc[i1] - s);
            }
        }
    }
    // ---- dP = c^T ds : A[m = node][k = h] = c[h][node]
    uint32_t ch[4], cl[4];
    {
        const float* cs = c + (size_t)slab * H * N;
        float a[8];
#pragma unroll
        for (int i = 0; i < 8; ++i) {
            // i: 0,1 -> (node g,   h 2t, 2t+1)   2,3 -> (node g+8, h 2t, 2t+1)   4,5 -> (node g, h 2t+8, 2t+9)   6,7 -> (node g+8, ..)
            const int h = 2 * t + (i & 1) + ((i & 4) ? 8 : 0);
            const int n = n0 + g + ((i & 2) ? 8 : 0);
            a[i] = (h < H && n < N) ? cs[(size_t)h * N + n] : 0.f;
        }
        split_h2<PREC>(a[0], a[1], ch[0], cl[0]);
        split_h2<PREC>(a[2], a[3], ch[1], cl[1]);
        split_h2<PREC>(a[4], a[5], ch[2], cl[2]);
        split_h2<PREC>(a[6], a[7], ch[3], cl[3]);
    }
    // one (16 nodes x 8 columns) tile of dP; B = ds planes transposed (k = h)
    auto dp_tiles = [&](int jp, float (&t0)[4], float (&t1)[4]) {
        uint32_t bh[4], bq[4] = {0u, 0u, 0u, 0u};
        const uint32_t baddr = smem_u32(dsp + (size_t)(8 * ((lane >> 3) & 1) + (lane & 7)) * ROWB + (16 * jp + 8 * (lane >> 4)) * 2);
        ldsm_x4_t(bh, baddr);
        if (PREC == PREC_3XTF32) ldsm_x4_t(bq, baddr + LO);
        t0[0] = t0[1] = t0[2] = t0[3] = 0.f;
        t1[0] = t1[1] = t1[2] = t1[3] = 0.f;
        mma3<PREC>(t0, ch, cl, bh[0], bh[1], bq[0], bq[1]);
        mma3<PREC>(t1, ch, cl, bh[2], bh[3], bq[2], bq[3]);
    };
    // pass 1: row dots <Z, dP>   (dP tiles are recomputed in pass 2: 24 more MMAs instead of 32 more live registers)
    float d0 = 0.f, d1 = 0.f;
#pragma unroll
    for (int jp = 0; jp < 4; ++jp) {
        float t0[4], t1[4];
        dp_tiles(jp, t0, t1);
        d0 += z[2 * jp][0] * t0[0] + z[2 * jp][1] * t0[1] + z[2 * jp + 1][0] * t1[0] + z[2 * jp + 1][1] * t1[1];
        d1 += z[2 * jp][2] * t0[2] + z[2 * jp][3] * t0[3] + z[2 * jp + 1][2] * t1[2] + z[2 * jp + 1][3] * t1[3];
    }
    d0 += __shfl_xor_sync(0xffffffffu, d0, 1); d0 += __shfl_xor_sync(0xffffffffu, d0, 2);
    d1 += __shfl_xor_sync(0xffffffffu, d1, 1); d1 += __shfl_xor_sync(0xffffffffu, d1, 2);
    // dZ = f dP + 2 Z f' <Z,dP>   (dP carries the ds scale: undo with sc.y)
    const float g0 = f0 * sc.y, g1 = f1 * sc.y;
    const float e0 = 2.f * fp0 * d0 * sc.y, e1 = 2.f * fp1 * d1 * sc.y;
    float* dzs = dZ + (size_t)slab * N * D;
#pragma unroll
    for (int jp = 0; jp < 4; ++jp) {
        float t0[4], t1[4];
        dp_tiles(jp, t0, t1);
        const int col = 16 * jp + 2 * t;
        if (va) {
            *reinterpret_cast<float2*>(dzs + (size_t)ra * D + col) =
                make_float2(fmaf(g0, t0[0], e0 * z[2 * jp][0]), fmaf(g0, t0[1], e0 * z[2 * jp][1]));
            *reinterpret_cast<float2*>(dzs + (size_t)ra * D + col + 8) =
                make_float2(fmaf(g0, t1[0], e0 * z[2 * jp + 1][0]), fmaf(g0, t1[1], e0 * z[2 * jp + 1][1]));
        }
        if (vb) {
            *reinterpret_cast<float2*>(dzs + (size_t)rb * D + col) =
                make_float2(fmaf(g1, t0[2], e1 * z[2 * jp][2]), fmaf(g1, t0[3], e1 * z[2 * jp][3]));
            *reinterpret_cast<float2*>(dzs + (size_t)rb * D + col + 8) =
                make_float2(fmaf(g1, t1[2], e1 * z[2 * jp + 1][2]), fmaf(g1, t1[3], e1 * z[2 * jp + 1][3]));
        }
    }
}

template <int NW, int MINB, int PREC, bool FROMZ = false>
static cudaError_t launch_bwd_dz(const float* x, const float* Wp, const float* bp, const float* c, const float* ds,
                                 const float* dcr, float* dZ, float* ddadj, int BT, int N, int H, cudaStream_t st) {
    const size_t smem = (FROMZ ? 0 : (size_t)NW * 16 * ROWB + 64 * ROWB) + 16 * ROWB + (64 + NW) * 4;
    auto kern = cap_route2_bwd_dz_kernel<NW, MINB, PREC, FROMZ>;
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return e;
    kern<<<BT, NW * 32, smem, st>>>(x, Wp, bp, c, ds, dcr, dZ, ddadj, N, H);
    return cudaGetLastError();
}

template <int PREC>
static cudaError_t dispatch_bwd_dz(const float* x, const float* Wp, const float* bp, const float* c, const float* ds,
                                   const float* dcr, float* dZ, float* ddadj, int BT, int N, int H, cudaStream_t st) {
    if (N <= 64) return launch_bwd_dz<4, 4, PREC>(x, Wp, bp, c, ds, dcr, dZ, ddadj, BT, N, H, st);
    if (N <= 128) return launch_bwd_dz<8, 3, PREC>(x, Wp, bp, c, ds, dcr, dZ, ddadj, BT, N, H, st);
    if (N <= 176) return launch_bwd_dz<11, 2, PREC>(x, Wp, bp, c, ds, dcr, dZ, ddadj, BT, N, H, st);
    if (N <= 208) return launch_bwd_dz<13, 2, PREC>(x, Wp, bp, c, ds, dcr, dZ, ddadj, BT, N, H, st);
    return launch_bwd_dz<16, 1, PREC>(x, Wp, bp, c, ds, dcr, dZ, ddadj, BT, N, H, st);
}

template <int PREC>
static cudaError_t dispatch_bwd_dz_z(const float* z, const float* c, const float* ds, const float* dcr, float* dZ, float* ddadj,
                                     int BT, int N, int H, cudaStream_t st) {
    if (N <= 64) return launch_bwd_dz<4, 4, PREC, true>(z, nullptr, nullptr, c, ds, dcr, dZ, ddadj, BT, N, H, st);
    if (N <= 128) return launch_bwd_dz<8, 3, PREC, true>(z, nullptr, nullptr, c, ds, dcr, dZ, ddadj, BT, N, H, st);
    if (N <= 176) return launch_bwd_dz<11, 2, PREC, true>(z, nullptr, nullptr, c, ds, dcr, dZ, ddadj, BT, N, H, st);
    if (N <= 208) return launch_bwd_dz<13, 2, PREC, true>(z, nullptr, nullptr, c, ds, dcr, dZ, ddadj, BT, N, H, st);
    return launch_bwd_dz<16, 1, PREC, true>(z, nullptr, nullptr, c, ds, dcr, dZ, ddadj, BT, N, H, st);
}

}  // namespace r2

bool route2_supported(int N, int D, int H);

}  // namespace gptst

using namespace gptst;

extern "C" int gptst_cap_route2_supported(int N, int D, int H) { return route2_supported(N, D, H) ? 1 : 0; }

extern "C" int gptst_cap_route_bwd_dz(const float* x, const float* Wp, const float* bp, const float* c, const float* ds,
                                      const float* dcr, float* dZ, float* ddadj, int B, int T, int N, int D, int H, int prec,
                                      void* stream) {
    if (!x || !Wp || !bp || !c || !ds || !dcr || !dZ || !ddadj || B <= 0 || T <= 0 || N <= 0) return -1;
    if (!route2_supported(N, D, H) || (prec != 1 && prec != 3)) return -2;
    cudaStream_t st = (cudaStream_t)stream;
    if (prec == 3) return (int)r2::dispatch_bwd_dz<PREC_3XTF32>(x, Wp, bp, c, ds, dcr, dZ, ddadj, B * T, N, H, st);
    return (int)r2::dispatch_bwd_dz<PREC_TF32>(x, Wp, bp, c, ds, dcr, dZ, ddadj, B * T, N, H, st);
}

// Same, from the Z the forward stored (gptst_cap_route_fwd_z): z (B,T,N,D) = x Wp^T + bp.
extern "C" int gptst_cap_route_bwd_dz_z(const float* z, const float* c, const float* ds, const float* dcr, float* dZ,
                                        float* ddadj, int B, int T, int N, int D, int H, int prec, void* stream) {
    if (!z || !c || !ds || !dcr || !dZ || !ddadj || B <= 0 || T <= 0 || N <= 0) return -1;
    if (!route2_supported(N, D, H) || (prec != 1 && prec != 3)) return -2;
    cudaStream_t st = (cudaStream_t)stream;
    if (prec == 3) return (int)r2::dispatch_bwd_dz_z<PREC_3XTF32>(z, c, ds, dcr, dZ, ddadj, B * T, N, H, st);
    return (int)r2::dispatch_bwd_dz_z<PREC_TF32>(z, c, ds, dcr, dZ, ddadj, B * T, N, H, st);
}
