// cap: inter-cluster hop (reference GPTST.py:125-134) + hyperedge -> node reconstruction (GPTST.py:135), split so
// that no launch is limited to one CTA per sample:
//
//   hop_e1     : E1[b] = LReLU(dyn_b (s_b + tau))   (HT x D per sample; the only part that mixes the 12 slabs of a
//                sample).  Column-separable in D, so the grid is (B, D/16) CTAs.
//   recon_hop  : per (b,t) slab:  r = LReLU(dyn_b[:, t-block]^T E1[b]) + s ;  v = squash(r) -> v (kept for backward) ;
//                recon[n,:] = sum_h c[h,n] v[h,:].
// Replaces cap_hop_fwd (one CTA per sample, 26 us at B=64) + cap_recon on the forward path.
#include "cap_common.cuh"

namespace gptst {

constexpr int kE1Cols = 16;

// grid (B, D/16), 256 threads: thread = (ht = tid/16 (+16 per round), column tid%16)
__global__ void __launch_bounds__(256) cap_hop_e1_kernel(const float* __restrict__ s, const float* __restrict__ dyn,
                                                         float* __restrict__ e1, int T, int D, int H, int HT) {
    extern __shared__ __align__(16) float smem[];
    const int K = T * H;
    float* Ss = smem;                        // [K][16]   s + tau
    float* dy = Ss + (size_t)K * kE1Cols;    // [HT][K+1]
    const int tid = threadIdx.x, b = blockIdx.x, c0 = blockIdx.y * kE1Cols;
    const float* sb = s + (size_t)b * K * D + c0;
    for (int i = tid; i < K * (kE1Cols / 4); i += 256) {
        const int k = i / (kE1Cols / 4), q = i % (kE1Cols / 4);
        float4 v = *reinterpret_cast<const float4*>(sb + (size_t)k * D + 4 * q);
        const float tau = (float)(k / H + 1) / 12.f;
        v.x += tau; v.y += tau; v.z += tau; v.w += tau;
        *reinterpret_cast<float4*>(Ss + k * kE1Cols + 4 * q) = v;
    }
    for (int i = tid; i < HT * K; i += 256) dy[(i / K) * (K + 1) + (i % K)] = dyn[(size_t)b * HT * K + i];
    __syncthreads();
    const int col = tid & (kE1Cols - 1);
    for (int ht = tid / kE1Cols; ht < HT; ht += 256 / kE1Cols) {
        const float* dr = dy + ht * (K + 1);
        float a0 = 0.f, a1 = 0.f, a2 = 0.f, a3 = 0.f;
        int k = 0;
        for (; k + 3 < K; k += 4) {
            a0 = fmaf(dr[k], Ss[k * kE1Cols + col], a0);
            a1 = fmaf(dr[k + 1], Ss[(k + 1) * kE1Cols + col], a1);
            a2 = fmaf(dr[k + 2], Ss[(k + 2) * kE1Cols + col], a2);
            a3 = fmaf(dr[k + 3], Ss[(k + 3) * kE1Cols + col], a3);
        }
        for (; k < K; ++k) a0 = fmaf(dr[k], Ss[k * kE1Cols + col], a0);
        e1[((size_t)b * HT + ht) * D + c0 + col] = lrelu((a0 + a1) + (a2 + a3));
    }
}

// grid (B*T, ychunks), 256 threads.  FUSE_E1: E1[b] is recomputed by every slab CTA of the sample (123k MACs, s_b read from
// L2) instead of being read from a separate hop_e1 launch; the t == 0 CTA writes it out for the backward pass.
template <int D, bool FUSE_E1>
__global__ void __launch_bounds__(256) cap_recon_hop_kernel(const float* __restrict__ c, const float* __restrict__ s,
                                                            const float* __restrict__ dyn, float* __restrict__ e1,
                                                            float* __restrict__ v, float* __restrict__ recon, int T, int N,
                                                            int H, int HT) {
    extern __shared__ __align__(16) float smem[];
    const int K = T * H;
    float* E1 = smem;                      // [HT][D]
    float* dy = E1 + (size_t)HT * D;       // FUSE_E1: [HT][K] (all of dyn_b) ; else [HT][H] = dyn[b][:, t*H .. t*H+H)
    float* vs = dy + (size_t)HT * (FUSE_E1 ? K : H);       // [H][D]
    float* cs = vs + (size_t)H * D;        // [H][N]  the slab's incidence (staged up front: its latency hides under the hop)
    const int slab = blockIdx.x, b = slab / T, tt = slab % T;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int dstride = FUSE_E1 ? K : H, doff = FUSE_E1 ? tt * H : 0;    // dy[ht * dstride + doff + h] = dyn[b][ht][t*H + h]
    for (int i = tid; i < H * N; i += 256) cs[i] = c[(size_t)slab * H * N + i];
    if (FUSE_E1) {
        for (int i = tid; i < HT * K; i += 256) dy[i] = dyn[(size_t)b * HT * K + i];
        __syncthreads();
        // E1[ht][d] = phi( sum_k dyn[ht][k] (s_b[k][d] + tau_k) ): thread = (column d, 4 rows ht = g, g + HT/4, ...)
        const int d = tid % D, grp = tid / D;          // 256 / D groups
        constexpr int G = 256 / D;
        const float* sb = s + (size_t)b * K * D + d;
        for (int h0 = grp; h0 < HT; h0 += 4 * G) {
            float a[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll 8
            for (int k = 0; k < K; ++k) {
                const float sv = sb[(size_t)k * D] + (float)(k / H + 1) / 12.f;
#pragma unroll
                for (int j = 0; j < 4; ++j)
                    if (h0 + j * G < HT) a[j] = fmaf(dy[(h0 + j * G) * K + k], sv, a[j]);
            }
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                const int ht = h0 + j * G;
                if (ht < HT) {
                    const float val = lrelu(a[j]);
                    E1[ht * D + d] = val;
                    if (tt == 0 && blockIdx.y == 0) e1[((size_t)b * HT + ht) * D + d] = val;
                }
            }
        }
    } else {
        for (int i = tid; i < HT * D / 4; i += 256)
            reinterpret_cast<float4*>(E1)[i] = reinterpret_cast<const float4*>(e1 + (size_t)b * HT * D)[i];
        for (int i = tid; i < HT * H; i += 256) dy[i] = dyn[((size_t)b * HT + i / H) * K + tt * H + (i % H)];
    }
    __syncthreads();
    for (int h = warp; h < H; h += 8) {
        float r[D / 32];
        float q = 0.f;
#pragma unroll
        for (int j = 0; j < D / 32; ++j) {
            const int d = lane + 32 * j;
            float a = 0.f;
            for (int ht = 0; ht < HT; ++ht) a = fmaf(dy[ht * dstride + doff + h], E1[ht * D + d], a);
            r[j] = lrelu(a) + s[((size_t)slab * H + h) * D + d];
            q += r[j] * r[j];
        }
        q = warp_sum(q);
        const float f = squash_f(q);
#pragma unroll
        for (int j = 0; j < D / 32; ++j) {
            const int d = lane + 32 * j;
            const float val = r[j] * f;
            vs[h * D + d] = val;
            if (blockIdx.y == 0) v[((size_t)slab * H + h) * D + d] = val;
        }
    }
    __syncthreads();
    constexpr int VPR = D / 4, NPC = 256 / VPR;
    const int nl = tid / VPR, cv = tid % VPR;
    for (int n = blockIdx.y * NPC + nl; n < N; n += gridDim.y * NPC) {
        float4 o = make_float4(0.f, 0.f, 0.f, 0.f);
        for (int h = 0; h < H; ++h) {
            const float cc = cs[h * N + n];
            const float4 vv = *reinterpret_cast<const float4*>(vs + h * D + cv * 4);
            o.x = fmaf(cc, vv.x, o.x); o.y = fmaf(cc, vv.y, o.y); o.z = fmaf(cc, vv.z, o.z); o.w = fmaf(cc, vv.w, o.w);
        }
        *reinterpret_cast<float4*>(recon + ((size_t)slab * N + n) * D + cv * 4) = o;
    }
}


// ---- backward of the hop, split the same way ---------------------------------------------
//   hop_bwd_rows (per slab)        : recompute pre2 = dyn_b[:, t-block]^T E1 and r ; dr = squash'(r, dv) ; dpre2 = dr * phi'(pre2)
//   hop_bwd_cols (per sample, 16 columns of D) : dE1 = dyn dpre2 ; dpre1 = dE1 * phi'(pre1) ; ds = dr + dyn^T dpre1 ;
//                                    ddyn_part[chunk] = E1 dpre2^T + dpre1 (s+tau)^T  restricted to the chunk's columns
template <int D>
__global__ void __launch_bounds__(256) cap_hop_bwd_rows_kernel(const float* __restrict__ s, const float* __restrict__ dyn,
                                                               const float* __restrict__ e1, const float* __restrict__ dv,
                                                               float* __restrict__ dr_out, float* __restrict__ dpre2_out,
                                                               int T, int H, int HT) {
    extern __shared__ __align__(16) float smem[];
    float* E1 = smem;                      // [HT][D]
    float* dy = E1 + (size_t)HT * D;       // [HT][H]
    const int slab = blockIdx.x, b = slab / T, tt = slab % T;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int K = T * H;
    for (int i = tid; i < HT * D / 4; i += 256)
        reinterpret_cast<float4*>(E1)[i] = reinterpret_cast<const float4*>(e1 + (size_t)b * HT * D)[i];
    for (int i = tid; i < HT * H; i += 256) dy[i] = dyn[((size_t)b * HT + i / H) * K + tt * H + (i % H)];
    __syncthreads();
    for (int h = warp; h < H; h += 8) {
        float r[D / 32], g[D / 32], p2[D / 32];
        float q = 0.f, rg = 0.f;
        const size_t row = ((size_t)slab * H + h) * D;
#pragma unroll
        for (int j = 0; j < D / 32; ++j) {
            const int d = lane + 32 * j;
            float a = 0.f;
            for (int ht = 0; ht < HT; ++ht) a = fmaf(dy[ht * H + h], E1[ht * D + d], a);
            p2[j] = a;
            r[j] = lrelu(a) + s[row + d];
            g[j] = dv[row + d];
            q += r[j] * r[j];
            rg += r[j] * g[j];
        }
        q = warp_sum(q);
        rg = warp_sum(rg);
        const float f = squash_f(q), fp = squash_df(q);
#pragma unroll
        for (int j = 0; j < D / 32; ++j) {
            const int d = lane + 32 * j;
            const float dr = f * g[j] + 2.f * r[j] * fp * rg;
            dr_out[row + d] = dr;
            dpre2_out[row + d] = lrelu_grad(p2[j], dr);
        }
    }
}

// grid (B, D/16), 256 threads
__global__ void __launch_bounds__(256) cap_hop_bwd_cols_kernel(const float* __restrict__ s, const float* __restrict__ dyn,
                                                               const float* __restrict__ e1, const float* __restrict__ dr,
                                                               const float* __restrict__ dpre2, float* __restrict__ ds,
                                                               float* __restrict__ ddyn_part, int B, int T, int D, int H,
                                                               int HT) {
    extern __shared__ __align__(16) float smem[];
    // Ss / P2 rows are padded to C + 1 floats: the ddyn pass below walks them with the lane index on k, and a stride of 16
    // floats put every other row on the same bank (16-way conflicts = most of this kernel's former 32 us)
    const int K = T * H, C = kE1Cols, CP = kE1Cols + 1;
    float* Ss = smem;                         // [K][CP]  s + tau
    float* P2 = Ss + (size_t)K * CP;          // [K][CP]  dpre2
    float* E1 = P2 + (size_t)K * CP;          // [HT][C]
    float* D1 = E1 + (size_t)HT * C;          // [HT][C]  dpre1
    float* dy = D1 + (size_t)HT * C;          // [HT][K+1]
    const int tid = threadIdx.x, b = blockIdx.x, c0 = blockIdx.y * C;
    for (int i = tid; i < K * (C / 4); i += 256) {
        const int k = i / (C / 4), q = i % (C / 4);
        const size_t off = ((size_t)b * K + k) * D + c0 + 4 * q;
        float4 v = *reinterpret_cast<const float4*>(s + off);
        const float tau = (float)(k / H + 1) / 12.f;
        const float4 p = *reinterpret_cast<const float4*>(dpre2 + off);
        float* sd = Ss + k * CP + 4 * q;
        float* pd = P2 + k * CP + 4 * q;
        sd[0] = v.x + tau; sd[1] = v.y + tau; sd[2] = v.z + tau; sd[3] = v.w + tau;
        pd[0] = p.x; pd[1] = p.y; pd[2] = p.z; pd[3] = p.w;
    }
    for (int i = tid; i < HT * (C / 4); i += 256) {
        const int ht = i / (C / 4), q = i % (C / 4);
        *reinterpret_cast<float4*>(E1 + ht * C + 4 * q) = *reinterpret_cast<const float4*>(e1 + ((size_t)b * HT + ht) * D + c0 + 4 * q);
    }
    for (int i = tid; i < HT * K; i += 256) dy[(i / K) * (K + 1) + (i % K)] = dyn[(size_t)b * HT * K + i];
    __syncthreads();
    const int col = tid & (C - 1);
    for (int ht = tid / C; ht < HT; ht += 256 / C) {
        const float* drow = dy + ht * (K + 1);
        float a0 = 0.f, a1 = 0.f, a2 = 0.f, a3 = 0.f;
        int k = 0;
        for (; k + 3 < K; k += 4) {
            a0 = fmaf(drow[k], P2[k * CP + col], a0);
            a1 = fmaf(drow[k + 1], P2[(k + 1) * CP + col], a1);
            a2 = fmaf(drow[k + 2], P2[(k + 2) * CP + col], a2);
            a3 = fmaf(drow[k + 3], P2[(k + 3) * CP + col], a3);
        }
        for (; k < K; ++k) a0 = fmaf(drow[k], P2[k * CP + col], a0);
        D1[ht * C + col] = lrelu_grad(E1[ht * C + col], (a0 + a1) + (a2 + a3));     // sign(E1) == sign(pre1)
    }
    __syncthreads();
    for (int i = tid; i < K * C; i += 256) {
        const int k = i / C, cc = i % C;
        const size_t off = ((size_t)b * K + k) * D + c0 + cc;
        float a = dr[off];
        for (int ht = 0; ht < HT; ++ht) a = fmaf(dy[ht * (K + 1) + k], D1[ht * C + cc], a);
        ds[off] = a;
    }
    float* dp = ddyn_part + ((size_t)blockIdx.y * B + b) * HT * K;
    for (int i = tid; i < HT * K; i += 256) {
        const int ht = i / K, k = i % K;
        float a = 0.f;
#pragma unroll
        for (int cc = 0; cc < C; ++cc) a = fmaf(E1[ht * C + cc], P2[k * CP + cc], fmaf(D1[ht * C + cc], Ss[k * CP + cc], a));
        dp[i] = a;
    }
}

}  // namespace gptst

using namespace gptst;

extern "C" int gptst_cap_hop_e1(const float* s, const float* dyn, float* e1, int B, int T, int D, int H, int HT,
                                void* stream) {
    if (!s || !dyn || !e1 || B <= 0 || T <= 0 || HT <= 0) return -1;
    if (H < 1 || H > kMaxH || D % kE1Cols != 0) return -2;
    const int K = T * H;
    const size_t smem = ((size_t)K * kE1Cols + (size_t)HT * (K + 1)) * 4;
    if (smem > kSmemMax) return -2;
    cudaError_t e = cudaFuncSetAttribute(cap_hop_e1_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return (int)e;
    cap_hop_e1_kernel<<<dim3(B, D / kE1Cols), 256, smem, (cudaStream_t)stream>>>(s, dyn, e1, T, D, H, HT);
    return (int)cudaGetLastError();
}

template <bool FUSE>
static int launch_recon_hop(const float* c, const float* s, const float* dyn, float* e1, float* v, float* recon, int B, int T,
                            int N, int D, int H, int HT, cudaStream_t st) {
    const int npc = 256 / (D / 4);
    int ychunks = (N + npc - 1) / npc;
    int want = (592 + B * T - 1) / (B * T);
    if (ychunks > want) ychunks = want;
    if (ychunks < 1) ychunks = 1;
    const size_t smem = ((size_t)HT * D + (size_t)HT * (FUSE ? T * H : H) + (size_t)H * D + (size_t)H * N) * 4;
    if (smem > kSmemMax) return -2;
    dim3 grid(B * T, ychunks);
    cudaError_t e;
    if (D == 64) {
        e = cudaFuncSetAttribute(cap_recon_hop_kernel<64, FUSE>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) return (int)e;
        cap_recon_hop_kernel<64, FUSE><<<grid, 256, smem, st>>>(c, s, dyn, e1, v, recon, T, N, H, HT);
    } else if (D == 128) {
        e = cudaFuncSetAttribute(cap_recon_hop_kernel<128, FUSE>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) return (int)e;
        cap_recon_hop_kernel<128, FUSE><<<grid, 256, smem, st>>>(c, s, dyn, e1, v, recon, T, N, H, HT);
    } else return -2;
    return (int)cudaGetLastError();
}

extern "C" int gptst_cap_recon_hop(const float* c, const float* s, const float* dyn, const float* e1, float* v,
                                   float* recon, int B, int T, int N, int D, int H, int HT, void* stream) {
    if (!c || !s || !dyn || !e1 || !v || !recon || B <= 0 || T <= 0 || N <= 0 || HT <= 0) return -1;
    if (H < 1 || H > kMaxH) return -2;
    return launch_recon_hop<false>(c, s, dyn, const_cast<float*>(e1), v, recon, B, T, N, D, H, HT, (cudaStream_t)stream);
}

// the same with hop_e1 folded in: e1 (B,HT,D) is an OUTPUT (kept for the backward pass); one launch instead of two
extern "C" int gptst_cap_recon_hop_fused(const float* c, const float* s, const float* dyn, float* e1, float* v, float* recon,
                                         int B, int T, int N, int D, int H, int HT, void* stream) {
    if (!c || !s || !dyn || !e1 || !v || !recon || B <= 0 || T <= 0 || N <= 0 || HT <= 0) return -1;
    if (H < 1 || H > kMaxH) return -2;
    return launch_recon_hop<true>(c, s, dyn, e1, v, recon, B, T, N, D, H, HT, (cudaStream_t)stream);
}

extern "C" int gptst_cap_hop_bwd_parts(int D) { return D / kE1Cols; }

// dr_tmp, dpre2_tmp: scratch (B,T,H,D) each; ddyn_part: (gptst_cap_hop_bwd_parts(D), B, HT, T*H), summed by the caller.
extern "C" int gptst_cap_hop_bwd2(const float* s, const float* dyn, const float* e1, const float* dv, float* dr_tmp,
                                  float* dpre2_tmp, float* ds, float* ddyn_part, int B, int T, int D, int H, int HT,
                                  void* stream) {
    if (!s || !dyn || !e1 || !dv || !dr_tmp || !dpre2_tmp || !ds || !ddyn_part || B <= 0 || T <= 0 || HT <= 0) return -1;
    if (H < 1 || H > kMaxH || D % kE1Cols != 0) return -2;
    cudaStream_t st = (cudaStream_t)stream;
    const int K = T * H;
    const size_t smem1 = ((size_t)HT * D + (size_t)HT * H) * 4;
    const size_t smem2 = ((size_t)2 * K * (kE1Cols + 1) + (size_t)2 * HT * kE1Cols + (size_t)HT * (K + 1)) * 4;
    if (smem1 > 48 * 1024 || smem2 > kSmemMax) return -2;
    if (D == 64) cap_hop_bwd_rows_kernel<64><<<B * T, 256, smem1, st>>>(s, dyn, e1, dv, dr_tmp, dpre2_tmp, T, H, HT);
    else if (D == 128) cap_hop_bwd_rows_kernel<128><<<B * T, 256, smem1, st>>>(s, dyn, e1, dv, dr_tmp, dpre2_tmp, T, H, HT);
    else return -2;
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) return (int)e;
    e = cudaFuncSetAttribute(cap_hop_bwd_cols_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem2);
    if (e != cudaSuccess) return (int)e;
    cap_hop_bwd_cols_kernel<<<dim3(B, D / kE1Cols), 256, smem2, st>>>(s, dyn, e1, dr_tmp, dpre2_tmp, ds, ddyn_part, B, T, D, H, HT);
    return (int)cudaGetLastError();
}

// column pass only (dr / dpre2 already produced, e.g. by gptst_cap_dv_dcr_hoprows)
extern "C" int gptst_cap_hop_bwd_cols(const float* s, const float* dyn, const float* e1, const float* dr, const float* dpre2,
                                      float* ds, float* ddyn_part, int B, int T, int D, int H, int HT, void* stream) {
    if (!s || !dyn || !e1 || !dr || !dpre2 || !ds || !ddyn_part || B <= 0 || T <= 0 || HT <= 0) return -1;
    if (H < 1 || H > kMaxH || D % kE1Cols != 0) return -2;
    const int K = T * H;
    const size_t smem2 = ((size_t)2 * K * (kE1Cols + 1) + (size_t)2 * HT * kE1Cols + (size_t)HT * (K + 1)) * 4;
    if (smem2 > kSmemMax) return -2;
    cudaError_t e = cudaFuncSetAttribute(cap_hop_bwd_cols_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem2);
    if (e != cudaSuccess) return (int)e;
    cap_hop_bwd_cols_kernel<<<dim3(B, D / kE1Cols), 256, smem2, (cudaStream_t)stream>>>(s, dyn, e1, dr, dpre2, ds, ddyn_part, B, T, D,
                                                                                      H, HT);
    return (int)cudaGetLastError();
}
