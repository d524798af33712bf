// Hierarchical intra/inter-cluster hypergraph block `cap` (reference GPTST.py:100-141), forward + backward.
//
// Forward, per (b,t) slab of x (N x D):
//   route_fwd : Z = x Wp^T + bp ; P = squash(Z) ; dynamic routing on P (R iterations) with the
//               data-dependent incidence logits dadj ; c = softmax_H(b + dadj) (H x N) ; s = c P (H x D)
//   hop_fwd   : per sample b, inter-cluster hop over k = (t,h):  r = LReLU(dyn^T LReLU(dyn (s + tau))) + s ; v = squash(r)
//   recon     : recon = c^T v (N x D)      (the node-adaptive projection + residual is gproj.cu, group = node)
// Backward (routing logits are constants of the graph, GPTST.py:108-109):
//   dv_dcr    : dv = c drecon (H x D) ; dc_r = v drecon^T (H x N)
//   hop_bwd   : dv -> ds (incl. the direct path through r = ... + s), ddyn
//   route_bwd : dc = dc_r + ds P^T ; dL = c*(dc - sum_h c dc) -> ddadj ; dP = c^T ds ; dZ = squash'(Z, dP) ;
//               dx = dy + dZ Wp ; dWp = dZ^T x ; dbp = sum dZ
//
// The N-reductions of route_fwd are done with lanes over nodes / lanes over D and a deterministic
// cross-warp + cross-CTA (thread-block cluster, DSMEM) tree; the D x D contractions use tensor cores.
#include "cap_common.cuh"

namespace gptst {


// ------------------------------------------------------------------------------------------------------
// inter-cluster hop (per sample b), GPTST.py:125-134
// ------------------------------------------------------------------------------------------------------
template <int D>
__global__ void __launch_bounds__(256) cap_hop_fwd_kernel(const float* __restrict__ s, const float* __restrict__ dyn,
                                                          float* __restrict__ v, int T, int H, int HT) {
    extern __shared__ __align__(16) float smem[];
    const int K = T * H, LD = D + 1;
    float* Ss = smem;                 // [K][LD]  raw s
    float* E1 = Ss + (size_t)K * LD;  // [HT][LD]
    float* dy = E1 + (size_t)HT * LD; // [HT][K+1]
    float* tau = dy + (size_t)HT * (K + 1); // [K]
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int b = blockIdx.x;
    const float* sb = s + (size_t)b * K * D;
    for (int i = tid; i < K * D; i += 256) Ss[(i / D) * LD + (i % D)] = sb[i];
    for (int i = tid; i < HT * K; i += 256) dy[(i / K) * (K + 1) + (i % K)] = dyn[(size_t)b * HT * K + i];
    for (int k = tid; k < K; k += 256) tau[k] = (float)(k / H + 1) / 12.f;
    __syncthreads();
    for (int i = tid; i < HT * D; i += 256) {
        const int h = i / D, d = i % D;
        float a0 = 0.f, a1 = 0.f;
        int k = 0;
        for (; k + 1 < K; k += 2) {
            a0 = fmaf(dy[h * (K + 1) + k], Ss[k * LD + d] + tau[k], a0);
            a1 = fmaf(dy[h * (K + 1) + k + 1], Ss[(k + 1) * LD + d] + tau[k + 1], a1);
        }
        if (k < K) a0 = fmaf(dy[h * (K + 1) + k], Ss[k * LD + d] + tau[k], a0);
        E1[h * LD + d] = lrelu(a0 + a1);
    }
    __syncthreads();
    float* vb = v + (size_t)b * K * D;
    for (int k = warp; k < K; k += kWarps) {
        float r[D / 32];
        float q = 0.f;
#pragma unroll
        for (int j = 0; j < D / 32; ++j) {
            const int d = lane + 32 * j;
            float a = 0.f;
            for (int h = 0; h < HT; ++h) a = fmaf(dy[h * (K + 1) + k], E1[h * LD + d], a);
            r[j] = lrelu(a) + Ss[k * LD + d];
            q += r[j] * r[j];
        }
        q = warp_sum(q);
        const float f = squash_f(q);
#pragma unroll
        for (int j = 0; j < D / 32; ++j) vb[(size_t)k * D + lane + 32 * j] = r[j] * f;
    }
}

template <int D>
__global__ void __launch_bounds__(256) cap_hop_bwd_kernel(const float* __restrict__ s, const float* __restrict__ dyn,
                                                          const float* __restrict__ dv, float* __restrict__ ds,
                                                          float* __restrict__ ddyn, int T, int H, int HT) {
    extern __shared__ __align__(16) float smem[];
    const int K = T * H, LD = D + 1, LK = K + 1;
    float* Ss = smem;                   // [K][LD] raw s
    float* P2 = Ss + (size_t)K * LD;    // [K][LD] pre2, later dpre2
    float* DR = P2 + (size_t)K * LD;    // [K][LD] dr
    float* P1 = DR + (size_t)K * LD;    // [HT][LD] pre1
    float* D1 = P1 + (size_t)HT * LD;   // [HT][LD] dpre1
    float* dy = D1 + (size_t)HT * LD;   // [HT][LK]
    float* tau = dy + (size_t)HT * LK;  // [K]
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int b = blockIdx.x;
    const float* sb = s + (size_t)b * K * D;
    for (int i = tid; i < K * D; i += 256) Ss[(i / D) * LD + (i % D)] = sb[i];
    for (int i = tid; i < HT * K; i += 256) dy[(i / K) * LK + (i % K)] = dyn[(size_t)b * HT * K + i];
    for (int k = tid; k < K; k += 256) tau[k] = (float)(k / H + 1) / 12.f;
    __syncthreads();
    for (int i = tid; i < HT * D; i += 256) {
        const int h = i / D, d = i % D;
        float a0 = 0.f, a1 = 0.f;
        int k = 0;
        for (; k + 1 < K; k += 2) {
            a0 = fmaf(dy[h * LK + k], Ss[k * LD + d] + tau[k], a0);
            a1 = fmaf(dy[h * LK + k + 1], Ss[(k + 1) * LD + d] + tau[k + 1], a1);
        }
        if (k < K) a0 = fmaf(dy[h * LK + k], Ss[k * LD + d] + tau[k], a0);
        P1[h * LD + d] = a0 + a1;
    }
    __syncthreads();
    const float* dvb = dv + (size_t)b * K * D;
    for (int k = warp; k < K; k += kWarps) {
        float r[D / 32], g[D / 32], p2[D / 32];
        float q = 0.f, rg = 0.f;
#pragma unroll
        for (int j = 0; j < D / 32; ++j) {
            const int d = lane + 32 * j;
            float a = 0.f;
            for (int h = 0; h < HT; ++h) a = fmaf(dy[h * LK + k], lrelu(P1[h * LD + d]), a);
            p2[j] = a;
            r[j] = lrelu(a) + Ss[k * LD + d];
            g[j] = dvb[(size_t)k * D + d];
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
            DR[k * LD + d] = dr;
            P2[k * LD + d] = lrelu_grad(p2[j], dr);  // dpre2
        }
    }
    __syncthreads();
    for (int i = tid; i < HT * D; i += 256) {
        const int h = i / D, d = i % D;
        float a = 0.f;
        for (int k = 0; k < K; ++k) a = fmaf(dy[h * LK + k], P2[k * LD + d], a);
        D1[h * LD + d] = lrelu_grad(P1[h * LD + d], a);  // dpre1
    }
    __syncthreads();
    for (int i = tid; i < HT * K; i += 256) {
        const int h = i / K, k = i % K;
        const float tk = tau[k];
        float a = 0.f;
        for (int d = 0; d < D; ++d)
            a = fmaf(lrelu(P1[h * LD + d]), P2[k * LD + d], fmaf(D1[h * LD + d], Ss[k * LD + d] + tk, a));
        ddyn[(size_t)b * HT * K + i] = a;
    }
    float* dsb = ds + (size_t)b * K * D;
    for (int i = tid; i < K * D; i += 256) {
        const int k = i / D, d = i % D;
        float a = DR[k * LD + d];
        for (int h = 0; h < HT; ++h) a = fmaf(dy[h * LK + k], D1[h * LD + d], a);
        dsb[i] = a;
    }
}

// recon[b,t,n,:] = sum_h c[b,t,h,n] v[b,t,h,:]            GPTST.py:135
template <int D>
__global__ void __launch_bounds__(256) cap_recon_kernel(const float* __restrict__ c, const float* __restrict__ v,
                                                        float* __restrict__ recon, int N, int H) {
    __shared__ __align__(16) float vs[kMaxH * D];
    const int slab = blockIdx.x;
    for (int i = threadIdx.x; i < H * D; i += 256) vs[i] = v[(size_t)slab * H * D + i];
    __syncthreads();
    constexpr int VPR = D / 4, NPC = 256 / VPR;
    const int nl = threadIdx.x / VPR, cv = threadIdx.x % VPR;
    for (int n = blockIdx.y * NPC + nl; n < N; n += gridDim.y * NPC) {
        float4 o = make_float4(0.f, 0.f, 0.f, 0.f);
        for (int h = 0; h < H; ++h) {
            const float cc = c[((size_t)slab * H + h) * N + n];
            const float4 vv = *reinterpret_cast<const float4*>(vs + h * D + cv * 4);
            o.x = fmaf(cc, vv.x, o.x); o.y = fmaf(cc, vv.y, o.y); o.z = fmaf(cc, vv.z, o.z); o.w = fmaf(cc, vv.w, o.w);
        }
        *reinterpret_cast<float4*>(recon + ((size_t)slab * N + n) * D + cv * 4) = o;
    }
}

// dv[b,t,h,:] = sum_n c[h,n] drecon[n,:]      dc_r[b,t,h,n] = v[h,:] . drecon[n,:]
template <int D, int HP>
__global__ void __launch_bounds__(256) cap_dv_dcr_kernel(const float* __restrict__ c, const float* __restrict__ v,
                                                         const float* __restrict__ drecon, float* __restrict__ dv,
                                                         float* __restrict__ dcr, int N, int H) {
    constexpr int VPR = D / 4, NL = 256 / VPR;   // threads per row, node lanes
    extern __shared__ __align__(16) float smem[];
    float* vs = smem;                 // [H][D]
    float* red = vs + kMaxH * D;      // [NL][H][D]
    const int slab = blockIdx.x;
    const int tid = threadIdx.x, nl = tid / VPR, cv = tid % VPR;
    for (int i = tid; i < H * D; i += 256) vs[i] = v[(size_t)slab * H * D + i];
    __syncthreads();
    float4 acc[HP];
#pragma unroll
    for (int h = 0; h < HP; ++h) acc[h] = make_float4(0.f, 0.f, 0.f, 0.f);
    const int niter = (N + NL - 1) / NL;
    for (int it = 0; it < niter; ++it) {
        const int n = it * NL + nl;
        const bool valid = n < N;
        float4 g = make_float4(0.f, 0.f, 0.f, 0.f);
        if (valid) g = *reinterpret_cast<const float4*>(drecon + ((size_t)slab * N + n) * D + cv * 4);
#pragma unroll
        for (int h = 0; h < HP; ++h) {
            if (h < H) {
                const float cc = valid ? c[((size_t)slab * H + h) * N + n] : 0.f;
                acc[h].x = fmaf(cc, g.x, acc[h].x); acc[h].y = fmaf(cc, g.y, acc[h].y);
                acc[h].z = fmaf(cc, g.z, acc[h].z); acc[h].w = fmaf(cc, g.w, acc[h].w);
                const float4 vv = *reinterpret_cast<const float4*>(vs + h * D + cv * 4);
                float dot = g.x * vv.x + g.y * vv.y + g.z * vv.z + g.w * vv.w;
                // reduce over the VPR threads of this row (VPR = 16 or 32, aligned inside a warp)
#pragma unroll
                for (int o = VPR / 2; o > 0; o >>= 1) dot += __shfl_xor_sync(0xffffffffu, dot, o);
                if (cv == 0 && valid) dcr[((size_t)slab * H + h) * N + n] = dot;
            }
        }
    }
#pragma unroll
    for (int h = 0; h < HP; ++h)
        if (h < H) *reinterpret_cast<float4*>(red + ((size_t)nl * H + h) * D + cv * 4) = acc[h];
    __syncthreads();
    for (int i = tid; i < H * D; i += 256) {
        float sacc = 0.f;
        for (int l = 0; l < NL; ++l) sacc += red[(size_t)l * H * D + i];
        dv[(size_t)slab * H * D + i] = sacc;
    }
}

bool route2_supported(int N, int D, int H);   // cap_route2_fwd.cu
cudaError_t dv_dcr2(const float* c, const float* v, const float* drecon, float* dv, float* dcr, int BT, int N, int H,
                    cudaStream_t st);         // cap_dvdcr2.cu

}  // namespace gptst

using namespace gptst;

extern "C" int gptst_cap_hop_fwd(const float* s, const float* dyn, float* v, int B, int T, int D, int H, int HT,
                                 void* stream) {
    if (!s || !dyn || !v || B <= 0) return -1;
    if (H > kMaxH) return -2;
    cudaStream_t st = (cudaStream_t)stream;
    const int K = T * H;
    size_t smem = ((size_t)K * (D + 1) + (size_t)HT * (D + 1) + (size_t)HT * (K + 1) + K) * 4;
    if (smem > kSmemMax) return -2;
    cudaError_t e;
    if (D == 64) {
        e = cudaFuncSetAttribute(cap_hop_fwd_kernel<64>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) return (int)e;
        cap_hop_fwd_kernel<64><<<B, 256, smem, st>>>(s, dyn, v, T, H, HT);
    } else if (D == 128) {
        e = cudaFuncSetAttribute(cap_hop_fwd_kernel<128>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) return (int)e;
        cap_hop_fwd_kernel<128><<<B, 256, smem, st>>>(s, dyn, v, T, H, HT);
    } else return -2;
    return (int)cudaGetLastError();
}

extern "C" int gptst_cap_hop_bwd(const float* s, const float* dyn, const float* dv, float* ds, float* ddyn, int B, int T,
                                 int D, int H, int HT, void* stream) {
    if (!s || !dyn || !dv || !ds || !ddyn || B <= 0) return -1;
    if (H > kMaxH) return -2;
    cudaStream_t st = (cudaStream_t)stream;
    const int K = T * H;
    size_t smem = (3 * (size_t)K * (D + 1) + 2 * (size_t)HT * (D + 1) + (size_t)HT * (K + 1) + K) * 4;
    if (smem > kSmemMax) return -2;
    cudaError_t e;
    if (D == 64) {
        e = cudaFuncSetAttribute(cap_hop_bwd_kernel<64>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) return (int)e;
        cap_hop_bwd_kernel<64><<<B, 256, smem, st>>>(s, dyn, dv, ds, ddyn, T, H, HT);
    } else if (D == 128) {
        e = cudaFuncSetAttribute(cap_hop_bwd_kernel<128>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) return (int)e;
        cap_hop_bwd_kernel<128><<<B, 256, smem, st>>>(s, dyn, dv, ds, ddyn, T, H, HT);
    } else return -2;
    return (int)cudaGetLastError();
}

extern "C" int gptst_cap_recon(const float* c, const float* v, float* recon, int B, int T, int N, int D, int H,
                               void* stream) {
    if (!c || !v || !recon || B <= 0 || N <= 0) return -1;
    if (H > kMaxH) return -2;
    cudaStream_t st = (cudaStream_t)stream;
    const int npc = 256 / (D / 4);
    int ychunks = (N + npc - 1) / npc;
    int want = (592 + B * T - 1) / (B * T);
    if (ychunks > want) ychunks = want;
    if (ychunks < 1) ychunks = 1;
    dim3 grid(B * T, ychunks);
    if (D == 64) cap_recon_kernel<64><<<grid, 256, 0, st>>>(c, v, recon, N, H);
    else if (D == 128) cap_recon_kernel<128><<<grid, 256, 0, st>>>(c, v, recon, N, H);
    else return -2;
    return (int)cudaGetLastError();
}

extern "C" int gptst_cap_dv_dcr(const float* c, const float* v, const float* drecon, float* dv, float* dcr, int B, int T,
                                int N, int D, int H, void* stream) {
    if (!c || !v || !drecon || !dv || !dcr || B <= 0 || N <= 0) return -1;
    if (H > kMaxH) return -2;
    cudaStream_t st = (cudaStream_t)stream;
    if (route2_supported(N, D, H)) return (int)dv_dcr2(c, v, drecon, dv, dcr, B * T, N, H, st);
    const int nlanes = 256 / (D / 4);
    size_t smem = ((size_t)kMaxH * D + (size_t)nlanes * H * D) * 4;
    cudaError_t e;
#define DV(DD, HH)                                                                                                       \
    do {                                                                                                                 \
        e = cudaFuncSetAttribute(cap_dv_dcr_kernel<DD, HH>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);     \
        if (e != cudaSuccess) return (int)e;                                                                             \
        cap_dv_dcr_kernel<DD, HH><<<B * T, 256, smem, st>>>(c, v, drecon, dv, dcr, N, H);                                \
    } while (0)
    if (D == 64 && H == 10) DV(64, 10);
    else if (D == 64) DV(64, 16);
    else if (D == 128 && H == 10) DV(128, 10);
    else if (D == 128) DV(128, 16);
    else return -2;
#undef DV
    return (int)cudaGetLastError();
}

