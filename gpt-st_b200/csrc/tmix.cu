// Temporal hypergraph two-hop of hyperTem (reference GPTST.py:156-158), restated as a per-node T x T mix:
//   A_n[h,t] = sum_k E[n,k] adj[k,h,t]      M_n = A_n^T A_n  (T x T, symmetric)
//   ret[b,t,n,:] = sum_t' M_n[t,t'] eb[b,t',n,:]
// M_n is built by the host side (N*T*T floats); this file holds the streaming kernels over (B,T,N,D):
//   tmix      y[b,t,n,:] (+)= sum_t' M[n][t][t'] x[b,t',n,:]     (transpose flag uses M[n][t'][t])
//   tmix_dM   dM[n][t][t'] = sum_{b,j} dy[b,t,n,j] x[b,t',n,j]   (gradient w.r.t. the mix matrix)
// Both are HBM-streaming: every element of x is read once and every element of y written once.
#include "common.cuh"

namespace gptst {

// thread <-> (node, float4 column group); each thread keeps the T input vectors of its (b, n, cols) in registers
template <int T>
__global__ void __launch_bounds__(256) tmix_kernel(const float* __restrict__ x, const float* __restrict__ M,
                                                   float* __restrict__ y, int B, int N, int D, int transpose,
                                                   int accumulate) {
    extern __shared__ float Ms[];  // [nodes_per_cta][T*T]
    const int vpr = D / 4;                    // float4 per row
    const int npc = blockDim.x / vpr;         // nodes per CTA
    const int n0 = blockIdx.x * npc;
    const int b = blockIdx.y;
    for (int i = threadIdx.x; i < npc * T * T; i += blockDim.x) {
        int nl = i / (T * T), e = i % (T * T);
        int tr = e / T, tc = e % T;
        float v = 0.f;
        if (n0 + nl < N) v = M[(size_t)(n0 + nl) * T * T + (transpose ? tc * T + tr : e)];
        Ms[i] = v;
    }
    __syncthreads();
    const int nl = threadIdx.x / vpr, cv = threadIdx.x % vpr;
    const int n = n0 + nl;
    if (n >= N) return;
    const size_t slab = (size_t)N * D;
    const float* xp = x + (size_t)b * T * slab + (size_t)n * D + cv * 4;
    float* yp = y + (size_t)b * T * slab + (size_t)n * D + cv * 4;
    float4 in[T];
#pragma unroll
    for (int t = 0; t < T; ++t) in[t] = *reinterpret_cast<const float4*>(xp + t * slab);
    const float* m = Ms + nl * T * T;
#pragma unroll
    for (int t = 0; t < T; ++t) {
        float4 o = make_float4(0.f, 0.f, 0.f, 0.f);
        if (accumulate) o = *reinterpret_cast<const float4*>(yp + t * slab);
#pragma unroll
        for (int s = 0; s < T; ++s) {
            const float w = m[t * T + s];
            o.x = fmaf(w, in[s].x, o.x);
            o.y = fmaf(w, in[s].y, o.y);
            o.z = fmaf(w, in[s].z, o.z);
            o.w = fmaf(w, in[s].w, o.w);
        }
        *reinterpret_cast<float4*>(yp + t * slab) = o;
    }
}

// two warps per node (each owns T/2 rows of the T x T result -> 72 accumulators, ~110 registers, 4x the occupancy of a
// one-warp-per-node layout); lanes cover D (VEC floats each); partial over a batch split.  out: [splits][N][T*T]
template <int T, int VEC>
__global__ void __launch_bounds__(128) tmix_dM_kernel(const float* __restrict__ dy, const float* __restrict__ x,
                                                      float* __restrict__ dM_part, int B, int N, int D) {
    constexpr int TH = T / 2;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int n = blockIdx.x * (blockDim.x >> 6) + (warp >> 1);
    const int half = warp & 1;
    if (n >= N) return;
    const size_t slab = (size_t)N * D;
    float acc[TH * T];
#pragma unroll
    for (int i = 0; i < TH * T; ++i) acc[i] = 0.f;
    for (int b = blockIdx.y; b < B; b += gridDim.y)
    for (int c0 = 0; c0 < D; c0 += 32 * VEC) {
        const float* dp = dy + (size_t)b * T * slab + (size_t)(half * TH) * slab + (size_t)n * D + c0 + lane * VEC;
        const float* xp = x + (size_t)b * T * slab + (size_t)n * D + c0 + lane * VEC;
        float dv[TH][VEC], xv[T][VEC];
#pragma unroll
        for (int t = 0; t < T; ++t) {
            if (VEC == 2) {
                float2 c = *reinterpret_cast<const float2*>(xp + t * slab);
                xv[t][0] = c.x; xv[t][VEC > 1 ? 1 : 0] = c.y;
            } else {
                xv[t][0] = xp[t * slab];
            }
        }
#pragma unroll
        for (int t = 0; t < TH; ++t) {
            if (VEC == 2) {
                float2 a = *reinterpret_cast<const float2*>(dp + t * slab);
                dv[t][0] = a.x; dv[t][VEC > 1 ? 1 : 0] = a.y;
            } else {
                dv[t][0] = dp[t * slab];
            }
        }
#pragma unroll
        for (int t = 0; t < TH; ++t)
#pragma unroll
            for (int s2 = 0; s2 < T; ++s2)
#pragma unroll
                for (int v = 0; v < VEC; ++v) acc[t * T + s2] = fmaf(dv[t][v], xv[s2][v], acc[t * T + s2]);
    }
    float* out = dM_part + ((size_t)blockIdx.y * N + n) * T * T + half * TH * T;
#pragma unroll
    for (int i = 0; i < TH * T; ++i) {
        float v = warp_sum(acc[i]);
        if (lane == (i & 31)) out[i] = v;
    }
}

}  // namespace gptst

using namespace gptst;

extern "C" int gptst_tmix(const float* x, const float* M, float* y, int B, int T, int N, int D, int transpose,
                          int accumulate, void* stream) {
    if (!x || !M || !y || B <= 0 || N <= 0) return -1;
    if (T != kMaxT || D % 4 != 0 || D > 1024) return -2;
    const int vpr = D / 4;
    if (256 % vpr != 0) return -2;
    const int npc = 256 / vpr;
    dim3 grid((N + npc - 1) / npc, B);
    size_t smem = (size_t)npc * T * T * sizeof(float);
    tmix_kernel<kMaxT><<<grid, 256, smem, (cudaStream_t)stream>>>(x, M, y, B, N, D, transpose, accumulate);
    return (int)cudaGetLastError();
}

extern "C" int gptst_tmix_dM_splits(int B, int N) {
    int ctas = (N + 1) / 2;
    int s = (592 + ctas - 1) / ctas;
    if (s > B) s = B;
    return s < 1 ? 1 : s;
}

extern "C" int gptst_tmix_dM(const float* dy, const float* x, float* dM_part, int B, int T, int N, int D, int splits,
                             void* stream) {
    if (!dy || !x || !dM_part || B <= 0 || N <= 0 || splits <= 0) return -1;
    if (T != kMaxT) return -2;
    dim3 grid((N + 1) / 2, splits);   // 2 nodes (4 warps) per CTA
    cudaStream_t st = (cudaStream_t)stream;
    if (D % 64 == 0) tmix_dM_kernel<kMaxT, 2><<<grid, 128, 0, st>>>(dy, x, dM_part, B, N, D);
    else if (D == 32) tmix_dM_kernel<kMaxT, 1><<<grid, 128, 0, st>>>(dy, x, dM_part, B, N, D);
    else return -2;
    return (int)cudaGetLastError();
}
