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

// grid (B*T, ychunks), 256 threads
template <int D>
__global__ void __launch_bounds__(256) cap_recon_hop_kernel(const float* __restrict__ c, const float* __restrict__ s,
                                                            const float* __restrict__ dyn, const float* __restrict__ e1,
                                                            float* __restrict__ v, float* __restrict__ recon, int T, int N,
                                                            int H, int HT) {
    extern __shared__ __align__(16) float smem[];
    float* E1 = smem;                      // [HT][D]
    float* dy = E1 + (size_t)HT * D;       // [HT][H]   dyn[b][:, t*H .. t*H+H)
    float* vs = dy + (size_t)HT * H;       // [H][D]
    const int slab = blockIdx.x, b = slab / T, tt = slab % T;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int K = T * H;
    for (int i = tid; i < HT * D / 4; i += 256)
        reinterpret_cast<float4*>(E1)[i] = reinterpret_cast<const float4*>(e1 + (size_t)b * HT * D)[i];
    for (int i = tid; i < HT * H; i += 256) dy[i] = dyn[((size_t)b * HT + i / H) * K + tt * H + (i % H)];
    __syncthreads();
    for (int h = warp; h < H; h += 8) {
        float r[D / 32];
        float q = 0.f;
#pragma unroll
        for (int j = 0; j < D / 32; ++j) {
            const int d = lane + 32 * j;
            float a = 0.f;
            for (int ht = 0; ht < HT; ++ht) a = fmaf(dy[ht * H + h], E1[ht * D + d], a);
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
            const float cc = c[((size_t)slab * H + h) * N + n];
            const float4 vv = *reinterpret_cast<const float4*>(vs + h * D + cv * 4);
            o.x = fmaf(cc, vv.x, o.x); o.y = fmaf(cc, vv.y, o.y); o.z = fmaf(cc, vv.z, o.z); o.w = fmaf(cc, vv.w, o.w);
        }
        *reinterpret_cast<float4*>(recon + ((size_t)slab * N + n) * D + cv * 4) = o;
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

extern "C" int gptst_cap_recon_hop(const float* c, const float* s, const float* dyn, const float* e1, float* v,
                                   float* recon, int B, int T, int N, int D, int H, int HT, void* stream) {
    if (!c || !s || !dyn || !e1 || !v || !recon || B <= 0 || T <= 0 || N <= 0 || HT <= 0) return -1;
    if (H < 1 || H > kMaxH) return -2;
    cudaStream_t st = (cudaStream_t)stream;
    const int npc = 256 / (D / 4);
    int ychunks = (N + npc - 1) / npc;
    int want = (592 + B * T - 1) / (B * T);
    if (ychunks > want) ychunks = want;
    if (ychunks < 1) ychunks = 1;
    const size_t smem = ((size_t)HT * D + (size_t)HT * H + (size_t)H * D) * 4;
    if (smem > 48 * 1024) return -2;
    dim3 grid(B * T, ychunks);
    if (D == 64) cap_recon_hop_kernel<64><<<grid, 256, smem, st>>>(c, s, dyn, e1, v, recon, T, N, H, HT);
    else if (D == 128) cap_recon_hop_kernel<128><<<grid, 256, smem, st>>>(c, s, dyn, e1, v, recon, T, N, H, HT);
    else return -2;
    return (int)cudaGetLastError();
}
