// EXPERIMENTAL variant of cap_recon_hop (cap_hop2.cu, reference GPTST.py:130-135), NOT on the default path; reachable only through
// gptst_cap_recon_hop3 (same arguments as gptst_cap_recon_hop, bit-identical results expected: same FMA order per output).
//
// The reconstruction recon[n,:] = sum_h c[h,n] v[h,:] of the default kernel re-reads the H float4 of v for every node: 200 bytes
// of shared-memory reads per 16 bytes of output, ~544 KB per slab, which at 5 slabs per SM is about half of its 22 us.  A thread's
// column quad is fixed, so its H float4 of v are loop-invariant over the nodes it walks: they are hoisted into registers here
// (kMaxH float4, fully unrolled), leaving one 4-byte broadcast read of c per (node, hyperedge).
#include "cap_common.cuh"

namespace gptst {
namespace h3 {

// grid (B*T, ychunks), 256 threads
template <int D>
__global__ void __launch_bounds__(256) cap_recon_hop3_kernel(const float* __restrict__ c, const float* __restrict__ s,
                                                             const float* __restrict__ dyn, const float* __restrict__ e1,
                                                             float* __restrict__ v, float* __restrict__ recon, int T, int N,
                                                             int H, int HT) {
    extern __shared__ __align__(16) float smem[];
    const int K = T * H;
    float* E1 = smem;                      // [HT][D]
    float* dy = E1 + (size_t)HT * D;       // [HT][H] = dyn[b][:, t*H .. t*H+H)
    float* vs = dy + (size_t)HT * H;       // [H][D]
    float* cs = vs + (size_t)H * D;        // [H][N]
    const int slab = blockIdx.x, b = slab / T, tt = slab % T;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    for (int i = tid; i < H * N; i += 256) cs[i] = c[(size_t)slab * H * N + i];
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
    float4 vr[kMaxH];                      // this thread's column quad of every v row: loop-invariant over the nodes
#pragma unroll
    for (int h = 0; h < kMaxH; ++h)
        vr[h] = (h < H) ? *reinterpret_cast<const float4*>(vs + h * D + cv * 4) : make_float4(0.f, 0.f, 0.f, 0.f);
    for (int n = blockIdx.y * NPC + nl; n < N; n += gridDim.y * NPC) {
        float4 o = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
        for (int h = 0; h < kMaxH; ++h) {
            if (h < H) {                   // uniform
                const float cc = cs[h * N + n];
                o.x = fmaf(cc, vr[h].x, o.x); o.y = fmaf(cc, vr[h].y, o.y); o.z = fmaf(cc, vr[h].z, o.z); o.w = fmaf(cc, vr[h].w, o.w);
            }
        }
        *reinterpret_cast<float4*>(recon + ((size_t)slab * N + n) * D + cv * 4) = o;
    }
}

}  // namespace h3
}  // namespace gptst

using namespace gptst;

extern "C" int gptst_cap_recon_hop3(const float* c, const float* s, const float* dyn, const float* e1, float* v, float* recon, int B,
                                    int T, int N, int D, int H, int HT, void* stream) {
    if (!c || !s || !dyn || !e1 || !v || !recon || B <= 0 || T <= 0 || N <= 0 || HT <= 0) return -1;
    if (H < 1 || H > kMaxH) return -2;
    cudaStream_t st = (cudaStream_t)stream;
    const int npc = 256 / (D / 4);
    int ychunks = (N + npc - 1) / npc;
    int want = (592 + B * T - 1) / (B * T);
    if (ychunks > want) ychunks = want;
    if (ychunks < 1) ychunks = 1;
    const size_t smem = ((size_t)HT * D + (size_t)HT * H + (size_t)H * D + (size_t)H * N) * 4;
    if (smem > kSmemMax) return -2;
    dim3 grid(B * T, ychunks);
    cudaError_t e;
    if (D == 64) {
        e = cudaFuncSetAttribute(h3::cap_recon_hop3_kernel<64>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) return (int)e;
        h3::cap_recon_hop3_kernel<64><<<grid, 256, smem, st>>>(c, s, dyn, e1, v, recon, T, N, H, HT);
    } else if (D == 128) {
        e = cudaFuncSetAttribute(h3::cap_recon_hop3_kernel<128>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) return (int)e;
        h3::cap_recon_hop3_kernel<128><<<grid, 256, smem, st>>>(c, s, dyn, e1, v, recon, T, N, H, HT);
    } else return -2;
    return (int)cudaGetLastError();
}
