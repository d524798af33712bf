// cap forward after the routing: the inter-cluster hop in one launch (cap_hop_ev_kernel, GPTST.py:125-134, bottom of this file)
// and the second half of the block in ONE launch (reference GPTST.py:135-141), D = 64:
//
//     recon[b,t,n,:] = sum_h c[b,t,h,n] v[b,t,h,:]                       hyperedge -> node reconstruction   (:135)
//     out[b,t,n,:]   = LReLU( recon[b,t,n,:] W_n + bias_n + x[b,t,n,:] )  node-adaptive GCN + residual        (:137-141)
//
// The reconstruction is slab-major (one v per (b,t)), the projection is node-major (one weight per node): the three-launch
// path wrote `recon` (33 MB at PEMS08 / batch 64) and read it back in node-grouped order.  Here a CTA owns
// (8 consecutive nodes) x (32 consecutive slabs): 256 rows that share 8 weights and 32 v's.
//   * build : thread = (slab, 4 columns) keeps its slab's H rows of v in registers (10 x float4) and walks the 8 nodes; the
//             8 x H incidence values of its slab come from a shared-memory copy of the c tile (32-byte global segments).
//             Exact fp32 FMA in hyperedge order (bit-identical to cap_recon_hop_kernel); |recon| <= 1 (a convex combination
//             of squashed rows), so the fp16 hi | lo split needs no scale.  Training keeps recon for the backward (coalesced
//             256-byte rows straight from registers); inference never writes it.
//   * MMA   : warp = node.  A = the node's 32 recon rows (ldmatrix from the hi | lo planes), B = W_n in FRAGMENT ORDER
//             (the table htem_pack_kernel writes once per call: coalesced 512-byte requests through L2, one k-block ahead,
//             never through shared memory), two column halves one after the other, 24 mma.sync per 4 fragment loads.
//   * epilogue in the accumulator layout: bias + residual (32-byte sectors of x) + LeakyReLU -> out.
// Two CTAs per SM (80 KB each) run out of phase, so the FMA pipe of one overlaps the tensor pipe of the other.
#include "common.cuh"
#include "mma_f16.cuh"

namespace gptst {
namespace crp {

using namespace hf;
constexpr int D = 64, NG = 8, MT = 2, ROWS = 16 * MT, NTH = 256;
constexpr float WSCALE = 64.f;                        // the pack kernel's weight scale
constexpr int WF_VEC_PER_GROUP = 2 * 4 * 4 * 32;      // uint4 per node: [half][kb][j][lane]

__device__ __forceinline__ void fma4(float4& a, float m, const float4& e) {
    const float2 mm = make_float2(m, m);
    const float2 lo = __ffma2_rn(mm, make_float2(e.x, e.y), make_float2(a.x, a.y));
    const float2 hi = __ffma2_rn(mm, make_float2(e.z, e.w), make_float2(a.z, a.w));
    a = make_float4(lo.x, lo.y, hi.x, hi.y);
}

template <int HMAX, bool WRITE_RECON>
__global__ void __launch_bounds__(NTH, 2)
cap_recon_proj_kernel(const float* __restrict__ c, const float* __restrict__ v, const float* __restrict__ x,
                      const uint4* __restrict__ wfrag, const float* __restrict__ bias, float* __restrict__ out,
                      float* __restrict__ recon, int BT, int N, int H) {
    // (PDL: the wait sits behind the incidence staging below)
    constexpr int HP = (HMAX + 3) / 4 * 4;
    extern __shared__ __align__(128) unsigned char sm[];
    unsigned char* As = sm;                                              // [NG][ROWS][ROWB]
    float* cs = reinterpret_cast<float*>(sm + (size_t)NG * ROWS * ROWB); // [ROWS][NG][HP]
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int n0 = blockIdx.x * NG, s0 = blockIdx.y * ROWS;

    {   // incidence tile: thread = (slab, node), its H values -> one padded row
        const int s = tid >> 3, j = tid & 7;
        const bool ok = (s0 + s < BT) && (n0 + j < N);
        const float* cp = c + ((size_t)(s0 + s) * H) * N + n0 + j;
        float cv[HP];
#pragma unroll
        for (int h = 0; h < HP; ++h) cv[h] = (ok && h < H) ? __ldg(cp + (size_t)h * N) : 0.f;
        float4* dst = reinterpret_cast<float4*>(cs + (size_t)(s * NG + j) * HP);
#pragma unroll
        for (int q = 0; q < HP / 4; ++q) dst[q] = make_float4(cv[4 * q], cv[4 * q + 1], cv[4 * q + 2], cv[4 * q + 3]);
    }
    // c comes from the routing kernel, TWO launches back: complete once the hop kernel (the direct predecessor) let this grid
    // start (pdl_enter waits before it triggers), so the tile above is staged while the hop kernel runs; v is the hop's output
    pdl_wait();
    pdl_trigger();
    __syncthreads();

    {   // build the 8 x 32 recon rows (A operands)
        const int r = tid >> 4, cg = tid & 15;
#pragma unroll
        for (int mt = 0; mt < MT; ++mt) {
            const int srow = 16 * mt + r, slab = s0 + srow;
            const bool ok = slab < BT;
            float4 vv[HMAX];
            const float4* vp = reinterpret_cast<const float4*>(v + ((size_t)slab * H) * D) + cg;
#pragma unroll
            for (int h = 0; h < HMAX; ++h) vv[h] = (ok && h < H) ? __ldg(vp + h * (D / 4)) : make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
            for (int j = 0; j < NG; ++j) {
                const float4* cj = reinterpret_cast<const float4*>(cs + (size_t)(srow * NG + j) * HP);
                float4 a = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
                for (int q = 0; q < HP / 4; ++q) {
                    const float4 cc = cj[q];
                    if (4 * q < HMAX) fma4(a, cc.x, vv[4 * q]);
                    if (4 * q + 1 < HMAX) fma4(a, cc.y, vv[4 * q + 1]);
                    if (4 * q + 2 < HMAX) fma4(a, cc.z, vv[4 * q + 2]);
                    if (4 * q + 3 < HMAX) fma4(a, cc.w, vv[4 * q + 3]);
                }
                if (WRITE_RECON && ok && n0 + j < N)
                    *reinterpret_cast<float4*>(recon + ((size_t)slab * N + n0 + j) * D + 4 * cg) = a;
                uint2 hi, lo;
                split_h2<PREC_3XTF32>(a.x, a.y, hi.x, lo.x);
                split_h2<PREC_3XTF32>(a.z, a.w, hi.y, lo.y);
                unsigned char* Arow = As + (size_t)(j * ROWS + srow) * ROWB;
                *reinterpret_cast<uint2*>(Arow + 8 * cg) = hi;
                *reinterpret_cast<uint2*>(Arow + LO + 8 * cg) = lo;
            }
        }
    }
    __syncthreads();

    const int n = n0 + warp;
    if (n >= N) return;
    const int g = lane >> 2, tg = lane & 3;
    const uint32_t a_base = smem_u32(As + (size_t)(warp * ROWS + 8 * ((lane >> 3) & 1) + (lane & 7)) * ROWB + 16 * (lane >> 4));
    constexpr float inv = 1.f / WSCALE;
#pragma unroll 1
    for (int half = 0; half < 2; ++half) {
        const uint4* wp = wfrag + ((size_t)n * 2 + half) * (4 * 4 * 32) + lane;
        uint4 cur[4];
#pragma unroll
        for (int j = 0; j < 4; ++j) cur[j] = __ldg(wp + j * 32);
        float acc[MT][4][4];
#pragma unroll
        for (int mt = 0; mt < MT; ++mt)
#pragma unroll
            for (int j = 0; j < 4; ++j) acc[mt][j][0] = acc[mt][j][1] = acc[mt][j][2] = acc[mt][j][3] = 0.f;
#pragma unroll
        for (int kb = 0; kb < 4; ++kb) {
            uint4 nxt[4];
            if (kb < 3) {
#pragma unroll
                for (int j = 0; j < 4; ++j) nxt[j] = __ldg(wp + ((kb + 1) * 4 + j) * 32);
            }
#pragma unroll
            for (int mt = 0; mt < MT; ++mt) {
                uint32_t ah[4], al[4];
                ldsm_x4(ah, a_base + mt * 16 * ROWB + 32 * kb);
                ldsm_x4(al, a_base + mt * 16 * ROWB + 32 * kb + LO);
#pragma unroll
                for (int j = 0; j < 4; ++j) mma3<PREC_3XTF32>(acc[mt][j], ah, al, cur[j].x, cur[j].y, cur[j].z, cur[j].w);
            }
            if (kb < 3) {
#pragma unroll
                for (int j = 0; j < 4; ++j) cur[j] = nxt[j];
            }
        }
        // epilogue: bias + residual + LeakyReLU, accumulator layout (rows g / g+8 of each 16-slab tile, column pairs)
        const float* bn = bias + (size_t)n * D + 32 * half + 2 * tg;
        float2 bj[4];
#pragma unroll
        for (int j = 0; j < 4; ++j) bj[j] = __ldg(reinterpret_cast<const float2*>(bn + 8 * j));
#pragma unroll
        for (int mt = 0; mt < MT; ++mt) {
#pragma unroll
            for (int hr = 0; hr < 2; ++hr) {
                const int slab = s0 + 16 * mt + g + 8 * hr;
                if (slab < BT) {
                    const size_t ro = ((size_t)slab * N + n) * D + 32 * half + 2 * tg;
                    float2 xr[4];
#pragma unroll
                    for (int j = 0; j < 4; ++j) xr[j] = __ldg(reinterpret_cast<const float2*>(x + ro + 8 * j));
#pragma unroll
                    for (int j = 0; j < 4; ++j) {
                        float y0 = fmaf(acc[mt][j][2 * hr], inv, bj[j].x) + xr[j].x;
                        float y1 = fmaf(acc[mt][j][2 * hr + 1], inv, bj[j].y) + xr[j].y;
                        *reinterpret_cast<float2*>(out + ro + 8 * j) = make_float2(lrelu(y0), lrelu(y1));
                    }
                }
            }
        }
    }
}

template <int HMAX>
static int launch(const float* c, const float* v, const float* x, const void* wfrag, const float* bias, float* out, float* recon,
                  int BT, int N, int H, cudaStream_t st) {
    constexpr int HP = (HMAX + 3) / 4 * 4;
    const size_t smem = (size_t)NG * ROWS * ROWB + (size_t)ROWS * NG * HP * 4;
    dim3 grid((N + NG - 1) / NG, (BT + ROWS - 1) / ROWS);
    cudaError_t e;
    if (recon) {
        auto k = cap_recon_proj_kernel<HMAX, true>;
        e = cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) return (int)e;
        launch_pdl(k, dim3(grid), dim3(NTH), smem, st, c, v, x, (const uint4*)wfrag, bias, out, recon, BT, N, H);
    } else {
        auto k = cap_recon_proj_kernel<HMAX, false>;
        e = cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) return (int)e;
        launch_pdl(k, dim3(grid), dim3(NTH), smem, st, c, v, x, (const uint4*)wfrag, bias, out, (float*)nullptr, BT, N, H);
    }
    return (int)cudaGetLastError();
}

}  // namespace crp

// ------------------------------------------------------------------------------------------------------------------
// Inter-cluster hop (GPTST.py:125-134), the only part of the block that mixes the 12 slabs of a sample, one CTA per sample:
//     E1[b] = LReLU(dyn_b (s_b + tau))          (HT x D, K = T*H)
//     v[b,t,h,:] = squash(LReLU(dyn_b[:, (t,h)]^T E1[b]) + s[b,t,h,:])
// 246 k MACs per sample: pure latency, so every global load is in flight before the first one is consumed and the inner
// products carry independent accumulators.  (Folding this into the routing kernel's tail -- the CTA that completes a sample
// runs it -- was measured: +15..28 us on the routing launch, the tails of the last wave run on an otherwise idle GPU; having
// the routing CTAs write their slab's share of the first contraction and a (B, 4)-CTA kernel finish: 57.4 + 6.3 us against
// 52.6 + 12.4 us here -- not worth a second routing flavour.)
// ------------------------------------------------------------------------------------------------------------------
namespace hop {
constexpr int NT = 512, D = 64;
__device__ __forceinline__ void fma4r(float4& a, float m, const float4& e) {
    a.x = fmaf(m, e.x, a.x); a.y = fmaf(m, e.y, a.y); a.z = fmaf(m, e.z, a.z); a.w = fmaf(m, e.w, a.w);
}
__device__ __forceinline__ float4 addt(const float4& a, float t) { return make_float4(a.x + t, a.y + t, a.z + t, a.w + t); }

__global__ void __launch_bounds__(NT) cap_hop_ev_kernel(const float* __restrict__ s, const float* __restrict__ dyn,
                                                         float* __restrict__ e1, float* __restrict__ v, int T, int H, int HT) {
    pdl_enter();
    extern __shared__ __align__(16) unsigned char hsm[];
    const int K = T * H, b = blockIdx.x, tid = threadIdx.x;
    float4* Ss4 = reinterpret_cast<float4*>(hsm);                   // [K][16]  raw s of the sample
    float4* E14 = Ss4 + (size_t)K * 16;                             // [HT][16]
    float* dy = reinterpret_cast<float*>(E14 + (size_t)HT * 16);    // [HT][K]
    float* taus = dy + (size_t)HT * K;                              // [K]
    const float4* sb4 = reinterpret_cast<const float4*>(s + (size_t)b * K * D);
    const float* db = dyn + (size_t)b * HT * K;
    constexpr int UN = 4;
    for (int i0 = 0; i0 < K * 16; i0 += NT * UN) {
        float4 t4[UN];
        float t1[UN];
#pragma unroll
        for (int u = 0; u < UN; ++u) {
            const int i = i0 + u * NT + tid;
            if (i < K * 16) t4[u] = __ldg(sb4 + i);
            if (i < HT * K) t1[u] = __ldg(db + i);
        }
#pragma unroll
        for (int u = 0; u < UN; ++u) {
            const int i = i0 + u * NT + tid;
            if (i < K * 16) Ss4[i] = t4[u];
            if (i < HT * K) dy[i] = t1[u];
        }
    }
    for (int i = K * 16 + tid; i < HT * K; i += NT) dy[i] = __ldg(db + i);     // HT > 16 only
    for (int k = tid; k < K; k += NT) taus[k] = (float)(k / H + 1) / 12.f;
    __syncthreads();
    // E1: item = (ht, 4 columns), two threads per item split the K range (adjacent lanes, combined by one shuffle)
    for (int it = tid; it < HT * 16 * 2; it += NT) {
        const int o = it >> 1, kh = it & 1, ht = o >> 4, cg = o & 15;
        const int kmid = (K + 1) >> 1, k0 = kh ? kmid : 0, k1 = kh ? K : kmid;
        const float* dr = dy + (size_t)ht * K;
        float4 a0 = make_float4(0.f, 0.f, 0.f, 0.f), a1 = a0;
        int k = k0;
#pragma unroll 4
        for (; k + 1 < k1; k += 2) {
            fma4r(a0, dr[k], addt(Ss4[k * 16 + cg], taus[k]));
            fma4r(a1, dr[k + 1], addt(Ss4[(k + 1) * 16 + cg], taus[k + 1]));
        }
        if (k < k1) fma4r(a0, dr[k], addt(Ss4[k * 16 + cg], taus[k]));
        float4 a = make_float4(a0.x + a1.x, a0.y + a1.y, a0.z + a1.z, a0.w + a1.w);
        a.x += __shfl_xor_sync(0xffffffffu, a.x, 1); a.y += __shfl_xor_sync(0xffffffffu, a.y, 1);
        a.z += __shfl_xor_sync(0xffffffffu, a.z, 1); a.w += __shfl_xor_sync(0xffffffffu, a.w, 1);
        if (kh == 0) {
            a.x = lrelu(a.x); a.y = lrelu(a.y); a.z = lrelu(a.z); a.w = lrelu(a.w);
            E14[o] = a;
            reinterpret_cast<float4*>(e1 + (size_t)b * HT * D)[o] = a;
        }
    }
    __syncthreads();
    // v: item = (row k = (t,h), 4 columns); the 16 lanes of a row reduce the squared norm with four shuffles
    for (int i = tid; i < K * 16; i += NT) {
        const int k = i >> 4, cg = i & 15;
        float4 a0 = make_float4(0.f, 0.f, 0.f, 0.f), a1 = a0;
        int ht = 0;
#pragma unroll 4
        for (; ht + 1 < HT; ht += 2) {
            fma4r(a0, dy[ht * K + k], E14[ht * 16 + cg]);
            fma4r(a1, dy[(ht + 1) * K + k], E14[(ht + 1) * 16 + cg]);
        }
        if (ht < HT) fma4r(a0, dy[ht * K + k], E14[ht * 16 + cg]);
        const float4 sv = Ss4[i];
        float4 r;
        r.x = lrelu(a0.x + a1.x) + sv.x; r.y = lrelu(a0.y + a1.y) + sv.y; r.z = lrelu(a0.z + a1.z) + sv.z; r.w = lrelu(a0.w + a1.w) + sv.w;
        float q = (r.x * r.x + r.y * r.y) + (r.z * r.z + r.w * r.w);
#pragma unroll
        for (int o = 8; o > 0; o >>= 1) q += __shfl_xor_sync(0xffffffffu, q, o);
        const float f = squash_f(q);
        reinterpret_cast<float4*>(v + (size_t)b * K * D)[i] = make_float4(r.x * f, r.y * f, r.z * f, r.w * f);
    }
}

}  // namespace hop

}  // namespace gptst

// c (B,T,H,N), v (B,T,H,D), x (B,T,N,D); wfrag = gptst_hypertem_pack_w(W_n, wfrag, NULL, N) (fragment-ordered fp16 hi/lo of the
// (N,D,D) [in][out] node-adaptive weights); bias (N,D); out (B,T,N,D); recon (B,T,N,D) is written when non-NULL (kept for backward).
extern "C" int gptst_cap_recon_proj(const float* c, const float* v, const float* x, const void* wfrag, const float* bias, float* out,
                                    float* recon, int B, int T, int N, int D, int H, void* stream) {
    if (!c || !v || !x || !wfrag || !bias || !out || B <= 0 || T <= 0 || N <= 0) return -1;
    if (D != 64 || H < 1 || H > 15) return -2;
    cudaStream_t st = (cudaStream_t)stream;
    if (H <= 10) return gptst::crp::launch<10>(c, v, x, wfrag, bias, out, recon, B * T, N, H, st);
    return gptst::crp::launch<15>(c, v, x, wfrag, bias, out, recon, B * T, N, H, st);
}

// s (B,T,H,D), dyn (B,HT,T*H) -> e1 (B,HT,D), v (B,T,H,D); one launch, one CTA per sample
extern "C" int gptst_cap_hop_ev(const float* s, const float* dyn, float* e1, float* v, int B, int T, int D, int H, int HT, void* stream) {
    if (!s || !dyn || !e1 || !v || B <= 0 || T <= 0 || HT <= 0) return -1;
    if (D != 64 || H < 1 || H > 15 || (T * H) % 2 != 0) return -2;
    const size_t K = (size_t)T * H;
    const size_t smem = (K * 16 + (size_t)HT * 16) * 16 + ((size_t)HT * K + K) * 4;
    if (smem > 227 * 1024) return -2;
    cudaError_t e = cudaFuncSetAttribute(gptst::hop::cap_hop_ev_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return (int)e;
    gptst::launch_pdl(gptst::hop::cap_hop_ev_kernel, dim3(B), dim3(gptst::hop::NT), smem, (cudaStream_t)stream, s, dyn, e1, v, T, H, HT);
    return (int)cudaGetLastError();
}
