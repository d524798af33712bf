// cap backward, first piece, second generation for D = 64, N <= 256:
//     dv[h,:]  = sum_n c[h,n] drecon[n,:]        (H x D, an aggregation over the nodes)
//     dcr[h,n] = v[h,:] . drecon[n,:]            (H x N, a logit-type product)
// Same decomposition as cap_route2_fwd.cu: one CTA per (b,t) slab, one warp per 16 nodes, fp16-split mma.sync m16n8k16
// with ldmatrix operands.  drecon is a gradient of arbitrary magnitude, so every warp first brings ITS 16 rows into fp16
// range with a power-of-two scale taken from their max |.| (exact, undone on the fp32 results); c and v are bounded by 1.
// Replaces cap_dv_dcr_kernel (FMA + one shuffle reduction per (h, node), 69 us at B=64/N=170).
#include "cap_common.cuh"
#include "mma_f16.cuh"

namespace gptst {
namespace r2 {

using namespace hf;

template <int NW, int MINB, int PREC>
__global__ void __launch_bounds__(NW * 32, MINB)
cap_dv_dcr2_kernel(const float* __restrict__ c, const float* __restrict__ v, const float* __restrict__ drecon,
                   float* __restrict__ dv, float* __restrict__ dcr, int N, int H, const float* __restrict__ s,
                   const float* __restrict__ dyn, const float* __restrict__ e1, float* __restrict__ dr_out,
                   float* __restrict__ dpre2_out, int T, int HT) {
    constexpr int D = 64;
    extern __shared__ __align__(128) unsigned char smraw[];
    unsigned char* Grow = smraw;                                   // [NW*16][ROWB] drecon rows: fp32, then hi|lo planes
    unsigned char* vpl = Grow + (size_t)NW * 16 * ROWB;            // [16][ROWB]
    float* red = reinterpret_cast<float*>(vpl + 16 * ROWB);        // [NW][H][REDLD]

    constexpr int NT = NW * 32;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int g = lane >> 2, t = lane & 3;
    const int slab = blockIdx.x;
    const float* gs = drecon + (size_t)slab * N * D;
    for (int i = tid; i < NW * 16 * 16; i += NT) {
        const int r = i >> 4, ch = i & 15;
        unsigned char* dst = Grow + (size_t)r * ROWB + ch * 16;
        if (r < N) cp_async16(dst, gs + (size_t)r * D + ch * 4);
        else *reinterpret_cast<float4*>(dst) = make_float4(0.f, 0.f, 0.f, 0.f);
    }
    // v -> hi|lo planes (rows >= H zero)
    for (int i = tid; i < 16 * (D / 2); i += NT) {
        const int h = i / (D / 2), p = i % (D / 2);
        uint32_t hi = 0u, lo = 0u;
        if (h < H) {
            const float2 vv = *reinterpret_cast<const float2*>(v + ((size_t)slab * H + h) * D + 2 * p);
            split_h2<PREC>(vv.x, vv.y, hi, lo);
        }
        *reinterpret_cast<uint32_t*>(vpl + (size_t)h * ROWB + p * 4) = hi;
        *reinterpret_cast<uint32_t*>(vpl + (size_t)h * ROWB + LO + p * 4) = lo;
    }
    const int n0 = warp * 16;
    const int h0 = g, h1 = g + 8;
    const int na = n0 + 2 * t, nb = na + 8;
    float cc[8];
    {
        const float* cs = c + (size_t)slab * H * N;
#pragma unroll
        for (int i = 0; i < 8; ++i) {
            const int h = (i & 2) ? h1 : h0;
            const int n = ((i & 4) ? nb : na) + (i & 1);
            cc[i] = (h < H && n < N) ? cs[(size_t)h * N + n] : 0.f;
        }
    }
    cp_async_wait_all();
    __syncthreads();

    // ---- this warp's 16 rows: max |.|, power-of-two scale, fp16 hi|lo planes in place
    float2 sc;
    {
        float4 f[8];
        float m = 0.f;
#pragma unroll
        for (int k = 0; k < 8; ++k) {
            const int id = lane + 32 * k, r = id >> 4, ch = id & 15;
            f[k] = *reinterpret_cast<const float4*>(Grow + (size_t)(n0 + r) * ROWB + ch * 16);
            m = fmaxf(m, fmaxf(fmaxf(fabsf(f[k].x), fabsf(f[k].y)), fmaxf(fabsf(f[k].z), fabsf(f[k].w))));
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, o));
        sc = pow2_scale_for_fp16(m);
        __syncwarp();
#pragma unroll
        for (int k = 0; k < 8; ++k) {
            const int id = lane + 32 * k, r = id >> 4, ch = id & 15;
            uint32_t hi0, lo0, hi1, lo1;
            split_h2<PREC>(f[k].x * sc.x, f[k].y * sc.x, hi0, lo0);
            split_h2<PREC>(f[k].z * sc.x, f[k].w * sc.x, hi1, lo1);
            unsigned char* row = Grow + (size_t)(n0 + r) * ROWB + ch * 8;
            *reinterpret_cast<uint2*>(row) = make_uint2(hi0, hi1);
            *reinterpret_cast<uint2*>(row + LO) = make_uint2(lo0, lo1);
        }
        __syncwarp();
    }
    // ---- dcr = v . drecon^T for the warp's nodes
    {
        float z[8];
#pragma unroll
        for (int i = 0; i < 8; ++i) z[i] = 0.f;
        warp_logits<PREC>(z, vpl, Grow, n0, lane, sc.y);
        float* dc = dcr + (size_t)slab * H * N;
#pragma unroll
        for (int i = 0; i < 8; ++i) {
            const int h = (i & 2) ? h1 : h0;
            const int n = ((i & 4) ? nb : na) + (i & 1);
            if (h < H && n < N) dc[(size_t)h * N + n] = z[i];
        }
    }
    // ---- dv = c . drecon: per-warp partial, then the deterministic cross-warp sum
    warp_aggregate<PREC>(cc, H, Grow, n0, red + (size_t)warp * H * REDLD, lane, sc.y);
    __syncthreads();
    const int b = slab / (T > 0 ? T : 1), tt = slab % (T > 0 ? T : 1);
    for (int h = warp; h < H; h += NW) {
        float2 acc = make_float2(0.f, 0.f);
#pragma unroll
        for (int w = 0; w < NW; ++w) {
            const float2 p = *reinterpret_cast<const float2*>(red + ((size_t)w * H + h) * REDLD + 2 * lane);
            acc.x += p.x; acc.y += p.y;
        }
        const size_t row = ((size_t)slab * H + h) * D + 2 * lane;
        if (dv) *reinterpret_cast<float2*>(dv + row) = acc;
        if (dr_out) {
            // row pass of the hop backward (cap_hop_bwd_rows_kernel) on the row this warp just finished:
            // pre2 = dyn_b[:, t-block]^T E1 ; r = phi(pre2) + s ; dr = squash'(r, dv) ; dpre2 = dr * phi'(pre2)
            float2 p2 = make_float2(0.f, 0.f);
            const int K = T * H;
            for (int ht = 0; ht < HT; ++ht) {
                const float dd = dyn[((size_t)b * HT + ht) * K + tt * H + h];
                const float2 ee = *reinterpret_cast<const float2*>(e1 + ((size_t)b * HT + ht) * D + 2 * lane);
                p2.x = fmaf(dd, ee.x, p2.x); p2.y = fmaf(dd, ee.y, p2.y);
            }
            const float2 sv = *reinterpret_cast<const float2*>(s + row);
            const float2 r = make_float2(lrelu(p2.x) + sv.x, lrelu(p2.y) + sv.y);
            const float q = warp_sum(r.x * r.x + r.y * r.y);
            const float rg = warp_sum(r.x * acc.x + r.y * acc.y);
            const float f = squash_f(q), fp = squash_df(q);
            const float2 dr = make_float2(f * acc.x + 2.f * r.x * fp * rg, f * acc.y + 2.f * r.y * fp * rg);
            *reinterpret_cast<float2*>(dr_out + row) = dr;
            *reinterpret_cast<float2*>(dpre2_out + row) = make_float2(lrelu_grad(p2.x, dr.x), lrelu_grad(p2.y, dr.y));
        }
    }
}

template <int NW, int MINB, int PREC>
static cudaError_t launch_dvdcr(const float* c, const float* v, const float* drecon, float* dv, float* dcr, int BT, int N,
                                int H, const float* s, const float* dyn, const float* e1, float* dr, float* dp2, int T, int HT,
                                cudaStream_t st) {
    const size_t smem = (size_t)NW * 16 * ROWB + 16 * ROWB + (size_t)NW * H * REDLD * 4;
    auto kern = cap_dv_dcr2_kernel<NW, MINB, PREC>;
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return e;
    kern<<<BT, NW * 32, smem, st>>>(c, v, drecon, dv, dcr, N, H, s, dyn, e1, dr, dp2, T, HT);
    return cudaGetLastError();
}

template <int PREC>
static cudaError_t dispatch_dvdcr(const float* c, const float* v, const float* drecon, float* dv, float* dcr, int BT, int N,
                                  int H, const float* s, const float* dyn, const float* e1, float* dr, float* dp2, int T, int HT,
                                  cudaStream_t st) {
#define DVDCR_ARGS c, v, drecon, dv, dcr, BT, N, H, s, dyn, e1, dr, dp2, T, HT, st
    if (N <= 64) return launch_dvdcr<4, 4, PREC>(DVDCR_ARGS);
    if (N <= 128) return launch_dvdcr<8, 3, PREC>(DVDCR_ARGS);
    if (N <= 176) return launch_dvdcr<11, 2, PREC>(DVDCR_ARGS);
    if (N <= 208) return launch_dvdcr<13, 2, PREC>(DVDCR_ARGS);
    return launch_dvdcr<16, 1, PREC>(DVDCR_ARGS);
#undef DVDCR_ARGS
}

}  // namespace r2

// always the three-term split: this product is far from being the bottleneck of the backward pass
cudaError_t dv_dcr2(const float* c, const float* v, const float* drecon, float* dv, float* dcr, int BT, int N, int H,
                    cudaStream_t st) {
    return r2::dispatch_dvdcr<PREC_3XTF32>(c, v, drecon, dv, dcr, BT, N, H, nullptr, nullptr, nullptr, nullptr, nullptr, 1, 0, st);
}

bool route2_supported(int N, int D, int H);

}  // namespace gptst

using namespace gptst;

// dv/dcr fused with the row pass of the hop backward: dcr (B,T,H,N) and, instead of dv, directly
// dr = squash'(r, dv) and dpre2 = dr * phi'(pre2) (both (B,T,H,D)) -- the inputs of gptst_cap_hop_bwd_cols.  D = 64, N <= 256.
extern "C" int gptst_cap_dv_dcr_hoprows(const float* c, const float* v, const float* drecon, const float* s, const float* dyn,
                                        const float* e1, float* dcr, float* dr, float* dpre2, int B, int T, int N, int D,
                                        int H, int HT, void* stream) {
    if (!c || !v || !drecon || !s || !dyn || !e1 || !dcr || !dr || !dpre2 || B <= 0 || T <= 0 || N <= 0 || HT <= 0) return -1;
    if (!route2_supported(N, D, H)) return -2;
    return (int)r2::dispatch_dvdcr<PREC_3XTF32>(c, v, drecon, nullptr, dcr, B * T, N, H, s, dyn, e1, dr, dpre2, T, HT,
                                                (cudaStream_t)stream);
}
