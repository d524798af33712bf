// Output projection of the decoder, dim_flow_out = nn.Linear(D, ibd) with ibd = 1..4 output features (reference
// GPTST.py:454-458):   y[r][o] = <x[r,:], W[o,:]> + b[o].
//
// As library calls this is a GEMV forward and, backward, a K = 1 GEMM (dX), a split-K GEMM with K = B*T*N plus its
// reduction (dW) and a column sum (db): ~100 us at the head of the backward main chain of a PEMS08 step.  Both directions
// are pure streaming passes over x (HBM bound, 4*rows*D bytes forward, 8*rows*D backward):
//   proj_out_fwd : D/4 lanes per row (one float4 each), shuffle reduction, 4 rows of loads in flight per lane;
//   proj_out_bwd : one pass that writes dX[r,:] = sum_o dy[r][o] W[o,:] and accumulates the CTA's partial of
//                  dW[o,:] = sum_r dy[r][o] x[r,:] and db[o] = sum_r dy[r][o] over its contiguous row range
//                  (fixed reduction order, no atomics: deterministic).  part: (parts, O*D + O), summed by the caller.
#include "common.cuh"

namespace gptst {
namespace po {

constexpr int kMaxO = 4;

template <int D>
__global__ void __launch_bounds__(256) proj_out_fwd_kernel(const float* __restrict__ x, const float* __restrict__ W,
                                                           const float* __restrict__ b, float* __restrict__ y, long rows, int O) {
    constexpr int LPR = D / 4;            // lanes per row
    constexpr int RPC = 256 / LPR;        // rows per CTA pass
    const int tid = threadIdx.x, q = tid % LPR, rl = tid / LPR;
    float4 w[kMaxO];
    float bo[kMaxO];
#pragma unroll
    for (int o = 0; o < kMaxO; ++o) {
        w[o] = (o < O) ? *reinterpret_cast<const float4*>(W + (size_t)o * D + 4 * q) : make_float4(0.f, 0.f, 0.f, 0.f);
        bo[o] = (o < O) ? b[o] : 0.f;
    }
    const long step = (long)gridDim.x * RPC;
    // the loop bound is CTA-uniform (the shuffles below need converged warps); rows past the end are guarded per access
    for (long b0 = (long)blockIdx.x * RPC; b0 < rows; b0 += 4 * step) {
        const long base = b0 + rl;
        float4 xv[4];
#pragma unroll
        for (int u = 0; u < 4; ++u) {
            const long r = base + u * step;
            xv[u] = (r < rows) ? *reinterpret_cast<const float4*>(x + r * D + 4 * q) : make_float4(0.f, 0.f, 0.f, 0.f);
        }
#pragma unroll
        for (int u = 0; u < 4; ++u) {
            const long r = base + u * step;
#pragma unroll
            for (int o = 0; o < kMaxO; ++o) {
                if (o < O) {       // uniform over the grid
                    float a = fmaf(xv[u].x, w[o].x, fmaf(xv[u].y, w[o].y, fmaf(xv[u].z, w[o].z, xv[u].w * w[o].w)));
#pragma unroll
                    for (int s = LPR / 2; s > 0; s >>= 1) a += __shfl_xor_sync(0xffffffffu, a, s);
                    if (q == 0 && r < rows) y[r * O + o] = a + bo[o];
                }
            }
        }
    }
}

template <int D>
__global__ void __launch_bounds__(256) proj_out_bwd_kernel(const float* __restrict__ dy, const float* __restrict__ x,
                                                           const float* __restrict__ W, float* __restrict__ dX,
                                                           float* __restrict__ part, long rows, int O, long rows_per_cta) {
    constexpr int LPR = D / 4;
    constexpr int RL = 256 / LPR;         // row lanes
    __shared__ __align__(16) float red[RL][kMaxO * D];
    __shared__ float redb[RL][kMaxO];
    const int tid = threadIdx.x, q = tid % LPR, rl = tid / LPR;
    float4 w[kMaxO], aw[kMaxO];
    float ab[kMaxO];
#pragma unroll
    for (int o = 0; o < kMaxO; ++o) {
        w[o] = (o < O) ? *reinterpret_cast<const float4*>(W + (size_t)o * D + 4 * q) : make_float4(0.f, 0.f, 0.f, 0.f);
        aw[o] = make_float4(0.f, 0.f, 0.f, 0.f);
        ab[o] = 0.f;
    }
    const long r0 = (long)blockIdx.x * rows_per_cta;
    long r1 = r0 + rows_per_cta;
    if (r1 > rows) r1 = rows;
    for (long base = r0 + rl; base < r1; base += 4 * RL) {
        float4 xv[4];
        float g[4][kMaxO];
#pragma unroll
        for (int u = 0; u < 4; ++u) {
            const long r = base + u * RL;
            const bool on = r < r1;
            xv[u] = on ? *reinterpret_cast<const float4*>(x + r * D + 4 * q) : make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
            for (int o = 0; o < kMaxO; ++o) g[u][o] = (on && o < O) ? dy[r * O + o] : 0.f;
        }
#pragma unroll
        for (int u = 0; u < 4; ++u) {
            const long r = base + u * RL;
            float4 d = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
            for (int o = 0; o < kMaxO; ++o) {
                const float gv = g[u][o];
                d.x = fmaf(gv, w[o].x, d.x); d.y = fmaf(gv, w[o].y, d.y); d.z = fmaf(gv, w[o].z, d.z); d.w = fmaf(gv, w[o].w, d.w);
                aw[o].x = fmaf(gv, xv[u].x, aw[o].x); aw[o].y = fmaf(gv, xv[u].y, aw[o].y);
                aw[o].z = fmaf(gv, xv[u].z, aw[o].z); aw[o].w = fmaf(gv, xv[u].w, aw[o].w);
                ab[o] += gv;
            }
            if (dX && r < r1) *reinterpret_cast<float4*>(dX + r * D + 4 * q) = d;
        }
    }
#pragma unroll
    for (int o = 0; o < kMaxO; ++o) {
        *reinterpret_cast<float4*>(&red[rl][o * D + 4 * q]) = aw[o];
        if (q == 0) redb[rl][o] = ab[o];
    }
    __syncthreads();
    float* po = part + (size_t)blockIdx.x * ((size_t)O * D + O);
    for (int i = tid; i < O * D; i += 256) {
        float s = 0.f;
#pragma unroll
        for (int k = 0; k < RL; ++k) s += red[k][i];
        po[i] = s;
    }
    if (tid < O) {
        float s = 0.f;
#pragma unroll
        for (int k = 0; k < RL; ++k) s += redb[k][tid];
        po[(size_t)O * D + tid] = s;
    }
}

}  // namespace po
}  // namespace gptst

using namespace gptst;

// x (rows, D), W (O, D) as nn.Linear stores it, b (O), y (rows, O).  D = 64 or 128, 1 <= O <= 4.
extern "C" int gptst_proj_out_fwd(const float* x, const float* W, const float* b, float* y, long rows, int D, int O, void* stream) {
    if (!x || !W || !b || !y || rows <= 0) return -1;
    if ((D != 64 && D != 128) || O < 1 || O > po::kMaxO) return -2;
    const int rpc = 256 / (D / 4);
    long blocks = (rows + 4L * rpc - 1) / (4L * rpc);
    if (blocks > 148 * 8) blocks = 148 * 8;
    if (D == 64) po::proj_out_fwd_kernel<64><<<(unsigned)blocks, 256, 0, (cudaStream_t)stream>>>(x, W, b, y, rows, O);
    else po::proj_out_fwd_kernel<128><<<(unsigned)blocks, 256, 0, (cudaStream_t)stream>>>(x, W, b, y, rows, O);
    return (int)cudaGetLastError();
}

extern "C" int gptst_proj_out_bwd_parts(long rows) {
    long want = 2 * 148;
    if (want > (rows + 63) / 64) want = (rows + 63) / 64;
    return (int)(want < 1 ? 1 : want);
}

// dy (rows, O), dX (rows, D) or NULL, part (parts, O*D + O): [p][o*D + d] = partial dW[o][d], [p][O*D + o] = partial db[o]
extern "C" int gptst_proj_out_bwd(const float* dy, const float* x, const float* W, float* dX, float* part, long rows, int D, int O,
                                  int parts, void* stream) {
    if (!dy || !x || !W || !part || rows <= 0 || parts <= 0) return -1;
    if ((D != 64 && D != 128) || O < 1 || O > po::kMaxO) return -2;
    const long rpc = (rows + parts - 1) / parts;
    if (D == 64) po::proj_out_bwd_kernel<64><<<parts, 256, 0, (cudaStream_t)stream>>>(dy, x, W, dX, part, rows, O, rpc);
    else po::proj_out_bwd_kernel<128><<<parts, 256, 0, (cudaStream_t)stream>>>(dy, x, W, dX, part, rows, O, rpc);
    return (int)cudaGetLastError();
}
