// Eval-path pieces next to the encoder, fp32 FMA kernels:
//
//  * STGCN's GLU temporal convolution (reference model/STGCN/stgcn.py:25-53, TemporalConvLayer act = "GLU", with the
//    Align of stgcn.py:10-23 folded in):
//        conv = Conv2d(c_in, 2 c_out, (kt, 1), padding (kt-1)/2)(x)            x : (B, c_in, T, N), N fastest
//        out  = (conv[:, :c_out] + align(x)) * sigmoid(conv[:, c_out:])         out : (B, c_out, T, N)
//    One CTA = (sample, 32 consecutive nodes): the whole (c_in, T + 2 pad, 32) input tile is staged in shared memory
//    once (every global row is one coalesced 128-byte line), a warp owns output channels, a lane a node; the kt taps
//    slide over the time steps held in registers, so each staged value feeds 2 x kt x OPW FMAs per shared-memory read.
//    The same kernel without the gate (GLU = false) is a plain temporal convolution: the backward uses it for
//    dx = conv_transpose(dconv) with the flipped / transposed weights prepared by the caller.
//  * glu_gate_bwd      : (dout, P, S) -> dconv = [dout * S ; dout * P * S * (1 - S)]          (elementwise)
//  * glu_conv_dw       : dW[o, i, k] = sum_{b,t,n} dconv[b,o,t,n] x[b,i,t+k-pad,n], db[o] = sum dconv   (partials per split)
//  * gate_blend / gate_bwd : the fusion gate of model/Model.py:12-17 as elementwise kernels (forward for widths the fused
//    epilogue of gptst_gate_fwd does not cover; backward for both).
#include "common.cuh"

namespace gptst {
namespace glu {

constexpr int TMAX = kMaxT;      // 12 time steps (the reference hard-wires input_window = 12)
constexpr int NTILE = 32;        // nodes per CTA = lanes
constexpr int NWARP = 8;

__device__ __forceinline__ float sigmoidf_(float v) { return 1.f / (1.f + expf(-v)); }

// align_mode: 0 = identity (c_in == c_out), 1 = zero-padded channels (c_in < c_out), 2 = 1x1 conv (aw, ab), 3 = none
template <int KT, bool GLU>
__global__ void __launch_bounds__(NWARP * 32)
tconv_fwd_kernel(const float* __restrict__ x, const float* __restrict__ W, const float* __restrict__ bias,
                 const float* __restrict__ aw, const float* __restrict__ ab, float* __restrict__ out,
                 float* __restrict__ Psave, float* __restrict__ Ssave, int CinTot, int c0, int Cin, int Cout, int T, int N,
                 int align_mode, int accumulate) {
    // input channels [c0, c0 + Cin) of CinTot are staged and contracted by this launch; accumulate: out += (plain mode only)
    extern __shared__ __align__(16) float xs[];            // [Cin][T + KT - 1][32]
    constexpr int PAD = (KT - 1) / 2;
    const int TP = T + KT - 1;
    const int b = blockIdx.y, n0 = blockIdx.x * NTILE;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int n = n0 + lane;
    const float* xb = x + ((size_t)b * CinTot + c0) * T * N;
    for (int r = warp; r < Cin * TP; r += NWARP) {
        const int i = r / TP, tp = r - i * TP, t = tp - PAD;
        xs[r * NTILE + lane] = (t >= 0 && t < T && n < N) ? xb[((size_t)i * T + t) * N + n] : 0.f;
    }
    __syncthreads();
    for (int o0 = warp * 2; o0 < Cout; o0 += NWARP * 2) {
        const bool two = o0 + 1 < Cout;
        const int o1 = two ? o0 + 1 : o0;
        float aP[2][TMAX], aQ[2][TMAX];
        {
            float bp0 = (bias && !accumulate) ? bias[o0] : 0.f, bp1 = (bias && !accumulate) ? bias[o1] : 0.f;
            if (GLU && align_mode == 2) { bp0 += ab[o0]; bp1 += ab[o1]; }
            const float bq0 = (GLU && bias) ? bias[Cout + o0] : 0.f, bq1 = (GLU && bias) ? bias[Cout + o1] : 0.f;
#pragma unroll
            for (int t = 0; t < TMAX; ++t) { aP[0][t] = bp0; aP[1][t] = bp1; aQ[0][t] = bq0; aQ[1][t] = bq1; }
        }
        const float* wp0 = W + ((size_t)o0 * CinTot + c0) * KT;
        const float* wp1 = W + ((size_t)o1 * CinTot + c0) * KT;
        const float* wq0 = W + ((size_t)(Cout + o0) * CinTot + c0) * KT;
        const float* wq1 = W + ((size_t)(Cout + o1) * CinTot + c0) * KT;
        for (int i = 0; i < Cin; ++i) {
            float xv[TMAX + KT - 1];
            const float* xr = xs + (size_t)i * TP * NTILE + lane;
#pragma unroll
            for (int tp = 0; tp < TMAX + KT - 1; ++tp) xv[tp] = tp < TP ? xr[tp * NTILE] : 0.f;
#pragma unroll
            for (int k = 0; k < KT; ++k) {
                float w0 = __ldg(wp0 + i * KT + k), w1 = __ldg(wp1 + i * KT + k);
                if (GLU && align_mode == 2 && k == PAD) { w0 += __ldg(aw + (size_t)o0 * CinTot + c0 + i); w1 += __ldg(aw + (size_t)o1 * CinTot + c0 + i); }
#pragma unroll
                for (int t = 0; t < TMAX; ++t) { aP[0][t] = fmaf(w0, xv[t + k], aP[0][t]); aP[1][t] = fmaf(w1, xv[t + k], aP[1][t]); }
                if (GLU) {
                    const float q0 = __ldg(wq0 + i * KT + k), q1 = __ldg(wq1 + i * KT + k);
#pragma unroll
                    for (int t = 0; t < TMAX; ++t) { aQ[0][t] = fmaf(q0, xv[t + k], aQ[0][t]); aQ[1][t] = fmaf(q1, xv[t + k], aQ[1][t]); }
                }
            }
        }
        if (n < N) {
#pragma unroll
            for (int u = 0; u < 2; ++u) {
                if (u == 1 && !two) break;
                const int o = u ? o1 : o0;
                const bool res = GLU && (align_mode == 0 || (align_mode == 1 && o < Cin));
#pragma unroll
                for (int t = 0; t < TMAX; ++t) {
                    if (t < T) {
                        const size_t idx = (((size_t)b * Cout + o) * T + t) * N + n;
                        float p = aP[u][t];
                        if (res) p += xs[((size_t)o * TP + t + PAD) * NTILE + lane];
                        if (GLU) {
                            const float s = sigmoidf_(aQ[u][t]);
                            if (Psave) { Psave[idx] = p; Ssave[idx] = s; }
                            out[idx] = p * s;
                        } else {
                            out[idx] = accumulate ? out[idx] + p : p;
                        }
                    }
                }
            }
        }
    }
}

// dconv (B, 2 Cout, T, N) from dout, P, S (B, Cout, T, N)
__global__ void __launch_bounds__(256) glu_gate_bwd_kernel(const float* __restrict__ dout, const float* __restrict__ P,
                                                           const float* __restrict__ S, float* __restrict__ dconv, long per_b,
                                                           long total) {
    for (long i = (long)blockIdx.x * 256 + threadIdx.x; i < total; i += (long)gridDim.x * 256) {
        const long b = i / per_b, r = i - b * per_b;
        const float g = dout[i], p = P[i], s = S[i];
        dconv[b * 2 * per_b + r] = g * s;
        dconv[b * 2 * per_b + per_b + r] = g * p * s * (1.f - s);
    }
}

// dW / db partials.  grid (ceil(C2 / 8), ceil(Cin / IC), splits); warp = channel o of dconv, lane = node; the thread keeps
// IC x KT accumulators over the (sample, node tile) positions of its split and reduces them over the lanes once at the end.
constexpr int IC = 8;
template <int KT>
__global__ void __launch_bounds__(NWARP * 32)
tconv_dw_kernel(const float* __restrict__ dconv, const float* __restrict__ x, float* __restrict__ dWp, float* __restrict__ dbp,
                int B, int C2, int Cin, int T, int N, int tiles_per_split) {
    constexpr int PAD = (KT - 1) / 2;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int o = blockIdx.x * NWARP + warp, i0 = blockIdx.y * IC, split = blockIdx.z;
    const int ntile = (N + NTILE - 1) / NTILE, total = B * ntile;
    float acc[IC][KT], accb = 0.f;
#pragma unroll
    for (int a = 0; a < IC; ++a)
#pragma unroll
        for (int k = 0; k < KT; ++k) acc[a][k] = 0.f;
    if (o < C2) {
        const int p0 = split * tiles_per_split, p1 = min(total, p0 + tiles_per_split);
        for (int p = p0; p < p1; ++p) {
            const int b = p / ntile, n = (p - b * ntile) * NTILE + lane;
            if (n >= N) continue;
            float g[TMAX];
            const float* gp = dconv + (((size_t)b * C2 + o) * T) * N + n;
#pragma unroll
            for (int t = 0; t < TMAX; ++t) g[t] = t < T ? gp[(size_t)t * N] : 0.f;
            if (blockIdx.y == 0) {
#pragma unroll
                for (int t = 0; t < TMAX; ++t) accb += g[t];
            }
#pragma unroll
            for (int a = 0; a < IC; ++a) {
                if (i0 + a < Cin) {
                    float xv[TMAX + KT - 1];
                    const float* xp = x + (((size_t)b * Cin + i0 + a) * T) * N + n;
#pragma unroll
                    for (int tp = 0; tp < TMAX + KT - 1; ++tp) {
                        const int t = tp - PAD;
                        xv[tp] = (t >= 0 && t < T) ? xp[(size_t)t * N] : 0.f;
                    }
#pragma unroll
                    for (int k = 0; k < KT; ++k)
#pragma unroll
                        for (int t = 0; t < TMAX; ++t) acc[a][k] = fmaf(g[t], xv[t + k], acc[a][k]);
                }
            }
        }
    }
#pragma unroll
    for (int a = 0; a < IC; ++a)
#pragma unroll
        for (int k = 0; k < KT; ++k) acc[a][k] = warp_sum(acc[a][k]);
    accb = warp_sum(accb);
    if (lane == 0 && o < C2) {
        float* dw = dWp + ((size_t)split * C2 + o) * Cin * KT;
#pragma unroll
        for (int a = 0; a < IC; ++a)
            if (i0 + a < Cin)
#pragma unroll
                for (int k = 0; k < KT; ++k) dw[(size_t)(i0 + a) * KT + k] = acc[a][k];
        if (blockIdx.y == 0) dbp[(size_t)split * C2 + o] = accb;
    }
}

// fusion gate, elementwise:  z = sigmoid(xs + xt) ; h = z * x + (1 - z) * y
__global__ void __launch_bounds__(256) gate_blend_kernel(const float4* __restrict__ xs, const float4* __restrict__ xt,
                                                         const float4* __restrict__ x, const float4* __restrict__ y,
                                                         float4* __restrict__ h, float4* __restrict__ z, long n4) {
    for (long i = (long)blockIdx.x * 256 + threadIdx.x; i < n4; i += (long)gridDim.x * 256) {
        const float4 a = xs[i], b = xt[i], f = x[i], t = y[i];
        float4 zz, hh;
        zz.x = sigmoidf_(a.x + b.x); zz.y = sigmoidf_(a.y + b.y); zz.z = sigmoidf_(a.z + b.z); zz.w = sigmoidf_(a.w + b.w);
        hh.x = zz.x * f.x + (1.f - zz.x) * t.x; hh.y = zz.y * f.y + (1.f - zz.y) * t.y;
        hh.z = zz.z * f.z + (1.f - zz.z) * t.z; hh.w = zz.w * f.w + (1.f - zz.w) * t.w;
        h[i] = hh;
        if (z) z[i] = zz;
    }
}
// dh, z, x, y -> dpre = dh (x - y) z (1 - z) ; dx = dh z ; dy = dh (1 - z)
__global__ void __launch_bounds__(256) gate_bwd_kernel(const float4* __restrict__ dh, const float4* __restrict__ z,
                                                       const float4* __restrict__ x, const float4* __restrict__ y,
                                                       float4* __restrict__ dpre, float4* __restrict__ dx, float4* __restrict__ dy,
                                                       long n4) {
    for (long i = (long)blockIdx.x * 256 + threadIdx.x; i < n4; i += (long)gridDim.x * 256) {
        const float4 g = dh[i], zz = z[i], f = x[i], t = y[i];
        float4 p, a, b;
        p.x = g.x * (f.x - t.x) * zz.x * (1.f - zz.x); a.x = g.x * zz.x; b.x = g.x * (1.f - zz.x);
        p.y = g.y * (f.y - t.y) * zz.y * (1.f - zz.y); a.y = g.y * zz.y; b.y = g.y * (1.f - zz.y);
        p.z = g.z * (f.z - t.z) * zz.z * (1.f - zz.z); a.z = g.z * zz.z; b.z = g.z * (1.f - zz.z);
        p.w = g.w * (f.w - t.w) * zz.w * (1.f - zz.w); a.w = g.w * zz.w; b.w = g.w * (1.f - zz.w);
        dpre[i] = p; dx[i] = a; dy[i] = b;
    }
}

template <int KT, bool GLU>
static int launch_fwd(const float* x, const float* W, const float* bias, const float* aw, const float* ab, float* out, float* P,
                      float* S, int B, int Cin, int Cout, int T, int N, int align_mode, cudaStream_t st) {
    const size_t per_ch = (size_t)(T + KT - 1) * NTILE * 4;
    const int cmax = (int)((size_t)227 * 1024 / per_ch);
    if (GLU && Cin > cmax) return -2;                  // the gated form keeps every input channel resident (residual, saves)
    auto kern = tconv_fwd_kernel<KT, GLU>;
    for (int c0 = 0; c0 < Cin; c0 += cmax) {           // plain mode: wider inputs are contracted chunk by chunk, out accumulated
        const int cc = Cin - c0 < cmax ? Cin - c0 : cmax;
        const size_t smem = per_ch * cc;
        cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) return (int)e;
        kern<<<dim3((N + NTILE - 1) / NTILE, B), NWARP * 32, smem, st>>>(x, W, bias, aw, ab, out, P, S, Cin, c0, cc, Cout, T, N,
                                                                        align_mode, c0 > 0);
        e = cudaGetLastError();
        if (e != cudaSuccess) return (int)e;
    }
    return 0;
}

}  // namespace glu
}  // namespace gptst

using namespace gptst;

// out = (conv(x)[:, :Cout] + align(x)) * sigmoid(conv(x)[:, Cout:]).  W (2 Cout, Cin, kt) = Conv2d weight with its trailing 1 dropped,
// bias (2 Cout); aw (Cout, Cin), ab (Cout) only when Cin > Cout (stgcn.py:15-16).  P, S (B, Cout, T, N): optional saves for the backward.
extern "C" int gptst_glu_tconv_fwd(const float* x, const float* W, const float* bias, const float* aw, const float* ab, float* out,
                                   float* P, float* S, int B, int Cin, int Cout, int T, int N, int kt, void* stream) {
    if (!x || !W || !bias || !out || B <= 0 || Cin <= 0 || Cout <= 0 || N <= 0 || (P == nullptr) != (S == nullptr)) return -1;
    if (T < 1 || T > glu::TMAX || kt < 1 || (kt & 1) == 0) return -2;     // an even kt changes the length: the reference's add fails too
    const int mode = Cin == Cout ? 0 : (Cin < Cout ? 1 : 2);
    if (mode == 2 && (!aw || !ab)) return -1;
    cudaStream_t st = (cudaStream_t)stream;
    switch (kt) {
        case 1: return glu::launch_fwd<1, true>(x, W, bias, aw, ab, out, P, S, B, Cin, Cout, T, N, mode, st);
        case 3: return glu::launch_fwd<3, true>(x, W, bias, aw, ab, out, P, S, B, Cin, Cout, T, N, mode, st);
        case 5: return glu::launch_fwd<5, true>(x, W, bias, aw, ab, out, P, S, B, Cin, Cout, T, N, mode, st);
        case 7: return glu::launch_fwd<7, true>(x, W, bias, aw, ab, out, P, S, B, Cin, Cout, T, N, mode, st);
    }
    return -2;
}

// plain temporal convolution out[b,o,t,n] = sum_{i,k} W[o,i,k] x[b,i,t+k-pad,n] (+ bias[o] when given), "same" zero padding
extern "C" int gptst_tconv_fwd(const float* x, const float* W, const float* bias, float* out, int B, int Cin, int Cout, int T, int N,
                               int kt, void* stream) {
    if (!x || !W || !out || B <= 0 || Cin <= 0 || Cout <= 0 || N <= 0) return -1;
    if (T < 1 || T > glu::TMAX || kt < 1 || (kt & 1) == 0) return -2;
    cudaStream_t st = (cudaStream_t)stream;
    switch (kt) {
        case 1: return glu::launch_fwd<1, false>(x, W, bias, nullptr, nullptr, out, nullptr, nullptr, B, Cin, Cout, T, N, 3, st);
        case 3: return glu::launch_fwd<3, false>(x, W, bias, nullptr, nullptr, out, nullptr, nullptr, B, Cin, Cout, T, N, 3, st);
        case 5: return glu::launch_fwd<5, false>(x, W, bias, nullptr, nullptr, out, nullptr, nullptr, B, Cin, Cout, T, N, 3, st);
        case 7: return glu::launch_fwd<7, false>(x, W, bias, nullptr, nullptr, out, nullptr, nullptr, B, Cin, Cout, T, N, 3, st);
    }
    return -2;
}

extern "C" int gptst_glu_gate_bwd(const float* dout, const float* P, const float* S, float* dconv, int B, int Cout, int T, int N,
                                  void* stream) {
    if (!dout || !P || !S || !dconv || B <= 0 || Cout <= 0 || T <= 0 || N <= 0) return -1;
    const long per_b = (long)Cout * T * N, total = per_b * B;
    long blocks = (total + 255) / 256;
    if (blocks > 148 * 16) blocks = 148 * 16;
    glu::glu_gate_bwd_kernel<<<(unsigned)blocks, 256, 0, (cudaStream_t)stream>>>(dout, P, S, dconv, per_b, total);
    return (int)cudaGetLastError();
}

// number of (dW, db) partials gptst_tconv_dw writes for this geometry
extern "C" int gptst_tconv_dw_splits(int B, int C2, int Cin, int N) {
    const int tiles = B * ((N + glu::NTILE - 1) / glu::NTILE);
    const int ctas = ((C2 + glu::NWARP - 1) / glu::NWARP) * ((Cin + glu::IC - 1) / glu::IC);
    int s = (148 * 8 + ctas - 1) / ctas;
    if (s > tiles) s = tiles;
    return s < 1 ? 1 : s;
}
// dW_part (splits, C2, Cin, kt), db_part (splits, C2): partial sums of dconv (B, C2, T, N) against the zero-padded x (B, Cin, T, N)
extern "C" int gptst_tconv_dw(const float* dconv, const float* x, float* dW_part, float* db_part, int B, int C2, int Cin, int T, int N,
                              int kt, int splits, void* stream) {
    if (!dconv || !x || !dW_part || !db_part || B <= 0 || C2 <= 0 || Cin <= 0 || N <= 0 || splits <= 0) return -1;
    if (T < 1 || T > glu::TMAX || kt < 1 || (kt & 1) == 0) return -2;
    const int tiles = B * ((N + glu::NTILE - 1) / glu::NTILE);
    const int tps = (tiles + splits - 1) / splits;
    dim3 grid((C2 + glu::NWARP - 1) / glu::NWARP, (Cin + glu::IC - 1) / glu::IC, splits);
    cudaStream_t st = (cudaStream_t)stream;
    switch (kt) {
        case 1: glu::tconv_dw_kernel<1><<<grid, glu::NWARP * 32, 0, st>>>(dconv, x, dW_part, db_part, B, C2, Cin, T, N, tps); break;
        case 3: glu::tconv_dw_kernel<3><<<grid, glu::NWARP * 32, 0, st>>>(dconv, x, dW_part, db_part, B, C2, Cin, T, N, tps); break;
        case 5: glu::tconv_dw_kernel<5><<<grid, glu::NWARP * 32, 0, st>>>(dconv, x, dW_part, db_part, B, C2, Cin, T, N, tps); break;
        case 7: glu::tconv_dw_kernel<7><<<grid, glu::NWARP * 32, 0, st>>>(dconv, x, dW_part, db_part, B, C2, Cin, T, N, tps); break;
        default: return -2;
    }
    return (int)cudaGetLastError();
}

// z = sigmoid(xs + xt), h = z * x + (1 - z) * y on n elements (n % 4 == 0); z optional
extern "C" int gptst_gate_blend(const float* xs, const float* xt, const float* x, const float* y, float* h, float* z, long n,
                                void* stream) {
    if (!xs || !xt || !x || !y || !h || n <= 0) return -1;
    if (n % 4) return -2;
    long blocks = (n / 4 + 255) / 256;
    if (blocks > 148 * 16) blocks = 148 * 16;
    glu::gate_blend_kernel<<<(unsigned)blocks, 256, 0, (cudaStream_t)stream>>>(
        reinterpret_cast<const float4*>(xs), reinterpret_cast<const float4*>(xt), reinterpret_cast<const float4*>(x),
        reinterpret_cast<const float4*>(y), reinterpret_cast<float4*>(h), reinterpret_cast<float4*>(z), n / 4);
    return (int)cudaGetLastError();
}
extern "C" int gptst_gate_bwd(const float* dh, const float* z, const float* x, const float* y, float* dpre, float* dx, float* dy, long n,
                              void* stream) {
    if (!dh || !z || !x || !y || !dpre || !dx || !dy || n <= 0) return -1;
    if (n % 4) return -2;
    long blocks = (n / 4 + 255) / 256;
    if (blocks > 148 * 16) blocks = 148 * 16;
    glu::gate_bwd_kernel<<<(unsigned)blocks, 256, 0, (cudaStream_t)stream>>>(
        reinterpret_cast<const float4*>(dh), reinterpret_cast<const float4*>(z), reinterpret_cast<const float4*>(x),
        reinterpret_cast<const float4*>(y), reinterpret_cast<float4*>(dpre), reinterpret_cast<float4*>(dx), reinterpret_cast<float4*>(dy),
        n / 4);
    return (int)cudaGetLastError();
}
