// Small parameter-side pieces of the pre-training step that sat on the critical tail of the captured graph as chains of
// tiny library kernels (reference GPTST.py:187-219 time_feature / time_feature_spg, and the one-feature input embeddings
// dim_in_flow / MLP_RL.ln1, GPTST.py:22, :298).
//
//   time_mlp_fwd / _bwd : the 5-linear time-embedding MLP   h0 = a Wd^T + bd + b Ww^T + bw ;  z1 = h0 W1^T + b1 ;
//                         z2 = relu(z1) W2^T + b2 ;  out = relu(z2) W3^T + b3      on R rows (R = B*T or B), F inputs per
//                         branch (1 or 12), e hidden units (16 or 4).  Forward: one launch (was ~12); backward: one launch
//                         that writes per-row-chunk partials of every parameter gradient (was ~25 launches).
//   affine1_bwd         : y = x w + b with one input feature: dw = sum_i dy[i,:] x[i], db = sum_i dy[i,:] in one pass over
//                         dy (row-range partials, summed by the caller).
// Everything is deterministic (fixed partial order, no floating-point atomics).
#include "common.cuh"

namespace gptst {
namespace sm {

constexpr int kE = 16;       // max hidden width
constexpr int kF = 12;       // max inputs per branch
constexpr int kRows = 32;    // rows per CTA

struct MlpParams {
    const float *Wd, *bd, *Ww, *bw, *W1, *b1, *W2, *b2, *W3, *b3;
};

// grid = ceil(R / kRows), block = kRows * kE threads: thread = (row, unit)
__global__ void __launch_bounds__(kRows* kE) time_mlp_fwd_kernel(const float* __restrict__ a, const float* __restrict__ b, MlpParams p,
                                                                float* __restrict__ h0, float* __restrict__ z1,
                                                                float* __restrict__ z2, float* __restrict__ out, int R, int F,
                                                                int e, long in_stride) {
    __shared__ float W[3][kE * kE], Wi[2][kE * kF], bias[5][kE], act[kRows][kE + 1];
    const int tid = threadIdx.x;
    for (int i = tid; i < e * e; i += blockDim.x) { W[0][i] = p.W1[i]; W[1][i] = p.W2[i]; W[2][i] = p.W3[i]; }
    for (int i = tid; i < e * F; i += blockDim.x) { Wi[0][i] = p.Wd[i]; Wi[1][i] = p.Ww[i]; }
    if (tid < e) { bias[0][tid] = p.bd[tid]; bias[1][tid] = p.bw[tid]; bias[2][tid] = p.b1[tid]; bias[3][tid] = p.b2[tid]; bias[4][tid] = p.b3[tid]; }
    __syncthreads();
    const int rl = tid / kE, u = tid % kE;
    const int r = blockIdx.x * kRows + rl;
    const bool on = r < R && u < e;
    float v = 0.f;
    if (on) {
        v = bias[0][u] + bias[1][u];
        for (int f = 0; f < F; ++f) v = fmaf(a[(long)r * in_stride + f], Wi[0][u * F + f], fmaf(b[(long)r * in_stride + f], Wi[1][u * F + f], v));
        h0[(long)r * e + u] = v;
    }
    act[rl][u] = v;
    __syncthreads();
    float y = 0.f;
    if (on) {
        y = bias[2][u];
        for (int k = 0; k < e; ++k) y = fmaf(act[rl][k], W[0][u * e + k], y);
        z1[(long)r * e + u] = y;
    }
    __syncthreads();
    act[rl][u] = fmaxf(y, 0.f);
    __syncthreads();
    if (on) {
        y = bias[3][u];
        for (int k = 0; k < e; ++k) y = fmaf(act[rl][k], W[1][u * e + k], y);
        z2[(long)r * e + u] = y;
    }
    __syncthreads();
    act[rl][u] = fmaxf(y, 0.f);
    __syncthreads();
    if (on) {
        y = bias[4][u];
        for (int k = 0; k < e; ++k) y = fmaf(act[rl][k], W[2][u * e + k], y);
        out[(long)r * e + u] = y;
    }
}

// packed gradient layout (per row chunk): dW3 e*e | db3 e | dW2 e*e | db2 e | dW1 e*e | db1 e | dWd e*F | dbd e | dWw e*F | dbw e
__host__ __device__ inline int mlp_grad_floats(int e, int F) { return 3 * (e * e + e) + 2 * (e * F + e); }

__global__ void __launch_bounds__(kRows* kE) time_mlp_bwd_kernel(const float* __restrict__ a, const float* __restrict__ b, MlpParams p,
                                                                const float* __restrict__ h0, const float* __restrict__ z1,
                                                                const float* __restrict__ z2, const float* __restrict__ g,
                                                                float* __restrict__ part, int R, int F, int e, long in_stride) {
    __shared__ float W[3][kE * kE];
    __shared__ float G[kRows][kE + 1], A2[kRows][kE + 1], D2[kRows][kE + 1], A1[kRows][kE + 1], D1[kRows][kE + 1], H0[kRows][kE + 1],
        DH[kRows][kE + 1];
    __shared__ float Ain[kRows][kF], Bin[kRows][kF];
    const int tid = threadIdx.x;
    for (int i = tid; i < e * e; i += blockDim.x) { W[0][i] = p.W1[i]; W[1][i] = p.W2[i]; W[2][i] = p.W3[i]; }
    const int rl = tid / kE, u = tid % kE;
    const int r = blockIdx.x * kRows + rl;
    const bool on = r < R && u < e;
    float zz1 = 0.f, zz2 = 0.f;
    if (on) { zz1 = z1[(long)r * e + u]; zz2 = z2[(long)r * e + u]; }
    G[rl][u] = on ? g[(long)r * e + u] : 0.f;
    A2[rl][u] = fmaxf(zz2, 0.f);
    A1[rl][u] = fmaxf(zz1, 0.f);
    H0[rl][u] = on ? h0[(long)r * e + u] : 0.f;
    if (u < F) {
        Ain[rl][u] = (r < R) ? a[(long)r * in_stride + u] : 0.f;
        Bin[rl][u] = (r < R) ? b[(long)r * in_stride + u] : 0.f;
    }
    __syncthreads();
    // dz2 = (g W3) * (z2 > 0)
    float d = 0.f;
    if (on) { for (int k = 0; k < e; ++k) d = fmaf(G[rl][k], W[2][k * e + u], d); d = zz2 > 0.f ? d : 0.f; }
    D2[rl][u] = d;
    __syncthreads();
    d = 0.f;
    if (on) { for (int k = 0; k < e; ++k) d = fmaf(D2[rl][k], W[1][k * e + u], d); d = zz1 > 0.f ? d : 0.f; }
    D1[rl][u] = d;
    __syncthreads();
    d = 0.f;
    if (on) for (int k = 0; k < e; ++k) d = fmaf(D1[rl][k], W[0][k * e + u], d);
    DH[rl][u] = d;
    __syncthreads();
    // parameter-gradient partials of this row chunk: one output element per thread (looped), fixed row order
    float* out = part + (size_t)blockIdx.x * mlp_grad_floats(e, F);
    const int nW = e * e, nI = e * F;
    const int total = mlp_grad_floats(e, F);
    for (int i = tid; i < total; i += blockDim.x) {
        int o = i;
        float s = 0.f;
        // segment decoding
        const float (*L)[kE + 1] = nullptr;      // left factor (delta), indexed [row][unit i]
        const float (*Rt)[kE + 1] = nullptr;     // right factor (activation), indexed [row][unit j]
        int mode = -1, ii = 0, jj = 0;           // 0: e x e outer product, 1: bias, 2: input a, 3: input b
        if (o < nW) { L = G; Rt = A2; mode = 0; ii = o / e; jj = o % e; }
        else if ((o -= nW) < e) { L = G; mode = 1; ii = o; }
        else if ((o -= e) < nW) { L = D2; Rt = A1; mode = 0; ii = o / e; jj = o % e; }
        else if ((o -= nW) < e) { L = D2; mode = 1; ii = o; }
        else if ((o -= e) < nW) { L = D1; Rt = H0; mode = 0; ii = o / e; jj = o % e; }
        else if ((o -= nW) < e) { L = D1; mode = 1; ii = o; }
        else if ((o -= e) < nI) { L = DH; mode = 2; ii = o / F; jj = o % F; }
        else if ((o -= nI) < e) { L = DH; mode = 1; ii = o; }
        else if ((o -= e) < nI) { L = DH; mode = 3; ii = o / F; jj = o % F; }
        else { o -= nI; L = DH; mode = 1; ii = o; }
        for (int rr = 0; rr < kRows; ++rr) {
            const float l = L[rr][ii];
            const float rv = (mode == 0) ? Rt[rr][jj] : (mode == 1) ? 1.f : (mode == 2) ? Ain[rr][jj] : Bin[rr][jj];
            s = fmaf(l, rv, s);
        }
        out[i] = s;
    }
}

// dw/db partials of y = x w + b (one input feature): grid CTAs over row ranges, 256 threads = 4 row lanes x 64 columns
__global__ void __launch_bounds__(256) affine1_bwd_kernel(const float* __restrict__ dy, const float* __restrict__ x,
                                                          float* __restrict__ part, long n, int D, long rows_per_cta) {
    __shared__ float red[2][4][128];
    const int tid = threadIdx.x, lane_r = tid / 64, c = tid % 64;
    const long r0 = (long)blockIdx.x * rows_per_cta;
    long r1 = r0 + rows_per_cta;
    if (r1 > n) r1 = n;
    for (int c0 = 0; c0 < D; c0 += 64) {
        float sw = 0.f, sb = 0.f;
        if (c0 + c < D) {
            for (long r = r0 + lane_r; r < r1; r += 4) {
                const float v = dy[r * D + c0 + c];
                sw = fmaf(v, x[r], sw);
                sb += v;
            }
        }
        red[0][lane_r][c] = sw;
        red[1][lane_r][c] = sb;
        __syncthreads();
        if (lane_r == 0 && c0 + c < D) {
            part[((size_t)blockIdx.x * 2 + 0) * D + c0 + c] = (red[0][0][c] + red[0][1][c]) + (red[0][2][c] + red[0][3][c]);
            part[((size_t)blockIdx.x * 2 + 1) * D + c0 + c] = (red[1][0][c] + red[1][1][c]) + (red[1][2][c] + red[1][3][c]);
        }
        __syncthreads();
    }
}

}  // namespace sm
}  // namespace gptst

using namespace gptst;

extern "C" int gptst_time_mlp_chunks(int R) { return (R + sm::kRows - 1) / sm::kRows; }
extern "C" int gptst_time_mlp_grad_floats(int e, int F) { return sm::mlp_grad_floats(e, F); }

// a, b: (R, F) with row stride in_stride floats; weights as nn.Linear stores them ([out][in]); h0, z1, z2, out: (R, e)
extern "C" int gptst_time_mlp_fwd(const float* a, const float* b, const float* Wd, const float* bd, const float* Ww,
                                  const float* bw, const float* W1, const float* b1, const float* W2, const float* b2,
                                  const float* W3, const float* b3, float* h0, float* z1, float* z2, float* out, int R, int F,
                                  int e, long in_stride, void* stream) {
    if (!a || !b || !Wd || !bd || !Ww || !bw || !W1 || !b1 || !W2 || !b2 || !W3 || !b3 || !h0 || !z1 || !z2 || !out || R <= 0) return -1;
    if (e < 1 || e > sm::kE || F < 1 || F > sm::kF) return -2;
    sm::MlpParams p{Wd, bd, Ww, bw, W1, b1, W2, b2, W3, b3};
    sm::time_mlp_fwd_kernel<<<gptst_time_mlp_chunks(R), sm::kRows * sm::kE, 0, (cudaStream_t)stream>>>(a, b, p, h0, z1, z2, out, R, F,
                                                                                                     e, in_stride);
    return (int)cudaGetLastError();
}

// g: (R, e) upstream gradient; part: (gptst_time_mlp_chunks(R), gptst_time_mlp_grad_floats(e, F)) partials, summed by the caller
extern "C" int gptst_time_mlp_bwd(const float* a, const float* b, const float* W1, const float* W2, const float* W3,
                                  const float* h0, const float* z1, const float* z2, const float* g, float* part, int R, int F,
                                  int e, long in_stride, void* stream) {
    if (!a || !b || !W1 || !W2 || !W3 || !h0 || !z1 || !z2 || !g || !part || R <= 0) return -1;
    if (e < 1 || e > sm::kE || F < 1 || F > sm::kF) return -2;
    sm::MlpParams p{nullptr, nullptr, nullptr, nullptr, W1, nullptr, W2, nullptr, W3, nullptr};
    sm::time_mlp_bwd_kernel<<<gptst_time_mlp_chunks(R), sm::kRows * sm::kE, 0, (cudaStream_t)stream>>>(a, b, p, h0, z1, z2, g, part, R,
                                                                                                     F, e, in_stride);
    return (int)cudaGetLastError();
}

extern "C" int gptst_affine1_bwd_parts(long n) {
    long want = 2 * 148;
    if (want > (n + 255) / 256) want = (n + 255) / 256;
    return (int)(want < 1 ? 1 : want);
}
// part: (parts, 2, D): [p][0] = partial dw, [p][1] = partial db
extern "C" int gptst_affine1_bwd(const float* dy, const float* x, float* part, long n, int D, int parts, void* stream) {
    if (!dy || !x || !part || n <= 0 || parts <= 0) return -1;
    if (D < 1 || D > 1024) return -2;
    const long rpc = (n + parts - 1) / parts;
    sm::affine1_bwd_kernel<<<parts, 256, 0, (cudaStream_t)stream>>>(dy, x, part, n, D, rpc);
    return (int)cudaGetLastError();
}

// ------------------------------------------------------------------------------------------------------------------
// Backward of the low-rank table generators  Tab = te . pool  (te (G,d), pool (d,C), d <= 16): the adaptive weights
// W_bt / W_n, their biases, the incidence logits dadj and the inter-cluster adjacency dyn are all of this form
// (GPTST.py:104, :129, :137-138, :160-161, :24-31).  cuBLAS runs these skinny products (K = 768 or 170, M = 16) with SIMT
// split-K kernels at 15-90 us each; as plain streaming kernels they are a few microseconds:
//     dpool[k][c] = sum_g te[g][k] dTab[g][c]      (column chunks of 32, 8 row lanes, te staged in shared memory)
//     dte[g][k]   = sum_c dTab[g][c] pool[k][c]    (4 rows per CTA, threads stride over the columns)
// ------------------------------------------------------------------------------------------------------------------
namespace gptst {
namespace sm {

constexpr int kDmax = 16;

// 1024 threads = 32 columns x 32 row lanes (the first version had 8 row lanes: ~7 warps per SM, latency-bound at 24-37 us)
__global__ void __launch_bounds__(1024) table_dpool_kernel(const float* __restrict__ te, const float* __restrict__ dtab,
                                                           float* __restrict__ dpool, int G, int d, int C) {
    extern __shared__ __align__(16) float tes[];   // [G][16] (rows padded to 16 floats), then red[32][32][17]
    float* red = tes + (size_t)G * kDmax;
    const int tid = threadIdx.x, cl = tid & 31, rl = tid >> 5;
    for (int i = tid; i < G * kDmax; i += 1024) {
        const int g = i / kDmax, k = i % kDmax;
        tes[i] = (k < d) ? te[(size_t)g * d + k] : 0.f;
    }
    __syncthreads();
    const int c = blockIdx.x * 32 + cl;
    float acc[kDmax];
#pragma unroll
    for (int k = 0; k < kDmax; ++k) acc[k] = 0.f;
    if (c < C) {
#pragma unroll 4
        for (int g = rl; g < G; g += 32) {
            const float v = dtab[(size_t)g * C + c];
            const float4* t4 = reinterpret_cast<const float4*>(tes + (size_t)g * kDmax);
#pragma unroll
            for (int q = 0; q < kDmax / 4; ++q) {
                const float4 t = t4[q];
                acc[4 * q] = fmaf(t.x, v, acc[4 * q]); acc[4 * q + 1] = fmaf(t.y, v, acc[4 * q + 1]);
                acc[4 * q + 2] = fmaf(t.z, v, acc[4 * q + 2]); acc[4 * q + 3] = fmaf(t.w, v, acc[4 * q + 3]);
            }
        }
    }
#pragma unroll
    for (int k = 0; k < kDmax; ++k) red[((size_t)rl * 32 + cl) * (kDmax + 1) + k] = acc[k];
    __syncthreads();
    for (int i = tid; i < 32 * d; i += 1024) {
        const int k = i / 32, cc = i % 32;
        if (blockIdx.x * 32 + cc < C) {
            float s = 0.f;
#pragma unroll 8
            for (int r = 0; r < 32; ++r) s += red[((size_t)r * 32 + cc) * (kDmax + 1) + k];
            dpool[(size_t)k * C + blockIdx.x * 32 + cc] = s;
        }
    }
}

__global__ void __launch_bounds__(256) table_dte_kernel(const float* __restrict__ pool, const float* __restrict__ dtab,
                                                        float* __restrict__ dte, int G, int d, int C) {
    __shared__ float red[8][4][kDmax];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int g0 = blockIdx.x * 4;
    float acc[4][kDmax];
#pragma unroll
    for (int r = 0; r < 4; ++r)
#pragma unroll
        for (int k = 0; k < kDmax; ++k) acc[r][k] = 0.f;
    for (int c = tid; c < C; c += 256) {
        float v[4];
#pragma unroll
        for (int r = 0; r < 4; ++r) v[r] = (g0 + r < G) ? dtab[(size_t)(g0 + r) * C + c] : 0.f;
#pragma unroll
        for (int k = 0; k < kDmax; ++k) {
            if (k < d) {
                const float p = pool[(size_t)k * C + c];
#pragma unroll
                for (int r = 0; r < 4; ++r) acc[r][k] = fmaf(v[r], p, acc[r][k]);
            }
        }
    }
#pragma unroll
    for (int r = 0; r < 4; ++r)
#pragma unroll
        for (int k = 0; k < kDmax; ++k) {
            float s = acc[r][k];
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
            if (lane == 0) red[warp][r][k] = s;
        }
    __syncthreads();
    if (tid < 4 * kDmax) {
        const int r = tid / kDmax, k = tid % kDmax;
        if (g0 + r < G && k < d) {
            float s = 0.f;
#pragma unroll
            for (int w = 0; w < 8; ++w) s += red[w][r][k];
            dte[(size_t)(g0 + r) * d + k] = s;
        }
    }
}

}  // namespace sm
}  // namespace gptst

// te (G,d), pool (d,C), dtab (G,C) -> dpool (d,C), dte (G,d); either output may be NULL
extern "C" int gptst_table_bwd(const float* te, const float* pool, const float* dtab, float* dpool, float* dte, int G, int d,
                               int C, void* stream) {
    if (!te || !pool || !dtab || G <= 0 || C <= 0) return -1;
    if (d < 1 || d > gptst::sm::kDmax || (size_t)G * gptst::sm::kDmax * 4 > 140 * 1024) return -2;
    cudaStream_t st = (cudaStream_t)stream;
    if (dpool) {
        const size_t smem = ((size_t)G * gptst::sm::kDmax + 32 * 32 * (gptst::sm::kDmax + 1)) * 4;
        cudaError_t e = cudaFuncSetAttribute(gptst::sm::table_dpool_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) return (int)e;
        gptst::sm::table_dpool_kernel<<<(C + 31) / 32, 1024, smem, st>>>(te, dtab, dpool, G, d, C);
    }
    if (dte) gptst::sm::table_dte_kernel<<<(G + 3) / 4, 256, 0, st>>>(pool, dtab, dte, G, d, C);
    return (int)cudaGetLastError();
}

// ------------------------------------------------------------------------------------------------------------------
// Fused backward of a low-rank table on the tensor cores: dTab is read ONCE.  CTA tile = 64 rows x 128 columns of dTab
// (cp.async into shared memory); 3xTF32 mma.sync m16n8k8 (truncating split, common.cuh):
//     dte partial  (rows of the tile, over its 128 columns) : M = 64 rows, N = 16, K = 128 -> dte_part[col chunk][G][16]
//     dpool partial (columns of the tile, over its 64 rows) : M = 16,      N = 128, K = 64 -> dpool_part[row chunk][16][C]
// The caller sums the partials (fixed order).  pool is re-read once per 64 rows (was once per 4), te once per 128 columns.
// ------------------------------------------------------------------------------------------------------------------
namespace gptst {
namespace sm {

constexpr int TR = 64, TC = 128, LDT = TC + 4, LDE = 20;   // LDT % 32 == 4, LDE: rows 2t hit banks 8t (+g)

__global__ void __launch_bounds__(256) table_bwd_fused_kernel(const float* __restrict__ te, const float* __restrict__ pool,
                                                              const float* __restrict__ dtab, float* __restrict__ dpool_part,
                                                              float* __restrict__ dte_part, int G, int d, int C) {
    extern __shared__ __align__(16) float smf[];
    float* tile = smf;                          // [TR][LDT]
    float* tes = tile + TR * LDT;               // [TR][LDE]  te rows of this chunk (columns >= d zero)
    float* red = tes + TR * LDE;                // [4][16][17] second k-half of the dte product
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int g = lane >> 2, t = lane & 3;
    const int r0 = blockIdx.x * TR, c0 = blockIdx.y * TC;
    const bool vec_ok = (C % 4 == 0);
    for (int i = tid; i < TR * (TC / 4); i += 256) {
        const int r = i / (TC / 4), q = i % (TC / 4);
        float* dst = tile + r * LDT + 4 * q;
        const int gr = r0 + r, gc = c0 + 4 * q;
        if (gr < G && gc + 3 < C && vec_ok) {
            asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"((uint32_t)__cvta_generic_to_shared(dst)),
                         "l"(dtab + (size_t)gr * C + gc) : "memory");
        } else {
#pragma unroll
            for (int e = 0; e < 4; ++e) dst[e] = (gr < G && gc + e < C) ? dtab[(size_t)gr * C + gc + e] : 0.f;
        }
    }
    for (int i = tid; i < TR * 16; i += 256) {
        const int r = i >> 4, k = i & 15;
        tes[r * LDE + k] = (r0 + r < G && k < d) ? te[(size_t)(r0 + r) * d + k] : 0.f;
    }
    asm volatile("cp.async.commit_group;\n\tcp.async.wait_group 0;" ::: "memory");
    __syncthreads();

    // ---- dte partial: warp = (m-tile mt = warp & 3 : rows 16mt.., k-half kh = warp >> 2 : columns 64kh..)
    {
        const int mt = warp & 3, kh = warp >> 2;
        float acc[2][4];
#pragma unroll
        for (int j = 0; j < 2; ++j) acc[j][0] = acc[j][1] = acc[j][2] = acc[j][3] = 0.f;
        const float* arow0 = tile + (16 * mt + g) * LDT + 64 * kh;
        const float* arow1 = arow0 + 8 * LDT;
#pragma unroll
        for (int ks = 0; ks < 8; ++ks) {
            const int k0 = 8 * ks;
            uint32_t ah[4], al[4];
            split_tf32<PREC_3XTF32>(arow0[k0 + t], ah[0], al[0]);
            split_tf32<PREC_3XTF32>(arow1[k0 + t], ah[1], al[1]);
            split_tf32<PREC_3XTF32>(arow0[k0 + t + 4], ah[2], al[2]);
            split_tf32<PREC_3XTF32>(arow1[k0 + t + 4], ah[3], al[3]);
            const int gc = c0 + 64 * kh + k0 + t;
#pragma unroll
            for (int j = 0; j < 2; ++j) {
                const int kk = 8 * j + g;
                const float p0 = (kk < d && gc < C) ? pool[(size_t)kk * C + gc] : 0.f;
                const float p1 = (kk < d && gc + 4 < C) ? pool[(size_t)kk * C + gc + 4] : 0.f;
                uint32_t bh[2], bl[2];
                split_tf32<PREC_3XTF32>(p0, bh[0], bl[0]);
                split_tf32<PREC_3XTF32>(p1, bh[1], bl[1]);
                mma_split<PREC_3XTF32>(acc[j], ah, al, bh, bl);
            }
        }
        // C fragment: (row g / g+8 of the m-tile, kk = 8j + 2t, 2t+1)
        if (kh == 1) {
#pragma unroll
            for (int j = 0; j < 2; ++j) {
                red[(mt * 16 + g) * 17 + 8 * j + 2 * t] = acc[j][0];
                red[(mt * 16 + g) * 17 + 8 * j + 2 * t + 1] = acc[j][1];
                red[(mt * 16 + g + 8) * 17 + 8 * j + 2 * t] = acc[j][2];
                red[(mt * 16 + g + 8) * 17 + 8 * j + 2 * t + 1] = acc[j][3];
            }
        }
        __syncthreads();
        if (kh == 0) {
            float* out = dte_part + (size_t)blockIdx.y * G * 16;
#pragma unroll
            for (int j = 0; j < 2; ++j) {
#pragma unroll
                for (int h = 0; h < 2; ++h) {
                    const int rl = mt * 16 + g + 8 * h, gr = r0 + rl;
                    if (gr < G) {
                        out[(size_t)gr * 16 + 8 * j + 2 * t] = acc[j][2 * h] + red[rl * 17 + 8 * j + 2 * t];
                        out[(size_t)gr * 16 + 8 * j + 2 * t + 1] = acc[j][2 * h + 1] + red[rl * 17 + 8 * j + 2 * t + 1];
                    }
                }
            }
        }
    }
    // ---- dpool partial: warp owns column tiles 2*warp, 2*warp+1 (8 columns each); K = the tile's 64 rows, k index of an
    //      8-step permuted as {t, t+4} <-> rows {2t, 2t+1} so that the B reads of the fp32 tile are conflict-free
    {
        float acc[2][4];
#pragma unroll
        for (int j = 0; j < 2; ++j) acc[j][0] = acc[j][1] = acc[j][2] = acc[j][3] = 0.f;
#pragma unroll
        for (int ks = 0; ks < 8; ++ks) {
            const int ra = 8 * ks + 2 * t, rb = ra + 1;
            uint32_t ah[4], al[4];
            split_tf32<PREC_3XTF32>(tes[ra * LDE + g], ah[0], al[0]);         // (m = kk g,     k = t)
            split_tf32<PREC_3XTF32>(tes[ra * LDE + g + 8], ah[1], al[1]);     // (m = kk g + 8, k = t)
            split_tf32<PREC_3XTF32>(tes[rb * LDE + g], ah[2], al[2]);         // (m = kk g,     k = t + 4)
            split_tf32<PREC_3XTF32>(tes[rb * LDE + g + 8], ah[3], al[3]);
#pragma unroll
            for (int j = 0; j < 2; ++j) {
                const int cl = 8 * (2 * warp + j) + g;
                uint32_t bh[2], bl[2];
                split_tf32<PREC_3XTF32>(tile[ra * LDT + cl], bh[0], bl[0]);
                split_tf32<PREC_3XTF32>(tile[rb * LDT + cl], bh[1], bl[1]);
                mma_split<PREC_3XTF32>(acc[j], ah, al, bh, bl);
            }
        }
        float* out = dpool_part + (size_t)blockIdx.x * 16 * C;
#pragma unroll
        for (int j = 0; j < 2; ++j) {
            const int gc = c0 + 8 * (2 * warp + j) + 2 * t;      // C fragment: (kk = g / g+8, column 2t, 2t+1 of the tile)
            if (gc < C) { out[(size_t)g * C + gc] = acc[j][0]; out[(size_t)(g + 8) * C + gc] = acc[j][2]; }
            if (gc + 1 < C) { out[(size_t)g * C + gc + 1] = acc[j][1]; out[(size_t)(g + 8) * C + gc + 1] = acc[j][3]; }
        }
    }
}

}  // namespace sm
}  // namespace gptst

extern "C" int gptst_table_bwd2_chunks(int G, int C, int* row_chunks, int* col_chunks) {
    if (!row_chunks || !col_chunks) return -1;
    *row_chunks = (G + gptst::sm::TR - 1) / gptst::sm::TR;
    *col_chunks = (C + gptst::sm::TC - 1) / gptst::sm::TC;
    return 0;
}
// dpool_part: (row_chunks, 16, C), dte_part: (col_chunks, G, 16); rows kk >= d of dpool_part and columns >= d of dte_part are zero
extern "C" int gptst_table_bwd2(const float* te, const float* pool, const float* dtab, float* dpool_part, float* dte_part, int G,
                                int d, int C, void* stream) {
    if (!te || !pool || !dtab || !dpool_part || !dte_part || G <= 0 || C <= 0) return -1;
    if (d < 1 || d > 16) return -2;
    using namespace gptst::sm;
    const size_t smem = ((size_t)TR * LDT + TR * LDE + 4 * 16 * 17) * 4;
    cudaError_t e = cudaFuncSetAttribute(table_bwd_fused_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return (int)e;
    dim3 grid((G + TR - 1) / TR, (C + TC - 1) / TC);
    table_bwd_fused_kernel<<<grid, 256, smem, (cudaStream_t)stream>>>(te, pool, dtab, dpool_part, dte_part, G, d, C);
    return (int)cudaGetLastError();
}

// ------------------------------------------------------------------------------------------------------------------
// Per-node T x T mix matrix of hyperTem (GPTST.py:156-158):  M_n = A_n^T A_n,  A_n (Ht x T).  The batched 8x12 / 12x12
// products were 30 us cuBLAS launches each way; here: one thread per output element.
//     fwd: M[n][t][s] = sum_h A[n][h][t] A[n][h][s]          bwd: dA[n][h][t] = sum_s A[n][h][s] (dM[n][t][s] + dM[n][s][t])
// ------------------------------------------------------------------------------------------------------------------
namespace gptst {
namespace sm {
__global__ void __launch_bounds__(256) mn_fwd_kernel(const float* __restrict__ A, float* __restrict__ M, int N, int Ht, int T) {
    const int i = blockIdx.x * 256 + threadIdx.x;
    if (i >= N * T * T) return;
    const int n = i / (T * T), t = (i / T) % T, s = i % T;
    const float* a = A + (size_t)n * Ht * T;
    float acc = 0.f;
    for (int h = 0; h < Ht; ++h) acc = fmaf(a[h * T + t], a[h * T + s], acc);
    M[i] = acc;
}
__global__ void __launch_bounds__(256) mn_bwd_kernel(const float* __restrict__ A, const float* __restrict__ dM, float* __restrict__ dA,
                                                     int N, int Ht, int T) {
    const int i = blockIdx.x * 256 + threadIdx.x;
    if (i >= N * Ht * T) return;
    const int n = i / (Ht * T), h = (i / T) % Ht, t = i % T;
    const float* a = A + ((size_t)n * Ht + h) * T;
    const float* m = dM + (size_t)n * T * T;
    float acc = 0.f;
    for (int s = 0; s < T; ++s) acc = fmaf(a[s], m[t * T + s] + m[s * T + t], acc);
    dA[i] = acc;
}
}  // namespace sm
}  // namespace gptst

extern "C" int gptst_mn_fwd(const float* A, float* M, int N, int Ht, int T, void* stream) {
    if (!A || !M || N <= 0 || Ht <= 0 || T <= 0) return -1;
    gptst::sm::mn_fwd_kernel<<<(N * T * T + 255) / 256, 256, 0, (cudaStream_t)stream>>>(A, M, N, Ht, T);
    return (int)cudaGetLastError();
}
extern "C" int gptst_mn_bwd(const float* A, const float* dM, float* dA, int N, int Ht, int T, void* stream) {
    if (!A || !dM || !dA || N <= 0 || Ht <= 0 || T <= 0) return -1;
    gptst::sm::mn_bwd_kernel<<<(N * Ht * T + 255) / 256, 256, 0, (cudaStream_t)stream>>>(A, dM, dA, N, Ht, T);
    return (int)cudaGetLastError();
}

// ------------------------------------------------------------------------------------------------------------------
// Forward of a low-rank table  Tab[g][c] = sum_k te[g][k] pool[k][c]  as a streaming kernel: a thread keeps the d x 4
// pool block of its four columns in registers and walks 16 rows (te rows broadcast from shared memory); the output
// (12.6 MB for the time-adaptive weights) is written with 16-byte stores.  cuBLAS needs ~30 us for this K = 16 product.
// ------------------------------------------------------------------------------------------------------------------
namespace gptst {
namespace sm {
constexpr int kTabRows = 16;
__global__ void __launch_bounds__(256) table_fwd_kernel(const float* __restrict__ te, const float* __restrict__ pool,
                                                        float* __restrict__ tab, int G, int d, int C) {
    __shared__ float tes[kTabRows][kDmax];
    const int tid = threadIdx.x;
    const int r0 = blockIdx.y * kTabRows;
    {
        const int r = tid / kDmax, k = tid % kDmax;       // 256 threads = 16 rows x 16
        tes[r][k] = (r0 + r < G && k < d) ? te[(size_t)(r0 + r) * d + k] : 0.f;
    }
    __syncthreads();
    const int c = (blockIdx.x * 256 + tid) * 4;
    if (c >= C) return;
    const bool vec = (C % 4 == 0) && (c + 3 < C);
    float4 p[kDmax];
#pragma unroll
    for (int k = 0; k < kDmax; ++k) {
        if (k < d) {
            if (vec) p[k] = *reinterpret_cast<const float4*>(pool + (size_t)k * C + c);
            else p[k] = make_float4(pool[(size_t)k * C + c], c + 1 < C ? pool[(size_t)k * C + c + 1] : 0.f,
                                    c + 2 < C ? pool[(size_t)k * C + c + 2] : 0.f, c + 3 < C ? pool[(size_t)k * C + c + 3] : 0.f);
        } else p[k] = make_float4(0.f, 0.f, 0.f, 0.f);
    }
    for (int r = 0; r < kTabRows && r0 + r < G; ++r) {
        float4 o = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
        for (int k = 0; k < kDmax; ++k) {
            const float t = tes[r][k];
            o.x = fmaf(t, p[k].x, o.x); o.y = fmaf(t, p[k].y, o.y); o.z = fmaf(t, p[k].z, o.z); o.w = fmaf(t, p[k].w, o.w);
        }
        float* dst = tab + (size_t)(r0 + r) * C + c;
        if (vec) *reinterpret_cast<float4*>(dst) = o;
        else { dst[0] = o.x; if (c + 1 < C) dst[1] = o.y; if (c + 2 < C) dst[2] = o.z; if (c + 3 < C) dst[3] = o.w; }
    }
}
}  // namespace sm
}  // namespace gptst

extern "C" int gptst_table_fwd(const float* te, const float* pool, float* tab, int G, int d, int C, void* stream) {
    if (!te || !pool || !tab || G <= 0 || C <= 0) return -1;
    if (d < 1 || d > gptst::sm::kDmax) return -2;
    dim3 grid((C + 1023) / 1024, (G + gptst::sm::kTabRows - 1) / gptst::sm::kTabRows);
    gptst::sm::table_fwd_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(te, pool, tab, G, d, C);
    return (int)cudaGetLastError();
}

// ------------------------------------------------------------------------------------------------------------------
// Score head of the mask scorer (GPTST.py:33 ln3 + the caller's softmax, :332/:343): prob = softmax(h W3^T + b3) per cell.
// It sits on the critical front of the adaptive phase (the encoder waits for the mask); as library calls it was a
// 64 -> 10 GEMM + bias epilogue + softmax (48 us).  CTA = 256 rows staged with 16-byte row chunks, thread = row.
// ------------------------------------------------------------------------------------------------------------------
namespace gptst {
namespace sm {
template <int D>
__global__ void __launch_bounds__(256) score_head_kernel(const float* __restrict__ h, const float* __restrict__ W3,
                                                         const float* __restrict__ b3, float* __restrict__ prob, long rows, int H) {
    constexpr int LDS = D + 1;
    extern __shared__ __align__(16) float sh[];
    float* tile = sh;                       // [256][D+1]
    float* Ws = tile + 256 * LDS + ((256 * LDS) % 4 ? 4 - (256 * LDS) % 4 : 0);   // [D][16] (transposed, 16-byte aligned)
    float* bs = Ws + kMaxH * D;             // [H]
    float* ps = bs + kMaxH;                 // [256][H] output staging
    const int tid = threadIdx.x;
    const long r0 = (long)blockIdx.x * 256;
    for (int i = tid; i < kMaxH * D; i += 256) { const int d = i / kMaxH, j = i % kMaxH; Ws[i] = (j < H) ? W3[j * D + d] : 0.f; }
    if (tid < H) bs[tid] = b3[tid];
    for (int i = tid; i < 256 * (D / 4); i += 256) {
        const int r = i / (D / 4), q = i % (D / 4);
        float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
        if (r0 + r < rows) v = *reinterpret_cast<const float4*>(h + (r0 + r) * D + 4 * q);
        float* dst = tile + r * LDS + 4 * q;
        dst[0] = v.x; dst[1] = v.y; dst[2] = v.z; dst[3] = v.w;
    }
    __syncthreads();
    float z[kMaxH];
#pragma unroll
    for (int j = 0; j < kMaxH; ++j) z[j] = (j < H) ? bs[j] : -INFINITY;
    const float* row = tile + tid * LDS;
    // W3 is held transposed, [d][16]: four 16-byte broadcast loads per column instead of one 4-byte load per (column, class)
#pragma unroll 4
    for (int d = 0; d < D; ++d) {
        const float x = row[d];
        const float4* w4 = reinterpret_cast<const float4*>(Ws + d * kMaxH);
#pragma unroll
        for (int q = 0; q < kMaxH / 4; ++q) {
            if (4 * q < H) {
                const float4 w = w4[q];
                z[4 * q] = fmaf(x, w.x, z[4 * q]); z[4 * q + 1] = fmaf(x, w.y, z[4 * q + 1]);
                z[4 * q + 2] = fmaf(x, w.z, z[4 * q + 2]); z[4 * q + 3] = fmaf(x, w.w, z[4 * q + 3]);
            }
        }
    }
#pragma unroll
    for (int j = 0; j < kMaxH; ++j) if (j >= H) z[j] = -INFINITY;
    float m = z[0];
#pragma unroll
    for (int j = 1; j < kMaxH; ++j) m = fmaxf(m, z[j]);
    float ssum = 0.f;
#pragma unroll
    for (int j = 0; j < kMaxH; ++j) { z[j] = (j < H) ? expf(z[j] - m) : 0.f; ssum += z[j]; }
    const float inv = 1.f / ssum;
#pragma unroll
    for (int j = 0; j < kMaxH; ++j)
        if (j < H) ps[tid * H + j] = z[j] * inv;
    __syncthreads();
    long nvalid = rows - r0;
    if (nvalid > 256) nvalid = 256;
    for (long i = tid; i < nvalid * H; i += 256) prob[r0 * H + i] = ps[i];
}
}  // namespace sm
}  // namespace gptst

extern "C" int gptst_score_head_fwd(const float* h, const float* W3, const float* b3, float* prob, long rows, int D, int H,
                                    void* stream) {
    if (!h || !W3 || !b3 || !prob || rows <= 0) return -1;
    if (H < 1 || H > gptst::kMaxH || (D != 64 && D != 128)) return -2;
    const size_t smem = ((size_t)256 * (D + 1) + (size_t)gptst::kMaxH * D + gptst::kMaxH + 256 * (size_t)H) * 4;
    const unsigned grid = (unsigned)((rows + 255) / 256);
    cudaError_t e;
    if (D == 64) {
        e = cudaFuncSetAttribute(gptst::sm::score_head_kernel<64>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) return (int)e;
        gptst::sm::score_head_kernel<64><<<grid, 256, smem, (cudaStream_t)stream>>>(h, W3, b3, prob, rows, H);
    } else {
        e = cudaFuncSetAttribute(gptst::sm::score_head_kernel<128>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) return (int)e;
        gptst::sm::score_head_kernel<128><<<grid, 256, smem, (cudaStream_t)stream>>>(h, W3, b3, prob, rows, H);
    }
    return (int)cudaGetLastError();
}

// y[i][:] = x[i] * w[:] + b[:]   (a linear layer with one input feature, GPTST.py:22 / :298) -- a write-bound streaming kernel
namespace gptst {
namespace sm {
__global__ void __launch_bounds__(256) affine1_fwd_kernel(const float* __restrict__ x, const float* __restrict__ w,
                                                          const float* __restrict__ b, float* __restrict__ y, long n, int D) {
    const int q4 = D / 4;
    const long total = n * q4;
    for (long i = (long)blockIdx.x * 256 + threadIdx.x; i < total; i += (long)gridDim.x * 256) {
        const long r = i / q4;
        const int q = (int)(i % q4);
        const float xv = x[r];
        const float4 wv = *reinterpret_cast<const float4*>(w + 4 * q), bv = *reinterpret_cast<const float4*>(b + 4 * q);
        *reinterpret_cast<float4*>(y + r * D + 4 * q) =
            make_float4(fmaf(xv, wv.x, bv.x), fmaf(xv, wv.y, bv.y), fmaf(xv, wv.z, bv.z), fmaf(xv, wv.w, bv.w));
    }
}
}  // namespace sm
}  // namespace gptst

extern "C" int gptst_affine1_fwd(const float* x, const float* w, const float* b, float* y, long n, int D, void* stream) {
    if (!x || !w || !b || !y || n <= 0) return -1;
    if (D < 4 || D % 4 != 0) return -2;
    long blocks = (n * (D / 4) + 255) / 256;
    if (blocks > 148 * 16) blocks = 148 * 16;
    gptst::sm::affine1_fwd_kernel<<<(unsigned)blocks, 256, 0, (cudaStream_t)stream>>>(x, w, b, y, n, D);
    return (int)cudaGetLastError();
}

// Masked input embedding of the encoder (reference GPTST.py:419-421): xm[i] = mask[i] == 0 ? fill : mask[i] * flow[i];
// y[i][:] = xm[i] * w[:] + b[:].  One launch instead of four elementwise library kernels in front of the affine one; xm is kept
// for the backward (dw = sum_i dy[i,:] xm[i]).
namespace gptst {
namespace sm {
__global__ void __launch_bounds__(256) masked_affine1_fwd_kernel(const float* __restrict__ flow, long flow_stride,
                                                                 const long long* __restrict__ mask, float fill,
                                                                 const float* __restrict__ w, const float* __restrict__ b,
                                                                 float* __restrict__ xm, float* __restrict__ y, long n, int D) {
    const int q4 = D / 4;
    const long total = n * q4;
    for (long i = (long)blockIdx.x * 256 + threadIdx.x; i < total; i += (long)gridDim.x * 256) {
        const long r = i / q4;
        const int q = (int)(i % q4);
        const long long m = mask[r];
        const float xv = (m == 0) ? fill : (float)m * flow[r * flow_stride];
        if (q == 0) xm[r] = xv;
        const float4 wv = *reinterpret_cast<const float4*>(w + 4 * q), bv = *reinterpret_cast<const float4*>(b + 4 * q);
        *reinterpret_cast<float4*>(y + r * D + 4 * q) =
            make_float4(fmaf(xv, wv.x, bv.x), fmaf(xv, wv.y, bv.y), fmaf(xv, wv.z, bv.z), fmaf(xv, wv.w, bv.w));
    }
}
}  // namespace sm
}  // namespace gptst

// flow: element r at flow[r * flow_stride] (the flow channel of `source` read in place); mask (n,) int64; xm (n,), y (n, D)
extern "C" int gptst_masked_affine1_fwd(const float* flow, long flow_stride, const long long* mask, float fill, const float* w,
                                        const float* b, float* xm, float* y, long n, int D, void* stream) {
    if (!flow || !mask || !w || !b || !xm || !y || n <= 0 || flow_stride <= 0) return -1;
    if (D < 4 || D % 4 != 0) return -2;
    long blocks = (n * (D / 4) + 255) / 256;
    if (blocks > 148 * 16) blocks = 148 * 16;
    gptst::sm::masked_affine1_fwd_kernel<<<(unsigned)blocks, 256, 0, (cudaStream_t)stream>>>(flow, flow_stride, mask, fill, w, b, xm,
                                                                                            y, n, D);
    return (int)cudaGetLastError();
}

// ------------------------------------------------------------------------------------------------------------------
// Sum of per-CTA / per-split gradient partials for up to 8 tensors in ONE launch:  out_s[i] = sum_p in_s[p * numel_s + i]
// (fixed order -> deterministic).  A cap backward ends with five such reductions (dW_n, db_n, ddyn, dWp, dbp); as separate
// library reductions they cost ~7 us each on the main chain.
// ------------------------------------------------------------------------------------------------------------------
namespace gptst {
namespace sm {
struct SumSegs {
    const float* in[8];
    float* out[8];
    long numel[8];
    int parts[8];
    long start[9];     // prefix sums of numel (scalar elements, or float4 units on the vectorised path)
    int n;
};
// thread = one element (VEC = 1) or one float4 of elements (VEC = 4; start[] / numel[] are then in float4 units)
template <int VEC>
__global__ void __launch_bounds__(256) sum_partials_kernel(SumSegs s) {
    const long total = s.start[s.n];
    for (long i = (long)blockIdx.x * 256 + threadIdx.x; i < total; i += (long)gridDim.x * 256) {
        int k = 0;
#pragma unroll
        for (int j = 1; j < 8; ++j) k += (j < s.n && i >= s.start[j]);
        const long e = i - s.start[k];
        const long stride = s.numel[k];
        const int parts = s.parts[k];
        if (VEC == 4) {
            const float4* p = reinterpret_cast<const float4*>(s.in[k]) + e;
            float4 a0 = make_float4(0.f, 0.f, 0.f, 0.f), a1 = a0;
            int q = 0;
            for (; q + 1 < parts; q += 2) {
                const float4 u = p[(long)q * stride], v = p[(long)(q + 1) * stride];
                a0.x += u.x; a0.y += u.y; a0.z += u.z; a0.w += u.w;
                a1.x += v.x; a1.y += v.y; a1.z += v.z; a1.w += v.w;
            }
            if (q < parts) { const float4 u = p[(long)q * stride]; a0.x += u.x; a0.y += u.y; a0.z += u.z; a0.w += u.w; }
            reinterpret_cast<float4*>(s.out[k])[e] = make_float4(a0.x + a1.x, a0.y + a1.y, a0.z + a1.z, a0.w + a1.w);
        } else {
            const float* p = s.in[k] + e;
            float a0 = 0.f, a1 = 0.f, a2 = 0.f, a3 = 0.f;
            int q = 0;
            for (; q + 3 < parts; q += 4) {
                a0 += p[(long)q * stride]; a1 += p[(long)(q + 1) * stride]; a2 += p[(long)(q + 2) * stride]; a3 += p[(long)(q + 3) * stride];
            }
            for (; q < parts; ++q) a0 += p[(long)q * stride];
            s.out[k][e] = (a0 + a1) + (a2 + a3);
        }
    }
}
// few elements, many partials (per-CTA partials of a small parameter gradient): one WARP per element, lanes stride over the
// partials, shuffle tree at the end -- a thread walking hundreds of partials serially is pure load latency (38 us for the
// score head's 650 x 510 case)
__global__ void __launch_bounds__(256) sum_partials_warp_kernel(SumSegs s) {
    const long total = s.start[s.n];
    const int lane = threadIdx.x & 31;
    for (long i = ((long)blockIdx.x * 256 + threadIdx.x) >> 5; i < total; i += (long)gridDim.x * 8) {   // warp-uniform
        int k = 0;
#pragma unroll
        for (int j = 1; j < 8; ++j) k += (j < s.n && i >= s.start[j]);
        const long e = i - s.start[k];
        const float* p = s.in[k] + e;
        const long stride = s.numel[k];
        const int parts = s.parts[k];
        float a = 0.f;
        for (int q = lane; q < parts; q += 32) a += p[(long)q * stride];
        a = warp_sum(a);
        if (lane == 0) s.out[k][e] = a;
    }
}
}  // namespace sm
}  // namespace gptst

// ins[k]: (parts[k], numel[k]) contiguous partials, outs[k]: (numel[k]); n <= 8 segments.
// Segments with few elements and many partials (the 255 per-CTA partials of a cap's dWp / dbp) take the warp-per-element
// kernel, everything else the float4 lanes; a mixed call becomes two launches (a float4 lane walking 255 partials serially
// is ~100 us of dependent L2 misses -- it used to be the tail of every cap backward's side stream).
static int sum_partials_launch(const float* const* ins, float* const* outs, const long* numel, const int* parts, int n,
                               bool warp_path, cudaStream_t stream) {
    gptst::sm::SumSegs s;
    bool vec = !warp_path;
    for (int k = 0; k < n; ++k)
        if (numel[k] % 4 != 0 || ((uintptr_t)ins[k] & 15) != 0 || ((uintptr_t)outs[k] & 15) != 0) vec = false;
    long tot = 0;
    for (int k = 0; k < 8; ++k) {
        s.start[k] = tot;
        if (k < n) {
            s.in[k] = ins[k]; s.out[k] = outs[k]; s.numel[k] = vec ? numel[k] / 4 : numel[k]; s.parts[k] = parts[k];
            tot += s.numel[k];
        } else { s.in[k] = nullptr; s.out[k] = nullptr; s.numel[k] = 0; s.parts[k] = 0; }
    }
    s.start[8] = tot;
    for (int k = n; k < 9; ++k) s.start[k] = tot;
    s.n = n;
    if (warp_path) {
        long blocks = (tot + 7) / 8;
        if (blocks > 148 * 8) blocks = 148 * 8;
        gptst::sm::sum_partials_warp_kernel<<<(unsigned)blocks, 256, 0, stream>>>(s);
        return (int)cudaGetLastError();
    }
    long blocks = (tot + 255) / 256;
    if (blocks > 148 * 8) blocks = 148 * 8;
    if (vec) gptst::sm::sum_partials_kernel<4><<<(unsigned)blocks, 256, 0, stream>>>(s);
    else gptst::sm::sum_partials_kernel<1><<<(unsigned)blocks, 256, 0, stream>>>(s);
    return (int)cudaGetLastError();
}

extern "C" int gptst_sum_partials(const float* const* ins, float* const* outs, const long* numel, const int* parts, int n,
                                  void* stream) {
    if (!ins || !outs || !numel || !parts || n <= 0) return -1;
    if (n > 8) return -2;
    for (int k = 0; k < n; ++k)
        if (!ins[k] || !outs[k] || numel[k] <= 0 || parts[k] <= 0) return -1;
    const float* gi[2][8]; float* go[2][8]; long gn[2][8]; int gp[2][8]; int cnt[2] = {0, 0};
    for (int k = 0; k < n; ++k) {
        const int w = (numel[k] <= 16384 && parts[k] >= 64) ? 1 : 0;
        gi[w][cnt[w]] = ins[k]; go[w][cnt[w]] = outs[k]; gn[w][cnt[w]] = numel[k]; gp[w][cnt[w]] = parts[k]; ++cnt[w];
    }
    for (int w = 0; w < 2; ++w) {
        if (!cnt[w]) continue;
        const int rc = sum_partials_launch(gi[w], go[w], gn[w], gp[w], cnt[w], w == 1, (cudaStream_t)stream);
        if (rc) return rc;
    }
    return 0;
}

// Backward of the score head:  dz = prob * (dprob - <prob, dprob>) ;  dh = dz W3 ;  dW3_part[cta] = dz^T h ;  db3_part[cta] = sum dz
// (library version: two 187 us split-K GEMMs with K = B*T*N).  CTA = 256 rows, thread = row for dz / dh, then the CTA's
// (H x D + H) partial with h and dz staged in shared memory.  part: (ctas, H*D + H), summed by the caller.
namespace gptst {
namespace sm {
template <int D>
__global__ void __launch_bounds__(256) score_head_bwd_kernel(const float* __restrict__ h, const float* __restrict__ W3,
                                                             const float* __restrict__ prob, const float* __restrict__ dprob,
                                                             float* __restrict__ dh, float* __restrict__ part, long rows, int H) {
    constexpr int LDS = D + 1;
    extern __shared__ __align__(16) float sh[];
    float* tile = sh;                       // [256][D+1]  h rows
    float* Ws = tile + 256 * LDS;           // [H][D]
    float* dzs = Ws + kMaxH * D;            // [256][kMaxH+1]
    const int tid = threadIdx.x;
    const long r0 = (long)blockIdx.x * 256;
    for (int i = tid; i < H * D; i += 256) Ws[i] = W3[i];
    for (int i = tid; i < 256 * (D / 4); i += 256) {
        const int r = i / (D / 4), q = i % (D / 4);
        float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
        if (r0 + r < rows) v = *reinterpret_cast<const float4*>(h + (r0 + r) * D + 4 * q);
        float* dst = tile + r * LDS + 4 * q;
        dst[0] = v.x; dst[1] = v.y; dst[2] = v.z; dst[3] = v.w;
    }
    // dz of this thread's row
    float dz[kMaxH];
    {
        const long r = r0 + tid;
        float dot = 0.f;
#pragma unroll
        for (int j = 0; j < kMaxH; ++j) {
            float p = 0.f, g = 0.f;
            if (j < H && r < rows) { p = prob[r * H + j]; g = dprob[r * H + j]; }
            dz[j] = p;
            dot = fmaf(p, g, dot);
            dzs[tid * (kMaxH + 1) + j] = g;          // temporarily dprob
        }
#pragma unroll
        for (int j = 0; j < kMaxH; ++j) {
            dz[j] = dz[j] * (dzs[tid * (kMaxH + 1) + j] - dot);
            dzs[tid * (kMaxH + 1) + j] = dz[j];
        }
    }
    __syncthreads();
    // dh row = dz . W3  (written through the row's own slot of the h tile? no: h is still needed -> straight to global)
    if (dh && r0 + tid < rows) {
        float* out = dh + (r0 + tid) * D;
        for (int d0 = 0; d0 < D; d0 += 4) {
            float4 o = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
            for (int j = 0; j < kMaxH; ++j) {
                if (j < H) {
                    const float z = dz[j];
                    o.x = fmaf(z, Ws[j * D + d0], o.x); o.y = fmaf(z, Ws[j * D + d0 + 1], o.y);
                    o.z = fmaf(z, Ws[j * D + d0 + 2], o.z); o.w = fmaf(z, Ws[j * D + d0 + 3], o.w);
                }
            }
            *reinterpret_cast<float4*>(out + d0) = o;
        }
    }
    // partial dW3 (H x D) and db3 (H) of this CTA
    float* po = part + (size_t)blockIdx.x * (H * D + H);
    for (int i = tid; i < H * D + H; i += 256) {
        float s = 0.f;
        if (i < H * D) {
            const int j = i / D, d = i % D;
            for (int r = 0; r < 256; ++r) s = fmaf(dzs[r * (kMaxH + 1) + j], tile[r * LDS + d], s);
        } else {
            const int j = i - H * D;
            for (int r = 0; r < 256; ++r) s += dzs[r * (kMaxH + 1) + j];
        }
        po[i] = s;
    }
}
}  // namespace sm
}  // namespace gptst

extern "C" int gptst_score_head_bwd_parts(long rows) { return (int)((rows + 255) / 256); }
extern "C" int gptst_score_head_bwd(const float* h, const float* W3, const float* prob, const float* dprob, float* dh, float* part,
                                    long rows, int D, int H, void* stream) {
    if (!h || !W3 || !prob || !dprob || !part || rows <= 0) return -1;
    if (H < 1 || H > gptst::kMaxH || (D != 64 && D != 128)) return -2;
    const size_t smem = ((size_t)256 * (D + 1) + (size_t)gptst::kMaxH * D + 256 * (size_t)(gptst::kMaxH + 1)) * 4;
    const unsigned grid = (unsigned)((rows + 255) / 256);
    cudaError_t e;
    if (D == 64) {
        e = cudaFuncSetAttribute(gptst::sm::score_head_bwd_kernel<64>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) return (int)e;
        gptst::sm::score_head_bwd_kernel<64><<<grid, 256, smem, (cudaStream_t)stream>>>(h, W3, prob, dprob, dh, part, rows, H);
    } else {
        e = cudaFuncSetAttribute(gptst::sm::score_head_bwd_kernel<128>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) return (int)e;
        gptst::sm::score_head_bwd_kernel<128><<<grid, 256, smem, (cudaStream_t)stream>>>(h, W3, prob, dprob, dh, part, rows, H);
    }
    return (int)cudaGetLastError();
}
