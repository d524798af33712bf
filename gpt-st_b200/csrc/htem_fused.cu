// Fused hyperTem (reference GPTST.py:154-163) for D = 64, T = 12: one persistent, TMA-fed kernel per
// direction on the main chain of the step.
//
//   forward   out = LReLU( (M_n o eb) W_bt + bias_bt + eb )                       reads eb, writes out (+ sign mask, + ret)
//   backward  deb = dy + M_n^T o (dy W_bt^T),  dy = dOut . LReLU'(mask)           reads dOut, writes deb (+ dret)
//
// The temporal mix is per NODE (it contracts the 12 time steps of one node), the projection has one weight per (b, t)
// (it wants 16-row tiles of nodes of ONE time step).  A task is (sample b, 16 consecutive nodes, all 12 time steps):
// a 12 x 16 x 64 fp32 tile of 48 KB, fetched by ONE cp.async.bulk.tensor (4-D tensor map over (B,T,N,D), rows past N are
// zero-filled by the TMA unit) into a landing buffer while the CTA still computes the previous task.
//
//   * mix: thread = (node, 4 columns) keeps its 12 x float4 in registers, M_n rows are broadcast reads of a
//     TMA-prefetched shared-memory copy; exact fp32 FMA.
//   * projection: mma.sync m16n8k16, three-term fp16 split (mma_f16.cuh).  The A operand (16 nodes x 64) of each time
//     step is written by the mix threads as hi | lo planes with ONE power-of-two scale per row (exact, undone on the fp32
//     accumulator), so any activation / gradient magnitude keeps 22 significant bits.  The B operand W_bt is used by
//     exactly one warp of one CTA per task, so it never goes through shared memory: `htem_pack_kernel` stores W_bt (x 64)
//     once per call in FRAGMENT ORDER (one 16-byte vector per lane = hi/lo of both B registers of one k-block x n-tile) and
//     the warps read it with fully coalesced 512-byte requests, one k-block ahead.
//   * work is pipelined in three rounds of four time steps: [mix/epilogue phase] -> barrier -> [MMA phase] -> barrier; two CTAs
//     per SM run out of phase, so the FMA pipe of one overlaps the tensor pipe of the other.
//   * the epilogue runs in the mix layout (coalesced 16-byte stores of full rows); the LeakyReLU sign mask (8 bytes per row,
//     bit 16*(c & 3) + (c >> 2) = out[c] > 0: the order four warp ballots deliver) costs four votes and two byte permutes.
// What is NOT here: the parameter-side gradients (dW_bt = ret^T dy, dM_n = dret . eb) -- nothing on the main chain reads
// them, so they are computed by side-stream kernels from `ret` / `dret` (gproj3 dW-only mode, tmix dM-only mode).
#include <cuda.h>

#include "common.cuh"
#include "mma_f16.cuh"
#include "tma.cuh"

namespace gptst {
namespace htf {

using namespace hf;
constexpr int T = 12, D = 64, NC = 16, RT = 4, NR = 3;
constexpr int NTHREADS = 256;
constexpr float WSCALE = 64.f;
constexpr int TILE_T_BYTES = NC * ROWB;                 // one time step of an operand tile: 16 rows x 272 B
constexpr int OPER_BYTES = RT * TILE_T_BYTES;           // 17408
constexpr int LAND_BYTES = T * NC * D * 4;              // 49152
constexpr int MN_FLOATS = NC * T * T;                   // 2304
constexpr int WF_VEC_PER_GROUP = 2 * 4 * 4 * 32;        // uint4 per (b,t): [half][kb][j][lane]

// ---- shared-memory maps ------------------------------------------------------------------------------------------
struct FwdSmem {
    static constexpr int land = 0;
    static constexpr int A = land + LAND_BYTES;
    static constexpr int C = A + OPER_BYTES;
    static constexpr int Mn = C + OPER_BYTES;                       // [2][2304] f32
    static constexpr int bias = Mn + 2 * MN_FLOATS * 4;             // [2][768] f32
    static constexpr int scl = bias + 2 * T * D * 4;                // [4][16] f32
    static constexpr int bar = scl + RT * NC * 4;
    static constexpr int total = bar + 16;
};
struct BwdSmem {
    static constexpr int land = 0;                                  // 3 boxes of [4][16][64] f32
    static constexpr int A = land + LAND_BYTES;
    static constexpr int C = A + OPER_BYTES;
    static constexpr int Mn = C + OPER_BYTES;                       // [2][2304] f32
    static constexpr int msk = Mn + 2 * MN_FLOATS * 4;              // [3][4][16] uint2
    static constexpr int scl = msk + NR * RT * NC * 8;              // [4][16] f32
    static constexpr int bar = scl + RT * NC * 4;                   // 3 barriers
    static constexpr int total = bar + 32;
};

// ------------------------------------------------------------------------------------------------------------------
// W_bt (G, 64, 64) fp32 [in][out]  ->  fragment-ordered fp16 hi/lo tables (x 64)
//   wf : B[k = in][n = out]  (forward,  ret W)        wb : B[k = out][n = in]  (backward, dy W^T)
// slot s = ((half * 4 + kb) * 4 + j) * 32 + lane holds {hi(b0), hi(b1), lo(b0), lo(b1)} of n-tile 4*half + j, k-block kb
// ------------------------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) htem_pack_kernel(const float* __restrict__ W, uint4* __restrict__ wf, uint4* __restrict__ wb) {
    __shared__ float Ws[64][65];
    const int grp = blockIdx.x, tid = threadIdx.x;
    const float* Wg = W + (size_t)grp * 4096;
#pragma unroll
    for (int u = 0; u < 4; ++u) {
        const int i = tid + 256 * u;                                // float4 index
        const float4 v = *reinterpret_cast<const float4*>(Wg + 4 * i);
        const int r = i >> 4, c = (i & 15) * 4;
        Ws[r][c] = v.x * WSCALE; Ws[r][c + 1] = v.y * WSCALE; Ws[r][c + 2] = v.z * WSCALE; Ws[r][c + 3] = v.w * WSCALE;
    }
    __syncthreads();
#pragma unroll
    for (int u = 0; u < 4; ++u) {
        const int s = tid + 256 * u;
        const int lane = s & 31, j = (s >> 5) & 3, kb = (s >> 7) & 3, half = s >> 9;
        const int g = lane >> 2, tg = lane & 3;
        const int k0 = 16 * kb + 2 * tg, nn = 8 * (4 * half + j) + g;
        uint4 o;
        if (wf) {
            split_h2<PREC_3XTF32>(Ws[k0][nn], Ws[k0 + 1][nn], o.x, o.z);
            split_h2<PREC_3XTF32>(Ws[k0 + 8][nn], Ws[k0 + 9][nn], o.y, o.w);
            wf[(size_t)grp * WF_VEC_PER_GROUP + s] = o;
        }
        if (wb) {
            split_h2<PREC_3XTF32>(Ws[nn][k0], Ws[nn][k0 + 1], o.x, o.z);
            split_h2<PREC_3XTF32>(Ws[nn][k0 + 8], Ws[nn][k0 + 9], o.y, o.w);
            wb[(size_t)grp * WF_VEC_PER_GROUP + s] = o;
        }
    }
}

// ---- pieces shared by both directions ----------------------------------------------------------------------------
// max |.| over the warp = the two rows (nodes) it holds; non-negative floats order like their bit patterns, so ONE redux.sync
// replaces a four-step shuffle chain.  A scale shared by two rows is as good as a per-row one: the split keeps 22 bits for
// every element within 2^16 of the scaled maximum and an absolute 2^-25 (of 2^14) below that.
__device__ __forceinline__ float rowpair_absmax(const float4& v) {
    const float m = fmaxf(fmaxf(fabsf(v.x), fabsf(v.y)), fmaxf(fabsf(v.z), fabsf(v.w)));
    return __uint_as_float(__reduce_max_sync(0xffffffffu, __float_as_uint(m)));
}
// a += m * e on two packed fp32 lanes per instruction (FFMA2, sm_100); bit-identical to four fmaf
__device__ __forceinline__ void fma4(float4& a, float m, const float4& e) {
    const float2 mm = make_float2(m, m);
    const float2 lo = __ffma2_rn(mm, make_float2(e.x, e.y), make_float2(a.x, a.y));
    const float2 hi = __ffma2_rn(mm, make_float2(e.z, e.w), make_float2(a.z, a.w));
    a = make_float4(lo.x, lo.y, hi.x, hi.y);
}
// one row of the A operand: v (4 columns of this thread) scaled by a per-row power of two, hi | lo planes
__device__ __forceinline__ void put_operand_row(unsigned char* Arow, float* sclp, int cg, const float4& v) {
    const float2 sc = pow2_scale_for_fp16(rowpair_absmax(v));
    uint2 hi, lo;
    split_h2<PREC_3XTF32>(v.x * sc.x, v.y * sc.x, hi.x, lo.x);
    split_h2<PREC_3XTF32>(v.z * sc.x, v.w * sc.x, hi.y, lo.y);
    *reinterpret_cast<uint2*>(Arow + 8 * cg) = hi;
    *reinterpret_cast<uint2*>(Arow + LO + 8 * cg) = lo;
    if (cg == 0) *sclp = sc.y * (1.f / WSCALE);
}
// MMA phase of one round: warp = (local time step tl, column half); C rows = the 16 nodes, un-scaled fp32
__device__ __forceinline__ void mma_round(const unsigned char* Ab, unsigned char* Cb, const float* scl, const uint4* __restrict__ wp,
                                          uint4 (&cur)[4], int tl, int half, int lane) {
    const int g = lane >> 2, tg = lane & 3;
    float acc[4][4];
#pragma unroll
    for (int j = 0; j < 4; ++j) acc[j][0] = acc[j][1] = acc[j][2] = acc[j][3] = 0.f;
    const uint32_t a0 = smem_u32(Ab + (size_t)(tl * NC + 8 * ((lane >> 3) & 1) + (lane & 7)) * ROWB + 16 * (lane >> 4));
#pragma unroll
    for (int kb = 0; kb < 4; ++kb) {
        uint4 nxt[4];
        if (kb < 3) {
#pragma unroll
            for (int j = 0; j < 4; ++j) nxt[j] = __ldg(wp + ((kb + 1) * 4 + j) * 32);
        }
        uint32_t ah[4], al[4];
        ldsm_x4(ah, a0 + 32 * kb);
        ldsm_x4(al, a0 + 32 * kb + LO);
#pragma unroll
        for (int j = 0; j < 4; ++j) mma3<PREC_3XTF32>(acc[j], ah, al, cur[j].x, cur[j].y, cur[j].z, cur[j].w);
        if (kb < 3) {
#pragma unroll
            for (int j = 0; j < 4; ++j) cur[j] = nxt[j];
        }
    }
    const float s0 = scl[tl * NC + g], s1 = scl[tl * NC + g + 8];
    unsigned char* r0 = Cb + (size_t)(tl * NC + g) * ROWB + (32 * half + 2 * tg) * 4;
    unsigned char* r1 = r0 + 8 * ROWB;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
        *reinterpret_cast<float2*>(r0 + 32 * j) = make_float2(acc[j][0] * s0, acc[j][1] * s0);
        *reinterpret_cast<float2*>(r1 + 32 * j) = make_float2(acc[j][2] * s1, acc[j][3] * s1);
    }
}
__device__ __forceinline__ void load_kb0(uint4 (&cur)[4], const uint4* __restrict__ wp) {
#pragma unroll
    for (int j = 0; j < 4; ++j) cur[j] = __ldg(wp + j * 32);
}
// Task order: plain grid-stride.  (Giving the two co-resident CTAs of an SM neighbouring chunks of one sample, so that the second
// finds the first one's W_bt fragments in L1, was measured: no gain, 38.8 vs 38.1 us.)
__device__ __forceinline__ int first_task() { return blockIdx.x; }
__device__ __forceinline__ int task_stride() { return gridDim.x; }
__device__ __forceinline__ const uint4* wfrag_ptr(const uint4* __restrict__ w, int bt, int half, int lane) {
    return w + ((size_t)bt * 2 + half) * (4 * 4 * 32) + lane;
}

// ------------------------------------------------------------------------------------------------------------------
// forward
// ------------------------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(NTHREADS, 2)
htem_fwd_kernel(const __grid_constant__ CUtensorMap tm_eb, const __grid_constant__ CUtensorMap tm_mn,
                const uint4* __restrict__ wfrag, const float* __restrict__ bias, float* __restrict__ out,
                uint2* __restrict__ mask, float* __restrict__ ret, int N, int nch, int ntasks) {
    pdl_enter();
    const int Npad = nch * NC;                          // mask rows are padded to whole 16-node chunks per (b, t)
    extern __shared__ __align__(128) unsigned char sm[];
    const float* land = reinterpret_cast<const float*>(sm + FwdSmem::land);
    unsigned char* Ab = sm + FwdSmem::A;
    unsigned char* Cb = sm + FwdSmem::C;
    const float* Mns = reinterpret_cast<const float*>(sm + FwdSmem::Mn);
    const float* bss = reinterpret_cast<const float*>(sm + FwdSmem::bias);
    float* scl = reinterpret_cast<float*>(sm + FwdSmem::scl);
    uint64_t* bar = reinterpret_cast<uint64_t*>(sm + FwdSmem::bar);

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int n = tid >> 4, cg = tid & 15;
    const int tl = warp >> 1, half = warp & 1;
    constexpr uint32_t TX = LAND_BYTES + MN_FLOATS * 4 + T * D * 4;

    auto issue = [&](int task, int buf) {
        const int b = task / nch, n0 = (task - b * nch) * NC;
        tma::mbar_expect_tx(bar, TX);
        tma::load_4d(sm + FwdSmem::land, &tm_eb, bar, 0, n0, 0, b);
        tma::load_2d(sm + FwdSmem::Mn + buf * MN_FLOATS * 4, &tm_mn, bar, 0, n0);
        tma::load_bulk(sm + FwdSmem::bias + buf * T * D * 4, bias + (size_t)b * T * D, T * D * 4, bar);
    };
    if (tid == 0) {
        tma::prefetch_map(&tm_eb);
        tma::prefetch_map(&tm_mn);
        tma::mbar_init(bar, 1);
        tma::fence_barrier_init();
    }
    __syncthreads();
    int task = first_task();
    const int tstride = task_stride();
    if (tid == 0 && task < ntasks) issue(task, 0);

    for (int it = 0; task < ntasks; task += tstride, ++it) {
        const int b = task / nch, n0 = (task - b * nch) * NC;
        const int buf = it & 1;
        const bool row_ok = n0 + n < N;
        const float* Mrow = Mns + buf * MN_FLOATS + n * (T * T);
        const float* bsb = bss + buf * T * D + 4 * cg;
        const size_t slab = (size_t)N * D;
        const size_t gofs = ((size_t)b * T * N + n0 + n) * D + 4 * cg;     // element (b, 0, n0+n, 4cg)
        uint4 cur[4];
        load_kb0(cur, wfrag_ptr(wfrag, b * T + tl, half, lane));
        tma::mbar_wait(bar, it & 1);
        float4 e[T];
#pragma unroll
        for (int t = 0; t < T; ++t) e[t] = *reinterpret_cast<const float4*>(land + (t * NC + n) * D + 4 * cg);

#pragma unroll
        for (int r = 0; r <= NR; ++r) {
            if (r > 0) {
                // ---- epilogue of round r-1 in the mix layout: bias + residual + LeakyReLU, sign mask, coalesced stores
#pragma unroll
                for (int tt = 0; tt < RT; ++tt) {
                    const int t = RT * (r - 1) + tt;
                    const float4 c = *reinterpret_cast<const float4*>(Cb + (size_t)(tt * NC + n) * ROWB + 16 * cg);
                    const float4 bv = *reinterpret_cast<const float4*>(bsb + t * D);
                    float4 y;
                    y.x = c.x + bv.x + e[t].x; y.y = c.y + bv.y + e[t].y; y.z = c.z + bv.z + e[t].z; y.w = c.w + bv.w + e[t].w;
                    const uint32_t b0 = __ballot_sync(0xffffffffu, y.x > 0.f), b1 = __ballot_sync(0xffffffffu, y.y > 0.f);
                    const uint32_t b2 = __ballot_sync(0xffffffffu, y.z > 0.f), b3 = __ballot_sync(0xffffffffu, y.w > 0.f);
                    const uint32_t sel = (lane < 16) ? 0x5410u : 0x7632u;      // low halves = the warp's first row, high = second
                    if (row_ok) {
                        y.x = lrelu(y.x); y.y = lrelu(y.y); y.z = lrelu(y.z); y.w = lrelu(y.w);
                        *reinterpret_cast<float4*>(out + gofs + t * slab) = y;
                        if (cg == 0) mask[(size_t)(b * T + t) * Npad + n0 + n] = make_uint2(__byte_perm(b0, b1, sel), __byte_perm(b2, b3, sel));
                    }
                }
            }
            if (r < NR) {
                // ---- temporal mix of the round's four time steps -> A operand rows (+ ret for the side-stream dW kernel)
#pragma unroll
                for (int tt = 0; tt < RT; ++tt) {
                    const int t = RT * r + tt;
                    float4 a = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
                    for (int q = 0; q < 3; ++q) {
                        const float4 m = *reinterpret_cast<const float4*>(Mrow + t * T + 4 * q);
                        fma4(a, m.x, e[4 * q]);
                        fma4(a, m.y, e[4 * q + 1]);
                        fma4(a, m.z, e[4 * q + 2]);
                        fma4(a, m.w, e[4 * q + 3]);
                    }
                    if (ret != nullptr && row_ok) *reinterpret_cast<float4*>(ret + gofs + t * slab) = a;
                    put_operand_row(Ab + (size_t)(tt * NC + n) * ROWB, scl + tt * NC + n, cg, a);
                }
                __syncthreads();
                if (r == 0 && tid == 0 && task + tstride < ntasks) issue(task + tstride, buf ^ 1);
                mma_round(Ab, Cb, scl, wfrag_ptr(wfrag, b * T + RT * r + tl, half, lane), cur, tl, half, lane);
                if (r + 1 < NR) load_kb0(cur, wfrag_ptr(wfrag, b * T + RT * (r + 1) + tl, half, lane));
                __syncthreads();
            }
        }
    }
}

// ------------------------------------------------------------------------------------------------------------------
// backward (main chain):  deb = dy + M^T o (dy W^T)
// ------------------------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(NTHREADS, 2)
htem_bwd_kernel(const __grid_constant__ CUtensorMap tm_go, const __grid_constant__ CUtensorMap tm_mn,
                const uint4* __restrict__ wfrag, const uint2* __restrict__ mask, float* __restrict__ deb,
                float* __restrict__ dret, int N, int nch, int ntasks) {
    extern __shared__ __align__(128) unsigned char sm[];
    const float* land = reinterpret_cast<const float*>(sm + BwdSmem::land);
    unsigned char* Ab = sm + BwdSmem::A;
    unsigned char* Cb = sm + BwdSmem::C;
    const float* Mns = reinterpret_cast<const float*>(sm + BwdSmem::Mn);
    const uint2* msk = reinterpret_cast<const uint2*>(sm + BwdSmem::msk);
    float* scl = reinterpret_cast<float*>(sm + BwdSmem::scl);
    uint64_t* bar = reinterpret_cast<uint64_t*>(sm + BwdSmem::bar);

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int n = tid >> 4, cg = tid & 15;
    const int tl = warp >> 1, half = warp & 1;
    constexpr uint32_t BOX = RT * NC * D * 4;          // 16384

    // box r of a task: dOut rows of time steps 4r..4r+3 (+ their mask words; box 0 also brings M_n)
    auto issue = [&](int task, int r, int buf) {
        const int b = task / nch, n0 = (task - b * nch) * NC;
        tma::mbar_expect_tx(bar + r, BOX + RT * NC * 8 + (r == 0 ? MN_FLOATS * 4 : 0));
        tma::load_4d(sm + BwdSmem::land + r * BOX, &tm_go, bar + r, 0, n0, RT * r, b);
#pragma unroll
        for (int tt = 0; tt < RT; ++tt)
            tma::load_bulk(sm + BwdSmem::msk + (r * RT + tt) * NC * 8, mask + (size_t)(b * T + RT * r + tt) * (nch * NC) + n0, NC * 8, bar + r);
        if (r == 0) tma::load_2d(sm + BwdSmem::Mn + buf * MN_FLOATS * 4, &tm_mn, bar, 0, n0);
    };
    if (tid == 0) {
        tma::prefetch_map(&tm_go);
        tma::prefetch_map(&tm_mn);
        for (int r = 0; r < NR; ++r) tma::mbar_init(bar + r, 1);
        tma::fence_barrier_init();
    }
    __syncthreads();
    int task = first_task();
    const int tstride = task_stride();
    if (tid == 0 && task < ntasks) {
        for (int r = 0; r < NR; ++r) issue(task, r, 0);
    }

    for (int it = 0; task < ntasks; task += tstride, ++it) {
        const int b = task / nch, n0 = (task - b * nch) * NC;
        const int buf = it & 1;
        const bool row_ok = n0 + n < N;
        const bool has_next = task + tstride < ntasks;
        const float* Mrow = Mns + buf * MN_FLOATS + n * (T * T);
        const size_t slab = (size_t)N * D;
        const size_t gofs = ((size_t)b * T * N + n0 + n) * D + 4 * cg;
        uint4 cur[4];
        load_kb0(cur, wfrag_ptr(wfrag, b * T + tl, half, lane));
        float4 acc[T];
#pragma unroll
        for (int t = 0; t < T; ++t) acc[t] = make_float4(0.f, 0.f, 0.f, 0.f);

#pragma unroll
        for (int r = 0; r <= NR; ++r) {
            if (r > 0) {
                // ---- transposed mix of dret (round r-1) into every time step's accumulator
#pragma unroll
                for (int tt = 0; tt < RT; ++tt) {
                    const int t = RT * (r - 1) + tt;
                    const float4 dr = *reinterpret_cast<const float4*>(Cb + (size_t)(tt * NC + n) * ROWB + 16 * cg);
                    if (dret != nullptr && row_ok) *reinterpret_cast<float4*>(dret + gofs + t * slab) = dr;
#pragma unroll
                    for (int q = 0; q < 3; ++q) {
                        const float4 m = *reinterpret_cast<const float4*>(Mrow + t * T + 4 * q);
                        fma4(acc[4 * q], m.x, dr);
                        fma4(acc[4 * q + 1], m.y, dr);
                        fma4(acc[4 * q + 2], m.z, dr);
                        fma4(acc[4 * q + 3], m.w, dr);
                    }
                }
            }
            if (r < NR) {
                // ---- dy of the round's four time steps from the landed dOut box and the sign mask -> A operand rows
                tma::mbar_wait(bar + r, it & 1);
#pragma unroll
                for (int tt = 0; tt < RT; ++tt) {
                    const int t = RT * r + tt;
                    float4 dy = *reinterpret_cast<const float4*>(land + ((r * RT + tt) * NC + n) * D + 4 * cg);
                    const uint2 mw = msk[(r * RT + tt) * NC + n];
                    const uint32_t mx = mw.x >> cg, my = mw.y >> cg;           // bit 16*(c & 3) + (c >> 2) of the row's word
                    dy.x = (mx & 1u) ? dy.x : kSlope * dy.x;
                    dy.y = (mx & 0x10000u) ? dy.y : kSlope * dy.y;
                    dy.z = (my & 1u) ? dy.z : kSlope * dy.z;
                    dy.w = (my & 0x10000u) ? dy.w : kSlope * dy.w;
                    acc[t].x += dy.x; acc[t].y += dy.y; acc[t].z += dy.z; acc[t].w += dy.w;
                    put_operand_row(Ab + (size_t)(tt * NC + n) * ROWB, scl + tt * NC + n, cg, dy);
                }
                __syncthreads();
                if (tid == 0 && has_next) issue(task + tstride, r, buf ^ 1);
                mma_round(Ab, Cb, scl, wfrag_ptr(wfrag, b * T + RT * r + tl, half, lane), cur, tl, half, lane);
                if (r + 1 < NR) load_kb0(cur, wfrag_ptr(wfrag, b * T + RT * (r + 1) + tl, half, lane));
                __syncthreads();
            } else if (row_ok) {
#pragma unroll
                for (int t = 0; t < T; ++t) *reinterpret_cast<float4*>(deb + gofs + t * slab) = acc[t];
            }
        }
    }
}

// host: tensor maps for one (B, T, N, 64) activation (box of `bt` time steps x 16 nodes) and for M_n (N, 144)
static int make_maps(CUtensorMap* act, CUtensorMap* mn, const float* x, const float* Mn, int B, int N, int bt) {
    const uint64_t d4[4] = {(uint64_t)D, (uint64_t)N, (uint64_t)T, (uint64_t)B};
    const uint32_t b4[4] = {(uint32_t)D, (uint32_t)NC, (uint32_t)bt, 1u};
    int rc = tma::make_map_f32(act, x, 4, d4, b4);
    if (rc) return rc;
    const uint64_t d2[2] = {(uint64_t)(T * T), (uint64_t)N};
    const uint32_t b2[2] = {(uint32_t)(T * T), (uint32_t)NC};
    return tma::make_map_f32(mn, Mn, 2, d2, b2);
}

static int grid_for(int ntasks, const void* kernel, int smem) {
    int dev = 0, sms = 148, per = 2;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per, kernel, NTHREADS, smem) != cudaSuccess || per < 1) per = 1;
    const int g = sms * per;
    return ntasks < g ? ntasks : g;
}

}  // namespace htf
}  // namespace gptst

using namespace gptst;

// bytes of one fragment table for G = B*T weight matrices
extern "C" long gptst_hypertem_wfrag_bytes(int G) { return (long)G * htf::WF_VEC_PER_GROUP * 16; }

// rows of sign-mask words the backward may touch past the last row (bulk copies fetch 16 rows at a time)
extern "C" int gptst_hypertem_mask_pad_rows(void) { return htf::NC; }

extern "C" int gptst_hypertem_pack_w(const float* W, void* wf, void* wb, int G, void* stream) {
    if (!W || (!wf && !wb) || G <= 0) return -1;
    htf::htem_pack_kernel<<<G, 256, 0, (cudaStream_t)stream>>>(W, (uint4*)wf, (uint4*)wb);
    return (int)cudaGetLastError();
}

extern "C" int gptst_hypertem_fwd(const float* eb, const float* Mn, const void* wfrag, const float* bias, float* out, void* mask,
                                  float* ret, int B, int T, int N, int D, void* stream) {
    if (!eb || !Mn || !wfrag || !bias || !out || !mask || B <= 0 || N <= 0) return -1;
    if (T != htf::T || D != htf::D) return -2;
    CUtensorMap ta, tm;
    int rc = htf::make_maps(&ta, &tm, eb, Mn, B, N, htf::T);
    if (rc) return rc;
    cudaError_t e = cudaFuncSetAttribute(htf::htem_fwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, htf::FwdSmem::total);
    if (e != cudaSuccess) return (int)e;
    const int nch = (N + htf::NC - 1) / htf::NC, ntasks = B * nch;
    const int grid = htf::grid_for(ntasks, (const void*)htf::htem_fwd_kernel, htf::FwdSmem::total);
    launch_pdl(htf::htem_fwd_kernel, dim3(grid), dim3(htf::NTHREADS), (size_t)htf::FwdSmem::total, (cudaStream_t)stream, ta, tm, (const uint4*)wfrag, bias, out, (uint2*)mask,
                                                                                            ret, N, nch, ntasks);
    return (int)cudaGetLastError();
}

extern "C" int gptst_hypertem_bwd(const float* dout, const void* mask, const float* Mn, const void* wfrag_t, float* deb, float* dret,
                                  int B, int T, int N, int D, void* stream) {
    if (!dout || !mask || !Mn || !wfrag_t || !deb || B <= 0 || N <= 0) return -1;
    if (T != htf::T || D != htf::D) return -2;
    CUtensorMap ta, tm;
    int rc = htf::make_maps(&ta, &tm, dout, Mn, B, N, htf::RT);
    if (rc) return rc;
    cudaError_t e = cudaFuncSetAttribute(htf::htem_bwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, htf::BwdSmem::total);
    if (e != cudaSuccess) return (int)e;
    const int nch = (N + htf::NC - 1) / htf::NC, ntasks = B * nch;
    const int grid = htf::grid_for(ntasks, (const void*)htf::htem_bwd_kernel, htf::BwdSmem::total);
    htf::htem_bwd_kernel<<<grid, htf::NTHREADS, htf::BwdSmem::total, (cudaStream_t)stream>>>(ta, tm, (const uint4*)wfrag_t, (const uint2*)mask, deb,
                                                                                            dret, N, nch, ntasks);
    return (int)cudaGetLastError();
}

