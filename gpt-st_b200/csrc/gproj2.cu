// Grouped D x D projection with per-group weights, second generation for D = 64 (reference GPTST.py:160-163, :137-141,
// :24-32): forward  Y = act(X W_g + b_g (+ Res))  and backward  dX = dy W_g^T, dW_g = X^T dy, db_g = sum dy, dRes = dy.
//
// mma.sync m16n8k16 with three-term fp16-split operands (mma_f16.cuh), one warp per 16-row tile, one CTA per
// (group, chunk of NW tiles).  Rows are staged by the warp that owns them with cp.async (16-byte chunks of the 256-byte
// rows, so the node-grouped gather with row stride N*D is as sector-efficient as the contiguous time-grouped case).
//   forward : the A operand is fetched with ldmatrix straight from the fp32 rows (k permutation shared with the staged
//             W_g, see cap_route2_fwd.cu), W_g is pre-split once per CTA, the residual is staged like X.
//   backward: dY and X are converted in place to fp16 hi|lo planes with one power-of-two scale per CTA chunk (gradients
//             have arbitrary magnitude); dX is a per-warp product; dW_g is accumulated over the chunk's rows by letting
//             every warp own ~32/NW of the 32 (16 x 8) output tiles and sweep all row tiles, so the accumulators stay in
//             registers across the chunks of a split and no shared-memory reduction is needed.  dW_part / dbias_part
//             keep the (splits, G, ...) layout of the first generation; everything is deterministic.
#include <cstdlib>

#include "common.cuh"
#include "mma_f16.cuh"

namespace gptst {
namespace gp2 {

using namespace hf;
constexpr int D = 64;
constexpr float WSCALE = 64.f;

__device__ __forceinline__ void ldsm_x2_t(uint32_t& r0, uint32_t& r1, uint32_t addr) {
    asm volatile("ldmatrix.sync.aligned.m8n8.x2.trans.shared.b16 {%0,%1}, [%2];" : "=r"(r0), "=r"(r1) : "r"(addr) : "memory");
}

// the warp's 16 rows r0..r0+15 (fp32, 256 B each) -> its 16 slots; rows >= R are zero-filled
__device__ __forceinline__ void stage16(unsigned char* slots, const float* base, long rs, int r0, int R, int lane) {
#pragma unroll
    for (int k = 0; k < 8; ++k) {
        const int id = lane + 32 * k, r = id >> 4, ch = id & 15;
        unsigned char* dst = slots + (size_t)r * ROWB + ch * 16;
        if (r0 + r < R) cp_async16(dst, base + (long)(r0 + r) * rs + ch * 4);
        else *reinterpret_cast<float4*>(dst) = make_float4(0.f, 0.f, 0.f, 0.f);
    }
}

__device__ __forceinline__ int kperm(int j) {   // physical k (mod 16) -> logical MMA k of the ldmatrix-from-fp32 A operand
    return (j < 4) ? 2 * j : (j < 8) ? 2 * (j - 4) + 1 : (j < 12) ? 8 + 2 * (j - 8) : 8 + 2 * (j - 12) + 1;
}

// ------------------------------------------------------------------------------------------------------------------
// forward
// ------------------------------------------------------------------------------------------------------------------
template <int NW, int MINB, int PREC>
__global__ void __launch_bounds__(NW * 32, MINB)
gproj2_fwd_kernel(const float* __restrict__ X, const float* __restrict__ W, const float* __restrict__ bias,
                  const float* __restrict__ Res, float* __restrict__ Y, int R, long gs, long rs, int act, int chunks,
                  const float* __restrict__ Gate, float* __restrict__ Zout) {
    // act: 0 none, 1 LeakyReLU, 2 fusion gate (reference model/Model.py:12-17): z = sigmoid(X W + b + Res),
    //      Y = z * Gate + (1 - z) * X   (X = the operand rows themselves, still fp32 in shared memory), z -> Zout if given
    extern __shared__ __align__(128) unsigned char smraw[];
    unsigned char* Xs = smraw;                                   // [NW*16][ROWB]
    unsigned char* Wt = Xs + (size_t)NW * 16 * ROWB;             // [64][ROWB]  W_g planes, row = logical k
    float* bs = reinterpret_cast<float*>(Wt + (size_t)D * ROWB); // [64]
    unsigned char* Rs = reinterpret_cast<unsigned char*>(bs + D);// [NW*16][ROWB] (only when Res != nullptr)

    constexpr int NT = NW * 32;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int g = lane >> 2, t = lane & 3;
    const int grp = blockIdx.x / chunks, chunk = blockIdx.x % chunks;
    const int n0 = warp * 16;
    const int r0 = chunk * NW * 16 + n0;
    const float* Xg = X + (long)grp * gs;
    stage16(Xs + (size_t)n0 * ROWB, Xg, rs, r0, R, lane);
    if (Res) stage16(Rs + (size_t)n0 * ROWB, Res + (long)grp * gs, rs, r0, R, lane);
    const float* Wg = W + (size_t)grp * D * D;
    {   // all of this thread's W_g loads are issued before the first one is consumed (a rolled loop serialises the latencies)
        constexpr int WI = (D * 16 + NT - 1) / NT;
        float4 wv[WI];
#pragma unroll
        for (int u = 0; u < WI; ++u) {
            const int i = tid + u * NT;
            wv[u] = (i < D * 16) ? *reinterpret_cast<const float4*>(Wg + (size_t)(i >> 4) * D + (i & 15) * 4) : make_float4(0.f, 0.f, 0.f, 0.f);
        }
#pragma unroll
        for (int u = 0; u < WI; ++u) {
            const int i = tid + u * NT;
            if (i < D * 16) {
                const int k = i >> 4, q4 = i & 15;
                uint32_t h0, l0, h1, l1;
                split_h2<PREC>(wv[u].x * WSCALE, wv[u].y * WSCALE, h0, l0);
                split_h2<PREC>(wv[u].z * WSCALE, wv[u].w * WSCALE, h1, l1);
                unsigned char* row = Wt + (size_t)(16 * (k >> 4) + kperm(k & 15)) * ROWB + q4 * 8;
                *reinterpret_cast<uint2*>(row) = make_uint2(h0, h1);
                *reinterpret_cast<uint2*>(row + LO) = make_uint2(l0, l1);
            }
        }
    }
    for (int i = tid; i < D; i += NT) bs[i] = bias ? bias[(size_t)grp * D + i] : 0.f;
    cp_async_wait_all();
    __syncthreads();
    if (r0 >= R) return;   // whole tile out of range (no further block-wide barrier below)

    float acc[8][4];
#pragma unroll
    for (int j = 0; j < 8; ++j) acc[j][0] = acc[j][1] = acc[j][2] = acc[j][3] = 0.f;
#pragma unroll
    for (int b = 0; b < 4; ++b) {
        uint32_t f[4], ah[4], al[4];
        const uint32_t aaddr = smem_u32(Xs + (size_t)(n0 + (lane & 7)) * ROWB + (16 * b + 4 * (lane >> 3)) * 4);
        ldsm_x4(f, aaddr);
        split_h2<PREC>(__uint_as_float(f[0]), __uint_as_float(f[1]), ah[0], al[0]);
        split_h2<PREC>(__uint_as_float(f[2]), __uint_as_float(f[3]), ah[2], al[2]);
        ldsm_x4(f, aaddr + 8 * ROWB);
        split_h2<PREC>(__uint_as_float(f[0]), __uint_as_float(f[1]), ah[1], al[1]);
        split_h2<PREC>(__uint_as_float(f[2]), __uint_as_float(f[3]), ah[3], al[3]);
#pragma unroll
        for (int jp = 0; jp < 4; ++jp) {
            uint32_t bh[4], bl[4] = {0u, 0u, 0u, 0u};
            const uint32_t baddr =
                smem_u32(Wt + (size_t)(16 * b + 8 * ((lane >> 3) & 1) + (lane & 7)) * ROWB + (16 * jp + 8 * (lane >> 4)) * 2);
            ldsm_x4_t(bh, baddr);
            if (PREC == PREC_3XTF32) ldsm_x4_t(bl, baddr + LO);
            mma3<PREC>(acc[2 * jp], ah, al, bh[0], bh[1], bl[0], bl[1]);
            mma3<PREC>(acc[2 * jp + 1], ah, al, bh[2], bh[3], bl[2], bl[3]);
        }
    }
    constexpr float inv = 1.f / WSCALE;
    float* Yg = Y + (long)grp * gs;
#pragma unroll
    for (int half = 0; half < 2; ++half) {
        const int rl = n0 + g + 8 * half, rg = r0 + g + 8 * half;
        if (rg < R) {
#pragma unroll
            for (int j = 0; j < 8; ++j) {
                const int col = 8 * j + 2 * t;
                float y0 = fmaf(acc[j][2 * half], inv, bs[col]), y1 = fmaf(acc[j][2 * half + 1], inv, bs[col + 1]);
                if (Res) {
                    const float2 rr = *reinterpret_cast<const float2*>(Rs + (size_t)rl * ROWB + col * 4);
                    y0 += rr.x; y1 += rr.y;
                }
                if (act == 1) { y0 = lrelu(y0); y1 = lrelu(y1); }
                else if (act == 2) {
                    const float z0 = 1.f / (1.f + expf(-y0)), z1 = 1.f / (1.f + expf(-y1));
                    const float2 xr = *reinterpret_cast<const float2*>(Xs + (size_t)rl * ROWB + col * 4);
                    const float2 fr = *reinterpret_cast<const float2*>(Gate + (long)grp * gs + (long)rg * rs + col);
                    if (Zout) *reinterpret_cast<float2*>(Zout + (long)grp * gs + (long)rg * rs + col) = make_float2(z0, z1);
                    y0 = z0 * fr.x + (1.f - z0) * xr.x;
                    y1 = z1 * fr.y + (1.f - z1) * xr.y;
                }
                *reinterpret_cast<float2*>(Yg + (long)rg * rs + col) = make_float2(y0, y1);
            }
        }
    }
}

// ------------------------------------------------------------------------------------------------------------------
// backward
// ------------------------------------------------------------------------------------------------------------------
template <int NW, int MINB, int PREC>
__global__ void __launch_bounds__(NW * 32, MINB)
gproj2_bwd_kernel(const float* __restrict__ dY, const float* __restrict__ Y, const float* __restrict__ X,
                  const float* __restrict__ W, float* __restrict__ dX, float* __restrict__ dWp, float* __restrict__ dbp,
                  float* __restrict__ dRes, int G, int R, long gs, long rs, int act, int cps, int flags) {
    // flags: bit 0 = dX is accumulated in place (dX += dy W^T); bit 1 = W and dW are [out][in] (a shared nn.Linear weight)
    constexpr int TPW = (32 + NW - 1) / NW;          // (16 x 8) dW output tiles per warp
    extern __shared__ __align__(128) unsigned char smraw[];
    unsigned char* Xs = smraw;                                      // [NW*16][ROWB]
    unsigned char* Gs = Xs + (size_t)NW * 16 * ROWB;                // [NW*16][ROWB]
    unsigned char* Wt = Gs + (size_t)NW * 16 * ROWB;                // [64][ROWB]  row = in index, planes along out
    float* cmax = reinterpret_cast<float*>(Wt + (size_t)D * ROWB);  // [2*NW]
    float* dbred = reinterpret_cast<float*>(Xs);                    // [NW][64], aliases the X slots after the chunk loop

    constexpr int NT = NW * 32;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int g = lane >> 2, t = lane & 3;
    const int grp = blockIdx.x, split = blockIdx.y;
    const int n0 = warp * 16;
    const float* Xg = X + (long)grp * gs;
    const float* dYg = dY + (long)grp * gs;
    const float* Yg = Y ? Y + (long)grp * gs : nullptr;
    float* dXg = dX + (long)grp * gs;
    float* dRg = dRes ? dRes + (long)grp * gs : nullptr;

    // the first chunk's dY / X rows start streaming in before the weight tile is fetched and converted (two cp.async groups,
    // consumed in the chunk loop below): the weight's global latency used to sit in front of them
    if (split * cps * NW * 16 < R) {
        const int r0 = split * cps * NW * 16 + n0;
        stage16(Gs + (size_t)n0 * ROWB, dYg, rs, r0, R, lane);
        cp_async_commit();
        stage16(Xs + (size_t)n0 * ROWB, Xg, rs, r0, R, lane);
        cp_async_commit();
    }
    const float* Wg = W + (size_t)grp * D * D;
    {   // issue all W_g loads of this thread first, convert afterwards
        constexpr int WI = (D * 16 + NT - 1) / NT;
        float4 wv[WI];
#pragma unroll
        for (int u = 0; u < WI; ++u) {
            const int i = tid + u * NT;
            const int k = i >> 4, q4 = i & 15;        // k = in index, columns out = 4*q4 .. 4*q4+3
            if (i >= D * 16) wv[u] = make_float4(0.f, 0.f, 0.f, 0.f);
            else if (flags & 2) wv[u] = make_float4(Wg[(size_t)(4 * q4) * D + k], Wg[(size_t)(4 * q4 + 1) * D + k],
                                                    Wg[(size_t)(4 * q4 + 2) * D + k], Wg[(size_t)(4 * q4 + 3) * D + k]);
            else wv[u] = *reinterpret_cast<const float4*>(Wg + (size_t)k * D + q4 * 4);
        }
#pragma unroll
        for (int u = 0; u < WI; ++u) {
            const int i = tid + u * NT;
            if (i < D * 16) {
                const int k = i >> 4, q4 = i & 15;
                uint32_t h0, l0, h1, l1;
                split_h2<PREC>(wv[u].x * WSCALE, wv[u].y * WSCALE, h0, l0);
                split_h2<PREC>(wv[u].z * WSCALE, wv[u].w * WSCALE, h1, l1);
                unsigned char* row = Wt + (size_t)k * ROWB + q4 * 8;
                *reinterpret_cast<uint2*>(row) = make_uint2(h0, h1);
                *reinterpret_cast<uint2*>(row + LO) = make_uint2(l0, l1);
            }
        }
    }

    float dwm[TPW][4];
#pragma unroll
    for (int i = 0; i < TPW; ++i) dwm[i][0] = dwm[i][1] = dwm[i][2] = dwm[i][3] = 0.f;
    float dbl[4] = {0.f, 0.f, 0.f, 0.f};

    for (int it = 0; it < cps; ++it) {
        const int rbase = (split * cps + it) * NW * 16;
        if (rbase >= R) break;                                    // uniform over the CTA
        const int r0 = rbase + n0;
        unsigned char* Xw = Xs + (size_t)n0 * ROWB;
        unsigned char* Gw = Gs + (size_t)n0 * ROWB;
        // dY first, X second (two cp.async groups): the dy conversion and the dX product only need dY, so the X rows
        // keep streaming in underneath them; Y goes straight to registers so its latency overlaps the staging as well
        if (it > 0) {                                             // chunk 0 was issued ahead of the weight staging
            stage16(Gw, dYg, rs, r0, R, lane);
            cp_async_commit();
            stage16(Xw, Xg, rs, r0, R, lane);
            cp_async_commit();
        }
        constexpr bool PREY = NW * 32 * MINB <= 512;   // room for 32 more live registers (cap >= 128 per thread)
        float4 yv[PREY ? 8 : 1];
        if (PREY && act) {
#pragma unroll
            for (int k = 0; k < 8; ++k) {
                const int id = lane + 32 * k, r = id >> 4, ch = id & 15;
                yv[k] = (r0 + r < R) ? *reinterpret_cast<const float4*>(Yg + (long)(r0 + r) * rs + ch * 4)
                                     : make_float4(1.f, 1.f, 1.f, 1.f);
            }
        }
        cp_async_wait_group<1>();
        __syncwarp();
        // ---- max |dY| of the chunk -> one power-of-two scale
        {
            float mg = 0.f;
#pragma unroll
            for (int k = 0; k < 8; ++k) {
                const int id = lane + 32 * k, r = id >> 4, ch = id & 15;
                const float4 b = *reinterpret_cast<const float4*>(Gw + (size_t)r * ROWB + ch * 16);
                mg = fmaxf(mg, fmaxf(fmaxf(fabsf(b.x), fabsf(b.y)), fmaxf(fabsf(b.z), fabsf(b.w))));
            }
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) mg = fmaxf(mg, __shfl_xor_sync(0xffffffffu, mg, o));
            if (lane == 0) cmax[NW + warp] = mg;
        }
        __syncthreads();
        float2 sg;
        {
            float mg = 0.f;
#pragma unroll
            for (int w = 0; w < NW; ++w) mg = fmaxf(mg, cmax[NW + w]);
            sg = pow2_scale_for_fp16(mg);
        }
        // ---- dy = dY * act'(Y) -> dRes, column sums, planes
        {
            float4 f[8];
#pragma unroll
            for (int k = 0; k < 8; ++k) {
                const int id = lane + 32 * k, r = id >> 4, ch = id & 15;
                f[k] = *reinterpret_cast<const float4*>(Gw + (size_t)r * ROWB + ch * 16);
                if (r0 + r < R) {
                    if (act) {
                        const float4 y = PREY ? yv[PREY ? k : 0] : *reinterpret_cast<const float4*>(Yg + (long)(r0 + r) * rs + ch * 4);
                        f[k].x = lrelu_grad(y.x, f[k].x); f[k].y = lrelu_grad(y.y, f[k].y);
                        f[k].z = lrelu_grad(y.z, f[k].z); f[k].w = lrelu_grad(y.w, f[k].w);
                    }
                    if (dRg) *reinterpret_cast<float4*>(dRg + (long)(r0 + r) * rs + ch * 4) = f[k];
                    dbl[0] += f[k].x; dbl[1] += f[k].y; dbl[2] += f[k].z; dbl[3] += f[k].w;
                }
            }
            __syncwarp();
#pragma unroll
            for (int k = 0; k < 8; ++k) {
                const int id = lane + 32 * k, r = id >> 4, ch = id & 15;
                uint32_t h0, l0, h1, l1;
                split_h2<PREC>(f[k].x * sg.x, f[k].y * sg.x, h0, l0);
                split_h2<PREC>(f[k].z * sg.x, f[k].w * sg.x, h1, l1);
                unsigned char* row = Gw + (size_t)r * ROWB + ch * 8;
                *reinterpret_cast<uint2*>(row) = make_uint2(h0, h1);
                *reinterpret_cast<uint2*>(row + LO) = make_uint2(l0, l1);
            }
            __syncwarp();
        }
        // ---- dX = dy W_g^T for the warp's 16 rows (two halves of 4 column tiles to keep the accumulators small)
        {
            const float un = sg.y * (1.f / WSCALE);
#pragma unroll
            for (int hf2 = 0; hf2 < 2; ++hf2) {
                float acc[4][4];
#pragma unroll
                for (int j = 0; j < 4; ++j) acc[j][0] = acc[j][1] = acc[j][2] = acc[j][3] = 0.f;
#pragma unroll
                for (int b = 0; b < 4; ++b) {
                    uint32_t ah[4], al[4] = {0u, 0u, 0u, 0u};
                    const uint32_t aaddr =
                        smem_u32(Gw + (size_t)(8 * ((lane >> 3) & 1) + (lane & 7)) * ROWB + (16 * b + 8 * (lane >> 4)) * 2);
                    ldsm_x4(ah, aaddr);
                    if (PREC == PREC_3XTF32) ldsm_x4(al, aaddr + LO);
#pragma unroll
                    for (int jq = 0; jq < 2; ++jq) {
                        const int jp = 2 * hf2 + jq;
                        uint32_t bh[4], bl[4] = {0u, 0u, 0u, 0u};
                        const uint32_t baddr = smem_u32(Wt + (size_t)(16 * jp + 8 * (lane >> 4) + (lane & 7)) * ROWB +
                                                        (16 * b + 8 * ((lane >> 3) & 1)) * 2);
                        ldsm_x4(bh, baddr);
                        if (PREC == PREC_3XTF32) ldsm_x4(bl, baddr + LO);
                        mma3<PREC>(acc[2 * jq], ah, al, bh[0], bh[1], bl[0], bl[1]);
                        mma3<PREC>(acc[2 * jq + 1], ah, al, bh[2], bh[3], bl[2], bl[3]);
                    }
                }
#pragma unroll
                for (int half = 0; half < 2; ++half) {
                    const int rg = r0 + g + 8 * half;
                    if (rg < R) {
#pragma unroll
                        for (int j = 0; j < 4; ++j) {
                            float2* p = reinterpret_cast<float2*>(dXg + (long)rg * rs + 32 * hf2 + 8 * j + 2 * t);
                            float2 o = make_float2(acc[j][2 * half] * un, acc[j][2 * half + 1] * un);
                            if (flags & 1) { const float2 old = *p; o.x += old.x; o.y += old.y; }
                            *p = o;
                        }
                    }
                }
            }
        }
        // ---- X landed meanwhile: max |X| -> scale -> planes
        cp_async_wait_group<0>();
        __syncwarp();
        {
            float mx = 0.f;
#pragma unroll
            for (int k = 0; k < 8; ++k) {
                const int id = lane + 32 * k, r = id >> 4, ch = id & 15;
                const float4 a = *reinterpret_cast<const float4*>(Xw + (size_t)r * ROWB + ch * 16);
                mx = fmaxf(mx, fmaxf(fmaxf(fabsf(a.x), fabsf(a.y)), fmaxf(fabsf(a.z), fabsf(a.w))));
            }
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, o));
            if (lane == 0) cmax[warp] = mx;
        }
        __syncthreads();
        float2 sx;
        {
            float mx = 0.f;
#pragma unroll
            for (int w = 0; w < NW; ++w) mx = fmaxf(mx, cmax[w]);
            sx = pow2_scale_for_fp16(mx);
            float4 f[8];
#pragma unroll
            for (int k = 0; k < 8; ++k) {
                const int id = lane + 32 * k, r = id >> 4, ch = id & 15;
                f[k] = *reinterpret_cast<const float4*>(Xw + (size_t)r * ROWB + ch * 16);
            }
            __syncwarp();
#pragma unroll
            for (int k = 0; k < 8; ++k) {
                const int id = lane + 32 * k, r = id >> 4, ch = id & 15;
                uint32_t h0, l0, h1, l1;
                split_h2<PREC>(f[k].x * sx.x, f[k].y * sx.x, h0, l0);
                split_h2<PREC>(f[k].z * sx.x, f[k].w * sx.x, h1, l1);
                unsigned char* row = Xw + (size_t)r * ROWB + ch * 8;
                *reinterpret_cast<uint2*>(row) = make_uint2(h0, h1);
                *reinterpret_cast<uint2*>(row + LO) = make_uint2(l0, l1);
            }
        }
        __syncthreads();   // every warp's planes are in place
        // ---- dW_g += X^T dy over the chunk: this warp's output tiles, all row tiles
        {
            float dwt[TPW][4];
#pragma unroll
            for (int i = 0; i < TPW; ++i) dwt[i][0] = dwt[i][1] = dwt[i][2] = dwt[i][3] = 0.f;
            int nks = (R - rbase + 15) / 16;
            nks = nks > NW ? NW : nks;
            for (int ks = 0; ks < nks; ++ks) {
                uint32_t ah[4] = {0u, 0u, 0u, 0u}, al[4] = {0u, 0u, 0u, 0u};
#pragma unroll
                for (int i = 0; i < TPW; ++i) {
                    // a warp's output tiles are CONSECUTIVE ids (mt = id >> 3 = 16-row block of dW, j = id & 7): they share the X^T
                    // fragments of their row block, which are fetched once per row tile and block instead of once per output tile
                    // (the kernel is bound by shared-memory wavefronts; profiles/ncu_route_fwd_r02.md)
                    const int id = warp * TPW + i;
                    if (id < 32) {
                        const int mt = id >> 3, j = id & 7;
                        uint32_t b0, b1, q0 = 0u, q1 = 0u;
                        if (i == 0 || mt != ((id - 1) >> 3)) {
                            const uint32_t aaddr = smem_u32(Xs + (size_t)(16 * ks + 8 * (lane >> 4) + (lane & 7)) * ROWB +
                                                            (16 * mt + 8 * ((lane >> 3) & 1)) * 2);
                            ldsm_x4_t(ah, aaddr);
                            if (PREC == PREC_3XTF32) ldsm_x4_t(al, aaddr + LO);
                        }
                        const uint32_t baddr =
                            smem_u32(Gs + (size_t)(16 * ks + 8 * ((lane >> 3) & 1) + (lane & 7)) * ROWB + (8 * j) * 2);
                        ldsm_x2_t(b0, b1, baddr);
                        if (PREC == PREC_3XTF32) ldsm_x2_t(q0, q1, baddr + LO);
                        mma3<PREC>(dwt[i], ah, al, b0, b1, q0, q1);
                    }
                }
            }
            const float un = sx.y * sg.y;
#pragma unroll
            for (int i = 0; i < TPW; ++i) {
                dwm[i][0] = fmaf(dwt[i][0], un, dwm[i][0]); dwm[i][1] = fmaf(dwt[i][1], un, dwm[i][1]);
                dwm[i][2] = fmaf(dwt[i][2], un, dwm[i][2]); dwm[i][3] = fmaf(dwt[i][3], un, dwm[i][3]);
            }
        }
        __syncthreads();   // before the next chunk overwrites the slots / cmax
    }
    // ---- results of this split
    float* dWo = dWp + ((size_t)split * G + grp) * D * D;
#pragma unroll
    for (int i = 0; i < TPW; ++i) {
        const int id = warp * TPW + i;
        if (id < 32) {
            const int mt = id >> 3, j = id & 7;
            const int rin = 16 * mt + g, col = 8 * j + 2 * t;
            if (flags & 2) {
                dWo[(size_t)col * D + rin] = dwm[i][0];       dWo[(size_t)(col + 1) * D + rin] = dwm[i][1];
                dWo[(size_t)col * D + rin + 8] = dwm[i][2];   dWo[(size_t)(col + 1) * D + rin + 8] = dwm[i][3];
            } else {
                *reinterpret_cast<float2*>(dWo + (size_t)rin * D + col) = make_float2(dwm[i][0], dwm[i][1]);
                *reinterpret_cast<float2*>(dWo + (size_t)(rin + 8) * D + col) = make_float2(dwm[i][2], dwm[i][3]);
            }
        }
    }
#pragma unroll
    for (int e = 0; e < 4; ++e) dbl[e] += __shfl_xor_sync(0xffffffffu, dbl[e], 16);
    if (lane < 16) *reinterpret_cast<float4*>(dbred + (size_t)warp * D + 4 * lane) = make_float4(dbl[0], dbl[1], dbl[2], dbl[3]);
    __syncthreads();
    if (tid < D) {
        float s = 0.f;
#pragma unroll
        for (int w = 0; w < NW; ++w) s += dbred[(size_t)w * D + tid];
        dbp[((size_t)split * G + grp) * D + tid] = s;
    }
}

// ------------------------------------------------------------------------------------------------------------------
// launch policy
// ------------------------------------------------------------------------------------------------------------------
static int env_int(const char* name, int dflt) {
    const char* e = getenv(name);
    return e ? atoi(e) : dflt;
}
static int pick_nw(int R) {   // warps (= 16-row tiles) per CTA chunk
    static const int mid = env_int("GPTST_B200_GP2_MID", 11);    // 129..176 rows: 11 (one chunk) or 6 (two chunks, 3 CTAs/SM)
    static const int lng = env_int("GPTST_B200_GP2_LONG", 8);    // long groups: chunks of 128 rows (8) or 64 rows (4)
    if (R <= 64) return 4;
    if (R <= 96) return 6;
    if (R <= 128) return 8;
    if (R <= 176) return mid == 6 ? 6 : 11;
    if (R <= 208) return 13;
    if (R <= 256) return 16;
    return lng == 4 ? 4 : (lng == 16 ? 16 : 8);   // long groups (node-grouped: R = B*T)
}

template <int NW, int MINB, int PREC>
static cudaError_t launch_fwd(const float* X, const float* W, const float* bias, const float* Res, float* Y, int G, int R,
                              long gs, long rs, int act, const float* Gate, float* Zout, cudaStream_t st) {
    const int chunks = (R + NW * 16 - 1) / (NW * 16);
    const size_t smem = (size_t)NW * 16 * ROWB * (Res ? 2 : 1) + (size_t)D * ROWB + D * 4;
    auto kern = gproj2_fwd_kernel<NW, MINB, PREC>;
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return e;
    kern<<<(unsigned)((size_t)G * chunks), NW * 32, smem, st>>>(X, W, bias, Res, Y, R, gs, rs, act, chunks, Gate, Zout);
    return cudaGetLastError();
}

template <int NW, int MINB, int PREC>
static cudaError_t launch_bwd(const float* dY, const float* Y, const float* X, const float* W, float* dX, float* dWp,
                              float* dbp, float* dRes, int G, int R, long gs, long rs, int act, int splits, int flags, cudaStream_t st) {
    const int chunks = (R + NW * 16 - 1) / (NW * 16);
    const int cps = (chunks + splits - 1) / splits;
    const size_t smem = (size_t)2 * NW * 16 * ROWB + (size_t)D * ROWB + (size_t)(2 * NW) * 4;
    auto kern = gproj2_bwd_kernel<NW, MINB, PREC>;
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return e;
    kern<<<dim3(G, splits), NW * 32, smem, st>>>(dY, Y, X, W, dX, dWp, dbp, dRes, G, R, gs, rs, act, cps, flags);
    return cudaGetLastError();
}

#define GP2_DISPATCH(FN, ...)                                         \
    switch (pick_nw(R)) {                                             \
        case 4: return FN<4, 4, PREC>(__VA_ARGS__);                   \
        case 6: return FN<6, 3, PREC>(__VA_ARGS__);                   \
        case 8: return FN<8, 2, PREC>(__VA_ARGS__);                   \
        case 11: return FN<11, 2, PREC>(__VA_ARGS__);                 \
        case 13: return FN<13, 1, PREC>(__VA_ARGS__);                 \
        default: return FN<16, 1, PREC>(__VA_ARGS__);                 \
    }

template <int PREC>
static cudaError_t fwd_p(const float* X, const float* W, const float* bias, const float* Res, float* Y, int G, int R, long gs,
                         long rs, int act, const float* Gate, float* Zout, cudaStream_t st) {
    GP2_DISPATCH(launch_fwd, X, W, bias, Res, Y, G, R, gs, rs, act, Gate, Zout, st)
}
template <int PREC>
static cudaError_t bwd_p(const float* dY, const float* Y, const float* X, const float* W, float* dX, float* dWp, float* dbp,
                         float* dRes, int G, int R, long gs, long rs, int act, int splits, int flags, cudaStream_t st) {
    GP2_DISPATCH(launch_bwd, dY, Y, X, W, dX, dWp, dbp, dRes, G, R, gs, rs, act, splits, flags, st)
}

}  // namespace gp2

// number of dW / dbias partials per group the backward writes: the chunks-per-split that minimises
// (rounds of 2 CTAs per SM) x (chunks per CTA), ties broken towards fewer partials
int gproj2_splits(int G, int R) {
    const int nw = gp2::pick_nw(R);
    const int chunks = (R + nw * 16 - 1) / (nw * 16);
    const long slots = 2 * 148;
    long best_cost = -1;
    int best = 1;
    for (int cps = chunks; cps >= 1; --cps) {
        const int splits = (chunks + cps - 1) / cps;
        const long rounds = ((long)G * splits + slots - 1) / slots;
        const long cost = rounds * cps * 64 + splits;     // partials cost a little (one more (D,D) write + the final sum)
        if (best_cost < 0 || cost < best_cost) { best_cost = cost; best = splits; }
    }
    return best;
}

cudaError_t gproj2_fwd(const float* X, const float* W, const float* bias, const float* Res, float* Y, int G, int R, long gs,
                       long rs, int act, int prec, cudaStream_t st, const float* Gate, float* Zout) {
    if (act == 2 && !Gate) return cudaErrorInvalidValue;
    if (prec == PREC_3XTF32) return gp2::fwd_p<PREC_3XTF32>(X, W, bias, Res, Y, G, R, gs, rs, act, Gate, Zout, st);
    return gp2::fwd_p<PREC_TF32>(X, W, bias, Res, Y, G, R, gs, rs, act, Gate, Zout, st);
}

cudaError_t gproj2_bwd(const float* dY, const float* Y, const float* X, const float* W, float* dX, float* dWp, float* dbp,
                       float* dRes, int G, int R, long gs, long rs, int act, int prec, int splits, int flags, cudaStream_t st) {
    if (prec == PREC_3XTF32) return gp2::bwd_p<PREC_3XTF32>(dY, Y, X, W, dX, dWp, dbp, dRes, G, R, gs, rs, act, splits, flags, st);
    return gp2::bwd_p<PREC_TF32>(dY, Y, X, W, dX, dWp, dbp, dRes, G, R, gs, rs, act, splits, flags, st);
}

}  // namespace gptst

using namespace gptst;

// ---- backward of a plain linear layer with a shared (D,D) weight, dX accumulated in place -------------------------
//   dX_io += dY W ;  dW_part[s] = partial dY^T X ([out][in], like nn.Linear.weight) ;  db_part[s] = partial sum dY
// (rows, D) row-major operands.  Used for the ln_p layer of cap (GPTST.py:102): dx += dZ Wp, dWp = dZ^T x, dbp = sum dZ.
extern "C" int gptst_linear_bwd_acc_splits(long rows, int D) {
    if (D != 64 || rows <= 0 || rows > 0x7fffffffL) return -2;
    return gproj2_splits(1, (int)rows);
}
extern "C" int gptst_linear_bwd_acc(const float* dY, const float* X, const float* W, float* dX_io, float* dW_part,
                                    float* db_part, long rows, int D, int prec, int splits, void* stream) {
    if (!dY || !X || !W || !dX_io || !dW_part || !db_part || rows <= 0 || splits <= 0) return -1;
    if (D != 64 || rows > 0x7fffffffL || (prec != 1 && prec != 3)) return -2;
    return (int)gproj2_bwd(dY, nullptr, X, W, dX_io, dW_part, db_part, nullptr, 1, (int)rows, 0, D, 0, prec, splits, 3,
                           (cudaStream_t)stream);
}
