// Sign-mask backward of the grouped D x D projection (D = 64): same decomposition as gproj2.cu (one warp per 16-row tile,
// three-term fp16 split, cp.async staged rows) with two differences that profiles/ncu_gproj2_bwd_r01.md asked for:
//   * LeakyReLU's derivative comes from a packed SIGN MASK of the forward output (8 bytes per row) instead of re-reading Y:
//     traffic 5A -> 4A and no exposed global latency on the Y rows;
//   * ONE power-of-two scale per WARP tile instead of per CTA chunk (no block-wide max exchange): for dW the un-scaling moves
//     into the row-tile loop (dW += dW_tile * (sx_tile * sg_tile)).
// Default use (round 2): flags 4|8 = the dW_bt / db_bt kernel of the fused hyperTem backward (gptst_hypertem_dw), fed by the
// mask the fused forward writes.  The general entry point gptst_gproj3_bwd is kept for the stand-alone checks.
#include <cstdlib>

#include "common.cuh"
#include "mma_f16.cuh"

namespace gptst {
namespace gp3 {

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

__device__ __forceinline__ void prefetch_l2(const void* p) { asm volatile("prefetch.global.L2 [%0];" ::"l"(p)); }

// the two 128-byte lines of each of the warp's 16 rows (the read-modify-write of dX in the epilogue then finds them in L2)
__device__ __forceinline__ void prefetch_rows16(const float* base, long rs, int r0, int R, int lane) {
    const int r = lane >> 1, half = lane & 1;
    if (r0 + r < R) prefetch_l2(base + (long)(r0 + r) * rs + 32 * half);
}

__device__ __forceinline__ int kperm(int j) {   // physical k (mod 16) -> logical MMA k of the ldmatrix-from-fp32 A operand
    return (j < 4) ? 2 * j : (j < 8) ? 2 * (j - 4) + 1 : (j < 12) ? 8 + 2 * (j - 8) : 8 + 2 * (j - 12) + 1;
}

// ------------------------------------------------------------------------------------------------------------------
// backward
// ------------------------------------------------------------------------------------------------------------------
template <int NW, int MINB, int PREC>
__global__ void __launch_bounds__(NW * 32, MINB)
gproj3_bwd_kernel(const float* __restrict__ dY, const uint2* __restrict__ Mask, const float* __restrict__ X,
                  const float* __restrict__ W, float* __restrict__ dX, float* __restrict__ dWp, float* __restrict__ dbp,
                  float* __restrict__ dRes, int G, int R, long gs, long rs, int act, int cps, int flags, int mrs) {
    // flags: bit 0 = dX is accumulated in place (dX += dy W^T); bit 1 = W and dW are [out][in] (a shared nn.Linear weight);
    //        bit 2 = dW / db only (no W, no dX: the side-stream half of the fused hyperTem backward, csrc/htem_fused.cu);
    //        bit 3 = the mask comes from the fused hyperTem forward: `mrs` words per group, bit 16*(c & 3) + (c >> 2) = column c
    constexpr int TPW = (32 + NW - 1) / NW;          // (16 x 8) dW output tiles per warp
    extern __shared__ __align__(128) unsigned char smraw[];
    unsigned char* Xs = smraw;                                      // [NW*16][ROWB]
    unsigned char* Gs = Xs + (size_t)NW * 16 * ROWB;                // [NW*16][ROWB]
    unsigned char* Wt = Gs + (size_t)NW * 16 * ROWB;                // [64][ROWB]  row = in index, planes along out
    float* tsc = reinterpret_cast<float*>(Wt + (size_t)D * ROWB);   // [NW]  un-scale (1/sx * 1/sg) of every warp tile
    float* dbred = reinterpret_cast<float*>(Xs);                    // [NW][64], aliases the X slots after the chunk loop

    constexpr int NT = NW * 32;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int g = lane >> 2, t = lane & 3;
    const int grp = blockIdx.x, split = blockIdx.y;
    const int n0 = warp * 16;
    const float* Xg = X + (long)grp * gs;
    const float* dYg = dY + (long)grp * gs;
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
        if (flags & 1) prefetch_rows16(dXg, rs, r0, R, lane);
    }
    const float* Wg = W + (size_t)grp * D * D;
    if (!(flags & 4)) {   // issue all W_g loads of this thread first, convert afterwards
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

    __syncthreads();   // the W_g planes are read by every warp's dX product (gproj2 got this ordering from its cmax barrier)

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
            if (flags & 1) prefetch_rows16(dXg, rs, r0, R, lane);
        }
        // sign mask of the warp's 16 rows: lane r < 16 holds row r0 + r (one 8-byte load, issued before the dY wait)
        uint32_t mlo = 0xffffffffu, mhi = 0xffffffffu;
        if (act && lane < 16 && r0 + lane < R) {
            const uint2 m = (flags & 8) ? Mask[(long)grp * mrs + r0 + lane] : Mask[((long)grp * gs + (long)(r0 + lane) * rs) / D];
            mlo = m.x; mhi = m.y;
        }
        cp_async_wait_group<1>();
        __syncwarp();
        // ---- max |dY| of the WARP tile -> its power-of-two scale (no block-wide exchange)
        float2 sg;
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
            sg = pow2_scale_for_fp16(mg);
        }
        // ---- dy = dY * act'(Y) -> dRes, column sums, planes
        {
            float4 f[8];
#pragma unroll
            for (int k = 0; k < 8; ++k) {
                const int id = lane + 32 * k, r = id >> 4, ch = id & 15;
                f[k] = *reinterpret_cast<const float4*>(Gw + (size_t)r * ROWB + ch * 16);
                // the row's mask words come from lane r (every lane takes part in the shuffles); bits 4ch .. 4ch+3 are this chunk's
                const uint32_t lo_r = __shfl_sync(0xffffffffu, mlo, r), hi_r = __shfl_sync(0xffffffffu, mhi, r);
                uint32_t bits = ((ch < 8 ? lo_r : hi_r) >> ((4 * ch) & 31)) & 15u;
                if (flags & 8) {
                    const uint32_t mx = lo_r >> ch, my = hi_r >> ch;
                    bits = (mx & 1u) | ((mx >> 15) & 2u) | ((my & 1u) << 2) | ((my >> 13) & 8u);
                }
                if (r0 + r < R) {
                    if (act) {
                        f[k].x = (bits & 1u) ? f[k].x : kSlope * f[k].x; f[k].y = (bits & 2u) ? f[k].y : kSlope * f[k].y;
                        f[k].z = (bits & 4u) ? f[k].z : kSlope * f[k].z; f[k].w = (bits & 8u) ? f[k].w : kSlope * f[k].w;
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
        if (!(flags & 4)) {
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
        // ---- X landed meanwhile: max |X| of the warp tile -> scale -> planes
        cp_async_wait_group<0>();
        __syncwarp();
        {
            float mx = 0.f;
            float4 f[8];
#pragma unroll
            for (int k = 0; k < 8; ++k) {
                const int id = lane + 32 * k, r = id >> 4, ch = id & 15;
                f[k] = *reinterpret_cast<const float4*>(Xw + (size_t)r * ROWB + ch * 16);
                mx = fmaxf(mx, fmaxf(fmaxf(fabsf(f[k].x), fabsf(f[k].y)), fmaxf(fabsf(f[k].z), fabsf(f[k].w))));
            }
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, o));
            const float2 sx = pow2_scale_for_fp16(mx);
            if (lane == 0) tsc[warp] = sx.y * sg.y;
            __syncwarp();                    // every lane has read its fp32 values before the planes overwrite the rows
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
        __syncthreads();   // every warp's planes and tile scales are in place
        // ---- dW_g += X^T dy over the chunk: this warp's output tiles, all row tiles
        {
            int nks = (R - rbase + 15) / 16;
            nks = nks > NW ? NW : nks;
            for (int ks = 0; ks < nks; ++ks) {
                const float un = tsc[ks];
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
                        float t4[4] = {0.f, 0.f, 0.f, 0.f};
                        mma3<PREC>(t4, ah, al, b0, b1, q0, q1);
                        dwm[i][0] = fmaf(t4[0], un, dwm[i][0]); dwm[i][1] = fmaf(t4[1], un, dwm[i][1]);
                        dwm[i][2] = fmaf(t4[2], un, dwm[i][2]); dwm[i][3] = fmaf(t4[3], un, dwm[i][3]);
                    }
                }
            }
        }
        __syncthreads();   // before the next chunk overwrites the slots / tile scales
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
static cudaError_t launch_bwd(const float* dY, const uint2* Mask, const float* X, const float* W, float* dX, float* dWp,
                              float* dbp, float* dRes, int G, int R, long gs, long rs, int act, int splits, int flags, int mrs, cudaStream_t st) {
    const int chunks = (R + NW * 16 - 1) / (NW * 16);
    const int cps = (chunks + splits - 1) / splits;
    const size_t smem = (size_t)2 * NW * 16 * ROWB + (size_t)D * ROWB + (size_t)(2 * NW) * 4;
    auto kern = gproj3_bwd_kernel<NW, MINB, PREC>;
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return e;
    kern<<<dim3(G, splits), NW * 32, smem, st>>>(dY, Mask, X, W, dX, dWp, dbp, dRes, G, R, gs, rs, act, cps, flags, mrs);
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
static cudaError_t bwd_p(const float* dY, const uint2* Mask, const float* X, const float* W, float* dX, float* dWp, float* dbp,
                         float* dRes, int G, int R, long gs, long rs, int act, int splits, int flags, int mrs, cudaStream_t st) {
    GP2_DISPATCH(launch_bwd, dY, Mask, X, W, dX, dWp, dbp, dRes, G, R, gs, rs, act, splits, flags, mrs, st)
}

}  // namespace gp3

int gproj2_splits(int G, int R);     // gproj2.cu: the split policy is shared (same chunking, same pick_nw)

}  // namespace gptst

using namespace gptst;

// flags: bit 0 = dX accumulated in place, bit 1 = W / dW are [out][in] (as gptst_linear_bwd_acc); mask is required when act != 0
extern "C" int gptst_gproj3_bwd(const float* dY, const void* mask, const float* X, const float* W, float* dX, float* dW_part,
                                float* dbias_part, float* dRes, int G, int R, long group_stride, long row_stride, int D, int act,
                                int prec, int splits, int flags, void* stream) {
    if (!dY || !X || !W || !dX || !dW_part || !dbias_part || G <= 0 || R <= 0 || splits <= 0) return -1;
    if (act && !mask) return -1;
    if (D != 64 || (prec != 1 && prec != 3) || group_stride % D != 0 || row_stride % D != 0 || (flags & ~3)) return -2;
    cudaStream_t st = (cudaStream_t)stream;
    if (prec == PREC_3XTF32)
        return (int)gp3::bwd_p<PREC_3XTF32>(dY, (const uint2*)mask, X, W, dX, dW_part, dbias_part, dRes, G, R, group_stride, row_stride,
                                            act, splits, flags, 0, st);
    return (int)gp3::bwd_p<PREC_TF32>(dY, (const uint2*)mask, X, W, dX, dW_part, dbias_part, dRes, G, R, group_stride, row_stride, act,
                                      splits, flags, 0, st);
}

// Side-stream half of the fused hyperTem backward (csrc/htem_fused.cu): dW_bt = ret_bt^T dy_bt and db_bt = sum_n dy_bt for every
// (b, t), dy = dOut . LeakyReLU'(mask).  mask: the fused forward's sign words, `mask_rows` per (b, t).  Partials: (splits, B*T, ..)
// with splits = gptst_gproj_splits(B*T, N, 64).   Reference: GPTST.py:160-162, SURVEY.md appendix A (G_bt, sigma_bt).
extern "C" int gptst_hypertem_dw(const float* dout, const void* mask, const float* ret, float* dW_part, float* dbias_part, int B, int T,
                                 int N, int D, int mask_rows, int splits, void* stream) {
    if (!dout || !mask || !ret || !dW_part || !dbias_part || B <= 0 || T <= 0 || N <= 0 || splits <= 0) return -1;
    if (D != 64 || mask_rows < N) return -2;
    return (int)gp3::bwd_p<PREC_3XTF32>(dout, (const uint2*)mask, ret, ret, dW_part, dW_part, dbias_part, nullptr, B * T, N, (long)N * D,
                                        (long)D, 1, splits, 4 | 8, mask_rows, (cudaStream_t)stream);
}
