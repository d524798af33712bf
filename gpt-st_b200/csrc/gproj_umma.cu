// Grouped D x D projection (see gproj.cu for the contract) on the 5th-generation tensor cores, D = 64:
//   Y[g][r][:] = act( X[g][r][:] . W[g] + bias[g] (+ Res[g][r][:]) )
//
// One CTA = one group g and a range of 128-row tiles.  Per tile:
//   1. all 8 warps load the X rows (coalesced float4), split every value ONCE into tf32 hi / lo and store both as
//      core-matrix images in shared memory (conflict-free 16-byte stores);
//   2. one thread issues tcgen05.mma.kind::tf32 (M=128, N=64, K=8 x 8 steps; 3 operand pairs for the 3xTF32 split)
//      with the accumulator in tensor memory, and commits to an mbarrier;
//   3. the warps read the accumulator back with tcgen05.ld 32x32b (one output row per thread), add bias and residual,
//      apply LeakyReLU and store the row.
// W[g] is transposed to the K-major image B[n=o][k=i] once per CTA.  Two CTAs per SM overlap each other's phases.
#include "common.cuh"
#include "umma.cuh"

namespace gptst {

constexpr int UD = 64;        // feature width handled by this kernel
constexpr int UBM = 128;      // rows per tile (UMMA M)

template <int PREC>
__global__ void __launch_bounds__(256, 2) gproj_fwd_umma_kernel(const float* __restrict__ X, const float* __restrict__ W,
                                                                const float* __restrict__ bias, const float* __restrict__ Res,
                                                                float* __restrict__ Y, int G, int R, long group_stride,
                                                                long row_stride, int act) {
    extern __shared__ __align__(128) float smem[];
    __shared__ uint32_t tmem_base_s;
    __shared__ __align__(8) uint64_t mbar;
    float* Bh = smem;                                         // [64 x 64] image of W^T (hi)
    float* Bl = Bh + UD * UD;                                 // (lo)
    float* Ah = Bl + (PREC == PREC_3XTF32 ? UD * UD : 0);     // [128 x 64] image of the X tile (hi)
    float* Al = Ah + UBM * UD;                                // (lo)
    float* bs = Ah + ((PREC == PREC_3XTF32) ? 2 * UBM * UD : UBM * (UD + 4));   // [64], after the A images / staging tile

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int g = blockIdx.x;
    if (warp == 0) umma::tmem_alloc(&tmem_base_s, 64);
    if (tid == 0) {
        umma::mbar_init(&mbar, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    // ---- W[g] ([i][o]) -> image of B[n=o][k=i]; each thread transposes one 4x4 block in registers
    {
        const float* Wg = W + (size_t)g * UD * UD;
        const int bi = tid >> 4, bo = tid & 15;
        float w[4][4];
#pragma unroll
        for (int r = 0; r < 4; ++r) {
            const float4 t = *reinterpret_cast<const float4*>(Wg + (size_t)(4 * bi + r) * UD + 4 * bo);
            w[r][0] = t.x; w[r][1] = t.y; w[r][2] = t.z; w[r][3] = t.w;
        }
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            const int o = 4 * bo + j;
            uint32_t hi[4], lo[4];
#pragma unroll
            for (int r = 0; r < 4; ++r) split_tf32<PREC>(w[r][j], hi[r], lo[r]);
            const int off = umma::img_off(o, 4 * bi, UD);
            *reinterpret_cast<uint4*>(Bh + off) = make_uint4(hi[0], hi[1], hi[2], hi[3]);
            if (PREC == PREC_3XTF32) *reinterpret_cast<uint4*>(Bl + off) = make_uint4(lo[0], lo[1], lo[2], lo[3]);
        }
        if (tid < UD) bs[tid] = bias ? bias[(size_t)g * UD + tid] : 0.f;
    }
    umma::fence_before_sync();
    __syncthreads();
    umma::fence_after_sync();
    const uint32_t tbase = tmem_base_s;
    const uint32_t idesc = umma::idesc_tf32(UBM, UD);
    const uint32_t sbo = (UD / 4) * 128;

    const float* Xg = X + (size_t)g * group_stride;
    const float* Rg = Res ? Res + (size_t)g * group_stride : nullptr;
    float* Yg = Y + (size_t)g * group_stride;
    const int ntiles = (R + UBM - 1) / UBM;
    uint32_t phase = 0;
    constexpr int LDS_ = UD + 4;                 // staging row pitch (floats): conflict-free 16-byte row writes
    float* stage = Ah;                           // the A images are dead once the MMAs of a tile have completed
    // thread <-> data mappings
    //   image mapping (staging X):   unit = it*8 + warp -> rows (unit>>2)*8 + (lane&7), chunk (unit&3)*4 + (lane>>3)
    //   coalesced mapping (res / Y): id = it*256 + tid  -> row id>>4, chunk id&15
    float4 xr[8];
    auto load_x = [&](int tile_) {
        const int r0_ = tile_ * UBM;
#pragma unroll
        for (int it = 0; it < 8; ++it) {
            const int unit = it * 8 + warp;
            const int r = (unit >> 2) * 8 + (lane & 7);
            const int c = (unit & 3) * 4 + (lane >> 3);
            xr[it] = make_float4(0.f, 0.f, 0.f, 0.f);
            if (tile_ < ntiles && r0_ + r < R) xr[it] = *reinterpret_cast<const float4*>(Xg + (size_t)(r0_ + r) * row_stride + 4 * c);
        }
    };
    load_x(blockIdx.y);
    for (int tile = blockIdx.y; tile < ntiles; tile += gridDim.y) {
        const int r0 = tile * UBM;
        // ---- 1. split once, store both images (conflict-free 16-byte stores)
#pragma unroll
        for (int it = 0; it < 8; ++it) {
            const int unit = it * 8 + warp;
            const int r = (unit >> 2) * 8 + (lane & 7);
            const int c = (unit & 3) * 4 + (lane >> 3);
            uint32_t hi[4], lo[4];
            split_tf32<PREC>(xr[it].x, hi[0], lo[0]); split_tf32<PREC>(xr[it].y, hi[1], lo[1]);
            split_tf32<PREC>(xr[it].z, hi[2], lo[2]); split_tf32<PREC>(xr[it].w, hi[3], lo[3]);
            const int off = umma::img_off(r, 4 * c, UD);
            *reinterpret_cast<uint4*>(Ah + off) = make_uint4(hi[0], hi[1], hi[2], hi[3]);
            if (PREC == PREC_3XTF32) *reinterpret_cast<uint4*>(Al + off) = make_uint4(lo[0], lo[1], lo[2], lo[3]);
        }
        umma::fence_smem_to_async();
        umma::fence_before_sync();
        __syncthreads();
        // ---- 2. one thread issues the MMAs (accumulator in tensor memory)
        if (tid == 0) {
            umma::fence_after_sync();
            const uint32_t ah = umma::smem_u32(Ah), al = umma::smem_u32(Al), bh = umma::smem_u32(Bh), bl = umma::smem_u32(Bl);
            uint32_t acc = 0;
            if (PREC == PREC_3XTF32) {
#pragma unroll
                for (int k = 0; k < UD / 8; ++k) {
                    umma::mma_tf32(tbase, umma::kmajor_desc(al + k * 256, 128, sbo), umma::kmajor_desc(bh + k * 256, 128, sbo), idesc, acc);
                    acc = 1;
                }
#pragma unroll
                for (int k = 0; k < UD / 8; ++k)
                    umma::mma_tf32(tbase, umma::kmajor_desc(ah + k * 256, 128, sbo), umma::kmajor_desc(bl + k * 256, 128, sbo), idesc, 1);
            }
#pragma unroll
            for (int k = 0; k < UD / 8; ++k) {
                umma::mma_tf32(tbase, umma::kmajor_desc(ah + k * 256, 128, sbo), umma::kmajor_desc(bh + k * 256, 128, sbo), idesc, acc);
                acc = 1;
            }
            umma::commit(&mbar);
        }
        // ---- 3. while the tensor core works: prefetch this tile's residual and the next tile's X (coalesced)
        float4 rr[8];
#pragma unroll
        for (int it = 0; it < 8; ++it) {
            const int id = it * 256 + tid, row = r0 + (id >> 4);
            rr[it] = make_float4(0.f, 0.f, 0.f, 0.f);
            if (Rg && row < R) rr[it] = *reinterpret_cast<const float4*>(Rg + (size_t)row * row_stride + 4 * (id & 15));
        }
        load_x(tile + gridDim.y);
        // ---- 4. accumulator -> registers (one row per thread) -> staging tile
        umma::mbar_wait(&mbar, phase);
        phase ^= 1;
        umma::fence_after_sync();
        {
            const int q = warp & 3, ch = (warp >> 2) * 32;
            float v[32];
            umma::tmem_ld32(tbase + ((uint32_t)(q * 32) << 16) + ch, v);
            float* dst = stage + (q * 32 + lane) * LDS_ + ch;
#pragma unroll
            for (int j = 0; j < 8; ++j)
                *reinterpret_cast<float4*>(dst + 4 * j) = make_float4(v[4 * j] + bs[ch + 4 * j], v[4 * j + 1] + bs[ch + 4 * j + 1],
                                                                      v[4 * j + 2] + bs[ch + 4 * j + 2], v[4 * j + 3] + bs[ch + 4 * j + 3]);
        }
        umma::fence_before_sync();
        __syncthreads();
        // ---- 5. coalesced epilogue
#pragma unroll
        for (int it = 0; it < 8; ++it) {
            const int id = it * 256 + tid, rl = id >> 4, row = r0 + rl;
            if (row < R) {
                float4 o = *reinterpret_cast<const float4*>(stage + rl * LDS_ + 4 * (id & 15));
                o.x += rr[it].x; o.y += rr[it].y; o.z += rr[it].z; o.w += rr[it].w;
                if (act) { o.x = lrelu(o.x); o.y = lrelu(o.y); o.z = lrelu(o.z); o.w = lrelu(o.w); }
                *reinterpret_cast<float4*>(Yg + (size_t)row * row_stride + 4 * (id & 15)) = o;
            }
        }
        __syncthreads();   // staging tile (aliases the A images) free before the next tile is staged
    }
    if (warp == 0) umma::tmem_dealloc(tbase, 64);
}

static size_t gproj_umma_smem(int prec) {
    const int planes = (prec == PREC_3XTF32) ? 2 : 1;
    const size_t a_floats = (planes == 2) ? 2 * UBM * UD : UBM * (UD + 4);   // A images, reused as the (128 x 68) staging tile
    return ((size_t)planes * UD * UD + a_floats + UD) * 4 + 128;
}

int gproj_fwd_umma_splits(int G, int R) {
    int ntiles = (R + UBM - 1) / UBM;
    int want = (592 + G - 1) / G;
    int s = want < ntiles ? want : ntiles;
    return s < 1 ? 1 : s;
}

cudaError_t gproj_fwd_umma(const float* X, const float* W, const float* bias, const float* Res, float* Y, int G, int R,
                           long gs, long rs, int act, int prec, cudaStream_t st) {
    size_t smem = gproj_umma_smem(prec);
    dim3 grid(G, gproj_fwd_umma_splits(G, R));
    cudaError_t e;
    if (prec == PREC_3XTF32) {
        e = cudaFuncSetAttribute(gproj_fwd_umma_kernel<PREC_3XTF32>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) return e;
        gproj_fwd_umma_kernel<PREC_3XTF32><<<grid, 256, smem, st>>>(X, W, bias, Res, Y, G, R, gs, rs, act);
    } else {
        e = cudaFuncSetAttribute(gproj_fwd_umma_kernel<PREC_TF32>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) return e;
        gproj_fwd_umma_kernel<PREC_TF32><<<grid, 256, smem, st>>>(X, W, bias, Res, Y, G, R, gs, rs, act);
    }
    return cudaGetLastError();
}

}  // namespace gptst

// =====================================================================================================
// Backward of the grouped projection, D = 64 (contract in gproj.cu):
//   dy = dY * act'(Y) ; dRes = dy ; dX = dy . W^T ; dW[g] = X^T dy ; dbias[g] = sum_r dy
// dX runs on tcgen05 (A = dy tile image, B = W[g] image, both K-major, accumulator in TMEM); dW (contraction over
// the rows, which would need MN-major operands) runs on mma.sync reading the SAME pre-split hi/lo images, so no
// fragment is ever re-split.  16 warps, one CTA per SM, next tile prefetched into registers while the tensor
// cores and the dW warps work.
// =====================================================================================================
namespace gptst {

template <int PREC>
__global__ void __launch_bounds__(512, 1) gproj_bwd_umma_kernel(const float* __restrict__ dY, const float* __restrict__ Y,
                                                                const float* __restrict__ X, const float* __restrict__ W,
                                                                float* __restrict__ dX, float* __restrict__ dWp,
                                                                float* __restrict__ dbp, float* __restrict__ dRes, int G, int R,
                                                                long group_stride, long row_stride, int act) {
    extern __shared__ __align__(128) float smem[];
    __shared__ uint32_t tmem_base_s;
    __shared__ __align__(8) uint64_t mbar;
    constexpr int PL = (PREC == PREC_3XTF32) ? 2 : 1;
    constexpr int LDS_ = UD + 4;            // dX staging pitch
    constexpr int LDR = UD + 8;             // raw tile pitch: conflict-free for both mma.sync fragment patterns
    float* Wh = smem;                       // image of W[g] ([i][o]) = B[n=i][k=o]   (tcgen05)
    float* Wl = Wh + UD * UD;
    float* Dh = Wh + PL * UD * UD;          // image of the dy tile (rows x o)         (tcgen05 A operand)
    float* Dl = Dh + UBM * UD;
    float* Xr = Dh + PL * UBM * UD;         // raw fp32 X tile  [128][72]              (mma.sync, read transposed)
    float* Dr = Xr + UBM * LDR;             // raw fp32 dy tile [128][72]              (mma.sync B operand)
    float* stage = Dr + UBM * LDR;          // [128][68] dX staging; at the end: dW cross-warp reduction
    float* red = stage + UBM * LDS_;        // [16 warps][16] column-sum exchange

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int g = blockIdx.x;
    if (warp == 0) umma::tmem_alloc(&tmem_base_s, 64);
    if (tid == 0) {
        umma::mbar_init(&mbar, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    {
        const float* Wg = W + (size_t)g * UD * UD;
#pragma unroll
        for (int it = 0; it < 2; ++it) {
            const int id = it * 512 + tid, r = id >> 4, c = id & 15;
            const float4 v = *reinterpret_cast<const float4*>(Wg + (size_t)r * UD + 4 * c);
            uint32_t hi[4], lo[4];
            split_tf32<PREC>(v.x, hi[0], lo[0]); split_tf32<PREC>(v.y, hi[1], lo[1]);
            split_tf32<PREC>(v.z, hi[2], lo[2]); split_tf32<PREC>(v.w, hi[3], lo[3]);
            const int off = umma::img_off(r, 4 * c, UD);
            *reinterpret_cast<uint4*>(Wh + off) = make_uint4(hi[0], hi[1], hi[2], hi[3]);
            if (PREC == PREC_3XTF32) *reinterpret_cast<uint4*>(Wl + off) = make_uint4(lo[0], lo[1], lo[2], lo[3]);
        }
    }
    umma::fence_before_sync();
    __syncthreads();
    umma::fence_after_sync();
    const uint32_t tbase = tmem_base_s;
    const uint32_t idesc = umma::idesc_tf32(UBM, UD);
    const uint32_t sbo = (UD / 4) * 128;

    const float* Xg = X + (size_t)g * group_stride;
    const float* Yg = Y ? Y + (size_t)g * group_stride : nullptr;
    const float* dYg = dY + (size_t)g * group_stride;
    float* dXg = dX + (size_t)g * group_stride;
    float* dRg = dRes ? dRes + (size_t)g * group_stride : nullptr;
    const int ntiles = (R + UBM - 1) / UBM;

    // image mapping with 16 warps: unit = it*16 + warp -> rows (unit>>2)*8 + (lane&7), chunk (unit&3)*4 + (lane>>3)
    float4 xr[4], dr[4], yr[4];
    auto load_tile = [&](int tile_) {
        const int r0_ = tile_ * UBM;
#pragma unroll
        for (int it = 0; it < 4; ++it) {
            const int unit = it * 16 + warp;
            const int r = (unit >> 2) * 8 + (lane & 7), c = (unit & 3) * 4 + (lane >> 3);
            xr[it] = dr[it] = make_float4(0.f, 0.f, 0.f, 0.f);
            yr[it] = make_float4(1.f, 1.f, 1.f, 1.f);
            if (tile_ < ntiles && r0_ + r < R) {
                const size_t off = (size_t)(r0_ + r) * row_stride + 4 * c;
                xr[it] = *reinterpret_cast<const float4*>(Xg + off);
                dr[it] = *reinterpret_cast<const float4*>(dYg + off);
                if (act) yr[it] = *reinterpret_cast<const float4*>(Yg + off);
            }
        }
    };
    // dW over 16 warps: m-tile (16 i's) = warp&3, K slice (32 rows of the tile) = warp>>2, all 8 n-tiles
    const int qm = warp & 3, ks = warp >> 2;
    float gacc[UD / 8][4];
#pragma unroll
    for (int j = 0; j < UD / 8; ++j) gacc[j][0] = gacc[j][1] = gacc[j][2] = gacc[j][3] = 0.f;
    float sig[4] = {0.f, 0.f, 0.f, 0.f};     // running column sums of dy for chunk (warp&3)*4 + (lane>>3)

    uint32_t phase = 0;
    load_tile(blockIdx.y);
    for (int tile = blockIdx.y; tile < ntiles; tile += gridDim.y) {
        const int r0 = tile * UBM;
        // ---- 1. dy = dY * act'(Y); dRes; dy -> tcgen05 images (split once) + raw tile; X -> raw tile
#pragma unroll
        for (int it = 0; it < 4; ++it) {
            const int unit = it * 16 + warp;
            const int r = (unit >> 2) * 8 + (lane & 7), c = (unit & 3) * 4 + (lane >> 3);
            float4 d = dr[it];
            if (act) {
                d.x = lrelu_grad(yr[it].x, d.x); d.y = lrelu_grad(yr[it].y, d.y);
                d.z = lrelu_grad(yr[it].z, d.z); d.w = lrelu_grad(yr[it].w, d.w);
            }
            if (dRg && r0 + r < R) *reinterpret_cast<float4*>(dRg + (size_t)(r0 + r) * row_stride + 4 * c) = d;
            sig[0] += d.x; sig[1] += d.y; sig[2] += d.z; sig[3] += d.w;
            uint32_t hi[4], lo[4];
            const int off = umma::img_off(r, 4 * c, UD);
            split_tf32<PREC>(d.x, hi[0], lo[0]); split_tf32<PREC>(d.y, hi[1], lo[1]);
            split_tf32<PREC>(d.z, hi[2], lo[2]); split_tf32<PREC>(d.w, hi[3], lo[3]);
            *reinterpret_cast<uint4*>(Dh + off) = make_uint4(hi[0], hi[1], hi[2], hi[3]);
            if (PREC == PREC_3XTF32) *reinterpret_cast<uint4*>(Dl + off) = make_uint4(lo[0], lo[1], lo[2], lo[3]);
            *reinterpret_cast<float4*>(Dr + r * LDR + 4 * c) = d;
            *reinterpret_cast<float4*>(Xr + r * LDR + 4 * c) = xr[it];
        }
        umma::fence_smem_to_async();
        umma::fence_before_sync();
        __syncthreads();
        // ---- 2. dX on the tensor cores (async)
        if (tid == 0) {
            umma::fence_after_sync();
            const uint32_t ah = umma::smem_u32(Dh), al = umma::smem_u32(Dl), bh = umma::smem_u32(Wh), bl = umma::smem_u32(Wl);
            uint32_t acc = 0;
            if (PREC == PREC_3XTF32) {
#pragma unroll
                for (int k = 0; k < UD / 8; ++k) {
                    umma::mma_tf32(tbase, umma::kmajor_desc(al + k * 256, 128, sbo), umma::kmajor_desc(bh + k * 256, 128, sbo), idesc, acc);
                    acc = 1;
                }
#pragma unroll
                for (int k = 0; k < UD / 8; ++k)
                    umma::mma_tf32(tbase, umma::kmajor_desc(ah + k * 256, 128, sbo), umma::kmajor_desc(bl + k * 256, 128, sbo), idesc, 1);
            }
#pragma unroll
            for (int k = 0; k < UD / 8; ++k) {
                umma::mma_tf32(tbase, umma::kmajor_desc(ah + k * 256, 128, sbo), umma::kmajor_desc(bh + k * 256, 128, sbo), idesc, acc);
                acc = 1;
            }
            umma::commit(&mbar);
        }
        // ---- 3. prefetch the next tile
        load_tile(tile + gridDim.y);
        // ---- 4. dW += X^T dy on mma.sync: A(m=i,k=row) = Xr[row][i] (transposed read), B(k=row,n=o) = Dr[row][o]
        warp_gemm<32, UD / 8, PREC, true, true>(gacc, Xr + (ks * 32) * LDR + qm * 16, LDR, Dr + (ks * 32) * LDR, LDR, lane);
        // ---- 5. dX accumulator -> staging -> coalesced store
        umma::mbar_wait(&mbar, phase);
        phase ^= 1;
        umma::fence_after_sync();
        {
            const int q = warp & 3, ch = (warp >> 2) * 16;
            uint32_t v[16];
            const uint32_t taddr = tbase + ((uint32_t)(q * 32) << 16) + ch;
            asm volatile("tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
                         : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]),
                           "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]) : "r"(taddr));
            asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
            float* dst = stage + (q * 32 + lane) * LDS_ + ch;
#pragma unroll
            for (int j = 0; j < 4; ++j) *reinterpret_cast<uint4*>(dst + 4 * j) = make_uint4(v[4 * j], v[4 * j + 1], v[4 * j + 2], v[4 * j + 3]);
        }
        umma::fence_before_sync();
        __syncthreads();
#pragma unroll
        for (int it = 0; it < 4; ++it) {
            const int id = it * 512 + tid, rl = id >> 4, row = r0 + rl;
            if (row < R)
                *reinterpret_cast<float4*>(dXg + (size_t)row * row_stride + 4 * (id & 15)) =
                    *reinterpret_cast<const float4*>(stage + rl * LDS_ + 4 * (id & 15));
        }
        __syncthreads();   // tiles and staging free for the next tile
    }
    // ---- dW: deterministic reduction over the 4 K slices through shared memory, then one store per CTA
    {
        const int gq = lane >> 2, tq = lane & 3;
        float* buf = stage;                     // [4 m-tiles][16][64] = 16 KB
        for (int round = 1; round < 4; ++round) {
            if (ks == round) {
#pragma unroll
                for (int j = 0; j < UD / 8; ++j) {
                    const int c = j * 8 + 2 * tq;
                    *reinterpret_cast<float2*>(buf + (qm * 16 + gq) * UD + c) = make_float2(gacc[j][0], gacc[j][1]);
                    *reinterpret_cast<float2*>(buf + (qm * 16 + gq + 8) * UD + c) = make_float2(gacc[j][2], gacc[j][3]);
                }
            }
            __syncthreads();
            if (ks == 0) {
#pragma unroll
                for (int j = 0; j < UD / 8; ++j) {
                    const int c = j * 8 + 2 * tq;
                    const float2 a = *reinterpret_cast<const float2*>(buf + (qm * 16 + gq) * UD + c);
                    const float2 b2 = *reinterpret_cast<const float2*>(buf + (qm * 16 + gq + 8) * UD + c);
                    gacc[j][0] += a.x; gacc[j][1] += a.y; gacc[j][2] += b2.x; gacc[j][3] += b2.y;
                }
            }
            __syncthreads();
        }
        if (ks == 0) {
            float* dWo = dWp + ((size_t)blockIdx.y * G + g) * UD * UD;
#pragma unroll
            for (int j = 0; j < UD / 8; ++j) {
                const int c = j * 8 + 2 * tq, r = qm * 16 + gq;
                *reinterpret_cast<float2*>(dWo + (size_t)r * UD + c) = make_float2(gacc[j][0], gacc[j][1]);
                *reinterpret_cast<float2*>(dWo + (size_t)(r + 8) * UD + c) = make_float2(gacc[j][2], gacc[j][3]);
            }
        }
        // column sums: reduce over the 8 row-lanes of each chunk, then over the 4 warps that share a chunk group
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            sig[j] += __shfl_xor_sync(0xffffffffu, sig[j], 1);
            sig[j] += __shfl_xor_sync(0xffffffffu, sig[j], 2);
            sig[j] += __shfl_xor_sync(0xffffffffu, sig[j], 4);
        }
        if ((lane & 7) == 0) {
#pragma unroll
            for (int j = 0; j < 4; ++j) red[warp * 16 + (lane >> 3) * 4 + j] = sig[j];
        }
        __syncthreads();
        if (tid < UD) {
            const int chunk = tid >> 2, cg = chunk >> 2, slot = (chunk & 3) * 4 + (tid & 3);
            float s = 0.f;
#pragma unroll
            for (int w = 0; w < 4; ++w) s += red[(w * 4 + cg) * 16 + slot];
            dbp[((size_t)blockIdx.y * G + g) * UD + tid] = s;
        }
    }
    umma::fence_before_sync();
    __syncthreads();
    if (warp == 0) umma::tmem_dealloc(tbase, 64);
}

static size_t gproj_bwd_umma_smem(int prec) {
    const int pl = (prec == PREC_3XTF32) ? 2 : 1;
    return ((size_t)pl * (UD * UD + UBM * UD) + 2 * UBM * (UD + 8) + UBM * (UD + 4) + 16 * 16) * 4 + 128;
}

int gproj_bwd_umma_splits(int G, int R) {
    int ntiles = (R + UBM - 1) / UBM;
    int want = (296 + G - 1) / G;
    int s = want < ntiles ? want : ntiles;
    return s < 1 ? 1 : s;
}

cudaError_t gproj_bwd_umma(const float* dY, const float* Y, const float* X, const float* W, float* dX, float* dWp, float* dbp,
                           float* dRes, int G, int R, long gs, long rs, int act, int prec, int splits, cudaStream_t st) {
    size_t smem = gproj_bwd_umma_smem(prec);
    dim3 grid(G, splits);
    cudaError_t e;
    if (prec == PREC_3XTF32) {
        e = cudaFuncSetAttribute(gproj_bwd_umma_kernel<PREC_3XTF32>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) return e;
        gproj_bwd_umma_kernel<PREC_3XTF32><<<grid, 512, smem, st>>>(dY, Y, X, W, dX, dWp, dbp, dRes, G, R, gs, rs, act);
    } else {
        e = cudaFuncSetAttribute(gproj_bwd_umma_kernel<PREC_TF32>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) return e;
        gproj_bwd_umma_kernel<PREC_TF32><<<grid, 512, smem, st>>>(dY, Y, X, W, dX, dWp, dbp, dRes, G, R, gs, rs, act);
    }
    return cudaGetLastError();
}

}  // namespace gptst
