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
