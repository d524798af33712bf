// Shared device helpers for the GPT-ST sm_100a kernels.
#pragma once
#include <cuda_runtime.h>
#include <cstdlib>
#include <stdint.h>

namespace gptst {

constexpr float kSlope = 0.01f;  // nn.LeakyReLU() default (reference GPTST.py:18,96,152)
constexpr int kMaxH = 16;        // hyperedge counts are padded to 16 in registers
constexpr int kMaxT = 12;        // the reference hard-wires 12 time steps (GPTST.py:97,208)

// Precision of the D x D tensor-core contractions:
//   1 = single TF32 pass (operands rounded to 10-bit mantissa, fp32 accumulate)
//   3 = 3xTF32 split (a_hi*b_hi + a_hi*b_lo + a_lo*b_hi): fp32-faithful, the default
enum { PREC_TF32 = 1, PREC_3XTF32 = 3 };

__device__ __forceinline__ float lrelu(float y) { return y > 0.f ? y : kSlope * y; }
// derivative selected by the sign of the *output* (sign(out) == sign(pre-activation));
// matches torch's leaky_relu_backward: slope 1 iff x > 0.
__device__ __forceinline__ float lrelu_grad(float out, float g) { return out > 0.f ? g : kSlope * g; }

// squash scale: P = Z * f(q), q = |Z|^2.  GPTST.py:36-39:  (q/(1+q)) * Z / (sqrt(q)+1e-8)
__device__ __forceinline__ float squash_f(float q) { return (q / (1.f + q)) / (sqrtf(q) + 1e-8f); }
// f'(q)
__device__ __forceinline__ float squash_df(float q) {
    float r = sqrtf(q), re = r + 1e-8f, den = (1.f + q) * re;
    // d/dq [ q / ((1+q)(r+eps)) ],  d(den)/dq = (r+eps) + (1+q)/(2r)
    float dden = re + (1.f + q) / (2.f * fmaxf(r, 1e-30f));
    return (den - q * dden) / (den * den);
}

// ---- programmatic dependent launch (PDL), FORWARD main chain only ---------------------------------------------------------
// A kernel launched with cudaLaunchAttributeProgrammaticStreamSerialization may be scheduled while its stream predecessor still
// runs; pdl_enter() as its first statement blocks until the predecessor has completed and its writes are visible (`wait`) and
// THEN lets its own successor be scheduled (`launch_dependents`): a successor that starts early therefore knows that everything
// older than its direct predecessor is complete, and may read such data before its own wait (cap_recon_proj stages the
// incidence tile of the routing kernel while the hop kernel runs).  Nothing else of a kernel runs early: launch latency and CTA
// scheduling overlap the predecessor.  No-ops without the attribute.  Measured (round 2): on EVERY
// kernel of the step this costs 120 us -- in the backward the early CTAs take the tail-wave slots the low-priority side-stream
// kernels live on (profiles/ab_pdl_slim_r02.md); the forward chain has no such tenants.  GPTST_B200_PDL=0 switches it off.
__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }
__device__ __forceinline__ void pdl_trigger() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }
__device__ __forceinline__ void pdl_enter() {
    pdl_wait();
    pdl_trigger();
}

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}

// TF32 operand preparation.  NOTE: on sm_100a `cvt.rna.tf32.f32` is emulated by a ~15-instruction sequence
// (observed in SASS: FSETP/SEL/LOP3/VIADD 0x1000), which made every kernel issue-bound.  The tensor cores ignore
// the low 13 mantissa bits of a tf32 operand, so:
//   1xTF32 : feed the fp32 bits as they are (truncation -- what cuBLAS' TF32 mode does as well);
//   3xTF32 : hi = x with the low 13 bits cleared (1 LOP3), lo = x - hi (exact in fp32, 1 FADD).  hi + lo == x exactly,
//            and lo itself is consumed with ~11 significant bits, so the dropped part is < 2^-21 |x|.
template <int PREC>
__device__ __forceinline__ void split_tf32(float x, uint32_t& hi, uint32_t& lo) {
    const uint32_t b = __float_as_uint(x);
    if (PREC == PREC_3XTF32) {
        hi = b & 0xffffe000u;
        lo = __float_as_uint(x - __uint_as_float(hi));
    } else {
        hi = b;
        lo = 0u;
    }
}

// D(16x8) += A(16x8, row) * B(8x8, col), tf32 operands, fp32 accumulate.
//   A: a0=(g,t) a1=(g+8,t) a2=(g,t+4) a3=(g+8,t+4)      g = lane>>2, t = lane&3
//   B: b0=(k=t,n=g) b1=(k=t+4,n=g)
//   C: c0=(g,2t) c1=(g,2t+1) c2=(g+8,2t) c3=(g+8,2t+1)
__device__ __forceinline__ void mma_tf32(float (&c)[4], const uint32_t (&a)[4], const uint32_t (&b)[2]) {
    asm volatile(
        "mma.sync.aligned.m16n8k8.row.col.f32.tf32.tf32.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
        : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3])
        : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b[0]), "r"(b[1]));
}

template <int PREC>
__device__ __forceinline__ void mma_split(float (&c)[4], const uint32_t (&ah)[4], const uint32_t (&al)[4],
                                          const uint32_t (&bh)[2], const uint32_t (&bl)[2]) {
    if (PREC == PREC_3XTF32) {
        mma_tf32(c, al, bh);
        mma_tf32(c, ah, bl);
    }
    mma_tf32(c, ah, bh);
}

// ---------------------------------------------------------------------------------------------
// Warp-level GEMM building block: acc[NT][4] += A(16 x K) * B(K x 8*NT).
// A is fp32 in shared memory, row-major: A(m,k) = As[m*lda + k]      (lda % 32 == 4 is conflict-free)
// A transposed:                          A(m,k) = As[k*lda + m]      (lda % 32 == 8 is conflict-free)
// B "n-major":                           B(k,n) = Bs[n*ldb + k]      (ldb % 32 == 4 is conflict-free)
// B "k-major":                           B(k,n) = Bs[k*ldb + n]      (ldb % 32 == 8 is conflict-free)
// Operands are split to tf32 hi/lo on the fly.
// ---------------------------------------------------------------------------------------------
template <int K, int NT, int PREC, bool A_TRANS, bool B_KMAJOR>
__device__ __forceinline__ void warp_gemm(float (&acc)[NT][4], const float* __restrict__ As, int lda,
                                          const float* __restrict__ Bs, int ldb, int lane) {
    const int g = lane >> 2, t = lane & 3;
#pragma unroll
    for (int k0 = 0; k0 < K; k0 += 8) {
        float af[4];
        if (!A_TRANS) {
            af[0] = As[g * lda + k0 + t];
            af[1] = As[(g + 8) * lda + k0 + t];
            af[2] = As[g * lda + k0 + t + 4];
            af[3] = As[(g + 8) * lda + k0 + t + 4];
        } else {
            af[0] = As[(k0 + t) * lda + g];
            af[1] = As[(k0 + t) * lda + g + 8];
            af[2] = As[(k0 + t + 4) * lda + g];
            af[3] = As[(k0 + t + 4) * lda + g + 8];
        }
        uint32_t ah[4], al[4];
#pragma unroll
        for (int i = 0; i < 4; ++i) split_tf32<PREC>(af[i], ah[i], al[i]);
#pragma unroll
        for (int nt = 0; nt < NT; ++nt) {
            float bf0, bf1;
            if (!B_KMAJOR) {
                bf0 = Bs[(nt * 8 + g) * ldb + k0 + t];
                bf1 = Bs[(nt * 8 + g) * ldb + k0 + t + 4];
            } else {
                bf0 = Bs[(k0 + t) * ldb + nt * 8 + g];
                bf1 = Bs[(k0 + t + 4) * ldb + nt * 8 + g];
            }
            uint32_t bh[2], bl[2];
            split_tf32<PREC>(bf0, bh[0], bl[0]);
            split_tf32<PREC>(bf1, bh[1], bl[1]);
            mma_split<PREC>(acc[nt], ah, al, bh, bl);
        }
    }
}

// Same, but B has been pre-split into tf32 hi / lo planes in shared memory (uint32 bit patterns).
template <int K, int NT, int PREC, bool B_KMAJOR>
__device__ __forceinline__ void warp_gemm_presplit(float (&acc)[NT][4], const float* __restrict__ As, int lda,
                                                   const uint32_t* __restrict__ Bh, const uint32_t* __restrict__ Bl,
                                                   int ldb, int lane) {
    const int g = lane >> 2, t = lane & 3;
#pragma unroll
    for (int k0 = 0; k0 < K; k0 += 8) {
        uint32_t ah[4], al[4];
        split_tf32<PREC>(As[g * lda + k0 + t], ah[0], al[0]);
        split_tf32<PREC>(As[(g + 8) * lda + k0 + t], ah[1], al[1]);
        split_tf32<PREC>(As[g * lda + k0 + t + 4], ah[2], al[2]);
        split_tf32<PREC>(As[(g + 8) * lda + k0 + t + 4], ah[3], al[3]);
#pragma unroll
        for (int nt = 0; nt < NT; ++nt) {
            uint32_t bh[2], bl[2] = {0u, 0u};
            if (!B_KMAJOR) {
                bh[0] = Bh[(nt * 8 + g) * ldb + k0 + t];
                bh[1] = Bh[(nt * 8 + g) * ldb + k0 + t + 4];
                if (PREC == PREC_3XTF32) {
                    bl[0] = Bl[(nt * 8 + g) * ldb + k0 + t];
                    bl[1] = Bl[(nt * 8 + g) * ldb + k0 + t + 4];
                }
            } else {
                bh[0] = Bh[(k0 + t) * ldb + nt * 8 + g];
                bh[1] = Bh[(k0 + t + 4) * ldb + nt * 8 + g];
                if (PREC == PREC_3XTF32) {
                    bl[0] = Bl[(k0 + t) * ldb + nt * 8 + g];
                    bl[1] = Bl[(k0 + t + 4) * ldb + nt * 8 + g];
                }
            }
            mma_split<PREC>(acc[nt], ah, al, bh, bl);
        }
    }
}

__device__ __forceinline__ uint32_t cluster_ctarank() {
    uint32_t r;
    asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
    return r;
}
__device__ __forceinline__ void cluster_sync_all() {
    asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
}
// generic address of `p` (a shared-memory pointer of this CTA) as seen in CTA `rank` of the cluster
__device__ __forceinline__ const float* cluster_map(const float* p, uint32_t rank) {
    uint64_t out;
    asm volatile("mapa.u64 %0, %1, %2;" : "=l"(out) : "l"(reinterpret_cast<uint64_t>(p)), "r"(rank));
    return reinterpret_cast<const float*>(out);
}


inline bool pdl_enabled() {
    static int v = -1;
    if (v < 0) {
        const char* e = getenv("GPTST_B200_PDL");
        v = (e && e[0] == '0') ? 0 : 1;
    }
    return v != 0;
}
// kern<<<grid, block, smem, st>>>(args...) with the programmatic-stream-serialization attribute (see pdl_enter)
template <typename... KArgs, typename... Args>
inline cudaError_t launch_pdl(void (*kern)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t st, Args&&... args) {
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = grid;
    cfg.blockDim = block;
    cfg.dynamicSmemBytes = smem;
    cfg.stream = st;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = attr;
    cfg.numAttrs = pdl_enabled() ? 1 : 0;
    return cudaLaunchKernelEx(&cfg, kern, static_cast<KArgs>(args)...);
}

}  // namespace gptst
