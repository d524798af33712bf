// Validation of the hand-built tcgen05 (UMMA) descriptors used by the kernels: no-swizzle core-matrix images,
// K-major and MN-major operands taken from the SAME shared-memory image, TMEM accumulator layout for M=128 / M=64.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O2 -o tools/umma_test tools/umma_test.cu && ./tools/umma_test
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <cuda_runtime.h>
#include <stdint.h>
#include <vector>

// element (r,c) of an R x C fp32 matrix in the blocked image: 8x4 core blocks of 128 B, column blocks contiguous
__host__ __device__ inline int img_off(int r, int c, int C) { return ((r >> 3) * (C >> 2) + (c >> 2)) * 32 + (r & 7) * 4 + (c & 3); }

// SW128 image: atoms of 8 rows x 32 cols (1024 B), 16-byte chunks XOR-swizzled with the row index; atom (rb, cb) at (rb*(C/32)+cb)*1024 B
__host__ __device__ inline int img128_off(int r, int c, int C) {
    return ((r >> 3) * (C >> 5) + (c >> 5)) * 256 + (r & 7) * 32 + ((((c & 31) >> 2) ^ (r & 7)) << 2) + (c & 3);
}
__device__ inline uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ inline uint64_t make_desc(uint32_t saddr, uint32_t lbo_bytes, uint32_t sbo_bytes) {
    uint64_t d = 0;
    d |= (uint64_t)((saddr >> 4) & 0x3FFF);
    d |= (uint64_t)((lbo_bytes >> 4) & 0x3FFF) << 16;
    d |= (uint64_t)((sbo_bytes >> 4) & 0x3FFF) << 32;
    d |= (uint64_t)1 << 46;   // descriptor version (Blackwell)
    return d;                 // layout_type = 0 (no swizzle), base_offset = 0
}
__device__ inline uint64_t make_desc128(uint32_t saddr, uint32_t lbo_bytes, uint32_t sbo_bytes) {
    return make_desc(saddr, lbo_bytes, sbo_bytes) | ((uint64_t)2 << 61);   // layout_type 2 = SWIZZLE_128B
}
__host__ __device__ inline uint32_t make_idesc(int M, int N, int a_mn_major, int b_mn_major) {
    return (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)a_mn_major << 15) | ((uint32_t)b_mn_major << 16) |
           ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
}

// mode 0: D[M=128][N] = A[128][K] * Bt[N][K]^T      A K-major (image of A rows x K), B K-major (image of Bt, N x K)
// mode 1: D[M=128][N] = A[128][K] * Bk[K][N]        B MN-major (image of Bk, K x N)
// mode 2: D[M][N]     = At[K][M]^T * Bk[K][N]       A MN-major (image of At, K x M), B MN-major; M = 64 or 128
__global__ void umma_kernel(const float* A, const float* B, float* Dout, int M, int N, int K, int mode, int variant) {
    extern __shared__ __align__(1024) float smem[];
    __shared__ uint32_t tmem_base;
    __shared__ __align__(8) uint64_t mbar;
    float* As = smem;
    const int a_rows = (mode % 10 == 2) ? K : M, a_cols = (mode % 10 == 2) ? M : K;
    const int b_rows = (mode % 10 == 0) ? N : K, b_cols = (mode % 10 == 0) ? K : N;
    float* Bs = As + a_rows * a_cols;
    const bool sw = mode >= 10;
    if (sw) mode -= 10;
    for (int i = threadIdx.x; i < a_rows * a_cols; i += blockDim.x)
        As[sw ? img128_off(i / a_cols, i % a_cols, a_cols) : img_off(i / a_cols, i % a_cols, a_cols)] = A[i];
    for (int i = threadIdx.x; i < b_rows * b_cols; i += blockDim.x)
        Bs[sw ? img128_off(i / b_cols, i % b_cols, b_cols) : img_off(i / b_cols, i % b_cols, b_cols)] = B[i];
    const int warp = threadIdx.x >> 5;
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_base)), "r"(128));
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
    }
    if (threadIdx.x == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(&mbar)), "r"(1));
    }
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");   // generic-proxy smem writes -> visible to the async (tensor core) proxy
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tbase = tmem_base;
    if (threadIdx.x == 0) {
        const uint32_t idesc = make_idesc(M, N, mode == 2, mode != 0);
        for (int k0 = 0; k0 < K; k0 += 8) {
            uint64_t da, db;
            if (sw) {
                if (mode != 2)  // A K-major SW128: 8-row groups at SBO; k advance inside the 128-B atom row, next K atom +1024
                    da = make_desc128(smem_u32(As) + (k0 / 32) * 1024 + (k0 % 32) * 4, 16, (K / 32) * 1024);
                else            // A MN-major SW128 from image of At (K x M): LBO = next 32 m, SBO = next 8 k
                    da = make_desc128(smem_u32(As) + (k0 / 8) * (M / 32) * 1024, 1024, (M / 32) * 1024);
                if (mode == 0)
                    db = make_desc128(smem_u32(Bs) + (k0 / 32) * 1024 + (k0 % 32) * 4, 16, (K / 32) * 1024);
                else
                    db = make_desc128(smem_u32(Bs) + (k0 / 8) * (N / 32) * 1024, 1024, (N / 32) * 1024);
            } else
            if (mode != 2) {   // A K-major: rows = M, cols = K.  SBO = next 8 rows, LBO = next 4 k
                da = make_desc(smem_u32(As) + (k0 / 4) * 128, 128, (K / 4) * 128);
            } else {           // A MN-major from image of At (K x M): SBO = next 4 m (col block), LBO = next 8 k (row block)
                da = (variant & 1) ? make_desc(smem_u32(As) + (k0 / 8) * (M / 4) * 128, 128, (M / 4) * 128)
                                   : make_desc(smem_u32(As) + (k0 / 8) * (M / 4) * 128, (M / 4) * 128, 128);
            }
            if (mode == 0) {   // B K-major from image of Bt (N x K)
                db = make_desc(smem_u32(Bs) + (k0 / 4) * 128, 128, (K / 4) * 128);
            } else {           // B MN-major from image of Bk (K x N)
                db = (variant & 1) ? make_desc(smem_u32(Bs) + (k0 / 8) * (N / 4) * 128, 128, (N / 4) * 128)
                                   : make_desc(smem_u32(Bs) + (k0 / 8) * (N / 4) * 128, (N / 4) * 128, 128);
            }
            const uint32_t acc = k0 > 0;
            asm volatile(
                "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
                "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}"
                ::"r"(tbase), "l"(da), "l"(db), "r"(idesc), "r"(acc));
        }
        asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(&mbar)));
    }
    // wait for the MMAs
    {
        uint32_t done = 0;
        while (!done) {
            asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                         : "=r"(done) : "r"(smem_u32(&mbar)), "r"(0));
        }
    }
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    // dump all 128 lanes x N columns: warp w reads lanes 32w..32w+31
    for (int c0 = 0; c0 < N; c0 += 8) {
        uint32_t r[8];
        const uint32_t taddr = tbase + ((uint32_t)(warp * 32) << 16) + c0;
        asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
                     : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]) : "r"(taddr));
        asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
        for (int j = 0; j < 8; ++j) Dout[(size_t)threadIdx.x * N + c0 + j] = __uint_as_float(r[j]);
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tbase), "r"(128));
}

static float val(int i, int j, int s) { uint32_t h = (uint32_t)(i * 7919 + j * 104729 + s * 31) * 2654435761u; return (float)((int)((h >> 20) % 17) - 8) * 0.25f; }

int run(int M, int N, int K, int mode, int variant = 0) {
    const int a_rows = (mode % 10 == 2) ? K : M, a_cols = (mode % 10 == 2) ? M : K;
    const int b_rows = (mode % 10 == 0) ? N : K, b_cols = (mode % 10 == 0) ? K : N;
    std::vector<float> A(a_rows * a_cols), B(b_rows * b_cols), D(128 * N, -777.f), ref((size_t)M * N);
    for (int i = 0; i < a_rows; ++i) for (int j = 0; j < a_cols; ++j) A[i * a_cols + j] = val(i, j, 1);
    for (int i = 0; i < b_rows; ++i) for (int j = 0; j < b_cols; ++j) B[i * b_cols + j] = val(i, j, 4);
    for (int m = 0; m < M; ++m) for (int n = 0; n < N; ++n) {
        double s = 0;
        for (int k = 0; k < K; ++k) {
            float a = (mode % 10 == 2) ? A[k * M + m] : A[m * K + k];
            float b = (mode % 10 == 0) ? B[n * K + k] : B[k * N + n];
            s += (double)a * b;
        }
        ref[(size_t)m * N + n] = (float)s;
    }
    float *dA, *dB, *dD;
    cudaMalloc(&dA, A.size() * 4); cudaMalloc(&dB, B.size() * 4); cudaMalloc(&dD, D.size() * 4);
    cudaMemcpy(dA, A.data(), A.size() * 4, cudaMemcpyHostToDevice);
    cudaMemcpy(dB, B.data(), B.size() * 4, cudaMemcpyHostToDevice);
    cudaMemcpy(dD, D.data(), D.size() * 4, cudaMemcpyHostToDevice);
    size_t smem = (A.size() + B.size()) * 4 + 2048;
    cudaFuncSetAttribute(umma_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    umma_kernel<<<1, 128, smem>>>(dA, dB, dD, M, N, K, mode, variant);
    cudaError_t e = cudaDeviceSynchronize();
    if (e != cudaSuccess) { printf("mode %d M=%d N=%d K=%d: CUDA error %s\n", mode, M, N, K, cudaGetErrorString(e)); return 1; }
    cudaMemcpy(D.data(), dD, D.size() * 4, cudaMemcpyDeviceToHost);
    // try the identity lane map first, then report which lane holds each row
    int bad = 0;
    for (int m = 0; m < M; ++m) for (int n = 0; n < N; ++n) if (D[(size_t)m * N + n] != ref[(size_t)m * N + n]) ++bad;
    printf("mode %d var %d M=%3d N=%3d K=%3d: identity lane map mismatches = %d / %d\n", mode, variant, M, N, K, bad, M * N);
    if (bad) {
        // find for rows 0,1,8,16,31,32,63 which lane matches
        int rows[] = {0, 1, 15, 16, 17, 31, 32, 33, 47, 48, 63};
        for (int ri = 0; ri < 11; ++ri) {
            int m = rows[ri]; if (m >= M) continue;
            int found = -1;
            for (int l = 0; l < 128 && found < 0; ++l) {
                bool ok = true;
                for (int n = 0; n < N && ok; ++n) ok = D[(size_t)l * N + n] == ref[(size_t)m * N + n];
                if (ok) found = l;
            }
            printf("   row %2d -> lane %d\n", m, found);
        }
        { int nz = 0; for (size_t i = 0; i < D.size(); ++i) nz += (D[i] != 0.f); printf("   nonzero outputs: %d of %zu\n", nz, D.size()); }
        printf("   D[lane0][0..7]: "); for (int n = 0; n < 8; ++n) printf("%g ", D[n]); printf("\n   ref[0][0..7]:   ");
        for (int n = 0; n < 8; ++n) printf("%g ", ref[n]); printf("\n");
    }
    cudaFree(dA); cudaFree(dB); cudaFree(dD);
    return bad != 0;
}

int main() {
    int rc = 0;
    rc |= run(128, 64, 64, 0);
    rc |= run(64, 64, 64, 0);
    rc |= run(128, 64, 64, 10);
    rc |= run(128, 16, 64, 10);
    rc |= run(128, 64, 64, 11);
    rc |= run(128, 32, 64, 11);
    rc |= run(128, 64, 128, 12);
    rc |= run(64, 64, 128, 12);
    rc |= run(64, 32, 128, 12);
    rc |= run(128, 128, 64, 10);
    printf(rc ? "SOME CASES FAILED\n" : "ALL OK\n");
    return 0;
}
