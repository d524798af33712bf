// Micro-benchmark: legacy mma.sync m16n8k8 tf32 throughput and fp32 FMA throughput on this GPU.
#include <cstdio>
#include <cuda_runtime.h>
#include <stdint.h>
__global__ void mma_loop(float* out, int iters) {
    float c[8][4] = {};
    uint32_t a[4] = {0x3f800000u, 0x3f800000u, 0x3f800000u, 0x3f800000u}, b[2] = {0x3f800000u, 0x3f000000u};
    for (int i = 0; i < iters; ++i) {
#pragma unroll
        for (int j = 0; j < 8; ++j)
            asm volatile("mma.sync.aligned.m16n8k8.row.col.f32.tf32.tf32.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
                         : "+f"(c[j][0]), "+f"(c[j][1]), "+f"(c[j][2]), "+f"(c[j][3])
                         : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b[0]), "r"(b[1]));
    }
    float s = 0; for (int j = 0; j < 8; ++j) s += c[j][0] + c[j][1] + c[j][2] + c[j][3];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
__global__ void mma_bf16_loop(float* out, int iters) {
    float c[8][4] = {};
    uint32_t a[4] = {0x3f803f80u, 0x3f803f80u, 0x3f803f80u, 0x3f803f80u}, b[2] = {0x3f803f80u, 0x3f003f00u};
    for (int i = 0; i < iters; ++i) {
#pragma unroll
        for (int j = 0; j < 8; ++j)
            asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
                         : "+f"(c[j][0]), "+f"(c[j][1]), "+f"(c[j][2]), "+f"(c[j][3])
                         : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b[0]), "r"(b[1]));
    }
    float s = 0; for (int j = 0; j < 8; ++j) s += c[j][0] + c[j][1] + c[j][2] + c[j][3];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
__global__ void fma_loop(float* out, int iters) {
    float c[16]; for (int j = 0; j < 16; ++j) c[j] = threadIdx.x * 1e-3f + j;
    float a = 1.0001f, b = 0.5f;
    for (int i = 0; i < iters; ++i) {
#pragma unroll
        for (int j = 0; j < 16; ++j) c[j] = fmaf(c[j], a, b);
    }
    float s = 0; for (int j = 0; j < 16; ++j) s += c[j];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
int main() {
    float* out; cudaMalloc(&out, 148 * 8 * 1024 * 4);
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    int iters = 20000;
    for (int warps : {4, 8, 16, 32}) {
        for (int rep = 0; rep < 2; ++rep) {
            cudaEventRecord(e0); mma_loop<<<148 * 2, warps * 32>>>(out, iters); cudaEventRecord(e1); cudaEventSynchronize(e1);
            float ms; cudaEventElapsedTime(&ms, e0, e1);
            double fl = 2.0 * 16 * 8 * 8 * 8.0 * iters * warps * 148 * 2;
            if (rep) printf("mma.sync tf32 m16n8k8: %2d warps/CTA x2 CTA/SM: %.1f TFLOP/s\n", warps, fl / ms / 1e9);
            cudaEventRecord(e0); mma_bf16_loop<<<148 * 2, warps * 32>>>(out, iters); cudaEventRecord(e1); cudaEventSynchronize(e1);
            cudaEventElapsedTime(&ms, e0, e1);
            fl = 2.0 * 16 * 8 * 16 * 8.0 * iters * warps * 148 * 2;
            if (rep) printf("mma.sync bf16 m16n8k16: %2d warps/CTA x2 CTA/SM: %.1f TFLOP/s\n", warps, fl / ms / 1e9);
        }
    }
    for (int rep = 0; rep < 2; ++rep) {
        cudaEventRecord(e0); fma_loop<<<148 * 4, 512>>>(out, iters); cudaEventRecord(e1); cudaEventSynchronize(e1);
        float ms; cudaEventElapsedTime(&ms, e0, e1);
        double fl = 2.0 * 16 * iters * 512.0 * 148 * 4;
        if (rep) printf("fp32 FMA: %.1f TFLOP/s\n", fl / ms / 1e9);
    }
    printf("err %s\n", cudaGetErrorString(cudaGetLastError()));
    return 0;
}
