// Stand-alone check of the fused hyperTem kernels (csrc/htem_fused.cu) against the verified unfused pair
// (tmix + gproj forward; gproj backward + tmix_bwd) through the C ABI, without Python.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O2 -o tools/htem_check tools/htem_check.cu -ldl
//   ./tools/htem_check [B N]            (from the repo root; loads gpt-st_b200/libgptst_b200.so)
// Prints max |difference| of every output next to its max |value|, mask bits that disagree with (out > 0), and launch times
// over ROT rotating buffer sets (> L2).
#include <cuda_runtime.h>
#include <dlfcn.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e_), __FILE__, __LINE__); return 2; } } while (0)

typedef int (*tmix_t)(const float*, const float*, float*, int, int, int, int, int, int, void*);
typedef int (*fwd2_t)(const float*, const float*, const float*, const float*, float*, int, int, long, long, int, int, int, void*);
typedef int (*bwd2_t)(const float*, const float*, const float*, const float*, float*, float*, float*, float*, int, int, long, long, int, int,
                      int, int, void*);
typedef int (*tmixb_t)(const float*, const float*, const float*, float*, float*, int, int, int, int, int, int, void*);
typedef int (*tsplits_t)(int, int);
typedef int (*splits_t)(int, int, int);
typedef long (*wbytes_t)(int);
typedef int (*pack_t)(const float*, void*, void*, int, void*);
typedef int (*hfwd_t)(const float*, const float*, const void*, const float*, float*, void*, float*, int, int, int, int, void*);
typedef int (*hbwd_t)(const float*, const void*, const float*, const void*, float*, float*, int, int, int, int, void*);

__global__ void fill(float* p, size_t n, uint32_t seed, float scale) {
    for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
        uint32_t h = (uint32_t)i * 2654435761u ^ seed;
        h ^= h >> 15; h *= 2246822519u; h ^= h >> 13; h *= 3266489917u; h ^= h >> 16;
        p[i] = ((float)(h & 0xffffff) / 8388608.f - 1.f) * scale;
    }
}
__global__ void diff(const float* a, const float* b, size_t n, unsigned int* out) {
    float md = 0.f, ma = 0.f;
    for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
        const float x = a[i], y = b[i];
        const float d = fabsf(x - y);
        md = (d > md || d != d) ? (d != d ? 3.0e38f : d) : md;
        ma = fmaxf(ma, fabsf(y));
    }
    atomicMax(out, __float_as_uint(md));
    atomicMax(out + 1, __float_as_uint(ma));
}
// mask rows are padded to Npad per (b, t)
__global__ void mask_check(const float* y, const uint2* mask, int BT, int N, int Npad, unsigned int* bad) {
    const size_t total = (size_t)BT * N * 64;
    for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < total; i += (size_t)gridDim.x * blockDim.x) {
        const size_t r = i >> 6;
        const int c = (int)(i & 63);
        const size_t bt = r / N, n = r % N;
        const uint2 m = mask[bt * Npad + n];
        const int p = 16 * (c & 3) + (c >> 2);          // ballot order of the fused kernels
        const unsigned bit = ((p < 32 ? m.x : m.y) >> (p & 31)) & 1u;
        if (bit != (y[i] > 0.f ? 1u : 0u)) atomicAdd(bad, 1u);
    }
}

static unsigned int* g_out;
static int report(const char* what, const float* a, const float* b, size_t n) {
    CK(cudaMemset(g_out, 0, 8));
    diff<<<592, 256>>>(a, b, n, g_out);
    unsigned int h[2];
    CK(cudaMemcpy(h, g_out, 8, cudaMemcpyDeviceToHost));
    float d, m;
    memcpy(&d, &h[0], 4); memcpy(&m, &h[1], 4);
    printf("    %-10s max|fused - ref| = %.3e   max|ref| = %.3e   rel = %.2e\n", what, d, m, m > 0 ? d / m : 0.0);
    return 0;
}

template <typename F>
static float time_it(F f, int iters) {
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0); cudaEventCreate(&e1);
    for (int i = 0; i < 3; ++i) f(i);
    cudaEventRecord(e0);
    for (int i = 0; i < iters; ++i) f(i);
    cudaEventRecord(e1);
    cudaEventSynchronize(e1);
    float ms = 0.f;
    cudaEventElapsedTime(&ms, e0, e1);
    return ms * 1000.f / iters;
}

static int run_case(void* L, int B, int N, float xscale, float gscale, bool timing) {
    tmix_t tmix = (tmix_t)dlsym(L, "gptst_tmix");
    fwd2_t fwd2 = (fwd2_t)dlsym(L, "gptst_gproj_fwd");
    bwd2_t bwd2 = (bwd2_t)dlsym(L, "gptst_gproj_bwd");
    tmixb_t tmixb = (tmixb_t)dlsym(L, "gptst_tmix_bwd");
    tsplits_t tsplits_f = (tsplits_t)dlsym(L, "gptst_tmix_bwd_splits");
    splits_t splits_f = (splits_t)dlsym(L, "gptst_gproj_splits");
    wbytes_t wbytes = (wbytes_t)dlsym(L, "gptst_hypertem_wfrag_bytes");
    pack_t pack = (pack_t)dlsym(L, "gptst_hypertem_pack_w");
    hfwd_t hfwd = (hfwd_t)dlsym(L, "gptst_hypertem_fwd");
    hbwd_t hbwd = (hbwd_t)dlsym(L, "gptst_hypertem_bwd");
    hfwd_t hfwd_ws = (hfwd_t)dlsym(L, "gptst_hypertem_fwd_ws");
    hbwd_t hbwd_ws = (hbwd_t)dlsym(L, "gptst_hypertem_bwd_ws");
    if (!tmix || !fwd2 || !bwd2 || !tmixb || !tsplits_f || !splits_f || !wbytes || !pack || !hfwd || !hbwd) { printf("missing symbol\n"); return 1; }
    const int T = 12, D = 64, ROT = timing ? 4 : 1;
    const int Npad = (N + 15) / 16 * 16;
    const size_t M = (size_t)B * T * N, A = M * D;
    printf("== B=%d N=%d  activations ~%.0e, gradients ~%.0e\n", B, N, xscale, gscale);
    float *eb[4], *dO[4], *out[4], *ret[4], *deb[4], *dret[4];
    for (int i = 0; i < ROT; ++i) {
        CK(cudaMalloc(&eb[i], A * 4)); CK(cudaMalloc(&dO[i], A * 4)); CK(cudaMalloc(&out[i], A * 4)); CK(cudaMalloc(&ret[i], A * 4));
        CK(cudaMalloc(&deb[i], A * 4)); CK(cudaMalloc(&dret[i], A * 4));
        fill<<<592, 256>>>(eb[i], A, 1u + 16 * i, xscale); fill<<<592, 256>>>(dO[i], A, 3u + 16 * i, gscale);
    }
    float *W, *bias, *Mn, *ret2, *out2, *dX2, *dR2, *dW2, *db2, *dM2;
    void *wf, *wb;
    uint2* mask;
    const int G = B * T;
    CK(cudaMalloc(&W, (size_t)G * D * D * 4)); CK(cudaMalloc(&bias, (size_t)G * D * 4)); CK(cudaMalloc(&Mn, (size_t)N * T * T * 4));
    CK(cudaMalloc(&wf, wbytes(G))); CK(cudaMalloc(&wb, wbytes(G)));
    CK(cudaMalloc(&mask, ((size_t)G * Npad + 16) * 8)); CK(cudaMalloc(&g_out, 16));
    CK(cudaMalloc(&ret2, A * 4)); CK(cudaMalloc(&out2, A * 4)); CK(cudaMalloc(&dX2, A * 4)); CK(cudaMalloc(&dR2, A * 4));
    const int sp = splits_f(G, N, D), spm = tsplits_f(B, N);
    CK(cudaMalloc(&dW2, (size_t)sp * G * D * D * 4)); CK(cudaMalloc(&db2, (size_t)sp * G * D * 4)); CK(cudaMalloc(&dM2, (size_t)spm * N * T * T * 4));
    fill<<<592, 256>>>(W, (size_t)G * D * D, 4u, 0.125f); fill<<<592, 256>>>(bias, (size_t)G * D, 6u, 0.5f * xscale);
    fill<<<592, 256>>>(Mn, (size_t)N * T * T, 8u, 0.2f);
    CK(cudaMemset(mask, 0, ((size_t)G * Npad + 16) * 8));
    CK(cudaDeviceSynchronize());

    // reference: tmix + time-grouped projection
    int rc = tmix(eb[0], Mn, ret2, B, T, N, D, 0, 0, 0);
    rc |= fwd2(ret2, W, bias, eb[0], out2, G, N, (long)N * D, (long)D, D, 1, 3, 0);
    int rcf = pack(W, wf, wb, G, 0);
    rcf |= hfwd(eb[0], Mn, wf, bias, out[0], mask, ret[0], B, T, N, D, 0);
    CK(cudaDeviceSynchronize());
    printf("  forward rc ref=%d fused=%d\n", rc, rcf);
    report("ret", ret[0], ret2, A);
    report("out", out[0], out2, A);
    CK(cudaMemset(g_out + 2, 0, 4));
    mask_check<<<592, 256>>>(out[0], mask, G, N, Npad, g_out + 2);
    unsigned int bad = 0;
    CK(cudaMemcpy(&bad, g_out + 2, 4, cudaMemcpyDeviceToHost));
    printf("    sign-mask bits that disagree with (out > 0): %u of %zu\n", bad, A);

    if (hfwd_ws) {       // warp-specialised variant against the same reference
        float *o3, *r3;
        uint2* m3;
        CK(cudaMalloc(&o3, A * 4)); CK(cudaMalloc(&r3, A * 4)); CK(cudaMalloc(&m3, ((size_t)G * Npad + 16) * 8));
        CK(cudaMemset(o3, 0xff, A * 4)); CK(cudaMemset(r3, 0xff, A * 4)); CK(cudaMemset(m3, 0, ((size_t)G * Npad + 16) * 8));
        const int rcw = hfwd_ws(eb[0], Mn, wf, bias, o3, m3, r3, B, T, N, D, 0);
        CK(cudaDeviceSynchronize());
        printf("  forward (warp-specialised) rc=%d\n", rcw);
        report("ret ws", r3, ret2, A);
        report("out ws", o3, out2, A);
        CK(cudaMemset(g_out + 2, 0, 4));
        mask_check<<<592, 256>>>(o3, m3, G, N, Npad, g_out + 2);
        CK(cudaMemcpy(&bad, g_out + 2, 4, cudaMemcpyDeviceToHost));
        printf("    sign-mask bits (ws) that disagree with (out > 0): %u of %zu\n", bad, A);
        cudaFree(o3); cudaFree(r3); cudaFree(m3);
    }

    // reference backward: projection backward (dret, dRes = dy) then the fused mix backward accumulates into dRes
    rc = bwd2(dO[0], out2, ret2, W, dX2, dW2, db2, dR2, G, N, (long)N * D, (long)D, D, 1, 3, sp, 0);
    CK(cudaDeviceSynchronize());
    rcf = hbwd(dO[0], mask, Mn, wb, deb[0], dret[0], B, T, N, D, 0);
    CK(cudaDeviceSynchronize());
    report("dret", dret[0], dX2, A);
    rc |= tmixb(dX2, eb[0], Mn, dR2, dM2, B, T, N, D, 3, spm, 0);
    CK(cudaDeviceSynchronize());
    printf("  backward rc ref=%d fused=%d\n", rc, rcf);
    report("deb", deb[0], dR2, A);
    if (hbwd_ws) {
        float *d3, *dr3;
        CK(cudaMalloc(&d3, A * 4)); CK(cudaMalloc(&dr3, A * 4));
        CK(cudaMemset(d3, 0xff, A * 4)); CK(cudaMemset(dr3, 0xff, A * 4));
        const int rcw = hbwd_ws(dO[0], mask, Mn, wb, d3, dr3, B, T, N, D, 0);
        CK(cudaDeviceSynchronize());
        printf("  backward (warp-specialised) rc=%d\n", rcw);
        report("dret ws", dr3, dret[0], A);
        report("deb ws", d3, dR2, A);
        cudaFree(d3); cudaFree(dr3);
    }

    if (timing) {
        const float tp = time_it([&](int) { pack(W, wf, wb, G, 0); }, 20);
        const float t_ref_f = time_it([&](int i) { tmix(eb[i % ROT], Mn, ret2, B, T, N, D, 0, 0, 0);
                                                   fwd2(ret2, W, bias, eb[i % ROT], out[i % ROT], G, N, (long)N * D, (long)D, D, 1, 3, 0); }, 20);
        const float t_f = time_it([&](int i) { hfwd(eb[i % ROT], Mn, wf, bias, out[i % ROT], mask, 0, B, T, N, D, 0); }, 40);
        const float t_fr = time_it([&](int i) { hfwd(eb[i % ROT], Mn, wf, bias, out[i % ROT], mask, ret[i % ROT], B, T, N, D, 0); }, 40);
        const float t_ref_b = time_it([&](int i) { bwd2(dO[i % ROT], out2, ret2, W, dX2, dW2, db2, deb[i % ROT], G, N, (long)N * D, (long)D, D, 1, 3, sp, 0);
                                                   tmixb(dX2, eb[i % ROT], Mn, deb[i % ROT], dM2, B, T, N, D, 3, spm, 0); }, 20);
        const float t_b = time_it([&](int i) { hbwd(dO[i % ROT], mask, Mn, wb, deb[i % ROT], 0, B, T, N, D, 0); }, 40);
        const float t_br = time_it([&](int i) { hbwd(dO[i % ROT], mask, Mn, wb, deb[i % ROT], dret[i % ROT], B, T, N, D, 0); }, 40);
        if (hfwd_ws) {
            const float t_w = time_it([&](int i) { hfwd_ws(eb[i % ROT], Mn, wf, bias, out[i % ROT], mask, 0, B, T, N, D, 0); }, 40);
            const float t_wr = time_it([&](int i) { hfwd_ws(eb[i % ROT], Mn, wf, bias, out[i % ROT], mask, ret[i % ROT], B, T, N, D, 0); }, 40);
            printf("  time (us): fwd warp-specialised %.1f (%.0f GB/s of 2A)  +ret %.1f\n", t_w, 2 * (double)A * 4 / t_w * 1e-3, t_wr);
        }
        if (hbwd_ws) {
            const float t_w = time_it([&](int i) { hbwd_ws(dO[i % ROT], mask, Mn, wb, deb[i % ROT], 0, B, T, N, D, 0); }, 40);
            const float t_wr = time_it([&](int i) { hbwd_ws(dO[i % ROT], mask, Mn, wb, deb[i % ROT], dret[i % ROT], B, T, N, D, 0); }, 40);
            printf("  time (us): bwd warp-specialised %.1f (%.0f GB/s of 2A)  +dret %.1f\n", t_w, 2 * (double)A * 4 / t_w * 1e-3, t_wr);
        }
        CK(cudaDeviceSynchronize());
        const double Ab = (double)A * 4;
        printf("  time (us): pack %.1f | fwd ref (tmix+gproj) %.1f  fused %.1f (%.0f GB/s of 2A)  fused+ret %.1f | bwd ref (gproj_bwd+tmix_bwd) %.1f  "
               "fused %.1f (%.0f GB/s of 2A)  fused+dret %.1f\n",
               tp, t_ref_f, t_f, 2 * Ab / t_f * 1e-3, t_fr, t_ref_b, t_b, 2 * Ab / t_b * 1e-3, t_br);
    }
    for (int i = 0; i < ROT; ++i) { cudaFree(eb[i]); cudaFree(dO[i]); cudaFree(out[i]); cudaFree(ret[i]); cudaFree(deb[i]); cudaFree(dret[i]); }
    cudaFree(W); cudaFree(bias); cudaFree(Mn); cudaFree(wf); cudaFree(wb); cudaFree(mask); cudaFree(ret2); cudaFree(out2);
    cudaFree(dX2); cudaFree(dR2); cudaFree(dW2); cudaFree(db2); cudaFree(dM2); cudaFree(g_out);
    printf("  last CUDA error: %s\n", cudaGetErrorString(cudaGetLastError()));
    return 0;
}

int main(int argc, char** argv) {
    void* L = dlopen("gpt-st_b200/libgptst_b200.so", RTLD_NOW);
    if (!L) { printf("dlopen failed: %s\n", dlerror()); return 1; }
    if (argc >= 3) return run_case(L, atoi(argv[1]), atoi(argv[2]), 1.f, 0.01f, true);
    int rc = run_case(L, 3, 37, 1.f, 0.01f, false);          // ragged chunk, fewer tasks than CTAs
    rc |= run_case(L, 5, 207, 1.f, 1e-6f, false);            // odd N (METR_LA), tiny gradients
    rc |= run_case(L, 2, 170, 3e4f, 1e3f, false);            // large activations (beyond fp16 range without the row scales)
    rc |= run_case(L, 64, 170, 1.f, 0.01f, true);            // PEMS08 geometry, timed
    return rc;
}
