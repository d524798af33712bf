// Stand-alone check of the EXPERIMENTAL projection kernels (csrc/gproj3.cu) against the second generation (csrc/gproj2.cu)
// through the C ABI, without Python: runs in ~2 s on a GPU box.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O2 -o tools/gproj3_check tools/gproj3_check.cu -ldl
//   ./tools/gproj3_check            (from the repo root; loads gpt-st_b200/libgptst_b200.so)
// PEMS08 geometry (B=64, T=12, N=170, D=64), three-term split.  For the time-grouped, the node-grouped and the shared-weight
// (linear_bwd_acc) flavour it prints: max |difference| of every output against gproj2 next to the output's max |value|, the
// number of sign-mask bits that disagree with (Y > 0), and the average launch time of both generations (same buffers every
// launch: ~170 MB of operands per launch, larger than L2 but not rotated -- a first look, not a bench number).
#include <cuda_runtime.h>
#include <dlfcn.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e_), __FILE__, __LINE__); return 2; } } while (0)

typedef int (*fwd2_t)(const float*, const float*, const float*, const float*, float*, int, int, long, long, int, int, int, void*);
typedef int (*bwd2_t)(const float*, const float*, const float*, const float*, float*, float*, float*, float*, int, int, long, long, int, int,
                      int, int, void*);
typedef int (*fwd3_t)(const float*, const float*, const float*, const float*, float*, void*, int, int, long, long, int, int, int, void*);
typedef int (*bwd3_t)(const float*, const void*, const float*, const float*, float*, float*, float*, float*, int, int, long, long, int, int,
                      int, int, int, void*);
typedef int (*tmixb_t)(const float*, const float*, const float*, float*, float*, int, int, int, int, int, int, void*);
typedef int (*tmix3b_t)(const float*, const float*, const float*, const float*, const void*, float*, float*, int, int, int, int, int, int,
                        void*);
typedef int (*tsplits_t)(int, int);
typedef int (*splits_t)(int, int, int);
typedef int (*lsplits_t)(long, int);
typedef int (*lacc_t)(const float*, const float*, const float*, float*, float*, float*, long, int, int, int, void*);

__global__ void fill(float* p, size_t n, uint32_t seed, float scale) {
    for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
        uint32_t h = (uint32_t)i * 2654435761u ^ seed;
        h ^= h >> 15; h *= 2246822519u; h ^= h >> 13; h *= 3266489917u; h ^= h >> 16;
        p[i] = ((float)(h & 0xffffff) / 8388608.f - 1.f) * scale;
    }
}
// out[0] = max |a - b| , out[1] = max |a|   (non-negative floats order like their bit patterns)
__global__ void diff(const float* a, const float* b, size_t n, unsigned int* out) {
    float md = 0.f, ma = 0.f;
    for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
        const float x = a[i], y = b[i];
        const float d = fabsf(x - y);
        md = (d > md || d != d) ? (d != d ? 3.0e38f : d) : md;
        ma = fmaxf(ma, fabsf(x));
    }
    atomicMax(out, __float_as_uint(md));
    atomicMax(out + 1, __float_as_uint(ma));
}
// counts mask bits that differ from (y > 0); row = element offset / 64
__global__ void mask_check(const float* y, const uint2* mask, size_t rows, unsigned int* bad) {
    for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < rows * 64; i += (size_t)gridDim.x * blockDim.x) {
        const size_t r = i >> 6;
        const int c = (int)(i & 63);
        const uint2 m = mask[r];
        const unsigned bit = ((c < 32 ? m.x : m.y) >> (c & 31)) & 1u;
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
    printf("    %-10s max|v3 - v2| = %.3e   max|v2| = %.3e   rel = %.2e\n", what, d, m, m > 0 ? d / m : 0.0);
    return 0;
}

template <typename F>
static float time_it(F f, int iters) {
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0); cudaEventCreate(&e1);
    for (int i = 0; i < 3; ++i) f();
    cudaEventRecord(e0);
    for (int i = 0; i < iters; ++i) f();
    cudaEventRecord(e1);
    cudaEventSynchronize(e1);
    float ms = 0.f;
    cudaEventElapsedTime(&ms, e0, e1);
    return ms * 1000.f / iters;
}

int main() {
    void* L = dlopen("gpt-st_b200/libgptst_b200.so", RTLD_NOW);
    if (!L) { printf("dlopen failed: %s\n", dlerror()); return 1; }
    fwd2_t fwd2 = (fwd2_t)dlsym(L, "gptst_gproj_fwd");
    bwd2_t bwd2 = (bwd2_t)dlsym(L, "gptst_gproj_bwd");
    fwd3_t fwd3 = (fwd3_t)dlsym(L, "gptst_gproj3_fwd");
    bwd3_t bwd3 = (bwd3_t)dlsym(L, "gptst_gproj3_bwd");
    splits_t splits_f = (splits_t)dlsym(L, "gptst_gproj_splits");
    lsplits_t lsplits_f = (lsplits_t)dlsym(L, "gptst_linear_bwd_acc_splits");
    lacc_t lacc = (lacc_t)dlsym(L, "gptst_linear_bwd_acc");
    tmixb_t tmixb = (tmixb_t)dlsym(L, "gptst_tmix_bwd");
    tmix3b_t tmix3b = (tmix3b_t)dlsym(L, "gptst_tmix3_bwd");
    tsplits_t tsplits_f = (tsplits_t)dlsym(L, "gptst_tmix_bwd_splits");
    if (!fwd2 || !bwd2 || !fwd3 || !bwd3 || !splits_f || !lsplits_f || !lacc || !tmixb || !tmix3b || !tsplits_f) { printf("missing symbol\n"); return 1; }

    const int B = 64, T = 12, N = 170, D = 64;
    const size_t M = (size_t)B * T * N, A = M * D;
    float *X, *Res, *dY, *Wt, *Wn, *bt, *bn, *Y2, *Y3, *dX2, *dX3, *dR2, *dR3, *dW2, *dW3, *db2, *db3;
    uint2* mask;
    const size_t wmax = (size_t)8 * 768 * D * D;      // room for up to 8 splits of 768 groups
    CK(cudaMalloc(&X, A * 4)); CK(cudaMalloc(&Res, A * 4)); CK(cudaMalloc(&dY, A * 4));
    CK(cudaMalloc(&Y2, A * 4)); CK(cudaMalloc(&Y3, A * 4)); CK(cudaMalloc(&dX2, A * 4)); CK(cudaMalloc(&dX3, A * 4));
    CK(cudaMalloc(&dR2, A * 4)); CK(cudaMalloc(&dR3, A * 4));
    CK(cudaMalloc(&Wt, (size_t)B * T * D * D * 4)); CK(cudaMalloc(&Wn, (size_t)N * D * D * 4));
    CK(cudaMalloc(&bt, (size_t)B * T * D * 4)); CK(cudaMalloc(&bn, (size_t)N * D * 4));
    CK(cudaMalloc(&dW2, wmax * 4)); CK(cudaMalloc(&dW3, wmax * 4)); CK(cudaMalloc(&db2, wmax / D * 4)); CK(cudaMalloc(&db3, wmax / D * 4));
    CK(cudaMalloc(&mask, M * sizeof(uint2))); CK(cudaMalloc(&g_out, 16));
    fill<<<592, 256>>>(X, A, 1u, 1.f); fill<<<592, 256>>>(Res, A, 2u, 1.f); fill<<<592, 256>>>(dY, A, 3u, 0.01f);
    fill<<<592, 256>>>(Wt, (size_t)B * T * D * D, 4u, 0.125f); fill<<<592, 256>>>(Wn, (size_t)N * D * D, 5u, 0.125f);
    fill<<<592, 256>>>(bt, (size_t)B * T * D, 6u, 0.5f); fill<<<592, 256>>>(bn, (size_t)N * D, 7u, 0.5f);
    CK(cudaDeviceSynchronize());

    // hyperTem backward pair: mix matrix and dM partial buffers
    const int spm = tsplits_f(B, N);
    float *Mn, *dM2, *dM3;
    CK(cudaMalloc(&Mn, (size_t)N * T * T * 4)); CK(cudaMalloc(&dM2, (size_t)spm * N * T * T * 4)); CK(cudaMalloc(&dM3, (size_t)spm * N * T * T * 4));
    fill<<<592, 256>>>(Mn, (size_t)N * T * T, 8u, 0.2f);
    CK(cudaDeviceSynchronize());

    struct Case { const char* name; int G, R; long gs, rs; const float *W, *b; };
    const Case cases[2] = {{"time-grouped (hyperTem)", B * T, N, (long)N * D, (long)D, Wt, bt},
                           {"node-grouped (cap)", N, B * T, (long)D, (long)N * D, Wn, bn}};
    for (const Case& c : cases) {
        const int sp = splits_f(c.G, c.R, D);
        printf("== %s: G=%d R=%d splits=%d\n", c.name, c.G, c.R, sp);
        if ((size_t)sp * c.G * D * D > wmax) { printf("partial buffer too small\n"); return 1; }
        int rc = fwd2(X, c.W, c.b, Res, Y2, c.G, c.R, c.gs, c.rs, D, 1, 3, 0);
        int rc3 = fwd3(X, c.W, c.b, Res, Y3, mask, c.G, c.R, c.gs, c.rs, D, 1, 3, 0);
        CK(cudaDeviceSynchronize());
        printf("  forward rc v2=%d v3=%d\n", rc, rc3);
        report("Y", Y3, Y2, A);
        CK(cudaMemset(g_out + 2, 0, 4));
        mask_check<<<592, 256>>>(Y3, mask, M, g_out + 2);
        unsigned int bad = 0;
        CK(cudaMemcpy(&bad, g_out + 2, 4, cudaMemcpyDeviceToHost));
        printf("    sign-mask bits that disagree with (Y > 0): %u of %zu\n", bad, A);
        rc = bwd2(dY, Y2, X, c.W, dX2, dW2, db2, dR2, c.G, c.R, c.gs, c.rs, D, 1, 3, sp, 0);
        rc3 = bwd3(dY, mask, X, c.W, dX3, dW3, db3, dR3, c.G, c.R, c.gs, c.rs, D, 1, 3, sp, 0, 0);
        CK(cudaDeviceSynchronize());
        printf("  backward rc v2=%d v3=%d\n", rc, rc3);
        report("dX", dX3, dX2, A); report("dRes", dR3, dR2, A);
        report("dW_part", dW3, dW2, (size_t)sp * c.G * D * D); report("db_part", db3, db2, (size_t)sp * c.G * D);
        const float tf2 = time_it([&] { fwd2(X, c.W, c.b, Res, Y2, c.G, c.R, c.gs, c.rs, D, 1, 3, 0); }, 20);
        const float tf3 = time_it([&] { fwd3(X, c.W, c.b, Res, Y3, mask, c.G, c.R, c.gs, c.rs, D, 1, 3, 0); }, 20);
        const float tb2 = time_it([&] { bwd2(dY, Y2, X, c.W, dX2, dW2, db2, dR2, c.G, c.R, c.gs, c.rs, D, 1, 3, sp, 0); }, 20);
        const float tb3 = time_it([&] { bwd3(dY, mask, X, c.W, dX3, dW3, db3, dR3, c.G, c.R, c.gs, c.rs, D, 1, 3, sp, 0, 0); }, 20);
        CK(cudaDeviceSynchronize());
        printf("  time per launch (us): fwd v2 %.1f  v3(+mask) %.1f | bwd v2 %.1f  v3 %.1f\n", tf2, tf3, tb2, tb3);
        if (c.G == B * T) {
            // the whole hyperTem backward: default pair (projection backward writes dRes, the mix backward accumulates into it) against
            // the experimental pair (no dRes store; the mix backward rebuilds dOut * act' from dOut and the sign mask).  eb := Res.
            auto pair2 = [&] { return bwd2(dY, Y2, X, c.W, dX2, dW2, db2, dR2, c.G, c.R, c.gs, c.rs, D, 1, 3, sp, 0) |
                                      tmixb(dX2, Res, Mn, dR2, dM2, B, T, N, D, 3, spm, 0); };
            auto pair3 = [&] { return bwd3(dY, mask, X, c.W, dX3, dW3, db3, 0, c.G, c.R, c.gs, c.rs, D, 1, 3, sp, 0, 0) |
                                      tmix3b(dX3, Res, Mn, dY, mask, dR3, dM3, B, T, N, D, 3, spm, 0); };
            rc = pair2(); rc3 = pair3();
            CK(cudaDeviceSynchronize());
            printf("  hyperTem backward pair rc v2=%d v3=%d (tmix splits %d)\n", rc, rc3, spm);
            report("d eb", dR3, dR2, A); report("dM_part", dM3, dM2, (size_t)spm * N * T * T);
            const float tp2 = time_it([&] { pair2(); }, 20), tp3 = time_it([&] { pair3(); }, 20);
            CK(cudaDeviceSynchronize());
            printf("  time per pair (us): default (gproj_bwd + tmix_bwd) %.1f  experimental (gproj3_bwd without dRes + tmix3_bwd) %.1f\n", tp2, tp3);
        }
    }
    {   // shared weight, dX accumulated in place (ln_p of cap): gptst_linear_bwd_acc vs gproj3 with flags = 3
        const int sp = lsplits_f((long)M, D);
        printf("== shared weight (linear_bwd_acc): rows=%zu splits=%d\n", M, sp);
        CK(cudaMemcpy(dX2, Res, A * 4, cudaMemcpyDeviceToDevice)); CK(cudaMemcpy(dX3, Res, A * 4, cudaMemcpyDeviceToDevice));
        int rc = lacc(dY, X, Wn, dX2, dW2, db2, (long)M, D, 3, sp, 0);
        int rc3 = bwd3(dY, 0, X, Wn, dX3, dW3, db3, 0, 1, (int)M, 0L, (long)D, D, 0, 3, sp, 3, 0);
        CK(cudaDeviceSynchronize());
        printf("  rc v2=%d v3=%d\n", rc, rc3);
        report("dX_io", dX3, dX2, A); report("dW_part", dW3, dW2, (size_t)sp * D * D); report("db_part", db3, db2, (size_t)sp * D);
        const float t2 = time_it([&] { lacc(dY, X, Wn, dX2, dW2, db2, (long)M, D, 3, sp, 0); }, 20);
        const float t3 = time_it([&] { bwd3(dY, 0, X, Wn, dX3, dW3, db3, 0, 1, (int)M, 0L, (long)D, D, 0, 3, sp, 3, 0); }, 20);
        CK(cudaDeviceSynchronize());
        printf("  time per launch (us): v2 %.1f  v3 %.1f\n", t2, t3);
    }
    {   // cap_recon_hop3 (v rows held in registers) against cap_recon_hop on real routing outputs
        typedef int (*route_t)(const float*, const float*, const float*, const float*, float*, float*, int, int, int, int, int, int, int, void*);
        typedef int (*e1_t)(const float*, const float*, float*, int, int, int, int, int, void*);
        typedef int (*rh_t)(const float*, const float*, const float*, const float*, float*, float*, int, int, int, int, int, int, void*);
        route_t route = (route_t)dlsym(L, "gptst_cap_route_fwd");
        e1_t hop_e1 = (e1_t)dlsym(L, "gptst_cap_hop_e1");
        rh_t rh2 = (rh_t)dlsym(L, "gptst_cap_recon_hop"), rh3 = (rh_t)dlsym(L, "gptst_cap_recon_hop3");
        if (!route || !hop_e1 || !rh2 || !rh3) { printf("missing symbol (recon_hop)\n"); return 1; }
        const int H = 10, HT = 16;
        const size_t C = M * H, S = (size_t)B * T * H * D;
        float *dadj, *dyn, *cc, *ss, *e1, *v2, *v3;
        CK(cudaMalloc(&dadj, C * 4)); CK(cudaMalloc(&cc, C * 4)); CK(cudaMalloc(&ss, S * 4)); CK(cudaMalloc(&v2, S * 4)); CK(cudaMalloc(&v3, S * 4));
        CK(cudaMalloc(&dyn, (size_t)B * HT * T * H * 4)); CK(cudaMalloc(&e1, (size_t)B * HT * D * 4));
        fill<<<592, 256>>>(dadj, C, 9u, 1.f); fill<<<592, 256>>>(dyn, (size_t)B * HT * T * H, 10u, 0.3f);
        int rc = route(X, Wn, bn, dadj, cc, ss, B, T, N, D, H, 2, 3, 0);          // Wn / bn: any (64,64) weight and (64) bias
        rc |= hop_e1(ss, dyn, e1, B, T, D, H, HT, 0);
        const int r2 = rh2(cc, ss, dyn, e1, v2, Y2, B, T, N, D, H, HT, 0), r3 = rh3(cc, ss, dyn, e1, v3, Y3, B, T, N, D, H, HT, 0);
        CK(cudaDeviceSynchronize());
        printf("== cap recon_hop: route/hop_e1 rc=%d, recon_hop rc v2=%d v3=%d\n", rc, r2, r3);
        report("recon", Y3, Y2, A); report("v", v3, v2, S);
        const float t2 = time_it([&] { rh2(cc, ss, dyn, e1, v2, Y2, B, T, N, D, H, HT, 0); }, 20);
        const float t3 = time_it([&] { rh3(cc, ss, dyn, e1, v3, Y3, B, T, N, D, H, HT, 0); }, 20);
        CK(cudaDeviceSynchronize());
        printf("  time per launch (us): v2 %.1f  v3 %.1f\n", t2, t3);
    }
    cudaError_t e = cudaGetLastError();
    printf("last CUDA error: %s\n", cudaGetErrorString(e));
    return e == cudaSuccess ? 0 : 2;
}
