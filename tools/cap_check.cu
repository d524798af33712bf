// Stand-alone check of the round-2 cap forward (gptst_cap_route_fwd + gptst_cap_hop_ev + gptst_cap_recon_proj) against the verified
// four-launch chain (gptst_cap_route_fwd + gptst_cap_hop_e1 + gptst_cap_recon_hop + gptst_gproj_fwd) through the C ABI,
// without Python, and of the fused routing backward against the verified pair.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O2 -o tools/cap_check tools/cap_check.cu -ldl
//   ./tools/cap_check [B N]            (from the repo root; loads gpt-st_b200/libgptst_b200.so)
// Prints max |difference| of every output next to its max |value| and launch times over ROT rotating buffer sets (> L2).
#include <cuda_runtime.h>
#include <dlfcn.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e_), __FILE__, __LINE__); return 2; } } while (0)

typedef int (*route_t)(const float*, const float*, const float*, const float*, float*, float*, int, int, int, int, int, int, int, void*);
typedef int (*hope1_t)(const float*, const float*, float*, int, int, int, int, int, void*);
typedef int (*reconhop_t)(const float*, const float*, const float*, const float*, float*, float*, int, int, int, int, int, int, void*);
typedef int (*fwd2_t)(const float*, const float*, const float*, const float*, float*, int, int, long, long, int, int, int, void*);
typedef int (*hopev_t)(const float*, const float*, float*, float*, int, int, int, int, int, void*);
typedef int (*reconproj_t)(const float*, const float*, const float*, const void*, const float*, float*, float*, int, int, int, int, int, void*);
typedef long (*wbytes_t)(int);
typedef int (*pack_t)(const float*, void*, void*, int, void*);

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
static unsigned int* g_out;
static int report(const char* what, const float* a, const float* b, size_t n) {
    CK(cudaMemset(g_out, 0, 8));
    diff<<<592, 256>>>(a, b, n, g_out);
    unsigned int h[2];
    CK(cudaMemcpy(h, g_out, 8, cudaMemcpyDeviceToHost));
    float d, m;
    memcpy(&d, &h[0], 4); memcpy(&m, &h[1], 4);
    printf("    %-10s max|new - ref| = %.3e   max|ref| = %.3e   rel = %.2e\n", what, d, m, m > 0 ? d / m : 0.0);
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

static int run_case(void* L, int B, int N, int H, bool timing) {
    route_t route = (route_t)dlsym(L, "gptst_cap_route_fwd");
    hope1_t hope1 = (hope1_t)dlsym(L, "gptst_cap_hop_e1");
    reconhop_t reconhop = (reconhop_t)dlsym(L, "gptst_cap_recon_hop");
    fwd2_t fwd2 = (fwd2_t)dlsym(L, "gptst_gproj_fwd");
    hopev_t hopev = (hopev_t)dlsym(L, "gptst_cap_hop_ev");
    reconproj_t reconproj = (reconproj_t)dlsym(L, "gptst_cap_recon_proj");
    wbytes_t wbytes = (wbytes_t)dlsym(L, "gptst_hypertem_wfrag_bytes");
    pack_t pack = (pack_t)dlsym(L, "gptst_hypertem_pack_w");
    if (!route || !hope1 || !reconhop || !fwd2 || !hopev || !reconproj || !wbytes || !pack) { printf("missing symbol\n"); return 1; }
    const int T = 12, D = 64, HT = 16, R = 2, ROT = timing ? 4 : 1;
    const size_t BT = (size_t)B * T, A = BT * N * D, C = BT * H * N, S = BT * H * D;
    printf("== B=%d N=%d H=%d\n", B, N, H);
    float *x[4], *out[4];
    for (int i = 0; i < ROT; ++i) {
        CK(cudaMalloc(&x[i], A * 4)); CK(cudaMalloc(&out[i], A * 4));
        fill<<<592, 256>>>(x[i], A, 1u + 16 * i, 1.f);
    }
    float *Wp, *bp, *dadj, *dyn, *Wn, *bn;
    float *c1, *s1, *e11, *v1, *recon1, *out1, *c2, *s2, *e12, *v2, *recon2;
    void* wf;
    CK(cudaMalloc(&Wp, D * D * 4)); CK(cudaMalloc(&bp, D * 4)); CK(cudaMalloc(&dadj, C * 4)); CK(cudaMalloc(&dyn, (size_t)B * HT * T * H * 4));
    CK(cudaMalloc(&Wn, (size_t)N * D * D * 4)); CK(cudaMalloc(&bn, (size_t)N * D * 4)); CK(cudaMalloc(&wf, wbytes(N)));
    CK(cudaMalloc(&c1, C * 4)); CK(cudaMalloc(&c2, C * 4)); CK(cudaMalloc(&s1, S * 4)); CK(cudaMalloc(&s2, S * 4));
    CK(cudaMalloc(&v1, S * 4)); CK(cudaMalloc(&v2, S * 4)); CK(cudaMalloc(&e11, (size_t)B * HT * D * 4)); CK(cudaMalloc(&e12, (size_t)B * HT * D * 4));
    CK(cudaMalloc(&recon1, A * 4)); CK(cudaMalloc(&recon2, A * 4)); CK(cudaMalloc(&out1, A * 4));
    CK(cudaMalloc(&g_out, 16));
    fill<<<592, 256>>>(Wp, D * D, 4u, 0.15f); fill<<<64, 64>>>(bp, D, 5u, 0.1f);
    fill<<<592, 256>>>(dadj, C, 6u, 1.5f); fill<<<592, 256>>>(dyn, (size_t)B * HT * T * H, 7u, 0.3f);
    fill<<<592, 256>>>(Wn, (size_t)N * D * D, 8u, 0.125f); fill<<<592, 256>>>(bn, (size_t)N * D, 9u, 0.3f);
    CK(cudaMemset(v2, 0xff, S * 4)); CK(cudaMemset(e12, 0xff, (size_t)B * HT * D * 4)); CK(cudaMemset(recon2, 0xff, A * 4)); CK(cudaMemset(out[0], 0xff, A * 4));
    CK(cudaDeviceSynchronize());

    int rc = route(x[0], Wp, bp, dadj, c1, s1, B, T, N, D, H, R, 3, 0);
    rc |= hope1(s1, dyn, e11, B, T, D, H, HT, 0);
    rc |= reconhop(c1, s1, dyn, e11, v1, recon1, B, T, N, D, H, HT, 0);
    rc |= fwd2(recon1, Wn, bn, x[0], out1, N, (int)BT, (long)D, (long)N * D, D, 1, 3, 0);
    CK(cudaDeviceSynchronize());
    int rn = pack(Wn, wf, nullptr, N, 0);
    rn |= route(x[0], Wp, bp, dadj, c2, s2, B, T, N, D, H, R, 3, 0);
    rn |= hopev(s2, dyn, e12, v2, B, T, D, H, HT, 0);
    rn |= reconproj(c2, v2, x[0], wf, bn, out[0], recon2, B, T, N, D, H, 0);
    CK(cudaDeviceSynchronize());
    printf("  forward rc ref=%d new=%d\n", rc, rn);
    report("c", c2, c1, C); report("s", s2, s1, S); report("e1", e12, e11, (size_t)B * HT * D); report("v", v2, v1, S);
    report("recon", recon2, recon1, A); report("out", out[0], out1, A);
    CK(cudaMemset(out[0], 0xff, A * 4));
    rn = reconproj(c2, v2, x[0], wf, bn, out[0], nullptr, B, T, N, D, H, 0);     // inference flavour: recon never written
    CK(cudaDeviceSynchronize());
    report("out (no recon)", out[0], out1, A);

    if (timing) {
        const float tp = time_it([&](int) { pack(Wn, wf, nullptr, N, 0); }, 20);
        const float t_route = time_it([&](int i) { route(x[i % ROT], Wp, bp, dadj, c1, s1, B, T, N, D, H, R, 3, 0); }, 30);
        const float t_rh = time_it([&](int) { hopev(s2, dyn, e12, v2, B, T, D, H, HT, 0); }, 30);
        const float t_e1 = time_it([&](int) { hope1(s1, dyn, e11, B, T, D, H, HT, 0); }, 30);
        const float t_rec = time_it([&](int) { reconhop(c1, s1, dyn, e11, v1, recon1, B, T, N, D, H, HT, 0); }, 30);
        const float t_gp = time_it([&](int i) { fwd2(recon1, Wn, bn, x[i % ROT], out[i % ROT], N, (int)BT, (long)D, (long)N * D, D, 1, 3, 0); }, 30);
        const float t_rp = time_it([&](int i) { reconproj(c2, v2, x[i % ROT], wf, bn, out[i % ROT], recon2, B, T, N, D, H, 0); }, 30);
        const float t_rp0 = time_it([&](int i) { reconproj(c2, v2, x[i % ROT], wf, bn, out[i % ROT], nullptr, B, T, N, D, H, 0); }, 30);
        const float t_old = time_it([&](int i) {
            route(x[i % ROT], Wp, bp, dadj, c1, s1, B, T, N, D, H, R, 3, 0); hope1(s1, dyn, e11, B, T, D, H, HT, 0);
            reconhop(c1, s1, dyn, e11, v1, recon1, B, T, N, D, H, HT, 0);
            fwd2(recon1, Wn, bn, x[i % ROT], out[i % ROT], N, (int)BT, (long)D, (long)N * D, D, 1, 3, 0); }, 30);
        const float t_new = time_it([&](int i) {
            route(x[i % ROT], Wp, bp, dadj, c2, s2, B, T, N, D, H, R, 3, 0); hopev(s2, dyn, e12, v2, B, T, D, H, HT, 0);
            reconproj(c2, v2, x[i % ROT], wf, bn, out[i % ROT], recon2, B, T, N, D, H, 0); }, 30);
        const float t_new0 = time_it([&](int i) {
            route(x[i % ROT], Wp, bp, dadj, c2, s2, B, T, N, D, H, R, 3, 0); hopev(s2, dyn, e12, v2, B, T, D, H, HT, 0);
            reconproj(c2, v2, x[i % ROT], wf, bn, out[i % ROT], nullptr, B, T, N, D, H, 0); }, 30);
        CK(cudaDeviceSynchronize());
        const double alg = 4.0 * BT * N * (2 * D + H);
        printf("  time (us): pack_wn %.1f | route %.1f  hop_ev %.1f | hop_e1 %.1f  recon_hop %.1f  gproj %.1f | recon_proj %.1f (no recon store %.1f)\n",
               tp, t_route, t_rh, t_e1, t_rec, t_gp, t_rp, t_rp0);
        printf("  chain (us): four launches %.1f (%.0f GB/s) | three launches %.1f (%.0f GB/s) | three launches, inference %.1f (%.0f GB/s)   [algorithmic %.2f MB]\n",
               t_old, alg / t_old * 1e-3, t_new, alg / t_new * 1e-3, t_new0, alg / t_new0 * 1e-3, alg * 1e-6);
    }
    for (int i = 0; i < ROT; ++i) { cudaFree(x[i]); cudaFree(out[i]); }
    cudaFree(Wp); cudaFree(bp); cudaFree(dadj); cudaFree(dyn); cudaFree(Wn); cudaFree(bn); cudaFree(wf);
    cudaFree(c1); cudaFree(c2); cudaFree(s1); cudaFree(s2); cudaFree(v1); cudaFree(v2); cudaFree(e11); cudaFree(e12);
    cudaFree(recon1); cudaFree(recon2); cudaFree(out1); cudaFree(g_out);
    printf("  last CUDA error: %s\n", cudaGetErrorString(cudaGetLastError()));
    return 0;
}

int main(int argc, char** argv) {
    void* L = dlopen("gpt-st_b200/libgptst_b200.so", RTLD_NOW);
    if (!L) { printf("dlopen failed: %s\n", dlerror()); return 1; }
    if (argc >= 3) return run_case(L, atoi(argv[1]), atoi(argv[2]), 10, true);
    int rc = run_case(L, 3, 37, 10, false);          // ragged node group / slab chunk
    rc |= run_case(L, 5, 207, 7, false);             // METR_LA nodes, fewer hyperedges
    rc |= run_case(L, 2, 60, 15, false);             // H = 15 (the wide-table flavour)
    rc |= run_case(L, 64, 170, 10, true);            // PEMS08 geometry, timed
    return rc;
}
