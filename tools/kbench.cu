// Per-kernel timing of the two heavy blocks (hyperTem, cap; forward and backward) through the C ABI, without Python:
// one process, ~3 s on a GPU box (a Python start-up alone costs more GPU-box time than this whole run).
//   nvcc -gencode arch=compute_100a,code=sm_100a -O2 -o tools/kbench tools/kbench.cu -ldl
//   ./tools/kbench [B=64] [N=170] [iters=30]        (from the repo root; loads gpt-st_b200/libgptst_b200.so)
// D = 64, T = 12, H = 10, HT = 16, two routing iterations, three-term split.  Every kernel is launched `iters` times over
// three rotating buffer sets (> L2 together), timed with CUDA events on the launching stream; the intermediate tensors of
// a set are real outputs of the forward chain (so squash / softmax see sane values), gradients are small random numbers.
// Prints microseconds per launch, the kernel's algorithmic bytes and the resulting GB/s.  Not a bench line.
#include <cuda_runtime.h>
#include <dlfcn.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>

#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e_), __FILE__, __LINE__); exit(2); } } while (0)
#define SYM(name) name##_t name = (name##_t)dlsym(L, "gptst_" #name); if (!name) { printf("missing symbol gptst_" #name "\n"); return 1; }

typedef const float* cf;
typedef int (*gproj_fwd_t)(cf, cf, cf, cf, float*, int, int, long, long, int, int, int, void*);
typedef int (*gproj_splits_t)(int, int, int);
typedef int (*gproj_bwd_t)(cf, cf, cf, cf, float*, float*, float*, float*, int, int, long, long, int, int, int, int, void*);
typedef int (*gproj3_bwd_t)(cf, const void*, cf, cf, float*, float*, float*, float*, int, int, long, long, int, int, int, int, int, void*);
typedef int (*tmix_t)(cf, cf, float*, int, int, int, int, int, int, void*);
typedef int (*tmix_bwd_splits_t)(int, int);
typedef int (*tmix_bwd_t)(cf, cf, cf, float*, float*, int, int, int, int, int, int, void*);
typedef int (*cap_route_fwd_t)(cf, cf, cf, cf, float*, float*, int, int, int, int, int, int, int, void*);
typedef int (*cap_hop_e1_t)(cf, cf, float*, int, int, int, int, int, void*);
typedef int (*cap_recon_hop_t)(cf, cf, cf, cf, float*, float*, int, int, int, int, int, int, void*);
typedef long (*hypertem_wfrag_bytes_t)(int);
typedef int (*hypertem_pack_w_t)(cf, void*, void*, int, void*);
typedef int (*hypertem_fwd_t)(cf, cf, const void*, cf, float*, void*, float*, int, int, int, int, void*);
typedef int (*hypertem_bwd_t)(cf, const void*, cf, const void*, float*, float*, int, int, int, int, void*);
typedef int (*hypertem_dw_t)(cf, const void*, cf, float*, float*, int, int, int, int, int, int, void*);
typedef int (*tmix_dM2_t)(cf, cf, float*, int, int, int, int, int, void*);
typedef int (*cap_dv_dcr_hoprows_t)(cf, cf, cf, cf, cf, cf, float*, float*, float*, int, int, int, int, int, int, void*);
typedef int (*cap_hop_bwd_parts_t)(int);
typedef int (*cap_hop_bwd_cols_t)(cf, cf, cf, cf, cf, float*, float*, int, int, int, int, int, void*);
typedef int (*cap_route_bwd_dz_t)(cf, cf, cf, cf, cf, cf, float*, float*, int, int, int, int, int, int, void*);
typedef int (*cap_route_fwd_z_t)(cf, cf, cf, cf, float*, float*, float*, int, int, int, int, int, int, int, void*);
typedef int (*cap_route_bwd_dz_z_t)(cf, cf, cf, cf, float*, float*, int, int, int, int, int, int, void*);
typedef int (*linear_bwd_acc_splits_t)(long, int);
typedef int (*linear_bwd_acc_t)(cf, cf, cf, float*, float*, float*, long, int, int, int, void*);
typedef int (*proj_out_fwd_t)(cf, cf, cf, float*, long, int, int, void*);
typedef int (*proj_out_bwd_parts_t)(long);
typedef int (*proj_out_bwd_t)(cf, cf, cf, float*, float*, long, int, int, int, void*);
typedef int (*score_head_fwd_t)(cf, cf, cf, float*, long, int, int, void*);

__global__ void fill(float* p, size_t n, uint32_t seed, float scale, float shift) {
    for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
        uint32_t h = (uint32_t)i * 2654435761u ^ seed;
        h ^= h >> 15; h *= 2246822519u; h ^= h >> 13; h *= 3266489917u; h ^= h >> 16;
        p[i] = ((float)(h & 0xffffff) / 8388608.f - 1.f) * scale + shift;
    }
}
static float* dev(size_t n, uint32_t seed = 0, float scale = 0.f, float shift = 0.f) {
    float* p;
    CK(cudaMalloc(&p, n * 4));
    if (seed) fill<<<592, 256>>>(p, n, seed, scale, shift);
    else CK(cudaMemset(p, 0, n * 4));
    return p;
}

constexpr int NSET = 3;
struct Set {   // one rotating buffer set: activations / gradients of one hyperTem and one cap block
    float *x, *dout, *ret, *out_t, *dret, *deb, *c, *s, *e1, *v, *recon, *out_n, *drecon, *dx, *dcr, *dr, *dp2, *ds, *dZ, *ddadj, *y1, *dy1;
    void* mask;
};

template <typename F>
static void bench(const char* name, double bytes, int iters, F f) {
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0); cudaEventCreate(&e1);
    int rc = 0;
    for (int i = 0; i < NSET; ++i) rc |= f(i);
    CK(cudaDeviceSynchronize());
    cudaEventRecord(e0);
    for (int i = 0; i < iters; ++i) rc |= f(i % NSET);
    cudaEventRecord(e1);
    CK(cudaEventSynchronize(e1));
    float ms = 0.f;
    cudaEventElapsedTime(&ms, e0, e1);
    const double us = ms * 1000.0 / iters;
    printf("%-44s %8.1f us   %7.1f MB   %7.0f GB/s   rc=%d\n", name, us, bytes / 1e6, bytes / us / 1e3, rc);
}

int main(int argc, char** argv) {
    const int B = argc > 1 ? atoi(argv[1]) : 64, N = argc > 2 ? atoi(argv[2]) : 170, iters = argc > 3 ? atoi(argv[3]) : 30;
    const int T = 12, D = 64, H = 10, HT = 16, RT = 2, prec = 3, K = T * H;
    void* L = dlopen("gpt-st_b200/libgptst_b200.so", RTLD_NOW);
    if (!L) { printf("dlopen failed: %s\n", dlerror()); return 1; }
    SYM(gproj_fwd) SYM(gproj_splits) SYM(gproj_bwd) SYM(gproj3_bwd) SYM(tmix) SYM(tmix_bwd_splits) SYM(tmix_bwd)
    SYM(cap_route_fwd) SYM(cap_hop_e1) SYM(cap_recon_hop) SYM(cap_dv_dcr_hoprows) SYM(cap_hop_bwd_parts) SYM(cap_hop_bwd_cols)
    SYM(cap_route_fwd_z) SYM(cap_route_bwd_dz_z)
    SYM(cap_route_bwd_dz) SYM(linear_bwd_acc_splits) SYM(linear_bwd_acc) SYM(proj_out_fwd) SYM(proj_out_bwd_parts) SYM(proj_out_bwd)
    SYM(score_head_fwd)
    SYM(hypertem_wfrag_bytes) SYM(hypertem_pack_w) SYM(hypertem_fwd) SYM(hypertem_bwd) SYM(hypertem_dw) SYM(tmix_dM2)

    const size_t M = (size_t)B * T * N, A = M * D, C = M * H, S = (size_t)B * T * H * D;
    const double Ab = A * 4.0, Cb = C * 4.0;
    // parameters / tables (shared by the sets)
    float* Mn = dev((size_t)N * T * T, 11, 0.2f);
    float* Wbt = dev((size_t)B * T * D * D, 12, 0.125f); float* bbt = dev((size_t)B * T * D, 13, 0.3f);
    float* Wn = dev((size_t)N * D * D, 14, 0.125f);      float* bn = dev((size_t)N * D, 15, 0.3f);
    float* Wp = dev((size_t)D * D, 16, 0.125f);          float* bp = dev(D, 17, 0.3f);
    float* dadj = dev(C, 18, 1.f);                       float* dyn = dev((size_t)B * HT * K, 19, 0.3f);
    float* W3 = dev((size_t)H * D, 20, 0.2f);            float* b3 = dev(H, 21, 0.3f);
    float* Wo = dev(D, 22, 0.2f);                        float* bo = dev(1, 23, 0.3f);
    // partial buffers
    const int sp_t = gproj_splits(B * T, N, D), sp_n = gproj_splits(N, B * T, D), sp_m = tmix_bwd_splits(B, N);
    const int sp_l = linear_bwd_acc_splits((long)M, D), hp = cap_hop_bwd_parts(D), pp = proj_out_bwd_parts((long)M);
    float* dWt = dev((size_t)sp_t * B * T * D * D); float* dbt = dev((size_t)sp_t * B * T * D);
    float* dWnp = dev((size_t)sp_n * N * D * D);    float* dbnp = dev((size_t)sp_n * N * D);
    float* dMp = dev((size_t)sp_m * N * T * T);     float* ddynp = dev((size_t)hp * B * HT * K);
    float* dWpp = dev((size_t)sp_l * D * D);        float* dbpp = dev((size_t)sp_l * D);
    float* pop = dev((size_t)pp * (D + 1));         float* prob = dev(C);
    printf("B=%d N=%d  splits: gproj time %d node %d, tmix_bwd %d, linear_bwd_acc %d, hop parts %d, proj_out parts %d\n", B, N, sp_t, sp_n,
           sp_m, sp_l, hp, pp);

    Set st[NSET];
    for (int i = 0; i < NSET; ++i) {
        Set& q = st[i];
        q.x = dev(A, 100 + i, 1.f); q.dout = dev(A, 200 + i, 0.01f);
        q.ret = dev(A); q.out_t = dev(A); q.dret = dev(A); q.deb = dev(A);
        q.c = dev(C); q.s = dev(S); q.e1 = dev((size_t)B * HT * D); q.v = dev(S); q.recon = dev(A); q.out_n = dev(A);
        q.drecon = dev(A); q.dx = dev(A); q.dcr = dev(C); q.dr = dev(S); q.dp2 = dev(S); q.ds = dev(S); q.dZ = dev(A); q.ddadj = dev(C);
        q.y1 = dev(M); q.dy1 = dev(M, 300 + i, 0.01f);
        CK(cudaMalloc(&q.mask, ((size_t)B * T * ((N + 15) / 16 * 16) + 16) * 8)); CK(cudaMemset(q.mask, 0xff, ((size_t)B * T * ((N + 15) / 16 * 16) + 16) * 8));
    }
    CK(cudaDeviceSynchronize());
    const long gsT = (long)N * D, rsT = D, gsN = D, rsN = (long)N * D;

    printf("---- hyperTem (GPTST.py:154-163)\n");
    bench("tmix                       fwd", 2 * Ab, iters, [&](int i) { return tmix(st[i].x, Mn, st[i].ret, B, T, N, D, 0, 0, 0); });
    bench("gproj time-grouped         fwd", 3 * Ab, iters, [&](int i) { return gproj_fwd(st[i].ret, Wbt, bbt, st[i].x, st[i].out_t, B * T, N, gsT, rsT, D, 1, prec, 0); });
    bench("gproj time-grouped         bwd", 5 * Ab, iters, [&](int i) { return gproj_bwd(st[i].dout, st[i].out_t, st[i].ret, Wbt, st[i].dret, dWt, dbt, st[i].deb, B * T, N, gsT, rsT, D, 1, prec, sp_t, 0); });
    bench("tmix_bwd (dx += M^T dy, dM) bwd", 4 * Ab, iters, [&](int i) { return tmix_bwd(st[i].dret, st[i].x, Mn, st[i].deb, dMp, B, T, N, D, prec, sp_m, 0); });

    {   // the fused block (csrc/htem_fused.cu): main-chain kernels + the two side-stream parameter-gradient kernels
        void *wf, *wb;
        CK(cudaMalloc(&wf, hypertem_wfrag_bytes(B * T))); CK(cudaMalloc(&wb, hypertem_wfrag_bytes(B * T)));
        const int npad = (N + 15) / 16 * 16;
        bench("hypertem pack_w (side stream)  ", 3 * (double)B * T * D * D * 4, iters, [&](int i) { return hypertem_pack_w(Wbt, wf, wb, B * T, 0); });
        bench("hypertem FUSED fwd (+mask,+ret)", 3 * Ab, iters, [&](int i) { return hypertem_fwd(st[i].x, Mn, wf, bbt, st[i].out_t, st[i].mask, st[i].ret, B, T, N, D, 0); });
        bench("hypertem FUSED fwd (no ret)    ", 2 * Ab, iters, [&](int i) { return hypertem_fwd(st[i].x, Mn, wf, bbt, st[i].out_t, st[i].mask, 0, B, T, N, D, 0); });
        bench("hypertem FUSED bwd (+dret)     ", 3 * Ab, iters, [&](int i) { return hypertem_bwd(st[i].dout, st[i].mask, Mn, wb, st[i].deb, st[i].dret, B, T, N, D, 0); });
        bench("hypertem dW_bt/db_bt (side)    ", 2 * Ab, iters, [&](int i) { return hypertem_dw(st[i].dout, st[i].mask, st[i].ret, dWt, dbt, B, T, N, D, npad, sp_t, 0); });
        bench("hypertem dM_n (side)           ", 2 * Ab, iters, [&](int i) { return tmix_dM2(st[i].dret, st[i].x, dMp, B, T, N, D, sp_m, 0); });
    }
    printf("---- cap (GPTST.py:100-141)\n");
    bench("cap_route_fwd (routing)    fwd", Ab + 2 * Cb, iters, [&](int i) { return cap_route_fwd(st[i].x, Wp, bp, dadj, st[i].c, st[i].s, B, T, N, D, H, RT, prec, 0); });
    bench("cap_hop_e1                 fwd", 0.1 * Ab, iters, [&](int i) { return cap_hop_e1(st[i].s, dyn, st[i].e1, B, T, D, H, HT, 0); });
    bench("cap_recon_hop              fwd", Ab + Cb, iters, [&](int i) { return cap_recon_hop(st[i].c, st[i].s, dyn, st[i].e1, st[i].v, st[i].recon, B, T, N, D, H, HT, 0); });
    bench("gproj node-grouped         fwd", 3 * Ab, iters, [&](int i) { return gproj_fwd(st[i].recon, Wn, bn, st[i].x, st[i].out_n, N, B * T, gsN, rsN, D, 1, prec, 0); });
    bench("gproj node-grouped         bwd", 5 * Ab, iters, [&](int i) { return gproj_bwd(st[i].dout, st[i].out_n, st[i].recon, Wn, st[i].drecon, dWnp, dbnp, st[i].dx, N, B * T, gsN, rsN, D, 1, prec, sp_n, 0); });
    bench("cap_dv_dcr_hoprows         bwd", Ab + 2 * Cb, iters, [&](int i) { return cap_dv_dcr_hoprows(st[i].c, st[i].v, st[i].drecon, st[i].s, dyn, st[i].e1, st[i].dcr, st[i].dr, st[i].dp2, B, T, N, D, H, HT, 0); });
    bench("cap_hop_bwd_cols           bwd", 0.3 * Ab, iters, [&](int i) { return cap_hop_bwd_cols(st[i].s, dyn, st[i].e1, st[i].dr, st[i].dp2, st[i].ds, ddynp, B, T, D, H, HT, 0); });
    bench("cap_route_fwd_z (+Z store) fwd", 2 * Ab + 2 * Cb, iters, [&](int i) { return cap_route_fwd_z(st[i].x, Wp, bp, dadj, st[i].c, st[i].s, st[i].dZ, B, T, N, D, H, RT, prec, 0); });
    bench("cap_route_bwd_dz_z (from Z) bwd", 2 * Ab + 3 * Cb, iters, [&](int i) { return cap_route_bwd_dz_z(st[i].x, st[i].c, st[i].ds, st[i].dcr, st[i].dZ, st[i].ddadj, B, T, N, D, H, prec, 0); });
    bench("cap_route_bwd_dz           bwd", 2 * Ab + 3 * Cb, iters, [&](int i) { return cap_route_bwd_dz(st[i].x, Wp, bp, st[i].c, st[i].ds, st[i].dcr, st[i].dZ, st[i].ddadj, B, T, N, D, H, prec, 0); });
    bench("linear_bwd_acc (ln_p)      bwd", 4 * Ab, iters, [&](int i) { return linear_bwd_acc(st[i].dZ, st[i].x, Wp, st[i].dx, dWpp, dbpp, (long)M, D, prec, sp_l, 0); });
    bench("gproj3 shared weight       bwd (sign-mask kernel)", 4 * Ab, iters, [&](int i) { return gproj3_bwd(st[i].dZ, 0, st[i].x, Wp, st[i].dx, dWpp, dbpp, 0, 1, (int)M, 0L, (long)D, D, 0, prec, sp_l, 3, 0); });

    printf("---- heads\n");
    bench("proj_out (dim_flow_out)    fwd", Ab, iters, [&](int i) { return proj_out_fwd(st[i].out_n, Wo, bo, st[i].y1, (long)M, D, 1, 0); });
    bench("proj_out                   bwd", 2 * Ab, iters, [&](int i) { return proj_out_bwd(st[i].dy1, st[i].out_n, Wo, st[i].dx, pop, (long)M, D, 1, pp, 0); });
    bench("score_head (ln3 + softmax) fwd", Ab + Cb, iters, [&](int i) { return score_head_fwd(st[i].out_t, W3, b3, prob, (long)M, D, H, 0); });
    cudaError_t e = cudaGetLastError();
    printf("last CUDA error: %s\n", cudaGetErrorString(e));
    return e == cudaSuccess ? 0 : 2;
}
