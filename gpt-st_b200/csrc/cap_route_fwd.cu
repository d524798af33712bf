// cap: intra-cluster routing forward (reference GPTST.py:102-123), one thread-block cluster per (b,t) slab.
//
//   P = squash(x Wp^T + bp)                                   (N x D, stays in shared memory)
//   pass A : u = squash(softmax_H(dadj) . P) ;  sumP = 1^T P   (first routing iteration has c = 1/H)
//   pass k : b += v P^T ; c = softmax_H(b) ; v' = squash(u * (c P))             (R-1 times)
//   final  : b += v P^T ; c = softmax_H(b + dadj) -> c_out ; s = c P -> s_out
//
// All contractions run on tensor cores (mma.sync m16n8k8 tf32, 1x or 3x split):
//   Z  = x Wp^T        : warp tile 16 nodes x D
//   b^T = v P^T        : M = 16 hyperedges (H padded), N = 8 nodes, K = D        "logit MMA"
//   acc = c P          : M = 16 hyperedges, N = D, K = 8 nodes                   "aggregation MMA"
// The logit MMA leaves c^T for nodes (2t, 2t+1) in lane (g,t); the aggregation MMA consumes exactly that as
// its A fragment when the K index is permuted as k=t <-> node 2t, k=t+4 <-> node 2t+1 (B rows are fetched
// with the same permutation), so the softmax output never leaves registers.
// The N-reduction is finished with a deterministic cross-warp tree in shared memory and, when the slab is
// split over a cluster, a DSMEM all-gather of the per-CTA partial (H+1) x D sums.
#include <cstdlib>

#include "cap_common.cuh"

namespace gptst {

constexpr int kHP = 16;  // hyperedge rows of the MMA tiles; row H carries the all-ones row in pass A => H <= 15

struct RF {
    float *Ps, *Wred, *bps, *dadj, *bl, *part, *tot, *u;
    uint32_t *vh, *vl;
    int ldp, ldv;
};

__host__ __device__ inline size_t rf_smem_floats(int D, int H, int RPC) {
    size_t wred = (size_t)D * (D + 4);
    size_t red = (size_t)kWarps * (H + 1) * D;
    if (red > wred) wred = red;
    size_t n = 0;
    n += (size_t)RPC * (D + 4);        // Ps
    n += wred;                         // Wp tile, later the cross-warp reduction buffer
    n += D;                            // bps
    n += 2 * (size_t)H * RPC;          // dadj, bl
    n += 2 * (size_t)kHP * (D + 4);    // vh, vl
    n += 2 * (size_t)(H + 1) * D;      // part (double buffered)
    n += (size_t)(H + 1) * D;          // tot
    n += (size_t)H * D;                // u
    return n;
}

// v (H x D fp32, row-major in `src`) -> tf32 hi/lo planes with zero rows for h >= H
template <int D, int PREC>
__device__ __forceinline__ void load_v_planes(const RF& sm, const float* src, int H) {
    for (int i = threadIdx.x; i < kHP * D; i += blockDim.x) {
        const int h = i / D, d = i % D;
        uint32_t hi = 0u, lo = 0u;
        if (h < H) split_tf32<PREC>(src[h * D + d], hi, lo);
        sm.vh[h * sm.ldv + d] = hi;
        sm.vl[h * sm.ldv + d] = lo;
    }
}

// mode 0: c = softmax(dadj), row H of c := 1 (ones row), logits untouched
// mode 1: bl += v.P ; c = softmax(bl)
// mode 2: bl += v.P (if use_v) ; c = softmax(bl + dadj) ; c -> global
template <int D, int PREC>
__device__ void route_pass_mma(const RF& sm, int mode, bool use_v, int H, int RPC, int nloc, int pass_idx,
                               float* __restrict__ c_out, int N_stride, cg::cluster_group& cluster, int CS) {
    constexpr int NT = D / 8;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int g = lane >> 2, tq = lane & 3;
    const int HA = (mode == 0) ? H + 1 : H;
    float acc[NT][4];
#pragma unroll
    for (int j = 0; j < NT; ++j) acc[j][0] = acc[j][1] = acc[j][2] = acc[j][3] = 0.f;
    const int ngroups = (nloc + 7) / 8;
    const int h0 = g, h1 = g + 8;
    const bool vh0 = h0 < H, vh1 = h1 < H;
    for (int grp = warp; grp < ngroups; grp += kWarps) {
        const int nb = grp * 8;
        float z[4] = {0.f, 0.f, 0.f, 0.f};   // (h0,na) (h0,nb) (h1,na) (h1,nb)
        if (mode != 0 && use_v) {
            // six independent accumulators (3 split terms x even/odd k-step) keep the dependent mma chain at D/16 instead
            // of 3*D/8: with 2-4 warps per scheduler the chain latency, not the tensor pipe, was the limiter
            float zz[6][4];
#pragma unroll
            for (int i = 0; i < 6; ++i) zz[i][0] = zz[i][1] = zz[i][2] = zz[i][3] = 0.f;
#pragma unroll
            for (int k0 = 0; k0 < D; k0 += 8) {
                const int par = (k0 >> 3) & 1;
                uint32_t ah[4], al[4], bh[2], bl2[2];
                ah[0] = sm.vh[h0 * sm.ldv + k0 + tq];     ah[1] = sm.vh[h1 * sm.ldv + k0 + tq];
                ah[2] = sm.vh[h0 * sm.ldv + k0 + tq + 4]; ah[3] = sm.vh[h1 * sm.ldv + k0 + tq + 4];
                split_tf32<PREC>(sm.Ps[(size_t)(nb + g) * sm.ldp + k0 + tq], bh[0], bl2[0]);
                split_tf32<PREC>(sm.Ps[(size_t)(nb + g) * sm.ldp + k0 + tq + 4], bh[1], bl2[1]);
                if (PREC == PREC_3XTF32) {
                    al[0] = sm.vl[h0 * sm.ldv + k0 + tq];     al[1] = sm.vl[h1 * sm.ldv + k0 + tq];
                    al[2] = sm.vl[h0 * sm.ldv + k0 + tq + 4]; al[3] = sm.vl[h1 * sm.ldv + k0 + tq + 4];
                    mma_tf32(zz[par], al, bh);
                    mma_tf32(zz[2 + par], ah, bl2);
                }
                mma_tf32(zz[4 + par], ah, bh);
            }
#pragma unroll
            for (int i = 0; i < 4; ++i) z[i] = ((zz[0][i] + zz[1][i]) + (zz[2][i] + zz[3][i])) + (zz[4][i] + zz[5][i]);
        }
        const int na = nb + 2 * tq, nbb = na + 1;
        const bool va = na < nloc, vb = nbb < nloc;
        if (mode != 0) {
            if (vh0) { const float2 t = *reinterpret_cast<const float2*>(sm.bl + h0 * RPC + na); z[0] += t.x; z[1] += t.y; }
            if (vh1) { const float2 t = *reinterpret_cast<const float2*>(sm.bl + h1 * RPC + na); z[2] += t.x; z[3] += t.y; }
            if (mode == 1) {
                if (vh0) *reinterpret_cast<float2*>(sm.bl + h0 * RPC + na) = make_float2(z[0], z[1]);
                if (vh1) *reinterpret_cast<float2*>(sm.bl + h1 * RPC + na) = make_float2(z[2], z[3]);
            }
        }
        if (mode != 1) {
            if (vh0) { const float2 t = *reinterpret_cast<const float2*>(sm.dadj + h0 * RPC + na); z[0] += t.x; z[1] += t.y; }
            if (vh1) { const float2 t = *reinterpret_cast<const float2*>(sm.dadj + h1 * RPC + na); z[2] += t.x; z[3] += t.y; }
        }
        // softmax over h: the 16 logits of node na live in z[0],z[2] of the 8 lanes sharing tq
        float ma = fmaxf(vh0 ? z[0] : -INFINITY, vh1 ? z[2] : -INFINITY);
        float mb = fmaxf(vh0 ? z[1] : -INFINITY, vh1 ? z[3] : -INFINITY);
#pragma unroll
        for (int o = 4; o < 32; o <<= 1) {
            ma = fmaxf(ma, __shfl_xor_sync(0xffffffffu, ma, o));
            mb = fmaxf(mb, __shfl_xor_sync(0xffffffffu, mb, o));
        }
        float c[4];
        c[0] = vh0 ? __expf(z[0] - ma) : 0.f; c[1] = vh0 ? __expf(z[1] - mb) : 0.f;
        c[2] = vh1 ? __expf(z[2] - ma) : 0.f; c[3] = vh1 ? __expf(z[3] - mb) : 0.f;
        float sa = c[0] + c[2], sb = c[1] + c[3];
#pragma unroll
        for (int o = 4; o < 32; o <<= 1) {
            sa += __shfl_xor_sync(0xffffffffu, sa, o);
            sb += __shfl_xor_sync(0xffffffffu, sb, o);
        }
        const float ia = va ? 1.f / sa : 0.f, ib = vb ? 1.f / sb : 0.f;
        c[0] *= ia; c[2] *= ia; c[1] *= ib; c[3] *= ib;
        if (mode == 2) {
            if (vh0) { if (va) c_out[(size_t)h0 * N_stride + na] = c[0]; if (vb) c_out[(size_t)h0 * N_stride + nbb] = c[1]; }
            if (vh1) { if (va) c_out[(size_t)h1 * N_stride + na] = c[2]; if (vb) c_out[(size_t)h1 * N_stride + nbb] = c[3]; }
        }
        if (mode == 0) {   // ones row at h == H
            if (h0 == H) { c[0] = va ? 1.f : 0.f; c[1] = vb ? 1.f : 0.f; }
            if (h1 == H) { c[2] = va ? 1.f : 0.f; c[3] = vb ? 1.f : 0.f; }
        }
        // aggregation MMA: A(m=h, k) with k=tq <-> node na, k=tq+4 <-> node nbb
        uint32_t ah[4], al[4];
        split_tf32<PREC>(c[0], ah[0], al[0]);   // (h0, na)
        split_tf32<PREC>(c[2], ah[1], al[1]);   // (h1, na)
        split_tf32<PREC>(c[1], ah[2], al[2]);   // (h0, nbb)
        split_tf32<PREC>(c[3], ah[3], al[3]);   // (h1, nbb)
        const float* pa = sm.Ps + (size_t)na * sm.ldp + g;
        const float* pb = pa + sm.ldp;
#pragma unroll
        for (int j = 0; j < NT; ++j) {
            uint32_t bh[2], bl2[2];
            split_tf32<PREC>(pa[8 * j], bh[0], bl2[0]);
            split_tf32<PREC>(pb[8 * j], bh[1], bl2[1]);
            mma_split<PREC>(acc[j], ah, al, bh, bl2);
        }
    }
    // cross-warp reduction (deterministic order): red[warp][h][d], h < HA
    float* red = sm.Wred;
#pragma unroll
    for (int j = 0; j < NT; ++j) {
        const int d = 8 * j + 2 * tq;
        if (h0 < HA) *reinterpret_cast<float2*>(red + ((size_t)warp * (H + 1) + h0) * D + d) = make_float2(acc[j][0], acc[j][1]);
        if (h1 < HA) *reinterpret_cast<float2*>(red + ((size_t)warp * (H + 1) + h1) * D + d) = make_float2(acc[j][2], acc[j][3]);
    }
    __syncthreads();
    float* part = sm.part + (size_t)(pass_idx & 1) * (H + 1) * D;
    for (int i = tid; i < HA * D; i += blockDim.x) {
        float s = 0.f;
#pragma unroll
        for (int w = 0; w < kWarps; ++w) s += red[(size_t)w * (H + 1) * D + i];
        part[i] = s;
    }
    if (CS > 1) {
        cluster.sync();
        for (int i = tid; i < HA * D; i += blockDim.x) {
            float s = 0.f;
            for (int r = 0; r < CS; ++r) s += cluster.map_shared_rank(part, r)[i];
            sm.tot[i] = s;
        }
    } else {
        __syncthreads();
        for (int i = tid; i < HA * D; i += blockDim.x) sm.tot[i] = part[i];
    }
    __syncthreads();
}

template <int D, int PREC>
__global__ void __launch_bounds__(256, 2) cap_route_fwd_kernel(const float* __restrict__ x, const float* __restrict__ Wp,
                                                               const float* __restrict__ bp, const float* __restrict__ dadj,
                                                               float* __restrict__ c_out, float* __restrict__ s_out, int N,
                                                               int H, int R, int CS, int RPC) {
    cg::cluster_group cluster = cg::this_cluster();
    extern __shared__ __align__(16) float smem[];
    RF sm;
    sm.ldp = D + 4; sm.ldv = D + 4;
    float* p = smem;
    sm.Ps = p; p += (size_t)RPC * sm.ldp;
    {
        size_t wred = (size_t)D * (D + 4), red = (size_t)kWarps * (H + 1) * D;
        sm.Wred = p; p += (red > wred ? red : wred);
    }
    sm.bps = p; p += D;
    sm.dadj = p; p += (size_t)H * RPC;
    sm.bl = p; p += (size_t)H * RPC;
    sm.vh = reinterpret_cast<uint32_t*>(p); p += (size_t)kHP * sm.ldv;
    sm.vl = reinterpret_cast<uint32_t*>(p); p += (size_t)kHP * sm.ldv;
    sm.part = p; p += 2 * (size_t)(H + 1) * D;
    sm.tot = p; p += (size_t)(H + 1) * D;
    sm.u = p; p += (size_t)H * D;

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int slab = blockIdx.x / CS;
    const int rank = (CS > 1) ? (int)cluster.block_rank() : 0;
    const int n0 = rank * RPC;
    int nloc = N - n0; nloc = nloc < 0 ? 0 : (nloc > RPC ? RPC : nloc);
    const float* xs = x + ((size_t)slab * N + n0) * D;
    constexpr int LDW = D + 4;
    float* Wps = sm.Wred;

    for (int i = tid; i < D * D / 4; i += 256) {
        int o = (i * 4) / D, k = (i * 4) % D;
        *reinterpret_cast<float4*>(Wps + o * LDW + k) = *reinterpret_cast<const float4*>(Wp + (size_t)i * 4);
    }
    for (int i = tid; i < D; i += 256) sm.bps[i] = bp[i];
    for (int i = tid; i < RPC * (D / 4); i += 256) {
        int r = i / (D / 4), cc = (i % (D / 4)) * 4;
        float4 v4 = make_float4(0.f, 0.f, 0.f, 0.f);
        if (r < nloc) v4 = *reinterpret_cast<const float4*>(xs + (size_t)r * D + cc);
        *reinterpret_cast<float4*>(sm.Ps + (size_t)r * sm.ldp + cc) = v4;
    }
    for (int i = tid; i < H * RPC; i += 256) {
        int h = i / RPC, n = i % RPC;
        sm.dadj[i] = (n < nloc) ? dadj[((size_t)slab * H + h) * N + n0 + n] : 0.f;
        sm.bl[i] = 0.f;
    }
    __syncthreads();
    // ---- P = squash(x Wp^T + bp), in place
    for (int mt = warp; mt < RPC / 16; mt += kWarps) {
        float acc[D / 8][4];
        float q0, q1;
        ztile<D, PREC>(sm.Ps, sm.ldp, Wps, LDW, sm.bps, mt, lane, acc, q0, q1);
        const int gq = lane >> 2, tq = lane & 3;
        const int r0 = mt * 16 + gq, r1 = r0 + 8;
        const float f0 = (r0 < nloc) ? squash_f(q0) : 0.f, f1 = (r1 < nloc) ? squash_f(q1) : 0.f;
        __syncwarp();
#pragma unroll
        for (int nt = 0; nt < D / 8; ++nt) {
            const int cc = nt * 8 + 2 * tq;
            *reinterpret_cast<float2*>(sm.Ps + (size_t)r0 * sm.ldp + cc) = make_float2(acc[nt][0] * f0, acc[nt][1] * f0);
            *reinterpret_cast<float2*>(sm.Ps + (size_t)r1 * sm.ldp + cc) = make_float2(acc[nt][2] * f1, acc[nt][3] * f1);
        }
    }
    __syncthreads();   // P complete; the Wp tile is dead from here on (its storage becomes `red`)
    float* cg_out = c_out + (size_t)slab * H * N + n0;
    int pass = 0;
    route_pass_mma<D, PREC>(sm, 0, false, H, RPC, nloc, pass++, cg_out, N, cluster, CS);
    for (int i = tid; i < H * D; i += 256) sm.u[i] = sm.tot[i];
    __syncthreads();
    squash_rows<D>(sm.u, H, warp, lane);
    __syncthreads();
    float* vtmp = sm.part + (size_t)(pass & 1) * (H + 1) * D;   // scratch: the part buffer the NEXT pass will overwrite
    if (R >= 1) {
        const float invH = 1.f / (float)H;
        for (int i = tid; i < H * D; i += 256) vtmp[i] = sm.u[i] * (sm.tot[H * D + (i % D)] * invH);
        __syncthreads();
        squash_rows<D>(vtmp, H, warp, lane);
        __syncthreads();
        load_v_planes<D, PREC>(sm, vtmp, H);
        __syncthreads();
        for (int it = 2; it <= R; ++it) {
            route_pass_mma<D, PREC>(sm, 1, true, H, RPC, nloc, pass++, cg_out, N, cluster, CS);
            vtmp = sm.part + (size_t)(pass & 1) * (H + 1) * D;
            for (int i = tid; i < H * D; i += 256) vtmp[i] = sm.u[i] * sm.tot[i];
            __syncthreads();
            squash_rows<D>(vtmp, H, warp, lane);
            __syncthreads();
            load_v_planes<D, PREC>(sm, vtmp, H);
            __syncthreads();
        }
    }
    route_pass_mma<D, PREC>(sm, 2, R >= 1, H, RPC, nloc, pass++, cg_out, N, cluster, CS);
    if (rank == 0) {
        float* so = s_out + (size_t)slab * H * D;
        for (int i = tid; i < H * D; i += 256) so[i] = sm.tot[i];
    }
    if (CS > 1) cluster.sync();  // peers may still be reading this CTA's `part` through DSMEM
}

static int min_cluster() {
    static int v = -1;
    if (v < 0) {
        const char* e = getenv("GPTST_B200_ROUTE_CLUSTER");   // tuning knob: smallest cluster size to consider
        v = e ? atoi(e) : 1;   // measured on B200 (N=170, D=64): cluster 1 -> 177 us, 2 -> 226 us, 4 -> 294 us for cap forward
        if (v < 1) v = 1;
    }
    return v;
}

static int pick_cluster(int N, int D, int H, int* rpc_out) {
    static const int sizes[5] = {1, 2, 4, 8, 16};
    for (int i = 0; i < 5; ++i) {
        int cs = sizes[i];
        if (cs < min_cluster() && N > 32 * cs) continue;   // more, smaller CTAs per slab: more warps in flight per SM
        int rpc = (N + cs - 1) / cs;
        rpc = (rpc + 15) / 16 * 16;
        if (rf_smem_floats(D, H, rpc) * 4 <= kSmemMax) {
            *rpc_out = rpc;
            return cs;
        }
    }
    return -1;
}

template <int D, int PREC>
static cudaError_t launch_route_fwd(const float* x, const float* Wp, const float* bp, const float* dadj, float* c,
                                    float* s, int BT, int N, int H, int R, cudaStream_t st) {
    int rpc = 0;
    int cs = pick_cluster(N, D, H, &rpc);
    if (cs < 0) return cudaErrorInvalidValue;
    size_t smem = rf_smem_floats(D, H, rpc) * 4;
    auto kern = cap_route_fwd_kernel<D, PREC>;
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return e;
    if (cs > 8) {
        e = cudaFuncSetAttribute(kern, cudaFuncAttributeNonPortableClusterSizeAllowed, 1);
        if (e != cudaSuccess) return e;
    }
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3((unsigned)BT * cs);
    cfg.blockDim = dim3(256);
    cfg.dynamicSmemBytes = smem;
    cfg.stream = st;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = cs;
    attr[0].val.clusterDim.y = 1;
    attr[0].val.clusterDim.z = 1;
    cfg.attrs = attr;
    cfg.numAttrs = 1;
    return cudaLaunchKernelEx(&cfg, kern, x, Wp, bp, dadj, c, s, N, H, R, cs, rpc);
}

// second generation (cap_route2_fwd.cu): D = 64, N <= 256, fp16-split tensor-core routing
bool route2_supported(int N, int D, int H);
cudaError_t route2_fwd(const float* x, const float* Wp, const float* bp, const float* dadj, float* c, float* s, float* z, int BT,
                       int N, int H, int R, int prec, cudaStream_t st);

}  // namespace gptst

using namespace gptst;

extern "C" int gptst_cap_route_fwd(const float* x, const float* Wp, const float* bp, const float* dadj, float* c,
                                   float* s, int B, int T, int N, int D, int H, int R, int prec, void* stream) {
    if (!x || !Wp || !bp || !dadj || !c || !s || B <= 0 || T <= 0 || N <= 0 || R < 0) return -1;
    if (H < 1 || H > 15) return -2;
    cudaStream_t st = (cudaStream_t)stream;
    if ((prec == 1 || prec == 3) && route2_supported(N, D, H)) return (int)route2_fwd(x, Wp, bp, dadj, c, s, nullptr, B * T, N, H, R, prec, st);
    if (D == 64 && prec == 1) return (int)launch_route_fwd<64, 1>(x, Wp, bp, dadj, c, s, B * T, N, H, R, st);
    if (D == 64 && prec == 3) return (int)launch_route_fwd<64, 3>(x, Wp, bp, dadj, c, s, B * T, N, H, R, st);
    if (D == 128 && prec == 1) return (int)launch_route_fwd<128, 1>(x, Wp, bp, dadj, c, s, B * T, N, H, R, st);
    if (D == 128 && prec == 3) return (int)launch_route_fwd<128, 3>(x, Wp, bp, dadj, c, s, B * T, N, H, R, st);
    return -2;
}

// Training flavour of the second-generation routing: additionally stores Z = x Wp^T + bp (B,T,N,D) for gptst_cap_route_bwd_dz_z.
// Only where gptst_cap_route2_supported(N, D, H); returns -2 otherwise (callers then use the recomputing pair).
extern "C" int gptst_cap_route_fwd_z(const float* x, const float* Wp, const float* bp, const float* dadj, float* c, float* s,
                                     float* z, int B, int T, int N, int D, int H, int R, int prec, void* stream) {
    if (!x || !Wp || !bp || !dadj || !c || !s || !z || B <= 0 || T <= 0 || N <= 0 || R < 0) return -1;
    if (!((prec == 1 || prec == 3) && route2_supported(N, D, H))) return -2;
    return (int)route2_fwd(x, Wp, bp, dadj, c, s, z, B * T, N, H, R, prec, (cudaStream_t)stream);
}
