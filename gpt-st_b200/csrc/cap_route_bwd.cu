// Hierarchical intra/inter-cluster hypergraph block `cap` (reference GPTST.py:100-141), forward + backward.
//
// Forward, per (b,t) slab of x (N x D):
//   route_fwd : Z = x Wp^T + bp ; P = squash(Z) ; dynamic routing on P (R iterations) with the
//               data-dependent incidence logits dadj ; c = softmax_H(b + dadj) (H x N) ; s = c P (H x D)
//   hop_fwd   : per sample b, inter-cluster hop over k = (t,h):  r = LReLU(dyn^T LReLU(dyn (s + tau))) + s ; v = squash(r)
//   recon     : recon = c^T v (N x D)      (the node-adaptive projection + residual is gproj.cu, group = node)
// Backward (routing logits are constants of the graph, GPTST.py:108-109):
//   dv_dcr    : dv = c drecon (H x D) ; dc_r = v drecon^T (H x N)
//   hop_bwd   : dv -> ds (incl. the direct path through r = ... + s), ddyn
//   route_bwd : dc = dc_r + ds P^T ; dL = c*(dc - sum_h c dc) -> ddadj ; dP = c^T ds ; dZ = squash'(Z, dP) ;
//               dx = dy + dZ Wp ; dWp = dZ^T x ; dbp = sum dZ
//
// The N-reductions of route_fwd are done with lanes over nodes / lanes over D and a deterministic
// cross-warp + cross-CTA (thread-block cluster, DSMEM) tree; the D x D contractions use tensor cores.
#include "cap_common.cuh"

namespace gptst {

// route backward.  grid = (row chunks, slab CTAs); each CTA walks slabs blockIdx.y, +gridDim.y, ... and keeps
// its dWp / dbp partial in registers.  dx_io holds dy = dOut*act'(out) on entry and receives dy + dZ Wp.
// ------------------------------------------------------------------------------------------------------
__host__ __device__ inline size_t route_bwd_smem_floats(int D, int H, int RPC) {
    return (size_t)2 * RPC * (D + 4) + (size_t)D * (D + 4) + D + (size_t)H * RPC + (size_t)H * D + 3 * (size_t)RPC + 256;
}

template <int D, int PREC, int HP>
__global__ void __launch_bounds__(256) cap_route_bwd_kernel(
    const float* __restrict__ x, const float* __restrict__ Wp, const float* __restrict__ bp, const float* __restrict__ c,
    const float* __restrict__ ds, const float* __restrict__ dcr, float* __restrict__ dx_io, float* __restrict__ ddadj,
    float* __restrict__ dWp_part, float* __restrict__ dbp_part, int nslab, int N, int H, int RPC) {
    extern __shared__ __align__(16) float smem[];
    constexpr int LD = D + 4;
    float* Xs = smem;                          // [RPC][LD]
    float* Zs = Xs + (size_t)RPC * LD;         // [RPC][LD]  Z, then dZ
    float* Wps = Zs + (size_t)RPC * LD;        // [D][LD]    Wps[o][i]
    float* bps = Wps + (size_t)D * LD;         // [D]
    float* cs = bps + D;                       // [H][RPC]
    float* dss = cs + (size_t)H * RPC;         // [H][D]
    float* fq = dss + (size_t)H * D;           // [RPC] f(q)
    float* fpq = fq + RPC;                     // [RPC] f'(q)
    float* zd = fpq + RPC;                     // [RPC] sum_h c[h,n] (ds_h . Z_n)
    float* red = zd + RPC;                     // [256]

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int n0 = blockIdx.x * RPC;
    int nloc = N - n0; nloc = nloc > RPC ? RPC : nloc;

    for (int i = tid; i < D * D / 4; i += 256) {
        int o = (i * 4) / D, k = (i * 4) % D;
        *reinterpret_cast<float4*>(Wps + o * LD + k) = *reinterpret_cast<const float4*>(Wp + (size_t)i * 4);
    }
    for (int i = tid; i < D; i += 256) bps[i] = bp[i];

    // dWp accumulators: dWp[o][i], M = o, N = i
    constexpr int MT = D / 16, NTT = D / 8;
    constexpr int WMG = (MT >= 8) ? 8 : MT, WNG = 8 / WMG, NT_W = NTT / WNG;
    const int gm = warp % WMG, gn = warp / WMG;
    float gacc[NT_W][4];
#pragma unroll
    for (int nt = 0; nt < NT_W; ++nt) gacc[nt][0] = gacc[nt][1] = gacc[nt][2] = gacc[nt][3] = 0.f;
    float sigma = 0.f;

    for (int slab = blockIdx.y; slab < nslab; slab += gridDim.y) {
        __syncthreads();
        const float* xs = x + ((size_t)slab * N + n0) * D;
        for (int i = tid; i < RPC * (D / 4); i += 256) {
            int r = i / (D / 4), cc = (i % (D / 4)) * 4;
            float4 v4 = make_float4(0.f, 0.f, 0.f, 0.f);
            if (r < nloc) v4 = *reinterpret_cast<const float4*>(xs + (size_t)r * D + cc);
            *reinterpret_cast<float4*>(Xs + (size_t)r * LD + cc) = v4;
        }
        for (int i = tid; i < H * RPC; i += 256) {
            int h = i / RPC, n = i % RPC;
            cs[i] = (n < nloc) ? c[((size_t)slab * H + h) * N + n0 + n] : 0.f;
        }
        for (int i = tid; i < H * D; i += 256) dss[i] = ds[(size_t)slab * H * D + i];
        __syncthreads();
        // ---- Z = x Wp^T + bp -> Zs ; row stats
        for (int mt = warp; mt < RPC / 16; mt += kWarps) {
            float acc[D / 8][4];
            float q0, q1;
            ztile<D, PREC>(Xs, LD, Wps, LD, bps, mt, lane, acc, q0, q1);
            const int gq = lane >> 2, tq = lane & 3;
            const int r0 = mt * 16 + gq, r1 = r0 + 8;
            const bool v0 = r0 < nloc, v1 = r1 < nloc;
#pragma unroll
            for (int nt = 0; nt < D / 8; ++nt) {
                const int cc = nt * 8 + 2 * tq;
                *reinterpret_cast<float2*>(Zs + (size_t)r0 * LD + cc) = v0 ? make_float2(acc[nt][0], acc[nt][1]) : make_float2(0.f, 0.f);
                *reinterpret_cast<float2*>(Zs + (size_t)r1 * LD + cc) = v1 ? make_float2(acc[nt][2], acc[nt][3]) : make_float2(0.f, 0.f);
            }
            if (tq == 0) {
                fq[r0] = v0 ? squash_f(q0) : 0.f; fpq[r0] = v0 ? squash_df(q0) : 0.f;
                fq[r1] = v1 ? squash_f(q1) : 0.f; fpq[r1] = v1 ? squash_df(q1) : 0.f;
            }
        }
        __syncthreads();
        // ---- per node: dsZ[h] = ds_h . Z_n ; dc, dL -> ddadj ; zd
        {
            const int nbatch = (nloc + kNPB - 1) / kNPB;
            for (int batch = warp; batch < nbatch; batch += kWarps) {
                const int nl = batch * kNPB + (lane / kLPN), q = lane % kLPN;
                const bool valid = nl < nloc;
                const int nrow = valid ? nl : 0;
                float dz[HP];
                node_dots<D, HP>(Zs + (size_t)nrow * LD, dss, H, q, dz);
                if (q == 0 && valid) {
                    const float f = fq[nrow];
                    float dc[HP], cdc = 0.f, zsum = 0.f;
#pragma unroll
                    for (int h = 0; h < HP; ++h) {
                        if (h < H) {
                            const float ch = cs[h * RPC + nrow];
                            dc[h] = dcr[((size_t)slab * H + h) * N + n0 + nl] + f * dz[h];
                            cdc = fmaf(ch, dc[h], cdc);
                            zsum = fmaf(ch, dz[h], zsum);
                        }
                    }
#pragma unroll
                    for (int h = 0; h < HP; ++h)
                        if (h < H) ddadj[((size_t)slab * H + h) * N + n0 + nl] = cs[h * RPC + nrow] * (dc[h] - cdc);
                    zd[nrow] = zsum;
                }
            }
        }
        __syncthreads();
        // ---- dZ = f dP + 2 f' zd Z  (in place), dP = c^T ds
        for (int r = warp; r < RPC; r += kWarps) {
            const float f = fq[r], g2 = 2.f * fpq[r] * ((r < nloc) ? zd[r] : 0.f);
            float ch[HP];
#pragma unroll
            for (int h = 0; h < HP; ++h) ch[h] = (h < H) ? cs[h * RPC + r] : 0.f;
#pragma unroll
            for (int j = 0; j < D / 32; ++j) {
                const int d = lane + 32 * j;
                float dP = 0.f;
#pragma unroll
                for (int h = 0; h < HP; ++h) if (h < H) dP = fmaf(ch[h], dss[h * D + d], dP);
                Zs[(size_t)r * LD + d] = f * dP + g2 * Zs[(size_t)r * LD + d];
            }
        }
        __syncthreads();
        // ---- dbp partial: column sums of dZ
        {
            constexpr int PARTS = 256 / D;
            const int cc = tid % D, part = tid / D;
            float sacc = 0.f;
            for (int r = part; r < RPC; r += PARTS) sacc += Zs[(size_t)r * LD + cc];
            red[part * D + cc] = sacc;
        }
        // ---- dx = dy + dZ Wp       B(k=o, n=i) = Wps[k][n]
        float* dxs = dx_io + ((size_t)slab * N + n0) * D;
        for (int mt = warp; mt < RPC / 16; mt += kWarps) {
            float acc[D / 8][4];
#pragma unroll
            for (int nt = 0; nt < D / 8; ++nt) acc[nt][0] = acc[nt][1] = acc[nt][2] = acc[nt][3] = 0.f;
            warp_gemm<D, D / 8, PREC, false, true>(acc, Zs + (size_t)mt * 16 * LD, LD, Wps, LD, lane);
            const int gq = lane >> 2, tq = lane & 3;
#pragma unroll
            for (int half = 0; half < 2; ++half) {
                const int r = mt * 16 + gq + half * 8;
                if (r < nloc) {
#pragma unroll
                    for (int nt = 0; nt < D / 8; ++nt) {
                        float2* pp = reinterpret_cast<float2*>(dxs + (size_t)r * D + nt * 8 + 2 * tq);
                        float2 old = *pp;
                        *pp = make_float2(old.x + acc[nt][half * 2], old.y + acc[nt][half * 2 + 1]);
                    }
                }
            }
        }
        // ---- dWp += dZ^T x        A(m=o,k=row) = Zs[row][o],  B(k=row,n=i) = Xs[row][i]
        warp_gemm_rt<0, NT_W, PREC, true, true>(gacc, Zs + gm * 16, LD, Xs + gn * NT_W * 8, LD, lane, RPC);
        __syncthreads();
        if (tid < D) {
            constexpr int PARTS = 256 / D;
#pragma unroll
            for (int p2 = 0; p2 < PARTS; ++p2) sigma += red[p2 * D + tid];
        }
    }
    const size_t pidx = (size_t)blockIdx.y * gridDim.x + blockIdx.x;
    float* dWo = dWp_part + pidx * D * D;
    const int gq = lane >> 2, tq = lane & 3;
#pragma unroll
    for (int nt = 0; nt < NT_W; ++nt) {
        const int cc = (gn * NT_W + nt) * 8 + 2 * tq;
        const int r = gm * 16 + gq;
        *reinterpret_cast<float2*>(dWo + (size_t)r * D + cc) = make_float2(gacc[nt][0], gacc[nt][1]);
        *reinterpret_cast<float2*>(dWo + (size_t)(r + 8) * D + cc) = make_float2(gacc[nt][2], gacc[nt][3]);
    }
    if (tid < D) dbp_part[pidx * D + tid] = sigma;
}

// ------------------------------------------------------------------------------------------------------
// host side
// ------------------------------------------------------------------------------------------------------
static void route_bwd_geometry(int BT, int N, int D, int H, int* rpc_out, int* chunks_out, int* slab_ctas_out) {
    // largest row chunk (multiple of 16) that fits; then enough slab-CTAs to cover the machine ~2x
    int rpc = (N + 15) / 16 * 16;
    while (route_bwd_smem_floats(D, H, rpc) * 4 > kSmemMax && rpc > 16) rpc -= 16;
    // prefer two CTAs per SM when a half-size chunk still amortises the Wp tile
    if (route_bwd_smem_floats(D, H, rpc) * 4 > kSmemMax / 2) {
        int half = ((N + 1) / 2 + 15) / 16 * 16;
        if (half >= 64 && half < rpc) rpc = half;
    }
    int chunks = (N + rpc - 1) / rpc;
    int ctas = (296 + chunks - 1) / chunks;
    if (ctas > BT) ctas = BT;
    if (ctas < 1) ctas = 1;
    *rpc_out = rpc; *chunks_out = chunks; *slab_ctas_out = ctas;
}

template <int D, int PREC, int HP>
static cudaError_t launch_route_bwd(const float* x, const float* Wp, const float* bp, const float* c, const float* ds,
                                    const float* dcr, float* dx_io, float* ddadj, float* dWp_part, float* dbp_part,
                                    int BT, int N, int H, cudaStream_t st) {
    int rpc, chunks, ctas;
    route_bwd_geometry(BT, N, D, H, &rpc, &chunks, &ctas);
    size_t smem = route_bwd_smem_floats(D, H, rpc) * 4;
    if (smem > kSmemMax) return cudaErrorInvalidValue;
    auto kern = cap_route_bwd_kernel<D, PREC, HP>;
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return e;
    kern<<<dim3(chunks, ctas), 256, smem, st>>>(x, Wp, bp, c, ds, dcr, dx_io, ddadj, dWp_part, dbp_part, BT, N, H, rpc);
    return cudaGetLastError();
}

}  // namespace gptst

using namespace gptst;

#define CAP_DISPATCH(D_, P_, H_, CALL)                                              \
    do {                                                                            \
        if ((H_) == 10) {                                                           \
            if ((D_) == 64 && (P_) == 1) { CALL(64, 1, 10); }                       \
            else if ((D_) == 64 && (P_) == 3) { CALL(64, 3, 10); }                  \
            else if ((D_) == 128 && (P_) == 1) { CALL(128, 1, 10); }                \
            else if ((D_) == 128 && (P_) == 3) { CALL(128, 3, 10); }                \
            else return -2;                                                         \
        } else if ((H_) >= 1 && (H_) <= 16) {                                       \
            if ((D_) == 64 && (P_) == 1) { CALL(64, 1, 16); }                       \
            else if ((D_) == 64 && (P_) == 3) { CALL(64, 3, 16); }                  \
            else if ((D_) == 128 && (P_) == 1) { CALL(128, 1, 16); }                \
            else if ((D_) == 128 && (P_) == 3) { CALL(128, 3, 16); }                \
            else return -2;                                                         \
        } else return -2;                                                           \
    } while (0)

extern "C" int gptst_cap_route_bwd_parts(int B, int T, int N, int D, int H) {
    int rpc, chunks, ctas;
    route_bwd_geometry(B * T, N, D, H, &rpc, &chunks, &ctas);
    return chunks * ctas;
}

extern "C" int gptst_cap_route_bwd(const float* x, const float* Wp, const float* bp, const float* c, const float* ds,
                                   const float* dcr, float* dx_io, float* ddadj, float* dWp_part, float* dbp_part, int B,
                                   int T, int N, int D, int H, int prec, void* stream) {
    if (!x || !Wp || !bp || !c || !ds || !dcr || !dx_io || !ddadj || !dWp_part || !dbp_part || B <= 0 || N <= 0) return -1;
    cudaStream_t st = (cudaStream_t)stream;
#define CALL(DD, PP, HH) \
    return (int)launch_route_bwd<DD, PP, HH>(x, Wp, bp, c, ds, dcr, dx_io, ddadj, dWp_part, dbp_part, B * T, N, H, st)
    CAP_DISPATCH(D, prec, H, CALL);
#undef CALL
    return -2;
}

