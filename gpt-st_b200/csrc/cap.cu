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
#include <cooperative_groups.h>

#include "common.cuh"

namespace cg = cooperative_groups;

namespace gptst {

constexpr int kWarps = 8;
constexpr int kLPN = 4;               // lanes per node in the dot-product phases
constexpr int kNPB = 32 / kLPN;       // nodes per warp batch
constexpr int kCW = 20;               // padded row length of the per-warp c staging (>= H+1, multiple of 4)

template <int K, int NT, int PREC, bool A_TRANS, bool B_KMAJOR>
__device__ __forceinline__ void warp_gemm_rt(float (&acc)[NT][4], const float* __restrict__ As, int lda,
                                             const float* __restrict__ Bs, int ldb, int lane, int Krt) {
    const int g = lane >> 2, t = lane & 3;
    for (int k0 = 0; k0 < Krt; k0 += 8) {
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

// Z = x Wp^T + bp for the 16-row tile `mt` (in place in Xs when Zs == Xs); returns per-row |Z|^2 for this
// lane's two rows (g and g+8), already reduced over the quad.
template <int D, int PREC>
__device__ __forceinline__ void ztile(const float* __restrict__ Xs, int ldx, const float* __restrict__ Wps, int ldw,
                                      const float* __restrict__ bps, int mt, int lane, float (&acc)[D / 8][4],
                                      float& q0, float& q1) {
#pragma unroll
    for (int nt = 0; nt < D / 8; ++nt) acc[nt][0] = acc[nt][1] = acc[nt][2] = acc[nt][3] = 0.f;
    warp_gemm<D, D / 8, PREC, false, false>(acc, Xs + mt * 16 * ldx, ldx, Wps, ldw, lane);
    const int tq = lane & 3;
    q0 = q1 = 0.f;
#pragma unroll
    for (int nt = 0; nt < D / 8; ++nt) {
        const float b0 = bps[nt * 8 + 2 * tq], b1 = bps[nt * 8 + 2 * tq + 1];
        acc[nt][0] += b0; acc[nt][1] += b1; acc[nt][2] += b0; acc[nt][3] += b1;
        q0 += acc[nt][0] * acc[nt][0] + acc[nt][1] * acc[nt][1];
        q1 += acc[nt][2] * acc[nt][2] + acc[nt][3] * acc[nt][3];
    }
    q0 += __shfl_xor_sync(0xffffffffu, q0, 1); q0 += __shfl_xor_sync(0xffffffffu, q0, 2);
    q1 += __shfl_xor_sync(0xffffffffu, q1, 1); q1 += __shfl_xor_sync(0xffffffffu, q1, 2);
}

// dots[h] = sum_d V[h][d] * Row[d] for one node handled by kLPN lanes (lane part `q` takes float4 index j = q mod kLPN)
template <int D, int HP>
__device__ __forceinline__ void node_dots(const float* __restrict__ row, const float* __restrict__ V, int H, int q,
                                          float (&dots)[HP]) {
#pragma unroll
    for (int h = 0; h < HP; ++h) dots[h] = 0.f;
#pragma unroll
    for (int jj = 0; jj < D / 4 / kLPN; ++jj) {
        const int j = jj * kLPN + q;
        const float4 p = *reinterpret_cast<const float4*>(row + 4 * j);
#pragma unroll
        for (int h = 0; h < HP; ++h) {
            if (h < H) {
                const float4 v = *reinterpret_cast<const float4*>(V + h * D + 4 * j);
                dots[h] = fmaf(p.x, v.x, fmaf(p.y, v.y, fmaf(p.z, v.z, fmaf(p.w, v.w, dots[h]))));
            }
        }
    }
#pragma unroll
    for (int h = 0; h < HP; ++h) {
        dots[h] += __shfl_xor_sync(0xffffffffu, dots[h], 1);
        dots[h] += __shfl_xor_sync(0xffffffffu, dots[h], 2);
    }
}

template <int HP>
__device__ __forceinline__ void softmax_h(float (&z)[HP], int H) {
    float m = -INFINITY;
#pragma unroll
    for (int h = 0; h < HP; ++h) if (h < H) m = fmaxf(m, z[h]);
    float s = 0.f;
#pragma unroll
    for (int h = 0; h < HP; ++h) {
        if (h < H) { z[h] = expf(z[h] - m); s += z[h]; } else z[h] = 0.f;
    }
    const float inv = 1.f / s;
#pragma unroll
    for (int h = 0; h < HP; ++h) z[h] *= inv;
}

// row-wise squash of an (rows x D) smem matrix in place, one warp per row
template <int D>
__device__ __forceinline__ void squash_rows(float* M, int rows, int warp, int lane) {
    for (int r = warp; r < rows; r += kWarps) {
        float q = 0.f;
        for (int d = lane; d < D; d += 32) { float v = M[r * D + d]; q += v * v; }
        q = warp_sum(q);
        const float f = squash_f(q);
        for (int d = lane; d < D; d += 32) M[r * D + d] *= f;
    }
}

struct RouteSmem {
    float *Ps, *Wps, *bps, *dadj, *bl, *cw, *red, *part, *tot, *u, *v;
    int ldp, ldw;
};

__host__ __device__ inline size_t route_fwd_smem_floats(int D, int H, int RPC) {
    size_t n = 0;
    n += (size_t)RPC * (D + 4);            // Ps
    n += (size_t)D * (D + 4);              // Wps
    n += D;                                // bps
    n += 2 * (size_t)H * RPC;              // dadj, bl
    n += (size_t)kWarps * kNPB * kCW;      // cw
    n += (size_t)kWarps * (H + 1) * D;     // red
    n += 2 * (size_t)(H + 1) * D;          // part (double buffered)
    n += (size_t)(H + 1) * D;              // tot
    n += 2 * (size_t)H * D;                // u, v
    return n;
}

// One pass over the CTA's nodes:  (optional) logits += V . P ;  c = softmax_H(logits (+ dadj)) ;
// acc[hh] += c[hh] * P  (hh < HA, row H == all-ones when HA == H+1) ; deterministic reduction to sm.tot.
// mode 0: first pass   c = softmax(dadj), extra ones-row (uniform routing iteration), logits untouched
// mode 1: middle pass  bl += v.P ; c = softmax(bl)
// mode 2: final pass   bl += v.P (if use_v) ; c = softmax(bl + dadj) ; c written to global
template <int D, int HP>
__device__ void route_pass(const RouteSmem& sm, int mode, bool use_v, int H, int RPC, int nloc, int pass_idx,
                           float* __restrict__ c_out /* c[b,t,:,n0:] with stride N */, int N_stride, cg::cluster_group& cluster,
                           int CS) {
    constexpr int VEC = D / 32;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int HA = (mode == 0) ? H + 1 : H;
    float acc[HP + 1][VEC];
#pragma unroll
    for (int h = 0; h <= HP; ++h)
#pragma unroll
        for (int v = 0; v < VEC; ++v) acc[h][v] = 0.f;
    float* cw = sm.cw + warp * kNPB * kCW;
    const int nbatch = (nloc + kNPB - 1) / kNPB;
    for (int batch = warp; batch < nbatch; batch += kWarps) {
        const int nl = batch * kNPB + (lane / kLPN), q = lane % kLPN;
        const bool valid = nl < nloc;
        const int nrow = valid ? nl : 0;
        float z[HP];
        if (mode != 0 && use_v) node_dots<D, HP>(sm.Ps + (size_t)nrow * sm.ldp, sm.v, H, q, z);
        else {
#pragma unroll
            for (int h = 0; h < HP; ++h) z[h] = 0.f;
        }
        if (mode != 0) {
#pragma unroll
            for (int h = 0; h < HP; ++h) if (h < H) z[h] += sm.bl[h * RPC + nrow];
            __syncwarp();
            if (mode == 1 && valid && q == 0) {
#pragma unroll
                for (int h = 0; h < HP; ++h) if (h < H) sm.bl[h * RPC + nrow] = z[h];
            }
        }
        if (mode != 1) {
#pragma unroll
            for (int h = 0; h < HP; ++h) if (h < H) z[h] += sm.dadj[h * RPC + nrow];
        }
        softmax_h<HP>(z, H);
        if (q == 0) {
#pragma unroll
            for (int h = 0; h < HP; ++h) if (h < H) cw[(lane / kLPN) * kCW + h] = valid ? z[h] : 0.f;
            if (mode == 0) cw[(lane / kLPN) * kCW + H] = valid ? 1.f : 0.f;
            if (mode == 2 && valid) {
#pragma unroll
                for (int h = 0; h < HP; ++h) if (h < H) c_out[(size_t)h * N_stride + nl] = z[h];
            }
        }
        __syncwarp();
        // accumulate: lanes over D
#pragma unroll
        for (int i = 0; i < kNPB; ++i) {
            const int n2 = batch * kNPB + i;
            if (n2 < nloc) {
                float p[VEC];
                if (VEC == 4) {
                    float4 t4 = *reinterpret_cast<const float4*>(sm.Ps + (size_t)n2 * sm.ldp + lane * 4);
                    p[0] = t4.x; p[1] = t4.y; p[VEC > 2 ? 2 : 0] = t4.z; p[VEC > 3 ? 3 : 0] = t4.w;
                } else if (VEC == 2) {
                    float2 t2 = *reinterpret_cast<const float2*>(sm.Ps + (size_t)n2 * sm.ldp + lane * 2);
                    p[0] = t2.x; p[VEC > 1 ? 1 : 0] = t2.y;
                } else {
                    p[0] = sm.Ps[(size_t)n2 * sm.ldp + lane];
                }
#pragma unroll
                for (int h = 0; h <= HP; ++h) {
                    if (h < HA) {
                        const float cv = cw[i * kCW + h];
#pragma unroll
                        for (int v = 0; v < VEC; ++v) acc[h][v] = fmaf(cv, p[v], acc[h][v]);
                    }
                }
            }
        }
        __syncwarp();
    }
    // cross-warp (deterministic order) -> part[pass parity]
#pragma unroll
    for (int h = 0; h <= HP; ++h)
        if (h < HA)
#pragma unroll
            for (int v = 0; v < VEC; ++v) sm.red[((size_t)warp * (H + 1) + h) * D + lane * VEC + v] = acc[h][v];
    __syncthreads();
    float* part = sm.part + (size_t)(pass_idx & 1) * (H + 1) * D;
    for (int i = tid; i < HA * D; i += blockDim.x) {
        float s = 0.f;
#pragma unroll
        for (int w = 0; w < kWarps; ++w) s += sm.red[(size_t)w * (H + 1) * D + i];
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

template <int D, int PREC, int HP>
__global__ void __launch_bounds__(256) cap_route_fwd_kernel(const float* __restrict__ x, const float* __restrict__ Wp,
                                                            const float* __restrict__ bp, const float* __restrict__ dadj,
                                                            float* __restrict__ c_out, float* __restrict__ s_out, int N,
                                                            int H, int R, int CS, int RPC) {
    cg::cluster_group cluster = cg::this_cluster();
    extern __shared__ __align__(16) float smem[];
    RouteSmem sm;
    sm.ldp = D + 4; sm.ldw = D + 4;
    float* p = smem;
    sm.Ps = p; p += (size_t)RPC * sm.ldp;
    sm.Wps = p; p += (size_t)D * sm.ldw;
    sm.bps = p; p += D;
    sm.dadj = p; p += (size_t)H * RPC;
    sm.bl = p; p += (size_t)H * RPC;
    sm.cw = p; p += (size_t)kWarps * kNPB * kCW;
    sm.red = p; p += (size_t)kWarps * (H + 1) * D;
    sm.part = p; p += 2 * (size_t)(H + 1) * D;
    sm.tot = p; p += (size_t)(H + 1) * D;
    sm.u = p; p += (size_t)H * D;
    sm.v = p; p += (size_t)H * D;

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int slab = blockIdx.x / CS;
    const int rank = (CS > 1) ? (int)cluster.block_rank() : 0;
    const int n0 = rank * RPC;
    int nloc = N - n0; nloc = nloc < 0 ? 0 : (nloc > RPC ? RPC : nloc);
    const float* xs = x + ((size_t)slab * N + n0) * D;

    for (int i = tid; i < D * D / 4; i += 256) {
        int o = (i * 4) / D, k = (i * 4) % D;
        *reinterpret_cast<float4*>(sm.Wps + o * sm.ldw + k) = *reinterpret_cast<const float4*>(Wp + (size_t)i * 4);
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
        ztile<D, PREC>(sm.Ps, sm.ldp, sm.Wps, sm.ldw, sm.bps, mt, lane, acc, q0, q1);
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
    __syncthreads();
    float* cg_out = c_out + (size_t)slab * H * N + n0;
    int pass = 0;
    // ---- pass A: u = squash(softmax_H(dadj) . P);  first routing iteration has uniform coupling 1/H
    route_pass<D, HP>(sm, 0, false, H, RPC, nloc, pass++, cg_out, N, cluster, CS);
    for (int i = tid; i < H * D; i += 256) sm.u[i] = sm.tot[i];
    __syncthreads();
    squash_rows<D>(sm.u, H, warp, lane);
    __syncthreads();
    if (R >= 1) {
        const float invH = 1.f / (float)H;
        for (int i = tid; i < H * D; i += 256) sm.v[i] = sm.u[i] * (sm.tot[H * D + (i % D)] * invH);
        __syncthreads();
        squash_rows<D>(sm.v, H, warp, lane);
        __syncthreads();
        for (int it = 2; it <= R; ++it) {
            route_pass<D, HP>(sm, 1, true, H, RPC, nloc, pass++, cg_out, N, cluster, CS);
            for (int i = tid; i < H * D; i += 256) sm.v[i] = sm.u[i] * sm.tot[i];
            __syncthreads();
            squash_rows<D>(sm.v, H, warp, lane);
            __syncthreads();
        }
    }
    // ---- final pass: c = softmax_H(b + dadj), s = c . P
    route_pass<D, HP>(sm, 2, R >= 1, H, RPC, nloc, pass++, cg_out, N, cluster, CS);
    if (rank == 0) {
        float* so = s_out + (size_t)slab * H * D;
        for (int i = tid; i < H * D; i += 256) so[i] = sm.tot[i];
    }
    if (CS > 1) cluster.sync();  // peers may still be reading this CTA's `part` through DSMEM
}

// ------------------------------------------------------------------------------------------------------
// inter-cluster hop (per sample b), GPTST.py:125-134
// ------------------------------------------------------------------------------------------------------
template <int D>
__global__ void __launch_bounds__(256) cap_hop_fwd_kernel(const float* __restrict__ s, const float* __restrict__ dyn,
                                                          float* __restrict__ v, int T, int H, int HT) {
    extern __shared__ __align__(16) float smem[];
    const int K = T * H, LD = D + 1;
    float* Ss = smem;                 // [K][LD]  raw s
    float* E1 = Ss + (size_t)K * LD;  // [HT][LD]
    float* dy = E1 + (size_t)HT * LD; // [HT][K+1]
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int b = blockIdx.x;
    const float* sb = s + (size_t)b * K * D;
    for (int i = tid; i < K * D; i += 256) Ss[(i / D) * LD + (i % D)] = sb[i];
    for (int i = tid; i < HT * K; i += 256) dy[(i / K) * (K + 1) + (i % K)] = dyn[(size_t)b * HT * K + i];
    __syncthreads();
    for (int i = tid; i < HT * D; i += 256) {
        const int h = i / D, d = i % D;
        float a = 0.f;
        for (int k = 0; k < K; ++k) a = fmaf(dy[h * (K + 1) + k], Ss[k * LD + d] + (float)(k / H + 1) / 12.f, a);
        E1[h * LD + d] = lrelu(a);
    }
    __syncthreads();
    float* vb = v + (size_t)b * K * D;
    for (int k = warp; k < K; k += kWarps) {
        float r[D / 32];
        float q = 0.f;
#pragma unroll
        for (int j = 0; j < D / 32; ++j) {
            const int d = lane + 32 * j;
            float a = 0.f;
            for (int h = 0; h < HT; ++h) a = fmaf(dy[h * (K + 1) + k], E1[h * LD + d], a);
            r[j] = lrelu(a) + Ss[k * LD + d];
            q += r[j] * r[j];
        }
        q = warp_sum(q);
        const float f = squash_f(q);
#pragma unroll
        for (int j = 0; j < D / 32; ++j) vb[(size_t)k * D + lane + 32 * j] = r[j] * f;
    }
}

template <int D>
__global__ void __launch_bounds__(256) cap_hop_bwd_kernel(const float* __restrict__ s, const float* __restrict__ dyn,
                                                          const float* __restrict__ dv, float* __restrict__ ds,
                                                          float* __restrict__ ddyn, int T, int H, int HT) {
    extern __shared__ __align__(16) float smem[];
    const int K = T * H, LD = D + 1, LK = K + 1;
    float* Ss = smem;                   // [K][LD] raw s
    float* P2 = Ss + (size_t)K * LD;    // [K][LD] pre2, later dpre2
    float* DR = P2 + (size_t)K * LD;    // [K][LD] dr
    float* P1 = DR + (size_t)K * LD;    // [HT][LD] pre1
    float* D1 = P1 + (size_t)HT * LD;   // [HT][LD] dpre1
    float* dy = D1 + (size_t)HT * LD;   // [HT][LK]
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int b = blockIdx.x;
    const float* sb = s + (size_t)b * K * D;
    for (int i = tid; i < K * D; i += 256) Ss[(i / D) * LD + (i % D)] = sb[i];
    for (int i = tid; i < HT * K; i += 256) dy[(i / K) * LK + (i % K)] = dyn[(size_t)b * HT * K + i];
    __syncthreads();
    for (int i = tid; i < HT * D; i += 256) {
        const int h = i / D, d = i % D;
        float a = 0.f;
        for (int k = 0; k < K; ++k) a = fmaf(dy[h * LK + k], Ss[k * LD + d] + (float)(k / H + 1) / 12.f, a);
        P1[h * LD + d] = a;
    }
    __syncthreads();
    const float* dvb = dv + (size_t)b * K * D;
    for (int k = warp; k < K; k += kWarps) {
        float r[D / 32], g[D / 32], p2[D / 32];
        float q = 0.f, rg = 0.f;
#pragma unroll
        for (int j = 0; j < D / 32; ++j) {
            const int d = lane + 32 * j;
            float a = 0.f;
            for (int h = 0; h < HT; ++h) a = fmaf(dy[h * LK + k], lrelu(P1[h * LD + d]), a);
            p2[j] = a;
            r[j] = lrelu(a) + Ss[k * LD + d];
            g[j] = dvb[(size_t)k * D + d];
            q += r[j] * r[j];
            rg += r[j] * g[j];
        }
        q = warp_sum(q);
        rg = warp_sum(rg);
        const float f = squash_f(q), fp = squash_df(q);
#pragma unroll
        for (int j = 0; j < D / 32; ++j) {
            const int d = lane + 32 * j;
            const float dr = f * g[j] + 2.f * r[j] * fp * rg;
            DR[k * LD + d] = dr;
            P2[k * LD + d] = lrelu_grad(p2[j], dr);  // dpre2
        }
    }
    __syncthreads();
    for (int i = tid; i < HT * D; i += 256) {
        const int h = i / D, d = i % D;
        float a = 0.f;
        for (int k = 0; k < K; ++k) a = fmaf(dy[h * LK + k], P2[k * LD + d], a);
        D1[h * LD + d] = lrelu_grad(P1[h * LD + d], a);  // dpre1
    }
    __syncthreads();
    for (int i = tid; i < HT * K; i += 256) {
        const int h = i / K, k = i % K;
        const float tau = (float)(k / H + 1) / 12.f;
        float a = 0.f;
        for (int d = 0; d < D; ++d)
            a = fmaf(lrelu(P1[h * LD + d]), P2[k * LD + d], fmaf(D1[h * LD + d], Ss[k * LD + d] + tau, a));
        ddyn[(size_t)b * HT * K + i] = a;
    }
    float* dsb = ds + (size_t)b * K * D;
    for (int i = tid; i < K * D; i += 256) {
        const int k = i / D, d = i % D;
        float a = DR[k * LD + d];
        for (int h = 0; h < HT; ++h) a = fmaf(dy[h * LK + k], D1[h * LD + d], a);
        dsb[i] = a;
    }
}

// recon[b,t,n,:] = sum_h c[b,t,h,n] v[b,t,h,:]            GPTST.py:135
template <int D>
__global__ void __launch_bounds__(256) cap_recon_kernel(const float* __restrict__ c, const float* __restrict__ v,
                                                        float* __restrict__ recon, int N, int H) {
    __shared__ __align__(16) float vs[kMaxH * D];
    const int slab = blockIdx.x;
    for (int i = threadIdx.x; i < H * D; i += 256) vs[i] = v[(size_t)slab * H * D + i];
    __syncthreads();
    constexpr int VPR = D / 4, NPC = 256 / VPR;
    const int nl = threadIdx.x / VPR, cv = threadIdx.x % VPR;
    for (int n = blockIdx.y * NPC + nl; n < N; n += gridDim.y * NPC) {
        float4 o = make_float4(0.f, 0.f, 0.f, 0.f);
        for (int h = 0; h < H; ++h) {
            const float cc = c[((size_t)slab * H + h) * N + n];
            const float4 vv = *reinterpret_cast<const float4*>(vs + h * D + cv * 4);
            o.x = fmaf(cc, vv.x, o.x); o.y = fmaf(cc, vv.y, o.y); o.z = fmaf(cc, vv.z, o.z); o.w = fmaf(cc, vv.w, o.w);
        }
        *reinterpret_cast<float4*>(recon + ((size_t)slab * N + n) * D + cv * 4) = o;
    }
}

// dv[b,t,h,:] = sum_n c[h,n] drecon[n,:]      dc_r[b,t,h,n] = v[h,:] . drecon[n,:]
template <int D, int HP>
__global__ void __launch_bounds__(256) cap_dv_dcr_kernel(const float* __restrict__ c, const float* __restrict__ v,
                                                         const float* __restrict__ drecon, float* __restrict__ dv,
                                                         float* __restrict__ dcr, int N, int H) {
    constexpr int VPR = D / 4, NL = 256 / VPR;   // threads per row, node lanes
    extern __shared__ __align__(16) float smem[];
    float* vs = smem;                 // [H][D]
    float* red = vs + kMaxH * D;      // [NL][H][D]
    const int slab = blockIdx.x;
    const int tid = threadIdx.x, nl = tid / VPR, cv = tid % VPR;
    for (int i = tid; i < H * D; i += 256) vs[i] = v[(size_t)slab * H * D + i];
    __syncthreads();
    float4 acc[HP];
#pragma unroll
    for (int h = 0; h < HP; ++h) acc[h] = make_float4(0.f, 0.f, 0.f, 0.f);
    const int niter = (N + NL - 1) / NL;
    for (int it = 0; it < niter; ++it) {
        const int n = it * NL + nl;
        const bool valid = n < N;
        float4 g = make_float4(0.f, 0.f, 0.f, 0.f);
        if (valid) g = *reinterpret_cast<const float4*>(drecon + ((size_t)slab * N + n) * D + cv * 4);
#pragma unroll
        for (int h = 0; h < HP; ++h) {
            if (h < H) {
                const float cc = valid ? c[((size_t)slab * H + h) * N + n] : 0.f;
                acc[h].x = fmaf(cc, g.x, acc[h].x); acc[h].y = fmaf(cc, g.y, acc[h].y);
                acc[h].z = fmaf(cc, g.z, acc[h].z); acc[h].w = fmaf(cc, g.w, acc[h].w);
                const float4 vv = *reinterpret_cast<const float4*>(vs + h * D + cv * 4);
                float dot = g.x * vv.x + g.y * vv.y + g.z * vv.z + g.w * vv.w;
                // reduce over the VPR threads of this row (VPR = 16 or 32, aligned inside a warp)
#pragma unroll
                for (int o = VPR / 2; o > 0; o >>= 1) dot += __shfl_xor_sync(0xffffffffu, dot, o);
                if (cv == 0 && valid) dcr[((size_t)slab * H + h) * N + n] = dot;
            }
        }
    }
#pragma unroll
    for (int h = 0; h < HP; ++h)
        if (h < H) *reinterpret_cast<float4*>(red + ((size_t)nl * H + h) * D + cv * 4) = acc[h];
    __syncthreads();
    for (int i = tid; i < H * D; i += 256) {
        float sacc = 0.f;
        for (int l = 0; l < NL; ++l) sacc += red[(size_t)l * H * D + i];
        dv[(size_t)slab * H * D + i] = sacc;
    }
}

// ------------------------------------------------------------------------------------------------------
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
constexpr size_t kSmemMax = 227 * 1024;

static int pick_cluster(int N, int D, int H, int* rpc_out) {
    static const int sizes[5] = {1, 2, 4, 8, 16};
    for (int i = 0; i < 5; ++i) {
        int cs = sizes[i];
        int rpc = (N + cs - 1) / cs;
        rpc = (rpc + 15) / 16 * 16;
        if (route_fwd_smem_floats(D, H, rpc) * 4 <= kSmemMax) {
            *rpc_out = rpc;
            return cs;
        }
    }
    return -1;
}

template <int D, int PREC, int HP>
static cudaError_t launch_route_fwd(const float* x, const float* Wp, const float* bp, const float* dadj, float* c,
                                    float* s, int BT, int N, int H, int R, cudaStream_t st) {
    int rpc = 0;
    int cs = pick_cluster(N, D, H, &rpc);
    if (cs < 0) return cudaErrorInvalidValue;
    size_t smem = route_fwd_smem_floats(D, H, rpc) * 4;
    auto kern = cap_route_fwd_kernel<D, PREC, HP>;
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

extern "C" int gptst_cap_route_fwd(const float* x, const float* Wp, const float* bp, const float* dadj, float* c,
                                   float* s, int B, int T, int N, int D, int H, int R, int prec, void* stream) {
    if (!x || !Wp || !bp || !dadj || !c || !s || B <= 0 || T <= 0 || N <= 0 || R < 0) return -1;
    cudaStream_t st = (cudaStream_t)stream;
#define CALL(DD, PP, HH) return (int)launch_route_fwd<DD, PP, HH>(x, Wp, bp, dadj, c, s, B * T, N, H, R, st)
    CAP_DISPATCH(D, prec, H, CALL);
#undef CALL
    return -2;
}

extern "C" int gptst_cap_hop_fwd(const float* s, const float* dyn, float* v, int B, int T, int D, int H, int HT,
                                 void* stream) {
    if (!s || !dyn || !v || B <= 0) return -1;
    if (H > kMaxH) return -2;
    cudaStream_t st = (cudaStream_t)stream;
    const int K = T * H;
    size_t smem = ((size_t)K * (D + 1) + (size_t)HT * (D + 1) + (size_t)HT * (K + 1)) * 4;
    if (smem > kSmemMax) return -2;
    cudaError_t e;
    if (D == 64) {
        e = cudaFuncSetAttribute(cap_hop_fwd_kernel<64>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) return (int)e;
        cap_hop_fwd_kernel<64><<<B, 256, smem, st>>>(s, dyn, v, T, H, HT);
    } else if (D == 128) {
        e = cudaFuncSetAttribute(cap_hop_fwd_kernel<128>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) return (int)e;
        cap_hop_fwd_kernel<128><<<B, 256, smem, st>>>(s, dyn, v, T, H, HT);
    } else return -2;
    return (int)cudaGetLastError();
}

extern "C" int gptst_cap_hop_bwd(const float* s, const float* dyn, const float* dv, float* ds, float* ddyn, int B, int T,
                                 int D, int H, int HT, void* stream) {
    if (!s || !dyn || !dv || !ds || !ddyn || B <= 0) return -1;
    if (H > kMaxH) return -2;
    cudaStream_t st = (cudaStream_t)stream;
    const int K = T * H;
    size_t smem = (3 * (size_t)K * (D + 1) + 2 * (size_t)HT * (D + 1) + (size_t)HT * (K + 1)) * 4;
    if (smem > kSmemMax) return -2;
    cudaError_t e;
    if (D == 64) {
        e = cudaFuncSetAttribute(cap_hop_bwd_kernel<64>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) return (int)e;
        cap_hop_bwd_kernel<64><<<B, 256, smem, st>>>(s, dyn, dv, ds, ddyn, T, H, HT);
    } else if (D == 128) {
        e = cudaFuncSetAttribute(cap_hop_bwd_kernel<128>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) return (int)e;
        cap_hop_bwd_kernel<128><<<B, 256, smem, st>>>(s, dyn, dv, ds, ddyn, T, H, HT);
    } else return -2;
    return (int)cudaGetLastError();
}

extern "C" int gptst_cap_recon(const float* c, const float* v, float* recon, int B, int T, int N, int D, int H,
                               void* stream) {
    if (!c || !v || !recon || B <= 0 || N <= 0) return -1;
    if (H > kMaxH) return -2;
    cudaStream_t st = (cudaStream_t)stream;
    const int npc = 256 / (D / 4);
    int ychunks = (N + npc - 1) / npc;
    int want = (592 + B * T - 1) / (B * T);
    if (ychunks > want) ychunks = want;
    if (ychunks < 1) ychunks = 1;
    dim3 grid(B * T, ychunks);
    if (D == 64) cap_recon_kernel<64><<<grid, 256, 0, st>>>(c, v, recon, N, H);
    else if (D == 128) cap_recon_kernel<128><<<grid, 256, 0, st>>>(c, v, recon, N, H);
    else return -2;
    return (int)cudaGetLastError();
}

extern "C" int gptst_cap_dv_dcr(const float* c, const float* v, const float* drecon, float* dv, float* dcr, int B, int T,
                                int N, int D, int H, void* stream) {
    if (!c || !v || !drecon || !dv || !dcr || B <= 0 || N <= 0) return -1;
    if (H > kMaxH) return -2;
    cudaStream_t st = (cudaStream_t)stream;
    const int nlanes = 256 / (D / 4);
    size_t smem = ((size_t)kMaxH * D + (size_t)nlanes * H * D) * 4;
    cudaError_t e;
#define DV(DD, HH)                                                                                                       \
    do {                                                                                                                 \
        e = cudaFuncSetAttribute(cap_dv_dcr_kernel<DD, HH>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);     \
        if (e != cudaSuccess) return (int)e;                                                                             \
        cap_dv_dcr_kernel<DD, HH><<<B * T, 256, smem, st>>>(c, v, drecon, dv, dcr, N, H);                                \
    } while (0)
    if (D == 64 && H == 10) DV(64, 10);
    else if (D == 64) DV(64, 16);
    else if (D == 128 && H == 10) DV(128, 10);
    else if (D == 128) DV(128, 16);
    else return -2;
#undef DV
    return (int)cudaGetLastError();
}

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
