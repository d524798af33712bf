// Fused pre-training losses, forward value AND analytic gradients in one pass.
//
//   probe   : mean |(o - x) * m|                                 (driver GPU-probe loss)
//   mask_mae: mean over { t*m > thr } of |t*m - p*m|,  p = o*std+mean, t = x*std+mean
//             (reference Run.py:91-101 + lib/metrics.py:11-18, without the data-dependent masked_select)
//   KL      : w * sum hs * (log hs - log prob)                    (Run.py:132 + BasicTrainer.py:84-86, w = 0.1)
//
// The loss is the root of the autograd graph (grad_output == 1), so the kernel also emits d loss / d o and
// d loss / d prob; the Python autograd.Function just hands them back (scaled by grad_output).  Two launches:
// per-block partial sums, then a finalise that also scales the gradients by 1/count (the count is data dependent).
#include "common.cuh"

namespace gptst {

// part[block] = {sum |e|, count, sum KL}
__global__ void __launch_bounds__(256) loss_partial_kernel(const float* __restrict__ o, const float* __restrict__ src,
                                                           const long long* __restrict__ inv_mask, const float* __restrict__ prob,
                                                           const float* __restrict__ hs, float* __restrict__ d_o,
                                                           float* __restrict__ d_prob, float* __restrict__ part, long n_cells,
                                                           int ibd, int src_stride, int H, int mode, float mean, float std_,
                                                           float thr, float kl_w) {
    float se = 0.f, cnt = 0.f, kl = 0.f;
    const long stride = (long)gridDim.x * blockDim.x;
    for (long i = (long)blockIdx.x * blockDim.x + threadIdx.x; i < n_cells * ibd; i += stride) {
        const long cell = i / ibd;
        const int ch = (int)(i % ibd);
        const float m = (float)inv_mask[i];
        const float x = src[cell * src_stride + ch];
        const float ov = o[i];
        float e, g;
        if (mode == 0) {                       // probe
            e = (ov - x) * m;
            g = (e > 0.f ? 1.f : (e < 0.f ? -1.f : 0.f)) * m;
            cnt += 1.f;
        } else {                               // mask_mae
            // the reference's inverse z-score is two rounded fp32 ops (lib/normalization.py: data * std + mean), NOT a fused
            // multiply-add: whether a gap cell (true == 0 after the transform) passes `t > thr` depends on that last bit
            const float p = __fadd_rn(__fmul_rn(ov, std_), mean) * m, t = __fadd_rn(__fmul_rn(x, std_), mean) * m;
            const bool sel = t > thr;
            e = sel ? (t - p) : 0.f;
            g = sel ? ((e > 0.f ? -1.f : (e < 0.f ? 1.f : 0.f)) * std_ * m) : 0.f;
            cnt += sel ? 1.f : 0.f;
        }
        se += fabsf(e);
        d_o[i] = g;                            // scaled by 1/count in the finalise pass
    }
    if (kl_w != 0.f) {
        for (long i = (long)blockIdx.x * blockDim.x + threadIdx.x; i < n_cells * H; i += stride) {
            const float t = hs[i], p = prob[i];
            const float lp = logf(p);
            kl += (t > 0.f) ? t * (logf(t) - lp) : 0.f;      // xlogy semantics of KLDivLoss
            d_prob[i] = -kl_w * t / p;
        }
    }
    __shared__ float red[3][8];
    se = warp_sum(se); cnt = warp_sum(cnt); kl = warp_sum(kl);
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    if (lane == 0) { red[0][warp] = se; red[1][warp] = cnt; red[2][warp] = kl; }
    __syncthreads();
    if (threadIdx.x < 3) {
        float s = 0.f;
#pragma unroll
        for (int w = 0; w < 8; ++w) s += red[threadIdx.x][w];
        part[blockIdx.x * 3 + threadIdx.x] = s;
    }
}

// out = {loss, mae, kl_sum}; d_o *= 1/count
__global__ void __launch_bounds__(256) loss_final_kernel(const float* __restrict__ part, int nparts, float* __restrict__ d_o,
                                                         long n, float kl_w, float* __restrict__ out) {
    __shared__ float tot[3];
    __shared__ float red[3][8];
    {   // every block sums the partials itself, in a fixed (thread-strided, then tree) order: deterministic
        float s0 = 0.f, s1 = 0.f, s2 = 0.f;
        for (int i = threadIdx.x; i < nparts; i += 256) { s0 += part[i * 3]; s1 += part[i * 3 + 1]; s2 += part[i * 3 + 2]; }
        s0 = warp_sum(s0); s1 = warp_sum(s1); s2 = warp_sum(s2);
        const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
        if (lane == 0) { red[0][warp] = s0; red[1][warp] = s1; red[2][warp] = s2; }
        __syncthreads();
        if (threadIdx.x < 3) {
            float s = 0.f;
#pragma unroll
            for (int w = 0; w < 8; ++w) s += red[threadIdx.x][w];
            tot[threadIdx.x] = s;
        }
    }
    __syncthreads();
    const float inv = tot[1] > 0.f ? 1.f / tot[1] : 0.f;
    if (blockIdx.x == 0 && threadIdx.x == 0) {
        const float mae = tot[0] * inv;
        out[0] = mae + kl_w * tot[2];
        out[1] = mae;
        out[2] = tot[2];
    }
    const long stride = (long)gridDim.x * blockDim.x;
    for (long i = (long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) d_o[i] *= inv;
}

}  // namespace gptst

using namespace gptst;

extern "C" int gptst_loss_parts(void) { return 4 * 148; }

extern "C" int gptst_pretrain_loss(const float* o, const float* src, const long long* inv_mask, const float* prob,
                                   const float* hs, float* d_o, float* d_prob, float* part, float* out, long n_cells, int ibd,
                                   int src_stride, int H, int mode, float mean, float std_, float thr, float kl_w, void* stream) {
    if (!o || !src || !inv_mask || !d_o || !part || !out || n_cells <= 0) return -1;
    if (kl_w != 0.f && (!prob || !hs || !d_prob)) return -1;
    if (mode != 0 && mode != 1) return -2;
    cudaStream_t st = (cudaStream_t)stream;
    const int nparts = gptst_loss_parts();
    loss_partial_kernel<<<nparts, 256, 0, st>>>(o, src, inv_mask, prob, hs, d_o, d_prob, part, n_cells, ibd, src_stride, H, mode,
                                                mean, std_, thr, kl_w);
    loss_final_kernel<<<nparts, 256, 0, st>>>(part, nparts, d_o, n_cells * ibd, kl_w, out);
    return (int)cudaGetLastError();
}
