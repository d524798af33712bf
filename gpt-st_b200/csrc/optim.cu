// Fused global-norm gradient clipping + Adam over a list of tensors (the optimiser half of the pre-training step,
// reference BasicTrainer.py:94-97: clip_grad_norm_(max_grad_norm) then Adam.step()).
//
// torch runs this as ~16 foreach launches over 135 small tensors (0.8 ms of a 6.3 ms step on B200, nothing to overlap
// with).  Here: a device table of (param, grad, exp_avg, exp_avg_sq, numel, step counter) entries plus a block map, and
// two launches: (1) per-block sum of squares of the gradients (+ the global step counter), (2) every block re-derives the
// global norm from the partials (fixed order, deterministic), forms the clip coefficient and applies Adam to its chunk.
// Math follows torch.optim.Adam (amsgrad=False, weight_decay=0): m = lerp(m, g, 1-b1); v = b2 v + (1-b2) g^2;
// p -= (lr / (1-b1^t)) * m / (sqrt(v)/sqrt(1-b2^t) + eps), with t counted PER PARAMETER ON THE DEVICE (one int32 each, bumped by the
// first launch): a parameter that gets its first gradient after thousands of graph replays still starts at t = 1, as
// torch.optim.Adam's per-parameter state['step'] does.
#include "common.cuh"

namespace gptst {

constexpr int kOptChunk = 2048;   // elements per block

struct OptEntry {     // 6 x int64 per tensor, filled by the host side (gptst_b200/optim.py); t = device int32* step counter
    long long p, g, m, v, n, t;
};

__global__ void __launch_bounds__(256) opt_sqnorm_kernel(const OptEntry* __restrict__ tab, const int2* __restrict__ blocks,
                                                         float* __restrict__ partial, int* __restrict__ step) {
    const int2 bm = blocks[blockIdx.x];
    const OptEntry e = tab[bm.x];
    const float* g = reinterpret_cast<const float*>(e.g);
    const long lo = (long)bm.y * kOptChunk, hi = min((long)e.n, lo + kOptChunk);
    float s = 0.f;
    for (long i = lo + threadIdx.x; i < hi; i += 256) { const float v = g[i]; s = fmaf(v, v, s); }
    __shared__ float red[8];
    s = warp_sum(s);
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = s;
    __syncthreads();
    if (threadIdx.x == 0) {
        float t = 0.f;
#pragma unroll
        for (int w = 0; w < 8; ++w) t += red[w];
        partial[blockIdx.x] = t;
        if (blockIdx.x == 0) step[0] += 1;
        if (bm.y == 0) *reinterpret_cast<int*>(e.t) += 1;            // this tensor's own step count (read by the second launch)
    }
}

// hyper = {lr, beta1, beta2, eps, max_norm (<=0: no clipping), grad_scale}.  grad_scale multiplies every gradient before the norm
// and the update: data-parallel training hands over the SUM of the ranks' gradients and 1/world here, so the averaging costs no
// pass of its own (gptst_b200/dp.py BucketedGradAllReduce).
__global__ void __launch_bounds__(256) opt_adam_kernel(const OptEntry* __restrict__ tab, const int2* __restrict__ blocks,
                                                       const float* __restrict__ partial, int nblocks, const int* __restrict__ step,
                                                       const float* __restrict__ hyper, float* __restrict__ norm_out) {
    __shared__ float red[8];
    __shared__ float coef_s;
    float s = 0.f;
    for (int i = threadIdx.x; i < nblocks; i += 256) s += partial[i];
    s = warp_sum(s);
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = s;
    __syncthreads();
    if (threadIdx.x == 0) {
        float t = 0.f;
#pragma unroll
        for (int w = 0; w < 8; ++w) t += red[w];
        const float norm = sqrtf(t) * hyper[5], max_norm = hyper[4];
        float c = 1.f;
        if (max_norm > 0.f) c = fminf(1.f, max_norm / (norm + 1e-6f));
        coef_s = c;
        if (blockIdx.x == 0 && norm_out) norm_out[0] = norm;
    }
    __syncthreads();
    const float coef = coef_s * hyper[5], lr = hyper[0], b1 = hyper[1], b2 = hyper[2], eps = hyper[3];
    const int2 bm = blocks[blockIdx.x];
    const OptEntry e = tab[bm.x];
    const float t = (float)(*reinterpret_cast<const int*>(e.t));
    const float bc1 = 1.f - powf(b1, t), bc2 = 1.f - powf(b2, t);
    const float step_size = lr / bc1, rs2 = rsqrtf(bc2);
    float* p = reinterpret_cast<float*>(e.p);
    const float* g = reinterpret_cast<const float*>(e.g);
    float* m = reinterpret_cast<float*>(e.m);
    float* v = reinterpret_cast<float*>(e.v);
    const long lo = (long)bm.y * kOptChunk, hi = min((long)e.n, lo + kOptChunk);
    for (long i = lo + threadIdx.x; i < hi; i += 256) {
        const float gi = g[i] * coef;
        const float mi = m[i] + (gi - m[i]) * (1.f - b1);
        const float vi = v[i] * b2 + gi * gi * (1.f - b2);
        m[i] = mi;
        v[i] = vi;
        p[i] -= step_size * mi / (sqrtf(vi) * rs2 + eps);
    }
}

}  // namespace gptst

using namespace gptst;

extern "C" int gptst_opt_chunk(void) { return kOptChunk; }

extern "C" int gptst_adam_clip(const void* table, const void* block_map, int nblocks, float* partial, int* step,
                               const float* hyper, float* norm_out, void* stream) {
    if (!table || !block_map || !partial || !step || !hyper || nblocks <= 0) return -1;
    cudaStream_t st = (cudaStream_t)stream;
    opt_sqnorm_kernel<<<nblocks, 256, 0, st>>>((const OptEntry*)table, (const int2*)block_map, partial, step);
    opt_adam_kernel<<<nblocks, 256, 0, st>>>((const OptEntry*)table, (const int2*)block_map, partial, nblocks, step, hyper, norm_out);
    return (int)cudaGetLastError();
}
