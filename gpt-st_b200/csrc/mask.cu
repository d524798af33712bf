// On-device construction of the pre-training masks (reference GPTST.py:312-413).
//
// The reference builds the adaptive mask with a host loop (one D2H sync per class), two full radix sorts of the
// B*T*N keys and several scatters.  What those sorts compute is only "zero the k largest keys, ties in index order"
// (torch.sort on CUDA is a stable radix sort), i.e. a selection, not a sort.  Here:
//
//   mask_labels : label = argmax_h prob (first maximum), class histogram (integer atomics: deterministic)
//   mask_select : ONE CTA.  role of every class from the shuffled class order + the two budgets (the reference's
//                 while-loop as a 10-element prefix sum), then two exact top-k selections by (key desc, index asc):
//                     m_ada : zero the (ada - n_full) largest of  part * u1 ; classes masked outright are zeroed too
//                     m_rnd : zero the rnd largest of  m_ada * u2
//                 final = m_ada * m_rnd  (int64, replicated over the input_base_dim channels)
//                 mode 1 (random phase, GPTST.py:316-323): one selection of the k largest of u1.
// Keys are the reference's own torch.rand draws (multiples of 2^-24 in [0,1)), mapped to integers by q = u * 2^32
// (exact, order preserving).  A selection is: 2048-bin histogram of the top 11 bits -> the bin holding the k-th key ->
// its (few) members are ranked exactly by (q desc, index asc) in shared memory -> the threshold is a (q*, index*) pair and
// every element compares itself against it lexicographically.  Bins that are too large for the candidate buffer (only
// the all-zero keys can do that) go down further radix levels and finally select on the index alone.
#include "common.cuh"

namespace gptst {
namespace mk {

constexpr int NT = 1024;        // threads of the selection CTA
constexpr int CAP = 2048;       // candidate buffer (elements of the threshold bin)
constexpr int BINS = 2048;

struct Thr { unsigned q; int idx; };   // select every element with (q, -idx) >= (thr.q, -thr.idx); idx = -1 selects nothing

__global__ void __launch_bounds__(256) mask_labels_kernel(const float* __restrict__ prob, unsigned char* __restrict__ label,
                                                          int* __restrict__ counts, long n, int H) {
    __shared__ int hist[kMaxH];
    if (threadIdx.x < kMaxH) hist[threadIdx.x] = 0;
    __syncthreads();
    const long i = (long)blockIdx.x * 256 + threadIdx.x;
    if (i < n) {
        const float* p = prob + i * H;
        float best = p[0];
        int arg = 0;
        for (int h = 1; h < H; ++h) {
            const float v = p[h];
            if (v > best) { best = v; arg = h; }
        }
        label[i] = (unsigned char)arg;
        atomicAdd(&hist[arg], 1);
    }
    __syncthreads();
    if (threadIdx.x < H && hist[threadIdx.x]) atomicAdd(&counts[threadIdx.x], hist[threadIdx.x]);
}

// class histogram of given labels (parity tests inject the reference's labels)
__global__ void __launch_bounds__(256) mask_count_kernel(const unsigned char* __restrict__ label, int* __restrict__ counts, long n,
                                                         int H) {
    __shared__ int hist[kMaxH];
    if (threadIdx.x < kMaxH) hist[threadIdx.x] = 0;
    __syncthreads();
    const long i = (long)blockIdx.x * 256 + threadIdx.x;
    if (i < n && label[i] < H) atomicAdd(&hist[label[i]], 1);
    __syncthreads();
    if (threadIdx.x < H && hist[threadIdx.x]) atomicAdd(&counts[threadIdx.x], hist[threadIdx.x]);
}

__device__ __forceinline__ unsigned key_q(float u) { return __float2uint_rz(u * 4294967296.f); }

// Exact selection of the k largest keys in (q desc, index asc) order.  Load4: (base, q[4]) -> the keys of elements
// base..base+3 (0 beyond n).  Block-wide; every thread returns the same threshold.  The CTA streams the keys with 16-byte
// loads, two groups in flight per thread.  sm: hist[BINS] ints, cq[CAP] unsigned, ci[CAP] ints, misc[8] ints.
template <class Load4>
__device__ Thr block_select(Load4 load4, int n, long k, int* hist, unsigned* cq, int* ci, int* misc) {
    const int tid = threadIdx.x, lane = tid & 31;
    if (k <= 0) return Thr{0xffffffffu, -1};
    if (k >= n) return Thr{0u, n};                 // everything (index n is beyond the last element)
    unsigned prefix = 0, pmask = 0;                // candidates so far: (q & pmask) == prefix
    long rem = k;
    const int shifts[3] = {21, 10, 0};
    const int widths[3] = {11, 11, 10};
    for (int lvl = 0; lvl < 3; ++lvl) {
        const int sh = shifts[lvl], nb = 1 << widths[lvl];
        for (int i = tid; i < BINS; i += NT) hist[i] = 0;
        __syncthreads();
        int zeros = 0;                             // q == 0 (masked-out elements) would serialise on one bin: count apart
#pragma unroll 2
        for (int base = tid * 4; base < n; base += NT * 4) {
            unsigned q[4];
            load4(base, q);
#pragma unroll
            for (int e = 0; e < 4; ++e) {
                if (base + e < n && (q[e] & pmask) == prefix) {
                    if (q[e] == 0u) ++zeros;
                    else atomicAdd(&hist[(q[e] >> sh) & (nb - 1)], 1);
                }
            }
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) zeros += __shfl_xor_sync(0xffffffffu, zeros, o);
        if (lane == 0 && zeros) atomicAdd(&hist[0], zeros);
        __syncthreads();
        if (tid == 0) {                            // walk the bins from the top: nb <= 2048 adds, once per level
            long above = 0;
            int b = nb - 1;
            for (; b > 0; --b) {
                if (above + hist[b] >= rem) break;
                above += hist[b];
            }
            misc[0] = b;
            misc[1] = hist[b];
            misc[2] = (int)(rem - above);           // how many of this bin are selected (>= 1, <= hist[b])
        }
        __syncthreads();
        const int b = misc[0], cnt = misc[1];
        rem = misc[2];
        prefix |= (unsigned)b << sh;
        pmask |= (unsigned)(nb - 1) << sh;
        __syncthreads();
        if (cnt <= CAP) {
            // collect the bin, rank its members exactly, the member of rank rem-1 is the threshold
            if (tid == 0) misc[3] = 0;
            __syncthreads();
#pragma unroll 2
            for (int base = tid * 4; base < n; base += NT * 4) {
                unsigned q[4];
                load4(base, q);
#pragma unroll
                for (int e = 0; e < 4; ++e) {
                    if (base + e < n && (q[e] & pmask) == prefix) {
                        const int slot = atomicAdd(&misc[3], 1);
                        cq[slot] = q[e];
                        ci[slot] = base + e;
                    }
                }
            }
            __syncthreads();
            for (int a = tid; a < cnt; a += NT) {
                const unsigned qa = cq[a];
                const int ia = ci[a];
                int rank = 0;
                for (int j = 0; j < cnt; ++j) rank += (cq[j] > qa) || (cq[j] == qa && ci[j] < ia);
                if (rank == (int)rem - 1) { misc[4] = (int)qa; misc[5] = ia; }
            }
            __syncthreads();
            Thr t{(unsigned)misc[4], misc[5]};
            __syncthreads();
            return t;
        }
    }
    // more than CAP elements share the threshold key exactly (the masked-out zeros): select the rem smallest indices
    // among them by bisection on the index (rare path)
    int lo = 0, hi = n - 1;                         // smallest index x with #{i <= x, q_i == prefix} >= rem
    while (lo < hi) {
        const int mid = (lo + hi) >> 1;
        int c = 0;
        for (int base = tid * 4; base <= mid; base += NT * 4) {
            unsigned q[4];
            load4(base, q);
#pragma unroll
            for (int e = 0; e < 4; ++e) c += (base + e <= mid && q[e] == prefix);
        }
        if (tid == 0) misc[6] = 0;
        __syncthreads();
        if (c) atomicAdd(&misc[6], c);
        __syncthreads();
        const int tot = misc[6];
        __syncthreads();
        if (tot >= rem) hi = mid; else lo = mid + 1;
    }
    return Thr{prefix, lo};
}

__device__ __forceinline__ bool selected(unsigned q, int i, Thr t) { return q > t.q || (q == t.q && i <= t.idx); }

__device__ __forceinline__ void load_u4(const float* __restrict__ u, int base, int n, float (&v)[4]) {
    if (base + 3 < n) {
        const float4 f = *reinterpret_cast<const float4*>(u + base);
        v[0] = f.x; v[1] = f.y; v[2] = f.z; v[3] = f.w;
    } else {
#pragma unroll
        for (int e = 0; e < 4; ++e) v[e] = (base + e < n) ? u[base + e] : 0.f;
    }
}
__device__ __forceinline__ void load_b4(const unsigned char* __restrict__ p, int base, int n, int (&v)[4]) {
    if (base + 3 < n) {
        const uchar4 c = *reinterpret_cast<const uchar4*>(p + base);
        v[0] = c.x; v[1] = c.y; v[2] = c.z; v[3] = c.w;
    } else {
#pragma unroll
        for (int e = 0; e < 4; ++e) v[e] = (base + e < n) ? p[base + e] : 0;
    }
}

// mode 2: adaptive phase.  mode 1: random phase (labels / counts / u2 unused, plan[0] = k).
__global__ void __launch_bounds__(NT) mask_select_kernel(const unsigned char* __restrict__ label, const int* __restrict__ counts,
                                                         const long long* __restrict__ plan, const float* __restrict__ u1,
                                                         const float* __restrict__ u2, unsigned char* __restrict__ m_ada,
                                                         long long* __restrict__ final_mask, int n, int H, int i0, int all_type,
                                                         int mode) {
    __shared__ int hist[BINS];
    __shared__ unsigned cq[CAP];
    __shared__ int ci[CAP];
    __shared__ int misc[8];
    __shared__ int lut[kMaxH];
    __shared__ long long ks[2];
    const int tid = threadIdx.x;
    if (mode == 1) {
        const long k = (long)plan[0];
        auto keyr = [&](int base, unsigned (&q)[4]) {
            float v[4];
            load_u4(u1, base, n, v);
#pragma unroll
            for (int e = 0; e < 4; ++e) q[e] = key_q(v[e]);
        };
        const Thr t = block_select(keyr, n, k, hist, cq, ci, misc);
        for (int base = tid * 4; base < n; base += NT * 4) {
            unsigned q[4];
            keyr(base, q);
            for (int e = 0; e < 4; ++e) {
                if (base + e < n) {
                    const long long m = selected(q[e], base + e, t) ? 0 : 1;
                    for (int c = 0; c < i0; ++c) final_mask[(long)(base + e) * i0 + c] = m;
                }
            }
        }
        return;
    }
    if (tid < kMaxH) lut[tid] = 0;
    __syncthreads();
    if (tid == 0) {
        // the reference's class-selection loop (GPTST.py:357-384): classes in shuffled order until the budget is covered
        const long long ada = plan[H], rnd = plan[H + 1];
        long long cum = 0, n_full = 0;
        int npick = 0;
        for (int j = 0; j < H; ++j) {
            if (cum < ada) npick = j + 1;
            cum += counts[(int)plan[j]];
        }
        for (int j = 0; j < H; ++j) {
            const int cls = (int)plan[j];
            const bool picked = j < npick;
            int role = 0;                                     // 1 = sub-sampled class, 2 = masked outright
            if (picked) {
                if (all_type && npick >= 2) role = (j == npick - 1) ? 1 : 2;
                else role = 1;
            }
            lut[cls] = role;
            if (role == 2) n_full += counts[cls];
        }
        ks[0] = ada - n_full;
        ks[1] = rnd;
    }
    __syncthreads();
    const long k1 = (long)ks[0], k2 = (long)ks[1];
    auto key1 = [&](int base, unsigned (&q)[4]) {
        float v[4];
        int lb[4];
        load_u4(u1, base, n, v);
        load_b4(label, base, n, lb);
#pragma unroll
        for (int e = 0; e < 4; ++e) q[e] = (lut[lb[e]] == 1) ? key_q(v[e]) : 0u;
    };
    const Thr t1 = block_select(key1, n, k1, hist, cq, ci, misc);
    for (int base = tid * 4; base < n; base += NT * 4) {
        float v[4];
        int lb[4];
        load_u4(u1, base, n, v);
        load_b4(label, base, n, lb);
#pragma unroll
        for (int e = 0; e < 4; ++e) {
            if (base + e < n) {
                const int role = lut[lb[e]];
                const unsigned q = role == 1 ? key_q(v[e]) : 0u;
                m_ada[base + e] = (role == 2 || selected(q, base + e, t1)) ? 0 : 1;
            }
        }
    }
    __syncthreads();
    auto key2 = [&](int base, unsigned (&q)[4]) {
        float v[4];
        int ma[4];
        load_u4(u2, base, n, v);
        load_b4(m_ada, base, n, ma);
#pragma unroll
        for (int e = 0; e < 4; ++e) q[e] = ma[e] ? key_q(v[e]) : 0u;
    };
    const Thr t2 = block_select(key2, n, k2, hist, cq, ci, misc);
    for (int base = tid * 4; base < n; base += NT * 4) {
        unsigned q[4];
        int ma[4];
        key2(base, q);
        load_b4(m_ada, base, n, ma);
#pragma unroll
        for (int e = 0; e < 4; ++e) {
            if (base + e < n) {
                const long long m = (ma[e] && !selected(q[e], base + e, t2)) ? 1 : 0;
                for (int c = 0; c < i0; ++c) final_mask[(long)(base + e) * i0 + c] = m;
            }
        }
    }
}


// ------------------------------------------------------------------------------------------------------------------
// Multi-CTA pipeline (the single-CTA kernel above streams 6 x 0.6 MB through one SM: ~220 us at B*T*N = 130k; it stays
// as the exact slow path and as the specification).  One selection = three grid-wide passes:
//     hist    : 2048-bin histogram of the top 11 key bits (global integer atomics; the all-zero keys are counted apart)
//     collect : every CTA finds the threshold bin with a parallel suffix scan of the histogram, members of that bin
//               are appended to a candidate list
//     apply   : every CTA ranks the (few) candidates exactly by (q desc, index asc) -> threshold (q*, index*), applies
//               it to its elements; for the first selection of the adaptive phase this pass also writes m_ada and
//               accumulates the histogram of the second selection
// If a threshold bin has more than CAP members (only the masked-out zero keys can do that) a one-CTA kernel between
// collect and apply runs the exact single-CTA selection instead.
// Workspace (int32): [0,16) counts | [16, 16+2048) hist A | [.., +2048) hist B | then per selection s in {0,1} 8 ints
// {bin, cnt, rem, ncand, thr_q, thr_idx, special, -} | candidate q[2*CAP] | candidate idx[2*CAP]
// ------------------------------------------------------------------------------------------------------------------
constexpr int WS_COUNTS = 0, WS_HIST = 16, WS_SEL = 16 + 2 * BINS, WS_CQ = WS_SEL + 16, WS_CI = WS_CQ + 2 * CAP,
              WS_INTS = WS_CI + 2 * CAP;
enum { SEL_RANDOM = 0, SEL_ADA1 = 1, SEL_ADA2 = 2 };

struct MaskArgs {
    const unsigned char* label;
    const long long* plan;
    const float* u1;
    const float* u2;
    unsigned char* m_ada;
    long long* final_mask;
    int* ws;
    int n, H, i0, all_type;
};

// class roles and the two budgets (every CTA recomputes them: 2 x H steps of one thread)
__device__ void class_roles(const MaskArgs& a, int* lut, long long* ks) {
    const int* counts = a.ws + WS_COUNTS;
    const int H = a.H;
    const long long ada = a.plan[H], rnd = a.plan[H + 1];
    long long cum = 0, n_full = 0;
    int npick = 0;
    for (int j = 0; j < H; ++j) {
        if (cum < ada) npick = j + 1;
        cum += counts[(int)a.plan[j]];
    }
    for (int j = 0; j < kMaxH; ++j) lut[j] = 0;
    for (int j = 0; j < H; ++j) {
        const int cls = (int)a.plan[j];
        int role = 0;
        if (j < npick) role = (a.all_type && npick >= 2) ? ((j == npick - 1) ? 1 : 2) : 1;
        lut[cls] = role;
        if (role == 2) n_full += counts[cls];
    }
    ks[0] = ada - n_full;
    ks[1] = rnd;
}

template <int SEL>
__device__ __forceinline__ void keys4(const MaskArgs& a, const int* lut, int base, unsigned (&q)[4]) {
    float v[4];
    int aux[4];
    if (SEL == SEL_ADA2) { load_u4(a.u2, base, a.n, v); load_b4(a.m_ada, base, a.n, aux); }
    else { load_u4(a.u1, base, a.n, v); if (SEL == SEL_ADA1) load_b4(a.label, base, a.n, aux); }
#pragma unroll
    for (int e = 0; e < 4; ++e) {
        const bool on = (SEL == SEL_RANDOM) ? true : (SEL == SEL_ADA1) ? (lut[aux[e]] == 1) : (aux[e] != 0);
        q[e] = on ? key_q(v[e]) : 0u;
    }
}

template <int SEL>
__device__ __forceinline__ long sel_k(const MaskArgs& a, const long long* ks) {
    return (SEL == SEL_RANDOM) ? (long)a.plan[0] : (SEL == SEL_ADA1) ? (long)ks[0] : (long)ks[1];
}

__device__ __forceinline__ void hist_add4(const unsigned (&q)[4], int base, int n, int* hist, int& zeros) {
#pragma unroll
    for (int e = 0; e < 4; ++e) {
        if (base + e < n) {
            if (q[e] == 0u) ++zeros;
            else atomicAdd(&hist[q[e] >> 21], 1);
        }
    }
}
__device__ __forceinline__ void hist_flush_zeros(int zeros, int* hist) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) zeros += __shfl_xor_sync(0xffffffffu, zeros, o);
    if ((threadIdx.x & 31) == 0 && zeros) atomicAdd(&hist[0], zeros);
}

template <int SEL>
__global__ void __launch_bounds__(256) mask_hist_kernel(MaskArgs a) {
    __shared__ int lut[kMaxH];
    __shared__ long long ks[2];
    if (SEL == SEL_ADA1) {
        if (threadIdx.x == 0) class_roles(a, lut, ks);
        __syncthreads();
    }
    int* hist = a.ws + WS_HIST;
    int zeros = 0;
    for (int base = (blockIdx.x * 256 + threadIdx.x) * 4; base < a.n; base += gridDim.x * 256 * 4) {
        unsigned q[4];
        keys4<SEL>(a, lut, base, q);
        hist_add4(q, base, a.n, hist, zeros);
    }
    hist_flush_zeros(zeros, hist);
}

// parallel search of the threshold bin: 256 threads x 8 bins, from the top.  Returns through sh[0..2] = {bin, cnt, rem}.
__device__ void find_bin(const int* __restrict__ hist, long k, int* sh) {
    __shared__ long wsum[8];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    int hb[8];
    long s = 0;
#pragma unroll
    for (int e = 0; e < 8; ++e) { hb[e] = hist[BINS - 1 - (8 * tid + e)]; s += hb[e]; }
    long incl = s;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        const long v = __shfl_up_sync(0xffffffffu, incl, o);
        if (lane >= o) incl += v;
    }
    if (lane == 31) wsum[warp] = incl;
    __syncthreads();
    long off = 0;
    for (int w = 0; w < warp; ++w) off += wsum[w];
    const long excl = off + incl - s;       // keys in bins above this thread's range
    if (excl < k && k <= excl + s) {
        long above = excl;
        int e = 0;
        for (; e < 7; ++e) {
            if (above + hb[e] >= k) break;
            above += hb[e];
        }
        sh[0] = BINS - 1 - (8 * tid + e);
        sh[1] = hb[e];
        sh[2] = (int)(k - above);
    }
    __syncthreads();
}

template <int SEL>
__global__ void __launch_bounds__(256) mask_collect_kernel(MaskArgs a) {
    __shared__ int lut[kMaxH];
    __shared__ long long ks[2];
    __shared__ int sh[4];
    const int slot = (SEL == SEL_ADA2) ? 1 : 0;
    if (threadIdx.x == 0) {
        if (SEL != SEL_RANDOM) class_roles(a, lut, ks);
        sh[0] = 0; sh[1] = 0; sh[2] = 0;
    }
    __syncthreads();
    const long k = sel_k<SEL>(a, ks);
    int* sel = a.ws + WS_SEL + 8 * slot;
    if (k <= 0 || k >= a.n) {               // nothing / everything: no threshold bin
        if (blockIdx.x == 0 && threadIdx.x == 0) { sel[0] = 0; sel[1] = 0; sel[2] = 0; sel[6] = (k <= 0) ? 1 : 2; }
        return;
    }
    const int* hist = a.ws + WS_HIST + ((SEL == SEL_ADA2) ? BINS : 0);
    find_bin(hist, k, sh);
    const int b = sh[0], cnt = sh[1];
    if (blockIdx.x == 0 && threadIdx.x == 0) { sel[0] = b; sel[1] = cnt; sel[2] = sh[2]; sel[6] = 0; }
    if (cnt > CAP) return;                  // the exact one-CTA selection takes over
    unsigned* cq = reinterpret_cast<unsigned*>(a.ws + WS_CQ) + slot * CAP;
    int* ci = a.ws + WS_CI + slot * CAP;
    for (int base = (blockIdx.x * 256 + threadIdx.x) * 4; base < a.n; base += gridDim.x * 256 * 4) {
        unsigned q[4];
        keys4<SEL>(a, lut, base, q);
#pragma unroll
        for (int e = 0; e < 4; ++e) {
            if (base + e < a.n && (int)(q[e] >> 21) == b) {
                const int p = atomicAdd(&sel[3], 1);
                cq[p] = q[e];
                ci[p] = base + e;
            }
        }
    }
}

// exact fallback: only does work when the threshold bin overflowed the candidate buffer
template <int SEL>
__global__ void __launch_bounds__(NT) mask_slow_kernel(MaskArgs a) {
    __shared__ int hist[BINS];
    __shared__ unsigned cq[CAP];
    __shared__ int ci[CAP];
    __shared__ int misc[8];
    __shared__ int lut[kMaxH];
    __shared__ long long ks[2];
    int* sel = a.ws + WS_SEL + 8 * ((SEL == SEL_ADA2) ? 1 : 0);
    if (sel[6] != 0 || sel[1] <= CAP) return;
    if (threadIdx.x == 0 && SEL != SEL_RANDOM) class_roles(a, lut, ks);
    __syncthreads();
    const long k = sel_k<SEL>(a, ks);
    auto ld = [&](int base, unsigned (&q)[4]) { keys4<SEL>(a, lut, base, q); };
    const Thr t = block_select(ld, a.n, k, hist, cq, ci, misc);
    if (threadIdx.x == 0) { sel[4] = (int)t.q; sel[5] = t.idx; }
}

template <int SEL>
__global__ void __launch_bounds__(256) mask_apply_kernel(MaskArgs a) {
    __shared__ int lut[kMaxH];
    __shared__ long long ks[2];
    __shared__ unsigned scq[CAP];
    __shared__ int sci[CAP];
    __shared__ int thr[2];
    const int slot = (SEL == SEL_ADA2) ? 1 : 0;
    const int* sel = a.ws + WS_SEL + 8 * slot;
    const int tid = threadIdx.x;
    if (tid == 0 && SEL != SEL_RANDOM) class_roles(a, lut, ks);
    const int cnt = sel[1], rem = sel[2], special = sel[6];
    Thr t;
    if (special == 1) t = Thr{0xffffffffu, -1};
    else if (special == 2) t = Thr{0u, a.n};
    else if (cnt > CAP) t = Thr{(unsigned)sel[4], sel[5]};
    else {
        const unsigned* cq = reinterpret_cast<const unsigned*>(a.ws + WS_CQ) + slot * CAP;
        const int* ci = a.ws + WS_CI + slot * CAP;
        for (int i = tid; i < cnt; i += 256) { scq[i] = cq[i]; sci[i] = ci[i]; }
        __syncthreads();
        for (int x = tid; x < cnt; x += 256) {
            const unsigned qa = scq[x];
            const int ia = sci[x];
            int rank = 0;
            for (int j = 0; j < cnt; ++j) rank += (scq[j] > qa) || (scq[j] == qa && sci[j] < ia);
            if (rank == rem - 1) { thr[0] = (int)qa; thr[1] = ia; }
        }
        __syncthreads();
        t = Thr{(unsigned)thr[0], thr[1]};
    }
    __syncthreads();
    int* hist2 = a.ws + WS_HIST + BINS;
    int zeros = 0;
    for (int base = (blockIdx.x * 256 + tid) * 4; base < a.n; base += gridDim.x * 256 * 4) {
        unsigned q[4];
        keys4<SEL>(a, lut, base, q);
        if (SEL == SEL_ADA1) {
            int lb[4];
            float v2[4];
            load_b4(a.label, base, a.n, lb);
            load_u4(a.u2, base, a.n, v2);
            unsigned q2[4];
            unsigned char ma[4];
#pragma unroll
            for (int e = 0; e < 4; ++e) {
                ma[e] = (lut[lb[e]] == 2 || selected(q[e], base + e, t)) ? 0 : 1;
                q2[e] = ma[e] ? key_q(v2[e]) : 0u;
            }
            if (base + 3 < a.n) *reinterpret_cast<uchar4*>(a.m_ada + base) = make_uchar4(ma[0], ma[1], ma[2], ma[3]);
            else for (int e = 0; e < 4; ++e) if (base + e < a.n) a.m_ada[base + e] = ma[e];
            hist_add4(q2, base, a.n, hist2, zeros);
        } else {
            int ma[4] = {1, 1, 1, 1};
            if (SEL == SEL_ADA2) load_b4(a.m_ada, base, a.n, ma);
#pragma unroll
            for (int e = 0; e < 4; ++e) {
                if (base + e < a.n) {
                    const long long m = (ma[e] && !selected(q[e], base + e, t)) ? 1 : 0;
                    for (int c = 0; c < a.i0; ++c) a.final_mask[(long)(base + e) * a.i0 + c] = m;
                }
            }
        }
    }
    if (SEL == SEL_ADA1) hist_flush_zeros(zeros, hist2);
}

}  // namespace mk
}  // namespace gptst

using namespace gptst;

// label: n bytes, counts: H int32 (zeroed here), from prob (n, H)
extern "C" int gptst_mask_labels(const float* prob, unsigned char* label, int* counts, long n, int H, void* stream) {
    if (!prob || !label || !counts || n <= 0) return -1;
    if (H < 1 || H > kMaxH || n > 0x7fffffffL) return -2;
    cudaStream_t st = (cudaStream_t)stream;
    cudaError_t e = cudaMemsetAsync(counts, 0, sizeof(int) * H, st);
    if (e != cudaSuccess) return (int)e;
    mk::mask_labels_kernel<<<(unsigned)((n + 255) / 256), 256, 0, st>>>(prob, label, counts, n, H);
    return (int)cudaGetLastError();
}

// mode 2: plan = {order[0..H), adaptive_num, random_num} (int64), u1/u2 the two torch.rand draws, m_ada n bytes scratch.
// mode 1: plan[0] = number of cells to mask, u1 the draw; label / counts / u2 / m_ada may be NULL.
// final_mask: (n, i0) int64, 1 = keep, 0 = masked.
extern "C" int gptst_mask_select(const unsigned char* label, const int* counts, const long long* plan, const float* u1,
                                 const float* u2, unsigned char* m_ada, long long* final_mask, long n, int H, int i0,
                                 int all_type, int mode, void* stream) {
    if (!plan || !u1 || !final_mask || n <= 0 || i0 <= 0) return -1;
    if (mode == 2 && (!label || !counts || !u2 || !m_ada)) return -1;
    if ((mode != 1 && mode != 2) || H < 1 || H > kMaxH || n > 0x7fffffffL) return -2;
    mk::mask_select_kernel<<<1, mk::NT, 0, (cudaStream_t)stream>>>(label, counts, plan, u1, u2, m_ada, final_mask, (int)n, H, i0,
                                                                  all_type, mode);
    return (int)cudaGetLastError();
}

// ---- the multi-CTA pipeline (what the model calls) ----------------------------------------------------------------------
extern "C" int gptst_mask_ws_ints(void) { return mk::WS_INTS; }

// adaptive phase: prob (n, H) -> final_mask (n, i0) int64; plan / u1 / u2 as for gptst_mask_select; label (n bytes), m_ada
// (n bytes) and ws (gptst_mask_ws_ints() int32) are scratch.  label_in != NULL skips the arg-max (labels given, uint8).
extern "C" int gptst_mask_adaptive(const float* prob, const unsigned char* label_in, const long long* plan, const float* u1,
                                   const float* u2, unsigned char* label, unsigned char* m_ada, int* ws,
                                   long long* final_mask, long n, int H, int i0, int all_type, void* stream) {
    if ((!prob && !label_in) || !plan || !u1 || !u2 || !label || !m_ada || !ws || !final_mask || n <= 0 || i0 <= 0) return -1;
    if (H < 1 || H > kMaxH || n > 0x7fffffffL) return -2;
    cudaStream_t st = (cudaStream_t)stream;
    cudaError_t e = cudaMemsetAsync(ws, 0, sizeof(int) * mk::WS_INTS, st);
    if (e != cudaSuccess) return (int)e;
    const int nb = (int)((n + 1023) / 1024);          // 256 threads x 4 elements
    const int grid = nb < 296 ? nb : 296;
    if (label_in) {
        e = cudaMemcpyAsync(label, label_in, (size_t)n, cudaMemcpyDeviceToDevice, st);
        if (e != cudaSuccess) return (int)e;
        // class histogram of the given labels: reuse the label kernel's shared-memory histogram through a tiny pass
    }
    mk::MaskArgs a{label, plan, u1, u2, m_ada, final_mask, ws, (int)n, H, i0, all_type};
    if (!label_in) mk::mask_labels_kernel<<<(unsigned)((n + 255) / 256), 256, 0, st>>>(prob, label, ws + mk::WS_COUNTS, n, H);
    else mk::mask_count_kernel<<<(unsigned)((n + 255) / 256), 256, 0, st>>>(label, ws + mk::WS_COUNTS, n, H);
    mk::mask_hist_kernel<mk::SEL_ADA1><<<grid, 256, 0, st>>>(a);
    mk::mask_collect_kernel<mk::SEL_ADA1><<<grid, 256, 0, st>>>(a);
    mk::mask_slow_kernel<mk::SEL_ADA1><<<1, mk::NT, 0, st>>>(a);
    mk::mask_apply_kernel<mk::SEL_ADA1><<<grid, 256, 0, st>>>(a);
    mk::mask_collect_kernel<mk::SEL_ADA2><<<grid, 256, 0, st>>>(a);
    mk::mask_slow_kernel<mk::SEL_ADA2><<<1, mk::NT, 0, st>>>(a);
    mk::mask_apply_kernel<mk::SEL_ADA2><<<grid, 256, 0, st>>>(a);
    return (int)cudaGetLastError();
}

// random phase: zero the plan[0] largest of u (n,), final_mask (n,) int64
extern "C" int gptst_mask_random(const long long* k_dev, const float* u, int* ws, long long* final_mask, long n, void* stream) {
    if (!k_dev || !u || !ws || !final_mask || n <= 0) return -1;
    if (n > 0x7fffffffL) return -2;
    cudaStream_t st = (cudaStream_t)stream;
    cudaError_t e = cudaMemsetAsync(ws, 0, sizeof(int) * mk::WS_INTS, st);
    if (e != cudaSuccess) return (int)e;
    const int nb = (int)((n + 1023) / 1024);
    const int grid = nb < 296 ? nb : 296;
    mk::MaskArgs a{nullptr, k_dev, u, nullptr, nullptr, final_mask, ws, (int)n, 1, 1, 0};
    mk::mask_hist_kernel<mk::SEL_RANDOM><<<grid, 256, 0, st>>>(a);
    mk::mask_collect_kernel<mk::SEL_RANDOM><<<grid, 256, 0, st>>>(a);
    mk::mask_slow_kernel<mk::SEL_RANDOM><<<1, mk::NT, 0, st>>>(a);
    mk::mask_apply_kernel<mk::SEL_RANDOM><<<grid, 256, 0, st>>>(a);
    return (int)cudaGetLastError();
}
