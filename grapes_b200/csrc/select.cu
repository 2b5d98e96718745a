// Per-hop node selection: Gumbel-top-k over the candidate frontier + Bernoulli log-probability of
// the chosen mask + sampler statistics.  Replaces sample_neighborhoods_from_probs
// (/root/reference/modules/utils.py:13-71) and the deterministic top-k of the mini-batch evaluator
// (/root/reference/eval.py:126-130), without the per-hop D2H of the mask (utils.py:60).
//
// The k-th largest key is found by an MSB-first radix select on order-preserving uint32 keys;
// warps build digit histograms with match-any aggregation.  Ties at the threshold go to the
// LOWEST candidate index (torch.topk leaves ties unspecified; the oracle uses the same rule).
#include "common.cuh"

#define SEL_THREADS 1024
#define SEL_WARPS (SEL_THREADS / 32)

__device__ __forceinline__ uint32_t float_to_ordered(float f) {
    if (f != f) return 0u;                                   // NaN ranks lowest
    const uint32_t u = __float_as_uint(f);
    return (u & 0x80000000u) ? ~u : (u | 0x80000000u);
}

// Philox4x32-10 (Salmon et al. 2011): counter-based, one call = 4 uniforms
__device__ __forceinline__ uint4 philox4x32_10(uint4 ctr, uint2 key) {
    const uint32_t M0 = 0xD2511F53u, M1 = 0xCD9E8D57u, W0 = 0x9E3779B9u, W1 = 0xBB67AE85u;
#pragma unroll
    for (int r = 0; r < 10; ++r) {
        const uint32_t hi0 = __umulhi(M0, ctr.x), lo0 = M0 * ctr.x;
        const uint32_t hi1 = __umulhi(M1, ctr.z), lo1 = M1 * ctr.z;
        ctr = make_uint4(hi1 ^ ctr.y ^ key.x, lo1, hi0 ^ ctr.w ^ key.y, lo0);
        key.x += W0; key.y += W1;
    }
    return ctr;
}
__device__ __forceinline__ float philox_uniform(unsigned long long seed, unsigned long long offset, int i) {
    uint4 ctr = make_uint4((uint32_t)(i >> 2), 0u, (uint32_t)offset, (uint32_t)(offset >> 32));
    const uint4 r = philox4x32_10(ctr, make_uint2((uint32_t)seed, (uint32_t)(seed >> 32)));
    const uint32_t x = (i & 3) == 0 ? r.x : (i & 3) == 1 ? r.y : (i & 3) == 2 ? r.z : r.w;
    return (float)(x & 0xffffffu) * (1.0f / 16777216.0f);   // 24-bit mantissa in [0,1), as torch.rand
}

__device__ __forceinline__ float sigmoidf_(float l) { return 1.0f / (1.0f + expf(-l)); }
// Bernoulli(logits=l).log_prob(y) = -BCEWithLogits(l, y)  (utils.py:71)
__device__ __forceinline__ float bern_log_prob(float l, float y) {
    return -((1.0f - y) * l + fmaxf(-l, 0.f) + log1pf(expf(-fabsf(l))));
}
__device__ __forceinline__ float entropy_bits(float p) {
    const float e = -(p * log2f(p) + (1.0f - p) * log2f(1.0f - p));
    return (e != e) ? 0.f : e;                                // NaN entropy -> 0 (utils.py:52-54)
}

template <typename T, typename Op>
__device__ __forceinline__ T block_reduce(T v, T* smem, Op op, T identity) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v = op(v, __shfl_xor_sync(GRAPES_FULL_MASK, v, o));
    if (lane_id() == 0) smem[threadIdx.x >> 5] = v;
    __syncthreads();
    T r = (threadIdx.x < SEL_WARPS) ? smem[threadIdx.x] : identity;
    if (threadIdx.x < 32) {
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) r = op(r, __shfl_xor_sync(GRAPES_FULL_MASK, r, o));
        if (threadIdx.x == 0) smem[0] = r;
    }
    __syncthreads();
    r = smem[0];
    __syncthreads();
    return r;
}
struct OpAdd { __device__ float operator()(float a, float b) const { return a + b; } };
struct OpMin { __device__ float operator()(float a, float b) const { return fminf(a, b); } };
struct OpMax { __device__ float operator()(float a, float b) const { return fmaxf(a, b); } };

// mode: GRAPES_NOISE_*
//
// One thread-block CLUSTER of SEL_CTAS CTAs x 1024 threads (distributed shared memory): every CTA
// scans a strided share of the candidates, digit histograms are merged into CTA 0's shared memory
// with DSMEM atomics, CTA 0 picks the digit and every CTA reads the running prefix back through
// DSMEM.  Integer atomics only -> the result does not depend on scheduling.
#include <cooperative_groups.h>
namespace cg = cooperative_groups;

#define SEL_CTAS 8

struct SelShared {
    int hist[256];                 // CTA 0: cluster-wide digit histogram of the current pass
    uint32_t prefix;               // CTA 0: bits of the threshold key fixed so far
    int kr;                        // CTA 0: rank still to be resolved inside the current prefix
    float part[SEL_CTAS][8];       // CTA 0: per-CTA partial reductions (min, max, esum, evar, lp, dl)
    long long counts[SEL_CTAS];    // CTA 0: per-CTA (n_eq << 32 | n_gt) of the final pass
};

__global__ void __cluster_dims__(SEL_CTAS, 1, 1) __launch_bounds__(SEL_THREADS) k_select(
    const float* __restrict__ logits_all, const int* __restrict__ nb_local, const int* __restrict__ nb_nodes,
    const int* __restrict__ c_dev, int cap_c, int k, int mode, const float* __restrict__ noise,
    unsigned long long* rng_state, uint32_t* __restrict__ ukeys, float* __restrict__ keys_out,
    int* __restrict__ sampled_out, int sampled_offset, int* __restrict__ s_dev, int* __restrict__ total_dev,
    uint8_t* __restrict__ mask_out, float* __restrict__ log_prob, float* tot_log_prob, float* __restrict__ stats,
    float* __restrict__ dl_all, float* sum_dl, uint32_t* bm_mark) {
    cg::cluster_group cluster = cg::this_cluster();
    const int rank = (int)cluster.block_rank();
    __shared__ SelShared sh;
    __shared__ int s_whist[SEL_WARPS / 4][256];      // 8 sub-histograms per CTA (4 warps share one)
    __shared__ float s_red[SEL_WARPS];
    __shared__ long long s_scan[SEL_WARPS + 2];
    SelShared* sh0 = cluster.map_shared_rank(&sh, 0);
    const int c = min(*c_dev, cap_c);
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int gtid = rank * SEL_THREADS + tid, gthreads = SEL_CTAS * SEL_THREADS;
    const bool take_all = (k >= c);                       // utils.py:31-33: no noise is drawn
    unsigned long long seed = 0ull, offset = 0ull;
    if (mode == GRAPES_NOISE_PHILOX) { seed = rng_state[0]; offset = rng_state[1]; }
    __shared__ float s_bcast_f;
    __shared__ uint32_t s_bcast_u;
    __shared__ long long s_bcast_ll;
    if (tid < 256) sh.hist[tid] = 0;
    if (tid == 0) { sh.prefix = 0u; sh.kr = k; }
    cluster.sync();                                        // every CTA of the cluster is running: DSMEM is safe

    // ---- pass 0: keys + probability statistics (coalesced, strided over the whole cluster) ----
    float pmin = INFINITY, pmax = -INFINITY, esum = 0.f;
    for (int i = gtid; i < c; i += gthreads) {
        const float l = logits_all[nb_local ? nb_local[i] : i];
        const float p = sigmoidf_(l);
        float key;
        if (mode == GRAPES_NOISE_KEYS) key = noise[i];
        else if (mode == GRAPES_NOISE_NONE_TOPK_PROBS) key = p;               // eval.py:126
        else {
            float g;
            if (mode == GRAPES_NOISE_GUMBEL) g = noise[i];
            else {
                float u = (mode == GRAPES_NOISE_UNIFORM) ? noise[i] : philox_uniform(seed, offset, i);
                if (mode == GRAPES_NOISE_PHILOX) u = u * ((1.0f - 1.1920929e-07f) - 1.17549435e-38f) + 1.17549435e-38f;
                g = -logf(-logf(u));
            }
            key = logf(p) + g;                                                // utils.py:42
        }
        if (!take_all) ukeys[i] = float_to_ordered(key);
        if (keys_out) keys_out[i] = key;
        pmin = fminf(pmin, p); pmax = fmaxf(pmax, p);
        esum += entropy_bits(p);
    }
    pmin = block_reduce(pmin, s_red, OpMin(), INFINITY);
    pmax = block_reduce(pmax, s_red, OpMax(), -INFINITY);
    esum = block_reduce(esum, s_red, OpAdd(), 0.f);
    if (tid == 0) { sh0->part[rank][0] = pmin; sh0->part[rank][1] = pmax; sh0->part[rank][2] = esum; }
    cluster.sync();                                        // also publishes ukeys[] to the other CTAs
    if (tid == 0) {
        float t = 0.f;
        for (int r = 0; r < SEL_CTAS; ++r) t += sh0->part[r][2];              // fixed order
        s_bcast_f = (c > 0) ? t / (float)c : 0.f;
    }
    __syncthreads();
    const float emean = s_bcast_f;
    float evar = 0.f;
    for (int i = gtid; i < c; i += gthreads) {
        const float d = entropy_bits(sigmoidf_(logits_all[nb_local ? nb_local[i] : i])) - emean;
        evar = fmaf(d, d, evar);
    }
    evar = block_reduce(evar, s_red, OpAdd(), 0.f);
    if (tid == 0) sh0->part[rank][3] = evar;

    // ---- radix select of the k-th largest key (MSB first, 8 bits per pass) -------------------
    uint32_t thr = 0u;
    int kr = 0;
    if (!take_all) {
        for (int shift = 24; shift >= 0; shift -= 8) {
            for (int b = tid; b < (SEL_WARPS / 4) * 256; b += SEL_THREADS) (&s_whist[0][0])[b] = 0;
            if (tid == 0) s_bcast_u = sh0->prefix;
            __syncthreads();
            const uint32_t prefix = s_bcast_u;
            const uint32_t himask = (shift == 24) ? 0u : (0xffffffffu << (shift + 8));
            for (int base = 0; base < c; base += gthreads) {
                const int i = base + gtid;
                bool valid = false; uint32_t digit = 0u;
                if (i < c) {
                    const uint32_t u = ukeys[i];
                    valid = (u & himask) == prefix;
                    digit = (u >> shift) & 255u;
                }
                const unsigned peers = __match_any_sync(GRAPES_FULL_MASK, valid ? digit : 0xffffffffu);
                if (valid && lane == (__ffs(peers) - 1)) atomicAdd(&s_whist[warp >> 2][digit], __popc(peers));
            }
            __syncthreads();
            if (tid < 256) {
                int a = 0;
#pragma unroll
                for (int w = 0; w < SEL_WARPS / 4; ++w) a += s_whist[w][tid];
                if (a) atomicAdd(&sh0->hist[tid], a);                          // DSMEM atomic into CTA 0
            }
            cluster.sync();
            if (rank == 0 && warp == 0) {
                // walk bins from the top; lane l owns bins [255-8l-7 .. 255-8l] (descending order)
                int mine[8], msum = 0;
#pragma unroll
                for (int t = 0; t < 8; ++t) { mine[t] = sh.hist[255 - (lane * 8 + t)]; msum += mine[t]; }
                const int incl = warp_scan_incl(msum);
                const int excl = incl - msum;
                const int krem = sh.kr;
                const bool here = (excl < krem) && (incl >= krem);
                __syncwarp();
                if (here) {
                    int cum = excl;
#pragma unroll
                    for (int t = 0; t < 8; ++t) {
                        if (cum + mine[t] >= krem) {
                            sh.prefix = prefix | ((uint32_t)(255 - (lane * 8 + t)) << shift);
                            sh.kr = krem - cum;
                            break;
                        }
                        cum += mine[t];
                    }
                }
                __syncwarp();
#pragma unroll
                for (int t = 0; t < 8; ++t) sh.hist[lane * 8 + t] = 0;        // clean for the next pass
            }
            cluster.sync();
        }
        thr = sh0->prefix;
        kr = sh0->kr;
    }

    // ---- ordered selection: CTA r owns a contiguous chunk, each thread a contiguous sub-range ----
    const int chunk = (c + SEL_CTAS - 1) / SEL_CTAS;
    const int cb = min(c, rank * chunk), ce = min(c, cb + chunk);
    const int ipt = (ce - cb + SEL_THREADS - 1) / SEL_THREADS;
    const int i0 = min(ce, cb + tid * ipt), i1 = min(ce, i0 + ipt);
    int n_eq = 0, n_gt = 0;
    if (!take_all) {
        for (int i = i0; i < i1; ++i) {
            const uint32_t u = ukeys[i];
            n_eq += (u == thr);
            n_gt += (u > thr);
        }
    }
    long long total;
    const long long packed = ((long long)n_eq << 32) | (long long)n_gt;
    const long long excl = block_scan_excl<long long>(packed, s_scan, &total);
    if (tid == 0) sh0->counts[rank] = total;
    cluster.sync();
    if (tid == 0) {
        long long b = 0;
        for (int r = 0; r < rank; ++r) b += sh0->counts[r];
        s_bcast_ll = b;
    }
    __syncthreads();
    const long long before = s_bcast_ll;
    int eq_before = (int)((before + excl) >> 32), gt_before = (int)((before + excl) & 0xffffffffll);
    float lp_sum = 0.f, dl_sum = 0.f;
    for (int i = i0; i < i1; ++i) {
        bool sel;
        int pos;
        if (take_all) { sel = true; pos = i; }
        else {
            const uint32_t u = ukeys[i];
            sel = false; pos = 0;
            if (u > thr) { sel = true; pos = gt_before + min(eq_before, kr); ++gt_before; }
            else if (u == thr) { if (eq_before < kr) { sel = true; pos = gt_before + eq_before; } ++eq_before; }
        }
        const int li = nb_local ? nb_local[i] : i;
        const float l = logits_all[li];
        const float y = sel ? 1.f : 0.f;
        const float lp = bern_log_prob(l, y);
        if (log_prob) log_prob[i] = lp;
        lp_sum += lp;
        const float d = y - sigmoidf_(l);
        if (dl_all) dl_all[li] = d;
        dl_sum += d;
        if (mask_out) mask_out[i] = sel ? 1 : 0;
        if (sel) {
            const int g = nb_nodes ? nb_nodes[i] : i;
            if (sampled_out) sampled_out[sampled_offset + pos] = g;
            if (bm_mark && nb_nodes) bitmap_set(bm_mark, g);
        }
    }
    lp_sum = block_reduce(lp_sum, s_red, OpAdd(), 0.f);
    dl_sum = block_reduce(dl_sum, s_red, OpAdd(), 0.f);
    if (tid == 0) { sh0->part[rank][4] = lp_sum; sh0->part[rank][5] = dl_sum; }
    cluster.sync();
    if (rank == 0 && tid == 0) {
        float mn = INFINITY, mx = -INFINITY, ev = 0.f, lp = 0.f, dl = 0.f;
        for (int r = 0; r < SEL_CTAS; ++r) {                                   // fixed order -> deterministic
            mn = fminf(mn, sh.part[r][0]); mx = fmaxf(mx, sh.part[r][1]);
            ev += sh.part[r][3]; lp += sh.part[r][4]; dl += sh.part[r][5];
        }
        const int s = take_all ? c : k;
        if (s_dev) *s_dev = s;
        if (total_dev) *total_dev = sampled_offset + s;
        if (tot_log_prob) *tot_log_prob += lp;
        if (sum_dl) *sum_dl += dl;
        if (stats) {
            if (take_all) { stats[0] = stats[1] = stats[2] = stats[3] = 0.f; }   // reference returns {}
            else {
                stats[0] = mn; stats[1] = mx; stats[2] = emean;
                stats[3] = (c > 1) ? sqrtf(ev / (float)(c - 1)) : 0.f;         // unbiased (utils.py:56)
            }
        }
        if (mode == GRAPES_NOISE_PHILOX && !take_all) rng_state[1] = offset + 1ull;
    }
    cluster.sync();                                        // keep CTA 0's shared memory alive until all are done
}

// ---------------------------------------------------------------------------------------
// k_select_reg: the same selection with every candidate key held ON CHIP (shared memory, blocked layout: CTA r owns
// a contiguous chunk, thread t a contiguous sub-range of <= SEL_MAX_IPT items, stored column-wise so accesses are
// conflict free), so keys are computed and read from global memory exactly once; 3 radix passes of 11 / 11 / 10
// bits.  Used when c <= 8 * 1024 * SEL_MAX_IPT, else k_select (keys in global memory).
// ---------------------------------------------------------------------------------------
#define SEL_MAX_IPT 24
#define SEL_BINS 2048

struct SelSharedR {
    int hist[SEL_BINS];
    uint32_t prefix;
    int kr;
    float part[SEL_CTAS][8];
    long long counts[SEL_CTAS];
};

__global__ void __cluster_dims__(SEL_CTAS, 1, 1) __launch_bounds__(SEL_THREADS, 1) k_select_reg(
    const float* __restrict__ logits_all, const int* __restrict__ nb_local, const int* __restrict__ nb_nodes,
    const int* __restrict__ c_dev, int cap_c, int k, int mode, const float* __restrict__ noise,
    unsigned long long* rng_state, float* __restrict__ keys_out, int* __restrict__ sampled_out, int sampled_offset,
    int* __restrict__ s_dev, int* __restrict__ total_dev, uint8_t* __restrict__ mask_out, float* __restrict__ log_prob,
    float* tot_log_prob, float* __restrict__ stats, float* __restrict__ dl_all, float* sum_dl, uint32_t* bm_mark) {
    cg::cluster_group cluster = cg::this_cluster();
    const int rank = (int)cluster.block_rank();
    __shared__ SelSharedR sh;
    __shared__ int s_lhist[SEL_BINS];
    __shared__ float s_red[SEL_WARPS];
    __shared__ long long s_scan[SEL_WARPS + 2];
    __shared__ uint32_t s_bcast_u;
    __shared__ long long s_bcast_ll;
    SelSharedR* sh0 = cluster.map_shared_rank(&sh, 0);
    const int c = min(*c_dev, cap_c);
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const bool take_all = (k >= c);
    unsigned long long seed = 0ull, offset = 0ull;
    if (mode == GRAPES_NOISE_PHILOX) { seed = rng_state[0]; offset = rng_state[1]; }
    for (int b = tid; b < SEL_BINS; b += SEL_THREADS) sh.hist[b] = 0;
    if (tid == 0) { sh.prefix = 0u; sh.kr = k; }
    cluster.sync();

    const int chunk = (c + SEL_CTAS - 1) / SEL_CTAS;
    const int cb = min(c, rank * chunk), ce = min(c, cb + chunk);
    const int ipt = (chunk + SEL_THREADS - 1) / SEL_THREADS;       // <= SEL_MAX_IPT (checked by the launcher)
    const int i0 = min(ce, cb + tid * ipt);
    const int cnt = max(0, min(ce, i0 + ipt) - i0);

    // ---- pass 0: load once, keys + statistics ----
    extern __shared__ uint32_t s_uk[];                             // [ipt][1024] keys, then [ipt][1024] logits
#define UK(j) s_uk[(j) * SEL_THREADS + tid]
    float* s_lg = reinterpret_cast<float*>(s_uk + ipt * SEL_THREADS);
#define LG(j) s_lg[(j) * SEL_THREADS + tid]
    // gather the logits first, 8 independent loads in flight per thread (the index -> logit chain is latency bound)
    for (int j0 = 0; j0 < cnt; j0 += 8) {
        int li[8];
#pragma unroll
        for (int u = 0; u < 8; ++u) li[u] = (j0 + u < cnt) ? (nb_local ? nb_local[i0 + j0 + u] : i0 + j0 + u) : 0;
        float lv[8];
#pragma unroll
        for (int u = 0; u < 8; ++u) lv[u] = (j0 + u < cnt) ? logits_all[li[u]] : 0.f;
#pragma unroll
        for (int u = 0; u < 8; ++u) if (j0 + u < cnt) LG(j0 + u) = lv[u];
    }
    float pmin = INFINITY, pmax = -INFINITY, esum = 0.f, esq = 0.f;
    for (int j = 0; j < cnt; ++j) {
        {
            const int i = i0 + j;
            const float l = LG(j);
            const float p = sigmoidf_(l);
            float key;
            if (mode == GRAPES_NOISE_KEYS) key = noise[i];
            else if (mode == GRAPES_NOISE_NONE_TOPK_PROBS) key = p;
            else {
                float g;
                if (mode == GRAPES_NOISE_GUMBEL) g = noise[i];
                else {
                    float u = (mode == GRAPES_NOISE_UNIFORM) ? noise[i] : philox_uniform(seed, offset, i);
                    if (mode == GRAPES_NOISE_PHILOX) u = u * ((1.0f - 1.1920929e-07f) - 1.17549435e-38f) + 1.17549435e-38f;
                    g = -logf(-logf(u));
                }
                key = logf(p) + g;
            }
            UK(j) = float_to_ordered(key);
            if (keys_out) keys_out[i] = key;
            pmin = fminf(pmin, p); pmax = fmaxf(pmax, p);
            const float e = entropy_bits(p);
            esum += e; esq = fmaf(e, e, esq);
        }
    }
    pmin = block_reduce(pmin, s_red, OpMin(), INFINITY);
    pmax = block_reduce(pmax, s_red, OpMax(), -INFINITY);
    esum = block_reduce(esum, s_red, OpAdd(), 0.f);
    esq = block_reduce(esq, s_red, OpAdd(), 0.f);
    if (tid == 0) { sh0->part[rank][0] = pmin; sh0->part[rank][1] = pmax; sh0->part[rank][2] = esum; sh0->part[rank][3] = esq; }

    // ---- radix select: 11 + 11 + 10 bits ----
    uint32_t thr = 0u;
    int kr = 0;
    if (!take_all) {
        const int shifts[3] = {21, 10, 0};
        const int nbits[3] = {11, 11, 10};
#pragma unroll
        for (int ps = 0; ps < 3; ++ps) {
            const int shift = shifts[ps];
            const uint32_t dmask = (1u << nbits[ps]) - 1u;
            for (int b = tid; b < SEL_BINS; b += SEL_THREADS) s_lhist[b] = 0;
            if (tid == 0) s_bcast_u = sh0->prefix;
            __syncthreads();
            const uint32_t prefix = s_bcast_u;
            const uint32_t himask = (ps == 0) ? 0u : (0xffffffffu << (shift + nbits[ps]));
            for (int j = 0; j < cnt; ++j) {
                const uint32_t u = UK(j);
                if ((u & himask) == prefix) atomicAdd(&s_lhist[(u >> shift) & dmask], 1);
            }
            __syncthreads();
            for (int b = tid; b < SEL_BINS; b += SEL_THREADS) {
                const int a = s_lhist[b];
                if (a) atomicAdd(&sh0->hist[b], a);
            }
            cluster.sync();
            if (rank == 0 && warp == 0) {
                // lane l owns bins [2047-64l-63 .. 2047-64l] in descending order
                int msum = 0;
                for (int t = 0; t < 64; ++t) msum += sh.hist[SEL_BINS - 1 - (lane * 64 + t)];
                const int incl = warp_scan_incl(msum);
                const int excl = incl - msum;
                const int krem = sh.kr;
                if (excl < krem && incl >= krem) {
                    int cum = excl;
                    for (int t = 0; t < 64; ++t) {
                        const int bin = SEL_BINS - 1 - (lane * 64 + t);
                        const int h = sh.hist[bin];
                        if (cum + h >= krem) { sh.prefix = prefix | ((uint32_t)bin << shift); sh.kr = krem - cum; break; }
                        cum += h;
                    }
                }
                __syncwarp();
                for (int t = 0; t < 64; ++t) sh.hist[lane * 64 + t] = 0;
            }
            cluster.sync();
        }
        thr = sh0->prefix;
        kr = sh0->kr;
    }

    // ---- ordered selection straight from registers ----
    int n_eq = 0, n_gt = 0;
    if (!take_all) {
        for (int j = 0; j < cnt; ++j) { const uint32_t u = UK(j); n_eq += (u == thr); n_gt += (u > thr); }
    }
    long long total;
    const long long packed = ((long long)n_eq << 32) | (long long)n_gt;
    const long long excl = block_scan_excl<long long>(packed, s_scan, &total);
    if (tid == 0) sh0->counts[rank] = total;
    cluster.sync();
    if (tid == 0) {
        long long b = 0;
        for (int r = 0; r < rank; ++r) b += sh0->counts[r];
        s_bcast_ll = b;
    }
    __syncthreads();
    const long long before = s_bcast_ll;
    int eq_before = (int)((before + excl) >> 32), gt_before = (int)((before + excl) & 0xffffffffll);
    float lp_sum = 0.f, dl_sum = 0.f;
    for (int j = 0; j < cnt; ++j) {
        {
            const int i = i0 + j;
            bool sel; int pos;
            if (take_all) { sel = true; pos = i; }
            else {
                const uint32_t u = UK(j);
                sel = false; pos = 0;
                if (u > thr) { sel = true; pos = gt_before + min(eq_before, kr); ++gt_before; }
                else if (u == thr) { if (eq_before < kr) { sel = true; pos = gt_before + eq_before; } ++eq_before; }
            }
            const float l = LG(j);
            const float y = sel ? 1.f : 0.f;
            const float lp = bern_log_prob(l, y);
            if (log_prob) log_prob[i] = lp;
            lp_sum += lp;
            const float d = y - sigmoidf_(l);
            if (dl_all) dl_all[nb_local ? nb_local[i] : i] = d;
            dl_sum += d;
            if (mask_out) mask_out[i] = sel ? 1 : 0;
            if (sel) {
                const int g = nb_nodes ? nb_nodes[i] : i;
                if (sampled_out) sampled_out[sampled_offset + pos] = g;
                if (bm_mark && nb_nodes) bitmap_set(bm_mark, g);
            }
        }
    }
    lp_sum = block_reduce(lp_sum, s_red, OpAdd(), 0.f);
    dl_sum = block_reduce(dl_sum, s_red, OpAdd(), 0.f);
    if (tid == 0) { sh0->part[rank][4] = lp_sum; sh0->part[rank][5] = dl_sum; }
    cluster.sync();
    if (rank == 0 && tid == 0) {
        float mn = INFINITY, mx = -INFINITY, lp = 0.f, dl = 0.f;
        double es = 0.0, eq = 0.0;
        for (int r = 0; r < SEL_CTAS; ++r) {
            mn = fminf(mn, sh.part[r][0]); mx = fmaxf(mx, sh.part[r][1]);
            es += (double)sh.part[r][2]; eq += (double)sh.part[r][3]; lp += sh.part[r][4]; dl += sh.part[r][5];
        }
        const int s = take_all ? c : k;
        if (s_dev) *s_dev = s;
        if (total_dev) *total_dev = sampled_offset + s;
        if (tot_log_prob) *tot_log_prob += lp;
        if (sum_dl) *sum_dl += dl;
        if (stats) {
            if (take_all) { stats[0] = stats[1] = stats[2] = stats[3] = 0.f; }
            else {
                const double mean = (c > 0) ? es / (double)c : 0.0;
                const double var = (c > 1) ? fmax(0.0, (eq - (double)c * mean * mean) / (double)(c - 1)) : 0.0;
                stats[0] = mn; stats[1] = mx; stats[2] = (float)mean; stats[3] = (float)sqrt(var);
            }
        }
        if (mode == GRAPES_NOISE_PHILOX && !take_all) rng_state[1] = offset + 1ull;
    }
    cluster.sync();
}

extern "C" {

int grapes_select_topk(grapes_ctx* ctx, const float* logits_all, const int* nb_local, const int* nb_nodes,
                       const int* c_dev, int cap_c, int k, int noise_mode, const float* noise,
                       unsigned long long* rng_state, uint32_t* ukeys_scratch, float* keys_out, int* sampled_out,
                       int sampled_offset, int* s_dev, int* total_dev, uint8_t* mask_out, float* log_prob,
                       float* tot_log_prob, float* stats, float* dl_all, float* sum_dl, uint32_t* bm_mark,
                       void* stream) {
    GRAPES_REQUIRE(ctx && logits_all && c_dev && ukeys_scratch, "null argument");
    GRAPES_REQUIRE(k > 0, "num_samples must be positive (utils.py:35)");
    GRAPES_REQUIRE(noise_mode >= 0 && noise_mode <= GRAPES_NOISE_NONE_TOPK_PROBS, "bad noise mode");
    GRAPES_REQUIRE(!(noise_mode == GRAPES_NOISE_GUMBEL || noise_mode == GRAPES_NOISE_UNIFORM ||
                     noise_mode == GRAPES_NOISE_KEYS) || noise, "noise array required");
    GRAPES_REQUIRE(noise_mode != GRAPES_NOISE_PHILOX || rng_state, "rng_state required");
    if (cap_c <= SEL_CTAS * SEL_THREADS * SEL_MAX_IPT) {
        const int chunk = (cap_c + SEL_CTAS - 1) / SEL_CTAS;
        const int ipt = (chunk + SEL_THREADS - 1) / SEL_THREADS;
        const int smem = 2 * ipt * SEL_THREADS * 4;
        static int attr = 0;
        if (smem > attr) {
            GRAPES_CUDA_OK(cudaFuncSetAttribute(k_select_reg, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
            attr = smem;
        }
        k_select_reg<<<SEL_CTAS, SEL_THREADS, smem, (cudaStream_t)stream>>>(
            logits_all, nb_local, nb_nodes, c_dev, cap_c, k, noise_mode, noise, rng_state, keys_out, sampled_out,
            sampled_offset, s_dev, total_dev, mask_out, log_prob, tot_log_prob, stats, dl_all, sum_dl, bm_mark);
        grapes_count_launches(1);
        GRAPES_LAUNCH_OK();
        return GRAPES_OK;
    }
    k_select<<<SEL_CTAS, SEL_THREADS, 0, (cudaStream_t)stream>>>(logits_all, nb_local, nb_nodes, c_dev, cap_c, k, noise_mode,
                                                          noise, rng_state, ukeys_scratch, keys_out, sampled_out,
                                                          sampled_offset, s_dev, total_dev, mask_out, log_prob,
                                                          tot_log_prob, stats, dl_all, sum_dl, bm_mark);
    grapes_count_launches(1);
    GRAPES_LAUNCH_OK();
    return GRAPES_OK;
}

}  // extern "C"
