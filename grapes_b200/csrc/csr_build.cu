// CSR construction on the device: edge_index [2, E] -> (indptr int64, indices int32), rows sorted, duplicate edges collapsed,
// self-loops kept -- the canonical form scipy gives /root/reference/main.py:134-136
// (`sp.csr_matrix((np.ones(E, bool), edge_index), shape=(N, N))`, then row slicing in modules/utils.py:78,89-91).
//
// HBM-bound integer work, one pass per phase, no global sort: a counting sort by row (degree histogram -> 64-bit scan ->
// scatter) groups the edges, then every row segment is sorted and deduplicated where it lies -- a warp in registers for
// rows of <= 32 entries, a warp in shared memory up to 256, a CTA in shared memory for rows up to 32768 entries, a CTA in global memory for longer hub rows
// (all three use the same ascending-only bitonic network, so tails never need padding) -- and a second scan + copy
// closes the gaps the duplicates left.  The scatter order inside a row depends on atomics, the result does not.
//
// Phases (E = edges in, N = nodes):                                algorithmic bytes
//   k_csr_degree     histogram of src                              16 E read (+ 4 E atomic)
//   scan64           deg -> raw_off                                 4 N read, 8 N write
//   k_csr_scatter    tmp[raw_off[src] + cursor++] = dst             16 E read, 4 E write
//   k_csr_sort_*     sort + unique per row, in place                4 E read, <= 4 E write
//   scan64           ucount -> indptr                               4 N read, 8 N write
//   k_csr_compact    indices[indptr[r] ..] = tmp[raw_off[r] ..]     4 nnz read + write
#define GRAPES_PDL_GROUP 1
#include "common.cuh"

#define CB_THREADS 256
#define CB_SCAN_THREADS 1024
#define CB_SCAN_ITEMS 8
#define CB_SCAN_TILE (CB_SCAN_THREADS * CB_SCAN_ITEMS)
#define CB_WARP_MAX 256              // rows of 33 .. 256 entries: one warp, a 1 KB strip of shared memory
#define CB_MED_MAX 2048              // rows of 257 .. 2048 entries: 256-thread CTA, 8 KB of shared memory
#define CB_LONG_SMEM 32768           // rows up to 32768 entries are sorted in shared memory (128 KB), longer ones in HBM
#define CB_MED_THREADS 256
#define CB_LONG_THREADS 1024

// ---------------------------------------------------------------------------------------------------------------------
// 64-bit exclusive scan of int32 counts: tile sums -> one CTA scans the tile sums -> tiles rescan with their offset.
// ---------------------------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(CB_SCAN_THREADS) k_scan64_tilesum(const int* __restrict__ cnt, int64_t n,
                                                                     long long* __restrict__ tile_sum) {
    __shared__ long long sm[CB_SCAN_THREADS / 32 + 2];
    const int64_t base = (int64_t)blockIdx.x * CB_SCAN_TILE + (int64_t)threadIdx.x * CB_SCAN_ITEMS;
    long long s = 0;
#pragma unroll
    for (int i = 0; i < CB_SCAN_ITEMS; ++i)
        if (base + i < n) s += cnt[base + i];
    long long total;
    block_scan_excl<long long>(s, sm, &total);
    if (threadIdx.x == 0) tile_sum[blockIdx.x] = total;
}

__global__ void __launch_bounds__(CB_SCAN_THREADS) k_scan64_tiles(long long* __restrict__ tile_sum, int tiles,
                                                                   long long* __restrict__ total_out) {
    __shared__ long long sm[CB_SCAN_THREADS / 32 + 2];
    long long run = 0;
    for (int t0 = 0; t0 < tiles; t0 += CB_SCAN_THREADS) {
        const int t = t0 + threadIdx.x;
        const long long v = t < tiles ? tile_sum[t] : 0;
        long long total;
        const long long ex = block_scan_excl<long long>(v, sm, &total);
        if (t < tiles) tile_sum[t] = run + ex;
        run += total;
    }
    if (threadIdx.x == 0 && total_out) *total_out = run;
}

__global__ void __launch_bounds__(CB_SCAN_THREADS) k_scan64_apply(const int* __restrict__ cnt, int64_t n,
                                                                   const long long* __restrict__ tile_off,
                                                                   long long* __restrict__ out /* [n + 1] */) {
    __shared__ long long sm[CB_SCAN_THREADS / 32 + 2];
    const int64_t base = (int64_t)blockIdx.x * CB_SCAN_TILE + (int64_t)threadIdx.x * CB_SCAN_ITEMS;
    int v[CB_SCAN_ITEMS];
    long long s = 0;
#pragma unroll
    for (int i = 0; i < CB_SCAN_ITEMS; ++i) {
        v[i] = (base + i < n) ? cnt[base + i] : 0;
        s += v[i];
    }
    long long total;
    long long run = block_scan_excl<long long>(s, sm, &total) + tile_off[blockIdx.x];
#pragma unroll
    for (int i = 0; i < CB_SCAN_ITEMS; ++i) {
        if (base + i < n) out[base + i] = run;
        run += v[i];
        if (base + i == n - 1) out[n] = run;
    }
}

static void cb_scan64(const int* cnt, int64_t n, long long* tile_sum, long long* out, long long* total_out,
                      cudaStream_t s) {
    const int tiles = (int)((n + CB_SCAN_TILE - 1) / CB_SCAN_TILE);
    pdl(k_scan64_tilesum, tiles, CB_SCAN_THREADS, 0, s)(cnt, n, tile_sum);
    pdl(k_scan64_tiles, 1, CB_SCAN_THREADS, 0, s)(tile_sum, tiles, total_out);
    pdl(k_scan64_apply, tiles, CB_SCAN_THREADS, 0, s)(cnt, n, tile_sum, out);
    grapes_count_launches(3);
}

// ---------------------------------------------------------------------------------------------------------------------
// counting sort by row
// ---------------------------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(CB_THREADS) k_csr_degree(const int64_t* __restrict__ src, const int64_t* __restrict__ dst,
                                                            int64_t E, int64_t N, int* __restrict__ deg,
                                                            int* __restrict__ err) {
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; e < E; e += stride) {
        const int64_t r = src[e], c = dst[e];
        if (r < 0 || r >= N || c < 0 || c >= N) { *err = 1; continue; }     // scipy: "row/column index exceeds matrix dimensions"
        atomicAdd(&deg[r], 1);
    }
}

__global__ void __launch_bounds__(CB_THREADS) k_csr_scatter(const int64_t* __restrict__ src, const int64_t* __restrict__ dst,
                                                             int64_t E, int64_t N, const long long* __restrict__ raw_off,
                                                             int* __restrict__ cursor, int* __restrict__ tmp) {
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; e < E; e += stride) {
        const int64_t r = src[e], c = dst[e];
        if (r < 0 || r >= N || c < 0 || c >= N) continue;
        const int k = atomicAdd(&cursor[r], 1);
        tmp[raw_off[r] + k] = (int)c;
    }
}

// ---------------------------------------------------------------------------------------------------------------------
// per-row sort + unique.  Short rows: one warp per row, the row in one register per lane.
// ---------------------------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(CB_THREADS) k_csr_sort_short(int64_t N, const long long* __restrict__ raw_off,
                                                                int* __restrict__ tmp, int* __restrict__ ucount,
                                                                int* __restrict__ worklist, int* __restrict__ wl_count) {
    __shared__ int s_rows[CB_THREADS / 32][CB_WARP_MAX];
    const int lane = lane_id();
    int* a = s_rows[threadIdx.x >> 5];
    const int64_t warps = ((int64_t)gridDim.x * blockDim.x) >> 5;
    for (int64_t r = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5; r < N; r += warps) {
        const long long o = raw_off[r];
        const int d = (int)(raw_off[r + 1] - o);
        if (d > CB_WARP_MAX) {                          // medium rows are listed from the front, long rows from the back
            if (lane == 0) {
                if (d <= CB_MED_MAX) worklist[atomicAdd(&wl_count[0], 1)] = (int)r;
                else worklist[N - 1 - atomicAdd(&wl_count[1], 1)] = (int)r;
            }
            continue;
        }
        if (d == 0) { if (lane == 0) ucount[r] = 0; continue; }
        if (d > 32) {
            // 33 .. CB_WARP_MAX entries: the warp sorts the row in its own shared-memory strip (same ascending-only
            // network as cb_bitonic, __syncwarp between stages), then writes the unique values back in order
            for (int i = lane; i < d; i += 32) a[i] = tmp[o + i];
            __syncwarp();
            int P = 64;
            while (P < d) P <<= 1;
            for (int size = 2; size <= P; size <<= 1) {
                const int half = size >> 1;
                for (int p = lane; p < (P >> 1); p += 32) {
                    const int blk = p / half, off = p - blk * half;
                    const int i = blk * size + off, j = blk * size + size - 1 - off;
                    if (j < d) { const int x = a[i], y = a[j]; if (x > y) { a[i] = y; a[j] = x; } }
                }
                __syncwarp();
                for (int st = size >> 2; st > 0; st >>= 1) {
                    for (int p = lane; p < (P >> 1); p += 32) {
                        const int i = 2 * st * (p / st) + (p % st), j = i + st;
                        if (j < d) { const int x = a[i], y = a[j]; if (x > y) { a[i] = y; a[j] = x; } }
                    }
                    __syncwarp();
                }
            }
            int base = 0;
            for (int i0 = 0; i0 < d; i0 += 32) {
                const int i = i0 + lane;
                const int v = i < d ? a[i] : 0;
                const bool keep = i < d && (i == 0 || v != a[i - 1]);
                const unsigned m = __ballot_sync(GRAPES_FULL_MASK, keep);
                if (keep) tmp[o + base + __popc(m & ((1u << lane) - 1u))] = v;
                base += __popc(m);
            }
            if (lane == 0) ucount[r] = base;
            __syncwarp();
            continue;
        }
        int v = lane < d ? tmp[o + lane] : 0x7fffffff;
#pragma unroll
        for (int size = 2; size <= 32; size <<= 1) {
            {   // flip step: partner = lane ^ (size - 1)
                const int p = __shfl_xor_sync(GRAPES_FULL_MASK, v, size - 1);
                const bool lower = (lane & (size - 1)) < (size >> 1);
                v = lower ? min(v, p) : max(v, p);
            }
#pragma unroll
            for (int st = size >> 2; st > 0; st >>= 1) {
                const int p = __shfl_xor_sync(GRAPES_FULL_MASK, v, st);
                v = (lane & st) ? max(v, p) : min(v, p);
            }
        }
        const int prev = __shfl_up_sync(GRAPES_FULL_MASK, v, 1);
        const bool keep = lane < d && (lane == 0 || v != prev);
        const unsigned m = __ballot_sync(GRAPES_FULL_MASK, keep);
        if (keep) tmp[o + __popc(m & ((1u << lane) - 1u))] = v;
        if (lane == 0) ucount[r] = __popc(m);
    }
}

// ascending-only bitonic network over a[0 .. n): every compare-exchange leaves the minimum at the lower index, so the
// virtual +inf padding up to the next power of two never moves and pairs whose partner is >= n are skipped.
template <typename Ptr>
__device__ __forceinline__ void cb_bitonic(Ptr a, int n) {
    int P = 1;
    while (P < n) P <<= 1;
    for (int size = 2; size <= P; size <<= 1) {
        const int half = size >> 1;
        for (int p = threadIdx.x; p < (P >> 1); p += blockDim.x) {
            const int blk = p / half, off = p - blk * half;
            const int i = blk * size + off, j = blk * size + size - 1 - off;
            if (j < n) { const int x = a[i], y = a[j]; if (x > y) { a[i] = y; a[j] = x; } }
        }
        __syncthreads();
        for (int st = size >> 2; st > 0; st >>= 1) {
            for (int p = threadIdx.x; p < (P >> 1); p += blockDim.x) {
                const int i = 2 * st * (p / st) + (p % st), j = i + st;
                if (j < n) { const int x = a[i], y = a[j]; if (x > y) { a[i] = y; a[j] = x; } }
            }
            __syncthreads();
        }
    }
}

// ordered unique of the sorted a[0 .. n) into out[0 ..]; `out` may alias `a` (writes trail the reads).  Returns the count.
template <typename Ptr>
__device__ __forceinline__ int cb_unique(Ptr a, int n, int* out, int* sm_scan) {
    __shared__ int s_last;
    int base = 0;
    for (int i0 = 0; i0 < n; i0 += blockDim.x) {
        const int i = i0 + threadIdx.x;
        int v = 0, keep = 0;
        if (i < n) {
            v = a[i];
            const int pv = (i == 0) ? 0 : ((threadIdx.x == 0) ? s_last : a[i - 1]);
            keep = (i == 0 || v != pv) ? 1 : 0;
        }
        __syncthreads();                                  // every read of this chunk (and of s_last) is done
        if (i < n && (threadIdx.x == blockDim.x - 1 || i == n - 1)) s_last = v;
        int total;
        const int ex = block_scan_excl<int>(keep, sm_scan, &total);
        if (keep) out[base + ex] = v;
        base += total;
        __syncthreads();
    }
    return base;
}

template <int THREADS, int SMEM_INTS, bool LONG>
__global__ void __launch_bounds__(THREADS) k_csr_sort_rows(int64_t N, const long long* __restrict__ raw_off,
                                                            int* __restrict__ tmp, int* __restrict__ ucount,
                                                            const int* __restrict__ worklist,
                                                            const int* __restrict__ wl_count) {
    extern __shared__ int cb_sm[];
    __shared__ int sm_scan[THREADS / 32 + 2];
    const int cnt = wl_count[LONG ? 1 : 0];
    for (int w = blockIdx.x; w < cnt; w += gridDim.x) {
        const int r = LONG ? worklist[N - 1 - w] : worklist[w];
        const long long o = raw_off[r];
        const int d = (int)(raw_off[r + 1] - o);
        int* row = tmp + o;
        int u;
        if (d <= SMEM_INTS) {
            for (int i = threadIdx.x; i < d; i += THREADS) cb_sm[i] = row[i];
            __syncthreads();
            cb_bitonic(cb_sm, d);
            u = cb_unique(cb_sm, d, row, sm_scan);
        } else {
            __syncthreads();
            cb_bitonic(row, d);                           // hub row: in place in HBM (L2-resident for anything realistic)
            u = cb_unique(row, d, row, sm_scan);
        }
        if (threadIdx.x == 0) ucount[r] = u;
        __syncthreads();
    }
}

__global__ void __launch_bounds__(CB_THREADS) k_csr_compact(int64_t N, const long long* __restrict__ raw_off,
                                                             const int* __restrict__ tmp,
                                                             const long long* __restrict__ indptr,
                                                             int* __restrict__ indices) {
    const int lane = lane_id();
    const int64_t warps = ((int64_t)gridDim.x * blockDim.x) >> 5;
    for (int64_t r = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5; r < N; r += warps) {
        const long long o = raw_off[r], p = indptr[r];
        const int u = (int)(indptr[r + 1] - p);
        for (int i = lane; i < u; i += 32) indices[p + i] = tmp[o + i];
    }
}

static inline size_t cb_align(size_t x) { return (x + 255) & ~(size_t)255; }

struct CbLayout {
    size_t deg, ucount, raw_off, tile_sum, worklist, flags, tmp, total;
};
static CbLayout cb_layout(int64_t N, int64_t E) {
    CbLayout L;
    size_t o = 0;
    const size_t tiles = (size_t)((N + CB_SCAN_TILE - 1) / CB_SCAN_TILE) + 1;
    L.deg = o; o += cb_align(sizeof(int) * (size_t)N);
    L.ucount = o; o += cb_align(sizeof(int) * (size_t)N);
    L.raw_off = o; o += cb_align(sizeof(long long) * (size_t)(N + 1));
    L.tile_sum = o; o += cb_align(sizeof(long long) * tiles);
    L.worklist = o; o += cb_align(sizeof(int) * (size_t)N);
    L.flags = o; o += cb_align(sizeof(int) * 4);
    L.tmp = o; o += cb_align(sizeof(int) * (size_t)(E > 0 ? E : 1));
    L.total = o;
    return L;
}

extern "C" {

int64_t grapes_csr_workspace_bytes(int64_t num_nodes, int64_t num_edges) {
    if (num_nodes <= 0 || num_edges < 0) return 0;
    return (int64_t)cb_layout(num_nodes, num_edges).total;
}

int grapes_csr_from_edges(const int64_t* src, const int64_t* dst, int64_t num_edges, int64_t num_nodes, int64_t* indptr,
                          int* indices, int64_t* nnz_dev, int* err_dev, void* workspace, int64_t workspace_bytes,
                          void* stream) {
    GRAPES_REQUIRE(num_nodes > 0 && num_nodes < (1ll << 31), "num_nodes must fit int32");
    GRAPES_REQUIRE(num_edges >= 0, "negative edge count");
    GRAPES_REQUIRE(indptr && nnz_dev && err_dev && workspace && (indices || num_edges == 0), "null argument");
    GRAPES_REQUIRE((src && dst) || num_edges == 0, "null edge list");
    const int64_t N = num_nodes, E = num_edges;
    const CbLayout L = cb_layout(N, E);
    GRAPES_REQUIRE(workspace_bytes >= (int64_t)L.total, "workspace smaller than grapes_csr_workspace_bytes()");
    int dev = 0, sms = 0;
    GRAPES_CUDA_OK(cudaGetDevice(&dev));
    GRAPES_CUDA_OK(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
    cudaStream_t s = (cudaStream_t)stream;
    unsigned char* ws = (unsigned char*)workspace;
    int* deg = (int*)(ws + L.deg);
    int* ucount = (int*)(ws + L.ucount);
    long long* raw_off = (long long*)(ws + L.raw_off);
    long long* tile_sum = (long long*)(ws + L.tile_sum);
    int* worklist = (int*)(ws + L.worklist);
    int* flags = (int*)(ws + L.flags);                    // [0] medium rows, [1] long rows
    int* tmp = (int*)(ws + L.tmp);
    GRAPES_CUDA_OK(cudaMemsetAsync(deg, 0, sizeof(int) * (size_t)N, s));
    GRAPES_CUDA_OK(cudaMemsetAsync(flags, 0, sizeof(int) * 4, s));
    GRAPES_CUDA_OK(cudaMemsetAsync(err_dev, 0, sizeof(int), s));
    const int edge_grid = (int)(E > 0 ? ((E + CB_THREADS - 1) / CB_THREADS < (int64_t)sms * 16 ? (E + CB_THREADS - 1) / CB_THREADS
                                                                                                : (int64_t)sms * 16)
                                      : 1);
    const int64_t row_ctas = (N * 32 + CB_THREADS - 1) / CB_THREADS;
    const int row_grid = (int)(row_ctas < (int64_t)sms * 16 ? row_ctas : (int64_t)sms * 16);
    if (E > 0) {
        pdl(k_csr_degree, edge_grid, CB_THREADS, 0, s)(src, dst, E, N, deg, err_dev);
        grapes_count_launches(1);
    }
    cb_scan64(deg, N, tile_sum, raw_off, nullptr, s);
    GRAPES_CUDA_OK(cudaMemsetAsync(deg, 0, sizeof(int) * (size_t)N, s));          // deg becomes the scatter cursor
    if (E > 0) {
        pdl(k_csr_scatter, edge_grid, CB_THREADS, 0, s)(src, dst, E, N, raw_off, deg, tmp);
        grapes_count_launches(1);
    }
    pdl(k_csr_sort_short, row_grid, CB_THREADS, 0, s)(N, raw_off, tmp, ucount, worklist, flags);
    {
        static bool configured = false;
        if (!configured) {
            GRAPES_CUDA_OK(cudaFuncSetAttribute(k_csr_sort_rows<CB_LONG_THREADS, CB_LONG_SMEM, true>,
                                                cudaFuncAttributeMaxDynamicSharedMemorySize, CB_LONG_SMEM * 4));
            configured = true;
        }
    }
    pdl((k_csr_sort_rows<CB_MED_THREADS, CB_MED_MAX, false>), sms * 8, CB_MED_THREADS, CB_MED_MAX * 4, s)(
        N, raw_off, tmp, ucount, worklist, flags);
    pdl((k_csr_sort_rows<CB_LONG_THREADS, CB_LONG_SMEM, true>), sms, CB_LONG_THREADS, CB_LONG_SMEM * 4, s)(
        N, raw_off, tmp, ucount, worklist, flags);
    grapes_count_launches(3);
    cb_scan64(ucount, N, tile_sum, (long long*)indptr, (long long*)nnz_dev, s);
    pdl(k_csr_compact, row_grid, CB_THREADS, 0, s)(N, raw_off, tmp, (const long long*)indptr, indices);
    grapes_count_launches(1);
    GRAPES_LAUNCH_OK();
    return GRAPES_OK;
}

}  // extern "C"
