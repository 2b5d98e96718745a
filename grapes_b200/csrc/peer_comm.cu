// Gradient all-reduce (mean) + both Adam updates as TWO launches over NVLink peer memory -- the only exchange step of
// the data-parallel path (one process per GPU, replicated graph + features, targets sharded; the reference is
// single-process, so this replaces nothing in /root/reference: it is the `optimizer_c.step(); optimizer_gf.step()` of
// main.py:268,289 applied to the mean of the per-rank gradients).
//
// Every rank owns one symmetric buffer (torch.distributed._symmetric_memory gives each rank the peers' device
// pointers): [slot 0 | slot 1 | flags].  Per step, sequence number s = *seq + 1:
//   k_peer_publish      copies the rank's flat gradient into its slot (s & 1); the last block to finish stores s into
//                       flags[rank] of EVERY peer (st.release.sys) -- the data is then visible system-wide.
//   k_peer_reduce_adam  waits until flags[r] >= s for all r (ld.acquire.sys, bounded spin -> error flag, never a hang),
//                       reads the W slots over NVLink, sums them in rank order (identical bits on every rank, run-to-run
//                       deterministic), divides by W, applies both Adam groups and writes the mean gradient back.
// Two slots suffice: a rank can publish step s + 2 only after its reduce of step s + 1, which needed every peer's
// publish of s + 1, which follows that peer's reduce of step s in stream order -- so nobody still reads slot s & 1.
// Both kernels take no per-step host argument (the sequence number lives on the device), so they are captured into the
// step's CUDA graph: the N > 1 step is ONE graph launch, no host-side collective call.
#define GRAPES_PDL_GROUP 8
#include "common.cuh"

#define PEER_MAX 8
struct PeerTable {
    float* buf[PEER_MAX];              // base of every rank's symmetric buffer (peer-mapped device pointers)
};

__device__ __forceinline__ void st_release_sys_u32(unsigned int* p, unsigned int v) {
    asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
__device__ __forceinline__ unsigned int ld_acquire_sys_u32(const unsigned int* p) {
    unsigned int v;
    asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ float ld_relaxed_sys_f32(const float* p) {
    float v;
    asm volatile("ld.relaxed.sys.global.f32 %0, [%1];" : "=f"(v) : "l"(p) : "memory");
    return v;
}

// layout of a symmetric buffer, in floats: [0, n_pad) slot 0, [n_pad, 2 n_pad) slot 1, then PEER_MAX flag words
__global__ void __launch_bounds__(256) k_peer_publish(const float* __restrict__ grads, int n, int n_pad, PeerTable pt,
                                                      int rank, int world, const unsigned int* __restrict__ seq_dev,
                                                      unsigned int* ticket) {
    pdl_begin();
    __shared__ int s_last;
    const unsigned int seq = *seq_dev + 1u;
    float* dst = pt.buf[rank] + (size_t)(seq & 1u) * n_pad;
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) dst[i] = grads[i];
    __threadfence_system();
    __syncthreads();
    if (threadIdx.x == 0) s_last = (atomicAdd(ticket, 1u) == gridDim.x - 1) ? 1 : 0;
    __syncthreads();
    if (s_last) {
        __threadfence_system();
        if (threadIdx.x < world) {
            unsigned int* flags = reinterpret_cast<unsigned int*>(pt.buf[threadIdx.x] + 2 * (size_t)n_pad);
            st_release_sys_u32(flags + rank, seq);
        }
        if (threadIdx.x == 0) *ticket = 0u;
    }
}

__global__ void __launch_bounds__(256) k_peer_reduce_adam(
    PeerTable pt, int rank, int world, int n, int n_pad, unsigned int* seq_dev, float* __restrict__ p,
    float* __restrict__ g_out, float* __restrict__ m, float* __restrict__ v, int off0, int n0, float lr0, int off1, int n1,
    float lr1, float beta1, float beta2, float eps, float* steps, unsigned int* ticket, int* err_flag,
    unsigned int spin_limit) {
    pdl_begin();
    __shared__ float s_step_size[2], s_bc2_sqrt[2];
    __shared__ int s_last;
    const unsigned int seq = *seq_dev + 1u;
    if (threadIdx.x < world) {
        const unsigned int* flags = reinterpret_cast<const unsigned int*>(pt.buf[rank] + 2 * (size_t)n_pad);
        unsigned int it = 0;
        while (ld_acquire_sys_u32(flags + threadIdx.x) < seq) {
            if (++it > spin_limit) { atomicOr(err_flag, GRAPES_OVF_PEER_TIMEOUT); break; }   // never hang the GPU
            __nanosleep(64);
        }
    }
    if (threadIdx.x >= 32 && threadIdx.x < 34) {
        const int q = threadIdx.x - 32;
        const double t = (double)steps[q] + 1.0;
        const double bc1 = 1.0 - pow((double)beta1, t), bc2 = 1.0 - pow((double)beta2, t);
        s_step_size[q] = (float)((double)(q ? lr1 : lr0) / bc1);
        s_bc2_sqrt[q] = (float)sqrt(bc2);
    }
    __syncthreads();
    const size_t slot = (size_t)(seq & 1u) * n_pad;
    const float inv_w = 1.0f / (float)world;
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
        float s = 0.f;
        for (int r = 0; r < world; ++r) s += ld_relaxed_sys_f32(pt.buf[r] + slot + i);     // rank order: same bits everywhere
        const float gi = s * inv_w;
        g_out[i] = gi;
        int grp = -1;
        if (i >= off0 && i < off0 + n0) grp = 0;
        else if (i >= off1 && i < off1 + n1) grp = 1;
        if (grp >= 0) {                                                  // same arithmetic as k_adam2
            const float mi = m[i] + (gi - m[i]) * (1.0f - beta1);
            const float vi = v[i] * beta2 + (1.0f - beta2) * gi * gi;
            m[i] = mi; v[i] = vi;
            const float denom = sqrtf(vi) / s_bc2_sqrt[grp] + eps;
            p[i] = p[i] - s_step_size[grp] * (mi / denom);
        }
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        __threadfence();
        s_last = (atomicAdd(ticket, 1u) == gridDim.x - 1) ? 1 : 0;
    }
    __syncthreads();
    if (s_last && threadIdx.x == 0) {
        if (n0 > 0) steps[0] += 1.0f;
        if (n1 > 0) steps[1] += 1.0f;
        *seq_dev = seq;
        *ticket = 0u;
    }
}

extern "C" {

int64_t grapes_peer_buffer_floats(int n) {
    const int64_t n_pad = ((int64_t)n + 63) / 64 * 64;
    return 2 * n_pad + 64;
}

// peer_bufs: HOST array of `world` device pointers (the peers' symmetric buffers, own buffer at index `rank`), each of
// grapes_peer_buffer_floats(n) floats, zero-initialised once.  state: 4 uint32 on the device, zero-initialised
// ([0] sequence number, [1] publish ticket, [2] reduce ticket).  err_flag: the engine's overflow word.
int grapes_allreduce_adam_peer(grapes_ctx* ctx, void* const* peer_bufs, int rank, int world, const float* grads, int n,
                               float* params, float* grads_mean_out, float* exp_avg, float* exp_avg_sq, int off0,
                               int n0, float lr0, int off1, int n1, float lr1, float beta1, float beta2, float eps,
                               float* steps_dev, unsigned int* state, int* err_flag, void* stream) {
    GRAPES_REQUIRE(ctx && peer_bufs && grads && params && grads_mean_out && exp_avg && exp_avg_sq && steps_dev && state &&
                       err_flag, "null argument");
    GRAPES_REQUIRE(world >= 1 && world <= PEER_MAX && rank >= 0 && rank < world, "1 <= world <= 8");
    GRAPES_REQUIRE(n > 0, "empty gradient");
    PeerTable pt;
    for (int r = 0; r < PEER_MAX; ++r) pt.buf[r] = (r < world) ? reinterpret_cast<float*>(peer_bufs[r]) : nullptr;
    const int n_pad = (n + 63) / 64 * 64;
    int blocks = (n + 255) / 256;
    if (blocks > ctx->sm_count * 2) blocks = ctx->sm_count * 2;
    cudaStream_t s = (cudaStream_t)stream;
    pdl((k_peer_publish), blocks, 256, 0, s)(grads, n, n_pad, pt, rank, world, state, state + 1);
    grapes_count_launches(1);
    // ~2 s of polling at 64 ns per try before a rank gives up on a peer and raises the error flag
    pdl((k_peer_reduce_adam), blocks, 256, 0, s)(pt, rank, world, n, n_pad, state, params, grads_mean_out, exp_avg,
                                                 exp_avg_sq, off0, n0, lr0, off1, n1, lr1, beta1, beta2, eps, steps_dev,
                                                 state + 2, err_flag, 30000000u);
    grapes_count_launches(1);
    GRAPES_LAUNCH_OK();
    return GRAPES_OK;
}

}  // extern "C"
