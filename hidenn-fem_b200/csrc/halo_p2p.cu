// Halo exchange over NVLink peer memory (multi-GPU): every rank's receive buffer lives in symmetric
// memory that all ranks of the box map, so the gradients of shared nodes and the loss partials are PUT straight into
// the peers' buffers by this rank's kernels (P2P stores over NVLink / NVSwitch) and completed by a flag per sender --
// no NCCL call, no host round trip, capturable in a CUDA graph.  Sums run in ascending rank order on every holder, so
// all holders of a node end with bit-identical gradients.
//
// Layout of every rank's buffer, in units of the real type R (smax = max shared nodes between any two ranks):
//   data  [2 parities][world senders][smax][4]      (gx.x, gx.y, gu.x, gu.y) of the k-th node the two ranks share
//   loss  [2 parities][world senders][4]            (loss, domain, edge, 0) partial of the sender
//   flags uint64 [2 channels][world senders]        step number the sender has completed on that channel
// The step counter lives in the caller's device memory and is advanced by the loss kernel, so a captured graph replays
// with fresh step numbers.
#include "../../include/hidenn_b200.h"
#include "common.cuh"
#include "halo_p2p.cuh"

namespace hidenn {

constexpr int kP2PBlock = 1024;

// one CTA: write the partial gradients of my shared nodes into every other holder's buffer, then raise my flag there
template <typename R>
__global__ void __launch_bounds__(kP2PBlock)
halo_p2p_push_kernel(const typename Real2<R>::type* __restrict__ gx, const typename Real2<R>::type* __restrict__ gu,
                     const int32_t* __restrict__ s_xrow, const int32_t* __restrict__ s_urow, const int32_t* __restrict__ s_peer,
                     const int32_t* __restrict__ s_k, const long long n_send, unsigned char* const* __restrict__ peer_bufs, const int me,
                     const int world, const long long smax, const unsigned long long* __restrict__ step_ptr, unsigned* wait_counter,
                     const unsigned wait_target) {
    using R2 = typename Real2<R>::type;
    if (wait_counter != nullptr) {      // the tile kernel on the other stream counts its finished shared tiles
        if (threadIdx.x == 0) {
            unsigned v;
            do { asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(wait_counter) : "memory"); } while (v < wait_target);
            *wait_counter = 0u;         // every increment of this step has happened: ready for the next one
        }
        __syncthreads();
    }
    const unsigned long long step = *step_ptr;
    const long long par = (long long)(step & 1ull);
    const P2PLayout L = p2p_layout<R>(world, smax);
    for (long long i = threadIdx.x; i < n_send; i += kP2PBlock) {
        const int xr = s_xrow[i], ur = s_urow[i];
        const R2 a = xr >= 0 ? __ldcg(gx + xr) : mk2<R>(R(0), R(0));
        const R2 b = ur >= 0 ? __ldcg(gu + ur) : mk2<R>(R(0), R(0));
        R* dst = reinterpret_cast<R*>(peer_bufs[s_peer[i]]) + L.data0 + ((par * world + me) * smax + s_k[i]) * 4;
        reinterpret_cast<R2*>(dst)[0] = a;
        reinterpret_cast<R2*>(dst)[1] = b;
    }
    __threadfence_system();
    __syncthreads();
    if ((int)threadIdx.x < world && (int)threadIdx.x != me) {
        unsigned long long* f = reinterpret_cast<unsigned long long*>(peer_bufs[threadIdx.x] + L.flags_bytes0) + (0 * world + me);
        st_release_sys(f, step);
    }
}

// wait for the senders, then complete my shared nodes: sum of all holders' partials in ascending rank order
template <typename R>
__global__ void __launch_bounds__(kP2PBlock)
halo_p2p_pull_kernel(typename Real2<R>::type* __restrict__ gx, typename Real2<R>::type* __restrict__ gu, const int32_t* __restrict__ n_xrow,
                     const int32_t* __restrict__ n_urow, const int32_t* __restrict__ n_off, const int32_t* __restrict__ src_rank,
                     const int32_t* __restrict__ src_k, const long long n_nodes, const int32_t* __restrict__ wait_ranks, const int n_wait,
                     unsigned char* __restrict__ my_buf, const int me, const int world, const long long smax,
                     unsigned long long* __restrict__ step_ptr) {
    using R2 = typename Real2<R>::type;
    const unsigned long long step = *step_ptr;
    const long long par = (long long)(step & 1ull);
    const P2PLayout L = p2p_layout<R>(world, smax);
    if ((int)threadIdx.x < n_wait) {
        const unsigned long long* f = reinterpret_cast<const unsigned long long*>(my_buf + L.flags_bytes0) + (0 * world + wait_ranks[threadIdx.x]);
        while (ld_acquire_sys(f) < step) {}
    }
    __syncthreads();
    const R* data = reinterpret_cast<const R*>(my_buf) + L.data0;
    for (long long j = threadIdx.x; j < n_nodes; j += kP2PBlock) {
        const int xr = n_xrow[j], ur = n_urow[j];
        const R2 ownx = xr >= 0 ? gx[xr] : mk2<R>(R(0), R(0));
        const R2 ownu = ur >= 0 ? gu[ur] : mk2<R>(R(0), R(0));
        R ax = R(0), ay = R(0), bx = R(0), by = R(0);
        for (int s = n_off[j]; s < n_off[j + 1]; ++s) {
            const int q = src_rank[s];
            if (q == me) {
                ax += ownx.x; ay += ownx.y; bx += ownu.x; by += ownu.y;
            } else {
                const volatile R* v = data + ((par * world + q) * smax + src_k[s]) * 4;
                ax += v[0]; ay += v[1]; bx += v[2]; by += v[3];
            }
        }
        if (xr >= 0) gx[xr] = mk2<R>(ax, ay);
        if (ur >= 0) gu[ur] = mk2<R>(bx, by);
    }
    __syncthreads();
    if (threadIdx.x == 0) *step_ptr = step + 1;      // the gradient channel has its own step counter (the loss channel another)
}

// loss exchange as a kernel of its own (the tile kernel can also run it in its tail: hidenn_tri_energy_overlap_*)
template <typename R>
__global__ void __launch_bounds__(64)
halo_p2p_loss_kernel(R* __restrict__ out, const P2PLossArgs A) {
    p2p_loss_exchange<R>(out, A, (int)threadIdx.x, [] { __syncthreads(); });
}

}  // namespace hidenn

using namespace hidenn;

extern "C" int64_t hidenn_halo_p2p_bytes(int world, int64_t smax, int real_bytes) {
    const P2PLayout L = real_bytes == 8 ? p2p_layout<double>(world, smax) : p2p_layout<float>(world, smax);
    return L.flags_bytes0 + 2LL * world * 8;
}

#define HIDENN_P2P_API(SUF, T)                                                                                                          \
    extern "C" int hidenn_halo_p2p_push_##SUF(const T* gx, const T* gu, const int32_t* s_xrow, const int32_t* s_urow, const int32_t* s_peer, \
                                              const int32_t* s_k, int64_t n_send, void* const* peer_bufs, int me, int world, int64_t smax,  \
                                              const uint64_t* step, uint32_t* wait_counter, uint32_t wait_target, void* stream) {        \
        HIDENN_REQUIRE(peer_bufs && step && world >= 1 && world <= 64, "halo_p2p_push: bad arguments");                                 \
        using T2 = Real2<T>::type;                                                                                                     \
        halo_p2p_push_kernel<T><<<1, kP2PBlock, 0, (cudaStream_t)stream>>>((const T2*)gx, (const T2*)gu, s_xrow, s_urow, s_peer, s_k, n_send, \
                                                                           (unsigned char* const*)peer_bufs, me, world, smax,          \
                                                                           (const unsigned long long*)step, wait_counter, wait_target); \
        HIDENN_CUDA_OK(cudaGetLastError());                                                                                            \
        return 0;                                                                                                                      \
    }                                                                                                                                  \
    extern "C" int hidenn_halo_p2p_pull_##SUF(T* gx, T* gu, const int32_t* n_xrow, const int32_t* n_urow, const int32_t* n_off,         \
                                              const int32_t* src_rank, const int32_t* src_k, int64_t n_nodes, const int32_t* wait_ranks, \
                                              int n_wait, void* my_buf, int me, int world, int64_t smax, uint64_t* step,                \
                                              void* stream) {                                                                          \
        HIDENN_REQUIRE(my_buf && step && n_wait <= kP2PBlock, "halo_p2p_pull: bad arguments");                                          \
        using T2 = Real2<T>::type;                                                                                                     \
        halo_p2p_pull_kernel<T><<<1, kP2PBlock, 0, (cudaStream_t)stream>>>((T2*)gx, (T2*)gu, n_xrow, n_urow, n_off, src_rank, src_k, n_nodes, \
                                                                           wait_ranks, n_wait, (unsigned char*)my_buf, me, world, smax, \
                                                                           (unsigned long long*)step);                                 \
        HIDENN_CUDA_OK(cudaGetLastError());                                                                                            \
        return 0;                                                                                                                      \
    }                                                                                                                                  \
    extern "C" int hidenn_halo_p2p_loss_##SUF(T* out, void* const* peer_bufs, void* my_buf, int me, int world, int64_t smax,            \
                                              uint64_t* step, void* stream) {                                                          \
        HIDENN_REQUIRE(out && peer_bufs && my_buf && step && world <= 64, "halo_p2p_loss: bad arguments");                              \
        const P2PLossArgs A{(unsigned char* const*)peer_bufs, (unsigned char*)my_buf, (unsigned long long*)step, (long long)smax, me, world}; \
        halo_p2p_loss_kernel<T><<<1, 64, 0, (cudaStream_t)stream>>>(out, A);                                                            \
        HIDENN_CUDA_OK(cudaGetLastError());                                                                                            \
        return 0;                                                                                                                      \
    }
HIDENN_P2P_API(f64, double)
HIDENN_P2P_API(f32, float)
