// Peer-memory halo exchange: buffer layout and the loss exchange shared by halo_p2p.cu and the tile kernel's tail.
#pragma once
#include "common.cuh"

namespace hidenn {

__device__ __forceinline__ void st_release_sys(unsigned long long* p, unsigned long long v) {
    asm volatile("st.release.sys.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}
__device__ __forceinline__ unsigned long long ld_acquire_sys(const unsigned long long* p) {
    unsigned long long v;
    asm volatile("ld.acquire.sys.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
    return v;
}

struct P2PLayout {
    long long data0, loss0, flags_bytes0;      // offsets: data / loss in reals, flags in bytes
};
template <typename R> __host__ __device__ inline P2PLayout p2p_layout(int world, long long smax) {
    P2PLayout L;
    L.data0 = 0;
    L.loss0 = 2LL * world * smax * 4;
    const long long reals = L.loss0 + 2LL * world * 4;
    L.flags_bytes0 = (reals * (long long)sizeof(R) + 15) / 16 * 16;
    return L;
}


// loss = sum over ranks (ascending) of the rank partials: the first `world` threads of the calling group put out[0..2]
// into every peer's buffer, raise their flag, wait for the peers' flags; `leader` then adds the partials in rank order,
// writes out[0..2] and advances *step_ptr.  sync() must be a barrier over the calling threads.
struct P2PLossArgs {
    unsigned char* const* peer_bufs;      // NULL: no exchange
    unsigned char* my_buf;
    unsigned long long* step_ptr;
    long long smax;
    int me, world;
};
template <typename R, typename Sync>
__device__ __forceinline__ void p2p_loss_exchange(R* out, const P2PLossArgs& A, const int t, Sync sync) {
    const unsigned long long step = *A.step_ptr;
    const long long par = (long long)(step & 1ull);
    const P2PLayout L = p2p_layout<R>(A.world, A.smax);
    if (t < A.world && t != A.me) {
        R* dst = reinterpret_cast<R*>(A.peer_bufs[t]) + L.loss0 + (par * A.world + A.me) * 4;
        dst[0] = out[0]; dst[1] = out[1]; dst[2] = out[2]; dst[3] = R(0);
        // the release store orders this thread's four data stores before the flag: no separate system fence
        st_release_sys(reinterpret_cast<unsigned long long*>(A.peer_bufs[t] + L.flags_bytes0) + (1 * A.world + A.me), step);
        const unsigned long long* f = reinterpret_cast<const unsigned long long*>(A.my_buf + L.flags_bytes0) + (1 * A.world + t);
        while (ld_acquire_sys(f) < step) {}
    }
    sync();
    if (t == 0) {
        const volatile R* loss = reinterpret_cast<const R*>(A.my_buf) + L.loss0 + par * A.world * 4;
        R a = R(0), b = R(0), c = R(0);
        for (int r = 0; r < A.world; ++r) {
            if (r == A.me) { a += out[0]; b += out[1]; c += out[2]; }
            else { a += loss[r * 4]; b += loss[r * 4 + 1]; c += loss[r * 4 + 2]; }
        }
        out[0] = a; out[1] = b; out[2] = c;
        *A.step_ptr = step + 1;
    }
}

}  // namespace hidenn
