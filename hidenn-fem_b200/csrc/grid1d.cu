// 1D model kernels: r-adaptive grid (scan), element lookup, interpolation + VJP, deterministic folds and the
// fused bar-energy forward+backward.  Reference: /root/reference/src/models.py:6-90 and
// /root/reference/examples/example3.py:16-70.  HBM/L2-bound scans and gathers: CUDA cores only.
#include "../../include/hidenn_b200_grid.h"
#include "common.cuh"

namespace hidenn {

constexpr int kScanThreads = 256;
constexpr int kScanItems = 8;
constexpr int kScanChunk = kScanThreads * kScanItems;   // items per block

template <typename R> __device__ __forceinline__ R softplus_inc(R p) {
    // torch.clamp(F.softplus(p), min=1e-6): softplus with beta=1, threshold=20
    const R sp = p > R(20) ? p : log1p(exp(p));
    return sp < R(1e-6) ? R(1e-6) : sp;
}
template <typename R> __device__ __forceinline__ R softplus_raw(R p) { return p > R(20) ? p : log1p(exp(p)); }

// exclusive block scan of one value per thread (fixed order: lanes, then warps); returns exclusive prefix,
// total in *total (all threads)
template <typename R> __device__ __forceinline__ R block_excl_scan(R v, R* s_warp /*[kScanThreads/32 + 1]*/, R* total) {
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    R inc = v;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        const R t = __shfl_up_sync(0xffffffffu, inc, o);
        if (lane >= o) inc += t;
    }
    if (lane == 31) s_warp[w] = inc;
    __syncthreads();
    if (w == 0) {
        R t = lane < kScanThreads / 32 ? s_warp[lane] : R(0);
        R ti = t;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const R u = __shfl_up_sync(0xffffffffu, ti, o);
            if (lane >= o) ti += u;
        }
        if (lane < kScanThreads / 32) s_warp[lane] = ti - t;      // exclusive warp offsets
        if (lane == kScanThreads / 32 - 1) s_warp[kScanThreads / 32] = ti;
    }
    __syncthreads();
    *total = s_warp[kScanThreads / 32];
    const R r = s_warp[w] + inc - v;
    __syncthreads();
    return r;
}

// inclusive prefix of item i as every kernel defines it: the running sum inside a block, the scanned offset of the next
// block at a block's last item (partial[nb] = S at the last item of the array)
template <typename R>
__device__ __forceinline__ R canonical_cum(R run, int64_t i, int64_t n, const R* __restrict__ partial, int64_t b) {
    return (i == n - 1 || ((i + 1) % kScanChunk) == 0) ? partial[b + 1] : run;
}

// ---- phase a: per-block sums of up to two channels -----------------------------------------------------
// MODE 0: channel0 = inc(p)                       (grid forward)
// MODE 1: channel0 = gamma = dgrid[i+1], channel1 = gamma*cum[i]     (grid backward)
template <typename R, int MODE>
__global__ void __launch_bounds__(kScanThreads)
scan_block_sums_kernel(const R* __restrict__ a, const R* __restrict__ b, int64_t n, R* __restrict__ partial, int64_t nb) {
    __shared__ R s_warp[kScanThreads / 32 + 1];
    const int64_t base = (int64_t)blockIdx.x * kScanChunk + (int64_t)threadIdx.x * kScanItems;
    R s0 = R(0), s1 = R(0);
#pragma unroll
    for (int k = 0; k < kScanItems; ++k) {
        const int64_t i = base + k;
        if (i < n) {
            if (MODE == 0) s0 += softplus_inc<R>(a[i]);
            else { const R g = a[i + 1]; s0 += g; s1 += g * b[i]; }
        }
    }
    R tot;
    block_excl_scan<R>(s0, s_warp, &tot);
    if (threadIdx.x == 0) partial[blockIdx.x] = tot;
    if (MODE == 1) {
        block_excl_scan<R>(s1, s_warp, &tot);
        if (threadIdx.x == 0) partial[nb + 1 + blockIdx.x] = tot;
    }
}

// ---- phase b: one block turns the per-block sums into exclusive offsets; totals at [nb] (and [2nb+1]) ---
template <typename R, int NCH>
__global__ void __launch_bounds__(kScanThreads) scan_partials_kernel(R* __restrict__ partial, int64_t nb) {
    __shared__ R s_warp[kScanThreads / 32 + 1];
    for (int ch = 0; ch < NCH; ++ch) {
        R* p = partial + ch * (nb + 1);
        R carry = R(0);
        for (int64_t base = 0; base < nb; base += kScanThreads) {
            const int64_t i = base + threadIdx.x;
            const R v = i < nb ? p[i] : R(0);
            R tot;
            const R ex = block_excl_scan<R>(v, s_warp, &tot);
            if (i < nb) p[i] = carry + ex;
            carry += tot;
        }
        if (threadIdx.x == 0) p[nb] = carry;
        __syncthreads();
    }
}

// ---- phase c (forward): grid and cum -----------------------------------------------------------------
template <typename R>
__global__ void __launch_bounds__(kScanThreads)
grid_fwd_final_kernel(const R* __restrict__ p, int64_t n, const R* __restrict__ x0p, const R* __restrict__ xNp,
                      const R* __restrict__ partial, int64_t nb, R* __restrict__ grid, R* __restrict__ cum) {
    __shared__ R s_warp[kScanThreads / 32 + 1];
    const int64_t base = (int64_t)blockIdx.x * kScanChunk + (int64_t)threadIdx.x * kScanItems;
    R v[kScanItems];
    R s = R(0);
#pragma unroll
    for (int k = 0; k < kScanItems; ++k) {
        const int64_t i = base + k;
        v[k] = i < n ? softplus_inc<R>(p[i]) : R(0);
        s += v[k];
    }
    R tot;
    R run = block_excl_scan<R>(s, s_warp, &tot) + partial[blockIdx.x];
    const R S = partial[nb], x0 = *x0p, L = *xNp - *x0p;
    if (blockIdx.x == 0 && threadIdx.x == 0) grid[0] = x0;
#pragma unroll
    for (int k = 0; k < kScanItems; ++k) {
        const int64_t i = base + k;
        run += v[k];
        if (i < n) {
            // canonical value at the end of a block (and of the array): the scanned offset of the next block, so that
            // every kernel that recomputes the prefix (bar_step_energy_kernel) sees the same bits, and cum[n-1] == S
            const R c = canonical_cum<R>(run, i, n, partial, blockIdx.x);
            cum[i] = c;
            grid[i + 1] = x0 + L * c / S;
        }
    }
}

// ---- phase c (backward): dp = (c*suffix_gamma - L*T/S^2) * [sp>=1e-6] * sigmoid(p) ----------------------
template <typename R>
__global__ void __launch_bounds__(kScanThreads)
grid_bwd_final_kernel(const R* __restrict__ dgrid, const R* __restrict__ p, const R* __restrict__ cum, int64_t n,
                      const R* __restrict__ x0p, const R* __restrict__ xNp, const R* __restrict__ partial, int64_t nb,
                      R* __restrict__ dp) {
    __shared__ R s_warp[kScanThreads / 32 + 1];
    const int64_t base = (int64_t)blockIdx.x * kScanChunk + (int64_t)threadIdx.x * kScanItems;
    R g[kScanItems];
    R s = R(0);
#pragma unroll
    for (int k = 0; k < kScanItems; ++k) {
        const int64_t i = base + k;
        g[k] = i < n ? dgrid[i + 1] : R(0);
        s += g[k];
    }
    R tot;
    R run = block_excl_scan<R>(s, s_warp, &tot) + partial[blockIdx.x];     // exclusive prefix of gamma
    const R G = partial[nb], T = partial[2 * nb + 1];
    const R S = cum[n - 1], L = *xNp - *x0p;
    const R c = L / S, corr = L * T / (S * S);
#pragma unroll
    for (int k = 0; k < kScanItems; ++k) {
        const int64_t i = base + k;
        if (i < n) {
            const R suffix = G - run;             // sum_{k>=i} gamma_k
            const R dinc = c * suffix - corr;
            const R pi = p[i];
            const R sp = softplus_raw<R>(pi);
            const R dsp = pi > R(20) ? R(1) : R(1) / (R(1) + exp(-pi));
            dp[i] = sp >= R(1e-6) ? dinc * dsp : R(0);
        }
        run += g[k];
    }
}

// ---- element lookup ------------------------------------------------------------------------------------
// searchsorted(grid, x) with right=False: first index i with grid[i] >= x; element = clamp(i-1, 0, N-2)
template <typename R> __device__ __forceinline__ int lookup_elem(const R* __restrict__ grid, int64_t N, R x) {
    int64_t lo = 0, hi = N;
    while (lo < hi) {
        const int64_t mid = (lo + hi) >> 1;
        if (__ldg(grid + mid) < x) lo = mid + 1; else hi = mid;     // NaN x: comparison false -> hi shrinks -> 0, as torch
    }
    int64_t e = lo - 1;
    e = e < 0 ? 0 : (e > N - 2 ? N - 2 : e);
    return (int)e;
}

template <typename R>
__global__ void __launch_bounds__(256)
lookup_kernel(const R* __restrict__ grid, int64_t N, const R* __restrict__ x, int64_t M, int32_t* __restrict__ elem) {
    for (int64_t m = (int64_t)blockIdx.x * 256 + threadIdx.x; m < M; m += (int64_t)gridDim.x * 256) elem[m] = lookup_elem<R>(grid, N, x[m]);
}

template <typename R>
__global__ void __launch_bounds__(256)
interp_fwd_kernel(const R* __restrict__ grid, int64_t N, const R* __restrict__ uf, const R* __restrict__ x, int64_t M,
                  R* __restrict__ u, int32_t* __restrict__ elem, R* __restrict__ slope) {
    for (int64_t m = (int64_t)blockIdx.x * 256 + threadIdx.x; m < M; m += (int64_t)gridDim.x * 256) {
        const R xv = x[m];
        const int e = lookup_elem<R>(grid, N, xv);
        const R ge = __ldg(grid + e), gp = __ldg(grid + e + 1), ue = __ldg(uf + e), up = __ldg(uf + e + 1);
        R h = gp - ge;
        h = h < R(1e-10) ? R(1e-10) : h;
        const R N1 = (gp - xv) / h, N2 = (xv - ge) / h;
        u[m] = ue * N1 + up * N2;
        if (elem) elem[m] = e;
        if (slope) slope[m] = (up - ue) / h;
    }
}

template <typename R>
__global__ void __launch_bounds__(256)
interp_bwd_kernel(const R* __restrict__ grid, const R* __restrict__ uf, const R* __restrict__ x, const int32_t* __restrict__ elem,
                  const R* __restrict__ r_u, const R* __restrict__ r_s, int64_t M, R* __restrict__ rows, R* __restrict__ dx) {
    for (int64_t m = (int64_t)blockIdx.x * 256 + threadIdx.x; m < M; m += (int64_t)gridDim.x * 256) {
        const int e = elem[m];
        const R xv = x[m];
        const R ge = __ldg(grid + e), gp = __ldg(grid + e + 1), ue = __ldg(uf + e), up = __ldg(uf + e + 1);
        const R hraw = gp - ge;
        const bool act = hraw >= R(1e-10);
        const R h = act ? hraw : R(1e-10);
        const R ru = r_u ? r_u[m] : R(0), rs = r_s ? r_s[m] : R(0);
        const R N1 = (gp - xv) / h, N2 = (xv - ge) / h;
        const R num = ue * (gp - xv) + up * (xv - ge);
        const R ih = R(1) / h;
        const R q = act ? num * ih * ih : R(0);              // d/dh through the clamp
        const R sl = act ? (up - ue) * ih * ih : R(0);
        rows[4 * m + 0] = ru * N1 - rs * ih;
        rows[4 * m + 1] = ru * N2 + rs * ih;
        rows[4 * m + 2] = ru * (-up * ih + q) + rs * sl;
        rows[4 * m + 3] = ru * (ue * ih - q) - rs * sl;
        if (dx) dx[m] = ru * (up - ue) * ih;
    }
}

template <typename R>
__global__ void __launch_bounds__(256)
fold1d_rows_kernel(const R* __restrict__ rows, const int64_t* __restrict__ order, const int64_t* __restrict__ seg, int64_t ne,
                   R* __restrict__ elem_tmp) {
    for (int64_t e = (int64_t)blockIdx.x * 256 + threadIdx.x; e < ne; e += (int64_t)gridDim.x * 256) {
        R a0 = R(0), a1 = R(0), a2 = R(0), a3 = R(0);
        for (int64_t r = seg[e]; r < seg[e + 1]; ++r) {
            const int64_t m = order[r];
            a0 += rows[4 * m]; a1 += rows[4 * m + 1]; a2 += rows[4 * m + 2]; a3 += rows[4 * m + 3];
        }
        elem_tmp[4 * e] = a0; elem_tmp[4 * e + 1] = a1; elem_tmp[4 * e + 2] = a2; elem_tmp[4 * e + 3] = a3;
    }
}

template <typename R>
__global__ void __launch_bounds__(256)
fold1d_nodes_kernel(const R* __restrict__ elem_tmp, int64_t N, R* __restrict__ du, R* __restrict__ dg) {
    for (int64_t k = (int64_t)blockIdx.x * 256 + threadIdx.x; k < N; k += (int64_t)gridDim.x * 256) {
        R a = R(0), b = R(0);
        if (k > 0) { a += elem_tmp[4 * (k - 1) + 1]; b += elem_tmp[4 * (k - 1) + 3]; }      // right node of element k-1
        if (k < N - 1) { a += elem_tmp[4 * k]; b += elem_tmp[4 * k + 2]; }                  // left node of element k
        du[k] = a; dg[k] = b;
    }
}

// ---- fused bar energy ------------------------------------------------------------------------------------
template <typename R> __device__ __forceinline__ R example3_b(R x) {
    // examples/example3.py:16-24: -N1/D1 - N2/D2 with D = exp(pi a^2), evaluated as N * exp(-pi a^2): one exponential and no
    // division per term (the step is bound by these FP64 exponentials); underflow -> 0 where the reference's D overflows -> N/inf = 0
    const R pi = R(3.14159265358979323846);
    const R a = x - R(2.5), b = x - R(7.5);
    const R N1 = R(4) * pi * pi * a * a - R(2) * pi, E1 = exp(-pi * a * a);
    const R N2 = R(8) * pi * pi * b * b - R(4) * pi, E2 = exp(-pi * b * b);
    return -N1 * E1 - N2 * E2;
}

constexpr int kBarBlock = 256;

template <typename R>
__global__ void __launch_bounds__(kBarBlock)
bar_energy_kernel(const R* __restrict__ grid, int64_t N, const R* __restrict__ uf, const R* __restrict__ xi, const R* __restrict__ wi,
                  int ng, R E, const R* __restrict__ b_table, int need_grad, R* __restrict__ partial, R* __restrict__ du,
                  R* __restrict__ dg, int32_t* __restrict__ flag) {
    __shared__ R s_ru[kBarBlock], s_rg[kBarBlock];
    __shared__ R s_red[kBarBlock / 32];
    // thread t handles element e = k0 - 1 + t and finalises node e (its left node) using thread t-1's right part
    const int64_t k0 = (int64_t)blockIdx.x * (kBarBlock - 1);
    const int t = threadIdx.x;
    const int64_t e = k0 - 1 + t;
    R lu = R(0), lg = R(0), ru = R(0), rg = R(0), en = R(0);
    if (e >= 0 && e <= N - 2) {
        const R ge = grid[e], gp = grid[e + 1], ue = uf[e], up = uf[e + 1];
        const R hd = gp - ge;                           // detached spacing used for xq, wq
        const bool act = hd >= R(1e-10);
        const R h = act ? hd : R(1e-10);
        const R ih = R(1) / h;
        const R dudx = (up - ue) * ih;
        for (int q = 0; q < ng; ++q) {
            const R xq = R(0.5) * hd * xi[q] + R(0.5) * (gp + ge);
            const R wq = R(0.5) * hd * wi[q];
            // the reference looks xq up again (models.py:73-74); the fused fold assumes it lands in element e
            const bool ok = (ge < xq || e == 0) && (xq <= gp || e == N - 2);
            if (!ok) *flag = 1;
            const R b = b_table ? b_table[e * ng + q] : example3_b<R>(xq);
            const R N1 = (gp - xq) * ih, N2 = (xq - ge) * ih;
            const R u = ue * N1 + up * N2;
            en += wq * (R(0.5) * E * dudx * dudx - b * u);
            const R r_u = -wq * b, r_s = wq * E * dudx;
            const R num = ue * (gp - xq) + up * (xq - ge);
            const R qq = act ? num * ih * ih : R(0);
            const R sl = act ? (up - ue) * ih * ih : R(0);
            lu += r_u * N1 - r_s * ih;
            ru += r_u * N2 + r_s * ih;
            lg += r_u * (-up * ih + qq) + r_s * sl;
            rg += r_u * (ue * ih - qq) - r_s * sl;
        }
    }
    // energy of elements owned by this block (t >= 1; thread 0 recomputes the previous block's last element)
    const R mine = (t >= 1) ? en : R(0);
    if (need_grad) {
        s_ru[t] = ru; s_rg[t] = rg;
        __syncthreads();
        if (t >= 1 && e >= 0 && e <= N - 1) {
            du[e] = lu + s_ru[t - 1];
            dg[e] = lg + s_rg[t - 1];
        }
    }
    const R tot = block_sum<R, kBarBlock>(mine, s_red);
    if (t == 0) partial[blockIdx.x] = tot;
}

template <typename R>
__global__ void __launch_bounds__(1024) sum_partials_kernel(const R* __restrict__ partial, int64_t nb, R* __restrict__ out) {
    __shared__ double s_red[32];
    double a = 0.0;
    for (int64_t i = threadIdx.x; i < nb; i += 1024) a += (double)partial[i];
    const double tot = block_sum<double, 1024>(a, s_red);
    if (threadIdx.x == 0) out[0] = (R)tot;
}

// =========================================================================================================
// Fused bar step (examples/example3.py:27-70 over models.py:45-90 for an r-adaptive model): grid, energy, d loss/d u and
// d loss/d increments in THREE launches --
//   1  block sums of the increments; the last block to finish (integer ticket) scans them          (bar_step_sums)
//   2  per element: prefix -> both grid values -> quadrature -> energy and the four nodal pieces; nodes finalised in
//      registers / shared memory / one recomputed overlap element per block; block sums of gamma = d loss/d grid[1:]
//      and gamma*cum; the last block scans them and adds the energy partials in fixed order          (bar_step_energy)
//   3  the softplus / cumsum / normalise chain: d loss/d increments                                  (grid_bwd_final)
// Scratch (reals): part0 [nb+1] | part1 [nb+1] | part2 [nb+1] | partE [nb] | 2 tickets (as reals' storage).
// u_free holds the trainable values; node k reads u_free[k - has_u0] (u0 / uN: fixed end values or NULL).
// =========================================================================================================
template <typename R> __device__ __forceinline__ void scan_partials_block(R* __restrict__ p, int64_t nb, R* s_warp) {
    R carry = R(0);
    for (int64_t base = 0; base < nb; base += kScanThreads) {
        const int64_t i = base + threadIdx.x;
        const R v = i < nb ? p[i] : R(0);
        R tot;
        const R ex = block_excl_scan<R>(v, s_warp, &tot);
        if (i < nb) p[i] = carry + ex;
        carry += tot;
    }
    if (threadIdx.x == 0) p[nb] = carry;
    __syncthreads();
}

template <typename R>
__global__ void __launch_bounds__(kScanThreads)
bar_step_sums_kernel(const R* __restrict__ p, int64_t n, R* __restrict__ part0, int64_t nb, unsigned* __restrict__ ticket,
                     int32_t* __restrict__ flag) {
    __shared__ R s_warp[kScanThreads / 32 + 1];
    __shared__ unsigned s_last;
    const int64_t base = (int64_t)blockIdx.x * kScanChunk + (int64_t)threadIdx.x * kScanItems;
    R s0 = R(0);
#pragma unroll
    for (int k = 0; k < kScanItems; ++k) {
        const int64_t i = base + k;
        if (i < n) s0 += softplus_inc<R>(p[i]);
    }
    R tot;
    block_excl_scan<R>(s0, s_warp, &tot);
    if (threadIdx.x == 0) {
        part0[blockIdx.x] = tot;
        if (blockIdx.x == 0) *flag = 0;
        __threadfence();
        s_last = (atomicAdd(ticket, 1u) == gridDim.x - 1) ? 1u : 0u;
    }
    __syncthreads();
    if (s_last) {
        __threadfence();
        scan_partials_block<R>(part0, nb, s_warp);
        if (threadIdx.x == 0) *ticket = 0u;
    }
}

template <typename R>
__device__ __forceinline__ void bar_element(const R ge, const R gp, const R ue, const R up, const int64_t e, const int64_t N,
                                            const R* __restrict__ xi, const R* __restrict__ wi, const int ng, const R E,
                                            const R* __restrict__ b_table, int32_t* __restrict__ flag, R& en, R& lu, R& lg, R& ru, R& rg) {
    const R hd = gp - ge;                           // detached spacing used for xq, wq (example3.py:41-52)
    const bool act = hd >= R(1e-10);
    const R h = act ? hd : R(1e-10);
    const R ih = R(1) / h;
    const R dudx = (up - ue) * ih;
    en = lu = lg = ru = rg = R(0);
    for (int q = 0; q < ng; ++q) {
        const R xq = R(0.5) * hd * xi[q] + R(0.5) * (gp + ge);
        const R wq = R(0.5) * hd * wi[q];
        const bool ok = (ge < xq || e == 0) && (xq <= gp || e == N - 2);
        if (!ok) *flag = 1;
        const R b = b_table ? b_table[e * ng + q] : example3_b<R>(xq);
        const R N1 = (gp - xq) * ih, N2 = (xq - ge) * ih;
        const R u = ue * N1 + up * N2;
        en += wq * (R(0.5) * E * dudx * dudx - b * u);
        const R r_u = -wq * b, r_s = wq * E * dudx;
        const R num = ue * (gp - xq) + up * (xq - ge);
        const R qq = act ? num * ih * ih : R(0);
        const R sl = act ? (up - ue) * ih * ih : R(0);
        lu += r_u * N1 - r_s * ih;
        ru += r_u * N2 + r_s * ih;
        lg += r_u * (-up * ih + qq) + r_s * sl;
        rg += r_u * (ue * ih - qq) - r_s * sl;
    }
}

template <typename R>
__global__ void __launch_bounds__(kScanThreads)
bar_step_energy_kernel(const R* __restrict__ p, int64_t n, const R* __restrict__ x0p, const R* __restrict__ xNp,
                       const R* __restrict__ u_free, const R* __restrict__ u0p, const R* __restrict__ uNp, const R* __restrict__ xi,
                       const R* __restrict__ wi, int ng, R E, const R* __restrict__ b_table, R* __restrict__ part0, int64_t nb,
                       R* __restrict__ gam, R* __restrict__ du_free, R* __restrict__ loss, unsigned* __restrict__ ticket,
                       int32_t* __restrict__ flag) {
    __shared__ R s_warp[kScanThreads / 32 + 1];
    __shared__ R s_c[kScanThreads], s_lu[kScanThreads], s_lg[kScanThreads];
    __shared__ double s_red[kScanThreads / 32];
    __shared__ unsigned s_last;
    R* part1 = part0 + (nb + 1);
    R* part2 = part1 + (nb + 1);
    R* partE = part2 + (nb + 1);
    const int64_t N = n + 1;                      // nodes
    const int t = threadIdx.x;
    const int64_t b = blockIdx.x;
    const int64_t base = b * kScanChunk + (int64_t)t * kScanItems;
    const int has_u0 = u0p != nullptr;
    auto u_at = [&](int64_t k) -> R {
        if (k == 0 && u0p) return *u0p;
        if (k == N - 1 && uNp) return *uNp;
        return u_free[k - has_u0];
    };
    R v[kScanItems], c[kScanItems];
    R s = R(0);
#pragma unroll
    for (int k = 0; k < kScanItems; ++k) {
        const int64_t i = base + k;
        v[k] = i < n ? softplus_inc<R>(p[i]) : R(0);
        s += v[k];
    }
    R tot;
    R run = block_excl_scan<R>(s, s_warp, &tot) + part0[b];
    const R S = part0[nb], x0 = *x0p, L = *xNp - *x0p;
#pragma unroll
    for (int k = 0; k < kScanItems; ++k) {
        const int64_t i = base + k;
        run += v[k];
        c[k] = i < n ? canonical_cum<R>(run, i, n, part0, b) : R(0);
    }
    s_c[t] = c[kScanItems - 1];
    __syncthreads();
    // left grid value of the thread's first element: canonical prefix of the item before it
    const R c_left = base == 0 ? R(0) : (t == 0 ? part0[b] : s_c[t - 1]);
    R gl = base == 0 ? x0 : x0 + L * c_left / S;
    R ul = base < n ? u_at(base) : R(0);
    R en_acc = R(0), g1 = R(0), g2 = R(0);
    R gamma[kScanItems];
    R first_lu = R(0), first_lg = R(0);
    R prev_ru = R(0), prev_rg = R(0);             // right pieces of the previous element, waiting for the next one's left pieces
#pragma unroll
    for (int k = 0; k < kScanItems; ++k) {
        const int64_t i = base + k;
        gamma[k] = R(0);
        if (i < n) {
            const R gr = x0 + L * c[k] / S;
            const R ur = u_at(i + 1);
            R en, lu, lg, ru, rg;
            bar_element<R>(gl, gr, ul, ur, i, N, xi, wi, ng, E, b_table, flag, en, lu, lg, ru, rg);
            en_acc += en;
            if (k == 0) { first_lu = lu; first_lg = lg; }
            else {      // node i (between elements i-1 and i) is complete
                gamma[k - 1] = lg + prev_rg;
                const R dun = lu + prev_ru;
                if (!(i == N - 1 && uNp)) du_free[i - has_u0] = dun;
            }
            prev_ru = ru; prev_rg = rg;
            gl = gr; ul = ur;
        }
    }
    s_lu[t] = first_lu; s_lg[t] = first_lg;
    __syncthreads();
    {
        // node after the thread's last element: left pieces from the next thread, from the recomputed first element of the
        // next block (last thread), or nothing (last element of the bar)
        const int64_t ilast = min(base + kScanItems, n) - 1;        // last element of this thread
        if (base < n) {
            R nlu = R(0), nlg = R(0);
            const int64_t inext = ilast + 1;
            if (inext < n) {
                if (t + 1 < kScanThreads && base + kScanItems < (b + 1) * (int64_t)kScanChunk && inext == base + kScanItems) {
                    nlu = s_lu[t + 1]; nlg = s_lg[t + 1];
                } else {
                    // first element of the next block: same canonical prefix as that block computes
                    const R cn = canonical_cum<R>(part0[b + 1] + softplus_inc<R>(p[inext]), inext, n, part0, b + 1);
                    R en, ru, rg;
                    bar_element<R>(gl, x0 + L * cn / S, ul, u_at(inext + 1), inext, N, xi, wi, ng, E, b_table, flag, en, nlu, nlg, ru, rg);
                }
            }
            const int kl = (int)(ilast - base);
            const R gnode = nlg + prev_rg, dun = nlu + prev_ru;
#pragma unroll
            for (int k = 0; k < kScanItems; ++k)
                if (k == kl) gamma[k] = gnode;
            const int64_t node = ilast + 1;
            if (!(node == N - 1 && uNp)) du_free[node - has_u0] = dun;
            if (base == 0 && !u0p) du_free[0] = first_lu;          // node 0 has one element
        }
    }
#pragma unroll
    for (int k = 0; k < kScanItems; ++k) {
        const int64_t i = base + k;
        if (i < n) { gam[i] = gamma[k]; g1 += gamma[k]; g2 += gamma[k] * c[k]; }
    }
    R t1, t2;
    block_excl_scan<R>(g1, s_warp, &t1);
    block_excl_scan<R>(g2, s_warp, &t2);
    const double eb = block_sum<double, kScanThreads>((double)en_acc, s_red);
    if (t == 0) {
        part1[b] = t1; part2[b] = t2; partE[b] = (R)eb;
        __threadfence();
        s_last = (atomicAdd(ticket, 1u) == gridDim.x - 1) ? 1u : 0u;
    }
    __syncthreads();
    if (s_last) {
        __threadfence();
        scan_partials_block<R>(part1, nb, s_warp);            // exclusive prefix of gamma per block, total at [nb]
        scan_partials_block<R>(part2, nb, s_warp);            // only the total T = sum gamma*cum at [nb] is used
        double a = 0.0;
        for (int64_t i = t; i < nb; i += kScanThreads) a += (double)partE[i];
        const double e_tot = block_sum<double, kScanThreads>(a, s_red);
        if (t == 0) { loss[0] = (R)e_tot; *ticket = 0u; }
    }
}

// phase 3: d loss / d increments from gamma, the prefix offsets of gamma (part1), T (part2[nb]) and S (part0[nb])
template <typename R>
__global__ void __launch_bounds__(kScanThreads)
bar_step_chain_kernel(const R* __restrict__ gam, const R* __restrict__ p, int64_t n, const R* __restrict__ x0p, const R* __restrict__ xNp,
                      const R* __restrict__ part0, int64_t nb, R* __restrict__ dp) {
    __shared__ R s_warp[kScanThreads / 32 + 1];
    const R* part1 = part0 + (nb + 1);
    const R* part2 = part1 + (nb + 1);
    const int64_t base = (int64_t)blockIdx.x * kScanChunk + (int64_t)threadIdx.x * kScanItems;
    R g[kScanItems];
    R s = R(0);
#pragma unroll
    for (int k = 0; k < kScanItems; ++k) {
        const int64_t i = base + k;
        g[k] = i < n ? gam[i] : R(0);
        s += g[k];
    }
    R tot;
    R run = block_excl_scan<R>(s, s_warp, &tot) + part1[blockIdx.x];
    const R G = part1[nb], T = part2[nb], S = part0[nb], L = *xNp - *x0p;
    const R c = L / S, corr = L * T / (S * S);
#pragma unroll
    for (int k = 0; k < kScanItems; ++k) {
        const int64_t i = base + k;
        if (i < n) {
            const R dinc = c * (G - run) - corr;
            const R pi = p[i];
            const R sp = softplus_raw<R>(pi);
            const R dsp = pi > R(20) ? R(1) : R(1) / (R(1) + exp(-pi));
            dp[i] = sp >= R(1e-6) ? dinc * dsp : R(0);
        }
        run += g[k];
    }
}

static inline int grid_for(int64_t n, int block = 256) { return (int)std::max<int64_t>(1, std::min<int64_t>((n + block - 1) / block, 148 * 32)); }
static inline int64_t nblocks_scan(int64_t n) { return (n + kScanChunk - 1) / kScanChunk; }

template <typename R>
static int grid_fwd(const R* p, int64_t n, const R* x0, const R* xN, R* grid, R* cum, R* scratch, void* s) {
    HIDENN_REQUIRE(p && x0 && xN && grid && cum && scratch && n >= 1, "1d_grid_fwd: bad arguments");
    cudaStream_t st = (cudaStream_t)s;
    const int64_t nb = nblocks_scan(n);
    scan_block_sums_kernel<R, 0><<<(int)nb, kScanThreads, 0, st>>>(p, nullptr, n, scratch, nb);
    scan_partials_kernel<R, 1><<<1, kScanThreads, 0, st>>>(scratch, nb);
    grid_fwd_final_kernel<R><<<(int)nb, kScanThreads, 0, st>>>(p, n, x0, xN, scratch, nb, grid, cum);
    HIDENN_CUDA_OK(cudaGetLastError());
    return 0;
}

template <typename R>
static int grid_bwd(const R* dgrid, const R* p, const R* cum, const R* x0, const R* xN, int64_t n, R* dp, R* scratch, void* s) {
    HIDENN_REQUIRE(dgrid && p && cum && x0 && xN && dp && scratch && n >= 1, "1d_grid_bwd: bad arguments");
    cudaStream_t st = (cudaStream_t)s;
    const int64_t nb = nblocks_scan(n);
    scan_block_sums_kernel<R, 1><<<(int)nb, kScanThreads, 0, st>>>(dgrid, cum, n, scratch, nb);
    scan_partials_kernel<R, 2><<<1, kScanThreads, 0, st>>>(scratch, nb);
    grid_bwd_final_kernel<R><<<(int)nb, kScanThreads, 0, st>>>(dgrid, p, cum, n, x0, xN, scratch, nb, dp);
    HIDENN_CUDA_OK(cudaGetLastError());
    return 0;
}

template <typename R> static int lookup(const R* grid, int64_t N, const R* x, int64_t M, int32_t* elem, void* s) {
    HIDENN_REQUIRE(N >= 2, "1d_lookup: the grid needs at least 2 nodes");
    if (M <= 0) return 0;
    HIDENN_REQUIRE(grid && x && elem, "1d_lookup: NULL");
    lookup_kernel<R><<<grid_for(M), 256, 0, (cudaStream_t)s>>>(grid, N, x, M, elem);
    HIDENN_CUDA_OK(cudaGetLastError());
    return 0;
}

template <typename R>
static int interp_fwd(const R* grid, int64_t N, const R* uf, const R* x, int64_t M, R* u, int32_t* elem, R* slope, void* s) {
    HIDENN_REQUIRE(N >= 2, "1d_interp_fwd: the grid needs at least 2 nodes");
    if (M <= 0) return 0;
    HIDENN_REQUIRE(grid && uf && x && u, "1d_interp_fwd: NULL");
    interp_fwd_kernel<R><<<grid_for(M), 256, 0, (cudaStream_t)s>>>(grid, N, uf, x, M, u, elem, slope);
    HIDENN_CUDA_OK(cudaGetLastError());
    return 0;
}

template <typename R>
static int interp_bwd(const R* grid, int64_t N, const R* uf, const R* x, const int32_t* elem, const R* r_u, const R* r_s, int64_t M,
                      R* rows, R* dx, void* s) {
    (void)N;
    if (M <= 0) return 0;
    HIDENN_REQUIRE(grid && uf && x && elem && rows, "1d_interp_bwd: NULL");
    interp_bwd_kernel<R><<<grid_for(M), 256, 0, (cudaStream_t)s>>>(grid, uf, x, elem, r_u, r_s, M, rows, dx);
    HIDENN_CUDA_OK(cudaGetLastError());
    return 0;
}

template <typename R>
static int fold_rows_1d(const R* rows, const int64_t* order, const int64_t* seg, int64_t N, R* elem_tmp, R* du, R* dg, void* s) {
    HIDENN_REQUIRE(N >= 2 && order && seg && elem_tmp && du && dg, "1d_fold_rows: bad arguments");
    fold1d_rows_kernel<R><<<grid_for(N - 1), 256, 0, (cudaStream_t)s>>>(rows, order, seg, N - 1, elem_tmp);
    fold1d_nodes_kernel<R><<<grid_for(N), 256, 0, (cudaStream_t)s>>>(elem_tmp, N, du, dg);
    HIDENN_CUDA_OK(cudaGetLastError());
    return 0;
}

template <typename R>
static int bar_energy(const R* grid, int64_t N, const R* uf, const R* xi, const R* wi, int ng, R E, const R* b_table, int need_grad,
                      R* loss, R* du, R* dg, int32_t* flag, R* scratch, void* s) {
    HIDENN_REQUIRE(N >= 2 && grid && uf && xi && wi && loss && flag && scratch, "1d_bar_energy: bad arguments");
    HIDENN_REQUIRE(ng >= 1 && ng <= 8, "1d_bar_energy: ng must be in [1,8]");
    HIDENN_REQUIRE(!need_grad || (du && dg), "1d_bar_energy: gradient outputs NULL");
    cudaStream_t st = (cudaStream_t)s;
    const int64_t nb = (N + (kBarBlock - 1) - 1) / (kBarBlock - 1);     // nodes 0..N-1, kBarBlock-1 nodes per block
    HIDENN_CUDA_OK(cudaMemsetAsync(flag, 0, sizeof(int32_t), st));
    bar_energy_kernel<R><<<(int)nb, kBarBlock, 0, st>>>(grid, N, uf, xi, wi, ng, E, b_table, need_grad, scratch, du, dg, flag);
    sum_partials_kernel<R><<<1, 1024, 0, st>>>(scratch, nb, loss);
    HIDENN_CUDA_OK(cudaGetLastError());
    return 0;
}

template <typename R>
static int bar_step(const R* p, int64_t n, const R* x0, const R* xN, const R* u_free, const R* u0, const R* uN, const R* xi, const R* wi,
                    int ng, R E, const R* b_table, R* loss, R* dp, R* du_free, R* gam, int32_t* flag, R* scratch, void* s) {
    HIDENN_REQUIRE(n >= 1 && p && x0 && xN && u_free && xi && wi && loss && dp && du_free && gam && flag && scratch, "1d_bar_step: bad arguments");
    HIDENN_REQUIRE(ng >= 1 && ng <= 8, "1d_bar_step: ng must be in [1,8]");
    cudaStream_t st = (cudaStream_t)s;
    const int64_t nb = nblocks_scan(n);
    unsigned* ticket = reinterpret_cast<unsigned*>(scratch + 3 * (nb + 1) + nb + 2);      // zero before the first call; kernels reset it
    bar_step_sums_kernel<R><<<(int)nb, kScanThreads, 0, st>>>(p, n, scratch, nb, ticket, flag);
    bar_step_energy_kernel<R><<<(int)nb, kScanThreads, 0, st>>>(p, n, x0, xN, u_free, u0, uN, xi, wi, ng, E, b_table, scratch, nb, gam, du_free,
                                                                loss, ticket + 1, flag);
    bar_step_chain_kernel<R><<<(int)nb, kScanThreads, 0, st>>>(gam, p, n, x0, xN, scratch, nb, dp);
    HIDENN_CUDA_OK(cudaGetLastError());
    return 0;
}

}  // namespace hidenn

using namespace hidenn;

extern "C" int64_t hidenn_1d_bar_step_scratch(int64_t n) { return 4 * (nblocks_scan(n) + 1) + 16; }

extern "C" int64_t hidenn_1d_scratch_size(int64_t n) {
    const int64_t a = 2 * (nblocks_scan(n) + 1) + 8;
    const int64_t b = (n + 1 + kBarBlock - 2) / (kBarBlock - 1) + 8;
    return a > b ? a : b;
}

#define HIDENN_GRID1D_API(SUF, T)                                                                                                  \
    extern "C" int hidenn_1d_grid_fwd_##SUF(const T* p, int64_t n, const T* x0, const T* xN, T* g, T* c, T* sc, void* s) {          \
        return grid_fwd<T>(p, n, x0, xN, g, c, sc, s);                                                                             \
    }                                                                                                                              \
    extern "C" int hidenn_1d_grid_bwd_##SUF(const T* dg, const T* p, const T* c, const T* x0, const T* xN, int64_t n, T* dp, T* sc,  \
                                            void* s) {                                                                             \
        return grid_bwd<T>(dg, p, c, x0, xN, n, dp, sc, s);                                                                        \
    }                                                                                                                              \
    extern "C" int hidenn_1d_lookup_##SUF(const T* g, int64_t N, const T* x, int64_t M, int32_t* e, void* s) {                     \
        return lookup<T>(g, N, x, M, e, s);                                                                                        \
    }                                                                                                                              \
    extern "C" int hidenn_1d_interp_fwd_##SUF(const T* g, int64_t N, const T* uf, const T* x, int64_t M, T* u, int32_t* e, T* sl,   \
                                              void* s) {                                                                           \
        return interp_fwd<T>(g, N, uf, x, M, u, e, sl, s);                                                                         \
    }                                                                                                                              \
    extern "C" int hidenn_1d_interp_bwd_##SUF(const T* g, int64_t N, const T* uf, const T* x, const int32_t* e, const T* ru,        \
                                              const T* rs, int64_t M, T* rows, T* dx, void* s) {                                   \
        return interp_bwd<T>(g, N, uf, x, e, ru, rs, M, rows, dx, s);                                                              \
    }                                                                                                                              \
    extern "C" int hidenn_1d_fold_rows_##SUF(const T* rows, const int64_t* o, const int64_t* sg, int64_t N, T* tmp, T* du, T* dg,   \
                                             void* s) {                                                                            \
        return fold_rows_1d<T>(rows, o, sg, N, tmp, du, dg, s);                                                                    \
    }                                                                                                                              \
    extern "C" int hidenn_1d_bar_energy_##SUF(const T* g, int64_t N, const T* uf, const T* xi, const T* wi, int ng, T E,            \
                                              const T* bt, int need, T* loss, T* du, T* dg, int32_t* flag, T* sc, void* s) {       \
        return bar_energy<T>(g, N, uf, xi, wi, ng, E, bt, need, loss, du, dg, flag, sc, s);                                        \
    }

HIDENN_GRID1D_API(f64, double)
HIDENN_GRID1D_API(f32, float)

extern "C" int hidenn_1d_bar_step_f64(const double* p, int64_t n, const double* x0, const double* xN, const double* u_free, const double* u0,
                                      const double* uN, const double* xi, const double* wi, int ng, double E, const double* b_table,
                                      double* loss, double* dp, double* du_free, double* gam, int32_t* flag, double* scratch, void* s) {
    return bar_step<double>(p, n, x0, xN, u_free, u0, uN, xi, wi, ng, E, b_table, loss, dp, du_free, gam, flag, scratch, s);
}
extern "C" int hidenn_1d_bar_step_f32(const float* p, int64_t n, const float* x0, const float* xN, const float* u_free, const float* u0,
                                      const float* uN, const float* xi, const float* wi, int ng, float E, const float* b_table, float* loss,
                                      float* dp, float* du_free, float* gam, int32_t* flag, float* scratch, void* s) {
    return bar_step<float>(p, n, x0, xN, u_free, u0, uN, xi, wi, ng, E, b_table, loss, dp, du_free, gam, flag, scratch, s);
}
