// Structured Q1 (tensor-product) model kernels: two 1D lookups, bilinear interpolation, per-row VJP pieces and
// the deterministic cell / node / grid-line folds.  Reference: /root/reference/src/models.py:180-212.
#include "../../include/hidenn_b200_grid.h"
#include "common.cuh"
#include <algorithm>

namespace hidenn {

__device__ __forceinline__ double rcp_rn(double x) { return __drcp_rn(x); }
__device__ __forceinline__ float rcp_rn(float x) { return __frcp_rn(x); }

template <typename R> __device__ __forceinline__ int lookup_line(const R* __restrict__ grid, int64_t N, R x) {
    int64_t lo = 0, hi = N;
    while (lo < hi) {
        const int64_t mid = (lo + hi) >> 1;
        if (__ldg(grid + mid) < x) lo = mid + 1; else hi = mid;
    }
    int64_t e = lo - 1;
    e = e < 0 ? 0 : (e > N - 2 ? N - 2 : e);
    return (int)e;
}

// clamp(searchsorted_left(grid, x) - 1, 0, N-2), bit-exact, but starting from the position a uniform grid would give:
// r-adaptive grids stay close to uniform, so a few neighbour comparisons replace the ~log2(N) dependent loads of a
// bisection (which remains the fallback when the guess is more than 4 lines off).  inv = (N-1) / (grid[N-1] - grid[0]).
template <typename R> __device__ __forceinline__ int lookup_line_smem(const R* grid, int N, R x, R g0, R inv) {
    const R t = (x - g0) * inv;
    int lo = t >= R(0) ? (t < (R)(N - 1) ? (int)t : N - 1) : 0;       // NaN -> 0, like the bisection
    if (lo > 0 && grid[lo - 1] >= x) {                                // first index with grid[i] >= x lies to the left
        int steps = 0;
        do { --lo; } while (lo > 0 && grid[lo - 1] >= x && ++steps < 4);
        if (lo > 0 && grid[lo - 1] >= x) {
            int a = 0, b = lo;
            while (a < b) { const int mid = (a + b) >> 1; if (grid[mid] < x) a = mid + 1; else b = mid; }
            lo = a;
        }
    } else {
        int steps = 0;
        while (lo < N && grid[lo] < x && steps < 4) { ++lo; ++steps; }
        if (lo < N && grid[lo] < x) {
            int a = lo + 1, b = N;
            while (a < b) { const int mid = (a + b) >> 1; if (grid[mid] < x) a = mid + 1; else b = mid; }
            lo = a;
        }
    }
    const int e = lo - 1;
    return e < 0 ? 0 : (e > N - 2 ? N - 2 : e);
}

// both grid lines staged in shared memory (used when they fit): the two binary searches then cost LDS latency only.
// L2 = true: fused L2-projection loss -- writes the residual weight 2 (u_h - target) / M instead of u_h and leaves
// the CTA's sum of squared residuals (fixed order: per-thread strided sum, warp tree, warp order) in partial[blockIdx.x].
template <typename R, bool L2, bool SMEM>
__global__ void __launch_bounds__(256)
q1_fwd_smem_kernel(const R* __restrict__ gx, int Nx, const R* __restrict__ gy, int Ny, const R* __restrict__ uf,
                   const typename Real2<R>::type* __restrict__ x, const R* __restrict__ target, int64_t M, R* __restrict__ u,
                   int32_t* __restrict__ ixo, int32_t* __restrict__ iyo, R* __restrict__ partial) {
    extern __shared__ __align__(16) unsigned char q1_smem[];
    __shared__ R s_red[8];
    const R* sx = gx;
    const R* sy = gy;
    if (SMEM) {
        R* wx = reinterpret_cast<R*>(q1_smem);
        R* wy = wx + Nx;
        for (int i = threadIdx.x; i < Nx; i += 256) wx[i] = gx[i];
        for (int i = threadIdx.x; i < Ny; i += 256) wy[i] = gy[i];
        __syncthreads();
        sx = wx; sy = wy;
    }
    const R scale = R(2) / (R)M;
    const R gx0 = sx[0], gy0 = sy[0];
    const R spanx = sx[Nx - 1] - gx0, spany = sy[Ny - 1] - gy0;
    const R invx = spanx > R(0) ? (R)(Nx - 1) / spanx : R(0), invy = spany > R(0) ? (R)(Ny - 1) / spany : R(0);
    R acc = R(0);
    // one sample: lookups, 4 nodal values (L2-resident gathers), bilinear interpolation.  Two samples per thread and
    // iteration: the kernel is bound by the latency of the dependent lookup -> gather -> divide chain (1.19 ms for 2.7 GB
    // at C3 with one sample in flight), not by bytes
    auto sample = [&](const int64_t m, R& out, int& ixr, int& iyr) {
        const typename Real2<R>::type p = x[m];
        const int ix = lookup_line_smem<R>(sx, Nx, p.x, gx0, invx), iy = lookup_line_smem<R>(sy, Ny, p.y, gy0, invy);
        const R u00 = __ldg(uf + (int64_t)ix * Ny + iy), u10 = __ldg(uf + (int64_t)(ix + 1) * Ny + iy);
        const R u01 = __ldg(uf + (int64_t)ix * Ny + iy + 1), u11 = __ldg(uf + (int64_t)(ix + 1) * Ny + iy + 1);
        const R x0 = sx[ix], x1 = sx[ix + 1], y0 = sy[iy], y1 = sy[iy + 1];
        R hx = x1 - x0, hy = y1 - y0;
        hx = hx < R(1e-10) ? R(1e-10) : hx;
        hy = hy < R(1e-10) ? R(1e-10) : hy;
        // two correctly rounded reciprocals instead of four divisions: the kernel is bound by the FP64 pipe (four IEEE
        // divisions were ~2/3 of its FP64 instructions); the backward (q1_row) forms the shape functions the same way
        const R ihx = rcp_rn(hx), ihy = rcp_rn(hy);
        const R N1x = (x1 - p.x) * ihx, N2x = (p.x - x0) * ihx, N1y = (y1 - p.y) * ihy, N2y = (p.y - y0) * ihy;
        const R uh = N1x * N1y * u00 + N2x * N1y * u10 + N1x * N2y * u01 + N2x * N2y * u11;
        out = L2 ? uh - target[m] : uh;
        ixr = ix; iyr = iy;
    };
    const int64_t stride = (int64_t)gridDim.x * 256;
    for (int64_t m = (int64_t)blockIdx.x * 256 + threadIdx.x; m < M; m += 2 * stride) {
        const int64_t m2 = m + stride;
        const bool two = m2 < M;
        R va, vb = R(0);
        int ixa, iya, ixb = 0, iyb = 0;
        sample(m, va, ixa, iya);
        if (two) sample(m2, vb, ixb, iyb);
        if (L2) {
            acc += va * va;            // same per-thread order as one sample per iteration: m, m + stride, ...
            u[m] = scale * va;
            if (two) { acc += vb * vb; u[m2] = scale * vb; }
        } else {
            u[m] = va;
            if (two) u[m2] = vb;
        }
        if (ixo) { ixo[m] = ixa; if (two) ixo[m2] = ixb; }
        if (iyo) { iyo[m] = iya; if (two) iyo[m2] = iyb; }
    }
    if (L2) {
        const R tot = block_sum<R, 256>(acc, s_red);
        if (threadIdx.x == 0) partial[blockIdx.x] = tot;
    }
}

template <typename R>
__global__ void __launch_bounds__(256) q1_l2_finish_kernel(const R* __restrict__ partial, int n, int64_t M, R* __restrict__ loss) {
    __shared__ double s_red[8];
    double a = 0.0;
    for (int i = threadIdx.x; i < n; i += 256) a += (double)partial[i];
    const double tot = block_sum<double, 256>(a, s_red);
    if (threadIdx.x == 0) loss[0] = (R)(tot / (double)M);
}

// VJP pieces of one sample in cell (ix,iy): o[0..3] = d u00,u10,u01,u11; o[4..5] = d gx_i, gx_{i+1}; o[6..7] = d gy_j, gy_{j+1}
template <typename R>
__device__ __forceinline__ void q1_row(const R x0, const R x1, const R y0, const R y1, const R u00, const R u10, const R u01,
                                       const R u11, const typename Real2<R>::type p, const R rv, R (&o)[8]) {
    const R hxr = x1 - x0, hyr = y1 - y0;
    const bool ax = hxr >= R(1e-10), ay = hyr >= R(1e-10);
    const R ihx = R(1) / (ax ? hxr : R(1e-10)), ihy = R(1) / (ay ? hyr : R(1e-10));
    const R N1x = (x1 - p.x) * ihx, N2x = (p.x - x0) * ihx, N1y = (y1 - p.y) * ihy, N2y = (p.y - y0) * ihy;
    o[0] = rv * N1x * N1y; o[1] = rv * N2x * N1y; o[2] = rv * N1x * N2y; o[3] = rv * N2x * N2y;
    const R A = N1y * u00 + N2y * u01, B = N1y * u10 + N2y * u11;
    const R numx = A * (x1 - p.x) + B * (p.x - x0);
    const R qx = ax ? numx * ihx * ihx : R(0);
    o[4] = rv * (-B * ihx + qx);
    o[5] = rv * (A * ihx - qx);
    const R Cc = N1x * u00 + N2x * u10, D = N1x * u01 + N2x * u11;
    const R numy = Cc * (y1 - p.y) + D * (p.y - y0);
    const R qy = ay ? numy * ihy * ihy : R(0);
    o[6] = rv * (-D * ihy + qy);
    o[7] = rv * (Cc * ihy - qy);
}

__global__ void __launch_bounds__(256) q1_bin_count_kernel(const int32_t* __restrict__ ix, const int32_t* __restrict__ iy, int64_t M,
                                                            int64_t cy, int32_t* __restrict__ cnt) {
    for (int64_t m = (int64_t)blockIdx.x * 256 + threadIdx.x; m < M; m += (int64_t)gridDim.x * 256)
        atomicAdd(cnt + (int64_t)ix[m] * cy + iy[m], 1);
}

__global__ void __launch_bounds__(256) q1_bin_scatter_kernel(const int32_t* __restrict__ ix, const int32_t* __restrict__ iy, int64_t M,
                                                              int64_t cy, const int32_t* __restrict__ seg, int32_t* __restrict__ cursor,
                                                              int32_t* __restrict__ order) {
    for (int64_t m = (int64_t)blockIdx.x * 256 + threadIdx.x; m < M; m += (int64_t)gridDim.x * 256) {
        const int64_t c = (int64_t)ix[m] * cy + iy[m];
        order[seg[c] + atomicAdd(cursor + c, 1)] = (int32_t)m;
    }
}

// canonical order inside every cell: ascending sample id (insertion sort; cells hold a handful of samples)
__global__ void __launch_bounds__(256) q1_bin_sort_kernel(int64_t ncell, const int32_t* __restrict__ seg, int32_t* __restrict__ order) {
    for (int64_t c = (int64_t)blockIdx.x * 256 + threadIdx.x; c < ncell; c += (int64_t)gridDim.x * 256) {
        const int32_t b = seg[c], e = seg[c + 1];
        for (int32_t i = b + 1; i < e; ++i) {
            const int32_t v = order[i];
            int32_t j = i - 1;
            while (j >= b && order[j] > v) { order[j + 1] = order[j]; --j; }
            order[j + 1] = v;
        }
    }
}

template <typename R>
__global__ void __launch_bounds__(256)
q1_fused_cells_kernel(const R* __restrict__ gx, const R* __restrict__ gy, int64_t Nx, int64_t Ny, const R* __restrict__ uf,
                      const typename Real2<R>::type* __restrict__ x, const R* __restrict__ r, const int32_t* __restrict__ seg,
                      const int32_t* __restrict__ order, R* __restrict__ cell_tmp) {
    const int64_t cy = Ny - 1, ncell = (Nx - 1) * cy;
    for (int64_t c = (int64_t)blockIdx.x * 256 + threadIdx.x; c < ncell; c += (int64_t)gridDim.x * 256) {
        R a[8];
#pragma unroll
        for (int k = 0; k < 8; ++k) a[k] = R(0);
        const int32_t b = seg[c], e = seg[c + 1];
        if (e > b) {
            const int64_t i = c / cy, j = c % cy;
            const R x0 = __ldg(gx + i), x1 = __ldg(gx + i + 1), y0 = __ldg(gy + j), y1 = __ldg(gy + j + 1);
            const R u00 = __ldg(uf + i * Ny + j), u10 = __ldg(uf + (i + 1) * Ny + j);
            const R u01 = __ldg(uf + i * Ny + j + 1), u11 = __ldg(uf + (i + 1) * Ny + j + 1);
            for (int32_t q = b; q < e; ++q) {
                const int32_t m = order[q];
                R o[8];
                q1_row<R>(x0, x1, y0, y1, u00, u10, u01, u11, x[m], r[m], o);
#pragma unroll
                for (int k = 0; k < 8; ++k) a[k] += o[k];
            }
        }
#pragma unroll
        for (int k = 0; k < 8; ++k) cell_tmp[(int64_t)k * ncell + c] = a[k];
    }
}

template <typename R>
__global__ void __launch_bounds__(256)
q1_bwd_kernel(const R* __restrict__ gx, const R* __restrict__ gy, int64_t Ny, const R* __restrict__ uf,
              const typename Real2<R>::type* __restrict__ x, const int32_t* __restrict__ ixs, const int32_t* __restrict__ iys,
              const R* __restrict__ r, int64_t M, R* __restrict__ rows) {
    for (int64_t m = (int64_t)blockIdx.x * 256 + threadIdx.x; m < M; m += (int64_t)gridDim.x * 256) {
        const typename Real2<R>::type p = x[m];
        const int ix = ixs[m], iy = iys[m];
        const R x0 = __ldg(gx + ix), x1 = __ldg(gx + ix + 1), y0 = __ldg(gy + iy), y1 = __ldg(gy + iy + 1);
        const R u00 = __ldg(uf + (int64_t)ix * Ny + iy), u10 = __ldg(uf + (int64_t)(ix + 1) * Ny + iy);
        const R u01 = __ldg(uf + (int64_t)ix * Ny + iy + 1), u11 = __ldg(uf + (int64_t)(ix + 1) * Ny + iy + 1);
        R o[8];
        q1_row<R>(x0, x1, y0, y1, u00, u10, u01, u11, p, r[m], o);
#pragma unroll
        for (int k = 0; k < 8; ++k) rows[8 * m + k] = o[k];
    }
}

template <typename R>
__global__ void __launch_bounds__(256)
q1_fold_cells_kernel(const R* __restrict__ rows, const int64_t* __restrict__ order, const int64_t* __restrict__ seg, int64_t ncell,
                     R* __restrict__ cell_tmp) {
    for (int64_t c = (int64_t)blockIdx.x * 256 + threadIdx.x; c < ncell; c += (int64_t)gridDim.x * 256) {
        R a[8];
#pragma unroll
        for (int k = 0; k < 8; ++k) a[k] = R(0);
        for (int64_t q = seg[c]; q < seg[c + 1]; ++q) {
            const int64_t m = order[q];
#pragma unroll
            for (int k = 0; k < 8; ++k) a[k] += rows[8 * m + k];
        }
#pragma unroll
        for (int k = 0; k < 8; ++k) cell_tmp[(int64_t)k * ncell + c] = a[k];      // [8, ncell]: every consumer reads one plane contiguously
    }
}

template <typename R>
__global__ void __launch_bounds__(256)
q1_fold_nodes_kernel(const R* __restrict__ cell_tmp, int64_t Nx, int64_t Ny, R* __restrict__ du) {
    const int64_t total = Nx * Ny, cy = Ny - 1, ncell = (Nx - 1) * cy;
    for (int64_t n = (int64_t)blockIdx.x * 256 + threadIdx.x; n < total; n += (int64_t)gridDim.x * 256) {
        const int64_t i = n / Ny, j = n % Ny;
        R a = R(0);
        if (i < Nx - 1 && j < Ny - 1) a += cell_tmp[0 * ncell + i * cy + j];
        if (i > 0 && j < Ny - 1) a += cell_tmp[1 * ncell + (i - 1) * cy + j];
        if (i < Nx - 1 && j > 0) a += cell_tmp[2 * ncell + i * cy + j - 1];
        if (i > 0 && j > 0) a += cell_tmp[3 * ncell + (i - 1) * cy + j - 1];
        du[n] = a;
    }
}

// x lines: one block per line, the cells of a line are contiguous in planes 4 / 5; fixed-order strided sums + block tree
template <typename R>
__global__ void __launch_bounds__(256) q1_fold_xlines_kernel(const R* __restrict__ cell_tmp, int64_t Nx, int64_t Ny, R* __restrict__ dgx) {
    __shared__ R s_red[8];
    const int64_t cy = Ny - 1, cx = Nx - 1, ncell = cx * cy;
    const int64_t i = blockIdx.x;
    R a = R(0);
    for (int64_t j = threadIdx.x; j < cy; j += 256) {
        if (i < cx) a += cell_tmp[4 * ncell + i * cy + j];
        if (i > 0) a += cell_tmp[5 * ncell + (i - 1) * cy + j];
    }
    const R tot = block_sum<R, 256>(a, s_red);
    if (threadIdx.x == 0) dgx[i] = tot;
}

// y lines: a block owns 32 neighbouring lines (threadIdx.x) and walks the cell rows 8 at a time (threadIdx.y), so every
// warp reads 32 consecutive values of planes 6 / 7; per-thread sums in ascending row order, then the 8 row groups in order
template <typename R>
__global__ void __launch_bounds__(256) q1_fold_ylines_kernel(const R* __restrict__ cell_tmp, int64_t Nx, int64_t Ny, R* __restrict__ dgy) {
    __shared__ R s_part[8][33];
    const int64_t cy = Ny - 1, cx = Nx - 1, ncell = cx * cy;
    const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
    const int64_t j = (int64_t)blockIdx.x * 32 + tx;
    R a = R(0);
    if (j < Ny) {
        for (int64_t i = ty; i < cx; i += 8) {
            if (j < cy) a += cell_tmp[6 * ncell + i * cy + j];
            if (j > 0) a += cell_tmp[7 * ncell + i * cy + j - 1];
        }
    }
    s_part[ty][tx] = a;
    __syncthreads();
    if (ty == 0 && j < Ny) {
        R tot = R(0);
#pragma unroll
        for (int k = 0; k < 8; ++k) tot += s_part[k][tx];
        dgy[j] = tot;
    }
}

template <typename R>
static void q1_fold_lines(const R* cell_tmp, int64_t Nx, int64_t Ny, R* dgx, R* dgy, cudaStream_t st) {
    q1_fold_xlines_kernel<R><<<(int)Nx, 256, 0, st>>>(cell_tmp, Nx, Ny, dgx);
    q1_fold_ylines_kernel<R><<<(int)((Ny + 31) / 32), 256, 0, st>>>(cell_tmp, Nx, Ny, dgy);
}

static inline int grid_for(int64_t n) { return (int)std::max<int64_t>(1, std::min<int64_t>((n + 255) / 256, 148 * 32)); }

constexpr int kQ1L2MaxCtas = 148 * 8;

// grid: every CTA stages the lines once (when they fit in 100 KB) and strides over the samples
template <typename R, bool L2>
static int q1_fwd_launch(const R* gx, int64_t Nx, const R* gy, int64_t Ny, const R* uf, const R* x, const R* target, int64_t M, R* out,
                         int32_t* ix, int32_t* iy, R* partial, cudaStream_t st, int* grid_out = nullptr) {
    using R2 = typename Real2<R>::type;
    HIDENN_REQUIRE(Nx < (1LL << 31) && Ny < (1LL << 31), "q1 forward: grid lines longer than 2^31");
    const size_t lines = (size_t)(Nx + Ny) * sizeof(R);
    const bool smem = lines <= 100 * 1024 && M >= 65536;
    int dev = 0, sms = 148;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    const int per_sm = smem ? (int)std::max<size_t>(1, std::min<size_t>(8, (200 * 1024) / (lines + 1024))) : 8;
    const int64_t want = (M + 255) / 256;
    const int grid = (int)std::max<int64_t>(1, std::min<int64_t>(want, std::min<int64_t>((int64_t)sms * per_sm, kQ1L2MaxCtas)));
    if (grid_out) *grid_out = grid;
    if (smem) {
        HIDENN_CUDA_OK(cudaFuncSetAttribute(q1_fwd_smem_kernel<R, L2, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, 100 * 1024));
        q1_fwd_smem_kernel<R, L2, true><<<grid, 256, lines, st>>>(gx, (int)Nx, gy, (int)Ny, uf, (const R2*)x, target, M, out, ix, iy, partial);
    } else {
        q1_fwd_smem_kernel<R, L2, false><<<grid, 256, 0, st>>>(gx, (int)Nx, gy, (int)Ny, uf, (const R2*)x, target, M, out, ix, iy, partial);
    }
    HIDENN_CUDA_OK(cudaGetLastError());
    return 0;
}

template <typename R>
static int q1_l2_fwd(const R* gx, int64_t Nx, const R* gy, int64_t Ny, const R* uf, const R* x, const R* target, int64_t M, R* r, int32_t* ix,
                     int32_t* iy, R* partial, R* loss, void* s) {
    HIDENN_REQUIRE(Nx >= 2 && Ny >= 2 && M > 0, "q1_l2_fwd: needs M > 0 samples and grids of at least 2 nodes");
    HIDENN_REQUIRE(gx && gy && uf && x && target && r && ix && iy && partial && loss, "q1_l2_fwd: NULL");
    int grid = 0;
    if (q1_fwd_launch<R, true>(gx, Nx, gy, Ny, uf, x, target, M, r, ix, iy, partial, (cudaStream_t)s, &grid)) return 1;
    q1_l2_finish_kernel<R><<<1, 256, 0, (cudaStream_t)s>>>(partial, grid, M, loss);
    HIDENN_CUDA_OK(cudaGetLastError());
    return 0;
}

template <typename R>
static int q1_fwd(const R* gx, int64_t Nx, const R* gy, int64_t Ny, const R* uf, const R* x, int64_t M, R* u, int32_t* ix, int32_t* iy, void* s) {
    HIDENN_REQUIRE(Nx >= 2 && Ny >= 2, "q1_interp_fwd: both grids need at least 2 nodes");
    if (M <= 0) return 0;
    HIDENN_REQUIRE(gx && gy && uf && x && u, "q1_interp_fwd: NULL");
    if (q1_fwd_launch<R, false>(gx, Nx, gy, Ny, uf, x, nullptr, M, u, ix, iy, nullptr, (cudaStream_t)s)) return 1;
    HIDENN_CUDA_OK(cudaGetLastError());
    return 0;
}

template <typename R>
static int q1_bwd(const R* gx, int64_t Nx, const R* gy, int64_t Ny, const R* uf, const R* x, const int32_t* ix, const int32_t* iy, const R* r,
                  int64_t M, R* rows, void* s) {
    (void)Nx;
    if (M <= 0) return 0;
    HIDENN_REQUIRE(gx && gy && uf && x && ix && iy && r && rows, "q1_interp_bwd: NULL");
    q1_bwd_kernel<R><<<grid_for(M), 256, 0, (cudaStream_t)s>>>(gx, gy, Ny, uf, (const typename Real2<R>::type*)x, ix, iy, r, M, rows);
    HIDENN_CUDA_OK(cudaGetLastError());
    return 0;
}

template <typename R>
static int q1_fold(const R* rows, const int64_t* order, const int64_t* seg, int64_t Nx, int64_t Ny, R* cell_tmp, R* du, R* dgx, R* dgy, void* s) {
    HIDENN_REQUIRE(Nx >= 2 && Ny >= 2 && order && seg && cell_tmp && du && dgx && dgy, "q1_fold_rows: bad arguments");
    cudaStream_t st = (cudaStream_t)s;
    const int64_t ncell = (Nx - 1) * (Ny - 1);
    q1_fold_cells_kernel<R><<<grid_for(ncell), 256, 0, st>>>(rows, order, seg, ncell, cell_tmp);
    q1_fold_nodes_kernel<R><<<grid_for(Nx * Ny), 256, 0, st>>>(cell_tmp, Nx, Ny, du);
    q1_fold_lines<R>(cell_tmp, Nx, Ny, dgx, dgy, st);
    HIDENN_CUDA_OK(cudaGetLastError());
    return 0;
}

template <typename R>
static int q1_bwd_fused(const R* gx, int64_t Nx, const R* gy, int64_t Ny, const R* uf, const R* x, const R* r, int64_t M, const int32_t* seg,
                        const int32_t* order, R* cell_tmp, R* du, R* dgx, R* dgy, void* s) {
    HIDENN_REQUIRE(Nx >= 2 && Ny >= 2 && gx && gy && uf && seg && cell_tmp && du && dgx && dgy, "q1_bwd_fused: bad arguments");
    HIDENN_REQUIRE(M == 0 || (x && r && order), "q1_bwd_fused: NULL sample arrays");
    cudaStream_t st = (cudaStream_t)s;
    const int64_t ncell = (Nx - 1) * (Ny - 1);
    q1_fused_cells_kernel<R><<<grid_for(ncell), 256, 0, st>>>(gx, gy, Nx, Ny, uf, (const typename Real2<R>::type*)x, r, seg, order, cell_tmp);
    q1_fold_nodes_kernel<R><<<grid_for(Nx * Ny), 256, 0, st>>>(cell_tmp, Nx, Ny, du);
    q1_fold_lines<R>(cell_tmp, Nx, Ny, dgx, dgy, st);
    HIDENN_CUDA_OK(cudaGetLastError());
    return 0;
}

}  // namespace hidenn

using namespace hidenn;

extern "C" int64_t hidenn_q1_l2_partials(void) { return kQ1L2MaxCtas; }
extern "C" int hidenn_q1_l2_fwd_f64(const double* gx, int64_t Nx, const double* gy, int64_t Ny, const double* uf, const double* x,
                                    const double* t, int64_t M, double* r, int32_t* ix, int32_t* iy, double* partial, double* loss, void* s) {
    return q1_l2_fwd<double>(gx, Nx, gy, Ny, uf, x, t, M, r, ix, iy, partial, loss, s);
}
extern "C" int hidenn_q1_l2_fwd_f32(const float* gx, int64_t Nx, const float* gy, int64_t Ny, const float* uf, const float* x,
                                    const float* t, int64_t M, float* r, int32_t* ix, int32_t* iy, float* partial, float* loss, void* s) {
    return q1_l2_fwd<float>(gx, Nx, gy, Ny, uf, x, t, M, r, ix, iy, partial, loss, s);
}

extern "C" int hidenn_q1_bin_count(const int32_t* ix, const int32_t* iy, int64_t M, int64_t Ny, int32_t* cnt, void* s) {
    if (M <= 0) return 0;
    HIDENN_REQUIRE(ix && iy && cnt && Ny >= 2, "q1_bin_count: bad arguments");
    q1_bin_count_kernel<<<grid_for(M), 256, 0, (cudaStream_t)s>>>(ix, iy, M, Ny - 1, cnt);
    HIDENN_CUDA_OK(cudaGetLastError());
    return 0;
}
extern "C" int hidenn_q1_bin_scatter(const int32_t* ix, const int32_t* iy, int64_t M, int64_t Nx, int64_t Ny, const int32_t* seg,
                                     int32_t* cursor, int32_t* order, void* s) {
    HIDENN_REQUIRE(Nx >= 2 && Ny >= 2 && seg, "q1_bin_scatter: bad arguments");
    if (M <= 0) return 0;
    HIDENN_REQUIRE(ix && iy && cursor && order, "q1_bin_scatter: NULL");
    q1_bin_scatter_kernel<<<grid_for(M), 256, 0, (cudaStream_t)s>>>(ix, iy, M, Ny - 1, seg, cursor, order);
    q1_bin_sort_kernel<<<grid_for((Nx - 1) * (Ny - 1)), 256, 0, (cudaStream_t)s>>>((Nx - 1) * (Ny - 1), seg, order);
    HIDENN_CUDA_OK(cudaGetLastError());
    return 0;
}
extern "C" int hidenn_q1_bwd_fused_f64(const double* gx, int64_t Nx, const double* gy, int64_t Ny, const double* uf, const double* x,
                                       const double* r, int64_t M, const int32_t* seg, const int32_t* order, double* tmp, double* du,
                                       double* dgx, double* dgy, void* s) {
    return q1_bwd_fused<double>(gx, Nx, gy, Ny, uf, x, r, M, seg, order, tmp, du, dgx, dgy, s);
}
extern "C" int hidenn_q1_bwd_fused_f32(const float* gx, int64_t Nx, const float* gy, int64_t Ny, const float* uf, const float* x,
                                       const float* r, int64_t M, const int32_t* seg, const int32_t* order, float* tmp, float* du,
                                       float* dgx, float* dgy, void* s) {
    return q1_bwd_fused<float>(gx, Nx, gy, Ny, uf, x, r, M, seg, order, tmp, du, dgx, dgy, s);
}

#define HIDENN_Q1_API(SUF, T)                                                                                                       \
    extern "C" int hidenn_q1_interp_fwd_##SUF(const T* gx, int64_t Nx, const T* gy, int64_t Ny, const T* uf, const T* x, int64_t M,  \
                                              T* u, int32_t* ix, int32_t* iy, void* s) {                                            \
        return q1_fwd<T>(gx, Nx, gy, Ny, uf, x, M, u, ix, iy, s);                                                                   \
    }                                                                                                                               \
    extern "C" int hidenn_q1_interp_bwd_##SUF(const T* gx, int64_t Nx, const T* gy, int64_t Ny, const T* uf, const T* x,             \
                                              const int32_t* ix, const int32_t* iy, const T* r, int64_t M, T* rows, void* s) {      \
        return q1_bwd<T>(gx, Nx, gy, Ny, uf, x, ix, iy, r, M, rows, s);                                                             \
    }                                                                                                                               \
    extern "C" int hidenn_q1_fold_rows_##SUF(const T* rows, const int64_t* o, const int64_t* sg, int64_t Nx, int64_t Ny, T* tmp,     \
                                             T* du, T* dgx, T* dgy, void* s) {                                                      \
        return q1_fold<T>(rows, o, sg, Nx, Ny, tmp, du, dgx, dgy, s);                                                               \
    }

HIDENN_Q1_API(f64, double)
HIDENN_Q1_API(f32, float)
