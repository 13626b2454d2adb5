/*
 * hidenn_b200_grid.h -- C-ABI of the 1D and structured-2D (tensor-product) paths.
 *
 * Reference: PiecewiseLinearShapeNN (src/models.py:6-90), the structured
 * PiecewiseLinearShapeNN2D (src/models.py:93-212), the 1D bar energy of
 * examples/example3.py:27-70 and the L2-projection losses of examples/example1.py:37-38,
 * examples/example2.py:45-46.  Same conventions as hidenn_b200.h (device pointers, status
 * codes, stream as void*, no CPU fallback).  `_f64` / `_f32` variants; R = double / float.
 */
#ifndef HIDENN_B200_GRID_H
#define HIDENN_B200_GRID_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

/* scratch reals needed by the scan-based grid kernels for n increments */
int64_t hidenn_1d_scratch_size(int64_t n);

/* r-adaptive grid (src/models.py:45-53, :146-155):
 *   inc = max(softplus(p), 1e-6); cum = cumsum(inc); grid = [x0, x0 + (xN-x0)*cum/cum[-1]]
 * p dev [n] (n = N-1 increments), x0/xN dev scalars, grid dev [n+1], cum dev [n] (kept for backward). */
int hidenn_1d_grid_fwd_f64(const double* p, int64_t n, const double* x0, const double* xN, double* grid, double* cum,
                           double* scratch, void* stream);
int hidenn_1d_grid_fwd_f32(const float* p, int64_t n, const float* x0, const float* xN, float* grid, float* cum,
                           float* scratch, void* stream);
/* chain rule d grid -> d p: dgrid dev [n+1] (entry 0 ignored), dp dev [n] */
int hidenn_1d_grid_bwd_f64(const double* dgrid, const double* p, const double* cum, const double* x0, const double* xN,
                           int64_t n, double* dp, double* scratch, void* stream);
int hidenn_1d_grid_bwd_f32(const float* dgrid, const float* p, const float* cum, const float* x0, const float* xN,
                           int64_t n, float* dp, float* scratch, void* stream);

/* element lookup, bit-exact with clamp(searchsorted(grid, x) - 1, 0, N-2) (src/models.py:73-74, Q14) */
int hidenn_1d_lookup_f64(const double* grid, int64_t N, const double* x, int64_t M, int32_t* elem, void* stream);
int hidenn_1d_lookup_f32(const float* grid, int64_t N, const float* x, int64_t M, int32_t* elem, void* stream);

/* forward (src/models.py:70-90): u[m] = u_e N1 + u_{e+1} N2; also returns elem [M] and slope du/dx [M] */
int hidenn_1d_interp_fwd_f64(const double* grid, int64_t N, const double* u_full, const double* x, int64_t M,
                             double* u, int32_t* elem, double* slope, void* stream);
int hidenn_1d_interp_fwd_f32(const float* grid, int64_t N, const float* u_full, const float* x, int64_t M,
                             float* u, int32_t* elem, float* slope, void* stream);
/* per-row VJP pieces for cotangents r_u (of u) and r_s (of the slope), either may be NULL:
 * rows dev [M,4] = (d u_e, d u_{e+1}, d g_e, d g_{e+1}),  dx dev [M] = r_u * slope (or NULL) */
int hidenn_1d_interp_bwd_f64(const double* grid, int64_t N, const double* u_full, const double* x, const int32_t* elem,
                             const double* r_u, const double* r_s, int64_t M, double* rows, double* dx, void* stream);
int hidenn_1d_interp_bwd_f32(const float* grid, int64_t N, const float* u_full, const float* x, const int32_t* elem,
                             const float* r_u, const float* r_s, int64_t M, float* rows, float* dx, void* stream);
/* deterministic fold: rows grouped by element through `order` (stable sort of elem) and seg [N] (row range of
 * element e is [seg[e], seg[e+1]) ), then node k = left part of element k + right part of element k-1 */
int hidenn_1d_fold_rows_f64(const double* rows, const int64_t* order, const int64_t* seg, int64_t N, double* elem_tmp,
                            double* du_full, double* dgrid, void* stream);
int hidenn_1d_fold_rows_f32(const float* rows, const int64_t* order, const int64_t* seg, int64_t N, float* elem_tmp,
                            float* du_full, float* dgrid, void* stream);

/* Fused 1D bar energy forward + backward (examples/example3.py:27-70):
 *   loss = sum_e sum_q wq (0.5 E u'^2 - b(xq) u),  xq,wq from the grid but detached (Q15)
 * grid dev [N], u_full dev [N], xi/wi dev [ng] (ng <= 8), b_table dev [N-1,ng] or NULL = built-in
 * b_force of examples/example3.py:16-24.  Outputs: loss dev [1], du_full dev [N], dgrid dev [N],
 * flag dev [1] int32: set to 1 if some Gauss point's lookup is not its own element (degenerate grid; the
 * gradients are then invalid and the caller must use the generic path).  scratch: hidenn_1d_scratch_size(N). */
int hidenn_1d_bar_energy_f64(const double* grid, int64_t N, const double* u_full, const double* xi, const double* wi,
                             int ng, double E, const double* b_table, int need_grad, double* loss, double* du_full,
                             double* dgrid, int32_t* flag, double* scratch, void* stream);
int hidenn_1d_bar_energy_f32(const float* grid, int64_t N, const float* u_full, const float* xi, const float* wi,
                             int ng, float E, const float* b_table, int need_grad, float* loss, float* du_full,
                             float* dgrid, int32_t* flag, float* scratch, void* stream);

/* Fused r-adaptive bar step (examples/example3.py:27-70 over src/models.py:45-90): from the increments p [n = N-1] and the
 * trainable values u_free (node k reads u_free[k - (u0 != NULL)]; u0 / uN = fixed end values on the device, or NULL) to
 * loss [1], d loss / d p [n] and d loss / d u_free in THREE launches (block sums + scan by the last block; elements, node
 * folds, gamma sums, energy by the last block; softplus / cumsum / normalise chain).  gam [n] receives d loss / d grid[1:].
 * scratch: hidenn_1d_bar_step_scratch(n) reals, ZERO before the first call (the kernels leave it reusable).
 * *flag != 0 afterwards: a Gauss point fell outside its own element (degenerate grid) -- the result is then invalid. */
int64_t hidenn_1d_bar_step_scratch(int64_t n);
int hidenn_1d_bar_step_f64(const double* p, int64_t n, const double* x0, const double* xN, const double* u_free, const double* u0,
                           const double* uN, const double* xi, const double* wi, int ng, double E, const double* b_table,
                           double* loss, double* dp, double* du_free, double* gam, int32_t* flag, double* scratch, void* stream);
int hidenn_1d_bar_step_f32(const float* p, int64_t n, const float* x0, const float* xN, const float* u_free, const float* u0,
                           const float* uN, const float* xi, const float* wi, int ng, float E, const float* b_table,
                           float* loss, float* dp, float* du_free, float* gam, int32_t* flag, float* scratch, void* stream);

/* structured Q1 forward (src/models.py:180-212): x dev [M,2] -> u dev [M], ix/iy dev [M] */
int hidenn_q1_interp_fwd_f64(const double* gx, int64_t Nx, const double* gy, int64_t Ny, const double* u_full,
                             const double* x, int64_t M, double* u, int32_t* ix, int32_t* iy, void* stream);
int hidenn_q1_interp_fwd_f32(const float* gx, int64_t Nx, const float* gy, int64_t Ny, const float* u_full,
                             const float* x, int64_t M, float* u, int32_t* ix, int32_t* iy, void* stream);
/* per-row VJP pieces: rows dev [M,8] = (d u00,d u10,d u01,d u11, d gx_i, d gx_{i+1}, d gy_j, d gy_{j+1}) */
int hidenn_q1_interp_bwd_f64(const double* gx, int64_t Nx, const double* gy, int64_t Ny, const double* u_full,
                             const double* x, const int32_t* ix, const int32_t* iy, const double* r, int64_t M,
                             double* rows, void* stream);
int hidenn_q1_interp_bwd_f32(const float* gx, int64_t Nx, const float* gy, int64_t Ny, const float* u_full,
                             const float* x, const int32_t* ix, const int32_t* iy, const float* r, int64_t M,
                             float* rows, void* stream);
/* deterministic fold: rows grouped by cell (ix*(Ny-1)+iy) through order/seg [ncell+1]; cell_tmp dev 8*ncell reals (scratch);
 * du_full dev [Nx,Ny]; dgx dev [Nx], dgy dev [Ny] */
int hidenn_q1_fold_rows_f64(const double* rows, const int64_t* order, const int64_t* seg, int64_t Nx, int64_t Ny,
                            double* cell_tmp, double* du_full, double* dgx, double* dgy, void* stream);
int hidenn_q1_fold_rows_f32(const float* rows, const int64_t* order, const int64_t* seg, int64_t Nx, int64_t Ny,
                            float* cell_tmp, float* du_full, float* dgx, float* dgy, void* stream);

/* Sort-free deterministic cell binning for the structured backward:
 *   count:   cell_count[c] += 1 for every sample (integer atomics; cell_count dev [ncell] int32, zeroed by the caller)
 *   scatter: order[seg[c] + k] = sample ids of cell c (seg dev [ncell+1] int32 = exclusive prefix of the counts,
 *            cursor dev [ncell] int32 zeroed by the caller), then every cell's ids are sorted ascending in place,
 *            so the summation order of the fold is the sample index order -- independent of the atomics' arrival order. */
int hidenn_q1_bin_count(const int32_t* ix, const int32_t* iy, int64_t M, int64_t Ny, int32_t* cell_count, void* stream);
int hidenn_q1_bin_scatter(const int32_t* ix, const int32_t* iy, int64_t M, int64_t Nx, int64_t Ny, const int32_t* seg,
                          int32_t* cursor, int32_t* order, void* stream);
/* Fused structured backward: per cell, the VJP pieces of its samples (recomputed from x, r; never materialised) are
 * summed in `order`, then folded to nodes / grid lines exactly like hidenn_q1_fold_rows. */
int hidenn_q1_bwd_fused_f64(const double* gx, int64_t Nx, const double* gy, int64_t Ny, const double* u_full,
                            const double* x, const double* r, int64_t M, const int32_t* seg, const int32_t* order,
                            double* cell_tmp, double* du_full, double* dgx, double* dgy, void* stream);
int hidenn_q1_bwd_fused_f32(const float* gx, int64_t Nx, const float* gy, int64_t Ny, const float* u_full,
                            const float* x, const float* r, int64_t M, const int32_t* seg, const int32_t* order,
                            float* cell_tmp, float* du_full, float* dgx, float* dgy, void* stream);

/* Fused structured L2-projection loss  mean((model(x) - target)^2)  (examples/example2.py:45-46; example1.py:37-38 for
 * the same expression): one pass computes the lookups, the interpolation, the residual weights
 *   r[m] = 2 (u_h(x_m) - target[m]) / M        (dev [M], = d loss / d u_h)
 * and per-CTA partial sums of the squared residuals; a second one-CTA kernel adds the partials in CTA order ->
 * loss dev [1].  partial dev [hidenn_q1_l2_partials()] reals.  The backward is hidenn_q1_bwd_fused_* on r
 * (scaled by autograd's grad_output through hidenn_scale_inplace_*). */
int64_t hidenn_q1_l2_partials(void);
int hidenn_q1_l2_fwd_f64(const double* gx, int64_t Nx, const double* gy, int64_t Ny, const double* u_full, const double* x,
                         const double* target, int64_t M, double* r, int32_t* ix, int32_t* iy, double* partial, double* loss,
                         void* stream);
int hidenn_q1_l2_fwd_f32(const float* gx, int64_t Nx, const float* gy, int64_t Ny, const float* u_full, const float* x,
                         const float* target, int64_t M, float* r, int32_t* ix, int32_t* iy, float* partial, float* loss,
                         void* stream);

#ifdef __cplusplus
}
#endif
#endif /* HIDENN_B200_GRID_H */
