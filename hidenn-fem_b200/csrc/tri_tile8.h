// Tile kernel v8 (FP64, tile-ordered numberings): launch interface used by tri_energy.cu.
#pragma once
#include <cuda_runtime.h>
#include <stddef.h>
struct hidenn_tri_plan;
namespace hidenn { struct P2PLossArgs; }

namespace hidenn {
size_t tile8_smem_bytes(const hidenn_tri_plan* p);
// tiles [tile_begin, tile_end): energies to scratch[tile] (domain) and scratch[n_tiles + tile] (edge), final gradient rows
// to gx / gu.  Without HIDENN_TILES_ONLY the last CTA also reduces ALL tile energies into out[0..3] (the range must then be
// the whole plan); `ticket` is a zero-initialised counter the kernel leaves at zero.
int tile8_launch(const hidenn_tri_plan* p, const double* x_free, const double* x_fixed, const double* u_free, const double* u_fixed,
                 const double* consts, const double* t_table, int flags, double* out, double* gx, double* gu, double* gt, double* scratch,
                 unsigned* ticket, cudaStream_t stream, int tile_begin, int tile_end);
// Warp-specialised variant (tri_tile9.cu): same arguments and results; one persistent CTA of 768 threads per SM
// (16 element warps at 112 registers, 7 fold warps + 1 loader warp at 32), needs tile9_fits(plan).
size_t tile9_smem_bytes(const hidenn_tri_plan* p);
bool tile9_fits(const hidenn_tri_plan* p);
// first_done (may be NULL): device counter every fold warp increments once per finished tile among the plan's first
// n_first_tiles (multi-GPU overlap: it reaches n_first_tiles * tile9_fold_warps(plan) when the shared rows are final);
// reserve_sms: leave that many SMs to kernels running beside this one.
int tile9_launch(const hidenn_tri_plan* p, const double* x_free, const double* x_fixed, const double* u_free, const double* u_fixed,
                 const double* consts, const double* t_table, int flags, double* out, double* gx, double* gu, double* gt, double* scratch,
                 unsigned* ticket, cudaStream_t stream, int tile_begin, int tile_end, unsigned* first_done = nullptr, int reserve_sms = 0,
                 const P2PLossArgs* loss_args = nullptr);      // loss_args: exchange the loss partials over peer memory in the kernel's tail
int tile9_fold_warps(const hidenn_tri_plan* p);
// fixed-order reduction of the tile energies alone (after ranged launches)
int tile8_reduce(const hidenn_tri_plan* p, double* scratch, double* out, cudaStream_t stream);
}  // namespace hidenn
