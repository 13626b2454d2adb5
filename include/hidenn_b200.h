/*
 * hidenn_b200.h -- C-ABI of the B200-native HiDeNN-FEM quadrature hot path.
 *
 * The reference (achraf-15/HiDeNN-FEM) has no FFI: its boundary for this path is the Python
 * class API (src/models.py, src/loss.py).  This header is the native boundary a binding of
 * that API calls; every entry point names the reference lines it replaces.  The Python
 * mirror (hidenn-fem_b200/models.py, loss.py) binds it with ctypes; INTEGRATION.md shows the
 * stub a reference maintainer would add.
 *
 * Conventions
 *   - extern "C", plain pointers and sizes, no torch types.
 *   - Every function returns 0 on success, non-zero on error; hidenn_last_error() then
 *     returns a thread-local message.
 *   - "host"  pointers are ordinary CPU memory; "dev" pointers are CUDA device memory on the
 *     plan's device.  No buffer ownership is transferred; nothing on the hot path allocates.
 *   - `stream` is a cudaStream_t passed as void*; all hot-path calls are asynchronous on it.
 *   - There is no CPU fallback: without a CUDA device every compute entry point fails.
 *   - Node coordinates / displacements are AoS pairs [N,2] exactly like the reference's
 *     Parameters `node_coords_free`, `u_free` (src/models.py:261,274).
 */
#ifndef HIDENN_B200_H
#define HIDENN_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define HIDENN_B200_VERSION 100

/* layout of the `consts` device array handed to the triangle energy kernels (in the compute type) */
#define HIDENN_TRI_C00 0      /* symmetrised plane-stress matrix 0.5*(C+C^T) of src/loss.py:29-32, 6 entries: */
#define HIDENN_TRI_C01 1      /*   C00 C01 C02 C11 C12 C22                                                     */
#define HIDENN_TRI_C02 2
#define HIDENN_TRI_C11 3
#define HIDENN_TRI_C12 4
#define HIDENN_TRI_C22 5
#define HIDENN_TRI_W 6        /* sum of the triangle Gauss weights (src/utils.py:13-81; 0.25 for order 4/6) */
#define HIDENN_TRI_FB 7       /* Fb[k][i] = sum_g w_g N_k(g) b_i(g), 6 entries (src/loss.py:80-81,86)         */
#define HIDENN_TRI_TX 13      /* uniform traction (src/loss.py:47-51): (1e5, 0)                              */
#define HIDENN_TRI_TY 14
#define HIDENN_TRI_NG1 15     /* number of 1D Gauss points on an edge (<= 8), stored as a real               */
#define HIDENN_TRI_XI1 16     /* 8 slots: raw leggauss points on [-1,1] (src/utils.py:4-11)                  */
#define HIDENN_TRI_W1 24      /* 8 slots: raw leggauss weights                                               */
#define HIDENN_TRI_NCONST 32

/* flags */
#define HIDENN_NEED_GX 1      /* write d loss / d node_coords_free */
#define HIDENN_NEED_GU 2      /* write d loss / d u_free           */
#define HIDENN_WITH_EDGES 4   /* include the Neumann edge term (src/loss.py:91-110) */
#define HIDENN_TILES_ONLY 16  /* measurement aid: launch only the tile kernel (per-tile energies stay in scratch) */
#define HIDENN_HINT_NO_BODY_FORCE 32   /* caller guarantees consts[HIDENN_TRI_FB..] == 0 (b_force=None): skips those terms */
#define HIDENN_HINT_C_PLANE_STRESS 64  /* caller guarantees C02 == C12 == 0 (the matrix form of src/loss.py:29-32)      */

const char* hidenn_last_error(void);
int hidenn_version(void);
/* number of CUDA devices visible; <0 on error (used by the loader to fail loudly) */
int hidenn_device_count(void);

/* Measurement aid: DFMA instructions per second (thread level) this GPU sustains on independent chains -- the measured
 * FP64-pipe peak bench.py prints the kernels' FP64 utilisation against.  Synchronises `stream`. */
int hidenn_fp64_peak(double* dfma_per_s, void* stream);

/* ------------------------------------------------------------------------------------------
 * Triangle plan: static topology preprocessing, once per mesh (connectivity never changes
 * during r-adaptation; src/models.py:252 keeps it as a buffer).  Replaces, for the hot path,
 * the per-call masked index_put assembly of src/models.py:292-305 (slot maps), the
 * connectivity gather of src/models.py:228-238 (tile-local uint packs) and autograd's
 * scatter-add backward (node->element CSR inside each tile).
 *
 *   conn            host [Ne,3] int64   (src/models.py:252)
 *   coords_init     host [Nn,2] double  initial coordinates, used ONLY for locality ordering
 *   boundary_mask   host [Nn] uint8     1 = coordinate fixed (src/models.py:256-263)
 *   dirichlet_mask  host [Nn] uint8     1 = displacement fixed (src/models.py:266-271)
 *   edges           host [Ned,2] int64  Neumann edges (src/models.py:280-282); may be NULL/0
 *   tile_nodes      owned nodes per tile; 0 = default for `real_bytes` (8 or 4)
 * ------------------------------------------------------------------------------------------ */
typedef struct hidenn_tri_plan hidenn_tri_plan;

int hidenn_tri_plan_create(const int64_t* conn, int64_t n_elems, int64_t n_nodes,
                           const double* coords_init,
                           const uint8_t* boundary_mask, const uint8_t* dirichlet_mask,
                           const int64_t* edges, int64_t n_edges,
                           int tile_nodes, int real_bytes, int device,
                           hidenn_tri_plan** out);
/* Same, with `first_nodes` [n_first] (may be NULL/0): tiles owning one of these nodes are listed first
 * (tiles [0, layout8[7])).  Multi-GPU callers pass the nodes shared with other ranks, run the tiles in two ranges
 * (hidenn_tri_energy_range_*) and exchange the finished shared rows while the second range computes. */
/* options: HIDENN_PLAN_JINV_TRANSPOSE = correct-math switch (default off): every kernel of this plan forms the physical
 * shape-function gradients with J^-T (dN/dx = J^-T dN/dxi) instead of the reference's J^-1 (src/models.py:339-351,
 * SURVEY Q1); gradients follow consistently. */
#define HIDENN_PLAN_JINV_TRANSPOSE 1
int hidenn_tri_plan_create_ex(const int64_t* conn, int64_t n_elems, int64_t n_nodes,
                              const double* coords_init,
                              const uint8_t* boundary_mask, const uint8_t* dirichlet_mask,
                              const int64_t* edges, int64_t n_edges,
                              const int64_t* first_nodes, int64_t n_first, int options,
                              int tile_nodes, int real_bytes, int device,
                              hidenn_tri_plan** out);
void hidenn_tri_plan_destroy(hidenn_tri_plan* plan);

/* Locality ordering for meshes with arbitrary numbering (the step before the plan; gmsh / meshzoo output,
 * src/mesh.py:125-153, 252-276): recursive coordinate bisection of the nodes into the same tiles the plan uses,
 * nodes numbered tile by tile -- inside a tile by class (coordinate free & displacement free, coordinate fixed &
 * displacement free, both fixed, coordinate free & displacement fixed), then by descending valence -- so that every
 * tile's rows of node_coords_free / u_free (and of their gradients) are ONE contiguous run per array.
 * hidenn_tri_plan_create recognises such a numbering ("tile-ordered") and then stages tiles with bulk copies.
 * new_to_old [Nn]: old index of new node i; elem_new_to_old [Ne]: elements listed by smallest new node id.
 * Corner order inside every element is the caller's to keep (the reference's results depend on it). */
int hidenn_tri_locality_order(const int64_t* conn, int64_t n_elems, int64_t n_nodes, const double* coords,
                              const uint8_t* boundary_mask, const uint8_t* dirichlet_mask,
                              const int64_t* edges, int64_t n_edges, int tile_nodes,
                              int64_t* new_to_old, int64_t* elem_new_to_old);

/* layout8[0]: bit 0 = the plan found a tile-ordered numbering (FP64 plans; bulk-copy tile kernels in use), bit 1 = the plan
 * is pairs-only (its tiles are sized for the paired layout: only the warp-specialised kernel can run it, and
 * hidenn_tri_plan_fold_tables / _bank_stats refuse it; HIDENN_PLAN_PAIRS=0 builds the other kind), [1]=max halo
 * nodes per tile, [2]=Neumann edge visits, [3]=dynamic smem bytes of the two-CTA kernel, [4]=element pairs of the global
 * matching, [5]=pair-or-single entries over all tiles, [6]=max fold slots per tile in the paired layout,
 * [7]=number of leading tiles that own the caller's first_nodes (hidenn_tri_plan_create_ex). */
int hidenn_tri_plan_layout(const hidenn_tri_plan* plan, int64_t* layout8);

/* Mesh ingestion: mask [Nn] <- 1 for the nodes on the topological boundary of the triangle mesh (edges that belong to
 * exactly one element: outer boundary and hole rims; what src/mesh.py:70-88, 202-215 take from gmsh / the hole test). */
int hidenn_mesh_boundary_nodes(const int64_t* conn, int64_t n_elems, int64_t n_nodes, uint8_t* mask);

/* info[0]=n_tiles [1]=tile element visits (incl. halo recompute) [2]=tile node visits
 * [3]=max local nodes/tile [4]=max fold entries/tile [5]=scratch elements needed
 * [6]=dynamic smem bytes f64 [7]=same f32 [8]=n_free_x [9]=n_free_u [10]=n_edges [11]=n_edge_nodes
 * [12]=plan bytes on device [13]=max elements/tile [14]=n_elems [15]=n_nodes */
int hidenn_tri_plan_info(const hidenn_tri_plan* plan, int64_t* info16);

/* Locality of the node numbering as the tile kernels see it: out2[0] = mean number of contiguous node_coords_free /
 * node_coords_fixed row runs per tile, out2[1] = mean local nodes per tile.  ~1-10 runs: tile-ordered; tens: a Z-curve
 * or natural numbering (fine); close to out2[1]: no locality (the kernels then run ~3x slower -- renumber the mesh with
 * hidenn_tri_locality_order first). */
int hidenn_tri_plan_locality(const hidenn_tri_plan* plan, double* out2);

/* Bit-exact integer views for tests (host copies): slot maps of src/models.py:292-305.
 * xslot[n] >= 0: row of node n in node_coords_free; < 0: ~row in node_coords_fixed. Same for uslot. */
int hidenn_tri_plan_slots(const hidenn_tri_plan* plan, int32_t* xslot_host, int32_t* uslot_host);
/* For every tile element visit: global element id and the three global node ids it will gather
 * (decoded from the tile packs) -- lets tests check connectivity indexing bit-exactly.
 * out_elem [visits], out_nodes [visits,3], out_owner [visits] (1 = this visit adds the energy). */
int hidenn_tri_plan_decode(const hidenn_tri_plan* plan, int64_t* out_elem, int64_t* out_nodes, uint8_t* out_owner);

/* Measurement aid: when dev_buf != NULL every tile CTA of later hidenn_tri_energy_* launches records
 * phase clocks into dev_buf[16*tile ..] (see profiles/phase_timing.py for the layout of the kernel variant in use). */
int hidenn_debug_tile_timing(long long* dev_buf);

/* Host-side model of the tile kernel's shared-memory passes for the plan's lane assignment:
 * out4 = {gather passes, ideal gather passes, partial-store passes, ideal store passes}. */
int hidenn_tri_plan_bank_stats(const hidenn_tri_plan* plan, int real_bytes, int64_t* out4);
/* Same model for the node staging / gradient flush (8 consecutive memory-order records per pass, each at its local id):
 * out2 = {passes, ideal passes}. */
int hidenn_tri_plan_stage_stats(const hidenn_tri_plan* plan, int64_t* out2);

/* Tile membership (tests): node_off [n_tiles+1] into nodes [info[2] = node_visits] (global node ids, owned nodes of a
 * tile first), n_owned [n_tiles].  Tiles are listed by ascending smallest owned node id. */
int hidenn_tri_plan_tiles(const hidenn_tri_plan* plan, int64_t* node_off, int32_t* n_owned, int32_t* nodes);
/* Raw fold tables (tests): per tile t, element visits [elem_off[t], elem_off[t+1]) with their 64-bit packs (3 x 10-bit
 * local ids, 3 x 11-bit fold-slot positions, bit 63 = owns the element's energy) and global element ids; per owned node
 * (owned_off[t] + l) the word  slot_start | valence << 16;  n_entries[t] = number of fold slots (= the dump position). */
int hidenn_tri_plan_fold_tables(const hidenn_tri_plan* plan, int64_t* elem_off, uint64_t* packs, int64_t* elems,
                                int64_t* owned_off, uint32_t* entry_off, int32_t* n_entries);
/* Paired layout of a tile-ordered plan (tests): per tile t, entries [pent_off[t], pent_off[t+1]) of two 64-bit words each
 * (elem_pack format; first word = element with the smaller id, second word = its partner of the global matching or the
 * null word with local ids 1023; corners of the second element whose partial is merged into the first carry the dump
 * position n_entries9[t]); entry_off9 per owned node (owned_off[t] + l) = slot_start | slots << 16;  mate [Ne] = partner
 * element of the global matching or -1. */
int hidenn_tri_plan_pair_tables(const hidenn_tri_plan* plan, int64_t* pent_off, uint64_t* packs, int64_t* owned_off,
                                uint32_t* entry_off9, int32_t* n_entries9, int32_t* mate);

/* Row-block tables of the host-buffer pipeline (hidenn_tri_energy_host_*): the free rows are cut into 64 blocks of
 * rows2[0] (node_coords_free) / rows2[1] (u_free) rows; first_need[b] = first tile reading a row of block b
 * (INT32_MAX: none), last_own[b] = last tile writing one (-1: none).  Each array has 64 entries. */
int hidenn_tri_plan_pipeline(const hidenn_tri_plan* plan, int32_t* rows2, int32_t* first_need_x, int32_t* last_own_x,
                             int32_t* first_need_u, int32_t* last_own_u);

/* ------------------------------------------------------------------------------------------
 * Fused energy + gradients:  EnergyLoss2D.__call__ (src/loss.py:113-116) =
 * domain_energy (src/loss.py:55-88) over PiecewiseLinearShapeNN2D.forward
 * (src/models.py:316-357) minus edge_energy (src/loss.py:91-110, src/models.py:359-376),
 * and what loss.backward() leaves in node_coords_free.grad / u_free.grad.
 *
 *   x_free  dev [Nfree,2]   x_fixed dev [Nfixed,2]   (src/models.py:261-262)
 *   u_free  dev [Nufree,2]  u_fixed dev [Nufixed,2]  (src/models.py:274-277, broadcast by caller)
 *   consts  dev [HIDENN_TRI_NCONST]
 *   t_table dev [Ned,ng1,2] traction at the physical edge Gauss points, or NULL = uniform
 *   out     dev [4]: loss, domain energy, edge energy, (unused)
 *   gx_free dev [Nfree,2], gu_free dev [Nufree,2]: overwritten (every row), not accumulated
 *   gt_out  dev [Ned,ng1,2] or NULL: d loss / d t_table (for tractions that depend on x)
 *   scratch dev [info[5]] reals, ZERO-INITIALISED before its first use (the kernels leave it reusable);
 *           one evaluation in flight per scratch buffer
 * ------------------------------------------------------------------------------------------ */
int hidenn_tri_energy_f64(const hidenn_tri_plan* plan,
                          const double* x_free, const double* x_fixed,
                          const double* u_free, const double* u_fixed,
                          const double* consts, const double* t_table, int flags,
                          double* out, double* gx_free, double* gu_free, double* gt_out,
                          double* scratch, void* stream);
int hidenn_tri_energy_f32(const hidenn_tri_plan* plan,
                          const float* x_free, const float* x_fixed,
                          const float* u_free, const float* u_fixed,
                          const float* consts, const float* t_table, int flags,
                          float* out, float* gx_free, float* gu_free, float* gt_out,
                          float* scratch, void* stream);

/* The same evaluation in pieces, for callers that overlap something with it (multi-GPU halo exchange): tiles
 * [tile_begin, tile_end) only -- their energies go to scratch, their owned gradient rows are final when the launch ends
 * -- and, after all ranges, hidenn_tri_energy_finish_* reduces the tile energies into out[0..3].  Tile-ordered FP64
 * plans only (the Neumann edge term is part of the tiles there); others return an error. */
int hidenn_tri_energy_range_f64(const hidenn_tri_plan* plan,
                                const double* x_free, const double* x_fixed,
                                const double* u_free, const double* u_fixed,
                                const double* consts, const double* t_table, int flags,
                                double* gx_free, double* gu_free, double* gt_out,
                                double* scratch, int tile_begin, int tile_end, void* stream);
int hidenn_tri_energy_finish_f64(const hidenn_tri_plan* plan, double* scratch, double* out, void* stream);
/* One launch with a progress signal: like hidenn_tri_energy_f64, and the kernel adds to the device counter *first_done
 * (zero before the launch) as the plan's leading tiles (hidenn_tri_plan_create_ex first_nodes) finish; when it reaches
 * hidenn_tri_plan_overlap_target(plan) the gradient rows of those nodes are final and a kernel on ANOTHER stream
 * (hidenn_halo_p2p_push_* with wait_counter) may read them while the remaining tiles still compute.  reserve_sms SMs are
 * left free for that kernel.  With peer_bufs != NULL (arguments of hidenn_halo_p2p_loss_*) the last CTA also exchanges
 * the rank's out[0..2] with the other ranks over peer memory, so out holds the GLOBAL sums when the launch ends.
 * hidenn_tri_plan_overlap_target returns 0 when the plan cannot signal (not tile-ordered, no first nodes, tiles too
 * large for the warp-specialised kernel): use the ranged calls above instead. */
int hidenn_tri_energy_overlap_f64(const hidenn_tri_plan* plan,
                                  const double* x_free, const double* x_fixed,
                                  const double* u_free, const double* u_fixed,
                                  const double* consts, const double* t_table, int flags,
                                  double* out, double* gx_free, double* gu_free, double* gt_out,
                                  double* scratch, uint32_t* first_done, int reserve_sms,
                                  void* const* peer_bufs, void* my_buf, int me, int world, int64_t smax, uint64_t* loss_step,
                                  void* stream);
int hidenn_tri_plan_overlap_target(const hidenn_tri_plan* plan);
/* Which gradient tile kernel hidenn_tri_energy_* launches for this plan: 7 = tri_tile_persistent_kernel (any numbering;
 * FP32), 8 = tri_tile8_kernel, 9 = tri_tile9_kernel (tile-ordered FP64 plans; 8 when the tiles do not fit the
 * warp-specialised kernel's shared memory or HIDENN_TILE_WS=0). */
int hidenn_tri_plan_kernel(const hidenn_tri_plan* plan);

/* Host-buffer convenience (the end-to-end drop-in for a CPU caller): copies the four parameter
 * arrays host->device, runs the fused step, copies loss and both gradients back and waits.
 * gx_free / gu_free may be NULL (gradients stay on the device, only out[4] comes back).
 * On large meshes the copies are pipelined with the tile kernels in row blocks over two plan-owned
 * copy streams (hidenn_tri_plan_pipeline); host buffers must be pinned for the copies to overlap
 * (pageable buffers work, serialised).  Device buffers live in a plan-owned arena: one call in
 * flight per plan.  Results are bit-identical to hidenn_tri_energy_* on resident buffers. */
int hidenn_tri_energy_host_f64(hidenn_tri_plan* plan,
                               const double* x_free_h, const double* x_fixed_h,
                               const double* u_free_h, const double* u_fixed_h,
                               const double* consts_h, int flags,
                               double* out_h, double* gx_free_h, double* gu_free_h, void* stream);
int hidenn_tri_energy_host_f32(hidenn_tri_plan* plan,
                               const float* x_free_h, const float* x_fixed_h,
                               const float* u_free_h, const float* u_fixed_h,
                               const float* consts_h, int flags,
                               float* out_h, float* gx_free_h, float* gu_free_h, void* stream);

/* g[i] *= *scale for i<n unless *scale == 1 (then every block exits at once): applies autograd's
 * grad_output to gradients computed in the forward launch without a host sync. */
int hidenn_scale_inplace2_f64(double* g1, int64_t n1, double* g2, int64_t n2, const double* scale_dev, void* stream);
int hidenn_scale_inplace2_f32(float* g1, int64_t n1, float* g2, int64_t n2, const float* scale_dev, void* stream);
int hidenn_scale_inplace_f64(double* g, int64_t n, const double* scale_dev, void* stream);
int hidenn_scale_inplace_f32(float* g, int64_t n, const float* scale_dev, void* stream);

/* ------------------------------------------------------------------------------------------
 * Generic pointwise forward(x_ref, elem_id) and its VJP (src/models.py:316-357; callers
 * src/loss.py:65, src/plots.py:183-187).
 *   elem_id dev [M] int64, x_ref dev [M,2]
 *   u_h dev [M,2], detJ dev [M] (signed), grad_u dev [M,2,2]
 * VJP: cotangents cu [M,2], cd [M], cG [M,2,2] (any may be NULL) -> per-row corner contributions
 *   row_gx dev [M,3,2], row_gu dev [M,3,2]  (folded to nodes by hidenn_tri_fold_rows)
 * ------------------------------------------------------------------------------------------ */
int hidenn_tri_eval_fwd_f64(const hidenn_tri_plan* plan, const double* x_free, const double* x_fixed,
                            const double* u_free, const double* u_fixed,
                            const double* x_ref, const int64_t* elem_id, int64_t M,
                            double* u_h, double* detJ, double* grad_u, void* stream);
int hidenn_tri_eval_fwd_f32(const hidenn_tri_plan* plan, const float* x_free, const float* x_fixed,
                            const float* u_free, const float* u_fixed,
                            const float* x_ref, const int64_t* elem_id, int64_t M,
                            float* u_h, float* detJ, float* grad_u, void* stream);
int hidenn_tri_eval_bwd_f64(const hidenn_tri_plan* plan, const double* x_free, const double* x_fixed,
                            const double* u_free, const double* u_fixed,
                            const double* x_ref, const int64_t* elem_id, int64_t M,
                            const double* cu, const double* cd, const double* cG,
                            double* row_gx, double* row_gu, void* stream);
int hidenn_tri_eval_bwd_f32(const hidenn_tri_plan* plan, const float* x_free, const float* x_fixed,
                            const float* u_free, const float* u_fixed,
                            const float* x_ref, const int64_t* elem_id, int64_t M,
                            const float* cu, const float* cd, const float* cG,
                            float* row_gx, float* row_gu, void* stream);
/* Deterministic fold of per-row corner contributions to the parameter gradients.
 *   order dev [M] int64: a permutation that sorts rows by elem_id (stable), seg dev [Ne+1] int64:
 *   row range of each element in that order.  Rows of one element are summed in that order,
 *   elements of one node in ascending element id (global node->element CSR held by the plan). */
int hidenn_tri_fold_rows_f64(const hidenn_tri_plan* plan, const double* row_gx, const double* row_gu,
                             const int64_t* order, const int64_t* seg, int64_t M,
                             double* elem_tmp, double* gx_free, double* gu_free, void* stream);
int hidenn_tri_fold_rows_f32(const hidenn_tri_plan* plan, const float* row_gx, const float* row_gu,
                             const int64_t* order, const int64_t* seg, int64_t M,
                             float* elem_tmp, float* gx_free, float* gu_free, void* stream);
/* Edge branch forward(x, edge_id, edge=True) (src/models.py:359-376): u_h [M,2], ds [M]. */
int hidenn_tri_edge_fwd_f64(const hidenn_tri_plan* plan, const double* x_free, const double* x_fixed,
                            const double* u_free, const double* u_fixed,
                            const double* xi, const int64_t* edge_id, int64_t M,
                            double* u_h, double* ds, void* stream);
int hidenn_tri_edge_fwd_f32(const hidenn_tri_plan* plan, const float* x_free, const float* x_fixed,
                            const float* u_free, const float* u_fixed,
                            const float* xi, const int64_t* edge_id, int64_t M,
                            float* u_h, float* ds, void* stream);

/* Gather full arrays: coords / u_full properties (src/models.py:292-305) without masked index_put. */
int hidenn_tri_assemble_f64(const hidenn_tri_plan* plan, int which /*0=coords,1=u*/,
                            const double* free_vals, const double* fixed_vals, double* full, void* stream);
int hidenn_tri_assemble_f32(const hidenn_tri_plan* plan, int which,
                            const float* free_vals, const float* fixed_vals, float* full, void* stream);

/* ------------------------------------------------------------------------------------------
 * Halo pack / unpack for element-block partitions (new capability, the
 * reference is single-device).  idx dev [n] int32 rows of a [N,2] gradient array.
 * pack:   buf[2*i..] = g[idx[i]]        unpack: g[idx[i]] = buf[2*i..]
 * ------------------------------------------------------------------------------------------ */
int hidenn_halo_pack_f64(const double* g, const int32_t* idx, int64_t n, double* buf, void* stream);
int hidenn_halo_unpack_f64(double* g, const int32_t* idx, int64_t n, const double* buf, void* stream);
int hidenn_halo_pack_f32(const float* g, const int32_t* idx, int64_t n, float* buf, void* stream);
int hidenn_halo_unpack_f32(float* g, const int32_t* idx, int64_t n, const float* buf, void* stream);

/* One-launch pack / unpack of the whole halo message  buf = [loss, 0 | gx pairs (S) | gu pairs (S)]:
 *   pack:   buf[0] = *loss;  bufx[xpos[i]] = gx[xrows[i]];  bufu[upos[i]] = gu[urows[i]]   (other entries untouched)
 *   unpack: *loss = buf[0];  gx[xrows[i]] = bufx[xpos[i]];  gu[urows[i]] = bufu[upos[i]]
 * gx / gu may be NULL (frozen parameter).  S = number of shared nodes; xpos/upos index pairs.
 * pack also clears `zero_other` (another message buffer of the same size, or NULL): with two alternating buffers the
 * in-place all-reduce never needs a separate copy or memset. */
int hidenn_halo_pack_all_f64(const double* gx, const int32_t* xrows, const int32_t* xpos, int64_t nx,
                             const double* gu, const int32_t* urows, const int32_t* upos, int64_t nu,
                             const double* loss, int64_t S, double* buf, double* zero_other, void* stream);
int hidenn_halo_unpack_all_f64(double* gx, const int32_t* xrows, const int32_t* xpos, int64_t nx,
                               double* gu, const int32_t* urows, const int32_t* upos, int64_t nu,
                               double* loss, int64_t S, const double* buf, void* stream);
int hidenn_halo_pack_all_f32(const float* gx, const int32_t* xrows, const int32_t* xpos, int64_t nx,
                             const float* gu, const int32_t* urows, const int32_t* upos, int64_t nu,
                             const float* loss, int64_t S, float* buf, float* zero_other, void* stream);
int hidenn_halo_unpack_all_f32(float* gx, const int32_t* xrows, const int32_t* xpos, int64_t nx,
                               float* gu, const int32_t* urows, const int32_t* upos, int64_t nu,
                               float* loss, int64_t S, const float* buf, void* stream);

/* ------------------------------------------------------------------------------------------
 * Halo exchange over NVLink peer memory (one box).  Every rank allocates one buffer of
 * hidenn_halo_p2p_bytes(world, smax, real_bytes) bytes in memory all ranks map (torch symmetric memory / CUDA IPC),
 * zero-initialised; peer_bufs [world] (device array) holds every rank's buffer address as mapped in THIS process.
 * smax = largest number of nodes any two ranks share; the k-th node (ascending global id) two ranks share uses slot k
 * in both directions.  Two device counters starting at 1 number the steps (graph-capturable): `grad_step` for the
 * gradient channel (read by push, advanced by pull) and `step` of the loss call for the loss channel (advanced by it).
 *   push:  (first spin until *wait_counter >= wait_target and reset it to 0, when wait_counter != NULL;) for i < n_send:
 *          put (gx[s_xrow[i]], gu[s_urow[i]]) (rows < 0: zeros) into rank s_peer[i]'s buffer, slot s_k[i] of sender
 *          `me`; then raise my flag on every peer.
 *   pull:  wait for the flags of wait_ranks [n_wait]; for shared node j < n_nodes with sources [n_off[j], n_off[j+1])
 *          (src_rank ascending, including `me` itself; src_k = slot in the pair's list): gradient rows n_xrow[j] /
 *          n_urow[j] (< 0: none) <- sum over the sources in that order.  Every holder computes the same bits.
 *   loss:  out[0..2] <- sum over ranks (ascending) of every rank's out[0..2]; advances *step.
 * One push -> pull pair and one loss call per step; the loss call is independent of the pair (it may also run in the
 * tail of the tile kernel: hidenn_tri_energy_overlap_*).
 * ------------------------------------------------------------------------------------------ */
int64_t hidenn_halo_p2p_bytes(int world, int64_t smax, int real_bytes);
int hidenn_halo_p2p_push_f64(const double* gx, const double* gu, const int32_t* s_xrow, const int32_t* s_urow,
                             const int32_t* s_peer, const int32_t* s_k, int64_t n_send, void* const* peer_bufs,
                             int me, int world, int64_t smax, const uint64_t* grad_step,
                             uint32_t* wait_counter, uint32_t wait_target, void* stream);
int hidenn_halo_p2p_pull_f64(double* gx, double* gu, const int32_t* n_xrow, const int32_t* n_urow, const int32_t* n_off,
                             const int32_t* src_rank, const int32_t* src_k, int64_t n_nodes, const int32_t* wait_ranks,
                             int n_wait, void* my_buf, int me, int world, int64_t smax, uint64_t* grad_step, void* stream);
int hidenn_halo_p2p_loss_f64(double* out, void* const* peer_bufs, void* my_buf, int me, int world, int64_t smax,
                             uint64_t* step, void* stream);
int hidenn_halo_p2p_push_f32(const float* gx, const float* gu, const int32_t* s_xrow, const int32_t* s_urow,
                             const int32_t* s_peer, const int32_t* s_k, int64_t n_send, void* const* peer_bufs,
                             int me, int world, int64_t smax, const uint64_t* grad_step,
                             uint32_t* wait_counter, uint32_t wait_target, void* stream);
int hidenn_halo_p2p_pull_f32(float* gx, float* gu, const int32_t* n_xrow, const int32_t* n_urow, const int32_t* n_off,
                             const int32_t* src_rank, const int32_t* src_k, int64_t n_nodes, const int32_t* wait_ranks,
                             int n_wait, void* my_buf, int me, int world, int64_t smax, uint64_t* grad_step, void* stream);
int hidenn_halo_p2p_loss_f32(float* out, void* const* peer_bufs, void* my_buf, int me, int world, int64_t smax,
                             uint64_t* step, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* HIDENN_B200_H */
