// Fused triangle energy forward + backward (sm_100a, FP64 / FP32 CUDA cores; HBM-bound gather/reduce,
// deliberately no tensor cores).
//
// One launch replaces EnergyLoss2D.__call__ + loss.backward() of the reference
// (/root/reference/src/loss.py:55-116 over /root/reference/src/models.py:292-376):
//   gather connectivity + nodal coordinates/values (staged in shared memory)
//   -> shape functions / J / det J / J^-1 / B-matrix -> plane-stress energy density
//   -> element energy (block reduction, fixed order) and the 12 nodal gradient partials
//   -> per-tile node->element CSR fold in shared memory (deterministic, no float atomics)
//   -> coalesced AoS pair stores straight into the Parameter-layout gradient arrays.
// Per-element formulas: SURVEY.md Appendix A.1 with  d psi/dJ = -(P Jinv)^T G  (DESIGN.md §4).
#include "../../include/hidenn_b200.h"
#include "common.cuh"
#include "tri_plan.h"
#include "tri_element.cuh"
#include "tri_tile8.h"
#include "halo_p2p.cuh"

#include <cstdlib>
#include <type_traits>

namespace hidenn {

// ---------------------------------------------------------------------------------------------
// The tile kernel is persistent: each CTA walks tiles blockIdx.x, +gridDim.x, ... and overlaps the
// whole load chain of the NEXT tile with the element / fold phases of the current one:
//   * node pairs of tile t+1 are gathered with cp.async (LDGSTS) into the second node buffer while tile t
//     computes (their slots were loaded one iteration earlier);
//   * packs / fold offsets of tile t+1 and the slots of tile t+2 are loaded into registers during the fold of
//     tile t, when register pressure is low.
// Two block barriers per tile; the tile-energy partials ride on the next barrier (double-buffered).
// ---------------------------------------------------------------------------------------------
template <typename R, bool BODY, bool ISO, int MINB, int BLOCK>
__global__ void __launch_bounds__(BLOCK, MINB)
tri_tile_persistent_kernel(const TriPlanDev P, const typename Real2<R>::type* __restrict__ x_free,
                           const typename Real2<R>::type* __restrict__ x_fixed, const typename Real2<R>::type* __restrict__ u_free,
                           const typename Real2<R>::type* __restrict__ u_fixed, const R* __restrict__ consts, const int flags,
                           typename Real2<R>::type* __restrict__ gx_free, typename Real2<R>::type* __restrict__ gu_free,
                           R* __restrict__ tile_energy, long long* __restrict__ timing) {
    using R2 = typename Real2<R>::type;
    extern __shared__ __align__(16) unsigned char smem_raw[];
    // shared layout: 2 x (xy | uv) node buffers, fold partial pairs gu | gx (+ dump slot), output staging (gu | gx per
    // owned node), 2 x 16 warp energy partials
    R2* s_node = reinterpret_cast<R2*>(smem_raw);
    const int nb = 2 * P.max_local;                       // pairs per node buffer
    const PartBuf<R> part(s_node + 2 * nb, P.max_entries + 1);
    const PartBuf<R> outb(s_node + 2 * nb + 2 * (P.max_entries + 1), P.max_owned);
    R* s_red = reinterpret_cast<R*>(s_node + 2 * nb + 2 * (P.max_entries + 1) + 2 * P.max_owned);      // [2][16]
    const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
    const int nct = gridDim.x;
    int tile = blockIdx.x;
    if (tile >= P.n_tiles) return;
    constexpr unsigned LM = (1u << kLidBits) - 1u, PM = (1u << kPosBits) - 1u;
    constexpr unsigned G = 8u;
    constexpr int PB = (int)sizeof(R2);
    constexpr int NPRE = 768 / BLOCK;      // element packs held in registers per thread (768 >= typical tile)
    constexpr int NW = BLOCK / 32;

    // node records come in memory order: record j of a tile is (slots, local id = shared-memory position)
    auto load_slots = [&](const int t, int2 (&sl)[2], unsigned (&ld)[2]) {
        const int2* __restrict__ src = P.t_slots + (size_t)t * P.stride_local;
        const uint16_t* __restrict__ lsrc = P.t_lid + (size_t)t * P.stride_local;
#pragma unroll
        for (int k = 0; k < 2; ++k) {
            const int i = tid + k * BLOCK;
            sl[k] = i < P.stride_local ? __ldg(src + i) : make_int2(0, 0);
            ld[k] = i < P.stride_local ? (unsigned)__ldg(lsrc + i) : 0xFFFFu;
        }
    };
    auto issue_gathers = [&](const int t, const int2 (&sl)[2], const unsigned (&ld)[2], R2* buf) {
        const NodeBuf<R> nbuf(buf, P.max_local);
#pragma unroll
        for (int k = 0; k < 2; ++k) {
            if (ld[k] != 0xFFFFu) {      // consecutive lanes read consecutive Parameter rows; the record lands at its local id
                cp_async_pair(nbuf.xy_ptr(ld[k]), sl[k].x >= 0 ? (const void*)(x_free + sl[k].x) : (const void*)(x_fixed + (~sl[k].x)), PB);
                cp_async_pair(nbuf.uv_ptr(ld[k]), sl[k].y >= 0 ? (const void*)(u_free + sl[k].y) : (const void*)(u_fixed + (~sl[k].y)), PB);
            }
        }
        const int2* __restrict__ src = P.t_slots + (size_t)t * P.stride_local;
        const uint16_t* __restrict__ lsrc = P.t_lid + (size_t)t * P.stride_local;
        for (int i = tid + 2 * BLOCK; i < P.stride_local; i += BLOCK) {     // tiles with more than 2*BLOCK local nodes
            const unsigned l2 = __ldg(lsrc + i);
            if (l2 == 0xFFFFu) continue;
            const int2 s2 = __ldg(src + i);
            cp_async_pair(nbuf.xy_ptr(l2), s2.x >= 0 ? (const void*)(x_free + s2.x) : (const void*)(x_fixed + (~s2.x)), PB);
            cp_async_pair(nbuf.uv_ptr(l2), s2.y >= 0 ? (const void*)(u_free + s2.y) : (const void*)(u_fixed + (~s2.y)), PB);
        }
    };
    // final gradient stores of a finished tile: records in memory order, values from the output staging buffer
    auto flush_outputs = [&](const int t, const int n_owned, const int2 (&sl)[2], const unsigned (&ld)[2]) {
#pragma unroll
        for (int k = 0; k < 2; ++k) {
            if (ld[k] < (unsigned)n_owned) {
                R2 gu, gx;
                outb.load(ld[k], gu, gx);
                if ((flags & HIDENN_NEED_GU) && sl[k].y >= 0) gu_free[sl[k].y] = gu;
                if ((flags & HIDENN_NEED_GX) && sl[k].x >= 0) gx_free[sl[k].x] = gx;
            }
        }
        const int2* __restrict__ src = P.t_slots + (size_t)t * P.stride_local;
        const uint16_t* __restrict__ lsrc = P.t_lid + (size_t)t * P.stride_local;
        for (int i = tid + 2 * BLOCK; i < P.stride_local; i += BLOCK) {
            const unsigned l2 = __ldg(lsrc + i);
            if (l2 >= (unsigned)n_owned) continue;
            const int2 s2 = __ldg(src + i);
            R2 gu, gx;
            outb.load(l2, gu, gx);
            if ((flags & HIDENN_NEED_GU) && s2.y >= 0) gu_free[s2.y] = gu;
            if ((flags & HIDENN_NEED_GX) && s2.x >= 0) gx_free[s2.x] = gx;
        }
    };
    auto load_meta = [&](const int t, unsigned long long (&pk)[NPRE], uint32_t (&of)[2]) {
        const unsigned long long* __restrict__ packs = P.elem_pack + (size_t)t * P.stride_elem;
        const uint32_t* __restrict__ offs = P.entry_off + (size_t)t * P.stride_owned;
#pragma unroll
        for (int k = 0; k < NPRE; ++k) {
            const int i = tid + k * BLOCK;
            pk[k] = i < P.stride_elem ? __ldg(packs + i) : 0ull;
        }
#pragma unroll
        for (int k = 0; k < 2; ++k) {
            const int i = tid + k * BLOCK;
            of[k] = i < P.stride_owned ? __ldg(offs + i) : 0u;
        }
    };

    // prologue: first tile's slots -> gathers; second tile's slots; first tile's packs / offsets / descriptor
    int2 slot_cur[2], slot_nxt[2];
    unsigned lid_cur[2], lid_nxt[2];
    unsigned long long pk[NPRE];
    uint32_t off[2];
    load_slots(tile, slot_cur, lid_cur);
    if (tile + nct < P.n_tiles) load_slots(tile + nct, slot_nxt, lid_nxt);
    else { slot_nxt[0] = slot_nxt[1] = make_int2(0, 0); lid_nxt[0] = lid_nxt[1] = 0xFFFFu; }
    load_meta(tile, pk, off);
    TileDesc td = P.tiles[tile];
    const TriConsts<R> K = load_consts<R, BODY>(consts);
    issue_gathers(tile, slot_cur, lid_cur, s_node);
    cp_async_wait_all();
    __syncthreads();

    int b = 0;
    int prev_tile = -1, prev_owned = 0;
    int2 slot_prev[2] = {make_int2(0, 0), make_int2(0, 0)};
    unsigned lid_prev[2] = {0xFFFFu, 0xFFFFu};
    for (;;) {
        // node buffer b holds this tile; the partial buffer is free (the previous fold ended before the barrier)
        if (tid == 0 && prev_tile >= 0) {        // energy of the previous tile, summed in fixed warp order
            const R* r = s_red + (b ^ 1) * 16;
            R acc = R(0);
#pragma unroll
            for (int w = 0; w < NW; ++w) acc += r[w];
            tile_energy[prev_tile] = acc;
        }
        long long tk0 = 0, tk1 = 0, tk2 = 0, tk3 = 0;
        if (timing) tk0 = clock64();
        const int tnext = tile + nct;
        const bool has_next = tnext < P.n_tiles;
        if (has_next) issue_gathers(tnext, slot_nxt, lid_nxt, s_node + (b ^ 1) * nb);      // lands during E + F of this tile
        if (prev_tile >= 0) flush_outputs(prev_tile, prev_owned, slot_prev, lid_prev);      // coalesced stores of the previous tile

        // E: elements -> energy + gradient partials at their fold slots
        const NodeBuf<R> nodes(s_node + b * nb, P.max_local);
        R e_acc = R(0);
        const unsigned dumpv = (unsigned)td.n_entries;
        auto do_element = [&](const unsigned long long w) {
            const unsigned lo = (unsigned)w, hi = (unsigned)(w >> 32);
            const unsigned l0 = lo & LM, l1 = (lo >> kLidBits) & LM, l2 = (lo >> (2 * kLidBits)) & LM;
            const unsigned p0 = (unsigned)(w >> (3 * kLidBits)) & PM, p1 = (hi >> (3 * kLidBits + kPosBits - 32)) & PM,
                           p2 = (hi >> (3 * kLidBits + 2 * kPosBits - 32)) & PM;
            R e;
            R2 gu[3], gx[3], v0, v1, v2, U0, U1, U2;
            nodes.load(l0, v0, U0); nodes.load(l1, v1, U1); nodes.load(l2, v2, U2);
            tri_element<R, BODY, ISO>(v0, v1, v2, U0, U1, U2, K, e, gu, gx, P.jinv_t != 0);
            e_acc += (hi >> 31) ? e : R(0);
            // halo corners carry the tile's dump position: skip their stores (predicated, no branch)
            if (p0 != dumpv) part.store(p0, gu[0], gx[0]);
            if (p1 != dumpv) part.store(p1, gu[1], gx[1]);
            if (p2 != dumpv) part.store(p2, gu[2], gx[2]);
        };
#pragma unroll
        for (int k = 0; k < NPRE; ++k)
            if (tid + k * BLOCK < td.n_elem) do_element(pk[k]);
        {
            const unsigned long long* __restrict__ packs = P.elem_pack + (size_t)tile * P.stride_elem;
            for (int i = tid + NPRE * BLOCK; i < td.n_elem; i += BLOCK) do_element(__ldg(packs + i));
        }
        e_acc = warp_sum(e_acc);
        if (lane == 0) s_red[b * 16 + wid] = e_acc;
        if (timing) tk1 = clock64();
        __syncthreads();
        if (timing) tk2 = clock64();

        // F: fold.  First put the next tile's metadata loads in flight (registers are cheap in this phase).
        unsigned long long pk_n[NPRE];
        uint32_t off_n[2];
        int2 slot_n2[2];
        unsigned lid_n2[2];
        TileDesc td_n = td;
        if (has_next) {
            load_meta(tnext, pk_n, off_n);
            td_n = P.tiles[tnext];
            if (tnext + nct < P.n_tiles) load_slots(tnext + nct, slot_n2, lid_n2);
            else { slot_n2[0] = slot_n2[1] = make_int2(0, 0); lid_n2[0] = lid_n2[1] = 0xFFFFu; }
        }
        // thread l folds owned node l (its slot rows are conflict-free) into the output staging buffer
        auto fold_node = [&](const uint32_t oc, const int l) {
            const unsigned fb = oc & 0xFFFFu, fe = fb + (oc >> 16) * G;
            R ax = R(0), ay = R(0), bx = R(0), by = R(0);
            for (unsigned k = fb; k < fe; k += G) {
                R2 u, x;
                part.load(k, u, x);
                ax += u.x; ay += u.y; bx += x.x; by += x.y;
            }
            outb.store(l, mk2<R>(ax, ay), mk2<R>(bx, by));
        };
#pragma unroll
        for (int k = 0; k < 2; ++k)
            if (tid + k * BLOCK < td.n_owned) fold_node(off[k], tid + k * BLOCK);
        {
            const uint32_t* __restrict__ offs = P.entry_off + (size_t)tile * P.stride_owned;
            for (int i = tid + 2 * BLOCK; i < td.n_owned; i += BLOCK) fold_node(__ldg(offs + i), i);
        }
        if (timing) tk3 = clock64();
        if (timing && (tid == 0 || tid == BLOCK - 1)) {     // per-tile phase clocks of the first and last warp (debug aid)
            unsigned smid;
            asm volatile("mov.u32 %0, %%smid;" : "=r"(smid));
            long long* tt = timing + 16 * (long long)tile + (tid == 0 ? 0 : 8);
            tt[0] = tk0; tt[1] = tk1; tt[2] = tk2; tt[3] = tk3; tt[4] = smid;
        }
        prev_tile = tile;
        prev_owned = td.n_owned;
#pragma unroll
        for (int k = 0; k < 2; ++k) { slot_prev[k] = slot_cur[k]; lid_prev[k] = lid_cur[k]; }
        if (!has_next) break;
        // rotate the pipeline registers
#pragma unroll
        for (int k = 0; k < 2; ++k) {
            slot_cur[k] = slot_nxt[k]; slot_nxt[k] = slot_n2[k]; off[k] = off_n[k];
            lid_cur[k] = lid_nxt[k]; lid_nxt[k] = lid_n2[k];
        }
#pragma unroll
        for (int k = 0; k < NPRE; ++k) pk[k] = pk_n[k];
        td = td_n;
        tile = tnext;
        b ^= 1;
        if (timing && (tid == 0 || tid == BLOCK - 1)) timing[16 * (long long)prev_tile + (tid == 0 ? 5 : 13)] = clock64();
        cp_async_wait_all();
        if (timing && (tid == 0 || tid == BLOCK - 1)) timing[16 * (long long)prev_tile + (tid == 0 ? 6 : 14)] = clock64();
        __syncthreads();      // next tile's nodes visible; fold reads of the partials and staging writes are done
        if (timing && (tid == 0 || tid == BLOCK - 1)) timing[16 * (long long)prev_tile + (tid == 0 ? 7 : 15)] = clock64();
    }
    __syncthreads();
    flush_outputs(prev_tile, prev_owned, slot_prev, lid_prev);
    if (tid == 0) {
        const R* r = s_red + b * 16;
        R acc = R(0);
#pragma unroll
        for (int w = 0; w < NW; ++w) acc += r[w];
        tile_energy[prev_tile] = acc;
    }
}

// Energy-only variant (torch.no_grad() evaluations, e.g. logging): no fold, no gradient traffic.
template <typename R>
__global__ void __launch_bounds__(kTileBlock)
tri_tile_energy_only_kernel(const TriPlanDev P, const typename Real2<R>::type* __restrict__ x_free,
                            const typename Real2<R>::type* __restrict__ x_fixed,
                            const typename Real2<R>::type* __restrict__ u_free,
                            const typename Real2<R>::type* __restrict__ u_fixed, const R* __restrict__ consts,
                            R* __restrict__ tile_energy) {
    using R2 = typename Real2<R>::type;
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const TileDesc td = P.tiles[blockIdx.x];
    R2* s_xy = reinterpret_cast<R2*>(smem_raw);
    R2* s_uv = s_xy + td.n_local;
    R* s_red = reinterpret_cast<R*>(s_uv + td.n_local);
    const int tid = threadIdx.x;
    const int2* __restrict__ slots = P.t_slots + (size_t)blockIdx.x * P.stride_local;
    const uint16_t* __restrict__ lids = P.t_lid + (size_t)blockIdx.x * P.stride_local;
    const unsigned long long* __restrict__ packs = P.elem_pack + (size_t)blockIdx.x * P.stride_elem;
    for (int i = tid; i < td.n_local; i += kTileBlock) {      // records in memory order land at their local id
        const int2 sl = __ldg(slots + i);
        const unsigned l = __ldg(lids + i);
        s_xy[l] = load_slot<R2>(x_free, x_fixed, sl.x);
        s_uv[l] = load_slot<R2>(u_free, u_fixed, sl.y);
    }
    const TriConsts<R> K = load_consts<R, true>(consts);
    __syncthreads();
    R e_acc = R(0);
    constexpr unsigned LM = (1u << kLidBits) - 1u;
    for (int i = tid; i < td.n_elem; i += kTileBlock) {
        const unsigned long long w = __ldg(packs + i);
        if (!((w >> kOwnerBit) & 1ull)) continue;
        const unsigned l0 = (unsigned)(w) & LM, l1 = (unsigned)(w >> kLidBits) & LM, l2 = (unsigned)(w >> (2 * kLidBits)) & LM;
        R e;
        R2 gu[3], gx[3];
        tri_element<R, true, false>(s_xy[l0], s_xy[l1], s_xy[l2], s_uv[l0], s_uv[l1], s_uv[l2], K, e, gu, gx, P.jinv_t != 0);
        e_acc += e;
    }
    const R tot = block_sum<R, kTileBlock>(e_acc, s_red);
    if (tid == 0) tile_energy[blockIdx.x] = tot;
}

// ---------------------------------------------------------------------------------------------
// Neumann edge term (src/loss.py:91-110, src/models.py:359-376) + final fixed-order reduction.
// One CTA: the edge count is O(sqrt(Ne)).  Edge nodes are folded node-centrically (each node by one
// thread, entries in ascending edge order) and *added* to the gradients the tile kernel stored.
// ---------------------------------------------------------------------------------------------
constexpr int kEdgeBlock = 256;
constexpr int kEdgeMaxCtas = 64;

template <typename R> struct EdgeEval {
    R ds, S;
    typename Real2<R>::type dir, x0, x1, U0, U1;
};

template <typename R>
__device__ __forceinline__ EdgeEval<R> edge_eval(const TriPlanDev& P, int e, const typename Real2<R>::type* x_free,
                                                 const typename Real2<R>::type* x_fixed,
                                                 const typename Real2<R>::type* u_free,
                                                 const typename Real2<R>::type* u_fixed, const R* consts,
                                                 const R* t_table, int ng1) {
    using R2 = typename Real2<R>::type;
    EdgeEval<R> E;
    const int4 s = __ldg(reinterpret_cast<const int4*>(P.e_slots) + e);
    E.x0 = load_slot<R2>(x_free, x_fixed, s.x); E.U0 = load_slot<R2>(u_free, u_fixed, s.y);
    E.x1 = load_slot<R2>(x_free, x_fixed, s.z); E.U1 = load_slot<R2>(u_free, u_fixed, s.w);
    const R dx = E.x1.x - E.x0.x, dy = E.x1.y - E.x0.y;
    E.ds = sqrt(dx * dx + dy * dy);
    E.dir = mk2<R>(dx / E.ds, dy / E.ds);
    R S = R(0);
    for (int q = 0; q < ng1; ++q) {
        const R xi = consts[HIDENN_TRI_XI1 + q], w = consts[HIDENN_TRI_W1 + q];
        const R ux = (R(1) - xi) * E.U0.x + xi * E.U1.x, uy = (R(1) - xi) * E.U0.y + xi * E.U1.y;
        R tx, ty;
        if (t_table) { tx = t_table[(e * ng1 + q) * 2]; ty = t_table[(e * ng1 + q) * 2 + 1]; }
        else { tx = consts[HIDENN_TRI_TX]; ty = consts[HIDENN_TRI_TY]; }
        S += w * (ux * tx + uy * ty);
    }
    E.S = S;
    return E;
}

// Multi-CTA: edges / edge nodes / tile energies are strided over the whole grid; every CTA leaves fixed-order
// FP64 partial sums in scratch, and the last CTA to finish (integer ticket) adds them in CTA order -> the result
// does not depend on which CTA happens to be last.  fin = scratch region after the tile energies:
// double part[2*kEdgeMaxCtas] | unsigned ticket (zero before the first launch, reset here).
template <typename R>
__global__ void __launch_bounds__(kEdgeBlock)
tri_edge_finalize_kernel(const TriPlanDev P, const typename Real2<R>::type* __restrict__ x_free,
                         const typename Real2<R>::type* __restrict__ x_fixed,
                         const typename Real2<R>::type* __restrict__ u_free,
                         const typename Real2<R>::type* __restrict__ u_fixed, const R* __restrict__ consts,
                         const R* __restrict__ t_table, const int flags, const R* __restrict__ tile_energy,
                         R* __restrict__ out, typename Real2<R>::type* __restrict__ gx_free,
                         typename Real2<R>::type* __restrict__ gu_free, R* __restrict__ gt_out, double* __restrict__ part,
                         unsigned* __restrict__ ticket, typename Real2<R>::type* __restrict__ en_final) {
    using R2 = typename Real2<R>::type;
    __shared__ double s_red[kEdgeBlock / 32];
    __shared__ unsigned s_last;
    const int tid = threadIdx.x;
    const int gtid = blockIdx.x * kEdgeBlock + tid, gstride = gridDim.x * kEdgeBlock;
    const bool with_edges = (flags & HIDENN_WITH_EDGES) && P.n_edges > 0;
    const int ng1 = with_edges ? (int)consts[HIDENN_TRI_NG1] : 0;

    double e_edge = 0.0;
    if (with_edges) {
        for (int e = gtid; e < P.n_edges; e += gstride) {
            const EdgeEval<R> E = edge_eval<R>(P, e, x_free, x_fixed, u_free, u_fixed, consts, t_table, ng1);
            e_edge += (double)(E.S * E.ds);
            if (gt_out) {   // d loss / d t_q = -w_q ds u_q
                for (int q = 0; q < ng1; ++q) {
                    const R xi = consts[HIDENN_TRI_XI1 + q], w = consts[HIDENN_TRI_W1 + q];
                    gt_out[(e * ng1 + q) * 2] = -w * E.ds * ((R(1) - xi) * E.U0.x + xi * E.U1.x);
                    gt_out[(e * ng1 + q) * 2 + 1] = -w * E.ds * ((R(1) - xi) * E.U0.y + xi * E.U1.y);
                }
            }
        }
        if (flags & (HIDENN_NEED_GX | HIDENN_NEED_GU)) {
            for (int k = gtid; k < P.n_enodes; k += gstride) {
                R gux = R(0), guy = R(0), gxx = R(0), gxy = R(0);
                for (int j = P.en_off[k]; j < P.en_off[k + 1]; ++j) {
                    const int ent = P.en_ent[j], e = ent >> 1, end = ent & 1;
                    const EdgeEval<R> E = edge_eval<R>(P, e, x_free, x_fixed, u_free, u_fixed, consts, t_table, ng1);
                    R fx = R(0), fy = R(0);
                    for (int q = 0; q < ng1; ++q) {
                        const R xi = consts[HIDENN_TRI_XI1 + q], w = consts[HIDENN_TRI_W1 + q];
                        const R nsh = end ? xi : (R(1) - xi);
                        R tx, ty;
                        if (t_table) { tx = t_table[(e * ng1 + q) * 2]; ty = t_table[(e * ng1 + q) * 2 + 1]; }
                        else { tx = consts[HIDENN_TRI_TX]; ty = consts[HIDENN_TRI_TY]; }
                        fx += w * nsh * tx; fy += w * nsh * ty;
                    }
                    gux -= E.ds * fx; guy -= E.ds * fy;
                    // d(-E_edge)/dx: end 0: +S*dir, end 1: -S*dir
                    const R sg = end ? -E.S : E.S;
                    gxx += sg * E.dir.x; gxy += sg * E.dir.y;
                }
                const int xs = P.en_xslot[k], us = P.en_uslot[k];
                // en_final (host-buffer pipeline): final rows of the edge nodes, compact, so the host can patch them in
                if ((flags & HIDENN_NEED_GU) && us >= 0) {
                    R2 g = gu_free[us]; g.x += gux; g.y += guy; gu_free[us] = g;
                    if (en_final) en_final[k] = g;
                }
                if ((flags & HIDENN_NEED_GX) && xs >= 0) {
                    R2 g = gx_free[xs]; g.x += gxx; g.y += gxy; gx_free[xs] = g;
                    if (en_final) en_final[P.n_enodes + k] = g;
                }
            }
        }
    }
    // domain energy: fixed-order strided + tree sum of the tile partials, accumulated in double
    double e_dom = 0.0;
    for (int t = gtid; t < P.n_tiles; t += gstride) e_dom += (double)tile_energy[t];
    const double dom = block_sum<double, kEdgeBlock>(e_dom, s_red);
    __syncthreads();
    const double edg = block_sum<double, kEdgeBlock>(e_edge, s_red);
    if (tid == 0) {
        part[2 * blockIdx.x] = dom;
        part[2 * blockIdx.x + 1] = edg;
        __threadfence();
        s_last = (atomicAdd(ticket, 1u) == gridDim.x - 1) ? 1u : 0u;
    }
    __syncthreads();
    if (s_last && tid == 0) {
        __threadfence();
        double d = 0.0, e = 0.0;
        const volatile double* vp = part;
        for (unsigned c = 0; c < gridDim.x; ++c) { d += vp[2 * c]; e += vp[2 * c + 1]; }
        out[0] = (R)(d - e);
        out[1] = (R)d;
        out[2] = (R)e;
        out[3] = R(0);
        *ticket = 0u;
    }
}

template <typename R>
__global__ void scale_inplace2_kernel(R* __restrict__ g1, int64_t n1, R* __restrict__ g2, int64_t n2, const R* __restrict__ scale) {
    const R s = __ldg(scale);
    if (s == R(1)) return;
    const int64_t i0 = (int64_t)blockIdx.x * blockDim.x + threadIdx.x, st = (int64_t)gridDim.x * blockDim.x;
    for (int64_t i = i0; i < n1; i += st) g1[i] *= s;
    for (int64_t i = i0; i < n2; i += st) g2[i] *= s;
}

template <typename R>
__global__ void scale_inplace_kernel(R* __restrict__ g, int64_t n, const R* __restrict__ scale) {
    const R s = __ldg(scale);
    if (s == R(1)) return;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) g[i] *= s;
}

template <typename R> [[maybe_unused]] static size_t smem_for(const hidenn_tri_plan* p) {
    return (size_t)p->dev.max_local * 4 * sizeof(R) + (size_t)(p->dev.max_entries + 1) * 4 * sizeof(R) + 128;
}

static long long* g_tile_timing = nullptr;     // set by hidenn_debug_tile_timing (measurement aid, not thread safe)

template <typename R, bool BODY, bool ISO, int MINB, int BLOCK>
static int launch_tile_persistent_mb(const hidenn_tri_plan* p, const R* x_free, const R* x_fixed, const R* u_free, const R* u_fixed,
                                     const R* consts, int flags, R* gx, R* gu, R* scratch, cudaStream_t stream, size_t smem,
                                     int tile_begin, int tile_end) {
    using R2 = typename Real2<R>::type;
    static size_t configured[kMaxDevices] = {};        // function attributes are per device
    size_t& cfg = configured[p->device % kMaxDevices];
    if (smem > cfg) {
        HIDENN_CUDA_OK(cudaFuncSetAttribute(tri_tile_persistent_kernel<R, BODY, ISO, MINB, BLOCK>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        HIDENN_CUDA_OK(cudaFuncSetAttribute(tri_tile_persistent_kernel<R, BODY, ISO, MINB, BLOCK>, cudaFuncAttributePreferredSharedMemoryCarveout, 100));
        cfg = smem;
    }
    const int n_sm = sm_count(p->device);
    // a tile range [tile_begin, tile_end) is run by shifting the fixed-stride record pointers: the kernel itself always
    // walks tiles 0 .. n_tiles-1 of the view it is given
    TriPlanDev P = p->dev;
    P.tiles += tile_begin;
    P.t_slots += (size_t)tile_begin * P.stride_local;
    P.t_lid += (size_t)tile_begin * P.stride_local;
    P.elem_pack += (size_t)tile_begin * P.stride_elem;
    P.entry_off += (size_t)tile_begin * P.stride_owned;
    P.n_tiles = tile_end - tile_begin;
    const int grid = std::min(P.n_tiles, n_sm * MINB);
    tri_tile_persistent_kernel<R, BODY, ISO, MINB, BLOCK><<<grid, BLOCK, smem, stream>>>(
        P, (const R2*)x_free, (const R2*)x_fixed, (const R2*)u_free, (const R2*)u_fixed, consts, flags, (R2*)gx, (R2*)gu,
        scratch + tile_begin, g_tile_timing ? g_tile_timing + 16 * (long long)tile_begin : nullptr);
    return 0;
}

// resident CTAs per SM the tile kernel is compiled for: as many as the tile's shared memory allows
// (227 KB per SM, 1 KB reserved per CTA), capped where the register budget would start to spill.
static int pick_minb(size_t smem, int real_bytes) {
    static const int env = [] { const char* e = getenv("HIDENN_TILE_MINB"); return e ? atoi(e) : 0; }();
    int mb = (int)((227 * 1024) / (smem + 1024));
    const int cap = real_bytes == 8 ? 4 : 5;
    mb = mb < 2 ? 2 : (mb > cap ? cap : mb);
    if (env >= 2 && env <= 6) mb = env;
    return mb;
}

template <typename R> static size_t smem_persistent_for(const hidenn_tri_plan* p) {
    return (size_t)p->dev.max_local * 8 * sizeof(R) + (size_t)(p->dev.max_entries + 1) * 4 * sizeof(R) +
           (size_t)p->dev.max_owned * 4 * sizeof(R) + 32 * sizeof(R) + 64;
}

template <typename R, bool BODY, bool ISO>
static int launch_tile(const hidenn_tri_plan* p, const R* x_free, const R* x_fixed, const R* u_free, const R* u_fixed, const R* consts,
                       int flags, R* gx, R* gu, R* scratch, cudaStream_t stream, int tile_begin, int tile_end) {
    const size_t smem = smem_persistent_for<R>(p);
#define HIDENN_LAUNCH_P(MB, BL) \
    return launch_tile_persistent_mb<R, BODY, ISO, MB, BL>(p, x_free, x_fixed, u_free, u_fixed, consts, flags, gx, gu, scratch, stream, smem, \
                                                           tile_begin, tile_end)
    switch (pick_minb(smem, (int)sizeof(R))) {      // 384- and 128-thread CTAs were measured no faster (profiles/README.md)
        case 2: HIDENN_LAUNCH_P(2, 256);
        case 3: HIDENN_LAUNCH_P(3, 256);
        case 4: HIDENN_LAUNCH_P(4, 256);
        default: HIDENN_LAUNCH_P(5, 256);
    }
#undef HIDENN_LAUNCH_P
}

// scratch layout: tile energies (one per tile; domain | edge, two per tile, for tile-ordered plans) | 8-byte aligned:
// double part[2*kEdgeMaxCtas] | unsigned ticket
static inline size_t scratch_energy_slots(const hidenn_tri_plan* p) { return (size_t)p->dev.n_tiles * (p->tile_order ? 2 : 1); }
template <typename R> static inline double* scratch_part(const hidenn_tri_plan* p, R* scratch) {
    const size_t off = ((scratch_energy_slots(p) + 8) * sizeof(R) + 7) / 8 * 8;
    return reinterpret_cast<double*>(reinterpret_cast<char*>(scratch) + off);
}

// kFinalizeOnly (internal): skip the tile kernels (the host-buffer pipeline has launched them range by range)
constexpr int kFinalizeOnly = 1 << 30;

template <typename R>
static int tri_energy_launch(const hidenn_tri_plan* p, const R* x_free, const R* x_fixed, const R* u_free, const R* u_fixed,
                             const R* consts, const R* t_table, int flags, R* out, R* gx, R* gu, R* gt, R* scratch,
                             void* stream_v, int tile_begin = 0, int tile_end = -1, R* en_final = nullptr, unsigned* first_done = nullptr,
                             int reserve_sms = 0, const P2PLossArgs* loss_args = nullptr) {
    using R2 = typename Real2<R>::type;
    HIDENN_REQUIRE(p != nullptr, "tri_energy: plan is NULL");
    HIDENN_REQUIRE(p->device >= 0, "tri_energy: host-only plan (device=-1) cannot run kernels; there is no CPU fallback");
    HIDENN_REQUIRE(consts && out && scratch, "tri_energy: NULL consts/out/scratch");
    HIDENN_REQUIRE(x_fixed || p->n_fixed_x == 0, "tri_energy: x_fixed NULL");
    HIDENN_REQUIRE(u_fixed || p->n_fixed_u == 0, "tri_energy: u_fixed NULL (the reference raises too when u_fixed=None)");
    HIDENN_REQUIRE(!(flags & HIDENN_NEED_GX) || gx, "tri_energy: gx_free NULL");
    HIDENN_REQUIRE(!(flags & HIDENN_NEED_GU) || gu, "tri_energy: gu_free NULL");
    cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_v);
    DeviceScope scope;
    HIDENN_CUDA_OK(scope.enter(p->device));
    const bool grad = flags & (HIDENN_NEED_GX | HIDENN_NEED_GU);
    if (tile_end < 0) tile_end = p->dev.n_tiles;
    if constexpr (std::is_same<R, double>::value) {
        if (grad && p->tile_order) {
            // tile-ordered numbering: bulk-copy tile kernel; edges and the final reduction are part of it
            unsigned* ticket = reinterpret_cast<unsigned*>(scratch_part<R>(p, scratch) + 2 * kEdgeMaxCtas);
            if (flags & kFinalizeOnly) return tile8_reduce(p, scratch, out, stream);
            const bool whole = tile_begin == 0 && tile_end == p->dev.n_tiles;
            HIDENN_REQUIRE(whole || (flags & HIDENN_TILES_ONLY), "tri_energy: a tile range needs HIDENN_TILES_ONLY");
            // HIDENN_TILE_WS=0: the two-CTA kernel of tri_tile8.cu instead of the warp-specialised one (A/B runs)
            static const bool ws_off = [] { const char* e = getenv("HIDENN_TILE_WS"); return e && atoi(e) == 0; }();
            if (tile_end > tile_begin) {
                HIDENN_REQUIRE(first_done == nullptr || (!ws_off && tile9_fits(p)),
                               "tri_energy_overlap: needs the warp-specialised tile kernel (hidenn_tri_plan_overlap_target > 0)");
                HIDENN_REQUIRE(p->unpaired_ok || (!ws_off && tile9_fits(p)),
                               "tri_energy: this plan's tiles are sized for the paired layout (kernel v9 only); build it with "
                               "HIDENN_PLAN_PAIRS=0 to run the other tile kernels");
                const int rc = (!ws_off && tile9_fits(p))
                                   ? tile9_launch(p, x_free, x_fixed, u_free, u_fixed, consts, t_table, flags, out, gx, gu, gt, scratch, ticket,
                                                  stream, tile_begin, tile_end, first_done, reserve_sms, loss_args)
                                   : tile8_launch(p, x_free, x_fixed, u_free, u_fixed, consts, t_table, flags, out, gx, gu, gt, scratch, ticket,
                                                  stream, tile_begin, tile_end);
                if (rc) return rc;
            }
            return 0;
        }
    }
    if (tile_end > tile_begin && !(flags & kFinalizeOnly)) {
        if (grad) {
            const bool body = !(flags & HIDENN_HINT_NO_BODY_FORCE), iso = (flags & HIDENN_HINT_C_PLANE_STRESS) != 0;
            int rc;
#define HIDENN_TILE_ARGS p, x_free, x_fixed, u_free, u_fixed, consts, flags, gx, gu, scratch, stream, tile_begin, tile_end
            if (body && iso) rc = launch_tile<R, true, true>(HIDENN_TILE_ARGS);
            else if (body) rc = launch_tile<R, true, false>(HIDENN_TILE_ARGS);
            else if (iso) rc = launch_tile<R, false, true>(HIDENN_TILE_ARGS);
            else rc = launch_tile<R, false, false>(HIDENN_TILE_ARGS);
#undef HIDENN_TILE_ARGS
            if (rc) return rc;
        } else {
            const size_t smem = (size_t)p->dev.max_local * 4 * sizeof(R) + 64 * 8;
            static size_t configured[kMaxDevices][2] = {};
            size_t& cfg = configured[p->device % kMaxDevices][sizeof(R) == 8 ? 0 : 1];
            if (smem > cfg) {
                HIDENN_CUDA_OK(cudaFuncSetAttribute(tri_tile_energy_only_kernel<R>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
                cfg = smem;
            }
            tri_tile_energy_only_kernel<R><<<p->dev.n_tiles, kTileBlock, smem, stream>>>(
                p->dev, (const R2*)x_free, (const R2*)x_fixed, (const R2*)u_free, (const R2*)u_fixed, consts, scratch);
        }
        HIDENN_CUDA_OK(cudaGetLastError());
    }
    if (flags & HIDENN_TILES_ONLY) return 0;
    {
        double* part = scratch_part<R>(p, scratch);
        unsigned* ticket = reinterpret_cast<unsigned*>(part + 2 * kEdgeMaxCtas);
        const int work = std::max(std::max(p->dev.n_edges, p->dev.n_enodes), p->dev.n_tiles / 8);
        const int grid = std::max(1, std::min(kEdgeMaxCtas, (work + kEdgeBlock - 1) / kEdgeBlock));
        tri_edge_finalize_kernel<R><<<grid, kEdgeBlock, 0, stream>>>(p->dev, (const R2*)x_free, (const R2*)x_fixed, (const R2*)u_free,
                                                                      (const R2*)u_fixed, consts, t_table, flags, scratch, out,
                                                                      (R2*)gx, (R2*)gu, gt, part, ticket, (R2*)en_final);
    }
    HIDENN_CUDA_OK(cudaGetLastError());
    return 0;
}

// Host-buffer entry point.  Large meshes are pipelined over the caller's stream and two copy streams (see tri_plan.h): while the tile kernels of
// chunk c run, the Parameter rows of chunk c+1 travel host->device and the finished gradient rows of chunk c-1 travel
// device->host, so the two PCIe directions overlap instead of adding up.  Rows touched by the Neumann edge term are
// patched from a compact buffer the finalize kernel fills.  Results are bit-identical to the resident entry point.
template <typename R>
static int tri_energy_host(hidenn_tri_plan* p, const R* xf, const R* xb, const R* uf, const R* ub, const R* consts_h, int flags,
                           R* out_h, R* gx_h, R* gu_h, void* stream_v) {
    HIDENN_REQUIRE(p != nullptr && p->device >= 0, "tri_energy_host: needs a device plan");
    HIDENN_REQUIRE(xf && uf && consts_h && out_h, "tri_energy_host: NULL argument");
    cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_v);
    DeviceScope scope;
    HIDENN_CUDA_OK(scope.enter(p->device));
    const size_t nfx = 2 * (size_t)p->n_free_x, nbx = 2 * (size_t)p->n_fixed_x, nfu = 2 * (size_t)p->n_free_u, nbu = 2 * (size_t)p->n_fixed_u;
    const size_t nsc = scratch_energy_slots(p) + 8 + 280;
    const size_t nen = 4 * (size_t)p->dev.n_enodes;
    auto al = [](size_t n) { return (n + 31) / 32 * 32; };
    const size_t total = al(nfx) * 2 + al(nbx) + al(nfu) * 2 + al(nbu) + al(HIDENN_TRI_NCONST) + al(4) + al(nsc) + al(nen);
    if (plan_ensure_arena(p, total * sizeof(R))) return 1;
    R* base = reinterpret_cast<R*>(p->arena);
    R* d_xf = base; base += al(nfx);
    R* d_xb = base; base += al(nbx);
    R* d_uf = base; base += al(nfu);
    R* d_ub = base; base += al(nbu);
    R* d_gx = base; base += al(nfx);
    R* d_gu = base; base += al(nfu);
    R* d_c = base; base += al(HIDENN_TRI_NCONST);
    R* d_out = base; base += al(4);
    R* d_sc = base; base += al(nsc);
    R* d_en = base;
    const bool want_gx = (flags & HIDENN_NEED_GX) && gx_h, want_gu = (flags & HIDENN_NEED_GU) && gu_h;
    const int n_tiles = p->dev.n_tiles;
    int chunks = 1;
    if ((flags & (HIDENN_NEED_GX | HIDENN_NEED_GU)) && n_tiles >= 512 && (nfx + nfu) * sizeof(R) >= ((size_t)8 << 20)) chunks = std::min(8, n_tiles / 256);
    if (const char* e = getenv("HIDENN_HOST_CHUNKS")) chunks = std::max(1, std::min(atoi(e), std::max(1, n_tiles)));
    if (!(flags & (HIDENN_NEED_GX | HIDENN_NEED_GU))) chunks = 1;      // the energy-only kernel runs all tiles in one launch

    HIDENN_CUDA_OK(cudaMemsetAsync(d_sc, 0, nsc * sizeof(R), stream));      // the finalize ticket must start at zero
    if (chunks <= 1) {
        HIDENN_CUDA_OK(cudaMemcpyAsync(d_xf, xf, nfx * sizeof(R), cudaMemcpyHostToDevice, stream));
        if (nbx) HIDENN_CUDA_OK(cudaMemcpyAsync(d_xb, xb, nbx * sizeof(R), cudaMemcpyHostToDevice, stream));
        HIDENN_CUDA_OK(cudaMemcpyAsync(d_uf, uf, nfu * sizeof(R), cudaMemcpyHostToDevice, stream));
        if (nbu) HIDENN_CUDA_OK(cudaMemcpyAsync(d_ub, ub, nbu * sizeof(R), cudaMemcpyHostToDevice, stream));
        HIDENN_CUDA_OK(cudaMemcpyAsync(d_c, consts_h, HIDENN_TRI_NCONST * sizeof(R), cudaMemcpyHostToDevice, stream));
        if (tri_energy_launch<R>(p, d_xf, d_xb, d_uf, d_ub, d_c, nullptr, flags, d_out, d_gx, d_gu, nullptr, d_sc, stream_v)) return 1;
        HIDENN_CUDA_OK(cudaMemcpyAsync(out_h, d_out, 4 * sizeof(R), cudaMemcpyDeviceToHost, stream));
        if (want_gx) HIDENN_CUDA_OK(cudaMemcpyAsync(gx_h, d_gx, nfx * sizeof(R), cudaMemcpyDeviceToHost, stream));
        if (want_gu) HIDENN_CUDA_OK(cudaMemcpyAsync(gu_h, d_gu, nfu * sizeof(R), cudaMemcpyDeviceToHost, stream));
        HIDENN_CUDA_OK(cudaStreamSynchronize(stream));
        return 0;
    }

    // side streams: one per PCIe direction (two per direction measured slower: same-direction transfers only
    // interleave, profiles/README.md)
    for (void*& st : p->pipe_streams)
        if (!st) {
            cudaStream_t a;
            HIDENN_CUDA_OK(cudaStreamCreateWithFlags(&a, cudaStreamNonBlocking));
            st = a;
        }
    while (p->pipe_events.size() < (size_t)(3 * chunks + 1)) {
        cudaEvent_t e;
        HIDENN_CUDA_OK(cudaEventCreateWithFlags(&e, cudaEventDisableTiming));
        p->pipe_events.push_back(e);
    }
    cudaStream_t s_inx = (cudaStream_t)p->pipe_streams[0], s_inu = s_inx;
    cudaStream_t s_outx = (cudaStream_t)p->pipe_streams[1], s_outu = s_outx;
    auto ev = [&](int i) { return (cudaEvent_t)p->pipe_events[i]; };
    // the side streams start after whatever the caller queued on `stream` (and after the scratch memset)
    HIDENN_CUDA_OK(cudaEventRecord(ev(3 * chunks), stream));
    for (void* st : p->pipe_streams) HIDENN_CUDA_OK(cudaStreamWaitEvent((cudaStream_t)st, ev(3 * chunks), 0));
    HIDENN_CUDA_OK(cudaMemcpyAsync(d_c, consts_h, HIDENN_TRI_NCONST * sizeof(R), cudaMemcpyHostToDevice, s_inx));
    if (nbx) HIDENN_CUDA_OK(cudaMemcpyAsync(d_xb, xb, nbx * sizeof(R), cudaMemcpyHostToDevice, s_inx));
    if (nbu) HIDENN_CUDA_OK(cudaMemcpyAsync(d_ub, ub, nbu * sizeof(R), cudaMemcpyHostToDevice, s_inu));
    // block -> chunk assignment: a block travels in with the chunk of the first tile that reads it and out with the
    // chunk of the last tile that writes it (blocks no tile touches ride with the last chunk)
    std::vector<int> tbound(chunks + 1);
    for (int c = 0; c <= chunks; ++c) tbound[c] = (int)((int64_t)n_tiles * c / chunks);
    auto chunk_of = [&](int32_t tile) {
        if (tile < 0 || tile >= n_tiles) return chunks - 1;
        int c = (int)std::min<int64_t>(chunks - 1, (int64_t)tile * chunks / n_tiles);
        while (tile < tbound[c]) --c;
        while (tile >= tbound[c + 1]) ++c;
        return c;
    };
    // copy the blocks whose key tile lies in chunk c, merging neighbouring blocks into one transfer
    auto copy_blocks = [&](const std::vector<int32_t>& key, int c, int32_t rows_per_block, size_t n_rows, R* dst, const R* src,
                           cudaMemcpyKind kind, cudaStream_t s) -> int {
        int b = 0;
        while (b < kPipeBlocks) {
            if ((size_t)b * rows_per_block >= n_rows) break;
            if (chunk_of(key[b]) != c) { ++b; continue; }
            int e = b + 1;
            while (e < kPipeBlocks && (size_t)e * rows_per_block < n_rows && chunk_of(key[e]) == c) ++e;
            const size_t r0 = (size_t)b * rows_per_block, r1 = std::min(n_rows, (size_t)e * rows_per_block);
            HIDENN_CUDA_OK(cudaMemcpyAsync(dst + 2 * r0, src + 2 * r0, 2 * (r1 - r0) * sizeof(R), kind, s));
            b = e;
        }
        return 0;
    };
    for (int c = 0; c < chunks; ++c) {
        const int t0 = tbound[c], t1 = tbound[c + 1];
        if (copy_blocks(p->first_need_x, c, p->pipe_rows_x, (size_t)p->n_free_x, d_xf, xf, cudaMemcpyHostToDevice, s_inx)) return 1;
        if (copy_blocks(p->first_need_u, c, p->pipe_rows_u, (size_t)p->n_free_u, d_uf, uf, cudaMemcpyHostToDevice, s_inu)) return 1;
        HIDENN_CUDA_OK(cudaEventRecord(ev(3 * c), s_inx));
        HIDENN_CUDA_OK(cudaEventRecord(ev(3 * c + 1), s_inu));
        HIDENN_CUDA_OK(cudaStreamWaitEvent(stream, ev(3 * c), 0));
        HIDENN_CUDA_OK(cudaStreamWaitEvent(stream, ev(3 * c + 1), 0));
        if (t1 > t0 &&
            tri_energy_launch<R>(p, d_xf, d_xb, d_uf, d_ub, d_c, nullptr, flags | HIDENN_TILES_ONLY, d_out, d_gx, d_gu, nullptr, d_sc,
                                 stream_v, t0, t1))
            return 1;
        HIDENN_CUDA_OK(cudaEventRecord(ev(3 * c + 2), stream));
        if (want_gx) {
            HIDENN_CUDA_OK(cudaStreamWaitEvent(s_outx, ev(3 * c + 2), 0));
            if (copy_blocks(p->last_own_x, c, p->pipe_rows_x, (size_t)p->n_free_x, gx_h, d_gx, cudaMemcpyDeviceToHost, s_outx)) return 1;
        }
        if (want_gu) {
            HIDENN_CUDA_OK(cudaStreamWaitEvent(s_outu, ev(3 * c + 2), 0));
            if (copy_blocks(p->last_own_u, c, p->pipe_rows_u, (size_t)p->n_free_u, gu_h, d_gu, cudaMemcpyDeviceToHost, s_outu)) return 1;
        }
    }
    // edge term + reduction; the edge-node rows it updates are fetched compactly and patched in below
    if (tri_energy_launch<R>(p, d_xf, d_xb, d_uf, d_ub, d_c, nullptr, flags | kFinalizeOnly, d_out, d_gx, d_gu, nullptr, d_sc, stream_v, 0, -1,
                             d_en))
        return 1;
    // rows the separate edge kernel updates after the tiles are patched in from a compact buffer; tile-ordered plans fold
    // the edge term inside the tiles, so their rows are already final
    const bool edges = !p->tile_order && (flags & HIDENN_WITH_EDGES) && p->dev.n_edges > 0 && (want_gx || want_gu);
    std::vector<R> en_h(edges ? nen : 0);
    HIDENN_CUDA_OK(cudaMemcpyAsync(out_h, d_out, 4 * sizeof(R), cudaMemcpyDeviceToHost, stream));
    if (edges) HIDENN_CUDA_OK(cudaMemcpyAsync(en_h.data(), d_en, nen * sizeof(R), cudaMemcpyDeviceToHost, stream));
    HIDENN_CUDA_OK(cudaStreamSynchronize(stream));
    HIDENN_CUDA_OK(cudaStreamSynchronize(s_outx));
    HIDENN_CUDA_OK(cudaStreamSynchronize(s_outu));
    if (edges) {
        const size_t ne = (size_t)p->dev.n_enodes;
        for (size_t k = 0; k < ne; ++k) {
            const int32_t us = p->en_uslot_h[k], xs = p->en_xslot_h[k];
            if (want_gu && us >= 0) { gu_h[2 * (size_t)us] = en_h[2 * k]; gu_h[2 * (size_t)us + 1] = en_h[2 * k + 1]; }
            if (want_gx && xs >= 0) { gx_h[2 * (size_t)xs] = en_h[2 * (ne + k)]; gx_h[2 * (size_t)xs + 1] = en_h[2 * (ne + k) + 1]; }
        }
    }
    return 0;
}

template <typename R> static int scale_launch(R* g, int64_t n, const R* s, void* stream_v) {
    HIDENN_REQUIRE(g && s, "scale_inplace: NULL");
    if (n <= 0) return 0;
    const int block = 256;
    const int grid = (int)std::min<int64_t>((n + block * 4 - 1) / (block * 4), 148 * 16);
    scale_inplace_kernel<R><<<grid, block, 0, reinterpret_cast<cudaStream_t>(stream_v)>>>(g, n, s);
    HIDENN_CUDA_OK(cudaGetLastError());
    return 0;
}

}  // namespace hidenn

using namespace hidenn;

extern "C" int hidenn_debug_tile_timing(long long* dev_buf) {
    hidenn::g_tile_timing = dev_buf;
    return 0;
}

extern "C" int hidenn_tri_energy_f64(const hidenn_tri_plan* plan, const double* x_free, const double* x_fixed,
                                     const double* u_free, const double* u_fixed, const double* consts,
                                     const double* t_table, int flags, double* out, double* gx_free, double* gu_free,
                                     double* gt_out, double* scratch, void* stream) {
    return tri_energy_launch<double>(plan, x_free, x_fixed, u_free, u_fixed, consts, t_table, flags, out, gx_free, gu_free,
                                     gt_out, scratch, stream);
}
extern "C" int hidenn_tri_energy_f32(const hidenn_tri_plan* plan, const float* x_free, const float* x_fixed,
                                     const float* u_free, const float* u_fixed, const float* consts, const float* t_table,
                                     int flags, float* out, float* gx_free, float* gu_free, float* gt_out, float* scratch,
                                     void* stream) {
    return tri_energy_launch<float>(plan, x_free, x_fixed, u_free, u_fixed, consts, t_table, flags, out, gx_free, gu_free,
                                    gt_out, scratch, stream);
}
extern "C" int hidenn_tri_energy_range_f64(const hidenn_tri_plan* plan, const double* x_free, const double* x_fixed, const double* u_free,
                                           const double* u_fixed, const double* consts, const double* t_table, int flags,
                                           double* gx_free, double* gu_free, double* gt_out, double* scratch, int tile_begin,
                                           int tile_end, void* stream) {
    HIDENN_REQUIRE(plan && plan->tile_order, "tri_energy_range: needs a tile-ordered FP64 plan");
    HIDENN_REQUIRE(flags & (HIDENN_NEED_GX | HIDENN_NEED_GU), "tri_energy_range: at least one gradient must be requested");
    HIDENN_REQUIRE(0 <= tile_begin && tile_begin <= tile_end && tile_end <= plan->dev.n_tiles, "tri_energy_range: bad tile range");
    double dummy_out[1];
    return tri_energy_launch<double>(plan, x_free, x_fixed, u_free, u_fixed, consts, t_table, flags | HIDENN_TILES_ONLY, dummy_out, gx_free,
                                     gu_free, gt_out, scratch, stream, tile_begin, tile_end);
}
extern "C" int hidenn_tri_energy_overlap_f64(const hidenn_tri_plan* plan, const double* x_free, const double* x_fixed, const double* u_free,
                                             const double* u_fixed, const double* consts, const double* t_table, int flags, double* out,
                                             double* gx_free, double* gu_free, double* gt_out, double* scratch, uint32_t* first_done,
                                             int reserve_sms, void* const* peer_bufs, void* my_buf, int me, int world, int64_t smax,
                                             uint64_t* loss_step, void* stream) {
    HIDENN_REQUIRE(plan && plan->tile_order && first_done, "tri_energy_overlap: needs a tile-ordered FP64 plan and a counter");
    HIDENN_REQUIRE(flags & (HIDENN_NEED_GX | HIDENN_NEED_GU), "tri_energy_overlap: at least one gradient must be requested");
    HIDENN_REQUIRE(!(flags & HIDENN_TILES_ONLY), "tri_energy_overlap: HIDENN_TILES_ONLY makes no sense here");
    HIDENN_REQUIRE(peer_bufs == nullptr || (my_buf && loss_step && world >= 1 && world <= 64), "tri_energy_overlap: bad peer-memory arguments");
    const P2PLossArgs A{(unsigned char* const*)peer_bufs, (unsigned char*)my_buf, (unsigned long long*)loss_step, (long long)smax, me, world};
    return tri_energy_launch<double>(plan, x_free, x_fixed, u_free, u_fixed, consts, t_table, flags, out, gx_free, gu_free, gt_out, scratch,
                                     stream, 0, -1, nullptr, first_done, reserve_sms, peer_bufs ? &A : nullptr);
}
// which gradient tile kernel hidenn_tri_energy_* launches for this plan: 7 = tri_tile_persistent_kernel (any numbering,
// FP64 / FP32), 8 = tri_tile8_kernel, 9 = tri_tile9_kernel (tile-ordered FP64 plans)
extern "C" int hidenn_tri_plan_kernel(const hidenn_tri_plan* plan) {
    static const bool ws_off = [] { const char* e = getenv("HIDENN_TILE_WS"); return e && atoi(e) == 0; }();
    if (!plan) return 0;
    if (!plan->tile_order) return 7;
    return ((!ws_off && tile9_fits(plan)) || !plan->unpaired_ok) ? 9 : 8;      // a pairs-only plan runs kernel v9 or nothing
}
extern "C" int hidenn_tri_plan_overlap_target(const hidenn_tri_plan* plan) {
    static const bool ws_off = [] { const char* e = getenv("HIDENN_TILE_WS"); return e && atoi(e) == 0; }();
    if (!plan || !plan->tile_order || plan->n_first_tiles <= 0 || ws_off || !tile9_fits(plan)) return 0;
    return plan->n_first_tiles * tile9_fold_warps(plan);
}
extern "C" int hidenn_tri_energy_finish_f64(const hidenn_tri_plan* plan, double* scratch, double* out, void* stream) {
    HIDENN_REQUIRE(plan && plan->tile_order && scratch && out, "tri_energy_finish: needs a tile-ordered FP64 plan, scratch and out");
    DeviceScope scope;
    HIDENN_CUDA_OK(scope.enter(plan->device));
    return tile8_reduce(plan, scratch, out, reinterpret_cast<cudaStream_t>(stream));
}
extern "C" int hidenn_tri_energy_host_f64(hidenn_tri_plan* plan, const double* a, const double* b, const double* c,
                                          const double* d, const double* k, int flags, double* out, double* gx, double* gu,
                                          void* stream) {
    return tri_energy_host<double>(plan, a, b, c, d, k, flags, out, gx, gu, stream);
}
extern "C" int hidenn_tri_energy_host_f32(hidenn_tri_plan* plan, const float* a, const float* b, const float* c,
                                          const float* d, const float* k, int flags, float* out, float* gx, float* gu,
                                          void* stream) {
    return tri_energy_host<float>(plan, a, b, c, d, k, flags, out, gx, gu, stream);
}
template <typename R> static int scale2_launch(R* g1, int64_t n1, R* g2, int64_t n2, const R* s, void* stream_v) {
    HIDENN_REQUIRE(s && (g1 || n1 == 0) && (g2 || n2 == 0), "scale_inplace2: NULL");
    const int64_t n = std::max(n1, n2);
    if (n <= 0) return 0;
    const int block = 256;
    const int grid = (int)std::min<int64_t>((n + block * 4 - 1) / (block * 4), 148 * 16);
    scale_inplace2_kernel<R><<<grid, block, 0, reinterpret_cast<cudaStream_t>(stream_v)>>>(g1, n1, g2, n2, s);
    HIDENN_CUDA_OK(cudaGetLastError());
    return 0;
}
extern "C" int hidenn_scale_inplace2_f64(double* g1, int64_t n1, double* g2, int64_t n2, const double* s, void* stream) {
    return scale2_launch<double>(g1, n1, g2, n2, s, stream);
}
extern "C" int hidenn_scale_inplace2_f32(float* g1, int64_t n1, float* g2, int64_t n2, const float* s, void* stream) {
    return scale2_launch<float>(g1, n1, g2, n2, s, stream);
}
extern "C" int hidenn_scale_inplace_f64(double* g, int64_t n, const double* s, void* stream) { return scale_launch<double>(g, n, s, stream); }
extern "C" int hidenn_scale_inplace_f32(float* g, int64_t n, const float* s, void* stream) { return scale_launch<float>(g, n, s, stream); }
