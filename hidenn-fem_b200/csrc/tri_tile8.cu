// Tile kernel v8 (FP64) for tile-ordered node numberings (tri_plan.h, hidenn_tri_locality_order).
//
// Same owner-computes fold as tri_energy.cu (one launch = EnergyLoss2D.__call__ + backward of the reference,
// /root/reference/src/loss.py:55-116 over /root/reference/src/models.py:292-376), with Blackwell tile movement:
//   * a tile's owned rows of node_coords_free / node_coords_fixed / u_free / u_fixed are contiguous runs, staged into
//     shared memory by ONE elected thread with cp.async.bulk (<= 5 copies per tile) completing on an mbarrier, one
//     tile ahead of the compute; only the halo rows (~20 %) are gathered per thread with cp.async;
//   * local id = row offset inside the run, so fold thread l holds the final gradient of row (first row + l):
//     consecutive lanes store consecutive rows -- no slot records, no output staging buffer, no flush pass;
//   * the Neumann edge term (src/loss.py:91-110, src/models.py:359-376) is folded by the tile that owns the edge node
//     (edge partials occupy the fold slots after the node's element slots), and the last CTA to finish (integer
//     ticket) adds the per-tile energies in fixed order: no finalize launch.
#include "../../include/hidenn_b200.h"
#include "common.cuh"
#include "tri_plan.h"
#include "tri_element.cuh"
#include "tri_tile8.h"

#include <algorithm>
#include <cstdlib>

namespace hidenn {

constexpr int kBlock8 = 256;

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t* bar, unsigned count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, unsigned bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
// global -> shared bulk copy (TMA unit, no LSU wavefronts); bytes and both addresses are multiples of 16
__device__ __forceinline__ void bulk_g2s(void* dst, const void* src, unsigned bytes, uint64_t* bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smem_u32(dst)),
                 "l"(src), "r"(bytes), "r"(smem_u32(bar))
                 : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, unsigned parity) {
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "WAIT_%=:\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
        "@p bra DONE_%=;\n"
        "bra WAIT_%=;\n"
        "DONE_%=:\n"
        "}\n" ::"r"(smem_u32(bar)),
        "r"(parity)
        : "memory");
}

// fixed-order reduction of the per-tile energies (domain | edge) by one CTA -> out = [loss, domain, edge, 0]
__device__ __forceinline__ void reduce_tile_energies(const double* __restrict__ e_dom, const double* __restrict__ e_edge, int n, double* out,
                                                     double* s_red /*[16]*/) {
    const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
    double d = 0.0, e = 0.0;
    for (int t = tid; t < n; t += kBlock8) { d += __ldcg(e_dom + t); e += __ldcg(e_edge + t); }
    d = warp_sum(d);
    e = warp_sum(e);
    __syncthreads();
    if (lane == 0) { s_red[wid] = d; s_red[8 + wid] = e; }
    __syncthreads();
    if (tid == 0) {
        double dd = 0.0, ee = 0.0;
#pragma unroll
        for (int w = 0; w < kBlock8 / 32; ++w) { dd += s_red[w]; ee += s_red[8 + w]; }
        out[0] = dd - ee;
        out[1] = dd;
        out[2] = ee;
        out[3] = 0.0;
    }
}

__global__ void __launch_bounds__(kBlock8) tri_reduce8_kernel(const double* e_dom, const double* e_edge, int n, double* out) {
    __shared__ double s_red[16];
    reduce_tile_energies(e_dom, e_edge, n, out, s_red);
}

template <bool BODY, bool ISO, int MINB>
__global__ void __launch_bounds__(kBlock8, MINB)
tri_tile8_kernel(const TriPlanDev P, const TriPlan8Dev P8, const double2* __restrict__ x_free, const double2* __restrict__ x_fixed,
                 const double2* __restrict__ u_free, const double2* __restrict__ u_fixed, const double* __restrict__ consts,
                 const double* __restrict__ t_table, const int flags, double2* __restrict__ gx_free, double2* __restrict__ gu_free,
                 double* __restrict__ gt_out, double* __restrict__ e_dom, double* __restrict__ e_edge, const double* e_dom_all,
                 const double* e_edge_all, const int n_tiles_total, double* __restrict__ out, unsigned* __restrict__ ticket) {
    using R = double;
    using R2 = double2;
    constexpr int BLOCK = kBlock8;
    extern __shared__ __align__(16) unsigned char smem_raw[];
    // shared layout: 2 x (xy | uv) node buffers, fold partial pairs gu | gx (+ dump slot), 2 x (8 domain + 8 edge) warp energy
    // partials, 2 mbarriers, last-CTA flag
    R2* s_node = reinterpret_cast<R2*>(smem_raw);
    const int nb = 2 * P.max_local;
    const PartBuf<R> part(s_node + 2 * nb, P.max_entries + 1);
    R* s_red = reinterpret_cast<R*>(s_node + 2 * nb + 2 * (P.max_entries + 1));      // [2][16]
    uint64_t* mbar = reinterpret_cast<uint64_t*>(s_red + 32);
    unsigned* s_flag = reinterpret_cast<unsigned*>(mbar + 2);
    const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
    const int nct = gridDim.x;
    int tile = blockIdx.x;
    constexpr unsigned LM = (1u << kLidBits) - 1u, PM = (1u << kPosBits) - 1u;
    constexpr unsigned G = 8u;
    constexpr int NPRE = 768 / BLOCK;
    constexpr int NW = BLOCK / 32;
    const bool need_gx = flags & HIDENN_NEED_GX, need_gu = flags & HIDENN_NEED_GU;

    if (tid == 0) {
        mbar_init(&mbar[0], 1);
        mbar_init(&mbar[1], 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();

    auto load_halo = [&](const int t) -> int2 {
        return tid < P8.stride_halo ? __ldg(P8.t_halo + (size_t)t * P8.stride_halo + tid) : make_int2(0, 0);
    };
    // put the node rows of tile t in flight into node buffer `buf`
    auto stage = [&](const int t, const TileDesc8& d, const int2 h, const int buf) {
        R2* xy = s_node + buf * nb;
        R2* uv = xy + P.max_local;
        if (tid == 0) {
            const int nBC = d.nB + d.nC, nAB = d.nA + d.nB, nCD = d.nC + d.nD;
            mbar_expect_tx(&mbar[buf], 32u * (unsigned)d.n_owned);
            if (d.nA) bulk_g2s(xy, x_free + d.rx_free, 16u * d.nA, &mbar[buf]);
            if (nBC) bulk_g2s(xy + d.nA, x_fixed + d.rx_fixed, 16u * nBC, &mbar[buf]);
            if (d.nD) bulk_g2s(xy + d.nA + nBC, x_free + d.rx_free + d.nA, 16u * d.nD, &mbar[buf]);
            if (nAB) bulk_g2s(uv, u_free + d.ru_free, 16u * nAB, &mbar[buf]);
            if (nCD) bulk_g2s(uv + nAB, u_fixed + d.ru_fixed, 16u * nCD, &mbar[buf]);
        }
        const int n_halo = d.n_local - d.n_owned;
        if (tid < n_halo) {      // halo rows: consecutive lanes land at consecutive local ids (conflict-free)
            cp_async_pair(xy + d.n_owned + tid, h.x >= 0 ? (const void*)(x_free + h.x) : (const void*)(x_fixed + (~h.x)), 16);
            cp_async_pair(uv + d.n_owned + tid, h.y >= 0 ? (const void*)(u_free + h.y) : (const void*)(u_fixed + (~h.y)), 16);
        }
        for (int j = tid + BLOCK; j < n_halo; j += BLOCK) {      // tiles with more than BLOCK halo nodes (strip meshes)
            const int2 h2 = __ldg(P8.t_halo + (size_t)t * P8.stride_halo + j);
            cp_async_pair(xy + d.n_owned + j, h2.x >= 0 ? (const void*)(x_free + h2.x) : (const void*)(x_fixed + (~h2.x)), 16);
            cp_async_pair(uv + d.n_owned + j, h2.y >= 0 ? (const void*)(u_free + h2.y) : (const void*)(u_fixed + (~h2.y)), 16);
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

    TileDesc8 td = P8.tiles[tile], td_n = td;
    int2 h_nxt = make_int2(0, 0);
    unsigned long long pk[NPRE];
    uint32_t off[2];
    stage(tile, td, load_halo(tile), 0);
    load_meta(tile, pk, off);
    if (tile + nct < P.n_tiles) { td_n = P8.tiles[tile + nct]; h_nxt = load_halo(tile + nct); }
    const TriConsts<R> K = load_consts<R, BODY>(consts);
    const bool with_edges = (flags & HIDENN_WITH_EDGES) != 0;
    cp_async_wait_all();
    __syncthreads();
    mbar_wait(&mbar[0], 0);

    int b = 0;
    unsigned ph = 1u;          // bit b = parity the NEXT wait on buffer b uses (buffer 0 has completed phase 0)
    for (;;) {
        const int tnext = tile + nct;
        const bool has_next = tnext < P.n_tiles;
        if (has_next) stage(tnext, td_n, h_nxt, b ^ 1);      // lands during E + F of this tile

        // E: elements -> energy + gradient partials at their fold slots
        const NodeBuf<R> nodes(s_node + b * nb, P.max_local);
        R e_acc = R(0), ee_acc = R(0);
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
        if (td.n_edge > 0) {
            // Neumann edges with an end owned by this tile (a handful of tiles): N = [1-xi, xi] on raw [-1,1] Gauss points
            const int ng1 = (int)consts[HIDENN_TRI_NG1];
            for (int i = tid; i < td.n_edge; i += BLOCK) {
                const unsigned long long w = __ldg(P8.edge_pack + td.edge_off + i);
                const int e = __ldg(P8.edge_id + td.edge_off + i);
                const unsigned lo = (unsigned)w;
                const unsigned l0 = lo & LM, l1 = (lo >> kLidBits) & LM;
                const unsigned p0 = (unsigned)(w >> (2 * kLidBits)) & PM, p1 = (unsigned)(w >> (2 * kLidBits + kPosBits)) & PM;
                const bool owner = (w >> kOwnerBit) & 1ull;
                R2 x0, x1, U0, U1;
                nodes.load(l0, x0, U0); nodes.load(l1, x1, U1);
                const R dx = x1.x - x0.x, dy = x1.y - x0.y;
                const R ds = sqrt(dx * dx + dy * dy);
                const R dirx = dx / ds, diry = dy / ds;
                R S = R(0), f0x = R(0), f0y = R(0), f1x = R(0), f1y = R(0);
                for (int q = 0; q < ng1; ++q) {
                    const R xi = consts[HIDENN_TRI_XI1 + q], wq = consts[HIDENN_TRI_W1 + q];
                    const R ux = (R(1) - xi) * U0.x + xi * U1.x, uy = (R(1) - xi) * U0.y + xi * U1.y;
                    R tx, ty;
                    if (t_table) { tx = t_table[((size_t)e * ng1 + q) * 2]; ty = t_table[((size_t)e * ng1 + q) * 2 + 1]; }
                    else { tx = consts[HIDENN_TRI_TX]; ty = consts[HIDENN_TRI_TY]; }
                    S += wq * (ux * tx + uy * ty);
                    f0x += wq * (R(1) - xi) * tx; f0y += wq * (R(1) - xi) * ty;
                    f1x += wq * xi * tx; f1y += wq * xi * ty;
                    if (owner && gt_out && with_edges) {      // d loss / d t_q = -w_q ds u_q
                        gt_out[((size_t)e * ng1 + q) * 2] = -wq * ds * ux;
                        gt_out[((size_t)e * ng1 + q) * 2 + 1] = -wq * ds * uy;
                    }
                }
                const R m = with_edges ? R(1) : R(0);      // the slots exist in the fold either way
                if (owner) ee_acc += m * S * ds;
                // d(-E_edge): dU_k = -ds sum_q w N_k t ;  dx0 = +S dir, dx1 = -S dir
                if (p0 != dumpv) part.store(p0, mk2<R>(-m * ds * f0x, -m * ds * f0y), mk2<R>(m * S * dirx, m * S * diry));
                if (p1 != dumpv) part.store(p1, mk2<R>(-m * ds * f1x, -m * ds * f1y), mk2<R>(-m * S * dirx, -m * S * diry));
            }
        }
        e_acc = warp_sum(e_acc);
        ee_acc = warp_sum(ee_acc);
        if (lane == 0) { s_red[b * 16 + wid] = e_acc; s_red[b * 16 + 8 + wid] = ee_acc; }
        __syncthreads();

        // F: fold.  First put the next tile's metadata loads in flight.
        unsigned long long pk_n[NPRE];
        uint32_t off_n[2];
        TileDesc8 td_n2 = td_n;
        int2 h_n2 = make_int2(0, 0);
        if (has_next) {
            load_meta(tnext, pk_n, off_n);
            if (tnext + nct < P.n_tiles) { td_n2 = P8.tiles[tnext + nct]; h_n2 = load_halo(tnext + nct); }
        }
        const int nA = td.nA, nAB = td.nA + td.nB, nABC = nAB + td.nC, nBC = td.nB + td.nC;
        // thread l folds owned node l (conflict-free slot rows) and stores its final gradient rows: lane -> consecutive rows
        auto fold_node = [&](const uint32_t oc, const int l) {
            const unsigned fb = oc & 0xFFFFu, fe = fb + (oc >> 16) * G;
            R ax = R(0), ay = R(0), bx = R(0), by = R(0);
            for (unsigned k = fb; k < fe; k += G) {
                R2 u, x;
                part.load(k, u, x);
                ax += u.x; ay += u.y; bx += x.x; by += x.y;
            }
            if (need_gu && l < nAB) gu_free[td.ru_free + l] = mk2<R>(ax, ay);
            if (need_gx && (l < nA || l >= nABC)) gx_free[td.rx_free + (l < nA ? l : l - nBC)] = mk2<R>(bx, by);
        };
#pragma unroll
        for (int k = 0; k < 2; ++k)
            if (tid + k * BLOCK < td.n_owned) fold_node(off[k], tid + k * BLOCK);
        {
            const uint32_t* __restrict__ offs = P.entry_off + (size_t)tile * P.stride_owned;
            for (int i = tid + 2 * BLOCK; i < td.n_owned; i += BLOCK) fold_node(__ldg(offs + i), i);
        }
        if (tid == 0) {        // tile energies, summed in fixed warp order
            const R* r = s_red + b * 16;
            R d = R(0), e = R(0);
#pragma unroll
            for (int w = 0; w < NW; ++w) { d += r[w]; e += r[8 + w]; }
            e_dom[tile] = d;
            e_edge[tile] = e;
        }
        if (!has_next) break;
#pragma unroll
        for (int k = 0; k < NPRE; ++k) pk[k] = pk_n[k];
        off[0] = off_n[0]; off[1] = off_n[1];
        td = td_n; td_n = td_n2; h_nxt = h_n2;
        tile = tnext;
        b ^= 1;
        cp_async_wait_all();
        __syncthreads();      // fold reads of the partials are done; this thread's halo rows have landed
        mbar_wait(&mbar[b], (ph >> b) & 1u);      // ... and the bulk copies of the owned rows
        ph ^= 1u << b;
    }

    if (flags & HIDENN_TILES_ONLY) return;
    // last CTA to finish adds the per-tile energies in fixed order (the result does not depend on which CTA is last)
    if (tid == 0) {
        __threadfence();
        *s_flag = (atomicAdd(ticket, 1u) == gridDim.x - 1) ? 1u : 0u;
    }
    __syncthreads();
    if (*s_flag) {
        __threadfence();
        reduce_tile_energies(e_dom_all, e_edge_all, n_tiles_total, out, s_red);
        if (tid == 0) *ticket = 0u;
    }
}

size_t tile8_smem_bytes(const hidenn_tri_plan* p) {
    return (size_t)p->dev.max_local * 64 + (size_t)(p->dev.max_entries + 1) * 32 + 32 * 8 + 16 + 16;
}

template <bool BODY, bool ISO, int MINB>
static int launch8(const hidenn_tri_plan* p, const double* x_free, const double* x_fixed, const double* u_free, const double* u_fixed,
                   const double* consts, const double* t_table, int flags, double* out, double* gx, double* gu, double* gt,
                   double* scratch, unsigned* ticket, cudaStream_t stream, int tile_begin, int tile_end) {
    const size_t smem = tile8_smem_bytes(p);
    static size_t configured[kMaxDevices] = {};
    size_t& cfg = configured[p->device % kMaxDevices];
    if (smem > cfg) {
        HIDENN_CUDA_OK(cudaFuncSetAttribute(tri_tile8_kernel<BODY, ISO, MINB>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        HIDENN_CUDA_OK(cudaFuncSetAttribute(tri_tile8_kernel<BODY, ISO, MINB>, cudaFuncAttributePreferredSharedMemoryCarveout, 100));
        cfg = smem;
    }
    TriPlanDev P = p->dev;
    TriPlan8Dev P8 = p->dev8;
    const int n_total = p->dev.n_tiles;
    P.elem_pack += (size_t)tile_begin * P.stride_elem;
    P.entry_off += (size_t)tile_begin * P.stride_owned;
    P.n_tiles = tile_end - tile_begin;
    P8.tiles += tile_begin;
    P8.t_halo += (size_t)tile_begin * P8.stride_halo;
    const int grid = std::min(P.n_tiles, sm_count(p->device) * MINB);
    tri_tile8_kernel<BODY, ISO, MINB><<<grid, kBlock8, smem, stream>>>(
        P, P8, (const double2*)x_free, (const double2*)x_fixed, (const double2*)u_free, (const double2*)u_fixed, consts, t_table, flags,
        (double2*)gx, (double2*)gu, gt, scratch + tile_begin, scratch + n_total + tile_begin, scratch, scratch + n_total, n_total, out, ticket);
    HIDENN_CUDA_OK(cudaGetLastError());
    return 0;
}

int tile8_launch(const hidenn_tri_plan* p, const double* x_free, const double* x_fixed, const double* u_free, const double* u_fixed,
                 const double* consts, const double* t_table, int flags, double* out, double* gx, double* gu, double* gt, double* scratch,
                 unsigned* ticket, cudaStream_t stream, int tile_begin, int tile_end) {
    const bool body = !(flags & HIDENN_HINT_NO_BODY_FORCE), iso = (flags & HIDENN_HINT_C_PLANE_STRESS) != 0;
    static const int env_mb = [] { const char* e = getenv("HIDENN_TILE_MINB"); return e ? atoi(e) : 0; }();
    const int mb = (env_mb >= 1 && env_mb <= 3) ? env_mb : ((227 * 1024) / (tile8_smem_bytes(p) + 1024) >= 3 ? 3 : 2);
#define HIDENN_L8(B_, I_, M_) \
    return launch8<B_, I_, M_>(p, x_free, x_fixed, u_free, u_fixed, consts, t_table, flags, out, gx, gu, gt, scratch, ticket, stream, tile_begin, tile_end)
    if (mb >= 3) {
        if (body && iso) HIDENN_L8(true, true, 3);
        if (body) HIDENN_L8(true, false, 3);
        if (iso) HIDENN_L8(false, true, 3);
        HIDENN_L8(false, false, 3);
    }
    if (body && iso) HIDENN_L8(true, true, 2);
    if (body) HIDENN_L8(true, false, 2);
    if (iso) HIDENN_L8(false, true, 2);
    HIDENN_L8(false, false, 2);
#undef HIDENN_L8
}

int tile8_reduce(const hidenn_tri_plan* p, double* scratch, double* out, cudaStream_t stream) {
    const int n = p->dev.n_tiles;
    tri_reduce8_kernel<<<1, kBlock8, 0, stream>>>(scratch, scratch + n, n, out);
    HIDENN_CUDA_OK(cudaGetLastError());
    return 0;
}

}  // namespace hidenn
