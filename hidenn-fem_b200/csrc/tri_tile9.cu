// Tile kernel v9 (FP64, tile-ordered numberings): warp-specialised, one persistent CTA per SM.
//
// Same math, fold order and results as tri_tile8.cu (EnergyLoss2D.__call__ + backward of the reference,
// /root/reference/src/loss.py:55-116 over /root/reference/src/models.py:292-376); what changes is who does what:
//   * 12 ELEMENT warps (3 warpgroups, raised to 120 registers with setmaxnreg) do nothing but gather -> element closed
//     form -> partial stores, tile after tile;
//   * 11 FOLD warps (40 registers; one thread per owned node of a 330-node tile) sum the fold slots of the PREVIOUS tile (second partial buffer) and store the final
//     gradient rows -- the latency-bound slot loops run in the issue slots the FP64 chains leave free instead of
//     stalling the element warps behind two block barriers per tile;
//   * 1 LOADER warp keeps a 3-deep ring of tile stages full: lane 0 issues the bulk copies (owned rows, element packs,
//     fold offsets, descriptor) on the stage's mbarrier, all lanes gather the halo rows with cp.async.
// Hand-overs are mbarriers (full / empty per stage, full / empty per partial buffer); there is no __syncthreads in the
// tile loop.  Register budget: the CTA is launched with 768 x 80 registers (its pool); setmaxnreg moves them to
// 384 x 120 + 384 x 40 = 61440 (HIDENN_WS_EWARPS=16: 512 x 96 + 256 x 48).
#include "../../include/hidenn_b200.h"
#include "common.cuh"
#include "tri_plan.h"
#include "tri_element.cuh"
#include "tri_tile8.h"
#include "halo_p2p.cuh"

#include <algorithm>
#include <cstdlib>

namespace hidenn {

#ifndef HIDENN_WS_FOLD
#define HIDENN_WS_FOLD 1      // slots in flight per fold step.  1 (u and x pair of one slot) is fastest for the WHOLE kernel:
#endif                        // 2: +2 %, two nodes per thread side by side: +24 %, one load at a time: +10 % (profiles/README.md)
#ifndef HIDENN_WS_EWARPS
#define HIDENN_WS_EWARPS 12
#endif
// halo rows as 16-byte bulk copies (the copy engine writes shared memory without passing the SM's LSU data pipe, which the
// element warps need) instead of per-lane cp.async
#ifndef HIDENN_WS_HALO_BULK
#define HIDENN_WS_HALO_BULK 0
#endif
// Measurement-only ablations (profiles/v9_ablate.py; results are WRONG on purpose, never shipped): fraction of the
// shared-memory exchange kept per element -- 1: two of three corners (what edge-sharing pairs would move), 2: one of
// three (triangle strips), 3: none (FP64 + issue + global traffic floor).  The FP64 work is unchanged.
#ifndef HIDENN_ABL
#define HIDENN_ABL 0
#endif
// HIDENN_PROF9 (measurement-only): per-warp clock64 totals of the mbarrier waits, written over the tile-energy scratch
// as doubles [cta][warp][4] = total, wait A, wait B, tiles (A/B: element = stage full / partial buffer free, fold = stage
// full / partials full, loader = stage empty / -)
#ifndef HIDENN_PROF9
#define HIDENN_PROF9 0
#endif
#if HIDENN_PROF9
#define PROF_DECL long long pf_t0 = clock64(), pf_a = 0, pf_b = 0
#define PROF_WAIT(acc, stmt) { const long long pf_s = clock64(); stmt; acc += clock64() - pf_s; }
#define PROF_END(ntiles) if (lane == 0) { double* o = e_dom + ((size_t)blockIdx.x * kWarps9 + wid) * 4; o[0] = (double)(clock64() - pf_t0); o[1] = (double)pf_a; o[2] = (double)pf_b; o[3] = (double)(ntiles); }
#else
#define PROF_DECL
#define PROF_WAIT(acc, stmt) stmt;
#define PROF_END(ntiles)
#endif
// Loader warps: the per-tile chain  descriptor -> halo records -> gathers  is 2-3 dependent global loads (~3000 cycles);
// with ONE loader warp that chain is the critical path of the whole CTA (HIDENN_PROF9: the loader never waits, the
// element warps wait 19 % of the time for a full stage), so the tiles are dealt round-robin to kLWarps loader warps.
#ifndef HIDENN_WS_LWARPS
#define HIDENN_WS_LWARPS 2
#endif
constexpr int kLWarps = HIDENN_WS_LWARPS;
// CTA size and register split (setmaxnreg works on warpgroups of 4 warps; the launch pool is threads x the launch
// register count, and the split may not exceed it):
//   24 warps (pool 768 x 80):  12 element warps x 120 + 12 x 40,  or 16 x 96 + 8 x 48
//   28 warps (pool 896 x 72):  16 x 96 + 12 x 40
//   32 warps (pool 1024 x 64): 16 x 80 + 16 x 40,  or 12 x 104 + 20 x 40
#ifndef HIDENN_WS_WARPS
#define HIDENN_WS_WARPS 24
#endif
constexpr int kWarps9 = HIDENN_WS_WARPS, kThreads9 = kWarps9 * 32;
constexpr int kPoolRegs = kWarps9 == 24 ? 80 : kWarps9 == 28 ? 72 : 64;
static_assert(kWarps9 == 24 || kWarps9 == 28 || kWarps9 == 32, "768, 896 or 1024 threads");
// Roles per layout.  One element per entry: 12 element warps (two passes over a tile's ~720 visits), 10 fold warps (one
// owned node per thread).  Pairs: an entry is two elements and needs ~95 registers, a tile has ~410 entries (classes padded
// to whole warps), so 16 element warps take it in one pass; the fold is a third shorter and runs on 6 warps (two passes).
#ifndef HIDENN_WS_EWARPS_PAIRS
#define HIDENN_WS_EWARPS_PAIRS 16
#endif
template <bool PAIRS> struct Roles9 {
    static constexpr int E = PAIRS ? HIDENN_WS_EWARPS_PAIRS : HIDENN_WS_EWARPS;
    static constexpr int F = kWarps9 - kLWarps - E;
    static constexpr int ERegs = kWarps9 == 24 ? (E == 16 ? 96 : 120) : kWarps9 == 28 ? 96 : (E == 16 ? 80 : 104);
    static constexpr int ORegs = kWarps9 == 24 ? (E == 16 ? 48 : 40) : 40;
    static constexpr int EnSlots = E * 32 + 16;
    static_assert(E % 4 == 0 && E <= 16 && F >= 1, "element warps come in warpgroups");
    static_assert(E * ERegs + (kWarps9 - E) * ORegs <= kWarps9 * kPoolRegs, "register split exceeds the launch pool");
};
#ifndef HIDENN_WS_SLEEP_SHORT
#define HIDENN_WS_SLEEP_SHORT 32
#endif
#ifndef HIDENN_WS_SLEEP_IDLE
#define HIDENN_WS_SLEEP_IDLE 256
#endif
constexpr unsigned kSleepShort = HIDENN_WS_SLEEP_SHORT, kSleepIdle = HIDENN_WS_SLEEP_IDLE;      // ns, mbarrier wait back-off
constexpr int kRedWarp0 = 16;      // warps 16..23 (small-register groups in every configuration) do the final reduction
constexpr int kMaxStages = 4;

namespace {
__device__ __forceinline__ uint32_t s32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void bar_init(uint64_t* bar, unsigned count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(s32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void bar_expect_tx(uint64_t* bar, unsigned bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(s32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void bar_arrive(uint64_t* bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(s32(bar)) : "memory");
}
// arrives on the barrier when all cp.async of this thread issued so far have landed (the barrier's count includes it)
__device__ __forceinline__ void bar_arrive_cp_async(uint64_t* bar) {
    asm volatile("cp.async.mbarrier.arrive.noinc.shared::cta.b64 [%0];" ::"r"(s32(bar)) : "memory");
}
__device__ __forceinline__ void bulk(void* dst, const void* src, unsigned bytes, uint64_t* bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(s32(dst)), "l"(src),
                 "r"(bytes), "r"(s32(bar))
                 : "memory");
}
// shared-memory load the compiler may not sink below later volatile loads: the fold issues the loads of two slots
// back to back, then adds in slot order
__device__ __forceinline__ double2 lds_pair(const double2* p) {
    double2 v;
    asm volatile("ld.shared.v2.f64 {%0, %1}, [%2];" : "=d"(v.x), "=d"(v.y) : "r"(s32(p)));
    return v;
}
// Spinning warps take issue slots from the working warps of their scheduler (the ncu source page of the first paired
// kernel: 22 M of 126 M warp instructions were try_wait retries), so a failed try backs off with nanosleep: a few tens of
// ns where the waiter is on the critical path (fold warps waiting for partials), longer where it is not (element warps
// without entries in this tile, loader warps waiting for an empty stage).
__device__ __forceinline__ void bar_wait(uint64_t* bar, unsigned parity, unsigned sleep_ns) {
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "WAIT_%=:\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
        "@p bra DONE_%=;\n"
        "nanosleep.u32 %2;\n"
        "bra WAIT_%=;\n"
        "DONE_%=:\n"
        "}\n" ::"r"(s32(bar)),
        "r"(parity), "r"(sleep_ns)
        : "memory");
}
}  // namespace

struct Smem9 {          // offsets (bytes) into the dynamic shared memory
    int stage_bytes, node_off, pack_off, offs_off, desc_off;      // inside a stage
    int part_bytes, part0, en0, bar0, total;
};
__host__ __device__ inline Smem9 smem9_layout(int max_local, int max_entries9, int pack_bytes, int stride_owned, int n_stages, int e_warps) {
    Smem9 L;
    L.node_off = 0;
    L.pack_off = max_local * 32;
    L.offs_off = L.pack_off + pack_bytes;
    L.desc_off = L.offs_off + ((stride_owned * 4 + 15) & ~15);
    L.stage_bytes = L.desc_off + 64;
    L.part_bytes = (max_entries9 + 1) * 32;
    L.part0 = n_stages * L.stage_bytes;
    L.en0 = L.part0 + 2 * L.part_bytes;
    L.bar0 = L.en0 + 2 * (e_warps * 32 + 16) * 8;      // per partial buffer: one energy slot per element lane + 16 edge-energy warp sums
    L.total = L.bar0 + 16 * 8 + 16;
    return L;
}

// PAIRS: paired layout (tri_plan.h; opt-in, HIDENN_PLAN_PAIRS=1) instead of one element per entry
// JT: correct-math switch jinv_transpose (compile-time here: the selects would sit at both ends of the FP64 chain)
template <bool BODY, bool ISO, bool PAIRS, bool JT>
__global__ void __launch_bounds__(kThreads9, 1)
tri_tile9_kernel(const TriPlanDev P, const TriPlan8Dev P8, const double2* __restrict__ x_free, const double2* __restrict__ x_fixed,
                 const double2* __restrict__ u_free, const double2* __restrict__ u_fixed, const double* __restrict__ consts,
                 const double* __restrict__ t_table, const int flags, double2* __restrict__ gx_free, double2* __restrict__ gu_free,
                 double* __restrict__ gt_out, double* __restrict__ e_dom, double* __restrict__ e_edge, const double* e_dom_all,
                 const double* e_edge_all, const int n_tiles_total, double* __restrict__ out, unsigned* __restrict__ ticket,
                 const int kStages, unsigned* __restrict__ first_done, const int n_first, const P2PLossArgs loss_args) {
    using R = double;
    using R2 = double2;
    extern __shared__ __align__(128) unsigned char smem[];
    const int max_entries = PAIRS ? P8.max_entries9 : P.max_entries;
    constexpr int kEWarps = Roles9<PAIRS>::E, kFWarps = Roles9<PAIRS>::F, kERegs = Roles9<PAIRS>::ERegs, kORegs = Roles9<PAIRS>::ORegs,
                  kEnSlots = Roles9<PAIRS>::EnSlots;
    const Smem9 L = smem9_layout(P.max_local, max_entries, PAIRS ? P8.stride_pent * 16 : P.stride_elem * 8, P.stride_owned, kStages, kEWarps);
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem + L.bar0);
    uint64_t* full_stage = bars;                  // [kStages] loader -> element / fold warps
    uint64_t* empty_stage = bars + kMaxStages;    // [kStages] fold warps -> loader
    uint64_t* part_full = bars + 2 * kMaxStages;  // [2] element warps -> fold warps
    uint64_t* part_empty = part_full + 2;      // [2] fold warps -> element warps
    unsigned* s_flag = reinterpret_cast<unsigned*>(bars + 16);
    const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
    const int nct = gridDim.x;
    const int n_mine = (P.n_tiles - (int)blockIdx.x + nct - 1) / nct;      // tiles blockIdx.x, +nct, ...
    constexpr unsigned LM = (1u << kLidBits) - 1u, PM = (1u << kPosBits) - 1u;

    if (tid == 0) {
        for (int s = 0; s < kStages; ++s) { bar_init(&full_stage[s], HIDENN_WS_HALO_BULK ? 1 : 1 + 32); bar_init(&empty_stage[s], kFWarps); }
        for (int s = 0; s < 2; ++s) { bar_init(&part_full[s], kEWarps); bar_init(&part_empty[s], kFWarps); }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();

    if (wid < kEWarps) {
        // ------------------------------------------------------------------ element warps
        asm volatile("setmaxnreg.inc.sync.aligned.u32 %0;" ::"n"(kERegs));
        const TriConsts<R> K = load_consts<R, BODY>(consts);
        const bool with_edges = (flags & HIDENN_WITH_EDGES) != 0;
        const int etid = tid;
        PROF_DECL;
        int st = 0;
        unsigned st_ph = 0;      // stage index / phase parity of tile k (k % kStages, (k / kStages) & 1 without the divisions)
        for (int k = 0; k < n_mine; ++k) {
            const int pb = k & 1;
            unsigned char* stage = smem + st * L.stage_bytes;
            const TileDesc8* d = reinterpret_cast<const TileDesc8*>(stage + L.desc_off);
            const ulonglong2* s_pack = reinterpret_cast<const ulonglong2*>(stage + L.pack_off);
            const unsigned long long* s_pack1 = reinterpret_cast<const unsigned long long*>(stage + L.pack_off);
            const NodeBuf<R> nodes(stage + L.node_off, P.max_local);
            const PartBuf<R> part(smem + L.part0 + pb * L.part_bytes, max_entries + 1);
            PROF_WAIT(pf_a, bar_wait(&full_stage[st], st_ph, kSleepShort))
            const int n_pent = PAIRS ? d->n_pent : d->n_elem, n_edge = d->n_edge;
            // the pack of the next pass is loaded one pass ahead (the first one before the wait for the partial buffer)
            unsigned long long pw_next = (!PAIRS && etid < n_pent) ? s_pack1[etid] : 0ull;
            PROF_WAIT(pf_b, bar_wait(&part_empty[pb], ((k >> 1) & 1) ^ 1, (wid * 32 < n_pent || n_edge > 0) ? kSleepShort : kSleepIdle))
            const unsigned dumpv = (unsigned)(PAIRS ? d->n_entries9 : d->n_entries);
            R e_acc = R(0), ee_acc = R(0);
            // one entry = an edge-sharing element pair (or a single element): both elements are evaluated by this thread
            // from FOUR gathered nodes; the partials of the two shared nodes are added in registers (first element +
            // second), so the pair costs 4 gathers, 4 partial stores and 4 fold reads instead of 6 / 6 / 6
            for (int i = etid; i < n_pent; i += kEWarps * 32) {
                R2 gu[3], gx[3];
                unsigned l0, l1, l2, p0, p1, p2;
                if (PAIRS) {
                    const ulonglong2 pw = s_pack[i];
                    const unsigned lo = (unsigned)pw.x, hi = (unsigned)(pw.x >> 32);
                    l0 = lo & LM; l1 = (lo >> kLidBits) & LM; l2 = (lo >> (2 * kLidBits)) & LM;
                    p0 = (unsigned)(pw.x >> (3 * kLidBits)) & PM; p1 = (hi >> (3 * kLidBits + kPosBits - 32)) & PM;
                    p2 = (hi >> (3 * kLidBits + 2 * kPosBits - 32)) & PM;
                    const unsigned w2 = (unsigned)pw.y;
                    const unsigned l3 = w2 & LM, p3 = (w2 >> kLidBits) & PM, cls = (w2 >> (kLidBits + kPosBits)) & 15u;
                    if (cls == (unsigned)kPairSkip) continue;      // padding up to the next class boundary
                    R2 Q0, Q1, Q2, Q3, V0, V1, V2, V3;
                    nodes.load(l0, Q0, V0); nodes.load(l1, Q1, V1); nodes.load(l2, Q2, V2);
                    if (cls != (unsigned)kPairSingle) nodes.load(l3, Q3, V3);
                    R e;
                    if (cls == (unsigned)kPairSingle) {
                        tri_element<R, BODY, ISO>(Q0, Q1, Q2, V0, V1, V2, K, e, gu, gx, JT);
                        e_acc += (hi >> 31) ? e : R(0);
                    } else {
                        R2 g3u, g3x;
                        R e2;
                        // second element of class 3 i + r: corner r = Q3, corner r+1 = Q(i+1), corner r+2 = Q(i).  Both
                        // elements are evaluated inside the case, so their two dependent FP64 chains interleave
#define HIDENN_PAIR_CASE(I_, R_)                                                                                              \
    case 3 * I_ + R_: {                                                                                                       \
        const R2 A = I_ == 0 ? Q1 : I_ == 1 ? Q2 : Q0, AU = I_ == 0 ? V1 : I_ == 1 ? V2 : V0;                                 \
        const R2 B = I_ == 0 ? Q0 : I_ == 1 ? Q1 : Q2, BU = I_ == 0 ? V0 : I_ == 1 ? V1 : V2;                                 \
        R2 hu[3], hx[3];                                                                                                      \
        tri_element<R, BODY, ISO>(Q0, Q1, Q2, V0, V1, V2, K, e, gu, gx, JT);                                                  \
        if (R_ == 0) tri_element<R, BODY, ISO>(Q3, A, B, V3, AU, BU, K, e2, hu, hx, JT);                                      \
        else if (R_ == 1) tri_element<R, BODY, ISO>(B, Q3, A, BU, V3, AU, K, e2, hu, hx, JT);                                 \
        else tri_element<R, BODY, ISO>(A, B, Q3, AU, BU, V3, K, e2, hu, hx, JT);                                              \
        constexpr int ia = (I_ + 1) % 3, ib = I_, ca = (R_ + 1) % 3, cb = (R_ + 2) % 3;                                       \
        gu[ia].x += hu[ca].x; gu[ia].y += hu[ca].y; gx[ia].x += hx[ca].x; gx[ia].y += hx[ca].y;                               \
        gu[ib].x += hu[cb].x; gu[ib].y += hu[cb].y; gx[ib].x += hx[cb].x; gx[ib].y += hx[cb].y;                               \
        g3u = hu[R_]; g3x = hx[R_];                                                                                           \
    } break;
                        switch (cls) {
                            HIDENN_PAIR_CASE(0, 0) HIDENN_PAIR_CASE(0, 1) HIDENN_PAIR_CASE(0, 2)
                            HIDENN_PAIR_CASE(1, 0) HIDENN_PAIR_CASE(1, 1) HIDENN_PAIR_CASE(1, 2)
                            HIDENN_PAIR_CASE(2, 0) HIDENN_PAIR_CASE(2, 1)
                            default: HIDENN_PAIR_CASE(2, 2)
                        }
#undef HIDENN_PAIR_CASE
                        e_acc += (hi >> 31) ? e : R(0);
                        e_acc += (pw.y >> 63) ? e2 : R(0);
                        if (p3 != dumpv) part.store(p3, g3u, g3x);
                    }
                } else {
                    const unsigned long long pw = pw_next;
                    pw_next = (i + kEWarps * 32 < n_pent) ? s_pack1[i + kEWarps * 32] : 0ull;
                    const unsigned lo = (unsigned)pw, hi = (unsigned)(pw >> 32);
                    l0 = lo & LM; l1 = (lo >> kLidBits) & LM; l2 = (lo >> (2 * kLidBits)) & LM;
                    p0 = (unsigned)(pw >> (3 * kLidBits)) & PM; p1 = (hi >> (3 * kLidBits + kPosBits - 32)) & PM;
                    p2 = (hi >> (3 * kLidBits + 2 * kPosBits - 32)) & PM;
                    R e;
                    R2 v0, v1, v2, U0, U1, U2;
#if HIDENN_ABL == 0
                    nodes.load(l0, v0, U0); nodes.load(l1, v1, U1); nodes.load(l2, v2, U2);
#elif HIDENN_ABL == 1
                    nodes.load(l0, v0, U0); nodes.load(l1, v1, U1);
                    v2 = mk2<R>(v0.x + 3e-4, v1.y - 1e-4); U2 = mk2<R>(U0.x * 0.5, U1.y * 1.5);
#elif HIDENN_ABL == 2
                    nodes.load(l0, v0, U0);
                    v1 = mk2<R>(v0.x + 1e-4 * (R)(l1 & 3u), v0.y + 3e-4); U1 = mk2<R>(U0.y, U0.x * 2.0);
                    v2 = mk2<R>(v0.x + 3e-4, v0.y - 1e-4 * (R)(l2 & 3u)); U2 = mk2<R>(U0.x * 0.5, U0.y * 1.5);
#else
                    v0 = mk2<R>(1e-3 * (R)l0, 2e-3 * (R)l1); U0 = mk2<R>(1e-5 * (R)l2, 2e-5 * (R)l0);
                    v1 = mk2<R>(v0.x + 1e-4 * (R)(l1 & 3u), v0.y + 3e-4); U1 = mk2<R>(U0.y, U0.x * 2.0);
                    v2 = mk2<R>(v0.x + 3e-4, v0.y - 1e-4 * (R)(l2 & 3u)); U2 = mk2<R>(U0.x * 0.5, U0.y * 1.5);
#endif
                    tri_element<R, BODY, ISO>(v0, v1, v2, U0, U1, U2, K, e, gu, gx, JT);
                    e_acc += (hi >> 31) ? e : R(0);
#if HIDENN_ABL >= 1      // keep every result live
                    gu[0].x += gu[2].x * 0.5; gu[0].y += gu[2].y * 0.5; gx[0].x += gx[2].x * 0.5; gx[0].y += gx[2].y * 0.5;
#endif
#if HIDENN_ABL >= 2
                    gu[0].x += gu[1].x * 0.25; gu[0].y += gu[1].y * 0.25; gx[0].x += gx[1].x * 0.25; gx[0].y += gx[1].y * 0.25;
#endif
#if HIDENN_ABL >= 3
                    e_acc += gu[0].x + gu[0].y + gx[0].x + gx[0].y;
#endif
                }
#if HIDENN_ABL <= 2
                if (p0 != dumpv) part.store(p0, gu[0], gx[0]);
#endif
#if HIDENN_ABL <= 1
                if (p1 != dumpv) part.store(p1, gu[1], gx[1]);
#endif
#if HIDENN_ABL == 0
                if (p2 != dumpv) part.store(p2, gu[2], gx[2]);
#endif
            }
            if (n_edge > 0) {
                // Neumann edges with an end owned by this tile (a handful of tiles): N = [1-xi, xi] on raw [-1,1] Gauss points
                const int ng1 = (int)consts[HIDENN_TRI_NG1];
                const int edge_off = d->edge_off;
                for (int i = etid; i < n_edge; i += kEWarps * 32) {
                    const unsigned long long w = __ldg((PAIRS ? P8.edge_pack9 : P8.edge_pack) + edge_off + i);
                    const int e = __ldg(P8.edge_id + edge_off + i);
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
                    if (p0 != dumpv) part.store(p0, mk2<R>(-m * ds * f0x, -m * ds * f0y), mk2<R>(m * S * dirx, m * S * diry));
                    if (p1 != dumpv) part.store(p1, mk2<R>(-m * ds * f1x, -m * ds * f1y), mk2<R>(-m * S * dirx, -m * S * diry));
                }
            }
            // tile energy: every lane leaves its sum in its own slot, the last fold warp adds the slots in fixed order (a
            // shuffle tree here cost each element warp ~400 cycles per tile); edge energies only where the tile has edges
            R* s_en = reinterpret_cast<R*>(smem + L.en0) + pb * kEnSlots;
            s_en[etid] = e_acc;
            if (n_edge > 0) {
                ee_acc = warp_sum(ee_acc);
                if (lane == 0) s_en[kEWarps * 32 + wid] = ee_acc;
            }
            __syncwarp();
            if (lane == 0) bar_arrive(&part_full[pb]);      // release: this warp's partials and energy are visible
            if (++st == kStages) { st = 0; st_ph ^= 1u; }
        }
        PROF_END(n_mine)
    } else {
        asm volatile("setmaxnreg.dec.sync.aligned.u32 %0;" ::"n"(kORegs));
        // which of the non-element warps load: the schedulers (warp id mod 4) that carry four busy element warps of a
        // paired tile (~14 of 16 warps have entries: schedulers 0 and 1) get a loader and one fold warp, the others two
        // fold warps
#ifndef HIDENN_WS_LPOS
#define HIDENN_WS_LPOS 4
#endif
        // (one element per entry: loaders on warps 12, 13 = schedulers 0, 1, and the fold warps that take the nodes with the
        // most slots on schedulers 2, 3, whose element warps run out of second-pass entries first: 172.6 -> 166.2 us)
#ifndef HIDENN_WS_LPOS_SINGLE
#define HIDENN_WS_LPOS_SINGLE 0
#endif
        constexpr int kLoaderPos = (PAIRS && kWarps9 - kEWarps == 8 && kLWarps == 2) ? HIDENN_WS_LPOS : HIDENN_WS_LPOS_SINGLE;
        const int oj = wid - kEWarps;
        const bool is_loader = oj >= kLoaderPos && oj < kLoaderPos + kLWarps;
        const int fwarp = oj < kLoaderPos ? oj : oj - kLWarps;      // fold warp index
        if (is_loader) {
            // -------------------------------------------------------------- loader warp
            // The per-tile chain  descriptor -> bulk copies, halo records -> gathers  is two dependent global loads; the
            // records of the tile kAhead iterations later are pulled into L2 now, so the chain costs L2 hits, not DRAM misses.
            constexpr int kAhead = 4;
            auto prefetch_tile = [&](const int t) {
                const char* rec = reinterpret_cast<const char*>(P8.t_halo + (size_t)t * P8.stride_halo);
                if (lane == 31) asm volatile("prefetch.global.L2 [%0];" ::"l"(P8.tiles + t));
                else if (lane * 128 < P8.stride_halo * 8) asm volatile("prefetch.global.L2 [%0];" ::"l"(rec + lane * 128));
            };
            const int lw = oj - kLoaderPos;
            for (int k = lw; k < kAhead * kLWarps && k < n_mine; k += kLWarps) prefetch_tile(blockIdx.x + k * nct);
            PROF_DECL;
            int st = lw % kStages;
            unsigned st_ph = (lw / kStages) & 1;
            for (int k = lw; k < n_mine; k += kLWarps) {
                const int tile = blockIdx.x + k * nct;
                unsigned char* stage = smem + st * L.stage_bytes;
                if (k + kAhead * kLWarps < n_mine) prefetch_tile(tile + kAhead * kLWarps * nct);
                const TileDesc8 d = P8.tiles[tile];
                const int n_halo = d.n_local - d.n_owned;
                const int2* __restrict__ hrec = P8.t_halo + (size_t)tile * P8.stride_halo;
                int2 h0 = make_int2(0, 0), h1 = h0, h2 = h0;              // the first 96 halo records, in flight during the wait
                if (lane < n_halo) h0 = __ldg(hrec + lane);
                if (lane + 32 < n_halo) h1 = __ldg(hrec + lane + 32);
                if (lane + 64 < n_halo) h2 = __ldg(hrec + lane + 64);
                PROF_WAIT(pf_a, bar_wait(&empty_stage[st], st_ph ^ 1u, kSleepIdle))
                R2* xy = reinterpret_cast<R2*>(stage + L.node_off);
                R2* uv = xy + P.max_local;
                if (lane == 0) {
                    const int nBC = d.nB + d.nC, nAB = d.nA + d.nB, nCD = d.nC + d.nD;
                    const unsigned pack_bytes = PAIRS ? 16u * (unsigned)d.n_pent : 8u * (unsigned)((d.n_elem + 1) & ~1);
                    const unsigned off_bytes = 4u * (unsigned)((d.n_owned + 3) & ~3);
                    bar_expect_tx(&full_stage[st], 32u * (unsigned)(HIDENN_WS_HALO_BULK ? d.n_local : d.n_owned) + pack_bytes + off_bytes + 64u);
                    bulk(stage + L.desc_off, P8.tiles + tile, 64u, &full_stage[st]);
                    const void* pack_src = PAIRS ? (const void*)(P8.pair_pack + (size_t)tile * P8.stride_pent * 2)
                                                 : (const void*)(P.elem_pack + (size_t)tile * P.stride_elem);
                    if (pack_bytes) bulk(stage + L.pack_off, pack_src, pack_bytes, &full_stage[st]);
                    bulk(stage + L.offs_off, (PAIRS ? P8.entry_off9 : P.entry_off) + (size_t)tile * P.stride_owned, off_bytes, &full_stage[st]);
                    if (d.nA) bulk(xy, x_free + d.rx_free, 16u * d.nA, &full_stage[st]);
                    if (nBC) bulk(xy + d.nA, x_fixed + d.rx_fixed, 16u * nBC, &full_stage[st]);
                    if (d.nD) bulk(xy + d.nA + nBC, x_free + d.rx_free + d.nA, 16u * d.nD, &full_stage[st]);
                    if (nAB) bulk(uv, u_free + d.ru_free, 16u * nAB, &full_stage[st]);
                    if (nCD) bulk(uv + nAB, u_fixed + d.ru_fixed, 16u * nCD, &full_stage[st]);
                }
#if HIDENN_WS_HALO_BULK
                __syncwarp();      // lane 0's expect_tx (with the halo bytes) is posted before any halo copy can complete
#endif
                auto gather = [&](const int j, const int2 h) {      // consecutive lanes land at consecutive local ids
#if HIDENN_WS_HALO_BULK
                    bulk(xy + d.n_owned + j, h.x >= 0 ? (const void*)(x_free + h.x) : (const void*)(x_fixed + (~h.x)), 16u, &full_stage[st]);
                    bulk(uv + d.n_owned + j, h.y >= 0 ? (const void*)(u_free + h.y) : (const void*)(u_fixed + (~h.y)), 16u, &full_stage[st]);
#else
                    cp_async_pair(xy + d.n_owned + j, h.x >= 0 ? (const void*)(x_free + h.x) : (const void*)(x_fixed + (~h.x)), 16);
                    cp_async_pair(uv + d.n_owned + j, h.y >= 0 ? (const void*)(u_free + h.y) : (const void*)(u_fixed + (~h.y)), 16);
#endif
                };
                if (lane < n_halo) gather(lane, h0);
                if (lane + 32 < n_halo) gather(lane + 32, h1);
                if (lane + 64 < n_halo) gather(lane + 64, h2);
                for (int j = lane + 96; j < n_halo; j += 32) gather(j, __ldg(hrec + j));
#if !HIDENN_WS_HALO_BULK
                bar_arrive_cp_async(&full_stage[st]);
#endif
                st += kLWarps;
                while (st >= kStages) { st -= kStages; st_ph ^= 1u; }
            }
            PROF_END(n_mine)
        } else {
            // -------------------------------------------------------------- fold warps
            const int ftid = fwarp * 32 + lane;
            const bool need_gx = flags & HIDENN_NEED_GX, need_gu = flags & HIDENN_NEED_GU;
            constexpr unsigned G = 8u;
            PROF_DECL;
            int st = 0;
            unsigned st_ph = 0;
            for (int k = 0; k < n_mine; ++k) {
                const int tile = blockIdx.x + k * nct;
                const int pb = k & 1;
                unsigned char* stage = smem + st * L.stage_bytes;
                const TileDesc8* d = reinterpret_cast<const TileDesc8*>(stage + L.desc_off);
                const uint32_t* s_off = reinterpret_cast<const uint32_t*>(stage + L.offs_off);
                const PartBuf<R> part(smem + L.part0 + pb * L.part_bytes, max_entries + 1);
                PROF_WAIT(pf_a, bar_wait(&full_stage[st], st_ph, kSleepShort))
                PROF_WAIT(pf_b, bar_wait(&part_full[pb], (k >> 1) & 1, kSleepShort))
                const int n_owned = d->n_owned, nA = d->nA, nAB = nA + d->nB, nBC = d->nB + d->nC, nABC = nAB + d->nC;
                const int rx = d->rx_free, ru = d->ru_free;
                // one node per thread and pass, ONE slot (its u and x pair) in flight per step.  More loads in flight make the
                // WHOLE kernel slower: the fold warps have slack, the element warps do not, and both queue at the SM's
                // load/store pipe (two slots +2 %, two nodes side by side +24 %; one load at a time is too slow, +10 %;
                // even the same loop written with a pass counter cost 4 % -- profiles/README.md)
                for (int l = ftid; l < n_owned; l += kFWarps * 32) {
                    const uint32_t oc = s_off[l];
#if HIDENN_ABL == 0
                    const unsigned fb = oc & 0xFFFFu, fe = fb + (oc >> 16) * G;
#else
                    const unsigned fb = oc & 0xFFFFu, fe = fb + (((oc >> 16) * (3u - HIDENN_ABL) + 1u) / 3u) * G;
#endif
                    R ax = R(0), ay = R(0), bx = R(0), by = R(0);
                    unsigned q = fb;
#if HIDENN_WS_FOLD == 2
#pragma unroll 1
                    for (; q + G < fe; q += 2 * G) {
                        const R2 u0 = lds_pair(part.pu + q), u1 = lds_pair(part.pu + q + G);
                        const R2 x0 = lds_pair(part.px + q), x1 = lds_pair(part.px + q + G);
                        ax += u0.x + u1.x; ay += u0.y + u1.y; bx += x0.x + x1.x; by += x0.y + x1.y;
                    }
#endif
#pragma unroll 1
                    for (; q < fe; q += G) {
#if HIDENN_WS_FOLD == 0
                        // ONE shared-memory load in flight: the second address is made to depend on the first value
                        const R2 u = lds_pair(part.pu + q);
                        const R2 x = lds_pair(part.px + q + (__double2hiint(u.x) == 0x7ff8dead ? 1 : 0));
#else
                        R2 u, x;
                        part.load(q, u, x);
#endif
                        ax += u.x; ay += u.y; bx += x.x; by += x.y;
                    }
                    if (need_gu && l < nAB) gu_free[ru + l] = mk2<R>(ax, ay);
                    if (need_gx && (l < nA || l >= nABC)) gx_free[rx + (l < nA ? l : l - nBC)] = mk2<R>(bx, by);
                }
                if (fwarp == kFWarps - 1 && !HIDENN_PROF9) {
                    // tile energies in fixed order: lane j adds the slots of lane j of the element warps 0, 1, ..., then a
                    // shuffle tree (this fold warp has the fewest second-pass nodes)
                    const R* s_en = reinterpret_cast<const R*>(smem + L.en0) + pb * kEnSlots;
                    R dd = R(0);
#pragma unroll
                    for (int w = 0; w < kEWarps; ++w) dd += s_en[w * 32 + lane];
                    dd = warp_sum(dd);
                    if (lane == 0) {
                        R ee = R(0);
                        if (d->n_edge > 0)
                            for (int w = 0; w < kEWarps; ++w) ee += s_en[kEWarps * 32 + w];
                        e_dom[tile] = dd;
                        e_edge[tile] = ee;
                    }
                }
                __syncwarp();
                if (lane == 0) {
                    bar_arrive(&part_empty[pb]);
                    bar_arrive(&empty_stage[st]);
                    // multi-GPU overlap: the first n_first tiles own the nodes shared with other ranks; a kernel on another
                    // stream waits until first_done == n_first * kFWarps and then puts their rows into the peers' memory
                    if (first_done != nullptr && tile < n_first) {
                        __threadfence();
                        atomicAdd(first_done, 1u);
                    }
                }
                if (++st == kStages) { st = 0; st_ph ^= 1u; }
            }
            PROF_END(n_mine)
        }
    }

    if (flags & HIDENN_TILES_ONLY) return;
    // last CTA to finish adds the per-tile energies in fixed order (the result does not depend on which CTA is last)
    __syncthreads();
    if (tid == 0) {
        __threadfence();
        *s_flag = (atomicAdd(ticket, 1u) == gridDim.x - 1) ? 1u : 0u;
    }
    __syncthreads();
    if (*s_flag && wid >= kRedWarp0 && wid < kRedWarp0 + 8) {        // 8 warps of the small-register groups do the reduction (256 threads)
        __threadfence();
        const int t0 = tid - kRedWarp0 * 32;
        double dsum = 0.0, esum = 0.0;
        for (int t = t0; t < n_tiles_total; t += 256) { dsum += __ldcg(e_dom_all + t); esum += __ldcg(e_edge_all + t); }
        dsum = warp_sum(dsum);
        esum = warp_sum(esum);
        double* s_red = reinterpret_cast<double*>(smem + L.en0);
        if (lane == 0) { s_red[wid - kRedWarp0] = dsum; s_red[8 + wid - kRedWarp0] = esum; }
        asm volatile("bar.sync 1, 256;" ::: "memory");
        if (t0 == 0) {
            double dd = 0.0, ee = 0.0;
#pragma unroll
            for (int w = 0; w < 8; ++w) { dd += s_red[w]; ee += s_red[8 + w]; }
            out[0] = dd - ee;
            out[1] = dd;
            out[2] = ee;
            out[3] = 0.0;
            *ticket = 0u;
        }
        if (loss_args.peer_bufs != nullptr) {      // multi-GPU: exchange the rank partials over peer memory right here
            asm volatile("bar.sync 1, 256;" ::: "memory");
            p2p_loss_exchange<double>(out, loss_args, t0, [] { asm volatile("bar.sync 1, 256;" ::: "memory"); });
        }
    }
}

static bool pairs9(const hidenn_tri_plan* p) { return p->dev8.pair_pack != nullptr; }
static size_t smem9_for(const hidenn_tri_plan* p, int n_stages) {
    return pairs9(p) ? (size_t)smem9_layout(p->dev.max_local, p->dev8.max_entries9, p->dev8.stride_pent * 16, p->dev.stride_owned, n_stages, Roles9<true>::E).total
                     : (size_t)smem9_layout(p->dev.max_local, p->dev.max_entries, p->dev.stride_elem * 8, p->dev.stride_owned, n_stages, Roles9<false>::E).total;
}
// as many tile stages as fit (4 if possible: one more tile of slack between the bulk copies and the element warps)
static int stages9_for(const hidenn_tri_plan* p) {
    static const int env = [] { const char* e = getenv("HIDENN_WS_STAGES"); return e ? atoi(e) : 0; }();
    if (env >= 2 && env <= kMaxStages) return env;
    return smem9_for(p, kMaxStages) <= (size_t)227 * 1024 ? kMaxStages : 3;
}
size_t tile9_smem_bytes(const hidenn_tri_plan* p) { return smem9_for(p, stages9_for(p)); }

template <bool BODY, bool ISO, bool PAIRS, bool JT>
static int launch9(const hidenn_tri_plan* p, const double* x_free, const double* x_fixed, const double* u_free, const double* u_fixed,
                   const double* consts, const double* t_table, int flags, double* out, double* gx, double* gu, double* gt,
                   double* scratch, unsigned* ticket, cudaStream_t stream, int tile_begin, int tile_end, unsigned* first_done,
                   int reserve_sms, const P2PLossArgs* loss_args) {
    const size_t smem = tile9_smem_bytes(p);
    static size_t configured[kMaxDevices] = {};      // per instantiation (static local of a function template)
    size_t& cfg = configured[p->device % kMaxDevices];
    if (smem > cfg) {
        HIDENN_CUDA_OK(cudaFuncSetAttribute(tri_tile9_kernel<BODY, ISO, PAIRS, JT>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        cfg = smem;
    }
    TriPlanDev P = p->dev;
    TriPlan8Dev P8 = p->dev8;
    const int n_total = p->dev.n_tiles;
    P.n_tiles = tile_end - tile_begin;
    P8.tiles += tile_begin;
    P8.t_halo += (size_t)tile_begin * P8.stride_halo;
    if (PAIRS) {
        P8.pair_pack += (size_t)tile_begin * P8.stride_pent * 2;
        P8.entry_off9 += (size_t)tile_begin * P.stride_owned;
    } else {
        P.elem_pack += (size_t)tile_begin * P.stride_elem;
        P.entry_off += (size_t)tile_begin * P.stride_owned;
    }
    const int grid = std::min(P.n_tiles, std::max(1, sm_count(p->device) - reserve_sms));
    tri_tile9_kernel<BODY, ISO, PAIRS, JT><<<grid, kThreads9, smem, stream>>>(
        P, P8, (const double2*)x_free, (const double2*)x_fixed, (const double2*)u_free, (const double2*)u_fixed, consts, t_table, flags,
        (double2*)gx, (double2*)gu, gt, scratch + tile_begin, scratch + n_total + tile_begin, scratch, scratch + n_total, n_total, out, ticket,
        stages9_for(p), first_done, tile_begin == 0 ? p->n_first_tiles : 0, loss_args ? *loss_args : P2PLossArgs{nullptr, nullptr, nullptr, 0, 0, 0});
    HIDENN_CUDA_OK(cudaGetLastError());
    return 0;
}

int tile9_fold_warps(const hidenn_tri_plan* p) { return pairs9(p) ? Roles9<true>::F : Roles9<false>::F; }
bool tile9_fits(const hidenn_tri_plan* p) { return tile9_smem_bytes(p) <= (size_t)227 * 1024; }

int tile9_launch(const hidenn_tri_plan* p, const double* x_free, const double* x_fixed, const double* u_free, const double* u_fixed,
                 const double* consts, const double* t_table, int flags, double* out, double* gx, double* gu, double* gt, double* scratch,
                 unsigned* ticket, cudaStream_t stream, int tile_begin, int tile_end, unsigned* first_done, int reserve_sms,
                 const P2PLossArgs* loss_args) {
    const bool body = !(flags & HIDENN_HINT_NO_BODY_FORCE), iso = (flags & HIDENN_HINT_C_PLANE_STRESS) != 0;
#define HIDENN_L9(B_, I_, P_) \
    do { \
        if (p->dev.jinv_t) \
            return launch9<B_, I_, P_, true>(p, x_free, x_fixed, u_free, u_fixed, consts, t_table, flags, out, gx, gu, gt, scratch, ticket, stream, \
                                             tile_begin, tile_end, first_done, reserve_sms, loss_args); \
        return launch9<B_, I_, P_, false>(p, x_free, x_fixed, u_free, u_fixed, consts, t_table, flags, out, gx, gu, gt, scratch, ticket, stream, \
                                          tile_begin, tile_end, first_done, reserve_sms, loss_args); \
    } while (0)
    if (pairs9(p)) {
        if (body && iso) HIDENN_L9(true, true, true);
        if (body) HIDENN_L9(true, false, true);
        if (iso) HIDENN_L9(false, true, true);
        HIDENN_L9(false, false, true);
    }
    if (body && iso) HIDENN_L9(true, true, false);
    if (body) HIDENN_L9(true, false, false);
    if (iso) HIDENN_L9(false, true, false);
    HIDENN_L9(false, false, false);
#undef HIDENN_L9
}

}  // namespace hidenn
