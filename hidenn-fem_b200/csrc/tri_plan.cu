// Host-side plan builder for the triangle path (static topology, once per mesh).
//
// Replaces, for the hot path, what the reference redoes on every call:
//   * masked index_put assembly of coords / u_full   (/root/reference/src/models.py:292-305)  -> slot maps
//   * coords[connectivity[idx]] gathers               (/root/reference/src/models.py:228-238)  -> tile-local packs
//   * autograd's scatter-add backward of those gathers                                         -> per-tile node->element CSR
//
// Tiling: recursive coordinate bisection of the nodes (initial coordinates) into compact blobs of
// ~tile_nodes owned nodes; a tile visits every element incident to an owned node ("owner computes",
// halo elements recomputed), so each nodal gradient is folded and written by exactly one CTA in a
// fixed order: deterministic, no float atomics, no second pass.
#include <climits>
#include <cstdio>
#include <cstdlib>
#include "../../include/hidenn_b200.h"
#include "common.cuh"
#include "tri_plan.h"

#include <algorithm>
#include <atomic>
#include <cstring>
#include <cstdlib>
#include <functional>
#include <future>
#include <memory>
#include <numeric>
#include <thread>

namespace hidenn {

// default owned nodes per tile: close to the largest tile whose fold slots fit the 11-bit position fields at valence ~6
constexpr int kDefaultTileNodes = 320;
// with the paired layout a node has ~4 fold slots instead of 6 and an entry is two elements: tiles are sized so that a
// tile's entries (pairs + singles + class padding) fill the 512 element lanes of kernel v9 in ONE pass -- the time of a tile
// hardly depends on how full that pass is, so fewer, fuller tiles are faster (C4: 160 us at 320 nodes, 173 us at 288)
#ifndef HIDENN_TILE_NODES_PAIRS
#define HIDENN_TILE_NODES_PAIRS 352
#endif
constexpr int kDefaultTileNodesPairs = HIDENN_TILE_NODES_PAIRS;

static thread_local std::string g_err;
void set_error(const std::string& msg) { g_err = msg; }

template <typename T> static int upload(hidenn_tri_plan* p, const std::vector<T>& h, const T** d) {
    void* ptr = nullptr;
    size_t bytes = std::max<size_t>(h.size(), 1) * sizeof(T);
    HIDENN_CUDA_OK(cudaMalloc(&ptr, bytes));
    p->dev_allocs.push_back(ptr);
    p->dev_bytes += bytes;
    if (!h.empty()) HIDENN_CUDA_OK(cudaMemcpy(ptr, h.data(), h.size() * sizeof(T), cudaMemcpyHostToDevice));
    *d = reinterpret_cast<const T*>(ptr);
    return 0;
}

struct Rcb {
    const double* xy;
    std::vector<int32_t>& idx;
    std::vector<int64_t>& tile_begin;   // size n_tiles+1 filled by leaves
    void run(int64_t lo, int64_t hi, int64_t t_lo, int64_t t_hi, int depth) {
        const int64_t nt = t_hi - t_lo;
        if (nt <= 1) {
            tile_begin[t_lo] = lo;
            return;
        }
        double mn[2] = {1e300, 1e300}, mx[2] = {-1e300, -1e300};
        for (int64_t i = lo; i < hi; ++i) {
            const double* p = xy + 2 * (int64_t)idx[i];
            mn[0] = std::min(mn[0], p[0]); mx[0] = std::max(mx[0], p[0]);
            mn[1] = std::min(mn[1], p[1]); mx[1] = std::max(mx[1], p[1]);
        }
        const int ax = (mx[1] - mn[1] > mx[0] - mn[0]) ? 1 : 0;
        const int64_t tl = nt / 2;
        const int64_t nl = (hi - lo) * tl / nt;
        const double* c = xy;
        // ties are broken by the other coordinate (then by id): the leaves do not depend on the node numbering, so a mesh
        // renumbered by hidenn_tri_locality_order is cut into the same tiles again by the plan
        auto cmp = [c, ax](int32_t a, int32_t b) {
            const double va = c[2 * (int64_t)a + ax], vb = c[2 * (int64_t)b + ax];
            if (va != vb) return va < vb;
            const double wa = c[2 * (int64_t)a + 1 - ax], wb = c[2 * (int64_t)b + 1 - ax];
            return wa < wb || (wa == wb && a < b);
        };
        std::nth_element(idx.begin() + lo, idx.begin() + lo + nl, idx.begin() + hi, cmp);
        if (depth < 3 && hi - lo > 200000) {
            auto f = std::async(std::launch::async, [=] { this->run(lo, lo + nl, t_lo, t_lo + tl, depth + 1); });
            run(lo + nl, hi, t_lo + tl, t_hi, depth + 1);
            f.get();
        } else {
            run(lo, lo + nl, t_lo, t_lo + tl, depth + 1);
            run(lo + nl, hi, t_lo + tl, t_hi, depth + 1);
        }
    }
};


struct TileBuild {
    std::vector<int32_t> nodes;     // owned (ascending) then halo (ascending)
    int32_t n_owned = 0;
    std::vector<int32_t> elems;     // ascending global element id
    std::vector<unsigned long long> pack;
    std::vector<uint32_t> off;      // n_owned: fold-slot start | count << 16
    int32_t n_entries = 0;          // padded slot count (= dump slot index)
    int err = 0;
    std::vector<unsigned long long> epack;   // tile-ordered layout: Neumann edge visits
    std::vector<int32_t> eid;
    bool unpaired_overflow = false;                  // one-element-per-entry packs unusable (plan is pairs-only)
    std::vector<unsigned long long> pack9, epack9;   // paired layout (2 words per entry)
    std::vector<uint32_t> off9;
    int32_t n_entries9 = 0;
};

// Lane assignment inside a tile.  Which thread handles which element is free (every fold slot is written exactly
// once and read in slot order), so elements are greedily grouped such that the lanes that access shared memory
// in the same pass -- 8 lanes for 16-byte pairs (FP64), 16 lanes for 8-byte pairs (FP32) -- hit different bank
// groups in the 3 gathers and the 3 partial stores.  Random placement costs ~2.2 wavefronts per pass.
static void reorder_for_banks(TileBuild& B, int real_bytes) {
    const int E = (int)B.pack.size();
    (void)real_bytes;
    const int G = 8;          // lanes per 128-bit shared-memory pass = number of 16-byte bank groups (both precisions)
    if (E <= G) return;
    constexpr unsigned LM = (1u << kLidBits) - 1u, PM = (1u << kPosBits) - 1u;
    const unsigned dump = (unsigned)B.n_entries;
    struct Item { uint16_t lid[3], pos[3]; };
    std::vector<Item> it(E);
    for (int i = 0; i < E; ++i) {
        const unsigned long long w = B.pack[i];
        for (int c = 0; c < 3; ++c) {
            it[i].lid[c] = (uint16_t)((w >> (kLidBits * c)) & LM);
            it[i].pos[c] = (uint16_t)((w >> (3 * kLidBits + kPosBits * c)) & PM);
        }
    }
    std::vector<char> used(E, 0);
    std::vector<int> ord;
    ord.reserve(E);
    static const int W = [] { const char* e = getenv("HIDENN_PLAN_WINDOW"); return e ? atoi(e) : 512; }();
    if (W <= 0) return;
    int head = 0;
    while ((int)ord.size() < E) {
        int cg[3][16] = {}, cp[3][16] = {};
        int lidat[3][16];
        for (int c = 0; c < 3; ++c) for (int b = 0; b < 16; ++b) lidat[c][b] = -1;
        for (int slot = 0; slot < G && (int)ord.size() < E; ++slot) {
            int best = -1, best_cost = 1 << 30, seen = 0;
            for (int j = head; j < E && seen < W; ++j) {
                if (used[j]) continue;
                ++seen;
                int cost = 0;
                for (int c = 0; c < 3; ++c) {
                    const int b = it[j].lid[c] % G;
                    if (cg[c][b] > 0 && lidat[c][b] != (int)it[j].lid[c]) cost += cg[c][b];     // same address = broadcast
                    if (it[j].pos[c] != dump) cost += cp[c][it[j].pos[c] % G];     // == lid % G for owned corners
                }
                if (cost < best_cost) { best_cost = cost; best = j; if (cost == 0) break; }
            }
            used[best] = 1;
            ord.push_back(best);
            for (int c = 0; c < 3; ++c) {
                const int b = it[best].lid[c] % G;
                if (lidat[c][b] != (int)it[best].lid[c]) { cg[c][b]++; lidat[c][b] = it[best].lid[c]; }
                if (it[best].pos[c] != dump) cp[c][it[best].pos[c] % G]++;
            }
            while (head < E && used[head]) ++head;
        }
    }
    std::vector<unsigned long long> np(E);
    std::vector<int32_t> ne(E);
    for (int i = 0; i < E; ++i) { np[i] = B.pack[ord[i]]; ne[i] = B.elems[ord[i]]; }
    B.pack.swap(np);
    B.elems.swap(ne);
}

// Same greedy lane assignment for the pair entries of the paired layout (2 words per entry): six gather columns (corners
// of the first and second element) and six store columns; a null second word contributes nothing.
// Class of a pair entry (tri_plan.h): 3 * i + r for a pair whose second element has its new corner at position r and
// shares the edge (corner i -> corner i+1) of the first element, run through in the opposite direction; kPairSingle for
// a single element; -1 if the two words are not such a pair.
int pair_class(unsigned long long w1, unsigned long long w2) {
    constexpr unsigned LM = (1u << kLidBits) - 1u;
    if (((unsigned)w1 & 0x3FFFFFFFu) == 0x3FFFFFFFu) return kPairSkip;
    if (((unsigned)w2 & 0x3FFFFFFFu) == 0x3FFFFFFFu) return kPairSingle;
    unsigned l[3], m[3];
    for (int c = 0; c < 3; ++c) { l[c] = (unsigned)(w1 >> (kLidBits * c)) & LM; m[c] = (unsigned)(w2 >> (kLidBits * c)) & LM; }
    for (int r = 0; r < 3; ++r) {
        if (m[r] == l[0] || m[r] == l[1] || m[r] == l[2]) continue;
        for (int i = 0; i < 3; ++i)
            if (m[(r + 1) % 3] == l[(i + 1) % 3] && m[(r + 2) % 3] == l[i]) return 3 * i + r;
        return -1;
    }
    return -1;
}

// (pairs of one class: the kernel gathers the three corners of the first element and the new corner of the second,
// four columns; singles: three)
static void reorder_pairs_for_banks(std::vector<unsigned long long>& pack9, unsigned dump) {
    const int E = (int)(pack9.size() / 2);
    const int G = 8;
    if (E <= G) return;
    static const int W = [] { const char* e = getenv("HIDENN_PLAN_WINDOW"); return e ? atoi(e) : 512; }();
    if (W <= 0) return;
    constexpr unsigned LM = (1u << kLidBits) - 1u, PM = (1u << kPosBits) - 1u;
    struct Item { uint16_t lid[6], pos[6]; int nc; };
    std::vector<Item> it(E);
    for (int i = 0; i < E; ++i) {
        const int cls = pair_class(pack9[2 * i], pack9[2 * i + 1]);
        it[i].nc = cls == kPairSingle ? 3 : 4;
        const unsigned long long w = pack9[2 * i], w2 = pack9[2 * i + 1];
        for (int c = 0; c < 3; ++c) {
            it[i].lid[c] = (uint16_t)((w >> (kLidBits * c)) & LM);
            it[i].pos[c] = (uint16_t)((w >> (3 * kLidBits + kPosBits * c)) & PM);
        }
        if (cls != kPairSingle) {
            const int r = cls < 0 ? 0 : cls % 3;
            it[i].lid[3] = (uint16_t)((w2 >> (kLidBits * r)) & LM);
            it[i].pos[3] = (uint16_t)((w2 >> (3 * kLidBits + kPosBits * r)) & PM);
        }
    }
    std::vector<char> used(E, 0);
    std::vector<int> ord;
    ord.reserve(E);
    int head = 0;
    while ((int)ord.size() < E) {
        int cg[6][8] = {}, cp[6][8] = {};
        int lidat[6][8];
        for (int c = 0; c < 6; ++c) for (int b = 0; b < 8; ++b) lidat[c][b] = -1;
        for (int slot = 0; slot < G && (int)ord.size() < E; ++slot) {
            int best = -1, best_cost = 1 << 30, seen = 0;
            for (int j = head; j < E && seen < W; ++j) {
                if (used[j]) continue;
                ++seen;
                int cost = 0;
                for (int c = 0; c < it[j].nc; ++c) {
                    const int b = it[j].lid[c] % G;
                    if (cg[c][b] > 0 && lidat[c][b] != (int)it[j].lid[c]) cost += cg[c][b];     // same address = broadcast
                    if (it[j].pos[c] != dump) cost += cp[c][it[j].pos[c] % G];
                }
                if (cost < best_cost) { best_cost = cost; best = j; if (cost == 0) break; }
            }
            used[best] = 1;
            ord.push_back(best);
            for (int c = 0; c < it[best].nc; ++c) {
                const int b = it[best].lid[c] % G;
                if (lidat[c][b] != (int)it[best].lid[c]) { cg[c][b]++; lidat[c][b] = it[best].lid[c]; }
                if (it[best].pos[c] != dump) cp[c][it[best].pos[c] % G]++;
            }
            while (head < E && used[head]) ++head;
        }
    }
    std::vector<unsigned long long> np(pack9.size());
    for (int i = 0; i < E; ++i) { np[2 * i] = pack9[2 * ord[i]]; np[2 * i + 1] = pack9[2 * ord[i] + 1]; }
    pack9.swap(np);
}

int plan_ensure_generic(hidenn_tri_plan* p) {
    if (p->generic_uploaded) return 0;
    HIDENN_REQUIRE(p->device >= 0, "host-only plan (device=-1) cannot run kernels");
    DeviceScope scope;
    HIDENN_CUDA_OK(scope.enter(p->device));
    if (upload(p, p->conn32, &p->dev.conn32)) return 1;
    if (upload(p, p->xslot, &p->dev.xslot)) return 1;
    if (upload(p, p->uslot, &p->dev.uslot)) return 1;
    if (upload(p, p->n2e_off, &p->dev.n2e_off)) return 1;
    if (upload(p, p->n2e_ent, &p->dev.n2e_ent)) return 1;
    if (upload(p, p->edges32, &p->dev.edges32)) return 1;
    p->generic_uploaded = true;
    return 0;
}

int plan_ensure_arena(hidenn_tri_plan* p, size_t bytes) {
    if (p->arena_bytes >= bytes) return 0;
    DeviceScope scope;
    HIDENN_CUDA_OK(scope.enter(p->device));
    if (p->arena) cudaFree(p->arena);
    p->arena = nullptr;
    p->arena_bytes = 0;
    HIDENN_CUDA_OK(cudaMalloc(&p->arena, bytes));
    p->arena_bytes = bytes;
    return 0;
}

}  // namespace hidenn

using namespace hidenn;

extern "C" const char* hidenn_last_error(void) { return hidenn::g_err.c_str(); }
extern "C" int hidenn_version(void) { return HIDENN_B200_VERSION; }
extern "C" int hidenn_device_count(void) {
    int n = 0;
    cudaError_t e = cudaGetDeviceCount(&n);
    if (e != cudaSuccess) {
        set_error(std::string("cudaGetDeviceCount: ") + cudaGetErrorString(e));
        return -1;
    }
    return n;
}

// class of a node in the tile-ordered numbering (tri_plan.h): A 0, B 1, C 2, D 3
static inline int node_class(uint8_t bmask, uint8_t dmask) { return bmask ? (dmask ? 2 : 1) : (dmask ? 3 : 0); }

// Fold slots a tile needs when its owned nodes are listed by (class, descending slot count) and padded in groups of 8:
// the numbering-independent bound both the locality ordering and the plan test against the pack limit, so that they
// settle on the same tile size.  cs = (class, slots) per owned node; sorted in place.
static int64_t canon_padded_entries(std::vector<std::pair<int32_t, int32_t>>& cs) {
    std::sort(cs.begin(), cs.end(), [](const std::pair<int32_t, int32_t>& a, const std::pair<int32_t, int32_t>& b) {
        return a.first != b.first ? a.first < b.first : a.second > b.second;
    });
    int64_t acc = 0;
    for (size_t g0 = 0; g0 < cs.size(); g0 += 8) {
        int32_t mx = 0;
        for (size_t g = g0; g < std::min(cs.size(), g0 + 8); ++g) mx = std::max(mx, cs[g].second);
        acc += (int64_t)mx * 8;
    }
    return acc;
}

struct EdgeEnds {       // Neumann edge ends by node: (node, edge*2+end), sorted
    std::vector<std::pair<int32_t, int32_t>> v;
    void build(const int64_t* edges, int64_t Ned) {
        v.resize(2 * Ned);
        for (int64_t i = 0; i < 2 * Ned; ++i) v[i] = {(int32_t)edges[i], (int32_t)i};
        std::sort(v.begin(), v.end());
    }
    std::pair<const std::pair<int32_t, int32_t>*, const std::pair<int32_t, int32_t>*> of(int32_t n) const {
        auto lo = std::lower_bound(v.begin(), v.end(), std::make_pair(n, (int32_t)INT32_MIN));
        auto hi = lo;
        while (hi != v.end() && hi->first == n) ++hi;
        return {v.data() + (lo - v.begin()), v.data() + (hi - v.begin())};
    }
};

// Global matching of the elements into edge-sharing pairs whose partners run through the shared edge in opposite
// directions (tri_plan.h).  Depends only on the mesh (connectivity + node->element lists), never on the tiling.
// HIDENN_PLAN_PAIRS: unset = automatic (kept when at least 80 % of the elements find a partner inside at most three wiring
// classes), 0 = never, 1 = always.
// Returns the number of pairs; `mate` stays empty when the paired layout is not used.
static int64_t match_elements(const int32_t* c32g, int64_t Ne, const int64_t* n2o, const int32_t* n2e, std::vector<int32_t>& mate,
                              int* n_classes_out = nullptr) {
    mate.clear();
    if (n_classes_out) *n_classes_out = 0;
    const char* pairs_env = getenv("HIDENN_PLAN_PAIRS");
    const bool want_pairs = pairs_env == nullptr || atoi(pairs_env) != 0;
    const bool force_pairs = pairs_env != nullptr && atoi(pairs_env) != 0;
    if (!want_pairs || Ne <= 0) return 0;
    int64_t n_pairs = 0;
    {
        std::vector<int32_t> nb(3 * Ne, -1);      // neighbour across the edge (corner c -> c+1): f * 4 + (corner of f at the edge's start node)
        auto nb_range = [&](int64_t e0, int64_t e1) {
            for (int64_t e = e0; e < e1; ++e)
                for (int c = 0; c < 3; ++c) {
                    // only a neighbour that runs through the shared edge in the opposite direction (b -> a; both elements
                    // counter-clockwise or both clockwise) can be this element's partner: the kernel has one register wiring
                    // per (edge of the first element, position of the new corner in the second) = 9 classes
                    const int32_t a = c32g[3 * e + c], b = c32g[3 * e + (c + 1) % 3];
                    for (int64_t k = n2o[a]; k < n2o[a + 1]; ++k) {
                        const int32_t f = n2e[k] >> 2;
                        const int cf = n2e[k] & 3;      // corner of f that is a
                        if (f != e && c32g[3 * (int64_t)f + (cf + 2) % 3] == b) { nb[3 * e + c] = f * 4 + cf; break; }
                    }
                }
        };
        {
            unsigned nthr = std::max(1u, std::min(16u, std::thread::hardware_concurrency()));
            if (Ne < 100000) nthr = 1;
            std::vector<std::thread> th;
            const int64_t per = (Ne + nthr - 1) / nthr;
            for (unsigned i = 0; i < nthr; ++i) {
                const int64_t a = i * per, b = std::min<int64_t>(Ne, a + per);
                if (a < b) th.emplace_back(nb_range, a, b);
            }
            for (auto& t : th) t.join();
        }
        // wiring class of the pair (e, neighbour across e's edge c), first element = smaller id (tri_plan.h)
        auto cls_of = [&](int64_t e, int c) -> int {
            const int32_t v = nb[3 * e + c];
            if (v < 0) return -1;
            const int64_t f = v >> 2;
            const int cf = v & 3;
            return e < f ? 3 * c + (cf + 1) % 3 : 3 * ((cf + 2) % 3) + (c + 2) % 3;
        };
        // Which classes?  Meshes from generators number their corners by a few patterns, and a tile's entries are listed
        // class by class with every class padded to whole warps, so the matching should live in as few classes as
        // possible: the PAIR of classes whose greedy matching (element order, free neighbour of an allowed class) covers
        // most of a sample of the mesh goes first, then the remaining classes one at a time, most frequent first.
        int64_t hist[9] = {};
        for (int64_t e = 0; e < Ne; ++e)
            for (int c = 0; c < 3; ++c) {
                const int32_t v = nb[3 * e + c];
                if (v >= 0 && e < (v >> 2)) hist[cls_of(e, c)]++;
            }
        mate.assign(Ne, -1);
        auto greedy = [&](unsigned mask, int64_t e_end, std::vector<int32_t>& mt) -> int64_t {
            int64_t n = 0;
            for (int64_t e = 0; e < e_end; ++e) {
                if (mt[e] >= 0) continue;
                for (int c = 0; c < 3; ++c) {
                    const int32_t v = nb[3 * e + c];
                    if (v < 0 || (v >> 2) >= e_end || mt[v >> 2] >= 0 || !((mask >> cls_of(e, c)) & 1u)) continue;
                    mt[e] = v >> 2; mt[v >> 2] = (int32_t)e; ++n;
                    break;
                }
            }
            return n;
        };
        unsigned best_mask = 0;
        {
            const int64_t ns = std::min<int64_t>(Ne, 1 << 20);
            std::vector<int32_t> tmp;
            int64_t best = -1;
            for (int x = 0; x < 9; ++x)
                for (int y = x; y < 9; ++y) {
                    if (hist[x] == 0 || hist[y] == 0) continue;
                    tmp.assign(ns, -1);
                    const int64_t n = greedy((1u << x) | (1u << y), ns, tmp);
                    if (n > best) { best = n; best_mask = (1u << x) | (1u << y); }
                }
        }
        static const int max_classes = [] { const char* e = getenv("HIDENN_PLAN_PAIR_CLASSES"); return e ? std::max(1, std::min(9, atoi(e))) : 9; }();
        n_pairs += greedy(best_mask, Ne, mate);
        int corder[9];
        std::iota(corder, corder + 9, 0);
        std::stable_sort(corder, corder + 9, [&](int x, int y) { return hist[x] > hist[y]; });
        int used = __builtin_popcount(best_mask);
        for (int ci = 0; ci < 9 && used < max_classes; ++ci) {
            const int cls = corder[ci];
            if (hist[cls] == 0 || ((best_mask >> cls) & 1u)) continue;
            n_pairs += greedy(1u << cls, Ne, mate);
            ++used;
        }
        // classes that hold at least 1 % of the pairs: each costs a tile up to 31 padding entries and the kernel one more
        // wiring to keep in the instruction cache
        int n_classes = 0;
        {
            int64_t cnt[9] = {};
            for (int64_t e = 0; e < Ne; ++e)
                if (mate[e] > e)
                    for (int c = 0; c < 3; ++c)
                        if (nb[3 * e + c] >= 0 && (nb[3 * e + c] >> 2) == mate[e]) { cnt[cls_of(e, c)]++; break; }
            for (int c = 0; c < 9; ++c) n_classes += (cnt[c] * 100 >= n_pairs && cnt[c] > 0) ? 1 : 0;
        }
        if (n_classes_out) *n_classes_out = n_classes;
        // Automatic mode keeps the paired layout only where it wins: at least 80 % of the elements paired, inside at most
        // three classes.  Measured on B200 (C4 geometry, 10 M elements): generator mesh (two classes) 147 us paired /
        // 173 us one element per entry; the same mesh with every element's corners rotated at random (all nine classes,
        // 240-node tiles) 284 us paired / 176 us one element per entry.
        if (!force_pairs && (n_pairs * 10 < Ne * 4 || n_classes > 3)) {
            mate.clear();
            n_pairs = 0;
        }
    }
    return n_pairs;
}

// Owned nodes per tile for the paired layout: a tile of T nodes visits ~2.25 T elements, i.e. ~2.25 T (1 - pf / 2) entries
// when a fraction pf of the elements is paired (a few per cent more: partners the tile does not visit), plus ~16 padding
// entries per class and for the singles; the entries should
// fill the 512 element lanes of kernel v9 in ONE pass (352 nodes for a generator mesh with two classes, ~290 when the
// corner order is random and all nine classes occur).
static int pairs_tile_nodes(int64_t Ne, int64_t n_pairs, int n_classes) {
    const double pf = Ne > 0 ? 2.0 * (double)n_pairs / (double)Ne : 0.0;
    const double t = (500.0 - 16.0 * (n_classes + 1)) / (2.25 * (1.0 - 0.5 * pf + 0.06));      // + 6 %: partners outside the tile
    return (int)std::max(128.0, std::min((double)kDefaultTileNodesPairs, t)) / 8 * 8;
}

// fold slots of node n in the paired layout: one per incident element whose partial is not merged into its (smaller-id)
// partner's, i.e. one per pair or single
static inline int32_t paired_elem_slots(int32_t n, const int32_t* c32, const int64_t* n2o, const int32_t* n2e, const std::vector<int32_t>& mate) {
    int32_t c = 0;
    for (int64_t k = n2o[n]; k < n2o[n + 1]; ++k) {
        const int32_t e = n2e[k] >> 2, f = mate[e];
        const bool merged = f >= 0 && f < e && (c32[3 * (int64_t)f] == n || c32[3 * (int64_t)f + 1] == n || c32[3 * (int64_t)f + 2] == n);
        c += merged ? 0 : 1;
    }
    return c;
}

// Does a leaf of the bisection fit the pack limits (local ids <= 1022, fold slots <= 2047)?  Depends only on the set of
// owned nodes, not on the numbering: local nodes = owned + nodes of incident elements + other ends of incident Neumann
// edges; slots by canon_padded_entries.
static bool leaf_feasible(const int32_t* ids, int64_t n, const int32_t* c32, const int64_t* n2o, const int32_t* n2e, const uint8_t* bmask,
                          const uint8_t* dmask, const EdgeEnds& ee, const int64_t* edges, std::vector<int32_t>& owned,
                          std::vector<int32_t>& loc, std::vector<std::pair<int32_t, int32_t>>& cs, const std::vector<int32_t>& mate) {
    owned.assign(ids, ids + n);
    std::sort(owned.begin(), owned.end());
    loc.clear();
    cs.clear();
    for (int32_t nd : owned) {
        for (int64_t k = n2o[nd]; k < n2o[nd + 1]; ++k) {
            const int64_t e = n2e[k] >> 2;
            for (int c = 0; c < 3; ++c) loc.push_back(c32[3 * e + c]);
        }
        auto r = ee.of(nd);
        for (auto it = r.first; it != r.second; ++it) loc.push_back((int32_t)edges[it->second ^ 1]);
        // with a matching only the paired layout has to fit (the one-element-per-entry packs of such a plan may overflow
        // and are then marked unusable)
        const int32_t es = mate.empty() ? (int32_t)(n2o[nd + 1] - n2o[nd]) : paired_elem_slots(nd, c32, n2o, n2e, mate);
        cs.push_back({node_class(bmask[nd], dmask[nd]), (int32_t)(es + (r.second - r.first))});
        loc.push_back(nd);
    }
    std::sort(loc.begin(), loc.end());
    loc.erase(std::unique(loc.begin(), loc.end()), loc.end());
    if ((int64_t)loc.size() > kMaxLocal) return false;
    return canon_padded_entries(cs) <= kMaxEntries;
}

extern "C" int hidenn_tri_locality_order(const int64_t* conn, int64_t Ne, int64_t Nn, const double* coords, const uint8_t* bmask,
                                         const uint8_t* dmask, const int64_t* edges, int64_t Ned, int tile_nodes, int64_t* new_to_old,
                                         int64_t* elem_new_to_old) {
    HIDENN_REQUIRE(conn && coords && bmask && dmask && new_to_old && elem_new_to_old, "locality_order: NULL argument");
    HIDENN_REQUIRE(Ne >= 0 && Nn > 0 && Nn < (int64_t)2147483000, "locality_order: sizes out of range");
    HIDENN_REQUIRE(Ned == 0 || edges != nullptr, "locality_order: edges NULL");
    const bool tile_nodes_given = tile_nodes > 0;
    if (tile_nodes <= 0) tile_nodes = kDefaultTileNodes;
    HIDENN_REQUIRE(tile_nodes >= 8 && tile_nodes <= 2048, "locality_order: tile_nodes must be in [8,2048]");
    for (int64_t i = 0; i < 3 * Ne; ++i) HIDENN_REQUIRE(conn[i] >= 0 && conn[i] < Nn, "locality_order: connectivity index out of range");
    for (int64_t i = 0; i < 2 * Ned; ++i) HIDENN_REQUIRE(edges[i] >= 0 && edges[i] < Nn, "locality_order: edge node index out of range");
    // node -> element lists (same as the plan's)
    std::vector<int32_t> c32(3 * Ne);
    for (int64_t i = 0; i < 3 * Ne; ++i) c32[i] = (int32_t)conn[i];
    std::vector<int64_t> n2o(Nn + 1, 0);
    for (int64_t i = 0; i < 3 * Ne; ++i) n2o[c32[i] + 1]++;
    for (int64_t n = 0; n < Nn; ++n) n2o[n + 1] += n2o[n];
    std::vector<int32_t> n2e(3 * Ne);
    {
        std::vector<int64_t> cur(n2o.begin(), n2o.end() - 1);
        for (int64_t e = 0; e < Ne; ++e)
            for (int c = 0; c < 3; ++c) n2e[cur[c32[3 * e + c]]++] = (int32_t)(e * 4 + c);
    }
    EdgeEnds ee;
    ee.build(edges, Ned);
    // the same global matching the plan will compute: with it the tiles are sized for the paired layout
    std::vector<int32_t> mate;
    int n_classes = 0;
    int64_t n_pairs_lo = 0;
    if (!getenv("HIDENN_PLAN_NO_V8")) n_pairs_lo = match_elements(c32.data(), Ne, n2o.data(), n2e.data(), mate, &n_classes);
    if (!mate.empty() && !tile_nodes_given) tile_nodes = pairs_tile_nodes(Ne, n_pairs_lo, n_classes);
    std::vector<int32_t> slots(Nn);
    for (int64_t n = 0; n < Nn; ++n) {
        auto r = ee.of((int32_t)n);
        const int32_t es = mate.empty() ? (int32_t)(n2o[n + 1] - n2o[n]) : paired_elem_slots((int32_t)n, c32.data(), n2o.data(), n2e.data(), mate);
        slots[n] = (int32_t)(es + (r.second - r.first));
    }
    int64_t n_tiles = 0;
    std::vector<int32_t> order(Nn);
    std::vector<int64_t> tile_begin;
    auto run_parallel = [&](const std::function<void(int64_t, int64_t)>& f) {
        unsigned nthr = std::max(1u, std::min(16u, std::thread::hardware_concurrency()));
        if (n_tiles < 64) nthr = 1;
        std::vector<std::thread> th;
        const int64_t per = (n_tiles + nthr - 1) / nthr;
        for (unsigned i = 0; i < nthr; ++i) {
            const int64_t a = i * per, b = std::min<int64_t>(n_tiles, a + per);
            if (a < b) th.emplace_back(f, a, b);
        }
        for (auto& t : th) t.join();
    };
    // the plan's own search: bisect, test every leaf against the pack limits, shrink the tiles by 3/4 until all fit
    for (int attempt = 0;; ++attempt) {
        n_tiles = (Nn + tile_nodes - 1) / tile_nodes;
        std::iota(order.begin(), order.end(), 0);
        tile_begin.assign(n_tiles + 1, 0);
        tile_begin[n_tiles] = Nn;
        {
            Rcb r{coords, order, tile_begin};
            r.run(0, Nn, 0, n_tiles, 0);
        }
        std::atomic<int> bad{0};
        run_parallel([&](int64_t t0, int64_t t1) {
            std::vector<int32_t> owned, loc;
            std::vector<std::pair<int32_t, int32_t>> cs;
            for (int64_t t = t0; t < t1 && !bad.load(std::memory_order_relaxed); ++t)
                if (!leaf_feasible(order.data() + tile_begin[t], tile_begin[t + 1] - tile_begin[t], c32.data(), n2o.data(), n2e.data(), bmask,
                                   dmask, ee, edges, owned, loc, cs, mate))
                    bad.store(1);
        });
        if (!bad.load()) break;
        HIDENN_REQUIRE(attempt < 12 && tile_nodes > 8, "locality_order: cannot tile this mesh within the pack limits");
        tile_nodes = std::max(8, tile_nodes * 3 / 4);
    }
    // Order of the tiles: a boustrophedon sweep along the long axis of the mesh (strips about one tile wide across the
    // short axis) instead of the bisection order.  A tile's neighbours are then at most one strip away in the numbering:
    // the row window a range of tiles reads is barely wider than the rows it owns, which is what lets the host-buffer
    // entry point stream rows in and gradient rows out at the same time (tri_plan.h), and neighbouring CTAs still share
    // their halo rows through L2.
    {
        std::vector<double> cx(n_tiles), cy(n_tiles);
        double mn[2] = {1e300, 1e300}, mx[2] = {-1e300, -1e300};
        for (int64_t t = 0; t < n_tiles; ++t) {
            double sx = 0.0, sy = 0.0;
            for (int64_t i = tile_begin[t]; i < tile_begin[t + 1]; ++i) { sx += coords[2 * (int64_t)order[i]]; sy += coords[2 * (int64_t)order[i] + 1]; }
            const double cnt = (double)std::max<int64_t>(1, tile_begin[t + 1] - tile_begin[t]);
            cx[t] = sx / cnt; cy[t] = sy / cnt;
            mn[0] = std::min(mn[0], cx[t]); mx[0] = std::max(mx[0], cx[t]);
            mn[1] = std::min(mn[1], cy[t]); mx[1] = std::max(mx[1], cy[t]);
        }
        const int lg = (mx[1] - mn[1] > mx[0] - mn[0]) ? 1 : 0;                 // long axis
        const double ext_l = std::max(1e-300, mx[lg] - mn[lg]), ext_s = std::max(1e-300, mx[1 - lg] - mn[1 - lg]);
        const double per_strip = std::max(1.0, std::sqrt((double)n_tiles * ext_s / ext_l));
        const int64_t n_strips = std::max<int64_t>(1, (int64_t)std::llround((double)n_tiles / per_strip));
        std::vector<int64_t> tord(n_tiles), strip(n_tiles);
        std::iota(tord.begin(), tord.end(), 0);
        const std::vector<double>& cl = lg ? cy : cx;
        const std::vector<double>& cs = lg ? cx : cy;
        for (int64_t t = 0; t < n_tiles; ++t)
            strip[t] = std::min<int64_t>(n_strips - 1, (int64_t)((cl[t] - mn[lg]) / ext_l * (double)n_strips));
        std::stable_sort(tord.begin(), tord.end(), [&](int64_t a, int64_t b) {
            if (strip[a] != strip[b]) return strip[a] < strip[b];
            const double ka = (strip[a] & 1) ? -cs[a] : cs[a], kb = (strip[b] & 1) ? -cs[b] : cs[b];
            return ka < kb;
        });
        std::vector<int32_t> order2(Nn);
        std::vector<int64_t> begin2(n_tiles + 1, 0);
        int64_t pos = 0;
        for (int64_t k = 0; k < n_tiles; ++k) {
            const int64_t t = tord[k];
            begin2[k] = pos;
            for (int64_t i = tile_begin[t]; i < tile_begin[t + 1]; ++i) order2[pos++] = order[i];
        }
        begin2[n_tiles] = pos;
        order.swap(order2);
        tile_begin.swap(begin2);
    }
    // inside a tile: by class, then descending slot count (the fold groups of 8 consecutive nodes then have nearly equal
    // slot counts), then by old id
    run_parallel([&](int64_t t0, int64_t t1) {
        for (int64_t t = t0; t < t1; ++t)
            std::sort(order.begin() + tile_begin[t], order.begin() + tile_begin[t + 1], [&](int32_t a, int32_t b) {
                const int ca = node_class(bmask[a], dmask[a]), cb = node_class(bmask[b], dmask[b]);
                if (ca != cb) return ca < cb;
                if (slots[a] != slots[b]) return slots[a] > slots[b];
                return a < b;
            });
    });
    std::vector<int32_t> old_to_new(Nn);
    for (int64_t i = 0; i < Nn; ++i) { new_to_old[i] = order[i]; old_to_new[order[i]] = (int32_t)i; }
    // elements by their smallest new node id (stable): neighbouring elements stay close in memory
    std::vector<int32_t> key(Ne);
    for (int64_t e = 0; e < Ne; ++e)
        key[e] = std::min(old_to_new[conn[3 * e]], std::min(old_to_new[conn[3 * e + 1]], old_to_new[conn[3 * e + 2]]));
    std::vector<int64_t> eo(Ne);
    std::iota(eo.begin(), eo.end(), 0);
    std::stable_sort(eo.begin(), eo.end(), [&](int64_t a, int64_t b) { return key[a] < key[b]; });
    std::copy(eo.begin(), eo.end(), elem_new_to_old);
    return 0;
}

// Mesh ingestion (the step before the plan; /root/reference/src/mesh.py:70-88, 202-215 get the geometric boundary from
// gmsh entities / hole rims): topological boundary of a triangle mesh = nodes of the edges that belong to exactly one
// element.  The neighbour across an edge is looked up through the node -> element lists; host threads, no sort.
extern "C" int hidenn_mesh_boundary_nodes(const int64_t* conn, int64_t Ne, int64_t Nn, uint8_t* mask) {
    HIDENN_REQUIRE(conn && mask && Ne >= 0 && Nn > 0 && Nn < (int64_t)2147483000, "mesh_boundary_nodes: bad arguments");
    for (int64_t i = 0; i < 3 * Ne; ++i) HIDENN_REQUIRE(conn[i] >= 0 && conn[i] < Nn, "mesh_boundary_nodes: connectivity index out of range");
    std::vector<int64_t> off(Nn + 1, 0);
    for (int64_t i = 0; i < 3 * Ne; ++i) off[conn[i] + 1]++;
    for (int64_t n = 0; n < Nn; ++n) off[n + 1] += off[n];
    std::vector<int32_t> ent(3 * Ne);
    {
        std::vector<int64_t> cur(off.begin(), off.end() - 1);
        for (int64_t e = 0; e < Ne; ++e)
            for (int c = 0; c < 3; ++c) ent[cur[conn[3 * e + c]]++] = (int32_t)e;
    }
    std::vector<std::atomic<uint8_t>> m(Nn);
    for (auto& v : m) v.store(0, std::memory_order_relaxed);
    auto range = [&](int64_t e0, int64_t e1) {
        for (int64_t e = e0; e < e1; ++e)
            for (int c = 0; c < 3; ++c) {
                const int64_t a = conn[3 * e + c], b = conn[3 * e + (c + 1) % 3];
                bool shared = false;
                for (int64_t k = off[a]; k < off[a + 1] && !shared; ++k) {
                    const int64_t f = ent[k];
                    if (f != e && (conn[3 * f] == b || conn[3 * f + 1] == b || conn[3 * f + 2] == b)) shared = true;
                }
                if (!shared) { m[a].store(1, std::memory_order_relaxed); m[b].store(1, std::memory_order_relaxed); }
            }
    };
    unsigned nthr = std::max(1u, std::min(16u, std::thread::hardware_concurrency()));
    if (Ne < 100000) nthr = 1;
    std::vector<std::thread> th;
    const int64_t per = (Ne + nthr - 1) / nthr;
    for (unsigned i = 0; i < nthr; ++i) {
        const int64_t a = i * per, b = std::min<int64_t>(Ne, a + per);
        if (a < b) th.emplace_back(range, a, b);
    }
    for (auto& t : th) t.join();
    for (int64_t n = 0; n < Nn; ++n) mask[n] = m[n].load(std::memory_order_relaxed);
    return 0;
}

extern "C" int hidenn_tri_plan_create(const int64_t* conn, int64_t Ne, int64_t Nn, const double* coords,
                                      const uint8_t* bmask, const uint8_t* dmask, const int64_t* edges, int64_t Ned,
                                      int tile_nodes, int real_bytes, int device, hidenn_tri_plan** out) {
    return hidenn_tri_plan_create_ex(conn, Ne, Nn, coords, bmask, dmask, edges, Ned, nullptr, 0, 0, tile_nodes, real_bytes, device, out);
}

extern "C" int hidenn_tri_plan_create_ex(const int64_t* conn, int64_t Ne, int64_t Nn, const double* coords,
                                         const uint8_t* bmask, const uint8_t* dmask, const int64_t* edges, int64_t Ned,
                                         const int64_t* first_nodes, int64_t n_first, int options, int tile_nodes, int real_bytes,
                                         int device, hidenn_tri_plan** out) {
    HIDENN_REQUIRE(out != nullptr, "plan_create: out is NULL");
    *out = nullptr;
    HIDENN_REQUIRE(conn && coords && bmask && dmask, "plan_create: NULL input");
    HIDENN_REQUIRE(Ne >= 0 && Nn > 0 && Nn < (int64_t)2147483000 && Ne < (int64_t)500000000, "plan_create: sizes out of range");
    HIDENN_REQUIRE(real_bytes == 8 || real_bytes == 4, "plan_create: real_bytes must be 8 or 4");
    HIDENN_REQUIRE(Ned == 0 || edges != nullptr, "plan_create: edges NULL");
    HIDENN_REQUIRE(n_first == 0 || first_nodes != nullptr, "plan_create: first_nodes NULL");
    for (int64_t i = 0; i < n_first; ++i) HIDENN_REQUIRE(first_nodes[i] >= 0 && first_nodes[i] < Nn, "plan_create: first_nodes index out of range");
    const bool tile_nodes_given = tile_nodes > 0;
    if (tile_nodes <= 0) tile_nodes = kDefaultTileNodes;
    HIDENN_REQUIRE(tile_nodes >= 8 && tile_nodes <= 2048, "plan_create: tile_nodes must be in [8,2048]");

    for (int64_t i = 0; i < 3 * Ne; ++i)
        HIDENN_REQUIRE(conn[i] >= 0 && conn[i] < Nn, "plan_create: connectivity index out of range");
    for (int64_t i = 0; i < 2 * Ned; ++i)
        HIDENN_REQUIRE(edges[i] >= 0 && edges[i] < Nn, "plan_create: edge node index out of range");

    std::unique_ptr<hidenn_tri_plan> p(new hidenn_tri_plan());
    p->device = device;
    p->real_bytes = real_bytes;
    p->n_elems = Ne;
    p->n_nodes = Nn;

    // slot maps (src/models.py:261-262, 274: free rows in ascending node order)
    p->xslot.resize(Nn);
    p->uslot.resize(Nn);
    {
        int32_t fx = 0, bx = 0, fu = 0, bu = 0;
        for (int64_t n = 0; n < Nn; ++n) {
            p->xslot[n] = bmask[n] ? ~(bx++) : fx++;
            p->uslot[n] = dmask[n] ? ~(bu++) : fu++;
        }
        p->n_free_x = fx; p->n_fixed_x = bx; p->n_free_u = fu; p->n_fixed_u = bu;
    }
    p->conn32.resize(3 * Ne);
    for (int64_t i = 0; i < 3 * Ne; ++i) p->conn32[i] = (int32_t)conn[i];

    // global node -> (element, corner) CSR, ascending element id inside each node
    p->n2e_off.assign(Nn + 1, 0);
    for (int64_t i = 0; i < 3 * Ne; ++i) p->n2e_off[p->conn32[i] + 1]++;
    for (int64_t n = 0; n < Nn; ++n) p->n2e_off[n + 1] += p->n2e_off[n];
    p->n2e_ent.resize(3 * Ne);
    {
        std::vector<int64_t> cur(p->n2e_off.begin(), p->n2e_off.end() - 1);
        for (int64_t e = 0; e < Ne; ++e)
            for (int c = 0; c < 3; ++c) p->n2e_ent[cur[p->conn32[3 * e + c]]++] = (int32_t)(e * 4 + c);
    }
    for (int64_t n = 0; n < Nn; ++n)
        HIDENN_REQUIRE(p->n2e_off[n + 1] - p->n2e_off[n] < kMaxValence, "plan_create: node valence >= 255 not supported");

    // Neumann edge ends by node: in the tile-ordered layout the tile that owns an edge node also folds the edge term into
    // its gradient; the other end of such an edge is a local node of the tile in every layout
    EdgeEnds edge_ends;
    edge_ends.build(edges, Ned);
    auto ends_of = [&](int32_t n) { return edge_ends.of(n); };
    const bool no_v8 = getenv("HIDENN_PLAN_NO_V8") != nullptr;      // A/B: treat a tile-ordered mesh like any other
    bool tile_order = false, force_generic = false;
    // Global matching of the elements into edge-sharing pairs (paired layout of kernel v9; match_elements above).  It
    // depends only on the mesh, so hidenn_tri_locality_order computes the same one and sizes the tiles by it.
    const int32_t* c32g = p->conn32.data();
    auto elem_has = [c32g](int32_t f, int32_t n) { return c32g[3 * (int64_t)f] == n || c32g[3 * (int64_t)f + 1] == n || c32g[3 * (int64_t)f + 2] == n; };
    int n_classes = 0;
    if (real_bytes == 8 && !no_v8 && Ne > 0) p->n_pairs = match_elements(c32g, Ne, p->n2e_off.data(), p->n2e_ent.data(), p->mate, &n_classes);
    bool pairs_on = !p->mate.empty();
    if (pairs_on && !tile_nodes_given) tile_nodes = pairs_tile_nodes(Ne, p->n_pairs, n_classes);

    // RCB tiling of the nodes
    int64_t n_tiles = 0;
    std::vector<int32_t> order(Nn);
    std::vector<int64_t> tile_begin;
    std::vector<TileBuild> tb;
    for (int attempt = 0;; ++attempt) {
    n_tiles = (Nn + tile_nodes - 1) / tile_nodes;
    std::iota(order.begin(), order.end(), 0);
    tile_begin.assign(n_tiles + 1, 0);
    tile_begin[n_tiles] = Nn;
    {
        Rcb r{coords, order, tile_begin};
        r.run(0, Nn, 0, n_tiles, 0);
    }

    // Tile-ordered numbering?  Every leaf is one contiguous id range and lists its nodes by class (tri_plan.h).
    tile_order = (real_bytes == 8) && !no_v8 && !force_generic;
    for (int64_t t = 0; t < n_tiles && tile_order; ++t) {
        int32_t mn = INT32_MAX, mx = INT32_MIN;
        for (int64_t i = tile_begin[t]; i < tile_begin[t + 1]; ++i) { mn = std::min(mn, order[i]); mx = std::max(mx, order[i]); }
        if ((int64_t)mx - mn + 1 != tile_begin[t + 1] - tile_begin[t]) { tile_order = false; break; }
        for (int32_t n = mn; n < mx; ++n)
            if (node_class(bmask[n], dmask[n]) > node_class(bmask[n + 1], dmask[n + 1])) { tile_order = false; break; }
    }

    if (pairs_on && !tile_order) {
        // the paired layout only serves the tile-ordered kernel: without such a numbering drop it and size the tiles for
        // one element per entry
        p->mate.clear();
        p->n_pairs = 0;
        pairs_on = false;
        if (!tile_nodes_given) tile_nodes = kDefaultTileNodes;
        continue;
    }

    // per-tile packs, in parallel over tiles
    tb.assign(n_tiles, TileBuild());
    const int32_t* c32 = p->conn32.data();
    const int64_t* n2o = p->n2e_off.data();
    const int32_t* n2e = p->n2e_ent.data();
    auto build_range = [&](int64_t t0, int64_t t1) {
        std::vector<int32_t> cand, halo, owned_sorted, perm, lid_of, halo_lid, pool_of;
        std::vector<std::pair<int32_t, int32_t>> pools, cs;
        std::vector<std::vector<int32_t>> free_res;
        const int G = 8;
        for (int64_t t = t0; t < t1; ++t) {
            TileBuild& B = tb[t];
            // owned nodes: looked up through an id-sorted list, numbered locally by descending valence so that
            // every group of G consecutive local nodes has (nearly) equal valence -- see the slot layout below
            owned_sorted.assign(order.begin() + tile_begin[t], order.begin() + tile_begin[t + 1]);
            std::sort(owned_sorted.begin(), owned_sorted.end());
            B.n_owned = (int32_t)owned_sorted.size();
            perm.resize(B.n_owned);
            std::iota(perm.begin(), perm.end(), 0);
            std::stable_sort(perm.begin(), perm.end(), [&](int32_t a, int32_t b) {
                const int64_t va = n2o[owned_sorted[a] + 1] - n2o[owned_sorted[a]], vb = n2o[owned_sorted[b] + 1] - n2o[owned_sorted[b]];
                return va > vb;
            });
            lid_of.resize(B.n_owned);
            B.nodes.resize(B.n_owned);
            for (int32_t l = 0; l < B.n_owned; ++l) { lid_of[perm[l]] = l; B.nodes[l] = owned_sorted[perm[l]]; }
            cand.clear();
            for (int32_t n : owned_sorted)
                for (int64_t k = n2o[n]; k < n2o[n + 1]; ++k) cand.push_back(n2e[k] >> 2);
            std::sort(cand.begin(), cand.end());
            cand.erase(std::unique(cand.begin(), cand.end()), cand.end());
            B.elems = cand;
            halo.clear();
            auto owned_id = [&](int32_t n) -> int32_t {
                auto it = std::lower_bound(owned_sorted.begin(), owned_sorted.end(), n);
                return (it != owned_sorted.end() && *it == n) ? lid_of[it - owned_sorted.begin()] : -1;
            };
            for (int32_t e : B.elems)
                for (int c = 0; c < 3; ++c) {
                    int32_t n = c32[3 * (int64_t)e + c];
                    if (owned_id(n) < 0) halo.push_back(n);
                }
            if (Ned > 0)
                for (int32_t n : owned_sorted) {
                    auto r = ends_of(n);
                    for (auto it = r.first; it != r.second; ++it) {
                        const int32_t other = (int32_t)edges[it->second ^ 1];
                        if (owned_id(other) < 0) halo.push_back(other);
                    }
                }
            std::sort(halo.begin(), halo.end());
            halo.erase(std::unique(halo.begin(), halo.end()), halo.end());
            const int32_t n_halo = (int32_t)halo.size(), n_local = B.n_owned + n_halo;
            if (n_local > kMaxLocal) { B.err = 1; continue; }
            {       // numbering-independent slot bound (same test as hidenn_tri_locality_order)
                cs.clear();
                for (int32_t n : owned_sorted) {
                    auto r = ends_of(n);
                    const int32_t es = pairs_on ? paired_elem_slots(n, c32, n2o, n2e, p->mate) : (int32_t)(n2o[n + 1] - n2o[n]);
                    cs.push_back({node_class(bmask[n], dmask[n]), (int32_t)(es + (r.second - r.first))});
                }
                if (canon_padded_entries(cs) > kMaxEntries) { B.err = 2; continue; }
            }
            if (tile_order) {
                // tile-ordered layout: local id = memory order (bulk copies land the owned rows at their ids, fold
                // thread l stores row l of the tile's run); halo nodes follow in ascending id
                halo_lid.assign(n_halo, 0);
                B.nodes.assign(n_local, 0);
                for (int32_t l = 0; l < B.n_owned; ++l) { lid_of[l] = l; B.nodes[l] = owned_sorted[l]; }
                for (int32_t j = 0; j < n_halo; ++j) { halo_lid[j] = B.n_owned + j; B.nodes[B.n_owned + j] = halo[j]; }
            } else
            // Local ids within a valence class (and among the halo nodes) are free.  The node records are staged and
            // the gradients flushed in MEMORY order, 8 consecutive records per 128-byte shared-memory pass, each
            // landing at / read from its local id: choose the ids so that the 8 records of a pass have 8 different
            // ids mod 8 (different bank groups) wherever the class still has such an id free.
            {
                pools.clear();
                pool_of.assign(B.n_owned, 0);
                for (int32_t l = 0; l < B.n_owned;) {             // classes = runs of equal valence in the sorted order
                    const int64_t v = n2o[owned_sorted[perm[l]] + 1] - n2o[owned_sorted[perm[l]]];
                    int32_t e = l;
                    while (e < B.n_owned && n2o[owned_sorted[perm[e]] + 1] - n2o[owned_sorted[perm[e]]] == v) { pool_of[perm[e]] = (int32_t)pools.size(); ++e; }
                    pools.push_back({l, e});
                    l = e;
                }
                pools.push_back({B.n_owned, n_local});              // halo pool
                free_res.assign(pools.size() * 8, std::vector<int32_t>());
                for (size_t q = 0; q < pools.size(); ++q)
                    for (int32_t l = pools[q].second - 1; l >= pools[q].first; --l) free_res[q * 8 + (l & 7)].push_back(l);   // pop_back = lowest id
                halo_lid.assign(n_halo, 0);
                B.nodes.assign(n_local, 0);
                int32_t io = 0, ih = 0, j = 0;
                unsigned used = 0;
                while (io < B.n_owned || ih < n_halo) {
                    const bool take_owned = ih >= n_halo || (io < B.n_owned && owned_sorted[io] < halo[ih]);
                    const size_t q = take_owned ? (size_t)pool_of[io] : pools.size() - 1;
                    if ((j & 7) == 0) used = 0;
                    int best = -1;
                    size_t best_free = 0;
                    for (int r = 0; r < 8; ++r) {                   // unused residue with the most ids left in this class
                        const size_t nf = free_res[q * 8 + r].size();
                        if (nf > best_free && !((used >> r) & 1u)) { best = r; best_free = nf; }
                    }
                    if (best < 0)                                   // class exhausted for the free residues: accept a conflict
                        for (int r = 0; r < 8; ++r)
                            if (free_res[q * 8 + r].size() > best_free) { best = r; best_free = free_res[q * 8 + r].size(); }
                    const int32_t l = free_res[q * 8 + best].back();
                    free_res[q * 8 + best].pop_back();
                    used |= 1u << best;
                    if (take_owned) { lid_of[io] = l; B.nodes[l] = owned_sorted[io]; ++io; }
                    else { halo_lid[ih] = l; B.nodes[l] = halo[ih]; ++ih; }
                    ++j;
                }
            }
            // Fold-slot layout: local nodes in groups of G (the lanes one shared-memory pass serves: 8 x 16 B for FP64,
            // 16 x 8 B for FP32).  Slot k of node l lives at  base[group] + k*G + (l % G): the G lanes that fold a group
            // read one contiguous 128-byte row per step (conflict-free), and the bank group of any partial store is
            // l % G -- the same as for the gather of that node, so the lane assignment below fixes both at once.
            B.off.resize(B.n_owned);
            int64_t acc = 0;
            // slots of a node: one per incident element, in the tile-ordered layout followed by one per Neumann edge end
            auto slots_of = [&](int32_t n) -> int64_t {
                int64_t c = n2o[n + 1] - n2o[n];
                if (tile_order && Ned > 0) { auto r = ends_of(n); c += r.second - r.first; }
                return c;
            };
            for (int32_t g0 = 0; g0 < B.n_owned; g0 += G) {
                int64_t mx = 0;
                for (int32_t l = g0; l < std::min(B.n_owned, g0 + G); ++l) mx = std::max<int64_t>(mx, slots_of(B.nodes[l]));
                for (int32_t l = g0; l < std::min(B.n_owned, g0 + G); ++l) {
                    const int64_t cnt = slots_of(B.nodes[l]);
                    B.off[l] = (uint32_t)(acc + (l - g0)) | ((uint32_t)cnt << 16);
                }
                acc += mx * G;
                if (acc > kMaxEntries) {
                    // a plan with a paired layout sizes its tiles for that layout: the one-element-per-entry packs of this
                    // tile do not fit their 11-bit position fields and the plan is marked pairs-only (kernel v9 only);
                    // otherwise only a tile-ordered numbering that is not slot-sorted can get here
                    if (pairs_on && tile_order) B.unpaired_overflow = true;
                    else { B.err = 3; break; }
                }
            }
            if (B.err) continue;
            if (B.unpaired_overflow) acc = kMaxEntries;      // positions below are masked garbage, never used
            B.n_entries = (int32_t)acc;
            B.pack.resize(B.elems.size());
            for (size_t i = 0; i < B.elems.size(); ++i) {
                const int32_t e = B.elems[i];
                unsigned long long w = 0;
                for (int c = 0; c < 3; ++c) {
                    const int32_t n = c32[3 * (int64_t)e + c];
                    int32_t lid = owned_id(n);
                    unsigned long long pos = (unsigned long long)acc;     // dump slot for halo corners
                    if (lid >= 0) {
                        const int32_t key = e * 4 + c;
                        for (int64_t k = n2o[n]; k < n2o[n + 1]; ++k)
                            if (n2e[k] == key) { pos = (unsigned long long)((B.off[lid] & 0xFFFFu) + (k - n2o[n]) * G); break; }
                    } else {
                        lid = halo_lid[std::lower_bound(halo.begin(), halo.end(), n) - halo.begin()];
                    }
                    w |= (unsigned long long)lid << (kLidBits * c);
                    w |= (pos & (unsigned long long)kMaxEntries) << (3 * kLidBits + kPosBits * c);
                }
                if (owned_id(c32[3 * (int64_t)e]) >= 0) w |= 1ull << kOwnerBit;
                B.pack[i] = w;
            }
            // (a plan with the paired layout runs kernel v9 on the pair entries: its one-element packs serve only the
            // energy-only kernel and the fall-backs, and are left in element order)
            if (!(pairs_on && tile_order)) reorder_for_banks(B, real_bytes);
            if (tile_order && Ned > 0) {
                // Neumann edge visits: every edge with an owned end; the partial of an owned end goes to the slot after
                // the node's element slots (rank among the node's edge ends), a halo end to the dump slot
                std::vector<int32_t> ev;
                for (int32_t n : owned_sorted) {
                    auto r = ends_of(n);
                    for (auto it = r.first; it != r.second; ++it) ev.push_back(it->second >> 1);
                }
                std::sort(ev.begin(), ev.end());
                ev.erase(std::unique(ev.begin(), ev.end()), ev.end());
                for (int32_t e : ev) {
                    unsigned long long w = 0;
                    for (int k = 0; k < 2; ++k) {
                        const int32_t n = (int32_t)edges[2 * (int64_t)e + k];
                        int32_t lid = owned_id(n);
                        unsigned long long pos = (unsigned long long)acc;
                        if (lid >= 0) {
                            auto r = ends_of(n);
                            int64_t rank = 0;
                            for (auto it = r.first; it != r.second; ++it, ++rank)
                                if (it->second == 2 * e + k) break;
                            pos = (unsigned long long)((B.off[lid] & 0xFFFFu) + (n2o[n + 1] - n2o[n] + rank) * G);
                            if (k == 0) w |= 1ull << kOwnerBit;
                        } else {
                            lid = halo_lid[std::lower_bound(halo.begin(), halo.end(), n) - halo.begin()];
                        }
                        w |= (unsigned long long)lid << (kLidBits * k);
                        w |= (pos & (unsigned long long)kMaxEntries) << (2 * kLidBits + kPosBits * k);
                    }
                    B.epack.push_back(w);
                    B.eid.push_back(e);
                }
            }
            if (tile_order && !p->mate.empty()) {
                // ---- paired layout (kernel v9, opt-in) ----
                const std::vector<int32_t>& mate = p->mate;
                // the partial of element e at node n is merged away when e's partner is the smaller id and holds n too
                auto merged_away = [&](int32_t e, int32_t n) { const int32_t f = mate[e]; return f >= 0 && f < e && elem_has(f, n); };
                auto elem_slots9 = [&](int32_t n) -> int64_t {
                    int64_t c = 0;
                    for (int64_t k = n2o[n]; k < n2o[n + 1]; ++k) c += merged_away(n2e[k] >> 2, n) ? 0 : 1;
                    return c;
                };
                auto slots9 = [&](int32_t n) -> int64_t {
                    int64_t c = elem_slots9(n);
                    if (Ned > 0) { auto r = ends_of(n); c += r.second - r.first; }
                    return c;
                };
                B.off9.resize(B.n_owned);
                int64_t acc9 = 0;
                for (int32_t g0 = 0; g0 < B.n_owned; g0 += G) {
                    int64_t mx = 0;
                    for (int32_t l = g0; l < std::min(B.n_owned, g0 + G); ++l) mx = std::max<int64_t>(mx, slots9(B.nodes[l]));
                    for (int32_t l = g0; l < std::min(B.n_owned, g0 + G); ++l)
                        B.off9[l] = (uint32_t)(acc9 + (l - g0)) | ((uint32_t)slots9(B.nodes[l]) << 16);
                    acc9 += mx * G;
                }
                if (acc9 > kMaxEntries) { B.err = 2; continue; }      // the bound above is by class and count, the layout by id
                B.n_entries9 = (int32_t)acc9;
                auto word9 = [&](int32_t e) {
                    unsigned long long w = 0;
                    for (int c = 0; c < 3; ++c) {
                        const int32_t n = c32[3 * (int64_t)e + c];
                        int32_t lid = owned_id(n);
                        unsigned long long pos = (unsigned long long)acc9;
                        if (lid >= 0) {
                            if (!merged_away(e, n)) {
                                int64_t rank = 0;
                                for (int64_t k = n2o[n]; k < n2o[n + 1]; ++k) {
                                    if (n2e[k] == e * 4 + c) break;
                                    rank += merged_away(n2e[k] >> 2, n) ? 0 : 1;
                                }
                                pos = (unsigned long long)((B.off9[lid] & 0xFFFFu) + rank * G);
                            }
                        } else {
                            lid = halo_lid[std::lower_bound(halo.begin(), halo.end(), n) - halo.begin()];
                        }
                        w |= (unsigned long long)lid << (kLidBits * c);
                        w |= pos << (3 * kLidBits + kPosBits * c);
                    }
                    if (owned_id(c32[3 * (int64_t)e]) >= 0) w |= 1ull << kOwnerBit;
                    return w;
                };
                // B.elems is in lane order after reorder_for_banks: membership through a sorted copy
                std::vector<int32_t> visited(B.elems);
                std::sort(visited.begin(), visited.end());
                auto is_visited = [&](int32_t f) { return f >= 0 && std::binary_search(visited.begin(), visited.end(), f); };
                const unsigned long long null_word = kNullPack | ((unsigned long long)acc9 << (3 * kLidBits)) |
                                                     ((unsigned long long)acc9 << (3 * kLidBits + kPosBits)) |
                                                     ((unsigned long long)acc9 << (3 * kLidBits + 2 * kPosBits));
                std::vector<unsigned long long> singles;
                for (int32_t e : visited) {
                    const int32_t f = mate[e];
                    if (is_visited(f)) {
                        if (f < e) continue;                       // listed with its partner
                        B.pack9.push_back(word9(e));
                        B.pack9.push_back(word9(f));
                    } else {
                        singles.push_back(word9(e));
                    }
                }
                // entries are listed class by class (a warp whose lanes are all of one class runs one wiring of the second
                // element; a warp across a class boundary just runs both), bank-aware inside each class; singles at the end
                {
                    std::vector<unsigned long long> by_cls[9];
                    for (size_t i = 0; i + 1 < B.pack9.size(); i += 2) {
                        const int cls = pair_class(B.pack9[i], B.pack9[i + 1]);
                        if (cls < 0 || cls >= 9) { B.err = 3; break; }
                        by_cls[cls].push_back(B.pack9[i]);
                        by_cls[cls].push_back(B.pack9[i + 1]);
                    }
                    B.pack9.clear();
                    // every class starts at a multiple of 32 entries (padding = skip entries): a warp that ran two wirings
                    // one after the other would take twice as long, and the tile waits for its slowest warp
                    static const bool pad_cls = [] { const char* e = getenv("HIDENN_PLAN_PAIR_PAD"); return !e || atoi(e) != 0; }();
                    for (int c = 0; c < 9; ++c) {
                        reorder_pairs_for_banks(by_cls[c], (unsigned)acc9);
                        B.pack9.insert(B.pack9.end(), by_cls[c].begin(), by_cls[c].end());
                        while (pad_cls && (B.pack9.size() / 2) % 32 != 0) { B.pack9.push_back(null_word); B.pack9.push_back(null_word); }
                    }
                }
                {
                    std::vector<unsigned long long> sp;
                    for (unsigned long long w : singles) { sp.push_back(w); sp.push_back(null_word); }
                    reorder_pairs_for_banks(sp, (unsigned)acc9);
                    B.pack9.insert(B.pack9.end(), sp.begin(), sp.end());
                }
                for (size_t i = 0; i < B.epack.size(); ++i) {      // edge visits: same ends, positions of the paired layout
                    const int32_t e = B.eid[i];
                    unsigned long long w = B.epack[i] & ((1ull << (2 * kLidBits)) - 1ull);
                    w |= B.epack[i] & (1ull << kOwnerBit);
                    for (int k = 0; k < 2; ++k) {
                        const int32_t n = (int32_t)edges[2 * (int64_t)e + k];
                        const int32_t lid = owned_id(n);
                        unsigned long long pos = (unsigned long long)acc9;
                        if (lid >= 0) {
                            auto r = ends_of(n);
                            int64_t rank = 0;
                            for (auto it = r.first; it != r.second; ++it, ++rank)
                                if (it->second == 2 * e + k) break;
                            pos = (unsigned long long)((B.off9[lid] & 0xFFFFu) + (elem_slots9(n) + rank) * G);
                        }
                        w |= pos << (2 * kLidBits + kPosBits * k);
                    }
                    B.epack9.push_back(w);
                }
            }
        }
    };
    {
        unsigned nthr = std::max(1u, std::min(16u, std::thread::hardware_concurrency()));
        if (n_tiles < 64) nthr = 1;
        std::vector<std::thread> th;
        const int64_t per = (n_tiles + nthr - 1) / nthr;
        for (unsigned i = 0; i < nthr; ++i) {
            const int64_t a = i * per, b = std::min<int64_t>(n_tiles, a + per);
            if (a < b) th.emplace_back(build_range, a, b);
        }
        for (auto& t : th) t.join();
    }
    bool bad = false, bad_layout = false;
    for (auto& B : tb) { bad |= (B.err == 1 || B.err == 2); bad_layout |= (B.err == 3); }
    if (!bad && bad_layout) {       // keep the tiling, fall back to the generic layout (its slot count is below the bound)
        HIDENN_REQUIRE(!force_generic, "plan_create: internal error (generic layout exceeds the slot bound)");
        force_generic = true;
        continue;
    }
    if (!bad) break;
    // a tile exceeded the pack limits (1023 local nodes / 2047 fold slots): shrink the tiles and retry
    HIDENN_REQUIRE(attempt < 12 && tile_nodes > 8, "plan_create: cannot tile this mesh within the pack limits");
    tile_nodes = std::max(8, tile_nodes * 3 / 4);
    }
    int32_t max_local = 0, max_entries = 0, max_owned = 0, max_elem = 0;
    int64_t node_visits = 0, elem_visits = 0, off_total = 0;
    bool any_unpaired_overflow = false;
    for (auto& B : tb) any_unpaired_overflow |= B.unpaired_overflow;
    for (auto& B : tb) {
        node_visits += (int64_t)B.nodes.size();
        elem_visits += (int64_t)B.elems.size();
        off_total += B.n_owned;
    }
    HIDENN_REQUIRE(node_visits < 2147483000LL && elem_visits < 2147483000LL, "plan_create: mesh too large for int32 tile offsets");
    p->tiles.resize(n_tiles);
    p->t_node.reserve(node_visits);
    p->t_elem.reserve(elem_visits);
    p->elem_pack.reserve(elem_visits);
    p->entry_off.reserve(off_total);
    // tiles are laid out by ascending smallest owned node id (= ascending Parameter rows): neighbouring CTAs then work
    // on neighbouring memory, and the host-buffer entry point can stream rows in and gradients out chunk by chunk
    std::vector<int64_t> tord(n_tiles);
    std::iota(tord.begin(), tord.end(), 0);
    {
        std::vector<int32_t> key(n_tiles, INT32_MAX);
        for (int64_t t = 0; t < n_tiles; ++t)
            for (int32_t l = 0; l < tb[t].n_owned; ++l) key[t] = std::min(key[t], tb[t].nodes[l]);
        // tiles that own one of the caller's `first_nodes` (multi-GPU: nodes shared with another rank) are listed first,
        // so that the launch can be split into [0, n_first_tiles) and the rest and the halo exchange overlaps the second part
        std::vector<char> prio(n_tiles, 0);
        if (n_first > 0) {
            std::vector<char> mark(Nn, 0);
            for (int64_t i = 0; i < n_first; ++i) mark[first_nodes[i]] = 1;
            for (int64_t t = 0; t < n_tiles; ++t)
                for (int32_t l = 0; l < tb[t].n_owned && !prio[t]; ++l) prio[t] = mark[tb[t].nodes[l]];
            for (int64_t t = 0; t < n_tiles; ++t) p->n_first_tiles += prio[t];
        }
        if (!getenv("HIDENN_PLAN_RCB_ORDER"))      // debug: keep the recursive-bisection order
            std::stable_sort(tord.begin(), tord.end(), [&](int64_t a, int64_t b) {
                return prio[a] != prio[b] ? prio[a] > prio[b] : key[a] < key[b];
            });
    }
    for (int64_t t = 0; t < n_tiles; ++t) {
        TileBuild& B = tb[tord[t]];
        TileDesc d{};
        d.node_off = (int32_t)p->t_node.size();
        d.n_owned = B.n_owned;
        d.n_local = (int32_t)B.nodes.size();
        d.elem_off = (int32_t)p->t_elem.size();
        d.n_elem = (int32_t)B.elems.size();
        d.off_off = (int32_t)p->entry_off.size();
        d.n_entries = B.n_entries;
        p->tiles[t] = d;
        p->t_node.insert(p->t_node.end(), B.nodes.begin(), B.nodes.end());
        p->t_elem.insert(p->t_elem.end(), B.elems.begin(), B.elems.end());
        p->elem_pack.insert(p->elem_pack.end(), B.pack.begin(), B.pack.end());
        p->entry_off.insert(p->entry_off.end(), B.off.begin(), B.off.end());
        if (tile_order) {
            TileDesc8 d8{};
            d8.n_owned = d.n_owned; d8.n_local = d.n_local; d8.n_elem = d.n_elem; d8.n_entries = d.n_entries;
            const int32_t first = B.nodes[0];
            int32_t cnt[4] = {0, 0, 0, 0};
            for (int32_t l = 0; l < d.n_owned; ++l) cnt[node_class(bmask[first + l], dmask[first + l])]++;
            d8.nA = cnt[0]; d8.nB = cnt[1]; d8.nC = cnt[2]; d8.nD = cnt[3];
            // first rows: the class segments are contiguous in each array (free rows ascend with the node id)
            auto first_row = [&](const std::vector<int32_t>& slot, bool want_free, int32_t l0, int32_t l1) -> int32_t {
                for (int32_t l = l0; l < l1; ++l) {
                    const int32_t sl = slot[first + l];
                    if (want_free == (sl >= 0)) return sl >= 0 ? sl : ~sl;
                }
                return 0;
            };
            d8.rx_free = first_row(p->xslot, true, 0, d.n_owned);
            d8.rx_fixed = first_row(p->xslot, false, 0, d.n_owned);
            d8.ru_free = first_row(p->uslot, true, 0, d.n_owned);
            d8.ru_fixed = first_row(p->uslot, false, 0, d.n_owned);
            d8.edge_off = (int32_t)p->edge_pack.size();
            d8.n_edge = (int32_t)B.epack.size();
            d8.n_pent = (int32_t)(B.pack9.size() / 2);
            d8.n_entries9 = B.n_entries9;
            p->pair_pack.insert(p->pair_pack.end(), B.pack9.begin(), B.pack9.end());
            p->entry_off9.insert(p->entry_off9.end(), B.off9.begin(), B.off9.end());
            p->edge_pack9.insert(p->edge_pack9.end(), B.epack9.begin(), B.epack9.end());
            p->pair_entries += d8.n_pent;
            p->edge_pack.insert(p->edge_pack.end(), B.epack.begin(), B.epack.end());
            p->edge_id.insert(p->edge_id.end(), B.eid.begin(), B.eid.end());
            p->tiles8.push_back(d8);
        }
        max_local = std::max(max_local, d.n_local);
        max_entries = std::max(max_entries, d.n_entries);
        max_owned = std::max(max_owned, d.n_owned);
        max_elem = std::max(max_elem, d.n_elem);
        std::vector<int32_t>().swap(B.nodes);
        std::vector<int32_t>().swap(B.elems);
        std::vector<unsigned long long>().swap(B.pack);
    }
    p->node_visits = node_visits;
    p->elem_visits = elem_visits;
    p->tile_order = tile_order;
    p->unpaired_ok = !any_unpaired_overflow;

    // device layout: fixed-stride records per tile
    const int32_t SL = (max_local + 1) & ~1, SE = (max_elem + 1) & ~1, SO = (max_owned + 3) & ~3;
    // padding slots must stay loadable: row 0 of the fixed buffer if it exists, else row 0 of the free Parameter
    std::vector<int2> t_slots((size_t)n_tiles * SL, make_int2(p->n_fixed_x > 0 ? -1 : 0, p->n_fixed_u > 0 ? -1 : 0));
    std::vector<uint16_t> t_lid((size_t)n_tiles * SL, (uint16_t)0xFFFF);
    std::vector<unsigned long long> d_pack((size_t)n_tiles * SE, 0ull);
    std::vector<uint32_t> d_off((size_t)n_tiles * SO, 0u);
    {
        std::vector<std::pair<int32_t, int32_t>> byid;
        for (int64_t t = 0; t < n_tiles; ++t) {
            const TileDesc& d = p->tiles[t];
            byid.resize(d.n_local);
            for (int32_t i = 0; i < d.n_local; ++i) byid[i] = {p->t_node[d.node_off + i], i};
            std::sort(byid.begin(), byid.end());                 // memory order
            for (int32_t j = 0; j < d.n_local; ++j) {
                const int32_t n = byid[j].first;
                t_slots[(size_t)t * SL + j] = make_int2(p->xslot[n], p->uslot[n]);
                t_lid[(size_t)t * SL + j] = (uint16_t)byid[j].second;
            }
        }
        {       // locality of the numbering: contiguous Parameter-row runs per tile (1-2 for a tile-ordered numbering,
                // ~n_local for a random one); hidenn_tri_plan_locality, the Python plan warns when it is poor
            int64_t runs = 0, nl = 0;
            for (int64_t t = 0; t < n_tiles; ++t) {
                const TileDesc& d = p->tiles[t];
                int prev = INT32_MIN;
                for (int32_t j = 0; j < d.n_local; ++j) {
                    const int c = t_slots[(size_t)t * SL + j].x;
                    if (!(c == prev + 1 && c > 0) && !(c < 0 && c == prev - 1)) ++runs;
                    prev = c;
                }
                nl += d.n_local;
            }
            p->runs_per_tile = n_tiles ? (double)runs / n_tiles : 0.0;
            p->local_per_tile = n_tiles ? (double)nl / n_tiles : 0.0;
        }
        if (getenv("HIDENN_PLAN_RUNSTATS")) {       // debug: contiguous-row runs per tile (bulk-copy feasibility)
            int64_t rx = 0, ru = 0, ro = 0, nl = 0, no = 0;
            for (int64_t t = 0; t < n_tiles; ++t) {
                const TileDesc& d = p->tiles[t];
                int2 prev = make_int2(INT32_MIN, INT32_MIN);
                int prev_ox = INT32_MIN, prev_ou = INT32_MIN;
                for (int32_t j = 0; j < d.n_local; ++j) {
                    const int2 c = t_slots[(size_t)t * SL + j];
                    if (!(c.x == prev.x + 1 && c.x > 0) && !(c.x < 0 && c.x == prev.x - 1)) ++rx;
                    if (!(c.y == prev.y + 1 && c.y > 0) && !(c.y < 0 && c.y == prev.y - 1)) ++ru;
                    prev = c;
                    if (t_lid[(size_t)t * SL + j] < d.n_owned) {
                        if (c.x >= 0) { if (c.x != prev_ox + 1) ++ro; prev_ox = c.x; }
                        if (c.y >= 0) { if (c.y != prev_ou + 1) ++ro; prev_ou = c.y; }
                        ++no;
                    }
                }
                nl += d.n_local;
            }
            fprintf(stderr, "[plan runstats] tiles %lld  local/tile %.1f owned/tile %.1f  x-runs/tile %.1f u-runs/tile %.1f out-runs/tile %.1f\n",
                    (long long)n_tiles, (double)nl / n_tiles, (double)no / n_tiles, (double)rx / n_tiles, (double)ru / n_tiles, (double)ro / n_tiles);
        }
    }
    for (int64_t t = 0; t < n_tiles; ++t) {
        const TileDesc& d = p->tiles[t];
        std::copy(p->elem_pack.begin() + d.elem_off, p->elem_pack.begin() + d.elem_off + d.n_elem, d_pack.begin() + (size_t)t * SE);
        std::copy(p->entry_off.begin() + d.off_off, p->entry_off.begin() + d.off_off + d.n_owned, d_off.begin() + (size_t)t * SO);
    }

    // block tables of the host-buffer pipeline (see tri_plan.h)
    {
        auto rows_per_block = [](int64_t n) { return (int32_t)std::max<int64_t>(16, ((n + kPipeBlocks - 1) / kPipeBlocks + 15) / 16 * 16); };
        p->pipe_rows_x = rows_per_block(p->n_free_x);
        p->pipe_rows_u = rows_per_block(p->n_free_u);
        p->first_need_x.assign(kPipeBlocks, INT32_MAX); p->first_need_u.assign(kPipeBlocks, INT32_MAX);
        p->last_own_x.assign(kPipeBlocks, -1); p->last_own_u.assign(kPipeBlocks, -1);
        for (int64_t t = 0; t < n_tiles; ++t) {
            const TileDesc& d = p->tiles[t];
            for (int32_t i = 0; i < d.n_local; ++i) {
                const int32_t n = p->t_node[d.node_off + i];
                const int32_t xs = p->xslot[n], us = p->uslot[n];
                if (xs >= 0) {
                    const int32_t b = xs / p->pipe_rows_x;
                    p->first_need_x[b] = std::min(p->first_need_x[b], (int32_t)t);
                    if (i < d.n_owned) p->last_own_x[b] = (int32_t)t;
                }
                if (us >= 0) {
                    const int32_t b = us / p->pipe_rows_u;
                    p->first_need_u[b] = std::min(p->first_need_u[b], (int32_t)t);
                    if (i < d.n_owned) p->last_own_u[b] = (int32_t)t;
                }
            }
        }
    }

    // Neumann edges: slot quads + node-centric CSR (each edge node folded by one thread)
    p->edges32.resize(2 * Ned);
    std::vector<int32_t> e_slots(4 * Ned);
    for (int64_t e = 0; e < Ned; ++e) {
        const int32_t a = (int32_t)edges[2 * e], b = (int32_t)edges[2 * e + 1];
        p->edges32[2 * e] = a; p->edges32[2 * e + 1] = b;
        e_slots[4 * e + 0] = p->xslot[a]; e_slots[4 * e + 1] = p->uslot[a];
        e_slots[4 * e + 2] = p->xslot[b]; e_slots[4 * e + 3] = p->uslot[b];
    }
    std::vector<int32_t> enodes(p->edges32);
    std::sort(enodes.begin(), enodes.end());
    enodes.erase(std::unique(enodes.begin(), enodes.end()), enodes.end());
    const int32_t n_en = (int32_t)enodes.size();
    std::vector<int32_t> en_xslot(n_en), en_uslot(n_en), en_off(n_en + 1, 0), en_ent(2 * Ned);
    for (int32_t k = 0; k < n_en; ++k) { en_xslot[k] = p->xslot[enodes[k]]; en_uslot[k] = p->uslot[enodes[k]]; }
    auto en_id = [&](int32_t n) { return (int32_t)(std::lower_bound(enodes.begin(), enodes.end(), n) - enodes.begin()); };
    for (int64_t i = 0; i < 2 * Ned; ++i) en_off[en_id(p->edges32[i]) + 1]++;
    for (int32_t k = 0; k < n_en; ++k) en_off[k + 1] += en_off[k];
    {
        std::vector<int32_t> cur(en_off.begin(), en_off.end() - 1);
        for (int64_t i = 0; i < 2 * Ned; ++i) en_ent[cur[en_id(p->edges32[i])]++] = (int32_t)i;   // edge*2+end, ascending
    }

    p->en_xslot_h = en_xslot; p->en_uslot_h = en_uslot;
    // upload (device == -1: host-only plan for index tests, no compute possible)
    TriPlanDev& D0 = p->dev;
    D0.n_tiles = (int32_t)n_tiles;
    D0.max_local = max_local; D0.max_entries = max_entries; D0.max_owned = max_owned; D0.max_elem = max_elem;
    D0.n_edges = (int32_t)Ned; D0.n_enodes = n_en;
    if (device == -1) {
        *out = p.release();
        return 0;
    }
    int ndev = 0;
    HIDENN_CUDA_OK(cudaGetDeviceCount(&ndev));
    HIDENN_REQUIRE(device >= 0 && device < ndev, "plan_create: no such CUDA device (this library has no CPU fallback)");
    DeviceScope scope;
    HIDENN_CUDA_OK(scope.enter(device));
    TriPlanDev& D = p->dev;
    D.n_tiles = (int32_t)n_tiles;
    D.max_local = max_local; D.max_entries = max_entries; D.max_owned = max_owned; D.max_elem = max_elem;
    D.n_edges = (int32_t)Ned; D.n_enodes = n_en;
    D.n_elems = Ne; D.n_nodes = Nn; D.n_free_x = p->n_free_x; D.n_free_u = p->n_free_u;
    D.jinv_t = (options & HIDENN_PLAN_JINV_TRANSPOSE) ? 1 : 0;
    hidenn_tri_plan* pp = p.get();
    int rc = 0;
    rc |= upload(pp, p->tiles, &D.tiles);
    rc |= upload(pp, t_slots, &D.t_slots);
    rc |= upload(pp, t_lid, &D.t_lid);
    rc |= upload(pp, d_pack, &D.elem_pack);
    rc |= upload(pp, d_off, &D.entry_off);
    D.stride_local = SL; D.stride_elem = SE; D.stride_owned = SO;
    if (tile_order) {
        int32_t max_halo = 0;
        for (const TileDesc8& d8 : p->tiles8) max_halo = std::max(max_halo, d8.n_local - d8.n_owned);
        const int32_t SH = std::max(2, (max_halo + 1) & ~1);
        std::vector<int2> t_halo((size_t)n_tiles * SH, make_int2(p->n_fixed_x > 0 ? -1 : 0, p->n_fixed_u > 0 ? -1 : 0));
        for (int64_t t = 0; t < n_tiles; ++t) {
            const TileDesc& d = p->tiles[t];
            for (int32_t j = d.n_owned; j < d.n_local; ++j) {
                const int32_t n = p->t_node[d.node_off + j];
                t_halo[(size_t)t * SH + (j - d.n_owned)] = make_int2(p->xslot[n], p->uslot[n]);
            }
        }
        TriPlan8Dev& D8 = p->dev8;
        rc |= upload(pp, p->tiles8, &D8.tiles);
        rc |= upload(pp, t_halo, &D8.t_halo);
        rc |= upload(pp, p->edge_pack, &D8.edge_pack);
        rc |= upload(pp, p->edge_id, &D8.edge_id);
        D8.stride_halo = SH; D8.max_halo = max_halo; D8.n_edge_visits = (int32_t)p->edge_pack.size();
        // paired layout (opt-in): fixed-stride records like elem_pack / entry_off
        if (!p->mate.empty()) {
        int32_t max_pent = 0, max_entries9 = 0;
        for (const TileDesc8& d8 : p->tiles8) { max_pent = std::max(max_pent, d8.n_pent); max_entries9 = std::max(max_entries9, d8.n_entries9); }
        const int32_t SP = std::max(1, max_pent);
        std::vector<unsigned long long> d_pair((size_t)n_tiles * SP * 2, 0ull);
        std::vector<uint32_t> d_off9((size_t)n_tiles * SO, 0u);
        size_t pp_off = 0, po_off = 0;
        for (int64_t t = 0; t < n_tiles; ++t) {
            const TileDesc8& d8 = p->tiles8[t];
            for (int32_t i = 0; i < d8.n_pent; ++i) {      // device format of the second word (tri_plan.h)
                const unsigned long long w1 = p->pair_pack[pp_off + 2 * (size_t)i], w2 = p->pair_pack[pp_off + 2 * (size_t)i + 1];
                const int cls = pair_class(w1, w2);
                unsigned long long dw = (unsigned long long)cls << (kLidBits + kPosBits);
                if (cls < kPairSingle) {
                    const int r = cls % 3;
                    dw = ((w2 >> (kLidBits * r)) & ((1ull << kLidBits) - 1ull)) |
                         (((w2 >> (3 * kLidBits + kPosBits * r)) & ((1ull << kPosBits) - 1ull)) << kLidBits) |
                         ((unsigned long long)cls << (kLidBits + kPosBits)) | (w2 & (1ull << kOwnerBit));
                }
                d_pair[((size_t)t * SP + i) * 2] = w1;
                d_pair[((size_t)t * SP + i) * 2 + 1] = dw;
            }
            std::copy(p->entry_off9.begin() + po_off, p->entry_off9.begin() + po_off + d8.n_owned, d_off9.begin() + (size_t)t * SO);
            pp_off += 2 * (size_t)d8.n_pent;
            po_off += d8.n_owned;
        }
        rc |= upload(pp, d_pair, &D8.pair_pack);
        rc |= upload(pp, d_off9, &D8.entry_off9);
        rc |= upload(pp, p->edge_pack9, &D8.edge_pack9);
        D8.stride_pent = SP; D8.max_entries9 = max_entries9;
        }
    }
    rc |= upload(pp, e_slots, &D.e_slots);
    rc |= upload(pp, en_xslot, &D.en_xslot);
    rc |= upload(pp, en_uslot, &D.en_uslot);
    rc |= upload(pp, en_off, &D.en_off);
    rc |= upload(pp, en_ent, &D.en_ent);
    if (rc) {
        for (void* q : p->dev_allocs) cudaFree(q);
        return 1;
    }
    *out = p.release();
    return 0;
}

extern "C" void hidenn_tri_plan_destroy(hidenn_tri_plan* p) {
    if (!p) return;
    DeviceScope scope;
    if (p->device >= 0) scope.enter(p->device);
    for (void* q : p->dev_allocs) cudaFree(q);
    if (p->arena) cudaFree(p->arena);
    for (void* e : p->pipe_events) cudaEventDestroy((cudaEvent_t)e);
    for (void* st : p->pipe_streams)
        if (st) cudaStreamDestroy((cudaStream_t)st);
    delete p;
}

static size_t tile_smem_bytes(const hidenn_tri_plan* p, int rb) {
    // node pairs (xy, uv) + fold partial pairs (gu, gx) + entry offsets + block-reduce scratch
    return (size_t)p->dev.max_local * 4 * rb + (size_t)(p->dev.max_entries + 1) * 4 * rb + 128;
}

extern "C" int hidenn_tri_plan_info(const hidenn_tri_plan* p, int64_t* info) {
    HIDENN_REQUIRE(p && info, "plan_info: NULL");
    info[0] = p->dev.n_tiles;
    info[1] = p->elem_visits;
    info[2] = p->node_visits;
    info[3] = p->dev.max_local;
    info[4] = p->dev.max_entries;
    info[5] = (int64_t)p->dev.n_tiles * (p->tile_order ? 2 : 1) + 8 + 280;     // tile energies + finalize partials (64 x 2 doubles) + ticket
    info[6] = (int64_t)tile_smem_bytes(p, 8);
    info[7] = (int64_t)tile_smem_bytes(p, 4);
    info[8] = p->n_free_x;
    info[9] = p->n_free_u;
    info[10] = p->dev.n_edges;
    info[11] = p->dev.n_enodes;
    info[12] = (int64_t)p->dev_bytes;
    info[13] = p->dev.max_elem;
    info[14] = p->n_elems;
    info[15] = p->n_nodes;
    return 0;
}

extern "C" int hidenn_tri_plan_layout(const hidenn_tri_plan* p, int64_t* out8) {
    HIDENN_REQUIRE(p && out8, "plan_layout: NULL");
    int32_t max_halo = 0;
    for (const TileDesc8& d : p->tiles8) max_halo = std::max(max_halo, d.n_local - d.n_owned);
    out8[0] = (p->tile_order ? 1 : 0) | (p->unpaired_ok ? 0 : 2);
    out8[1] = max_halo;
    out8[2] = (int64_t)p->edge_pack.size();
    out8[3] = p->tile_order ? (int64_t)((size_t)p->dev.max_local * 64 + (size_t)(p->dev.max_entries + 1) * 32 + 256) : 0;   // smem bytes, FP64 kernel
    int32_t max_entries9 = 0;
    for (const TileDesc8& d : p->tiles8) max_entries9 = std::max(max_entries9, d.n_entries9);
    out8[4] = p->n_pairs;            // matched element pairs of the mesh
    out8[5] = p->pair_entries;       // pair-or-single entries over all tiles
    out8[6] = max_entries9;          // max fold slots per tile in the paired layout
    out8[7] = p->n_first_tiles;      // tiles owning the caller's first_nodes: they are tiles [0, n_first_tiles)
    return 0;
}

extern "C" int hidenn_tri_plan_locality(const hidenn_tri_plan* p, double* out2) {
    HIDENN_REQUIRE(p && out2, "plan_locality: NULL");
    out2[0] = p->runs_per_tile;
    out2[1] = p->local_per_tile;
    return 0;
}

extern "C" int hidenn_tri_plan_slots(const hidenn_tri_plan* p, int32_t* xs, int32_t* us) {
    HIDENN_REQUIRE(p && xs && us, "plan_slots: NULL");
    std::memcpy(xs, p->xslot.data(), p->xslot.size() * sizeof(int32_t));
    std::memcpy(us, p->uslot.data(), p->uslot.size() * sizeof(int32_t));
    return 0;
}

extern "C" int hidenn_tri_plan_decode(const hidenn_tri_plan* p, int64_t* out_elem, int64_t* out_nodes, uint8_t* out_owner) {
    HIDENN_REQUIRE(p && out_elem && out_nodes && out_owner, "plan_decode: NULL");
    for (size_t t = 0; t < p->tiles.size(); ++t) {
        const TileDesc& d = p->tiles[t];
        for (int32_t i = 0; i < d.n_elem; ++i) {
            const unsigned long long w = p->elem_pack[d.elem_off + i];
            const int64_t v = d.elem_off + i;
            out_elem[v] = p->t_elem[v];
            for (int c = 0; c < 3; ++c) {
                const int lid = (int)((w >> (kLidBits * c)) & ((1u << kLidBits) - 1));
                out_nodes[3 * v + c] = p->t_node[d.node_off + lid];
            }
            out_owner[v] = (uint8_t)((w >> kOwnerBit) & 1ull);
        }
    }
    return 0;
}

// Simulated shared-memory passes of the tile kernel's gathers and partial stores for the current lane assignment
// (model: one pass serves lanes whose 16-byte (FP64) / 8-byte (FP32) words fall in different bank groups; equal
// addresses broadcast).  out[0] = gather passes, out[1] = ideal gather passes, out[2] = store passes, out[3] = ideal.
extern "C" int hidenn_tri_plan_tiles(const hidenn_tri_plan* p, int64_t* node_off, int32_t* n_owned, int32_t* nodes) {
    HIDENN_REQUIRE(p && node_off && n_owned && nodes, "plan_tiles: NULL");
    const size_t nt = p->tiles.size();
    for (size_t t = 0; t < nt; ++t) { node_off[t] = p->tiles[t].node_off; n_owned[t] = p->tiles[t].n_owned; }
    node_off[nt] = (int64_t)p->t_node.size();
    std::copy(p->t_node.begin(), p->t_node.end(), nodes);
    return 0;
}

extern "C" int hidenn_tri_plan_fold_tables(const hidenn_tri_plan* p, int64_t* elem_off, uint64_t* packs, int64_t* elems, int64_t* owned_off,
                                           uint32_t* entry_off, int32_t* n_entries) {
    HIDENN_REQUIRE(p && elem_off && packs && elems && owned_off && entry_off && n_entries, "plan_fold_tables: NULL");
    HIDENN_REQUIRE(p->unpaired_ok, "plan_fold_tables: the plan is pairs-only (tiles sized for the paired layout); use hidenn_tri_plan_pair_tables or HIDENN_PLAN_PAIRS=0");
    const size_t nt = p->tiles.size();
    for (size_t t = 0; t < nt; ++t) {
        elem_off[t] = p->tiles[t].elem_off;
        owned_off[t] = p->tiles[t].off_off;
        n_entries[t] = p->tiles[t].n_entries;
    }
    elem_off[nt] = (int64_t)p->elem_pack.size();
    owned_off[nt] = (int64_t)p->entry_off.size();
    for (size_t i = 0; i < p->elem_pack.size(); ++i) { packs[i] = p->elem_pack[i]; elems[i] = p->t_elem[i]; }
    std::copy(p->entry_off.begin(), p->entry_off.end(), entry_off);
    return 0;
}

extern "C" int hidenn_tri_plan_pair_tables(const hidenn_tri_plan* p, int64_t* pent_off, uint64_t* packs, int64_t* owned_off,
                                           uint32_t* entry_off9, int32_t* n_entries9, int32_t* mate) {
    HIDENN_REQUIRE(p && pent_off && packs && owned_off && entry_off9 && n_entries9 && mate, "plan_pair_tables: NULL");
    HIDENN_REQUIRE(p->tile_order && !p->mate.empty(),
                   "plan_pair_tables: the plan has no paired layout (needs HIDENN_PLAN_PAIRS=1, a tile-ordered numbering and FP64)");
    const size_t nt = p->tiles8.size();
    int64_t a = 0, b = 0;
    for (size_t t = 0; t < nt; ++t) {
        pent_off[t] = a; owned_off[t] = b; n_entries9[t] = p->tiles8[t].n_entries9;
        a += p->tiles8[t].n_pent; b += p->tiles8[t].n_owned;
    }
    pent_off[nt] = a; owned_off[nt] = b;
    for (size_t i = 0; i < p->pair_pack.size(); ++i) packs[i] = p->pair_pack[i];
    std::copy(p->entry_off9.begin(), p->entry_off9.end(), entry_off9);
    std::copy(p->mate.begin(), p->mate.end(), mate);
    return 0;
}

extern "C" int hidenn_tri_plan_pipeline(const hidenn_tri_plan* p, int32_t* rows2, int32_t* first_need_x, int32_t* last_own_x,
                                        int32_t* first_need_u, int32_t* last_own_u) {
    HIDENN_REQUIRE(p && rows2 && first_need_x && last_own_x && first_need_u && last_own_u, "plan_pipeline: NULL");
    rows2[0] = p->pipe_rows_x; rows2[1] = p->pipe_rows_u;
    std::copy(p->first_need_x.begin(), p->first_need_x.end(), first_need_x);
    std::copy(p->last_own_x.begin(), p->last_own_x.end(), last_own_x);
    std::copy(p->first_need_u.begin(), p->first_need_u.end(), first_need_u);
    std::copy(p->last_own_u.begin(), p->last_own_u.end(), last_own_u);
    return 0;
}

extern "C" int hidenn_tri_plan_stage_stats(const hidenn_tri_plan* p, int64_t* out2) {
    HIDENN_REQUIRE(p && out2, "plan_stage_stats: NULL");
    // shared-memory passes of the staging writes / flush reads: 8 consecutive memory-order records per pass
    std::vector<int32_t> byid;
    int64_t passes = 0, ideal = 0;
    for (const TileDesc& d : p->tiles) {
        byid.resize(d.n_local);
        std::vector<std::pair<int32_t, int32_t>> rec(d.n_local);
        for (int32_t i = 0; i < d.n_local; ++i) rec[i] = {p->t_node[d.node_off + i], i};
        std::sort(rec.begin(), rec.end());
        for (int32_t j0 = 0; j0 < d.n_local; j0 += 8) {
            int cnt[8] = {0, 0, 0, 0, 0, 0, 0, 0}, mx = 0;
            for (int32_t j = j0; j < std::min(d.n_local, j0 + 8); ++j) mx = std::max(mx, ++cnt[rec[j].second & 7]);
            passes += mx;
            ideal += 1;
        }
    }
    out2[0] = passes; out2[1] = ideal;
    return 0;
}

extern "C" int hidenn_tri_plan_bank_stats(const hidenn_tri_plan* p, int real_bytes, int64_t* out4) {
    HIDENN_REQUIRE(p && out4 && (real_bytes == 8 || real_bytes == 4), "plan_bank_stats: bad arguments");
    HIDENN_REQUIRE(p->unpaired_ok, "plan_bank_stats: the plan is pairs-only (tiles sized for the paired layout)");
    (void)real_bytes;
    const int G = 8;
    constexpr unsigned LM = (1u << kLidBits) - 1u, PM = (1u << kPosBits) - 1u;
    int64_t gp = 0, gi = 0, sp = 0, si = 0;
    for (const TileDesc& d : p->tiles) {
        for (int32_t base = 0; base < d.n_elem; base += G) {
            const int n = std::min<int32_t>(G, d.n_elem - base);
            for (int c = 0; c < 3; ++c) {
                int cnt[16] = {}, cntp[16] = {};
                std::vector<unsigned> seen[16];
                for (int j = 0; j < n; ++j) {
                    const unsigned long long w = p->elem_pack[d.elem_off + base + j];
                    const unsigned lid = (unsigned)(w >> (kLidBits * c)) & LM;
                    const unsigned pos = (unsigned)(w >> (3 * kLidBits + kPosBits * c)) & PM;
                    auto& sv = seen[lid % G];
                    if (std::find(sv.begin(), sv.end(), lid) == sv.end()) { sv.push_back(lid); cnt[lid % G]++; }
                    if ((int)pos != d.n_entries) cntp[pos % G]++; else cntp[pos % G] = std::max(cntp[pos % G], 1);
                }
                gp += 2 * *std::max_element(cnt, cnt + 16);       // xy and uv gathers
                sp += 2 * *std::max_element(cntp, cntp + 16);     // gu and gx stores
                gi += 2; si += 2;
            }
        }
    }
    out4[0] = gp; out4[1] = gi; out4[2] = sp; out4[3] = si;
    return 0;
}
