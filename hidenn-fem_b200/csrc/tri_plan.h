// Triangle plan: tile packs for the owner-computes fused energy kernel.
#pragma once
#include <stdint.h>
#include <vector>

namespace hidenn {

// One CTA processes one tile: a compact blob of "owned" nodes (it alone writes their gradients)
// plus every element incident to them (halo elements are recomputed by the neighbouring tile too).
struct TileDesc {          // 32 B
    int32_t node_off;      // first entry of this tile in t_xslot / t_uslot
    int32_t n_owned;       // local ids [0,n_owned) are owned, [n_owned,n_local) are halo (read only)
    int32_t n_local;
    int32_t elem_off;      // first entry in elem_pack
    int32_t n_elem;
    int32_t off_off;       // first entry in the compact host entry_off (n_owned values)
    int32_t n_entries;     // sum of owned valences = shared-memory partial slots
    int32_t pad;
};

// elem_pack bit layout (uint64): local node ids 3 x 10 bit, fold-slot positions 3 x 11 bit (precomputed
// off[node] + rank; halo corners point at the tile's dump slot = n_entries), bit 63 = this visit owns the
// element's energy.
constexpr int kLidBits = 10;
constexpr int kPosBits = 11;
constexpr int kOwnerBit = 63;
constexpr int kMaxLocal = (1 << kLidBits) - 1;       // local ids 0..1022
constexpr int kMaxEntries = (1 << kPosBits) - 1;     // fold slots 0..2046 + dump slot <= 2047
constexpr int kMaxValence = 255;
constexpr int kPipeBlocks = 64;

// ---- "tile-ordered" layout (kernel v8) ----------------------------------------------------------------------------
// When the node numbering lists every tile's owned nodes as ONE contiguous id range, and inside the range by class
//   A: coordinate free, displacement free   B: coordinate fixed, displacement free
//   C: coordinate fixed, displacement fixed  D: coordinate free, displacement fixed
// (hidenn_tri_locality_order produces such a numbering), the owned rows of a tile are at most two contiguous runs of
// each Parameter / fixed buffer.  The tile kernel then stages them with bulk copies (cp.async.bulk + mbarrier, one
// elected thread), uses local id = id - first id (memory order), and the fold threads store the final gradient rows
// directly (consecutive lanes -> consecutive rows): no per-node slot records, no output staging buffer.
struct TileDesc8 {           // 64 B
    int32_t n_owned, n_local, n_elem, n_entries;
    int32_t nA, nB, nC, nD;                              // owned class sizes, in local-id order
    int32_t rx_free, rx_fixed, ru_free, ru_fixed;        // first row of the tile's run in each array
    int32_t edge_off, n_edge;                            // Neumann edge visits (edge_pack / edge_id)
    int32_t n_pent, n_entries9;                          // pair entries / fold slots of the paired layout (kernel v9)
};
// edge_pack bit layout (uint64): local ids of the two ends 2 x 10 bit, fold-slot positions 2 x 11 bit (dump slot for a
// halo end), bit 63 = this visit owns the edge's energy (and its d loss / d traction row)
// Paired layout (kernel v9): the elements of the mesh are matched once, globally, into edge-sharing pairs (greedy
// matching of the dual graph); a thread evaluates both elements of a pair and adds the two partials of each shared node
// in registers, so that node receives ONE partial for the pair: 4 instead of 6 partial stores and fold reads per pair.
// pair_pack holds two 64-bit words of the elem_pack format per entry (first element = smaller element id; corners of
// the second element that are merged carry the dump position; a single element has the null word kNullPack second).
// Partners run through their shared edge in opposite directions (both counter-clockwise or both clockwise), so with the
// first element's corners Q0 Q1 Q2 and the second element's new corner Q3 a pair is one of 9 classes
//   cls = 3 i + r:   second element: corner r = Q3, corner r+1 = Q(i+1), corner r+2 = Q(i)      (indices mod 3)
// and the kernel has one compile-time register wiring per class (the reference's J^-1 D_N quirk makes an element's result
// depend on its own corner order, so the corners cannot be rotated into a canonical position).  Entries are listed class
// by class.  On the DEVICE the second word is  lid(Q3) | pos(Q3) << 10 | cls << 21 | owner << 63  (cls = kPairSingle for
// a single element); the host copy keeps the two elem_pack words (tests replay them).
struct TriPlan8Dev {
    const TileDesc8* tiles;
    const int2* t_halo;                    // [n_tiles, stride_halo]: (xslot, uslot) of the halo nodes, ascending node id
    const unsigned long long* edge_pack;   // [n_edge_visits]
    const int32_t* edge_id;                // [n_edge_visits]
    int32_t stride_halo, max_halo, n_edge_visits, pad;
    const unsigned long long* pair_pack;   // [n_tiles, stride_pent, 2]
    const uint32_t* entry_off9;            // [n_tiles, stride_owned]: fold-slot start | count << 16 (paired layout)
    const unsigned long long* edge_pack9;  // [n_edge_visits]: edge_pack with the slot positions of the paired layout
    int32_t stride_pent, max_entries9, pad2, pad3;
};
constexpr unsigned long long kNullPack = 0x3FFFFFFFull;      // local ids 1023,1023,1023: no element
constexpr int kPairSingle = 9, kPairSkip = 10;      // skip: padding entry (both words null), no work
int pair_class(unsigned long long w1, unsigned long long w2);

struct TriPlanDev {
    const TileDesc* tiles;
    int32_t n_tiles;
    // fixed-stride tile records (tile t starts at t*stride): every load address depends only on blockIdx
    // node records are listed in MEMORY order (ascending node id = ascending Parameter row): the staging gathers
    // and the final gradient stores of a warp then touch consecutive 16-byte pairs; t_lid gives the tile-local id
    // (shared-memory position; owned nodes < n_owned) of each record
    const int2* t_slots;                   // [n_tiles, stride_local]: (xslot, uslot), padded with a loadable row
    const uint16_t* t_lid;                 // [n_tiles, stride_local]: local id, padding = 0xFFFF
    const unsigned long long* elem_pack;   // [n_tiles, stride_elem], padded with 0
    const uint32_t* entry_off;             // [n_tiles, stride_owned]: fold-slot start | count << 16 per owned node
    int32_t stride_local, stride_elem, stride_owned;
    int32_t max_local, max_entries, max_owned, max_elem;
    // Neumann edges
    int32_t n_edges;
    const int32_t* e_slots;      // [Ned,4]: xslot0, uslot0, xslot1, uslot1
    int32_t n_enodes;
    const int32_t* en_xslot;     // [n_enodes]
    const int32_t* en_uslot;
    const int32_t* en_off;       // [n_enodes+1]
    const int32_t* en_ent;       // edge*2 + end
    // global views (generic forward / fold); uploaded lazily
    const int32_t* conn32;       // [Ne,3]
    const int32_t* xslot;        // [Nn]
    const int32_t* uslot;
    const int64_t* n2e_off;      // [Nn+1]
    const int32_t* n2e_ent;      // element*4 + corner, ascending element id per node
    const int32_t* edges32;      // [Ned,2]
    int64_t n_elems, n_nodes, n_free_x, n_free_u;
    // correct-math switch (default 0 = the reference's J^-1 quirk): 1 = physical gradients use J^-T
    int32_t jinv_t, pad_opt;
};

}  // namespace hidenn

struct hidenn_tri_plan {
    int device = 0;
    int real_bytes = 8;
    int64_t n_elems = 0, n_nodes = 0, n_free_x = 0, n_free_u = 0, n_fixed_x = 0, n_fixed_u = 0;
    int64_t elem_visits = 0, node_visits = 0;
    // host copies (tests / decode)
    std::vector<hidenn::TileDesc> tiles;
    std::vector<int32_t> t_node;            // global node id of every tile-local node
    std::vector<int32_t> t_elem;            // global element id of every tile element visit
    std::vector<unsigned long long> elem_pack;
    std::vector<uint32_t> entry_off;        // compact: start | count << 16 per owned node, tiles back to back
    std::vector<int32_t> xslot, uslot;
    std::vector<int32_t> conn32;
    std::vector<int64_t> n2e_off;
    std::vector<int32_t> n2e_ent;
    std::vector<int32_t> edges32;
    hidenn::TriPlanDev dev{};
    bool tile_order = false;                // numbering is tile-ordered -> kernel v8 (FP64)
    bool unpaired_ok = true;                // false: tiles sized for the paired layout, the one-element-per-entry packs overflow
                                            // their position fields -> only kernel v9 (paired) can run this plan
    hidenn::TriPlan8Dev dev8{};
    std::vector<hidenn::TileDesc8> tiles8;
    std::vector<unsigned long long> edge_pack, edge_pack9;
    std::vector<int32_t> edge_id;
    std::vector<unsigned long long> pair_pack;      // compact: 2 words per entry, tiles back to back
    std::vector<uint32_t> entry_off9;               // compact, like entry_off
    std::vector<int32_t> mate;                      // global matching: partner element or -1
    int64_t n_pairs = 0, pair_entries = 0;
    double runs_per_tile = 0.0, local_per_tile = 0.0;      // numbering locality (hidenn_tri_plan_locality)
    int32_t n_first_tiles = 0;                      // tiles owning the caller's first_nodes, listed first
    std::vector<void*> dev_allocs;
    size_t dev_bytes = 0;
    bool generic_uploaded = false;
    // arena for the host-buffer entry points
    void* arena = nullptr;
    size_t arena_bytes = 0;
    // Host-buffer pipeline (tri_energy_host).  Tiles are listed by ascending smallest owned node id and the Parameter
    // rows are cut into kPipeBlocks blocks; first_need_*[b] is the first tile that reads a row of block b, last_own_*[b]
    // the last tile that writes one.  A chunk of tiles then needs the blocks whose first_need falls in it (sent
    // host->device just before) and completes the blocks whose last_own falls in it (fetched device->host right after),
    // so on side streams the copies of chunk c+1 and c-1 overlap the kernels of chunk c in both PCIe directions.
    int32_t pipe_rows_x = 0, pipe_rows_u = 0;                                    // rows per block
    std::vector<int32_t> first_need_x, last_own_x, first_need_u, last_own_u;     // [kPipeBlocks]
    std::vector<int32_t> en_xslot_h, en_uslot_h;              // host copies of the Neumann edge-node rows
    void* pipe_streams[2] = {nullptr, nullptr};               // cudaStream_t: rows in, gradient rows out
    std::vector<void*> pipe_events;                           // cudaEvent_t
};

namespace hidenn {
int plan_ensure_generic(hidenn_tri_plan* p);
int plan_ensure_arena(hidenn_tri_plan* p, size_t bytes);
}
