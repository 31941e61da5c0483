// Device code of the Seismic query hot path for sm_100a (B200).
//
// Data layout in HBM (built once by sgpu_index_create, see sgpu_api.cu):
//   lists[dim]        ListHdr: bases of the list's slices in the arrays below
//   postings[P]       u64 (rec_start32 << 16) | nnz     — same packing as the reference's
//                     PackedPostingBlock (src/posting_list.rs:38-52) but `start` counts 32-byte
//                     units of the record buffer instead of elements
//   blk_post_off      u32, per list B+1 entries (block_offsets, relative to the list's postings)
//   blk_min/blk_quant f32 per block (QuantizedSummary::minimums / quants)
//   sc_comp           u32 ascending summary component ids per list (component_ids)
//   sc_run_off        u32, per list n_sc+1 run bounds into the entry arrays (offsets)
//   ent_blk/ent_code  u16 block id / u8 code per summary entry (summaries_ids / values)
//   fwd records       one record per document, 32-byte aligned, made of 32-byte chunks
//                     [8 x u16 component | 8 x f16 value]; the tail chunk is padded with
//                     (component 0, value +0.0).  A group of 8 lanes reads 8 consecutive chunks
//                     = 256 contiguous bytes with one 256-bit load per lane; every fetched
//                     32-byte sector is fully used.
//   rec_start[N+1]    u32 record start (32-byte units) per doc, for key -> doc id mapping
//
// Kernels (all hand written; no tensor cores — this is a gather/reduce path bounded by HBM):
//   k_prep    one warp per query: validation + top-`query_cut` term selection
//             (reference src/inverted_index.rs:172-175,187-190)
//   k_est     one warp per (query, term): QuantizedSummary::distances (Loop A,
//             reference src/quantized_summary.rs:64-160), bit-exact accumulation order
//   k_order   one CTA per query: block order of the first list by estimate, descending
//             (reference sort_and_search, src/posting_list.rs:162-166); emits one 16-byte selection entry
//             per position
//   k_search  persistent (as many CTAs as fit the SMs, work counter): block traversal with skip test,
//             forward-index scoring, bounded top-k, Knn::refine (Loop B; reference src/posting_list.rs:115-215,
//             src/utils.rs:12-66, src/inverted_index.rs:551-593) — search.cuh
//   k_finish  key -> doc id (id_from_range, reference src/inverted_index.rs:227-233)
#pragma once
#include "summary.cuh"
#include "types.cuh"

namespace sgpu {

// ------------------------------------------------------------------------------------------
// k_prep: one warp per query.
// ------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(128) k_prep(Batch b, uint32_t dim, uint32_t query_cut, uint32_t max_nnz_ok,
                                              uint32_t* nterms, uint32_t* status, uint32_t* prep) {
    // prep[1] largest nterms, [2] invalid queries, [3] queries longer than any kernel of this index can stage,
    // [6] largest query nnz
    const uint32_t lane = threadIdx.x & 31;
    const uint32_t q = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    if (q >= b.nq) return;
    const uint64_t o = b.q_off[b.q_base + q];
    const uint64_t n = b.q_off[b.q_base + q + 1] - o;
    bool bad = false;
    for (uint64_t i = lane; i < n; i += 32) {
        uint32_t c = b.q_comps[o + i];
        if (c >= dim) bad = true;
        if (i > 0 && b.q_comps[o + i - 1] > c) bad = true;  // is_sorted(): non-decreasing
    }
    bad = __any_sync(0xffffffffu, bad);
    const bool too_long = n > (uint64_t)max_nnz_ok;
    if (lane == 0) {
        uint32_t nt = (bad || too_long) ? 0u : (uint32_t)(n < (uint64_t)query_cut ? n : (uint64_t)query_cut);
        nterms[q] = nt;
        status[q] = bad ? 1u : (too_long ? 2u : 0u);
        atomicMax(&prep[1], nt);
        if (bad) atomicAdd(&prep[2], 1u);
        else if (too_long) atomicAdd(&prep[3], 1u);
        atomicMax(&prep[6], (uint32_t)(n < 0xffffffffull ? n : 0xffffffffull));
    }
}

// Term selection: descending value under total_cmp, ties by position (components are sorted, so
// this is "smaller component first").  One warp per query; `cut_eff` rounds of warp arg-max.
// The same warp then routes the query: if it has <= hq_max_nnz components and a collision-free multiplier for
// the HQ_SLOTS-slot hash table is found within HQ_TRIES attempts it goes to the hash-query kernel, else to the
// dense-query kernel.
__global__ void __launch_bounds__(128) k_terms(Batch b, Scratch sc, const ListHdr* __restrict__ lists,
                                               uint32_t hq_max_nnz, uint32_t hq_log2_slots, uint32_t hq_tries) {
    __shared__ uint32_t s_bm[4][128];  // one 4096-bit occupancy map per warp
    const uint32_t lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    const uint32_t q = blockIdx.x * (blockDim.x >> 5) + w;
    if (q >= b.nq) return;
    const uint32_t nt = sc.nterms[q];
    const uint64_t o = b.q_off[b.q_base + q];
    const uint32_t n = (uint32_t)(b.q_off[b.q_base + q + 1] - o);
    // (key, pos) of the previously selected term; next term is the max strictly after it in the order
    uint64_t last = ~0ull;  // composite (key << 32) | (~pos): larger == earlier in the order
    for (uint32_t t = 0; t < nt; ++t) {
        uint64_t best = 0;
        bool have = false;
        for (uint32_t i = lane; i < n; i += 32) {
            uint64_t c = ((uint64_t)total_key(b.q_vals[o + i]) << 32) | (uint32_t)(~i);
            if ((t == 0 || c < last) && (!have || c > best)) best = c, have = true;
        }
        for (int s = 16; s > 0; s >>= 1) {
            uint64_t ob = __shfl_xor_sync(0xffffffffu, best, s);
            bool oh = __shfl_xor_sync(0xffffffffu, (int)have, s);
            if (oh && (!have || ob > best)) best = ob, have = true;
        }
        last = best;
        if (lane == 0) sc.terms[(uint64_t)q * sc.cut_eff + t] = b.q_comps[o + (uint32_t)(~(uint32_t)best)];
    }
    // ---- routing + perfect-hash multiplier
    uint32_t mult = 0;
    if (hq_max_nnz > 0 && n <= hq_max_nnz) {
        if (nt == 0 || hq_tries == 0) {
            mult = 0x9E3779B1u;  // nothing will be staged / byte-indexed query: no hashing needed
        } else {
            const uint32_t words = (1u << hq_log2_slots) >> 5;
            for (uint32_t attempt = 0; attempt < hq_tries && mult == 0; ++attempt) {
                const uint32_t m = (2u * attempt + 1u) * 0x9E3779B1u;
                for (uint32_t i = lane; i < words; i += 32) s_bm[w][i] = 0;
                __syncwarp();
                bool coll = false;
                for (uint32_t i = lane; i < n; i += 32) {
                    const uint32_t c = b.q_comps[o + i];
                    if (i > 0 && b.q_comps[o + i - 1] == c) continue;  // duplicates share a slot
                    const uint32_t slot = (c * m) >> (32 - hq_log2_slots);
                    const uint32_t bit = 1u << (slot & 31);
                    if (atomicOr(&s_bm[w][slot >> 5], bit) & bit) coll = true;
                }
                if (!__any_sync(0xffffffffu, coll)) mult = m;
                __syncwarp();
            }
        }
    }
    if (lane == 0) {
        sc.hmult[q] = mult;
        // cost proxy for longest-first scheduling: postings of the lists the query will walk
        uint32_t cost = 0;
        for (uint32_t t = 0; t < nt; ++t) cost += lists[sc.terms[(uint64_t)q * sc.cut_eff + t]].n_post;
        sc.cost[q] = cost;
    }
}

// Longest-first work lists: a counting sort of the chunk's queries by descending cost (64 buckets), one list per
// k_search instantiation (compact / dense).  Single CTA; the order inside a bucket is irrelevant for results.
constexpr int ROUTE_THREADS = 1024;
constexpr int ROUTE_BUCKETS = 64;
__global__ void __launch_bounds__(ROUTE_THREADS) k_route(Scratch sc, uint32_t nq, uint32_t cost_max) {
    __shared__ uint32_t hist[2][ROUTE_BUCKETS], base[2][ROUTE_BUCKETS];
    const uint32_t tid = threadIdx.x;
    if (tid < 2 * ROUTE_BUCKETS) (&hist[0][0])[tid] = 0;
    __syncthreads();
    const uint32_t div = cost_max / ROUTE_BUCKETS + 1;
    for (uint32_t q = tid; q < nq; q += ROUTE_THREADS) {
        const uint32_t bkt = min(sc.cost[q] / div, (uint32_t)ROUTE_BUCKETS - 1);
        atomicAdd(&hist[sc.hmult[q] ? 0 : 1][bkt], 1u);
    }
    __syncthreads();
    if (tid < 2) {
        uint32_t acc = 0;
        for (int bkt = ROUTE_BUCKETS - 1; bkt >= 0; --bkt) {
            base[tid][bkt] = acc;
            acc += hist[tid][bkt];
        }
        sc.counters[4 + tid] = acc;
    }
    __syncthreads();
    for (uint32_t q = tid; q < nq; q += ROUTE_THREADS) {
        const uint32_t kind = sc.hmult[q] ? 0 : 1;
        const uint32_t bkt = min(sc.cost[q] / div, (uint32_t)ROUTE_BUCKETS - 1);
        const uint32_t slot = atomicAdd(&base[kind][bkt], 1u);
        (kind ? sc.qlist_dense : sc.qlist_hq)[slot] = q;
    }
}

// ------------------------------------------------------------------------------------------
// k_est / k_order: Loop A as stand-alone kernels (summary.cuh: est_task, order_task) — used for the queries that the
// dense-query / sorted-query kernels take; k_search computes Loop A for its own queries inside its CTAs (`fuse_est`).
// `skip_fused`: leave out the queries routed to the compact-query kernel (hmult != 0).
// ------------------------------------------------------------------------------------------
constexpr int EST_WARPS = 4;
constexpr int EST_SMEM = 1024;  // blocks per warp kept in shared memory (4 KB); larger lists accumulate in global
constexpr int EST_STAGE = 512;  // staged (block id, addend) pairs per warp and pass
// (Builds with less staging and more resident warps — 256 pairs / 640 accumulators at 8 or 10 CTAs per SM, 128 / 640 at
// 12 — were measured against this one, 6 CTAs per SM at 80 registers: 0.53 / 0.65 / 0.77 ms vs 0.51 ms for k_est +
// k_order per 10 k queries at cut 3; the deep staging matters more than the occupancy.)

__global__ void __launch_bounds__(EST_WARPS * 32) k_est(DevIndex ix, Batch b, Scratch sc, int skip_fused) {
    constexpr int SLICE = EST_STAGE * 6 + EST_AUX_BYTES + 16 + EST_SMEM * 4;
    __shared__ __align__(16) unsigned char s_raw[EST_WARPS][SLICE];
    const uint32_t lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    const uint32_t task = blockIdx.x * EST_WARPS + w;
    const uint32_t q = task / sc.cut_eff, t = task % sc.cut_eff;
    if (q >= b.nq || t >= sc.nterms[q]) return;  // warp-uniform; no block-wide barrier below
    if (skip_fused && sc.hmult[q] != 0) return;
    EstScratch es;
    es.carve(s_raw[w], SLICE, EST_STAGE);
    est_task(ix, b, sc, q, t, lane, es);
}

constexpr int ORDER_THREADS = 256;
constexpr int ORDER_SMEM = 4096;
constexpr uint32_t ORDER_WARP_MAX = 1024;

// `big_only`: only the queries whose first list has more than ORDER_WARP_MAX blocks (the others went to k_order_warp)
__global__ void __launch_bounds__(ORDER_THREADS) k_order(DevIndex ix, Batch b, Scratch sc, int big_only) {
    __shared__ uint64_t s_key[ORDER_SMEM];
    const uint32_t q = blockIdx.x;
    if (q >= b.nq || sc.nterms[q] == 0) return;
    if (big_only && ix.lists[sc.terms[(uint64_t)q * sc.cut_eff]].n_blk <= ORDER_WARP_MAX) return;
    order_task<ORDER_THREADS>(ix, sc, q, threadIdx.x, s_key, ORDER_SMEM);
}

// one warp per query, composites in registers (summary.cuh: order_warp); first lists of <= ORDER_WARP_MAX blocks.
// BIG = false: lists of <= 512 blocks (<= 16 composites per lane); BIG = true: 513 .. 1024 blocks (32 per lane, a
// separate kernel so that its 64 key registers do not set the occupancy of the common case).
constexpr int ORDER_WARPS = 4;
template <bool BIG>
__global__ void __launch_bounds__(ORDER_WARPS * 32) k_order_warp(DevIndex ix, Batch b, Scratch sc) {
    const uint32_t lane = threadIdx.x & 31;
    const uint32_t q = blockIdx.x * ORDER_WARPS + (threadIdx.x >> 5);
    if (q >= b.nq || sc.nterms[q] == 0) return;  // warp-uniform
    const uint32_t B = ix.lists[sc.terms[(uint64_t)q * sc.cut_eff]].n_blk;
    if constexpr (BIG) {
        if (B > 512 && B <= ORDER_WARP_MAX) order_warp<32>(ix, sc, q, lane);
    } else {
        if (B <= 128) order_warp<4>(ix, sc, q, lane);
        else if (B <= 256) order_warp<8>(ix, sc, q, lane);
        else if (B <= 512) order_warp<16>(ix, sc, q, lane);
    }
}

// ------------------------------------------------------------------------------------------
// k_finish: key (record start) -> document index.  upper_bound(rec_start, key) - 1, which resolves a start
// shared with preceding EMPTY documents to the non-empty one (pinned by reference src/inverted_index.rs:716-772).
// ------------------------------------------------------------------------------------------
__global__ void k_finish(const uint32_t* rec_start, uint64_t n_docs, const uint32_t* keys, const uint32_t* counts,
                         uint32_t k, uint32_t nq, uint64_t* out_ids) {
    const uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= (uint64_t)nq * k) return;
    const uint32_t q = (uint32_t)(i / k), r = (uint32_t)(i % k);
    if (r >= counts[q]) {
        out_ids[i] = ~0ull;
        return;
    }
    const uint32_t key = keys[i];
    uint64_t lo = 0, hi = n_docs + 1;  // first index with rec_start > key
    while (lo < hi) {
        const uint64_t mid = (lo + hi) >> 1;
        if (__ldg(rec_start + mid) <= key) lo = mid + 1;
        else hi = mid;
    }
    out_ids[i] = lo - 1;
}

// ------------------------------------------------------------------------------------------
// k_knn_posts: neighbour document ids -> postings of the record image (Knn::refine calls range_from_id per
// neighbour, src/inverted_index.rs:584; here that is done once per graph).  The nnz field is the record's chunk
// count * 8 (padding pairs are (0, +0.0) and leave every partial sum bit-identical).
// ------------------------------------------------------------------------------------------
__global__ void k_knn_posts(const uint64_t* __restrict__ ids, uint64_t n, const uint32_t* __restrict__ rec_start,
                            uint64_t n_docs, uint32_t chunk_units, uint64_t* __restrict__ out) {
    const uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const uint64_t d = ids[i];
    uint64_t post = ~0ull;
    if (d < n_docs) {
        const uint32_t r0 = rec_start[d], nch = (rec_start[d + 1] - r0) / chunk_units;
        if (nch > 0) post = ((uint64_t)r0 << 16) | (uint64_t)min(nch * 8u, 65535u);
    }
    out[i] = post;
}

// ------------------------------------------------------------------------------------------
// Image construction helpers (run once per index).
// ------------------------------------------------------------------------------------------
// directory of a list's sorted summary components: the last id of every group of 32 (one warp per list)
__global__ void k_build_skip(const ListHdr* __restrict__ lists, uint32_t dim, const uint32_t* __restrict__ sc_comp,
                             uint32_t* __restrict__ sc_skip) {
    const uint32_t l = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    if (l >= dim) return;
    const ListHdr h = lists[l];
    const uint32_t n_skip = (h.n_sc + 31) >> 5;
    for (uint32_t i = threadIdx.x & 31; i < n_skip; i += 32)
        sc_skip[h.skip_base + i] = sc_comp[h.sc_base + min(32 * i + 31, h.n_sc - 1)];
}

// one warp per document: element arrays -> 32-byte chunk records, tail padded with zeros
__global__ void k_pack_records(const uint64_t* fwd_off, const uint16_t* comps, const uint16_t* vals,
                               const uint32_t* rec_start, uint64_t doc0, uint64_t n_docs_chunk, uint64_t elem0,
                               uint16_t* records /* base of the whole record buffer */) {
    const uint64_t d = doc0 + (uint64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    if (d >= doc0 + n_docs_chunk) return;
    const uint32_t lane = threadIdx.x & 31;
    const uint64_t e0 = fwd_off[d], len = fwd_off[d + 1] - e0;
    const uint32_t nch = (uint32_t)((len + 7) >> 3);
    uint16_t* rec = records + (uint64_t)rec_start[d] * 16;
    for (uint32_t i = lane; i < nch * 8; i += 32) {
        const uint32_t ch = i >> 3, j = i & 7;
        const bool ok = i < len;
        const uint32_t sw = (!SGPU_LD256 && (ch & 4u)) ? 8u : 0u;  // chunks 4..7 of a round: [values | components] (search.cuh, ld_chunk)
        rec[ch * 16 + sw + j] = ok ? comps[e0 - elem0 + i] : (uint16_t)0;
        rec[ch * 16 + (8 - sw) + j] = ok ? vals[e0 - elem0 + i] : (uint16_t)0;
    }
}

// u32 components: chunk = 8 x u32 components followed by 8 x f16 values (48 bytes); rec_start counts 16-byte units
__global__ void k_pack_records32(const uint64_t* fwd_off, const uint32_t* comps, const uint16_t* vals,
                                 const uint32_t* rec_start, uint64_t doc0, uint64_t n_docs_chunk, uint64_t elem0,
                                 uint32_t* records /* base of the whole record buffer */) {
    const uint64_t d = doc0 + (uint64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    if (d >= doc0 + n_docs_chunk) return;
    const uint32_t lane = threadIdx.x & 31;
    const uint64_t e0 = fwd_off[d], len = fwd_off[d + 1] - e0;
    const uint32_t nch = (uint32_t)((len + 7) >> 3);
    uint32_t* rec = records + (uint64_t)rec_start[d] * 4;
    for (uint32_t i = lane; i < nch * 8; i += 32) {
        const uint32_t ch = i >> 3, j = i & 7;
        const bool ok = i < len;
        rec[ch * 12 + j] = ok ? comps[e0 - elem0 + i] : 0u;
        reinterpret_cast<uint16_t*>(rec + ch * 12 + 8)[j] = ok ? vals[e0 - elem0 + i] : (uint16_t)0;
    }
}

// any plain layout: chunk = 8 components (comp_bytes each) followed by 8 values (val_bytes each); rec_start counts
// unit_bytes units.  Byte-wise, one warp per document (runs once per index).
__global__ void k_pack_records_any(const uint64_t* fwd_off, const uint8_t* comps, const uint8_t* vals,
                                   const uint32_t* rec_start, uint64_t doc0, uint64_t n_docs_chunk, uint64_t elem0,
                                   uint8_t* records, uint32_t comp_bytes, uint32_t val_bytes, uint32_t unit_bytes) {
    const uint64_t d = doc0 + (uint64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    if (d >= doc0 + n_docs_chunk) return;
    const uint32_t lane = threadIdx.x & 31;
    const uint64_t e0 = fwd_off[d], len = fwd_off[d + 1] - e0;
    const uint32_t nch = (uint32_t)((len + 7) >> 3), chunk = 8 * (comp_bytes + val_bytes);
    uint8_t* rec = records + (uint64_t)rec_start[d] * unit_bytes;
    for (uint32_t i = lane; i < nch * 8; i += 32) {
        const uint32_t ch = i >> 3, j = i & 7;
        const bool ok = i < len;
        const uint64_t e = e0 - elem0 + i;
        for (uint32_t b = 0; b < comp_bytes; ++b) rec[ch * chunk + j * comp_bytes + b] = ok ? comps[e * comp_bytes + b] : 0;
        for (uint32_t b = 0; b < val_bytes; ++b)
            rec[ch * chunk + 8 * comp_bytes + j * val_bytes + b] = ok ? vals[e * val_bytes + b] : 0;
    }
}

// posting (element start << 16 | len) -> (record start << 16 | len)
__global__ void k_translate_postings(const uint64_t* fwd_off, const uint32_t* rec_start, uint64_t n_docs,
                                     uint64_t* postings, uint64_t n) {
    const uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const uint64_t pk = postings[i];
    const uint64_t start = pk >> 16;
    uint64_t lo = 0, hi = n_docs + 1;  // upper_bound(fwd_off, start)
    while (lo < hi) {
        const uint64_t mid = (lo + hi) >> 1;
        if (fwd_off[mid] <= start) lo = mid + 1;
        else hi = mid;
    }
    const uint64_t doc = lo - 1;
    postings[i] = ((uint64_t)rec_start[doc] << 16) | (pk & 0xffffu);
}

}  // namespace sgpu
