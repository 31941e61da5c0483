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
//                     = 256 contiguous bytes with two 128-bit loads per lane; every fetched
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
// k_est: Loop A.  One warp per (query, term), four independent warps per CTA.  est[s] accumulates, for the query
// components present in the list's summaries IN ASCENDING COMPONENT ORDER, ((code * quant[s]) + min[s]) * qv with
// four separate roundings (Rust does not contract to FMA).
//
// A task touches one to a few thousand summary entries (the list's own component alone occurs in almost every block
// summary) and is bound by dependent memory round trips, not by bytes.  The addends do not depend on the accumulation
// order, only the additions do, so a batch of up to 64 query components is handled in three steps: (1) every lane
// searches two components in the list's sorted summary components (5-ary search: four independent probes per step,
// then one 8-element probe) and fetches the run bounds; a warp scan lays the runs of the matched components end to
// end; (2) the lanes walk that flat entry list, EST_U positions per lane and step, all loads of a step issued before
// the first use — one memory latency covers 32 * EST_U entries instead of one run — and stage (block id, addend)
// pairs in shared memory; (3) the staged pairs are added run by run (= component by component, ascending; a summary id
// occurs at most once per component, so the lanes of one step never collide), __syncwarp() between runs.
// The accumulators live in shared memory when the list has <= EST_SMEM blocks, else in the global scratch.
// ------------------------------------------------------------------------------------------
constexpr int EST_WARPS = 4;
constexpr int EST_SMEM = 1024;  // blocks per warp kept in shared memory (4 KB); larger lists accumulate in global
constexpr int EST_STAGE = 512;  // staged (block id, addend) pairs per warp and pass
constexpr int EST_QB = 64;      // query components per batch (two per lane)
constexpr int EST_U = 4;        // flat positions per lane and staging step

// lower_bound(a[0, n), c): four independent probes per step, one aligned-size probe of up to 8 elements at the end
__device__ __forceinline__ uint32_t lower_bound5(const uint32_t* __restrict__ a, uint32_t n, uint32_t c) {
    uint32_t lo = 0, hi = n;
    while (hi - lo > 8) {
        const uint32_t st = (hi - lo + 4) / 5;  // five pieces of st elements; probe the last element of the first four
        uint32_t v[4];
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            const uint32_t at = lo + (k + 1) * st - 1;
            v[k] = at < hi ? __ldg(a + at) : 0xffffffffu;
        }
        uint32_t cnt = 0;
#pragma unroll
        for (int k = 0; k < 4; ++k) cnt += v[k] < c;
        lo += cnt * st;
        hi = min(hi, lo + st);
    }
    uint32_t v[8];
#pragma unroll
    for (int k = 0; k < 8; ++k) v[k] = lo + k < hi ? __ldg(a + lo + k) : 0xffffffffu;
    uint32_t cnt = 0;
#pragma unroll
    for (int k = 0; k < 8; ++k) cnt += v[k] < c;
    return lo + cnt;
}

__global__ void __launch_bounds__(EST_WARPS * 32) k_est(DevIndex ix, Batch b, Scratch sc) {
    __shared__ float s_est[EST_WARPS][EST_SMEM];
    __shared__ float s_add[EST_WARPS][EST_STAGE];
    __shared__ uint16_t s_blk[EST_WARPS][EST_STAGE];
    __shared__ uint32_t s_off[EST_WARPS][EST_QB + 1];  // first flat position of every component's run (exclusive scan)
    __shared__ uint32_t s_e0[EST_WARPS][EST_QB];
    __shared__ float s_qv[EST_WARPS][EST_QB];
    __shared__ uint8_t s_own[EST_WARPS][EST_STAGE / 32];  // owner (component slot) of every 32nd position of the pass
    const uint32_t lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    const uint32_t task = blockIdx.x * EST_WARPS + w;
    const uint32_t q = task / sc.cut_eff, t = task % sc.cut_eff;
    if (q >= b.nq || t >= sc.nterms[q]) return;  // warp-uniform; no block-wide barrier below
    const uint32_t l = sc.terms[(uint64_t)q * sc.cut_eff + t];
    const ListHdr h = ix.lists[l];
    const uint32_t B = h.n_blk;
    float* g_est = sc.est + ((uint64_t)q * sc.cut_eff + t) * sc.est_stride;
    const bool in_smem = B <= EST_SMEM;
    float* acc = in_smem ? s_est[w] : g_est;
    for (uint32_t i = lane; i < B; i += 32) acc[i] = 0.f;
    const uint64_t o = b.q_off[b.q_base + q];
    const uint32_t n = (uint32_t)(b.q_off[b.q_base + q + 1] - o);
    const uint32_t* scomp = ix.sc_comp + h.sc_base;
    const uint32_t* skip = ix.sc_skip + h.skip_base;
    const uint32_t n_skip = (h.n_sc + 31) >> 5;
    const uint32_t* run = ix.sc_run_off + h.sc_base + l;
    const uint16_t* eb = ix.ent_blk + h.ent_base;
    const uint8_t* ec = ix.ent_code + h.ent_base;
    const float* mins = ix.blk_min + h.blk_base;
    const float* quants = ix.blk_quant + h.blk_base;
    uint32_t* off = s_off[w];
    uint32_t* e0s = s_e0[w];
    float* qvs = s_qv[w];
    float* st_add = s_add[w];
    uint16_t* st_blk = s_blk[w];
    uint8_t* own = s_own[w];
    __syncwarp();
    if (!in_smem) __threadfence_block();
    for (uint32_t base = 0; base < n; base += EST_QB) {
        // ---- (1) lane handles query components base + 2 * lane and base + 2 * lane + 1 (ascending across lanes)
        uint32_t e0[2] = {0, 0}, len[2] = {0, 0};
        float qv[2] = {0.f, 0.f};
#pragma unroll
        for (int u = 0; u < 2; ++u) {
            const uint32_t i = base + 2 * lane + u;
            if (i < n) {
                const uint32_t c = b.q_comps[o + i];
                qv[u] = b.q_vals[o + i];
                if (!(i > 0 && b.q_comps[o + i - 1] == c)) {  // the merge consumes the first duplicate only
                    // directory first (the 40 searches of a task share its ~16 sectors), then one group of 32
                    const uint32_t g = lower_bound5(skip, n_skip, c);
                    const uint32_t glen = g < n_skip ? min(32u, h.n_sc - 32 * g) : 0u;
                    const uint32_t lo = 32 * g + lower_bound5(scomp + 32 * g, glen, c);
                    if (g < n_skip && lo < h.n_sc && __ldg(scomp + lo) == c) {
                        e0[u] = __ldg(run + lo);
                        len[u] = __ldg(run + lo + 1) - e0[u];
                    }
                }
            }
        }
        // flat entry list of the batch: exclusive prefix sum of the run lengths (slot 2 * lane + u)
        const uint32_t mine = len[0] + len[1];
        uint32_t incl = mine;
#pragma unroll
        for (int sft = 1; sft < 32; sft <<= 1) {
            const uint32_t up = __shfl_up_sync(0xffffffffu, incl, sft);
            if (lane >= (uint32_t)sft) incl += up;
        }
        const uint32_t total = __shfl_sync(0xffffffffu, incl, 31);
        off[2 * lane] = incl - mine;
        off[2 * lane + 1] = incl - mine + len[0];
        e0s[2 * lane] = e0[0], e0s[2 * lane + 1] = e0[1];
        qvs[2 * lane] = qv[0], qvs[2 * lane + 1] = qv[1];
        if (lane == 0) off[EST_QB] = total;
        __syncwarp();
        for (uint32_t w0 = 0; w0 < total; w0 += EST_STAGE) {
            const uint32_t w1 = min(total, w0 + EST_STAGE);
            // owner of every 32nd position of the pass: last slot whose first position is <= p
            if (lane * 32 < w1 - w0) {
                const uint32_t p = w0 + lane * 32;
                uint32_t j = 0;
#pragma unroll
                for (int step = EST_QB / 2; step > 0; step >>= 1)
                    if (off[j + step] <= p) j += step;
                own[lane] = (uint8_t)j;
            }
            __syncwarp();
            // ---- (2) stage the addends of flat positions [w0, w1)
            for (uint32_t p0 = w0; p0 < w1; p0 += 32 * EST_U) {
                uint32_t pp[EST_U], ee[EST_U], ss[EST_U];
                float wq[EST_U], code[EST_U], qn[EST_U], mn[EST_U];
#pragma unroll
                for (int u = 0; u < EST_U; ++u) {
                    pp[u] = p0 + u * 32 + lane;
                    if (pp[u] < w1) {
                        uint32_t j = own[(pp[u] - w0) >> 5];
                        while (off[j + 1] <= pp[u]) ++j;  // off[EST_QB] = total > p ends the walk
                        ee[u] = e0s[j] + (pp[u] - off[j]);
                        wq[u] = qvs[j];
                    }
                }
#pragma unroll
                for (int u = 0; u < EST_U; ++u)
                    if (pp[u] < w1) ss[u] = __ldg(eb + ee[u]), code[u] = (float)__ldg(ec + ee[u]);
#pragma unroll
                for (int u = 0; u < EST_U; ++u)
                    if (pp[u] < w1) qn[u] = __ldg(quants + ss[u]), mn[u] = __ldg(mins + ss[u]);
#pragma unroll
                for (int u = 0; u < EST_U; ++u)
                    if (pp[u] < w1) {
                        st_blk[pp[u] - w0] = (uint16_t)ss[u];
                        st_add[pp[u] - w0] = __fmul_rn(__fadd_rn(__fmul_rn(code[u], qn[u]), mn[u]), wq[u]);
                    }
            }
            __syncwarp();
            // ---- (3) add, one run (= one component) at a time, in ascending component order
            for (uint32_t j = 0; j < EST_QB; ++j) {
                const uint32_t r0 = off[j], r1 = off[j + 1];
                if (r1 <= w0 || r0 >= w1 || r0 == r1) continue;  // warp-uniform
                const uint32_t a0 = max(r0, w0), a1 = min(r1, w1);
                for (uint32_t p = a0 + lane; p < a1; p += 32) {
                    const uint32_t s = st_blk[p - w0];
                    const float add = st_add[p - w0];
                    if (in_smem) {
                        acc[s] = __fadd_rn(acc[s], add);
                    } else {
                        const float cur = __ldcg(acc + s);
                        __stcg(acc + s, __fadd_rn(cur, add));
                    }
                }
                __syncwarp();
            }
        }
        __syncwarp();
    }
    if (in_smem)
        for (uint32_t i = lane; i < B; i += 32) g_est[i] = acc[i];
}

// ------------------------------------------------------------------------------------------
// k_order: blocks of the FIRST list of each query sorted by (estimate desc under total_cmp, block id asc).
// One CTA per query.  Bitonic sort of 64-bit composites in shared memory when B <= ORDER_SMEM, else a
// rank sort straight from global memory (correct for any B <= 65535, slow; never hit by sane configs).
// ------------------------------------------------------------------------------------------
constexpr int ORDER_THREADS = 256;
constexpr int ORDER_SMEM = 4096;

__global__ void __launch_bounds__(ORDER_THREADS) k_order(DevIndex ix, Batch b, Scratch sc) {
    __shared__ uint64_t s_key[ORDER_SMEM];
    const uint32_t q = blockIdx.x;
    if (q >= b.nq || sc.nterms[q] == 0) return;
    const uint32_t l = sc.terms[(uint64_t)q * sc.cut_eff];
    const uint32_t B = ix.lists[l].n_blk;
    const float* est = sc.est + (uint64_t)q * sc.cut_eff * sc.est_stride;
    uint4* out = sc.sel + (uint64_t)q * sc.est_stride;
    const uint32_t* boff = ix.blk_post_off + ix.lists[l].blk_base + l;
    // the search kernel reads one 16-byte entry per position: no dependent order -> estimate -> offsets chain
    auto emit = [&](uint32_t pos, uint32_t blk) {
        const uint32_t p0 = boff[blk];
        out[pos] = make_uint4(__float_as_uint(est[blk]), p0, boff[blk + 1] - p0, blk);
    };
    if (B <= ORDER_SMEM) {
        uint32_t n2 = 1;
        while (n2 < B) n2 <<= 1;
        // ascending sort of ((~key) << 32 | id): smallest composite == largest estimate, then smallest id
        for (uint32_t i = threadIdx.x; i < n2; i += ORDER_THREADS)
            s_key[i] = i < B ? (((uint64_t)(~total_key(est[i])) << 32) | i) : ~0ull;
        __syncthreads();
        // thread t owns elements t, t + 256, ...: for j < 32 both partners of an exchange belong to the same warp
        // (same 32-aligned group of elements), so only the steps with j >= 32 need a block-wide barrier
        for (uint32_t ksz = 2; ksz <= n2; ksz <<= 1)
            for (uint32_t j = ksz >> 1; j > 0; j >>= 1) {
                for (uint32_t i = threadIdx.x; i < n2; i += ORDER_THREADS) {
                    uint32_t p = i ^ j;
                    if (p > i) {
                        uint64_t a = s_key[i], c = s_key[p];
                        bool up = (i & ksz) == 0;
                        if ((a > c) == up) s_key[i] = c, s_key[p] = a;
                    }
                }
                if (j >= 32 || j == 1 && (ksz << 1) > 32) __syncthreads();  // also before the next step's wide exchange
                else __syncwarp();
            }
        for (uint32_t i = threadIdx.x; i < B; i += ORDER_THREADS) emit(i, (uint32_t)(s_key[i] & 0xffffu));
    } else {
        for (uint32_t i = threadIdx.x; i < B; i += ORDER_THREADS) {
            const uint64_t mine = ((uint64_t)(~total_key(est[i])) << 32) | i;
            uint32_t rank = 0;
            for (uint32_t j = 0; j < B; ++j) rank += ((((uint64_t)(~total_key(est[j])) << 32) | j) < mine);
            emit(rank, i);
        }
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
        const uint32_t sw = (ch & 4u) ? 8u : 0u;  // chunks 4..7 of a round: [values | components] (search.cuh, ld_chunk)
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
