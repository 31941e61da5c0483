// Loop A of the Seismic query path as device routines shared by the stand-alone kernels (k_est / k_order, kernels.cuh)
// and by k_search, which runs them inside its own CTAs (search.cuh, `fuse_est`): QuantizedSummary::distances
// (reference src/quantized_summary.rs:64-160) and the block order of the first list (sort_and_search,
// src/posting_list.rs:162-166).
#pragma once
#include "types.cuh"

namespace sgpu {

// ------------------------------------------------------------------------------------------
// est_task: Loop A for one (query, term), run by ONE warp.  est[s] accumulates, for the query
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
constexpr int EST_QB = 64;  // query components per batch (two per lane)
constexpr int EST_U = 4;    // flat positions per lane and staging step
constexpr int EST_AUX_BYTES = (EST_QB + 1) * 4 + EST_QB * 4 + EST_QB * 4 + 32;  // off / e0 / qv / own

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

// Shared-memory scratch of one warp: accumulators for acc_cap blocks (lists with more blocks accumulate in the global
// scratch), stage_cap staged pairs (a multiple of 32, <= 1024), EST_AUX_BYTES of run bookkeeping.
struct EstScratch {
    float* acc;
    uint32_t acc_cap;
    float* st_add;
    uint16_t* st_blk;
    uint32_t stage_cap;
    uint32_t* off;   // [EST_QB + 1] first flat position of every component's run (exclusive scan)
    uint32_t* e0s;   // [EST_QB]
    float* qvs;      // [EST_QB]
    uint8_t* own;    // [32] owner (component slot) of every 32nd position of the pass
    // carve a scratch out of `bytes` bytes at `base` (16-byte aligned): stage_cap pairs, the rest accumulators
    __device__ __forceinline__ void carve(unsigned char* base, uint32_t bytes, uint32_t stage) {
        stage_cap = stage;
        st_add = reinterpret_cast<float*>(base);
        st_blk = reinterpret_cast<uint16_t*>(base + stage * 4);
        unsigned char* p = base + stage * 6;
        off = reinterpret_cast<uint32_t*>(p);
        e0s = off + EST_QB + 1;
        qvs = reinterpret_cast<float*>(e0s + EST_QB);
        own = reinterpret_cast<uint8_t*>(qvs + EST_QB);
        p += EST_AUX_BYTES;
        p = reinterpret_cast<unsigned char*>((reinterpret_cast<uintptr_t>(p) + 3) & ~(uintptr_t)3);
        acc = reinterpret_cast<float*>(p);
        acc_cap = (uint32_t)((base + bytes - p) / 4);
    }
};

__device__ __forceinline__ void est_task(const DevIndex& ix, const Batch& b, const Scratch& sc, uint32_t q, uint32_t t,
                                         uint32_t lane, const EstScratch& es) {
    const uint32_t l = sc.terms[(uint64_t)q * sc.cut_eff + t];
    const ListHdr h = ix.lists[l];
    const uint32_t B = h.n_blk;
    float* g_est = sc.est + ((uint64_t)q * sc.cut_eff + t) * sc.est_stride;
    const bool in_smem = B <= es.acc_cap;
    float* acc = in_smem ? es.acc : g_est;
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
    uint32_t* off = es.off;
    uint32_t* e0s = es.e0s;
    float* qvs = es.qvs;
    float* st_add = es.st_add;
    uint16_t* st_blk = es.st_blk;
    uint8_t* own = es.own;
    const uint32_t stage_cap = es.stage_cap;
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
        for (uint32_t w0 = 0; w0 < total; w0 += stage_cap) {
            const uint32_t w1 = min(total, w0 + stage_cap);
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
// order_task: blocks of the FIRST list of a query sorted by (estimate desc under total_cmp, block id asc), run by a
// whole CTA of T threads.  Bitonic sort of 64-bit composites in `s_key` (key_cap entries of shared memory) when the
// padded block count fits, else a rank sort straight from global memory (correct for any B <= 65535, slow; never hit
// by sane configs).  Emits one 16-byte selection entry {estimate, first posting, postings, block} per position so that
// a selection pass of k_search is a single coalesced load.  All threads must call it; ends without a barrier.
// ------------------------------------------------------------------------------------------
template <int T>
__device__ __forceinline__ void order_task(const DevIndex& ix, const Scratch& sc, uint32_t q, uint32_t tid, uint64_t* s_key,
                                           uint32_t key_cap) {
    const uint32_t l = sc.terms[(uint64_t)q * sc.cut_eff];
    const uint32_t B = ix.lists[l].n_blk;
    const float* est = sc.est + (uint64_t)q * sc.cut_eff * sc.est_stride;
    uint4* out = sc.sel + (uint64_t)q * sc.est_stride;
    const uint32_t* boff = ix.blk_post_off + ix.lists[l].blk_base + l;
    auto emit = [&](uint32_t pos, uint32_t blk) {
        const uint32_t p0 = boff[blk];
        out[pos] = make_uint4(__float_as_uint(__ldcg(est + blk)), p0, boff[blk + 1] - p0, blk);
    };
    uint32_t n2 = 1;
    while (n2 < B) n2 <<= 1;
    if (n2 <= key_cap) {
        // ascending sort of ((~key) << 32 | id): smallest composite == largest estimate, then smallest id
        for (uint32_t i = tid; i < n2; i += T)
            s_key[i] = i < B ? (((uint64_t)(~total_key(__ldcg(est + i))) << 32) | i) : ~0ull;
        __syncthreads();
        // thread t owns elements t, t + T, ...: for j < 32 both partners of an exchange belong to the same warp
        // (same 32-aligned group of elements), so only the steps with j >= 32 need a block-wide barrier
        for (uint32_t ksz = 2; ksz <= n2; ksz <<= 1)
            for (uint32_t j = ksz >> 1; j > 0; j >>= 1) {
                for (uint32_t i = tid; i < n2; i += T) {
                    uint32_t p = i ^ j;
                    if (p > i) {
                        uint64_t a = s_key[i], c = s_key[p];
                        bool up = (i & ksz) == 0;
                        if ((a > c) == up) s_key[i] = c, s_key[p] = a;
                    }
                }
                if (j >= 32 || (j == 1 && (ksz << 1) > 32)) __syncthreads();  // also before the next step's wide exchange
                else __syncwarp();
            }
        for (uint32_t i = tid; i < B; i += T) emit(i, (uint32_t)(s_key[i] & 0xffffu));
    } else {
        for (uint32_t i = tid; i < B; i += T) {
            const uint64_t mine = ((uint64_t)(~total_key(__ldcg(est + i))) << 32) | i;
            uint32_t rank = 0;
            for (uint32_t j = 0; j < B; ++j) rank += ((((uint64_t)(~total_key(__ldcg(est + j))) << 32) | j) < mine);
            emit(rank, i);
        }
    }
}


// ------------------------------------------------------------------------------------------
// order_warp<E>: the same order for a list of B <= 32 * E blocks, by ONE warp with the composites in registers
// (element i = slot i / 32 of lane i % 32): bitonic network, exchanges at distance >= 32 stay inside a lane, the others
// are one 64-bit shuffle — no shared memory, no barrier.  A third of the instructions of the CTA-wide version.
// ------------------------------------------------------------------------------------------
template <int E, int JJ>
__device__ __forceinline__ void order_inlane(uint64_t (&key)[E], uint32_t ksz) {
#pragma unroll
    for (int e = 0; e < E; ++e) {
        if ((e & JJ) == 0 && (e | JJ) < E) {
            const uint64_t a = key[e], c = key[e | JJ];
            const bool up = (((uint32_t)e << 5) & ksz) == 0;  // ksz >= 64 here: the direction bit is a slot bit
            const bool sw = (a > c) == up;
            key[e] = sw ? c : a;
            key[e | JJ] = sw ? a : c;
        }
    }
}
template <int E>
__device__ __forceinline__ void order_warp(const DevIndex& ix, const Scratch& sc, uint32_t q, uint32_t lane) {
    const uint32_t l = sc.terms[(uint64_t)q * sc.cut_eff];
    const uint32_t B = ix.lists[l].n_blk;
    const float* est = sc.est + (uint64_t)q * sc.cut_eff * sc.est_stride;
    uint4* out = sc.sel + (uint64_t)q * sc.est_stride;
    const uint32_t* boff = ix.blk_post_off + ix.lists[l].blk_base + l;
    uint64_t key[E];
#pragma unroll
    for (int e = 0; e < E; ++e) {
        const uint32_t i = e * 32 + lane;
        key[e] = i < B ? (((uint64_t)(~total_key(__ldcg(est + i))) << 32) | i) : ~0ull;
    }
    constexpr uint32_t N2 = 32 * E;
    for (uint32_t ksz = 2; ksz <= N2; ksz <<= 1) {
        for (uint32_t j = ksz >> 1; j >= 32; j >>= 1) {  // partner = another slot of the same lane
            switch (j >> 5) {
                case 16: order_inlane<E, 16>(key, ksz); break;
                case 8: order_inlane<E, 8>(key, ksz); break;
                case 4: order_inlane<E, 4>(key, ksz); break;
                case 2: order_inlane<E, 2>(key, ksz); break;
                default: order_inlane<E, 1>(key, ksz); break;
            }
        }
        for (uint32_t j = min(ksz >> 1, 16u); j > 0; j >>= 1) {  // partner = the same slot of lane ^ j
            const bool lower = (lane & j) == 0;
#pragma unroll
            for (int e = 0; e < E; ++e) {
                const uint64_t a = key[e];
                const uint64_t o = __shfl_xor_sync(0xffffffffu, a, j);
                const bool up = ((((uint32_t)e << 5) | lane) & ksz) == 0;
                const bool take_min = lower == up;
                key[e] = ((o < a) == take_min) ? o : a;
            }
        }
    }
#pragma unroll
    for (int e = 0; e < E; ++e) {
        const uint32_t pos = e * 32 + lane;
        if (pos < B) {
            const uint32_t blk = (uint32_t)(key[e] & 0xffffffffu);
            const uint32_t t = ~(uint32_t)(key[e] >> 32);  // total_key of the estimate
            const uint32_t bits = (t & 0x80000000u) ? (t & 0x7fffffffu) : ~t;
            const uint32_t p0 = boff[blk];
            out[pos] = make_uint4(bits, p0, boff[blk + 1] - p0, blk);
        }
    }
}

}  // namespace sgpu
