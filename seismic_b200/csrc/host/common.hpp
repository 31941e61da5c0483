// Shared host-side helpers: error string, f16 conversion, parallel_for, counter-based RNG.
#pragma once
#include <algorithm>
#include <atomic>
#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cstring>
#include <functional>
#include <string>
#include <thread>
#include <vector>

#include "../../../include/seismic_b200.h"

namespace shost {

void set_error(const std::string& msg);  // defined in host_api.cpp (thread-local, shared with sgpu_*)

// ---- IEEE binary16 <-> binary32, round-to-nearest-even (what the `half` crate does) ----
static inline uint16_t f32_to_f16_bits(float f) {
    uint32_t x;
    std::memcpy(&x, &f, 4);
    uint32_t sign = (x >> 16) & 0x8000u;
    uint32_t absx = x & 0x7fffffffu;
    if (absx >= 0x7f800000u) {  // inf / nan
        return (uint16_t)(sign | 0x7c00u | (absx > 0x7f800000u ? 0x0200u | ((absx >> 13) & 0x3ffu) : 0));
    }
    if (absx >= 0x477ff000u) {  // >= 65520 rounds to inf
        return (uint16_t)(sign | 0x7c00u);
    }
    if (absx < 0x38800000u) {  // subnormal half or zero
        if (absx < 0x33000000u) return (uint16_t)sign;  // < 2^-25 -> 0
        int exp = (int)(absx >> 23);
        uint32_t mant = (absx & 0x7fffffu) | 0x800000u;
        int shift = 126 - exp;  // 14..24
        uint32_t half_mant = mant >> shift;
        uint32_t rem = mant & ((1u << shift) - 1);
        uint32_t halfway = 1u << (shift - 1);
        if (rem > halfway || (rem == halfway && (half_mant & 1))) half_mant++;
        return (uint16_t)(sign | half_mant);
    }
    uint32_t exp = (absx >> 23) - 112;
    uint32_t mant = absx & 0x7fffffu;
    uint32_t h = (exp << 10) | (mant >> 13);
    uint32_t rem = mant & 0x1fffu;
    if (rem > 0x1000u || (rem == 0x1000u && (h & 1))) h++;
    return (uint16_t)(sign | h);
}

#if defined(__F16C__)
}  // namespace shost
#include <immintrin.h>
namespace shost {
static inline float f16_bits_to_f32(uint16_t h) { return _cvtsh_ss(h); }
#else
static inline float f16_bits_to_f32(uint16_t h) {
    uint32_t sign = ((uint32_t)h & 0x8000u) << 16;
    uint32_t exp = (h >> 10) & 0x1fu;
    uint32_t mant = h & 0x3ffu;
    uint32_t x;
    if (exp == 0) {
        if (mant == 0) {
            x = sign;
        } else {
            int e = -1;
            do {
                mant <<= 1;
                e++;
            } while (!(mant & 0x400u));
            x = sign | ((uint32_t)(112 - e) << 23) | ((mant & 0x3ffu) << 13);
        }
    } else if (exp == 31) {
        x = sign | 0x7f800000u | (mant << 13);
    } else {
        x = sign | ((exp + 112) << 23) | (mant << 13);
    }
    float f;
    std::memcpy(&f, &x, 4);
    return f;
}
#endif

static inline uint16_t f32_to_bf16_bits(float f) {  // round-to-nearest-even
    uint32_t x;
    std::memcpy(&x, &f, 4);
    if ((x & 0x7fffffffu) > 0x7f800000u) return (uint16_t)((x >> 16) | 0x40u);
    uint32_t lsb = (x >> 16) & 1u;
    x += 0x7fffu + lsb;
    return (uint16_t)(x >> 16);
}
static inline float bf16_bits_to_f32(uint16_t h) {
    uint32_t x = (uint32_t)h << 16;
    float f;
    std::memcpy(&f, &x, 4);
    return f;
}

// monotone key of an f32 under IEEE total order (Rust f32::total_cmp): larger float -> larger key
static inline uint32_t f32_total_key(float f) {
    uint32_t x;
    std::memcpy(&x, &f, 4);
    return (x & 0x80000000u) ? ~x : (x | 0x80000000u);
}

static inline unsigned hw_threads(unsigned requested) {
    if (requested) return requested;
    unsigned n = std::thread::hardware_concurrency();
    return n ? n : 1;
}

// Dynamic-chunk parallel loop over [0, n). fn(begin, end, thread_id).
template <class F>
void parallel_for(uint64_t n, uint64_t chunk, unsigned n_threads, F&& fn) {
    n_threads = hw_threads(n_threads);
    if (n == 0) return;
    if (chunk == 0) chunk = 1;
    uint64_t n_chunks = (n + chunk - 1) / chunk;
    if (n_threads > n_chunks) n_threads = (unsigned)n_chunks;
    if (n_threads <= 1) {
        fn((uint64_t)0, n, 0u);
        return;
    }
    std::atomic<uint64_t> next{0};
    std::vector<std::thread> th;
    th.reserve(n_threads);
    for (unsigned t = 0; t < n_threads; ++t) {
        th.emplace_back([&, t]() {
            for (;;) {
                uint64_t c = next.fetch_add(1, std::memory_order_relaxed);
                if (c >= n_chunks) break;
                uint64_t b = c * chunk, e = std::min(n, b + chunk);
                fn(b, e, t);
            }
        });
    }
    for (auto& x : th) x.join();
}

// Static partition into exactly `parts` contiguous ranges; fn(part, begin, end).
template <class F>
void parallel_parts(uint64_t n, unsigned parts, F&& fn) {
    if (parts <= 1) {
        fn(0u, (uint64_t)0, n);
        return;
    }
    std::vector<std::thread> th;
    th.reserve(parts);
    for (unsigned p = 0; p < parts; ++p) {
        uint64_t b = n * p / parts, e = n * (p + 1) / parts;
        th.emplace_back([&, p, b, e]() { fn(p, b, e); });
    }
    for (auto& x : th) x.join();
}

// ---- counter-based RNG: splitmix64 stream keyed by (seed, stream id) ----
struct Rng {
    uint64_t s;
    Rng(uint64_t seed, uint64_t stream) {
        s = seed * 0x9E3779B97F4A7C15ull + stream * 0xD1B54A32D192ED03ull + 0x2545F4914F6CDD1Dull;
        next();
    }
    inline uint64_t next() {
        uint64_t z = (s += 0x9E3779B97F4A7C15ull);
        z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
        z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
        return z ^ (z >> 31);
    }
    inline uint32_t u32() { return (uint32_t)(next() >> 32); }
    inline double uniform() { return (double)(next() >> 11) * (1.0 / 9007199254740992.0); }  // [0,1)
    inline uint64_t below(uint64_t n) { return (uint64_t)(((__uint128_t)next() * n) >> 64); }
    inline double normal() {
        double u1 = uniform(), u2 = uniform();
        if (u1 < 1e-300) u1 = 1e-300;
        return std::sqrt(-2.0 * std::log(u1)) * std::cos(6.283185307179586 * u2);
    }
};

// Walker alias table for O(1) sampling from a discrete distribution.
struct Alias {
    std::vector<float> prob;
    std::vector<uint32_t> alias;
    void build(const std::vector<double>& w) {
        size_t n = w.size();
        prob.assign(n, 0.f);
        alias.assign(n, 0);
        double sum = 0;
        for (double x : w) sum += x;
        std::vector<double> p(n);
        std::vector<uint32_t> small, large;
        for (size_t i = 0; i < n; ++i) {
            p[i] = w[i] * n / sum;
            (p[i] < 1.0 ? small : large).push_back((uint32_t)i);
        }
        while (!small.empty() && !large.empty()) {
            uint32_t s = small.back(), l = large.back();
            small.pop_back();
            prob[s] = (float)p[s];
            alias[s] = l;
            p[l] = (p[l] + p[s]) - 1.0;
            if (p[l] < 1.0) {
                large.pop_back();
                small.push_back(l);
            }
        }
        for (uint32_t i : large) prob[i] = 1.f, alias[i] = i;
        for (uint32_t i : small) prob[i] = 1.f, alias[i] = i;
    }
    inline uint32_t sample(Rng& r) const {
        uint64_t x = r.next();
        uint32_t i = (uint32_t)(((__uint128_t)(x >> 24) * prob.size()) >> 40);
        float u = (float)(x & 0xffffff) * (1.0f / 16777216.0f);
        return u < prob[i] ? i : alias[i];
    }
};

}  // namespace shost
