// Device-side pieces of the exchange protocol (uem_exchange.cu has the description) shared with the region-max kernel,
// whose last CTA can carry the id part of a step's send (no separate launch between the region pass and its consumers).
#pragma once
#include "uem_common.cuh"
#include <limits.h>
#include <stddef.h>

namespace {

constexpr int kMaxWorld = 16;
constexpr int kMaxDepth = 4;
constexpr int kSendChunks = 6;
constexpr int kOffAcks = 1280, kOffSlots = 2048;
constexpr unsigned long long kSpinLimitNs = 2000000000ull;

struct XHeader {
    unsigned seq_send[kMaxDepth];
    unsigned seq_recv[kMaxDepth];
    unsigned arrive_all[kMaxDepth];
    unsigned fold_arrive[kMaxDepth];
    int status;
    unsigned region_done;   // arrival counter of the region-max kernel when it carries the id part of a send (zero between launches)
};
static_assert(sizeof(XHeader) <= kOffAcks, "header overflows its page");

struct Peers {
    char* base[kMaxWorld];
};

// slot geometry in 8-byte LL words
__host__ __device__ inline int64_t sum_words(int c, int k) { return (((int64_t)c * k) + 3) & ~(int64_t)3; }
__host__ __device__ inline int64_t tail_words(int c) { return (int64_t)(2 * c + 2) * 2; }   // int64 values as (lo, hi)
__host__ __device__ inline int64_t slot_bytes(int c, int k) { return ((sum_words(c, k) + tail_words(c)) * 8 + 127) & ~(int64_t)127; }

__device__ __forceinline__ unsigned ld_acquire_sys(const unsigned* p) {
    unsigned v;
    asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ void st_release_sys(unsigned* p, unsigned v) {
    asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
__device__ __forceinline__ unsigned long long globaltimer_ns() {
    unsigned long long t;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
    return t;
}
// bounded spin until *flag >= want; returns false on timeout
__device__ __forceinline__ bool spin_until(const unsigned* flag, unsigned want) {
    if ((int)(ld_acquire_sys(flag) - want) >= 0) return true;
    const unsigned long long t0 = globaltimer_ns();
    while ((int)(ld_acquire_sys(flag) - want) < 0) {
        __nanosleep(64);
        if (globaltimer_ns() - t0 > kSpinLimitNs) return false;
    }
    return true;
}
// LL word: one 8-byte transaction {payload, sequence number}
__device__ __forceinline__ void ll_store(char* dst, unsigned payload, unsigned seq) {
    asm volatile("st.volatile.global.v2.u32 [%0], {%1,%2};" ::"l"(dst), "r"(payload), "r"(seq) : "memory");
}
// polls (bounded) until the word carries `seq`; *ok = false on timeout
__device__ __forceinline__ unsigned ll_load(const char* src, unsigned seq, bool* ok) {
    unsigned v, f;
    asm volatile("ld.volatile.global.v2.u32 {%0,%1}, [%2];" : "=r"(v), "=r"(f) : "l"(src) : "memory");
    if (f == seq) return v;
    const unsigned long long t0 = globaltimer_ns();
    do {
        __nanosleep(32);
        asm volatile("ld.volatile.global.v2.u32 {%0,%1}, [%2];" : "=r"(v), "=r"(f) : "l"(src) : "memory");
        if (globaltimer_ns() - t0 > kSpinLimitNs) { *ok = false; return v; }
    } while (f != seq);
    return v;
}

__device__ __forceinline__ int64_t ll_load_i64(const char* src, unsigned seq, bool* ok) {
    const unsigned lo = ll_load(src, seq, ok), hi = ll_load(src + 8, seq, ok);
    return (int64_t)(((uint64_t)hi << 32) | lo);
}

// One element of the rank-ordered fold: the words of ALL ranks are requested first (independent loads: one memory latency
// instead of `world` in a row -- the slots were written over NVLink, so every first touch misses), then checked; a word
// whose sequence number is not there yet falls back to the bounded poll.  Same additions in the same order as a plain loop.
__device__ __forceinline__ void ll_raw(const char* src, unsigned& v, unsigned& f) {
    asm volatile("ld.volatile.global.v2.u32 {%0,%1}, [%2];" : "=r"(v), "=r"(f) : "l"(src) : "memory");
}
__device__ __forceinline__ void ll_fold_ranks(const char* sum0, const char* cnt0, int64_t sb, int world, unsigned seq, float* s_out,
                                              int64_t* n_out, bool* ok) {
    float s = 0.f;
    int64_t n = 0;
    for (int r0 = 0; r0 < world; r0 += 8) {
        unsigned v[8], vf[8], lo[8], lf[8], hi[8], hf[8];
#pragma unroll
        for (int u = 0; u < 8; ++u)
            if (r0 + u < world) {
                const int64_t off = (int64_t)(r0 + u) * sb;
                ll_raw(sum0 + off, v[u], vf[u]);
                ll_raw(cnt0 + off, lo[u], lf[u]);
                ll_raw(cnt0 + off + 8, hi[u], hf[u]);
            }
#pragma unroll
        for (int u = 0; u < 8; ++u)
            if (r0 + u < world) {
                const int r = r0 + u;
                const int64_t off = (int64_t)r * sb;
                if (vf[u] != seq) v[u] = ll_load(sum0 + off, seq, ok);
                if (lf[u] != seq) lo[u] = ll_load(cnt0 + off, seq, ok);
                if (hf[u] != seq) hi[u] = ll_load(cnt0 + off + 8, seq, ok);
                const float x = __uint_as_float(v[u]);
                s = r ? s + x : x;
                n += (int64_t)(((uint64_t)hi[u] << 32) | lo[u]);
            }
    }
    *s_out = s;
    *n_out = n;
}

// The id part of a step's send, issued by ONE CTA (every thread calls this; blockDim >= 32): the rank-local max superpixel id
// goes into slot [slot][rank] of every rank, tagged with the slot's current sequence number like xchg_send_kernel's parts = 2;
// with global_id_out the first warp then polls the other ranks' ids of the same step and leaves the batch-global id
// (alignment.py:241).  world == 0: disabled.
struct RegionXchg {
    Peers peers;
    int rank, world, slot, c, k;
    long long* global_id_out;
};
__device__ __forceinline__ void xchg_send_id_from_cta(const RegionXchg& x, long long id) {
    char* const mine = x.peers.base[x.rank];
    XHeader* hdr = reinterpret_cast<XHeader*>(mine);
    const unsigned my = *reinterpret_cast<volatile unsigned*>(&hdr->seq_send[x.slot]) + 1u;
    const int64_t sb = slot_bytes(x.c, x.k);
    const int64_t id_off = kOffSlots + sum_words(x.c, x.k) * 8 + (int64_t)(2 * x.c + 1) * 16;
    if ((int)threadIdx.x < x.world) {
        const int p = threadIdx.x;
        // peer p must have folded the previous contents of this slot (its ack lands in MY region)
        const unsigned* ack = reinterpret_cast<const unsigned*>(mine + kOffAcks) + x.slot * kMaxWorld + p;
        if (!spin_until(ack, my - 1u)) atomicOr(&hdr->status, 8);
        char* dst = x.peers.base[p] + id_off + ((int64_t)x.slot * x.world + x.rank) * sb;
        ll_store(dst, (unsigned)((uint64_t)id & 0xffffffffu), my);
        ll_store(dst + 8, (unsigned)((uint64_t)id >> 32), my);
    }
    if (x.global_id_out && threadIdx.x < 32) {
        const int r = threadIdx.x;
        long long g = LLONG_MIN;
        if (r < x.world) {
            bool ok = true;
            g = ll_load_i64(mine + id_off + ((int64_t)x.slot * x.world + r) * sb, my, &ok);
            if (!ok) atomicOr(&hdr->status, 8);
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            const long long other = __shfl_xor_sync(0xffffffffu, g, o);
            g = other > g ? other : g;
        }
        if (r == 0) x.global_id_out[0] = g;
    }
}

inline int fill_peers(Peers* P, const void* const* peer_regions, int rank, int world, int depth, int slot, const char* who) {
    UEM_REQUIRE(peer_regions && world >= 1 && world <= kMaxWorld && rank >= 0 && rank < world, "%s: bad rank/world (%d/%d, at most %d ranks)",
                who, rank, world, kMaxWorld);
    UEM_REQUIRE(depth >= 1 && depth <= kMaxDepth && slot >= 0 && slot < depth, "%s: bad slot/depth (%d/%d, depth at most %d)", who, slot,
                depth, kMaxDepth);
    for (int r = 0; r < kMaxWorld; ++r) P->base[r] = r < world ? (char*)peer_regions[r] : nullptr;
    for (int r = 0; r < world; ++r) UEM_REQUIRE(P->base[r], "%s: peer region %d is NULL", who, r);
    return 0;
}

}  // namespace
