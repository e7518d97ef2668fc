// Hot path C: demux record router (reference loop F:774-810, write_reads F:726-730).
// Per chunk pair: R2 headers -> packed keys (scan kernel, demux rule F:778) -> sink id by a
// device lookup table built from the results CSV -> stable partition of whole records of both
// mates into per-sink regions (radix sort of (sink, record#) is stable, then a prefix sum of
// record lengths in that order gives every record its output offset) -> record copy.
#pragma once
#include "common.cuh"

namespace frb {

struct RouteBufs {
    unsigned char *in1 = nullptr, *in2 = nullptr, *out1 = nullptr, *out2 = nullptr;
    size_t cap_bytes = 0;
    unsigned long long *key2 = nullptr, *off1 = nullptr, *off2 = nullptr;
    unsigned long long *len1 = nullptr, *len2 = nullptr, *pos1 = nullptr, *pos2 = nullptr;
    unsigned *sink = nullptr, *sink_sorted = nullptr, *idx = nullptr, *idx_sorted = nullptr;
    unsigned long long *sink_off1 = nullptr, *sink_off2 = nullptr;
    size_t cap_recs = 0;
    unsigned cap_sinks = 0;
};

inline void route_free(RouteBufs& b) {
    cudaFree(b.in1), cudaFree(b.in2), cudaFree(b.out1), cudaFree(b.out2);
    cudaFree(b.key2), cudaFree(b.off1), cudaFree(b.off2), cudaFree(b.len1), cudaFree(b.len2);
    cudaFree(b.pos1), cudaFree(b.pos2), cudaFree(b.sink), cudaFree(b.sink_sorted), cudaFree(b.idx);
    cudaFree(b.idx_sorted), cudaFree(b.sink_off1), cudaFree(b.sink_off2);
    b = RouteBufs{};
}

// results-table build: key -> sink id stored in Slot.count (keys are unique: the host keeps the
// last CSV row of a repeated key, as a Python dict does, F:660)
__global__ void __launch_bounds__(256) route_build_kernel(Slot* tab, unsigned long long mask,
                                                          const unsigned long long* __restrict__ keys,
                                                          const unsigned* __restrict__ sinks, unsigned long long n,
                                                          DevState* st) {
    const unsigned long long i = blockIdx.x * static_cast<unsigned long long>(blockDim.x) + threadIdx.x;
    if (i >= n) return;
    const unsigned long long key = keys[i];
    unsigned long long h = hash64(key) & mask;
    for (unsigned probe = 0; probe < kMaxProbe; ++probe) {
        const unsigned long long k = atomicCAS(&tab[h].key, kEmpty, key);
        if (k == kEmpty || k == key) {
            tab[h].count = sinks[i];
            return;
        }
        h = (h + 1) & mask;
    }
    raise_error(st, FRB_ERR_TABLE_FULL, key);
}

// sink[i] for pair i; the smallest i with an unknown key goes to *first_bad (atomicMin)
__global__ void __launch_bounds__(256) route_lookup_kernel(const Slot* __restrict__ tab, unsigned long long mask,
                                                           const unsigned long long* __restrict__ key2,
                                                           unsigned long long n, unsigned* __restrict__ sink,
                                                           unsigned* __restrict__ idx, unsigned long long* first_bad) {
    const unsigned long long i = blockIdx.x * static_cast<unsigned long long>(blockDim.x) + threadIdx.x;
    if (i >= n) return;
    const unsigned long long key = key2[i];
    unsigned long long h = hash64(key) & mask;
    unsigned s = 0xFFFFFFFFu;
    for (unsigned probe = 0; probe < kMaxProbe; ++probe) {
        const unsigned long long k = tab[h].key;
        if (k == key) {
            s = static_cast<unsigned>(tab[h].count);
            break;
        }
        if (k == kEmpty) break;
        h = (h + 1) & mask;
    }
    if (s == 0xFFFFFFFFu) {
        atomicMin(first_bad, i);
        s = 0;
    }
    sink[i] = s;
    idx[i] = static_cast<unsigned>(i);
}

// record lengths of both mates in sink-sorted order
__global__ void __launch_bounds__(256) route_len_kernel(const unsigned* __restrict__ idx_sorted,
                                                        const unsigned long long* __restrict__ off1,
                                                        const unsigned long long* __restrict__ off2,
                                                        unsigned long long n, unsigned long long* __restrict__ len1,
                                                        unsigned long long* __restrict__ len2) {
    const unsigned long long j = blockIdx.x * static_cast<unsigned long long>(blockDim.x) + threadIdx.x;
    if (j >= n) return;
    const unsigned i = idx_sorted[j];
    len1[j] = off1[i + 1] - off1[i];
    len2[j] = off2[i + 1] - off2[i];
}

// sink_off[s] = output offset of the first record of sink s (or of the next non-empty sink)
__global__ void __launch_bounds__(256) route_sink_off_kernel(const unsigned* __restrict__ sink_sorted,
                                                             const unsigned long long* __restrict__ pos1,
                                                             const unsigned long long* __restrict__ pos2,
                                                             const unsigned long long* __restrict__ len1,
                                                             const unsigned long long* __restrict__ len2,
                                                             unsigned long long n, unsigned n_sinks,
                                                             unsigned long long* __restrict__ so1,
                                                             unsigned long long* __restrict__ so2) {
    const unsigned s = blockIdx.x * blockDim.x + threadIdx.x;
    if (s > n_sinks) return;
    unsigned long long lo = 0, hi = n;  // first j with sink_sorted[j] >= s
    while (lo < hi) {
        const unsigned long long mid = (lo + hi) >> 1;
        if (sink_sorted[mid] < s) lo = mid + 1;
        else hi = mid;
    }
    if (lo < n) {
        so1[s] = pos1[lo];
        so2[s] = pos2[lo];
    } else {
        so1[s] = n ? pos1[n - 1] + len1[n - 1] : 0;
        so2[s] = n ? pos2[n - 1] + len2[n - 1] : 0;
    }
}

// one warp per (record, mate): byte copy into the sink region
__global__ void __launch_bounds__(256) route_copy_kernel(const unsigned* __restrict__ idx_sorted,
                                                         const unsigned long long* __restrict__ off,
                                                         const unsigned long long* __restrict__ pos,
                                                         const unsigned long long* __restrict__ len,
                                                         unsigned long long n, const unsigned char* __restrict__ in,
                                                         unsigned char* __restrict__ out) {
    const int lane = threadIdx.x & 31;
    const unsigned long long wstride = static_cast<unsigned long long>(gridDim.x) * (blockDim.x >> 5);
    for (unsigned long long j = blockIdx.x * static_cast<unsigned long long>(blockDim.x >> 5) + (threadIdx.x >> 5);
         j < n; j += wstride) {
        const unsigned char* src = in + off[idx_sorted[j]];
        unsigned char* dst = out + pos[j];
        const unsigned long long L = len[j];
        // head bytes up to 4-byte alignment of dst, then 4-byte words assembled from src bytes
        unsigned long long q = lane;
        const unsigned mis = static_cast<unsigned>(reinterpret_cast<uintptr_t>(dst) & 3u);
        const unsigned long long head = mis ? (4 - mis < L ? 4 - mis : L) : 0;
        if (q < head) dst[q] = src[q];
        const unsigned long long words = (L - head) >> 2;
        const unsigned char* s2 = src + head;
        unsigned* d4 = reinterpret_cast<unsigned*>(dst + head);
        const unsigned smis = static_cast<unsigned>(reinterpret_cast<uintptr_t>(s2) & 3u);
        const unsigned* s4 = reinterpret_cast<const unsigned*>(s2 - smis);
        for (unsigned long long wd = lane; wd < words; wd += 32) {
            unsigned v;
            if (smis == 0) {
                v = s4[wd];
            } else {
                v = __funnelshift_r(s4[wd], s4[wd + 1], 8 * smis);
            }
            d4[wd] = v;
        }
        const unsigned long long done = head + (words << 2);
        if (done + lane < L) dst[done + lane] = src[done + lane];
    }
}

}  // namespace frb
