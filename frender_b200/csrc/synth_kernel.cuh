// Device twin of frender_b200/synth.py: counter-based synthetic FASTQ (bench / tests only).
// Byte-identical to the numpy generator for any slice [g0, g1) -- checked in tests.
#pragma once
#include "common.cuh"

namespace frb {

constexpr unsigned long long kGold = 0x9E3779B97F4A7C15ULL;
constexpr int kJKind = 0, kJPartner = 1, kJRand7 = 2, kJRand5 = 3, kJErr0 = 4, kJCoord = 16, kJSeq1 = 17, kJSeq2 = 22;
constexpr int kSynthPrefixLen = 21;  // "@A00123:45:HXXXXXXXX:"

struct SynthArgs {
    unsigned long long seed;
    const unsigned* emit_i7;          // [S] 2-bit codes, base p at bits 2p
    const unsigned* emit_i5;          // [S]
    const unsigned long long* cdf;    // [S] cumulative thresholds (2^32 units)
    unsigned long long rand_t, hop_t;
    unsigned long long g0;
    unsigned long long n;
    unsigned l1, l2, n_samples, lane, read_len, sub_t, n_t;
    int read_no;
};

__host__ __device__ __forceinline__ unsigned long long mix64(unsigned long long z) {
    z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ULL;
    z = (z ^ (z >> 27)) * 0x94D049BB133111EBULL;
    return z ^ (z >> 31);
}
__host__ __device__ __forceinline__ unsigned long long draw(unsigned long long seed, unsigned long long g, int j) {
    return mix64(seed + (g * 32ULL + static_cast<unsigned long long>(j + 1)) * kGold);
}
__device__ __forceinline__ unsigned digits_of(unsigned v) {
    return v >= 100000 ? 6 : v >= 10000 ? 5 : v >= 1000 ? 4 : v >= 100 ? 3 : v >= 10 ? 2 : 1;
}
__device__ __forceinline__ void coords(const SynthArgs& a, unsigned long long g, unsigned* x, unsigned* y,
                                       unsigned* tile) {
    const unsigned long long w = draw(a.seed, g, kJCoord);
    *x = 1000u + static_cast<unsigned>((w & 0xFFFFFFFFULL) % 31000ULL);
    *y = 1000u + static_cast<unsigned>((w >> 32) % 199000ULL);
    *tile = 1101u + static_cast<unsigned>(g % 78ULL);
}
__device__ __forceinline__ unsigned header_len(const SynthArgs& a, unsigned x, unsigned y) {
    return kSynthPrefixLen + 1 + 1 + 4 + 1 + digits_of(x) + 1 + digits_of(y) + 1 + 1 + 5 + a.l1 +
           (a.l2 ? 1 + a.l2 : 0) + 1;
}

__global__ void __launch_bounds__(256) synth_len_kernel(const SynthArgs a, unsigned long long* len) {
    const unsigned long long i = blockIdx.x * static_cast<unsigned long long>(blockDim.x) + threadIdx.x;
    if (i >= a.n) return;
    unsigned x, y, tile;
    coords(a, a.g0 + i, &x, &y, &tile);
    len[i] = header_len(a, x, y) + 2ULL * a.read_len + 4ULL;
}

__device__ __forceinline__ unsigned pick_sample(const SynthArgs& a, unsigned long long u) {
    unsigned lo = 0, hi = a.n_samples;  // first index with cdf[idx] > u  (searchsorted side='right')
    while (lo < hi) {
        const unsigned mid = (lo + hi) >> 1;
        if (a.cdf[mid] <= u) lo = mid + 1;
        else hi = mid;
    }
    return lo < a.n_samples ? lo : a.n_samples - 1;
}

// one warp per record
__global__ void __launch_bounds__(256) synth_write_kernel(const SynthArgs a, const unsigned long long* __restrict__ off,
                                                          unsigned char* __restrict__ out) {
    __shared__ unsigned char s_hdr[8][96];
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    const unsigned long long wstride = static_cast<unsigned long long>(gridDim.x) * 8ULL;
    for (unsigned long long i = blockIdx.x * 8ULL + w; i < a.n; i += wstride) {
        const unsigned long long g = a.g0 + i;
        unsigned x, y, tile;
        coords(a, g, &x, &y, &tile);
        const unsigned hlen = header_len(a, x, y);
        if (lane == 0) {
            unsigned char* h = s_hdr[w];
            const char* pre = "@A00123:45:HXXXXXXXX:";
            unsigned p = 0;
            for (int k = 0; k < kSynthPrefixLen; ++k) h[p++] = pre[k];
            h[p++] = '0' + a.lane;
            h[p++] = ':';
            for (int d = 1000; d > 0; d /= 10) h[p++] = '0' + (tile / d) % 10;
            h[p++] = ':';
            for (int k = digits_of(x) - 1, d = 1; k >= 0; --k) {
                d = 1;
                for (int q = 0; q < k; ++q) d *= 10;
                h[p++] = '0' + (x / d) % 10;
            }
            h[p++] = ':';
            for (int k = digits_of(y) - 1, d = 1; k >= 0; --k) {
                d = 1;
                for (int q = 0; q < k; ++q) d *= 10;
                h[p++] = '0' + (y / d) % 10;
            }
            h[p++] = ' ';
            h[p++] = '0' + a.read_no;
            h[p++] = ':', h[p++] = 'N', h[p++] = ':', h[p++] = '0', h[p++] = ':';
            // index symbols
            const unsigned long long w0 = draw(a.seed, g, kJKind);
            const unsigned sample = pick_sample(a, w0 & 0xFFFFFFFFULL);
            const unsigned long long kind = w0 >> 32;
            const bool is_rand = kind < a.rand_t;
            const bool is_hop = !is_rand && kind < a.rand_t + a.hop_t;
            const unsigned partner = pick_sample(a, draw(a.seed, g, kJPartner) & 0xFFFFFFFFULL);
            unsigned long long c7 = a.emit_i7[sample];
            unsigned long long c5 = a.l2 ? a.emit_i5[is_hop ? partner : sample] : 0;
            if (is_rand) {
                c7 = draw(a.seed, g, kJRand7);
                c5 = draw(a.seed, g, kJRand5);
            }
            for (unsigned q = 0; q < a.l1 + a.l2; ++q) {
                unsigned base = q < a.l1 ? static_cast<unsigned>(c7 >> (2 * q)) & 3u
                                         : static_cast<unsigned>(c5 >> (2 * (q - a.l1))) & 3u;
                const unsigned long long word = draw(a.seed, g, kJErr0 + q / 2) >> (32 * (q & 1));
                const unsigned e16 = static_cast<unsigned>(word) & 0xFFFFu;
                const unsigned s16 = static_cast<unsigned>(word >> 16) & 0xFFFFu;
                if (e16 < a.sub_t) base = (base + 1 + s16 % 3) & 3u;
                else if (e16 < a.sub_t + a.n_t) base = 4;
                if (q == a.l1) h[p++] = '+';
                h[p++] = "ACGTN"[base];
            }
            h[p++] = '\n';
        }
        __syncwarp();
        const unsigned total = hlen + 2 * a.read_len + 4;
        unsigned char* dst = out + off[i];
        const int jseq = a.read_no == 1 ? kJSeq1 : kJSeq2;
        for (unsigned q = lane; q < total; q += 32) {
            unsigned char c;
            if (q < hlen) {
                c = s_hdr[w][q];
            } else {
                const unsigned r = q - hlen;
                if (r < a.read_len) c = "ACGT"[(draw(a.seed, g, jseq + r / 32) >> (2 * (r & 31))) & 3ULL];
                else if (r == a.read_len || r == a.read_len + 2 || r == 2 * a.read_len + 3) c = '\n';
                else if (r == a.read_len + 1) c = '+';
                else c = 'F';
            }
            dst[q] = c;
        }
        __syncwarp();
    }
}

}  // namespace frb
