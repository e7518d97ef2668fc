// Second half of the speculative scan path (scan_ws_kernel.cuh, WS_LEAN): the line phase of every tile by COUNT,
// as the reference defines it (F:161-169 takes every 4th line from the start of the file), checked against the
// phase the kernel guessed from the text.
//
//   spec_sum_kernel     newlines per block of 1024 tiles
//   spec_scan_kernel    one block: exclusive scan of the block sums; lines and reads of the whole chunk
//                       (n_reads, line_carry of the file)
//   spec_verify_kernel  per tile: newlines in front of it -> its line phase; a tile without a guess goes to the
//                       redo list, a tile whose guess differs from its phase marks the chunk bad (spec_bad: the
//                       negate pass and scan_redo_kernel then redo the whole chunk by count); first read ordinal
//                       of the tile (tile_first, what turns composite positions into read ordinals when the
//                       file's table is compacted); status[] becomes the inclusive newline prefix that
//                       scan_redo_kernel reads
//   spec_reset_first_kernel  bad chunk only: forget every `first` the chunk may have set
#pragma once
#include "scan_ws_kernel.cuh"

namespace frb {

constexpr int kVerifyTiles = 1024;  // tiles per block = threads per block

__device__ __forceinline__ unsigned info_total(unsigned long long w) { return static_cast<unsigned>(w) & 0xFFFFFu; }
__device__ __forceinline__ unsigned info_vnl(unsigned long long w) { return static_cast<unsigned>(w >> 20) & 1u; }
__device__ __forceinline__ unsigned info_guess(unsigned long long w) { return static_cast<unsigned>(w >> 24) & 0xFFu; }

// inclusive scan over the 1024 threads of a block
__device__ __forceinline__ unsigned long long block_scan_incl(unsigned long long v, unsigned long long* s_warp,
                                                              unsigned long long* block_total) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
        const unsigned long long n = __shfl_up_sync(0xFFFFFFFFu, v, d);
        if (lane >= d) v += n;
    }
    if (lane == 31) s_warp[warp] = v;
    __syncthreads();
    if (warp == 0) {
        unsigned long long w = s_warp[lane];
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) {
            const unsigned long long n = __shfl_up_sync(0xFFFFFFFFu, w, d);
            if (lane >= d) w += n;
        }
        s_warp[lane] = w;
    }
    __syncthreads();
    if (block_total) *block_total = s_warp[31];
    const unsigned long long before = warp ? s_warp[warp - 1] : 0ULL;
    __syncthreads();
    return v + before;
}

__global__ void __launch_bounds__(kVerifyTiles) spec_sum_kernel(const unsigned long long* __restrict__ info,
                                                                unsigned n_tiles, unsigned long long* block_sums) {
    __shared__ unsigned long long s_warp[32];
    const unsigned t = blockIdx.x * kVerifyTiles + threadIdx.x;
    unsigned long long total;
    block_scan_incl(t < n_tiles ? info_total(info[t]) : 0u, s_warp, &total);
    if (threadIdx.x == 0) block_sums[blockIdx.x] = total;
}

__global__ void __launch_bounds__(kVerifyTiles) spec_scan_kernel(unsigned long long* block_sums, unsigned n_blocks,
                                                                 const unsigned long long* __restrict__ info,
                                                                 unsigned n_tiles, unsigned long long line_base,
                                                                 int use_carry, DevState* st) {
    __shared__ unsigned long long s_warp[32];
    unsigned long long carry = 0;
    for (unsigned b0 = 0; b0 < n_blocks; b0 += kVerifyTiles) {
        const unsigned b = b0 + threadIdx.x;
        const unsigned long long v = b < n_blocks ? block_sums[b] : 0ULL;
        unsigned long long total;
        const unsigned long long incl = block_scan_incl(v, s_warp, &total);
        if (b < n_blocks) block_sums[b] = carry + incl - v;  // newlines in front of the block
        carry += total;
    }
    if (threadIdx.x == 0) {
        const unsigned long long L0 = use_carry ? st->line_carry : line_base;
        const unsigned long long lines = carry + info_vnl(info[n_tiles - 1]);  // a last line without '\n' is a line
        st->chunk_l0 = L0;
        st->n_reads += ((L0 + lines + 3) >> 2) - ((L0 + 3) >> 2);  // header lines = lines k with k % 4 == 0
        st->line_carry = L0 + lines;
    }
}

__global__ void __launch_bounds__(kVerifyTiles) spec_verify_kernel(unsigned long long* info, unsigned n_tiles,
                                                                   const unsigned long long* __restrict__ block_excl,
                                                                   DevState* st, unsigned long long* tile_first,
                                                                   unsigned int* redo) {
    __shared__ unsigned long long s_warp[32];
    const unsigned t = blockIdx.x * kVerifyTiles + threadIdx.x;
    const unsigned long long w = t < n_tiles ? info[t] : 0ULL;
    const unsigned total = info_total(w);
    const unsigned long long incl = block_excl[blockIdx.x] + block_scan_incl(total, s_warp, nullptr);
    if (t >= n_tiles) return;
    const unsigned long long K0 = st->chunk_l0 + incl - total;  // index of the tile's first line end
    tile_first[t] = (K0 + 3) >> 2;
    info[t] = kFlagInc | incl;
    const unsigned guess = info_guess(w);
    const unsigned lines = total + info_vnl(w);
    if (guess == kNoGuess) {
        if (lines) redo[atomicAdd(&st->redo_n, 1ULL)] = t;
    } else if (guess != static_cast<unsigned>((4 - (K0 & 3)) & 3)) {
        st->spec_bad = 1;
    }
}

__global__ void __launch_bounds__(256) spec_reset_first_kernel(Slot* tab, unsigned long long cap,
                                                               unsigned long long threshold, const DevState* st) {
    if (!st->spec_bad) return;
    const unsigned long long stride = static_cast<unsigned long long>(gridDim.x) * blockDim.x;
    for (unsigned long long i = blockIdx.x * static_cast<unsigned long long>(blockDim.x) + threadIdx.x; i < cap;
         i += stride)
        if (tab[i].key != kEmpty && tab[i].first >= threshold) tab[i].first = ~0ULL;
}

}  // namespace frb
