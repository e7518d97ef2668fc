// Warp-specialised form of the scan kernel (hot path A).  Same arithmetic, same results and the
// same helpers as scan_kernel.cuh; the difference is the schedule inside a CTA.
//
//   warps 0-3  COUNTERS    for tile i: wait for its bytes (TMA mbarrier), 256 bytes per thread ->
//                          newline mask, group scan, ordered newline-position list in shared memory,
//                          publish the tile's newline count, signal `counted[stage]`.
//   warps 4-6  EXTRACTORS  for tile i: one thread per header line: key extraction, warp fold, deferred
//                          table update; thread 0 then refills the stage (bulk copy of tile i+3).
//   warp  7    HELPER      runs one tile ahead of the extractors: start of the line that straddles the
//                          tile start, look-back over the published counts (line number of the tile's
//                          first newline -- the only wait on other CTAs), ticket for the next refill.
//
// Counters and the parser side only meet through shared-memory mbarriers (full -> counted per stage);
// helper and extractors meet once per tile at a named barrier.  Counting tile i+1, preparing tile i+1
// and extracting tile i overlap, and a helper that waits on another CTA's count never stops its own
// CTA's counters or extractors.
#pragma once
#include "scan_kernel.cuh"

namespace frb {

constexpr int kWsGroup = 128;                 // counter threads
constexpr unsigned kNoTile = 0xFFFFFFFFu;

// Tile geometry.  SEG = 16-byte segments per counter thread, PW = parser warps, CTAS = resident CTAs
// per SM the shared memory and registers are budgeted for.
template <int SEG, int PW, int CTAS, int NLCAP>
struct WsGeom {
    static constexpr int seg = SEG;
    static constexpr int per_thread = SEG * 16;
    static constexpr int tile = kWsGroup * per_thread;
    static constexpr int buf = tile + kHalo;
    static constexpr int nl_cap = NLCAP;
    static constexpr int pwarps = PW;
    static constexpr int pgroup = PW * 32;
    static constexpr int threads = kWsGroup + pgroup;
    static constexpr int ctas = CTAS;
    static constexpr int maxreg = (65536 / (CTAS * threads)) / 8 * 8;  // per-thread budget that keeps CTAS resident
    static constexpr int smem = kStages * buf + kStages * nl_cap * (int)sizeof(uint16_t);
};
using WsWide = WsGeom<16, 4, 2, 2048>;   // 32 KiB tiles, 2 x 8 warps per SM
using WsDense = WsGeom<10, 3, 3, 1024>;  // 20 KiB tiles, 3 x 7 warps per SM
constexpr int kWsTile = WsWide::tile;
constexpr int kWsThreads = WsWide::threads;
constexpr int kWsSmem = WsWide::smem;

template <int N>
__device__ __forceinline__ void group_sync(int id) {
    asm volatile("bar.sync %0, %1;" ::"r"(id), "n"(N) : "memory");
}

template <class G>
__global__ void __launch_bounds__(G::threads) __maxnreg__(G::maxreg) scan_ws_kernel(const ScanArgs a) {
    constexpr int kWsTile = G::tile, kWsBuf = G::buf, kWsNlCap = G::nl_cap, kWsPerThread = G::per_thread;
    constexpr int kPGroup = G::pgroup, kPWarps = G::pwarps;
    extern __shared__ __align__(128) unsigned char smem[];
    uint16_t* const s_nl = reinterpret_cast<uint16_t*>(smem + kStages * kWsBuf);
    __shared__ __align__(8) unsigned long long s_full[kStages], s_counted[kStages];
    __shared__ unsigned s_tile[kStages], s_total[kStages], s_valid[kStages], s_vnl[kStages];
    __shared__ unsigned s_cwarp[kWsGroup / 32];
    __shared__ unsigned long long s_prefix[2];
    __shared__ volatile unsigned long long s_ticket[2];
    __shared__ unsigned s_halo_start[2];
    __shared__ unsigned char s_lut[256];

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const unsigned long long L0 =
        a.use_carry ? *reinterpret_cast<volatile unsigned long long*>(&a.st->line_carry) : a.line_base;
    const unsigned long long chunk_first_read = (L0 + 3) >> 2;
    volatile unsigned long long* status = a.status + 1;

    for (int i = tid; i < 256; i += G::threads) s_lut[i] = lut_entry(i, a.rule);
    if (tid == 0) {
        s_ticket[0] = ~0ULL, s_ticket[1] = ~0ULL;
#pragma unroll
        for (int i = 0; i < kStages; ++i) {
            mbar_init(&s_full[i], 1);
            mbar_init(&s_counted[i], 1);
        }
        mbar_fence_init();
    }
    __syncthreads();

    if (warp < kWsGroup / 32) {
        // =============================== COUNTERS ================================================
        const int ct = tid;
        unsigned full_parity = 0;  // bit s
        long long tm = (a.timing && ct == 0) ? clock64() : 0;
        unsigned long long t_wait = 0, t_work = 0;
        for (unsigned i = 0;; ++i) {
            const int s = i % kStages;
            mbar_wait(&s_full[s], (full_parity >> s) & 1u);
            full_parity ^= 1u << s;
            if (a.timing && ct == 0) { const long long now = clock64(); t_wait += now - tm; tm = now; }
            const unsigned t = s_tile[s];
            if (t == kNoTile) {
                if (ct == 0) mbar_arrive(&s_counted[s]);
                break;
            }
            unsigned char* const buf = smem + s * kWsBuf;
            const unsigned long long tile_off = static_cast<unsigned long long>(t) * kWsTile;
            const unsigned long long left = a.nbytes - tile_off;
            const unsigned valid = static_cast<unsigned>(left < kWsTile ? left : kWsTile);
            {   // bytes past the last 16-byte multiple of the bulk copy (final tile only)
                const unsigned halo = t ? kHalo : 0;
                const unsigned avail = valid + halo, bulk = avail & ~15u;
                if (avail != bulk) {
                    if (ct < static_cast<int>(avail - bulk))
                        buf[(kHalo - halo) + bulk + ct] = a.data[tile_off - halo + bulk + ct];
                    group_sync<kWsGroup>(1);
                }
            }
            // newline masks of this thread's bytes as 32-bit words in byte order (bit k of word q = byte
            // 32q + k); segments are read rotated so that the eight lanes of a 16-byte load phase hit
            // eight different bank groups.
            constexpr int kWords = kWsPerThread / 32;
            unsigned w[kWords];
            const uint4* t4 = reinterpret_cast<const uint4*>(buf + kHalo) + ct * G::seg;
            if constexpr (G::seg == 16) {
                unsigned long long m00 = 0, m01 = 0, m10 = 0, m11 = 0;
#pragma unroll
                for (int j = 0; j < 8; ++j) {
                    const int r = (j + ct) & 7;
                    const unsigned long long ma = newline_mask16(t4[r]);
                    const unsigned long long mb = newline_mask16(t4[8 + r]);
                    const int sh = (r & 3) * 16;
                    if (r < 4) m00 |= ma << sh, m10 |= mb << sh;
                    else m01 |= ma << sh, m11 |= mb << sh;
                }
                w[0] = static_cast<unsigned>(m00), w[1] = static_cast<unsigned>(m00 >> 32);
                w[2] = static_cast<unsigned>(m01), w[3] = static_cast<unsigned>(m01 >> 32);
                w[4] = static_cast<unsigned>(m10), w[5] = static_cast<unsigned>(m10 >> 32);
                w[6] = static_cast<unsigned>(m11), w[7] = static_cast<unsigned>(m11 >> 32);
            } else {
                // thread stride 160 bytes = 10 bank groups: lanes 4-7 of a phase start one segment later
                static_assert(G::seg == 10, "rotation below is written for 10 segments per thread");
                const bool rot = (ct >> 2) & 1;
                unsigned m[G::seg];
#pragma unroll
                for (int j = 0; j < G::seg; ++j) {
                    const int r = (j == G::seg - 1) ? (rot ? 0 : j) : j + (rot ? 1 : 0);
                    m[j] = newline_mask16(t4[r]);
                }
#pragma unroll
                for (int q = 0; q < kWords; ++q) {
                    const unsigned lo = rot ? m[(2 * q + G::seg - 1) % G::seg] : m[2 * q];
                    const unsigned hi = rot ? m[2 * q] : m[2 * q + 1];
                    w[q] = lo | (hi << 16);
                }
            }
            {
                const int nv = static_cast<int>(valid) - ct * kWsPerThread;
                if (nv < kWsPerThread) {
#pragma unroll
                    for (int q = 0; q < kWords; ++q) {
                        const int n = nv - 32 * q;
                        w[q] = n <= 0 ? 0u : (n >= 32 ? w[q] : (w[q] & ((1u << n) - 1u)));
                    }
                }
            }
            unsigned cnt = 0;
#pragma unroll
            for (int q = 0; q < kWords; ++q) cnt += __popc(w[q]);
            unsigned incl = cnt;
#pragma unroll
            for (int d = 1; d < 32; d <<= 1) {
                const unsigned n = __shfl_up_sync(0xFFFFFFFFu, incl, d);
                if (lane >= d) incl += n;
            }
            if (lane == 31) s_cwarp[warp] = incl;
            group_sync<kWsGroup>(1);
            unsigned wbase = 0, total = 0;
#pragma unroll
            for (int k = 0; k < kWsGroup / 32; ++k) {
                const unsigned v = s_cwarp[k];
                if (k < warp) wbase += v;
                total += v;
            }
            // a last line without '\n' still is a line (F:161 iterates it; F:169 rstrip)
            const unsigned vnl = (t == a.n_tiles - 1 && valid > 0 && buf[kHalo + valid - 1] != '\n') ? 1u : 0u;
            {   // ordered list of newline positions; line numbers come later, from the look-back
                uint16_t* const nl = s_nl + s * kWsNlCap;
                unsigned idx = wbase + incl - cnt;
                unsigned pos0 = kHalo + ct * kWsPerThread;
#pragma unroll
                for (int q = 0; q < kWords; ++q) {
                    unsigned m = w[q];
                    while (m) {
                        if (idx < static_cast<unsigned>(kWsNlCap)) nl[idx] = static_cast<uint16_t>(pos0 + (__ffs(m) - 1));
                        m &= m - 1;
                        ++idx;
                    }
                    pos0 += 32;
                }
                if (ct == 0 && vnl && total < static_cast<unsigned>(kWsNlCap)) nl[total] = static_cast<uint16_t>(kHalo + valid);
            }
            if (ct == 0) {
                s_total[s] = total, s_valid[s] = valid, s_vnl[s] = vnl;
                status[t] = (t == 0 ? kFlagInc : kFlagAgg) | total;
            }
            group_sync<kWsGroup>(1);  // list + meta complete (also protects s_cwarp)
            if (ct == 0) mbar_arrive(&s_counted[s]);
            if (a.timing && ct == 0) { const long long now = clock64(); t_work += now - tm; tm = now; }
        }
        if (a.timing && ct == 0) atomicAdd(&a.timing[0], t_wait), atomicAdd(&a.timing[1], t_work);
    } else {
        // =============================== PARSERS =================================================
        const int pt = tid - kWsGroup;
        const int pwarp = warp - kWsGroup / 32;
        constexpr int kLast = kPWarps - 1;  // helper warp: halo scan, look-back, tickets
        // deferred table update, three steps (see scan_kernel.cuh)
        unsigned long long p_key = 0, p_pos = 0, p_slot = 0, p_seen = 0;
        unsigned p_cnt = 0;
        unsigned long long q_key = 0, q_pos = 0, q_slot = 0, q_old = 0;
        unsigned q_cnt = 0;
        auto bump = [&](unsigned long long slot, unsigned cnt, unsigned long long pos) {
            atomicAdd(&a.table[slot].count, static_cast<unsigned long long>(cnt));
            atomicMin(&a.table[slot].first, pos);
        };
        auto finish_pending = [&]() {
            if (q_cnt) {
                if (q_old == kEmpty) {
                    atomicAdd(&a.st->occupied, 1ULL);
                    bump(q_slot, q_cnt, q_pos);
                } else if (q_old == q_key) {
                    bump(q_slot, q_cnt, q_pos);
                } else {
                    table_add(a.table, a.table_mask, q_key, q_cnt, q_pos, &a.st->occupied, a.st);
                }
                q_cnt = 0;
            }
            if (p_cnt) {
                if (p_seen == p_key) {
                    bump(p_slot, p_cnt, p_pos);
                } else if (p_seen == kEmpty) {
                    q_key = p_key, q_pos = p_pos, q_slot = p_slot, q_cnt = p_cnt;
                    q_old = atomicCAS(&a.table[p_slot].key, kEmpty, p_key);
                } else {
                    table_add(a.table, a.table_mask, p_key, p_cnt, p_pos, &a.st->occupied, a.st);
                }
                p_cnt = 0;
            }
        };
        auto emit = [&](unsigned long long o, unsigned long long key, unsigned long long start_g) {
            const unsigned long long slot = o - chunk_first_read;
            if (slot < a.out_cap) {
                if (a.keys_out) a.keys_out[slot] = key;
                if (a.rec_off_out) a.rec_off_out[slot] = start_g;
            }
        };
        // A stage is refilled the moment the extractors are done with it; the ticket for it was drawn
        // by the helper warp a tile earlier.  The counters never wait for anything but bytes.
        auto issue = [&](int s, unsigned ticket) {  // ticket -> stage s, start its bulk copy
            const unsigned t = ticket < a.n_tiles ? ticket : kNoTile;
            s_tile[s] = t;
            if (t == kNoTile) {
                mbar_arrive(&s_full[s]);
                return;
            }
            const unsigned long long off = static_cast<unsigned long long>(t) * kWsTile;
            const unsigned halo = t ? kHalo : 0;
            const unsigned long long left = a.nbytes - off;
            const unsigned avail = static_cast<unsigned>(left < kWsTile ? left : kWsTile) + halo;
            const unsigned bulk = avail & ~15u;
            if (bulk) {
                mbar_expect_tx(&s_full[s], bulk);
                bulk_g2s(smem + s * kWsBuf + (kHalo - halo), a.data + off - halo, bulk, &s_full[s]);
            } else {
                mbar_arrive(&s_full[s]);
            }
        };
        unsigned counted_parity = 0;
        if (pwarp == kLast) {
            // ------------------------- helper warp: runs one tile ahead of the extractors -------------
            // For tile i: start of the line that straddles the tile start (halo scan), line number of the
            // tile's first newline (look-back, the only place a CTA waits on other CTAs), then the ticket
            // for the refill of this tile's stage.  All of it overlaps the key extraction of tile i-1.
            if (lane == 0) {
                for (int s = 0; s < kStages; ++s) issue(s, static_cast<unsigned>(atomicAdd(&a.status[0], 1ULL)));
            }
            long long tm = (a.timing && lane == 0) ? clock64() : 0;
            unsigned long long h_wait = 0, h_look = 0, h_bar = 0;
            auto htick = [&](unsigned long long& acc) {
                if (a.timing && lane == 0) { const long long now = clock64(); acc += now - tm; tm = now; }
            };
            for (unsigned i = 0;; ++i) {
                const int s = i % kStages;
                mbar_wait(&s_counted[s], (counted_parity >> s) & 1u);
                counted_parity ^= 1u << s;
                htick(h_wait);
                const unsigned t = s_tile[s];
                if (t != kNoTile) {
                    if (t == 0) {
                        if (lane == 0) s_halo_start[i & 1] = kHalo;
                    } else {
                        const unsigned m = newline_mask16(reinterpret_cast<const uint4*>(smem + s * kWsBuf)[lane]);
                        const unsigned any = __ballot_sync(0xFFFFFFFFu, m != 0);
                        if (any == 0) {
                            if (lane == 0) s_halo_start[i & 1] = kUnknown;
                        } else if (lane == 31 - __clz(any)) {
                            s_halo_start[i & 1] = lane * 16 + (31 - __clz(m)) + 1;
                        }
                    }
                    unsigned long long excl;
                    tile_prefix<true>(status, t, s_total[s], lane, &excl);
                    if (lane == 0) s_prefix[i & 1] = excl;
                }
                htick(h_look);
                group_sync<kPGroup>(2);  // tile i is prepared AND the extractors have finished tile i-1
                htick(h_bar);
                if (t == kNoTile) break;
                if (lane == 0) {  // ticket for the refill at the end of tile i, tagged with i
                    const unsigned long long ticket = atomicAdd(&a.status[0], 1ULL) & 0xFFFFFFFFULL;
                    s_ticket[i & 1] = (static_cast<unsigned long long>(i) << 32) | ticket;
                }
            }
            if (a.timing && lane == 0) {
                atomicAdd(&a.timing[3], h_look), atomicAdd(&a.timing[4], h_wait), atomicAdd(&a.timing[5], h_bar);
            }
        } else {
            // ------------------------- extractor warps -------------------------------------------------
            constexpr int kExt = kPGroup - 32;
            unsigned long long my_reads = 0;
            long long tm = (a.timing && pt == 0) ? clock64() : 0;
            unsigned long long tp[8] = {0, 0, 0, 0, 0, 0, 0, 0};
            auto tick = [&](int k) {
                if (a.timing && pt == 0) { const long long now = clock64(); tp[k] += now - tm; tm = now; }
            };
            for (unsigned i = 0;; ++i) {
                const int s = i % kStages;
                group_sync<kPGroup>(2);
                mbar_wait(&s_counted[s], (counted_parity >> s) & 1u);  // complete already: the helper saw it
                counted_parity ^= 1u << s;
                tick(0);
                const unsigned t = s_tile[s];
                if (t == kNoTile) break;
                unsigned char* const buf = smem + s * kWsBuf;
                const unsigned total = s_total[s], valid = s_valid[s], vnl = s_vnl[s];
                const unsigned long long tile_off = static_cast<unsigned long long>(t) * kWsTile;
                const uint16_t* const nl = s_nl + s * kWsNlCap;
                const unsigned long long K0 = L0 + s_prefix[i & 1];
                const unsigned halo_start = s_halo_start[i & 1];
                const unsigned long long o_first = (K0 + 3) >> 2;
                const unsigned long long o_end = (K0 + total + vnl + 3) >> 2;
                const unsigned n_owned = static_cast<unsigned>(o_end - o_first);
                const unsigned j0 = static_cast<unsigned>((4 - (K0 & 3)) & 3);
                if (n_owned == 0) finish_pending();
                if (total + vnl <= static_cast<unsigned>(kWsNlCap)) {
#pragma unroll 1
                    for (unsigned h0 = 0; h0 < n_owned; h0 += kExt) {
                        // table updates of the previous tile (or pass): issued here, far from the next
                        // barrier, so that their atomics never stall one
                        finish_pending();
                        tick(1);
                        const unsigned h = h0 + pt;
                        const unsigned long long o = o_first + h;
                        bool have = (h < n_owned) && (o < a.read_limit);
                        unsigned long long key = 0, start_g = 0;
                        if (have) {
                            const unsigned j = j0 + 4 * h;
                            const unsigned sb = j ? nl[j - 1] + 1u : halo_start;
                            const int rc = parse_header(buf, s_lut, sb, nl[j], a, tile_off, &key, &start_g);
                            if (rc) {
                                raise_error(a.st, rc, o);
                                have = false;
                            }
                        }
                        const unsigned grp = __ballot_sync(0xFFFFFFFFu, have);
                        tick(6);
                        if (have) {
                            const unsigned same = __match_any_sync(grp, key);
                            if (a.table && lane == __ffs(same) - 1) {  // lowest lane = lowest read ordinal
                                p_key = key, p_pos = a.pos_base + o, p_cnt = __popc(same);
                                p_slot = hash64(key) & a.table_mask;
                                p_seen = *reinterpret_cast<volatile unsigned long long*>(&a.table[p_slot].key);
                            }
                            emit(o, key, start_g);
                        }
                        tick(7);
                    }
                } else if (pt == 0) {
                    // more newlines than the list holds (lines < 16 bytes on average): exact but serial
                    unsigned long long k = K0;
                    unsigned prev = halo_start;
                    for (unsigned p = 0; p <= valid; ++p) {
                        const bool is_end = (p < valid) ? (buf[kHalo + p] == '\n') : (vnl != 0);
                        if (!is_end) continue;
                        if ((k & 3) == 0 && (k >> 2) < a.read_limit) {
                            unsigned long long key = 0, start_g = 0;
                            const int rc = parse_header(buf, s_lut, prev, kHalo + p, a, tile_off, &key, &start_g);
                            if (rc) raise_error(a.st, rc, k >> 2);
                            else {
                                if (a.table) table_add(a.table, a.table_mask, key, 1, a.pos_base + (k >> 2), &a.st->occupied, a.st);
                                emit(k >> 2, key, start_g);
                            }
                        }
                        prev = kHalo + p + 1;
                        ++k;
                    }
                }
                group_sync<kExt>(3);  // every extractor is done reading stage s
                tick(2);
                if (pt == 0) {
                    unsigned long long tk;
                    do {
                        tk = s_ticket[i & 1];
                    } while (static_cast<unsigned>(tk >> 32) != i);
                    issue(s, static_cast<unsigned>(tk));
                    const unsigned long long c_hi = o_end < a.read_limit ? o_end : a.read_limit;
                    if (c_hi > o_first) my_reads += c_hi - o_first;
                    if (t == a.n_tiles - 1) a.st->line_carry = K0 + total + vnl;
                    tp[3] += 1;
                }
                tick(4);
            }
            if (a.timing && pt == 0) {
                atomicAdd(&a.timing[2], tp[0]), atomicAdd(&a.timing[6], tp[1]), atomicAdd(&a.timing[7], tp[6]);
                atomicAdd(&a.timing[8], tp[7]), atomicAdd(&a.timing[10], tp[2]), atomicAdd(&a.timing[11], tp[4]);
                atomicAdd(&a.timing[9], tp[3]);
            }
            if (pt == 0 && my_reads) atomicAdd(&a.st->n_reads, my_reads);
        }
#pragma unroll 1
        for (int k = 0; k < 2; ++k) finish_pending();  // the second call retires a CAS issued by the first
    }
}

}  // namespace frb
