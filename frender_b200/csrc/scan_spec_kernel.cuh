// Hot path A, the tally of whole files: fused read-name parse + 3-bit key packing + unique-combination count.
// Replaces scan_file's loop (reference frender.py:161-177): "every 4th line from the start of the file",
// key = 2nd space token's last ':' field (F:169).
//
// SPECULATIVE: the reference defines a header line by COUNT (every 4th line).  This kernel takes the line phase
// of a tile from the text itself (count_tile's guess), extracts and commits the keys at once under composite
// positions (tile << 13 | header index, which need no prefix over earlier tiles), and leaves every tile's newline
// count and guess in status[].  scan_verify.cuh then sums the counts, checks every guess against the count,
// gives tiles without a guess to scan_redo_kernel, and -- when a guess was wrong, i.e. the input is not
// well-formed FASTQ -- has the chunk taken out of the table again (this kernel with `negate`) and redone strictly
// by count.  The result is the reference's for every input; what the speculation buys is that no CTA ever waits
// for another CTA.
//
// Persistent CTAs (three per SM) claim 30 KiB tiles from a global ticket counter (SMs differ in speed by some
// ten percent -- a static tile-to-CTA map was measured 12 % slower).  Inside a CTA:
//
//   warps 0-3  COUNTERS    wait for the tile's bytes (TMA mbarrier); newline mask, scan, newline list, guess
//                          (count_tile); signal counted[stage]; leave count + guess in status[].
//   warps 4-6  EXTRACTORS  one thread per header line: LOAD what the line needs (registers + a private
//                          scratch), arrive on released[stage], then count spaces, extract + pack the key, fold
//                          equal keys of the warp, retire the table updates of the previous tile (atomics), and
//                          load the key's home slot for the next tile's update.  The three warps never wait for
//                          each other.
//   warp  7    DRIVER      one thread: when a stage is released, bulk-copy the next tile (ticket drawn one
//                          step ahead) into it.
//
// A stage is busy from the start of its copy until the extractors have LOADED their lines, not until they
// have parsed them: the copy of the next tile (2-3 thousand cycles) runs under the parsing.  Table updates are
// deferred by one tile per step (slot load -> RED or CAS -> RED), so no L2 / DRAM round trip is waited for in
// line.  The arrival on released[] is a relaxed one ordered by a register dependency on the stage loads: a
// releasing arrival would wait for the atomics and slot loads the thread has in flight.
#pragma once
#include "scan_count.cuh"

namespace frb {

// kProbe: clock instrumentation of the stage cycle (issue -> landed+counters free -> counted -> released -> issue)
template <class G, bool kProbe = false>
__global__ void __launch_bounds__(G::threads) __maxnreg__(G::maxreg) scan_spec_kernel(const ScanArgs a) {
    constexpr int kStages = G::stages;
    constexpr int kWsTile = G::tile, kWsBuf = G::buf, kWsNlCap = G::nl_cap;
    constexpr int kExt = G::xgroup, kXWarps = G::xwarps;
    extern __shared__ __align__(128) unsigned char smem[];
    uint16_t* const s_nl = reinterpret_cast<uint16_t*>(smem + kStages * kWsBuf);
    __shared__ __align__(8) unsigned long long s_full[kStages], s_counted[kStages], s_released[kStages];
    __shared__ unsigned s_tile[kStages];
    __shared__ TileMeta s_meta[kStages];
    __shared__ unsigned s_cwarp[kWsGroup / 32];
    __shared__ unsigned char s_lut[256];
    __shared__ uint4 s_tail[kExt][3];
    __shared__ long long s_clk[kStages][3];  // kProbe: issue, counted, released

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    // negate pass of a chunk whose guesses all held: nothing to take back
    if (a.negate && *reinterpret_cast<volatile int*>(&a.st->spec_bad) == 0) return;
    volatile unsigned long long* status = a.status + 1;

    for (int i = tid; i < 256; i += G::threads) s_lut[i] = lut_entry(i, FRB_RULE_SCAN);
    if (tid == 0) {
#pragma unroll
        for (int i = 0; i < kStages; ++i) {
            mbar_init(&s_full[i], 1);
            mbar_init(&s_counted[i], 1);
            mbar_init(&s_released[i], kXWarps);
        }
        mbar_fence_init();
    }
    __syncthreads();

    if (warp < kWsGroup / 32) {
        if constexpr (G::split_regs) asm volatile("setmaxnreg.dec.sync.aligned.u32 %0;" ::"n"(G::regs_count));
        const int ct = tid;
        // =============================== COUNTERS ================================================
        unsigned full_parity = 0;  // bit s
        for (unsigned i = 0;; ++i) {
            const int s = i % kStages;
            mbar_wait(&s_full[s], (full_parity >> s) & 1u);
            full_parity ^= 1u << s;
            const unsigned t = s_tile[s];
            if (t == kNoTile) {
                if (ct == 0) mbar_arrive(&s_counted[s]);
                break;
            }
            long long c0 = 0;
            if (kProbe && ct == 0) {
                c0 = clock64();
                atomicAdd(&a.timing[0], static_cast<unsigned long long>(c0 - s_clk[s][0]));  // issue -> counters start
                const unsigned long long d = static_cast<unsigned long long>(c0 - s_clk[s][0]);
                atomicAdd(&a.timing[16 + 96 + (d / 500 < 31 ? d / 500 : 31)], 1ULL);
            }
            const TileMeta m = count_tile<G, false>(a, smem + s * kWsBuf, s_nl + s * kWsNlCap, s_cwarp, t, true, nullptr, ct);
            if (ct == 0) {
                if (kProbe) {
                    const long long c1 = clock64();
                    atomicAdd(&a.timing[1], static_cast<unsigned long long>(c1 - c0));  // count
                    atomicAdd(&a.timing[9], 1ULL);
                    s_clk[s][1] = c1;
                }
                s_meta[s] = m;
                mbar_arrive(&s_counted[s]);
                // count and guess of the tile for scan_verify.cuh (nobody waits for it)
                if (!a.negate) status[t] = spec_info(m.total, m.vnl, m.guess);
            }
        }
    } else {
        if constexpr (G::split_regs) asm volatile("setmaxnreg.inc.sync.aligned.u32 %0;" ::"n"(G::regs_work));
        const int pt = tid - kWsGroup;
        const int pwarp = warp - kWsGroup / 32;
        if (pwarp == kXWarps) {
            // =============================== DRIVER ==============================================
            if (lane != 0) return;
            auto issue = [&](int s, unsigned long long ticket) -> bool {  // ticket -> stage s, start its bulk copy
                const unsigned t = ticket < a.n_tiles ? static_cast<unsigned>(ticket) : kNoTile;
                s_tile[s] = t;
                if (t == kNoTile) {
                    mbar_arrive(&s_full[s]);
                    return false;
                }
                const unsigned long long off = static_cast<unsigned long long>(t) * kWsTile;
                const unsigned halo = t ? kHalo : 0;
                const unsigned long long left = a.nbytes - off;
                const unsigned avail = static_cast<unsigned>(left < kWsTile ? left : kWsTile) + halo;
                const unsigned bulk = avail & ~15u;
                if (kProbe) s_clk[s][0] = clock64();
                if (bulk) {
                    mbar_expect_tx(&s_full[s], bulk);
                    bulk_g2s(smem + s * kWsBuf + (kHalo - halo), a.data + off - halo, bulk, &s_full[s]);
                } else {
                    mbar_arrive(&s_full[s]);
                }
                return true;
            };
            bool more = true;
            for (int s = 0; s < kStages && more; ++s) more = issue(s, atomicAdd(&a.status[0], 1ULL));
            unsigned long long next = more ? atomicAdd(&a.status[0], 1ULL) : 0ULL;
            unsigned rel_parity = 0;
            for (unsigned i = 0; more; ++i) {
                const int s = i % kStages;
                mbar_wait_one(&s_released[s], (rel_parity >> s) & 1u);
                rel_parity ^= 1u << s;
                if (kProbe) atomicAdd(&a.timing[3], static_cast<unsigned long long>(clock64() - s_clk[s][2]));  // released -> driver
                more = issue(s, next);
                // the ticket after this one: its round trip is over long before the next stage is released
                if (more) asm volatile("atom.global.add.u64 %0, [%1], 1;" : "=l"(next) : "l"(a.status) : "memory");
            }
            return;
        }
        auto give_back = [&](int s, unsigned i, unsigned dep) {
            if (lane == 0) mbar_arrive_relaxed_after(&s_released[s], dep);
        };
        // =============================== EXTRACTORS ==============================================
        // Deferred table update, three steps, each using a memory result requested one tile earlier:
        //   p: key -> load of its home slot's key
        //   home slot holds the key -> RED.ADD + RED.MIN;  free -> compare-and-swap (result dropped: the
        //   instruction returns it in its compare register, and copying it out would wait for the L2 round trip on
        //   the spot) followed by a load of the same slot (same thread, same address: it sees the slot after the
        //   swap);  taken by another key -> load of the next slot            (q)
        //   q: the loaded key is ours -> count / first update; anything else (lost race, two other keys in a row)
        //   -> the in-line probe loop.  Nobody knows who won a slot: occupied slots are counted when the file ends.
        unsigned long long p_key = 0, p_pos = 0, p_slot = 0, p_seen = 0, p_first = 0;
        unsigned long long q_key = 0, q_pos = 0, q_slot = 0, q_old = 0;
        unsigned p_cnt = 0, q_cnt = 0;
        // seen_first: a value `first` has had (it only ever falls): an update that cannot lower it is left out
        auto bump = [&](unsigned long long slot, unsigned cnt, unsigned long long pos, unsigned long long seen_first) {
            if (a.negate) {  // take the keys of a mis-guessed chunk out again
                atomicAdd(&a.table[slot].count, 0ULL - static_cast<unsigned long long>(cnt));
                return;
            }
            atomicAdd(&a.table[slot].count, static_cast<unsigned long long>(cnt));
            if (pos < seen_first) atomicMin(&a.table[slot].first, pos);
        };
        auto finish = [&]() {
            if (q_cnt) {
                if (q_old == q_key) bump(q_slot, q_cnt, q_pos, ~0ULL);
                else table_add(a.table, a.table_mask, q_key, a.negate ? 0ULL - q_cnt : static_cast<unsigned long long>(q_cnt),
                               a.negate ? ~0ULL : q_pos, &a.st->occupied, a.st);
                q_cnt = 0;
            }
            const bool claim = p_cnt && p_seen == kEmpty;
            const bool look = p_cnt && p_seen != p_key;  // claim or re-probe: a q set is born
            if (p_cnt) {
                if (p_seen == p_key) {
                    bump(p_slot, p_cnt, p_pos, p_first);
                } else {
                    q_key = p_key, q_pos = p_pos, q_cnt = p_cnt;
                    q_slot = claim ? p_slot : ((p_slot + 1) & a.table_mask);
                }
                p_cnt = 0;
            }
            asm volatile(
                "{\n\t.reg .pred p, q;\n\t.reg .b64 t;\n\t"
                "setp.ne.u32 p, %4, 0;\n\t"
                "setp.ne.u32 q, %5, 0;\n\t"
                "@p atom.global.cas.b64 t, [%1], %2, %3;\n\t"
                "@q ld.volatile.global.u64 %0, [%1];\n\t}"
                : "+l"(q_old)
                : "l"(&a.table[q_slot & a.table_mask].key), "l"(kEmpty), "l"(q_key), "r"(claim ? 1u : 0u),
                  "r"(look ? 1u : 0u)
                : "memory");
        };
        unsigned counted_parity = 0;
        unsigned long long ph[4] = {0, 0, 0, 0};  // kProbe: per-phase clocks of this warp
        long long pc = 0;
        unsigned long long ph_wait = 0;
        for (unsigned i = 0;; ++i) {
            const int s = i % kStages;
            const long long wait_start = kProbe ? clock64() : 0;
            mbar_wait(&s_counted[s], (counted_parity >> s) & 1u);
            counted_parity ^= 1u << s;
            const unsigned t = s_tile[s];
            if (t == kNoTile) break;
            if (kProbe && lane == 0) {  // wake-up lag: seen - max(counted, start of the wait), histogram per warp
                const long long now = clock64();
                const long long from = s_clk[s][1] > wait_start ? s_clk[s][1] : wait_start;
                const unsigned long long d = static_cast<unsigned long long>(now - from);
                if (pt == 0) atomicAdd(&a.timing[4], static_cast<unsigned long long>(now - s_clk[s][1]));
                atomicAdd(&a.timing[16 + 32 * pwarp + (d / 500 < 31 ? d / 500 : 31)], 1ULL);
                ph_wait += static_cast<unsigned long long>(now - wait_start);
            }
            unsigned char* const buf = smem + s * kWsBuf;
            const TileMeta m = s_meta[s];
            const unsigned lines = m.total + m.vnl, guess = m.guess;
            const unsigned long long tile_off = static_cast<unsigned long long>(t) * kWsTile;
            const unsigned long long pos0 = (a.tile_base + t) << kCompositeShift;
            const uint16_t* const nl = s_nl + s * kWsNlCap;
            // no guess (or more newlines than the list holds): scan_redo_kernel takes the tile
            const unsigned n_owned = (guess != kNoGuess && lines > guess) ? (lines - guess + 3) / 4 : 0;
            bool released = false;
#pragma unroll 1
            for (unsigned h0 = 0; h0 < n_owned; h0 += kExt) {
                const unsigned h = h0 + pt;
                const bool have = h < n_owned;
                unsigned sb = 0, eb = 0;
                if (have) {
                    const unsigned j = guess + 4 * h;
                    sb = j ? nl[j - 1] + 1u : m.halo;
                    eb = nl[j];
                }
                const unsigned segs = (have && sb != kUnknown) ? (eb - (sb & ~15u) + 15u) >> 4 : 0u;
                const bool need5 = __any_sync(0xFFFFFFFFu, segs >= 6u && segs <= 7u);
                const bool need6 = __any_sync(0xFFFFFFFFu, segs == 7u);
                auto ptick = [&](int k) {
                    if (kProbe && lane == 0) {
                        const long long now = clock64();
                        ph[k] += static_cast<unsigned long long>(now - pc);
                        pc = now;
                    }
                };
                if (kProbe && lane == 0) pc = clock64();
                HeaderRegs hr;
                unsigned dep = header_load(buf, sb, eb, have, need5, need6, hr, s_tail[pt]);
                if (h0 + kExt >= n_owned) {  // last pass over this tile: this warp is done reading the stage
                    dep = __reduce_xor_sync(0xFFFFFFFFu, dep);  // every lane's loads
                    if (kProbe && pt == 0) {
                        const long long r = clock64();
                        atomicAdd(&a.timing[2], static_cast<unsigned long long>(r - s_clk[s][1]));  // counted -> released
                        s_clk[s][2] = r;
                    }
                    give_back(s, i, dep);
                    released = true;
                }
                ptick(0);
                unsigned long long key = kEmpty;
                if (have) {
                    const int rc = header_key(hr, s_tail[pt], s_lut, a, tile_off, need5, need6, &key);
                    if (rc) {  // an error of a guessed tile counts once the guesses are confirmed
                        key = kEmpty;
                        if (atomicCAS(&a.st->spec_err_code, 0, rc) == 0) a.st->spec_err_pos = pos0 + h;
                    }
                }
                ptick(1);
                // table updates of the previous pass / tile: their slot loads were issued a whole tile ago
                finish();
                ptick(2);
                // fold equal keys of the warp: the lowest lane (lowest read ordinal) carries the count
                const unsigned grp = __ballot_sync(0xFFFFFFFFu, key != kEmpty);
                if (key != kEmpty) {
                    const unsigned same = __match_any_sync(grp, key);
                    if (lane == __ffs(same) - 1) {
                        p_key = key, p_pos = pos0 + h, p_cnt = __popc(same);
                        p_slot = hash64(key) & a.table_mask;
                        // key and first of the home slot in one 16-byte load
                        asm volatile("ld.volatile.global.v2.u64 {%0, %1}, [%2];"
                                     : "=l"(p_seen), "=l"(p_first)
                                     : "l"(&a.table[p_slot].key)
                                     : "memory");
                    }
                }
                ptick(3);
            }
            if (!released) give_back(s, i, 0u);
        }
        finish();
        finish();  // retires a compare-and-swap issued by the call above
        if (kProbe && lane == 0)
        {
            for (int k = 0; k < 4; ++k) atomicAdd(&a.timing[160 + 8 * pwarp + k], ph[k]);
            atomicAdd(&a.timing[160 + 8 * pwarp + 4], ph_wait);
        }
    }
}

}  // namespace frb
