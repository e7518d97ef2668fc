// Hot path A: fused read-name parse + 3-bit key packing + unique-combination count.
// Replaces scan_file's loop (reference frender.py:161-177): "every 4th line from the start of
// the file", key = 2nd space token's last ':' field (F:169) or, for demux, the last ':' field of
// the whole line (F:778).
//
// One persistent CTA per (SM x 3) walks 32 KiB tiles of the decompressed FASTQ:
//   * the tile (+ the 512 B before it) is brought into shared memory by one TMA bulk copy
//     (cp.async.bulk + mbarrier), double buffered so the next tile streams in while this one
//     is processed -- every input byte crosses HBM once;
//   * each thread turns its 128 contiguous bytes into a 128-bit newline mask with 16-byte
//     shared loads and SIMD-in-word byte compares; a block scan plus a decoupled look-back
//     over per-tile newline counts (single pass, no second read of the data) gives every
//     newline its line number, hence "line % 4 == 0" and the read ordinal;
//   * a header line is owned by the tile holding its terminating newline; one thread per
//     owned header extracts and packs the key (branch-light backward scan over <= 22 bytes +
//     a word-parallel space count) and the warp folds equal keys (__match_any_sync) before a
//     single atomic update of the 32-byte table slot (count += n, first = min).
#pragma once
#include "common.cuh"

namespace frb {

constexpr int kHalo = 512;
constexpr int kPerThread = 128;  // bytes of a tile owned by one thread
constexpr unsigned kUnknown = 0xFFFFu;
constexpr int kRuleOffsetsOnly = 2;  // internal: no key, record offsets only
constexpr int kStages = 3;           // shared-memory tile buffers per CTA

// Geometry of one scan CTA: NT threads walk tiles of NT * 128 bytes.
template <int NT>
struct ScanCfg {
    static constexpr int threads = NT;
    static constexpr int tile = NT * kPerThread;
    static constexpr int buf = tile + kHalo;
    static constexpr int nl_cap = NT * 8;  // newline positions kept per tile (more: serial fallback)
    static constexpr int smem = kStages * buf + 2 * nl_cap * (int)sizeof(uint16_t);
    static constexpr int ctas_per_sm = (220 * 1024) / (smem + 2048) > 8 ? 8 : (220 * 1024) / (smem + 2048);
};

#define kFlagAgg (1ULL << 62)
#define kFlagInc (2ULL << 62)
#define kValMask ((1ULL << 62) - 1)

struct ScanArgs {
    const unsigned char* data;      // chunk, 16-byte aligned, begins at a line start
    unsigned long long nbytes;
    unsigned long long line_base;   // lines of the file before this chunk (ignored if use_carry)
    unsigned long long read_limit;  // reads with ordinal >= limit are not tallied (-s, F:163-165)
    unsigned long long pos_base;    // added to the read ordinal to form `first`
    Slot* table;
    unsigned long long table_mask;
    unsigned long long* status;     // [0] tile counter, [1 + t] look-back word of tile t
    DevState* st;
    unsigned long long* keys_out;     // optional: key of read r at [r - first read of chunk]
    unsigned long long* rec_off_out;  // optional: chunk offset of the record start, same index
    unsigned long long out_cap;       // entries available in keys_out / rec_off_out
    unsigned int n_tiles;
    int use_carry;
    int rule;
    unsigned long long* timing;  // optional: per-stage clock64 sums of thread 0 of every CTA
    int dbg_flags;  // experiments: 1 = no table update, 2 = no atomicMin, 4 = no parse
    int no_tma;     // debug / A-B: stage tiles with plain 16-byte loads instead of the bulk copy
    unsigned int* redo;       // warp-specialised kernel: tiles left to scan_redo_kernel, capacity n_tiles
    unsigned int tile_bytes;  // tile size of the kernel that filled status[] (for scan_redo_kernel)
    int no_guess;             // A-B: never guess the line phase from the text
    // '\n' and ' ' replicated over a word.  Kernel parameters, not literals: with the pattern in a register the
    // byte compare is three instructions per word (LOP3 takes one immediate; as a literal next to the 0x7f..
    // mask the pattern costs a fourth).
    unsigned int pat_nl, pat_sp;
};

// ---- PTX helpers: mbarrier + TMA bulk copy ------------------------------------------------
__device__ __forceinline__ unsigned smem_addr(const void* p) {
    return static_cast<unsigned>(__cvta_generic_to_shared(p));
}
__device__ __forceinline__ void mbar_init(unsigned long long* bar, unsigned count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_addr(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_fence_init() {
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void mbar_expect_tx(unsigned long long* bar, unsigned bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_addr(bar)), "r"(bytes)
                 : "memory");
}
__device__ __forceinline__ void mbar_arrive(unsigned long long* bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_addr(bar)) : "memory");
}
__device__ __forceinline__ void mbar_wait(unsigned long long* bar, unsigned parity) {
    // whole spin loop in one asm block (the form ptxas knows), then an explicit warp reconvergence:
    // lanes leave the loop at different times and the warp-collective code that follows
    // (shuffles, ballots) must not run on a partial warp.
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "FRB_WAIT_%=:\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t"
        "@p bra FRB_DONE_%=;\n\t"
        "bra FRB_WAIT_%=;\n\t"
        "FRB_DONE_%=:\n\t}"
        ::"r"(smem_addr(bar)), "r"(parity)
        : "memory");
    // not __syncwarp(): nvcc sees straight-line code here and drops it
    asm volatile("bar.warp.sync 0xffffffff;" ::: "memory");
}
// non-blocking probe: has the phase with this parity completed?
__device__ __forceinline__ bool mbar_test(unsigned long long* bar, unsigned parity) {
    unsigned done;
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.test_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(done)
        : "r"(smem_addr(bar)), "r"(parity)
        : "memory");
    return done != 0;
}
// single-thread wait (no warp reconvergence: the caller is one elected lane)
__device__ __forceinline__ void mbar_wait_one(unsigned long long* bar, unsigned parity) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "FRB_WAIT1_%=:\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t"
        "@p bra FRB_DONE1_%=;\n\t"
        "bra FRB_WAIT1_%=;\n\t"
        "FRB_DONE1_%=:\n\t}"
        ::"r"(smem_addr(bar)), "r"(parity)
        : "memory");
}
__device__ __forceinline__ void bulk_g2s(void* dst, const void* src, unsigned bytes, unsigned long long* bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                     smem_addr(dst)),
                 "l"(src), "r"(bytes), "r"(smem_addr(bar))
                 : "memory");
}

#define kFoldLsb3 0x1249249249249249ULL  // bit 0 of every 3-bit group

// ---- byte-compare primitives ---------------------------------------------------------------
// bit 7 of every byte of w that equals the byte replicated in `pat` (pattern bytes < 0x80); exact (no borrow
// leaks).  (w & 0x7f..) ^ pat is the low seven bits of w ^ pat, and bit 7 of w ^ pat is bit 7 of w.
__device__ __forceinline__ unsigned eq_flags(unsigned w, unsigned pat) {
    const unsigned t = ((w & 0x7F7F7F7Fu) ^ pat) + 0x7F7F7F7Fu;
    return ~(t | w) & 0x80808080u;
}
// 16-bit mask (bit i = byte i) of the bytes of v equal to the byte replicated in `pat`.
// The flags (0x80 per equal byte) are gathered with byte dot products: weights 1,2,4,8 for the first word
// of a pair and 16,32,64,128 for the second give mask8 * 128, on the multiply pipe instead of the ALU.
__device__ __forceinline__ unsigned eq_mask16(const uint4 v, unsigned pat) {
    unsigned lo = __dp4a(eq_flags(v.x, pat), 0x08040201u, 0u);
    lo = __dp4a(eq_flags(v.y, pat), 0x80402010u, lo);
    unsigned hi = __dp4a(eq_flags(v.z, pat), 0x08040201u, 0u);
    hi = __dp4a(eq_flags(v.w, pat), 0x80402010u, hi);
    return (lo >> 7) | (hi << 1);
}
__device__ __forceinline__ unsigned newline_mask16(const uint4 v, unsigned pat_nl) { return eq_mask16(v, pat_nl); }

// Number of ' ' in buf[sb, eb) for lines that end within 112 bytes of their 16-byte-aligned start,
// else -1.  Fully unrolled and branch-free: seven independent 16-byte shared loads in flight.
__device__ __forceinline__ int count_spaces_fast(const unsigned char* buf, unsigned sb, unsigned eb, unsigned pat_sp) {
    const unsigned a0 = sb & ~15u;
    if (eb - a0 > 112u) return -1;
    unsigned cnt = 0;
#pragma unroll
    for (int i = 0; i < 7; ++i) {
        const unsigned p = a0 + 16u * i;
        if (i >= 4 && p >= eb) break;  // ordinary header lines span 5 segments; the loads stay batched
        const uint4 v = *reinterpret_cast<const uint4*>(buf + (p < eb ? p : a0));
        unsigned m = eq_mask16(v, pat_sp);
        const unsigned lo_cut = sb > p ? sb - p : 0u;                  // bytes of this segment before the line
        const unsigned hi_cut = eb > p ? (eb - p < 16u ? eb - p : 16u) : 0u;  // bytes of it inside [.., eb)
        m &= (0xFFFFu << lo_cut) & ((1u << hi_cut) - 1u);
        cnt += __popc(m);
    }
    return static_cast<int>(cnt);
}

// shared-memory LUT entry of byte c for parse_header: bits 0-2 symbol code, bit 3 = c ends the key
// when walking back from the end of the line (':' always, ' ' under the scan rule).  Nothing above bit 3.
__device__ __forceinline__ unsigned char lut_entry(unsigned c, int rule) {
    const bool delim = (c == ':') || (rule == FRB_RULE_SCAN && c == ' ');
    return static_cast<unsigned char>(enc_read(c) | (delim ? 0x08u : 0u));
}

// Exact, byte-serial statement of both key rules over one header line (no trailing newline).
// Used for the rare lines the fast path declines (second space, key > 21, line not in smem).
__host__ __device__ inline int parse_serial(const unsigned char* s, unsigned long long len, int rule,
                                            unsigned long long* key_out) {
    unsigned long long i = 0;
    if (rule == FRB_RULE_SCAN) {
        while (i < len && s[i] != ' ') ++i;
        if (i >= len) return FRB_ERR_BAD_HEADER;  // split(" ")[1] -> IndexError, F:169
        ++i;
    }
    unsigned long long k = 0;
    int n = 0;
    bool bad = false, too_long = false;
    for (; i < len; ++i) {
        unsigned c = s[i];
        if (rule == FRB_RULE_SCAN && c == ' ') break;
        if (c == ':') {
            k = 0, n = 0, bad = false, too_long = false;
            continue;
        }
        unsigned code = enc_read(c);
        bad |= (code == 0);
        if (n >= kMaxSyms) too_long = true;
        else k |= static_cast<unsigned long long>(code) << (3 * n);
        ++n;
    }
    if (bad) return FRB_ERR_BAD_ALPHABET;
    if (too_long) return FRB_ERR_KEY_TOO_LONG;
    *key_out = k;
    return 0;
}

// Key of the header line occupying buffer positions [sb, eb) (sb == kUnknown: starts before
// the staged bytes).  Buffer position p is chunk offset tile_off + p - kHalo.
__device__ __forceinline__ int parse_header(const unsigned char* buf, const unsigned char* lut, unsigned sb,
                                            unsigned eb, const ScanArgs& a, unsigned long long tile_off,
                                            unsigned long long* key_out, unsigned long long* start_out) {
    const bool scan_rule = (a.rule == FRB_RULE_SCAN);
    if (a.rule == kRuleOffsetsOnly) {  // record boundaries only (R1 side of the demux router)
        if (sb != kUnknown) {
            *start_out = tile_off + sb - kHalo;
        } else {
            unsigned long long s_g = tile_off + eb - kHalo;
            while (s_g > 0 && a.data[s_g - 1] != '\n') --s_g;
            *start_out = s_g;
        }
        *key_out = 0;
        return 0;
    }
    if (sb != kUnknown) *start_out = tile_off + sb - kHalo;
    if (sb != kUnknown && eb - sb > static_cast<unsigned>(kMaxSyms)) {
        // Fast path, branch-free and latency-flat.  With exactly one ' ' in the line the key is the
        // text after the last ':' or ' ' (the 2nd space token runs to the end of the line).  The line
        // has at least 22 bytes, so the 22 bytes before its end all belong to it.
        const int spaces = scan_rule ? count_spaces_fast(buf, sb, eb, a.pat_sp) : 1;
        const unsigned char* const e = buf + eb;
        unsigned delim = 0, lo = 0, hi = 0, top = 0;
#pragma unroll
        for (int j = 0; j < kMaxSyms + 1; ++j) {  // closest to the end of the line first
            const unsigned v = lut[e[-1 - j]];
            const unsigned code = v & 7u;
            delim += (v >> 3) << j;
            if (j < 10) lo += code << (3 * j);
            else if (j < 20) hi += code << (3 * (j - 10));
            else if (j == 20) top = code;
        }
        if (spaces == 0) return FRB_ERR_BAD_HEADER;
        if (spaces == 1 && delim != 0) {
            const int len = __ffs(delim) - 1;  // symbols in the key, <= 21
            // krev holds the key backwards (symbol j = j-th char from the end)
            const unsigned long long krev = static_cast<unsigned long long>(lo) |
                                            (static_cast<unsigned long long>(hi) << 30) |
                                            (static_cast<unsigned long long>(top) << 60);
            // every one of the len symbols must have a non-zero code
            const unsigned long long want = len ? (kFoldLsb3 & ((1ULL << (3 * len)) - 1ULL)) : 0ULL;
            if (((krev | (krev >> 1) | (krev >> 2)) & want) != want) return FRB_ERR_BAD_ALPHABET;
            // reverse the order of the 3-bit groups: bit-reverse the word, then put the bits of every
            // group back in order
            const unsigned long long r = __brevll(krev) >> 1;  // group j now at group index 20 - j, bits mirrored
            const unsigned long long g = ((r & kFoldLsb3) << 2) | (r & (kFoldLsb3 << 1)) | ((r >> 2) & kFoldLsb3);
            *key_out = len ? (g >> (3 * (kMaxSyms - len))) : 0ULL;
            return 0;
        }
    }
    const unsigned long long e_g = tile_off + eb - kHalo;
    unsigned long long s_g;
    if (sb != kUnknown) {
        s_g = tile_off + sb - kHalo;
    } else {
        s_g = e_g;
        while (s_g > 0 && a.data[s_g - 1] != '\n') --s_g;
    }
    *start_out = s_g;
    return parse_serial(a.data + s_g, e_g - s_g, a.rule, key_out);
}

// Decoupled look-back over per-tile newline counts.  The tile's own count was published by the
// count stage (one pipeline step earlier, see scan_kernel); returns the number of newlines before
// tile t and publishes the tile's inclusive prefix.  Called by one full warp.
// kBlocking = false: a single pass that gives up (returns false) as soon as a needed count is not
// published yet -- used to take the look-back off the critical path without ever stalling a CTA on
// another CTA's progress.
template <bool kBlocking>
__device__ __forceinline__ bool tile_prefix(volatile unsigned long long* status, unsigned t, unsigned total,
                                            int lane, unsigned long long* out) {
    // Windows of 128 predecessors per step (four independent loads per lane, one L2 round trip): with several
    // hundred tiles in flight the nearest inclusive word is usually more than 32 tiles back.  Measured effect
    // on the kernel: +0.7 %; what the look-back mostly waits for is predecessors that are not counted yet.
    constexpr int kSub = 4;
    *out = 0;
    if (t == 0) return true;  // published as inclusive by the count stage
    unsigned long long acc = 0;  // lane-local partial sum, reduced once at the end
    long long idx = static_cast<long long>(t) - 1;
    for (;;) {
        bool done = false;
        for (;;) {  // until every word this step needs has been published
            unsigned long long s[kSub];
#pragma unroll
            for (int m = 0; m < kSub; ++m) {
                const long long j = idx - 32 * m - lane;
                s[m] = (j >= 0) ? status[j] : kFlagInc;
            }
            unsigned long long part = 0;
            bool ready = true;
#pragma unroll
            for (int m = 0; m < kSub; ++m) {
                if (ready && !done) {
                    const unsigned none = __ballot_sync(0xFFFFFFFFu, (s[m] >> 62) == 0);
                    const unsigned inc = __ballot_sync(0xFFFFFFFFu, (s[m] >> 62) == 2);
                    const int first_inc = inc ? (__ffs(inc) - 1) : 32;
                    const unsigned relevant = (first_inc >= 31) ? 0xFFFFFFFFu : ((2u << first_inc) - 1u);
                    if (none & relevant) {
                        ready = false;
                    } else {
                        part += (lane <= first_inc) ? (s[m] & kValMask) : 0ULL;
                        done = first_inc < 32;
                    }
                }
            }
            if (ready) {
                acc += part;
                break;
            }
            done = false;
            if (!kBlocking) return false;
        }
        if (done) break;
        idx -= 32 * kSub;
    }
#pragma unroll
    for (int d = 16; d > 0; d >>= 1) acc += __shfl_xor_sync(0xFFFFFFFFu, acc, d);
    if (lane == 0) status[t] = kFlagInc | (acc + total);
    *out = acc;
    return true;
}

// Software pipeline of one CTA over its tiles T0, T1, ... (claimed from a global counter), three
// shared-memory buffers deep:
//   iteration i:  claim Ti+2 and start its bulk copy          (lands during this iteration)
//                 C  Ti+1 (copy started one iteration ago): newline masks, block scan, publish count
//                 A  Ti: look-back over the published counts -> line number of its first newline
//                 B  Ti: header positions, key extraction, table update
// Stage C depends on nothing but the tile's bytes, so every tile's count is published about one
// tile-time before any CTA looks back over it: the look-back never waits on a chain of CTAs.
struct TileState {
    unsigned total;             // newlines in the tile
    unsigned valid;             // bytes in the tile
    unsigned vnl;               // 1 if the chunk ends in this tile without a final '\n'
};

template <int NT>
__global__ void __launch_bounds__(NT, ScanCfg<NT>::ctas_per_sm) scan_kernel(const ScanArgs a) {
    constexpr int kThreads = NT, kTile = ScanCfg<NT>::tile, kBuf = ScanCfg<NT>::buf, kNlCap = ScanCfg<NT>::nl_cap;
    extern __shared__ __align__(128) unsigned char smem[];
    // positions (buffer offsets) of the newlines of a tile, in order; two lists: the tile being parsed
    // and the tile counted one stage ahead
    uint16_t* const s_nl = reinterpret_cast<uint16_t*>(smem + kStages * kBuf);
    __shared__ __align__(8) unsigned long long s_bar[kStages];
    __shared__ unsigned long long s_prefix[2];
    __shared__ unsigned s_have_prefix[2];  // early look-back of the tile parsed in iteration (i & 1) succeeded
    __shared__ unsigned s_tile[kStages];
    __shared__ unsigned s_warp[kThreads / 32];
    __shared__ unsigned s_halo_start;
    __shared__ unsigned char s_lut[256];

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const unsigned long long L0 =
        a.use_carry ? *reinterpret_cast<volatile unsigned long long*>(&a.st->line_carry) : a.line_base;
    const unsigned long long chunk_first_read = (L0 + 3) >> 2;
    volatile unsigned long long* status = a.status + 1;

    // stage timing (thread 0 only, when a.timing is set): 0 wait, 1 count, 2 look-back, 3 positions,
    // 4 parse, 5 insert+sync, 6 end-of-tile atomics+sync, 7 claim+issue+sync, 9 tiles
    long long tmark = 0;
    unsigned long long tacc[10] = {0, 0, 0, 0, 0, 0, 0, 0, 0, 0};
    auto tick = [&](int slot) {
        if (a.timing && tid == 0) {
            const long long now = clock64();
            tacc[slot] += static_cast<unsigned long long>(now - tmark);
            tmark = now;
        }
    };
    if (a.timing && tid == 0) tmark = clock64();

    // Tile tickets.  __syncthreads() waits for the arriving thread's outstanding global atomics, so the
    // ticket for the tile after next is drawn at the start of the barrier-free key-extraction phase
    // (claim_next) and only consumed one iteration later (issue): its L2 round trip is never exposed.
    unsigned claimed = 0;
    auto claim_next = [&]() { claimed = static_cast<unsigned>(atomicAdd(&a.status[0], 1ULL)); };  // thread 0
    auto issue = [&](int b) {  // thread 0 only: start the bulk copy of tile `claimed` into buffer b
        const unsigned t = claimed;
        s_tile[b] = t;
        if (t >= a.n_tiles) return;
        const unsigned long long off = static_cast<unsigned long long>(t) * kTile;
        const unsigned halo = t ? kHalo : 0;
        const unsigned long long left = a.nbytes - off;
        const unsigned avail = static_cast<unsigned>(left < kTile ? left : kTile) + halo;
        const unsigned bulk = avail & ~15u;
        if (bulk && !a.no_tma) {
            mbar_expect_tx(&s_bar[b], bulk);
            bulk_g2s(smem + b * kBuf + (kHalo - halo), a.data + off - halo, bulk, &s_bar[b]);
        } else {
            mbar_arrive(&s_bar[b]);
        }
    };

    unsigned parity = 0;  // bit b = phase parity of s_bar[b]

    // stage C: newline masks + block scan of the tile staged in buffer b, publish its count
    auto count_stage = [&](int b, TileState& ts, unsigned lsel) {
        const unsigned t = s_tile[b];
        if (t >= a.n_tiles) return;  // uniform
        unsigned char* const buf = smem + b * kBuf;
        mbar_wait(&s_bar[b], (parity >> b) & 1u);
        parity ^= 1u << b;
        tick(0);
        const unsigned long long tile_off = static_cast<unsigned long long>(t) * kTile;
        const unsigned long long left = a.nbytes - tile_off;
        ts.valid = static_cast<unsigned>(left < kTile ? left : kTile);
        {   // bytes past the last 16-byte multiple of the bulk copy (final tile only)
            const unsigned halo = t ? kHalo : 0;
            const unsigned avail = ts.valid + halo, bulk = avail & ~15u;
            if (a.no_tma) {
                const uint4* src = reinterpret_cast<const uint4*>(a.data + tile_off - halo);
                uint4* dst = reinterpret_cast<uint4*>(buf + (kHalo - halo));
                for (unsigned i = tid; i < bulk / 16; i += kThreads) dst[i] = src[i];
                if (avail == bulk) __syncthreads();
            }
            if (avail != bulk) {
                if (tid < static_cast<int>(avail - bulk))
                    buf[(kHalo - halo) + bulk + tid] = a.data[tile_off - halo + bulk + tid];
                __syncthreads();
            }
        }
        unsigned long long lo = 0, hi = 0;
        const uint4* t4 = reinterpret_cast<const uint4*>(buf + kHalo) + tid * (kPerThread / 16);
#pragma unroll
        for (int j = 0; j < kPerThread / 16; ++j) {
            const int seg = (j + tid) & 7;  // rotate so a quarter-warp hits 8 distinct bank groups
            const unsigned long long m = newline_mask16(t4[seg], a.pat_nl);
            const int sh = (seg & 3) * 16;
            if (seg < 4) lo |= m << sh;
            else hi |= m << sh;
        }
        const int nv = static_cast<int>(ts.valid) - tid * kPerThread;
        if (nv < kPerThread) {
            if (nv <= 0) lo = 0, hi = 0;
            else if (nv < 64) lo &= (1ULL << nv) - 1, hi = 0;
            else hi &= (1ULL << (nv - 64)) - 1;
        }
        const unsigned cnt = __popcll(lo) + __popcll(hi);
        unsigned incl = cnt;
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) {
            const unsigned n = __shfl_up_sync(0xFFFFFFFFu, incl, d);
            if (lane >= d) incl += n;
        }
        if (lane == 31) s_warp[warp] = incl;
        __syncthreads();
        unsigned wbase = 0, total = 0;
#pragma unroll
        for (int w = 0; w < kThreads / 32; ++w) {
            const unsigned v = s_warp[w];
            if (w < warp) wbase += v;
            total += v;
        }
        ts.total = total;
        if (tid == 0) status[t] = (t == 0 ? kFlagInc : kFlagAgg) | total;
        // a last line without '\n' still is a line (F:161 iterates it; F:169 rstrip)
        ts.vnl = (t == a.n_tiles - 1 && ts.valid > 0 && buf[kHalo + ts.valid - 1] != '\n') ? 1u : 0u;
        {   // ordered list of newline positions; the line numbers come later, from the look-back
            uint16_t* const nl = s_nl + lsel * kNlCap;
            unsigned idx = wbase + incl - cnt;
            unsigned pos0 = kHalo + tid * kPerThread;
            unsigned long long m = lo;
#pragma unroll
            for (int half = 0; half < 2; ++half) {
                while (m) {
                    if (idx < kNlCap) nl[idx] = static_cast<uint16_t>(pos0 + (__ffsll(static_cast<long long>(m)) - 1));
                    m &= m - 1;
                    ++idx;
                }
                m = hi;
                pos0 += 64;
            }
            if (tid == 0 && ts.vnl && total < kNlCap) nl[total] = static_cast<uint16_t>(kHalo + ts.valid);
        }
        __syncthreads();  // s_warp is reused by the next count
        tick(1);
    };

    s_lut[tid & 255] = lut_entry(tid & 255, a.rule);
    if (kThreads < 256) s_lut[(tid + kThreads) & 255] = lut_entry((tid + kThreads) & 255, a.rule);
    if (tid == 0) {
#pragma unroll
        for (int i = 0; i < kStages; ++i) mbar_init(&s_bar[i], 1);
        mbar_fence_init();
        claim_next();
        issue(0);
        claim_next();
        issue(1);
        claim_next();
    }
    __syncthreads();
    TileState cur{}, nxt{};
    count_stage(0, cur, 0);
    // look-back of the first tile; later tiles are looked back one iteration ahead (see stage B)
    if (warp == kThreads / 32 - 1 && lane == 0) s_have_prefix[0] = 0, s_have_prefix[1] = 0;
    __syncthreads();

    // Deferred table update, three steps, each consuming a memory result issued one tile earlier so
    // that no DRAM / L2 round trip of the (random) probe sits on the tile's critical path:
    //   step 1 (tile i)    key extracted -> load of its home slot issued
    //   step 2 (tile i+1)  slot holds the key -> RED.ADD/RED.MIN;  slot empty -> CAS issued
    //   step 3 (tile i+2)  CAS won (or another thread inserted the same key) -> RED.ADD/RED.MIN
    // anything else (collision) falls back to the synchronous probing loop.
    unsigned long long p_key = 0, p_pos = 0, p_slot = 0, p_seen = 0;  // step-1 state
    unsigned p_cnt = 0;                                               // 0 = nothing pending
    unsigned long long q_key = 0, q_pos = 0, q_slot = 0, q_old = 0;   // step-2 state (CAS in flight)
    unsigned q_cnt = 0;
    auto bump = [&](unsigned long long slot, unsigned cnt, unsigned long long pos) {
        atomicAdd(&a.table[slot].count, static_cast<unsigned long long>(cnt));
        if (!(a.dbg_flags & 2)) atomicMin(&a.table[slot].first, pos);
    };
    auto finish_pending = [&]() {
        if (q_cnt) {  // step 3
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
        if (p_cnt) {  // step 2
            if (a.dbg_flags & 1) {
            } else if (p_seen == p_key) {
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

    int c = 0, n = 1, nn = 2;
    unsigned long long my_reads = 0;  // thread 0: reads tallied by this CTA
    for (unsigned iter = 0;; ++iter) {
        const unsigned t = s_tile[c];
        if (t >= a.n_tiles) break;
        if (tid == 0) issue(nn);  // buffer nn was released at the end of the last iteration
        tick(8);
        __syncthreads();          // s_tile[nn] visible; also orders the stages
        tick(7);
        count_stage(n, nxt, (iter + 1) & 1);

        unsigned char* const buf = smem + c * kBuf;
        const unsigned long long tile_off = static_cast<unsigned long long>(t) * kTile;
        const bool is_last = (t == a.n_tiles - 1);
        const unsigned total = cur.total, vnl = cur.vnl, valid = cur.valid;

        // ---- stage A: tile prefix unless the early attempt of the previous iteration already has it
        //      (warp 0), start of the line that straddles the tile start (last warp) -----------------
        if (warp == 0 && !s_have_prefix[iter & 1]) {
            unsigned long long excl;
            tile_prefix<true>(status, t, total, lane, &excl);
            if (lane == 0) s_prefix[iter & 1] = excl;
        }
        if (warp == kThreads / 32 - 1) {
            if (t == 0) {
                if (lane == 0) s_halo_start = kHalo;
            } else {
                const unsigned m = newline_mask16(reinterpret_cast<const uint4*>(buf)[lane], a.pat_nl);
                const unsigned any = __ballot_sync(0xFFFFFFFFu, m != 0);
                if (any == 0) {
                    if (lane == 0) s_halo_start = kUnknown;
                } else if (lane == 31 - __clz(any)) {
                    s_halo_start = lane * 16 + (31 - __clz(m)) + 1;
                }
            }
        }
        __syncthreads();
        tick(2);

        const unsigned long long K0 = L0 + s_prefix[iter & 1];  // index of the first newline of the tile
        const unsigned long long o_first = (K0 + 3) >> 2;
        const unsigned long long o_end = (K0 + total + vnl + 3) >> 2;
        const unsigned n_owned = static_cast<unsigned>(o_end - o_first);
        const unsigned j0 = static_cast<unsigned>((4 - (K0 & 3)) & 3);  // list index of the first header end
        const uint16_t* const nl = s_nl + (iter & 1) * kNlCap;

        // Work that issues global atomics or polls other CTAs runs once per tile at the start of the
        // longest barrier-free stretch (key extraction), so that no __syncthreads() waits for an L2
        // round trip: last tile's table updates, the ticket for the tile after next, and a non-blocking
        // look-back attempt for the NEXT tile (its count was published by this iteration's stage C) on
        // the last warp, which has no headers to parse in ordinary tiles.
        finish_pending();
        if (tid == 0) claim_next();
        if (warp == kThreads / 32 - 1) {
            unsigned long long excl = 0;
            const bool ok = s_tile[n] < a.n_tiles && tile_prefix<false>(status, s_tile[n], nxt.total, lane, &excl);
            if (lane == 0) s_prefix[(iter + 1) & 1] = excl, s_have_prefix[(iter + 1) & 1] = ok ? 1u : 0u;
        }

        // ---- stage B: one thread per header line owned by the tile ---------------------------------
        // header h ends at list entry j0 + 4h and starts right after entry j0 + 4h - 1 (or in the halo)
        auto emit = [&](unsigned long long o, unsigned long long key, unsigned long long start_g) {
            const unsigned long long slot = o - chunk_first_read;
            if (slot < a.out_cap) {
                if (a.keys_out) a.keys_out[slot] = key;
                if (a.rec_off_out) a.rec_off_out[slot] = start_g;
            }
        };
        if (total + vnl <= static_cast<unsigned>(kNlCap)) {
            for (unsigned h0 = 0; h0 < n_owned; h0 += kThreads) {
                const unsigned h = h0 + tid;
                const unsigned long long o = o_first + h;
                bool have = (h < n_owned) && (o < a.read_limit) && !(a.dbg_flags & 4);
                unsigned long long key = 0, start_g = 0;
                if (have) {
                    const unsigned j = j0 + 4 * h;
                    const unsigned sb = j ? nl[j - 1] + 1u : s_halo_start;
                    const int rc = parse_header(buf, s_lut, sb, nl[j], a, tile_off, &key, &start_g);
                    if (rc) {
                        raise_error(a.st, rc, o);
                        have = false;
                    }
                }
                const unsigned grp = __ballot_sync(0xFFFFFFFFu, have);
                tick(4);
                if (h0) finish_pending();  // more than one header per thread: make room for the next key
                if (have) {
                    const unsigned same = __match_any_sync(grp, key);
                    if (a.table && !(a.dbg_flags & 8) && lane == __ffs(same) - 1) {  // lowest lane = lowest ordinal
                        p_key = key, p_pos = a.pos_base + o, p_cnt = __popc(same);
                        p_slot = hash64(key) & a.table_mask;
                        p_seen = *reinterpret_cast<volatile unsigned long long*>(&a.table[p_slot].key);
                    }
                    emit(o, key, start_g);
                }
            }
        } else if (tid == 0) {
            // More newlines than the list holds (lines shorter than 16 bytes on average): exact but serial.
            unsigned long long k = K0;
            unsigned prev = s_halo_start;
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
        __syncthreads();
        tick(5);

        if (tid == 0) {
            const unsigned long long c_hi = o_end < a.read_limit ? o_end : a.read_limit;
            if (c_hi > o_first) my_reads += c_hi - o_first;
            if (is_last) a.st->line_carry = K0 + total + vnl;
            tacc[9] += 1;
        }
        __syncthreads();  // everyone is done with buffer c and the header lists
        tick(6);
        cur = nxt;
        const int old = c;
        c = n, n = nn, nn = old;
    }
    finish_pending();
    finish_pending();  // a CAS issued by the call above
    if (tid == 0 && my_reads) atomicAdd(&a.st->n_reads, my_reads);
    if (a.timing && tid == 0) {
#pragma unroll
        for (int i = 0; i < 10; ++i) atomicAdd(&a.timing[i], tacc[i]);
    }
}

}  // namespace frb
