// Hot path B: Hamming matcher + classifier + reverse-complement orientation sums.
// One thread per UNIQUE key (F:610 passes the unique dict, not the reads); the packed sample
// sheet sits in shared memory and every thread streams over its rows.
//   mismatches(idx)  = popc( fold3(key ^ row) & field_mask )          (F:226-230, per symbol)
//   fold3(d)         = (d | d>>1 | d>>2) & 0b001001...  -> one bit per differing 3-bit symbol
// N is an ordinary symbol (N==N equal, N!=A); a sheet symbol outside ACGTN is code 7 and can
// never equal a read symbol.  The classifier keeps only what analyze_barcode needs (F:256-284):
// the FIRST matching row per index, the number of rows matching both, and that row.
#pragma once
#include "common.cuh"

namespace frb {

constexpr unsigned long long kFoldLsb = 0x1249249249249249ULL;  // bit 0 of each 3-bit field
constexpr int kMatchThreads = 256;

struct MatchArgs {
    const unsigned long long* keys;    // [n] packed unique keys
    const unsigned long long* counts;  // [n] reads per key
    unsigned long long n;
    const unsigned long long* sheet_fwd;  // [rows] pack(idx1+idx2)
    const unsigned long long* sheet_rc;   // [rows] pack(idx1+revcomp(idx2))
    const int* group;                     // [rows] sample-name group
    const unsigned char* use_rc;          // [rows] or null: pick sheet_rc for that row (F:618-623)
    unsigned rows, l1, l2, n_subs;
    int rc_mode;
    int *m1, *m2, *srow, *m2rc, *srowrc;  // per-key outputs (device, may be null)
    unsigned char *type, *typerc;
    unsigned long long *f_sum, *rc_sum;   // [rows] per group (rc_mode only)
    unsigned* work;                       // [n] indices of the keys with an idx1 match
    unsigned long long* work_n;           // entries in work[]
    DevState* st;
};

__device__ __forceinline__ unsigned long long fold3(unsigned long long d) {
    return (d | (d >> 1) | (d >> 2)) & kFoldLsb;
}

__device__ __forceinline__ void classify_one(int first1, int first2, unsigned both, int both_row, int* m1, int* m2,
                                             int* type, int* srow) {
    if (first1 >= 0 && first2 >= 0) {  // F:259-278
        *m1 = first1;
        *m2 = first2;
        *type = both == 0 ? FRB_TYPE_INDEX_HOP : (both == 1 ? FRB_TYPE_DEMUXABLE : FRB_TYPE_AMBIGUOUS);
        *srow = both == 1 ? both_row : -1;
    } else {  // F:280-284
        *m1 = -1, *m2 = -1, *type = FRB_TYPE_UNDETERMINED, *srow = -1;
    }
}

// ---- kernel 1: idx1 sweep over ALL unique keys, fully convergent --------------------------------
// A key whose idx1 matches no row is undetermined whatever its idx2 does (F:259, F:280-284) -- the fate
// of almost every unique key of a real lane (random index pairs dominate the unique set).  Those keys
// are finished here; the others go to a compact work list so that the expensive idx2 / reverse-
// complement sweep (kernel 2) runs with full warps instead of a few live lanes per warp.
__global__ void __launch_bounds__(kMatchThreads) match_idx1_kernel(const MatchArgs a) {
    extern __shared__ __align__(16) unsigned char msmem[];
    unsigned long long* s_a = reinterpret_cast<unsigned long long*>(msmem);
    for (unsigned r = threadIdx.x; r < a.rows; r += blockDim.x) s_a[r] = a.sheet_fwd[r];
    __syncthreads();
    const unsigned long long B1 = ((1ULL << (3 * a.l1)) - 1) & kFoldLsb;
    const unsigned n_subs = a.n_subs;
    const int lane = threadIdx.x & 31;
    const unsigned long long stride = static_cast<unsigned long long>(gridDim.x) * blockDim.x;
    const unsigned long long n_up = (a.n + 31) & ~31ULL;
    for (unsigned long long i = blockIdx.x * static_cast<unsigned long long>(blockDim.x) + threadIdx.x; i < n_up;
         i += stride) {
        int first1 = -1;
        if (i < a.n) {
            const unsigned long long key = a.keys[i];
            // shape check: idx1 of l1 symbols in 1..5, then '+' and idx2 of l2 symbols in 1..5, then end
            // or another '+' (extra parts ignored, F:306); single index: l1 symbols then end.
            if (a.rows) {
                bool ok = true;
                const unsigned total = a.l2 ? a.l1 + 1 + a.l2 : a.l1;
                for (unsigned s = 0; s < total; ++s) {
                    const unsigned sym = static_cast<unsigned>(key >> (3 * s)) & 7u;
                    if (a.l2 && s == a.l1) ok &= (sym == 6);
                    else ok &= (sym >= 1 && sym <= 5);
                }
                if (total < kMaxSyms) {
                    const unsigned nxt = static_cast<unsigned>(key >> (3 * total)) & 7u;
                    ok &= a.l2 ? (nxt == 0 || nxt == 6) : (nxt == 0);
                }
                if (!ok) raise_error(a.st, FRB_ERR_BAD_LENGTH, key);
            }
            if (a.l1 <= 10) {  // idx1 fits 30 bits: 32-bit compare, one LOP3 folds the three bit planes
                const unsigned m = (1u << (3 * a.l1)) - 1u;
                const unsigned k1 = static_cast<unsigned>(key) & m, b1 = static_cast<unsigned>(B1);
                for (unsigned r = 0; r < a.rows; ++r) {
                    const unsigned d = k1 ^ (static_cast<unsigned>(s_a[r]) & m);
                    const bool h1 = static_cast<unsigned>(__popc((d | (d >> 1) | (d >> 2)) & b1)) <= n_subs;
                    if (h1 && first1 < 0) first1 = r;
                }
            } else {
                for (unsigned r = 0; r < a.rows; ++r) {
                    const bool h1 = static_cast<unsigned>(__popcll(fold3(key ^ s_a[r]) & B1)) <= n_subs;
                    if (h1 && first1 < 0) first1 = r;
                }
            }
            a.m1[i] = first1;
            if (first1 < 0) {
                a.m2[i] = -1, a.srow[i] = -1, a.type[i] = FRB_TYPE_UNDETERMINED;
                if (a.rc_mode) a.m2rc[i] = -1, a.srowrc[i] = -1, a.typerc[i] = FRB_TYPE_UNDETERMINED;
            }
        }
        const unsigned need = __ballot_sync(0xFFFFFFFFu, first1 >= 0);
        if (need) {
            unsigned long long base = 0;
            const int leader = __ffs(need) - 1;
            if (lane == leader) base = atomicAdd(a.work_n, static_cast<unsigned long long>(__popc(need)));
            base = __shfl_sync(0xFFFFFFFFu, base, leader);
            if (first1 >= 0) a.work[base + __popc(need & ((1u << lane) - 1u))] = static_cast<unsigned>(i);
        }
    }
}

// ---- kernel 2: full classification of the keys on the work list ---------------------------------
__global__ void __launch_bounds__(kMatchThreads) match_kernel(const MatchArgs a) {
    extern __shared__ __align__(16) unsigned char msmem[];
    unsigned long long* s_a = reinterpret_cast<unsigned long long*>(msmem);  // idx2 as supplied / oriented
    unsigned long long* s_b = s_a + a.rows;                                  // idx2 reverse-complemented
    unsigned long long* s_f = s_b + a.rows;                                  // per-group forward reads
    unsigned long long* s_r = s_f + a.rows;                                  // per-group rc reads
    int* s_g = reinterpret_cast<int*>(s_r + a.rows);

    for (unsigned r = threadIdx.x; r < a.rows; r += blockDim.x) {
        const bool flip = a.use_rc && a.use_rc[r];
        s_a[r] = flip ? a.sheet_rc[r] : a.sheet_fwd[r];
        s_b[r] = a.sheet_rc[r];
        s_f[r] = 0, s_r[r] = 0;
        s_g[r] = a.group[r];
    }
    __syncthreads();

    // field masks in "one bit per symbol" space
    const unsigned long long B1 = ((1ULL << (3 * a.l1)) - 1) & kFoldLsb;
    const unsigned long long B2 = a.l2 ? ((((1ULL << (3 * a.l2)) - 1) & kFoldLsb) << (3 * (a.l1 + 1))) : 0ULL;
    const unsigned n_subs = a.n_subs;
    const unsigned long long n_work = *a.work_n;

    const unsigned long long stride = static_cast<unsigned long long>(gridDim.x) * blockDim.x;
    for (unsigned long long j = blockIdx.x * static_cast<unsigned long long>(blockDim.x) + threadIdx.x; j < n_work;
         j += stride) {
        const unsigned long long i = a.work[j];
        const unsigned long long key = a.keys[i];
        // first idx1 row: from kernel 1, or from the rc pass when this is the oriented second pass of a
        // scan (F:628) -- orientation only changes idx2, and a key the rc pass left without an
        // idx1+idx2 match in either orientation cannot get one from a per-row choice between the two.
        const int first1 = a.m1[i];
        int firstf = -1, firstr = -1, rowf = -1, rowr = -1;
        unsigned nf = 0, nr = 0;
        if (first1 >= 0) {
            for (unsigned r = 0; r < a.rows; ++r) {
                const unsigned long long df = fold3(key ^ s_a[r]);
                const bool h1 = static_cast<unsigned>(__popcll(df & B1)) <= n_subs;
                const bool h2 = a.l2 ? (static_cast<unsigned>(__popcll(df & B2)) <= n_subs) : true;
                if (h2 && firstf < 0) firstf = r;
                if (h1 && h2) {
                    ++nf;
                    if (rowf < 0) rowf = r;
                }
                if (a.rc_mode) {
                    const unsigned long long dr = fold3(key ^ s_b[r]);
                    const bool h3 = static_cast<unsigned>(__popcll(dr & B2)) <= n_subs;
                    if (h3 && firstr < 0) firstr = r;
                    if (h1 && h3) {
                        ++nr;
                        if (rowr < 0) rowr = r;
                    }
                }
            }
        }
        int m1, m2, type, srow;
        classify_one(first1, firstf, nf, rowf, &m1, &m2, &type, &srow);
        if (a.rc_mode) {
            int m1r, m2r, typer, srowr;
            classify_one(first1, firstr, nr, rowr, &m1r, &m2r, &typer, &srowr);
            if (m1 < 0) m1 = m1r;  // F:319-323
            if (type == FRB_TYPE_DEMUXABLE && typer == FRB_TYPE_DEMUXABLE && s_g[srow] != s_g[srowr]) {
                type = typer = FRB_TYPE_AMBIGUOUS;  // F:336-349
                srow = srowr = -1;
            }
            const unsigned long long c = a.counts[i];
            if (srow >= 0) atomicAdd(&s_f[s_g[srow]], c);      // F:370-371
            if (srowr >= 0) atomicAdd(&s_r[s_g[srowr]], c);    // F:372-373
            a.m2rc[i] = m2r;
            a.typerc[i] = static_cast<unsigned char>(typer);
            a.srowrc[i] = srowr;
        }
        a.m1[i] = m1;
        a.m2[i] = m2;
        a.type[i] = static_cast<unsigned char>(type);
        a.srow[i] = srow;
    }
    if (a.rc_mode) {
        __syncthreads();
        for (unsigned g = threadIdx.x; g < a.rows; g += blockDim.x) {
            if (s_f[g]) atomicAdd(&a.f_sum[g], s_f[g]);
            if (s_r[g]) atomicAdd(&a.rc_sum[g], s_r[g]);
        }
    }
}

}  // namespace frb
