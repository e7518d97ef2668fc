/* CPU oracle in plain C for the frender scan hot path -- TEST INFRASTRUCTURE ONLY.
 *
 * A restatement of the reference's algorithms (/root/reference/frender.py, "F:") that is fast
 * enough to check the CUDA path at millions of reads.  It is checked against the Python oracle
 * (itself pinned by fixtures the reference wrote, tests/golden) in tests/test_oracle_c.py.  Only
 * tests/, __graft_entry__.smoke() and bench.py's cpu_baseline leg may load it; the product never
 * does.  Build: `make -C oracle` -> oracle/_build/liboracle.so.
 *
 *   oracle_tally     every 4th line -> key -> count, in first-appearance order    (F:154-181)
 *   oracle_classify  Hamming match + classification of one key list              (F:214-351)
 */
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#define KEY_MAX 63

typedef struct {
    char (*keys)[KEY_MAX + 1]; /* first-appearance order */
    uint64_t* counts;
    uint64_t n, cap, reads;
    uint64_t* slots;           /* open addressing: index + 1 into keys, 0 = empty */
    uint64_t nslots;
} tally_t;

static uint64_t fnv(const char* s, size_t n) {
    uint64_t h = 1469598103934665603ULL;
    for (size_t i = 0; i < n; ++i) h = (h ^ (unsigned char)s[i]) * 1099511628211ULL;
    return h;
}

static void grow(tally_t* t) {
    uint64_t ns = t->nslots ? t->nslots * 2 : 1024;
    free(t->slots);
    t->slots = calloc(ns, sizeof(uint64_t));
    t->nslots = ns;
    for (uint64_t i = 0; i < t->n; ++i) {
        uint64_t h = fnv(t->keys[i], strlen(t->keys[i])) & (ns - 1);
        while (t->slots[h]) h = (h + 1) & (ns - 1);
        t->slots[h] = i + 1;
    }
}

static void add(tally_t* t, const char* key, size_t len) {
    if (t->n * 2 >= t->nslots) grow(t);
    uint64_t h = fnv(key, len) & (t->nslots - 1);
    while (t->slots[h]) {
        const char* k = t->keys[t->slots[h] - 1];
        if (strlen(k) == len && memcmp(k, key, len) == 0) {
            t->counts[t->slots[h] - 1]++;
            return;
        }
        h = (h + 1) & (t->nslots - 1);
    }
    if (t->n == t->cap) {
        t->cap = t->cap ? t->cap * 2 : 1024;
        t->keys = realloc(t->keys, t->cap * sizeof(*t->keys));
        t->counts = realloc(t->counts, t->cap * sizeof(uint64_t));
    }
    memcpy(t->keys[t->n], key, len);
    t->keys[t->n][len] = 0;
    t->counts[t->n] = 1;
    t->slots[h] = ++t->n;
}

/* rule 0: line.rstrip("\n").split(" ")[1].split(":")[-1]   (F:169)
 * rule 1: line.split(":")[-1].rstrip("\n")                 (F:778)
 * returns 0, -1 = header without a second token (IndexError), -2 = key longer than KEY_MAX */
static int key_of(const unsigned char* s, size_t len, int rule, const unsigned char** k, size_t* klen) {
    const unsigned char *a = s, *e = s + len;
    if (rule == 0) {
        const unsigned char* sp = memchr(s, ' ', len);
        if (!sp) return -1;
        a = sp + 1;
        const unsigned char* sp2 = memchr(a, ' ', (size_t)(e - a));
        if (sp2) e = sp2;
    }
    for (const unsigned char* p = e; p > a; --p)
        if (p[-1] == ':') {
            a = p;
            break;
        }
    if ((size_t)(e - a) > KEY_MAX) return -2;
    *k = a;
    *klen = (size_t)(e - a);
    return 0;
}

void* oracle_tally(const unsigned char* data, uint64_t nbytes, int rule, uint64_t sample, int* err) {
    tally_t* t = calloc(1, sizeof(tally_t));
    uint64_t pos = 0, line = 0;
    *err = 0;
    while (pos < nbytes) {
        const unsigned char* nl = memchr(data + pos, '\n', nbytes - pos);
        uint64_t end = nl ? (uint64_t)(nl - data) : nbytes;
        if ((line & 3) == 0) {
            if (sample && t->reads >= sample) break; /* F:163-165 */
            t->reads++;
            const unsigned char* k;
            size_t klen;
            int rc = key_of(data + pos, end - pos, rule, &k, &klen);
            if (rc) {
                *err = rc;
                return t;
            }
            add(t, (const char*)k, klen);
        }
        pos = end + 1;
        line++;
    }
    return t;
}
uint64_t oracle_tally_size(void* h) { return ((tally_t*)h)->n; }
uint64_t oracle_tally_reads(void* h) { return ((tally_t*)h)->reads; }
const char* oracle_tally_key(void* h, uint64_t i) { return ((tally_t*)h)->keys[i]; }
uint64_t oracle_tally_count(void* h, uint64_t i) { return ((tally_t*)h)->counts[i]; }
void oracle_tally_free(void* h) {
    tally_t* t = h;
    free(t->keys), free(t->counts), free(t->slots), free(t);
}

static int lower(int c) { return (c >= 'A' && c <= 'Z') ? c + 32 : c; }
static int within(const char* q, const char* c, int len, int max_subs) { /* F:226-230 */
    int d = 0;
    for (int i = 0; i < len; ++i) d += lower(q[i]) != lower(c[i]);
    return d <= max_subs;
}

/* One key ("idx1+idx2[+...]") against the sheet (rows of fixed-width strings l1 / l2).
 * out[0] = first idx1 match row or -1, out[1] = first idx2 match row or -1,
 * out[2] = read type (0 undetermined, 1 index_hop, 2 demuxable, 3 ambiguous), out[3] = sample row.
 * returns 0, -1 when a length differs from the sheet (AssertionError F:227) or no '+' (F:306). */
int oracle_classify(const char* key, const char* idx1, const char* idx2, int rows, int l1, int l2, int max_subs,
                    int* out) {
    const char* plus = strchr(key, '+');
    if (!plus || (int)(plus - key) != l1) return -1;
    const char* b = plus + 1;
    const char* plus2 = strchr(b, '+');
    int blen = plus2 ? (int)(plus2 - b) : (int)strlen(b);
    if (blen != l2) return -1;
    int first1 = -1, first2 = -1, both = 0, both_row = -1;
    for (int r = 0; r < rows; ++r) {
        int m1 = within(key, idx1 + (size_t)r * l1, l1, max_subs);
        int m2 = within(b, idx2 + (size_t)r * l2, l2, max_subs);
        if (m1 && first1 < 0) first1 = r;
        if (m2 && first2 < 0) first2 = r;
        if (m1 && m2) {
            both++;
            if (both_row < 0) both_row = r;
        }
    }
    if (first1 >= 0 && first2 >= 0) { /* F:259-278 */
        out[0] = first1, out[1] = first2;
        out[2] = both == 0 ? 1 : (both == 1 ? 2 : 3);
        out[3] = both == 1 ? both_row : -1;
    } else { /* F:280-284 */
        out[0] = out[1] = out[3] = -1;
        out[2] = 0;
    }
    return 0;
}

/* Reverse complement of a fixed-width string (F:210-211): translate ATGCNatgcn, reverse; other bytes unchanged. */
static void revcomp(const char* in, int len, char* out) {
    for (int i = 0; i < len; ++i) {
        char c = in[len - 1 - i];
        switch (c) {
            case 'A': c = 'T'; break;
            case 'T': c = 'A'; break;
            case 'G': c = 'C'; break;
            case 'C': c = 'G'; break;
            case 'a': c = 't'; break;
            case 't': c = 'a'; break;
            case 'g': c = 'c'; break;
            case 'c': c = 'g'; break;
            default: break; /* N, n and everything else */
        }
        out[i] = c;
    }
}

/* analyze_barcodes_with_rc (F:294-351) for one key: forward classification into out[0..3] as oracle_classify,
 * reverse-complement classification into out[4] (first rc-idx2 match row or -1), out[5] (rc read type), out[6]
 * (rc sample row); out[0] takes the rc pass's idx1 row when the forward pass left it empty (F:319-323).
 * `group[r]` = dense id of row r's sample NAME: two demuxable verdicts with different names turn both
 * ambiguous (F:336-349).  `idx2_rc` = rows of reverse-complemented idx2 (oracle_revcomp_sheet).            */
int oracle_classify_rc(const char* key, const char* idx1, const char* idx2, const char* idx2_rc, const int* group,
                       int rows, int l1, int l2, int max_subs, int* out) {
    int fwd[4], rc[4];
    if (oracle_classify(key, idx1, idx2, rows, l1, l2, max_subs, fwd)) return -1;
    if (oracle_classify(key, idx1, idx2_rc, rows, l1, l2, max_subs, rc)) return -1;
    if (fwd[0] < 0) fwd[0] = rc[0];
    if (fwd[2] == 2 && rc[2] == 2 && group[fwd[3]] != group[rc[3]]) {
        fwd[2] = rc[2] = 3;
        fwd[3] = rc[3] = -1;
    }
    out[0] = fwd[0], out[1] = fwd[1], out[2] = fwd[2], out[3] = fwd[3];
    out[4] = rc[1], out[5] = rc[2], out[6] = rc[3];
    return 0;
}

void oracle_revcomp_sheet(const char* idx2, int rows, int l2, char* out) {
    for (int r = 0; r < rows; ++r) revcomp(idx2 + (size_t)r * l2, l2, out + (size_t)r * l2);
}

/* ---- demux loop (F:774-810) at scale: per-sink digests ------------------------------------------------------
 * Walks R1 and R2 as the reference does -- groups of four lines (F:719-723: a trailing partial group still is a
 * record), zipped, so the shorter file ends the loop (F:777) -- takes the key from the R2 header (text behind its
 * last ':' up to the line end, F:778), looks the key up (keys are given as strings, sink ids beside them) and
 * "writes" the eight lines to the sink: here that is a byte count and an order-sensitive FNV-1a digest per sink and
 * mate, which a run of the router at ANY chunking must reproduce exactly.  Returns 0, or -1 with *bad_record = the
 * first record whose key is not in the table (F:807-810). */
typedef struct {
    uint64_t bytes1, bytes2, hash1, hash2, records;
} route_sum_t;

static uint64_t fnv_more(uint64_t h, const unsigned char* p, size_t n) {
    for (size_t i = 0; i < n; i++) h = (h ^ p[i]) * 1099511628211ULL;
    return h;
}

/* one FNV-1a step per sink: hashes[s] over base[off[s], off[s + 1]) */
void oracle_fnv1a_segments(uint64_t* hashes, const unsigned char* base, const uint64_t* off, uint32_t n_sinks) {
    for (uint32_t s = 0; s < n_sinks; s++) hashes[s] = fnv_more(hashes[s], base + off[s], (size_t)(off[s + 1] - off[s]));
}

static size_t record_end(const unsigned char* d, size_t n, size_t pos) { /* behind four lines, or n */
    for (int k = 0; k < 4 && pos < n; k++) {
        const unsigned char* nl = memchr(d + pos, '\n', n - pos);
        pos = nl ? (size_t)(nl - d) + 1 : n;
    }
    return pos;
}

int oracle_route(const unsigned char* r1, uint64_t n1, const unsigned char* r2, uint64_t n2, const char* keys,
                 const uint32_t* key_off, const uint32_t* sinks, uint64_t n_keys, uint32_t n_sinks, route_sum_t* out,
                 uint64_t* bad_record) {
    uint64_t cap = 16;
    while (cap < 2 * n_keys + 2) cap <<= 1;
    int64_t* slot = malloc(cap * sizeof(int64_t));
    if (!slot) return -3;
    for (uint64_t i = 0; i < cap; i++) slot[i] = -1;
    for (uint64_t i = 0; i < n_keys; i++) { /* a repeated key keeps its last row, as a dict does (F:660) */
        const char* k = keys + key_off[i];
        const size_t len = key_off[i + 1] - key_off[i];
        uint64_t h = fnv(k, len) & (cap - 1);
        while (slot[h] >= 0) {
            const uint64_t j = (uint64_t)slot[h];
            if (key_off[j + 1] - key_off[j] == len && memcmp(keys + key_off[j], k, len) == 0) break;
            h = (h + 1) & (cap - 1);
        }
        slot[h] = (int64_t)i;
    }
    for (uint32_t s = 0; s < n_sinks; s++) {
        out[s].bytes1 = out[s].bytes2 = out[s].records = 0;
        out[s].hash1 = out[s].hash2 = 14695981039346656037ULL;
    }
    size_t p1 = 0, p2 = 0;
    uint64_t rec = 0;
    int rc = 0;
    while (p1 < n1 && p2 < n2) {
        const size_t e1 = record_end(r1, n1, p1), e2 = record_end(r2, n2, p2);
        const unsigned char* nl = memchr(r2 + p2, '\n', e2 - p2);
        size_t hend = nl ? (size_t)(nl - r2) : e2; /* header line without its '\n' */
        size_t ks = hend;
        while (ks > p2 && r2[ks - 1] != ':') ks--;
        const size_t len = hend - ks;
        uint64_t h = fnv((const char*)r2 + ks, len) & (cap - 1);
        int64_t hit = -1;
        while (slot[h] >= 0) {
            const uint64_t j = (uint64_t)slot[h];
            if (key_off[j + 1] - key_off[j] == len && memcmp(keys + key_off[j], r2 + ks, len) == 0) {
                hit = (int64_t)j;
                break;
            }
            h = (h + 1) & (cap - 1);
        }
        if (hit < 0) {
            *bad_record = rec;
            rc = -1;
            break;
        }
        route_sum_t* s = &out[sinks[hit]];
        s->bytes1 += e1 - p1, s->bytes2 += e2 - p2, s->records++;
        s->hash1 = fnv_more(s->hash1, r1 + p1, e1 - p1);
        s->hash2 = fnv_more(s->hash2, r2 + p2, e2 - p2);
        p1 = e1, p2 = e2;
        rec++;
    }
    free(slot);
    return rc;
}
