// Native writer of the scan-results CSV (included by frb_lib.cu).  A lane has ~10^7 unique index pairs and
// Python's csv module spends about 2 us per row on them; this writes the same bytes at memory speed.
// Dialect of csv.writer's default ("excel"): fields separated by ',', rows ended by "\r\n", a field is
// quoted only if it holds ',', '"', '\r' or '\n' (quotes doubled) -- F:499 uses csv.DictWriter defaults.

namespace {

struct CsvOut {
    FILE* fh;
    std::vector<char> buf;
    size_t n = 0;
    explicit CsvOut(FILE* f) : fh(f), buf(8u << 20) {}
    bool flush() {
        const bool ok = fwrite(buf.data(), 1, n, fh) == n;
        n = 0;
        return ok;
    }
    bool room(size_t want) { return n + want <= buf.size() || flush(); }
    void put(char ch) { buf[n++] = ch; }
    void put(const char* s, size_t len) {
        memcpy(buf.data() + n, s, len);
        n += len;
    }
    void field(const char* s) {  // QUOTE_MINIMAL
        const size_t len = strlen(s);
        if (!strpbrk(s, ",\"\r\n")) {
            put(s, len);
            return;
        }
        put('"');
        for (size_t i = 0; i < len; ++i) {
            if (s[i] == '"') put('"');
            put(s[i]);
        }
        put('"');
    }
    void number(unsigned long long v) {
        char tmp[24];
        int k = 0;
        do {
            tmp[k++] = static_cast<char>('0' + v % 10);
            v /= 10;
        } while (v);
        while (k) put(tmp[--k]);
    }
};

}  // namespace

extern "C" {

// idx1,idx2,matched_idx1,matched_idx2,read_type,sample_name,reads,demux_ok -- one row per unique key, in
// the order given (first appearance).  m1/m2/srow index the string tables (-1 = empty field).
int frb_write_scan_csv(const char* path, const uint64_t* keys, const uint64_t* counts, const int32_t* m1,
                       const int32_t* m2, const uint8_t* type, const int32_t* srow, const uint8_t* ok, uint64_t n,
                       const char* const* idx1_tab, const char* const* idx2_tab, const char* const* id_tab,
                       uint32_t rows, int single_index) {
    static const char* const kTypes[4] = {"undetermined", "index_hop", "demuxable", "ambiguous"};
    FILE* fh = fopen(path, "wb");
    if (!fh) return fail(nullptr, FRB_ERR_IO, "cannot open %s for writing", path);
    size_t longest = 64;  // worst-case bytes of one row: the string tables decide
    for (uint32_t r = 0; r < rows; ++r) {
        longest = std::max(longest, 2 * strlen(idx1_tab[r]) + 2 * strlen(single_index ? "" : idx2_tab[r]) +
                                        2 * strlen(id_tab[r]) + 8);
    }
    const size_t row_max = longest + 2 * kMaxSyms + 64;
    CsvOut out(fh);
    static const char kHeader[] = "idx1,idx2,matched_idx1,matched_idx2,read_type,sample_name,reads,demux_ok\r\n";
    out.put(kHeader, sizeof(kHeader) - 1);
    bool good = true;
    for (uint64_t i = 0; i < n && good; ++i) {
        good = out.room(row_max);
        char txt[24];
        const int len = frb_unpack_key(keys[i], txt);
        // key.split("+"): first part, second part (further parts are ignored)
        int p = 0;
        while (p < len && txt[p] != '+') ++p;
        out.put(txt, p);
        out.put(',');
        if (p < len) {
            int q = p + 1;
            while (q < len && txt[q] != '+') ++q;
            out.put(txt + p + 1, q - p - 1);
        }
        out.put(',');
        if (m1[i] >= 0 && static_cast<uint32_t>(m1[i]) < rows) out.field(idx1_tab[m1[i]]);
        out.put(',');
        if (!single_index && m2[i] >= 0 && static_cast<uint32_t>(m2[i]) < rows) out.field(idx2_tab[m2[i]]);
        out.put(',');
        const char* kind = kTypes[type[i] & 3];
        out.put(kind, strlen(kind));
        out.put(',');
        if (srow[i] >= 0 && static_cast<uint32_t>(srow[i]) < rows) out.field(id_tab[srow[i]]);
        out.put(',');
        out.number(counts[i]);
        out.put(',');
        if (ok[i]) out.put("True", 4);
        else out.put("False", 5);
        out.put('\r');
        out.put('\n');
    }
    good = good && out.flush();
    good = (fclose(fh) == 0) && good;
    if (!good) return fail(nullptr, FRB_ERR_IO, "write to %s failed", path);
    return FRB_OK;
}

}  // extern "C"
