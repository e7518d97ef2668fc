// Host side of the device inflate (gz_kernels.cuh), included by frb_lib.cu.
//
// A .gz file is read in pieces of compressed bytes; every piece goes to the device as it is (a fifth of the
// bytes a host-inflated file would send), is inflated there by thousands of warps at once and handed on where
// it lies -- the scan kernel reads it from HBM.  Anything the device path does not take (a block start that
// turned out false, blocks larger than the chunk stride, '\r' in the text, staging overflow) makes
// gz_device_inflate return FRB_GZ_RETRY_HOST and the caller inflates that file with zlib as before.
namespace {

constexpr int FRB_GZ_RETRY_HOST = 1000;   // internal status: use the host (zlib) path for this file
constexpr int FRB_GZ_RETRY_SPACE = 1001;  // internal status: a chunk's staging area was too small

// FRB_GZ_VERBOSE: why the device path handed a file back
int gz_decline(const char* why) {
    if (getenv("FRB_GZ_VERBOSE")) fprintf(stderr, "gz: device inflate declined: %s\n", why);
    return FRB_GZ_RETRY_HOST;
}

struct GzBuffers {
    unsigned char* comp[2] = {nullptr, nullptr};  // pieces of the compressed file (+ overlap, padded to words)
    unsigned char* host[2] = {nullptr, nullptr};  // pinned twins: a reader thread fills one while the other is in use
    cudaEvent_t copied[2] = {nullptr, nullptr};   // H2D of the piece done
    size_t comp_cap = 0;
    unsigned short* stage = nullptr;      // 16-bit symbols, one area per chunk
    size_t stage_syms = 0;
    gz::Chunk* chunks = nullptr;
    gz::Chunk* chunks_host = nullptr;     // pinned (first / last entries only are looked at)
    size_t chunk_cap = 0;
    unsigned char* windows = nullptr;     // 32 KiB in front of every chunk
    unsigned short* maps = nullptr;       // per chunk: window in front of its group -> window behind it
    unsigned char* group_win = nullptr;   // 32 KiB in front of every group of chunks
    int* prev_found = nullptr;
    unsigned char* win[2] = {nullptr, nullptr};  // 32 KiB in front of / behind the piece
    unsigned char* out[2] = {nullptr, nullptr};  // inflated text, two pieces in flight (carry area in front)
    size_t out_cap = 0;
    gz::Trailer* trailers = nullptr;      // member trailers the decoder passed in this piece
    size_t trailer_cap = 0;
    unsigned long long* tr_end[2] = {nullptr, nullptr};  // their ends in the text, unsorted / sorted
    unsigned* tr_idx[2] = {nullptr, nullptr};
    unsigned* tr_crc = nullptr;           // crc32 of the text up to every trailer end
    unsigned* crc_blk = nullptr;          // crc32 of every 4 KiB block of the piece's text, and their prefix
    unsigned* crc_pre = nullptr;
    // device: [0] total symbols, [1] last newline end; [2..3] as unsigned: bad, cr, fail; [4] as unsigned: trailers;
    // [5..6] as unsigned: members whose CRC32 / ISIZE disagree, crc and length (lo, hi) of the open member
    unsigned long long* scalars = nullptr;
    unsigned long long* scalars_host = nullptr;  // pinned
};

struct GzConfig {
    size_t piece = 128u << 20;   // compressed bytes per piece
    size_t stride = 48u << 10;   // compressed bytes per chunk (one warp): the finder's time goes with the number of chunks,
                                 // the decode kernel's does not until there are fewer chunks than warp slots (measured: 16 KiB
                                 // 10.6, 32 KiB 13.4, 48 KiB 14.9, 64 KiB 13.3, 96 KiB 12.0 GB/s inflated)
    size_t expand = 12;          // staging symbols per compressed byte
    size_t carry = 4u << 20;     // room for an unfinished line in front of a piece's text
    bool stride_fixed = false;   // FRB_GZ_STRIDE_KB given: no per-file choice
};

GzConfig gz_config() {
    GzConfig g;
    if (const char* e = getenv("FRB_GZ_PIECE_MB")) g.piece = static_cast<size_t>(atoi(e)) << 20;
    if (const char* e = getenv("FRB_GZ_STRIDE_KB")) g.stride = static_cast<size_t>(atoi(e)) << 10, g.stride_fixed = true;
    if (const char* e = getenv("FRB_GZ_EXPAND")) g.expand = static_cast<size_t>(atoi(e));
    return g;
}

void gz_free(GzBuffers& b) {
    for (int i = 0; i < 2; ++i) {
        cudaFree(b.comp[i]), cudaFreeHost(b.host[i]);
        if (b.copied[i]) cudaEventDestroy(b.copied[i]);
    }
    cudaFree(b.stage), cudaFree(b.chunks), cudaFreeHost(b.chunks_host);
    cudaFree(b.maps), cudaFree(b.group_win), cudaFree(b.prev_found);
    cudaFree(b.windows), cudaFree(b.win[0]), cudaFree(b.win[1]), cudaFree(b.out[0]), cudaFree(b.out[1]);
    cudaFree(b.trailers), cudaFree(b.tr_end[0]), cudaFree(b.tr_end[1]), cudaFree(b.tr_idx[0]), cudaFree(b.tr_idx[1]);
    cudaFree(b.tr_crc), cudaFree(b.crc_blk), cudaFree(b.crc_pre);
    cudaFree(b.scalars), cudaFreeHost(b.scalars_host);
    b = GzBuffers{};
}

int gz_ensure(frb_ctx* c, GzBuffers& b, const GzConfig& g, size_t piece_bytes) {
    // sizes in steps of 16 MiB, so that a run of files of similar size keeps one set of buffers (allocating and
    // freeing pinned and device memory synchronises the device and costs more than a small file's inflate)
    piece_bytes = (piece_bytes + (16u << 20) - 1) & ~static_cast<size_t>((16u << 20) - 1);
    const size_t comp_cap = piece_bytes + 3 * std::max<size_t>(g.stride, 32u << 10) + 64;
    const size_t n_chunks = (piece_bytes + g.stride - 1) / g.stride + 1;
    const size_t stage_syms = std::max<size_t>((n_chunks - 1) * g.stride * g.expand, 32u << 20);  // >= 64 MB of symbols
    const size_t out_cap = g.carry + stage_syms + 64;
    if (comp_cap <= b.comp_cap && n_chunks <= b.chunk_cap && stage_syms <= b.stage_syms && out_cap <= b.out_cap) return FRB_OK;
    CU(c, cudaStreamSynchronize(c->compute));
    gz_free(b);
    for (int i = 0; i < 2; ++i) {
        CU(c, cudaMalloc(&b.comp[i], comp_cap));
        CU(c, cudaMallocHost(&b.host[i], comp_cap));
        CU(c, cudaEventCreateWithFlags(&b.copied[i], cudaEventDisableTiming));
    }
    CU(c, cudaMalloc(&b.stage, stage_syms * 2));
    CU(c, cudaMalloc(&b.chunks, n_chunks * sizeof(gz::Chunk)));
    CU(c, cudaMallocHost(&b.chunks_host, n_chunks * sizeof(gz::Chunk)));
    CU(c, cudaMalloc(&b.windows, n_chunks * gz::kWindow));
    CU(c, cudaMalloc(&b.maps, n_chunks * gz::kWindow * 2));
    CU(c, cudaMalloc(&b.group_win, (n_chunks / gz::kGroup + 1) * gz::kWindow));
    CU(c, cudaMalloc(&b.prev_found, n_chunks * sizeof(int)));
    for (int i = 0; i < 2; ++i) {
        CU(c, cudaMalloc(&b.win[i], gz::kWindow));
        CU(c, cudaMalloc(&b.out[i], out_cap));
    }
    const size_t trailer_cap = 4 * n_chunks + 4096;  // BGZF: about two members per 32 KiB of compressed bytes
    CU(c, cudaMalloc(&b.trailers, trailer_cap * sizeof(gz::Trailer)));
    for (int i = 0; i < 2; ++i) {
        CU(c, cudaMalloc(&b.tr_end[i], trailer_cap * 8));
        CU(c, cudaMalloc(&b.tr_idx[i], trailer_cap * 4));
    }
    CU(c, cudaMalloc(&b.tr_crc, trailer_cap * 4));
    CU(c, cudaMalloc(&b.crc_blk, (out_cap / gz::kCrcBlock + 2) * 4));
    CU(c, cudaMalloc(&b.crc_pre, (out_cap / gz::kCrcBlock + 2) * 4));
    b.trailer_cap = trailer_cap;
    CU(c, cudaMalloc(&b.scalars, 64));
    CU(c, cudaMallocHost(&b.scalars_host, 64));
    b.comp_cap = comp_cap, b.chunk_cap = n_chunks, b.stage_syms = stage_syms, b.out_cap = out_cap;
    CU(c, cudaFuncSetAttribute(gz::gz_tailmap_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 4 * gz::kWindow));
    CU(c, cudaFuncSetAttribute(gz::gz_groupwin_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 2 * gz::kWindow));
    return FRB_OK;
}

// FRB_GZ_TIMING=1: device time of every step of a piece on stderr
struct GzTimer {
    frb_ctx* c;
    bool on;
    cudaEvent_t ev[8];
    explicit GzTimer(frb_ctx* ctx) : c(ctx), on(getenv("FRB_GZ_TIMING") != nullptr) {
        if (on)
            for (auto& e : ev) cudaEventCreate(&e);
    }
    ~GzTimer() {
        if (on)
            for (auto& e : ev) cudaEventDestroy(e);
    }
    void mark(int k) {
        if (on) cudaEventRecord(ev[k], c->compute);
    }
    void report(int piece, size_t n_in, uint64_t n_sym, unsigned n_chunks) {
        if (!on) return;
        float t[7] = {0, 0, 0, 0, 0, 0, 0};
        for (int k = 0; k < 7; ++k)
            if (k != 5) cudaEventElapsedTime(&t[k], ev[k], ev[k + 1]);
        fprintf(stderr, "gz piece %d: %zu -> %llu bytes, %u chunks: find %.2f link %.2f decode %.2f offsets %.2f window %.2f "
                        "resolve %.2f ms\n", piece, n_in, (unsigned long long)n_sym, n_chunks, t[0], t[1], t[2], t[3], t[4], t[6]);
    }
};

// x^(2^k) mod P, k < 32, for the CRC-32 kernels (reflected polynomial, bit 31 = x^0)
gz::CrcPowers gz_crc_powers() {
    auto mul = [](unsigned a, unsigned b) {
        unsigned p = 0;
        for (int k = 0; k < 32; ++k) {
            if ((a >> (31 - k)) & 1u) p ^= b;
            b = (b & 1u) ? (b >> 1) ^ 0xEDB88320u : b >> 1;
        }
        return p;
    };
    gz::CrcPowers pw;
    pw.x2n[0] = 0x40000000u;  // x^1
    for (int k = 1; k < 32; ++k) pw.x2n[k] = mul(pw.x2n[k - 1], pw.x2n[k - 1]);
    return pw;
}

// gzip member header in host memory (RFC 1952): bytes of the header, 0 = not gzip / incomplete
size_t gz_host_header(const unsigned char* d, size_t n) {
    if (n < 10 || d[0] != 0x1f || d[1] != 0x8b || d[2] != 8 || (d[3] & 0xE0)) return 0;
    const unsigned flg = d[3];
    size_t pos = 10;
    if (flg & 4) {
        if (pos + 2 > n) return 0;
        pos += 2 + (d[pos] | (d[pos + 1] << 8));
    }
    for (int k = 0; k < 2; ++k)
        if (flg & (8 << k)) {
            while (pos < n && d[pos]) ++pos;
            ++pos;
        }
    if (flg & 2) pos += 2;
    return pos <= n ? pos : 0;
}

// One or more .gz files read as ONE stream: gzip files laid end to end are a gzip stream of several members whose
// text is the files' texts laid end to end (RFC 1952), so a run of small files is inflated by one set of launches.
struct GzSource {
    std::vector<std::string> paths;
    std::vector<uint64_t> start;  // start[i] = stream offset of file i, start[n] = total bytes
    int open(frb_ctx* c) {
        start.assign(1, 0);
        for (const auto& p : paths) {
            FILE* fh = fopen(p.c_str(), "rb");
            if (!fh) return fail(c, FRB_ERR_IO, "cannot open %s", p.c_str());
            fseek(fh, 0, SEEK_END);
            start.push_back(start.back() + static_cast<uint64_t>(ftell(fh)));
            fclose(fh);
        }
        return FRB_OK;
    }
    uint64_t size() const { return start.back(); }
    bool read(uint64_t off, size_t n, unsigned char* dst) const {
        for (size_t i = 0; i < paths.size() && n; ++i) {
            if (off >= start[i + 1]) continue;
            const uint64_t in_file = off - start[i];
            const size_t take = static_cast<size_t>(std::min<uint64_t>(n, start[i + 1] - off));
            FILE* fh = fopen(paths[i].c_str(), "rb");
            if (!fh) return false;
            const bool ok = fseek(fh, static_cast<long>(in_file), SEEK_SET) == 0 && fread(dst, 1, take, fh) == take;
            fclose(fh);
            if (!ok) return false;
            off += take, dst += take, n -= take;
        }
        return n == 0;
    }
};

// a member of the stream that ended in the piece just inflated: where its text ends and where its trailer ends
// (both counted from the start of the stream)
struct GzMemberEnd {
    uint64_t text_end, comp_end;
};

// Inflate `path` on the device piece by piece; every piece of text (cut behind its last complete line, the rest
// carried into the next piece; the last piece whole) goes to `sink(dev_ptr, nbytes, last)`, 16-byte aligned,
// valid until the sink of the piece after next is called.  FRB_GZ_RETRY_HOST: nothing usable happened (the
// sink may already have been called -- the caller starts the file over).
// on_members(vector<GzMemberEnd>) is called for every piece before its sink (a batch of files needs to know which
// text belongs to which file); `path` names the stream in messages.
template <typename Sink, typename Members>
int gz_device_inflate_once(frb_ctx* c, GzBuffers& b, const GzConfig& g_in, const GzSource& src, uint64_t* raw_bytes,
                           Sink&& sink, Members&& on_members, bool want_members) {
    const char* const path = src.paths[0].c_str();
    const uint64_t file_bytes = src.size();
    if (file_bytes < 18) return gz_decline("smaller than an empty member");  // smaller than an empty member: let zlib say what it is
    // One warp decodes one chunk from end to end, so the decode of a piece takes as long as ONE chunk does: a small
    // file is cut into smaller chunks (down to 8 KiB) until there are a few thousand of them.
    GzConfig g = g_in;
    const uint64_t overlap = 3 * std::max<uint64_t>(g_in.stride, 32u << 10);  // compressed bytes read behind a piece
    if (!g.stride_fixed)
        while (g.stride / 2 >= (8u << 10) && file_bytes / g.stride < 2048) g.stride >>= 1;
    // Pieces begin at fixed file offsets k * piece (so that they can be read ahead); the last one takes up to a
    // piece and a half rather than leaving a small rest for a launch of its own.
    const size_t piece = std::min<uint64_t>(g.piece, (file_bytes + 3) & ~3ull);
    std::vector<std::pair<uint64_t, uint64_t>> plan;  // (file offset, nominal bytes)
    for (uint64_t off = 0; off < file_bytes;) {
        const uint64_t left = file_bytes - off;
        const uint64_t take = left <= piece + piece / 2 ? left : piece;
        plan.emplace_back(off, take);
        off += take;
    }
    uint64_t largest = 0;
    for (const auto& p : plan) largest = std::max(largest, p.second);
    TRY(gz_ensure(c, b, g, static_cast<size_t>((largest + 3) & ~3ull)));
    unsigned char head[1024];
    const size_t got_head = static_cast<size_t>(std::min<uint64_t>(sizeof head, file_bytes));
    if (!src.read(0, got_head, head)) return fail(c, FRB_ERR_IO, "%s: read error", path);
    const size_t hdr = gz_host_header(head, got_head);
    if (!hdr) return gz_decline("no gzip header");

    // ---- reader thread: piece k of the file into pinned buffer k & 1, then to the device on the copy stream ---
    std::mutex mu;
    std::condition_variable cv;
    int filled = 0;        // pieces read and queued for copying
    int released = 0;      // pieces whose buffers the consumer is done with
    bool quit = false;
    std::string io_err;
    std::thread reader([&] {
        cudaSetDevice(c->device);
        for (size_t k = 0; k < plan.size(); ++k) {
            {
                std::unique_lock<std::mutex> lk(mu);
                cv.wait(lk, [&] { return quit || static_cast<int>(k) < released + 2; });
                if (quit) return;
            }
            const uint64_t off = plan[k].first;
            const uint64_t end = std::min<uint64_t>(off + plan[k].second + overlap, file_bytes);
            const size_t n = static_cast<size_t>(end - off);
            unsigned char* const hb = b.host[k & 1];
            bool ok = src.read(off, n, hb);
            memset(hb + n, 0, 64);
            if (ok) {
                ok = cudaMemcpyAsync(b.comp[k & 1], hb, n + 64, cudaMemcpyHostToDevice, c->copy) == cudaSuccess &&
                     cudaEventRecord(b.copied[k & 1], c->copy) == cudaSuccess;
            }
            {
                std::lock_guard<std::mutex> lk(mu);
                if (!ok) io_err = "read error";
                ++filled;
            }
            cv.notify_all();
            if (!ok) return;
        }
    });
    struct Joiner {
        std::thread& t;
        std::mutex& mu;
        std::condition_variable& cv;
        bool& quit;
        ~Joiner() {
            {
                std::lock_guard<std::mutex> lk(mu);
                quit = true;
            }
            cv.notify_all();
            t.join();
        }
    } joiner{reader, mu, cv, quit};

    uint64_t start_bit_abs = hdr * 8ull;   // where the next piece's first chunk starts (absolute bit in the file)
    bool member_start = true;
    size_t carry_len = 0;                  // bytes of an unfinished line in front of the next piece's text
    int ob = 0;                            // output buffer in use
    uint64_t total_out = 0;
    unsigned member_crc = 0;               // crc32 / bytes of the member that is open at the start of the next piece
    uint64_t member_len = 0;
    static const gz::CrcPowers crc_pw = gz_crc_powers();
    CU(c, cudaMemsetAsync(b.win[0], 0, gz::kWindow, c->compute));
    for (int piece_no = 0;; ++piece_no) {
        // ---- compressed bytes of the piece (+ overlap) -----------------------------------------------------
        const uint64_t file_pos = plan[piece_no].first;
        const bool last_piece = static_cast<size_t>(piece_no) + 1 == plan.size();
        const size_t n_in = static_cast<size_t>(std::min<uint64_t>(file_pos + plan[piece_no].second + overlap, file_bytes) - file_pos);
        {
            std::unique_lock<std::mutex> lk(mu);
            cv.wait(lk, [&] { return filled > piece_no || !io_err.empty(); });
            if (!io_err.empty()) return fail(c, FRB_ERR_IO, "%s: %s", path, io_err.c_str());
        }
        unsigned char* const comp = b.comp[piece_no & 1];
        CU(c, cudaStreamWaitEvent(c->compute, b.copied[piece_no & 1], 0));
        const size_t body = last_piece ? n_in : static_cast<size_t>(plan[piece_no].second);
        const unsigned n_chunks = static_cast<unsigned>((body + g.stride - 1) / g.stride);
        const uint64_t rel_start = start_bit_abs - file_pos * 8;
        const unsigned first_chunk = static_cast<unsigned>((rel_start >> 3) / g.stride);  // chunk holding the start
        if (first_chunk >= n_chunks) return gz_decline("first block start behind the piece");
        // ---- chunk table --------------------------------------------------------------------------------
        CU(c, cudaMemsetAsync(b.chunks, 0, (n_chunks + 1) * sizeof(gz::Chunk), c->compute));
        gz::Chunk first{};
        first.start_bit = rel_start, first.found = 1, first.member_start = member_start ? 1u : 0u;
        b.chunks_host[0] = first;
        CU(c, cudaMemcpyAsync(b.chunks + first_chunk, b.chunks_host, sizeof(gz::Chunk), cudaMemcpyHostToDevice, c->compute));
        CU(c, cudaMemsetAsync(b.scalars, 0, 64, c->compute));
        unsigned* const flags = reinterpret_cast<unsigned*>(b.scalars + 2);  // bad, cr, fail
        const unsigned search_from = first_chunk + 1, search_to = n_chunks + (last_piece ? 0u : 1u);
        GzTimer tm(c);
        {
            ProfScope ps(c, FRB_K_INFLATE);
            tm.mark(0);
            if (search_to > search_from)
                gz::gz_find_kernel<<<search_to - search_from, gz::kFindThreads, 0, c->compute>>>(
                    comp, n_in, b.chunks, search_to, g.stride, 0, search_from, last_piece ? 0xFFFFFFFFu : n_chunks,
                    overlap - 8192);
            tm.mark(1);
            // every chunk of this piece an equal share of the staging area
            gz::gz_link_kernel<<<1, 32, 0, c->compute>>>(b.chunks, n_chunks, b.stage_syms / n_chunks / 16 * 16,
                                                         last_piece ? 0 : 1, flags + 2);
            tm.mark(2);
            gz::gz_decode_kernel<<<(n_chunks + gz::kDecodeWarps - 1) / gz::kDecodeWarps, gz::kDecodeWarps * 32, 0, c->compute>>>(
                comp, n_in, b.chunks, n_chunks, b.stage, b.trailers, static_cast<unsigned>(b.trailer_cap),
                reinterpret_cast<unsigned*>(b.scalars + 4), flags + 2);
            tm.mark(3);
            gz::gz_offsets_kernel<<<1, 1024, 0, c->compute>>>(b.chunks, n_chunks, b.scalars, flags);
            tm.mark(4);
            gz::gz_tailmap_kernel<<<(n_chunks + gz::kGroup - 1) / gz::kGroup, 1024, 4 * gz::kWindow, c->compute>>>(
                b.chunks, n_chunks, b.stage, b.maps, b.prev_found);
            gz::gz_groupwin_kernel<<<1, 1024, 2 * gz::kWindow, c->compute>>>(b.chunks, n_chunks, b.maps, b.win[piece_no & 1],
                                                                            b.group_win, b.win[(piece_no + 1) & 1]);
            gz::gz_windows_kernel<<<n_chunks, 256, 0, c->compute>>>(b.chunks, b.maps, b.prev_found, b.group_win, b.windows);
            tm.mark(5);
            c->launches += 7;
        }
        CU(c, cudaMemcpyAsync(b.scalars_host, b.scalars, 64, cudaMemcpyDeviceToHost, c->compute));
        CU(c, cudaMemcpyAsync(b.chunks_host + 1, b.chunks + n_chunks, sizeof(gz::Chunk), cudaMemcpyDeviceToHost, c->compute));
        CU(c, cudaStreamSynchronize(c->compute));
        {   // the decode kernel was the last reader of the compressed piece: its buffers may be refilled
            std::lock_guard<std::mutex> lk(mu);
            released = piece_no + 1;
        }
        cv.notify_all();
        const unsigned* hflags = reinterpret_cast<const unsigned*>(b.scalars_host + 2);
        const uint64_t n_sym = b.scalars_host[0];
        if (getenv("FRB_GZ_VERBOSE")) {  // how long the chunks really are: the longest one is the decode kernel's time
            std::vector<gz::Chunk> all(n_chunks);
            cudaMemcpy(all.data(), b.chunks, n_chunks * sizeof(gz::Chunk), cudaMemcpyDeviceToHost);
            unsigned found = 0, longest_at = 0;
            uint64_t longest = 0, longest_out = 0;
            for (unsigned k = 0; k < n_chunks; ++k) {
                if (!all[k].found) continue;
                ++found;
                const uint64_t len = (all[k].end_bit - all[k].start_bit) / 8;
                if (len > longest) longest = len, longest_at = k, longest_out = all[k].n_out;
            }
            fprintf(stderr, "gz: piece %d: %u of %u chunks have a block start; longest chunk %llu compressed bytes -> %llu "
                            "symbols (chunk %u), stride %zu\n", piece_no, found, n_chunks, (unsigned long long)longest,
                    (unsigned long long)longest_out, longest_at, g.stride);
        }
        if (hflags[2]) return gz_decline("no block start found behind the piece");  // (nothing was decoded)
        if (hflags[0] != 0xFFFFFFFFu) {  // a chunk failed: which way?
            gz::Chunk bad;
            CU(c, cudaMemcpy(&bad, b.chunks + hflags[0], sizeof bad, cudaMemcpyDeviceToHost));
            if (getenv("FRB_GZ_VERBOSE"))
                fprintf(stderr, "gz: piece %d chunk %u of %u: status %d, start %llu stop %llu end %llu, %u symbols (cap %u)\n",
                        piece_no, hflags[0], n_chunks, bad.status, bad.start_bit, bad.stop_bit, bad.end_bit, bad.n_out,
                        bad.stage_cap);
            // corrupt or truncated data in the piece the stream really ends in is an error of the file, not of the
            // chunking (every earlier chunk met its successor exactly)
            if (bad.status == gz::GZ_ERR_TRUNC && last_piece) return fail(c, FRB_ERR_IO, "%s: unexpected end of file", path);
            return bad.status == gz::GZ_ERR_SPACE ? FRB_GZ_RETRY_SPACE : gz_decline("a chunk did not decode to its successor's start");
        }
        const unsigned n_tr = *reinterpret_cast<const unsigned*>(b.scalars_host + 4);
        if (n_tr > b.trailer_cap) return gz_decline("too many members in one piece");
        if (carry_len + n_sym + 64 > b.out_cap) return gz_decline("text larger than the output buffer");
        // ---- symbols -> text behind the carried line ---------------------------------------------------------
        unsigned char* const text = b.out[ob] + carry_len;
        {
            ProfScope ps(c, FRB_K_INFLATE);
            const unsigned per_chunk_blocks = static_cast<unsigned>(std::max<size_t>(1, g.stride * 4 / 4096));
            tm.mark(6);
            gz::gz_resolve_kernel<<<dim3(per_chunk_blocks, n_chunks), 256, 0, c->compute>>>(b.chunks, b.stage, b.windows, text,
                                                                                         n_sym);
            tm.mark(7);
            // CRC32 + ISIZE of every member that ends in this piece (RFC 1952), on the text where it lies
            unsigned* const crc_res = reinterpret_cast<unsigned*>(b.scalars + 5);
            gz::gz_crc_blocks_kernel<<<c->sm_count * 8, 256, 0, c->compute>>>(text, n_sym, b.crc_blk, crc_pw);
            gz::gz_crc_scan_kernel<<<1, 1024, 0, c->compute>>>(b.crc_blk, n_sym, b.crc_pre, crc_pw);
            if (n_tr) {
                gz::gz_trailer_ends_kernel<<<(n_tr + 255) / 256, 256, 0, c->compute>>>(b.trailers, n_tr, b.chunks, b.tr_end[0],
                                                                                   b.tr_idx[0]);
                size_t tmp = 0;
                CU(c, cub::DeviceRadixSort::SortPairs(nullptr, tmp, b.tr_end[0], b.tr_end[1], b.tr_idx[0], b.tr_idx[1],
                                                      static_cast<int>(n_tr), 0, 40, c->compute));
                TRY(ensure_cub_tmp(c, tmp));
                CU(c, cub::DeviceRadixSort::SortPairs(c->cub_tmp, tmp, b.tr_end[0], b.tr_end[1], b.tr_idx[0], b.tr_idx[1],
                                                      static_cast<int>(n_tr), 0, 40, c->compute));
                gz::gz_crc_ends_kernel<<<(n_tr + 127) / 128, 128, 0, c->compute>>>(text, b.tr_end[1], n_tr, b.crc_pre, b.tr_crc,
                                                                               crc_pw);
            }
            gz::gz_crc_check_kernel<<<std::max(1u, (n_tr + 127) / 128), 128, 0, c->compute>>>(
                b.trailers, b.tr_idx[1], b.tr_end[1], b.tr_crc, n_tr, b.crc_pre, n_sym, member_crc, member_len, crc_res, crc_pw);
            c->launches += 5;
            gz::gz_text_kernel<<<c->sm_count * 2, 1024, 0, c->compute>>>(text, n_sym, b.scalars + 1, flags + 1);
            c->launches += 2;
        }
        CU(c, cudaMemcpyAsync(b.scalars_host, b.scalars, 64, cudaMemcpyDeviceToHost, c->compute));
        CU(c, cudaStreamSynchronize(c->compute));
        tm.report(piece_no, n_in, n_sym, n_chunks);
        {
            const unsigned* const crc_res = reinterpret_cast<const unsigned*>(b.scalars_host + 5);
            if (crc_res[0]) return fail(c, FRB_ERR_IO, "%s: CRC check failed (gzip member trailer)", path);
            member_crc = crc_res[1];
            member_len = static_cast<uint64_t>(crc_res[2]) | (static_cast<uint64_t>(crc_res[3]) << 32);
            if (last_piece && member_len) return fail(c, FRB_ERR_IO, "%s: unexpected end of file", path);
        }
        if (hflags[1]) return gz_decline("'\\r' in the text (universal newlines are the host path's job)");
        if (want_members && n_tr) {  // which text ends where a member (and so, maybe, a file of the batch) ends
            std::vector<unsigned long long> ends(n_tr);
            std::vector<unsigned> order(n_tr);
            std::vector<gz::Trailer> tr(n_tr);
            CU(c, cudaMemcpy(ends.data(), b.tr_end[1], n_tr * 8ull, cudaMemcpyDeviceToHost));
            CU(c, cudaMemcpy(order.data(), b.tr_idx[1], n_tr * 4ull, cudaMemcpyDeviceToHost));
            CU(c, cudaMemcpy(tr.data(), b.trailers, n_tr * sizeof(gz::Trailer), cudaMemcpyDeviceToHost));
            std::vector<GzMemberEnd> members(n_tr);
            for (unsigned i = 0; i < n_tr; ++i) members[i] = GzMemberEnd{total_out + ends[i], file_pos + tr[order[i]].comp_end};
            TRY(on_members(members));
        }
        total_out += n_sym;
        const uint64_t have = carry_len + n_sym;
        uint64_t usable = have;
        if (!last_piece) {
            const uint64_t nl = b.scalars_host[1];  // end of the last complete line inside the new text
            if (nl == 0) {
                if (have > g.carry) return gz_decline("a line longer than the carry area");
                usable = 0;
            } else {
                usable = carry_len + nl;
            }
        }
        if (usable || last_piece) TRY(sink(b.out[ob], usable, last_piece));
        if (last_piece) break;
        // ---- next piece ---------------------------------------------------------------------------------
        const size_t tail = static_cast<size_t>(have - usable);
        if (tail > g.carry) return gz_decline("unfinished line longer than the carry area");
        if (tail) CU(c, cudaMemcpyAsync(b.out[ob ^ 1], b.out[ob] + usable, tail, cudaMemcpyDeviceToDevice, c->compute));
        carry_len = tail;
        ob ^= 1;
        start_bit_abs = file_pos * 8 + b.chunks_host[1].start_bit;
        member_start = false;
    }
    if (raw_bytes) *raw_bytes = total_out;
    return FRB_OK;
}

// A stream that expands more than the staging areas allow (long runs) is tried again with larger ones while the
// memory for them is there; the sink must be able to start over (`restart()` is called before every new attempt).
template <typename Sink, typename Restart, typename Members>
int gz_device_inflate_source(frb_ctx* c, GzBuffers& b, const GzSource& src, uint64_t* raw_bytes, Sink&& sink, Restart&& restart,
                             Members&& on_members, bool want_members) {
    GzConfig g = gz_config();
    for (;;) {
        const int rc = gz_device_inflate_once(c, b, g, src, raw_bytes, sink, on_members, want_members);
        if (rc != FRB_GZ_RETRY_SPACE) return rc;
        g.expand *= 8;
        size_t free_b = 0, total_b = 0;
        cudaMemGetInfo(&free_b, &total_b);
        if (g.expand > 1100 || g.piece * g.expand * 3 > free_b + b.stage_syms * 2) return gz_decline("staging areas would not fit");
        TRY(restart());
    }
}

template <typename Sink, typename Restart>
int gz_device_inflate(frb_ctx* c, GzBuffers& b, const char* path, uint64_t* raw_bytes, Sink&& sink, Restart&& restart) {
    GzSource src;
    src.paths.emplace_back(path);
    TRY(src.open(c));
    return gz_device_inflate_source(c, b, src, raw_bytes, sink, restart, [](const std::vector<GzMemberEnd>&) { return FRB_OK; }, false);
}

}  // namespace
