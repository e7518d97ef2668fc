// Host side of hot path C (included by frb_lib.cu).
namespace {

int route_ensure(frb_ctx* c, size_t bytes, unsigned n_sinks) {
    RouteBufs& b = c->rb;
    if (bytes > b.cap_bytes || n_sinks + 1 > b.cap_sinks) {
        CU(c, cudaStreamSynchronize(c->compute));
        route_free(b);
        const size_t cap = std::max<size_t>(bytes, 1 << 20);
        const size_t recs = cap / 16 + 16;  // a record shorter than 16 bytes is reported as an error
        CU(c, cudaMalloc(&b.in1, cap + 64));
        CU(c, cudaMalloc(&b.in2, cap + 64));
        CU(c, cudaMalloc(&b.out1, cap + 64));
        CU(c, cudaMalloc(&b.out2, cap + 64));
        CU(c, cudaMalloc(&b.key2, recs * 8));
        CU(c, cudaMalloc(&b.off1, (recs + 1) * 8));
        CU(c, cudaMalloc(&b.off2, (recs + 1) * 8));
        CU(c, cudaMalloc(&b.len1, recs * 8));
        CU(c, cudaMalloc(&b.len2, recs * 8));
        CU(c, cudaMalloc(&b.pos1, recs * 8));
        CU(c, cudaMalloc(&b.pos2, recs * 8));
        CU(c, cudaMalloc(&b.sink, recs * 4));
        CU(c, cudaMalloc(&b.sink_sorted, recs * 4));
        CU(c, cudaMalloc(&b.idx, recs * 4));
        CU(c, cudaMalloc(&b.idx_sorted, recs * 4));
        CU(c, cudaMalloc(&b.sink_off1, (n_sinks + 2) * 8));
        CU(c, cudaMalloc(&b.sink_off2, (n_sinks + 2) * 8));
        b.cap_bytes = cap;
        b.cap_recs = recs;
        b.cap_sinks = n_sinks + 1;
    }
    return FRB_OK;
}

// header lines (= records, a trailing partial one included) and lines of a chunk
int route_parse(frb_ctx* c, const unsigned char* dev, uint64_t nbytes, int rule, unsigned long long* keys,
                unsigned long long* offs, uint64_t* n_records, uint64_t* n_lines) {
    CU(c, cudaMemsetAsync(c->st, 0, offsetof(DevState, occupied_total), c->compute));
    const uint64_t saved_limit = c->cur_limit;
    c->cur_limit = ~0ULL;
    int rc = launch_scan(c, dev, nbytes, 0, rule, keys, offs, nullptr, 0, c->rb.cap_recs);
    c->cur_limit = saved_limit;
    TRY(rc);
    TRY(device_error_check(c));
    *n_records = c->st_host->n_reads;
    *n_lines = c->st_host->line_carry;
    if (*n_records > c->rb.cap_recs) return fail(c, FRB_ERR_ARG, "records shorter than 16 bytes are not supported");
    CU(c, cudaMemcpyAsync(offs + *n_records, &nbytes, 8, cudaMemcpyHostToDevice, c->compute));
    CU(c, cudaStreamSynchronize(c->compute));
    return FRB_OK;
}

}  // namespace

extern "C" {

int frb_route_load(frb_ctx* c, const uint64_t* keys, const uint32_t* sink_ids, uint64_t n, uint32_t n_sinks) {
    CU(c, cudaSetDevice(c->device));
    if (n_sinks == 0 || n_sinks > (1u << 20)) return fail(c, FRB_ERR_ARG, "bad sink count");
    for (uint64_t i = 0; i < n; ++i)
        if (sink_ids[i] >= n_sinks) return fail(c, FRB_ERR_ARG, "sink id out of range");
    CU(c, cudaStreamSynchronize(c->compute));
    uint64_t cap = 1024;
    while (cap < 2 * n + 2) cap <<= 1;
    if (c->route_tab) CU(c, cudaFree(c->route_tab));
    c->route_tab = nullptr;
    CU(c, cudaMalloc(&c->route_tab, cap * sizeof(Slot)));
    c->route_cap = cap;
    c->n_sinks = n_sinks;
    CU(c, cudaMemsetAsync(c->route_tab, 0xFF, cap * sizeof(Slot), c->compute));
    if (n) {
        unsigned long long* dk = nullptr;
        unsigned* ds = nullptr;
        CU(c, cudaMalloc(&dk, n * 8));
        CU(c, cudaMalloc(&ds, n * 4));
        CU(c, cudaMemcpyAsync(dk, keys, n * 8, cudaMemcpyHostToDevice, c->compute));
        CU(c, cudaMemcpyAsync(ds, sink_ids, n * 4, cudaMemcpyHostToDevice, c->compute));
        route_build_kernel<<<static_cast<unsigned>((n + 255) / 256), 256, 0, c->compute>>>(c->route_tab, cap - 1, dk, ds,
                                                                                           n, c->st);
        c->launches++;
        CU(c, cudaGetLastError());
        CU(c, cudaStreamSynchronize(c->compute));
        CU(c, cudaFree(dk));
        CU(c, cudaFree(ds));
    }
    return device_error_check(c);
}

int frb_route_pair(frb_ctx* c, const void* r1, uint64_t r1_bytes, const void* r2, uint64_t r2_bytes, int final_chunk,
                   void* out_r1, void* out_r2, uint64_t* off_r1, uint64_t* off_r2, uint64_t* n_pairs,
                   uint64_t* consumed_r1, uint64_t* consumed_r2, uint64_t* bad_key) {
    CU(c, cudaSetDevice(c->device));
    if (!c->route_tab) return fail(c, FRB_ERR_STATE, "frb_route_pair: frb_route_load first");
    TRY(route_ensure(c, std::max(r1_bytes, r2_bytes), c->n_sinks));
    RouteBufs& b = c->rb;
    const unsigned S = c->n_sinks;
    if (r1_bytes) CU(c, cudaMemcpyAsync(b.in1, r1, r1_bytes, cudaMemcpyHostToDevice, c->compute));
    if (r2_bytes) CU(c, cudaMemcpyAsync(b.in2, r2, r2_bytes, cudaMemcpyHostToDevice, c->compute));
    uint64_t n1 = 0, n2 = 0, lines1 = 0, lines2 = 0;
    TRY(route_parse(c, b.in2, r2_bytes, FRB_RULE_DEMUX, b.key2, b.off2, &n2, &lines2));
    TRY(route_parse(c, b.in1, r1_bytes, kRuleOffsetsOnly, nullptr, b.off1, &n1, &lines1));
    // zip() of the 4-line groupers stops at the shorter mate (F:777); a trailing partial record
    // only exists at the end of a file (F:719-723)
    const uint64_t e1 = (final_chunk & 1) ? n1 : lines1 / 4, e2 = (final_chunk & 2) ? n2 : lines2 / 4;
    const uint64_t n = std::min(e1, e2);
    *n_pairs = n;
    unsigned long long c1 = 0, c2 = 0;
    if (r1_bytes) CU(c, cudaMemcpyAsync(&c1, b.off1 + n, 8, cudaMemcpyDeviceToHost, c->compute));
    if (r2_bytes) CU(c, cudaMemcpyAsync(&c2, b.off2 + n, 8, cudaMemcpyDeviceToHost, c->compute));
    CU(c, cudaStreamSynchronize(c->compute));
    if (n == 0) c1 = 0, c2 = 0;
    *consumed_r1 = c1, *consumed_r2 = c2;
    if (n >= (1ULL << 31)) return fail(c, FRB_ERR_ARG, "too many records in one chunk");
    {
        ProfScope ps(c, FRB_K_ROUTE);
        CU(c, cudaMemsetAsync(&c->st->scratch, 0xFF, 8, c->compute));
        if (n) {
            const unsigned blocks = static_cast<unsigned>((n + 255) / 256);
            route_lookup_kernel<<<blocks, 256, 0, c->compute>>>(c->route_tab, c->route_cap - 1, b.key2, n, b.sink, b.idx,
                                                               &c->st->scratch);
            int bits = 1;
            while ((1u << bits) < S) ++bits;
            size_t tmp = 0, tmp2 = 0;
            CU(c, cub::DeviceRadixSort::SortPairs(nullptr, tmp, b.sink, b.sink_sorted, b.idx, b.idx_sorted,
                                                  static_cast<int>(n), 0, bits, c->compute));
            CU(c, cub::DeviceScan::ExclusiveSum(nullptr, tmp2, b.len1, b.pos1, static_cast<int>(n), c->compute));
            TRY(ensure_cub_tmp(c, std::max(tmp, tmp2)));
            CU(c, cub::DeviceRadixSort::SortPairs(c->cub_tmp, tmp, b.sink, b.sink_sorted, b.idx, b.idx_sorted,
                                                  static_cast<int>(n), 0, bits, c->compute));
            route_len_kernel<<<blocks, 256, 0, c->compute>>>(b.idx_sorted, b.off1, b.off2, n, b.len1, b.len2);
            CU(c, cub::DeviceScan::ExclusiveSum(c->cub_tmp, tmp2, b.len1, b.pos1, static_cast<int>(n), c->compute));
            CU(c, cub::DeviceScan::ExclusiveSum(c->cub_tmp, tmp2, b.len2, b.pos2, static_cast<int>(n), c->compute));
            c->launches += 8;
        }
        route_sink_off_kernel<<<(S + 1 + 255) / 256, 256, 0, c->compute>>>(b.sink_sorted, b.pos1, b.pos2, b.len1, b.len2,
                                                                          n, S, b.sink_off1, b.sink_off2);
        c->launches++;
        if (n) {
            const int grid = grid_for(n * 32, 256, c->sm_count, 8);
            route_copy_kernel<<<grid, 256, 0, c->compute>>>(b.idx_sorted, b.off1, b.pos1, b.len1, n, b.in1, b.out1);
            route_copy_kernel<<<grid, 256, 0, c->compute>>>(b.idx_sorted, b.off2, b.pos2, b.len2, n, b.in2, b.out2);
            c->launches += 2;
        }
        CU(c, cudaGetLastError());
    }
    TRY(device_error_check(c));
    const unsigned long long first_bad = c->st_host->scratch;
    if (first_bad != ~0ULL) {  // F:807-810
        unsigned long long k = 0;
        CU(c, cudaMemcpy(&k, b.key2 + first_bad, 8, cudaMemcpyDeviceToHost));
        if (bad_key) *bad_key = k;
        char txt[24];
        frb_unpack_key(k, txt);
        return fail(c, FRB_ERR_KEY_NOT_FOUND, "Couldn't find barcode %s in supplied frender result file!", txt);
    }
    if (c1) CU(c, cudaMemcpyAsync(out_r1, b.out1, c1, cudaMemcpyDeviceToHost, c->compute));
    if (c2) CU(c, cudaMemcpyAsync(out_r2, b.out2, c2, cudaMemcpyDeviceToHost, c->compute));
    CU(c, cudaMemcpyAsync(off_r1, b.sink_off1, (S + 1) * 8, cudaMemcpyDeviceToHost, c->compute));
    CU(c, cudaMemcpyAsync(off_r2, b.sink_off2, (S + 1) * 8, cudaMemcpyDeviceToHost, c->compute));
    CU(c, cudaStreamSynchronize(c->compute));
    prof_collect(c);
    return FRB_OK;
}

}  // extern "C"
