// Host side of hot path C (included by frb_lib.cu): the demux stream.
//
// frb_route_push queues one chunk pair -- H2D on the copy stream, the kernels of route_kernel.cuh on the compute
// stream, nothing waits for the host -- and frb_route_pop hands out the routed chunk: it learns the sizes (a few
// hundred bytes of state), starts the D2H of exactly the bytes written on a third stream and waits for it.  With
// the next chunk pushed before the previous one is popped, the H2D of chunk k+1, the kernels of chunk k+1 and the
// D2H of chunk k run at the same time (PCIe is full duplex).  Two slots of buffers alternate.
namespace {

struct RouteSlot {
    unsigned char *in1 = nullptr, *in2 = nullptr;      // [carry area | new bytes]
    unsigned char *out1 = nullptr, *out2 = nullptr;    // routed records, sink after sink
    unsigned char *hout1 = nullptr, *hout2 = nullptr;  // pinned twins of out1 / out2
    unsigned char *hin1 = nullptr, *hin2 = nullptr;    // pinned staging for pageable input
    unsigned long long* sink_off = nullptr;            // device: [2][n_sinks + 1]
    unsigned char* hstate = nullptr;                   // pinned: RouteState + sink offsets
    cudaEvent_t copied = nullptr, kernels = nullptr, state_done = nullptr, d2h_done = nullptr;
    uint64_t end1 = 0, end2 = 0;                       // bytes in the chunk buffers (carry area included)
    bool busy = false;
};

struct RouteStream {
    RouteSlot slot[2];
    RouteState* rs = nullptr;
    unsigned long long *key2 = nullptr, *off1 = nullptr, *off2 = nullptr, *cell = nullptr;
    unsigned *sink = nullptr, *local1 = nullptr, *local2 = nullptr;
    size_t cap = 0, rec_cap = 0;
    unsigned n_sinks = 0, n_blocks = 0;
    cudaStream_t d2h = nullptr;
    uint64_t pushed = 0, popped = 0;
};

void route_stream_free(RouteStream& r) {
    for (auto& s : r.slot) {
        cudaFree(s.in1), cudaFree(s.in2), cudaFree(s.out1), cudaFree(s.out2), cudaFree(s.sink_off);
        cudaFreeHost(s.hout1), cudaFreeHost(s.hout2), cudaFreeHost(s.hin1), cudaFreeHost(s.hin2), cudaFreeHost(s.hstate);
        if (s.copied) cudaEventDestroy(s.copied), cudaEventDestroy(s.kernels), cudaEventDestroy(s.state_done), cudaEventDestroy(s.d2h_done);
    }
    cudaFree(r.rs), cudaFree(r.key2), cudaFree(r.off1), cudaFree(r.off2), cudaFree(r.cell), cudaFree(r.sink), cudaFree(r.local1),
        cudaFree(r.local2);
    if (r.d2h) cudaStreamDestroy(r.d2h);
    r = RouteStream{};
}

int route_stream_alloc(frb_ctx* c, RouteStream& r, size_t cap, unsigned n_sinks) {
    r.rec_cap = 2 * cap / 16 + 16;  // a record shorter than 16 bytes is reported as an error
    r.n_blocks = static_cast<unsigned>((r.rec_cap + kRouteBlock - 1) / kRouteBlock);
    CU(c, cudaStreamCreateWithFlags(&r.d2h, cudaStreamNonBlocking));
    for (auto& s : r.slot) {
        CU(c, cudaMalloc(&s.in1, 2 * cap + 256));
        CU(c, cudaMalloc(&s.in2, 2 * cap + 256));
        CU(c, cudaMalloc(&s.out1, 2 * cap + 256));
        CU(c, cudaMalloc(&s.out2, 2 * cap + 256));
        CU(c, cudaMallocHost(&s.hout1, 2 * cap + 256));
        CU(c, cudaMallocHost(&s.hout2, 2 * cap + 256));
        CU(c, cudaMallocHost(&s.hin1, cap));
        CU(c, cudaMallocHost(&s.hin2, cap));
        CU(c, cudaMalloc(&s.sink_off, 2 * (n_sinks + 1) * 8));
        CU(c, cudaMallocHost(&s.hstate, sizeof(RouteState) + 2 * (n_sinks + 1) * 8));
        CU(c, cudaEventCreateWithFlags(&s.copied, cudaEventDisableTiming));
        CU(c, cudaEventCreateWithFlags(&s.kernels, cudaEventDisableTiming));
        CU(c, cudaEventCreateWithFlags(&s.state_done, cudaEventDisableTiming));
        CU(c, cudaEventCreateWithFlags(&s.d2h_done, cudaEventDisableTiming));
    }
    CU(c, cudaMalloc(&r.rs, sizeof(RouteState)));
    CU(c, cudaMalloc(&r.key2, r.rec_cap * 8));
    CU(c, cudaMalloc(&r.off1, (r.rec_cap + 1) * 8));
    CU(c, cudaMalloc(&r.off2, (r.rec_cap + 1) * 8));
    CU(c, cudaMalloc(&r.sink, r.rec_cap * 4));
    CU(c, cudaMalloc(&r.local1, r.rec_cap * 4));
    CU(c, cudaMalloc(&r.local2, r.rec_cap * 4));
    CU(c, cudaMalloc(&r.cell, (2ull * n_sinks * r.n_blocks + 2ull * n_sinks) * 8));  // + per-sink totals
    RouteState init{};
    init.skip1 = init.skip2 = cap;  // nothing carried: the text begins where the host puts the new bytes
    CU(c, cudaMemcpy(r.rs, &init, sizeof init, cudaMemcpyHostToDevice));
    // the histogram keeps two counters per sink in shared memory
    CU(c, cudaFuncSetAttribute(route_hist_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                               static_cast<int>(std::max<size_t>(2 * n_sinks * sizeof(unsigned), 48 << 10))));
    return FRB_OK;
}

int route_stream_ensure(frb_ctx* c, RouteStream& r, size_t bytes, unsigned n_sinks) {
    if (bytes <= r.cap && n_sinks == r.n_sinks) return FRB_OK;
    if (r.pushed != r.popped) return fail(c, FRB_ERR_STATE, "demux stream: chunks in flight while its buffers would grow");
    CU(c, cudaStreamSynchronize(c->compute));
    const size_t cap = std::max<size_t>(std::max(bytes, r.cap), 1 << 20);
    route_stream_free(r);
    const int rc = route_stream_alloc(c, r, cap, n_sinks);
    if (rc != FRB_OK) {  // out of memory half way: nothing of the stream stays behind
        route_stream_free(r);
        return rc;
    }
    r.cap = cap;
    r.n_sinks = n_sinks;
    return FRB_OK;
}

}  // namespace

struct RouteStreamHolder {
    RouteStream r;
};

static void route_stream_destroy(frb_ctx* c) {
    if (!c->route) return;
    route_stream_free(c->route->r);
    delete c->route;
    c->route = nullptr;
}

extern "C" {

int frb_route_load(frb_ctx* c, const uint64_t* keys, const uint32_t* sink_ids, uint64_t n, uint32_t n_sinks) {
    CU(c, cudaSetDevice(c->device));
    if (n_sinks == 0 || n_sinks > (1u << 14)) return fail(c, FRB_ERR_ARG, "bad sink count");
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

// A new stream of chunk pairs begins (nothing carried over from an earlier one).
int frb_route_reset(frb_ctx* c) {
    CU(c, cudaSetDevice(c->device));
    if (!c->route) return FRB_OK;
    RouteStream& r = c->route->r;
    CU(c, cudaStreamSynchronize(c->compute));
    if (r.d2h) CU(c, cudaStreamSynchronize(r.d2h));
    r.pushed = r.popped = 0;
    for (auto& s : r.slot) s.busy = false;
    if (r.rs) {
        RouteState init{};
        init.skip1 = init.skip2 = r.cap;
        CU(c, cudaMemcpy(r.rs, &init, sizeof init, cudaMemcpyHostToDevice));
    }
    return FRB_OK;
}

// Buffers for chunks of up to `chunk_bytes` per mate (otherwise the first chunk pushed sets the size, and later
// chunks of the stream may not be larger).  Between streams only.
int frb_route_reserve(frb_ctx* c, uint64_t chunk_bytes) {
    CU(c, cudaSetDevice(c->device));
    if (!c->route_tab) return fail(c, FRB_ERR_STATE, "frb_route_reserve: frb_route_load first");
    if (!c->route) c->route = new RouteStreamHolder();
    RouteStream& r = c->route->r;
    if (r.pushed != r.popped) return fail(c, FRB_ERR_STATE, "frb_route_reserve: chunks in flight");
    TRY(route_stream_ensure(c, r, chunk_bytes, c->n_sinks));
    return frb_route_reset(c);
}

int frb_route_push(frb_ctx* c, const void* r1, uint64_t n1, const void* r2, uint64_t n2, int final_chunk) {
    CU(c, cudaSetDevice(c->device));
    if (!c->route_tab) return fail(c, FRB_ERR_STATE, "frb_route_push: frb_route_load first");
    if (!c->route) c->route = new RouteStreamHolder();
    RouteStream& r = c->route->r;
    if (r.pushed - r.popped >= 2) return fail(c, FRB_ERR_STATE, "frb_route_push: two chunks in flight, pop one first");
    if (std::max(n1, n2) > r.cap || c->n_sinks != r.n_sinks) {
        if (r.pushed != r.popped) return fail(c, FRB_ERR_ARG, "frb_route_push: chunk larger than the stream's buffers (%zu)", r.cap);
        // a larger chunk size starts a new set of buffers; what was carried must move along
        if (r.pushed) return fail(c, FRB_ERR_ARG, "frb_route_push: chunks of a stream may not grow beyond its first chunk's size");
        TRY(route_stream_ensure(c, r, std::max(n1, n2), c->n_sinks));
    }
    RouteSlot& s = r.slot[r.pushed & 1];
    const unsigned S = r.n_sinks;
    // ---- H2D of the new bytes behind the carry area (copy stream) --------------------------------------------
    auto stage = [&](const void* src, uint64_t n, unsigned char* pinned, unsigned char* dst) -> int {
        if (!n) return FRB_OK;
        cudaPointerAttributes at{};
        const bool is_pinned = cudaPointerGetAttributes(&at, src) == cudaSuccess && at.type == cudaMemoryTypeHost;
        cudaGetLastError();
        const void* from = src;
        if (!is_pinned) {  // pageable memory would make the copy synchronous: go through the slot's pinned buffer
            memcpy(pinned, src, n);
            from = pinned;
        }
        CU(c, cudaMemcpyAsync(dst, from, n, cudaMemcpyHostToDevice, c->copy));
        return FRB_OK;
    };
    // the slot's input buffers were last read by the kernels of chunk k-2 (and its carry went into the other slot)
    CU(c, cudaStreamWaitEvent(c->copy, s.kernels, 0));
    TRY(stage(r1, n1, s.hin1, s.in1 + r.cap));
    TRY(stage(r2, n2, s.hin2, s.in2 + r.cap));
    CU(c, cudaEventRecord(s.copied, c->copy));
    s.end1 = r.cap + n1, s.end2 = r.cap + n2;
    // ---- kernels (compute stream) --------------------------------------------------------------------------------
    CU(c, cudaStreamWaitEvent(c->compute, s.copied, 0));
    CU(c, cudaStreamWaitEvent(c->compute, s.d2h_done, 0));  // out buffers of this slot: D2H of chunk k-2 done
    RouteSlot& nxt = r.slot[(r.pushed + 1) & 1];
    {
        const uint64_t saved_limit = c->cur_limit;
        c->cur_limit = ~0ULL;
        CU(c, cudaMemsetAsync(c->st, 0, offsetof(DevState, occupied_total), c->compute));
        int rc = launch_scan(c, s.in2, s.end2, 0, FRB_RULE_DEMUX, r.key2, r.off2, nullptr, 0, r.rec_cap, &r.rs->skip2);
        if (rc == FRB_OK) {
            route_grab_kernel<<<1, 32, 0, c->compute>>>(c->st, r.rs, 2);
            rc = launch_scan(c, s.in1, s.end1, 0, kRuleOffsetsOnly, nullptr, r.off1, nullptr, 0, r.rec_cap, &r.rs->skip1);
        }
        c->cur_limit = saved_limit;
        TRY(rc);
        route_grab_kernel<<<1, 32, 0, c->compute>>>(c->st, r.rs, 1);
    }
    {
        ProfScope ps(c, FRB_K_ROUTE);
        unsigned long long* const tot = r.cell + 2ull * S * r.n_blocks;
        route_plan_kernel<<<1, 32, 0, c->compute>>>(r.rs, c->st, r.off1, r.off2, s.in1, s.in2, s.end1, s.end2, final_chunk,
                                               r.rec_cap, r.cap, tot, 2 * S);
        route_hist_kernel<<<r.n_blocks, kRouteBlock, 2 * S * sizeof(unsigned), c->compute>>>(
            c->route_tab, c->route_cap - 1, r.key2, r.off1, r.off2, r.rs, S, r.n_blocks, r.sink, r.local1, r.local2, r.cell, tot);
        route_scan_kernel<<<dim3(S, 2), kRouteBlock, 0, c->compute>>>(r.cell, S, r.n_blocks, r.rs, tot, s.sink_off);
        const int grid = c->sm_count * 8;
        route_copy_kernel<<<grid, 256, 0, c->compute>>>(r.rs, r.sink, r.local1, r.cell, r.n_blocks, r.off1, s.in1, s.out1);
        route_copy_kernel<<<grid, 256, 0, c->compute>>>(r.rs, r.sink, r.local2,
                                                        r.cell + static_cast<unsigned long long>(S) * r.n_blocks, r.n_blocks,
                                                        r.off2, s.in2, s.out2);
        route_carry_kernel<<<dim3(64, 2), 1024, 0, c->compute>>>(r.rs, s.in1, s.in2, s.end1, s.end2, nxt.in1, nxt.in2, r.cap);
        c->launches += 8;
        CU(c, cudaGetLastError());
    }
    // ---- the chunk's state to the host, in stream order behind its kernels (a few hundred bytes; on the D2H stream
    // it would queue the routed bytes of the chunk before, which frb_route_pop sends there, behind these kernels)
    CU(c, cudaMemcpyAsync(s.hstate, r.rs, sizeof(RouteState), cudaMemcpyDeviceToHost, c->compute));
    CU(c, cudaMemcpyAsync(s.hstate + sizeof(RouteState), s.sink_off, 2 * (S + 1) * 8, cudaMemcpyDeviceToHost, c->compute));
    CU(c, cudaEventRecord(s.kernels, c->compute));
    CU(c, cudaEventRecord(s.state_done, c->compute));
    s.busy = true;
    r.pushed++;
    return FRB_OK;
}

// The oldest chunk in flight.  *out_r1 / *out_r2 point into pinned memory of the library, valid until the chunk
// after next is pushed; sink s of mate m is out_m[off_m[s], off_m[s + 1]).  carry_r1 / carry_r2: bytes of the
// pushed input not routed yet (they are part of the next chunk).
int frb_route_pop(frb_ctx* c, void** out_r1, void** out_r2, uint64_t* off_r1, uint64_t* off_r2, uint64_t* n_pairs,
                  uint64_t* carry_r1, uint64_t* carry_r2, uint64_t* bad_key) {
    CU(c, cudaSetDevice(c->device));
    if (!c->route || c->route->r.pushed == c->route->r.popped) return fail(c, FRB_ERR_STATE, "frb_route_pop: nothing pushed");
    RouteStream& r = c->route->r;
    RouteSlot& s = r.slot[r.popped & 1];
    const unsigned S = r.n_sinks;
    CU(c, cudaEventSynchronize(s.state_done));
    RouteState st;
    memcpy(&st, s.hstate, sizeof st);
    r.popped++;
    s.busy = false;
    if (st.error) {
        CU(c, cudaEventRecord(s.d2h_done, r.d2h));
        if (st.error == FRB_ERR_BAD_ALPHABET)
            return fail(c, st.error, "record %llu of the chunk: index field holds a symbol outside ACGTN+", st.bad_key);
        if (st.error == FRB_ERR_KEY_TOO_LONG)
            return fail(c, st.error, "record %llu of the chunk: index field longer than 21 symbols", st.bad_key);
        return fail(c, st.error, "demux stream: a record (or the lead of one mate over the other) exceeds the chunk size, "
                                 "or records are shorter than 16 bytes");
    }
    if (st.bad != ~0ull) {  // F:807-810
        unsigned long long k = 0;
        CU(c, cudaMemcpy(&k, r.key2 + st.bad, 8, cudaMemcpyDeviceToHost));
        CU(c, cudaEventRecord(s.d2h_done, r.d2h));
        if (bad_key) *bad_key = k;
        if (k > kBadKeyBase && k != kEmpty) {  // the header of a routed record did not parse
            const int code = -static_cast<int>(k - kBadKeyBase);
            if (code == FRB_ERR_BAD_ALPHABET)
                return fail(c, code, "record %llu of the chunk: index field holds a symbol outside ACGTN+", st.bad);
            if (code == FRB_ERR_KEY_TOO_LONG)
                return fail(c, code, "record %llu of the chunk: index field longer than 21 symbols", st.bad);
            return fail(c, code, "record %llu of the chunk: no index field in the header line", st.bad);
        }
        char txt[24];
        frb_unpack_key(k, txt);
        return fail(c, FRB_ERR_KEY_NOT_FOUND, "Couldn't find barcode %s in supplied frender result file!", txt);
    }
    // exactly the bytes written, on the third stream: they cross while the next chunk's H2D and kernels run
    if (st.out1) CU(c, cudaMemcpyAsync(s.hout1, s.out1, st.out1, cudaMemcpyDeviceToHost, r.d2h));
    if (st.out2) CU(c, cudaMemcpyAsync(s.hout2, s.out2, st.out2, cudaMemcpyDeviceToHost, r.d2h));
    CU(c, cudaEventRecord(s.d2h_done, r.d2h));
    CU(c, cudaEventSynchronize(s.d2h_done));
    const unsigned long long* so = reinterpret_cast<const unsigned long long*>(s.hstate + sizeof(RouteState));
    memcpy(off_r1, so, (S + 1) * 8);
    memcpy(off_r2, so + (S + 1), (S + 1) * 8);
    *out_r1 = s.hout1, *out_r2 = s.hout2;
    *n_pairs = st.pairs;
    if (carry_r1) *carry_r1 = s.end1 - st.used1;
    if (carry_r2) *carry_r2 = s.end2 - st.used2;
    return FRB_OK;
}

// One chunk pair, synchronously (the round-1 entry point, now a stream of one): both inputs begin at a record
// start; the caller carries what was not consumed.
int frb_route_pair(frb_ctx* c, const void* r1, uint64_t r1_bytes, const void* r2, uint64_t r2_bytes, int final_chunk,
                   void* out_r1, void* out_r2, uint64_t* off_r1, uint64_t* off_r2, uint64_t* n_pairs,
                   uint64_t* consumed_r1, uint64_t* consumed_r2, uint64_t* bad_key) {
    TRY(frb_route_reset(c));
    TRY(frb_route_push(c, r1, r1_bytes, r2, r2_bytes, final_chunk));
    void *o1 = nullptr, *o2 = nullptr;
    uint64_t carry1 = 0, carry2 = 0;
    const int rc = frb_route_pop(c, &o1, &o2, off_r1, off_r2, n_pairs, &carry1, &carry2, bad_key);
    if (rc != FRB_OK) {
        frb_route_reset(c);
        return rc;
    }
    const unsigned S = c->n_sinks;
    if (off_r1[S]) memcpy(out_r1, o1, off_r1[S]);
    if (off_r2[S]) memcpy(out_r2, o2, off_r2[S]);
    *consumed_r1 = r1_bytes - carry1;
    *consumed_r2 = r2_bytes - carry2;
    return frb_route_reset(c);  // the caller carries the rest itself
}

}  // extern "C"
