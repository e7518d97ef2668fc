// Multi-GPU merge of per-rank totals over NCCL (included by frb_lib.cu).  NCCL is resolved with
// dlopen so the library loads on hosts without it; one process (or thread) per GPU.
#include <nccl.h>

namespace {

struct NcclApi {
    void* lib = nullptr;
    ncclResult_t (*GetUniqueId)(ncclUniqueId*) = nullptr;
    ncclResult_t (*CommInitRank)(ncclComm_t*, int, ncclUniqueId, int) = nullptr;
    ncclResult_t (*CommDestroy)(ncclComm_t) = nullptr;
    ncclResult_t (*AllGather)(const void*, void*, size_t, ncclDataType_t, ncclComm_t, cudaStream_t) = nullptr;
    ncclResult_t (*Send)(const void*, size_t, ncclDataType_t, int, ncclComm_t, cudaStream_t) = nullptr;
    ncclResult_t (*Recv)(void*, size_t, ncclDataType_t, int, ncclComm_t, cudaStream_t) = nullptr;
    ncclResult_t (*AllReduce)(const void*, void*, size_t, ncclDataType_t, ncclRedOp_t, ncclComm_t, cudaStream_t) = nullptr;
    ncclResult_t (*GroupStart)() = nullptr;
    ncclResult_t (*GroupEnd)() = nullptr;
    const char* (*GetErrorString)(ncclResult_t) = nullptr;
};

NcclApi* nccl_api(std::string* why) {
    static NcclApi api;
    static std::once_flag once;
    static std::string err;
    std::call_once(once, [] {
        const char* names[] = {"libnccl.so.2", "libnccl.so"};
        for (const char* n : names) {
            api.lib = dlopen(n, RTLD_NOW | RTLD_GLOBAL);
            if (api.lib) break;
        }
        if (!api.lib) {
            err = std::string("dlopen(libnccl.so.2): ") + dlerror();
            return;
        }
        api.GetUniqueId = reinterpret_cast<decltype(api.GetUniqueId)>(dlsym(api.lib, "ncclGetUniqueId"));
        api.CommInitRank = reinterpret_cast<decltype(api.CommInitRank)>(dlsym(api.lib, "ncclCommInitRank"));
        api.CommDestroy = reinterpret_cast<decltype(api.CommDestroy)>(dlsym(api.lib, "ncclCommDestroy"));
        api.AllGather = reinterpret_cast<decltype(api.AllGather)>(dlsym(api.lib, "ncclAllGather"));
        api.GetErrorString = reinterpret_cast<decltype(api.GetErrorString)>(dlsym(api.lib, "ncclGetErrorString"));
        api.Send = reinterpret_cast<decltype(api.Send)>(dlsym(api.lib, "ncclSend"));
        api.Recv = reinterpret_cast<decltype(api.Recv)>(dlsym(api.lib, "ncclRecv"));
        api.AllReduce = reinterpret_cast<decltype(api.AllReduce)>(dlsym(api.lib, "ncclAllReduce"));
        api.GroupStart = reinterpret_cast<decltype(api.GroupStart)>(dlsym(api.lib, "ncclGroupStart"));
        api.GroupEnd = reinterpret_cast<decltype(api.GroupEnd)>(dlsym(api.lib, "ncclGroupEnd"));
        if (!api.GetUniqueId || !api.CommInitRank || !api.AllGather || !api.GetErrorString || !api.Send || !api.Recv ||
            !api.GroupStart || !api.GroupEnd || !api.AllReduce)
            err = "NCCL symbols missing";
    });
    if (!err.empty()) {
        *why = err;
        return nullptr;
    }
    return &api;
}

#define NC(c, api, call)                                                                              \
    do {                                                                                              \
        ncclResult_t r_ = (call);                                                                     \
        if (r_ != ncclSuccess) return fail(c, FRB_ERR_NCCL, "%s failed: %s", #call, (api)->GetErrorString(r_)); \
    } while (0)

// grow-only device buffer with a stable address between calls (NCCL caches per-address state for send/recv)
int xchg_reserve(frb_ctx* c, int which, size_t bytes) {
    if (bytes <= c->xchg_cap[which]) return FRB_OK;
    if (c->xchg[which]) CU(c, cudaFree(c->xchg[which]));
    c->xchg[which] = nullptr;
    c->xchg_cap[which] = 0;
    const size_t want = bytes + bytes / 4 + 4096;
    CU(c, cudaMalloc(&c->xchg[which], want));
    c->xchg_cap[which] = want;
    return FRB_OK;
}

}  // namespace

extern "C" {

int frb_nccl_unique_id(char id128[128]) {
    std::string why;
    NcclApi* api = nccl_api(&why);
    if (!api) return fail(nullptr, FRB_ERR_NCCL, "%s", why.c_str());
    static_assert(sizeof(ncclUniqueId) == 128, "ncclUniqueId size");
    ncclUniqueId id;
    NC(nullptr, api, api->GetUniqueId(&id));
    memcpy(id128, &id, 128);
    return FRB_OK;
}

int frb_nccl_init(frb_ctx* c, const char id128[128], int rank, int n_ranks) {
    std::string why;
    NcclApi* api = nccl_api(&why);
    if (!api) return fail(c, FRB_ERR_NCCL, "%s", why.c_str());
    CU(c, cudaSetDevice(c->device));
    ncclUniqueId id;
    memcpy(&id, id128, 128);
    ncclComm_t comm;
    NC(c, api, api->CommInitRank(&comm, n_ranks, id, rank));
    c->nccl_comm = comm;
    c->rank = rank;
    c->n_ranks = n_ranks;
    return FRB_OK;
}

// Every rank contributes its sorted total list; all ranks rebuild the same merged total
// (count: +, first: min -- a commutative monoid, so the result is rank-count invariant).
int frb_allmerge(frb_ctx* c, uint64_t* n_unique) {
    CU(c, cudaSetDevice(c->device));
    TRY(frb_total_finish(c, nullptr));
    if (c->n_ranks <= 1 || !c->nccl_comm) {
        if (n_unique) *n_unique = c->total.n;
        return FRB_OK;
    }
    std::string why;
    NcclApi* api = nccl_api(&why);
    ncclComm_t comm = static_cast<ncclComm_t>(c->nccl_comm);
    const int R = c->n_ranks;
    unsigned long long* d_sizes = nullptr;
    TRY(dmalloc(c, &d_sizes, (R + 1) * 8));
    unsigned long long mine = c->total.n;
    CU(c, cudaMemcpyAsync(d_sizes + R, &mine, 8, cudaMemcpyHostToDevice, c->compute));
    NC(c, api, api->AllGather(d_sizes + R, d_sizes, 1, ncclUint64, comm, c->compute));
    std::vector<unsigned long long> sizes(R);
    CU(c, cudaMemcpyAsync(sizes.data(), d_sizes, R * 8, cudaMemcpyDeviceToHost, c->compute));
    CU(c, cudaStreamSynchronize(c->compute));
    unsigned long long mx = 1;
    for (auto s : sizes) mx = std::max(mx, s);
    // one padded buffer per field: [R][mx]
    unsigned long long *send = nullptr, *recv = nullptr;
    TRY(dmalloc(c, &send, 3 * mx * 8));
    TRY(dmalloc(c, &recv, 3ULL * R * mx * 8));
    CU(c, cudaMemsetAsync(send, 0, 3 * mx * 8, c->compute));
    if (mine) {
        CU(c, cudaMemcpyAsync(send, c->total.keys, mine * 8, cudaMemcpyDeviceToDevice, c->compute));
        CU(c, cudaMemcpyAsync(send + mx, c->total.counts, mine * 8, cudaMemcpyDeviceToDevice, c->compute));
        CU(c, cudaMemcpyAsync(send + 2 * mx, c->total.first, mine * 8, cudaMemcpyDeviceToDevice, c->compute));
    }
    NC(c, api, api->AllGather(send, recv, 3 * mx, ncclUint64, comm, c->compute));
    CU(c, cudaMemsetAsync(&c->st->occupied_total, 0, 8, c->compute));
    TRY(clear_table(c, c->total_tab));
    c->total_tab_clean = false;
    c->merged_upto = c->files.size();
    for (int r = 0; r < R; ++r) {
        if (!sizes[r]) continue;
        const unsigned long long* base = recv + 3ULL * r * mx;
        ProfScope ps(c, FRB_K_EXPORT);
        merge_list_kernel<<<static_cast<unsigned>((sizes[r] + 255) / 256), 256, 0, c->compute>>>(
            c->total_tab, c->cap - 1, base, base + mx, base + 2 * mx, sizes[r], 0ULL, &c->st->occupied_total, c->st);
        c->launches++;
    }
    CU(c, cudaGetLastError());
    TRY(dfree(c, send));
    TRY(dfree(c, recv));
    TRY(dfree(c, d_sizes));
    // the merged table replaces the local one: rebuild the sorted list from it
    CU(c, cudaStreamSynchronize(c->compute));
    TRY(device_error_check(c));
    TRY(free_list(c, c->total));
    TRY(table_to_sorted_list(c, c->total_tab, c->st_host->occupied_total, &c->total));
    c->total_ready = true;
    c->total_gen++;
    if (n_unique) *n_unique = c->total.n;
    return FRB_OK;
}

// Sharded merge: key k belongs to rank key_owner(k).  Every rank partitions its total list by owner and
// sends each part to its owner (one grouped ncclSend/ncclRecv exchange over NVLink); the owner folds what
// it receives into its cleared total table.  Afterwards the ranks hold disjoint shares whose union is what
// frb_allmerge would leave on every rank, each share in first-appearance order; matcher work and table
// size per rank stay constant as ranks are added.

int frb_shardmerge(frb_ctx* c, uint64_t* n_unique) {
    CU(c, cudaSetDevice(c->device));
    static const bool timing = getenv("FRB_MERGE_TIMING") != nullptr;
    auto t_prev = std::chrono::steady_clock::now();
    auto lap = [&](const char* what) {
        if (!timing) return;
        cudaStreamSynchronize(c->compute);
        const auto now = std::chrono::steady_clock::now();
        fprintf(stderr, "[shardmerge rank %d] %-22s %.3f ms\n", c->rank, what, std::chrono::duration<double, std::milli>(now - t_prev).count());
        t_prev = now;
    };
    if (c->in_file) return fail(c, FRB_ERR_STATE, "frb_shardmerge: a file is still open");
    TRY(total_build(c, false));  // the entries are regrouped by owner: their order does not matter here
    lap("local total");
    if (c->n_ranks <= 1 || !c->nccl_comm) {
        if (n_unique) *n_unique = c->total.n;
        return FRB_OK;
    }
    std::string why;
    NcclApi* api = nccl_api(&why);
    ncclComm_t comm = static_cast<ncclComm_t>(c->nccl_comm);
    const int R = c->n_ranks;
    if (R > 256) return fail(c, FRB_ERR_ARG, "frb_shardmerge: at most 256 ranks");
    const unsigned long long n = c->total.n;
    // 1. owner of every entry, histogram, entries grouped by owner
    TRY(xchg_reserve(c, 2, (static_cast<size_t>(R) + static_cast<size_t>(R) * R) * 8));
    unsigned long long* d_hist = c->xchg[2];
    unsigned long long* d_all = d_hist + R;
    CU(c, cudaMemsetAsync(d_hist, 0, R * 8, c->compute));
    unsigned *own = nullptr, *own_sorted = nullptr, *idx = nullptr, *idx_sorted = nullptr;
    TRY(xchg_reserve(c, 0, 3 * std::max<unsigned long long>(n, 1) * 8));
    unsigned long long* part = c->xchg[0];  // [3][n]: keys, counts, first grouped by owner
    const unsigned long long n1 = std::max<unsigned long long>(n, 1);
    TRY(dmalloc(c, &own, n1 * 4));
    TRY(dmalloc(c, &own_sorted, n1 * 4));
    TRY(dmalloc(c, &idx, n1 * 4));
    TRY(dmalloc(c, &idx_sorted, n1 * 4));
    if (n) {
        ProfScope ps(c, FRB_K_EXPORT);
        const unsigned grid = static_cast<unsigned>((n + 255) / 256);
        owner_kernel<<<grid, 256, 0, c->compute>>>(c->total.keys, n, static_cast<unsigned>(R), own, d_hist);
        iota_kernel<<<grid, 256, 0, c->compute>>>(idx, n);
        size_t tmp = 0;
        CU(c, cub::DeviceRadixSort::SortPairs(nullptr, tmp, own, own_sorted, idx, idx_sorted, static_cast<int>(n), 0, 8,
                                              c->compute));
        TRY(ensure_cub_tmp(c, tmp));
        CU(c, cub::DeviceRadixSort::SortPairs(c->cub_tmp, tmp, own, own_sorted, idx, idx_sorted, static_cast<int>(n), 0, 8,
                                              c->compute));
        gather3_kernel<<<grid, 256, 0, c->compute>>>(idx_sorted, c->total.keys, c->total.counts, c->total.first, part,
                                                     part + n, part + 2 * n, n);
        c->launches += 5;
        CU(c, cudaGetLastError());
    }
    lap("partition by owner");
    // 2. everybody learns everybody's histogram
    NC(c, api, api->AllGather(d_hist, d_all, R, ncclUint64, comm, c->compute));
    std::vector<unsigned long long> all(static_cast<size_t>(R) * R);
    CU(c, cudaMemcpyAsync(all.data(), d_all, all.size() * 8, cudaMemcpyDeviceToHost, c->compute));
    CU(c, cudaStreamSynchronize(c->compute));
    std::vector<unsigned long long> send_off(R + 1, 0), recv_off(R + 1, 0);
    for (int r = 0; r < R; ++r) {
        send_off[r + 1] = send_off[r] + all[static_cast<size_t>(c->rank) * R + r];  // my entries owned by r
        recv_off[r + 1] = recv_off[r] + all[static_cast<size_t>(r) * R + c->rank];  // r's entries owned by me
    }
    if (send_off[R] != n) return fail(c, FRB_ERR_STATE, "frb_shardmerge: owner histogram does not add up");
    const unsigned long long m = recv_off[R];
    const unsigned long long m1 = std::max<unsigned long long>(m, 1);
    TRY(xchg_reserve(c, 1, 3 * m1 * 8));
    unsigned long long* got = c->xchg[1];  // [3][m]
    lap("histogram exchange");
    // 3. the exchange
    NC(c, api, api->GroupStart());
    for (int r = 0; r < R; ++r) {
        const unsigned long long ns = send_off[r + 1] - send_off[r], nr = recv_off[r + 1] - recv_off[r];
        for (int f = 0; f < 3; ++f) {
            if (ns) NC(c, api, api->Send(part + f * n + send_off[r], ns, ncclUint64, r, comm, c->compute));
            if (nr) NC(c, api, api->Recv(got + f * m + recv_off[r], nr, ncclUint64, r, comm, c->compute));
        }
    }
    NC(c, api, api->GroupEnd());
    lap("send/recv");
    // 4. fold my share: sort what arrived by key, one entry per run of equal keys (count: +, first: min),
    //    then the usual sort by first appearance.  No table involved: sequential traffic only.
    unsigned long long *k0 = nullptr, *c0 = nullptr, *f0 = nullptr;
    unsigned *pi = nullptr, *po = nullptr, *h_in = nullptr, *h_sorted = nullptr;
    TRY(dmalloc(c, &k0, m1 * 8));
    TRY(dmalloc(c, &c0, m1 * 8));
    TRY(dmalloc(c, &f0, m1 * 8));
    TRY(dmalloc(c, &h_in, m1 * 4));
    TRY(dmalloc(c, &h_sorted, m1 * 4));
    TRY(dmalloc(c, &pi, m1 * 4));
    TRY(dmalloc(c, &po, m1 * 4));
    CU(c, cudaMemsetAsync(&c->st->scratch, 0, 8, c->compute));
    CU(c, cudaMemsetAsync(&c->st->occupied_total, 0, 8, c->compute));  // borrowed: largest `first` received
    if (m) {
        if (m >= (1ULL << 31)) return fail(c, FRB_ERR_ARG, "frb_shardmerge: share too large");
        ProfScope ps(c, FRB_K_EXPORT);
        const unsigned grid = static_cast<unsigned>((m + 255) / 256);
        iota_kernel<<<grid, 256, 0, c->compute>>>(pi, m);
        hash32_kernel<<<grid, 256, 0, c->compute>>>(got, got + 2 * m, m, h_in, &c->st->occupied_total);
        size_t tmp = 0;
        CU(c, cub::DeviceRadixSort::SortPairs(nullptr, tmp, h_in, h_sorted, pi, po, static_cast<int>(m), 0, 32, c->compute));
        TRY(ensure_cub_tmp(c, tmp));
        CU(c, cub::DeviceRadixSort::SortPairs(c->cub_tmp, tmp, h_in, h_sorted, pi, po, static_cast<int>(m), 0, 32, c->compute));
        fold_hash_runs_kernel<<<grid, 256, 0, c->compute>>>(h_sorted, po, got, got + m, got + 2 * m, m, k0, c0, f0,
                                                            &c->st->scratch);
        c->launches += 7;
        CU(c, cudaGetLastError());
    }
    unsigned long long u = 0, max_first = 0;
    CU(c, cudaMemcpyAsync(&u, &c->st->scratch, 8, cudaMemcpyDeviceToHost, c->compute));
    CU(c, cudaMemcpyAsync(&max_first, &c->st->occupied_total, 8, cudaMemcpyDeviceToHost, c->compute));
    CU(c, cudaStreamSynchronize(c->compute));
    lap("sort by key + fold");
    TRY(dfree(c, own));
    TRY(dfree(c, own_sorted));
    TRY(dfree(c, idx));
    TRY(dfree(c, idx_sorted));
    TRY(dfree(c, h_in));
    TRY(dfree(c, h_sorted));
    TRY(dfree(c, pi));
    TRY(dfree(c, po));
    TRY(device_error_check(c));
    TRY(free_list(c, c->total));
    int first_bits = 1;  // significant bits of (file ordinal << 40 | read ordinal) among what arrived
    while (first_bits < 64 && (max_first >> first_bits)) ++first_bits;
    TRY(unsorted_to_sorted_list(c, k0, c0, f0, u, &c->total, first_bits));
    c->merged_upto = c->files.size();
    c->sharded = true;
    lap("sorted share");
    c->total_ready = true;
    c->total_gen++;
    if (n_unique) *n_unique = c->total.n;
    return FRB_OK;
}

// Element-wise sum of a small u64 host array over all ranks (the per-sample orientation sums of the first
// matcher pass when the total is sharded: the f < rc call of F:354-388 is over ALL reads of the job).
int frb_allreduce_u64(frb_ctx* c, uint64_t* host_inout, uint64_t n) {
    CU(c, cudaSetDevice(c->device));
    if (c->n_ranks <= 1 || !c->nccl_comm || n == 0) return FRB_OK;
    std::string why;
    NcclApi* api = nccl_api(&why);
    ncclComm_t comm = static_cast<ncclComm_t>(c->nccl_comm);
    unsigned long long* d = nullptr;
    TRY(dmalloc(c, &d, n * 8));
    CU(c, cudaMemcpyAsync(d, host_inout, n * 8, cudaMemcpyHostToDevice, c->compute));
    NC(c, api, api->AllReduce(d, d, n, ncclUint64, ncclSum, comm, c->compute));
    CU(c, cudaMemcpyAsync(host_inout, d, n * 8, cudaMemcpyDeviceToHost, c->compute));
    CU(c, cudaStreamSynchronize(c->compute));
    TRY(dfree(c, d));
    return FRB_OK;
}

}  // extern "C"
