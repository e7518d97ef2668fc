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
        if (!api.GetUniqueId || !api.CommInitRank || !api.AllGather || !api.GetErrorString) err = "NCCL symbols missing";
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

}  // extern "C"
