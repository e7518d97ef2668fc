// Microbenchmark: what read bandwidth does a persistent bulk-copy pipeline reach on B200 as a function of the
// bytes it keeps in flight per SM?  One thread per CTA drives `stages` shared-memory buffers of `tile` bytes:
// wait for a buffer's copy, hold it for `hold` cycles (stands in for the processing time during which the buffer
// cannot be refilled), re-issue.  Also reports the mean issue->arrival latency of a copy.
//   nvcc -O3 -gencode arch=compute_100a,code=sm_100a -o tma_probe tma_probe.cu
#include <cstdio>
#include <cstdlib>
#include <cuda_runtime.h>

__device__ __forceinline__ unsigned smem_addr(const void* p) { return static_cast<unsigned>(__cvta_generic_to_shared(p)); }

__global__ void probe(const unsigned char* data, unsigned long long n_tiles, unsigned tile, int stages, int hold,
                      unsigned long long* ticket, unsigned long long* lat_sum, unsigned long long* lat_n) {
    extern __shared__ __align__(128) unsigned char smem[];
    __shared__ __align__(8) unsigned long long bar[8];
    if (threadIdx.x != 0) return;
    for (int s = 0; s < stages; ++s)
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_addr(&bar[s])) : "memory");
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    long long t_issue[8];
    bool live[8];
    unsigned parity = 0;
    unsigned long long sum = 0, cnt = 0;
    auto issue = [&](int s) {
        const unsigned long long t = atomicAdd(ticket, 1ULL);
        live[s] = t < n_tiles;
        if (!live[s]) return;
        asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_addr(&bar[s])), "r"(tile) : "memory");
        asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                         smem_addr(smem + static_cast<size_t>(s) * tile)),
                     "l"(data + t * tile), "r"(tile), "r"(smem_addr(&bar[s]))
                     : "memory");
        t_issue[s] = clock64();
    };
    for (int s = 0; s < stages; ++s) issue(s);
    for (int i = 0;; ++i) {
        const int s = i % stages;
        if (!live[s]) break;
        asm volatile(
            "{\n\t.reg .pred p;\n\tW_%=:\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t@p bra D_%=;\n\tbra W_%=;\n\tD_%=:\n\t}"
            ::"r"(smem_addr(&bar[s])), "r"((parity >> s) & 1u) : "memory");
        parity ^= 1u << s;
        const long long now = clock64();
        sum += now - t_issue[s];
        ++cnt;
        while (clock64() - now < hold) {}
        issue(s);
    }
    atomicAdd(lat_sum, sum);
    atomicAdd(lat_n, cnt);
}

int main(int argc, char** argv) {
    const size_t bytes = 8ULL << 30;
    unsigned char* d;
    cudaMalloc(&d, bytes);
    cudaMemset(d, 1, bytes);
    unsigned long long* ctr;
    cudaMalloc(&ctr, 24);
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0), cudaEventCreate(&e1);
    struct Cfg { int ctas, stages; unsigned tile; int hold; };
    const Cfg cfgs[] = {
        {3, 2, 30720, 0},    {3, 2, 30720, 1500},  {3, 2, 30720, 3000},  {3, 2, 30720, 4500},
        {3, 3, 20480, 0},    {3, 3, 20480, 2000},  {3, 3, 20480, 3000},
        {3, 4, 15360, 0},    {3, 4, 15360, 1500},  {3, 4, 15360, 2250},
        {3, 6, 10240, 0},    {3, 6, 10240, 1000},  {3, 6, 10240, 1500},
        {1, 6, 30720, 0},    {1, 6, 30720, 1000},  {2, 3, 30720, 1500},  {2, 3, 30720, 3000},
        {6, 2, 15360, 0},    {6, 2, 15360, 1500},  {6, 2, 15360, 2250},
        {1, 2, 98304, 0},    {1, 2, 98304, 3000},  {4, 2, 24576, 2400},  {4, 3, 16384, 1600},
    };
    cudaFuncSetAttribute(probe, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
    for (const Cfg& c : cfgs) {
        const unsigned long long n_tiles = bytes / c.tile;
        float best = 1e30f;
        unsigned long long h[3];
        for (int rep = 0; rep < 3; ++rep) {
            cudaMemset(ctr, 0, 24);
            cudaEventRecord(e0);
            probe<<<148 * c.ctas, 32, static_cast<size_t>(c.stages) * c.tile>>>(d, n_tiles, c.tile, c.stages, c.hold, ctr,
                                                                                  ctr + 1, ctr + 2);
            cudaEventRecord(e1);
            cudaEventSynchronize(e1);
            float ms;
            cudaEventElapsedTime(&ms, e0, e1);
            if (ms < best) best = ms;
            cudaMemcpy(h, ctr, 24, cudaMemcpyDeviceToHost);
        }
        cudaError_t err = cudaGetLastError();
        printf("ctas/SM %d stages %d tile %6u hold %5d : %7.1f GB/s  staged %4zu KB/SM  mean copy latency %6.0f cycles %s\n", c.ctas,
               c.stages, c.tile, c.hold, n_tiles * (double)c.tile / best / 1e6, (size_t)c.ctas * c.stages * c.tile / 1024,
               (double)h[1] / (double)h[2], err == cudaSuccess ? "" : cudaGetErrorString(err));
    }
    return 0;
}
