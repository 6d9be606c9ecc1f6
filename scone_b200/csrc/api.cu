// api.cu -- library-level entry points: version, thread-local error string, launch counter.
#include <sys/mman.h>

#include <cstring>
#include <thread>
#include <vector>

#include "common.cuh"

namespace scone {

static thread_local char t_error[512] = "";
std::atomic<int64_t> g_launches{0};

void set_error(const char *fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(t_error, sizeof t_error, fmt, ap);
    va_end(ap);
}

}  // namespace scone

extern "C" {

int scone_version(void) { return SCONE_B200_VERSION; }

const char *scone_last_error(void) { return scone::t_error; }

int64_t scone_launch_count(void) { return scone::g_launches.load(); }

// Offloaded tier: table storage in host memory that the GPU reads directly.  Plain cudaHostAlloc faults and pins 4 KB pages
// one by one (measured: 1.4 GB/s, 30 s for a 41 GB table); here the region is an anonymous mapping advised to use transparent
// huge pages, first-touched by `nthreads` host threads in parallel, then registered with the driver (mapped + portable):
// 512 times fewer pages to pin, and far fewer TLB misses for the host-side row gather of the staged variant.
int scone_host_alloc(int64_t bytes, int32_t nthreads, void **out_host, void **out_device) {
    SCONE_REQUIRE(out_host && out_device, "scone_host_alloc: NULL output");
    *out_host = *out_device = nullptr;
    SCONE_REQUIRE(bytes > 0, "scone_host_alloc: bytes must be positive");
    const size_t huge = size_t(2) << 20;
    const size_t len = ((size_t)bytes + huge - 1) / huge * huge;
    void *p = mmap(nullptr, len, PROT_READ | PROT_WRITE, MAP_PRIVATE | MAP_ANONYMOUS, -1, 0);
    if (p == MAP_FAILED) {
        scone::set_error("scone_host_alloc: mmap of %llu bytes failed", (unsigned long long)len);
        return SCONE_E_NOMEM;
    }
    madvise(p, len, MADV_HUGEPAGE);  // advisory: without THP the region simply stays on 4 KB pages
    if (nthreads < 1) nthreads = 1;
    if (nthreads > 256) nthreads = 256;
    {
        std::vector<std::thread> pool;
        const size_t per = ((len / huge) + nthreads - 1) / nthreads * huge;
        for (int t = 0; t < nthreads; ++t) {
            const size_t lo = (size_t)t * per, hi = lo + per < len ? lo + per : len;
            if (lo < hi) pool.emplace_back([=] { std::memset(static_cast<char *>(p) + lo, 0, hi - lo); });  // first touch
        }
        for (auto &th : pool) th.join();
    }
    cudaError_t e = cudaHostRegister(p, len, cudaHostRegisterMapped | cudaHostRegisterPortable);
    void *d = nullptr;
    if (e == cudaSuccess) e = cudaHostGetDevicePointer(&d, p, 0);
    if (e != cudaSuccess) {
        scone::set_error("scone_host_alloc: registering %llu bytes failed: %s", (unsigned long long)len, cudaGetErrorString(e));
        cudaGetLastError();
        munmap(p, len);
        return e == cudaErrorMemoryAllocation ? SCONE_E_NOMEM : SCONE_E_CUDA;
    }
    *out_host = p;
    *out_device = d;
    return SCONE_OK;
}

int scone_host_free(void *host, int64_t bytes) {
    if (!host) return SCONE_OK;
    const size_t huge = size_t(2) << 20;
    const size_t len = ((size_t)bytes + huge - 1) / huge * huge;
    cudaError_t e = cudaHostUnregister(host);
    munmap(host, len);
    if (e != cudaSuccess) {
        scone::set_error("scone_host_free: cudaHostUnregister failed: %s", cudaGetErrorString(e));
        return SCONE_E_CUDA;
    }
    return SCONE_OK;
}

int scone_host_gather_rows(const void *h_rows, int64_t row_stride, int64_t num_rows, const int32_t *h_row_ids, int64_t k,
                           void *h_staging, int32_t nthreads) {
    SCONE_REQUIRE(k >= 0 && row_stride > 0, "scone_host_gather_rows: bad sizes");
    if (k == 0) return SCONE_OK;
    SCONE_REQUIRE(h_rows && h_row_ids && h_staging, "scone_host_gather_rows: NULL buffer");
    if (nthreads < 1) nthreads = 1;
    if (nthreads > 256) nthreads = 256;
    if ((int64_t)nthreads > k) nthreads = (int)k;
    std::atomic<int> bad{0};
    auto work = [&](int64_t lo, int64_t hi) {
        const uint8_t *src = static_cast<const uint8_t *>(h_rows);
        uint8_t *dst = static_cast<uint8_t *>(h_staging);
        for (int64_t r = lo; r < hi; ++r) {
            if (r + 4 < hi) {  // the rows are random: start the next ones' first lines towards the cache while this one is copied
                const int64_t nx = h_row_ids[r + 4];
                if (nx >= 0 && nx < num_rows) {
                    const uint8_t *q = src + nx * row_stride;
                    __builtin_prefetch(q);
                    __builtin_prefetch(q + 64);
                    __builtin_prefetch(q + 128);
                    __builtin_prefetch(q + 192);
                }
            }
            const int64_t id = h_row_ids[r];
            if (id < 0 || id >= num_rows) {
                bad.store(1);
                continue;
            }
            std::memcpy(dst + r * row_stride, src + id * row_stride, (size_t)row_stride);
        }
    };
    if (nthreads == 1) {
        work(0, k);
    } else {
        std::vector<std::thread> pool;
        const int64_t per = (k + nthreads - 1) / nthreads;
        for (int t = 0; t < nthreads; ++t) {
            const int64_t lo = t * per, hi = lo + per < k ? lo + per : k;
            if (lo < hi) pool.emplace_back(work, lo, hi);
        }
        for (auto &th : pool) th.join();
    }
    SCONE_REQUIRE(bad.load() == 0, "scone_host_gather_rows: row id outside [0, %lld)", (long long)num_rows);
    return SCONE_OK;
}

}  // extern "C"
