// api.cu -- library-level entry points: version, thread-local error string, launch counter.
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
