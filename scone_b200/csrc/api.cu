// api.cu -- library-level entry points: version, thread-local error string, launch counter.
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

}  // extern "C"
