// index.cu -- the device-resident f-gram index: build, audit, longest-match lookup, match-all.
//
// Takes over the Python set / dict of int tuples of the reference's NGramExtractor
// (scone/tokenization/n_gram_extractor.py:41-44, :96-99, :121-122) and the tuple -> id lookup of
// scone/inference/embedding_cache.py:173.
#include <cstdlib>
#include <cstring>

#include "common.cuh"
#include "match.cuh"

namespace scone {

// audit counters read back once at create time
struct BuildStats {
    unsigned long long bad_len;
    unsigned long long bad_tok;
    unsigned long long dup;
    unsigned int len_mask;
    int max_probe;
    int max_tok;
};

// First pass: validate the vocabulary and find the largest token (decides the slot format).
__global__ void __launch_bounds__(256) index_scan_kernel(const int32_t *__restrict__ tokens, const uint8_t *__restrict__ lens, int64_t n,
                                                         int32_t max_n, BuildStats *stats) {
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const int len = lens[i];
    if (len < 1 || len > max_n) {
        atomicAdd(&stats->bad_len, 1ull);
        return;
    }
    int mx = 0;
    bool bad = false;
    for (int k = 0; k < len; ++k) {
        const int32_t t = tokens[i * max_n + k];
        bad |= t < 0;
        mx = t > mx ? t : mx;
    }
    if (bad) atomicAdd(&stats->bad_tok, 1ull);
    else atomicMax(&stats->max_tok, mx);
}

// One thread per f-gram: claim the first free slot of its probe sequence with a CAS on the id
// word, then fill in the key.  No lookup runs concurrently with the build.
__global__ void __launch_bounds__(256) index_build_kernel(const int32_t *__restrict__ tokens, const uint8_t *__restrict__ lens,
                                                          int64_t n, int32_t max_n, Slot *slots, uint64_t cap, int compact,
                                                          uint32_t *filter, uint32_t filter_mask, BuildStats *stats) {
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const int len = lens[i];
    if (len < 1 || len > max_n) return;  // counted by the scan
    int32_t key[7];
    uint64_t h = hash_seed();
    bool bad = false;
#pragma unroll
    for (int k = 0; k < 7; ++k) {
        key[k] = -1;
        if (k < len) {
            int32_t t = tokens[i * max_n + (len - 1 - k)];
            bad |= t < 0;
            key[k] = t;
            h = hash_roll(h, (uint32_t)t);
        }
    }
    if (bad) return;  // counted by the scan
    h = hash_finish(h, len);
    if (filter) atomicOr(&filter[filter_word(h, filter_mask)], filter_bits(h));
    int probes = 1;
    if (compact == kSlotCompact20) {
        Slot20 *s20 = reinterpret_cast<Slot20 *>(slots);
        uint32_t k20[4];
        pack_key20(key, k20);  // validity was established by the scan pass (max token, max_n)
        uint64_t s = home_slot16(h, cap);
        for (;;) {
            uint32_t old = atomicCAS(&s20[s].w[0], kEmpty20, (uint32_t)i | k20[0]);
            if (old == kEmpty20) break;
            if (++s == cap) s = 0;
            ++probes;
        }
#pragma unroll
        for (int k = 1; k < 4; ++k) s20[s].w[k] = k20[k];
    } else if (compact == kSlotCompact16) {
        Slot16 *s16 = reinterpret_cast<Slot16 *>(slots);
        uint64_t s = home_slot16(h, cap);
        for (;;) {
            int32_t old = atomicCAS(&s16[s].id, -1, (int32_t)i);
            if (old == -1) break;
            if (++s == cap) s = 0;
            ++probes;
        }
        uint32_t k16[3];
        pack_key16(key, k16);
#pragma unroll
        for (int k = 0; k < 3; ++k) s16[s].t[k] = k16[k];
    } else {
        uint64_t s = home_slot(h, cap);
        for (;;) {
            int32_t old = atomicCAS(&slots[s].w[0], -1, (int32_t)i);
            if (old == -1) break;
            if (++s == cap) s = 0;
            ++probes;
        }
#pragma unroll
        for (int k = 0; k < 7; ++k) slots[s].w[k + 1] = key[k];
    }
    atomicOr(&stats->len_mask, 1u << (len - 1));
    atomicMax(&stats->max_probe, probes);
}

// After the build: every f-gram must find ITSELF.  Two rows with the same key resolve to the
// same (first) slot, so the later row sees a different id -> counted as a duplicate.
__global__ void __launch_bounds__(256) index_audit_kernel(const int32_t *__restrict__ tokens, const uint8_t *__restrict__ lens,
                                                          int64_t n, int32_t max_n, IndexView ix, BuildStats *stats) {
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const int len = lens[i];
    if (len < 1 || len > max_n) return;
    int32_t key[7];
    uint64_t h = hash_seed();
#pragma unroll
    for (int k = 0; k < 7; ++k) {
        key[k] = -1;
        if (k < len) {
            int32_t t = tokens[i * max_n + (len - 1 - k)];
            if (t < 0) return;
            key[k] = t;
            h = hash_roll(h, (uint32_t)t);
        }
    }
    h = hash_finish(h, len);
    if (probe_any(ix, h, key) != (int32_t)i) atomicAdd(&stats->dup, 1ull);
}

// ---------------------------------------------------------------------------------------------
// standalone match kernels (the fused path in embed.cu uses the same match_window())
// ---------------------------------------------------------------------------------------------
template <int P>
__global__ void __launch_bounds__(256) lookup_kernel(IndexView ix, const int64_t *__restrict__ ids, int64_t T, int64_t L,
                                                     int32_t *__restrict__ out_id, uint8_t *__restrict__ out_len) {
    constexpr int G = 32 / P;
    const int lane = threadIdx.x & 31;
    const int64_t warp = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int64_t base = warp * G;
    if (base >= T) return;
    const int back = ix.max_n - 1;
    WindowMatch m = match_window<P>(ix, load_window_token<P>(ids, T, base, lane, back), T, L, base, lane, back);
    const int j = lane / P;
    if ((lane % P) == 0 && base + j < T) {
        if (out_id) out_id[base + j] = m.fid;
        if (out_len) out_len[base + j] = (uint8_t)m.len;
    }
}

template <int P>
__global__ void __launch_bounds__(256) match_all_kernel(IndexView ix, const int64_t *__restrict__ ids, int64_t T, int64_t L,
                                                        int32_t *__restrict__ out) {
    constexpr int G = 32 / P;
    const int lane = threadIdx.x & 31;
    const int64_t warp = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int64_t base = warp * G;
    if (base >= T) return;
    const int back = ix.max_n - 1;
    const int j = lane / P, n1 = lane % P;  // dense mapping: candidate n1 is length n1 + 1
    const bool present = n1 < ix.max_n && ((ix.len_mask >> n1) & 1u);
    int32_t fid = candidate_id<P>(ix, load_window_token<P>(ids, T, base, lane, back), T, L, base, lane, present ? n1 + 1 : 0, back);
    if (n1 < ix.max_n && base + j < T) out[(base + j) * ix.max_n + n1] = fid;
}

static int lanes_dense(int max_n) { return max_n <= 1 ? 1 : max_n <= 2 ? 2 : max_n <= 4 ? 4 : 8; }

}  // namespace scone

using namespace scone;

extern "C" {

int scone_index_create(const int32_t *d_tokens, const uint8_t *d_lens, int64_t n, int32_t max_n, double load_factor,
                       void *stream_, scone_index_t **out) {
    cudaStream_t stream = (cudaStream_t)stream_;
    SCONE_REQUIRE(out != nullptr, "scone_index_create: out is NULL");
    *out = nullptr;
    SCONE_REQUIRE(n >= 0 && n < (int64_t)0x7FFFFFFF, "scone_index_create: n = %lld outside [0, 2^31-1)", (long long)n);
    SCONE_REQUIRE(max_n >= 1 && max_n <= SCONE_MAX_N, "scone_index_create: max_n = %d outside [1, %d]", max_n, SCONE_MAX_N);
    SCONE_REQUIRE(n == 0 || (d_tokens && d_lens), "scone_index_create: NULL vocabulary arrays");
    SCONE_REQUIRE(!(load_factor > 0.9), "scone_index_create: load_factor %.3f > 0.9", load_factor);

    scone_index_impl *ix = new (std::nothrow) scone_index_impl();
    if (!ix) {
        set_error("scone_index_create: out of host memory");
        return SCONE_E_NOMEM;
    }
    ix->n = n;
    ix->max_n = max_n;
    BuildStats *d_stats = nullptr;
    BuildStats h{};
    const unsigned blocks = (unsigned)((n + 255) / 256);
    int rc = [&]() -> int {
        SCONE_CUDA(cudaGetDevice(&ix->device));
        SCONE_CUDA(cudaMalloc(&d_stats, sizeof(BuildStats)));
        SCONE_CUDA(cudaMemsetAsync(d_stats, 0, sizeof(BuildStats), stream));
        // pass 1: validate, find the largest token -> slot format
        if (n > 0) {
            index_scan_kernel<<<blocks, 256, 0, stream>>>(d_tokens, d_lens, n, max_n, d_stats);
            SCONE_LAUNCHED();
        }
        SCONE_CUDA(cudaMemcpyAsync(&h, d_stats, sizeof h, cudaMemcpyDeviceToHost, stream));
        SCONE_CUDA(cudaStreamSynchronize(stream));
        if (h.bad_len || h.bad_tok) return SCONE_OK;  // reported below
        const char *fmt = getenv("SCONE_INDEX_FORMAT");  // "wide" / "compact" / "compact20" override the automatic choice (testing)
        const bool fits16 = max_n <= 6 && h.max_tok < 0xFFFF;
        const bool fits20 = max_n <= 5 && h.max_tok < (int)kPad20 && n < (int64_t)kIdMask20;
        ix->compact = fits16 ? kSlotCompact16 : fits20 ? kSlotCompact20 : kSlotWide;
        if (fmt && !strcmp(fmt, "wide")) ix->compact = kSlotWide;
        if (fmt && !strcmp(fmt, "compact")) {
            if (!fits16) {
                set_error("scone_index_create: SCONE_INDEX_FORMAT=compact needs max_n <= 6 and tokens < 65535");
                return SCONE_E_INVALID;
            }
            ix->compact = kSlotCompact16;
        }
        if (fmt && !strcmp(fmt, "prefer-compact20") && fits20) ix->compact = kSlotCompact20;  // testing: small vocabularies too
        if (fmt && !strcmp(fmt, "compact20")) {
            if (!fits20) {
                set_error("scone_index_create: SCONE_INDEX_FORMAT=compact20 needs max_n <= 5, tokens < 1048575 and n < 2^28 - 1");
                return SCONE_E_INVALID;
            }
            ix->compact = kSlotCompact20;
        }
        // default load factor 0.25 for both formats (measured: compact at 0.5 costs config 1 a microsecond of probe chain)
        const double lf = load_factor > 0.0 ? load_factor : 0.25;
        uint64_t cap = (uint64_t)((double)n / lf) + 1;
        if (cap < 64) cap = 64;
        cap = (cap + 3) & ~3ull;  // whole 64-byte blocks of either format
        ix->cap = cap;
        const size_t bytes = cap * (ix->compact ? sizeof(Slot16) : sizeof(Slot));
        cudaError_t e = cudaMalloc(&ix->slots, bytes);
        if (e != cudaSuccess) {
            set_error("scone_index_create: cudaMalloc of %llu slot bytes failed: %s", (unsigned long long)bytes, cudaGetErrorString(e));
            ix->slots = nullptr;
            return e == cudaErrorMemoryAllocation ? SCONE_E_NOMEM : SCONE_E_CUDA;
        }
        SCONE_CUDA(cudaMemsetAsync(ix->slots, 0xFF, bytes, stream));
        // pre-filter (common.cuh: filter_pass): 16 bits per f-gram up to 12 M f-grams (24 MB), 8 bits up to 48 M (48 MB), none
        // above -- it only pays while it stays in L2; SCONE_INDEX_FILTER=never / always override (testing)
        const char *fenv = getenv("SCONE_INDEX_FILTER");
        const bool f_never = fenv && !strcmp(fenv, "never"), f_always = fenv && !strcmp(fenv, "always");
        if (n > 0 && !f_never && (f_always || n <= 48000000ll)) {
            uint64_t words = (uint64_t)(n <= 12000000ll ? (n + 1) / 2 : (n + 3) / 4);
            if (words < 1024) words = 1024;
            words = (words + 31) & ~31ull;
            e = cudaMalloc(&ix->filter, words * sizeof(uint32_t));
            if (e != cudaSuccess) {
                set_error("scone_index_create: cudaMalloc of %llu filter bytes failed: %s", (unsigned long long)(words * 4), cudaGetErrorString(e));
                ix->filter = nullptr;
                return e == cudaErrorMemoryAllocation ? SCONE_E_NOMEM : SCONE_E_CUDA;
            }
            ix->filter_mask = (uint32_t)words;
            ix->filter_always = f_always ? 1 : 0;
            SCONE_CUDA(cudaMemsetAsync(ix->filter, 0, words * sizeof(uint32_t), stream));
        }
        // pass 2: insert, then every f-gram looks itself up (duplicate audit)
        if (n > 0) {
            index_build_kernel<<<blocks, 256, 0, stream>>>(d_tokens, d_lens, n, max_n, ix->slots, cap, ix->compact, ix->filter,
                                                           ix->filter_mask, d_stats);
            SCONE_LAUNCHED();
            // the audit goes through the filter too: an f-gram the filter rejected would not find itself
            IndexView v{ix->slots, cap, 0xFFFFFFFFu, max_n, ix->compact, ix->filter, ix->filter_mask};
            index_audit_kernel<<<blocks, 256, 0, stream>>>(d_tokens, d_lens, n, max_n, v, d_stats);
            SCONE_LAUNCHED();
        }
        SCONE_CUDA(cudaMemcpyAsync(&h, d_stats, sizeof h, cudaMemcpyDeviceToHost, stream));
        SCONE_CUDA(cudaStreamSynchronize(stream));
        return SCONE_OK;
    }();
    if (d_stats) cudaFree(d_stats);
    if (rc == SCONE_OK && (h.bad_len || h.bad_tok || h.dup)) {
        set_error("scone_index_create: vocabulary rejected: %llu duplicated f-grams, %llu lengths outside [1, %d], %llu rows with negative tokens",
                  h.dup, h.bad_len, max_n, h.bad_tok);
        rc = SCONE_E_VOCAB;
    }
    if (rc != SCONE_OK) {
        if (ix->slots) cudaFree(ix->slots);
        if (ix->filter) cudaFree(ix->filter);
        delete ix;
        return rc;
    }
    ix->len_mask = h.len_mask;
    ix->max_probe = h.max_probe;
    *out = reinterpret_cast<scone_index_t *>(ix);
    return SCONE_OK;
}

int scone_index_destroy(scone_index_t *index) {
    if (!index) return SCONE_OK;
    scone_index_impl *ix = reinterpret_cast<scone_index_impl *>(index);
    cudaError_t e = cudaFree(ix->slots);
    if (ix->filter) {
        cudaError_t e2 = cudaFree(ix->filter);
        if (e == cudaSuccess) e = e2;
    }
    delete ix;
    if (e != cudaSuccess) {
        set_error("scone_index_destroy: cudaFree failed: %s", cudaGetErrorString(e));
        return SCONE_E_CUDA;
    }
    return SCONE_OK;
}

int scone_index_info(const scone_index_t *index, scone_index_info_t *info) {
    SCONE_REQUIRE(index && info, "scone_index_info: NULL argument");
    const scone_index_impl *ix = reinterpret_cast<const scone_index_impl *>(index);
    info->num_fgrams = ix->n;
    info->capacity = (int64_t)ix->cap;
    info->filter_bytes = ix->filter ? (int64_t)ix->filter_mask * 4 : 0;
    info->bytes = (int64_t)(ix->cap * (ix->compact ? sizeof(Slot16) : sizeof(Slot))) + info->filter_bytes;
    info->max_n = ix->max_n;
    info->len_mask = ix->len_mask;
    info->max_probe = ix->max_probe;
    info->slot_bytes = (int32_t)(ix->compact ? sizeof(Slot16) : sizeof(Slot));
    info->slot_format = ix->compact;
    return SCONE_OK;
}

int scone_index_lookup(const scone_index_t *index, const int64_t *d_ids, int64_t B, int64_t L, int32_t *d_out_id,
                       uint8_t *d_out_len, void *stream_) {
    cudaStream_t stream = (cudaStream_t)stream_;
    SCONE_REQUIRE(index, "scone_index_lookup: NULL index");
    SCONE_REQUIRE(B >= 0 && L >= 0, "scone_index_lookup: negative shape");
    const int64_t T = B * L;
    if (T == 0) return SCONE_OK;
    SCONE_REQUIRE(d_ids, "scone_index_lookup: NULL ids");
    SCONE_REQUIRE(T < (1ll << 40), "scone_index_lookup: batch too large");
    const scone_index_impl *ix = reinterpret_cast<const scone_index_impl *>(index);
    IndexView v = view_of(ix, T);
    const int P = lanes_per_position(ix->len_mask, ix->max_n);
    const int64_t windows = (T + (32 / P) - 1) / (32 / P);
    SCONE_GRID(blocks, (windows + 7) / 8, "scone_index_lookup");
    switch (P) {
        case 1: lookup_kernel<1><<<blocks, 256, 0, stream>>>(v, d_ids, T, L, d_out_id, d_out_len); break;
        case 2: lookup_kernel<2><<<blocks, 256, 0, stream>>>(v, d_ids, T, L, d_out_id, d_out_len); break;
        case 4: lookup_kernel<4><<<blocks, 256, 0, stream>>>(v, d_ids, T, L, d_out_id, d_out_len); break;
        default: lookup_kernel<8><<<blocks, 256, 0, stream>>>(v, d_ids, T, L, d_out_id, d_out_len); break;
    }
    SCONE_LAUNCHED();
    return SCONE_OK;
}

int scone_index_match_all(const scone_index_t *index, const int64_t *d_ids, int64_t B, int64_t L, int32_t *d_out,
                          void *stream_) {
    cudaStream_t stream = (cudaStream_t)stream_;
    SCONE_REQUIRE(index, "scone_index_match_all: NULL index");
    SCONE_REQUIRE(B >= 0 && L >= 0, "scone_index_match_all: negative shape");
    const int64_t T = B * L;
    if (T == 0) return SCONE_OK;
    SCONE_REQUIRE(d_ids && d_out, "scone_index_match_all: NULL buffer");
    const scone_index_impl *ix = reinterpret_cast<const scone_index_impl *>(index);
    IndexView v = view_of(ix, T);
    const int P = lanes_dense(ix->max_n);
    const int64_t windows = (T + (32 / P) - 1) / (32 / P);
    SCONE_GRID(blocks, (windows + 7) / 8, "scone_index_match_all");
    switch (P) {
        case 1: match_all_kernel<1><<<blocks, 256, 0, stream>>>(v, d_ids, T, L, d_out); break;
        case 2: match_all_kernel<2><<<blocks, 256, 0, stream>>>(v, d_ids, T, L, d_out); break;
        case 4: match_all_kernel<4><<<blocks, 256, 0, stream>>>(v, d_ids, T, L, d_out); break;
        default: match_all_kernel<8><<<blocks, 256, 0, stream>>>(v, d_ids, T, L, d_out); break;
    }
    SCONE_LAUNCHED();
    return SCONE_OK;
}

}  // extern "C"
