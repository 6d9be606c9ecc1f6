// fit.cu -- vocabulary construction on the device: NGramExtractor.fit (scone/tokenization/n_gram_extractor.py:72-104).
//
// The reference counts every n-gram (n = 1..max_n, never across texts) of a tokenised corpus in a Python Counter (:58-70,
// :86-88), keeps the max_f_grams most frequent (ties: first seen, in its enumeration order text -> n -> start; :91), THEN drops
// those below min_freq (:92-94); id = rank (:98-99).  Same result here, without materialising the n-grams:
//
//   1. one 64-bit hash per n-gram occurrence, written in the reference's enumeration order, with (position, n) as its payload;
//   2. a stable radix sort of the hashes (cub::DeviceRadixSort -- the one library call): equal n-grams become adjacent and the
//      first member of each run is its first occurrence in enumeration order;
//   3. run heads, run lengths (= counts) and a VERIFY pass: every member of a run must spell the same n-gram as its
//      predecessor -- two different n-grams behind one hash are detected (the caller retries with another seed), not trusted;
//   4. runs ranked by (count descending, first occurrence ascending): two stable sorts;
//   5. the top max_f_grams with count >= min_freq are written out as tokens int32 [n, max_n] (-1 padded) + lens.
// Device memory: 32 bytes per n-gram occurrence (hash + payload, double-buffered by the sort) + the sort's scratch: a
// 10^8-token corpus with max_n = 5 needs ~17 GB of the 180 GB.
#include <cub/cub.cuh>

#include "common.cuh"

namespace scone {

// text t covers tokens [off[t], off[t+1]); it contributes sum_n max(0, len - n + 1) occurrences, laid out n-major
__device__ __forceinline__ int64_t items_before_n(int64_t len, int n) {  // occurrences of lengths 1 .. n-1 in a text of `len` tokens
    int64_t s = 0;
    for (int m = 1; m < n; ++m) s += len - m + 1 > 0 ? len - m + 1 : 0;
    return s;
}

__global__ void __launch_bounds__(256) fit_text_items_kernel(const int64_t *__restrict__ off, int64_t num_texts, int max_n, int64_t *items) {
    const int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (t < num_texts) items[t] = items_before_n(off[t + 1] - off[t], max_n + 1);
}

__device__ __forceinline__ int64_t text_of(const int64_t *__restrict__ off, int64_t num_texts, int64_t p) {  // largest t with off[t] <= p
    int64_t lo = 0, hi = num_texts - 1;
    while (lo < hi) {
        const int64_t mid = (lo + hi + 1) >> 1;
        if (off[mid] <= p) lo = mid;
        else hi = mid - 1;
    }
    return lo;
}

__device__ __forceinline__ uint64_t gram_hash(const int32_t *__restrict__ tok, int64_t p, int n, uint64_t seed) {
    uint64_t h = hash_seed() ^ seed;
    for (int k = 0; k < n; ++k) h = hash_roll(h, (uint32_t)tok[p + k]);
    return hash_finish(h, n);
}

// one thread per corpus position p: the n-grams STARTING at p, each at its slot of the enumeration order
__global__ void __launch_bounds__(256) fit_emit_kernel(const int32_t *__restrict__ tok, int64_t M, const int64_t *__restrict__ off, int64_t num_texts,
                                                      const int64_t *__restrict__ text_base, int max_n, uint64_t seed, uint64_t *keys,
                                                      uint64_t *vals, unsigned int *bad_token) {
    const int64_t p = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (p >= M) return;
    if (tok[p] < 0) atomicOr(bad_token, 1u);
    const int64_t t = text_of(off, num_texts, p);
    const int64_t start = p - off[t], len = off[t + 1] - off[t];
    for (int n = 1; n <= max_n && start + n <= len; ++n) {
        const int64_t item = text_base[t] + items_before_n(len, n) + start;
        keys[item] = gram_hash(tok, p, n, seed);
        vals[item] = ((uint64_t)p << 3) | (uint64_t)n;
    }
}

// after the sort: head flags, and the verification that a run holds ONE n-gram
__global__ void __launch_bounds__(256) fit_heads_kernel(const int32_t *__restrict__ tok, const uint64_t *__restrict__ keys,
                                                       const uint64_t *__restrict__ vals, int64_t I, uint8_t *head, unsigned int *collisions) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= I) return;
    const bool h = i == 0 || keys[i] != keys[i - 1];
    head[i] = h ? 1 : 0;
    if (!h) {
        const uint64_t a = vals[i], b = vals[i - 1];
        const int n = (int)(a & 7), nb = (int)(b & 7);
        bool same = n == nb;
        for (int k = 0; same && k < n; ++k) same = tok[(a >> 3) + k] == tok[(b >> 3) + k];
        if (!same) atomicAdd(collisions, 1u);
    }
}

// first occurrence (enumeration index) of each distinct n-gram, from its representative's (position, n)
__global__ void __launch_bounds__(256) fit_first_kernel(const uint64_t *__restrict__ rep, int64_t R, const int64_t *__restrict__ off, int64_t num_texts,
                                                       const int64_t *__restrict__ text_base, uint64_t *first, uint32_t *run_index) {
    const int64_t r = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (r >= R) return;
    const int64_t p = (int64_t)(rep[r] >> 3);
    const int n = (int)(rep[r] & 7);
    const int64_t t = text_of(off, num_texts, p);
    first[r] = (uint64_t)(text_base[t] + items_before_n(off[t + 1] - off[t], n) + (p - off[t]));
    run_index[r] = (uint32_t)r;
}

__global__ void __launch_bounds__(256) fit_gather_count_kernel(const uint32_t *__restrict__ order, const uint32_t *__restrict__ counts, int64_t R,
                                                              uint32_t *neg_count) {
    const int64_t r = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (r < R) neg_count[r] = ~counts[order[r]];  // ascending sort of ~count = descending count
}

__global__ void __launch_bounds__(256) fit_write_kernel(const int32_t *__restrict__ tok, const uint32_t *__restrict__ order, const uint64_t *__restrict__ rep,
                                                       const uint32_t *__restrict__ counts, int64_t top, int64_t min_freq, int max_n,
                                                       int32_t *out_tokens, uint8_t *out_lens, int64_t *out_counts, unsigned long long *n_out) {
    const int64_t j = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (j >= top) return;
    const uint32_t r = order[j];
    if ((int64_t)counts[r] < min_freq) return;  // counts are descending in j: the survivors are a prefix
    const int64_t p = (int64_t)(rep[r] >> 3);
    const int n = (int)(rep[r] & 7);
    for (int k = 0; k < max_n; ++k) out_tokens[j * max_n + k] = k < n ? tok[p + k] : -1;
    out_lens[j] = (uint8_t)n;
    if (out_counts) out_counts[j] = counts[r];
    atomicAdd(n_out, 1ull);
}

struct DeviceBuf {  // RAII for the builder's scratch (stream-ordered allocations)
    void *p = nullptr;
    cudaStream_t s;
    explicit DeviceBuf(cudaStream_t stream) : s(stream) {}
    cudaError_t alloc(size_t bytes) { return cudaMallocAsync(&p, bytes ? bytes : 16, s); }
    ~DeviceBuf() {
        if (p) cudaFreeAsync(p, s);
    }
    template <typename T>
    T *as() {
        return static_cast<T *>(p);
    }
};

}  // namespace scone

using namespace scone;

#define FIT_ALLOC(buf, bytes)                                                                                  \
    do {                                                                                                       \
        cudaError_t _e = (buf).alloc(bytes);                                                                   \
        if (_e != cudaSuccess) {                                                                               \
            set_error("scone_fit_vocab: allocating %llu bytes failed: %s", (unsigned long long)(bytes), cudaGetErrorString(_e)); \
            cudaGetLastError();                                                                                \
            return _e == cudaErrorMemoryAllocation ? SCONE_E_NOMEM : SCONE_E_CUDA;                             \
        }                                                                                                      \
    } while (0)

extern "C" int scone_fit_vocab(const int32_t *d_tokens, int64_t num_tokens, const int64_t *d_text_offsets, int64_t num_texts, int32_t max_n,
                               int64_t min_freq, int64_t max_f_grams, uint64_t seed, int32_t *d_out_tokens, uint8_t *d_out_lens,
                               int64_t *d_out_counts, int64_t *out_n, int64_t *out_distinct, void *stream_) {
    cudaStream_t stream = (cudaStream_t)stream_;
    SCONE_REQUIRE(out_n, "scone_fit_vocab: out_n is NULL");
    *out_n = 0;
    if (out_distinct) *out_distinct = 0;
    SCONE_REQUIRE(max_n >= 1 && max_n <= SCONE_MAX_N, "scone_fit_vocab: max_n = %d outside [1, %d]", max_n, SCONE_MAX_N);
    SCONE_REQUIRE(num_tokens >= 0 && num_texts >= 0 && max_f_grams >= 0, "scone_fit_vocab: negative size");
    SCONE_REQUIRE(num_tokens < (1ll << 40), "scone_fit_vocab: corpus too large");
    if (num_tokens == 0 || num_texts == 0 || max_f_grams == 0) return SCONE_OK;
    SCONE_REQUIRE(d_tokens && d_text_offsets && d_out_tokens && d_out_lens, "scone_fit_vocab: NULL buffer");
    const int max_items_per_token = max_n;
    SCONE_REQUIRE(num_tokens * max_items_per_token < (1ll << 32) * 64, "scone_fit_vocab: too many n-gram occurrences");

    // ---- 1. enumeration layout: occurrences per text, exclusive scan --------------------------------------------------
    DeviceBuf items(stream), base(stream), scan_tmp(stream), flags(stream);
    FIT_ALLOC(items, (size_t)num_texts * 8);
    FIT_ALLOC(base, (size_t)(num_texts + 1) * 8);
    FIT_ALLOC(flags, 64);
    SCONE_CUDA(cudaMemsetAsync(flags.p, 0, 64, stream));
    unsigned int *d_bad = flags.as<unsigned int>(), *d_coll = d_bad + 1;
    unsigned long long *d_nout = reinterpret_cast<unsigned long long *>(d_bad + 2);
    int *d_runs = reinterpret_cast<int *>(d_bad + 4);
    {
        SCONE_GRID(blocks, (num_texts + 255) / 256, "scone_fit_vocab");
        fit_text_items_kernel<<<blocks, 256, 0, stream>>>(d_text_offsets, num_texts, max_n, items.as<int64_t>());
        SCONE_LAUNCHED();
    }
    size_t tmp_bytes = 0;
    SCONE_CUDA(cub::DeviceScan::ExclusiveSum(nullptr, tmp_bytes, items.as<int64_t>(), base.as<int64_t>(), (int)num_texts, stream));
    FIT_ALLOC(scan_tmp, tmp_bytes);
    SCONE_REQUIRE(num_texts < (1ll << 31), "scone_fit_vocab: too many texts");
    SCONE_CUDA(cub::DeviceScan::ExclusiveSum(scan_tmp.p, tmp_bytes, items.as<int64_t>(), base.as<int64_t>(), (int)num_texts, stream));
    int64_t last_base = 0, last_items = 0;
    SCONE_CUDA(cudaMemcpyAsync(&last_base, base.as<int64_t>() + (num_texts - 1), 8, cudaMemcpyDeviceToHost, stream));
    SCONE_CUDA(cudaMemcpyAsync(&last_items, items.as<int64_t>() + (num_texts - 1), 8, cudaMemcpyDeviceToHost, stream));
    SCONE_CUDA(cudaStreamSynchronize(stream));
    const int64_t I = last_base + last_items;  // n-gram occurrences in the corpus
    if (I == 0) return SCONE_OK;
    SCONE_REQUIRE(I < (1ll << 31), "scone_fit_vocab: %lld n-gram occurrences exceed 2^31 - 1 (split the corpus)", (long long)I);

    // ---- 2. hash every occurrence, sort ----------------------------------------------------------------------------------
    DeviceBuf k0(stream), k1(stream), v0(stream), v1(stream), sort_tmp(stream);
    FIT_ALLOC(k0, (size_t)I * 8);
    FIT_ALLOC(k1, (size_t)I * 8);
    FIT_ALLOC(v0, (size_t)I * 8);
    FIT_ALLOC(v1, (size_t)I * 8);
    {
        SCONE_GRID(blocks, (num_tokens + 255) / 256, "scone_fit_vocab");
        fit_emit_kernel<<<blocks, 256, 0, stream>>>(d_tokens, num_tokens, d_text_offsets, num_texts, base.as<int64_t>(), max_n, seed,
                                                    k0.as<uint64_t>(), v0.as<uint64_t>(), d_bad);
        SCONE_LAUNCHED();
    }
    cub::DoubleBuffer<uint64_t> dk(k0.as<uint64_t>(), k1.as<uint64_t>()), dv(v0.as<uint64_t>(), v1.as<uint64_t>());
    SCONE_CUDA(cub::DeviceRadixSort::SortPairs(nullptr, tmp_bytes, dk, dv, (int)I, 0, 64, stream));
    FIT_ALLOC(sort_tmp, tmp_bytes);
    SCONE_CUDA(cub::DeviceRadixSort::SortPairs(sort_tmp.p, tmp_bytes, dk, dv, (int)I, 0, 64, stream));  // LSD radix sort: stable
    const uint64_t *keys = dk.Current(), *vals = dv.Current();
    uint64_t *spare_k = dk.Alternate(), *spare_v = dv.Alternate();

    // ---- 3. run heads + verification, run lengths, representatives ---------------------------------------------------------
    DeviceBuf head(stream), counts(stream), rle_tmp(stream), sel_tmp(stream);
    FIT_ALLOC(head, (size_t)I);
    {
        SCONE_GRID(blocks, (I + 255) / 256, "scone_fit_vocab");
        fit_heads_kernel<<<blocks, 256, 0, stream>>>(d_tokens, keys, vals, I, head.as<uint8_t>(), d_coll);
        SCONE_LAUNCHED();
    }
    FIT_ALLOC(counts, (size_t)I * 4);
    // unique hashes go to the spare key buffer (not needed afterwards), counts to `counts`
    SCONE_CUDA(cub::DeviceRunLengthEncode::Encode(nullptr, tmp_bytes, keys, spare_k, counts.as<uint32_t>(), d_runs, (int)I, stream));
    FIT_ALLOC(rle_tmp, tmp_bytes);
    SCONE_CUDA(cub::DeviceRunLengthEncode::Encode(rle_tmp.p, tmp_bytes, keys, spare_k, counts.as<uint32_t>(), d_runs, (int)I, stream));
    SCONE_CUDA(cub::DeviceSelect::Flagged(nullptr, tmp_bytes, vals, head.as<uint8_t>(), spare_v, d_runs + 1, (int)I, stream));
    FIT_ALLOC(sel_tmp, tmp_bytes);
    SCONE_CUDA(cub::DeviceSelect::Flagged(sel_tmp.p, tmp_bytes, vals, head.as<uint8_t>(), spare_v, d_runs + 1, (int)I, stream));
    unsigned int h_flags[6] = {0, 0, 0, 0, 0, 0};
    SCONE_CUDA(cudaMemcpyAsync(h_flags, flags.p, sizeof h_flags, cudaMemcpyDeviceToHost, stream));
    SCONE_CUDA(cudaStreamSynchronize(stream));
    SCONE_REQUIRE(h_flags[0] == 0, "scone_fit_vocab: token ids must be in [0, 2^31)");
    if (h_flags[1] != 0) {
        set_error("scone_fit_vocab: %u n-gram occurrences share a 64-bit hash with a different n-gram (seed %llu): retry with another seed", h_flags[1],
                  (unsigned long long)seed);
        return SCONE_E_VOCAB;
    }
    const int64_t R = (int)h_flags[4];  // distinct n-grams
    if (out_distinct) *out_distinct = R;
    const uint64_t *rep = spare_v;      // (position << 3 | n) of each distinct n-gram's first occurrence

    // ---- 4. rank: count descending, first occurrence ascending ----------------------------------------------------------------
    DeviceBuf first0(stream), first1(stream), ord0(stream), ord1(stream), neg0(stream), neg1(stream), rank_tmp(stream), rank_tmp2(stream);
    FIT_ALLOC(first0, (size_t)R * 8);
    FIT_ALLOC(first1, (size_t)R * 8);
    FIT_ALLOC(ord0, (size_t)R * 4);
    FIT_ALLOC(ord1, (size_t)R * 4);
    FIT_ALLOC(neg0, (size_t)R * 4);
    FIT_ALLOC(neg1, (size_t)R * 4);
    {
        SCONE_GRID(blocks, (R + 255) / 256, "scone_fit_vocab");
        fit_first_kernel<<<blocks, 256, 0, stream>>>(rep, R, d_text_offsets, num_texts, base.as<int64_t>(), first0.as<uint64_t>(), ord0.as<uint32_t>());
        SCONE_LAUNCHED();
    }
    cub::DoubleBuffer<uint64_t> df(first0.as<uint64_t>(), first1.as<uint64_t>());
    cub::DoubleBuffer<uint32_t> dord(ord0.as<uint32_t>(), ord1.as<uint32_t>());
    SCONE_CUDA(cub::DeviceRadixSort::SortPairs(nullptr, tmp_bytes, df, dord, (int)R, 0, 40, stream));
    FIT_ALLOC(rank_tmp, tmp_bytes);
    SCONE_CUDA(cub::DeviceRadixSort::SortPairs(rank_tmp.p, tmp_bytes, df, dord, (int)R, 0, 40, stream));
    {
        SCONE_GRID(blocks, (R + 255) / 256, "scone_fit_vocab");
        fit_gather_count_kernel<<<blocks, 256, 0, stream>>>(dord.Current(), counts.as<uint32_t>(), R, neg0.as<uint32_t>());
        SCONE_LAUNCHED();
    }
    cub::DoubleBuffer<uint32_t> dneg(neg0.as<uint32_t>(), neg1.as<uint32_t>());
    SCONE_CUDA(cub::DeviceRadixSort::SortPairs(nullptr, tmp_bytes, dneg, dord, (int)R, 0, 32, stream));
    FIT_ALLOC(rank_tmp2, tmp_bytes);
    SCONE_CUDA(cub::DeviceRadixSort::SortPairs(rank_tmp2.p, tmp_bytes, dneg, dord, (int)R, 0, 32, stream));  // stable: ties keep first-seen order

    // ---- 5. truncate to max_f_grams, THEN drop counts below min_freq (reference :91-94), write the vocabulary ----------------------
    const int64_t top = R < max_f_grams ? R : max_f_grams;
    {
        SCONE_GRID(blocks, (top + 255) / 256, "scone_fit_vocab");
        fit_write_kernel<<<blocks, 256, 0, stream>>>(d_tokens, dord.Current(), rep, counts.as<uint32_t>(), top, min_freq, max_n, d_out_tokens,
                                                     d_out_lens, d_out_counts, d_nout);
        SCONE_LAUNCHED();
    }
    unsigned long long n_out = 0;
    SCONE_CUDA(cudaMemcpyAsync(&n_out, d_nout, 8, cudaMemcpyDeviceToHost, stream));
    SCONE_CUDA(cudaStreamSynchronize(stream));
    *out_n = (int64_t)n_out;
    return SCONE_OK;
}
