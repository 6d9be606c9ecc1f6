// match.cuh -- warp-cooperative f-gram matching shared by the standalone lookup and the fused path.
//
// A warp owns a window of G = 32 / P consecutive positions of the flattened [B * L] batch, where
// P = lanes per position = the power of two >= max_n.  Lane j*P + (n-1) probes the n-gram ENDING
// at position base + j; the P candidates of a position are resolved concurrently and the longest
// hit is picked with one ballot (Algorithm 2: n = max_n .. 1, first hit wins).
#pragma once
#include "common.cuh"

namespace scone {

struct WindowMatch {
    int32_t fid;  // -1 = no f-gram ends here
    int32_t len;  // 0 = none
};

// position of flat index i inside its row of length L
__device__ __forceinline__ int64_t pos_in_row(int64_t i, int64_t L, int64_t T) {
    if (T <= 0xFFFFFFFFll) return (int64_t)((uint32_t)i % (uint32_t)L);
    return i % L;
}

// Lanes 0 .. G+P-2 of the warp that owns window `base` hold the tokens base-(P-1) .. base+G-1 (as int32;
// vocabulary tokens are non-negative int32, anything else is mapped to -1 and can only miss).
template <int P>
__device__ __forceinline__ int32_t load_window_token(const int64_t *__restrict__ ids, int64_t T, int64_t base, int lane) {
    constexpr int G = 32 / P;
    const int64_t gi = base - (P - 1) + lane;
    int64_t t64 = -1;
    if (lane < G + P - 1 && gi >= 0 && gi < T) t64 = __ldg(ids + gi);
    return (t64 >= 0 && t64 <= 0x7FFFFFFFll) ? (int32_t)t64 : -1;
}

// token of the lane's own position (lane / P) out of the window registers
template <int P>
__device__ __forceinline__ int32_t own_token(int32_t tok, int lane) {
    return __shfl_sync(0xFFFFFFFFu, tok, lane / P + (P - 1));
}

// id of the candidate this lane is responsible for, or -1.  `tok` from load_window_token.
template <int P>
__device__ __forceinline__ int32_t candidate_id(const IndexView &ix, int32_t tok, int64_t T, int64_t L, int64_t base, int lane,
                                                bool use_len_mask) {
    constexpr unsigned FULL = 0xFFFFFFFFu;
    const int j = lane / P, n = lane % P + 1;
    const int64_t i = base + j;
    bool cand = n <= ix.max_n && i < T;
    if (cand) {
        cand = (int64_t)n <= pos_in_row(i, L, T) + 1;  // the n-gram must not cross the row start
        if (use_len_mask) cand = cand && ((ix.len_mask >> (n - 1)) & 1u);
    }
    int32_t key[7];
    uint64_t h = hash_seed();
#pragma unroll
    for (int k = 0; k < 7; ++k) key[k] = -1;
#pragma unroll
    for (int k = 0; k < (P < 7 ? P : 7); ++k) {
        const int32_t t = __shfl_sync(FULL, tok, j + (P - 1) - k);  // token at position i - k
        if (k < n) {
            key[k] = t;
            cand = cand && t >= 0;
            h = hash_roll(h, (uint32_t)t);
        }
    }
    if (!cand) return -1;
    return probe(ix, hash_finish(h, n), key);
}

// Longest hit of the lane's position; every lane of a P-lane group returns the same value.
template <int P>
__device__ __forceinline__ WindowMatch match_window(const IndexView &ix, int32_t tok, int64_t T, int64_t L, int64_t base, int lane) {
    constexpr unsigned FULL = 0xFFFFFFFFu;
    const int32_t cid = candidate_id<P>(ix, tok, T, L, base, lane, true);
    const unsigned hit = __ballot_sync(FULL, cid >= 0);
    const int j = lane / P;
    const unsigned bits = (hit >> (j * P)) & ((1u << P) - 1u);
    WindowMatch m{-1, 0};
    int src = lane;
    if (bits) {
        const int top = 31 - __clz((int)bits);
        m.len = top + 1;
        src = j * P + top;
    }
    const int32_t f = __shfl_sync(FULL, cid, src);
    if (bits) m.fid = f;
    return m;
}

}  // namespace scone
