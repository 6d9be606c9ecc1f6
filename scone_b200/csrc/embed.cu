// embed.cu -- the fused hot path: longest-match lookup + cache-row gather + dequant + fallback,
// one kernel, output written once in the transformer's inputs_embeds layout.
//
// Takes over, per position, get_token_f_grams + f_gram_to_id + get_embeddings + the engine's
// assemble loop + the wte fallback of the reference (scone/tokenization/n_gram_extractor.py:106-126,
// scone/inference/embedding_cache.py:149-181, scone/inference/engine.py:235-266,
// scone/models/language_model.py:239-243), with Algorithm-2 (replace-or-fallback) semantics.
//
// Shape of the kernel (HBM-bound gather, no tensor cores):
//   * persistent CTAs (a multiple of the 148 SMs), each 1 MATCHER warp + 8 GATHER warps.
//   * the matcher walks the CTA's tiles (a tile = the G = 32/P consecutive positions one warp can match
//     at once, match.cuh), resolves their f-gram ids and pushes (row id, fallback token) into a small
//     shared-memory ring guarded by mbarriers.  It runs up to kRing tiles ahead, so the dependent
//     chain ids -> hash -> slot -> (re-probe) is off the streaming warps' critical path.
//   * a gather warp pops one position at a time and streams its row: each lane owns 8 consecutive
//     elements (one 16 B output vector) per 256-element step, U steps are loaded back to back before
//     any is converted.  Cache / fallback rows are read with ld.global.nc.L1::no_allocate (touched
//     once), the output is written with 128-bit st.global.cs; slots use default caching so the much
//     smaller index stays L2-resident.
#include <cstdlib>

#include "common.cuh"
#include "match.cuh"

namespace scone {

struct EmbedParams {
    IndexView ix;
    const int32_t *fgram_in;  // not NULL: ids already resolved, the matcher only forwards them
    const uint8_t *rows;
    int64_t row_stride;
    int64_t num_rows;
    const uint8_t *base;  // [V, D] 16-bit
    int64_t V;
    const uint8_t *pos;  // [>= L, D] 16-bit or NULL
    const int64_t *ids;
    int64_t T, L;
    uint8_t *out;
    int32_t *out_id;
    uint8_t *out_len;
    uint32_t *status;
    int64_t num_tiles;
    int32_t D;
    int32_t scale_off;
    int32_t group_shift;  // log2(group / 8): chunk index >> group_shift = group index (INT4)
};

constexpr int kGatherWarps = 8;
constexpr int kThreads = 32 * (1 + kGatherWarps);
constexpr int kRing = 8;  // tiles the matcher may run ahead

template <int OUT>
__device__ __forceinline__ void decode16x8(uint4 raw, float (&x)[8]) {
    if (OUT == SCONE_OUT_BF16) decode_bf16x8(raw, x);
    else decode_fp16x8(raw, x);
}
template <int OUT>
__device__ __forceinline__ uint4 pack16x8(const float (&x)[8]) {
    return OUT == SCONE_OUT_BF16 ? pack_bf16x8(x) : pack_fp16x8(x);
}

// ---- streaming one position -------------------------------------------------------------------------

// Hit: dequantise table row `fid` into dst.  U 256-element steps in flight per lane.
template <int QUANT, int OUT, int U>
__device__ __forceinline__ void stream_hit(const EmbedParams &p, int32_t fid, uint8_t *__restrict__ dst, int lane) {
    const uint8_t *__restrict__ row = p.rows + (int64_t)fid * p.row_stride;
    const int nchunks = p.D >> 3;
    float rs = 1.0f;
    if (QUANT == SCONE_QUANT_INT8) rs = __ldg(reinterpret_cast<const float *>(row + p.scale_off));
    for (int c0 = lane; c0 < nchunks; c0 += 32 * U) {
        uint4 raw[U];
        float sc[U];
#pragma unroll
        for (int u = 0; u < U; ++u) {
            const int c = c0 + 32 * u;
            sc[u] = rs;
            if (c < nchunks) {
                if (QUANT == SCONE_QUANT_FP16) {
                    raw[u] = ldg_stream_16(row + c * 16);
                } else if (QUANT == SCONE_QUANT_INT8) {
                    const uint2 v = ldg_stream_8(row + c * 8);
                    raw[u].x = v.x;
                    raw[u].y = v.y;
                } else {
                    raw[u].x = ldg_stream_4(row + c * 4);
                    sc[u] = __half2float(__ldg(reinterpret_cast<const __half *>(row + p.scale_off) + (c >> p.group_shift)));
                }
            }
        }
#pragma unroll
        for (int u = 0; u < U; ++u) {
            const int c = c0 + 32 * u;
            if (c < nchunks) {
                uint4 o;
                if (QUANT == SCONE_QUANT_FP16 && OUT == SCONE_OUT_FP16) {
                    o = raw[u];
                } else {
                    float x[8];
                    if (QUANT == SCONE_QUANT_FP16) decode_fp16x8(raw[u], x);
                    else if (QUANT == SCONE_QUANT_INT8) decode_int8x8(make_uint2(raw[u].x, raw[u].y), sc[u], x);
                    else decode_int4x8(raw[u].x, sc[u], x);
                    o = pack16x8<OUT>(x);
                }
                stg_stream_16(dst + c * 16, o);
            }
        }
    }
}

// Miss: the fallback row is already in the output type -> 16 B copy.
template <int U>
__device__ __forceinline__ void stream_miss(const EmbedParams &p, int32_t tok, uint8_t *__restrict__ dst, int lane) {
    const uint8_t *__restrict__ row = p.base + (int64_t)tok * p.D * 2;
    const int nchunks = p.D >> 3;
    for (int c0 = lane; c0 < nchunks; c0 += 32 * U) {
        uint4 raw[U];
#pragma unroll
        for (int u = 0; u < U; ++u) {
            const int c = c0 + 32 * u;
            if (c < nchunks) raw[u] = ldg_stream_16(row + c * 16);
        }
#pragma unroll
        for (int u = 0; u < U; ++u) {
            const int c = c0 + 32 * u;
            if (c < nchunks) stg_stream_16(dst + c * 16, raw[u]);
        }
    }
}

// Everything else (position add, out-of-range token): correctness path, not tuned.
template <int QUANT, int OUT>
__device__ __noinline__ void stream_general(const EmbedParams &p, int32_t fid, int32_t tok, int64_t t, uint8_t *__restrict__ dst,
                                            int lane) {
    const int nchunks = p.D >> 3;
    const uint8_t *row = fid >= 0 ? p.rows + (int64_t)fid * p.row_stride : nullptr;
    const uint8_t *brow = (fid < 0 && tok >= 0) ? p.base + (int64_t)tok * p.D * 2 : nullptr;
    const uint8_t *prow = p.pos ? p.pos + pos_in_row(t, p.L, p.T) * p.D * 2 : nullptr;
    for (int c = lane; c < nchunks; c += 32) {
        float x[8];
        if (row) {
            if (QUANT == SCONE_QUANT_FP16) {
                decode_fp16x8(ldg_stream_16(row + c * 16), x);
            } else if (QUANT == SCONE_QUANT_INT8) {
                decode_int8x8(ldg_stream_8(row + c * 8), __ldg(reinterpret_cast<const float *>(row + p.scale_off)), x);
            } else {
                const __half hs = __ldg(reinterpret_cast<const __half *>(row + p.scale_off) + (c >> p.group_shift));
                decode_int4x8(ldg_stream_4(row + c * 4), __half2float(hs), x);
            }
        } else if (brow) {
            decode16x8<OUT>(ldg_stream_16(brow + c * 16), x);
        } else {
#pragma unroll
            for (int k = 0; k < 8; ++k) x[k] = 0.0f;
        }
        if (prow) {
            float y[8];
            decode16x8<OUT>(__ldg(reinterpret_cast<const uint4 *>(prow + c * 16)), y);
#pragma unroll
            for (int k = 0; k < 8; ++k) x[k] = __fadd_rn(x[k], y[k]);
        }
        stg_stream_16(dst + c * 16, pack16x8<OUT>(x));
    }
}

// ---- the kernel --------------------------------------------------------------------------------------

template <int QUANT, int OUT, int P, int U, int MINB>
__global__ void __launch_bounds__(kThreads, MINB) embed_kernel(const EmbedParams p) {
    constexpr int G = 32 / P;
    __shared__ __align__(8) uint64_t full_bar[kRing];
    __shared__ __align__(8) uint64_t empty_bar[kRing];
    __shared__ int2 ring[kRing][G];  // (row id or <0, fallback token or -1)

    const int lane = threadIdx.x & 31;
    const int warp = threadIdx.x >> 5;
    if (threadIdx.x == 0) {
#pragma unroll
        for (int q = 0; q < kRing; ++q) {
            mbar_init(&full_bar[q], 1);
            mbar_init(&empty_bar[q], G);
        }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();

    if (warp == 0) {
        // ===== matcher: resolve tiles, run ahead of the gather warps =====
        const int j = lane / P;
        int it = 0;
        for (int64_t tile = blockIdx.x; tile < p.num_tiles; tile += gridDim.x, ++it) {
            const int q = it % kRing;
            const int64_t base = tile * G;
            const int64_t i = base + j;
            int32_t fid = -1;
            if (p.fgram_in) {
                if (i < p.T) fid = __ldg(p.fgram_in + i);
                if (fid >= p.num_rows) fid = -2;  // caller error: zero row + status
            } else {
                const WindowMatch m = match_window<P>(p.ix, p.ids, p.T, p.L, base, lane);
                fid = m.fid;
                if ((lane % P) == 0 && i < p.T) {
                    if (p.out_id) p.out_id[i] = m.fid;
                    if (p.out_len) p.out_len[i] = (uint8_t)m.len;
                }
            }
            int32_t tok = -1;
            if (fid == -1 && i < p.T) {
                const int64_t t64 = __ldg(p.ids + i);
                if (t64 >= 0 && t64 < p.V) tok = (int32_t)t64;
            }
            mbar_wait(&empty_bar[q], ((it / kRing) & 1) ^ 1);
            if ((lane % P) == 0) ring[q][j] = make_int2(fid, tok);
            __syncwarp();
            if (lane == 0) mbar_arrive(&full_bar[q]);
        }
    } else {
        // ===== gather warps: pop positions round-robin, stream their rows =====
        bool flagged = false;
        const bool general = p.pos != nullptr;
        for (int64_t s = warp - 1;; s += kGatherWarps) {
            const int64_t itl = s / G;
            const int j = (int)(s - itl * G);
            const int64_t tile = blockIdx.x + itl * gridDim.x;
            if (tile >= p.num_tiles) break;
            const int q = (int)(itl % kRing);
            mbar_wait(&full_bar[q], (uint32_t)((itl / kRing) & 1));
            const int2 e = ring[q][j];
            const int64_t t = tile * G + j;
            if (t < p.T) {
                uint8_t *dst = p.out + t * p.D * 2;
                if (general || (e.x < 0 && e.y < 0)) {
                    stream_general<QUANT, OUT>(p, e.x, e.y, t, dst, lane);
                    flagged |= (e.x < 0 && e.y < 0);
                } else if (e.x >= 0) {
                    stream_hit<QUANT, OUT, U>(p, e.x, dst, lane);
                } else {
                    stream_miss<U>(p, e.y, dst, lane);
                }
            }
            __syncwarp();
            if (lane == 0) mbar_arrive(&empty_bar[q]);
        }
        if (flagged && p.status && lane == 0) atomicOr(p.status, SCONE_STATUS_TOKEN_OOR);
    }
}

static int lanes_per_token(int max_n) { return max_n <= 1 ? 1 : max_n <= 2 ? 2 : max_n <= 4 ? 4 : 8; }

static int num_sms() {
    static int n = 0;
    if (n == 0) {
        int dev = 0;
        if (cudaGetDevice(&dev) != cudaSuccess || cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || n <= 0)
            n = kNumSMsB200;
    }
    return n;
}

// Tuning hook (tools/tune_embed.py): SCONE_EMBED_VARIANT="U:MINB" selects another instantiation of the same
// kernel for the combinations compiled below; anything else runs the default.
static void variant(int &u, int &minb) {
    u = 0;
    minb = 0;
    if (const char *e = getenv("SCONE_EMBED_VARIANT")) sscanf(e, "%d:%d", &u, &minb);
}

template <int QUANT, int OUT, int P, int U, int MINB>
static void launch_one(EmbedParams &p, cudaStream_t stream) {
    constexpr int G = 32 / P;
    p.num_tiles = (p.T + G - 1) / G;
    const int64_t resident = (int64_t)num_sms() * MINB;
    const unsigned blocks = (unsigned)(p.num_tiles < resident ? p.num_tiles : resident);
    embed_kernel<QUANT, OUT, P, U, MINB><<<blocks, kThreads, 0, stream>>>(p);
}

template <int QUANT, int OUT, int P>
static void launch(EmbedParams &p, cudaStream_t stream) {
    if constexpr (OUT == SCONE_OUT_BF16 && (P == 4 || P == 8)) {
        int u, minb;
        variant(u, minb);
#define SCONE_V(UU, MM)                                        \
    if (u == UU && minb == MM) {                               \
        launch_one<QUANT, OUT, P, UU, MM>(p, stream);          \
        return;                                                \
    }
        SCONE_V(4, 3) SCONE_V(4, 5) SCONE_V(4, 6) SCONE_V(8, 3) SCONE_V(8, 4) SCONE_V(2, 6) SCONE_V(2, 4)
#undef SCONE_V
    }
    launch_one<QUANT, OUT, P, 4, 4>(p, stream);
}

template <int QUANT, int OUT>
static void launch_p(int P, EmbedParams &p, cudaStream_t stream) {
    switch (P) {
        case 1: launch<QUANT, OUT, 1>(p, stream); break;
        case 2: launch<QUANT, OUT, 2>(p, stream); break;
        case 4: launch<QUANT, OUT, 4>(p, stream); break;
        default: launch<QUANT, OUT, 8>(p, stream); break;
    }
}

static int dispatch(int P, EmbedParams &p, int quant, int out_dtype, cudaStream_t stream) {
    if (out_dtype == SCONE_OUT_BF16) {
        if (quant == SCONE_QUANT_FP16) launch_p<SCONE_QUANT_FP16, SCONE_OUT_BF16>(P, p, stream);
        else if (quant == SCONE_QUANT_INT8) launch_p<SCONE_QUANT_INT8, SCONE_OUT_BF16>(P, p, stream);
        else launch_p<SCONE_QUANT_INT4, SCONE_OUT_BF16>(P, p, stream);
    } else {
        if (quant == SCONE_QUANT_FP16) launch_p<SCONE_QUANT_FP16, SCONE_OUT_FP16>(P, p, stream);
        else if (quant == SCONE_QUANT_INT8) launch_p<SCONE_QUANT_INT8, SCONE_OUT_FP16>(P, p, stream);
        else launch_p<SCONE_QUANT_INT4, SCONE_OUT_FP16>(P, p, stream);
    }
    SCONE_LAUNCHED();
    return SCONE_OK;
}

int check_table(const scone_table_desc_t *t, const char *who);  // table.cu

static int fill_table(EmbedParams &p, const scone_table_desc_t *t, const char *who) {
    int rc = check_table(t, who);
    if (rc != SCONE_OK) return rc;
    p.rows = static_cast<const uint8_t *>(t->d_rows);
    p.row_stride = t->row_stride;
    p.num_rows = t->num_rows;
    p.D = t->dim;
    p.scale_off = t->scale_offset;
    p.group_shift = 0;
    if (t->quant == SCONE_QUANT_INT4) {
        int g8 = t->group / 8, sh = 0;
        while ((1 << sh) < g8) ++sh;
        p.group_shift = sh;
    }
    return SCONE_OK;
}

}  // namespace scone

using namespace scone;

extern "C" {

int scone_embed_forward(const scone_index_t *index, const scone_table_desc_t *table, const void *d_base_emb, int64_t base_rows,
                        const void *d_pos_emb, const int64_t *d_ids, int64_t B, int64_t L, void *d_out, int32_t out_dtype,
                        int32_t *d_out_id, uint8_t *d_out_len, uint32_t *d_status, void *stream_) {
    cudaStream_t stream = (cudaStream_t)stream_;
    SCONE_REQUIRE(index && table, "scone_embed_forward: NULL index or table");
    SCONE_REQUIRE(out_dtype == SCONE_OUT_BF16 || out_dtype == SCONE_OUT_FP16, "scone_embed_forward: out_dtype must be bf16 or fp16");
    SCONE_REQUIRE(B >= 0 && L >= 0, "scone_embed_forward: negative shape");
    const int64_t T = B * L;
    if (T == 0) return SCONE_OK;
    SCONE_REQUIRE(T < (1ll << 40), "scone_embed_forward: batch too large");
    SCONE_REQUIRE(d_ids && d_out, "scone_embed_forward: NULL ids or out");
    SCONE_REQUIRE(d_base_emb && base_rows > 0, "scone_embed_forward: base embedding table required (fallback rows)");
    SCONE_REQUIRE((((uintptr_t)d_base_emb | (uintptr_t)d_out | (uintptr_t)d_pos_emb) & 15) == 0,
                  "scone_embed_forward: base_emb, pos_emb and out must be 16-byte aligned");
    const scone_index_impl *ix = reinterpret_cast<const scone_index_impl *>(index);
    SCONE_REQUIRE(ix->n <= table->num_rows, "scone_embed_forward: index has %lld f-grams but the table only %lld rows",
                  (long long)ix->n, (long long)table->num_rows);
    EmbedParams p{};
    int rc = fill_table(p, table, "scone_embed_forward");
    if (rc != SCONE_OK) return rc;
    p.ix = IndexView{ix->slots, ix->cap, ix->len_mask, ix->max_n};
    p.fgram_in = nullptr;
    p.base = static_cast<const uint8_t *>(d_base_emb);
    p.V = base_rows;
    p.pos = static_cast<const uint8_t *>(d_pos_emb);
    p.ids = d_ids;
    p.T = T;
    p.L = L;
    p.out = static_cast<uint8_t *>(d_out);
    p.out_id = d_out_id;
    p.out_len = d_out_len;
    p.status = d_status;
    return dispatch(lanes_per_token(ix->max_n), p, table->quant, out_dtype, stream);
}

int scone_embed_gather(const scone_table_desc_t *table, const void *d_base_emb, int64_t base_rows, const void *d_pos_emb,
                       int64_t L, const int64_t *d_ids, const int32_t *d_fgram_id, int64_t T, void *d_out, int32_t out_dtype,
                       uint32_t *d_status, void *stream_) {
    cudaStream_t stream = (cudaStream_t)stream_;
    SCONE_REQUIRE(table, "scone_embed_gather: NULL table");
    SCONE_REQUIRE(out_dtype == SCONE_OUT_BF16 || out_dtype == SCONE_OUT_FP16, "scone_embed_gather: out_dtype must be bf16 or fp16");
    SCONE_REQUIRE(T >= 0, "scone_embed_gather: negative T");
    if (T == 0) return SCONE_OK;
    SCONE_REQUIRE(d_ids && d_out && d_fgram_id, "scone_embed_gather: NULL buffer");
    SCONE_REQUIRE(d_base_emb && base_rows > 0, "scone_embed_gather: base embedding table required (fallback rows)");
    SCONE_REQUIRE(!d_pos_emb || L > 0, "scone_embed_gather: L required with pos_emb");
    SCONE_REQUIRE((((uintptr_t)d_base_emb | (uintptr_t)d_out | (uintptr_t)d_pos_emb) & 15) == 0,
                  "scone_embed_gather: base_emb, pos_emb and out must be 16-byte aligned");
    EmbedParams p{};
    int rc = fill_table(p, table, "scone_embed_gather");
    if (rc != SCONE_OK) return rc;
    p.fgram_in = d_fgram_id;
    p.base = static_cast<const uint8_t *>(d_base_emb);
    p.V = base_rows;
    p.pos = static_cast<const uint8_t *>(d_pos_emb);
    p.ids = d_ids;
    p.T = T;
    p.L = L > 0 ? L : T;
    p.out = static_cast<uint8_t *>(d_out);
    p.status = d_status;
    return dispatch(4, p, table->quant, out_dtype, stream);
}

}  // extern "C"
