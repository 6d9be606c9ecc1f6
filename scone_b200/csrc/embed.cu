// embed.cu -- the fused hot path: longest-match lookup + cache-row gather + dequant + fallback,
// one kernel, output written once in the transformer's inputs_embeds layout.
//
// Takes over, per position, get_token_f_grams + f_gram_to_id + get_embeddings + the engine's
// assemble loop + the wte fallback of the reference (scone/tokenization/n_gram_extractor.py:106-126,
// scone/inference/embedding_cache.py:149-181, scone/inference/engine.py:235-266,
// scone/models/language_model.py:239-243), with Algorithm-2 (replace-or-fallback) semantics.
//
// Shape of the kernel (HBM-bound gather, no tensor cores):
//   * a warp owns G = 32/P consecutive positions; phase 1 resolves their f-gram ids in registers
//     (match.cuh), phase 2 streams their rows.
//   * phase 2 flattens the warp's work into items (position, 256-element step): each lane owns 8
//     consecutive elements of an item = one 16 B output vector.  U items are loaded back to back
//     before any is converted, so a warp keeps U x 32 vector loads in flight.
//   * cache rows / fallback rows are read with ld.global.nc.L1::no_allocate (touched once),
//     the output is written with st.global.cs 128-bit stores; slots use default caching so the
//     (much smaller) index stays L2-resident.
#include <cstdlib>

#include "common.cuh"
#include "match.cuh"

namespace scone {

struct EmbedParams {
    IndexView ix;
    const int32_t *fgram_in;  // not NULL: ids already resolved, skip phase 1
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
    int32_t D;
    int32_t scale_off;
    int32_t group_shift;  // log2(group / 8): chunk index >> group_shift = group index (INT4)
};

enum : int { kInactive = 0, kHit = 1, kMiss = 2, kZero = 3 };

template <int OUT>
__device__ __forceinline__ void decode16x8(uint4 raw, float (&x)[8]) {
    if (OUT == SCONE_OUT_BF16) decode_bf16x8(raw, x);
    else decode_fp16x8(raw, x);
}
template <int OUT>
__device__ __forceinline__ uint4 pack16x8(const float (&x)[8]) {
    return OUT == SCONE_OUT_BF16 ? pack_bf16x8(x) : pack_fp16x8(x);
}

template <int QUANT, int OUT, int P, int U, int MINB>
__global__ void __launch_bounds__(256, MINB) embed_kernel(const EmbedParams p) {
    constexpr int G = 32 / P;
    constexpr unsigned FULL = 0xFFFFFFFFu;
    const int lane = threadIdx.x & 31;
    const int64_t warp = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int64_t base = warp * G;
    if (base >= p.T) return;

    // ---- phase 1: which row feeds each of the G positions --------------------------------------
    const int j = lane / P;
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
    int32_t tok = -1;  // fallback row, only meaningful when fid < 0
    if (fid == -1 && i < p.T) {
        const int64_t t64 = __ldg(p.ids + i);
        if (t64 >= 0 && t64 < p.V) tok = (int32_t)t64;
    }

    // ---- phase 2: stream the rows ----------------------------------------------------------------
    const int D = p.D;
    const int nchunks = D >> 3;
    const int nsteps = (nchunks + 31) >> 5;
    const int ntok = (int)((p.T - base) < (int64_t)G ? (p.T - base) : (int64_t)G);
    const int total = ntok * nsteps;
    const bool has_pos = p.pos != nullptr;
    bool flagged = false;

    int jj = 0, s = 0;
    for (int w0 = 0; w0 < total; w0 += U) {
        uint4 raw[U];
        float sc[U];
        int kind[U];
        int coord[U];
#pragma unroll
        for (int u = 0; u < U; ++u) {
            const bool act = (w0 + u) < total;       // warp-uniform
            const int jsrc = (act ? jj : 0) * P;
            const int32_t f = __shfl_sync(FULL, fid, jsrc);
            const int32_t tk = __shfl_sync(FULL, tok, jsrc);
            const int c = (s << 5) + lane;
            kind[u] = kInactive;
            coord[u] = (jj << 20) | c;
            sc[u] = 1.0f;
            raw[u] = make_uint4(0u, 0u, 0u, 0u);
            if (act && c < nchunks) {
                if (f >= 0) {
                    kind[u] = kHit;
                    const uint8_t *r = p.rows + (int64_t)f * p.row_stride;
                    if (QUANT == SCONE_QUANT_FP16) {
                        raw[u] = ldg_stream_16(r + c * 16);
                    } else if (QUANT == SCONE_QUANT_INT8) {
                        const uint2 v = ldg_stream_8(r + c * 8);
                        raw[u].x = v.x;
                        raw[u].y = v.y;
                        sc[u] = __ldg(reinterpret_cast<const float *>(r + p.scale_off));
                    } else {
                        raw[u].x = ldg_stream_4(r + c * 4);
                        const __half hs = __ldg(reinterpret_cast<const __half *>(r + p.scale_off) + (c >> p.group_shift));
                        sc[u] = __half2float(hs);
                    }
                } else if (tk >= 0) {
                    kind[u] = kMiss;
                    raw[u] = ldg_stream_16(p.base + ((int64_t)tk * D + c * 8) * 2);
                } else {
                    kind[u] = kZero;
                }
            }
            if (act && ++s == nsteps) {
                s = 0;
                ++jj;
            }
        }
#pragma unroll
        for (int u = 0; u < U; ++u) {
            if (kind[u] == kInactive) continue;
            const int tj = coord[u] >> 20, c = coord[u] & 0xFFFFF;
            const int64_t t = base + tj;
            uint4 o;
            if (kind[u] == kMiss && !has_pos) {
                o = raw[u];  // fallback rows are already in the output type
            } else if (kind[u] == kZero && !has_pos) {
                o = make_uint4(0u, 0u, 0u, 0u);
                flagged = true;
            } else {
                float x[8];
                if (kind[u] == kHit) {
                    if (QUANT == SCONE_QUANT_FP16) decode_fp16x8(raw[u], x);
                    else if (QUANT == SCONE_QUANT_INT8) decode_int8x8(make_uint2(raw[u].x, raw[u].y), sc[u], x);
                    else decode_int4x8(raw[u].x, sc[u], x);
                } else if (kind[u] == kMiss) {
                    decode16x8<OUT>(raw[u], x);
                } else {
#pragma unroll
                    for (int k = 0; k < 8; ++k) x[k] = 0.0f;
                    flagged = true;
                }
                if (has_pos) {
                    const int64_t pr = pos_in_row(t, p.L, p.T);
                    const uint4 pv = __ldg(reinterpret_cast<const uint4 *>(p.pos + (pr * D + c * 8) * 2));
                    float y[8];
                    decode16x8<OUT>(pv, y);
#pragma unroll
                    for (int k = 0; k < 8; ++k) x[k] = __fadd_rn(x[k], y[k]);
                }
                if (QUANT == SCONE_QUANT_FP16 && OUT == SCONE_OUT_FP16 && kind[u] == kHit && !has_pos) o = raw[u];
                else o = pack16x8<OUT>(x);
            }
            stg_stream_16(p.out + (t * D + c * 8) * 2, o);
        }
    }
    if (flagged && p.status) atomicOr(p.status, SCONE_STATUS_TOKEN_OOR);
}

static int lanes_per_token(int max_n) { return max_n <= 1 ? 1 : max_n <= 2 ? 2 : max_n <= 4 ? 4 : 8; }

// Tuning hook (tools/tune_embed.py): SCONE_EMBED_VARIANT="U:MINB" selects another instantiation of the same
// kernel for the combinations compiled below; anything else runs the default.
static void variant(int &u, int &minb) {
    u = 0;
    minb = 0;
    if (const char *e = getenv("SCONE_EMBED_VARIANT")) sscanf(e, "%d:%d", &u, &minb);
}

template <int QUANT, int OUT, int P>
static void launch(const EmbedParams &p, cudaStream_t stream) {
    constexpr int G = 32 / P;
    const int64_t windows = (p.T + G - 1) / G;
    const unsigned blocks = (unsigned)((windows + 7) / 8);
    if constexpr (OUT == SCONE_OUT_BF16 && (P == 4 || P == 8)) {
        int u, minb;
        variant(u, minb);
#define SCONE_V(UU, MM)                                                                  \
    if (u == UU && minb == MM) {                                                         \
        embed_kernel<QUANT, OUT, P, UU, MM><<<blocks, 256, 0, stream>>>(p);              \
        return;                                                                          \
    }
        SCONE_V(4, 3) SCONE_V(4, 6) SCONE_V(4, 8) SCONE_V(8, 2) SCONE_V(8, 3) SCONE_V(8, 4) SCONE_V(2, 8) SCONE_V(16, 2)
#undef SCONE_V
    }
    embed_kernel<QUANT, OUT, P, 4, 4><<<blocks, 256, 0, stream>>>(p);
}

template <int QUANT, int OUT>
static void launch_p(int P, const EmbedParams &p, cudaStream_t stream) {
    switch (P) {
        case 1: launch<QUANT, OUT, 1>(p, stream); break;
        case 2: launch<QUANT, OUT, 2>(p, stream); break;
        case 4: launch<QUANT, OUT, 4>(p, stream); break;
        default: launch<QUANT, OUT, 8>(p, stream); break;
    }
}

static int dispatch(int P, const EmbedParams &p, int quant, int out_dtype, cudaStream_t stream) {
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
