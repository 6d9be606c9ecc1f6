/*
 * c_oracle.c -- plain-C restatement of the reference's input-embedding lookup
 * over flat arrays.  TEST INFRASTRUCTURE ONLY (see oracle/__init__.py): the
 * checker for sizes where the Python oracle's ~360 B/f-gram dictionaries do
 * not fit, and a timed CPU baseline.  Never linked or loaded by scone_b200.
 *
 * It is validated bit-exactly against oracle/py_oracle.py (itself pinned to
 * fixtures generated from the unmodified reference) in tests/test_oracle.py.
 * Parity status is the same as py_oracle.py: match / fp32 gather / fp16 cast
 * pinned; INT8 / INT4 row dequant "parity unpinned" (formulas defined by us).
 *
 * Reference lines restated (llmsresearch/scone):
 *   membership test of an n-gram tuple  scone/tokenization/n_gram_extractor.py:121-122
 *   tuple -> id                          scone/inference/embedding_cache.py:173
 *   table[ids] gather                    scone/inference/embedding_cache.py:127-135
 *   .half() cast                         scone/inference/engine.py:265-266
 *   fallback row wte(ids)                scone/models/language_model.py:239
 *   longest f-gram ending at i           assets/algorithm.png (Algorithm 2)
 *
 * Deliberately NOT the GPU design: FNV-1a hash, id-only slots, keys compared
 * through the id-indexed token array -- an independent implementation.
 */
#include <pthread.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

typedef struct {
    int32_t max_n;
    int64_t n;
    uint64_t cap;        /* power of two */
    int32_t *slot_id;    /* cap entries, -1 = empty */
    int32_t *toks;       /* n * max_n, copy */
    uint8_t *lens;       /* n */
    uint32_t len_mask;   /* bit (len-1) set if some f-gram has that length */
} oracle_index;

static uint64_t fnv1a(const int32_t *t, int len) {
    uint64_t h = 1469598103934665603ULL;
    h = (h ^ (uint64_t)len) * 1099511628211ULL;
    for (int k = 0; k < len; ++k) {
        uint32_t v = (uint32_t)t[k];
        for (int b = 0; b < 4; ++b) {
            h = (h ^ ((v >> (8 * b)) & 0xFF)) * 1099511628211ULL;
        }
    }
    h ^= h >> 29;
    return h;
}

/* id of the f-gram equal to t[0..len), or -1 */
static int32_t find(const oracle_index *ix, const int32_t *t, int len) {
    uint64_t m = ix->cap - 1, s = fnv1a(t, len) & m;
    for (;;) {
        int32_t id = ix->slot_id[s];
        if (id < 0) return -1;
        if (ix->lens[id] == len && memcmp(ix->toks + (int64_t)id * ix->max_n, t, (size_t)len * 4) == 0) return id;
        s = (s + 1) & m;
    }
}

void oracle_index_destroy(void *p) {
    oracle_index *ix = (oracle_index *)p;
    if (!ix) return;
    free(ix->slot_id); free(ix->toks); free(ix->lens); free(ix);
}

/* toks: int32 [n, max_n] padded with -1; lens: uint8 [n]; id = row number.
 * Returns NULL on allocation failure, invalid length, or a duplicated key. */
void *oracle_index_create(const int32_t *toks, const uint8_t *lens, int64_t n, int32_t max_n) {
    oracle_index *ix = (oracle_index *)calloc(1, sizeof(oracle_index));
    if (!ix) return NULL;
    ix->max_n = max_n; ix->n = n;
    uint64_t cap = 16;
    while (cap < (uint64_t)n * 2) cap <<= 1;
    ix->cap = cap;
    ix->slot_id = (int32_t *)malloc(cap * sizeof(int32_t));
    ix->toks = (int32_t *)malloc((size_t)(n > 0 ? n : 1) * max_n * sizeof(int32_t));
    ix->lens = (uint8_t *)malloc((size_t)(n > 0 ? n : 1));
    if (!ix->slot_id || !ix->toks || !ix->lens) { oracle_index_destroy(ix); return NULL; }
    memset(ix->slot_id, 0xFF, cap * sizeof(int32_t));
    memcpy(ix->toks, toks, (size_t)n * max_n * sizeof(int32_t));
    memcpy(ix->lens, lens, (size_t)n);
    for (int64_t i = 0; i < n; ++i) {
        int len = lens[i];
        if (len < 1 || len > max_n) { oracle_index_destroy(ix); return NULL; }
        const int32_t *t = ix->toks + i * max_n;
        if (find(ix, t, len) >= 0) { oracle_index_destroy(ix); return NULL; }
        uint64_t m = cap - 1, s = fnv1a(t, len) & m;
        while (ix->slot_id[s] >= 0) s = (s + 1) & m;
        ix->slot_id[s] = (int32_t)i;
        ix->len_mask |= 1u << (len - 1);
    }
    return ix;
}

/* ---- 16-bit casts -------------------------------------------------------- */
static uint16_t f32_to_bf16(float f) {
    uint32_t u; memcpy(&u, &f, 4);
    if ((u & 0x7FFFFFFFu) > 0x7F800000u) return (uint16_t)((u >> 16) | 0x0040);
    return (uint16_t)((u + 0x7FFFu + ((u >> 16) & 1u)) >> 16);
}

static uint16_t f32_to_f16(float f) { /* IEEE RNE, handles subnormals/inf/nan */
    uint32_t u; memcpy(&u, &f, 4);
    uint32_t sign = (u >> 16) & 0x8000u, a = u & 0x7FFFFFFFu;
    if (a > 0x7F800000u) return (uint16_t)(sign | 0x7E00u | ((a >> 13) & 0x1FFu)); /* nan */
    if (a >= 0x47800000u) return (uint16_t)(sign | 0x7C00u);                        /* >= 65536 -> inf (65520.. handled below) */
    if (a < 0x33000001u) return (uint16_t)sign;                                     /* <= 2^-25 -> 0 (tie to even) */
    int32_t e = (int32_t)(a >> 23) - 127;
    uint32_t mant = (a & 0x7FFFFFu) | 0x800000u;
    int shift; uint32_t base;
    if (e < -14) { shift = 13 + (-14 - e); base = 0; }            /* subnormal result */
    else { shift = 13; base = (uint32_t)(e + 15) << 10; mant &= 0x7FFFFFu; }
    uint32_t q = mant >> shift, rem = mant & ((1u << shift) - 1), half = 1u << (shift - 1);
    uint32_t r = base + q;
    if (rem > half || (rem == half && (r & 1u))) r += 1;          /* carries propagate into exponent / inf */
    return (uint16_t)(sign | r);
}

static float f16_to_f32(uint16_t h) {
    uint32_t sign = ((uint32_t)h & 0x8000u) << 16, e = (h >> 10) & 0x1F, m = h & 0x3FF, u;
    if (e == 0) {
        if (m == 0) u = sign;
        else { int sh = 0; while (!(m & 0x400)) { m <<= 1; ++sh; } u = sign | ((uint32_t)(113 - sh) << 23) | ((m & 0x3FF) << 13); }
    } else if (e == 31) u = sign | 0x7F800000u | (m << 13);
    else u = sign | ((e + 112) << 23) | (m << 13);
    float f; memcpy(&f, &u, 4); return f;
}

/* exported for the tests that pin these helpers against numpy */
uint16_t oracle_f32_to_f16(float f) { return f32_to_f16(f); }
uint16_t oracle_f32_to_bf16(float f) { return f32_to_bf16(f); }
float oracle_f16_to_f32(uint16_t h) { return f16_to_f32(h); }

/* ---- the path ------------------------------------------------------------ */
enum { Q_FP16 = 0, Q_INT8 = 1, Q_INT4 = 2, Q_FP32 = 3 };  /* Q_FP32: unquantised rows, embedding_cache.py:84-91,132-135 */
enum { OUT_BF16 = 0, OUT_FP16 = 1 };

typedef struct {
    const oracle_index *ix;
    const int64_t *ids; int64_t B, L;
    int32_t *out_id; uint8_t *out_len; int32_t *out_all;
    /* embed */
    int do_embed, quant, D, group, out_dtype;
    const uint8_t *payload; int64_t row_stride;
    const uint8_t *scales; int64_t scale_stride;
    const uint16_t *base; int64_t V;
    const uint16_t *pos;  /* optional [>= L, D] in out_dtype: the wpe term (language_model.py:253-254) */
    int additive;         /* reference-code combine: wte row + f-gram row (language_model.py:239-243) */
    uint16_t *out;
    int64_t t0, t1; int err;
} job;

static void longest_at(const oracle_index *ix, const int64_t *row, int64_t i, int32_t *oid, uint8_t *olen) {
    int32_t t[16];
    int nmax = ix->max_n < i + 1 ? ix->max_n : (int)(i + 1);
    *oid = -1; *olen = 0;
    for (int n = nmax; n >= 1; --n) {
        if (!(ix->len_mask >> (n - 1) & 1u)) continue;
        int ok = 1;
        for (int k = 0; k < n; ++k) {
            int64_t v = row[i - n + 1 + k];
            if (v < 0 || v > 0x7FFFFFFFLL) { ok = 0; break; }
            t[k] = (int32_t)v;
        }
        if (!ok) continue;
        int32_t id = find(ix, t, n);
        if (id >= 0) { *oid = id; *olen = (uint8_t)n; return; }
    }
}

static float widen16(const job *j, uint16_t b) {
    if (j->out_dtype == OUT_BF16) { uint32_t u = (uint32_t)b << 16; float f; memcpy(&f, &u, 4); return f; }
    return f16_to_f32(b);
}

static float table_elem(const job *j, const uint8_t *p, const uint8_t *sp, int d) {
    if (j->quant == Q_FP32) { float f; memcpy(&f, p + 4 * d, 4); return f; }
    if (j->quant == Q_FP16) { uint16_t h; memcpy(&h, p + 2 * d, 2); return f16_to_f32(h); }
    if (j->quant == Q_INT8) { float s; memcpy(&s, sp, 4); return (float)(int8_t)p[d] * s; }
    uint16_t h; memcpy(&h, sp + 2 * (d / j->group), 2);
    int q = (int)((p[d >> 1] >> ((d & 1) * 4)) & 0xF) - 8;
    return (float)q * f16_to_f32(h);
}

/* position add and/or additive combine: fp32 adds in the reference's order (wte + row) + wpe, ONE rounding
 * (oracle/py_oracle.py embed_forward with pos_emb_bits / additive). */
static void emit_row_general(const job *j, int32_t fid, int64_t tok, int64_t i, uint16_t *o, int *err) {
    const int D = j->D;
    const int tok_ok = tok >= 0 && tok < j->V;
    const uint8_t *p = fid >= 0 ? j->payload + (int64_t)fid * j->row_stride : NULL;
    const uint8_t *sp = (fid >= 0 && j->scales) ? j->scales + (int64_t)fid * j->scale_stride : NULL;
    if (!tok_ok && (fid < 0 || j->additive)) *err = 1;
    for (int d = 0; d < D; ++d) {
        float x;
        if (fid >= 0) {
            x = table_elem(j, p, sp, d);
            if (j->additive && tok_ok) x = widen16(j, j->base[tok * D + d]) + x;
        } else {
            x = tok_ok ? widen16(j, j->base[tok * D + d]) : 0.0f;
        }
        if (j->pos) x = x + widen16(j, j->pos[i * D + d]);
        o[d] = j->out_dtype == OUT_BF16 ? f32_to_bf16(x) : f32_to_f16(x);
    }
}

static void emit_row(const job *j, int32_t fid, int64_t tok, uint16_t *o, int *err) {
    const int D = j->D;
    if (fid < 0) {
        if (tok < 0 || tok >= j->V) { memset(o, 0, (size_t)D * 2); *err = 1; return; }
        memcpy(o, j->base + tok * D, (size_t)D * 2);
        return;
    }
    const uint8_t *p = j->payload + (int64_t)fid * j->row_stride;
    const uint8_t *sp = j->scales ? j->scales + (int64_t)fid * j->scale_stride : NULL;
    for (int d = 0; d < D; ++d) {
        float x;
        if (j->quant == Q_FP32) memcpy(&x, p + 4 * d, 4);
        else if (j->quant == Q_FP16) { uint16_t h; memcpy(&h, p + 2 * d, 2); x = f16_to_f32(h); }
        else if (j->quant == Q_INT8) { float s; memcpy(&s, sp, 4); x = (float)(int8_t)p[d] * s; }
        else { uint16_t h; memcpy(&h, sp + 2 * (d / j->group), 2);
               int q = (int)((p[d >> 1] >> ((d & 1) * 4)) & 0xF) - 8; x = (float)q * f16_to_f32(h); }
        o[d] = j->out_dtype == OUT_BF16 ? f32_to_bf16(x) : f32_to_f16(x);
    }
}

static void *worker(void *arg) {
    job *j = (job *)arg;
    const oracle_index *ix = j->ix;
    for (int64_t t = j->t0; t < j->t1; ++t) {
        int64_t b = t / j->L, i = t % j->L;
        const int64_t *row = j->ids + b * j->L;
        if (j->out_all) {
            int32_t tk[16];
            for (int n = 1; n <= ix->max_n; ++n) {
                int32_t id = -1;
                if (n <= i + 1) {
                    int ok = 1;
                    for (int k = 0; k < n; ++k) { int64_t v = row[i - n + 1 + k]; if (v < 0 || v > 0x7FFFFFFFLL) { ok = 0; break; } tk[k] = (int32_t)v; }
                    if (ok) id = find(ix, tk, n);
                }
                j->out_all[t * ix->max_n + (n - 1)] = id;
            }
            continue;
        }
        int32_t fid; uint8_t fl;
        longest_at(ix, row, i, &fid, &fl);
        if (j->out_id) j->out_id[t] = fid;
        if (j->out_len) j->out_len[t] = fl;
        if (j->do_embed) {
            if (j->pos || j->additive) emit_row_general(j, fid, row[i], i, j->out + t * j->D, &j->err);
            else emit_row(j, fid, row[i], j->out + t * j->D, &j->err);
        }
    }
    return NULL;
}

static int run(job *proto, int nthreads) {
    int64_t T = proto->B * proto->L;
    if (nthreads < 1) nthreads = 1;
    if (nthreads > 256) nthreads = 256;
    if ((int64_t)nthreads > T) nthreads = T > 0 ? (int)T : 1;
    pthread_t th[256]; job jobs[256];
    int64_t per = (T + nthreads - 1) / nthreads;
    for (int k = 0; k < nthreads; ++k) {
        jobs[k] = *proto;
        jobs[k].t0 = k * per; jobs[k].t1 = (k + 1) * per < T ? (k + 1) * per : T;
        if (jobs[k].t0 > T) jobs[k].t0 = T;
        jobs[k].err = 0;
        if (nthreads == 1) worker(&jobs[k]); else pthread_create(&th[k], NULL, worker, &jobs[k]);
    }
    int err = 0;
    for (int k = 0; k < nthreads; ++k) { if (nthreads > 1) pthread_join(th[k], NULL); err |= jobs[k].err; }
    return err;
}

void oracle_match(const void *ix, const int64_t *ids, int64_t B, int64_t L, int32_t *out_id, uint8_t *out_len, int nthreads) {
    job j; memset(&j, 0, sizeof j);
    j.ix = (const oracle_index *)ix; j.ids = ids; j.B = B; j.L = L; j.out_id = out_id; j.out_len = out_len;
    run(&j, nthreads);
}

void oracle_match_all(const void *ix, const int64_t *ids, int64_t B, int64_t L, int32_t *out_all, int nthreads) {
    job j; memset(&j, 0, sizeof j);
    j.ix = (const oracle_index *)ix; j.ids = ids; j.B = B; j.L = L; j.out_all = out_all;
    run(&j, nthreads);
}

/* Algorithm 2 end to end.  Returns 0, or 1 if some missed token id was outside [0, V). */
int oracle_embed(const void *ix, int quant, int D, int group,
                 const uint8_t *payload, int64_t row_stride, const uint8_t *scales, int64_t scale_stride,
                 const uint16_t *base, int64_t V, const int64_t *ids, int64_t B, int64_t L,
                 int out_dtype, uint16_t *out, int32_t *out_id, uint8_t *out_len, int nthreads) {
    job j; memset(&j, 0, sizeof j);
    j.ix = (const oracle_index *)ix; j.ids = ids; j.B = B; j.L = L; j.out_id = out_id; j.out_len = out_len;
    j.do_embed = 1; j.quant = quant; j.D = D; j.group = group; j.out_dtype = out_dtype;
    j.payload = payload; j.row_stride = row_stride; j.scales = scales; j.scale_stride = scale_stride;
    j.base = base; j.V = V; j.out = out;
    return run(&j, nthreads);
}

/* Same with the optional fused position add (pos: [>= L, D] in out_dtype, or NULL) and the additive combine. */
int oracle_embed_ex(const void *ix, int quant, int D, int group,
                    const uint8_t *payload, int64_t row_stride, const uint8_t *scales, int64_t scale_stride,
                    const uint16_t *base, int64_t V, const uint16_t *pos, int additive, const int64_t *ids, int64_t B, int64_t L,
                    int out_dtype, uint16_t *out, int32_t *out_id, uint8_t *out_len, int nthreads) {
    job j; memset(&j, 0, sizeof j);
    j.ix = (const oracle_index *)ix; j.ids = ids; j.B = B; j.L = L; j.out_id = out_id; j.out_len = out_len;
    j.do_embed = 1; j.quant = quant; j.D = D; j.group = group; j.out_dtype = out_dtype;
    j.payload = payload; j.row_stride = row_stride; j.scales = scales; j.scale_stride = scale_stride;
    j.base = base; j.V = V; j.pos = pos; j.additive = additive; j.out = out;
    return run(&j, nthreads);
}
