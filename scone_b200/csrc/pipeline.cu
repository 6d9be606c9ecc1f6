// pipeline.cu -- host-fed, multi-slot pipeline around the fused kernel (native runtime code: streams, events, copies).
//
// Replaces the reference engine's per-position host loop + per-position pageable H2D copy
// (scone/inference/engine.py:235-259, scone/inference/embedding_cache.py:143-145) for callers whose ids are on the host.
#include <vector>

#include "common.cuh"

namespace scone {

struct Pipeline {
    const scone_index_t *index;
    scone_table_desc_t table;
    const void *base;
    int64_t base_rows;
    const void *pos;
    int64_t B, L;
    int32_t out_dtype;
    int32_t slots;
    std::vector<void *> d_ids, d_out, d_meta, h_meta;
    uint32_t *status;
    // two compute streams, used alternately: consecutive batches write different slots, so their kernels are independent and
    // batch k+1's CTAs move onto the SMs as batch k's retire (one stream would serialise them: the event record / waits between
    // two launches of a stream also end the programmatic-dependent-launch overlap, and each batch would cost an isolated launch)
    cudaStream_t s_in = nullptr, s_run[2] = {nullptr, nullptr}, s_out = nullptr;
    std::vector<cudaEvent_t> ev_in, ev_run, ev_out;
    std::vector<char> busy;
    int64_t k = 0;
};

static void pipeline_free(Pipeline *p) {
    for (auto e : p->ev_in) cudaEventDestroy(e);
    for (auto e : p->ev_run) cudaEventDestroy(e);
    for (auto e : p->ev_out) cudaEventDestroy(e);
    if (p->s_in) cudaStreamDestroy(p->s_in);
    for (int i = 0; i < 2; ++i)
        if (p->s_run[i]) cudaStreamDestroy(p->s_run[i]);
    if (p->s_out) cudaStreamDestroy(p->s_out);
    delete p;
}

}  // namespace scone

using namespace scone;

extern "C" {

int scone_pipeline_create(const scone_index_t *index, const scone_table_desc_t *table, const void *d_base_emb, int64_t base_rows,
                          const void *d_pos_emb, int64_t B, int64_t L, int32_t out_dtype, int32_t slots, void *const *d_ids_slots,
                          void *const *d_out_slots, void *const *d_meta_slots, void *const *h_meta_slots, uint32_t *d_status,
                          scone_pipeline_t **out) {
    SCONE_REQUIRE(out, "scone_pipeline_create: out is NULL");
    *out = nullptr;
    SCONE_REQUIRE(index && table && d_base_emb, "scone_pipeline_create: NULL index, table or base embedding");
    SCONE_REQUIRE(B > 0 && L > 0, "scone_pipeline_create: batch shape must be positive");
    SCONE_REQUIRE(slots >= 1 && slots <= 16, "scone_pipeline_create: slots %d outside [1, 16]", slots);
    SCONE_REQUIRE(d_ids_slots && d_out_slots && d_meta_slots && h_meta_slots, "scone_pipeline_create: NULL slot buffer table");
    Pipeline *p = new (std::nothrow) Pipeline();
    if (!p) {
        set_error("scone_pipeline_create: out of host memory");
        return SCONE_E_NOMEM;
    }
    p->index = index;
    p->table = *table;
    p->base = d_base_emb;
    p->base_rows = base_rows;
    p->pos = d_pos_emb;
    p->B = B;
    p->L = L;
    p->out_dtype = out_dtype;
    p->slots = slots;
    p->status = d_status;
    for (int s = 0; s < slots; ++s) {
        if (!d_ids_slots[s] || !d_out_slots[s] || !d_meta_slots[s] || !h_meta_slots[s] || ((uintptr_t)d_meta_slots[s] & 3)) {
            set_error("scone_pipeline_create: slot %d has a NULL or misaligned buffer", s);
            delete p;
            return SCONE_E_INVALID;
        }
        p->d_ids.push_back(d_ids_slots[s]);
        p->d_out.push_back(d_out_slots[s]);
        p->d_meta.push_back(d_meta_slots[s]);
        p->h_meta.push_back(h_meta_slots[s]);
    }
    p->busy.assign(slots, 0);
    auto fail = [&](cudaError_t e, const char *what) {
        set_error("scone_pipeline_create: %s failed: %s", what, cudaGetErrorString(e));
        pipeline_free(p);
        return SCONE_E_CUDA;
    };
    cudaError_t e;
    if ((e = cudaStreamCreateWithFlags(&p->s_in, cudaStreamNonBlocking)) != cudaSuccess) return fail(e, "cudaStreamCreate");
    for (int i = 0; i < 2; ++i)
        if ((e = cudaStreamCreateWithFlags(&p->s_run[i], cudaStreamNonBlocking)) != cudaSuccess) return fail(e, "cudaStreamCreate");
    if ((e = cudaStreamCreateWithFlags(&p->s_out, cudaStreamNonBlocking)) != cudaSuccess) return fail(e, "cudaStreamCreate");
    for (int s = 0; s < slots; ++s) {
        cudaEvent_t a, b, c;
        if ((e = cudaEventCreateWithFlags(&a, cudaEventDisableTiming)) != cudaSuccess) return fail(e, "cudaEventCreate");
        p->ev_in.push_back(a);
        if ((e = cudaEventCreateWithFlags(&b, cudaEventDisableTiming)) != cudaSuccess) return fail(e, "cudaEventCreate");
        p->ev_run.push_back(b);
        if ((e = cudaEventCreateWithFlags(&c, cudaEventDisableTiming)) != cudaSuccess) return fail(e, "cudaEventCreate");
        p->ev_out.push_back(c);
    }
    *out = reinterpret_cast<scone_pipeline_t *>(p);
    return SCONE_OK;
}

int scone_pipeline_submit(scone_pipeline_t *pp, const int64_t *h_ids_pinned, int32_t *slot_out) {
    SCONE_REQUIRE(pp && h_ids_pinned, "scone_pipeline_submit: NULL argument");
    Pipeline *p = reinterpret_cast<Pipeline *>(pp);
    const int s = (int)(p->k % p->slots);
    const int64_t T = p->B * p->L;
    if (p->busy[s]) SCONE_CUDA(cudaEventSynchronize(p->ev_out[s]));  // caller never collected it: do not overwrite live results
    // copy-in: after the previous kernel that read this slot's ids
    SCONE_CUDA(cudaStreamWaitEvent(p->s_in, p->ev_run[s], 0));
    SCONE_CUDA(cudaMemcpyAsync(p->d_ids[s], h_ids_pinned, (size_t)T * 8, cudaMemcpyHostToDevice, p->s_in));
    SCONE_CUDA(cudaEventRecord(p->ev_in[s], p->s_in));
    // compute: after the ids are in and the slot's previous results have left the device
    cudaStream_t run = p->s_run[p->k & 1];
    SCONE_CUDA(cudaStreamWaitEvent(run, p->ev_in[s], 0));
    SCONE_CUDA(cudaStreamWaitEvent(run, p->ev_out[s], 0));
    uint8_t *meta = static_cast<uint8_t *>(p->d_meta[s]);
    // The kernel that precedes this one on its stream is the pipeline's own batch k-2, which writes only another slot's
    // outputs; this batch's ids arrive by the copy above (an event dependency, not a kernel) and the tables are static
    // while the pipeline runs (scone_pipeline_follow orders it behind a caller's update): SCONE_EMBED_INPUTS_STABLE holds.
    scone_embed_opts_t opts{};
    opts.flags = SCONE_EMBED_INPUTS_STABLE;
    int rc = scone_embed_forward_ex(p->index, &p->table, p->base, p->base_rows, p->pos, static_cast<const int64_t *>(p->d_ids[s]), p->B,
                                    p->L, p->d_out[s], p->out_dtype, reinterpret_cast<int32_t *>(meta), meta + 4 * T, p->status, &opts, run);
    if (rc != SCONE_OK) return rc;
    SCONE_CUDA(cudaEventRecord(p->ev_run[s], run));
    // copy-out
    SCONE_CUDA(cudaStreamWaitEvent(p->s_out, p->ev_run[s], 0));
    SCONE_CUDA(cudaMemcpyAsync(p->h_meta[s], p->d_meta[s], (size_t)T * 5, cudaMemcpyDeviceToHost, p->s_out));
    SCONE_CUDA(cudaEventRecord(p->ev_out[s], p->s_out));
    p->busy[s] = 1;
    p->k += 1;
    if (slot_out) *slot_out = s;
    return SCONE_OK;
}

int scone_pipeline_follow(scone_pipeline_t *pp, void *stream) {
    SCONE_REQUIRE(pp, "scone_pipeline_follow: NULL pipeline");
    Pipeline *p = reinterpret_cast<Pipeline *>(pp);
    cudaEvent_t ev;
    SCONE_CUDA(cudaEventCreateWithFlags(&ev, cudaEventDisableTiming));
    cudaError_t e = cudaEventRecord(ev, (cudaStream_t)stream);
    if (e == cudaSuccess) e = cudaStreamWaitEvent(p->s_in, ev, 0);
    if (e == cudaSuccess) e = cudaStreamWaitEvent(p->s_run[0], ev, 0);
    if (e == cudaSuccess) e = cudaStreamWaitEvent(p->s_run[1], ev, 0);
    if (e == cudaSuccess) e = cudaStreamWaitEvent(p->s_out, ev, 0);
    cudaEventDestroy(ev);  // released once the recorded work has completed
    if (e != cudaSuccess) {
        set_error("scone_pipeline_follow failed: %s", cudaGetErrorString(e));
        return SCONE_E_CUDA;
    }
    return SCONE_OK;
}

int scone_pipeline_wait(scone_pipeline_t *pp, int32_t slot) {
    SCONE_REQUIRE(pp, "scone_pipeline_wait: NULL pipeline");
    Pipeline *p = reinterpret_cast<Pipeline *>(pp);
    SCONE_REQUIRE(slot >= 0 && slot < p->slots, "scone_pipeline_wait: slot %d outside [0, %d)", slot, p->slots);
    if (!p->busy[slot]) return SCONE_OK;
    SCONE_CUDA(cudaEventSynchronize(p->ev_out[slot]));
    p->busy[slot] = 0;
    return SCONE_OK;
}

int scone_pipeline_destroy(scone_pipeline_t *pp) {
    if (!pp) return SCONE_OK;
    Pipeline *p = reinterpret_cast<Pipeline *>(pp);
    cudaStreamSynchronize(p->s_in);
    cudaStreamSynchronize(p->s_run[0]);
    cudaStreamSynchronize(p->s_run[1]);
    cudaStreamSynchronize(p->s_out);
    pipeline_free(p);
    return SCONE_OK;
}

}  // extern "C"
