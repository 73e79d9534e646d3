// Ensemble generation driver: "dates x ensemble members" through the graph-replayed sampler.
// Replaces the loop around DiffusionUtils.sample in the reference's generation scripts
// (DDPM_DANRA_conditional/generation_DANRA_conditional.py:369-441, DDPM_clean_application/test/generation_ddpm.py:371-439):
// there ONE dataloader batch is moved to the device, sampled, and moved back; an ensemble needs the same conditioning fields
// for every member of a date, so the driver takes the per-DATE fields once (host memory) and schedules the flattened
// (date-major) member list in sub-batches:
//   host gather of the sub-batch's conditioning rows -> pinned staging [2]  -> H2D on a copy stream -> device staging [2]
//   compute stream: x_T drawn on the device (Philox keyed by global member index) -> conditioning pre-pass -> T-1 graph-replayed
//   reverse steps -> fields parked in an output stage [2] -> D2H on the copy stream straight into the caller's array
// so the uploads of sub-batch k+1 and the download of sub-batch k run under the reverse loop of the neighbouring sub-batch.
// Included inside the extern "C" block of b200ddpm.cu (it uses its static helpers).

}  // extern "C"
namespace b2d {
struct EnsembleRes {
    int sub_batch = 0;
    size_t cond_elems = 0;
    float *d_lsm[2] = {nullptr, nullptr}, *d_topo[2] = {nullptr, nullptr}, *d_cond[2] = {nullptr, nullptr}, *d_out[2] = {nullptr, nullptr};
    int* d_y[2] = {nullptr, nullptr};
    float *h_lsm[2] = {nullptr, nullptr}, *h_topo[2] = {nullptr, nullptr}, *h_cond[2] = {nullptr, nullptr}, *h_out[2] = {nullptr, nullptr};
    int* h_y[2] = {nullptr, nullptr};
    cudaStream_t copy = nullptr;
    cudaEvent_t ev_h2d[2] = {nullptr, nullptr}, ev_done[2] = {nullptr, nullptr}, ev_d2h[2] = {nullptr, nullptr};
};
void ensemble_release(EnsembleRes* r) {
    if (!r) return;
    for (int s = 0; s < 2; ++s) {
        cudaFree(r->d_lsm[s]); cudaFree(r->d_topo[s]); cudaFree(r->d_cond[s]); cudaFree(r->d_out[s]); cudaFree(r->d_y[s]);
        cudaFreeHost(r->h_lsm[s]); cudaFreeHost(r->h_topo[s]); cudaFreeHost(r->h_cond[s]); cudaFreeHost(r->h_out[s]); cudaFreeHost(r->h_y[s]);
        if (r->ev_h2d[s]) cudaEventDestroy(r->ev_h2d[s]);
        if (r->ev_done[s]) cudaEventDestroy(r->ev_done[s]);
        if (r->ev_d2h[s]) cudaEventDestroy(r->ev_d2h[s]);
    }
    if (r->copy) cudaStreamDestroy(r->copy);
    delete r;
}
static int ensemble_resources(b2d_handle* h, int sub_batch, size_t cond_row_elems) {
    const b2d_config& c = h->cfg;
    const size_t plane = (size_t)c.img_size * c.img_size;
    const size_t cond_elems = (size_t)sub_batch * std::max<size_t>(cond_row_elems, 1);
    if (h->ens && h->ens->sub_batch >= sub_batch && h->ens->cond_elems >= cond_elems) return 0;
    B2D_CUDA(cudaDeviceSynchronize());
    ensemble_release(h->ens);
    h->ens = new EnsembleRes();
    EnsembleRes* r = h->ens;
    r->sub_batch = sub_batch;
    r->cond_elems = cond_elems;
    const size_t nx = (size_t)sub_batch * c.c_hr * plane;
    for (int s = 0; s < 2; ++s) {
        B2D_CUDA(cudaMalloc(&r->d_lsm[s], (size_t)sub_batch * plane * 4));
        B2D_CUDA(cudaMalloc(&r->d_topo[s], (size_t)sub_batch * plane * 4));
        B2D_CUDA(cudaMalloc(&r->d_cond[s], cond_elems * 4));
        B2D_CUDA(cudaMalloc(&r->d_out[s], nx * 4));
        B2D_CUDA(cudaMalloc(&r->d_y[s], (size_t)sub_batch * 4));
        B2D_CUDA(cudaMallocHost(&r->h_lsm[s], (size_t)sub_batch * plane * 4));
        B2D_CUDA(cudaMallocHost(&r->h_topo[s], (size_t)sub_batch * plane * 4));
        B2D_CUDA(cudaMallocHost(&r->h_cond[s], cond_elems * 4));
        B2D_CUDA(cudaMallocHost(&r->h_out[s], nx * 4));
        B2D_CUDA(cudaMallocHost(&r->h_y[s], (size_t)sub_batch * 4));
        B2D_CUDA(cudaEventCreateWithFlags(&r->ev_h2d[s], cudaEventDisableTiming));
        B2D_CUDA(cudaEventCreateWithFlags(&r->ev_done[s], cudaEventDisableTiming));
        B2D_CUDA(cudaEventCreateWithFlags(&r->ev_d2h[s], cudaEventDisableTiming));
    }
    B2D_CUDA(cudaStreamCreateWithFlags(&r->copy, cudaStreamNonBlocking));
    return 0;
}
}  // namespace b2d
extern "C" {

int b2d_ensemble_run(b2d_handle* h, const b2d_ensemble_job* job, b2d_ensemble_stats* stats) {
    B2D_CHECK(h && job && (job->out || job->count == 0), "null argument");
    B2D_CHECK(h->T >= 2, "b2d_set_schedule has not been called");
    const b2d_config& c = h->cfg;
    B2D_CHECK(job->n_dates >= 1 && job->members >= 1 && job->sub_batch >= 1, "bad ensemble shape");
    B2D_CHECK(job->sub_batch <= c.max_batch, "sub_batch exceeds max_batch of the handle");
    const long long total = (long long)job->n_dates * job->members;
    B2D_CHECK(job->first >= 0 && job->count >= 0 && (long long)job->first + job->count <= total, "slice outside the member list");
    const int H = c.img_size;
    const size_t plane = (size_t)H * H, per_sample = (size_t)c.c_hr * plane;
    const bool is_d = c.family == B2D_FAMILY_D;
    const bool use_lsm = !is_d && c.has_lsm, use_topo = !is_d && c.has_topo, use_cond = c.cond_channels > 0 && job->cond != nullptr;
    B2D_CHECK(!use_lsm || job->lsm, "model was built with lsm conditioning: per-date lsm fields are required");
    B2D_CHECK(!use_topo || job->topo, "model was built with topography conditioning: per-date topo fields are required");
    B2D_CHECK(is_d || (c.cond_channels > 0) == (job->cond != nullptr), "cond fields must match cond_on_img of the model");
    const size_t cond_row = !use_cond ? 0 : (is_d ? (size_t)c.cond_channels * job->cond_h * job->cond_w : (size_t)c.cond_channels * plane);
    if (job->y) {
        B2D_CHECK(c.num_classes > 0, "y given but the model has no label embedding");
        for (int d = 0; d < job->n_dates; ++d)
            B2D_CHECK(job->y[d] >= 0 && job->y[d] < c.num_classes, "class label out of range");
    }
    B2D_TRY(ensemble_resources(h, job->sub_batch, cond_row));
    EnsembleRes* r = h->ens;
    cudaStream_t cs = h->own_stream;
    // the caller's output array is the D2H target when it can be page-locked for the duration of the job
    const size_t out_bytes = (size_t)job->count * per_sample * 4;
    const bool out_pinned = job->count > 0 && cudaHostRegister(job->out, out_bytes, cudaHostRegisterDefault) == cudaSuccess;
    if (!out_pinned) cudaGetLastError();
    const auto t0 = std::chrono::steady_clock::now();
    double gather_ms = 0;
    int64_t launches = 0;
    int nsub = 0, rc = 0;
    struct Pending { int s; size_t off, n; } pend[2] = {{-1, 0, 0}, {-1, 0, 0}};
    auto flush_bounce = [&](int s) {      // un-pinned output: copy the bounce buffer of stage s into the caller's array
        if (!out_pinned && pend[s].s >= 0) {
            cudaEventSynchronize(r->ev_d2h[s]);
            memcpy(job->out + pend[s].off, r->h_out[s], pend[s].n * 4);
            pend[s].s = -1;
        }
    };
    for (long long j0 = job->first; j0 < (long long)job->first + job->count && rc == 0; j0 += job->sub_batch, ++nsub) {
        const int s = nsub & 1;
        const int Bk = (int)std::min<long long>(job->sub_batch, (long long)job->first + job->count - j0);
        // pinned staging of stage s is free once its previous upload (sub-batch k-2) has been consumed by the copy engine
        if (nsub >= 2) cudaEventSynchronize(r->ev_h2d[s]);
        const auto g0 = std::chrono::steady_clock::now();
        for (int i = 0; i < Bk; ++i) {
            const long long d = (j0 + i) / job->members;      // date-major member list
            if (use_lsm) memcpy(r->h_lsm[s] + (size_t)i * plane, job->lsm + (size_t)d * plane, plane * 4);
            if (use_topo) memcpy(r->h_topo[s] + (size_t)i * plane, job->topo + (size_t)d * plane, plane * 4);
            if (use_cond) memcpy(r->h_cond[s] + (size_t)i * cond_row, job->cond + (size_t)d * cond_row, cond_row * 4);
            if (job->y) r->h_y[s][i] = (int)job->y[d];
        }
        gather_ms += std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - g0).count();
        auto body = [&]() -> int {
            // device staging of stage s was last read by the conditioning pre-pass of sub-batch k-2 (long finished: the compute
            // stream has since run a whole sampling loop), uploads go through the copy stream
            if (nsub >= 2) B2D_CUDA(cudaStreamWaitEvent(r->copy, r->ev_done[s], 0));
            if (use_lsm) B2D_CUDA(cudaMemcpyAsync(r->d_lsm[s], r->h_lsm[s], (size_t)Bk * plane * 4, cudaMemcpyHostToDevice, r->copy));
            if (use_topo) B2D_CUDA(cudaMemcpyAsync(r->d_topo[s], r->h_topo[s], (size_t)Bk * plane * 4, cudaMemcpyHostToDevice, r->copy));
            if (use_cond) B2D_CUDA(cudaMemcpyAsync(r->d_cond[s], r->h_cond[s], (size_t)Bk * cond_row * 4, cudaMemcpyHostToDevice, r->copy));
            if (job->y) B2D_CUDA(cudaMemcpyAsync(r->d_y[s], r->h_y[s], (size_t)Bk * 4, cudaMemcpyHostToDevice, r->copy));
            B2D_CUDA(cudaEventRecord(r->ev_h2d[s], r->copy));
            B2D_TRY(ensure_program(h, Bk));
            B2D_CUDA(cudaStreamWaitEvent(cs, r->ev_h2d[s], 0));
            const size_t n = (size_t)Bk * per_sample;
            const int blocks = (int)std::min<size_t>((n / 4 + 255) / 256, (size_t)h->num_sms * 8);
            B2D_CUDA(launch_k(init_noise_kernel, dim3(blocks), dim3(256), 0, cs, h->d_x_work, n, per_sample, (unsigned long long)job->seed,
                              (unsigned long long)j0, job->xT_scale));
            B2D_TRY(set_conditioning_impl(h, use_lsm ? r->d_lsm[s] : nullptr, use_topo ? r->d_topo[s] : nullptr,
                                          use_cond ? r->d_cond[s] : nullptr, job->cond_h, job->cond_w, nullptr, nullptr, Bk, cs,
                                          job->y ? r->d_y[s] : nullptr));
            B2D_TRY(sample_core(h, nullptr, job->seed, (uint64_t)j0, job->noise_scale, Bk, cs));
            launches += h->last_launches + 1;
            // park the fields so that the next sub-batch can start while they travel to the host
            if (nsub >= 2) B2D_CUDA(cudaStreamWaitEvent(cs, r->ev_d2h[s], 0));
            B2D_CUDA(cudaMemcpyAsync(r->d_out[s], h->d_x_work, n * 4, cudaMemcpyDeviceToDevice, cs));
            B2D_CUDA(cudaEventRecord(r->ev_done[s], cs));
            B2D_CUDA(cudaStreamWaitEvent(r->copy, r->ev_done[s], 0));
            flush_bounce(s);
            const size_t off = (size_t)(j0 - job->first) * per_sample;
            float* dst = out_pinned ? job->out + off : r->h_out[s];
            B2D_CUDA(cudaMemcpyAsync(dst, r->d_out[s], n * 4, cudaMemcpyDeviceToHost, r->copy));
            B2D_CUDA(cudaEventRecord(r->ev_d2h[s], r->copy));
            pend[s] = {s, off, n};
            return 0;
        };
        rc = body();
    }
    cudaError_t e1 = cudaStreamSynchronize(cs), e2 = cudaStreamSynchronize(r->copy);
    flush_bounce(0);
    flush_bounce(1);
    if (out_pinned) cudaHostUnregister(job->out);
    if (rc) return rc;
    B2D_CUDA(e1);
    B2D_CUDA(e2);
    h->last_launches = launches;
    if (stats) {
        stats->sub_batches = nsub;
        stats->launches = launches;
        stats->gather_ms = gather_ms;
        stats->wall_ms = std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t0).count();
        stats->out_pinned = out_pinned ? 1 : 0;
    }
    return 0;
}
