// Multi-device context: ONE process, a list of GPUs, the coset-sharded commit of SURVEY.md section 8e
// driven from inside the library -- what a Rust `PolynomialBatch::from_values` behind the cargo feature
// (one process, plonky2/src/fri/oracle.rs:168-175) needs in order to reach every GPU of the box.
// Included at the end of qp_plonky2.cu (single translation unit: it uses that file's internals).
//
// Decomposition (same as qp-plonky2_b200/dist.py, which is the one-process-per-GPU form):
//   * columns are sharded across the devices for the upload and the inverse transform;
//   * a piece of <= 8 coefficient columns is PULLED from its owner by every other device as a peer copy
//     (cudaMemcpyPeerAsync over NVLink / NVSwitch on the pulling device's transfer stream, copy engines:
//     no SM is taken from the hashing, no NCCL) once its inverse transform is done, pieces in global
//     column order, a few ahead of the compute stream;
//   * device e extends every piece to ITS cosets (leaf blocks [e 2^r / D, (e + 1) 2^r / D) = whole cap
//     subtrees) and advances its leaf sponges over the column prefix while later pieces are in flight;
//   * the cap is the concatenation of the shards' caps (no data-path collective besides the peer copies).
// Two host threads per device: a producer (stages pageable columns -- host-blocking --, uploads, inverse
// transforms on the device's producer context, issues the peer copies) and a consumer (extends / hashes
// on the device's main context).  Cross-thread ordering: a piece's events are recorded by its producer,
// which then publishes `issued[k]`; consumers wait for the flag before they make their stream wait on the
// event (an event that has not been recorded yet would not block anything).

// A piece travels as a COPY KERNEL on the pulling device: 16-byte loads straight out of the owner's memory over
// NVLink (peer access), grid-stride, a few dozen blocks -- measured here, cudaMemcpyPeerAsync blocked the issuing
// host thread for milliseconds per call when several devices pulled at once, which serialised the whole pipeline;
// a kernel launch does not.  (Without peer access the driver-staged cudaMemcpyPeerAsync remains the fallback.)
__global__ void __launch_bounds__(256) peer_copy_kernel(const ulonglong2* __restrict__ src, ulonglong2* __restrict__ dst,
                                                        size_t n16) {
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n16; i += (size_t)gridDim.x * blockDim.x)
        dst[i] = src[i];
}

struct qp_mctx {
    std::vector<int> devices;
    std::vector<qp_ctx*> main_ctx;   // extend + hash
    std::vector<qp_ctx*> prod_ctx;   // upload + inverse transform
    std::vector<cudaStream_t> xfer;  // peer copies out of device d
    // the producer's stream: HIGH priority -- an inverse transform launched while a leaf-hash kernel (thousands of
    // blocks) is running must get SM slots as blocks retire, not after all of that kernel's blocks
    std::vector<cudaStream_t> prod_stream;
    std::vector<uint8_t> peer_ok;    // [D * D]: device a can load from device b's memory
    std::string err;
};

struct qp_mbatch {
    qp_mctx* m = nullptr;
    std::vector<qp_batch*> shards;
    unsigned cap_height = 0;
};

extern "C" void qp_mctx_destroy(qp_mctx* m) {
    if (!m) return;
    for (size_t d = 0; d < m->devices.size(); d++) {
        cudaSetDevice(m->devices[d]);
        if (d < m->xfer.size() && m->xfer[d]) cudaStreamDestroy(m->xfer[d]);
        if (d < m->main_ctx.size()) qp_ctx_destroy(m->main_ctx[d]);
        if (d < m->prod_ctx.size()) qp_ctx_destroy(m->prod_ctx[d]);
        if (d < m->prod_stream.size() && m->prod_stream[d]) cudaStreamDestroy(m->prod_stream[d]);
        peer_buf_trim(m->devices[d]);   // cached peer-readable buffers nobody uses go back to the driver
    }
    delete m;
}

extern "C" int qp_mctx_create(const int* devices, unsigned n_devices, unsigned max_lde_log, qp_mctx** out) {
    if (!out) return QP_ERR_BAD_ARG;
    *out = nullptr;
    if (!devices || n_devices == 0 || (n_devices & (n_devices - 1))) return QP_ERR_BAD_ARG;  // coset sharding: power of two
    for (unsigned a = 0; a < n_devices; a++)
        for (unsigned b = a + 1; b < n_devices; b++)
            if (devices[a] == devices[b]) return QP_ERR_BAD_ARG;
    qp_mctx* m = new qp_mctx();
    m->devices.assign(devices, devices + n_devices);
    m->main_ctx.assign(n_devices, nullptr);
    m->prod_ctx.assign(n_devices, nullptr);
    m->xfer.assign(n_devices, nullptr);
    m->prod_stream.assign(n_devices, nullptr);
    m->peer_ok.assign((size_t)n_devices * n_devices, 0);
    for (unsigned d = 0; d < n_devices; d++) {
        int rc = qp_ctx_create(devices[d], nullptr, max_lde_log, &m->main_ctx[d]);
        int lo_prio = 0, hi_prio = 0;
        if (!rc && (cudaSetDevice(devices[d]) != cudaSuccess ||
                    cudaDeviceGetStreamPriorityRange(&lo_prio, &hi_prio) != cudaSuccess ||
                    cudaStreamCreateWithPriority(&m->prod_stream[d], cudaStreamNonBlocking, hi_prio) != cudaSuccess ||
                    cudaStreamCreateWithPriority(&m->xfer[d], cudaStreamNonBlocking, hi_prio) != cudaSuccess))
            rc = QP_ERR_CUDA;
        if (!rc) rc = qp_ctx_create(devices[d], m->prod_stream[d], max_lde_log, &m->prod_ctx[d]);
        if (rc) {
            qp_mctx_destroy(m);
            return rc;
        }
        // direct peer access where the topology has it (NVLink / NVSwitch); without it cudaMemcpyPeerAsync
        // still works, staged by the driver
        for (unsigned p = 0; p < n_devices; p++) {
            if (p == d) continue;
            int can = 0;
            cudaSetDevice(devices[d]);
            if (cudaDeviceCanAccessPeer(&can, devices[d], devices[p]) == cudaSuccess && can) {
                cudaError_t e = cudaDeviceEnablePeerAccess(devices[p], 0);
                if (e != cudaSuccess) cudaGetLastError();  // already enabled is fine
                // (the matrices the peers read are plain cudaMalloc buffers, peer_buf_acquire: device-level peer access
                //  covers them; the stream-ordered pools stay private to their device)
                const bool ok = (e == cudaSuccess || e == cudaErrorPeerAccessAlreadyEnabled);
                m->peer_ok[(size_t)d * n_devices + p] = ok ? 1 : 0;
                if (getenv("QP_TRACE")) fprintf(stderr, "[qp_mctx] peer access %d -> %d: %s\n", devices[d], devices[p], cudaGetErrorString(e));
            } else if (getenv("QP_TRACE")) {
                fprintf(stderr, "[qp_mctx] peer access %d -> %d: not available (copies are staged by the driver)\n", devices[d], devices[p]);
            }
        }
    }
    *out = m;
    return QP_OK;
}

extern "C" unsigned qp_mctx_num_devices(const qp_mctx* m) { return m ? (unsigned)m->devices.size() : 0; }
extern "C" qp_ctx* qp_mctx_ctx(qp_mctx* m, unsigned i) { return (m && i < m->main_ctx.size()) ? m->main_ctx[i] : nullptr; }
extern "C" const char* qp_mctx_last_error(const qp_mctx* m) { return m ? m->err.c_str() : "no context"; }

extern "C" void qp_mbatch_free(qp_mbatch* b) {
    if (!b) return;
    for (qp_batch* s : b->shards) qp_batch_free(s);
    delete b;
}
extern "C" unsigned qp_mbatch_num_shards(const qp_mbatch* b) { return b ? (unsigned)b->shards.size() : 0; }
extern "C" qp_batch* qp_mbatch_shard(qp_mbatch* b, unsigned i) { return (b && i < b->shards.size()) ? b->shards[i] : nullptr; }

extern "C" int qp_mbatch_cap(const qp_mbatch* b, uint64_t* out) {
    if (!b || !out) return QP_ERR_BAD_ARG;
    size_t off = 0;
    for (qp_batch* s : b->shards) {
        int rc = qp_batch_cap(s, out + off, QP_HOST);
        if (rc) {
            b->m->err = qp_last_error(s->ctx);
            return rc;
        }
        off += s->tree.n_cap() * 4;
    }
    return QP_OK;
}

namespace multi {
struct Piece {
    unsigned owner;
    size_t c0, c1;
};
static void column_shard(size_t n_cols, unsigned world, unsigned rank, size_t* lo, size_t* hi) {
    const size_t base = n_cols / world, extra = n_cols % world;
    *lo = rank * base + std::min<size_t>(rank, extra);
    *hi = *lo + base + (rank < extra ? 1 : 0);
}
}  // namespace multi

// Source of the columns: host column vectors (sharded across the devices for upload + inverse transform), or a
// matrix [n_cols][n] already resident on devices[0] (values or coefficients: the Z / partial-product columns and
// the quotient chunks of a proof are computed there) -- then device 0 owns every piece.
static int mbatch_build(qp_mctx* m, const uint64_t* const* cols, const uint64_t* dev0_data, int dev0_is_coeffs,
                        size_t n_cols, unsigned degree_log, unsigned rate_bits, int blinding, unsigned cap_height,
                        const uint64_t* salt, qp_mbatch** out);

extern "C" int qp_mbatch_from_values_cols(qp_mctx* m, const uint64_t* const* cols, size_t n_cols, unsigned degree_log,
                                          unsigned rate_bits, int blinding, unsigned cap_height, const uint64_t* salt,
                                          qp_mbatch** out) {
    if (!m) return QP_ERR_BAD_ARG;
    if (!cols) {
        m->err = "null column table";
        return QP_ERR_BAD_ARG;
    }
    for (size_t c = 0; c < n_cols; c++)
        if (!cols[c]) {
            m->err = "null column";
            return QP_ERR_BAD_ARG;
        }
    return mbatch_build(m, cols, nullptr, 0, n_cols, degree_log, rate_bits, blinding, cap_height, salt, out);
}

// from_values / from_coeffs on a matrix [n_cols][2^degree_log] resident on devices[0] (complete before the call);
// `salt` is host memory.
extern "C" int qp_mbatch_from_device(qp_mctx* m, const uint64_t* dev0_data, int is_coeffs, size_t n_cols,
                                     unsigned degree_log, unsigned rate_bits, int blinding, unsigned cap_height,
                                     const uint64_t* salt, qp_mbatch** out) {
    if (!m) return QP_ERR_BAD_ARG;
    if (!dev0_data) {
        m->err = "null device matrix";
        return QP_ERR_BAD_ARG;
    }
    return mbatch_build(m, nullptr, dev0_data, is_coeffs, n_cols, degree_log, rate_bits, blinding, cap_height, salt, out);
}

static int mbatch_build(qp_mctx* m, const uint64_t* const* cols, const uint64_t* dev0_data, int dev0_is_coeffs,
                        size_t n_cols, unsigned degree_log, unsigned rate_bits, int blinding, unsigned cap_height,
                        const uint64_t* salt, qp_mbatch** out) {
    if (!m) return QP_ERR_BAD_ARG;
    auto mfail = [&](int code, const char* msg) {
        m->err = msg;
        return code;
    };
    if (!out) return mfail(QP_ERR_BAD_ARG, "null out");
    *out = nullptr;
    const unsigned D = (unsigned)m->devices.size();
    if (D > (1u << rate_bits) || D > (1u << cap_height))
        return mfail(QP_ERR_BAD_ARG, "coset sharding needs #devices <= 2^rate_bits and <= 2^cap_height");
    const unsigned blocks = (1u << rate_bits) / D;
    const size_t n = (size_t)1 << degree_log;
    const size_t PIECE = 8;
    std::vector<multi::Piece> pieces;
    auto shard_of = [&](unsigned d, size_t* lo, size_t* hi) {
        if (dev0_data) {  // everything lives on device 0
            *lo = 0;
            *hi = d == 0 ? n_cols : 0;
        } else {
            multi::column_shard(n_cols, D, d, lo, hi);
        }
    };
    for (unsigned d = 0; d < D; d++) {
        size_t lo, hi;
        shard_of(d, &lo, &hi);
        for (size_t c0 = lo; c0 < hi; c0 += PIECE) pieces.push_back({d, c0, std::min(c0 + PIECE, hi)});
    }
    const size_t K = pieces.size();
    qp_mbatch* mb = new qp_mbatch();
    mb->m = m;
    mb->cap_height = cap_height;
    mb->shards.assign(D, nullptr);
    // shards: the full coefficient matrix on every device (each extends ALL columns to its own cosets)
    int rc = QP_OK;
    for (unsigned d = 0; d < D && !rc; d++) {
        qp_ctx* ctx = m->main_ctx[d];
        rc = check_batch_args(ctx, n_cols, degree_log, rate_bits, blinding, cap_height, salt, d * blocks, blocks,
                              &mb->shards[d]);
        uint64_t* d_coeffs = nullptr;
        // with peers, the matrix is read by them: a peer-readable buffer instead of a pool allocation
        if (!rc) rc = D > 1 ? peer_buf_acquire(ctx, n_cols * n, &d_coeffs) : dev_alloc(ctx, &d_coeffs, n_cols * n);
        if (!rc) {
            rc = batch_create(ctx, d_coeffs, n_cols, degree_log, rate_bits, blinding, cap_height, d * blocks, blocks,
                              &mb->shards[d]);
            if (mb->shards[d]) mb->shards[d]->coeffs_peer_buf = D > 1;   // (the batch owns the matrix from here on)
        }
        if (!rc) {
            cudaEventRecord(ctx->ev[1], ctx->stream);
            if (cudaStreamSynchronize(ctx->stream) != cudaSuccess) rc = QP_ERR_CUDA;  // peers write into the matrix
        }
        if (rc) m->err = ctx->err;
    }
    // events: ev_ready[k] on the owner's producer stream (inverse transform done), ev_arrived[k * D + p] on
    // the owner's transfer stream (copy to device p done)
    const auto t_setup = std::chrono::steady_clock::now();
    std::vector<cudaEvent_t> ev_ready(K, nullptr), ev_arrived(K * D, nullptr);
    for (size_t k = 0; k < K && !rc; k++) {
        cudaSetDevice(m->devices[pieces[k].owner]);
        if (cudaEventCreateWithFlags(&ev_ready[k], cudaEventDisableTiming) != cudaSuccess) rc = QP_ERR_CUDA;
        for (unsigned p = 0; p < D && !rc; p++) {
            if (p == pieces[k].owner) continue;
            cudaSetDevice(m->devices[p]);   // recorded on the PULLING device's transfer stream
            if (cudaEventCreateWithFlags(&ev_arrived[k * D + p], cudaEventDisableTiming) != cudaSuccess) rc = QP_ERR_CUDA;
        }
    }
    const bool trace = getenv("QP_TRACE") != nullptr;
    const auto t_begin = std::chrono::steady_clock::now();
    if (trace)
        fprintf(stderr, "[qp_mbatch] event creation took %.2f ms\n",
                std::chrono::duration<double, std::milli>(t_begin - t_setup).count());
    auto since = [&]() { return std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t_begin).count(); };
    std::vector<double> t_prod_done(D, 0), t_cons_issued(D, 0), t_cons_done(D, 0), t_first_piece(D, 0);
    std::vector<std::atomic<int>> issued(K);
    for (auto& f : issued) f.store(0);
    std::atomic<int> abort_flag{0};
    std::vector<int> rcs(2 * D, QP_OK);

    auto producer = [&](unsigned d) -> int {
        qp_ctx* ctx = m->prod_ctx[d];
        if (cudaSetDevice(m->devices[d]) != cudaSuccess) return QP_ERR_CUDA;
        size_t lo, hi;
        shard_of(d, &lo, &hi);
        if (lo == hi) return QP_OK;
        TempScope tmp(ctx);
        uint64_t* d_values = nullptr;
        int r = QP_OK;
        if (!dev0_data) {
            r = tmp.alloc(&d_values, (hi - lo) * n);
            if (!r) r = ensure_ring(ctx, std::min(PIECE, hi - lo) * n);
            if (r) return r;
            cudaEventRecord(ctx->ready_ev, ctx->stream);
            cudaStreamWaitEvent(ctx->copy_stream, ctx->ready_ev, 0);
        }
        int g = 0;
        for (size_t k = 0; k < K; k++) {
            if (pieces[k].owner != d) continue;
            if (abort_flag.load()) return QP_OK;
            const size_t c0 = pieces[k].c0, c1 = pieces[k].c1;
            uint64_t* own = mb->shards[d]->coeffs + c0 * n;
            if (dev0_data) {
                if (dev0_is_coeffs)
                    CUDA_TRY(ctx, cudaMemcpyAsync(own, dev0_data + c0 * n, (c1 - c0) * n * 8, cudaMemcpyDeviceToDevice, ctx->stream));
                else
                    r = ifft_device(ctx, dev0_data + c0 * n, c1 - c0, degree_log, own, nullptr);
            } else {
                uint64_t* dv = d_values + (c0 - lo) * n;
                // (columns the caller has pinned -- cudaHostRegister / cudaMallocHost -- skip the staging ring)
                r = upload_columns(ctx, cols, host_pointer_is_pinned(cols[c0]) && host_pointer_is_pinned(cols[c1 - 1]), c0,
                                   c1, n, dv, g);
                if (r) return r;
                cudaStreamWaitEvent(ctx->stream, ctx->copy_ev[g % qp_ctx::MAX_GROUPS], 0);
                g++;
                r = ifft_device(ctx, dv, c1 - c0, degree_log, own, dv);
            }
            if (r) return r;
            const double t_a = since();
            CUDA_TRY(ctx, cudaEventRecord(ev_ready[k], ctx->stream));
            issued[k].store(1, std::memory_order_release);   // the consumers PULL the piece (below)
            if (trace && d == D - 1)
                fprintf(stderr, "[qp_mbatch]     producer %u piece %zu: upload + inverse transform queued at %.2f ms\n", d, k, t_a);
        }
        // the staging buffer and the ring must outlive the copies that read them
        CUDA_TRY(ctx, cudaStreamSynchronize(ctx->stream));
        t_prod_done[d] = since();
        return QP_OK;
    };
    auto consumer = [&](unsigned d) -> int {
        qp_ctx* ctx = m->main_ctx[d];
        qp_batch* b = mb->shards[d];
        if (cudaSetDevice(m->devices[d]) != cudaSuccess) return QP_ERR_CUDA;
        // pieces of other owners are PULLED by this device on its own transfer stream (peer copy out of the
        // owner's matrix once its inverse transform is done), a few pieces ahead of the compute stream
        const size_t LOOKAHEAD = 4;
        size_t pulled = 0;
        auto pull_up_to = [&](size_t upto) -> int {
            for (; pulled < upto && pulled < K; pulled++) {
                const size_t k = pulled;
                while (!issued[k].load(std::memory_order_acquire)) {
                    if (abort_flag.load()) return QP_OK;
                    std::this_thread::yield();
                }
                const unsigned o = pieces[k].owner;
                if (o == d) continue;
                const size_t c0 = pieces[k].c0, c1 = pieces[k].c1;
                CUDA_TRY(ctx, cudaStreamWaitEvent(m->xfer[d], ev_ready[k], 0));
                if (m->peer_ok[(size_t)d * D + o]) {
                    const size_t n16 = (c1 - c0) * n / 2;
                    peer_copy_kernel<<<64, 256, 0, m->xfer[d]>>>(
                        reinterpret_cast<const ulonglong2*>(mb->shards[o]->coeffs + c0 * n),
                        reinterpret_cast<ulonglong2*>(b->coeffs + c0 * n), n16);
                    ctx->launches++;
                    CUDA_TRY(ctx, cudaGetLastError());
                } else {
                    CUDA_TRY(ctx, cudaMemcpyPeerAsync(b->coeffs + c0 * n, m->devices[d], mb->shards[o]->coeffs + c0 * n,
                                                      m->devices[o], (c1 - c0) * n * 8, m->xfer[d]));
                }
                CUDA_TRY(ctx, cudaEventRecord(ev_arrived[k * D + d], m->xfer[d]));
            }
            return QP_OK;
        };
        for (size_t k = 0; k < K; k++) {
            int r = pull_up_to(k + 1 + LOOKAHEAD);
            if (r) return r;
            if (abort_flag.load()) return QP_OK;
            cudaEvent_t e = pieces[k].owner == d ? ev_ready[k] : ev_arrived[k * D + d];
            if (k == 0) t_first_piece[d] = since();
            CUDA_TRY(ctx, cudaStreamWaitEvent(ctx->stream, e, 0));
            r = batch_extend(b, pieces[k].c0, pieces[k].c1 - pieces[k].c0, true);
            if (r) return r;
        }
        t_cons_issued[d] = since();
        TempScope tmp(ctx);
        const uint64_t* d_salt = nullptr;
        if (blinding) {
            uint64_t* salt_owned = nullptr;
            int r = to_device(ctx, salt, QP_HOST, (size_t)QP_SALT_SIZE << (degree_log + rate_bits), &d_salt, &salt_owned);
            tmp.adopt(salt_owned);
            if (r) return r;
        }
        int r = batch_finish(b, d_salt, b->chunks_done, b->sponge_state);
        t_cons_done[d] = since();
        dev_free(ctx, b->sponge_state);
        b->sponge_state = nullptr;
        return r;
    };
    if (!rc) {
        std::vector<std::thread> th;
        for (unsigned d = 0; d < D; d++) {
            th.emplace_back([&, d] {
                rcs[2 * d] = producer(d);
                if (rcs[2 * d]) abort_flag.store(1);
            });
            th.emplace_back([&, d] {
                rcs[2 * d + 1] = consumer(d);
                if (rcs[2 * d + 1]) abort_flag.store(1);
            });
        }
        for (auto& t : th) t.join();
        for (unsigned d = 0; d < D && !rc; d++) {
            if (rcs[2 * d]) {
                rc = rcs[2 * d];
                m->err = m->prod_ctx[d]->err;
            } else if (rcs[2 * d + 1]) {
                rc = rcs[2 * d + 1];
                m->err = m->main_ctx[d]->err;
            }
        }
    }
    // every stream that touched the events is idle before they go
    for (unsigned d = 0; d < D; d++) {
        cudaSetDevice(m->devices[d]);
        cudaStreamSynchronize(m->xfer[d]);
        cudaStreamSynchronize(m->prod_ctx[d]->stream);
        cudaStreamSynchronize(m->main_ctx[d]->stream);
    }
    if (trace) {
        fprintf(stderr, "[qp_mbatch] %zu cols, %u devices, %zu pieces: setup done at t=0, all done at %.2f ms\n", n_cols, D, K, since());
        for (unsigned d = 0; d < D; d++)
            fprintf(stderr, "[qp_mbatch]   device %u: producer drained %.2f ms | consumer: first piece seen %.2f, all issued %.2f, "
                            "finished %.2f ms (LDE+sponge phase %.2f ms, final hash + tree %.2f ms on the device)\n",
                    d, t_prod_done[d], t_first_piece[d], t_cons_issued[d], t_cons_done[d], mb->shards[d] ? mb->shards[d]->ms[1] : 0.f,
                    mb->shards[d] ? mb->shards[d]->ms[3] : 0.f);
    }
    for (cudaEvent_t e : ev_ready)
        if (e) cudaEventDestroy(e);
    for (cudaEvent_t e : ev_arrived)
        if (e) cudaEventDestroy(e);
    if (rc) {
        qp_mbatch_free(mb);
        return rc;
    }
    *out = mb;
    return QP_OK;
}
