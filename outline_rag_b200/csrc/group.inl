// A multi-GPU index inside ONE process (orx_create_multi): the row-sharded table of SURVEY.md 8(e) behind the SAME
// C-ABI handle as a single-GPU index, so that the drop-in object bound to `rag.vector_store` (reference app/rag.py:28,
// one per uvicorn worker process, entrypoint.sh:16) can own all the GPUs of a box without torchrun ranks.
//
// Included by orx_api.cu inside its translation unit (it drives the shards through the same internal functions as the
// one-process-per-GPU path: stage_queries / scan_pass / finalize-with-publish / merge_wait).
//
//   * shard g = one ordinary orx_index on devices[g]; rows are placed by mix64(id) mod G (the shard function of
//     outline_rag_b200/sharded.py), so upserts and deletes stay balanced and need no data-path exchange;
//   * one worker thread per distinct device launches that device's chain (prep -> scan -> finalize); the finalize
//     kernel of every shard pushes its k candidates per query STRAIGHT into the root GPU's gather buffer with peer
//     stores over NVLink and raises its arrival word there; the root's merge_wait kernel merges the G lists and writes
//     the answer into mapped host memory (or the caller's device arrays); the calling thread polls the completion word;
//   * shards that share a device (tests on one GPU, devices = [0, 0, 0]) share ONE stream, so the merge is ordered
//     behind their finalize kernels by the stream and never waits for a kernel that cannot run.

#ifdef ORX_GROUP_TYPES      // first inclusion (global scope): the types

constexpr int WORKER_SPIN_MS = 8;      // longer than one search of a 10M-row table on one GPU

struct Worker {
    int device = 0;
    std::vector<int> shards;                 // indices into Group::shards
    std::thread th;
    std::mutex m;
    std::condition_variable cv;
    std::function<int()> job;
    std::atomic<uint64_t> posted{0}, finished{0};
    std::atomic<bool> sleeping{false}, quit{false};
    int rc = ORX_OK;
    std::string err;

    void run() {
        cudaSetDevice(device);
        uint64_t seen = 0;
        for (;;) {
            // spin for a few milliseconds (a loop of searches re-posts one search time later -- 0.75 ms at 8 GPUs, 5.6 ms
            // on one -- and a condition-variable wake-up costs 30-60 us of every search it hits), then sleep
            int spins = 0;
            std::chrono::steady_clock::time_point t_idle{};
            while (posted.load(std::memory_order_acquire) == seen && !quit.load(std::memory_order_relaxed)) {
                if ((++spins & 63) != 0) {
#if defined(__x86_64__)
                    __builtin_ia32_pause();
#endif
                    continue;
                }
                const auto now = std::chrono::steady_clock::now();
                if (spins == 64) t_idle = now;
                if (now - t_idle < std::chrono::milliseconds(WORKER_SPIN_MS)) continue;
                std::unique_lock<std::mutex> lk(m);
                sleeping.store(true, std::memory_order_seq_cst);
                cv.wait(lk, [&] { return posted.load(std::memory_order_acquire) != seen || quit.load(); });
                sleeping.store(false, std::memory_order_seq_cst);
            }
            if (quit.load() && posted.load(std::memory_order_acquire) == seen) return;
            seen = posted.load(std::memory_order_acquire);
            g_err.clear();
            rc = job ? job() : ORX_OK;
            err = g_err;
            finished.store(seen, std::memory_order_release);
        }
    }
    void post(std::function<int()> fn) {
        job = std::move(fn);
        posted.fetch_add(1, std::memory_order_release);
        if (sleeping.load(std::memory_order_seq_cst)) {
            std::lock_guard<std::mutex> lk(m);
            cv.notify_one();
        }
    }
    void wait() {
        const uint64_t want = posted.load(std::memory_order_acquire);
        while (finished.load(std::memory_order_acquire) != want) {
#if defined(__x86_64__)
            __builtin_ia32_pause();
#endif
        }
    }
};

struct Group {
    int dtype = ORX_DTYPE_F32;
    std::vector<int> devices;
    std::vector<orx_index *> shards;
    std::vector<std::unique_ptr<Worker>> workers;       // one per distinct device, in order of first appearance
    std::vector<int> worker_of_shard;
    std::vector<cudaStream_t> streams;                  // per worker; shared by the shards of that device
    uint32_t seq = 0;
    PinBuf<float> h_q;                                  // the query batch, staged once for all devices
    std::mutex mu;
    uint64_t searches = 0, queries = 0;
    float last_search_ms = 0.f;
    int last_path = 0;
};

#else                       // second inclusion (inside orx_api.cu's anonymous namespace): the functions

uint32_t shard_of_id(const orx_id &id, uint32_t world) {      // == outline_rag_b200/sharded.py shard_of
    uint64_t z = id.hi * 0x9E3779B97F4A7C15ull ^ id.lo;
    z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
    z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
    z = z ^ (z >> 31);
    return (uint32_t)(z % world);
}

// run fn(shard index) for every shard, the shards of one device in order on that device's worker, devices in parallel
int group_run(Group *g, const std::function<int(int)> &fn) {
    for (auto &w : g->workers) {
        Worker *wp = w.get();
        wp->post([wp, &fn] {
            for (int s : wp->shards) {
                int rc = fn(s);
                if (rc != ORX_OK) return rc;
            }
            return (int)ORX_OK;
        });
    }
    int rc = ORX_OK;
    std::string why;
    for (auto &w : g->workers) {
        w->wait();
        if (w->rc != ORX_OK && rc == ORX_OK) {
            rc = w->rc;
            why = w->err;
        }
    }
    if (rc != ORX_OK) g_err = why;
    return rc;
}

void group_destroy(orx_index *h) {
    Group *g = h->group;
    if (g) {
        for (auto &w : g->workers) {
            w->quit.store(true);
            {
                std::lock_guard<std::mutex> lk(w->m);
                w->cv.notify_one();
            }
            if (w->th.joinable()) w->th.join();
        }
        for (orx_index *s : g->shards)
            if (s) {
                // the shards' stream belongs to the group: detach it before the shard is destroyed
                orx_destroy(s);
            }
        for (size_t i = 0; i < g->streams.size(); ++i)
            if (g->streams[i]) {
                DeviceGuard dg(g->workers[i]->device);
                cudaStreamDestroy(g->streams[i]);
            }
        g->h_q.release();
        delete g;
    }
    cudaGetLastError();
    delete h;
}

int group_create(orx_index **out, int dtype, uint64_t capacity_rows, const int *devices, int n) {
    int ndev = 0;
    if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0) {
        cudaGetLastError();
        return fail(ORX_ERR_CUDA, "no CUDA device: this engine has no CPU fallback");
    }
    for (int i = 0; i < n; ++i)
        if (devices[i] < 0 || devices[i] >= ndev)
            return fail(ORX_ERR_INVALID, "device %d out of range (%d present)", devices[i], ndev);
    orx_index *h = new orx_index();
    Group *g = new Group();
    h->group = g;
    h->dtype = dtype;
    h->device = devices[0];
    g->dtype = dtype;
    g->devices.assign(devices, devices + n);
    g->shards.assign(n, nullptr);
    g->worker_of_shard.assign(n, -1);
    auto bail = [&](int rc) {
        const std::string why = g_err;
        group_destroy(h);
        g_err = why;
        return rc;
    };
    // ---- workers and streams: one per distinct device
    for (int i = 0; i < n; ++i) {
        int w = -1;
        for (size_t j = 0; j < g->workers.size(); ++j)
            if (g->workers[j]->device == devices[i]) w = (int)j;
        if (w < 0) {
            g->workers.emplace_back(new Worker());
            w = (int)g->workers.size() - 1;
            g->workers[w]->device = devices[i];
            g->streams.push_back(nullptr);
        }
        g->workers[w]->shards.push_back(i);
        g->worker_of_shard[i] = w;
    }
    // ---- peer access between every pair of distinct devices (finalize stores into the root's gather buffer; the
    //      shards read a device-resident query batch from the GPU it lives on)
    for (auto &a : g->workers)
        for (auto &b : g->workers) {
            if (a->device == b->device) continue;
            DeviceGuard dg(a->device);
            int can = 0;
            cudaDeviceCanAccessPeer(&can, a->device, b->device);
            if (!can) return bail(fail(ORX_ERR_CUDA, "GPU %d cannot access GPU %d's memory (NVLink / P2P peer access is required "
                                                    "for a multi-GPU index)", a->device, b->device));
            cudaError_t e = cudaDeviceEnablePeerAccess(b->device, 0);
            if (e != cudaSuccess && e != cudaErrorPeerAccessAlreadyEnabled)
                return bail(fail(ORX_ERR_CUDA, "cudaDeviceEnablePeerAccess(%d -> %d): %s", a->device, b->device, cudaGetErrorString(e)));
            cudaGetLastError();
        }
    // ---- the shards
    const uint64_t per = capacity_rows / n + capacity_rows / (8ull * n) + 4096;
    for (int i = 0; i < n; ++i) {
        int rc = orx_create(&g->shards[i], ORX_DIM, dtype, per, devices[i]);
        if (rc != ORX_OK) return bail(rc);
        const int w = g->worker_of_shard[i];
        DeviceGuard dg(devices[i]);
        if (!g->streams[w]) {
            cudaError_t e = cudaStreamCreateWithFlags(&g->streams[w], cudaStreamNonBlocking);
            if (e != cudaSuccess) return bail(fail(ORX_ERR_CUDA, "cudaStreamCreate: %s", cudaGetErrorString(e)));
        }
        g->shards[i]->stream = g->streams[w];
    }
    // ---- the gather buffer lives on the root (shard 0's device); every shard's exchange has ONE target: the root
    const size_t slot_bytes = slot_layout(XQ_MAX, 32).bytes;
    const size_t set_bytes = slot_bytes * n;
    const size_t flags_off = 2 * set_bytes;
    const size_t total = flags_off + 2 * (size_t)n * XFLAG_STRIDE;
    char *gbase = nullptr;
    {
        DeviceGuard dg(devices[0]);
        cudaError_t e = cudaMalloc(&gbase, total);
        if (e == cudaSuccess) e = cudaMemset(gbase, 0, total);
        if (e != cudaSuccess) {
            if (gbase) cudaFree(gbase);
            return bail(fail(ORX_ERR_CUDA, "gather buffer: %s", cudaGetErrorString(e)));
        }
    }
    for (int i = 0; i < n; ++i) {
        Exchange *x = new Exchange();
        x->world = n;
        x->rank = i;
        x->n_targets = 1;
        x->slot_bytes = slot_bytes;
        x->set_bytes = set_bytes;
        x->flags_off = flags_off;
        x->total_bytes = total;
        x->base = i == 0 ? gbase : nullptr;             // the root shard owns (and frees) the buffer
        g->shards[i]->xchg = x;
        DeviceGuard dg(devices[i]);
        for (int s = 0; s < 2; ++s) {
            void *slot = gbase + (size_t)s * set_bytes + (size_t)i * slot_bytes;
            uint32_t *flag = reinterpret_cast<uint32_t *>(gbase + flags_off + ((size_t)s * n + i) * XFLAG_STRIDE);
            cudaError_t e = cudaMalloc(&x->d_peer_slot[s], sizeof(void *));
            if (e == cudaSuccess) e = cudaMalloc(&x->d_peer_flag[s], sizeof(uint32_t *));
            if (e == cudaSuccess) e = cudaMemcpy(x->d_peer_slot[s], &slot, sizeof(void *), cudaMemcpyHostToDevice);
            if (e == cudaSuccess) e = cudaMemcpy(x->d_peer_flag[s], &flag, sizeof(uint32_t *), cudaMemcpyHostToDevice);
            if (e != cudaSuccess) return bail(fail(ORX_ERR_CUDA, "exchange tables: %s", cudaGetErrorString(e)));
        }
        x->connected = true;
    }
    for (auto &w : g->workers) {
        Worker *wp = w.get();
        wp->th = std::thread([wp] { wp->run(); });
    }
    *out = h;
    return ORX_OK;
}

// ------------------------------------------------------------------------------- reads
uint64_t group_size(Group *g) {
    uint64_t n = 0;
    for (orx_index *s : g->shards) n += orx_size(s);
    return n;
}

// exact answer of ONE query from every shard's own exact search, merged on the root (rare path: a shard could not
// prove its candidates -- dense near-ties, NaN rows, a zero query)
int group_exact_query(Group *g, const float *q_host, int k, orx_id *out_ids, double *out_dist, int *out_count) {
    const int G = (int)g->shards.size();
    std::vector<orx_id> ids((size_t)G * k);
    std::vector<double> dist((size_t)G * k);
    std::vector<int> cnt(G);
    int rc = group_run(g, [&](int s) {
        orx_index *ix = g->shards[s];
        std::lock_guard<std::mutex> lk(ix->mu);
        return search_locked(ix, q_host, 1, k, ids.data() + (size_t)s * k, dist.data() + (size_t)s * k, cnt.data() + s);
    });
    if (rc != ORX_OK) return rc;
    return orx_merge_topk(g->shards[0], G, 1, k, ids.data(), dist.data(), cnt.data(), out_ids, out_dist, out_count);
}

int group_search_chunk(Group *g, const float *queries, int nq, int k, orx_id *out_ids, double *out_dist, int *out_counts) {
    const auto t_begin = std::chrono::steady_clock::now();
    const int G = (int)g->shards.size();
    orx_index *root = g->shards[0];
    const bool out_on_dev = is_device_ptr(out_ids);
    if (out_on_dev != is_device_ptr(out_dist) || out_on_dev != is_device_ptr(out_counts))
        return fail(ORX_ERR_INVALID, "out_ids, out_dist and out_counts must all be host or all be device");
    if (out_on_dev && ptr_device(out_ids) != root->device)
        return fail(ORX_ERR_INVALID, "device outputs of a multi-GPU index must live on its first device (%d)", root->device);
    const size_t nk = (size_t)nq * k;
    const SlotLayout L = slot_layout(nq, k);
    const float *q_all = queries;
    bool pinned = false;
    if (!is_device_ptr(queries)) {
        CK(g->h_q.ensure((size_t)nq * ORX_DIM));
        memcpy(g->h_q.p, queries, (size_t)nq * ORX_DIM * sizeof(float));
        q_all = g->h_q.p;
        pinned = true;
    }
    {
        DeviceGuard dg(root->device);
        CK(root->cur->h_flags.ensure(nq));
        CK(root->cur->h_myflags.ensure(nq));
        CK(root->cur->h_redo.ensure(1));
        if (!out_on_dev) {
            CK(root->cur->h_ids.ensure(nk));
            CK(root->cur->h_dist.ensure(nk));
            CK(root->cur->h_counts.ensure(nq));
        }
        int rc0 = ensure_signalling(root);
        if (rc0 != ORX_OK) return rc0;
    }
    const orx::ResultOut final_out{out_on_dev ? out_ids : root->cur->h_ids.p, out_on_dev ? out_dist : root->cur->h_dist.p,
                                   out_on_dev ? out_counts : root->cur->h_counts.p, root->cur->h_flags.p};
    *root->cur->h_redo.p = 0;
    root->cur->h_done.p[1] = 0;
    const uint32_t seq = ++g->seq;
    const uint32_t token = next_token(root);
    std::vector<int> paths(G, 1);
    // every device launches its shards' chains; the root's worker appends the merge behind its own shards' kernels
    int rc = group_run(g, [&](int s) {
        orx_index *ix = g->shards[s];
        const float *q_src = nullptr;
        int r = shard_round1(ix, ix->xchg, q_all, pinned, nq, k, L, seq, &q_src, &paths[s]);
        const bool last_on_root_device = g->worker_of_shard[s] == g->worker_of_shard[0] &&
                                         s == g->workers[g->worker_of_shard[0]]->shards.back();
        if (last_on_root_device) {
            // (also after a failed shard: its error block was published, the merge completes and reports it)
            int rm = launch_shard_merge(root, root->xchg, nq, k, final_out, L, seq, token);
            if (r == ORX_OK) r = rm;
        }
        return r;
    });
    {
        DeviceGuard dg(root->device);
        const int rw = wait_done(root, token);
        if (rc == ORX_OK) rc = rw;
    }
    if (rc != ORX_OK) return rc;
    if (root->cur->h_done.p[1]) return fail(ORX_ERR_CUDA, "multi-GPU search: a shard did not publish its candidates within 10 s");
    bool any_unproven = false;
    for (int j = 0; j < nq; ++j) {
        const int f = root->cur->h_flags.p[j];
        if (f & 8) return fail(ORX_ERR_CUDA, "multi-GPU search: a shard failed (query %d)", j);
        if (f & 2) return fail(ORX_ERR_NONFINITE, "NaN or infinite value not allowed in vector (query %d)", j);
        any_unproven |= (f & 1) != 0;
    }
    if (!out_on_dev) {
        memcpy(out_ids, root->cur->h_ids.p, nk * sizeof(orx_id));
        memcpy(out_dist, root->cur->h_dist.p, nk * sizeof(double));
        memcpy(out_counts, root->cur->h_counts.p, nq * sizeof(int));
    }
    if (any_unproven) {
        std::vector<float> qh(ORX_DIM);
        std::vector<orx_id> ti(k);
        std::vector<double> td(k);
        for (int j = 0; j < nq; ++j) {
            if (!(root->cur->h_flags.p[j] & 1)) continue;
            const float *qj = queries + (size_t)j * ORX_DIM;
            if (is_device_ptr(queries)) {
                CK(cudaMemcpy(qh.data(), qj, ORX_DIM * sizeof(float), cudaMemcpyDeviceToHost));
                qj = qh.data();
            }
            int cj = 0;
            rc = group_exact_query(g, qj, k, ti.data(), td.data(), &cj);
            if (rc != ORX_OK) return rc;
            if (out_on_dev) {
                CK(cudaMemcpy(out_ids + (size_t)j * k, ti.data(), k * sizeof(orx_id), cudaMemcpyHostToDevice));
                CK(cudaMemcpy(out_dist + (size_t)j * k, td.data(), k * sizeof(double), cudaMemcpyHostToDevice));
                CK(cudaMemcpy(out_counts + j, &cj, sizeof(int), cudaMemcpyHostToDevice));
            } else {
                memcpy(out_ids + (size_t)j * k, ti.data(), k * sizeof(orx_id));
                memcpy(out_dist + (size_t)j * k, td.data(), k * sizeof(double));
                out_counts[j] = cj;
            }
        }
    }
    for (int s = 0; s < G; ++s) {
        g->shards[s]->stats.last_path = paths[s];
        if (g->shards[s]->cur->scan_ev_used == 0) continue;       // scan timing is off: nothing to read on that device
        DeviceGuard dg(g->shards[s]->device);
        harvest_scan_events(g->shards[s]);
    }
    g->last_search_ms = std::chrono::duration<float, std::milli>(std::chrono::steady_clock::now() - t_begin).count();
    g->last_path = paths[0];
    g->searches += 1;
    g->queries += nq;
    return ORX_OK;
}

int group_search(Group *g, const float *queries, int nq, int k, orx_id *out_ids, double *out_dist, int *out_counts) {
    std::lock_guard<std::mutex> lk(g->mu);
    if ((size_t)g->shards.size() * k > 1024) return fail(ORX_ERR_INVALID, "multi-GPU search: devices * k must be <= 1024");
    const int qchunk = k <= 32 ? XQ_MAX : XQ_MAX * 32 / k;
    for (int q0 = 0; q0 < nq; q0 += qchunk) {
        const int m = std::min(qchunk, nq - q0);
        int rc = group_search_chunk(g, queries + (size_t)q0 * ORX_DIM, m, k, out_ids + (size_t)q0 * k,
                                    out_dist + (size_t)q0 * k, out_counts + q0);
        if (rc != ORX_OK) return rc;
    }
    return ORX_OK;
}

// `WHERE langchain_id IN (...)` on a multi-GPU index: every shard answers for the ids it owns (its own exact filtered
// search, host buffers), the root merges the G lists.
int group_search_filtered(Group *g, const float *queries, int nq, int k, const orx_id *allow_ids, uint64_t n_allow,
                          orx_id *out_ids, double *out_dist, int *out_counts) {
    std::lock_guard<std::mutex> lk(g->mu);
    const int G = (int)g->shards.size();
    if ((size_t)G * k > 1024) return fail(ORX_ERR_INVALID, "multi-GPU search: devices * k must be <= 1024");
    if (is_device_ptr(queries) || is_device_ptr(out_ids))
        return fail(ORX_ERR_INVALID, "filtered search on a multi-GPU index takes host buffers");
    std::vector<std::vector<orx_id>> allow(G);
    for (uint64_t i = 0; i < n_allow; ++i) allow[shard_of_id(allow_ids[i], G)].push_back(allow_ids[i]);
    const size_t nk = (size_t)nq * k;
    std::vector<orx_id> ids((size_t)G * nk);
    std::vector<double> dist((size_t)G * nk);
    std::vector<int> cnt((size_t)G * nq);
    int rc = group_run(g, [&](int s) {
        return orx_search_filtered(g->shards[s], queries, nq, ORX_DIM, k, allow[s].data(), allow[s].size(),
                                   ids.data() + (size_t)s * nk, dist.data() + (size_t)s * nk, cnt.data() + (size_t)s * nq);
    });
    if (rc != ORX_OK) return rc;
    g->searches += 1;
    g->queries += nq;
    return orx_merge_topk(g->shards[0], G, nq, k, ids.data(), dist.data(), cnt.data(), out_ids, out_dist, out_counts);
}

// ------------------------------------------------------------------------------- writes
int validate_locked(orx_index *ix, const float *vecs, uint64_t n) {      // pgvector's element check, whole batch
    cudaStream_t st = ix->stream;
    const bool on_dev = is_device_ptr(vecs);
    const uint64_t chunk = std::min<uint64_t>(n, STAGE_ROWS);
    CK(ix->d_flag.ensure(1));
    CK(ix->h_flag.ensure(1));
    if (!on_dev) CK(ix->stage.ensure(chunk * ORX_DIM));
    CK(cudaMemsetAsync(ix->d_flag.p, 0, sizeof(int), st));
    for (uint64_t s = 0; s < n; s += chunk) {
        const uint64_t m = std::min(chunk, n - s);
        const float *src = vecs + s * ORX_DIM;
        if (!on_dev) {
            CK(cudaMemcpyAsync(ix->stage.p, src, m * ORX_DIM * sizeof(float), cudaMemcpyHostToDevice, st));
            src = ix->stage.p;
        }
        orx::launch_validate_rows(src, m, ix->d_flag.p, st);
        ix->stats.kernel_launches += 1;
    }
    CK(cudaMemcpyAsync(ix->h_flag.p, ix->d_flag.p, sizeof(int), cudaMemcpyDeviceToHost, st));
    CK(cudaStreamSynchronize(st));
    CK(cudaGetLastError());
    if (*ix->h_flag.p) return fail(ORX_ERR_NONFINITE, "NaN or infinite value not allowed in vector");
    return ORX_OK;
}

int group_upsert(Group *g, const orx_id *ids, const float *vecs, uint64_t n) {
    std::lock_guard<std::mutex> lk(g->mu);
    const int G = (int)g->shards.size();
    const bool on_dev = is_device_ptr(vecs);
    {   // all-or-nothing: the whole batch is checked before any shard writes
        orx_index *root = g->shards[0];
        std::lock_guard<std::mutex> lr(root->mu);
        DeviceGuard dg(root->device);
        int rc = validate_locked(root, vecs, n);
        if (rc != ORX_OK) return rc;
    }
    std::vector<std::vector<uint32_t>> idx(G);
    for (uint64_t i = 0; i < n; ++i) idx[shard_of_id(ids[i], G)].push_back((uint32_t)i);
    return group_run(g, [&](int s) {
        orx_index *ix = g->shards[s];
        const std::vector<uint32_t> &sel = idx[s];
        if (sel.empty()) return (int)ORX_OK;
        std::lock_guard<std::mutex> ls(ix->mu);
        cudaStream_t st = ix->stream;
        std::vector<orx_id> sub_ids;
        std::vector<float> sub_host;
        for (size_t off = 0; off < sel.size(); off += STAGE_ROWS) {
            const size_t m = std::min<size_t>(STAGE_ROWS, sel.size() - off);
            sub_ids.resize(m);
            for (size_t i = 0; i < m; ++i) sub_ids[i] = ids[sel[off + i]];
            const float *rows = nullptr;
            if (on_dev) {
                // gather this shard's rows out of the batch (which may live on another GPU: read over NVLink)
                CK(ix->fb_dist.ensure(m * ORX_DIM / 2));
                CK(ix->d_src_idx.ensure(m));
                CK(cudaMemcpyAsync(ix->d_src_idx.p, sel.data() + off, m * sizeof(uint32_t), cudaMemcpyHostToDevice, st));
                float *dst = reinterpret_cast<float *>(ix->fb_dist.p);
                orx::launch_gather_rows(ORX_DTYPE_F32, vecs, ix->d_src_idx.p, (uint32_t)m, dst, st);
                ix->stats.kernel_launches += 1;
                CK(cudaStreamSynchronize(st));          // d_src_idx is reused by the commit that follows
                rows = dst;
            } else {
                sub_host.resize(m * ORX_DIM);
                for (size_t i = 0; i < m; ++i)
                    memcpy(sub_host.data() + i * ORX_DIM, vecs + (size_t)sel[off + i] * ORX_DIM, ORX_DIM * sizeof(float));
                rows = sub_host.data();
            }
            int rc = upsert_locked(ix, sub_ids.data(), rows, m, /*validate=*/false);
            if (rc != ORX_OK) return rc;
        }
        return (int)ORX_OK;
    });
}

int group_delete(Group *g, const orx_id *ids, uint64_t n, uint64_t *n_removed) {
    std::lock_guard<std::mutex> lk(g->mu);
    const int G = (int)g->shards.size();
    std::vector<std::vector<orx_id>> sub(G);
    for (uint64_t i = 0; i < n; ++i) sub[shard_of_id(ids[i], G)].push_back(ids[i]);
    std::vector<uint64_t> removed(G, 0);
    int rc = group_run(g, [&](int s) {
        return sub[s].empty() ? (int)ORX_OK : orx_delete(g->shards[s], sub[s].data(), sub[s].size(), &removed[s]);
    });
    if (n_removed) {
        *n_removed = 0;
        for (uint64_t r : removed) *n_removed += r;
    }
    return rc;
}

int group_fetch(Group *g, const orx_id *ids, uint64_t n, float *out_vecs, int *out_found) {
    std::lock_guard<std::mutex> lk(g->mu);
    const int G = (int)g->shards.size();
    std::vector<std::vector<uint32_t>> idx(G);
    for (uint64_t i = 0; i < n; ++i) idx[shard_of_id(ids[i], G)].push_back((uint32_t)i);
    return group_run(g, [&](int s) {
        const size_t m = idx[s].size();
        if (!m) return (int)ORX_OK;
        std::vector<orx_id> sid(m);
        std::vector<float> v(m * ORX_DIM);
        std::vector<int> f(m);
        for (size_t i = 0; i < m; ++i) sid[i] = ids[idx[s][i]];
        int rc = orx_fetch(g->shards[s], sid.data(), m, v.data(), f.data());
        if (rc != ORX_OK) return rc;
        for (size_t i = 0; i < m; ++i) {
            memcpy(out_vecs + (size_t)idx[s][i] * ORX_DIM, v.data() + i * ORX_DIM, ORX_DIM * sizeof(float));
            out_found[idx[s][i]] = f[i];
        }
        return (int)ORX_OK;
    });
}

// rows of the whole group numbered shard after shard (shard 0's live rows first)
int group_export_rows(Group *g, uint64_t row_start, uint64_t n, orx_id *ids_out, void *rows_out) {
    std::lock_guard<std::mutex> lk(g->mu);
    const size_t rb = ORX_DIM * elem_size(g->dtype);
    uint64_t base = 0, done = 0;
    for (orx_index *s : g->shards) {
        const uint64_t sz = orx_size(s);
        const uint64_t lo = std::max(row_start, base), hi = std::min(row_start + n, base + sz);
        if (lo < hi) {
            int rc = orx_export_rows(s, lo - base, hi - lo, ids_out + (lo - row_start),
                                     static_cast<char *>(rows_out) + (lo - row_start) * rb);
            if (rc != ORX_OK) return rc;
            done += hi - lo;
        }
        base += sz;
    }
    if (done != n) return fail(ORX_ERR_INVALID, "rows [%llu, %llu) exceed the %llu live rows", (unsigned long long)row_start,
                               (unsigned long long)(row_start + n), (unsigned long long)base);
    return ORX_OK;
}

int group_import_rows(Group *g, const orx_id *ids, const void *rows_raw, uint64_t n) {
    std::lock_guard<std::mutex> lk(g->mu);
    if (is_device_ptr(rows_raw)) return fail(ORX_ERR_INVALID, "orx_import_rows on a multi-GPU index takes host rows");
    const int G = (int)g->shards.size();
    const size_t rb = ORX_DIM * elem_size(g->dtype);
    std::vector<std::vector<uint32_t>> idx(G);
    for (uint64_t i = 0; i < n; ++i) idx[shard_of_id(ids[i], G)].push_back((uint32_t)i);
    return group_run(g, [&](int s) {
        const size_t m = idx[s].size();
        if (!m) return (int)ORX_OK;
        std::vector<orx_id> sid(m);
        std::vector<char> rows(m * rb);
        for (size_t i = 0; i < m; ++i) {
            sid[i] = ids[idx[s][i]];
            memcpy(rows.data() + i * rb, static_cast<const char *>(rows_raw) + (size_t)idx[s][i] * rb, rb);
        }
        return orx_import_rows(g->shards[s], sid.data(), rows.data(), m);
    });
}

int group_stats(Group *g, orx_stats *out) {
    std::lock_guard<std::mutex> lk(g->mu);
    orx_stats t{};
    const int G = (int)g->shards.size();
    double scan_ms = 0.0;
    float last_scan = 0.f;
    for (orx_index *s : g->shards) {
        orx_stats a;
        orx_get_stats(s, &a);
        t.kernel_launches += a.kernel_launches;
        t.fallback_gemv += a.fallback_gemv;
        t.fallback_exhaustive += a.fallback_exhaustive;
        t.rows_moved += a.rows_moved;
        scan_ms += a.scan_ms_total;
        last_scan = std::max(last_scan, a.last_scan_ms);
        t.scan_launches = std::max(t.scan_launches, a.scan_launches);
    }
    t.scan_ms_total = scan_ms / G;            // mean over the shards of the per-shard scan time (they run in parallel)
    t.last_scan_ms = last_scan;
    t.searches = g->searches;
    t.queries = g->queries;
    t.last_search_ms = g->last_search_ms;
    t.last_path = g->last_path;
    *out = t;
    return ORX_OK;
}

#endif
