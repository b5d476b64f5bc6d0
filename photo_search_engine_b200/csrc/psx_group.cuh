// psx_group.cuh -- one handle, G devices, ONE process ("one Python process, G devices ... keeps the
// single-VectorStore-instance contract", main.py:59-68 constructs one store that core/searcher.py and
// core/indexer.py share).  Not a standalone header: included at the end of psx_api.cu, whose static helpers
// (launch_scan, launch_batch, launch_query, upload_host_rows, read_rows_locked ...) it drives per child.
//
// Layout.  The group owns one complete single-device index ("child") per device entry and no rows itself.  Rows are
// split in contiguous id ranges: child s holds the rows [s*quota, (s+1)*quota) that are in HBM, the last child also
// everything beyond; global id = first row of the child + local row (`id_base` inside the kernels), so results are
// bit-identical to one device holding everything.  Appends land on the tail; when the tail child outgrows the others by
// 25 % the layout is re-split evenly with device-to-device copies (amortised like arena doubling).
//
// A query.  The host-pinned query goes to every child (one async copy each, each on the child's own stream), every
// child scans its rows, and the k best of every shard meet on the HOME device (child 0):
//   * single queries on fp32 / bf16 stores: the exchange is fused into the kernels -- the scan's last CTA stores its k
//     keys straight into home's receive buffer over NVLink peer mappings and raises a flag there; a one-CTA kernel on
//     home waits for the flags and selects the global top-k (K4; no event, no collective, no host round trip);
//   * query batches (tensor-core path per shard), k > PSX_K_PASS_MAX paging, the bf16+fp32-master tier and groups of
//     more than 8 shards: every child writes its sorted key lists into home's list buffer (peer stores from the
//     kernels' epilogues), stream events order home's merge kernel behind them.
// If the fused wait times out (a device that never publishes) the status word says so and the query is re-run over the
// event-ordered path -- the process and the CUDA context stay healthy.
#pragma once

static long long group_base(const psx_index* g, size_t s) {
    long long b = 0;
    for (size_t i = 0; i < s; ++i) b += g->shards[i]->n;
    return b;
}

static int group_enable_peer(int from_dev, int to_dev) {
    if (from_dev == to_dev) return PSX_OK;
    int can = 0;
    CU(cudaDeviceCanAccessPeer(&can, from_dev, to_dev));
    if (!can) return fail(PSX_ERR_CUDA, "device %d cannot map the memory of device %d (no peer access): a multi-device handle needs NVLink / PCIe P2P", from_dev, to_dev);
    DeviceGuard g(from_dev);
    cudaError_t e = cudaDeviceEnablePeerAccess(to_dev, 0);
    if (e == cudaErrorPeerAccessAlreadyEnabled) {
        cudaGetLastError();
        e = cudaSuccess;
    }
    if (e != cudaSuccess) return fail(PSX_ERR_CUDA, "cudaDeviceEnablePeerAccess(%d -> %d): %s", from_dev, to_dev, cudaGetErrorString(e));
    return PSX_OK;
}

static void group_destroy_parts(psx_index* g) {
    for (psx_index* c : g->shards) psx_destroy(c);
    g->shards.clear();
    DeviceGuard dg(g->device);
    cudaFree(g->gx);
    cudaFree(g->gkeys);
    cudaFreeHost(g->g_hq);
    cudaFreeHost(g->xstatus_host);
    g->xstatus_host = nullptr;
    for (cudaEvent_t& e : g->g_merged)
        if (e) cudaEventDestroy(e);
    for (cudaEvent_t e : g->g_done)
        if (e) cudaEventDestroy(e);
    for (cudaEvent_t e : g->g_h2d)
        if (e) cudaEventDestroy(e);
}

extern "C" int psx_create_sharded(int d, int metric, int store_dtype, int n_devices, const int* devices, psx_index** out) {
    if (!out) return fail(PSX_ERR_INVALID, "out is null");
    *out = nullptr;
    if (n_devices < 1 || n_devices > 64 || !devices) return fail(PSX_ERR_INVALID, "need 1..64 devices");
    if (n_devices == 1) return psx_create(d, metric, store_dtype, devices[0], out);
    psx_index* g = new (std::nothrow) psx_index();
    if (!g) return fail(PSX_ERR_OOM, "host allocation failed");
    g->d = d;
    g->metric = metric;
    g->dtype = store_dtype;
    g->device = devices[0];
    int rc = PSX_OK;
    for (int i = 0; i < n_devices && rc == PSX_OK; ++i) {
        psx_index* c = nullptr;
        rc = psx_create(d, metric, store_dtype, devices[i], &c);
        if (rc == PSX_OK) g->shards.push_back(c);
    }
    auto init = [&]() -> int {
        // every child's kernels store into (and, for page ceilings, read from) home's memory
        for (psx_index* c : g->shards) {
            int r = group_enable_peer(c->device, g->device);
            if (r) return r;
        }
        DeviceGuard dg(g->device);
        CU(cudaMalloc(&g->gx, (size_t)psx_exchange_bytes()));
        CU(cudaMemset(g->gx, 0, (size_t)psx_exchange_bytes()));
        for (cudaEvent_t& e : g->g_merged) CU(cudaEventCreateWithFlags(&e, cudaEventDisableTiming));
        g->g_done.assign(g->shards.size(), nullptr);
        g->g_h2d.assign(g->shards.size(), nullptr);
        g->g_h2d_busy.assign(g->shards.size(), 0);
        for (size_t s = 0; s < g->shards.size(); ++s) {
            DeviceGuard cg(g->shards[s]->device);
            CU(cudaEventCreateWithFlags(&g->g_done[s], cudaEventDisableTiming));
            CU(cudaEventCreateWithFlags(&g->g_h2d[s], cudaEventDisableTiming));
        }
        int* unused = nullptr;
        return ensure_xstatus(g, &unused);
    };
    if (rc == PSX_OK) rc = init();
    if (rc != PSX_OK) {
        const std::string keep = g_err;
        group_destroy_parts(g);
        delete g;
        g_err = keep;
        return rc;
    }
    *out = g;
    return PSX_OK;
}

extern "C" int psx_device_count(const psx_index* h) { return h ? (is_group(h) ? (int)h->shards.size() : 1) : 0; }

extern "C" int psx_shard_rows(psx_index* h, int64_t* rows, int* devices, int capacity) {
    if (!h || capacity < psx_device_count(h)) return fail(PSX_ERR_INVALID, "bad arguments to psx_shard_rows");
    std::lock_guard<std::mutex> lk(h->mu);
    if (!is_group(h)) {
        if (rows) rows[0] = h->n;
        if (devices) devices[0] = h->device;
        return PSX_OK;
    }
    for (size_t s = 0; s < h->shards.size(); ++s) {
        if (rows) rows[s] = h->shards[s]->n;
        if (devices) devices[s] = h->shards[s]->device;
    }
    return PSX_OK;
}

extern "C" int psx_group_stats(psx_index* h, int64_t* fused, int64_t* keyed, int64_t* timeouts) {
    if (!h) return fail(PSX_ERR_INVALID, "null handle");
    if (fused) *fused = h->g_fused;
    if (keyed) *keyed = h->g_keyed;
    if (timeouts) *timeouts = h->g_timeouts;
    return PSX_OK;
}

static int group_reset(psx_index* g) {
    for (psx_index* c : g->shards) {
        int rc = psx_reset(c);
        if (rc) return rc;
    }
    g->n = 0;
    g->quota = 0;
    std::lock_guard<std::mutex> pk(g->pmu);
    g->pending.clear();
    g->pending.shrink_to_fit();
    g->pending_n = 0;
    return PSX_OK;
}

// ---- layout -------------------------------------------------------------------------------------------------------
static long long group_even_quota(const psx_index* g, long long total) {
    const long long G = (long long)g->shards.size();
    return std::max<long long>((total + G - 1) / G, std::max<long long>(1, g->shard_min_rows));
}

// Re-split the rows in HBM so that child s holds [s*q, (s+1)*q) (the last child: everything from (G-1)*q on).
// Device-to-device copies into fresh arenas; the largest stored row norm (certificate bound) becomes the maximum
// over all children on every child.
static int group_rebalance(psx_index* g, long long q) {
    const size_t G = g->shards.size();
    const long long N = g->n;
    std::vector<long long> ob(G + 1, 0), nb(G + 1, 0);
    for (size_t s = 0; s < G; ++s) ob[s + 1] = ob[s] + g->shards[s]->n;
    for (size_t s = 0; s <= G; ++s) nb[s] = s == G ? N : std::min<long long>((long long)s * q, N);
    bool same = true;
    for (size_t s = 0; s <= G; ++s) same = same && ob[s] == nb[s];
    g->quota = q;
    if (same) return PSX_OK;
    struct Fresh {
        unsigned char *x = nullptr, *xm = nullptr;
        uint64_t* attrs = nullptr;
        long long cap = 0;
    };
    std::vector<Fresh> fresh(G);
    auto drop = [&]() {
        for (size_t s = 0; s < G; ++s) {
            DeviceGuard dg(g->shards[s]->device);
            cudaFree(fresh[s].x);
            cudaFree(fresh[s].xm);
            cudaFree(fresh[s].attrs);
        }
    };
    float max_sumsq = 0.f;
    bool attrs_set = false;
    for (size_t s = 0; s < G; ++s) {
        psx_index* c = g->shards[s];
        DeviceGuard dg(c->device);
        cudaStreamSynchronize(c->stream);
        if (c->has_last) cudaEventSynchronize(c->last_ev);
        float m = 0.f;
        if (cudaMemcpy(&m, c->dmax_sumsq, sizeof(float), cudaMemcpyDeviceToHost) == cudaSuccess) max_sumsq = std::max(max_sumsq, m);
        attrs_set = attrs_set || c->attrs_set;
        const long long rows = nb[s + 1] - nb[s];
        const long long cap = std::max<long long>(std::max(rows, s + 1 < G ? q : rows), 1024);
        cudaError_t e = cudaMalloc(&fresh[s].x, (size_t)cap * c->row_bytes);
        if (e == cudaSuccess) e = cudaMalloc(&fresh[s].attrs, (size_t)cap * sizeof(uint64_t));
        if (e == cudaSuccess && c->mrow_bytes) e = cudaMalloc(&fresh[s].xm, (size_t)cap * c->mrow_bytes);
        if (e == cudaSuccess) e = cudaMemsetAsync(fresh[s].attrs, 0, (size_t)cap * sizeof(uint64_t), c->stream);
        if (e != cudaSuccess) {
            cudaGetLastError();
            drop();
            return fail(PSX_ERR_OOM, "re-splitting the corpus over %zu devices: allocation of %lld rows on device %d failed: %s", G, cap,
                        c->device, cudaGetErrorString(e));
        }
        fresh[s].cap = cap;
        for (size_t t = 0; t < G; ++t) {  // pieces of the old children that now belong to child s
            const long long lo = std::max(nb[s], ob[t]), hi = std::min(nb[s + 1], ob[t + 1]);
            if (lo >= hi) continue;
            const psx_index* o = g->shards[t];
            const size_t doff = (size_t)(lo - nb[s]), soff = (size_t)(lo - ob[t]), cnt = (size_t)(hi - lo);
            e = cudaMemcpyPeerAsync(fresh[s].x + doff * c->row_bytes, c->device, o->x + soff * o->row_bytes, o->device, cnt * c->row_bytes, c->stream);
            if (e == cudaSuccess)
                e = cudaMemcpyPeerAsync(fresh[s].attrs + doff, c->device, o->attrs + soff, o->device, cnt * sizeof(uint64_t), c->stream);
            if (e == cudaSuccess && c->mrow_bytes)
                e = cudaMemcpyPeerAsync(fresh[s].xm + doff * c->mrow_bytes, c->device, o->xm + soff * o->mrow_bytes, o->device, cnt * c->mrow_bytes,
                                        c->stream);
            if (e != cudaSuccess) {
                cudaGetLastError();
                for (psx_index* w : g->shards) {
                    DeviceGuard wg(w->device);
                    cudaStreamSynchronize(w->stream);
                }
                drop();
                return fail(PSX_ERR_CUDA, "re-splitting the corpus: peer copy %d -> %d failed: %s", o->device, c->device, cudaGetErrorString(e));
            }
        }
    }
    for (size_t s = 0; s < G; ++s) {
        psx_index* c = g->shards[s];
        DeviceGuard dg(c->device);
        cudaError_t e = cudaStreamSynchronize(c->stream);
        if (e != cudaSuccess) {
            drop();
            return fail(PSX_ERR_CUDA, "re-splitting the corpus: %s", cudaGetErrorString(e));
        }
    }
    for (size_t s = 0; s < G; ++s) {
        psx_index* c = g->shards[s];
        DeviceGuard dg(c->device);
        cudaFree(c->x);
        cudaFree(c->xm);
        cudaFree(c->attrs);
        c->x = fresh[s].x;
        c->xm = fresh[s].xm;
        c->attrs = fresh[s].attrs;
        c->cap = fresh[s].cap;
        c->n = nb[s + 1] - nb[s];
        c->attrs_set = attrs_set;
        c->max_norm = 0.f;
        cudaMemcpy(c->dmax_sumsq, &max_sumsq, sizeof(float), cudaMemcpyHostToDevice);
    }
    return PSX_OK;
}

// Decide the layout for `total` rows (the rows in HBM plus the ones about to be appended): keep the current split
// while the tail child stays within 25 % of the others, else re-split evenly.
static int group_prepare_layout(psx_index* g, long long total) {
    const long long G = (long long)g->shards.size();
    if (g->quota <= 0 || g->n == 0) {
        g->quota = std::max(g->quota, group_even_quota(g, total));
        return PSX_OK;
    }
    if (total * 4 > (4 * G + 1) * g->quota) return group_rebalance(g, group_even_quota(g, total));
    return PSX_OK;
}

// child that receives global row `gi` under the current layout, and how many consecutive rows fit there
static size_t group_target(const psx_index* g, long long gi, long long want, long long* room) {
    const long long G = (long long)g->shards.size();
    long long s = g->quota > 0 ? gi / g->quota : 0;
    if (s >= G - 1) {
        *room = want;
        return (size_t)(G - 1);
    }
    *room = std::min(want, (s + 1) * g->quota - gi);
    return (size_t)s;
}

static int group_place_host_rows(psx_index* g, const float* rows, long long rows_n) {
    int rc = group_prepare_layout(g, g->n + rows_n);
    if (rc) return rc;
    long long done = 0;
    while (done < rows_n) {
        long long room = 0;
        psx_index* c = g->shards[group_target(g, g->n + done, rows_n - done, &room)];
        DeviceGuard dg(c->device);
        if ((rc = upload_host_rows(c, rows + (size_t)done * g->d, room))) {
            // flush_pending keeps the whole batch staged: roll the children back to the rows that were in HBM before
            // this call (the rows placed so far are the highest ids, i.e. they sit at the tail of the layout)
            long long extra = done;
            for (size_t s = g->shards.size(); s-- > 0 && extra > 0;) {
                const long long take = std::min<long long>(extra, g->shards[s]->n);
                g->shards[s]->n -= take;
                extra -= take;
            }
            return rc;
        }
        c->n += room;
        done += room;
    }
    return PSX_OK;
}

static int group_reserve(psx_index* g, long long n) {
    const long long q = group_even_quota(g, n);
    int rc = PSX_OK;
    if (g->n > 0 && q > g->quota) rc = group_rebalance(g, q);
    if (rc) return rc;
    g->quota = std::max(g->quota, q);
    for (size_t s = 0; s < g->shards.size(); ++s) {
        psx_index* c = g->shards[s];
        // rows this child will hold when n rows are stored
        const long long lo = std::min<long long>((long long)s * g->quota, n);
        const long long hi = s + 1 == g->shards.size() ? n : std::min<long long>((long long)(s + 1) * g->quota, n);
        if (hi <= lo) continue;
        DeviceGuard dg(c->device);
        if ((rc = ensure_capacity(c, hi - lo, true))) return rc;
    }
    return PSX_OK;
}

static int pointer_device(const void* p, int fallback) {
    cudaPointerAttributes a;
    if (cudaPointerGetAttributes(&a, p) != cudaSuccess) {
        cudaGetLastError();
        return fallback;
    }
    return a.type == cudaMemoryTypeDevice || a.type == cudaMemoryTypeManaged ? a.device : fallback;
}

static int group_add_device(psx_index* g, const float* x_dev, long long n, int normalize, cudaStream_t st) {
    if ((unsigned long long)(g->n + n) >= 0xffffffffull) return fail(PSX_ERR_RANGE, "more than 2^32-1 rows");
    const int src_dev = pointer_device(x_dev, g->device);
    {   // the caller's stream produced the rows: everything below runs on the children's own streams
        DeviceGuard dg(src_dev);
        CU(cudaStreamSynchronize(st));
    }
    int rc = group_prepare_layout(g, g->n + n);
    if (rc) return rc;
    long long done = 0;
    while (done < n) {
        long long room = 0;
        psx_index* c = g->shards[group_target(g, g->n, n - done, &room)];
        DeviceGuard dg(c->device);
        if ((rc = ensure_capacity(c, c->n + room, false))) break;
        const float* src = x_dev + (size_t)done * g->d;
        if (c->device == src_dev) {
            if ((rc = launch_pack(c, src, c->n, room, normalize, c->stream))) break;
            CU(cudaStreamSynchronize(c->stream));
        } else {
            const long long chunk = std::max<long long>(1, (256ll << 20) / ((long long)g->d * 4));
            float* bounce = nullptr;
            const long long brows = std::min(chunk, room);
            CU(cudaMalloc(&bounce, (size_t)brows * g->d * sizeof(float)));
            for (long long off = 0; off < room && rc == PSX_OK; off += brows) {
                const long long m = std::min(brows, room - off);
                cudaError_t e = cudaMemcpyPeerAsync(bounce, c->device, src + (size_t)off * g->d, src_dev, (size_t)m * g->d * sizeof(float), c->stream);
                if (e != cudaSuccess) rc = fail(PSX_ERR_CUDA, "peer copy %d -> %d failed: %s", src_dev, c->device, cudaGetErrorString(e));
                if (rc == PSX_OK) rc = launch_pack(c, bounce, c->n + off, m, normalize, c->stream);
                if (rc == PSX_OK && cudaStreamSynchronize(c->stream) != cudaSuccess) rc = fail(PSX_ERR_CUDA, "bulk ingest failed");
            }
            cudaFree(bounce);
            if (rc) break;
        }
        c->n += room;
        g->n += room;
        done += room;
    }
    return rc;
}

static int group_set_attrs(psx_index* g, long long row0, const uint64_t* attrs, long long n, bool from_device, cudaStream_t st) {
    if (row0 + n > g->n) return fail(PSX_ERR_RANGE, "attribute rows [%lld,%lld) exceed ntotal %lld", row0, row0 + n, g->n.load());
    int src_dev = g->device;
    if (from_device) {
        src_dev = pointer_device(attrs, g->device);
        DeviceGuard dg(src_dev);
        CU(cudaStreamSynchronize(st));
    }
    long long base = 0;
    for (psx_index* c : g->shards) {
        const long long lo = std::max(row0, base), hi = std::min(row0 + n, base + c->n);
        if (lo < hi) {
            DeviceGuard dg(c->device);
            if (from_device)
                CU(cudaMemcpyPeerAsync(c->attrs + (lo - base), c->device, attrs + (lo - row0), src_dev, (size_t)(hi - lo) * sizeof(uint64_t), c->stream));
            else
                CU(cudaMemcpyAsync(c->attrs + (lo - base), attrs + (lo - row0), (size_t)(hi - lo) * sizeof(uint64_t), cudaMemcpyHostToDevice, c->stream));
            CU(cudaStreamSynchronize(c->stream));
        }
        c->attrs_set = true;  // every child: a predicate must find words (zero = "no EXIF") on all of them
        base += c->n;
    }
    g->attrs_set = true;
    return PSX_OK;
}

static int group_read_rows(psx_index* g, long long row0, long long n, float* out) {
    long long base = 0;
    for (psx_index* c : g->shards) {
        const long long lo = std::max(row0, base), hi = std::min(row0 + n, base + c->n);
        if (lo < hi) {
            DeviceGuard dg(c->device);
            int rc = read_rows_locked(c, lo - base, hi - lo, out + (size_t)(lo - row0) * g->d);
            if (rc) return rc;
        }
        base += c->n;
    }
    return PSX_OK;
}

static int group_set_tunable(psx_index* g, const char* key, int value) {
    if (!strcmp(key, "shard_min_rows")) {
        g->shard_min_rows = value <= 0 ? 8192 : value;
        return PSX_OK;
    }
    if (!strcmp(key, "xchg_timeout_ms")) return PSX_OK;  // a property of the merging side, kept on the group itself
    for (psx_index* c : g->shards) {
        int rc = psx_set_tunable(c, key, value);
        if (rc) return rc;
    }
    return PSX_OK;
}

// ---- search -------------------------------------------------------------------------------------------------------
struct GroupActive {
    psx_index* c;
    uint32_t id_base;
    size_t shard;
};

static int group_ensure_keys(psx_index* g, size_t words) {
    if (words <= g->gkeys_cap) return PSX_OK;
    DeviceGuard dg(g->device);
    if (g->gkeys) {
        // kernels of any child may still be storing into the old buffer
        for (psx_index* c : g->shards) {
            DeviceGuard cg(c->device);
            CU(cudaStreamSynchronize(c->stream));
        }
        cudaFree(g->gkeys);
        g->gkeys = nullptr;
        g->gkeys_cap = 0;
    }
    CU(cudaMalloc(&g->gkeys, words * sizeof(uint64_t)));
    g->gkeys_cap = words;
    return PSX_OK;
}

// Queries [q0, q0+gq): ONE copy from the pinned staging to the home device; every other child reads home's copy through
// its peer mapping (a query is 4 KB per CTA: nothing next to the rows it scans) and only orders its stream behind the
// copy with an event.  Per query and child that is one cudaStreamWaitEvent instead of a memcpy + an event record.
static int group_send_queries(psx_index* g, const std::vector<GroupActive>& act, const float* q, int64_t gq) {
    psx_index* home = g->shards[0];
    const size_t floats = (size_t)gq * g->d;
    if (floats > g->g_hq_cap) {
        cudaFreeHost(g->g_hq);
        g->g_hq = nullptr;
        g->g_hq_cap = 0;
        CU(cudaHostAlloc(&g->g_hq, floats * sizeof(float), cudaHostAllocPortable));
        g->g_hq_cap = floats;
    }
    // the staging may still be the source of the previous step's copy (paged calls issue several steps per sync)
    if (g->g_h2d_busy[0]) {
        CU(cudaEventSynchronize(g->g_h2d[0]));
        g->g_h2d_busy[0] = 0;
    }
    memcpy(g->g_hq, q, floats * sizeof(float));
    {
        DeviceGuard dg(home->device);
        int rc = ensure_io(home, floats, 1);
        if (rc) return rc;
        CU(cudaMemcpyAsync(home->dq, g->g_hq, floats * sizeof(float), cudaMemcpyHostToDevice, home->stream));
        CU(cudaEventRecord(g->g_h2d[0], home->stream));
        g->g_h2d_busy[0] = 1;
    }
    for (const GroupActive& a : act) {
        if (a.c == home) continue;
        DeviceGuard dg(a.c->device);
        CU(cudaStreamWaitEvent(a.c->stream, g->g_h2d[0], 0));
    }
    return PSX_OK;
}

// home's merge of `nq` queries x act.size() lists in g->gkeys, ordered behind every child's stream by events
static int group_merge_lists(psx_index* g, const std::vector<GroupActive>& act, int64_t nq, int kp, float* out_scores, long long* out_ids,
                             uint64_t* out_keys) {
    psx_index* home = g->shards[0];
    for (const GroupActive& a : act) {
        if (a.c == home) continue;
        DeviceGuard dg(a.c->device);
        CU(cudaEventRecord(g->g_done[a.shard], a.c->stream));
    }
    DeviceGuard dg(home->device);
    for (const GroupActive& a : act)
        if (a.c != home) CU(cudaStreamWaitEvent(home->stream, g->g_done[a.shard], 0));
    const int kpad = (int)psx_kpad(kp);
    const int nlists = (int)act.size();
    int cap_lists = (int)std::min<int64_t>(nlists + 1, (200 * 1024 / 8) / kpad);
    if (cap_lists < 2) cap_lists = 2;
    const size_t smem = (size_t)cap_lists * kpad * 8;
    merge_keys_kernel<<<(unsigned)nq, 256, smem, home->stream>>>(g->gkeys, nlists, kp, kpad, cap_lists, g->metric, out_scores, out_ids, out_keys);
    g_launches++;
    CU(cudaGetLastError());
    return PSX_OK;
}

// One step of the event-ordered path: every active child writes the sorted keys of queries [0, gq) (page `pg` of them)
// into home's list buffer, home merges.  `ceil_base` (home memory, or nullptr): per query, the key every hit must be
// strictly below (the last key of the previous page).
static int group_keyed_step(psx_index* g, const std::vector<GroupActive>& act, int64_t gq, int kp, const psx_filter* filter, bool batched,
                            const uint64_t* ceil_base, size_t ceil_stride, float* out_scores, long long* out_ids, uint64_t* out_keys) {
    psx_index* home = g->shards[0];
    const int kpad = (int)psx_kpad(kp);
    const size_t A = act.size();
    int rc = group_ensure_keys(g, (size_t)gq * A * kpad);
    if (rc) return rc;
    if (ceil_base) {  // the ceilings were written by home's previous merge
        DeviceGuard dg(home->device);
        CU(cudaEventRecord(g->g_merged[0], home->stream));
        for (const GroupActive& a : act) {
            if (a.c == home) continue;
            DeviceGuard cg(a.c->device);
            CU(cudaStreamWaitEvent(a.c->stream, g->g_merged[0], 0));
        }
    }
    std::vector<char> pending_flags(A, 0);
    for (size_t i = 0; i < A; ++i) {
        psx_index* c = act[i].c;
        DeviceGuard dg(c->device);
        uint64_t* slot0 = g->gkeys + i * kpad;  // query qi's list of this child: slot0 + qi * A * kpad
        if (batched && !ceil_base && batch_shape_ok(c, kp)) {
            if ((rc = ensure_batch_scratch(c, 0))) return rc;
            if ((rc = launch_batch(c, home->dq, (int)gq, kp, filter, act[i].id_base, 0.f, nullptr, nullptr, slot0, c->bflags, c->stream,
                                   (long long)(A * kpad))))
                return rc;
            CU(cudaMemcpyAsync(c->hflags, c->bflags, (size_t)gq * sizeof(int), cudaMemcpyDeviceToHost, c->stream));
            pending_flags[i] = 1;
            c->batch_queries += gq;
        } else {
            for (int64_t qi = 0; qi < gq; ++qi) {
                const uint64_t* ceil_ptr = ceil_base ? ceil_base + (size_t)qi * ceil_stride : nullptr;
                if ((rc = launch_query(c, home->dq + (size_t)qi * g->d, kp, filter, act[i].id_base, ceil_ptr, nullptr, nullptr,
                                       slot0 + (size_t)qi * A * kpad, c->stream)))
                    return rc;
            }
        }
    }
    // certificates of the tensor-core path: the unproven queries are re-run by that child's scan into the same slots
    for (size_t i = 0; i < A; ++i) {
        if (!pending_flags[i]) continue;
        psx_index* c = act[i].c;
        DeviceGuard dg(c->device);
        CU(cudaStreamSynchronize(c->stream));
        for (int64_t qi = 0; qi < gq; ++qi) {
            if (!c->hflags[qi]) continue;
            c->batch_fallbacks++;
            if ((rc = launch_exact_scan(c, home->dq + (size_t)qi * g->d, kp, filter, act[i].id_base, nullptr, nullptr, nullptr,
                                        g->gkeys + ((size_t)qi * A + i) * kpad, c->stream)))
                return rc;
        }
    }
    g->g_keyed += gq;
    return group_merge_lists(g, act, gq, kp, out_scores, out_ids, out_keys);
}

// One query through the fused exchange: scan + publish on every child, wait + merge on home.
static int group_fused_query(psx_index* g, const std::vector<GroupActive>& act, int64_t qi, int kp, const psx_filter* filter, float* out_scores,
                             long long* out_ids) {
    psx_index* home = g->shards[0];
    const int A = (int)act.size();
    if (++g->gseq == 0) {
        g->gseq = 1;
        g->g_synced_seq = 0;
    }
    const uint32_t seq = g->gseq;
    const uint64_t bases[1] = {(uint64_t)(uintptr_t)g->gx};
    int rc;
    // slot (seq & 1) of the receive buffer was last used by query seq - 2: its merge must be over before anyone overwrites it
    // (home's own stream is ordered anyway; after a host synchronisation the event is long complete)
    const bool inline_merge = g->xchg_inline;
    for (int i = A - 1; i >= 0; --i) {  // home last: its last CTA also merges
        psx_index* c = act[i].c;
        DeviceGuard dg(c->device);
        // (nothing to wait for when the host has synchronised with home since that merge -- the one-query-per-call case)
        if (c != home && seq >= 3 && seq - 2 > g->g_synced_seq) CU(cudaStreamWaitEvent(c->stream, g->g_merged[seq & 1u], 0));
        XchgArgs xa{A, i, seq, bases, 1};
        if (g->fault_skip_publish == (int)act[i].shard) xa.targets = -1;  // test hook: this shard stays silent
        if (c == home && inline_merge) {
            // the merging side fused into home's scan: its last CTA waits for the A lists and selects the global top-k
            xa.my_base = bases[0];
            xa.out_scores = out_scores;
            xa.out_ids = out_ids;
            xa.status = g->xstatus_dev;
            xa.spin_limit = xchg_spin_limit(g->xchg_timeout_ms);
        }
        if ((rc = launch_scan(c, home->dq + (size_t)qi * g->d, kp, filter, act[i].id_base, nullptr, nullptr, nullptr, nullptr, c->stream, &xa)))
            return rc;
    }
    DeviceGuard dg(home->device);
    if (!inline_merge) {
        const int kpad = (int)psx_kpad(kp);
        int np = kpad;
        while (np < A * kpad) np <<= 1;
        static std::atomic<bool> ready[64];
        if (home->device < 64 && !ready[home->device].load()) {
            CU(cudaFuncSetAttribute(merge_wait_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, PSX_SMEM_LIMIT - 1024));
            ready[home->device].store(true);
        }
        merge_wait_kernel<<<1, 256, (size_t)np * 8, home->stream>>>((const uint64_t*)g->gx, (const uint32_t*)(g->gx + xchg_flag_offset()), A, seq, kp,
                                                                   kpad, np, g->metric, out_scores, out_ids, nullptr, g->xstatus_dev,
                                                                   xchg_spin_limit(g->xchg_timeout_ms));
        g_launches++;
        CU(cudaGetLastError());
    }
    CU(cudaEventRecord(g->g_merged[seq & 1u], home->stream));
    g->g_fused++;
    return PSX_OK;
}

static int group_search(psx_index* g, const float* q, int64_t nq, int64_t k, const psx_filter* filter, float* out_scores, int64_t* out_ids) {
    const float empty = g->metric == PSX_METRIC_L2 ? INFINITY : -INFINITY;
    const long long N = g->n;
    if (N == 0) {
        for (int64_t i = 0; i < nq * k; ++i) {
            out_scores[i] = empty;
            out_ids[i] = -1;
        }
        return PSX_OK;
    }
    std::vector<GroupActive> act;
    for (size_t s = 0; s < g->shards.size(); ++s)
        if (g->shards[s]->n > 0) act.push_back({g->shards[s], (uint32_t)group_base(g, s), s});
    psx_index* home = g->shards[0];
    if (act.size() == 1) {  // everything lives on one device (a corpus below shard_min_rows): no exchange at all
        DeviceGuard dg(act[0].c->device);
        return search_single(act[0].c, q, nq, k, filter, act[0].id_base, out_scores, out_ids);
    }
    int rc;
    const int64_t kk = std::min<int64_t>(k, N);
    const int64_t pages = (kk + PSX_K_PASS_MAX - 1) / PSX_K_PASS_MAX;
    const int64_t kslot = pages == 1 ? psx_kpad(kk) : pages * PSX_K_PASS_MAX;
    const bool batched = g->batch_min > 0 && nq >= g->batch_min && pages == 1;
    const bool fused = !batched && pages == 1 && g->dtype != PSX_STORE_BF16_MASTER && act.size() <= PSX_XCHG_MAX_WORLD;
    const int64_t group = std::max<int64_t>(1, std::min<int64_t>(std::min<int64_t>(nq, BATCH_MAX_Q), (8ll << 20) / kslot));
    for (const GroupActive& a : act) a.c->call_first = true;  // (children only ever use their own stream: no cross-stream event)
    {
        DeviceGuard dg(home->device);
        if ((rc = ensure_io(home, (size_t)group * g->d, (size_t)group * kslot))) return rc;
    }
    for (int64_t q0 = 0; q0 < nq; q0 += group) {
        const int64_t gq = std::min(group, nq - q0);
        if ((rc = group_send_queries(g, act, q + q0 * g->d, gq))) return rc;
        bool redo_keyed = false;
        if (fused) {
            *g->xstatus_host = 0;
            for (int64_t qi = 0; qi < gq; ++qi)
                if ((rc = group_fused_query(g, act, qi, (int)kk, filter, home->dscores + qi * kslot, home->dids + qi * kslot))) return rc;
        } else {
            for (int64_t pg = 0; pg < pages; ++pg) {
                const int kp = (int)std::min<int64_t>(PSX_K_PASS_MAX, kk - pg * PSX_K_PASS_MAX);
                // page layout of home's outputs: query qi at qi * kslot, page pg at + pg * PSX_K_PASS_MAX (as in search_single).
                // merge_keys_kernel writes query qi at qi * kp / qi * kpad, so paged calls go one query at a time.
                if (pages == 1) {
                    if ((rc = group_keyed_step(g, act, gq, kp, filter, batched, nullptr, 0, home->dscores, home->dids, nullptr))) return rc;
                } else {
                    for (int64_t qi = 0; qi < gq; ++qi) {
                        const size_t off = (size_t)qi * kslot + (size_t)pg * PSX_K_PASS_MAX;
                        std::vector<GroupActive> one = act;
                        // the children read query qi: shift their query pointer by sending it alone
                        if ((rc = group_send_queries(g, act, q + (q0 + qi) * g->d, 1))) return rc;
                        if ((rc = group_keyed_step(g, one, 1, kp, filter, false, pg ? home->dkeys + off - 1 : nullptr, 0, home->dscores + off,
                                                   home->dids + off, home->dkeys + off)))
                            return rc;
                    }
                }
            }
        }
        DeviceGuard dg(home->device);
        const size_t per_q = pages == 1 && !fused ? (size_t)kk : (size_t)kslot;  // stride of the merged outputs on home
        CU(cudaMemcpyAsync(home->hscores, home->dscores, (size_t)gq * per_q * sizeof(float), cudaMemcpyDeviceToHost, home->stream));
        CU(cudaMemcpyAsync(home->hids, home->dids, (size_t)gq * per_q * sizeof(long long), cudaMemcpyDeviceToHost, home->stream));
        CU(cudaStreamSynchronize(home->stream));
        g->g_synced_seq = g->gseq;  // every fused merge issued so far has completed
        if (fused && *g->xstatus_host) {
            // a device never published: say which, and answer the group of queries over the event-ordered path
            g->g_timeouts++;
            redo_keyed = true;
            *g->xstatus_host = 0;
        }
        size_t stride = per_q;
        if (redo_keyed) {
            if ((rc = group_keyed_step(g, act, gq, (int)kk, filter, false, nullptr, 0, home->dscores, home->dids, nullptr))) return rc;
            stride = (size_t)kk;
            CU(cudaMemcpyAsync(home->hscores, home->dscores, (size_t)gq * stride * sizeof(float), cudaMemcpyDeviceToHost, home->stream));
            CU(cudaMemcpyAsync(home->hids, home->dids, (size_t)gq * stride * sizeof(long long), cudaMemcpyDeviceToHost, home->stream));
            CU(cudaStreamSynchronize(home->stream));
        }
        for (int64_t qi = 0; qi < gq; ++qi) {
            float* os = out_scores + (q0 + qi) * k;
            int64_t* oi = out_ids + (q0 + qi) * k;
            memcpy(os, home->hscores + qi * stride, (size_t)kk * sizeof(float));
            memcpy(oi, home->hids + qi * stride, (size_t)kk * sizeof(long long));
            for (int64_t i = kk; i < k; ++i) {
                os[i] = empty;
                oi[i] = -1;
            }
        }
    }
    return PSX_OK;
}
