/*
 * psx.h -- C ABI of the B200-native dense-recall engine ("psx" = photo-search exact scan).
 *
 * This is the drop-in boundary below the reference's Python class
 * `utils/vector_store.py::VectorStore`.  In the reference that class binds the third-party
 * faiss-cpu SWIG module; each entry point below names the FAISS call (and the reference
 * file:line that makes it) which it replaces.  Signatures are plain C: pointers, sizes and
 * integer codes only -- no torch / C++ types.
 *
 * Conventions
 *   - every function returns PSX_OK (0) or a negative psx_status; psx_last_error() gives a
 *     thread-local human readable message for the last failure on the calling thread;
 *   - "host" pointers are ordinary CPU memory owned by the caller, "dev" pointers are CUDA
 *     device memory on the index' device, `stream` is a cudaStream_t passed as void* and is
 *     used literally (NULL = the CUDA default stream, exactly as in the runtime API); the
 *     host-buffer entry points run on a private non-blocking stream of the index;
 *   - row ids are 0-based insertion order, exactly as FAISS labels (utils/vector_store.py:194-197);
 *     unfilled result slots are id -1 with score -inf (inner product) / +inf (L2), as FAISS
 *     leaves them (utils/vector_store.py:195 skips label -1);
 *   - results are ordered best first; equal scores are ordered by lower id first;
 *   - a handle may be used from several threads; calls on one handle are serialised internally.
 *   - there is NO CPU fallback: every compute entry point fails with PSX_ERR_CUDA when no
 *     sm_100 device is usable.
 */
#ifndef PSX_H_
#define PSX_H_

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define PSX_ABI_VERSION 2

typedef enum psx_status {
    PSX_OK = 0,
    PSX_ERR_INVALID = -1,   /* bad argument (dimension mismatch, k <= 0, null pointer ...) */
    PSX_ERR_CUDA = -2,      /* CUDA runtime / driver failure, or no usable device */
    PSX_ERR_OOM = -3,       /* device or host allocation failed */
    PSX_ERR_RANGE = -4,     /* row id out of range */
    PSX_ERR_STATE = -5      /* call not valid in the current state */
} psx_status;

/* faiss.METRIC_INNER_PRODUCT / faiss.METRIC_L2 as used at utils/vector_store.py:73-81,92-101 */
#define PSX_METRIC_IP 0
#define PSX_METRIC_L2 1

/* storage precision of the corpus in HBM; arithmetic is always fp32 accumulate */
#define PSX_STORE_F32 0
#define PSX_STORE_BF16 1
/* bf16 rows for the scan PLUS an fp32 master copy (1.5x the fp32 footprint): every search streams
 * the bf16 rows (half the bytes) for k' = 4k+64 candidates and re-scores them on the master with
 * the fp32 scan's reduction tree; a per-query proof obligation (k-th exact score >= k'-th bf16
 * score + rounding bound) certifies the result, otherwise the fp32 master is scanned.  Results are
 * bit-identical to a PSX_STORE_F32 index. */
#define PSX_STORE_BF16_MASTER 2

/* hard upper bound of k for ONE scan pass; larger k is served by paging inside psx_search */
#define PSX_K_PASS_MAX 2048

typedef struct psx_index psx_index; /* opaque */

/* ---- EXIF predicate (core/searcher.py:1884-1950 `_check_time_match_v2`) -----------------
 * One 64-bit attribute word per row:
 *   bits  0..38  dt      0 = no parseable datetime, else 1 + seconds since 0001-01-01T00:00:00
 *                         of `time_info.datetime_str or exif_data.datetime`
 *   bits 39..42  month   time_info.month  (0 = None, 1..12, 15 = not representable)
 *   bits 43..56  year    time_info.year   (0 = None, 1..16382, 16383 = not representable)
 *   bits 57..59  period  time_info.time_period code (0 = None, 1..7 = the 7 day parts of
 *                         core/indexer.py:583-598, in that order)
 *   bits 60..62  season  time_info.season code (0 = None, 1 春天 2 夏天 3 秋天 4 冬天, 7 = other)
 *   bit  63      exif    exif_data.datetime is truthy
 * A row passes iff every active clause holds; a row without `exif` fails any of the
 * season/period/year/month clauses, a row with dt == 0 fails an active range clause.
 */
#define PSX_F_SEASON 0x01u
#define PSX_F_PERIOD 0x02u
#define PSX_F_YEAR 0x04u
#define PSX_F_MONTH 0x08u
#define PSX_F_NEED_DT 0x10u /* start_date or end_date given (even if unparseable) */
#define PSX_F_START 0x20u   /* `start` is valid */
#define PSX_F_END 0x40u     /* `end` is valid   */

typedef struct psx_filter {
    uint32_t flags;
    uint32_t season;
    uint32_t period;
    uint32_t year;
    uint32_t month;
    uint32_t reserved;
    uint64_t start; /* dt encoding, inclusive */
    uint64_t end;   /* dt encoding, inclusive */
} psx_filter;

/* ---- lifecycle ---------------------------------------------------------------------------- */

/* Replaces faiss.IndexFlatIP(d) / faiss.IndexFlatL2(d) (utils/vector_store.py:72-81).
 * `device` is a CUDA ordinal.  Fails with PSX_ERR_CUDA if the device is absent or not sm_100. */
int psx_create(int d, int metric, int store_dtype, int device, psx_index** out);
/* ONE handle over several GPUs of one box, for the reference's process model: main.py:59-68 constructs a single
 * VectorStore that core/searcher.py:78 and core/indexer.py:57 share, so the corpus is row-sharded BEHIND the handle
 * instead of across processes.  `devices` lists n_devices CUDA ordinals (an ordinal may repeat: several shards on one
 * GPU, used by the tests); devices[0] is the "home" that merges.  Rows are split in contiguous id ranges (re-split with
 * device-to-device copies when appends unbalance them), every host-buffer entry point below works on the handle
 * (add / sync / reserve / set_attrs / search / reconstruct / read_rows / reset / tunables), results are bit-identical
 * to a single-device index.  The device-pointer entry points (psx_search_device, psx_search_batch_device,
 * psx_search_exchange_device, psx_storage_device) are per-device calls and fail with PSX_ERR_STATE on such a handle;
 * psx_add_device / psx_set_attrs_device accept a pointer on any device.  Needs peer access devices[i] -> devices[0].
 * n_devices == 1 is psx_create. */
int psx_create_sharded(int d, int metric, int store_dtype, int n_devices, const int* devices, psx_index** out);
/* number of shards behind the handle (1 for psx_create), rows and device of each (arrays of `capacity` >= that number) */
int psx_device_count(const psx_index* h);
int psx_shard_rows(psx_index* h, int64_t* rows, int* devices, int capacity);
/* multi-device handles: queries answered through the fused NVLink exchange / through event-ordered key lists, and
 * fused exchanges that timed out (a device that never published; the query was re-run over the key-list path) */
int psx_group_stats(psx_index* h, int64_t* fused, int64_t* keyed, int64_t* timeouts);
/* Replaces dropping the faiss index object. */
int psx_destroy(psx_index* h);
/* Replaces re-creating the index in VectorStore.clear() (utils/vector_store.py:273-280). */
int psx_reset(psx_index* h);
/* Replaces index.ntotal / index.d (utils/vector_store.py:183,188,255-258,271). */
int64_t psx_ntotal(const psx_index* h);
int psx_dim(const psx_index* h);
int psx_metric(const psx_index* h);
const char* psx_last_error(void);
int psx_abi_version(void);

/* ---- write side --------------------------------------------------------------------------- */

/* Replaces index.add(np.float32 (n,d)) (utils/vector_store.py:163-164).  `x` is n*d host
 * floats, row major, ALREADY normalised by the caller when the metric is cosine (the
 * reference normalises in Python, utils/vector_store.py:83-90, before calling FAISS).
 * Rows are staged on the host and uploaded lazily by the next search / psx_sync, so
 * one-vector-at-a-time appends (core/indexer.py:858) do not launch anything. */
int psx_add(psx_index* h, const float* x, int64_t n);
/* Bulk ingest of rows already resident on the device (fp32, row major, n*d).  When
 * `normalize` is non-zero every row is L2-normalised on the device first (zero rows are kept
 * unchanged, as utils/vector_store.py:88-89).  Runs on `stream`. */
int psx_add_device(psx_index* h, const float* x_dev, int64_t n, int normalize, void* stream);
/* Host-to-HBM throughput (GB/s of fp32 host bytes over wall time) of the last large psx_add / psx_sync upload: blocks of
 * 32 MB or more skip the staging vector and are pipelined through two pinned staging buffers (multi-threaded host copy
 * of chunk i+1 overlaps the PCIe transfer and the pack kernel of chunk i). */
double psx_upload_gbps(psx_index* h);
/* Pre-size the HBM arena for `n` rows in total (avoids regrowth copies). */
int psx_reserve(psx_index* h, int64_t n);
/* Upload everything staged by psx_add. */
int psx_sync(psx_index* h);
/* Attribute words for rows [row0, row0+n) (host pointer).  Rows never given a word hold 0
 * (no EXIF), i.e. fail every active clause -- the reference's behaviour for photos without
 * EXIF (core/searcher.py:1903-1927). */
int psx_set_attrs(psx_index* h, int64_t row0, const uint64_t* attrs, int64_t n);
int psx_set_attrs_device(psx_index* h, int64_t row0, const uint64_t* attrs_dev, int64_t n, void* stream);

/* ---- read side ---------------------------------------------------------------------------- */

/* Replaces index.search(np.float32 (nq,d), k) -> (D float32 (nq,k), I int64 (nq,k))
 * (utils/vector_store.py:190-191).  Host buffers; `q` already normalised for cosine.
 * `filter` may be NULL (no predicate).  Scores are inner products (best = largest) or
 * squared L2 distances (best = smallest).  Any k >= 1 is accepted. */
int psx_search(psx_index* h, const float* q, int64_t nq, int64_t k, const psx_filter* filter,
               float* out_scores, int64_t* out_ids);

/* Same scan with every buffer on the device and no host synchronisation: q_dev (nq*d fp32),
 * out_scores_dev (nq*k), out_ids_dev (nq*k int64), out_keys_dev (nq*kpad uint64, may be NULL)
 * with kpad = psx_kpad(k).  `id_base` is added to every row id (row-sharded corpora).
 * k <= PSX_K_PASS_MAX.  The 64-bit keys are the sortable form of (score, id):
 *   key = (orderable(score) << 32) | ~(uint32)(id_base + row),  0 = empty slot,
 * so that a larger key is a better hit and a plain integer sort reproduces the result order;
 * they are what row shards exchange (psx_merge_keys_device). */
int psx_search_device(psx_index* h, const float* q_dev, int64_t nq, int64_t k, const psx_filter* filter,
                      uint32_t id_base, float* out_scores_dev, int64_t* out_ids_dev,
                      uint64_t* out_keys_dev, void* stream);
int64_t psx_kpad(int64_t k);

/* Batched queries on the tensor cores (replaces nq consecutive index.search calls, e.g. the
 * expansion alternatives of core/searcher.py:1392-1412): S = Q X^T as tcgen05 TF32 GEMM tiles with
 * the selection fused into the epilogue (the score matrix never reaches HBM), then an exact fp32
 * re-score of the survivors.  flags_dev[i] == 0 certifies that query i's result is the exact
 * top-k (bit-identical to psx_search_device); a non-zero flag means "not proven", the caller
 * re-runs that query with psx_search_device.  psx_search does both steps itself for nq >= the
 * "batch_min" tunable.  `qnorm_max` is ignored (kept for ABI stability): the rounding bound of every query is
 * computed on the device from that query's own norm and the largest stored row norm.
 * `filter` (nullable) is the same EXIF predicate as in psx_search, shared by the whole batch: it is
 * applied to the survivors in the epilogue and to the sample the thresholds come from.
 * Inner-product indexes with fp32 rows (PSX_STORE_F32 / PSX_STORE_BF16_MASTER), >= 65536 rows, d >= 32,
 * k <= PSX_K_PASS_MAX (the reference's call sites ask for 500..1333, core/searcher.py:771-820). */
int psx_search_batch_device(psx_index* h, const float* q_dev, int64_t nq, int64_t k, const psx_filter* filter, float qnorm_max,
                            uint32_t id_base, float* out_scores_dev, int64_t* out_ids_dev, uint64_t* out_keys_dev,
                            int* flags_dev, void* stream);
/* 1 when psx_search_batch_device accepts this index for this k (metric, rows, dimension, k -- see above), else 0 */
int psx_batch_supported(psx_index* h, int64_t k);
/* queries served by the batched path so far, and how many of them had to be re-run on the scan */
int psx_batch_stats(psx_index* h, int64_t* queries, int64_t* fallbacks);
/* Final merge of `nlists` sorted key lists per query (layout [nq][nlists][kpad], e.g. the
 * all-gathered per-shard results) into the global top-k.  Pure device work on `stream`. */
int psx_merge_keys_device(int device, const uint64_t* keys_dev, int64_t nq, int64_t nlists, int64_t k,
                          int metric, float* out_scores_dev, int64_t* out_ids_dev, void* stream);

/* Row-sharded search with the exchange fused into the kernels (no collective library call on the
 * query path): the scan's last CTA stores this shard's k keys straight into every rank's receive
 * buffer over NVLink peer mappings and raises a per-source flag there; a one-CTA kernel on every
 * rank waits for `world` flags and selects the global top-k.  `peer_bases` is a HOST array of
 * `world` device addresses: the exchange buffer of every rank as mapped into THIS process
 * (psx_exchange_bytes() bytes each, zero-initialised once, e.g. torch symmetric memory or CUDA IPC);
 * peer_bases[rank] is this rank's own buffer.  `seq` >= 1 must increase by one per query and be
 * the same on all ranks; every rank must issue the same sequence of calls.  world <= 8.
 * `phases`: 1 = scan + publish, 2 = wait + merge, 3 = both (the normal call; the split exists so
 * that several ranks can be emulated on one GPU without kernels waiting on one another). */
int64_t psx_exchange_bytes(void);
int psx_search_exchange_device(psx_index* h, const float* q_dev, int64_t k, const psx_filter* filter, uint32_t id_base,
                               int rank, int world, const uint64_t* peer_bases, uint32_t seq, int phases,
                               float* out_scores_dev, int64_t* out_ids_dev, void* stream);
/* The wait of the fused exchange is bounded (tunable "xchg_timeout_ms", default 20 s): a rank that never publishes
 * does not hang or fault this GPU.  After the stream has been synchronised, *status is 0 when every exchange since
 * the last call completed, else 1 + the first silent rank -- the outputs of that query are then undefined and the
 * caller re-runs it over the collective path (psx_search_device keys + all-gather + psx_merge_keys_device).  Reading
 * clears the word. */
int psx_exchange_status(psx_index* h, int* status);

/* On-device form of the numeric core of Searcher._hybrid_search (core/searcher.py:893-986) and
 * Searcher._distance_to_score (core/searcher.py:605-625) over the merged vector candidates
 * (vec_dist/vec_ids [nq][kv], id -1 = empty) and the keyword hits (kw_ids/kw_scores [nq][kw], scores
 * in [0,1] as utils/keyword_store.py:270-279 normalises them).  Python-float (IEEE double)
 * arithmetic and round(x, 6) are reproduced bit for bit.  *_boost are the per-hit metadata
 * boosts of core/searcher.py:435-449 (NULL = 1.0).  `keyword_filtered` = the Elasticsearch filter
 * branch was taken (es_filtered_paths is not None).  Outputs, per query, kv+kw slots sorted by
 * fused score descending then id ascending: ids (-1 beyond out_count), fused / vector / keyword
 * scores (rounded to 6 places).  kv + kw <= 2048. */
int psx_hybrid_fuse_device(int device, int64_t nq, int64_t kv, const float* vec_dist_dev, const int64_t* vec_ids_dev,
                           const double* vec_boost_dev, int64_t kw, const int64_t* kw_ids_dev, const double* kw_scores_dev,
                           const double* kw_boost_dev, double vector_weight, double keyword_weight, int metric,
                           int allow_keyword_only, int keyword_filtered, int64_t* out_ids_dev, double* out_fused_dev,
                           double* out_vscore_dev, double* out_kscore_dev, int* out_count_dev, void* stream);

/* On-device form of the numeric part of Searcher._finalize_results (core/searcher.py:1497-1526): per query, over its
 * fused scores in candidate order (scores_dev [nq][m] descending, counts_dev [nq] valid entries -- the out_fused /
 * out_count of psx_hybrid_fuse_device), the dynamic threshold of Searcher._calculate_dynamic_threshold (:627-674;
 * np.percentile's linear interpolation, np.median and round(x, 6) reproduced bit for bit), the strict / broad
 * thresholds from the round's floors (Searcher._get_round_score_floors, :822-826, computed by the caller) and the
 * confidence bucket of every hit by score (3 reliable, 2 generalised, 1 rest; :828-840 without the term matching, which
 * stays on the host).  out_counts_dev [nq][2] = hits in bucket 3 and in bucket 2. */
int psx_finalize_device(int device, int64_t nq, int64_t m, const double* scores_dev, const int* counts_dev, int top_k,
                        double strict_floor, double broad_floor, double threshold_floor, double* out_strict_dev,
                        double* out_broad_dev, int* out_bucket_dev, int* out_counts_dev, void* stream);

/* Replaces index.reconstruct(i) (utils/vector_store.py:207): the stored row as fp32. */
int psx_reconstruct(psx_index* h, int64_t id, float* out);
/* Replaces the flat payload of faiss.write_index (utils/vector_store.py:234): rows
 * [row0,row0+n) as fp32 into a host buffer, in id order. */
int psx_read_rows(psx_index* h, int64_t row0, int64_t n, float* out);
/* Device address/stride of the stored rows (after psx_sync), for zero-copy consumers. */
int psx_storage_device(psx_index* h, const void** rows_dev, int64_t* ld_elems, int* store_dtype);

/* ---- tuning / introspection ---------------------------------------------------------------- */

/* key: "warps" (consumer warps per CTA), "stages" (ring slots per warp), "ctas_per_sm"
 * (values <= 0 restore the default), "batch_min" (smallest nq routed to the tensor-core path
 * by psx_search; 0 disables it, < 0 restores the default of 4), "batch_pair" (CTA-pair GEMM kernel
 * for 129..256 queries, default 1), "filter_mode" (0 = default (3 up to 4M rows, 2 beyond), 1 = EXIF predicate evaluated inside
 * the scan group by group, 2 = predicate compacted into a row list by a kernel ahead of the scan, 3 = compacted into
 * the row list by the first phase of the scan launch itself -- one launch per filtered query), "deal" (1 = rows
 * dealt to the warps as units with a dynamically scheduled tail, the default; 0 = static groups),
 * "dyn_tail" (0 = deal everything statically), "static_batch" (units per dealt batch, default 8), "pdl"
 * (programmatic dependent launch: the next scan starts streaming while the previous one sorts and merges.  0 = plain
 * stream order; 1 = between the scans of ONE call, the default; 2 = also across calls: only valid when the query of a
 * call is never written by the kernel enqueued right before that call on its stream -- queries already resident on
 * the device or delivered by a memcpy).
 * Tunables change launch geometry only, never results. */
int psx_set_tunable(psx_index* h, const char* key, int value);
/* Diagnostics: while `trace_dev` is non-NULL every scan launch writes, per CTA, 8 uint64 into
 * trace_dev[cta*8 ..]: %globaltimer (ns) at kernel entry, prologue done, stream drained, list
 * published, lists merged (last CTA only), results emitted (last CTA only), then the number of
 * candidate-buffer compactions; a filtered launch that compacts its own row list adds a second block at
 * trace_dev[(grid + cta)*8 ..]: first ticket known, words evaluated, list space reserved, entries written, every ticket
 * finished.  The buffer (>= 16 * SM count * ctas_per_sm words, device memory) stays owned by the caller; NULL switches
 * tracing off. */
int psx_set_trace_device(psx_index* h, uint64_t* trace_dev);
/* Number of kernels launched by this library in the calling process so far. */
int64_t psx_launch_count(void);

#ifdef __cplusplus
}
#endif
#endif /* PSX_H_ */
