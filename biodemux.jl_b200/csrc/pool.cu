// pool.cu -- bdx_pool: the dispatcher a single host process uses to spread batches over several GPUs
// (include/bdx.h).  Built on the public stream calls only.
#include <new>

#include "api_internal.h"

// ---------------------------------------------------------------------------
// dispatcher over several GPUs: round-robin over streams, results in submission order
// ---------------------------------------------------------------------------
struct bdx_pool {
    std::vector<bdx_stream *> streams;
    std::vector<int> order;   // stream index of every batch in flight, oldest first
    size_t next = 0;          // stream whose turn it is
};

extern "C" void bdx_pool_destroy(bdx_pool *p)
{
    if (!p) return;
    for (bdx_stream *s : p->streams) bdx_stream_destroy(s);
    delete p;
}

extern "C" int bdx_pool_create(const bdx_config *cfg, const int *devices, int n_devices, int streams_per_device,
                               int32_t max_reads, int64_t max_bytes, bdx_pool **out)
{
    if (!cfg || !devices || !out || n_devices <= 0 || streams_per_device <= 0) return bdx_fail(BDX_ERR_INVALID, "bad argument");
    *out = nullptr;
    bdx_pool *p = new (std::nothrow) bdx_pool();
    if (!p) return bdx_fail(BDX_ERR_NOMEM, "out of memory");
    // stream k of every device before stream k + 1 of any: consecutive batches land on different GPUs
    for (int k = 0; k < streams_per_device; k++)
        for (int d = 0; d < n_devices; d++) {
            bdx_stream *s = nullptr;
            const int rc = bdx_stream_create(cfg, devices[d], max_reads, max_bytes, &s);
            if (rc) {
                std::string keep = bdx_error_text();
                bdx_pool_destroy(p);
                bdx_error_text() = keep;
                return rc;
            }
            p->streams.push_back(s);
        }
    *out = p;
    return BDX_OK;
}

// The next batch goes to the stream with the fewest batches in flight (ties: the one after the last used, i.e.
// round-robin over equally loaded streams): GPUs that drain their queues faster -- a shorter path to the host
// memory the batches come from, a less loaded device -- take more of the work, and the pool's rate approaches the
// sum of the GPUs' rates instead of N times the slowest.  Results still come back in submission order.
template <typename Submit>
static int pool_submit(bdx_pool *p, Submit submit)
{
    if (!p) return bdx_fail(BDX_ERR_INVALID, "null pool");
    const size_t n = p->streams.size();
    size_t best = n;
    for (size_t k = 0; k < n; k++) {
        const size_t i = (p->next + k) % n;
        if (p->streams[i]->in_flight >= BDX_MAX_IN_FLIGHT) continue;
        if (best == n || p->streams[i]->in_flight < p->streams[best]->in_flight) best = i;
    }
    if (best == n) return bdx_fail(BDX_ERR_STATE, "every stream of the pool is full; call bdx_pool_fetch");
    const int rc = submit(p->streams[best]);
    if (rc) return rc;
    p->order.push_back((int)best);
    p->next = (best + 1) % n;
    return BDX_OK;
}

extern "C" int bdx_pool_submit(bdx_pool *p, const uint8_t *seq, const int32_t *offsets, int32_t n, uint64_t tag)
{
    return pool_submit(p, [&](bdx_stream *s) { return bdx_submit(s, seq, offsets, n, tag); });
}

extern "C" int bdx_pool_submit_pinned(bdx_pool *p, const uint8_t *seq, const int32_t *offsets, int32_t n, uint64_t tag)
{
    return pool_submit(p, [&](bdx_stream *s) { return bdx_submit_pinned(s, seq, offsets, n, tag); });
}

extern "C" int bdx_pool_fetch(bdx_pool *p, uint64_t *tag, int32_t *n_reads, bdx_result *results, bdx_pass_detail *details)
{
    if (!p) return bdx_fail(BDX_ERR_INVALID, "null pool");
    if (p->order.empty()) return bdx_fail(BDX_ERR_STATE, "nothing in flight");
    const int rc = bdx_fetch(p->streams[(size_t)p->order.front()], tag, n_reads, results, details);
    if (rc == BDX_OK) p->order.erase(p->order.begin());
    return rc;
}

extern "C" int bdx_pool_fetch_view(bdx_pool *p, uint64_t *tag, int32_t *n_reads, const bdx_result **results,
                                   const bdx_pass_detail **details)
{
    if (!p) return bdx_fail(BDX_ERR_INVALID, "null pool");
    if (p->order.empty()) return bdx_fail(BDX_ERR_STATE, "nothing in flight");
    const int rc = bdx_fetch_view(p->streams[(size_t)p->order.front()], tag, n_reads, results, details);
    if (rc == BDX_OK) p->order.erase(p->order.begin());
    return rc;
}

extern "C" int bdx_pool_in_flight(const bdx_pool *p) { return p ? (int)p->order.size() : 0; }

extern "C" int bdx_pool_stats_fetch(bdx_pool *p, int64_t *out, int64_t out_len)
{
    if (!p || !out) return bdx_fail(BDX_ERR_INVALID, "null argument");
    const int64_t L = p->streams[0]->cfg->lay.total_len;
    if (out_len < L) return bdx_fail(BDX_ERR_INVALID, "stats buffer too small");
    std::vector<int64_t> tmp((size_t)L);
    std::fill(out, out + L, 0);
    for (bdx_stream *s : p->streams) {
        const int rc = bdx_stats_fetch(s, tmp.data(), L);
        if (rc) return rc;
        for (int64_t k = 0; k < L; k++) out[k] += tmp[(size_t)k];
    }
    return BDX_OK;
}
