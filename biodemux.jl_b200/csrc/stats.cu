// stats.cu -- DemuxStats counters (classification.jl:736-767): layout, retrieval, the report bridge
// (include/bdx.h).
#include <algorithm>
#include <cmath>
#include <cstring>

#include "api_internal.h"

using namespace bdx;

// ---------------------------------------------------------------------------
// stats
// ---------------------------------------------------------------------------
extern "C" int bdx_stats_layout_get(const bdx_config *cfg, bdx_stats_layout *out)
{
    if (!cfg || !out) return bdx_fail(BDX_ERR_INVALID, "null argument");
    *out = cfg->lay;
    return BDX_OK;
}

extern "C" int bdx_stats_fetch(bdx_stream *s, int64_t *out, int64_t out_len)
{
    if (!s || !out) return bdx_fail(BDX_ERR_INVALID, "null argument");
    if (!s->d_stats) return bdx_fail(BDX_ERR_STATE, "config was created without want_stats");
    if (out_len < s->cfg->lay.total_len) return bdx_fail(BDX_ERR_INVALID, "stats buffer too small");
    CU(cudaSetDevice(s->device));
    CU(cudaStreamSynchronize(s->st_comp));
    CU(cudaMemcpy(out, s->d_stats, (size_t)s->cfg->lay.total_len * 8, cudaMemcpyDeviceToHost));
    return BDX_OK;
}

// DemuxStats dictionaries from a (summed) counter buffer: what match_barcode_pass stores per matched pass
// (classification.jl:827-865) -- keys are alignment start, alignment length and round(score, digits=2).
extern "C" int64_t bdx_stats_entries(const bdx_config *cfg, const int64_t *counters, bdx_stats_entry *out, int64_t cap)
{
    if (!cfg || !counters) return bdx_fail(BDX_ERR_INVALID, "null argument");
    const bdx_stats_layout &L = cfg->lay;
    int64_t n = 0;
    auto emit = [&](int pass, int kind, int bc, int64_t key, double score, int64_t count) {
        if (out && n < cap) out[n] = bdx_stats_entry{pass, kind, bc, 0, key, score, count};
        n++;
    };
    const int passes = cfg->base.is_dual ? 2 : 1;
    for (int p = 0; p < passes; p++) {
        const HostSet &hs = cfg->set[p];
        const int nb = p == 0 ? L.b1 : L.b2;
        std::map<double, int64_t> global_score;   // the global score Dict is keyed by the rounded score, so it
                                                  // has to be re-binned from the per-barcode distances
        for (int b = 0; b <= nb; b++) {
            const int64_t *pos = counters + L.pos_off[p] + (int64_t)b * L.pos_bins;
            const int64_t *len = counters + L.len_off[p] + (int64_t)b * L.len_bins;
            const int64_t *dst = counters + L.dist_off[p] + (int64_t)b * L.dist_bins;
            for (int k = 0; k < L.pos_bins; k++)
                if (pos[k]) emit(p + 1, BDX_STATS_POS, b, k - L.pos_bias, 0.0, pos[k]);
            for (int k = 0; k < L.len_bins; k++)
                if (len[k]) emit(p + 1, BDX_STATS_LEN, b, k, 0.0, len[k]);
            if (b == 0) continue;
            // normalisation as in the kernels: bc_lengths_no_N under NScoring, else the barcode length
            const int norm = cfg->base.algo == BDX_SEMIGLOBAL ? hs.norm[b - 1] : hs.off[b] - hs.off[b - 1];
            for (int k = 0; k < L.dist_bins; k++) {
                if (!dst[k]) continue;
                const double score = (double)(k - L.dist_bias) / (double)norm;
                // Base.round(x, digits=2): round-half-even of x * 100, divided by 100; x itself if that is not finite
                volatile double scaled = score * 100.0;
                double r = std::nearbyint(scaled) / 100.0;
                if (!std::isfinite(r)) r = score;
                emit(p + 1, BDX_STATS_SCORE, b, 0, r, dst[k]);
                global_score[r] += dst[k];
            }
        }
        for (auto &kv : global_score) emit(p + 1, BDX_STATS_SCORE, 0, 0, kv.first, kv.second);
    }
    return n;
}

// ---- overflow list: exact records of matched passes whose start / length lie outside the histograms ----
// A pass of a read appends at most one record, and only reads longer than the histograms' 1024 positions can.
// Before a batch is enqueued the host reserves room for the worst case (2 records per read, none at all when the
// batch's longest read fits the histograms); when the room is not there it first drains the device list to the
// host (one synchronisation, amortised over many batches) and, for a single huge batch, grows the list.
int bdx_stats_drain_overflow(bdx_stream *s)
{
    if (!s->d_stats) return BDX_OK;
    CU(cudaSetDevice(s->device));
    CU(cudaStreamSynchronize(s->st_comp));
    unsigned int h[2];
    CU(cudaMemcpy(h, s->d_n_ovf, sizeof(h), cudaMemcpyDeviceToHost));
    const int64_t kept = std::min<int64_t>(h[0], s->ovf_cap);
    if (kept > 0) {
        const size_t old = s->h_ovf.size();
        s->h_ovf.resize(old + (size_t)kept);
        CU(cudaMemcpy(s->h_ovf.data() + old, s->d_ovf, (size_t)kept * sizeof(bdx_stats_overflow), cudaMemcpyDeviceToHost));
        const unsigned int zero = 0;
        CU(cudaMemcpy(s->d_n_ovf, &zero, sizeof(zero), cudaMemcpyHostToDevice));     // `lost` (h[1]) is kept
    }
    s->ovf_bound = 0;
    return BDX_OK;
}

int bdx_stats_reserve_overflow(bdx_stream *s, int64_t n_reads, int64_t max_len)
{
    if (!s->d_stats || n_reads <= 0) return BDX_OK;
    if (max_len >= 0 && max_len <= 1024) return BDX_OK;       // starts <= n and lengths <= n + m fit the histograms
    const int64_t need = 2 * n_reads;
    if (s->ovf_bound + need > s->ovf_cap) {
        const int rc = bdx_stats_drain_overflow(s);
        if (rc) return rc;
        if (need > s->ovf_cap) {
            cudaFree(s->d_ovf);
            s->d_ovf = nullptr;
            s->ovf_cap = 0;
            bdx_stream_drop_graphs(s);
            CU(cudaMalloc(&s->d_ovf, (size_t)need * sizeof(bdx_stats_overflow)));
            s->ovf_cap = need;
        }
    }
    s->ovf_bound += need;
    return BDX_OK;
}

extern "C" int bdx_stats_overflow_fetch(bdx_stream *s, bdx_stats_overflow *out, int64_t cap, int64_t *n, int64_t *lost)
{
    if (!s || !n || cap < 0 || (cap > 0 && !out)) return bdx_fail(BDX_ERR_INVALID, "bad argument");
    *n = 0;
    if (lost) *lost = 0;
    if (!s->d_stats) return bdx_fail(BDX_ERR_STATE, "config was created without want_stats");
    const int rc = bdx_stats_drain_overflow(s);
    if (rc) return rc;
    unsigned int h[2];
    CU(cudaMemcpy(h, s->d_n_ovf, sizeof(h), cudaMemcpyDeviceToHost));
    *n = (int64_t)s->h_ovf.size();
    if (lost) *lost = h[1];
    const int64_t take = std::min<int64_t>((int64_t)s->h_ovf.size(), cap);
    if (take > 0) memcpy(out, s->h_ovf.data(), (size_t)take * sizeof(bdx_stats_overflow));
    return BDX_OK;
}

extern "C" void *bdx_stats_device_ptr(bdx_stream *s) { return s ? (void *)s->d_stats : nullptr; }

extern "C" int bdx_stats_reset(bdx_stream *s)
{
    if (!s) return bdx_fail(BDX_ERR_INVALID, "null stream");
    if (!s->d_stats) return BDX_OK;
    CU(cudaSetDevice(s->device));
    CU(cudaMemsetAsync(s->d_stats, 0, (size_t)s->cfg->lay.total_len * 8, s->st_comp));
    CU(cudaMemsetAsync(s->d_n_ovf, 0, 2 * sizeof(unsigned int), s->st_comp));
    s->h_ovf.clear();
    s->ovf_bound = 0;
    return BDX_OK;
}
