// bdx_api.cu -- C ABI of libbdx (include/bdx.h): configuration, per-worker streams with
// pinned double-buffered staging, batch submission / retrieval, DemuxStats counters.
// Replaces the body of the reference's worker_task (src/core.jl:226-279).
#include <algorithm>
#include <cmath>
#include <cstdlib>
#include <cstring>
#include <map>
#include <mutex>
#include <new>
#include <string>
#include <vector>

#include "bdx_internal.h"
#include "demux.h"

using namespace bdx;

namespace bdx {
size_t filter_smem_bytes_for(const DevSet &S);
}

// ---------------------------------------------------------------------------
// errors
// ---------------------------------------------------------------------------
static thread_local std::string g_err;

static int fail(int code, const std::string &msg)
{
    g_err = msg;
    return code;
}
static int cuda_fail(cudaError_t e, const char *what)
{
    g_err = std::string(what) + ": " + cudaGetErrorString(e);
    return BDX_ERR_CUDA;
}
#define CU(call)                                              \
    do {                                                      \
        cudaError_t e__ = (call);                             \
        if (e__ != cudaSuccess) return cuda_fail(e__, #call); \
    } while (0)

extern "C" const char *bdx_last_error(void) { return g_err.c_str(); }
extern "C" int bdx_abi_version(void) { return BDX_ABI_VERSION; }
extern "C" int bdx_device_count(void)
{
    int n = 0;
    if (cudaGetDeviceCount(&n) != cudaSuccess) {
        cudaGetLastError();
        return 0;
    }
    return n;
}

// ---------------------------------------------------------------------------
// configuration
// ---------------------------------------------------------------------------
struct HostSet {
    int n_bc = 0, n_bc_pad = 0, max_m = 0, trim_side = 0, words = 0, n_classes = 1, use_filter = 0;
    DevRange rs{}, bs{}, be{};
    std::vector<uint8_t> bytes;
    std::vector<int> off, norm, filt_allowed, allowed0;
    std::vector<uint32_t> peq;
    uint8_t class_of[256];
    // perfect-occurrence prefilter
    int pf_enabled = 0, pf_seed = 0, pf_log2 = 0, pf_bm_log2 = 0;
    std::vector<uint32_t> pf_bitmap;
    uint32_t pf_pow = 0;
    std::vector<uint32_t> pf_keys, pf_vals;
    std::vector<uint8_t> bc_cls;
    // :hamming pigeonhole seeds
    int hs_enabled = 0, hs_q = 0, hs_log2 = 0, hs_max_off = 0;
    uint32_t hs_pow = 0;
    std::vector<uint32_t> hs_bstart, hs_entries;
    // :semiglobal depth-limited seeds
    // :hamming packed scan (hamming.cu)
    int hp_enabled = 0, hp_m = 0, hp_allowed = 0, hp_n_seg = 0;
    int hp_off[8] = {}, hp_q[8] = {}, hp_base[8] = {};
    std::vector<uint16_t> hp_bstart, hp_entries;
    std::vector<uint2> hp_bcw;
    struct HostSeedLevel {
        int k = 0, q = 0, log2 = 0, bm_log2 = 0;
        uint32_t pow = 0;
        std::vector<uint32_t> bstart, entries, ekeys, bitmap;
    };
    int sd_levels = 0, sd_m = 0;
    HostSeedLevel sd[2];
    int sdd_n = 0, sdd_k = 0;      // deepest level (seed_deep.cu): one table per seed length
    HostSeedLevel sdd[2];
    // variable lengths / constrained geometries (seed_var.cu)
    struct HostSeedVar {
        int q = 0, complete = 0, group_reads = 0, hit_rows = 0;
        double sigma_min = 0.0;
        std::vector<uint16_t> bstart;
        std::vector<uint32_t> entries;
        std::vector<uint8_t> kdepth;
    };
    int sv_levels = 0;
    HostSeedVar sv[2];
};

struct DeviceTables {
    DevParams P;
    std::vector<void *> allocs;
    int sm_count = 0;
};

struct bdx_config {
    DevParams base{};  // device pointers unset
    HostSet set[2];
    bdx_stats_layout lay{};
    std::mutex mu;
    std::map<int, DeviceTables *> per_device;
};

static int narrow_range(const bdx_range &in, DevRange &out, const char *name)
{
    const int64_t lim = 1ll << 30;
    if (in.start_offset > lim || in.start_offset < -lim || in.end_offset > lim || in.end_offset < -lim)
        return fail(BDX_ERR_INVALID, std::string(name) + ": range offset out of bounds");
    out.start_off = (int)in.start_offset;
    out.start_from_end = in.start_from_end ? 1 : 0;
    out.end_off = (int)in.end_offset;
    out.end_from_end = in.end_from_end ? 1 : 0;
    return BDX_OK;
}

static int host_allowed(double max_error, int norm)
{
    // floor(Int, max_error * normalization_length), classification.jl:254 (same clamp as the device)
    volatile double prod = max_error * (double)norm;
    double x = std::floor(prod);
    if (!(x < 268435456.0)) return 268435456;
    if (x < -268435456.0) return -268435456;
    return (int)x;
}

static int build_set(const bdx_params &p, const bdx_barcode_set &in, HostSet &hs, bool disable_filter,
                     const char *name)
{
    if (in.n_barcodes <= 0 || !in.bytes || !in.offsets)
        return fail(BDX_ERR_INVALID, std::string(name) + ": empty barcode set");
    if (in.n_barcodes > 65535) return fail(BDX_ERR_INVALID, std::string(name) + ": more than 65535 barcodes");
    if (in.trim_side != 0 && in.trim_side != 3 && in.trim_side != 5)
        return fail(BDX_ERR_INVALID, "trim_side must be 3 or 5");  // core.jl:308-313
    if (p.has_nindel && !in.lengths_no_n)
        return fail(BDX_ERR_INVALID, std::string(name) + ": lengths_no_n required with nindel");
    hs.n_bc = in.n_barcodes;
    hs.trim_side = in.trim_side;
    int rc;
    if ((rc = narrow_range(in.ref_search_range, hs.rs, name))) return rc;
    if ((rc = narrow_range(in.barcode_start_range, hs.bs, name))) return rc;
    if ((rc = narrow_range(in.barcode_end_range, hs.be, name))) return rc;
    if (in.offsets[0] != 0) return fail(BDX_ERR_INVALID, std::string(name) + ": offsets[0] must be 0");
    hs.off.assign(in.offsets, in.offsets + in.n_barcodes + 1);
    hs.max_m = 0;
    for (int b = 0; b < hs.n_bc; b++) {
        const int m = hs.off[b + 1] - hs.off[b];
        if (m <= 0) return fail(BDX_ERR_INVALID, std::string(name) + ": empty barcode (not supported)");
        if (m > kMaxBarcodeLen) return fail(BDX_ERR_INVALID, std::string(name) + ": barcode longer than 256");
        hs.max_m = std::max(hs.max_m, m);
    }
    hs.bytes.assign(in.bytes, in.bytes + hs.off[hs.n_bc]);
    hs.norm.resize(hs.n_bc);
    for (int b = 0; b < hs.n_bc; b++) {
        const int m = hs.off[b + 1] - hs.off[b];
        // semiglobal: m, or bc_lengths_no_N under NScoring (classification.jl:460, :476, :647);
        // hamming: m (:567, :607)
        hs.norm[b] = (p.algorithm == BDX_SEMIGLOBAL && p.has_nindel) ? in.lengths_no_n[b] : m;
        if (hs.norm[b] < 0) return fail(BDX_ERR_INVALID, std::string(name) + ": negative lengths_no_n");
    }

    // ---- bit-parallel filter tables (semiglobal only) ----
    const int groups = (hs.n_bc + 31) / 32;
    int gpad = groups;
    if (groups > 4) {
        const int m3 = (groups + 2) / 3 * 3, m4 = (groups + 3) / 4 * 4;
        gpad = m4 <= m3 ? m4 : m3;
    }
    hs.n_bc_pad = gpad * 32;
    memset(hs.class_of, 0, sizeof(hs.class_of));
    hs.n_classes = 1;
    for (uint8_t c : hs.bytes)
        if (!hs.class_of[c]) hs.class_of[c] = (uint8_t)hs.n_classes++;
    hs.words = hs.max_m <= 32 ? 1 : (hs.max_m <= 32 * kMaxFilterWords ? 2 : 0);
    const bool benign = p.match >= 0 && p.mismatch >= 1 && p.indel >= 1 && (!p.has_nindel || p.nindel >= p.indel);
    // The unit-cost filter is also a superset filter for :hamming (Hamming distance >= edit
    // distance; a barcode N is a wildcard there, classification.jl:597) and :exact (distance 0).
    // Tiny sets are cheaper to scan with the literal kernel than to spread over 32 lanes: they get the tables
    // (for the thread-per-read prefilter / seed kernels) but not the filter kernel.
    const bool sg = p.algorithm == BDX_SEMIGLOBAL;
    if ((sg && !benign) || disable_filter || hs.n_classes > 64) hs.words = 0;
    hs.use_filter = hs.words > 0 && hs.n_bc >= 8;

    hs.allowed0.assign(hs.n_bc_pad, -1);
    hs.filt_allowed.assign(hs.n_bc_pad, -1);
    int64_t min_cost = std::min(p.mismatch, p.indel);
    if (p.has_nindel) min_cost = std::min(min_cost, p.nindel);
    for (int b = 0; b < hs.n_bc; b++) {
        hs.allowed0[b] = host_allowed(p.max_error_rate, hs.norm[b]);
        if (p.algorithm == BDX_EXACT)
            hs.filt_allowed[b] = p.max_error_rate >= 0.0 ? 0 : -1;          // score 0.0 <= thr (:658, :696)
        else if (p.algorithm == BDX_HAMMING)
            hs.filt_allowed[b] = hs.allowed0[b] < 0 ? -1 : hs.allowed0[b];  // floor(thr * m) (:567)
        else if (benign)
            hs.filt_allowed[b] = hs.allowed0[b] < 0 ? -1 : (int)(hs.allowed0[b] / min_cost);
    }
    if (hs.words) {
        const int W = hs.words;
        const size_t plane = (size_t)hs.n_classes * hs.n_bc_pad;
        hs.peq.assign((size_t)W * plane, 0u);
        for (int b = 0; b < hs.n_bc_pad; b++) {
            const int m = b < hs.n_bc ? hs.off[b + 1] - hs.off[b] : 0;
            const int first_row_bit = W * 32 - m;  // bit of barcode row 1; lower bits are phantom rows
            for (int c = 0; c < hs.n_classes; c++) {
                uint64_t v = first_row_bit >= 64 ? ~0ull : ((1ull << first_row_bit) - 1);  // phantom rows match
                if (W == 1) v &= 0xFFFFFFFFull;
                for (int i = 0; i < m; i++) {
                    const uint8_t q = hs.bytes[hs.off[b] + i];
                    // wildcard rows: NScoring (:196-203) and hamming_align (:597); literal in :exact
                    const bool is_n = q == (uint8_t)'N' && ((sg && p.has_nindel) || p.algorithm == BDX_HAMMING);
                    if (is_n || (c != 0 && hs.class_of[q] == c)) v |= 1ull << (first_row_bit + i);
                }
                hs.peq[0 * plane + (size_t)c * hs.n_bc_pad + b] = (uint32_t)v;
                if (W == 2) hs.peq[1 * plane + (size_t)c * hs.n_bc_pad + b] = (uint32_t)(v >> 32);
            }
        }
    }
    // ---- perfect-occurrence prefilter table (semiglobal, no wildcard rows) ----
    hs.bc_cls.resize(hs.bytes.size());
    for (size_t k = 0; k < hs.bytes.size(); k++) hs.bc_cls[k] = hs.class_of[hs.bytes[k]];
    int min_m = hs.max_m;
    for (int b = 0; b < hs.n_bc; b++) min_m = std::min(min_m, hs.off[b + 1] - hs.off[b]);
    const bool ex = p.algorithm == BDX_EXACT;   // :exact keeps duplicates (each index is a candidate)
    // :hamming treats every barcode N as a wildcard (classification.jl:597): no table then
    const bool hm = p.algorithm == BDX_HAMMING &&
                    std::find(hs.bytes.begin(), hs.bytes.end(), (uint8_t)'N') == hs.bytes.end();
    if (hs.words && ((sg && !p.has_nindel) || ex || hm) && min_m >= kPfMinSeed && !getenv("BDX_DISABLE_PREFILTER")) {
        const int seed = std::min(min_m, kPfMaxSeed);
        hs.pf_seed = seed;
        uint32_t pw = 1;
        for (int i = 1; i < seed; i++) pw *= kPfBase;
        hs.pf_pow = pw;
        int lg = 4;
        while ((1 << lg) < 2 * hs.n_bc) lg++;
        hs.pf_log2 = lg;
        const uint32_t size = 1u << lg;
        hs.pf_keys.assign(size, 0u);
        hs.pf_vals.assign(size, kPfEmpty);
        int bl = 13;                                    // >= 512 bits per barcode, 8 KB .. 32 KB
        while (bl < 18 && (1 << bl) < 512 * hs.n_bc) bl++;
        hs.pf_bm_log2 = bl;
        hs.pf_bitmap.assign((size_t)1 << (bl - 5), 0u);
        for (int b = 0; b < hs.n_bc; b++) {            // ascending: the lowest index of identical sequences stays
            const int m = hs.off[b + 1] - hs.off[b];
            uint32_t h = 0;
            for (int i = 0; i < seed; i++) h = h * kPfBase + (uint32_t)hs.bytes[hs.off[b] + i];
            const uint32_t bit = pf_bit(h, bl);
            hs.pf_bitmap[bit >> 5] |= 1u << (bit & 31);
            uint32_t slot = pf_slot(h, lg);
            bool dup = false;
            while (hs.pf_vals[slot] != kPfEmpty) {
                const uint32_t v = hs.pf_vals[slot];
                const int ob = (int)(v & 0xFFFFu);
                if (!ex && (int)(v >> 16) == m &&
                    memcmp(&hs.bytes[hs.off[ob]], &hs.bytes[hs.off[b]], (size_t)m) == 0) {
                    dup = true;
                    break;
                }
                slot = (slot + 1) & (size - 1);
            }
            if (!dup) {
                hs.pf_keys[slot] = h;
                hs.pf_vals[slot] = ((uint32_t)m << 16) | (uint32_t)b;
            }
        }
        hs.pf_enabled = 1;
    }
    // ---- :semiglobal depth-limited seeds (seed.cu): uniform barcode length, no wildcard rows ----
    if (hs.pf_enabled && sg && hs.words >= 1 && min_m == hs.max_m && hs.allowed0[0] >= 1 && hs.n_bc < (1 << 14) &&
        !getenv("BDX_DISABLE_SEED")) {
        const int m = hs.max_m, allowed = hs.allowed0[0];
        const double alpha = std::max(2, hs.n_classes - 1);
        // chance hits per read column of level k: entries / alphabet^q with q = min(12, m / (k + 1))
        auto q_of = [&](int k) { return std::min(12, m / (k + 1)); };
        auto rate_of = [&](int k) { return (double)hs.n_bc * (k + 1) / std::pow(alpha, q_of(k)); };
        // deepest level whose seeds are long enough to be selective (q >= 6, <= 0.12 chance hits per column)
        int K = 0;
        for (int k = 1; k <= std::min(allowed, 7); k++)   // hit records keep the diagonal span in 3 bits
            if (q_of(k) >= 6 && rate_of(k) <= 0.12) K = k;
        // a shallower level with far fewer chance hits in front of it pays when the deep one has many
        int K0 = 0;
        if (K >= 2 && rate_of(K) > 0.02 && !getenv("BDX_SEED_ONE_LEVEL"))
            for (int k = 1; k < K; k++)
                if (rate_of(k) <= 0.01) K0 = k;
        hs.sd_m = m;
        for (int K_l : {K0, K}) {
            if (K_l < 1) continue;
            HostSet::HostSeedLevel &L = hs.sd[hs.sd_levels++];
            const int seg = m / (K_l + 1), q = std::min(12, seg);
            L.k = K_l;
            L.q = q;
            uint32_t pw = 1;
            for (int i = 1; i < q; i++) pw *= kPfBase;
            L.pow = pw;
            const size_t n_entries = (size_t)hs.n_bc * (K_l + 1);
            int lg = 8;
            while (lg < 14 && (size_t)(1 << lg) < n_entries) lg++;
            L.log2 = lg;
            int bl = 13;                                // ~64 bits per entry: 4 KB for 96 barcodes
            while (bl < 18 && ((size_t)1 << bl) < 64 * n_entries) bl++;
            L.bm_log2 = bl;
            L.bitmap.assign((size_t)1 << (bl - 5), 0u);
            std::vector<std::vector<std::pair<uint32_t, uint32_t>>> buckets((size_t)1 << lg);
            for (int b = 0; b < hs.n_bc; b++)
                for (int i = 0; i <= K_l; i++) {
                    const int o = i * seg;
                    uint32_t h = 0;
                    for (int k = 0; k < q; k++) h = h * kPfBase + (uint32_t)hs.bc_cls[hs.off[b] + o + k];
                    const uint32_t bit = pf_bit(h, bl);
                    L.bitmap[bit >> 5] |= 1u << (bit & 31);
                    buckets[pf_slot(h, lg)].emplace_back(((uint32_t)b << 8) | (uint32_t)o, h);
                }
            L.bstart.assign(((size_t)1 << lg) + 1, 0u);
            for (size_t k = 0; k < buckets.size(); k++) {
                L.bstart[k + 1] = L.bstart[k] + (uint32_t)buckets[k].size();
                for (auto &pr : buckets[k]) {
                    L.entries.push_back(pr.first);
                    L.ekeys.push_back(pr.second);
                }
            }
        }
    }
    // ---- deepest seed level (seed_deep.cu): depth beyond the regular levels, each segment hashed with its own
    // length.  Measured on B200: with 0.75 chance hits per column (96 x 24 nt at depth 4) it is slower than the
    // bit-parallel kernel it would replace (17 vs 11.5 ms per 10 M-read step), so it is used while the hits
    // stay rare (<= 0.25 per column) -- small sets, e.g. one adapter, whose alternative is k_literal ----
    if (hs.sd_levels > 0 && !getenv("BDX_DISABLE_SEED_DEEP")) {
        const int m = hs.max_m, allowed = hs.allowed0[0];
        const double alpha = std::max(2, hs.n_classes - 1);
        const char *dr = getenv("BDX_SEED_DEEP_RATE");      // experiments: chance hits per column the deep level accepts
        const double deep_rate = dr ? atof(dr) : 0.25;
        int KD = 0;
        for (int k = hs.sd[hs.sd_levels - 1].k + 1; k <= std::min(allowed, 7); k++) {
            const int n_seg = k + 1, base_len = m / n_seg, extra = m % n_seg;
            if (base_len < 4) break;
            double rate = 0.0;
            for (int i = 0; i < n_seg; i++) rate += hs.n_bc / std::pow(alpha, std::min(base_len + (i < extra ? 1 : 0), 8));
            if (rate <= deep_rate) KD = k;
        }
        if (KD > 0) {
            const int n_seg = KD + 1, base_len = m / n_seg, extra = m % n_seg;
            const int q_long = std::min(base_len + 1, 8), q_short = std::min(base_len, 8);
            hs.sdd_k = KD;
            // table 0: the longer seeds (if any segment is longer and that changes the seed length), table 1 / 0: the rest
            struct Seg { int off, q; };
            std::vector<Seg> segs[2];
            int o = 0;
            for (int i = 0; i < n_seg; i++) {
                const int len = base_len + (i < extra ? 1 : 0);
                const int q = std::min(len, 8);
                segs[(q == q_long && q_long != q_short) ? 0 : 1].push_back(Seg{o, q});
                o += len;
            }
            for (int t = 0; t < 2; t++) {
                if (segs[t].empty()) continue;
                HostSet::HostSeedLevel &L = hs.sdd[hs.sdd_n++];
                const int q = segs[t][0].q;
                L.k = KD;
                L.q = q;
                uint32_t pw = 1;
                for (int i = 1; i < q; i++) pw *= kPfBase;
                L.pow = pw;
                const size_t n_entries = (size_t)hs.n_bc * segs[t].size();
                int lg = 8;
                while (lg < 14 && (size_t)(1 << lg) < n_entries) lg++;
                L.log2 = lg;
                int bl = 13;
                while (bl < 18 && ((size_t)1 << bl) < 64 * n_entries) bl++;
                L.bm_log2 = bl;
                L.bitmap.assign((size_t)1 << (bl - 5), 0u);
                std::vector<std::vector<std::pair<uint32_t, uint32_t>>> buckets((size_t)1 << lg);
                for (int b = 0; b < hs.n_bc; b++)
                    for (const Seg &sg2 : segs[t]) {
                        uint32_t h = 0;
                        for (int k = 0; k < q; k++) h = h * kPfBase + (uint32_t)hs.bc_cls[hs.off[b] + sg2.off + k];
                        const uint32_t bit = pf_bit(h, bl);
                        L.bitmap[bit >> 5] |= 1u << (bit & 31);
                        buckets[pf_slot(h, lg)].emplace_back(((uint32_t)b << 8) | (uint32_t)sg2.off, h);
                    }
                L.bstart.assign(((size_t)1 << lg) + 1, 0u);
                for (size_t k = 0; k < buckets.size(); k++) {
                    L.bstart[k + 1] = L.bstart[k] + (uint32_t)buckets[k].size();
                    for (auto &pr : buckets[k]) {
                        L.entries.push_back(pr.first);
                        L.ekeys.push_back(pr.second);
                    }
                }
            }
        }
    }
    // ---- seed-and-verify for sets of different lengths and constrained start / end geometries (seed_var.cu):
    // K_b + 1 disjoint segments per barcode, K_b = min(m_b / q - 1, allowed_b), their first q bases in a
    // direct-address table.  Level 1: q = the shortest seed whose CHANCE hits on admissible diagonals (estimated
    // for a 150-base read) stay around two dozen per read -- position constraints keep short seeds selective.
    // Level 2 (reads level 1 could not decide): the longest seed that is COMPLETE (K_b = allowed_b for every
    // barcode, so the candidates are a superset and every verdict is final), used while verifying its chance
    // hits costs less than half the lane-per-barcode automaton over the whole range ----
    if (sg && hs.words >= 1 && !p.has_nindel && hs.n_classes - 1 <= 4 && hs.n_bc < (1 << 14) && hs.max_m <= 64 &&
        p.max_error_rate >= 0.0 && !getenv("BDX_DISABLE_SEED")) {
        auto resolve = [](const DevRange &dr, int len, int &first, int &last) {      // classification.jl:96-100
            const int s = dr.start_from_end ? len + dr.start_off : dr.start_off;
            const int e = dr.end_from_end ? len + dr.end_off : dr.end_off;
            first = std::max(1, s);
            last = std::min(len, e);
            if (last < first) last = first - 1;
        };
        const int n_nom = 150;
        int rf, rl, bf, bl, ef, el;
        resolve(hs.rs, n_nom, rf, rl);
        resolve(hs.bs, n_nom, bf, bl);
        resolve(hs.be, n_nom, ef, el);
        const int start_j = std::max(rf, std::max(bf, 1)), end_j = std::min(rl, std::min(el, n_nom));
        const int L = std::max(end_j - start_j + 1, 1), sbase = start_j - 1;
        const int min_end_rel = ef - sbase, max_start_rel = bl - sbase;
        struct Est { double chance, steps; size_t n_entries; bool complete; };
        auto estimate = [&](int q) {
            Est e{0.0, 0.0, 0, true};
            for (int b = 0; b < hs.n_bc; b++) {
                const int m = hs.off[b + 1] - hs.off[b], a0 = hs.allowed0[b];
                const int K = std::min(m / q - 1, a0);
                const int dlo = std::max(0, min_end_rel - m) - K, dhi = std::min(max_start_rel + a0, L - m + K);
                const double c = (double)(K + 1) * std::max(0, dhi - dlo + 1) / std::pow(4.0, q);
                e.chance += c;
                e.steps += c * (m + 2 * K);          // columns verified for those hits
                e.n_entries += (size_t)K + 1;
                if (K < a0) e.complete = false;
            }
            return e;
        };
        auto build = [&](int q, double chance) {
            HostSet::HostSeedVar &V = hs.sv[hs.sv_levels++];
            V.q = q;
            V.kdepth.assign((size_t)hs.n_bc, 0);
            std::vector<std::vector<uint32_t>> buckets((size_t)1 << (2 * q));
            V.sigma_min = 1e300;
            V.complete = 1;
            for (int b = 0; b < hs.n_bc; b++) {
                const int m = hs.off[b + 1] - hs.off[b], a0 = hs.allowed0[b];
                const int K = std::min(m / q - 1, a0);
                V.kdepth[(size_t)b] = (uint8_t)K;
                if (K < a0) V.complete = 0;
                V.sigma_min = std::min(V.sigma_min, (double)(K + 1) / (double)hs.norm[b]);
                const int seg = m / (K + 1);                       // >= q: the segments are disjoint
                for (int i = 0; i <= K; i++) {
                    const int o = i * seg;
                    uint32_t code = 0;
                    for (int k = 0; k < q; k++) code |= ((uint32_t)(hs.bc_cls[hs.off[b] + o + k] - 1) & 3u) << (2 * k);
                    buckets[code].push_back(((uint32_t)b << 8) | (uint32_t)o);
                }
            }
            V.bstart.assign(buckets.size() + 1, 0);
            for (size_t k = 0; k < buckets.size(); k++) {
                V.bstart[k + 1] = (uint16_t)(V.bstart[k] + buckets[k].size());
                V.entries.insert(V.entries.end(), buckets[k].begin(), buckets[k].end());
            }
            // 128 reads per group and a hit list of up to 64 rows x 128 records (seed_var.cu) that their hits -- chance
            // + a handful of true ones -- fill to about 70 %; denser levels take fewer reads per group instead
            const double per_read = chance + 6.0;
            V.hit_rows = std::min(64, std::max(32, (int)std::ceil(per_read / 0.7)));
            int R = 128;
            while (R > 8 && per_read * R > 0.7 * 128 * V.hit_rows) R -= 8;
            V.group_reads = R;
        };
        int q1 = 0;
        for (int q = 4; q <= 8 && !q1; q++) {
            if (min_m < q) break;
            const Est e = estimate(q);
            if (e.chance <= 24.0 && e.n_entries <= 65535) {
                q1 = q;
                build(q, e.chance);
            }
        }
        if (q1 && !hs.sv[0].complete && !getenv("BDX_SEED_ONE_LEVEL")) {
            int q2 = q1 - 1;
            for (int b = 0; b < hs.n_bc; b++) q2 = std::min(q2, (hs.off[b + 1] - hs.off[b]) / (hs.allowed0[b] + 1));
            if (q2 >= 3) {
                const Est e = estimate(q2);
                const double automaton_steps = (double)hs.n_bc * L;
                if (e.complete && e.n_entries <= 65535 && e.steps < 0.5 * automaton_steps && e.chance <= 200.0)
                    build(q2, e.chance);
            }
        }
    }
    // ---- :hamming on packed words (hamming.cu): uniform length <= 32, <= 4 distinct barcode bytes, no 'N' ----
    if (p.algorithm == BDX_HAMMING && min_m == hs.max_m && hs.max_m <= 32 && hs.n_classes - 1 <= 4 && hs.n_bc <= 65535 &&
        hs.allowed0[0] >= 0 && hs.allowed0[0] <= 7 && p.max_error_rate >= 0.0 &&
        std::find(hs.bytes.begin(), hs.bytes.end(), (uint8_t)'N') == hs.bytes.end()) {
        const int m = hs.max_m, n_seg = hs.allowed0[0] + 1;
        const int seg_len = m / n_seg, extra = m % n_seg;      // the first `extra` segments are one base longer
        if (seg_len >= 2) {
            hs.hp_m = m;
            hs.hp_allowed = hs.allowed0[0];
            hs.hp_n_seg = n_seg;
            int o = 0, base = 0;
            for (int i = 0; i < n_seg; i++) {
                const int len = seg_len + (i < extra ? 1 : 0);
                hs.hp_off[i] = o;
                hs.hp_q[i] = std::min(len, 6);                  // direct-address table of 4^q buckets
                hs.hp_base[i] = base;
                base += (1 << (2 * hs.hp_q[i])) + 1;
                o += len;
            }
            auto code_of = [&](int b, int pos) { return (uint32_t)(hs.bc_cls[hs.off[b] + pos] - 1) & 3u; };
            hs.hp_bstart.assign((size_t)base, 0);
            hs.hp_entries.assign((size_t)n_seg * hs.n_bc, 0);
            for (int i = 0; i < n_seg; i++) {
                const int nb = 1 << (2 * hs.hp_q[i]);
                std::vector<std::vector<uint16_t>> buckets((size_t)nb);
                for (int b = 0; b < hs.n_bc; b++) {
                    uint32_t gram = 0;           // bit plane 0 of the q bases, then bit plane 1
                    for (int k = 0; k < hs.hp_q[i]; k++) {
                        const uint32_t c = code_of(b, hs.hp_off[i] + k);
                        gram |= (c & 1u) << k;
                        gram |= (c >> 1) << (hs.hp_q[i] + k);
                    }
                    buckets[gram].push_back((uint16_t)b);
                }
                uint16_t run = 0;
                size_t w = (size_t)i * hs.n_bc;
                for (int gidx = 0; gidx < nb; gidx++) {
                    hs.hp_bstart[(size_t)hs.hp_base[i] + gidx] = run;
                    for (uint16_t b : buckets[(size_t)gidx]) hs.hp_entries[w++] = b;
                    run = (uint16_t)(run + buckets[(size_t)gidx].size());
                }
                hs.hp_bstart[(size_t)hs.hp_base[i] + nb] = run;
            }
            hs.hp_bcw.resize((size_t)hs.n_bc);
            for (int b = 0; b < hs.n_bc; b++) {
                uint32_t w0 = 0, w1 = 0;         // the two bit planes, base k in bit k
                for (int k = 0; k < m; k++) {
                    w0 |= (code_of(b, k) & 1u) << k;
                    w1 |= (code_of(b, k) >> 1) << k;
                }
                hs.hp_bcw[(size_t)b] = make_uint2(w0, w1);
            }
            hs.hp_enabled = 1;
        }
    }
    // ---- :hamming pigeonhole seeds: mismatches <= allowed_b leave one of allowed_b + 1 disjoint
    // segments of the barcode intact, so every acceptable placement contains an exact seed ----
    if (p.algorithm == BDX_HAMMING && hs.use_filter && p.max_error_rate >= 0.0 && !getenv("BDX_DISABLE_PREFILTER") &&
        std::find(hs.bytes.begin(), hs.bytes.end(), (uint8_t)'N') == hs.bytes.end()) {
        int q = 8;
        bool ok = true;
        size_t n_entries = 0;
        for (int b = 0; b < hs.n_bc && ok; b++) {
            const int m = hs.off[b + 1] - hs.off[b];
            const int a = hs.allowed0[b];
            if (a < 0 || a > 254) { ok = false; break; }
            q = std::min(q, m / (a + 1));
            n_entries += (size_t)a + 1;
        }
        if (ok && q >= 4 && n_entries <= (1u << 20)) {
            hs.hs_q = q;
            uint32_t pw = 1;
            for (int i = 1; i < q; i++) pw *= kPfBase;
            hs.hs_pow = pw;
            int lg = 8;
            while (lg < 13 && (size_t)(1 << lg) < n_entries) lg++;
            hs.hs_log2 = lg;
            const uint32_t nb = 1u << lg;
            std::vector<std::vector<uint32_t>> buckets(nb);
            for (int b = 0; b < hs.n_bc; b++) {
                const int m = hs.off[b + 1] - hs.off[b];
                const int a = hs.allowed0[b];
                const int seg = m / (a + 1);
                for (int i = 0; i <= a; i++) {
                    const int o = i * seg;
                    uint32_t h = 0;
                    for (int k = 0; k < q; k++) h = h * kPfBase + (uint32_t)hs.bytes[hs.off[b] + o + k];
                    buckets[pf_slot(h, lg)].push_back(((uint32_t)b << 8) | (uint32_t)o);
                    hs.hs_max_off = std::max(hs.hs_max_off, o);
                }
            }
            hs.hs_bstart.assign(nb + 1, 0u);
            for (uint32_t k = 0; k < nb; k++) {
                hs.hs_bstart[k + 1] = hs.hs_bstart[k] + (uint32_t)buckets[k].size();
                hs.hs_entries.insert(hs.hs_entries.end(), buckets[k].begin(), buckets[k].end());
            }
            hs.hs_enabled = hs.hs_max_off < 256;
        }
    }
    return BDX_OK;
}

extern "C" int bdx_config_create(const bdx_params *p, bdx_config **out)
{
    if (!p || !out) return fail(BDX_ERR_INVALID, "null argument");
    *out = nullptr;
    if (p->struct_size != sizeof(bdx_params) || p->abi_version != BDX_ABI_VERSION)
        return fail(BDX_ERR_INVALID, "bdx_params: struct_size / abi_version mismatch");
    if (p->algorithm < BDX_SEMIGLOBAL || p->algorithm > BDX_EXACT) return fail(BDX_ERR_INVALID, "unknown algorithm");
    const int64_t costs[4] = {p->match, p->mismatch, p->indel, p->has_nindel ? p->nindel : 1};
    for (int64_t c : costs)
        if (c > kMaxCost || c < -kMaxCost) return fail(BDX_ERR_INVALID, "cost magnitude above 2^20");
    if (p->algorithm == BDX_SEMIGLOBAL && (p->indel == 0 || (p->has_nindel && p->nindel == 0)))
        return fail(BDX_ERR_INVALID, "zero gap cost (the reference raises DivideError, classification.jl:170-176)");
    if (std::isnan(p->max_error_rate) || std::isnan(p->min_delta)) return fail(BDX_ERR_INVALID, "NaN option");

    bdx_config *cfg = new (std::nothrow) bdx_config();
    if (!cfg) return fail(BDX_ERR_NOMEM, "out of memory");
    const char *env = getenv("BDX_DISABLE_FILTER");
    const bool disable_filter = env && env[0] == '1';
    int rc = build_set(*p, p->set1, cfg->set[0], disable_filter, "set1");
    if (rc == BDX_OK && p->is_dual) rc = build_set(*p, p->set2, cfg->set[1], disable_filter, "set2");
    if (rc != BDX_OK) {
        delete cfg;
        return rc;
    }
    DevParams &P = cfg->base;
    P.max_error_rate = p->max_error_rate;
    P.min_delta = p->min_delta;
    P.match = (int)p->match;
    P.mismatch = (int)p->mismatch;
    P.indel = (int)p->indel;
    P.nindel = p->has_nindel ? (int)p->nindel : 0;
    P.has_n = p->has_nindel ? 1 : 0;
    P.algo = p->algorithm;
    P.is_dual = p->is_dual ? 1 : 0;
    P.want_stats = p->want_stats ? 1 : 0;
    P.filter_ok = cfg->set[0].use_filter && (!p->is_dual || cfg->set[1].use_filter);
    P.two = 2;
    P.unit_costs = p->match == 0 && p->mismatch == 1 && p->indel == 1 && (!p->has_nindel || p->nindel == 1);

    // stats layout (classification.jl:736-758)
    bdx_stats_layout &L = cfg->lay;
    const int b1 = cfg->set[0].n_bc, b2 = p->is_dual ? cfg->set[1].n_bc : 0;
    const int max_m = std::max(cfg->set[0].max_m, p->is_dual ? cfg->set[1].max_m : 0);
    int max_allowed = 0;
    for (int s = 0; s < (p->is_dual ? 2 : 1); s++)
        for (int b = 0; b < cfg->set[s].n_bc; b++) max_allowed = std::max(max_allowed, cfg->set[s].allowed0[b]);
    L.b1 = b1;
    L.b2 = b2;
    L.pos_bias = max_m;
    L.pos_bins = 1024 + max_m + 2;
    // e - s + 1 with s >= 1 - m (origin labels of the init column, classification.jl:281) and e <= n
    L.len_bins = 1024 + max_m + 2;
    L.dist_bias = (int)std::min<int64_t>(std::max<int64_t>(0, -p->match) * max_m, 4096);
    L.dist_bins = L.dist_bias + std::min(std::max(max_allowed, 0), 4095) + 1;
    int64_t o = 4;
    L.sample_off = o;
    o += (int64_t)(b1 + 1) * (b2 + 1);
    const int bsz[2] = {b1, b2};
    for (int s = 0; s < 2; s++) {
        L.pos_off[s] = o;
        o += (int64_t)(bsz[s] + 1) * L.pos_bins;
        L.len_off[s] = o;
        o += (int64_t)(bsz[s] + 1) * L.len_bins;
        L.dist_off[s] = o;
        o += (int64_t)(bsz[s] + 1) * L.dist_bins;
    }
    L.total_len = o;
    *out = cfg;
    return BDX_OK;
}

static void free_tables(DeviceTables *t)
{
    for (void *p : t->allocs) cudaFree(p);
    delete t;
}

extern "C" void bdx_config_destroy(bdx_config *cfg)
{
    if (!cfg) return;
    for (auto &kv : cfg->per_device) {
        cudaSetDevice(kv.first);
        free_tables(kv.second);
    }
    delete cfg;
}

template <typename T>
static cudaError_t upload(DeviceTables *t, const std::vector<T> &v, const T **out)
{
    *out = nullptr;
    if (v.empty()) return cudaSuccess;
    void *d = nullptr;
    cudaError_t e = cudaMalloc(&d, v.size() * sizeof(T));
    if (e != cudaSuccess) return e;
    t->allocs.push_back(d);
    e = cudaMemcpy(d, v.data(), v.size() * sizeof(T), cudaMemcpyHostToDevice);
    *out = (const T *)d;
    return e;
}

static int get_tables(bdx_config *cfg, int device, DeviceTables **out)
{
    std::lock_guard<std::mutex> lk(cfg->mu);
    auto it = cfg->per_device.find(device);
    if (it != cfg->per_device.end()) {
        *out = it->second;
        return BDX_OK;
    }
    CU(cudaSetDevice(device));
    cudaDeviceProp prop;
    CU(cudaGetDeviceProperties(&prop, device));
    if (prop.major < 10)
        return fail(BDX_ERR_CUDA, "libbdx is built for sm_100a only; device compute capability too low");
    DeviceTables *t = new DeviceTables();
    t->P = cfg->base;
    t->sm_count = prop.multiProcessorCount;
    for (int s = 0; s < (cfg->base.is_dual ? 2 : 1); s++) {
        HostSet &hs = cfg->set[s];
        DevSet &D = t->P.set[s];
        D.n_bc = hs.n_bc;
        D.n_bc_pad = hs.n_bc_pad;
        D.max_m = hs.max_m;
        D.trim_side = hs.trim_side;
        D.words = hs.words;
        D.use_filter = hs.use_filter;
        D.n_classes = hs.n_classes;
        D.rs = hs.rs;
        D.bs = hs.bs;
        D.be = hs.be;
        std::vector<uint8_t> cls(hs.class_of, hs.class_of + 256);
        cudaError_t e = upload(t, hs.bytes, &D.bc_bytes);
        if (e == cudaSuccess) e = upload(t, hs.off, &D.bc_off);
        if (e == cudaSuccess) e = upload(t, hs.norm, &D.norm);
        if (e == cudaSuccess) e = upload(t, hs.peq, &D.peq);
        if (e == cudaSuccess) e = upload(t, hs.filt_allowed, &D.filt_allowed);
        if (e == cudaSuccess) e = upload(t, hs.allowed0, &D.allowed0);
        if (e == cudaSuccess) e = upload(t, cls, &D.class_of);
        D.pf_enabled = hs.pf_enabled;
        D.pf_seed = hs.pf_seed;
        D.pf_pow = hs.pf_pow;
        D.pf_log2 = hs.pf_log2;
        D.pf_bm_log2 = hs.pf_bm_log2;
        if (e == cudaSuccess) e = upload(t, hs.pf_bitmap, &D.pf_bitmap);
        if (e == cudaSuccess) e = upload(t, hs.pf_keys, &D.pf_keys);
        if (e == cudaSuccess) e = upload(t, hs.pf_vals, &D.pf_vals);
        if (e == cudaSuccess) e = upload(t, hs.bc_cls, &D.bc_cls);
        D.sd_levels = hs.sd_levels;
        D.sd_m = hs.sd_m;
        for (int l = 0; l < hs.sd_levels; l++) {
            const HostSet::HostSeedLevel &H = hs.sd[l];
            SeedLevel &L = D.sd[l];
            L.k = H.k;
            L.q = H.q;
            L.pow = H.pow;
            L.log2 = H.log2;
            L.bm_log2 = H.bm_log2;
            L.n_entries = (int)H.entries.size();
            if (e == cudaSuccess) e = upload(t, H.bstart, &L.bstart);
            if (e == cudaSuccess) e = upload(t, H.entries, &L.entries);
            if (e == cudaSuccess) e = upload(t, H.ekeys, &L.ekeys);
            if (e == cudaSuccess) e = upload(t, H.bitmap, &L.bitmap);
        }
        D.sdd_n = hs.sdd_n;
        D.sdd_k = hs.sdd_k;
        for (int l = 0; l < hs.sdd_n; l++) {
            const HostSet::HostSeedLevel &H = hs.sdd[l];
            SeedLevel &L = D.sdd[l];
            L.k = H.k;
            L.q = H.q;
            L.pow = H.pow;
            L.log2 = H.log2;
            L.bm_log2 = H.bm_log2;
            L.n_entries = (int)H.entries.size();
            if (e == cudaSuccess) e = upload(t, H.bstart, &L.bstart);
            if (e == cudaSuccess) e = upload(t, H.entries, &L.entries);
            if (e == cudaSuccess) e = upload(t, H.ekeys, &L.ekeys);
            if (e == cudaSuccess) e = upload(t, H.bitmap, &L.bitmap);
        }
        D.sv_levels = hs.sv_levels;
        for (int l = 0; l < hs.sv_levels; l++) {
            const HostSet::HostSeedVar &H = hs.sv[l];
            SeedVar &V = D.sv[l];
            V.enabled = 1;
            V.q = H.q;
            V.n_buckets = 1 << (2 * H.q);
            V.n_entries = (int)H.entries.size();
            V.complete = H.complete;
            V.group_reads = H.group_reads;
            V.hit_rows = H.hit_rows;
            V.sigma_min = H.sigma_min;
            if (e == cudaSuccess) e = upload(t, H.bstart, &V.bstart);
            if (e == cudaSuccess) e = upload(t, H.entries, &V.entries);
            if (e == cudaSuccess) e = upload(t, H.kdepth, &V.kdepth);
        }
        D.hp.enabled = hs.hp_enabled;
        D.hp.m = hs.hp_m;
        D.hp.allowed = hs.hp_allowed;
        D.hp.n_seg = hs.hp_n_seg;
        D.hp.n_bstart = (int)hs.hp_bstart.size();
        for (int k = 0; k < 8; k++) {
            D.hp.seg_off[k] = hs.hp_off[k];
            D.hp.seg_q[k] = hs.hp_q[k];
            D.hp.seg_base[k] = hs.hp_base[k];
        }
        if (e == cudaSuccess) e = upload(t, hs.hp_bstart, &D.hp.bstart);
        if (e == cudaSuccess) e = upload(t, hs.hp_entries, &D.hp.entries);
        if (e == cudaSuccess) e = upload(t, hs.hp_bcw, &D.hp.bcw);
        D.hs_enabled = hs.hs_enabled;
        D.hs_q = hs.hs_q;
        D.hs_pow = hs.hs_pow;
        D.hs_log2 = hs.hs_log2;
        D.hs_n_entries = (int)hs.hs_entries.size();
        D.hs_max_off = hs.hs_max_off;
        if (e == cudaSuccess) e = upload(t, hs.hs_bstart, &D.hs_bstart);
        if (e == cudaSuccess) e = upload(t, hs.hs_entries, &D.hs_entries);
        if (e != cudaSuccess) {
            free_tables(t);
            return cuda_fail(e, "uploading barcode tables");
        }
        if (D.words && filter_smem_bytes_for(D) > (size_t)prop.sharedMemPerBlockOptin) {
            D.pf_enabled = 0;   // drop the prefilter table first, then the filter itself
            if (filter_smem_bytes_for(D) > (size_t)prop.sharedMemPerBlockOptin) {
                D.words = 0;
                D.use_filter = 0;
            }
        }
    }
    t->P.filter_ok = t->P.set[0].use_filter && (!t->P.is_dual || t->P.set[1].use_filter);
    cfg->per_device[device] = t;
    *out = t;
    return BDX_OK;
}

// ---------------------------------------------------------------------------
// streams
// ---------------------------------------------------------------------------
struct Slot {
    uint8_t *h_seq = nullptr;
    int32_t *h_off = nullptr;
    bdx_result *h_res = nullptr;
    bdx_pass_detail *h_det = nullptr;
    uint8_t *d_seq = nullptr;
    int32_t *d_off = nullptr;
    bdx_result *d_res = nullptr;
    bdx_pass_detail *d_det = nullptr;
    cudaEvent_t ev_h2d = nullptr, ev_kern = nullptr, ev_done = nullptr;
    int32_t n = 0;
    uint64_t tag = 0;
    bool busy = false;
    // the kernel sequence of a batch of graph_n reads on this slot's buffers, captured as a CUDA graph: small
    // batches (the reference's 4000-read chunks) are bound by launch gaps, not by the kernels
    cudaGraphExec_t graph = nullptr;
    int32_t graph_n = -1;         // batch size the graph was captured for
    bool graph_details = false;
    int graph_launches = 0;       // kernel launches it holds
    int uses = 0;                 // plain runs of this slot so far (the first warms the launch caches)
};

// stage of a profiled kernel launch (bdx_stream_profile_read_stages)
enum { kStPrefilter = 0, kStSeed = 1, kStSeedDeep = 2, kStFilter = 3, kStLiteral = 4, kStHamming = 5, kStFinalize = 6,
       kStOther = 7 };
struct ProfEvent {
    int kind;
    cudaEvent_t e0, e1;
};

struct bdx_stream {
    bdx_config *cfg = nullptr;
    DeviceTables *tab = nullptr;
    int device = 0;
    int32_t max_reads = 0;
    int64_t max_bytes = 0;
    bool details = false;
    cudaStream_t st_copy = nullptr, st_comp = nullptr, st_d2h = nullptr;
    Slot slot[BDX_MAX_IN_FLIGHT];
    bool host_staging = false;  // pinned h_seq / h_off are allocated on first use
    int head = 0;      // next slot to submit into
    int tail = 0;      // oldest in-flight slot
    int in_flight = 0;
    bool acquired = false;
    Scratch sc{};
    int64_t sc_cap = 0;
    unsigned long long *d_stats = nullptr;
    bdx_stats_overflow *d_ovf = nullptr;       // exact records of passes outside the pos / len histograms
    unsigned int *d_n_ovf = nullptr;           // [2] appended, lost
    unsigned long long *d_counters = nullptr;  // [0] reads resolved by the perfect-occurrence prefilter,
                                               // [1] reads that ran the bit-parallel automaton
    int64_t launches = 0;
    // optional per-kernel timing of the dominant (filter) kernel, for roofline reporting
    bool profile = false;
    bool graphs_ok = true;                     // cleared when a capture fails: plain launches from then on
    std::vector<ProfEvent> prof_events;
    DemuxState *demux = nullptr;               // device FASTQ block demultiplexer (demux.cu), created on first use
};

static int ensure_scratch(bdx_stream *s, int64_t n)
{
    if (n <= s->sc_cap) return BDX_OK;
    // in-order on the compute stream: earlier kernels still own the old buffers
    CU(cudaStreamSynchronize(s->st_comp));
    cudaFree(s->sc.pass[0]);
    cudaFree(s->sc.pass[1]);
    cudaFree(s->sc.cand);
    cudaFree(s->sc.cand_cnt);
    cudaFree(s->sc.worklist);
    cudaFree(s->sc.n_work);
    cudaFree(s->sc.worklist2);
    cudaFree(s->sc.n_work2);
    cudaFree(s->sc.wl_win);
    cudaFree(s->sc.wl_full);
    cudaFree(s->sc.n_lit);
    s->sc = Scratch{};
    s->sc_cap = 0;
    const int64_t cap = n + n / 8 + 1024;
    CU(cudaMalloc(&s->sc.pass[0], cap * sizeof(PassOut)));
    CU(cudaMalloc(&s->sc.pass[1], cap * sizeof(PassOut)));
    CU(cudaMalloc(&s->sc.cand, cap * kCandMax * sizeof(uint16_t)));
    CU(cudaMalloc(&s->sc.cand_cnt, cap));
    CU(cudaMalloc(&s->sc.worklist, cap * sizeof(int)));
    CU(cudaMalloc(&s->sc.n_work, sizeof(int)));
    CU(cudaMalloc(&s->sc.worklist2, cap * sizeof(int)));
    CU(cudaMalloc(&s->sc.n_work2, sizeof(int)));
    CU(cudaMalloc(&s->sc.wl_win, cap * sizeof(int)));
    CU(cudaMalloc(&s->sc.wl_full, cap * sizeof(int)));
    CU(cudaMalloc(&s->sc.n_lit, 2 * sizeof(int)));
    s->sc_cap = cap;
    return BDX_OK;
}

extern "C" void bdx_stream_destroy(bdx_stream *s)
{
    if (!s) return;
    cudaSetDevice(s->device);
    if (s->st_comp) cudaStreamSynchronize(s->st_comp);
    if (s->st_copy) cudaStreamSynchronize(s->st_copy);
    if (s->st_d2h) cudaStreamSynchronize(s->st_d2h);
    for (Slot &sl : s->slot) {
        cudaFreeHost(sl.h_seq);
        cudaFreeHost(sl.h_off);
        cudaFreeHost(sl.h_res);
        cudaFreeHost(sl.h_det);
        cudaFree(sl.d_seq);
        cudaFree(sl.d_off);
        cudaFree(sl.d_res);
        cudaFree(sl.d_det);
        if (sl.ev_h2d) cudaEventDestroy(sl.ev_h2d);
        if (sl.ev_kern) cudaEventDestroy(sl.ev_kern);
        if (sl.ev_done) cudaEventDestroy(sl.ev_done);
        if (sl.graph) cudaGraphExecDestroy(sl.graph);
    }
    cudaFree(s->sc.pass[0]);
    cudaFree(s->sc.pass[1]);
    cudaFree(s->sc.cand);
    cudaFree(s->sc.cand_cnt);
    cudaFree(s->sc.worklist);
    cudaFree(s->sc.n_work);
    cudaFree(s->sc.worklist2);
    cudaFree(s->sc.n_work2);
    cudaFree(s->sc.wl_win);
    cudaFree(s->sc.wl_full);
    cudaFree(s->sc.n_lit);
    cudaFree(s->d_stats);
    cudaFree(s->d_ovf);
    cudaFree(s->d_n_ovf);
    cudaFree(s->d_counters);
    demux_state_destroy(s->demux);
    for (auto &pr : s->prof_events) {
        cudaEventDestroy(pr.e0);
        cudaEventDestroy(pr.e1);
    }
    if (s->st_copy) cudaStreamDestroy(s->st_copy);
    if (s->st_comp) cudaStreamDestroy(s->st_comp);
    if (s->st_d2h) cudaStreamDestroy(s->st_d2h);
    delete s;
}

static int stream_create_impl(bdx_stream *s)
{
    CU(cudaSetDevice(s->device));
    CU(cudaStreamCreateWithFlags(&s->st_copy, cudaStreamNonBlocking));
    CU(cudaStreamCreateWithFlags(&s->st_comp, cudaStreamNonBlocking));
    CU(cudaStreamCreateWithFlags(&s->st_d2h, cudaStreamNonBlocking));
    if (s->max_reads > 0) {
        for (Slot &sl : s->slot) {
            CU(cudaHostAlloc(&sl.h_res, (size_t)s->max_reads * sizeof(bdx_result), cudaHostAllocDefault));
            CU(cudaHostAlloc(&sl.h_det, (size_t)s->max_reads * 2 * sizeof(bdx_pass_detail), cudaHostAllocDefault));
            CU(cudaMalloc(&sl.d_seq, (size_t)std::max<int64_t>(s->max_bytes, 16)));
            CU(cudaMalloc(&sl.d_off, ((size_t)s->max_reads + 1) * 4));
            CU(cudaMalloc(&sl.d_res, (size_t)s->max_reads * sizeof(bdx_result)));
            CU(cudaMalloc(&sl.d_det, (size_t)s->max_reads * 2 * sizeof(bdx_pass_detail)));
            CU(cudaEventCreateWithFlags(&sl.ev_h2d, cudaEventDisableTiming));
            CU(cudaEventCreateWithFlags(&sl.ev_kern, cudaEventDisableTiming));
            CU(cudaEventCreateWithFlags(&sl.ev_done, cudaEventDisableTiming));
        }
        int rc = ensure_scratch(s, s->max_reads);
        if (rc) return rc;
    }
    CU(cudaMalloc(&s->d_counters, 4 * sizeof(unsigned long long)));
    CU(cudaMemset(s->d_counters, 0, 4 * sizeof(unsigned long long)));
    if (s->cfg->base.want_stats) {
        CU(cudaMalloc(&s->d_stats, (size_t)s->cfg->lay.total_len * 8));
        CU(cudaMemset(s->d_stats, 0, (size_t)s->cfg->lay.total_len * 8));
        CU(cudaMalloc(&s->d_ovf, (size_t)kStatsOvfCap * sizeof(bdx_stats_overflow)));
        CU(cudaMalloc(&s->d_n_ovf, 2 * sizeof(unsigned int)));
        CU(cudaMemset(s->d_n_ovf, 0, 2 * sizeof(unsigned int)));
    }
    return BDX_OK;
}

extern "C" int bdx_stream_create(const bdx_config *cfg_c, int device, int32_t max_reads, int64_t max_bytes,
                                 bdx_stream **out)
{
    if (!cfg_c || !out || max_reads < 0 || max_bytes < 0) return fail(BDX_ERR_INVALID, "bad argument");
    *out = nullptr;
    if (max_bytes > 0x7FFFFFF0ll) return fail(BDX_ERR_TOO_LARGE, "max_bytes must stay below 2^31 (int32 offsets)");
    bdx_config *cfg = const_cast<bdx_config *>(cfg_c);
    int ndev = 0;
    cudaError_t e = cudaGetDeviceCount(&ndev);
    if (e != cudaSuccess || ndev == 0) {
        cudaGetLastError();
        return fail(BDX_ERR_CUDA, "no CUDA device available (libbdx has no CPU fallback)");
    }
    if (device < 0 || device >= ndev) return fail(BDX_ERR_INVALID, "device index out of range");
    DeviceTables *tab = nullptr;
    int rc = get_tables(cfg, device, &tab);
    if (rc) return rc;
    bdx_stream *s = new (std::nothrow) bdx_stream();
    if (!s) return fail(BDX_ERR_NOMEM, "out of memory");
    s->cfg = cfg;
    s->tab = tab;
    s->device = device;
    s->max_reads = max_reads;
    s->max_bytes = max_bytes;
    s->details = cfg->base.want_stats != 0;
    rc = stream_create_impl(s);
    if (rc) {
        std::string keep = g_err;
        bdx_stream_destroy(s);
        g_err = keep;
        return rc;
    }
    *out = s;
    return BDX_OK;
}

extern "C" int bdx_stream_enable_details(bdx_stream *s, int on)
{
    if (!s) return fail(BDX_ERR_INVALID, "null stream");
    if (s->in_flight) return fail(BDX_ERR_STATE, "batches in flight");
    s->details = on != 0;
    return BDX_OK;
}

// Enqueue the classification kernels for n reads resident on the device.
static int enqueue_classify(bdx_stream *s, const uint8_t *d_seq, const int32_t *d_off, int32_t n,
                            bdx_result *d_res, bdx_pass_detail *d_det)
{
    if (n == 0) return BDX_OK;
    int rc = ensure_scratch(s, n);
    if (rc) return rc;
    // every kernel launch goes through here: counted, and bracketed by CUDA events while profiling is on
    auto staged = [s](int kind, auto &&launch) -> int {
        cudaEvent_t e0 = nullptr, e1 = nullptr;
        if (s->profile) {
            CU(cudaEventCreate(&e0));
            CU(cudaEventCreate(&e1));
            CU(cudaEventRecord(e0, s->st_comp));
        }
        const cudaError_t e = launch();
        if (e != cudaSuccess) return cuda_fail(e, "kernel launch");
        s->launches++;
        if (s->profile) {
            CU(cudaEventRecord(e1, s->st_comp));
            s->prof_events.push_back(ProfEvent{kind, e0, e1});
        }
        return BDX_OK;
    };
    const DevParams &P = s->tab->P;
    const int passes = P.is_dual ? 2 : 1;
    for (int pass = 0; pass < passes; pass++) {
        if (hamming_packed_applies(P, pass)) {
            // :hamming -- packed pigeonhole scan over every start position, then the literal rules on the candidates
            if ((rc = staged(kStHamming, [&] { return launch_hamming_scan(P, pass, d_seq, d_off, n, s->sc, s->tab->sm_count, s->st_comp); }))) return rc;
            if ((rc = staged(kStLiteral, [&] { return launch_literal(P, pass, 1, d_seq, d_off, n, s->sc, s->st_comp); }))) return rc;
        } else if (hamming_seed_applies(P, pass)) {
            // :hamming -- pigeonhole seeds + in-place verification, then the literal rules on the candidates
            if ((rc = staged(kStHamming, [&] { return launch_seed_hamming(P, pass, d_seq, d_off, n, s->sc, s->tab->sm_count, s->st_comp); }))) return rc;
            if ((rc = staged(kStLiteral, [&] { return launch_literal(P, pass, 1, d_seq, d_off, n, s->sc, s->st_comp); }))) return rc;
        } else if (exact_hash_applies(P, pass)) {
            // :exact -- rolling-hash candidate generation, then the literal rules on the candidates
            if ((rc = staged(kStPrefilter, [&] { return launch_prefilter(P, pass, d_seq, d_off, n, s->sc, s->tab->sm_count, s->d_counters, s->st_comp); }))) return rc;
            if ((rc = staged(kStLiteral, [&] { return launch_literal(P, pass, 1, d_seq, d_off, n, s->sc, s->st_comp); }))) return rc;
        } else if (P.set[pass].words > 0) {
            const bool pre = prefilter_applies(P, pass);
            CU(cudaMemsetAsync(s->sc.n_lit, 0, 2 * sizeof(int), s->st_comp));   // k_literal's two read lists
            if (pre) {
                if ((rc = staged(kStPrefilter, [&] { return launch_prefilter(P, pass, d_seq, d_off, n, s->sc, s->tab->sm_count, s->d_counters, s->st_comp); }))) return rc;
            }
            int wl = pre ? 1 : 0;
            {
                // seed levels hand the reads they cannot finish from one worklist to the other; without a
                // prefilter in front (min_delta != 0) the first level takes every read of the batch
                // barcodes of different lengths / constrained start or end: k_seed_var instead of the levels
                const int sv_levels = seed_var_levels(P, pass);
                for (int l = 0; l < sv_levels; l++) {
                    const int *wl_in = wl == 0 ? nullptr : (wl == 1 ? s->sc.worklist : s->sc.worklist2);
                    const int *n_in = wl == 0 ? nullptr : (wl == 1 ? s->sc.n_work : s->sc.n_work2);
                    const bool to2 = wl != 2;
                    if ((rc = staged(l == 0 ? kStSeed : kStSeedDeep, [&] { return launch_seed_var(P, pass, l, d_seq, d_off, n, s->sc, wl_in, n_in,
                                   to2 ? s->sc.worklist2 : s->sc.worklist, to2 ? s->sc.n_work2 : s->sc.n_work,
                                   s->tab->sm_count, s->d_counters, s->st_comp); }))) return rc;
                    wl = to2 ? 2 : 1;
                }
                const int levels = sv_levels ? 0 : seed_levels(P, pass);
                for (int l = 0; l < levels; l++) {
                    const int *wl_in = wl == 0 ? nullptr : (wl == 1 ? s->sc.worklist : s->sc.worklist2);
                    const int *n_in = wl == 0 ? nullptr : (wl == 1 ? s->sc.n_work : s->sc.n_work2);
                    const bool to2 = wl != 2;
                    if ((rc = staged(kStSeed, [&] { return launch_seed(P, pass, l, d_seq, d_off, n, s->sc, wl_in, n_in, to2 ? s->sc.worklist2 : s->sc.worklist,
                                   to2 ? s->sc.n_work2 : s->sc.n_work, s->tab->sm_count, s->d_counters, s->st_comp); }))) return rc;
                    wl = to2 ? 2 : 1;
                }
                if (levels > 0 && seed_deep_applies(P, pass)) {
                    const bool to2 = wl != 2;
                    if ((rc = staged(kStSeedDeep, [&] { return launch_seed_deep(P, pass, d_seq, d_off, n, s->sc, wl == 1 ? s->sc.worklist : s->sc.worklist2,
                                        wl == 1 ? s->sc.n_work : s->sc.n_work2, to2 ? s->sc.worklist2 : s->sc.worklist,
                                        to2 ? s->sc.n_work2 : s->sc.n_work, s->tab->sm_count, s->d_counters, s->st_comp); }))) return rc;
                    wl = to2 ? 2 : 1;
                }
            }
            if (!P.set[pass].use_filter) {
                // tiny set: no filter kernel.  What the shortcut stages left goes to k_literal over every barcode.
                if (wl == 0) {
                    if ((rc = staged(kStLiteral, [&] { return launch_literal(P, pass, 0, d_seq, d_off, n, s->sc, s->st_comp); }))) return rc;
                } else {
                    const int *rest = wl == 1 ? s->sc.worklist : s->sc.worklist2;
                    const int *n_rest = wl == 1 ? s->sc.n_work : s->sc.n_work2;
                    if ((rc = staged(kStOther, [&] { return launch_mark_pending(P, pass, n, s->sc, rest, n_rest, s->st_comp); }))) return rc;
                    // seed winners (windowed) and the scan-everything rest as separate, compacted launches
                    if ((rc = staged(kStLiteral, [&] { return launch_literal(P, pass, 1, d_seq, d_off, n, s->sc, s->st_comp, s->sc.wl_win, s->sc.n_lit); }))) return rc;
                    if ((rc = staged(kStLiteral, [&] { return launch_literal(P, pass, 1, d_seq, d_off, n, s->sc, s->st_comp, rest, n_rest); }))) return rc;
                }
                continue;
            }
            if ((rc = staged(kStFilter, [&] { return launch_filter(P, pass, d_seq, d_off, n, s->sc, s->tab->sm_count, s->d_counters, wl, s->st_comp); }))) return rc;
            // the exact regime finishes inside the filter kernel; anything else leaves
            // kBcPending reads with candidate lists for the literal kernel
            const bool may_finish = P.algo == BDX_SEMIGLOBAL && P.unit_costs && P.set[pass].trim_side == 0 &&
                                    !P.want_stats;
            const DevRange &bs = P.set[pass].bs, &be = P.set[pass].be;
            const bool default_geometry = bs.start_off <= 1 && !bs.start_from_end && bs.end_from_end &&
                                          bs.end_off >= 0 && be.start_off <= 1 && !be.start_from_end;
            if (!(may_finish && default_geometry)) {
                // seed winners (windowed DP) and k_filter's candidate reads as separate, compacted launches:
                // a warp costs as much as its most expensive lane
                if (wl != 0 && seed_levels(P, pass) > 0) {
                    if ((rc = staged(kStLiteral, [&] { return launch_literal(P, pass, 1, d_seq, d_off, n, s->sc, s->st_comp, s->sc.wl_win, s->sc.n_lit); }))) return rc;
                }
                if ((rc = staged(kStLiteral, [&] { return launch_literal(P, pass, 1, d_seq, d_off, n, s->sc, s->st_comp, s->sc.wl_full, s->sc.n_lit + 1); }))) return rc;
            }
        } else {
            if ((rc = staged(kStLiteral, [&] { return launch_literal(P, pass, 0, d_seq, d_off, n, s->sc, s->st_comp); }))) return rc;
        }
    }
    StatsDev sd{s->d_stats, s->cfg->lay, s->d_ovf, s->d_n_ovf};
    if ((rc = staged(kStFinalize, [&] { return launch_finalize(P, d_off, n, s->sc, d_res, d_det, sd, s->st_comp); }))) return rc;
    return BDX_OK;
}

// The kernels of one staged batch.  Batches up to kGraphMaxReads reads replay a CUDA graph captured on the
// slot's second use with that size (the first use runs plainly and warms the per-kernel launch caches, so the
// capture holds stream operations only); anything else -- other sizes, profiling, a failed capture -- launches
// the kernels one by one.
constexpr int32_t kGraphMaxReads = 100000;

static int enqueue_batch(bdx_stream *s, Slot &sl)
{
    static const bool graphs_off = getenv("BDX_DISABLE_GRAPHS") != nullptr;
    const int32_t n = sl.n;
    bdx_pass_detail *det = s->details ? sl.d_det : nullptr;
    const bool eligible = !graphs_off && s->graphs_ok && !s->profile && n > 0 && n <= kGraphMaxReads;
    if (eligible && sl.graph && sl.graph_n == n && sl.graph_details == s->details) {
        CU(cudaGraphLaunch(sl.graph, s->st_comp));
        s->launches += sl.graph_launches;
        return BDX_OK;
    }
    if (eligible && sl.uses >= 1 && n == s->max_reads) {          // full-size chunks are the ones that repeat
        int rc = ensure_scratch(s, n);                            // allocation is not capturable
        if (rc) return rc;
        if (sl.graph) {
            cudaGraphExecDestroy(sl.graph);
            sl.graph = nullptr;
        }
        const int64_t l0 = s->launches;
        cudaGraph_t g = nullptr;
        bool ok = cudaStreamBeginCapture(s->st_comp, cudaStreamCaptureModeThreadLocal) == cudaSuccess;
        if (ok) {
            rc = enqueue_classify(s, sl.d_seq, sl.d_off, n, sl.d_res, det);
            const cudaError_t ce = cudaStreamEndCapture(s->st_comp, &g);
            ok = rc == BDX_OK && ce == cudaSuccess && g != nullptr;
        }
        if (ok) ok = cudaGraphInstantiate(&sl.graph, g, 0) == cudaSuccess;
        if (g) cudaGraphDestroy(g);
        if (ok) {
            sl.graph_n = n;
            sl.graph_details = s->details;
            sl.graph_launches = (int)(s->launches - l0);
            s->launches = l0;
            CU(cudaGraphLaunch(sl.graph, s->st_comp));
            s->launches += sl.graph_launches;
            return BDX_OK;
        }
        // capture failed: clear the error state and never try again on this stream
        cudaGetLastError();
        s->launches = l0;
        sl.graph = nullptr;
        s->graphs_ok = false;
    }
    sl.uses++;
    return enqueue_classify(s, sl.d_seq, sl.d_off, n, sl.d_res, det);
}

static int launch_slot(bdx_stream *s, Slot &sl, const uint8_t *h_seq, const int32_t *h_off)
{
    CU(cudaSetDevice(s->device));
    const int32_t n = sl.n;
    const size_t bytes = n ? (size_t)h_off[n] : 0;
    if (n) {
        CU(cudaMemcpyAsync(sl.d_off, h_off, ((size_t)n + 1) * 4, cudaMemcpyHostToDevice, s->st_copy));
        if (bytes) CU(cudaMemcpyAsync(sl.d_seq, h_seq, bytes, cudaMemcpyHostToDevice, s->st_copy));
    }
    CU(cudaEventRecord(sl.ev_h2d, s->st_copy));
    CU(cudaStreamWaitEvent(s->st_comp, sl.ev_h2d, 0));
    int rc = enqueue_batch(s, sl);
    if (rc) return rc;
    CU(cudaEventRecord(sl.ev_kern, s->st_comp));
    CU(cudaStreamWaitEvent(s->st_d2h, sl.ev_kern, 0));
    if (n) {
        CU(cudaMemcpyAsync(sl.h_res, sl.d_res, (size_t)n * sizeof(bdx_result), cudaMemcpyDeviceToHost, s->st_d2h));
        if (s->details)
            CU(cudaMemcpyAsync(sl.h_det, sl.d_det, (size_t)n * 2 * sizeof(bdx_pass_detail),
                               cudaMemcpyDeviceToHost, s->st_d2h));
    }
    CU(cudaEventRecord(sl.ev_done, s->st_d2h));
    sl.busy = true;
    s->head = (s->head + 1) % BDX_MAX_IN_FLIGHT;
    s->in_flight++;
    return BDX_OK;
}

// pinned input staging (only bdx_submit / bdx_acquire need it; bdx_submit_pinned does not)
static int ensure_host_staging(bdx_stream *s)
{
    if (s->host_staging) return BDX_OK;
    CU(cudaSetDevice(s->device));
    for (Slot &sl : s->slot) {
        CU(cudaHostAlloc(&sl.h_seq, (size_t)std::max<int64_t>(s->max_bytes, 16), cudaHostAllocDefault));
        CU(cudaHostAlloc(&sl.h_off, ((size_t)s->max_reads + 1) * 4, cudaHostAllocDefault));
    }
    s->host_staging = true;
    return BDX_OK;
}

static int check_batch(bdx_stream *s, const int32_t *offsets, int32_t n)
{
    if (n < 0) return fail(BDX_ERR_INVALID, "negative n_reads");
    if (n > s->max_reads) return fail(BDX_ERR_TOO_LARGE, "batch exceeds max_reads of the stream");
    if (n && offsets[0] != 0) return fail(BDX_ERR_INVALID, "offsets[0] must be 0");
    if (n && (offsets[n] < 0 || (int64_t)offsets[n] > s->max_bytes))
        return fail(BDX_ERR_TOO_LARGE, "batch exceeds max_bytes of the stream");
    return BDX_OK;
}

extern "C" int bdx_submit(bdx_stream *s, const uint8_t *seq, const int32_t *offsets, int32_t n, uint64_t tag)
{
    if (!s || (n > 0 && (!seq || !offsets))) return fail(BDX_ERR_INVALID, "null argument");
    if (s->max_reads <= 0) return fail(BDX_ERR_STATE, "stream has no staging (max_reads = 0)");
    if (s->acquired) return fail(BDX_ERR_STATE, "bdx_acquire pending; call bdx_commit");
    if (s->in_flight >= BDX_MAX_IN_FLIGHT) return fail(BDX_ERR_STATE, "BDX_MAX_IN_FLIGHT batches already in flight; call bdx_fetch");
    int rc = check_batch(s, offsets, n);
    if (rc) return rc;
    if ((rc = ensure_host_staging(s))) return rc;
    Slot &sl = s->slot[s->head];
    if (n) {
        memcpy(sl.h_off, offsets, ((size_t)n + 1) * 4);
        memcpy(sl.h_seq, seq, (size_t)offsets[n]);
    }
    sl.n = n;
    sl.tag = tag;
    return launch_slot(s, sl, sl.h_seq, sl.h_off);
}

extern "C" int bdx_submit_pinned(bdx_stream *s, const uint8_t *seq, const int32_t *offsets, int32_t n,
                                 uint64_t tag)
{
    if (!s || (n > 0 && (!seq || !offsets))) return fail(BDX_ERR_INVALID, "null argument");
    if (s->max_reads <= 0) return fail(BDX_ERR_STATE, "stream has no staging (max_reads = 0)");
    if (s->acquired) return fail(BDX_ERR_STATE, "bdx_acquire pending; call bdx_commit");
    if (s->in_flight >= BDX_MAX_IN_FLIGHT) return fail(BDX_ERR_STATE, "BDX_MAX_IN_FLIGHT batches already in flight; call bdx_fetch");
    int rc = check_batch(s, offsets, n);
    if (rc) return rc;
    Slot &sl = s->slot[s->head];
    sl.n = n;
    sl.tag = tag;
    return launch_slot(s, sl, seq, offsets);
}

extern "C" int bdx_acquire(bdx_stream *s, uint8_t **seq, int32_t **offsets)
{
    if (!s || !seq || !offsets) return fail(BDX_ERR_INVALID, "null argument");
    if (s->max_reads <= 0) return fail(BDX_ERR_STATE, "stream has no staging (max_reads = 0)");
    if (s->acquired) return fail(BDX_ERR_STATE, "already acquired");
    if (s->in_flight >= BDX_MAX_IN_FLIGHT) return fail(BDX_ERR_STATE, "BDX_MAX_IN_FLIGHT batches already in flight; call bdx_fetch");
    int rc = ensure_host_staging(s);
    if (rc) return rc;
    Slot &sl = s->slot[s->head];
    *seq = sl.h_seq;
    *offsets = sl.h_off;
    s->acquired = true;
    return BDX_OK;
}

extern "C" int bdx_commit(bdx_stream *s, int32_t n, uint64_t tag)
{
    if (!s) return fail(BDX_ERR_INVALID, "null stream");
    if (!s->acquired) return fail(BDX_ERR_STATE, "bdx_commit without bdx_acquire");
    Slot &sl = s->slot[s->head];
    int rc = check_batch(s, sl.h_off, n);
    if (rc) return rc;
    s->acquired = false;
    sl.n = n;
    sl.tag = tag;
    return launch_slot(s, sl, sl.h_seq, sl.h_off);
}

extern "C" int bdx_fetch(bdx_stream *s, uint64_t *tag, int32_t *n_reads, bdx_result *results,
                         bdx_pass_detail *details)
{
    if (!s) return fail(BDX_ERR_INVALID, "null stream");
    if (s->in_flight == 0) return fail(BDX_ERR_STATE, "nothing in flight");
    if (details && !s->details) return fail(BDX_ERR_STATE, "details not enabled on this stream");
    Slot &sl = s->slot[s->tail];
    CU(cudaEventSynchronize(sl.ev_done));
    if (tag) *tag = sl.tag;
    if (n_reads) *n_reads = sl.n;
    if (results && sl.n) memcpy(results, sl.h_res, (size_t)sl.n * sizeof(bdx_result));
    if (details && sl.n) memcpy(details, sl.h_det, (size_t)sl.n * 2 * sizeof(bdx_pass_detail));
    sl.busy = false;
    s->tail = (s->tail + 1) % BDX_MAX_IN_FLIGHT;
    s->in_flight--;
    return BDX_OK;
}

// zero-copy retrieval: pointers into the pinned result staging of the oldest batch,
// valid until the next bdx_submit / bdx_commit that reuses the slot
extern "C" int bdx_fetch_view(bdx_stream *s, uint64_t *tag, int32_t *n_reads, const bdx_result **results,
                              const bdx_pass_detail **details)
{
    if (!s) return fail(BDX_ERR_INVALID, "null stream");
    if (s->in_flight == 0) return fail(BDX_ERR_STATE, "nothing in flight");
    Slot &sl = s->slot[s->tail];
    CU(cudaEventSynchronize(sl.ev_done));
    if (tag) *tag = sl.tag;
    if (n_reads) *n_reads = sl.n;
    if (results) *results = sl.h_res;
    if (details) *details = s->details ? sl.h_det : nullptr;
    sl.busy = false;
    s->tail = (s->tail + 1) % BDX_MAX_IN_FLIGHT;
    s->in_flight--;
    return BDX_OK;
}

extern "C" int bdx_classify(bdx_stream *s, const uint8_t *seq, const int32_t *offsets, int32_t n,
                            bdx_result *results, bdx_pass_detail *details)
{
    if (s && s->in_flight) return fail(BDX_ERR_STATE, "batches in flight");
    int rc = bdx_submit(s, seq, offsets, n, 0);
    if (rc) return rc;
    return bdx_fetch(s, nullptr, nullptr, results, details);
}

extern "C" int bdx_classify_device(bdx_stream *s, const uint8_t *d_seq, const int32_t *d_off, int32_t n,
                                   bdx_result *d_res, bdx_pass_detail *d_det)
{
    if (!s || n < 0 || (n > 0 && (!d_seq || !d_off || !d_res))) return fail(BDX_ERR_INVALID, "bad argument");
    CU(cudaSetDevice(s->device));
    return enqueue_classify(s, d_seq, d_off, n, d_res, d_det);
}

extern "C" int bdx_stream_sync(bdx_stream *s)
{
    if (!s) return fail(BDX_ERR_INVALID, "null stream");
    CU(cudaSetDevice(s->device));
    CU(cudaStreamSynchronize(s->st_comp));
    return BDX_OK;
}

extern "C" int bdx_stream_profile(bdx_stream *s, int on)
{
    if (!s) return fail(BDX_ERR_INVALID, "null stream");
    s->profile = on != 0;
    return BDX_OK;
}

extern "C" int bdx_stream_profile_read_stages(bdx_stream *s, double ms[BDX_PROFILE_STAGES], int32_t n_launches[BDX_PROFILE_STAGES])
{
    if (!s || !ms || !n_launches) return fail(BDX_ERR_INVALID, "null argument");
    CU(cudaSetDevice(s->device));
    CU(cudaStreamSynchronize(s->st_comp));
    for (int k = 0; k < BDX_PROFILE_STAGES; k++) {
        ms[k] = 0.0;
        n_launches[k] = 0;
    }
    for (auto &pr : s->prof_events) {
        float t = 0.f;
        CU(cudaEventElapsedTime(&t, pr.e0, pr.e1));
        ms[pr.kind] += t;
        n_launches[pr.kind]++;
        cudaEventDestroy(pr.e0);
        cudaEventDestroy(pr.e1);
    }
    s->prof_events.clear();
    return BDX_OK;
}

extern "C" int bdx_stream_profile_read(bdx_stream *s, double *filter_ms, int32_t *n_launches)
{
    if (!s || !filter_ms || !n_launches) return fail(BDX_ERR_INVALID, "null argument");
    double ms[BDX_PROFILE_STAGES];
    int32_t nl[BDX_PROFILE_STAGES];
    const int rc = bdx_stream_profile_read_stages(s, ms, nl);
    if (rc) return rc;
    *filter_ms = ms[kStFilter];
    *n_launches = nl[kStFilter];
    return BDX_OK;
}

extern "C" int bdx_stream_path_counters(bdx_stream *s, int64_t *prefilter_reads, int64_t *seed_reads,
                                        int64_t *automaton_reads, int reset)
{
    if (!s) return fail(BDX_ERR_INVALID, "null stream");
    CU(cudaSetDevice(s->device));
    CU(cudaStreamSynchronize(s->st_comp));
    unsigned long long h[4];
    CU(cudaMemcpy(h, s->d_counters, sizeof(h), cudaMemcpyDeviceToHost));
    if (prefilter_reads) *prefilter_reads = (int64_t)h[0];
    if (seed_reads) *seed_reads = (int64_t)h[2];
    if (automaton_reads) *automaton_reads = (int64_t)h[1];
    if (reset) CU(cudaMemset(s->d_counters, 0, sizeof(h)));
    return BDX_OK;
}

extern "C" void *bdx_stream_cuda_stream(bdx_stream *s) { return s ? (void *)s->st_comp : nullptr; }
extern "C" int64_t bdx_stream_launch_count(const bdx_stream *s) { return s ? s->launches : 0; }

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
    if (!cfg || !devices || !out || n_devices <= 0 || streams_per_device <= 0) return fail(BDX_ERR_INVALID, "bad argument");
    *out = nullptr;
    bdx_pool *p = new (std::nothrow) bdx_pool();
    if (!p) return fail(BDX_ERR_NOMEM, "out of memory");
    // stream k of every device before stream k + 1 of any: consecutive batches land on different GPUs
    for (int k = 0; k < streams_per_device; k++)
        for (int d = 0; d < n_devices; d++) {
            bdx_stream *s = nullptr;
            const int rc = bdx_stream_create(cfg, devices[d], max_reads, max_bytes, &s);
            if (rc) {
                std::string keep = g_err;
                bdx_pool_destroy(p);
                g_err = keep;
                return rc;
            }
            p->streams.push_back(s);
        }
    *out = p;
    return BDX_OK;
}

template <typename Submit>
static int pool_submit(bdx_pool *p, Submit submit)
{
    if (!p) return fail(BDX_ERR_INVALID, "null pool");
    bdx_stream *s = p->streams[p->next];
    if (s->in_flight >= BDX_MAX_IN_FLIGHT) return fail(BDX_ERR_STATE, "the next stream of the pool is full; call bdx_pool_fetch");
    const int rc = submit(s);
    if (rc) return rc;
    p->order.push_back((int)p->next);
    p->next = (p->next + 1) % p->streams.size();
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
    if (!p) return fail(BDX_ERR_INVALID, "null pool");
    if (p->order.empty()) return fail(BDX_ERR_STATE, "nothing in flight");
    const int rc = bdx_fetch(p->streams[(size_t)p->order.front()], tag, n_reads, results, details);
    if (rc == BDX_OK) p->order.erase(p->order.begin());
    return rc;
}

extern "C" int bdx_pool_fetch_view(bdx_pool *p, uint64_t *tag, int32_t *n_reads, const bdx_result **results,
                                   const bdx_pass_detail **details)
{
    if (!p) return fail(BDX_ERR_INVALID, "null pool");
    if (p->order.empty()) return fail(BDX_ERR_STATE, "nothing in flight");
    const int rc = bdx_fetch_view(p->streams[(size_t)p->order.front()], tag, n_reads, results, details);
    if (rc == BDX_OK) p->order.erase(p->order.begin());
    return rc;
}

extern "C" int bdx_pool_in_flight(const bdx_pool *p) { return p ? (int)p->order.size() : 0; }

extern "C" int bdx_pool_stats_fetch(bdx_pool *p, int64_t *out, int64_t out_len)
{
    if (!p || !out) return fail(BDX_ERR_INVALID, "null argument");
    const int64_t L = p->streams[0]->cfg->lay.total_len;
    if (out_len < L) return fail(BDX_ERR_INVALID, "stats buffer too small");
    std::vector<int64_t> tmp((size_t)L);
    std::fill(out, out + L, 0);
    for (bdx_stream *s : p->streams) {
        const int rc = bdx_stats_fetch(s, tmp.data(), L);
        if (rc) return rc;
        for (int64_t k = 0; k < L; k++) out[k] += tmp[(size_t)k];
    }
    return BDX_OK;
}

// ---------------------------------------------------------------------------
// device FASTQ block demultiplexer (demux.cu)
// ---------------------------------------------------------------------------
extern "C" int bdx_demux_block(bdx_stream *s, const uint8_t *fq1, int64_t len1, const uint8_t *fq2, int64_t len2,
                               int final_block, int mode, bdx_demux_out *out)
{
    if (!s || !out) return fail(BDX_ERR_INVALID, "null argument");
    if ((mode & 3) > BDX_DEMUX_BOTH || (mode & ~(3 | BDX_DEMUX_DEVICE_IO))) return fail(BDX_ERR_INVALID, "unknown demux mode");
    if (s->in_flight) return fail(BDX_ERR_STATE, "batches in flight");
    CU(cudaSetDevice(s->device));
    if (!s->demux && !(s->demux = demux_state_create())) return fail(BDX_ERR_NOMEM, "out of memory");
    std::string err;
    const int rc = demux_run(
        s->demux, s->tab->P, s->st_comp,
        [s](const uint8_t *d_seq, const int *d_off, int n, bdx_result *d_res) {
            return enqueue_classify(s, d_seq, d_off, n, d_res, nullptr);
        },
        fq1, len1, fq2, len2, final_block, mode, &s->launches, out, err);
    if (rc && !err.empty()) g_err = err;
    return rc;
}

extern "C" int bdx_demux_stage_ms(const bdx_stream *s, float ms[8])
{
    if (!s || !ms) return fail(BDX_ERR_INVALID, "null argument");
    if (!s->demux) return fail(BDX_ERR_STATE, "no bdx_demux_block call yet");
    memcpy(ms, demux_stage_ms(s->demux), 8 * sizeof(float));
    return BDX_OK;
}

// ---------------------------------------------------------------------------
// stats
// ---------------------------------------------------------------------------
extern "C" int bdx_stats_layout_get(const bdx_config *cfg, bdx_stats_layout *out)
{
    if (!cfg || !out) return fail(BDX_ERR_INVALID, "null argument");
    *out = cfg->lay;
    return BDX_OK;
}

extern "C" int bdx_stats_fetch(bdx_stream *s, int64_t *out, int64_t out_len)
{
    if (!s || !out) return fail(BDX_ERR_INVALID, "null argument");
    if (!s->d_stats) return fail(BDX_ERR_STATE, "config was created without want_stats");
    if (out_len < s->cfg->lay.total_len) return fail(BDX_ERR_INVALID, "stats buffer too small");
    CU(cudaSetDevice(s->device));
    CU(cudaStreamSynchronize(s->st_comp));
    CU(cudaMemcpy(out, s->d_stats, (size_t)s->cfg->lay.total_len * 8, cudaMemcpyDeviceToHost));
    return BDX_OK;
}

// DemuxStats dictionaries from a (summed) counter buffer: what match_barcode_pass stores per matched pass
// (classification.jl:827-865) -- keys are alignment start, alignment length and round(score, digits=2).
extern "C" int64_t bdx_stats_entries(const bdx_config *cfg, const int64_t *counters, bdx_stats_entry *out, int64_t cap)
{
    if (!cfg || !counters) return fail(BDX_ERR_INVALID, "null argument");
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

extern "C" int bdx_stats_overflow_fetch(bdx_stream *s, bdx_stats_overflow *out, int64_t cap, int64_t *n, int64_t *lost)
{
    if (!s || !n || cap < 0 || (cap > 0 && !out)) return fail(BDX_ERR_INVALID, "bad argument");
    *n = 0;
    if (lost) *lost = 0;
    if (!s->d_stats) return fail(BDX_ERR_STATE, "config was created without want_stats");
    CU(cudaSetDevice(s->device));
    CU(cudaStreamSynchronize(s->st_comp));
    unsigned int h[2];
    CU(cudaMemcpy(h, s->d_n_ovf, sizeof(h), cudaMemcpyDeviceToHost));
    const int64_t kept = std::min<int64_t>(h[0], kStatsOvfCap);
    *n = kept;
    if (lost) *lost = h[1];
    const int64_t take = std::min(kept, cap);
    if (take > 0) CU(cudaMemcpy(out, s->d_ovf, (size_t)take * sizeof(bdx_stats_overflow), cudaMemcpyDeviceToHost));
    return BDX_OK;
}

extern "C" void *bdx_stats_device_ptr(bdx_stream *s) { return s ? (void *)s->d_stats : nullptr; }

extern "C" int bdx_stats_reset(bdx_stream *s)
{
    if (!s) return fail(BDX_ERR_INVALID, "null stream");
    if (!s->d_stats) return BDX_OK;
    CU(cudaSetDevice(s->device));
    CU(cudaMemsetAsync(s->d_stats, 0, (size_t)s->cfg->lay.total_len * 8, s->st_comp));
    CU(cudaMemsetAsync(s->d_n_ovf, 0, 2 * sizeof(unsigned int), s->st_comp));
    return BDX_OK;
}

// ---------------------------------------------------------------------------
// utilities
// ---------------------------------------------------------------------------
extern "C" int bdx_synth_reads_device(bdx_stream *s, const bdx_synth_spec *spec, int32_t n, uint8_t *d_seq,
                                      int32_t *d_off)
{
    if (!s || !spec || n < 0 || !d_seq || !d_off) return fail(BDX_ERR_INVALID, "bad argument");
    if (spec->read_len <= 0 || (int64_t)spec->read_len * n > 0x7FFFFFF0ll)
        return fail(BDX_ERR_TOO_LARGE, "n_reads * read_len must stay below 2^31");
    if (spec->start_lo < 1 || spec->start_hi < spec->start_lo || spec->end_hi < spec->end_lo)
        return fail(BDX_ERR_INVALID, "bad plant range");
    CU(cudaSetDevice(s->device));
    CU(launch_synth(s->tab->P, *spec, n, d_seq, d_off, s->st_comp));
    s->launches++;
    return BDX_OK;
}

extern "C" int bdx_int_alu_peak(int device, double *ops_per_second)
{
    if (!ops_per_second) return fail(BDX_ERR_INVALID, "null argument");
    cudaError_t e = run_int_alu_peak(device, ops_per_second);
    if (e != cudaSuccess) return cuda_fail(e, "int alu peak microbenchmark");
    return BDX_OK;
}

extern "C" void *bdx_host_alloc(size_t bytes)
{
    void *p = nullptr;
    if (cudaHostAlloc(&p, bytes, cudaHostAllocDefault) != cudaSuccess) {
        cudaGetLastError();
        return nullptr;
    }
    return p;
}
extern "C" void bdx_host_free(void *p)
{
    if (p) cudaFreeHost(p);
}
