// kernels.cu -- literal (reference-order) classification kernel, result
// finalisation + DemuxStats counters, synthetic read generator, integer-ALU peak
// microbenchmark.  sm_100a only.
#include <algorithm>
#include <map>
#include <mutex>
#include <tuple>
#include <cstdio>
#include <math_constants.h>

#include "bdx_internal.h"
#include "literal.cuh"

namespace bdx {

cudaError_t blocks_per_sm_cached(const void *kern, int threads, size_t smem, int *per_sm)
{
    static std::mutex mu;
    static std::map<std::tuple<int, const void *, int, size_t>, int> occupancy;
    static std::map<std::pair<int, const void *>, bool> limit_raised;
    int dev = 0;
    cudaError_t e = cudaGetDevice(&dev);
    if (e != cudaSuccess) return e;
    const auto key = std::make_tuple(dev, kern, threads, smem);
    bool raised;
    {
        std::lock_guard<std::mutex> lk(mu);
        auto it = occupancy.find(key);
        if (it != occupancy.end()) {
            *per_sm = it->second;
            return cudaSuccess;
        }
        raised = limit_raised.count({dev, kern}) != 0;
    }
    if (!raised) {
        // once per device and kernel, to the device maximum: configs with different table sizes share kernels
        int optin = 0;
        e = cudaDeviceGetAttribute(&optin, cudaDevAttrMaxSharedMemoryPerBlockOptin, dev);
        if (e != cudaSuccess) return e;
        e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, optin);
        if (e != cudaSuccess) return e;
    }
    int n = 0;
    e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&n, kern, threads, smem);
    if (e != cudaSuccess) return e;
    if (n < 1) n = 1;
    std::lock_guard<std::mutex> lk(mu);
    limit_raised[{dev, kern}] = true;
    occupancy[key] = n;
    *per_sm = n;
    return cudaSuccess;
}

// ---------------------------------------------------------------------------
// k_literal: one thread per read walks barcodes in file order under the running
// threshold -- the body of find_best_matching_bc (classification.jl:632-728)
// inside match_barcode_pass (:776-824).
//   from_filter = 0 : every barcode of the set is evaluated
//   from_filter = 1 : only reads the filter kernel left kBcPending; their
//                     candidate list (or the whole set on overflow) is evaluated
// ---------------------------------------------------------------------------
constexpr int kLiteralThreads = 128;

// SMEM_WS: the DP / origin columns live in shared memory as [row][thread] instead of in
// thread-local (i.e. off-chip, L1-cached) arrays.
// list / n_list (optional): the reads to evaluate, compacted by the kernels that queued them -- the lanes of a
// warp then all have work of the same kind (a warp costs as much as its most expensive lane).
template <int MAXM, bool SMEM_WS>
__device__ __forceinline__ void literal_one(const DevParams &P, const int pass, const int from_filter,
                                            const uint8_t *__restrict__ seq, const int *__restrict__ off, const int i,
                                            PassOut *__restrict__ out, const PassOut *__restrict__ prev_pass,
                                            const uint16_t *__restrict__ cand, const uint8_t *__restrict__ cand_cnt,
                                            int *ws_smem)
{
    if (from_filter) {
        if (out[i].bc != kBcPending) return;
    } else if (pass == 1) {
        // pass 2 runs only when pass 1 matched (classification.jl:879-888)
        if (prev_pass[i].bc <= 0) {
            PassOut o{kBcNotRun, 0, -1, -1};
            out[i] = o;
            return;
        }
    }
    const DevSet &S = P.set[pass];
    const int base = off[i];
    const int n = off[i + 1] - base;
    const uint8_t *r = seq + base;
    const Geometry g = pass_geometry(S, n);
    if (!g.valid) {
        PassOut o{kBcUnknown, 0, -1, -1};
        out[i] = o;
        return;
    }
    int dp_local[SMEM_WS ? 1 : MAXM + 2];
    int or_local[SMEM_WS ? 1 : MAXM + 2];
    const WsCol DP{SMEM_WS ? ws_smem + threadIdx.x : dp_local, SMEM_WS ? kLiteralThreads : 1};
    const WsCol OR{SMEM_WS ? ws_smem + (MAXM + 2) * kLiteralThreads + threadIdx.x : or_local,
                   SMEM_WS ? kLiteralThreads : 1};
    const Costs c{P.match, P.mismatch, P.indel, P.nindel, P.has_n};
    const bool with_delta = P.min_delta != 0.0;                   // :723
    const bool need_tb = S.trim_side != 0 || P.want_stats;         // :812
    BestState bs;
    best_init(bs, P.max_error_rate);

    int n_iter = S.n_bc;
    bool use_list = false;
    int win_first = 0, win_last = 0x7FFFFFFF;
    if (from_filter) {
        const int cnt = cand_cnt[i];
        if (cnt == kCandWindow) {            // k_seed's winner with the columns that hold all its best alignments
            n_iter = 1;
            use_list = true;
            win_first = cand[(size_t)i * kCandMax + 1];
            win_last = cand[(size_t)i * kCandMax + 2];
        } else if (cnt != kCandOverflow) {
            n_iter = cnt;
            use_list = true;
        }
    }
    for (int k = 0; k < n_iter; k++) {
        const int b = use_list ? (int)cand[(size_t)i * kCandMax + k] : k;
        const int qo = S.bc_off[b];
        const int m = S.bc_off[b + 1] - qo;
        const uint8_t *q = S.bc_bytes + qo;
        int s = -1, e = -1, dist;
        double score;
        if (P.algo == BDX_HAMMING) {
            const int allowed = allowed_from(bs.thr, m);           // :567
            dist = hamming_literal(q, m, r, n, allowed, g.start_j, g.end_j, g.max_start_pos,
                                   g.min_end_pos, S.trim_side, s, e);
            score = dist >= kInf ? CUDART_INF : __ddiv_rn((double)dist, (double)m);
        } else if (P.algo == BDX_EXACT) {
            dist = exact_literal(q, m, r, n, g.start_j, g.end_j, g.max_start_pos, g.min_end_pos,
                                 S.trim_side, s, e);
            score = dist >= kInf ? CUDART_INF : 0.0;
        } else {
            const int norm = S.norm[b];
            const int allowed = allowed_from(bs.thr, norm);        // :254
            if (need_tb)
                dist = sg_literal<true>(DP, OR, q - 1, r - 1, m, n, allowed, c, S.trim_side, g.start_j,
                                        g.end_j, g.max_start_pos, g.min_end_pos, s, e, win_first, win_last);
            else
                dist = sg_literal<false>(DP, OR, q - 1, r - 1, m, n, allowed, c, S.trim_side, g.start_j,
                                         g.end_j, g.max_start_pos, g.min_end_pos, s, e);
            score = dist >= kInf ? CUDART_INF : __ddiv_rn((double)dist, (double)norm);
        }
        best_consider(bs, with_delta, score, dist, b + 1, s, e);
    }
    out[i] = best_finish(bs, with_delta, P.min_delta);
}

template <int MAXM, bool SMEM_WS>
__global__ void __launch_bounds__(kLiteralThreads)
k_literal(const __grid_constant__ DevParams P, const int pass, const int from_filter,
          const uint8_t *__restrict__ seq, const int *__restrict__ off, const int n_reads,
          PassOut *__restrict__ out, const PassOut *__restrict__ prev_pass,
          const uint16_t *__restrict__ cand, const uint8_t *__restrict__ cand_cnt,
          const int *__restrict__ list, const int *__restrict__ n_list)
{
    extern __shared__ int ws_smem[];
    const int n_items = list ? *n_list : n_reads;
    for (int item = blockIdx.x * blockDim.x + threadIdx.x; item < n_items; item += gridDim.x * blockDim.x)
        literal_one<MAXM, SMEM_WS>(P, pass, from_filter, seq, off, list ? list[item] : item, out, prev_pass, cand,
                                   cand_cnt, ws_smem);
}

cudaError_t launch_literal(const DevParams &P, int pass, int from_filter, const uint8_t *seq,
                           const int *off, int n, const Scratch &sc, cudaStream_t st, const int *list,
                           const int *n_list)
{
    if (n <= 0) return cudaSuccess;
    const int threads = kLiteralThreads;
    // a list's length is known on the device only: a capped grid strides over it
    const int blocks = list ? std::max(1, std::min((n + threads - 1) / threads, 148 * 16)) : (n + threads - 1) / threads;
    const int max_m = P.set[pass].max_m;
    PassOut *out = sc.pass[pass];
    const PassOut *prev = sc.pass[0];
    if (max_m <= 32) {
        const size_t smem = (size_t)2 * (32 + 2) * threads * sizeof(int);
        k_literal<32, true><<<blocks, threads, smem, st>>>(P, pass, from_filter, seq, off, n, out, prev, sc.cand,
                                                           sc.cand_cnt, list, n_list);
    } else if (max_m <= 64) {
        const size_t smem = (size_t)2 * (64 + 2) * threads * sizeof(int);
        int unused = 0;     // raises the dynamic shared-memory limit once per device
        cudaError_t e = blocks_per_sm_cached((const void *)k_literal<64, true>, threads, smem, &unused);
        if (e != cudaSuccess) return e;
        k_literal<64, true><<<blocks, threads, smem, st>>>(P, pass, from_filter, seq, off, n, out, prev, sc.cand,
                                                           sc.cand_cnt, list, n_list);
    } else {
        k_literal<kMaxBarcodeLen, false><<<blocks, threads, 0, st>>>(P, pass, from_filter, seq, off, n, out, prev,
                                                                     sc.cand, sc.cand_cnt, list, n_list);
    }
    return cudaGetLastError();
}

// ---------------------------------------------------------------------------
// k_mark_pending: the reads a worklist still holds after the shortcut stages of a set without filter
// kernel are queued for k_literal with "scan every barcode".
// ---------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
k_mark_pending(const int pass, const int *__restrict__ wl, const int *__restrict__ n_wl, PassOut *__restrict__ out,
               const PassOut *__restrict__ prev_pass, uint8_t *__restrict__ cand_cnt)
{
    const int n = *n_wl;
    for (int k = blockIdx.x * blockDim.x + threadIdx.x; k < n; k += gridDim.x * blockDim.x) {
        const int read = wl[k];
        if (pass == 1 && prev_pass[read].bc <= 0) {                  // classification.jl:879-888
            out[read] = PassOut{kBcNotRun, 0, -1, -1};
        } else {
            cand_cnt[read] = (uint8_t)kCandOverflow;
            out[read] = PassOut{kBcPending, 0, -1, -1};
        }
    }
}

cudaError_t launch_mark_pending(const DevParams &P, int pass, int n, const Scratch &sc, const int *wl, const int *n_wl,
                                cudaStream_t st)
{
    (void)P;
    if (n <= 0) return cudaSuccess;
    const int blocks = std::max(1, std::min((n + 255) / 256, 2048));
    k_mark_pending<<<blocks, 256, 0, st>>>(pass, wl, n_wl, sc.pass[pass], sc.pass[0], sc.cand_cnt);
    return cudaGetLastError();
}

// ---------------------------------------------------------------------------
// k_finalize: determine_filename[_and_stats] (classification.jl:871-1005) minus
// the string building -- status, indices, keep range -- plus the DemuxStats
// counters of match_barcode_pass (:827-865) and :942-978.
// ---------------------------------------------------------------------------
__device__ __forceinline__ void stat_add(unsigned long long *p) { atomicAdd(p, 1ull); }

__device__ __forceinline__ int pass_norm(const DevParams &P, const DevSet &S, int bc)
{
    if (P.algo == BDX_SEMIGLOBAL) return S.norm[bc - 1];
    return S.bc_off[bc] - S.bc_off[bc - 1];
}

__device__ void stats_pass(const DevParams &P, const StatsDev &st, int pass, const PassOut &o)
{
    const bdx_stats_layout &L = st.lay;
    const int pos = o.start + L.pos_bias;
    const int len = o.end - o.start + 1;
    const int d = min(max(o.dist + L.dist_bias, 0), L.dist_bins - 1);
    unsigned long long *bp = st.buf + L.pos_off[pass], *bl = st.buf + L.len_off[pass],
                       *bd = st.buf + L.dist_off[pass];
    stat_add(bd + d);
    stat_add(bd + (size_t)o.bc * L.dist_bins + d);
    if (pos < 0 || pos >= L.pos_bins || len < 0 || len >= L.len_bins) {
        // outside the histograms (a long read searched near its end): keep the exact record instead
        const unsigned int k = atomicAdd(st.n_ovf, 1u);
        if (k < st.ovf_cap)
            st.ovf[k] = bdx_stats_overflow{pass + 1, o.bc, o.start, len};
        else
            atomicAdd(st.n_ovf + 1, 1u);
        return;
    }
    stat_add(bp + pos);
    stat_add(bl + len);
    stat_add(bp + (size_t)o.bc * L.pos_bins + pos);
    stat_add(bl + (size_t)o.bc * L.len_bins + len);
}

__global__ void __launch_bounds__(256)
k_finalize(const __grid_constant__ DevParams P, const int *__restrict__ off, const int n_reads,
           const PassOut *__restrict__ p1, const PassOut *__restrict__ p2,
           bdx_result *__restrict__ res, bdx_pass_detail *__restrict__ det, const StatsDev st)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n_reads) return;
    const int n = off[i + 1] - off[i];
    const PassOut a = p1[i];
    PassOut b{kBcNotRun, 0, -1, -1};
    if (P.is_dual && a.bc > 0) b = p2[i];

    bdx_result r;
    r.bc1 = 0;
    r.bc2 = 0;
    r.keep_start = -1;
    r.keep_end = -1;
    if (a.bc <= 0) {
        r.status = (a.bc == kBcAmbiguous) ? BDX_AMBIGUOUS : BDX_UNKNOWN;         // :879-883
    } else if (P.is_dual && b.bc <= 0) {
        r.status = (b.bc == kBcAmbiguous) ? BDX_AMBIGUOUS : BDX_UNKNOWN;         // :890-894
    } else {
        r.status = BDX_MATCH;
        r.bc1 = a.bc;
        r.bc2 = P.is_dual ? b.bc : 0;
        int keep_start = 1, keep_end = n;                                        // :907-908
        const int t1 = P.set[0].trim_side, t2 = P.set[1].trim_side;
        if (t1 == 3) keep_end = max(1, a.start) - 1;                             // :914
        else if (t1 == 5) keep_start = a.end + 1;                                // :917
        if (P.is_dual && t2 != 0) {                                              // :921-929
            if (t2 == 3) keep_end = min(keep_end, max(1, b.start) - 1);
            else if (t2 == 5) keep_start = max(keep_start, b.end + 1);
        }
        if (keep_start > keep_end) { keep_start = 1; keep_end = 0; }             // :932-935
        r.keep_start = keep_start;
        r.keep_end = keep_end;
    }
    res[i] = r;

    if (det) {
        bdx_pass_detail d1;
        d1.status = a.bc > 0 ? BDX_MATCH : (a.bc == kBcAmbiguous ? BDX_AMBIGUOUS : BDX_UNKNOWN);
        d1.bc = a.bc > 0 ? a.bc : 0;
        d1.dist = a.bc > 0 ? a.dist : -1;
        d1.norm = a.bc > 0 ? pass_norm(P, P.set[0], a.bc) : 0;
        d1.start = a.bc > 0 ? a.start : -1;
        d1.end = a.bc > 0 ? a.end : -1;
        det[i] = d1;
        bdx_pass_detail d2;
        const bool ran2 = P.is_dual && a.bc > 0;
        d2.status = !ran2 ? -1 : (b.bc > 0 ? BDX_MATCH : (b.bc == kBcAmbiguous ? BDX_AMBIGUOUS : BDX_UNKNOWN));
        d2.bc = (ran2 && b.bc > 0) ? b.bc : 0;
        d2.dist = (ran2 && b.bc > 0) ? b.dist : -1;
        d2.norm = (ran2 && b.bc > 0) ? pass_norm(P, P.set[1], b.bc) : 0;
        d2.start = (ran2 && b.bc > 0) ? b.start : -1;
        d2.end = (ran2 && b.bc > 0) ? b.end : -1;
        det[(size_t)n_reads + i] = d2;
    }

    if (P.want_stats && st.buf) {
        stat_add(st.buf + 0);                                                    // :942 total_reads
        if (a.bc > 0) stats_pass(P, st, 0, a);                                   // :827-845 (even if pass 2 fails)
        if (P.is_dual && a.bc > 0 && b.bc > 0) stats_pass(P, st, 1, b);          // :846-864
        if (r.status == BDX_MATCH) {
            stat_add(st.buf + 1);                                                // :976
            stat_add(st.buf + st.lay.sample_off + (size_t)r.bc1 * (st.lay.b2 + 1) + r.bc2);  // :977-978
        } else if (r.status == BDX_UNKNOWN) {
            stat_add(st.buf + 2);                                                // :950, :963
        } else {
            stat_add(st.buf + 3);                                                // :953, :966
        }
    }
}

cudaError_t launch_finalize(const DevParams &P, const int *off, int n, const Scratch &sc, bdx_result *res,
                            bdx_pass_detail *det, StatsDev stats, cudaStream_t st)
{
    if (n <= 0) return cudaSuccess;
    k_finalize<<<(n + 255) / 256, 256, 0, st>>>(P, off, n, sc.pass[0], sc.pass[1], res, det, stats);
    return cudaGetLastError();
}

// ---------------------------------------------------------------------------
// Synthetic reads.  Philox4x32-10 keyed by the seed, counter =
// (read index, draw index): a read's bytes depend only on (seed, global index).
// ---------------------------------------------------------------------------
struct Philox {
    uint32_t key0, key1, c0, c1, ctr;
    uint32_t buf[4];
    int have;
    __device__ Philox(uint64_t seed, uint64_t idx)
        : key0((uint32_t)seed), key1((uint32_t)(seed >> 32)), c0((uint32_t)idx), c1((uint32_t)(idx >> 32)),
          ctr(0), have(0) {}
    __device__ void refill()
    {
        uint32_t x0 = c0, x1 = c1, x2 = ctr++, x3 = 0x42444D58u;  // "BDMX"
        uint32_t k0 = key0, k1 = key1;
#pragma unroll
        for (int round = 0; round < 10; round++) {
            const uint32_t hi0 = __umulhi(0xD2511F53u, x0), lo0 = 0xD2511F53u * x0;
            const uint32_t hi1 = __umulhi(0xCD9E8D57u, x2), lo1 = 0xCD9E8D57u * x2;
            const uint32_t y0 = hi1 ^ x1 ^ k0, y1 = lo1, y2 = hi0 ^ x3 ^ k1, y3 = lo0;
            x0 = y0; x1 = y1; x2 = y2; x3 = y3;
            k0 += 0x9E3779B9u;
            k1 += 0xBB67AE85u;
        }
        buf[0] = x0; buf[1] = x1; buf[2] = x2; buf[3] = x3;
        have = 4;
    }
    __device__ uint32_t next()
    {
        if (have == 0) refill();
        return buf[--have];
    }
    __device__ uint32_t below(uint32_t n) { return (uint32_t)(((uint64_t)next() * n) >> 32); }
};

__device__ int synth_edit_count(Philox &g)
{
    // P(k) = {0:.50, 1:.25, 2:.13, 3:.07, 4:.03, 5:.02}
    const uint32_t u = g.below(100);
    if (u < 50) return 0;
    if (u < 75) return 1;
    if (u < 88) return 2;
    if (u < 95) return 3;
    if (u < 98) return 4;
    return 5;
}

__device__ int synth_mutate(Philox &g, const DevSet &S, uint8_t *buf)
{
    const char base[4] = {'A', 'C', 'G', 'T'};
    const int b = (int)g.below((uint32_t)S.n_bc);
    int len = S.bc_off[b + 1] - S.bc_off[b];
    for (int k = 0; k < len; k++) buf[k] = S.bc_bytes[S.bc_off[b] + k];
    const int edits = synth_edit_count(g);
    for (int e = 0; e < edits; e++) {
        const uint32_t kind = g.below(3);
        if (kind == 0) {  // substitution by a different base
            const int p = (int)g.below((uint32_t)len);
            uint8_t nb = (uint8_t)base[g.below(4)];
            if (nb == buf[p]) nb = (uint8_t)base[(g.below(3) + 1 + (nb == 'A' ? 0 : nb == 'C' ? 1 : nb == 'G' ? 2 : 3)) & 3];
            buf[p] = nb;
        } else if (kind == 1) {  // insertion
            if (len < kMaxBarcodeLen + 7) {
                const int p = (int)g.below((uint32_t)len + 1);
                for (int k = len; k > p; k--) buf[k] = buf[k - 1];
                buf[p] = (uint8_t)base[g.below(4)];
                len++;
            }
        } else if (len > 1) {  // deletion
            const int p = (int)g.below((uint32_t)len);
            for (int k = p; k + 1 < len; k++) buf[k] = buf[k + 1];
            len--;
        }
    }
    return len;
}

__global__ void __launch_bounds__(128)
k_synth(const __grid_constant__ DevParams P, const bdx_synth_spec spec, const int n_reads,
        uint8_t *__restrict__ seq, int *__restrict__ off)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i > n_reads) return;
    const int L = spec.read_len;
    off[i] = i * L;  // includes the terminating offset at i == n_reads
    if (i == n_reads) return;
    Philox g(spec.seed, (uint64_t)spec.first_read + (uint64_t)i);
    uint8_t *r = seq + (size_t)i * L;
    const char base[4] = {'A', 'C', 'G', 'T'};
    for (int j = 0; j < L; j += 16) {
        uint32_t w = g.next();
        for (int k = 0; k < 16 && j + k < L; k++, w >>= 2) r[j + k] = (uint8_t)base[w & 3];
    }
    uint8_t buf[kMaxBarcodeLen + 8];
    if ((int)g.below(1000) < spec.plant_permille) {
        const int len = synth_mutate(g, P.set[0], buf);
        const int start = spec.start_lo + (int)g.below((uint32_t)(spec.start_hi - spec.start_lo + 1));
        for (int k = 0; k < len && start - 1 + k < L; k++)
            if (start - 1 + k >= 0) r[start - 1 + k] = buf[k];
        if (spec.set2_mode != 0 && P.is_dual) {
            const int len2 = synth_mutate(g, P.set[1], buf);
            const int v = spec.end_lo + (int)g.below((uint32_t)(spec.end_hi - spec.end_lo + 1));
            // mode 1: the barcode ends `v` bases before the read end; mode 2: it starts at 1-based position v
            const int s2 = spec.set2_mode == 1 ? L - v - len2 : v - 1;
            for (int k = 0; k < len2; k++)
                if (s2 + k >= 0 && s2 + k < L) r[s2 + k] = buf[k];
        }
    }
    if ((int)g.below(10000) < spec.n_permille_x10) r[g.below((uint32_t)L)] = (uint8_t)'N';
}

cudaError_t launch_synth(const DevParams &P, const bdx_synth_spec &spec, int n, uint8_t *seq, int *off,
                         cudaStream_t st)
{
    k_synth<<<(n + 1 + 127) / 128, 128, 0, st>>>(P, spec, n, seq, off);
    return cudaGetLastError();
}

// ---------------------------------------------------------------------------
// Integer-ALU peak: 8 independent LOP3 / IADD3 dependency chains per thread, the
// instruction mix of the bit-parallel column step, at full occupancy.
// ---------------------------------------------------------------------------
constexpr int kPeakIters = 2048;
constexpr int kPeakChains = 8;
constexpr int kPeakOpsPerIter = 4;  // per chain: lop3, add, lop3, add

__global__ void __launch_bounds__(1024)
k_int_peak(uint32_t *out, uint32_t seed)
{
    uint32_t v[kPeakChains], a = seed ^ threadIdx.x, b = seed * 2654435761u + blockIdx.x;
#pragma unroll
    for (int c = 0; c < kPeakChains; c++) v[c] = a + c * 0x9E3779B9u;
    for (int it = 0; it < kPeakIters; it++) {
#pragma unroll
        for (int c = 0; c < kPeakChains; c++) {
            asm volatile("lop3.b32 %0, %0, %1, %2, 0x96;" : "+r"(v[c]) : "r"(a), "r"(b));
            asm volatile("add.u32 %0, %0, %1;" : "+r"(v[c]) : "r"(b));
            asm volatile("lop3.b32 %0, %0, %1, %2, 0xE8;" : "+r"(v[c]) : "r"(b), "r"(a));
            asm volatile("add.u32 %0, %0, %1;" : "+r"(v[c]) : "r"(a));
        }
    }
    uint32_t acc = 0;
#pragma unroll
    for (int c = 0; c < kPeakChains; c++) acc ^= v[c];
    if (acc == 0x12345678u) out[0] = acc;  // keep the chains alive
}

cudaError_t run_int_alu_peak(int device, double *ops_per_second)
{
    cudaError_t e = cudaSetDevice(device);
    if (e != cudaSuccess) return e;
    cudaDeviceProp prop;
    if ((e = cudaGetDeviceProperties(&prop, device)) != cudaSuccess) return e;
    uint32_t *d = nullptr;
    if ((e = cudaMalloc(&d, 4)) != cudaSuccess) return e;
    const int blocks = prop.multiProcessorCount * 2, threads = 1024;
    cudaEvent_t t0, t1;
    cudaEventCreate(&t0);
    cudaEventCreate(&t1);
    double best = 0.0;
    for (int rep = 0; rep < 6; rep++) {
        cudaEventRecord(t0);
        k_int_peak<<<blocks, threads>>>(d, 12345u + rep);
        cudaEventRecord(t1);
        if ((e = cudaEventSynchronize(t1)) != cudaSuccess) break;
        float ms = 0.f;
        cudaEventElapsedTime(&ms, t0, t1);
        const double ops = (double)blocks * threads * kPeakIters * kPeakChains * kPeakOpsPerIter;
        if (rep >= 2 && ms > 0.f) best = fmax(best, ops / (ms * 1e-3));
    }
    cudaEventDestroy(t0);
    cudaEventDestroy(t1);
    cudaFree(d);
    if (e == cudaSuccess) e = cudaGetLastError();
    *ops_per_second = best;
    return e;
}

}  // namespace bdx
