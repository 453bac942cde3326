// seed.cu -- depth-limited seed-and-verify for :semiglobal in the exact regime.
//
// Setting: unit costs, score-only, no start/end constraint, min_delta = 0, a barcode set of one
// common length m without wildcard rows (the regime k_prefilter<0> works in).  k_prefilter has
// already resolved the reads with a verbatim barcode occurrence; this kernel resolves reads whose
// best barcode is within K = sd_k edits, without running the full-range automaton:
//
//   * Pigeonhole: cut every barcode into K + 1 disjoint segments.  An alignment with <= K edits
//     leaves at least one segment untouched, so its first q bytes occur verbatim in the read at
//     column p = s + o + shift, |shift| <= K (s = alignment start, o = segment offset).
//   * Every read column's q-mer is hashed (rolling polynomial hash over class codes) into a
//     first-level bitmap and a CSR bucket table of (barcode, o) entries; a hit fixes the
//     diagonal delta = p - o.
//   * Each hit is verified with the same Myers/Hyyro automaton as k_filter, but only over the
//     window of columns [delta - K, delta + m + 2K] that can hold such an alignment.  A window is
//     a sub-range of the search range with free start and end, so its minimum is >= the
//     full-range distance d_b, and for d_b <= K some hit window contains an optimal alignment:
//     min over the hit windows == d_b exactly whenever d_b <= K.
//   * Hence the set {b : d_b <= K} and those distances are known exactly.  If it is non-empty the
//     reference's answer (no min_delta: first barcode attaining the minimal score,
//     classification.jl:658) is the lowest index among the minimal d_b.  If it is empty and
//     K == allowed, no barcode is acceptable.  Otherwise (best distance in (K, allowed]) the read
//     goes to the bit-parallel kernel via worklist2, like any read this kernel cannot stage.
//
// One thread per read; hits are first collected, then verified in lock step (verifying inside
// the scan would serialise the lanes of a warp, see k_prefilter).
#include <algorithm>
#include <cstdio>
#include <cstdlib>
#include <math_constants.h>

#include "bdx_internal.h"
#include "literal.cuh"

namespace bdx {

constexpr int kSeedThreads = 128;
constexpr int kSeedSlot = 176;      // staged class codes per read (longer reads take the full path)
constexpr int kSeedMaxHits = 28;    // hits remembered per read (more => full path)
constexpr int kSeedMaxWins = 32;    // bitmap-passing columns remembered per read (more => full path)
constexpr int kSeedIlp = 4;         // hits verified concurrently per thread

__global__ void __launch_bounds__(kSeedThreads)
k_seed(const __grid_constant__ DevParams P, const int pass, const uint8_t *__restrict__ seq,
       const int *__restrict__ off, PassOut *__restrict__ out, const int *__restrict__ worklist,
       const int *__restrict__ n_work, int *__restrict__ worklist2, int *__restrict__ n_work2,
       unsigned long long *__restrict__ counters)
{
    extern __shared__ __align__(16) uint32_t smem[];
    const DevSet &S = P.set[pass];
    const int n_pad = S.n_bc_pad;
    const int n_buckets = 1 << S.sd_log2;
    const int bm_words = 1 << (S.sd_bm_log2 - 5);
    uint32_t *peq_s = smem;                                     // [n_classes][n_pad]
    uint32_t *bitmap_s = peq_s + S.n_classes * n_pad;
    uint32_t *bstart_s = bitmap_s + bm_words;                   // [n_buckets + 1]
    uint32_t *entries_s = bstart_s + n_buckets + 1;             // [sd_n_entries]
    uint32_t *ekeys_s = entries_s + S.sd_n_entries;             // [sd_n_entries] full hash of each entry
    uint32_t *hits_s = ekeys_s + S.sd_n_entries;                // [kSeedMaxHits][kSeedThreads]
    uint8_t *wins_s = reinterpret_cast<uint8_t *>(hits_s + kSeedMaxHits * kSeedThreads);   // [kSeedMaxWins][threads]
    uint8_t *class_s = wins_s + kSeedMaxWins * kSeedThreads;
    uint8_t *slot_s = class_s + 256;                            // [kSeedThreads][kSeedSlot] class codes

    for (int k = threadIdx.x; k < S.n_classes * n_pad; k += blockDim.x) peq_s[k] = S.peq[k];
    for (int k = threadIdx.x; k < bm_words; k += blockDim.x) bitmap_s[k] = S.sd_bitmap[k];
    for (int k = threadIdx.x; k <= n_buckets; k += blockDim.x) bstart_s[k] = S.sd_bstart[k];
    for (int k = threadIdx.x; k < S.sd_n_entries; k += blockDim.x) {
        entries_s[k] = S.sd_entries[k];
        ekeys_s[k] = S.sd_ekeys[k];
    }
    for (int k = threadIdx.x; k < 256; k += blockDim.x) class_s[k] = S.class_of[k];
    __syncthreads();

    const int lane = threadIdx.x & 31;
    const int warp = threadIdx.x >> 5;
    const int m = S.sd_m, K = S.sd_k, q = S.sd_q;
    const uint32_t pw = S.sd_pow;
    const int bm_log2 = S.sd_bm_log2;
    const int n_items = *n_work;
    const int n_groups = (n_items + kSeedThreads - 1) / kSeedThreads;
    const uint32_t row_mask = m >= 32 ? 0xFFFFFFFFu : (0xFFFFFFFFu << (32 - m));
    unsigned int n_done = 0;

    for (int grp = blockIdx.x; grp < n_groups; grp += gridDim.x) {
        const int item = grp * kSeedThreads + threadIdx.x;
        const bool have = item < n_items;
        const int read = have ? worklist[item] : 0;
        const int base = have ? off[read] : 0;
        const int n = have ? off[read + 1] - base : 0;

        // ---- stage the warp's 32 reads as class codes, one read at a time, coalesced ----
        uint8_t *my_slot = slot_s + (size_t)threadIdx.x * kSeedSlot;
        __syncwarp();
        for (int r = 0; r < 32; r++) {
            const int rb = __shfl_sync(0xFFFFFFFFu, base, r);
            const int rn = __shfl_sync(0xFFFFFFFFu, n, r);
            if (rn > kSeedSlot) continue;
            uint8_t *dst = slot_s + (size_t)(warp * 32 + r) * kSeedSlot;
            const uint8_t *src = seq + rb;
            for (int t = lane; t < rn; t += 32) dst[t] = class_s[src[t]];
        }
        __syncwarp();

        // ---- scan, phase 1: remember the columns whose q-mer passes the first-level bitmap.
        // (Doing the bucket walk right here would serialise the lanes of a warp: every lane hits
        // at different columns.  The divergent part is kept to one shared-memory store.) ----
        int n_wins = 0;
        bool punt = !have;       // true => this read goes to worklist2 (or is not a read at all)
        Geometry g{};
        if (have) {
            g = pass_geometry(S, n);
            // the regime test of k_filter's `fast` / k_prefilter<0>
            if (!(g.valid && g.max_start_pos >= n && g.min_end_pos <= g.start_j) || n > kSeedSlot) punt = true;
        }
        if (!punt) {
            const int p0 = g.start_j - 1;              // 0-based column of the first q-mer
            const int p1 = g.end_j - q;                // last one that lies inside the search range
            if (p1 >= p0) {
                uint32_t h = 0;
                for (int i = 0; i < q; i++) h = h * kPfBase + (uint32_t)my_slot[p0 + i];
                for (int p = p0;;) {
                    const uint32_t bit = pf_bit(h, bm_log2);
                    if ((bitmap_s[bit >> 5] >> (bit & 31)) & 1u) {
                        if (n_wins < kSeedMaxWins) wins_s[n_wins * kSeedThreads + threadIdx.x] = (uint8_t)p;
                        n_wins++;
                    }
                    if (++p > p1) break;
                    h = (h - (uint32_t)my_slot[p - 1] * pw) * kPfBase + (uint32_t)my_slot[p - 1 + q];
                }
            }
            if (n_wins > kSeedMaxWins) punt = true;
        }

        // ---- scan, phase 2 (lock step over the remembered columns): bucket walk, key check,
        // (barcode, diagonal) hits ----
        int n_hits = 0;
        {
            const int my_wins = punt ? 0 : n_wins;
            const int max_wins = __reduce_max_sync(0xFFFFFFFFu, my_wins);
            for (int k = 0; k < max_wins; k++) {
                if (k >= my_wins) continue;
                const int p = wins_s[k * kSeedThreads + threadIdx.x];
                uint32_t h = 0;
                for (int i = 0; i < q; i++) h = h * kPfBase + (uint32_t)my_slot[p + i];
                const uint32_t bucket = pf_slot(h, S.sd_log2);
                const uint32_t e1 = bstart_s[bucket + 1];
                for (uint32_t e = bstart_s[bucket]; e < e1; e++) {
                    if (ekeys_s[e] != h) continue;                      // bucket-mate with another q-mer
                    const uint32_t ent = entries_s[e];
                    // 32-bit hash collisions are harmless: a false hit only costs a verification
                    const int delta = p - (int)(ent & 0xFFu);           // 0-based diagonal
                    const uint32_t rec = ((ent >> 8) << 16) | (uint32_t)(delta + 256);
                    // (several intact segments of one alignment give the same record several times;
                    // searching the list for duplicates costs more than verifying them again)
                    if (n_hits < kSeedMaxHits) hits_s[n_hits * kSeedThreads + threadIdx.x] = rec;
                    n_hits++;
                }
            }
            if (n_hits > kSeedMaxHits) punt = true;
        }

        // ---- verify the hits in lock step: windowed bit-parallel automaton, kSeedIlp hits at a
        // time per thread (independent dependency chains hide the latency of the serial column
        // recurrence at the low occupancy the shared-memory tables allow) ----
        int best_d = kInf, best_b = 0x7FFFFFFF;
        const int my_hits = punt ? 0 : n_hits;
        const int max_hits = __reduce_max_sync(0xFFFFFFFFu, my_hits);
        const int win = m + 3 * K + 1;                       // columns [delta + 1 - K, delta + m + 2K]
        for (int k0 = 0; k0 < max_hits; k0 += kSeedIlp) {
            int hb[kSeedIlp], c0[kSeedIlp], c1[kSeedIlp], score[kSeedIlp], best[kSeedIlp];
            uint32_t pv[kSeedIlp], mv[kSeedIlp];
#pragma unroll
            for (int u = 0; u < kSeedIlp; u++) {
                const bool live = k0 + u < my_hits;
                const uint32_t rec = live ? hits_s[(k0 + u) * kSeedThreads + threadIdx.x] : 0u;
                hb[u] = (int)(rec >> 16);
                const int delta = (int)(rec & 0xFFFFu) - 256;
                // 1-based columns an alignment with <= K edits on this diagonal can occupy
                c0[u] = live ? max(g.start_j, delta + 1 - K) : 1;
                c1[u] = live ? min(g.end_j, delta + m + 2 * K) : 0;
                pv[u] = row_mask;
                mv[u] = 0u;
                score[u] = m;
                best[u] = kInf;
            }
            for (int t = 0; t < win; t++) {
#pragma unroll
                for (int u = 0; u < kSeedIlp; u++) {
                    const int c = c0[u] + t;
                    if (c <= c1[u]) {
                        const uint32_t eq = peq_s[(int)my_slot[c - 1] * n_pad + hb[u]];
                        const uint32_t xv = eq | mv[u];
                        const uint32_t xh = ((((eq & pv[u]) + pv[u]) ^ pv[u]) | eq);
                        const uint32_t ph = mv[u] | ~(xh | pv[u]);
                        const uint32_t mh = pv[u] & xh;
                        score[u] += (int)(ph >> 31) - (int)(mh >> 31);
                        const uint32_t phs = ph << 1, mhs = mh << 1;
                        pv[u] = mhs | ~(xv | phs);
                        mv[u] = phs & xv;
                        best[u] = min(best[u], score[u]);
                    }
                }
            }
#pragma unroll
            for (int u = 0; u < kSeedIlp; u++)
                if (best[u] <= K && (best[u] < best_d || (best[u] == best_d && hb[u] < best_b))) {
                    best_d = best[u];
                    best_b = hb[u];
                }
        }

#ifdef BDX_SEED_DEBUG
        if (item < 3) printf("item %d read %d n %d punt %d wins %d hits %d best_d %d best_b %d\n", item, read, n, (int)punt,
                             n_wins, n_hits, best_d, best_b);
#endif
        // ---- decide ----
        bool resolved = false;
        if (!punt) {
            if (best_d <= K) {
                // acceptance exactly as find_best_matching_bc_no_delta does it for this barcode
                // (classification.jl:254, :658); with one common length the earlier, worse
                // barcodes the reference may accept first cannot change the final winner
                const int norm = S.norm[best_b];
                const double sc = __ddiv_rn((double)best_d, (double)norm);
                if (best_d <= allowed_from(P.max_error_rate, norm) && sc <= P.max_error_rate) {
                    out[read] = PassOut{best_b + 1, best_d, -1, -1};
                    resolved = true;
                }
            } else if (K >= S.allowed0[0]) {
                out[read] = PassOut{kBcUnknown, 0, -1, -1};    // nothing within the allowed distance
                resolved = true;
            }
        }
        const bool todo = have && !resolved;
        const uint32_t mask = __ballot_sync(0xFFFFFFFFu, todo);
        int base_slot = 0;
        if (lane == 0 && mask) base_slot = atomicAdd(n_work2, __popc(mask));
        base_slot = __shfl_sync(0xFFFFFFFFu, base_slot, 0);
        if (todo) worklist2[base_slot + __popc(mask & ((1u << lane) - 1u))] = read;
        n_done += __popc(__ballot_sync(0xFFFFFFFFu, resolved));
    }
    if (lane == 0 && n_done && counters) atomicAdd(counters + 2, (unsigned long long)n_done);
}

static size_t seed_smem(const DevSet &S)
{
    size_t words = (size_t)S.n_classes * S.n_bc_pad + ((size_t)1 << (S.sd_bm_log2 - 5)) + ((size_t)1 << S.sd_log2) + 1 +
                   2 * (size_t)S.sd_n_entries + (size_t)kSeedMaxHits * kSeedThreads;
    return words * 4 + (size_t)kSeedMaxWins * kSeedThreads + 256 + (size_t)kSeedThreads * kSeedSlot + 16;
}

bool seed_applies(const DevParams &P, int pass)
{
    static const bool off = getenv("BDX_DISABLE_SEED") != nullptr;
    const DevSet &S = P.set[pass];
    return !off && prefilter_applies(P, pass) && P.algo == BDX_SEMIGLOBAL && S.sd_enabled && S.words == 1 &&
           seed_smem(S) <= 100 * 1024;
}

cudaError_t launch_seed(const DevParams &P, int pass, const uint8_t *seq, const int *off, int n, const Scratch &sc,
                        int sm_count, unsigned long long *counters, cudaStream_t st)
{
    const DevSet &S = P.set[pass];
    const size_t smem = seed_smem(S);
    cudaError_t e = cudaFuncSetAttribute(k_seed, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return e;
    int per_sm = 0;
    e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, k_seed, kSeedThreads, smem);
    if (e != cudaSuccess) return e;
    if (per_sm < 1) per_sm = 1;
    const int groups = (n + kSeedThreads - 1) / kSeedThreads;     // upper bound: the worklist is <= n
    const int blocks = std::max(1, std::min(groups, sm_count * per_sm));
    e = cudaMemsetAsync(sc.n_work2, 0, sizeof(int), st);
    if (e != cudaSuccess) return e;
    k_seed<<<blocks, kSeedThreads, smem, st>>>(P, pass, seq, off, sc.pass[pass], sc.worklist, sc.n_work,
                                               sc.worklist2, sc.n_work2, counters);
    return cudaGetLastError();
}

}  // namespace bdx
