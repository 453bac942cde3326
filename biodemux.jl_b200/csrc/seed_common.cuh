// seed_common.cuh -- pieces shared by the seed-and-verify kernels (seed.cu, seed_deep.cu): hit records,
// and the pooled, windowed bit-parallel verification of a warp's hits.
#pragma once

#include <type_traits>

#include "bdx_internal.h"
#include "literal.cuh"

namespace bdx {

constexpr int kSeedThreads = 128;
constexpr int kSeedSlot = 180;      // slot bytes per read: its search range + up to 3 bytes of skew (longer ranges take the
                                    // full path); 45 words:
                                    // an odd word stride keeps the lock-step scan free of bank conflicts
constexpr int kSeedMaxHits = 28;    // distinct (barcode, diagonal group) hits remembered per read (more => next stage)
constexpr int kSeedMaxWins = 32;    // bitmap-passing columns remembered per read (more => full path)

// Stages, for each of the warp's 32 reads, `my_len` bytes starting at seq + my_start as class codes into that read's
// slot (lane l describes read l).  Four reads per round; a lane loads up to two ALIGNED 32-bit words per read -- a
// range spans at most 46 words including its misaligned head -- and all eight loads of a round are issued before the
// first table lookup.  The slot keeps the SOURCE alignment: word w of the slot = the class codes of aligned source
// word w, one 32-bit store per word, so the read's first byte sits at slot[mis] with mis = (address of its first
// byte) & 3 (seed_slot_skew), and a range needs mis + len <= kSeedSlot.  The bytes in front of and behind the range
// inside its first / last word belong to neighbouring reads and are never looked at.  The aligned words never leave
// the allocation that holds the bytes (allocations start and end on coarser boundaries).
__device__ __forceinline__ int seed_slot_skew(const uint8_t *seq, long long start)
{
    return (int)(reinterpret_cast<uintptr_t>(seq + start) & 3u);
}

__device__ __forceinline__ void seed_stage_warp(const uint8_t *__restrict__ seq, long long my_start, int my_len,
                                                uint8_t *warp_slots, const uint8_t *class_s, int lane)
{
    for (int r0 = 0; r0 < 32; r0 += 4) {
        uint32_t w[4][2];
        int nw[4];
#pragma unroll
        for (int j = 0; j < 4; j++) {
            const long long start = __shfl_sync(0xFFFFFFFFu, my_start, r0 + j);
            const int len = __shfl_sync(0xFFFFFFFFu, my_len, r0 + j);
            const uintptr_t addr = reinterpret_cast<uintptr_t>(seq + start);
            const int mis = (int)(addr & 3u);
            nw[j] = len ? (mis + len + 3) >> 2 : 0;
            const uint32_t *base = reinterpret_cast<const uint32_t *>(addr - (uintptr_t)mis);
            w[j][0] = lane < nw[j] ? __ldg(base + lane) : 0u;
            w[j][1] = lane + 32 < nw[j] ? __ldg(base + lane + 32) : 0u;
        }
#pragma unroll
        for (int j = 0; j < 4; j++) {
            uint32_t *dst = reinterpret_cast<uint32_t *>(warp_slots + (size_t)(r0 + j) * kSeedSlot);
#pragma unroll
            for (int h = 0; h < 2; h++) {
                const int wi = lane + 32 * h;
                if (wi >= nw[j]) continue;
                const uint32_t v = w[j][h];
                dst[wi] = (uint32_t)class_s[v & 0xFFu] | ((uint32_t)class_s[(v >> 8) & 0xFFu] << 8) |
                          ((uint32_t)class_s[(v >> 16) & 0xFFu] << 16) | ((uint32_t)class_s[v >> 24] << 24);
            }
        }
    }
}

// Hit record: barcode << 13 | diagonal span << 10 | lowest diagonal + 256   (barcode < 2^14, span <= K <= 7)
__device__ __forceinline__ uint32_t hit_pack(uint32_t b, int span, int dmin)
{
    return (b << 13) | ((uint32_t)span << 10) | (uint32_t)(dmin + 256);
}

struct SeedVerifyCtx {
    uint32_t *hits;                 // [kSeedMaxHits][kSeedThreads], column warp * 32 + lane belongs to a lane;
                                    // bits 27..30 of a record receive the verified distance (15 = more than K)
    const uint32_t *peq_s;
    const uint8_t *warp_slots;      // slot of lane 0 of this warp
    int n_pad, plane, m, K, win, total;      // plane: words per 32-bit plane of the Peq table (barcodes > 32 nt use two)
};

// Verifies pooled hits [i0, i0 + 32 * ILP) of the warp, one per lane and chain: windowed Myers/Hyyro automaton
// over the columns the hit's alignment can occupy.  The hits of the warp's 32 reads form one pool (lane L owns
// the indices [excl(L), incl(L)) of the inclusive prefix sum `incl` of the per-lane hit counts), so every lane
// has work whatever its own read found; the distance goes back into the hit record for its owner.  Branch-free so that the ILP chains of a lane interleave: past the
// end of its window a chain keeps stepping on class 0 and its minimum is not updated.
template <int ILP, typename WT>      // WT: uint32_t for barcodes up to 32 nt, unsigned long long up to 64
__device__ __forceinline__ void seed_verify(const SeedVerifyCtx &v, int i0, int lane, int incl, int start_j, int end_j, int my_skew)
{
    constexpr int kMsb = (int)sizeof(WT) * 8 - 1;
    const WT row_mask = v.m > kMsb ? ~(WT)0 : (~(WT)0 << (kMsb + 1 - v.m));     // barcode rows top-aligned
    int hb[ILP], c0[ILP], c1[ILP], score[ILP], best[ILP], owner[ILP], at[ILP];
    WT pv[ILP], mv[ILP];
    uint32_t recs[ILP];
    const uint8_t *slot[ILP];
#pragma unroll
    for (int u = 0; u < ILP; u++) {
        const int i = i0 + 32 * u + lane;
        const bool live = i < v.total;
        // owner = first lane whose inclusive prefix exceeds i (binary search over the lanes by shuffles)
        int lo = 0;
#pragma unroll
        for (int step = 16; step >= 1; step >>= 1) {
            const int probe = __shfl_sync(0xFFFFFFFFu, incl, lo + step - 1);
            if (probe <= i) lo += step;
        }
        owner[u] = live ? lo : 0;
        const int prev_incl = __shfl_sync(0xFFFFFFFFu, incl, max(owner[u] - 1, 0));   // all lanes shuffle
        const int o_excl = owner[u] ? prev_incl : 0;
        at[u] = live ? (i - o_excl) * kSeedThreads + owner[u] : -1;
        const uint32_t rec = live ? v.hits[at[u]] : 0u;
        recs[u] = rec;
        hb[u] = (int)(rec >> 13);
        const int dmin = (int)(rec & 0x3FFu) - 256, span = (int)((rec >> 10) & 0x7u);
        const int sj = __shfl_sync(0xFFFFFFFFu, start_j, owner[u]);
        const int ej = __shfl_sync(0xFFFFFFFFu, end_j, owner[u]);
        // 1-based columns an alignment with <= K edits on these diagonals can occupy
        c0[u] = live ? max(sj, dmin + 1 - v.K) : 1;
        c1[u] = live ? min(ej, dmin + span + v.m + 2 * v.K) : 0;
        slot[u] = v.warp_slots + (size_t)owner[u] * kSeedSlot + __shfl_sync(0xFFFFFFFFu, my_skew, owner[u]);
        pv[u] = row_mask;
        mv[u] = 0;
        score[u] = v.m;
        best[u] = kInf;
    }
    for (int t = 0; t < v.win; t++) {
        WT eq[ILP];
        bool in[ILP];
#pragma unroll
        for (int u = 0; u < ILP; u++) {
            const int c = c0[u] + t;
            in[u] = c <= c1[u];
            // outside the window nothing staged may be read: class 0 then
            const uint32_t cls = in[u] ? (uint32_t)slot[u][c - 1] : 0u;
            eq[u] = v.peq_s[cls * v.n_pad + hb[u]];
            if (sizeof(WT) == 8) eq[u] |= (WT)v.peq_s[v.plane + cls * v.n_pad + hb[u]] << (kMsb - 31);
        }
#pragma unroll
        for (int u = 0; u < ILP; u++) {
            const WT xv = eq[u] | mv[u];
            const WT xh = ((((eq[u] & pv[u]) + pv[u]) ^ pv[u]) | eq[u]);
            const WT ph = mv[u] | ~(xh | pv[u]);
            const WT mh = pv[u] & xh;
            score[u] += (int)(ph >> kMsb) - (int)(mh >> kMsb);
            const WT phs = ph << 1, mhs = mh << 1;
            pv[u] = mhs | ~(xv | phs);
            mv[u] = phs & xv;
            best[u] = in[u] ? min(best[u], score[u]) : best[u];
        }
    }
#pragma unroll
    for (int u = 0; u < ILP; u++)
        if (at[u] >= 0) v.hits[at[u]] = recs[u] | ((uint32_t)(best[u] <= v.K ? best[u] : 15) << 27);
}

}  // namespace bdx
