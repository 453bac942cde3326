// demux.cu -- device-side FASTQ block demultiplexer: raw FASTQ text in, one contiguous run of
// (trimmed) records per output file out.  SURVEY.md section 8f-1 (reader_task's record parsing,
// reference src/core.jl:43-110) and 8f-3 (writer_task's trimming and per-file routing,
// src/core.jl:118-224) around the classification path.
//
// Stages, all on the stream's compute CUDA stream (every one is a byte/integer streaming kernel
// bound by HBM; none has a GEMM shape):
//   1 k_fq_count / k_fq_index   newline positions of the block: per-4 KB-tile counts, exclusive scan,
//                               then each tile writes the positions of its newlines at their rank
//   2 k_fq_records              record i = lines 4i .. 4i+3 ("four readlines", core.jl:96-101): start and
//                               length of each line, one trailing "\r" stripped like Julia's readline;
//                               at the end of the input missing lines read as ""
//   3 k_fq_pack                 sequence lines gathered into the packed (bytes, offsets) batch layout
//   4 classification            the kernels of bdx_classify_device, unchanged
//   5 k_part_keys               output-file key (unknown / ambiguous / (bc1, bc2)) and output size of every
//                               record, keep range applied as writer_task does (core.jl:155-173)
//   6 k_rs_hist / k_rs_scatter  stable LSD radix sort of (key, record index): records of one output file
//                               become contiguous and stay in input order (core.jl:146-148)
//   7 k_part_*                  byte offsets by exclusive scan of the sizes in sorted order, bucket table
//   8 k_part_copy               one warp per record writes header\n seq[keep]\n plus\n qual[keep]\n
//                               (write_entry, core.jl:135-137) to its final place
// The host appends each bucket's byte range to that bucket's file: one write per file and block
// instead of one per record.
#include <algorithm>
#include <cstdio>
#include <string>

#include "bdx_internal.h"
#include "demux.h"

namespace bdx {

// ---------------------------------------------------------------------------------------
// exclusive scan: tiles of 2048 elements, tile totals scanned recursively, then added back
// ---------------------------------------------------------------------------------------
constexpr int kScanThreads = 256;
constexpr int kScanItems = 8;
constexpr int kScanTile = kScanThreads * kScanItems;

template <typename T>
__global__ void __launch_bounds__(kScanThreads)
k_scan_tile(const T *in, T *out, T *tile_sums, const long long n_in, const long long n_out)
{
    __shared__ T warp_tot[kScanThreads / 32];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const long long base = (long long)blockIdx.x * kScanTile + (long long)threadIdx.x * kScanItems;
    T v[kScanItems];
    T sum = 0;
#pragma unroll
    for (int k = 0; k < kScanItems; k++) {
        v[k] = base + k < n_in ? in[base + k] : (T)0;
        sum += v[k];
    }
    T inc = sum;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        const T t = __shfl_up_sync(0xFFFFFFFFu, inc, o);
        if (lane >= o) inc += t;
    }
    if (lane == 31) warp_tot[warp] = inc;
    __syncthreads();
    T wbase = 0;
    for (int w = 0; w < warp; w++) wbase += warp_tot[w];
    T run = wbase + inc - sum;
#pragma unroll
    for (int k = 0; k < kScanItems; k++) {
        if (base + k < n_out) out[base + k] = run;
        run += v[k];
    }
    if (threadIdx.x == kScanThreads - 1) tile_sums[blockIdx.x] = run;
}

template <typename T>
__global__ void __launch_bounds__(kScanThreads)
k_scan_add(T *out, const T *tile_offs, const long long n_out)
{
    const T add = tile_offs[blockIdx.x];
    const long long base = (long long)blockIdx.x * kScanTile + (long long)threadIdx.x * kScanItems;
#pragma unroll
    for (int k = 0; k < kScanItems; k++)
        if (base + k < n_out) out[base + k] += add;
}

static size_t scan_ws_bytes(long long n_out, size_t elem)
{
    size_t total = 0;
    long long tiles = (n_out + kScanTile - 1) / kScanTile;
    while (true) {
        total += ((size_t)tiles * elem + 255) & ~(size_t)255;
        if (tiles <= 1) break;
        tiles = (tiles + kScanTile - 1) / kScanTile;
    }
    return total;
}

// out[i] = sum of in[0, i) for i < n_out (n_out may be n_in + 1: the last entry is the total).
// in == out is allowed.  ws must hold scan_ws_bytes(n_out, sizeof(T)).
template <typename T>
static cudaError_t scan_exclusive(const T *in, T *out, long long n_in, long long n_out, void *ws, cudaStream_t st)
{
    if (n_out <= 0) return cudaSuccess;
    const long long tiles = (n_out + kScanTile - 1) / kScanTile;
    T *sums = (T *)ws;
    k_scan_tile<T><<<(unsigned)tiles, kScanThreads, 0, st>>>(in, out, sums, n_in, n_out);
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess || tiles == 1) return e;
    void *next = (char *)ws + (((size_t)tiles * sizeof(T) + 255) & ~(size_t)255);
    e = scan_exclusive<T>(sums, sums, tiles, tiles, next, st);
    if (e != cudaSuccess) return e;
    k_scan_add<T><<<(unsigned)tiles, kScanThreads, 0, st>>>(out, sums, n_out);
    return cudaGetLastError();
}

// ---------------------------------------------------------------------------------------
// 1: newline index
// ---------------------------------------------------------------------------------------
constexpr int kFqTile = 4096;       // bytes per block: 256 threads x 16 bytes
constexpr int kFqThreads = 256;

// 16 bytes at text + pos; bytes at or beyond len read as 0.  text is 16-byte aligned.
__device__ __forceinline__ uint4 fq_load16(const uint8_t *__restrict__ text, long long pos, long long len)
{
    if (pos + 16 <= len) return __ldg(reinterpret_cast<const uint4 *>(text + pos));
    uint32_t w[4] = {0u, 0u, 0u, 0u};
    for (int k = 0; k < 16; k++)
        if (pos + k < len) w[k >> 2] |= (uint32_t)text[pos + k] << (8 * (k & 3));
    return make_uint4(w[0], w[1], w[2], w[3]);
}

__device__ __forceinline__ uint32_t nl_flags(uint32_t w) { return __vcmpeq4(w, 0x0A0A0A0Au) & 0x01010101u; }

__global__ void __launch_bounds__(kFqThreads)
k_fq_count(const uint8_t *__restrict__ text, const long long len, int *__restrict__ tile_cnt)
{
    __shared__ int warp_tot[kFqThreads / 32];
    const long long pos = (long long)blockIdx.x * kFqTile + threadIdx.x * 16;
    int c = 0;
    if (pos < len) {
        const uint4 v = fq_load16(text, pos, len);
        c = __popc(nl_flags(v.x)) + __popc(nl_flags(v.y)) + __popc(nl_flags(v.z)) + __popc(nl_flags(v.w));
    }
    c = __reduce_add_sync(0xFFFFFFFFu, c);
    if ((threadIdx.x & 31) == 0) warp_tot[threadIdx.x >> 5] = c;
    __syncthreads();
    if (threadIdx.x == 0) {
        int t = 0;
        for (int w = 0; w < kFqThreads / 32; w++) t += warp_tot[w];
        tile_cnt[blockIdx.x] = t;
    }
}

__global__ void __launch_bounds__(kFqThreads)
k_fq_index(const uint8_t *__restrict__ text, const long long len, const int *__restrict__ tile_base,
           int *__restrict__ nl_pos)
{
    __shared__ int warp_tot[kFqThreads / 32];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const long long pos = (long long)blockIdx.x * kFqTile + threadIdx.x * 16;
    uint32_t f[4] = {0u, 0u, 0u, 0u};
    if (pos < len) {
        const uint4 v = fq_load16(text, pos, len);
        f[0] = nl_flags(v.x); f[1] = nl_flags(v.y); f[2] = nl_flags(v.z); f[3] = nl_flags(v.w);
    }
    const int c = __popc(f[0]) + __popc(f[1]) + __popc(f[2]) + __popc(f[3]);
    int inc = c;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        const int t = __shfl_up_sync(0xFFFFFFFFu, inc, o);
        if (lane >= o) inc += t;
    }
    if (lane == 31) warp_tot[warp] = inc;
    __syncthreads();
    int rank = tile_base[blockIdx.x] + inc - c;
    for (int w = 0; w < warp; w++) rank += warp_tot[w];
    if (c) {
#pragma unroll
        for (int k = 0; k < 16; k++)
            if ((f[k >> 2] >> (8 * (k & 3))) & 1u) nl_pos[rank++] = (int)(pos + k);
    }
}

struct FqHdr {               // device -> host after stage 1 and again after stage 7
    int n_lines[2];          // newline count of each side
    int unterminated[2];     // the block does not end with '\n'
    int n_buckets;
    int pad;
    long long out_len[2];
    long long consumed[2];
};

__global__ void k_fq_hdr(const uint8_t *__restrict__ text, const long long len, const int *__restrict__ tile_cnt,
                         const long long n_tiles, FqHdr *hdr, const int side)
{
    hdr->n_lines[side] = n_tiles ? tile_cnt[n_tiles] : 0;
    hdr->unterminated[side] = len > 0 && text[len - 1] != '\n';
}

// ---------------------------------------------------------------------------------------
// 2, 3: records and the packed sequence batch
// ---------------------------------------------------------------------------------------
struct FqRec {
    int start[4];   // header, sequence, plus, quality line
    int len[4];     // terminators stripped
};

__global__ void __launch_bounds__(256)
k_fq_records(const uint8_t *__restrict__ text, const int len, const int *__restrict__ nl_pos, const int n_nl,
             const int unterminated, const int n_rec, FqRec *__restrict__ recs, int *__restrict__ seq_len)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n_rec) return;
    FqRec r;
#pragma unroll
    for (int k = 0; k < 4; k++) {
        const int L = 4 * i + k;
        int s = L == 0 ? 0 : (L - 1 < n_nl ? nl_pos[L - 1] + 1 : len);
        int e;
        if (L < n_nl) {
            e = nl_pos[L];
            if (e > s && text[e - 1] == '\r') e--;       // readline strips "\r\n" as well as "\n"
        } else if (L == n_nl && unterminated) {
            e = len;                                     // last line of the input, no terminator: kept as is
        } else {
            s = len;                                     // past the end of the input: ""
            e = len;
        }
        r.start[k] = s;
        r.len[k] = e - s;
    }
    recs[i] = r;
    if (seq_len) seq_len[i] = r.len[1];
}

// Warp-wide byte copy with 32-bit stores: the destination is brought to 4-byte alignment, then every lane
// moves one word per step, assembled from the two aligned source words that hold its bytes (funnel shift).
// The aligned source words read never reach past the word that holds the last source byte.
__device__ __forceinline__ void warp_copy(uint8_t *__restrict__ dst, const uint8_t *__restrict__ src, int len, int lane)
{
    const int head = min(len, (int)((4u - (uint32_t)(reinterpret_cast<uintptr_t>(dst) & 3u)) & 3u));
    if (lane < head) dst[lane] = src[lane];
    dst += head;
    src += head;
    len -= head;
    const int words = len >> 2;
    const uint32_t mis = (uint32_t)(reinterpret_cast<uintptr_t>(src) & 3u);
    const uint32_t *sw = reinterpret_cast<const uint32_t *>(src - mis);
    uint32_t *dw = reinterpret_cast<uint32_t *>(dst);
    // four words per lane are loaded before the first store, so a whole 512-byte stretch is in flight
    if (mis == 0) {
        for (int w0 = lane; w0 < words; w0 += 128) {
            uint32_t v[4];
#pragma unroll
            for (int u = 0; u < 4; u++)
                if (w0 + 32 * u < words) v[u] = sw[w0 + 32 * u];
#pragma unroll
            for (int u = 0; u < 4; u++)
                if (w0 + 32 * u < words) dw[w0 + 32 * u] = v[u];
        }
    } else {
        const uint32_t sh = mis * 8u;
        for (int w0 = lane; w0 < words; w0 += 128) {
            uint32_t v[4];
#pragma unroll
            for (int u = 0; u < 4; u++)
                if (w0 + 32 * u < words) v[u] = __funnelshift_r(sw[w0 + 32 * u], sw[w0 + 32 * u + 1], sh);
#pragma unroll
            for (int u = 0; u < 4; u++)
                if (w0 + 32 * u < words) dw[w0 + 32 * u] = v[u];
        }
    }
    const int done = words << 2;
    if (lane < len - done) dst[done + lane] = src[done + lane];
}

// One warp per 32 records: the lanes fetch the 32 records' metadata together, then the warp copies the
// records one after the other (a warp per record would pay the metadata latency once per record).
__global__ void __launch_bounds__(256)
k_fq_pack(const uint8_t *__restrict__ text, const FqRec *__restrict__ recs, const int *__restrict__ off,
          const int n_rec, uint8_t *__restrict__ seq)
{
    const int lane = threadIdx.x & 31;
    const int i0 = (blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5)) * 32;
    if (i0 >= n_rec) return;
    const int i = i0 + lane;
    int src = 0, len = 0, dst = 0;
    if (i < n_rec) {
        src = recs[i].start[1];
        len = recs[i].len[1];
        dst = off[i];
    }
    const int cnt = min(32, n_rec - i0);
    for (int r = 0; r < cnt; r++)
        warp_copy(seq + __shfl_sync(0xFFFFFFFFu, dst, r), text + __shfl_sync(0xFFFFFFFFu, src, r),
                  __shfl_sync(0xFFFFFFFFu, len, r), lane);
}

// ---------------------------------------------------------------------------------------
// 5: keys and sizes
// ---------------------------------------------------------------------------------------
// keep range of read 1 as writer_task applies it (core.jl:155-173): only when trimming is configured and
// the classification produced a range (keep_start != -1, core.jl:250-254); r = intersect(keep, 1:len).
// Returns the 0-based first kept column and sets n_seq / n_qual (the quality line is cut with the same range;
// a quality line shorter than the sequence -- malformed FASTQ, a BoundsError in the reference -- is clamped).
__device__ __forceinline__ int keep_of(const bdx_result &r, int do_trim, int seq_len, int qual_len, int &n_seq,
                                       int &n_qual)
{
    if (!do_trim || r.keep_start == -1) {
        n_seq = seq_len;
        n_qual = qual_len;
        return 0;
    }
    const int lo = max(r.keep_start, 1), hi = min(r.keep_end, seq_len);
    n_seq = max(hi - lo + 1, 0);
    n_qual = n_seq ? max(min(hi, qual_len) - lo + 1, 0) : 0;
    return n_seq ? lo - 1 : 0;
}

__global__ void __launch_bounds__(256)
k_part_keys(const bdx_result *__restrict__ res, const FqRec *__restrict__ recs1, const FqRec *__restrict__ recs2,
            const int n, const int do_trim, const int b2_eff, int *__restrict__ key, int *__restrict__ size1,
            int *__restrict__ size2)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const bdx_result r = res[i];
    // 0 = unknown, 1 = ambiguous_classification, 2 + (bc1-1) * max(B2,1) + (bc2-1)   (classification.jl:877-900)
    key[i] = r.status == BDX_UNKNOWN ? 0 : (r.status == BDX_AMBIGUOUS ? 1 : 2 + (r.bc1 - 1) * b2_eff + max(r.bc2 - 1, 0));
    if (size1) {
        const FqRec a = recs1[i];
        int ns, nq;
        keep_of(r, do_trim, a.len[1], a.len[3], ns, nq);
        size1[i] = a.len[0] + ns + a.len[2] + nq + 4;
    }
    if (size2) {
        const FqRec b = recs2[i];
        size2[i] = b.len[0] + b.len[1] + b.len[2] + b.len[3] + 4;
    }
}

// ---------------------------------------------------------------------------------------
// 6: stable LSD radix sort of (key, record index), 8 bits per pass
// ---------------------------------------------------------------------------------------
constexpr int kRsThreads = 256;
constexpr int kRsRounds = 8;
constexpr int kRsTile = kRsThreads * kRsRounds;   // warp w owns elements [w*256, (w+1)*256) of the tile

__global__ void __launch_bounds__(kRsThreads)
k_rs_hist(const int *__restrict__ keys, const int n, const int shift, int *__restrict__ hist, const int n_tiles)
{
    __shared__ int h[256];
    h[threadIdx.x] = 0;
    __syncthreads();
    const int base = blockIdx.x * kRsTile;
#pragma unroll
    for (int k = 0; k < kRsRounds; k++) {
        const int e = base + k * kRsThreads + threadIdx.x;
        if (e < n) atomicAdd(&h[(keys[e] >> shift) & 255], 1);
    }
    __syncthreads();
    hist[(size_t)threadIdx.x * n_tiles + blockIdx.x] = h[threadIdx.x];   // digit-major: one scan gives all bases
}

__global__ void __launch_bounds__(kRsThreads)
k_rs_scatter(const int *__restrict__ keys_in, const int *__restrict__ vals_in, int *__restrict__ keys_out,
             int *__restrict__ vals_out, const int n, const int shift, const int *__restrict__ hist_scanned,
             const int n_tiles)
{
    __shared__ int wh[kRsThreads / 32][256];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    for (int k = threadIdx.x; k < (kRsThreads / 32) * 256; k += kRsThreads) (&wh[0][0])[k] = 0;
    __syncthreads();
    int key[kRsRounds], val[kRsRounds], lr[kRsRounds];
    const int base = blockIdx.x * kRsTile + warp * (kRsRounds * 32);
#pragma unroll
    for (int r = 0; r < kRsRounds; r++) {
        const int e = base + r * 32 + lane;
        const bool valid = e < n;
        key[r] = valid ? keys_in[e] : 0;
        val[r] = valid ? (vals_in ? vals_in[e] : e) : 0;
        const int d = valid ? (key[r] >> shift) & 255 : 256;
        // lanes with the same digit: rank among them by lane = by input position (stable)
        const uint32_t mask = __match_any_sync(0xFFFFFFFFu, d);
        const int leader = __ffs(mask) - 1;
        int old = 0;
        if (lane == leader && valid) {
            old = wh[warp][d];
            wh[warp][d] = old + __popc(mask);
        }
        old = __shfl_sync(0xFFFFFFFFu, old, leader);
        lr[r] = old + __popc(mask & ((1u << lane) - 1u));
        __syncwarp();
    }
    __syncthreads();
    {
        const int d = threadIdx.x;   // one thread per digit: tile base + counts of the earlier warps
        int run = hist_scanned[(size_t)d * n_tiles + blockIdx.x];
#pragma unroll
        for (int w = 0; w < kRsThreads / 32; w++) {
            const int t = wh[w][d];
            wh[w][d] = run;
            run += t;
        }
    }
    __syncthreads();
#pragma unroll
    for (int r = 0; r < kRsRounds; r++) {
        const int e = base + r * 32 + lane;
        if (e < n) {
            const int dst = wh[warp][(key[r] >> shift) & 255] + lr[r];
            keys_out[dst] = key[r];
            vals_out[dst] = val[r];
        }
    }
}

// ---------------------------------------------------------------------------------------
// 7: offsets and buckets
// ---------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
k_part_gather(const int *__restrict__ sidx, const int *__restrict__ skey, const int n, const int *__restrict__ size1,
              const int *__restrict__ size2, long long *__restrict__ ss1, long long *__restrict__ ss2,
              int *__restrict__ flag, int *__restrict__ rank)
{
    const int p = blockIdx.x * blockDim.x + threadIdx.x;
    if (p >= n) return;
    const int i = sidx[p];
    rank[i] = p;                  // where record i stands in the output order
    if (ss1) ss1[p] = size1[i];
    if (ss2) ss2[p] = size2[i];
    flag[p] = p == 0 || skey[p] != skey[p - 1];
}

__global__ void __launch_bounds__(256)
k_part_bstart(const int *__restrict__ skey, const int *__restrict__ bid, const int n, int *__restrict__ bstart,
              int *__restrict__ bkey)
{
    const int p = blockIdx.x * blockDim.x + threadIdx.x;
    if (p >= n) return;
    if (p == 0 || skey[p] != skey[p - 1]) {
        bstart[bid[p]] = p;
        bkey[bid[p]] = skey[p];
    }
    if (p == n - 1) bstart[bid[n]] = n;
}

__global__ void __launch_bounds__(256)
k_part_buckets(const int *__restrict__ bstart, const int *__restrict__ bkey, const int *__restrict__ n_buckets_p,
               const long long *__restrict__ oo1, const long long *__restrict__ oo2, const int n, const int b2_eff,
               const int is_dual, bdx_demux_bucket *__restrict__ buckets, FqHdr *hdr)
{
    const int nb = *n_buckets_p;
    const int b = blockIdx.x * blockDim.x + threadIdx.x;
    if (b == 0) {
        hdr->n_buckets = nb;
        hdr->out_len[0] = oo1 ? oo1[n] : 0;
        hdr->out_len[1] = oo2 ? oo2[n] : 0;
    }
    if (b >= nb) return;
    const int p0 = bstart[b], p1 = bstart[b + 1], k = bkey[b];
    bdx_demux_bucket o;
    o.status = k == 0 ? BDX_UNKNOWN : (k == 1 ? BDX_AMBIGUOUS : BDX_MATCH);
    o.bc1 = k >= 2 ? (k - 2) / b2_eff + 1 : 0;
    o.bc2 = (k >= 2 && is_dual) ? (k - 2) % b2_eff + 1 : 0;
    o.n_records = p1 - p0;
    o.offset1 = oo1 ? oo1[p0] : 0;
    o.length1 = oo1 ? oo1[p1] - oo1[p0] : 0;
    o.offset2 = oo2 ? oo2[p0] : 0;
    o.length2 = oo2 ? oo2[p1] - oo2[p0] : 0;
    buckets[b] = o;
}

// ---------------------------------------------------------------------------------------
// 8: write_entry (core.jl:135-137) for every record, at its final place
// ---------------------------------------------------------------------------------------
// One warp per 32 consecutive INPUT records: every lane prepares one record (its source runs and where it goes),
// then the warp copies the records one after the other.  Input order keeps the reads of the text and of the
// per-record tables sequential; the writes advance one frontier per output file, which the L2 absorbs.
__global__ void __launch_bounds__(256)
k_part_copy(const uint8_t *__restrict__ text, const int text_len, const FqRec *__restrict__ recs,
            const bdx_result *__restrict__ res, const int do_trim, const int *__restrict__ rank,
            const long long *__restrict__ ooff, const int n, uint8_t *__restrict__ out)
{
    const int lane = threadIdx.x & 31;
    const int p0 = (blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5)) * 32;
    if (p0 >= n) return;
    const int p = p0 + lane;      // input record of this lane
    // the four output lines as source runs; a run whose "\n" sits right behind it in the text takes it along,
    // and runs that are adjacent in the text merge: an untrimmed "\n"-terminated record is ONE copy.
    // run_len: bytes to copy (incl. a trailing "\n" taken from the text); run_nl: 1 = write the "\n" separately
    int run_src[4] = {0, 0, 0, 0}, run_len[4] = {0, 0, 0, 0}, run_nl[4] = {0, 0, 0, 0}, n_runs = 0;
    long long dst0 = 0;
    if (p < n) {
        const int i = p;
        const FqRec r = recs[i];
        int ns = r.len[1], nq = r.len[3], first = 0;
        if (res) first = keep_of(res[i], do_trim, r.len[1], r.len[3], ns, nq);
        const int src[4] = {r.start[0], r.start[1] + first, r.start[2], r.start[3] + first};
        const int len[4] = {r.len[0], ns, r.len[2], nq};
        bool nl[4];      // the text has the line's own "\n" directly behind the run
        nl[0] = r.start[1] == r.start[0] + r.len[0] + 1;
        nl[1] = first + ns == r.len[1] && r.start[2] == r.start[1] + r.len[1] + 1;
        nl[2] = r.start[3] == r.start[2] + r.len[2] + 1;
        const int e = r.start[3] + r.len[3];
        nl[3] = first + nq == r.len[3] && e < text_len && text[e] == '\n';
        dst0 = ooff[rank[i]];
#pragma unroll
        for (int k = 0; k < 4; k++) {
            // line k extends the current run when it starts right after that run's "\n"
            const bool joins = k > 0 && run_nl[n_runs - 1] == 0 && src[k] == run_src[n_runs - 1] + run_len[n_runs - 1];
            if (joins) {
                run_len[n_runs - 1] += len[k] + (nl[k] ? 1 : 0);
                run_nl[n_runs - 1] = nl[k] ? 0 : 1;
            } else {
                run_src[n_runs] = src[k];
                run_len[n_runs] = len[k] + (nl[k] ? 1 : 0);
                run_nl[n_runs] = nl[k] ? 0 : 1;
                n_runs++;
            }
        }
    }
    const int cnt = min(32, n - p0);
    for (int r = 0; r < cnt; r++) {
        uint8_t *dst = out + __shfl_sync(0xFFFFFFFFu, dst0, r);
        const int nr = __shfl_sync(0xFFFFFFFFu, n_runs, r);
#pragma unroll
        for (int k = 0; k < 4; k++) {
            const int rs = __shfl_sync(0xFFFFFFFFu, run_src[k], r);
            const int rl = __shfl_sync(0xFFFFFFFFu, run_len[k], r);
            const int rn = __shfl_sync(0xFFFFFFFFu, run_nl[k], r);
            if (k < nr) {
                warp_copy(dst, text + rs, rl, lane);
                if (rn && lane == 0) dst[rl] = '\n';
                dst += rl + rn;
            }
        }
    }
}

// ---------------------------------------------------------------------------------------
// host orchestration
// ---------------------------------------------------------------------------------------
struct DBuf {
    void *p = nullptr;
    size_t cap = 0;
    bool host = false;
    cudaError_t reserve(size_t bytes, cudaStream_t st)
    {
        if (bytes <= cap) return cudaSuccess;
        if (p) {
            cudaError_t e = cudaStreamSynchronize(st);   // queued kernels may still use the old block
            if (e != cudaSuccess) return e;
            host ? cudaFreeHost(p) : cudaFree(p);
            p = nullptr;
            cap = 0;
        }
        const size_t want = bytes + bytes / 4 + 4096;
        cudaError_t e = host ? cudaHostAlloc(&p, want, cudaHostAllocDefault) : cudaMalloc(&p, want);
        if (e == cudaSuccess) cap = want;
        return e;
    }
    void release()
    {
        if (p) host ? cudaFreeHost(p) : cudaFree(p);
        p = nullptr;
        cap = 0;
    }
    template <typename T> T *as() const { return (T *)p; }
};

struct DemuxSide {
    DBuf text, tile_cnt, nl_pos, recs, size, ssize, ooff, out, h_out;
};

struct DemuxState {
    DemuxSide side[2];
    DBuf hdr, h_hdr, seq_len, seq, res, h_res, key[2], idx[2], hist, flag, bstart, bkey, buckets, h_buckets, scan_ws;
    cudaEvent_t ev[10] = {};
    bool have_ev = false;
    float stage_ms[8] = {};
};

DemuxState *demux_state_create()
{
    DemuxState *d = new (std::nothrow) DemuxState();
    if (!d) return nullptr;
    d->h_hdr.host = d->h_res.host = d->h_buckets.host = true;
    d->side[0].h_out.host = d->side[1].h_out.host = true;
    return d;
}

void demux_state_destroy(DemuxState *d)
{
    if (!d) return;
    for (DemuxSide &s : d->side)
        for (DBuf *b : {&s.text, &s.tile_cnt, &s.nl_pos, &s.recs, &s.size, &s.ssize, &s.ooff, &s.out, &s.h_out}) b->release();
    for (DBuf *b : {&d->hdr, &d->h_hdr, &d->seq_len, &d->seq, &d->res, &d->h_res, &d->key[0], &d->key[1], &d->idx[0],
                    &d->idx[1], &d->hist, &d->flag, &d->bstart, &d->bkey, &d->buckets, &d->h_buckets, &d->scan_ws})
        b->release();
    if (d->have_ev)
        for (cudaEvent_t e : d->ev) cudaEventDestroy(e);
    delete d;
}

const float *demux_stage_ms(const DemuxState *d) { return d->stage_ms; }

#define DM(call)                                                                          \
    do {                                                                                  \
        cudaError_t e__ = (call);                                                         \
        if (e__ != cudaSuccess) {                                                         \
            err = std::string(#call) + ": " + cudaGetErrorString(e__);                    \
            return BDX_ERR_CUDA;                                                          \
        }                                                                                 \
    } while (0)

int demux_run(DemuxState *d, const DevParams &P, cudaStream_t st, const DemuxClassifyFn &classify,
              const uint8_t *fq1, int64_t len1, const uint8_t *fq2, int64_t len2, int final_block, int mode,
              int64_t *launches, bdx_demux_out *out, std::string &err)
{
    const bool dev_io = (mode & BDX_DEMUX_DEVICE_IO) != 0;
    const int kind = mode & 3;
    const int n_sides = kind == BDX_DEMUX_SINGLE ? 1 : 2;
    const bool emit1 = kind != BDX_DEMUX_MATES, emit2 = kind != BDX_DEMUX_SINGLE;
    // single-end: any non-zero value; paired: bit 0 = file 1 ends with this block, bit 1 = file 2
    const bool fin[2] = {kind == BDX_DEMUX_SINGLE ? final_block != 0 : (final_block & 1) != 0, (final_block & 2) != 0};
    const uint8_t *fq[2] = {fq1, fq2};
    const int64_t len[2] = {len1, n_sides == 2 ? len2 : 0};
    memset(out, 0, sizeof(*out));
    for (int s = 0; s < n_sides; s++) {
        if (len[s] < 0 || (len[s] > 0 && !fq[s])) { err = "bad FASTQ block argument"; return BDX_ERR_INVALID; }
        if (len[s] > 0x7FFFFF00ll) { err = "FASTQ block must stay below 2^31 bytes"; return BDX_ERR_TOO_LARGE; }
        if (dev_io && (reinterpret_cast<uintptr_t>(fq[s]) & 15)) { err = "device FASTQ block must be 16-byte aligned"; return BDX_ERR_INVALID; }
    }
    const int b2_eff = P.is_dual ? P.set[1].n_bc : 1;
    const long long n_keys = 2 + (long long)P.set[0].n_bc * b2_eff;
    if (n_keys > 0x7FFFFFFFll) { err = "too many output files for the device partitioner"; return BDX_ERR_INVALID; }
    const int do_trim = P.set[0].trim_side != 0 || (P.is_dual && P.set[1].trim_side != 0);   // core.jl:240

    if (!d->have_ev) {
        for (cudaEvent_t &e : d->ev) DM(cudaEventCreate(&e));
        d->have_ev = true;
    }
    DM(d->hdr.reserve(sizeof(FqHdr), st));
    DM(d->h_hdr.reserve(sizeof(FqHdr), st));
    FqHdr *hdr = d->hdr.as<FqHdr>(), *h_hdr = d->h_hdr.as<FqHdr>();
    DM(cudaMemsetAsync(hdr, 0, sizeof(FqHdr), st));
    DM(cudaEventRecord(d->ev[0], st));

    // ---- stage 1a: H2D, newline counts ----
    const uint8_t *text[2] = {nullptr, nullptr};
    long long n_tiles[2] = {0, 0};
    for (int s = 0; s < n_sides; s++) {
        DemuxSide &S = d->side[s];
        if (dev_io) {
            text[s] = fq[s];
        } else {
            DM(S.text.reserve((size_t)len[s] + 16, st));
            if (len[s]) DM(cudaMemcpyAsync(S.text.p, fq[s], (size_t)len[s], cudaMemcpyHostToDevice, st));
            text[s] = S.text.as<uint8_t>();
        }
    }
    DM(cudaEventRecord(d->ev[1], st));
    for (int s = 0; s < n_sides; s++) {
        DemuxSide &S = d->side[s];
        n_tiles[s] = (len[s] + kFqTile - 1) / kFqTile;
        DM(S.tile_cnt.reserve(((size_t)n_tiles[s] + 1) * 4, st));
        DM(d->scan_ws.reserve(scan_ws_bytes(n_tiles[s] + 1, 4), st));
        if (n_tiles[s]) {
            k_fq_count<<<(unsigned)n_tiles[s], kFqThreads, 0, st>>>(text[s], len[s], S.tile_cnt.as<int>());
            DM(cudaGetLastError());
            DM(scan_exclusive<int>(S.tile_cnt.as<int>(), S.tile_cnt.as<int>(), n_tiles[s], n_tiles[s] + 1, d->scan_ws.p, st));
            *launches += 2;
        }
        k_fq_hdr<<<1, 1, 0, st>>>(text[s], len[s], S.tile_cnt.as<int>(), n_tiles[s], hdr, s);
        DM(cudaGetLastError());
        *launches += 1;
    }
    DM(cudaMemcpyAsync(h_hdr, hdr, sizeof(FqHdr), cudaMemcpyDeviceToHost, st));
    DM(cudaStreamSynchronize(st));

    // ---- how many records (reader_task: `while !eof(io1) && !eof(io2)`, four readlines each) ----
    long long n_rec_side[2] = {0, 0};
    for (int s = 0; s < n_sides; s++) {
        const long long lines = (long long)h_hdr->n_lines[s] + (fin[s] ? h_hdr->unterminated[s] : 0);
        n_rec_side[s] = fin[s] ? (lines + 3) / 4 : lines / 4;
    }
    const long long n_ll = n_sides == 2 ? std::min(n_rec_side[0], n_rec_side[1]) : n_rec_side[0];
    const int n = (int)n_ll;
    out->n_records = n;
    if (n == 0) {
        for (int s = 0; s < n_sides; s++) (s ? out->consumed2 : out->consumed1) = (fin[s] && n_rec_side[s] == 0) ? len[s] : 0;
        return BDX_OK;
    }

    // ---- stage 1b + 2: newline positions, records ----
    for (int s = 0; s < n_sides; s++) {
        DemuxSide &S = d->side[s];
        DM(S.nl_pos.reserve(((size_t)h_hdr->n_lines[s] + 1) * 4, st));
        DM(S.recs.reserve((size_t)n * sizeof(FqRec), st));
    }
    DM(d->seq_len.reserve(((size_t)n + 1) * 4, st));
    DM(d->scan_ws.reserve(std::max(scan_ws_bytes((long long)n + 1, 8), scan_ws_bytes(256ll * ((n + kRsTile - 1) / kRsTile), 4)), st));
    for (int s = 0; s < n_sides; s++) {
        DemuxSide &S = d->side[s];
        k_fq_index<<<(unsigned)n_tiles[s], kFqThreads, 0, st>>>(text[s], len[s], S.tile_cnt.as<int>(), S.nl_pos.as<int>());
        DM(cudaGetLastError());
        k_fq_records<<<(n + 255) / 256, 256, 0, st>>>(text[s], (int)len[s], S.nl_pos.as<int>(), h_hdr->n_lines[s],
                                                      fin[s] ? h_hdr->unterminated[s] : 0, n, S.recs.as<FqRec>(),
                                                      s == 0 ? d->seq_len.as<int>() : nullptr);
        DM(cudaGetLastError());
        *launches += 2;
    }
    DM(cudaEventRecord(d->ev[2], st));

    // ---- stage 3: packed batch ----
    int *off = d->seq_len.as<int>();
    DM(scan_exclusive<int>(off, off, n, (long long)n + 1, d->scan_ws.p, st));
    DM(d->seq.reserve((size_t)len[0] + 16, st));
    k_fq_pack<<<(n + 255) / 256, 256, 0, st>>>(text[0], d->side[0].recs.as<FqRec>(), off, n, d->seq.as<uint8_t>());
    DM(cudaGetLastError());
    *launches += 2;
    DM(cudaEventRecord(d->ev[3], st));

    // ---- stage 4: classification ----
    DM(d->res.reserve((size_t)n * sizeof(bdx_result), st));
    int rc = classify(d->seq.as<uint8_t>(), off, n, d->res.as<bdx_result>());
    if (rc) return rc;
    DM(cudaEventRecord(d->ev[4], st));

    // ---- stage 5: keys and sizes ----
    for (int k = 0; k < 2; k++) {
        DM(d->key[k].reserve((size_t)n * 4, st));
        DM(d->idx[k].reserve((size_t)n * 4, st));
    }
    for (int s = 0; s < n_sides; s++) {
        DemuxSide &S = d->side[s];
        DM(S.size.reserve((size_t)n * 4, st));
        DM(S.ssize.reserve(((size_t)n + 1) * 8, st));
    }
    k_part_keys<<<(n + 255) / 256, 256, 0, st>>>(d->res.as<bdx_result>(), d->side[0].recs.as<FqRec>(),
                                                 n_sides == 2 ? d->side[1].recs.as<FqRec>() : nullptr, n, do_trim, b2_eff,
                                                 d->key[0].as<int>(), emit1 ? d->side[0].size.as<int>() : nullptr,
                                                 emit2 ? d->side[1].size.as<int>() : nullptr);
    DM(cudaGetLastError());
    *launches += 1;

    // ---- stage 6: stable sort by key ----
    int key_bits = 1;
    while ((1ll << key_bits) < n_keys) key_bits++;
    const int passes = (key_bits + 7) / 8;
    const int rs_tiles = (n + kRsTile - 1) / kRsTile;
    DM(d->hist.reserve((size_t)256 * rs_tiles * 4, st));
    int cur = 0;
    for (int pass = 0; pass < passes; pass++) {
        const int *kin = d->key[cur].as<int>(), *vin = pass == 0 ? nullptr : d->idx[cur].as<int>();
        int *kout = d->key[cur ^ 1].as<int>(), *vout = d->idx[cur ^ 1].as<int>();
        k_rs_hist<<<rs_tiles, kRsThreads, 0, st>>>(kin, n, 8 * pass, d->hist.as<int>(), rs_tiles);
        DM(cudaGetLastError());
        DM(scan_exclusive<int>(d->hist.as<int>(), d->hist.as<int>(), 256ll * rs_tiles, 256ll * rs_tiles, d->scan_ws.p, st));
        k_rs_scatter<<<rs_tiles, kRsThreads, 0, st>>>(kin, vin, kout, vout, n, 8 * pass, d->hist.as<int>(), rs_tiles);
        DM(cudaGetLastError());
        *launches += 3;
        cur ^= 1;
    }
    const int *skey = d->key[cur].as<int>(), *sidx = d->idx[cur].as<int>();
    DM(cudaEventRecord(d->ev[5], st));

    // ---- stage 7: byte offsets, buckets ----
    DM(d->flag.reserve(((size_t)n + 1) * 4, st));
    DM(d->bstart.reserve(((size_t)n + 1) * 4, st));
    DM(d->bkey.reserve((size_t)n * 4, st));
    const long long max_buckets = std::min<long long>(n, n_keys);
    DM(d->buckets.reserve((size_t)max_buckets * sizeof(bdx_demux_bucket), st));
    long long *ss1 = emit1 ? d->side[0].ssize.as<long long>() : nullptr;
    long long *ss2 = emit2 ? d->side[1].ssize.as<long long>() : nullptr;
    int *rank = d->idx[cur ^ 1].as<int>();        // the sort's spare buffer
    k_part_gather<<<(n + 255) / 256, 256, 0, st>>>(sidx, skey, n, d->side[0].size.as<int>(), d->side[1].size.as<int>(), ss1,
                                                   ss2, d->flag.as<int>(), rank);
    DM(cudaGetLastError());
    if (ss1) DM(scan_exclusive<long long>(ss1, ss1, n, (long long)n + 1, d->scan_ws.p, st));
    if (ss2) DM(scan_exclusive<long long>(ss2, ss2, n, (long long)n + 1, d->scan_ws.p, st));
    int *bid = d->flag.as<int>();
    DM(scan_exclusive<int>(bid, bid, n, (long long)n + 1, d->scan_ws.p, st));
    // bid[p] for a flagged p counts the flags before it: its bucket number; bid[n] = number of buckets
    k_part_bstart<<<(n + 255) / 256, 256, 0, st>>>(skey, bid, n, d->bstart.as<int>(), d->bkey.as<int>());
    DM(cudaGetLastError());
    k_part_buckets<<<(unsigned)((max_buckets + 255) / 256), 256, 0, st>>>(d->bstart.as<int>(), d->bkey.as<int>(), bid + n, ss1, ss2, n,
                                                                          b2_eff, P.is_dual, d->buckets.as<bdx_demux_bucket>(), hdr);
    DM(cudaGetLastError());
    *launches += 6;
    DM(cudaEventRecord(d->ev[6], st));

    // ---- stage 8: records to their final place ----
    // an output record is never longer than its input bytes plus the terminators a truncated final record lacks
    for (int s = 0; s < n_sides; s++) {
        if (!(s == 0 ? emit1 : emit2)) continue;
        DemuxSide &S = d->side[s];
        DM(S.out.reserve((size_t)len[s] + 16, st));
        k_part_copy<<<(n + 255) / 256, 256, 0, st>>>(text[s], (int)len[s], S.recs.as<FqRec>(), s == 0 ? d->res.as<bdx_result>() : nullptr, do_trim,
                                                 rank, s == 0 ? ss1 : ss2, n, S.out.as<uint8_t>());
        DM(cudaGetLastError());
        *launches += 1;
    }
    DM(cudaEventRecord(d->ev[7], st));
    DM(cudaMemcpyAsync(h_hdr, hdr, sizeof(FqHdr), cudaMemcpyDeviceToHost, st));
    DM(cudaStreamSynchronize(st));

    // ---- results to the host ----
    out->n_buckets = h_hdr->n_buckets;
    out->out1_len = h_hdr->out_len[0];
    out->out2_len = h_hdr->out_len[1];
    for (int s = 0; s < n_sides; s++) {
        // bytes of this side the n records covered; the caller re-presents the rest with the next block
        long long c;
        if (n == n_rec_side[s] && fin[s]) {
            c = len[s];
        } else {
            int last_nl = 0;   // newline that ends record n - 1
            DM(cudaMemcpyAsync(&last_nl, d->side[s].nl_pos.as<int>() + (4ll * n - 1), 4, cudaMemcpyDeviceToHost, st));
            DM(cudaStreamSynchronize(st));
            c = (long long)last_nl + 1;
        }
        (s ? out->consumed2 : out->consumed1) = c;
    }
    if (dev_io) {
        out->out1 = emit1 ? d->side[0].out.as<uint8_t>() : nullptr;
        out->out2 = emit2 ? d->side[1].out.as<uint8_t>() : nullptr;
        out->buckets = d->buckets.as<bdx_demux_bucket>();
        out->results = d->res.as<bdx_result>();
    } else {
        DM(d->h_buckets.reserve((size_t)std::max(out->n_buckets, 1) * sizeof(bdx_demux_bucket), st));
        DM(d->h_res.reserve((size_t)n * sizeof(bdx_result), st));
        DM(cudaMemcpyAsync(d->h_buckets.p, d->buckets.p, (size_t)out->n_buckets * sizeof(bdx_demux_bucket), cudaMemcpyDeviceToHost, st));
        DM(cudaMemcpyAsync(d->h_res.p, d->res.p, (size_t)n * sizeof(bdx_result), cudaMemcpyDeviceToHost, st));
        for (int s = 0; s < n_sides; s++) {
            if (!(s == 0 ? emit1 : emit2)) continue;
            DemuxSide &S = d->side[s];
            const long long ol = h_hdr->out_len[s];
            DM(S.h_out.reserve((size_t)ol + 16, st));
            if (ol) DM(cudaMemcpyAsync(S.h_out.p, S.out.p, (size_t)ol, cudaMemcpyDeviceToHost, st));
        }
        out->out1 = emit1 ? d->side[0].h_out.as<uint8_t>() : nullptr;
        out->out2 = emit2 ? d->side[1].h_out.as<uint8_t>() : nullptr;
        out->buckets = d->h_buckets.as<bdx_demux_bucket>();
        out->results = d->h_res.as<bdx_result>();
    }
    DM(cudaEventRecord(d->ev[8], st));
    DM(cudaStreamSynchronize(st));
    // stage times of this block: H2D | index+records | pack | classify | keys+sort | offsets+buckets | copy | D2H
    const int a[8] = {0, 1, 2, 3, 4, 5, 6, 7}, b[8] = {1, 2, 3, 4, 5, 6, 7, 8};
    for (int k = 0; k < 8; k++) {
        float ms = 0.f;
        if (cudaEventElapsedTime(&ms, d->ev[a[k]], d->ev[b[k]]) != cudaSuccess) {
            cudaGetLastError();
            ms = 0.f;
        }
        d->stage_ms[k] = ms;
    }
    return BDX_OK;
}

}  // namespace bdx
