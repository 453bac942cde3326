// fastq.cpp -- host-side FASTQ block scanner / packer.
// Keeps the record semantics of the reference's reader_task (src/core.jl:43-110): a record is
// four `readline`s; readline strips one trailing "\n" or "\r\n"; at end of input missing lines
// read as "".  Produces the packed (bytes, offsets) batch layout of bdx_submit directly, so a
// host can parse straight into the pinned staging returned by bdx_acquire.
#include <cstring>

#include "../../include/bdx.h"

namespace {

struct Line {
    int64_t off;
    int32_t len;
    int64_t next;   // offset just past the terminator
    bool complete;  // terminated by '\n'
};

inline Line next_line(const uint8_t *buf, int64_t pos, int64_t len)
{
    Line l;
    l.off = pos;
    const void *nl = pos < len ? memchr(buf + pos, '\n', (size_t)(len - pos)) : nullptr;
    if (nl) {
        const int64_t e = (const uint8_t *)nl - buf;
        l.next = e + 1;
        l.complete = true;
        int64_t end = e;
        if (end > pos && buf[end - 1] == '\r') end--;   // "\r\n"
        l.len = (int32_t)(end - pos);
    } else {
        l.next = len;
        l.complete = false;
        l.len = (int32_t)(len - pos);                    // last line without terminator: kept as is
    }
    return l;
}

}  // namespace

extern "C" int bdx_fastq_scan(const uint8_t *buf, int64_t len, int final_block, int32_t max_records,
                              bdx_fastq_record *recs, int32_t *n_records, int64_t *consumed)
{
    if ((!buf && len > 0) || len < 0 || max_records < 0 || !recs || !n_records || !consumed) return BDX_ERR_INVALID;
    int64_t pos = 0;
    int32_t n = 0;
    while (n < max_records && pos < len) {   // `while !eof(io)` (core.jl:85)
        Line l[4];
        int64_t p = pos;
        bool all_complete = true;
        for (int k = 0; k < 4; k++) {
            l[k] = next_line(buf, p, len);
            all_complete = all_complete && l[k].complete;
            p = l[k].next;
        }
        if (!all_complete && !final_block) break;   // the record continues in the next block
        bdx_fastq_record &r = recs[n++];
        r.header_off = l[0].off; r.header_len = l[0].len;
        r.seq_off = l[1].off;    r.seq_len = l[1].len;
        r.plus_off = l[2].off;   r.plus_len = l[2].len;
        r.qual_off = l[3].off;   r.qual_len = l[3].len;
        pos = p;
    }
    *n_records = n;
    *consumed = pos;
    return BDX_OK;
}

extern "C" int bdx_fastq_pack(const uint8_t *buf, const bdx_fastq_record *recs, int32_t n, uint8_t *seq_out,
                              int64_t seq_cap, int32_t *offsets_out)
{
    if (n < 0 || (n > 0 && (!buf || !recs)) || !seq_out || !offsets_out) return BDX_ERR_INVALID;
    int64_t o = 0;
    offsets_out[0] = 0;
    for (int32_t i = 0; i < n; i++) {
        const int32_t L = recs[i].seq_len;
        if (o + L > seq_cap || o + L > 0x7FFFFFF0ll) return BDX_ERR_TOO_LARGE;
        memcpy(seq_out + o, buf + recs[i].seq_off, (size_t)L);
        o += L;
        offsets_out[i + 1] = (int32_t)o;
    }
    return BDX_OK;
}
