// demux.h -- internal interface of the device FASTQ block demultiplexer (demux.cu)
#pragma once

#include <functional>
#include <string>

#include "bdx_internal.h"

namespace bdx {

struct DemuxState;
// enqueues the classification kernels for n packed reads on the stream the demultiplexer runs on
using DemuxClassifyFn = std::function<int(const uint8_t *d_seq, const int *d_off, int n, bdx_result *d_res)>;

DemuxState *demux_state_create();
void demux_state_destroy(DemuxState *d);
const float *demux_stage_ms(const DemuxState *d);
int demux_run(DemuxState *d, const DevParams &P, cudaStream_t st, const DemuxClassifyFn &classify,
              const uint8_t *fq1, int64_t len1, const uint8_t *fq2, int64_t len2, int final_block, int mode,
              int64_t *launches, bdx_demux_out *out, std::string &err);

}  // namespace bdx
