// packed.cu -- 4-bit packed read input (include/bdx.h, bdx_submit_packed4): the device side.
//
// The host link is what bounds the end-to-end path (154 B per 150-base read over PCIe), and all the path ever
// asks of a read byte is whether it EQUALS a barcode byte.  So a read can travel as codes: 0 = "a byte that
// occurs in no barcode of the config", 1..15 = the distinct barcode bytes -- two codes per byte, byte k of the
// batch's concatenated reads in nibble k of the packed stream (low nibble first), the offsets unchanged.  This
// kernel expands the codes to REPRESENTATIVE bytes (the barcode byte itself; for code 0 a byte outside every
// barcode) into the slot's ordinary sequence buffer, and the classification kernels run on it unchanged:
// every comparison gives what it would give on the original bytes.  HBM-bound: n / 2 bytes in, n bytes out.
#include "bdx_internal.h"

namespace bdx {

struct RepTable {
    uint8_t rep[16];        // code -> representative byte
};

__global__ void __launch_bounds__(256)
k_unpack4(const uint8_t *__restrict__ packed, uint8_t *__restrict__ seq, const int *__restrict__ off, const int n_reads,
          const RepTable T)
{
    __shared__ uint16_t pair_s[256];                 // packed byte -> its two representative bytes
    for (int k = threadIdx.x; k < 256; k += blockDim.x)
        pair_s[k] = (uint16_t)(T.rep[k & 15] | (T.rep[k >> 4] << 8));
    __syncthreads();
    const long long total = off[n_reads];            // bytes of the batch
    const long long n_vec = (total + 15) >> 4;       // 16 output bytes per thread and round: 8 packed bytes in
    for (long long v = (long long)blockIdx.x * blockDim.x + threadIdx.x; v < n_vec; v += (long long)gridDim.x * blockDim.x) {
        const uint2 in = __ldg(reinterpret_cast<const uint2 *>(packed) + v);
        uint4 o;
        o.x = pair_s[in.x & 0xFF] | ((uint32_t)pair_s[(in.x >> 8) & 0xFF] << 16);
        o.y = pair_s[(in.x >> 16) & 0xFF] | ((uint32_t)pair_s[in.x >> 24] << 16);
        o.z = pair_s[in.y & 0xFF] | ((uint32_t)pair_s[(in.y >> 8) & 0xFF] << 16);
        o.w = pair_s[(in.y >> 16) & 0xFF] | ((uint32_t)pair_s[in.y >> 24] << 16);
        reinterpret_cast<uint4 *>(seq)[v] = o;       // the buffers are sized to a multiple of 16 bytes
    }
}

cudaError_t launch_unpack4(const uint8_t *d_packed, uint8_t *d_seq, const int *d_off, int n_reads, long long max_bytes,
                           const uint8_t rep[16], int sm_count, cudaStream_t st)
{
    if (n_reads <= 0) return cudaSuccess;
    RepTable T;
    for (int k = 0; k < 16; k++) T.rep[k] = rep[k];
    const long long vecs = (max_bytes + 15) >> 4;
    const int blocks = (int)std::max<long long>(1, std::min<long long>((vecs + 255) / 256, (long long)sm_count * 8));
    k_unpack4<<<blocks, 256, 0, st>>>(d_packed, d_seq, d_off, n_reads, T);
    return cudaGetLastError();
}

}  // namespace bdx
