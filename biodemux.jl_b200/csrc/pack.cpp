// pack.cpp -- host side of the 4-bit packed read input (include/bdx.h): bytes -> codes, two per byte.
// What a reader does instead of copying sequence bytes into the staging buffer: it touches every byte once
// either way.  AVX2 path for configs whose barcode bytes are all in 0x40..0x5F (upper-case letters, the usual
// case): two 16-entry pshufb tables over the low five bits; scalar 256-entry table otherwise.
#include <cstdint>
#include <cstring>

#if defined(__x86_64__)
#include <immintrin.h>
#endif

#include "../../include/bdx.h"

namespace {

void pack_scalar(const uint8_t *code, const uint8_t *in, int64_t n, uint8_t *out)
{
    int64_t k = 0;
    for (; k + 1 < n; k += 2) out[k >> 1] = (uint8_t)(code[in[k]] | (code[in[k + 1]] << 4));
    if (k < n) out[k >> 1] = code[in[k]];
}

#if defined(__x86_64__)
__attribute__((target("avx2"))) void pack_avx2(const uint8_t *code, const uint8_t *in, int64_t n, uint8_t *out)
{
    // bytes 0x40..0x4F -> table lo, 0x50..0x5F -> table hi, anything else -> code 0
    alignas(32) uint8_t lo[32], hi[32];
    for (int k = 0; k < 16; k++) {
        lo[k] = lo[k + 16] = code[0x40 + k];
        hi[k] = hi[k + 16] = code[0x50 + k];
    }
    const __m256i tlo = _mm256_load_si256((const __m256i *)lo), thi = _mm256_load_si256((const __m256i *)hi);
    const __m256i m0f = _mm256_set1_epi8(0x0F), me0 = _mm256_set1_epi8((char)0xE0), c40 = _mm256_set1_epi8(0x40);
    const __m256i m10 = _mm256_set1_epi8(0x10);
    int64_t k = 0;
    for (; k + 64 <= n; k += 64) {
        __m256i v[2];
        for (int h = 0; h < 2; h++) {
            const __m256i b = _mm256_loadu_si256((const __m256i *)(in + k + 32 * h));
            const __m256i idx = _mm256_and_si256(b, m0f);
            const __m256i a = _mm256_shuffle_epi8(tlo, idx), c = _mm256_shuffle_epi8(thi, idx);
            const __m256i is_hi = _mm256_cmpeq_epi8(_mm256_and_si256(b, m10), m10);
            const __m256i valid = _mm256_cmpeq_epi8(_mm256_and_si256(b, me0), c40);
            v[h] = _mm256_and_si256(_mm256_blendv_epi8(a, c, is_hi), valid);
        }
        // codes c0 c1 c2 ... (one per byte) -> c0 | c1 << 4, c2 | c3 << 4, ...: maddubs with (1, 16) pairs gives
        // 16-bit sums, packus narrows them (values <= 255); packus works per 128-bit lane, so fix the lane order
        const __m256i w = _mm256_set1_epi16(0x1001);
        const __m256i s0 = _mm256_maddubs_epi16(v[0], w), s1 = _mm256_maddubs_epi16(v[1], w);
        const __m256i p = _mm256_permute4x64_epi64(_mm256_packus_epi16(s0, s1), 0xD8);
        _mm256_storeu_si256((__m256i *)(out + (k >> 1)), p);
    }
    pack_scalar(code, in + k, n - k, out + (k >> 1));
}
#endif

}  // namespace

// shared with bdx_api.cu (the stream's own staging copy uses the same routine)
void bdx_pack4_bytes(const uint8_t code[256], const uint8_t *in, int64_t n, uint8_t *out)
{
#if defined(__x86_64__)
    bool letters = true;                       // every non-zero code sits in 0x40..0x5F?
    for (int b = 0; b < 256; b++)
        if (code[b] && (b & 0xE0) != 0x40) letters = false;
    if (letters && n >= 64 && __builtin_cpu_supports("avx2")) {
        pack_avx2(code, in, n, out);
        return;
    }
#endif
    pack_scalar(code, in, n, out);
}
