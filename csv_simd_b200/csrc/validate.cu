// validate.cu -- K7: ASCII / UTF-8 validation of the input.
//
// The reference carries two validators that are not wired into the path: is_ascii (word-at-a-time,
// src/reader.rs:26-132, "Non-core") and a copied SIMD UTF-8 checker whose module line is commented out
// (src/avx/utf8check.rs, src/avx/mod.rs:3), while seek_record builds &str with from_utf8_unchecked
// (src/record_source.rs:97-101).  This kernel answers both questions in one HBM-bound pass:
//   is_ascii      = no byte >= 0x80                         (is_ascii's result)
//   valid_up_to   = what core::str::from_utf8 would report: the start of the first ill-formed
//                   sequence (UINT64_MAX when the whole input is well-formed UTF-8)
// Every well-formedness rule of UTF-8 is local to a 4-byte window, so each thread judges the
// sequences that START in its own 16 bytes (3 bytes of look-ahead, 3 of look-behind for stray
// continuation bytes) and the answer is an atomicMin.  16-byte chunks that are pure ASCII (the common
// case in CSV) cost one 128-bit load and a mask test.
#include "internal.h"

namespace csvb200 {

namespace {

__device__ __forceinline__ uint4 ldg_128(const void* p)
{
    uint4 r;
    asm volatile("ld.global.nc.v4.u32 {%0, %1, %2, %3}, [%4];" : "=r"(r.x), "=r"(r.y), "=r"(r.z), "=r"(r.w) : "l"(p));
    return r;
}

__device__ __forceinline__ bool is_cont(uint32_t b) { return (b & 0xC0u) == 0x80u; }

// length of the sequence a lead byte announces (0 = not a valid lead: continuation, C0, C1, F5..FF)
__device__ __forceinline__ uint32_t lead_len(uint32_t b)
{
    if (b < 0x80u) return 1u;
    if (b >= 0xC2u && b <= 0xDFu) return 2u;
    if (b >= 0xE0u && b <= 0xEFu) return 3u;
    if (b >= 0xF0u && b <= 0xF4u) return 4u;
    return 0u;
}

__global__ void __launch_bounds__(256) utf8_validate_kernel(const uint8_t* __restrict__ in, uint64_t n,
                                                            uint64_t* __restrict__ result)   // {valid_up_to, non-ascii flag}
{
    const uint64_t nchunks = (n + 15) / 16;
    const uint64_t stride = (uint64_t)gridDim.x * blockDim.x;
    uint64_t bad = UINT64_MAX;
    uint32_t nonascii = 0u;
    for (uint64_t c = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; c < nchunks; c += stride) {
        const uint64_t i0 = 16 * c;
        uint32_t w[4];
        if (i0 + 16 <= n) {
            const uint4 v = ldg_128(in + i0);
            w[0] = v.x; w[1] = v.y; w[2] = v.z; w[3] = v.w;
        } else {
#pragma unroll
            for (int k = 0; k < 4; ++k) {
                uint32_t x = 0;
                for (int j = 0; j < 4; ++j)
                    if (i0 + 4 * k + j < n) x |= (uint32_t)in[i0 + 4 * k + j] << (8 * j);
                w[k] = x;
            }
        }
        if (((w[0] | w[1] | w[2] | w[3]) & 0x80808080u) == 0u) continue;   // pure ASCII chunk
        nonascii = 1u;
        // bytes i0-3 .. i0+18 as b[0 .. 21]
        uint32_t b[22];
#pragma unroll
        for (int j = 0; j < 3; ++j) b[j] = i0 + j >= 3 ? in[i0 + j - 3] : 0u;
#pragma unroll
        for (int j = 0; j < 16; ++j) b[3 + j] = (w[j >> 2] >> (8 * (j & 3))) & 0xffu;
#pragma unroll
        for (int j = 0; j < 3; ++j) b[19 + j] = i0 + 16 + j < n ? in[i0 + 16 + j] : 0x100u;   // 0x100 = past the end
        const uint32_t live = (uint32_t)(n - i0 < 16 ? n - i0 : 16);
#pragma unroll
        for (int j = 0; j < 16; ++j) {
            if ((uint32_t)j >= live) break;
            const uint32_t x = b[3 + j];
            if (x < 0x80u) continue;
            bool ok;
            if (is_cont(x)) {
                // a continuation byte is fine iff some lead within the 3 bytes before it reaches it
                const uint32_t p1 = b[2 + j], p2 = b[1 + j], p3 = b[j];
                ok = lead_len(p1) >= 2u || (is_cont(p1) && (lead_len(p2) >= 3u || (is_cont(p2) && lead_len(p3) == 4u)));
                // (a lead that reaches it but is itself ill-formed is reported at the lead: a smaller position)
            } else {
                const uint32_t len = lead_len(x);
                ok = len != 0u;
                if (ok) {
                    const uint32_t n1 = b[4 + j];
                    // second byte: general range 80..BF, narrowed for E0 (no overlongs), ED (no surrogates),
                    // F0 (no overlongs), F4 (<= U+10FFFF)
                    uint32_t lo = 0x80u, hi = 0xBFu;
                    if (x == 0xE0u) lo = 0xA0u;
                    if (x == 0xEDu) hi = 0x9Fu;
                    if (x == 0xF0u) lo = 0x90u;
                    if (x == 0xF4u) hi = 0x8Fu;
                    ok = n1 >= lo && n1 <= hi;
                    if (ok && len >= 3u) ok = b[5 + j] < 0x100u && is_cont(b[5 + j]);
                    if (ok && len == 4u) ok = b[6 + j] < 0x100u && is_cont(b[6 + j]);
                }
            }
            if (!ok) {
                const uint64_t pos = i0 + j;
                if (pos < bad) bad = pos;
                break;   // later positions of this chunk cannot beat it
            }
        }
    }
    bad = [&] {
        uint64_t v = bad;
#pragma unroll
        for (int d = 16; d > 0; d >>= 1) {
            const uint64_t o = __shfl_xor_sync(0xffffffffu, v, d);
            v = o < v ? o : v;
        }
        return v;
    }();
    const uint32_t any = __ballot_sync(0xffffffffu, nonascii != 0u);
    if ((threadIdx.x & 31u) == 0u) {
        if (bad != UINT64_MAX) atomicMin(reinterpret_cast<unsigned long long*>(result), (unsigned long long)bad);
        if (any) atomicOr(reinterpret_cast<unsigned long long*>(result + 1), 1ull);
    }
}

}  // namespace

cudaError_t launch_utf8_validate(const uint8_t* in, uint64_t n, uint64_t* result, cudaStream_t stream)
{
    if (n == 0) return cudaSuccess;
    int dev = 0, sms = 148;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    uint64_t blocks = ((n + 15) / 16 + 255) / 256;
    const uint64_t max_blocks = (uint64_t)sms * 8;
    if (blocks > max_blocks) blocks = max_blocks;
    utf8_validate_kernel<<<(unsigned)blocks, 256, 0, stream>>>(in, n, result);
    return cudaGetLastError();
}

}  // namespace csvb200
