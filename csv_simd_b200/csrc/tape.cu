// tape.cu -- K5: device-side Tape validation.
//
// The reference accepts a file when (index.len() - 1) % jump == 0 (TapeCore::init, src/tape.rs:315-347)
// and otherwise returns InvalidCsvFormat without saying where the structure breaks; a file whose rows
// are ragged in a way that happens to cancel passes silently and every later seek_record / seek_field
// (src/record_source.rs:70-140) returns the wrong bytes.  This kernel checks what those seeks assume:
// entry s (s >= 1) of the index is the k-th separator of its record, k = (s - 1) % jump, so it must be
//   LF files   (jump = field_cnt)     : ',' for k < jump-1, CR or LF for k = jump-1
//   CRLF files (jump = field_cnt + 1) : ',' for k < jump-2, CR for k = jump-2, LF right after it for k = jump-1
// and reports the FIRST slot that is not (atomicMin), i.e. the first bad record.  HBM-bound: reads the
// index once (8E bytes) and one byte per entry of the input (every 32-byte sector of a dense file: N).
#include "internal.h"

namespace csvb200 {

namespace {

__device__ __forceinline__ uint64_t ldg_u64(const uint64_t* p)
{
    uint64_t v;
    asm volatile("ld.global.nc.u64 %0, [%1];" : "=l"(v) : "l"(p));
    return v;
}

// One row of 32 slots per warp step.  Everything that does not depend on the slot is hoisted: CRLF / LF is a
// template parameter, k = (s - 1) % jump lives in 32 bits and advances by 32 % jump, and the only 64-bit
// work left per slot is the position arithmetic (the first version spent ~100 instructions per slot row on
// 64-bit modulo / select chains and was issue-bound at 0.55 of the HBM peak).
template <bool kCrlf>
__global__ void __launch_bounds__(256) tape_validate_kernel(const TapeValidateParams p)
{
    const uint32_t lane = threadIdx.x & 31u;
    const uint64_t warps = ((uint64_t)gridDim.x * blockDim.x) >> 5;
    const uint64_t warp0 = ((uint64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const uint32_t jump = (uint32_t)p.jump;
    const uint32_t nl0 = jump - (kCrlf ? 2u : 1u);    // first line-end slot of a row (k >= nl0 is CR / LF territory)
    const uint64_t last = p.index_len - 1;            // slots are 1 .. last; i = s - 1 in [0, last)
    const uint32_t step = 32u % jump;                 // k advances by 32 slots per warp row
    const uint64_t* __restrict__ index = p.index + 1; // index[i] here is slot s = i + 1
    uint64_t bad = UINT64_MAX;                        // smallest failing i
    constexpr int kUnroll = 4;
    constexpr uint64_t kRun = 32 * 16;
    for (uint64_t run = warp0 * kRun; run < last; run += warps * kRun) {
        const uint64_t run_end = run + kRun < last ? run + kRun : last;
        uint32_t k = (uint32_t)((run + lane) % jump);
        for (uint64_t i = run + lane; i < run_end; i += 32 * kUnroll) {
            uint64_t pos[kUnroll];
            uint32_t byte[kUnroll], kk[kUnroll];
#pragma unroll
            for (int u = 0; u < kUnroll; ++u) {
                const uint64_t iu = i + 32ull * u;
                pos[u] = iu < run_end ? ldg_u64(index + iu) : UINT64_MAX;   // UINT64_MAX - bias >= n: treated as skipped below
                kk[u] = k;
                k += step;
                k = k >= jump ? k - jump : k;
            }
#pragma unroll
            for (int u = 0; u < kUnroll; ++u) {
                const uint64_t rel = pos[u] - p.pos_bias;
                byte[u] = rel < p.n ? (uint32_t)p.bytes[rel] : 0x100u;      // 0x100: no separator there (out of range)
            }
#pragma unroll
            for (int u = 0; u < kUnroll; ++u) {
                const uint64_t iu = i + 32ull * u;
                if (iu >= run_end) continue;
                const uint32_t b = byte[u];
                bool ok;
                if (kk[u] < nl0) {
                    ok = b == 0x2Cu;
                } else if (!kCrlf) {
                    ok = b == 0x0Au || b == 0x0Du;
                } else if (kk[u] == nl0) {
                    ok = b == 0x0Du;
                } else {
                    ok = b == 0x0Au && ldg_u64(index + iu - 1) + 1 == pos[u];   // the LF directly after its CR
                }
                if (!ok && iu < bad) bad = iu;
            }
        }
    }
    // warp minimum, one atomic per warp that saw a violation
#pragma unroll
    for (int d = 16; d > 0; d >>= 1) {
        const uint64_t o = __shfl_xor_sync(0xffffffffu, bad, d);
        bad = o < bad ? o : bad;
    }
    if (lane == 0 && bad != UINT64_MAX)
        atomicMin(reinterpret_cast<unsigned long long*>(p.first_bad_slot), (unsigned long long)(bad + 1));
}

__global__ void gather_slots_kernel(const uint64_t* __restrict__ index, uint64_t index_len,
                                    const uint64_t* __restrict__ slots, uint64_t n, uint64_t* __restrict__ out)
{
    const uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) out[i] = slots[i] < index_len ? index[slots[i]] : UINT64_MAX;
}

}  // namespace

cudaError_t launch_gather_slots(const uint64_t* index, uint64_t index_len, const uint64_t* slots, uint64_t n,
                                uint64_t* out, cudaStream_t stream)
{
    if (n == 0) return cudaSuccess;
    gather_slots_kernel<<<(unsigned)((n + 127) / 128), 128, 0, stream>>>(index, index_len, slots, n, out);
    return cudaGetLastError();
}

cudaError_t launch_tape_validate(const TapeValidateParams& p, cudaStream_t stream)
{
    if (p.index_len <= 1 || p.jump == 0) return cudaSuccess;
    int dev = 0, sms = 148;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    uint64_t blocks = (p.index_len + 32 * 16 * 8 - 1) / (32 * 16 * 8);   // 8 warps per block, one run each
    const uint64_t max_blocks = (uint64_t)sms * 8;
    if (blocks > max_blocks) blocks = max_blocks;
    if (blocks == 0) blocks = 1;
    if (p.jump > 0xffffffffull) return cudaErrorInvalidValue;
    if (p.crlf)
        tape_validate_kernel<true><<<(unsigned)blocks, 256, 0, stream>>>(p);
    else
        tape_validate_kernel<false><<<(unsigned)blocks, 256, 0, stream>>>(p);
    return cudaGetLastError();
}

}  // namespace csvb200
