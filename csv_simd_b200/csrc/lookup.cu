// lookup.cu -- K4: batched record / field lookups against a device-resident index.
//
// Replaces RecordSource::seek_record (src/record_source.rs:70-102) and
// RecordSource::seek_field (src/record_source.rs:104-140) for batches of
// independent queries.  One thread per query:
//   seek_field : None if r+1 >= record_cnt or f >= field_cnt;
//                s = (r+1)*row_size + f      (u32 arithmetic, as the reference);
//                start = index[s] + 1, end = index[s+1]
//   seek_record: None if r+1 >= record_cnt;
//                s = (r+1)*jump;  start = index[s] + 1, end = index[s + field_cnt]
// The (start, end) pair is written with one 16-byte store; None is
// (UINT64_MAX, UINT64_MAX).  A slot outside the index (where the reference
// would panic on the Vec bounds check) also yields None and bumps *oob.
// The access is a random 16-byte gather over a multi-GB index: HBM
// sector-bound, 40 algorithmic bytes per query (8 in, 16 gathered, 16 out).
#include "internal.h"

namespace csvb200 {

namespace {

__device__ __forceinline__ uint64_t ldg_u64(const uint64_t* p)
{
    uint64_t v;
    asm volatile("ld.global.nc.u64 %0, [%1];" : "=l"(v) : "l"(p));
    return v;
}

__global__ void __launch_bounds__(256) seek_kernel(const LookupParams p)
{
    const uint64_t stride = (uint64_t)gridDim.x * blockDim.x;
    for (uint64_t q = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; q < p.nq; q += stride) {
        const uint32_t r = p.rec[q];
        uint64_t start = UINT64_MAX, end = UINT64_MAX;
        bool live = (uint32_t)(r + 1u) < p.record_cnt;
        uint32_t s;
        uint64_t e_slot;
        if (p.fld != nullptr) {
            const uint32_t f = p.fld[q];
            live = live && f < p.field_cnt;
            s = (uint32_t)(r + 1u) * p.row_size + f;
            e_slot = (uint64_t)s + 1;
        } else {
            s = (uint32_t)(r + 1u) * p.row_size;
            e_slot = (uint64_t)s + p.field_cnt;
        }
        if (live) {
            if (e_slot < p.index_len) {
                start = ldg_u64(p.index + s) + 1;
                end = ldg_u64(p.index + e_slot);
            } else {
                atomicAdd(p.oob, 1u);
            }
        }
        asm volatile("st.global.v2.u64 [%0], {%1, %2};" ::"l"(p.ranges + 2 * q), "l"(start), "l"(end) : "memory");
    }
}

// One entry of an index whose segments live on different GPUs: a register-resident search over <= 16 bases, then a
// plain 8-byte load through the peer mapping (an NVLink read when the slot's owner is another GPU).
__device__ __forceinline__ uint64_t sharded_entry(const SegmentTable& seg, uint64_t s)
{
    uint32_t k = 0;
#pragma unroll
    for (uint32_t j = 1; j < kExMaxWorld; ++j)
        if (j < seg.nseg && s >= seg.base[j]) k = j;
    return ldg_u64(seg.ptr[k] + (s - seg.base[k]));
}

__global__ void __launch_bounds__(256) seek_sharded_kernel(const LookupParams p, const SegmentTable seg)
{
    const uint64_t stride = (uint64_t)gridDim.x * blockDim.x;
    for (uint64_t q = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; q < p.nq; q += stride) {
        const uint32_t r = p.rec[q];
        uint64_t start = UINT64_MAX, end = UINT64_MAX;
        bool live = (uint32_t)(r + 1u) < p.record_cnt;
        uint32_t s;
        uint64_t e_slot;
        if (p.fld != nullptr) {
            const uint32_t f = p.fld[q];
            live = live && f < p.field_cnt;
            s = (uint32_t)(r + 1u) * p.row_size + f;     // u32 arithmetic, as the reference (src/record_source.rs:129)
            e_slot = (uint64_t)s + 1;
        } else {
            s = (uint32_t)(r + 1u) * p.row_size;         // src/record_source.rs:83
            e_slot = (uint64_t)s + p.field_cnt;
        }
        if (live) {
            if (e_slot < p.index_len) {
                start = sharded_entry(seg, s) + 1;
                end = sharded_entry(seg, e_slot);
            } else {
                atomicAdd(p.oob, 1u);
            }
        }
        asm volatile("st.global.v2.u64 [%0], {%1, %2};" ::"l"(p.ranges + 2 * q), "l"(start), "l"(end) : "memory");
    }
}

__global__ void __launch_bounds__(256) range_lengths_kernel(const uint64_t* __restrict__ ranges, uint64_t nq,
                                                            uint64_t* __restrict__ lens)
{
    const uint64_t q = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (q >= nq) return;
    const uint64_t a = ranges[2 * q], b = ranges[2 * q + 1];
    lens[q] = (a == UINT64_MAX || b < a) ? 0 : b - a;
}

// eight lanes per range: fields are a few bytes long, so a whole warp per range left 3/4 of every request empty
// (ncu, round 1: 1.0 sectors per request, 0.15 of the HBM peak); four ranges per warp instruction keep four
// independent sectors in flight per request.  Positions are global: bytes[0] is global byte pos_bias, and a range
// that does not lie inside [pos_bias, pos_bias + n) (an index of another shard) is skipped, never dereferenced.
constexpr uint32_t kGatherLanes = 8;

__global__ void __launch_bounds__(256) gather_bytes_kernel(const uint8_t* __restrict__ bytes, uint64_t n, uint64_t pos_bias,
                                                           const uint64_t* __restrict__ ranges,
                                                           const uint64_t* __restrict__ out_off, uint64_t nq,
                                                           uint8_t* __restrict__ out)
{
    const uint32_t sub = threadIdx.x & (kGatherLanes - 1u);
    const uint64_t groups = ((uint64_t)gridDim.x * blockDim.x) / kGatherLanes;
    for (uint64_t q = ((uint64_t)blockIdx.x * blockDim.x + threadIdx.x) / kGatherLanes; q < nq; q += groups) {
        const uint64_t a = ranges[2 * q], b = ranges[2 * q + 1];
        if (a == UINT64_MAX || b <= a || a < pos_bias || b - pos_bias > n) continue;
        const uint8_t* src = bytes + (a - pos_bias);
        uint8_t* dst = out + out_off[q];
        for (uint64_t i = sub; i < b - a; i += kGatherLanes) dst[i] = src[i];
    }
}

}  // namespace

cudaError_t launch_seek(const LookupParams& p, cudaStream_t stream)
{
    if (p.nq == 0) return cudaSuccess;
    int dev = 0, sms = 148;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    uint64_t blocks = (p.nq + 255) / 256;
    const uint64_t max_blocks = (uint64_t)sms * 32;
    if (blocks > max_blocks) blocks = max_blocks;
    seek_kernel<<<(unsigned)blocks, 256, 0, stream>>>(p);
    return cudaGetLastError();
}

cudaError_t launch_seek_sharded(const LookupParams& p, const SegmentTable& seg, cudaStream_t stream)
{
    if (p.nq == 0) return cudaSuccess;
    int dev = 0, sms = 148;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    uint64_t blocks = (p.nq + 255) / 256;
    const uint64_t max_blocks = (uint64_t)sms * 32;
    if (blocks > max_blocks) blocks = max_blocks;
    seek_sharded_kernel<<<(unsigned)blocks, 256, 0, stream>>>(p, seg);
    return cudaGetLastError();
}

cudaError_t launch_range_lengths(const uint64_t* ranges, uint64_t nq, uint64_t* lens, cudaStream_t stream)
{
    if (nq == 0) return cudaSuccess;
    range_lengths_kernel<<<(unsigned)((nq + 255) / 256), 256, 0, stream>>>(ranges, nq, lens);
    return cudaGetLastError();
}

cudaError_t launch_gather_bytes(const uint8_t* bytes, uint64_t n, uint64_t pos_bias, const uint64_t* ranges,
                                const uint64_t* out_off, uint64_t nq, uint8_t* out, cudaStream_t stream)
{
    if (nq == 0) return cudaSuccess;
    int dev = 0, sms = 148;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    uint64_t blocks = (nq * kGatherLanes + 255) / 256;
    const uint64_t max_blocks = (uint64_t)sms * 16;
    if (blocks > max_blocks) blocks = max_blocks;
    gather_bytes_kernel<<<(unsigned)blocks, 256, 0, stream>>>(bytes, n, pos_bias, ranges, out_off, nq, out);
    return cudaGetLastError();
}

}  // namespace csvb200
