// materialize.cu -- K6: column materialisation.
//
// RecordSource::seek_field (src/record_source.rs:104-140) hands back the RAW slice of a field,
// "incl. surrounding quotes / padding; no unescape, no trim" (:135-139).  What callers do next is turn
// a column of such slices into values; this file does that on the device for a whole column at once:
//   for r in [first_record, first_record + nrec):   (start, end) = seek_field(r, field_idx)
//       value = raw slice, optionally trimmed of ASCII space / tab, optionally RFC-4180 unquoted
//               (outer quotes stripped when both are present, "" -> ")
//   offsets[r - first_record] = exclusive prefix sum of the value lengths, out = values back to back
// Two kernels: (1) lengths + exclusive scan fused (one pass, decoupled look-back over 62-bit sums),
// (2) the write.  Both read two index entries per record and the field's bytes: HBM sector-bound.
#include "internal.h"

namespace csvb200 {

namespace {

constexpr int kMatThreads = 256;
constexpr int kMatItems = 4;
constexpr int kMatTile = kMatThreads * kMatItems;
constexpr uint64_t kMatAgg = 1ull << 62, kMatPrefix = 2ull << 62, kMatMask = (1ull << 62) - 1;

__device__ __forceinline__ uint64_t ldg_u64(const uint64_t* p)
{
    uint64_t v;
    asm volatile("ld.global.nc.u64 %0, [%1];" : "=l"(v) : "l"(p));
    return v;
}
__device__ __forceinline__ uint64_t ld_relaxed(const uint64_t* p)
{
    uint64_t v;
    asm volatile("ld.relaxed.gpu.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ void st_relaxed(uint64_t* p, uint64_t v)
{
    asm volatile("st.relaxed.gpu.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}

// what the per-value helpers need of either parameter block
struct FieldView {
    const uint64_t* index;
    uint64_t index_len;
    const uint8_t* bytes;
    uint64_t n, pos_bias;
    uint32_t record_cnt, field_cnt, row_size, field_idx, flags;
};
__device__ __forceinline__ FieldView view_of(const MaterializeParams& p)
{
    return FieldView{p.index, p.index_len, p.bytes, p.n, p.pos_bias, p.record_cnt, p.field_cnt, p.row_size, p.field_idx, p.flags};
}
__device__ __forceinline__ FieldView view_of(const MaterializeMultiParams& p, uint32_t c)
{
    return FieldView{p.index, p.index_len, p.bytes, p.n, p.pos_bias, p.record_cnt, p.field_cnt, p.row_size, p.field_idx[c], p.flags};
}

// seek_field (src/record_source.rs:104-140) as byte offsets relative to p.bytes; false = Ok(None) or a
// slot the reference would panic on
__device__ __forceinline__ bool field_range(const FieldView& p, uint32_t r, uint64_t& a, uint64_t& b)
{
    if ((uint32_t)(r + 1u) >= p.record_cnt || p.field_idx >= p.field_cnt) return false;
    const uint32_t s = (uint32_t)(r + 1u) * p.row_size + p.field_idx;   // u32 arithmetic, as the reference
    if ((uint64_t)s + 1 >= p.index_len) return false;
    a = ldg_u64(p.index + s) + 1 - p.pos_bias;
    b = ldg_u64(p.index + s + 1) - p.pos_bias;
    return a <= b && b <= p.n;
}

// trims [a, b) in place and reports whether the value is a quoted field to unquote
__device__ __forceinline__ bool trim_and_test(const FieldView& p, uint64_t& a, uint64_t& b)
{
    const uint8_t* x = p.bytes;
    if (p.flags & 2u) {
        while (a < b && (x[a] == 0x20u || x[a] == 0x09u)) ++a;
        while (b > a && (x[b - 1] == 0x20u || x[b - 1] == 0x09u)) --b;
    }
    if ((p.flags & 1u) && b - a >= 2 && x[a] == 0x22u && x[b - 1] == 0x22u) {
        ++a;
        --b;
        return true;
    }
    return false;
}

__device__ __forceinline__ uint4 ldg_128(const void* p)
{
    uint4 r;
    asm volatile("ld.global.nc.v4.u32 {%0, %1, %2, %3}, [%4];" : "=r"(r.x), "=r"(r.y), "=r"(r.z), "=r"(r.w) : "l"(p));
    return r;
}

// Calls f(byte) for every byte of x[a, b) in order, fetching the range as aligned 16-byte words: two or three
// 128-bit loads per value instead of one byte load per character (with 64 warps per SM each walking a different
// 128-byte line, byte-at-a-time loads thrash L1 and every character becomes an L2 round trip).  x must be 16-byte
// aligned (the ABI requires it of device inputs); a chunk that would read past n is fetched bytewise.
template <class F>
__device__ __forceinline__ void scan_bytes(const uint8_t* __restrict__ x, uint64_t n, uint64_t a, uint64_t b, F&& f)
{
    for (uint64_t base = a & ~15ull; base < b; base += 16) {
        uint32_t w[4];
        if (base + 16 <= n) {
            const uint4 v = ldg_128(x + base);
            w[0] = v.x; w[1] = v.y; w[2] = v.z; w[3] = v.w;
        } else {
#pragma unroll
            for (int k = 0; k < 4; ++k) {
                uint32_t t = 0;
                for (int j = 0; j < 4; ++j)
                    if (base + 4 * k + j < n) t |= (uint32_t)x[base + 4 * k + j] << (8 * j);
                w[k] = t;
            }
        }
        const uint32_t lo = a > base ? (uint32_t)(a - base) : 0u;
        const uint32_t hi = b - base < 16 ? (uint32_t)(b - base) : 16u;
#pragma unroll
        for (int k = 0; k < 16; ++k)
            if ((uint32_t)k >= lo && (uint32_t)k < hi) f((w[k >> 2] >> (8 * (k & 3))) & 0xffu);
    }
}

// RFC-4180 unquoting as a streaming rule: a quote is held back one byte; a second quote right after it turns the
// pair into one '"', anything else releases the held quote as it is (same result as the scalar definition: '"'
// followed by '"' emits one and skips both, a lone '"' stays).
template <class Emit>
struct Unquoter {
    Emit emit;
    bool held = false;
    __device__ __forceinline__ void operator()(uint32_t c)
    {
        if (held) {
            held = false;
            emit(0x22u);
            if (c == 0x22u) return;
        }
        if (c == 0x22u)
            held = true;
        else
            emit(c);
    }
    __device__ __forceinline__ void finish()
    {
        if (held) emit(0x22u);
        held = false;
    }
};

__device__ __forceinline__ uint64_t value_len(const FieldView& p, uint32_t r)
{
    uint64_t a, b;
    if (!field_range(p, r, a, b)) return 0;
    if (!trim_and_test(p, a, b)) return b - a;
    uint64_t o = 0;
    auto count = [&](uint32_t) { ++o; };
    Unquoter<decltype(count)> u{count};
    scan_bytes(p.bytes, p.n, a, b, u);
    u.finish();
    return o;
}

__global__ void __launch_bounds__(kMatThreads) materialize_offsets_kernel(const MaterializeParams p)
{
    __shared__ uint64_t s_warp[kMatThreads / 32];
    __shared__ uint64_t s_prefix;
    __shared__ uint32_t s_tile;
    const uint32_t tid = threadIdx.x, lane = tid & 31u, warp = tid >> 5;
    if (tid == 0) s_tile = atomicAdd(p.ticket, 1u);   // tiles are taken in order: look-back only waits on started tiles
    __syncthreads();
    const uint32_t tile = s_tile;
    const uint64_t i0 = (uint64_t)tile * kMatTile + (uint64_t)tid * kMatItems;
    uint64_t len[kMatItems], tsum = 0;
#pragma unroll
    for (int k = 0; k < kMatItems; ++k) {
        len[k] = i0 + k < p.nrec ? value_len(view_of(p), p.first_record + (uint32_t)(i0 + k)) : 0;
        tsum += len[k];
    }
    // block exclusive scan of the thread sums
    uint64_t inc = tsum;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
        const uint64_t v = __shfl_up_sync(0xffffffffu, inc, d);
        if (lane >= (uint32_t)d) inc += v;
    }
    if (lane == 31) s_warp[warp] = inc;
    __syncthreads();
    uint64_t wbase = 0, total = 0;
#pragma unroll
    for (int w = 0; w < kMatThreads / 32; ++w) {
        if (w < (int)warp) wbase += s_warp[w];
        total += s_warp[w];
    }
    // decoupled look-back over the tile totals (status in the top two bits)
    if (tid == 0) {
        st_relaxed(p.tile_desc + tile, kMatAgg | (total & kMatMask));
        uint64_t prefix = 0;
        for (int64_t t = (int64_t)tile - 1; t >= 0;) {
            const uint64_t d = ld_relaxed(p.tile_desc + t);
            const uint64_t st = d >> 62;
            if (st == 0) {
                __nanosleep(40);
                continue;
            }
            prefix += d & kMatMask;
            if (st == 2) break;
            --t;
        }
        st_relaxed(p.tile_desc + tile, kMatPrefix | ((prefix + total) & kMatMask));
        s_prefix = prefix;
    }
    __syncthreads();
    uint64_t off = s_prefix + wbase + (inc - tsum);
#pragma unroll
    for (int k = 0; k < kMatItems; ++k) {
        if (i0 + k < p.nrec) p.offsets[i0 + k] = off;
        off += len[k];
    }
    if (i0 < p.nrec && i0 + kMatItems >= p.nrec) p.offsets[p.nrec] = off;   // the thread owning the last record
}

__global__ void __launch_bounds__(kMatThreads) materialize_write_kernel(const MaterializeParams p)
{
    const uint64_t stride = (uint64_t)gridDim.x * blockDim.x;
    // restrict-qualified local copies: without them every byte load has to wait for the previous byte store
    // (the compiler must assume out aliases bytes), which serialises the copy into load -> store round trips
    const uint8_t* __restrict__ x = p.bytes;
    uint8_t* __restrict__ out = p.out;
    const FieldView fv = view_of(p);
    for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < p.nrec; i += stride) {
        uint64_t a, b;
        if (!field_range(fv, p.first_record + (uint32_t)i, a, b)) continue;
        uint64_t o = p.offsets[i];
        const uint64_t o_end = p.offsets[i + 1];
        if (o_end > p.out_cap) continue;   // host form checks the capacity first; device form clips whole values
        auto put = [&](uint32_t c) { out[o++] = (uint8_t)c; };
        if (!trim_and_test(fv, a, b)) {
            if (b - a < 16) {   // short raw values: a handful of byte loads beat 16 predicated steps per chunk
                for (; a < b; ++a) put(x[a]);
            } else {
                scan_bytes(x, p.n, a, b, put);
            }
        } else {
            Unquoter<decltype(put)> u{put};
            scan_bytes(x, p.n, a, b, u);
            u.finish();
        }
    }
}

// ---- several columns in one sweep --------------------------------------------------------------------------------
// A tile is 256 consecutive records, one per thread; the thread walks the requested fields of ITS row, so the index
// slots and bytes of neighbouring columns come out of the sectors the first column fetched.  Per column: a block scan
// of the 256 lengths (warp w scans columns w, w + 8, ...), one decoupled look-back chain over the tile totals
// (thread c looks back for column c), then the exclusive offsets go out coalesced.
constexpr int kMultiThreads = 256;

__global__ void __launch_bounds__(kMultiThreads) materialize_multi_offsets_kernel(const MaterializeMultiParams p)
{
    extern __shared__ uint32_t s_len[];                    // [ncols][256] lengths, then exclusive prefixes in place
    __shared__ uint64_t s_total[kMatMaxCols], s_base[kMatMaxCols];
    __shared__ uint32_t s_tile;
    const uint32_t tid = threadIdx.x, lane = tid & 31u, warp = tid >> 5;
    if (tid == 0) s_tile = atomicAdd(p.ticket, 1u);        // tiles are taken in order: look-back only waits on started tiles
    __syncthreads();
    const uint32_t tile = s_tile;
    const uint64_t i = (uint64_t)tile * kMultiThreads + tid;
    for (uint32_t c = 0; c < p.ncols; ++c)
        s_len[c * kMultiThreads + tid] = i < p.nrec ? (uint32_t)value_len(view_of(p, c), p.first_record + (uint32_t)i) : 0u;
    __syncthreads();
    for (uint32_t c = warp; c < p.ncols; c += kMultiThreads / 32) {
        uint32_t* col = s_len + c * kMultiThreads;
        uint32_t v[8], sum = 0;
#pragma unroll
        for (int k = 0; k < 8; ++k) {
            v[k] = col[8 * lane + k];
            sum += v[k];
        }
        uint32_t inc = sum;
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) {
            const uint32_t t = __shfl_up_sync(0xffffffffu, inc, d);
            if (lane >= (uint32_t)d) inc += t;
        }
        uint32_t run = inc - sum;
#pragma unroll
        for (int k = 0; k < 8; ++k) {
            col[8 * lane + k] = run;
            run += v[k];
        }
        if (lane == 31) s_total[c] = inc;
    }
    __syncthreads();
    if (tid < p.ncols) {
        uint64_t* desc = p.tile_desc + (uint64_t)tid * p.tiles;
        const uint64_t total = s_total[tid];
        st_relaxed(desc + tile, kMatAgg | (total & kMatMask));
        uint64_t prefix = 0;
        for (int64_t t = (int64_t)tile - 1; t >= 0;) {
            const uint64_t d = ld_relaxed(desc + t);
            const uint64_t st = d >> 62;
            if (st == 0) {
                __nanosleep(40);
                continue;
            }
            prefix += d & kMatMask;
            if (st == 2) break;
            --t;
        }
        st_relaxed(desc + tile, kMatPrefix | ((prefix + total) & kMatMask));
        s_base[tid] = prefix;
    }
    __syncthreads();
    for (uint32_t c = 0; c < p.ncols; ++c) {
        const uint64_t off = s_base[c] + s_len[c * kMultiThreads + tid];
        if (i < p.nrec) p.offsets[c][i] = off;
        if (i + 1 == p.nrec) p.offsets[c][p.nrec] = s_base[c] + s_total[c];   // the thread owning the last record
    }
}

__global__ void __launch_bounds__(kMultiThreads) materialize_multi_write_kernel(const MaterializeMultiParams p)
{
    const uint8_t* __restrict__ x = p.bytes;
    const uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= p.nrec) return;
    for (uint32_t c = 0; c < p.ncols; ++c) {
        const FieldView fv = view_of(p, c);
        uint8_t* __restrict__ out = p.out[c];
        uint64_t a, b;
        if (!field_range(fv, p.first_record + (uint32_t)i, a, b)) continue;
        uint64_t o = p.offsets[c][i];
        if (p.offsets[c][i + 1] > p.out_cap[c]) continue;
        auto put = [&](uint32_t ch) { out[o++] = (uint8_t)ch; };
        if (!trim_and_test(fv, a, b)) {
            if (b - a < 16) {
                for (; a < b; ++a) put(x[a]);
            } else {
                scan_bytes(x, p.n, a, b, put);
            }
        } else {
            Unquoter<decltype(put)> u{put};
            scan_bytes(x, p.n, a, b, u);
            u.finish();
        }
    }
}

}  // namespace

size_t materialize_multi_scratch_bytes(uint32_t nrec, uint32_t ncols)
{
    const size_t tiles = ((size_t)nrec + kMultiThreads - 1) / kMultiThreads + 1;
    return 128 + tiles * ncols * sizeof(uint64_t);
}

cudaError_t launch_materialize_multi_offsets(const MaterializeMultiParams& p, cudaStream_t stream)
{
    if (p.tiles == 0 || p.ncols == 0) return cudaSuccess;
    const size_t smem = (size_t)p.ncols * kMultiThreads * sizeof(uint32_t);
    materialize_multi_offsets_kernel<<<p.tiles, kMultiThreads, smem, stream>>>(p);
    return cudaGetLastError();
}

cudaError_t launch_materialize_multi_write(const MaterializeMultiParams& p, cudaStream_t stream)
{
    if (p.nrec == 0 || p.ncols == 0) return cudaSuccess;
    materialize_multi_write_kernel<<<(p.nrec + kMultiThreads - 1) / kMultiThreads, kMultiThreads, 0, stream>>>(p);
    return cudaGetLastError();
}

size_t materialize_scratch_bytes(uint32_t nrec) { return 128 + ((size_t)(nrec + kMatTile - 1) / kMatTile + 1) * sizeof(uint64_t); }

cudaError_t launch_materialize_offsets(const MaterializeParams& p, cudaStream_t stream)
{
    const uint32_t tiles = (p.nrec + kMatTile - 1) / kMatTile;
    if (tiles == 0) return cudaSuccess;
    materialize_offsets_kernel<<<tiles, kMatThreads, 0, stream>>>(p);
    return cudaGetLastError();
}

cudaError_t launch_materialize_write(const MaterializeParams& p, cudaStream_t stream)
{
    if (p.nrec == 0) return cudaSuccess;
    int dev = 0, sms = 148;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    uint64_t blocks = ((uint64_t)p.nrec + kMatThreads - 1) / kMatThreads;
    const uint64_t max_blocks = (uint64_t)sms * 16;
    if (blocks > max_blocks) blocks = max_blocks;
    materialize_write_kernel<<<(unsigned)blocks, kMatThreads, 0, stream>>>(p);
    return cudaGetLastError();
}

}  // namespace csvb200
