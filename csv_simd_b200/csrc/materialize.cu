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

#include <cstdlib>
#include <mutex>

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

// Decoupled look-back by a whole warp: publishes this tile's total, sums the totals of the tiles before it back to the
// nearest published prefix, publishes prefix + total and returns the prefix (every lane).  Tiles are small, so tiles
// start faster than one 32-descriptor hop (~0.7 us of L2 latency) could follow: with one window per hop the nearest
// published prefix drifts away until every look-back walks all the tiles in flight.  Four windows (128 tiles) are
// fetched at once per hop, the loads overlapped.  Tiles must be taken in ticket order.
__device__ __forceinline__ uint64_t lookback_warp(uint64_t* desc, uint32_t tile, uint64_t total, uint32_t lane)
{
    if (lane == 0) st_relaxed(desc + tile, kMatAgg | (total & kMatMask));
    uint64_t prefix = 0;
    for (int64_t t = (int64_t)tile - 1; t >= 0;) {
        uint64_t d[4];
#pragma unroll
        for (int w = 0; w < 4; ++w) {
            const int64_t mine = t - 32 * w - (int64_t)lane;
            d[w] = mine >= 0 ? ld_relaxed(desc + mine) : kMatPrefix;          // before tile 0: prefix 0
        }
        bool done = false, retry = false;
#pragma unroll
        for (int w = 0; w < 4; ++w) {
            if (done || retry) break;
            const uint32_t st = (uint32_t)(d[w] >> 62);
            const uint32_t pref = __ballot_sync(0xffffffffu, st == 2u), wait = __ballot_sync(0xffffffffu, st == 0u);
            const uint32_t upto = pref ? (uint32_t)__ffs((int)pref) - 1u : 31u;  // lanes 0 .. upto count
            const uint32_t use = 0xffffffffu >> (31u - upto);
            if (wait & use) {
                retry = true;                                                 // resume from this window
                break;
            }
            uint64_t v = (use >> lane & 1u) ? (d[w] & kMatMask) : 0ull;
#pragma unroll
            for (int o = 16; o; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
            prefix += v;
            t -= 32;
            done = pref != 0u;
        }
        if (done) break;
        if (retry) __nanosleep(40);
    }
    if (lane == 0) st_relaxed(desc + tile, kMatPrefix | ((prefix + total) & kMatMask));
    return prefix;
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

// length of the value made of the raw slice [a, b)
__device__ __forceinline__ uint64_t value_len_range(const FieldView& p, uint64_t a, uint64_t b)
{
    if (!trim_and_test(p, a, b)) return b - a;
    uint64_t o = 0;
    auto count = [&](uint32_t) { ++o; };
    Unquoter<decltype(count)> u{count};
    scan_bytes(p.bytes, p.n, a, b, u);
    u.finish();
    return o;
}

__device__ __forceinline__ uint64_t value_len(const FieldView& p, uint32_t r)
{
    uint64_t a, b;
    if (!field_range(p, r, a, b)) return 0;
    return value_len_range(p, a, b);
}

constexpr int kMultiThreads = 256;
constexpr int kColBatch = 4;
constexpr uint32_t kColBatchMin = 8;   // columns from which the batched form pays

__device__ __forceinline__ uint32_t ldg_u8(const uint8_t* p)
{
    uint32_t v;
    asm volatile("ld.global.nc.u8 %0, [%1];" : "=r"(v) : "l"(p));
    return v;
}

// The dependent chain of one value is index slots -> first / last byte -> (only for padded or quoted values) a walk.
// A thread that resolves the columns of its row issues each level for a batch of kColBatch columns before it uses any of
// them: kColBatch loads in flight per thread instead of one round trip after another (1 GiB unquoted, 8 / 16 columns:
// 3.43 -> 3.00 ms, 6.83 -> 4.82 ms; batches of 8 cost 80 registers and lose it again; the single-column kernels are
// DRAM-sector-bound at 0.75-0.89 of peak and gain nothing from it).  ok bit k: the value exists;
// plain bit k: neither end is padding and it is not a quoted field, so the value IS the raw slice [a, b).
template <int kB>
struct RangeBatch {
    uint64_t a[kB], b[kB];
    uint32_t ok, plain, quoted;   // quoted bit k: not padded, both ends are quotes: the value is the unquoted [a + 1, b - 1)
};
// item(k, rec, fld) -> does item k exist, and which (record, field) it is
template <int kB, class Item>
__device__ __forceinline__ void range_batch_load(const FieldView& p, Item item, RangeBatch<kB>& rb)
{
    rb.ok = rb.plain = rb.quoted = 0;
#pragma unroll
    for (int k = 0; k < kB; ++k) {
        rb.a[k] = rb.b[k] = 0;
        uint32_t r = 0, f = 0;
        if (!item(k, r, f)) continue;
        if ((uint32_t)(r + 1u) >= p.record_cnt || f >= p.field_cnt) continue;
        const uint32_t s = (uint32_t)(r + 1u) * p.row_size + f;   // u32 arithmetic, as the reference
        if ((uint64_t)s + 1 >= p.index_len) continue;
        rb.a[k] = ldg_u64(p.index + s);
        rb.b[k] = ldg_u64(p.index + s + 1);
        rb.ok |= 1u << k;
    }
#pragma unroll
    for (int k = 0; k < kB; ++k) {
        if (!(rb.ok >> k & 1u)) continue;
        rb.a[k] = rb.a[k] + 1 - p.pos_bias;
        rb.b[k] = rb.b[k] - p.pos_bias;
        if (!(rb.a[k] <= rb.b[k] && rb.b[k] <= p.n)) rb.ok &= ~(1u << k);
    }
    if (p.flags == 0u) {
        rb.plain = rb.ok;
        return;
    }
    uint32_t fb[kB], lb[kB];
#pragma unroll
    for (int k = 0; k < kB; ++k) {
        fb[k] = lb[k] = 0;
        if ((rb.ok >> k & 1u) && rb.b[k] > rb.a[k]) {
            fb[k] = ldg_u8(p.bytes + rb.a[k]);
            lb[k] = ldg_u8(p.bytes + rb.b[k] - 1);
        }
    }
#pragma unroll
    for (int k = 0; k < kB; ++k) {
        if (!(rb.ok >> k & 1u)) continue;
        const bool padded = (p.flags & 2u) && (fb[k] == 0x20u || fb[k] == 0x09u || lb[k] == 0x20u || lb[k] == 0x09u);
        const bool quoted = (p.flags & 1u) && rb.b[k] - rb.a[k] >= 2 && fb[k] == 0x22u && lb[k] == 0x22u;
        if (rb.b[k] == rb.a[k] || (!padded && !quoted)) rb.plain |= 1u << k;
        else if (!padded) rb.quoted |= 1u << k;
    }
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
using ColBatch = RangeBatch<kColBatch>;
__device__ __forceinline__ void col_batch_load(const MaterializeMultiParams& p, uint32_t r, uint32_t c0, bool row_live, ColBatch& cb)
{
    range_batch_load(view_of(p, 0), [&](int k, uint32_t& rec, uint32_t& fld) {
        const uint32_t c = c0 + (uint32_t)k;
        if (!row_live || c >= p.ncols) return false;
        rec = r;
        fld = p.field_idx[c];
        return true;
    }, cb);
}

template <bool kBatched>   // separate kernels: the batched form needs 64 registers, the plain one 32
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
    // fewer than kColBatchMin columns: one value after the other (the batches cost registers: 4 columns 1.86 -> 2.01 ms
    // unquoted, 3.01 -> 3.29 ms quoted, where 8 / 16 unquoted columns go 3.43 -> 2.96 ms and 6.83 -> 4.74 ms)
    for (uint32_t c = 0; !kBatched && c < p.ncols; ++c)
        s_len[c * kMultiThreads + tid] = i < p.nrec ? (uint32_t)value_len(view_of(p, c), p.first_record + (uint32_t)i) : 0u;
    for (uint32_t c0 = 0; kBatched && c0 < p.ncols; c0 += kColBatch) {
        ColBatch cb;
        col_batch_load(p, p.first_record + (uint32_t)i, c0, i < p.nrec, cb);
#pragma unroll
        for (int k = 0; k < kColBatch; ++k) {
            const uint32_t c = c0 + (uint32_t)k;
            if (c >= p.ncols) break;
            uint32_t len = 0;
            if (cb.plain >> k & 1u)
                len = (uint32_t)(cb.b[k] - cb.a[k]);
            else if (cb.quoted >> k & 1u) {
                uint32_t o = 0;
                auto count = [&](uint32_t) { ++o; };
                Unquoter<decltype(count)> u{count};
                scan_bytes(p.bytes, p.n, cb.a[k] + 1, cb.b[k] - 1, u);
                u.finish();
                len = o;
            } else if (cb.ok >> k & 1u)
                len = (uint32_t)value_len_range(view_of(p, c), cb.a[k], cb.b[k]);          // padded: trim first
            s_len[c * kMultiThreads + tid] = len;
        }
    }
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
    // one chain per column, thread c walks column c's: all the chains of the tile advance together in one warp
    // (a warp per chain with 128-descriptor hops, as the sweep below does, measured slower here: 0.404 against 0.362 ms
    // for 8 columns, 0.697 against 0.513 ms for 16 -- with 256-row tiles the walks are short and the chains queue up)
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

template <bool kBatched>
__global__ void __launch_bounds__(kMultiThreads) materialize_multi_write_kernel(const MaterializeMultiParams p)
{
    const uint8_t* __restrict__ x = p.bytes;
    const uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= p.nrec) return;
    for (uint32_t c = 0; !kBatched && c < p.ncols; ++c) {
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
    for (uint32_t c0 = 0; kBatched && c0 < p.ncols; c0 += kColBatch) {
        ColBatch cb;
        col_batch_load(p, p.first_record + (uint32_t)i, c0, true, cb);
        uint64_t o[kColBatch], o_end[kColBatch];
#pragma unroll
        for (int k = 0; k < kColBatch; ++k) {
            o[k] = o_end[k] = 0;
            if (cb.ok >> k & 1u) {
                o[k] = p.offsets[c0 + k][i];
                o_end[k] = p.offsets[c0 + k][i + 1];
            }
        }
#pragma unroll
        for (int k = 0; k < kColBatch; ++k) {
            const uint32_t c = c0 + (uint32_t)k;
            if (c >= p.ncols) break;
            if (!(cb.ok >> k & 1u) || o_end[k] > p.out_cap[c]) continue;   // the device form clips whole values
            uint8_t* __restrict__ out = p.out[c];
            uint64_t at = o[k];
            auto put = [&](uint32_t ch) { out[at++] = (uint8_t)ch; };
            uint64_t a = cb.a[k], b = cb.b[k];
            if (cb.plain >> k & 1u) {
                if (b - a < 16) {   // short raw values: a handful of byte loads beat 16 predicated steps per chunk
                    for (; a < b; ++a) put(x[a]);
                } else {
                    scan_bytes(x, p.n, a, b, put);
                }
                continue;
            }
            if (cb.quoted >> k & 1u) {
                Unquoter<decltype(put)> u{put};
                scan_bytes(x, p.n, a + 1, b - 1, u);
                u.finish();
                continue;
            }
            const FieldView fv = view_of(p, c);
            if (!trim_and_test(fv, a, b)) {
                scan_bytes(x, p.n, a, b, put);
            } else {
                Unquoter<decltype(put)> u{put};
                scan_bytes(x, p.n, a, b, u);
                u.finish();
            }
        }
    }
}

// ---- several columns, row sweep -------------------------------------------------------------------------------------
// When the requested columns cover a fair share of every row, the cheapest way to fetch them is to stream the rows:
// a tile is R consecutive records (R = 32, 64 or 128, chosen on the host from the average row length); its bytes are
// ONE contiguous range of the input and its index slots one contiguous range of the index.  One kernel, three forms:
// offsets only, write only (the host form's second call), or both in ONE pass over the input (the device form called
// with its destinations: lengths, scan, look-back, offsets and values without reading anything twice).  The CTA
//   A. loads the row's slots (a warp per row: whole 128-byte lines of the index) as offsets relative to the tile's
//      first byte, and
//   B. copies the tile's bytes into shared memory with 16-byte loads (every input sector is fetched once, coalesced);
//   C. warp (g, c) owns records 32 g .. 32 g + 31 of column c, one value per lane: trim / quote test / "" collapse run
//      on shared memory (a byte walk costs an LDS, not an L2 round trip; quoted values are compacted in place), then
//      a warp scan of the lengths gives the value's offset inside the tile;
//   D. (offsets) one decoupled look-back chain per column over the tile totals, a 32-tile window per step;
//   E. (offsets) base + local offsets leave coalesced;
//   F. (write) values are packed per column into a staging area laid out so that its 16-byte chunks map to 16-byte
//      aligned destinations, and leave as 16-byte stores.
// A tile whose bytes do not fit (one huge quoted field), whose slot arithmetic wraps u32 or whose slots are not
// monotonic takes the per-value global path of the kernels above for steps C and F.
constexpr int kSweepThreads = 256;
constexpr int kSweepWarps = kSweepThreads / 32;
constexpr uint32_t kSweepMaxCap = 32u * 1024u - 16u; // input bytes staged per tile (relative offsets fit 15 bits)
constexpr uint32_t kSweepMaxSlotWords = 4096;        // R * slot pitch
constexpr uint32_t kSweepMaxPairs = 2048;            // R * ncols
constexpr uint32_t kSlotNone = 0xffffffffu;

__device__ __forceinline__ uint32_t sweep_width(uint32_t ncols) { return ncols | 1u; }           // odd pitches: lanes = rows
__device__ __forceinline__ uint32_t sweep_slot_pitch(uint32_t row_size) { return (row_size + 1u) | 1u; }   // hit 32 banks

template <bool kOffsets, bool kWrite>
__global__ void __launch_bounds__(kSweepThreads) materialize_sweep_kernel(const MaterializeMultiParams p)
{
    extern __shared__ __align__(16) uint8_t s_raw[];
    __shared__ uint64_t s_base[kMatMaxCols];
    __shared__ uint32_t s_gtot[kMatMaxCols][4], s_gpre[kMatMaxCols][4];
    __shared__ uint32_t s_colstart[kMatMaxCols + 1], s_mis[kMatMaxCols], s_keep[kMatMaxCols];
    __shared__ uint32_t s_tile, s_fits;
    const uint32_t tid = threadIdx.x, lane = tid & 31u, warp = tid >> 5;
    const uint32_t R = p.rows_per_tile, G = R >> 5, W = sweep_width(p.ncols), pitch = sweep_slot_pitch(p.row_size);
    const uint32_t cap = p.cap_bytes;
    const uint32_t stage_cap = cap + 16u * (kMatMaxCols + 1);
    uint8_t* s_in = s_raw;
    uint8_t* s_out = s_raw + cap + 16u;
    uint32_t* s_slot = reinterpret_cast<uint32_t*>(s_out + (kWrite ? stage_cap : 0u));
    uint32_t* s_val = s_slot + R * pitch;                     // [R][W] start | len << 16 of the (trimmed, unquoted) value
    uint32_t* s_loc = s_val + R * W;                          // [R][W] exclusive offset inside the 32-record group
    if (kOffsets) {
        if (tid == 0) s_tile = atomicAdd(p.ticket, 1u);       // tiles are taken in order: look-back only waits on started tiles
        __syncthreads();
    }
    const uint32_t tile = kOffsets ? s_tile : blockIdx.x;
    const uint32_t r0 = tile * R, nr = p.nrec - r0 < R ? p.nrec - r0 : R;
    // rows of the tile in the reference's numbering (the header is row 0); a row at or past record_cnt is Ok(None)
    const uint64_t first_row = (uint64_t)p.first_record + r0 + 1u;
    uint32_t nrv = first_row < p.record_cnt ? (p.record_cnt - first_row < nr ? (uint32_t)(p.record_cnt - first_row) : nr) : 0u;
    const uint64_t slot_first = first_row * p.row_size, slot_last = slot_first + (uint64_t)nrv * p.row_size;
    if (slot_first >= p.index_len) nrv = 0;
    const bool wraps = slot_last >= 0xffffffffull;            // the reference computes slots in u32: take the exact path
    // ---- A + B: slots and bytes, every global load of a step in flight before the first one is used ----
    // The tile's slots are one contiguous run of the index: the rows' own slots (nmain of them) + the one that ends the
    // last row (= the tile's last slot, fetched as `hi`).  (row, f) of slot j advance incrementally (no division in
    // the loop); the slot that ends a row is also the next row's first.
    constexpr int kSlotBatch = 8, kByteBatch = 4;
    const uint32_t rs = p.row_size, nmain = (nrv && !wraps) ? nrv * rs : 0u;
    const uint64_t* __restrict__ src = p.index + slot_first;
    const uint64_t avail = nmain ? p.index_len - slot_first : 0ull;
    uint64_t lo = 0, hi = 0, v[kSlotBatch];
    if (nmain) {
        lo = ldg_u64(src);
        hi = ldg_u64(p.index + (slot_last < p.index_len ? slot_last : p.index_len - 1u));
    }
#pragma unroll
    for (int k = 0; k < kSlotBatch; ++k) {
        const uint32_t j = tid + (uint32_t)k * kSweepThreads;
        v[k] = (j < nmain && j < avail) ? ldg_u64(src + j) : 0ull;
    }
    lo = lo + 1u - p.pos_bias;                                // first byte of the tile's first row
    hi = hi - p.pos_bias;                                     // the separator that ends its last row (or the last entry)
    const uint64_t lo16 = lo & ~15ull;
    int bad = (!nmain || hi < lo || hi > p.n || hi - lo16 > cap) ? (nrv ? 1 : 0) : 0;
    const bool want_bytes = nmain && !bad && (p.flags != 0u || kWrite);
    const uint32_t span0 = want_bytes ? (uint32_t)(hi - lo16) : 0u;   // (the slots are still to be checked)
    uint4 bytes[kByteBatch];
#pragma unroll
    for (int k = 0; k < kByteBatch; ++k) {
        const uint32_t j = (tid + (uint32_t)k * kSweepThreads) * 16u;
        bytes[k] = make_uint4(0, 0, 0, 0);
        if (j < span0 && lo16 + j + 16 <= p.n) bytes[k] = ldg_128(p.bytes + lo16 + j);
    }
    if (nmain && !bad) {
        const uint64_t sub = p.pos_bias + lo16 - 1u;
        const uint32_t drow = kSweepThreads / rs, df = kSweepThreads - drow * rs;
        uint32_t row = tid / rs, f = tid - row * rs;
        auto put_slot = [&](uint64_t raw, bool present) {
            uint32_t rel = kSlotNone;                         // stored: the byte AFTER the separator, relative to lo16
            if (present) {
                const uint64_t d = raw - sub;
                if (d > (uint64_t)cap + 1u) bad = 1;          // not inside the tile's range: a slot out of order
                rel = (uint32_t)d;
            }
            s_slot[row * pitch + f] = rel;
            if (f == 0u && row > 0u) s_slot[(row - 1u) * pitch + rs] = rel;
            row += drow;
            f += df;
            if (f >= rs) {
                f -= rs;
                ++row;
            }
        };
#pragma unroll
        for (int k = 0; k < kSlotBatch; ++k) {
            const uint32_t j = tid + (uint32_t)k * kSweepThreads;
            if (j < nmain) put_slot(v[k], j < avail);
        }
        for (uint32_t j = tid + kSlotBatch * kSweepThreads; j < nmain; j += kSweepThreads)   // long rows only
            put_slot(j < avail ? ldg_u64(src + j) : 0ull, j < avail);
        if (tid == 0)                                         // the slot that ends the last row
            s_slot[(nrv - 1u) * pitch + rs] = slot_last < p.index_len ? (uint32_t)(hi + 1u - lo16) : kSlotNone;
    }
    const bool staged = __syncthreads_or(bad) == 0;
    const uint32_t span = (staged && nmain) ? (uint32_t)(hi - lo16) : 0u;
    if (staged && want_bytes) {
#pragma unroll
        for (int k = 0; k < kByteBatch; ++k) {
            const uint32_t j = (tid + (uint32_t)k * kSweepThreads) * 16u;
            if (j < span) {
                if (lo16 + j + 16 > p.n) {                    // the last bytes of the input, one at a time
                    uint32_t w[4] = {0, 0, 0, 0};
                    for (uint32_t b = 0; b < 16 && lo16 + j + b < p.n; ++b) w[b >> 2] |= (uint32_t)p.bytes[lo16 + j + b] << (8 * (b & 3));
                    bytes[k] = make_uint4(w[0], w[1], w[2], w[3]);
                }
                *reinterpret_cast<uint4*>(s_in + j) = bytes[k];
            }
        }
        for (uint32_t j = (tid + kByteBatch * kSweepThreads) * 16u; j < span; j += kSweepThreads * 16u) {
            const uint64_t g = lo16 + j;
            uint4 b4;
            if (g + 16 <= p.n) {
                b4 = ldg_128(p.bytes + g);
            } else {
                uint32_t w[4] = {0, 0, 0, 0};
                for (uint32_t b = 0; b < 16 && g + b < p.n; ++b) w[b >> 2] |= (uint32_t)p.bytes[g + b] << (8 * (b & 3));
                b4 = make_uint4(w[0], w[1], w[2], w[3]);
            }
            *reinterpret_cast<uint4*>(s_in + j) = b4;
        }
    }
    if (staged && want_bytes) __syncthreads();
    // ---- C: one value per lane ----
    const uint32_t g = warp % G, row = 32u * g + lane;
    for (uint32_t c = warp / G; c < p.ncols; c += kSweepWarps / G) {
        const uint32_t f = p.field_idx[c];
        uint32_t a = 0, len = 0;
        if (staged) {
            bool ok = row < nrv && f < p.field_cnt;
            uint32_t b = 0;
            if (ok) {
                a = s_slot[row * pitch + f];
                const uint32_t nb = s_slot[row * pitch + f + 1u];
                b = nb - 1u;                                  // the next separator itself
                ok = a != kSlotNone && nb != kSlotNone && nb != 0u && a <= b;
            }
            if (ok) {
                if (p.flags & 2u) {
                    while (a < b && (s_in[a] == 0x20u || s_in[a] == 0x09u)) ++a;
                    while (b > a && (s_in[b - 1] == 0x20u || s_in[b - 1] == 0x09u)) --b;
                }
                if ((p.flags & 1u) && b - a >= 2u && s_in[a] == 0x22u && s_in[b - 1] == 0x22u) {
                    ++a;
                    --b;
                    if (kWrite) {                             // collapse "" in place: the output never passes the input
                        uint8_t* dst = s_in + a;
                        auto put = [&](uint32_t ch) { *dst++ = (uint8_t)ch; };
                        Unquoter<decltype(put)> u{put};
                        for (uint32_t k = a; k < b; ++k) u(s_in[k]);
                        u.finish();
                        len = (uint32_t)(dst - (s_in + a));
                    } else {
                        auto count = [&](uint32_t) { ++len; };
                        Unquoter<decltype(count)> u{count};
                        for (uint32_t k = a; k < b; ++k) u(s_in[k]);
                        u.finish();
                    }
                } else {
                    len = b - a;
                }
            }
        } else if (row < nr) {
            len = (uint32_t)value_len(view_of(p, c), p.first_record + r0 + row);
        }
        uint32_t inc = len;
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) {
            const uint32_t u = __shfl_up_sync(0xffffffffu, inc, d);
            if (lane >= (uint32_t)d) inc += u;
        }
        s_val[row * W + c] = staged ? a | (len << 16) : len;
        s_loc[row * W + c] = inc - len;
        if (lane == 31) s_gtot[c][g] = inc;
    }
    __syncthreads();
    // ---- D: column bases ----
    if (kOffsets) {
        for (uint32_t c = warp; c < p.ncols; c += kSweepWarps) {
            uint64_t* desc = p.tile_desc + (uint64_t)c * p.tiles;
            uint64_t total = 0;
            for (uint32_t k = 0; k < G; ++k) {
                if (lane == 0) s_gpre[c][k] = (uint32_t)total;
                total += s_gtot[c][k];
            }
            const uint64_t prefix = lookback_warp(desc, tile, total, lane);
            if (lane == 0) {
                s_base[c] = prefix;
                if (r0 + nr == p.nrec) p.offsets[c][p.nrec] = prefix + total;
            }
        }
    } else if (tid < p.ncols) {
        s_base[tid] = p.offsets[tid][r0];
        uint32_t run = 0;
        for (uint32_t k = 0; k < G; ++k) {
            s_gpre[tid][k] = run;
            run += s_gtot[tid][k];
        }
    }
    __syncthreads();
    // ---- E: offsets ----
    if (kOffsets && row < nr) {
        for (uint32_t c = warp / G; c < p.ncols; c += kSweepWarps / G) {
            p.offsets[c][r0 + row] = s_base[c] + s_gpre[c][g] + s_loc[row * W + c];
        }
    }
    if (!kWrite) return;
    // ---- F: values ----
    if (tid == 0) {
        uint32_t at = 0;
        bool fits = staged;
        for (uint32_t c = 0; c < p.ncols && fits; ++c) {
            uint32_t total = 0;
            for (uint32_t k = 0; k < G; ++k) total += s_gtot[c][k];
            const uint32_t mis = (uint32_t)((reinterpret_cast<uintptr_t>(p.out[c]) + s_base[c]) & 15u);
            s_colstart[c] = at;
            s_mis[c] = mis;
            s_keep[c] = 0;
            at += (mis + total + 15u) & ~15u;
            if (at > stage_cap) fits = false;                  // cannot happen with distinct columns (the host checks)
        }
        s_colstart[p.ncols] = at;
        s_fits = fits ? 1u : 0u;
    }
    __syncthreads();
    const bool fits = s_fits != 0u;
    for (uint32_t c = warp / G; c < p.ncols; c += kSweepWarps / G) {
        if (row >= nr) break;
        const uint32_t local = s_gpre[c][g] + s_loc[row * W + c];
        const uint32_t av = s_val[row * W + c];
        if (fits) {
            const uint32_t a = av & 0xffffu, len = av >> 16;
            if (len == 0u || s_base[c] + local + len > p.out_cap[c]) continue;          // the device form clips whole values
            atomicMax(&s_keep[c], local + len);
            uint8_t* dst = s_out + s_colstart[c] + s_mis[c] + local;
            const uint8_t* src = s_in + a;
            for (uint32_t k = 0; k < len; ++k) dst[k] = src[k];
        } else {                                               // per value, straight from and to global memory
            const FieldView fv = view_of(p, c);
            const uint8_t* __restrict__ x = p.bytes;
            uint8_t* __restrict__ out = p.out[c];
            uint64_t va, vb;
            if (!field_range(fv, p.first_record + r0 + row, va, vb)) continue;
            uint64_t o = s_base[c] + local;
            const uint64_t len = staged ? av >> 16 : av;       // staged, but the staging area overflowed: see s_fits
            if (o + len > p.out_cap[c]) continue;
            auto put = [&](uint32_t ch) { out[o++] = (uint8_t)ch; };
            if (!trim_and_test(fv, va, vb)) {
                scan_bytes(x, p.n, va, vb, put);
            } else {
                Unquoter<decltype(put)> u{put};
                scan_bytes(x, p.n, va, vb, u);
                u.finish();
            }
        }
    }
    if (!fits) return;
    __syncthreads();
    const uint32_t chunks = s_colstart[p.ncols] >> 4;
    for (uint32_t q = tid; q < chunks; q += kSweepThreads) {
        const uint32_t at = q << 4;
        uint32_t c = 0;
        while (c + 1 < p.ncols && s_colstart[c + 1] <= at) ++c;
        const uint32_t local = at - s_colstart[c];              // byte offset inside the column's region (mis included)
        const uint32_t first = s_mis[c], last = s_mis[c] + s_keep[c];   // live bytes of the region: [first, last)
        uint8_t* gp = p.out[c] + s_base[c] - s_mis[c] + local;
        if (local >= first && local + 16u <= last) {
            *reinterpret_cast<uint4*>(gp) = *reinterpret_cast<const uint4*>(s_out + at);
        } else {
            for (uint32_t b = 0; b < 16; ++b)
                if (local + b >= first && local + b < last) gp[b] = s_out[at + b];
        }
    }
}

size_t sweep_smem_bytes(const MaterializeMultiParams& p, bool write)
{
    const size_t R = p.rows_per_tile;
    return (size_t)p.cap_bytes + 16 + (write ? (size_t)p.cap_bytes + 16 * (kMatMaxCols + 1) : 0) +
           (R * (((size_t)p.row_size + 1) | 1) + 2 * R * (p.ncols | 1u)) * sizeof(uint32_t);
}

}  // namespace

// Plan of the row sweep: rows per tile (0 = use the per-row kernels) and the bytes of input staged per tile.
// OPT-IN (CSVB200_MAT_SWEEP): measured on a B200 (256 MiB of the 16-field workloads, UNQUOTE | TRIM, offsets + values in
// one call, ms; profiles/r02_mat_sweep_ab.jsonl):
//                         4 columns        8 columns        16 columns
//     unquoted  per-row   0.535            0.933            1.785
//               sweep     0.684            1.095            1.694
//     quoted    per-row   0.865            1.661            1.850
//               sweep     0.677            1.356            2.214
// The sweep reads every input byte and index slot once, coalesced (DRAM traffic 0.72 GB against 1.9 GB for 8 columns),
// but a tile publishes its column totals only after all of its work, so every tile ends up waiting in the look-back
// for the slowest of the ~100 tiles in flight before it (ncu: a third of the warp samples), and the kernel needs
// ~450 thread instructions per value.  It wins on quoted columns (the byte walks run on shared memory) and loses on
// short unquoted values, so the per-row kernels stay the default; what would make it win everywhere is the deferred
// look-back of the index build (classify tile t + 1 while tile t's prefix resolves).
uint32_t materialize_sweep_plan(uint64_t n, uint32_t record_cnt, uint32_t row_size, uint32_t ncols, uint32_t* cap_bytes)
{
    // CSVB200_MAT_SWEEP: unset / 0 = per-row kernels, 1 = sweep, 32 / 64 / 128 = sweep with that many rows per tile;
    // CSVB200_MAT_CAP: staged bytes per tile (tests: small values push tiles onto the per-value path).  Read per call.
    const char* env = getenv("CSVB200_MAT_SWEEP");
    const int force = env ? atoi(env) : -1;
    *cap_bytes = 0;
    if (force <= 0 || ncols == 0) return 0;
    const uint64_t avg_row = n / (record_cnt ? record_cnt : 1u) + 1;
    const uint64_t pitch = ((uint64_t)row_size + 1) | 1;
    uint32_t rows = 0;
    for (uint32_t r = 128; r >= 32; r >>= 1) {
        if (force > 1 && r != (uint32_t)force) continue;
        if (force <= 1 && (r * avg_row * 3 / 2 > kSweepMaxCap || r * pitch > kSweepMaxSlotWords || r * ncols > kSweepMaxPairs)) continue;
        rows = r;
        break;
    }
    if (rows == 0 || rows * pitch > kSweepMaxSlotWords || rows * ncols > kSweepMaxPairs) return 0;
    uint64_t cap = (rows * avg_row * 2 + 1023) & ~1023ull;
    if (cap < 4096) cap = 4096;
    if (cap > kSweepMaxCap) cap = kSweepMaxCap;
    if (const char* e = getenv("CSVB200_MAT_CAP")) {
        const long v = atol(e);
        if (v >= 64 && v <= (long)kSweepMaxCap) cap = (uint64_t)v & ~15ull;
    }
    *cap_bytes = (uint32_t)cap;
    return rows;
}

size_t materialize_multi_scratch_bytes(uint32_t nrec, uint32_t ncols, uint32_t rows_per_tile)
{
    const size_t per = rows_per_tile ? rows_per_tile : (uint32_t)kMultiThreads;
    const size_t tiles = ((size_t)nrec + per - 1) / per + 1;
    return 128 + tiles * ncols * sizeof(uint64_t);
}

template <bool kOffsets, bool kWrite>
static cudaError_t sweep_launch(const MaterializeMultiParams& p, cudaStream_t stream)
{
    static std::mutex mu;                                   // one attribute cache per instantiation and device
    static bool done[64] = {};
    constexpr size_t kMax = kSweepMaxCap + 16 + kSweepMaxCap + 16 * (kMatMaxCols + 1) +
                            (kSweepMaxSlotWords + 2 * (kSweepMaxPairs + 128)) * sizeof(uint32_t);
    int dev = 0;
    cudaError_t e = cudaGetDevice(&dev);
    if (e != cudaSuccess) return e;
    {
        std::lock_guard<std::mutex> lock(mu);
        if (dev >= 64 || !done[dev]) {
            e = cudaFuncSetAttribute(materialize_sweep_kernel<kOffsets, kWrite>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kMax);
            if (e != cudaSuccess) return e;
            if (dev < 64) done[dev] = true;
        }
    }
    materialize_sweep_kernel<kOffsets, kWrite><<<p.tiles, kSweepThreads, sweep_smem_bytes(p, kWrite), stream>>>(p);
    return cudaGetLastError();
}

// one launch: offsets, values, or both in one pass
cudaError_t launch_materialize_sweep(const MaterializeMultiParams& p, bool offsets, bool write, cudaStream_t stream)
{
    if (p.tiles == 0 || p.ncols == 0 || !(offsets || write)) return cudaSuccess;
    if (offsets && write) return sweep_launch<true, true>(p, stream);
    return offsets ? sweep_launch<true, false>(p, stream) : sweep_launch<false, true>(p, stream);
}

cudaError_t launch_materialize_multi_offsets(const MaterializeMultiParams& p, cudaStream_t stream)
{
    if (p.tiles == 0 || p.ncols == 0) return cudaSuccess;
    const size_t smem = (size_t)p.ncols * kMultiThreads * sizeof(uint32_t);
    if (p.ncols >= kColBatchMin)
        materialize_multi_offsets_kernel<true><<<p.tiles, kMultiThreads, smem, stream>>>(p);
    else
        materialize_multi_offsets_kernel<false><<<p.tiles, kMultiThreads, smem, stream>>>(p);
    return cudaGetLastError();
}

cudaError_t launch_materialize_multi_write(const MaterializeMultiParams& p, cudaStream_t stream)
{
    if (p.nrec == 0 || p.ncols == 0) return cudaSuccess;
    const unsigned blocks = (p.nrec + kMultiThreads - 1) / kMultiThreads;
    if (p.ncols >= kColBatchMin)
        materialize_multi_write_kernel<true><<<blocks, kMultiThreads, 0, stream>>>(p);
    else
        materialize_multi_write_kernel<false><<<blocks, kMultiThreads, 0, stream>>>(p);
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
