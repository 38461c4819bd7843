/*
 * csv_oracle.c -- TEST INFRASTRUCTURE ONLY (checker, never the product path).
 *
 * CPU restatement of the csv-simd hot path (csv -> structural index -> record /
 * field lookup).  Only tests/, __graft_entry__.smoke() and bench.py's
 * cpu_baseline / --impl reference legs may load this library.  The product
 * (csv_simd_b200/) never links, imports or calls anything in oracle/.
 *
 * The reference is a Rust crate; there is no rustc/cargo in this image and its
 * crates (memmap 0.7.0, thiserror 1.0.23, bytemuck 1.5.0) are not vendored, so
 * oracle/_ref cannot be built: the reference is "unbuildable here".  None of
 * the path's arithmetic lives in those crates -- it is all core::arch::x86_64
 * intrinsics, which exist 1:1 in <immintrin.h>, so this file restates the
 * reference instruction for instruction.
 *
 * PINNING STATUS: the reference's own tests pin only
 *   - reader::tests::mk_index (src/reader.rs:318-327): index[1]==4 and
 *     index[last]==95 on res/reader_test01.csv,
 *   - the `boundaries` doctest (src/tape.rs:362-384),
 *   - the blsr identity (src/lib.rs:139-152).
 * tests/test_oracle.py checks this file against all three.  Everything else
 * (Header, Tape, seek_record, seek_field, CRLF, quoted separators, the tail
 * block) is PARITY UNPINNED by the reference's tests; it is pinned here by the
 * literal restatement below cross-checked against an independent closed-form
 * scalar model (oracle_read_closed_form) on fixtures and fuzzed inputs.
 *
 * Each function cites the reference file:line it follows.
 */
#include <immintrin.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#include "csv_oracle.h"

/* ------------------------------------------------------------------------ */
/* Vec<usize> growth model: RawVec::grow_amortized (cap = max(2*cap, needed, */
/* 4)) -- only matters for the timed CPU baseline, not for results.          */
/* ------------------------------------------------------------------------ */
typedef struct {
    uint64_t *ptr;
    size_t len;
    size_t cap;
} vec_u64;

static int vec_reserve(vec_u64 *v, size_t additional)
{
    if (v->cap - v->len >= additional) return 0;
    size_t need = v->len + additional;
    size_t ncap = v->cap * 2;
    if (ncap < need) ncap = need;
    if (ncap < 4) ncap = 4;
    uint64_t *p = (uint64_t *)realloc(v->ptr, ncap * sizeof(uint64_t));
    if (!p) return -1;
    v->ptr = p;
    v->cap = ncap;
    return 0;
}

/* ------------------------------------------------------------------------ */
/* src/avx/stage1.rs:14-19 SimdInput {v0..v3}                               */
/* ------------------------------------------------------------------------ */
typedef struct {
    __m128i v0, v1, v2, v3;
} simd_input;

/* src/avx/stage1.rs:23-35 SimdInput::new -- four aligned 16-byte loads */
static inline simd_input simd_input_new(const __m128i *ptr)
{
    simd_input in;
    in.v0 = _mm_load_si128(ptr);
    in.v1 = _mm_load_si128(ptr + 1);
    in.v2 = _mm_load_si128(ptr + 2);
    in.v3 = _mm_load_si128(ptr + 3);
    return in;
}

/* src/avx/stage1.rs:37-94 SimdInput::new_with_padding -- remaining 0..3 full
 * vectors, then the <16-byte tail copied into a zeroed 16-byte buffer, then
 * zero vectors.  The two asserts (:45-52) become error returns. */
static int simd_input_new_with_padding(const __m128i *ptr, size_t load,
                                       const uint8_t *tail, size_t tail_len,
                                       simd_input *out)
{
    if (!(load < 4)) return ORACLE_ERR_PANIC;
    if (!(tail_len < 16)) return ORACLE_ERR_PANIC;
    uint8_t padded_tail[16];
    memset(padded_tail, 0, sizeof padded_tail);
    for (size_t i = 0; i < tail_len; ++i) padded_tail[i] = tail[i];
    const __m128i pt = _mm_loadu_si128((const __m128i *)padded_tail);
    const __m128i z = _mm_setzero_si128();
    switch (load) {
    case 3:
        out->v0 = _mm_load_si128(ptr);
        out->v1 = _mm_load_si128(ptr + 1);
        out->v2 = _mm_load_si128(ptr + 2);
        out->v3 = pt;
        break;
    case 2:
        out->v0 = _mm_load_si128(ptr);
        out->v1 = _mm_load_si128(ptr + 1);
        out->v2 = pt;
        out->v3 = z;
        break;
    case 1:
        out->v0 = _mm_load_si128(ptr);
        out->v1 = pt;
        out->v2 = z;
        out->v3 = z;
        break;
    default:
        out->v0 = pt;
        out->v1 = z;
        out->v2 = z;
        out->v3 = z;
        break;
    }
    return 0;
}

/* src/avx/stage1.rs:111-187 get_struct_positions: bit i = (class[i] & search)
 * != 0, bit 0 = lowest address; and + cmpeq(zero) + movemask x4, or, not. */
static inline uint64_t get_struct_positions(uint8_t search, __m128i res0,
                                            __m128i res1, __m128i res2,
                                            __m128i res3)
{
    const __m128i struct_mask = _mm_set1_epi8((char)search);
    const __m128i s0 = _mm_and_si128(res0, struct_mask);
    const __m128i s1 = _mm_and_si128(res1, struct_mask);
    const __m128i s2 = _mm_and_si128(res2, struct_mask);
    const __m128i s3 = _mm_and_si128(res3, struct_mask);
    const __m128i zero = _mm_setzero_si128();
    const uint64_t r0 = (uint32_t)_mm_movemask_epi8(_mm_cmpeq_epi8(s0, zero));
    const uint64_t r1 = (uint32_t)_mm_movemask_epi8(_mm_cmpeq_epi8(s1, zero));
    const uint64_t r2 = (uint32_t)_mm_movemask_epi8(_mm_cmpeq_epi8(s2, zero));
    const uint64_t r3 = (uint32_t)_mm_movemask_epi8(_mm_cmpeq_epi8(s3, zero));
    return ~(r0 | (r1 << 16) | (r2 << 32) | (r3 << 48));
}

/* src/stage1.rs:24-35 the two nibble look-up tables */
#define LOW_NIBBLE_MASK() \
    _mm_setr_epi8(4, 0, 16, 0, 0, 0, 0, 0, 0, 0, 1, 0, 10, 1, 0, 0)
#define HIGH_NIBBLE_MASK() \
    _mm_setr_epi8(1, 0, 22, 0, 0, 8, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0)

/* src/avx/stage1.rs:249-316 classify: class = LO[b & 15] & HI[b >> 4] */
static inline void classify4(const simd_input *in, __m128i *res0,
                             __m128i *res1, __m128i *res2, __m128i *res3)
{
    const __m128i lo_tbl = LOW_NIBBLE_MASK();
    const __m128i hi_tbl = HIGH_NIBBLE_MASK();
    const __m128i low_mask = _mm_set1_epi8(0xf);
    const __m128i nl0 = _mm_and_si128(in->v0, low_mask);
    const __m128i nl1 = _mm_and_si128(in->v1, low_mask);
    const __m128i nl2 = _mm_and_si128(in->v2, low_mask);
    const __m128i nl3 = _mm_and_si128(in->v3, low_mask);
    const __m128i nh0 = _mm_and_si128(_mm_srli_epi64(in->v0, 4), low_mask);
    const __m128i nh1 = _mm_and_si128(_mm_srli_epi64(in->v1, 4), low_mask);
    const __m128i nh2 = _mm_and_si128(_mm_srli_epi64(in->v2, 4), low_mask);
    const __m128i nh3 = _mm_and_si128(_mm_srli_epi64(in->v3, 4), low_mask);
    *res0 = _mm_and_si128(_mm_shuffle_epi8(lo_tbl, nl0), _mm_shuffle_epi8(hi_tbl, nh0));
    *res1 = _mm_and_si128(_mm_shuffle_epi8(lo_tbl, nl1), _mm_shuffle_epi8(hi_tbl, nh1));
    *res2 = _mm_and_si128(_mm_shuffle_epi8(lo_tbl, nl2), _mm_shuffle_epi8(hi_tbl, nh2));
    *res3 = _mm_and_si128(_mm_shuffle_epi8(lo_tbl, nl3), _mm_shuffle_epi8(hi_tbl, nh3));
}

/* src/avx/stage1.rs:193-407 Stage1::structure for SimdInput.
 *   string_mask = clmul(quote_bits, ~0)[63:0] ^ in_string       (:342-382,397)
 *   structure   = all_struct & !string_mask                     (:400-406)
 *   in_string'  = (string_mask as i64) >> 63                    (:407)       */
static inline void structure(const simd_input *in, uint64_t *structure_out,
                             int64_t *in_string)
{
    __m128i res0, res1, res2, res3;
    classify4(in, &res0, &res1, &res2, &res3);
    const __m128i zero = _mm_setzero_si128();
    const __m128i ones = _mm_cmpeq_epi32(zero, zero);
    const uint64_t quote_bits = get_struct_positions(16, res0, res1, res2, res3);
    const uint64_t all_struct = get_struct_positions(3, res0, res1, res2, res3);
    __m128i string_mask =
        _mm_clmulepi64_si128(_mm_set_epi64x(0, (int64_t)quote_bits), ones, 0);
    string_mask = _mm_xor_si128(string_mask, _mm_set_epi64x(0, *in_string));
    const __m128i result = _mm_and_si128(_mm_set_epi64x(0, (int64_t)all_struct),
                                         _mm_xor_si128(string_mask, ones));
    *structure_out = (uint64_t)_mm_cvtsi128_si64(result);
    *in_string = (int64_t)_mm_cvtsi128_si64(string_mask) >> 63;
}

/* src/stage1.rs:162-296 Stage1::crush_set_bits: popcount, reserve(64),
 * set_len(base+64), 8-unrolled (tzcnt, store, blsr) with the over-write trick
 * (garbage codepoint_cnt+64 slots truncated by set_len(next_base)), array_idx
 * kept as u32 exactly like the reference. */
static inline int crush_set_bits(vec_u64 *acc, uint64_t set_bits,
                                 size_t codepoint_cnt, uint32_t *array_idx)
{
    const uint32_t cnt = (uint32_t)__builtin_popcountll(set_bits);
    const size_t base = (size_t)*array_idx;
    const size_t next_base = (size_t)*array_idx + (size_t)cnt;
    if (vec_reserve(acc, 64)) return ORACLE_ERR_OOM;
    uint64_t *ptr = acc->ptr;
    acc->len = base + 64;
    size_t shift = 0;
    while (set_bits != 0) {
        for (int k = 0; k < 8; ++k) {
            /* u64::trailing_zeros(0) == 64 */
            const unsigned tz = set_bits ? (unsigned)__builtin_ctzll(set_bits) : 64u;
            ptr[base + (size_t)k + shift] = (uint64_t)(codepoint_cnt + tz);
            /* set_bits &= set_bits.saturating_sub(1) */
            set_bits &= (set_bits ? set_bits - 1 : 0);
        }
        *array_idx = *array_idx + 8;
        shift += 8;
    }
    acc->len = next_base;
    *array_idx = (uint32_t)next_base;
    return 0;
}

/* src/reader.rs:150-306 reader::read.
 * bytes must be 16-byte aligned (an mmap is page aligned, so align_to's
 * head_u8 is empty; the reference ignores head_u8 -- :180-181,192-199).
 * n < 64 makes the reference over-read and then trip assert!(load < 4)
 * (:220-229, avx/stage1.rs:45-48): reported as ORACLE_ERR_PANIC. */
int oracle_read_sse(const uint8_t *bytes, size_t n, uint64_t **out, size_t *out_len)
{
    if (((uintptr_t)bytes & 15u) != 0) return ORACLE_ERR_UNALIGNED;
    if (n < 64) return ORACLE_ERR_PANIC;
    const __m128i *body_vectors = (const __m128i *)bytes;
    const size_t num_vectors = n / 16;
    const uint8_t *tail_u8 = bytes + num_vectors * 16;
    const size_t tail_len = n - num_vectors * 16;

    size_t simdinput_cnt = 0;
    size_t codepoint_cnt = 0;
    uint64_t set_bits = 0;
    vec_u64 struct_acc = {0, 0, 0};
    if (vec_reserve(&struct_acc, 1)) return ORACLE_ERR_OOM;
    /* vec![0] allocates exactly one slot */
    struct_acc.cap = 1;
    struct_acc.ptr[0] = 0;
    struct_acc.len = 1;
    uint32_t array_idx = 1;
    int64_t inside_str = 0;

    const size_t iter_cnt = num_vectors < 4 ? 0 : num_vectors - 4;
    int rc = 0;
    while (simdinput_cnt <= iter_cnt) {
        const simd_input input = simd_input_new(body_vectors + simdinput_cnt);
        structure(&input, &set_bits, &inside_str);
        if ((rc = crush_set_bits(&struct_acc, set_bits, codepoint_cnt, &array_idx))) goto fail;
        simdinput_cnt += 4;
        codepoint_cnt += 64;
    }
    simd_input padded;
    if ((rc = simd_input_new_with_padding(body_vectors + simdinput_cnt,
                                          num_vectors - simdinput_cnt, tail_u8,
                                          tail_len, &padded)))
        goto fail;
    set_bits = 0;
    structure(&padded, &set_bits, &inside_str);
    if ((rc = crush_set_bits(&struct_acc, set_bits, codepoint_cnt, &array_idx))) goto fail;

    *out = struct_acc.ptr;
    *out_len = struct_acc.len;
    return 0;
fail:
    free(struct_acc.ptr);
    return rc;
}

/* Timed variant for the CPU baseline: same loop, result discarded except the
 * length and a checksum (sum of entries mod 2^64). */
int oracle_read_sse_timed(const uint8_t *bytes, size_t n, size_t *out_len, uint64_t *checksum)
{
    uint64_t *idx = NULL;
    size_t len = 0;
    int rc = oracle_read_sse(bytes, n, &idx, &len);
    if (rc) return rc;
    uint64_t s = 0;
    for (size_t i = 0; i < len; ++i) s += idx[i];
    *out_len = len;
    *checksum = s;
    free(idx);
    return 0;
}

/* Independent closed form (SURVEY.md TL;DR): index = [0] ++ [i : b[i] in
 * {',', CR, LF} and #quotes in b[0..i) is even].  Any n, any alignment.
 * start_parity / pos_bias let tests model a shard that starts inside a quoted
 * region at a global byte offset (pos_bias is added to every position; the
 * sentinel is emitted only when with_sentinel != 0). */
int oracle_read_closed_form(const uint8_t *bytes, size_t n, int start_parity,
                            uint64_t pos_bias, int with_sentinel,
                            uint64_t **out, size_t *out_len, int *end_parity)
{
    size_t cnt = with_sentinel ? 1 : 0;
    int par = start_parity & 1;
    for (size_t i = 0; i < n; ++i) {
        const uint8_t b = bytes[i];
        if (b == 0x22) par ^= 1;
        else if (!par && (b == 0x2C || b == 0x0D || b == 0x0A)) ++cnt;
    }
    uint64_t *idx = (uint64_t *)malloc((cnt ? cnt : 1) * sizeof(uint64_t));
    if (!idx) return ORACLE_ERR_OOM;
    size_t k = 0;
    if (with_sentinel) idx[k++] = 0;
    par = start_parity & 1;
    for (size_t i = 0; i < n; ++i) {
        const uint8_t b = bytes[i];
        if (b == 0x22) par ^= 1;
        else if (!par && (b == 0x2C || b == 0x0D || b == 0x0A)) idx[k++] = pos_bias + i;
    }
    *out = idx;
    *out_len = cnt;
    if (end_parity) *end_parity = par;
    return 0;
}

/* Shard summary used by the multi-GPU exchange: quote parity of
 * the shard, unquoted-separator count if the shard is entered outside quotes
 * (c0) and total separator count (s); c1 = s - c0. */
void oracle_shard_summary(const uint8_t *bytes, size_t n, uint64_t *parity,
                          uint64_t *c0, uint64_t *s)
{
    uint64_t par = 0, a = 0, t = 0;
    for (size_t i = 0; i < n; ++i) {
        const uint8_t b = bytes[i];
        if (b == 0x22) par ^= 1;
        else if (b == 0x2C || b == 0x0D || b == 0x0A) {
            ++t;
            if (!par) ++a;
        }
    }
    *parity = par;
    *c0 = a;
    *s = t;
}

void oracle_free(void *p) { free(p); }

/* src/structure.rs:10-58 structure::run -- class byte per input byte (16 B). */
void oracle_structure_run(const uint8_t *chunk, size_t at, uint8_t out16[16])
{
    const __m128i lo_tbl = LOW_NIBBLE_MASK();
    const __m128i hi_tbl = HIGH_NIBBLE_MASK();
    const __m128i low_mask = _mm_set1_epi8(0xf);
    const __m128i c = _mm_loadu_si128((const __m128i *)(chunk + at));
    const __m128i nib_lo = _mm_and_si128(c, low_mask);
    const __m128i nib_hi = _mm_and_si128(_mm_srli_epi64(c, 4), low_mask);
    const __m128i r = _mm_and_si128(_mm_shuffle_epi8(lo_tbl, nib_lo),
                                    _mm_shuffle_epi8(hi_tbl, nib_hi));
    _mm_storeu_si128((__m128i *)out16, r);
}

/* ------------------------------------------------------------------------ */
/* src/tape.rs:226-273 Header::new                                           */
/* ------------------------------------------------------------------------ */
static int is_rust_ascii_ws(uint8_t c)
{
    /* str::trim strips Unicode White_Space; for the ASCII range that is
     * U+0009..U+000D and U+0020. */
    return c == 0x20 || (c >= 0x09 && c <= 0x0D);
}

int oracle_header_new(const uint8_t *bytes, size_t n, oracle_header *h)
{
    size_t end = 0;
    while (end < n && bytes[end] != 0x0D && bytes[end] != 0x0A) ++end; /* :228-232 */
    if (end + 1 >= n) return ORACLE_ERR_PANIC; /* memmap[header_end_idx + 1] :236 */
    h->crlf = bytes[end + 1] == 0x0A;          /* :235-238 */
    size_t start = 0;                          /* :241-249 BOM skip */
    while (start < n && (bytes[start] == 0xEF || bytes[start] == 0xBB || bytes[start] == 0xBF)) ++start;
    if (start > end) return ORACLE_ERR_PANIC;  /* &memmap[start..end] :253 */
    uint32_t fields = 1;                       /* split(",") :259-262 */
    for (size_t i = start; i < end; ++i)
        if (bytes[i] == 0x2C) ++fields;
    h->field_cnt = fields;                     /* :264 */
    h->record_offset = (uint32_t)end;          /* :271 */
    h->header_start = start;
    h->header_end = end;
    return 0;
}

/* name i of the header: split(",").map(trim) (src/tape.rs:259-262) */
int oracle_header_name(const uint8_t *bytes, const oracle_header *h, uint32_t i,
                       size_t *name_start, size_t *name_end)
{
    if (i >= h->field_cnt) return ORACLE_ERR_PANIC;
    size_t s = h->header_start;
    uint32_t k = 0;
    for (size_t p = h->header_start; p <= h->header_end; ++p) {
        if (p == h->header_end || bytes[p] == 0x2C) {
            if (k == i) {
                size_t a = s, b = p;
                while (a < b && is_rust_ascii_ws(bytes[a])) ++a;
                while (b > a && is_rust_ascii_ws(bytes[b - 1])) --b;
                *name_start = a;
                *name_end = b;
                return 0;
            }
            ++k;
            s = p + 1;
        }
    }
    return ORACLE_ERR_PANIC;
}

/* src/tape.rs:315-347 TapeCore::init */
int oracle_tape_init(size_t index_len, uint32_t field_cnt, int crlf,
                     uint64_t *jump, uint32_t *record_cnt)
{
    const uint64_t j = crlf ? (uint64_t)field_cnt + 1 : (uint64_t)field_cnt; /* :318-321 */
    if (j == 0 || index_len == 0) return ORACLE_ERR_PANIC;
    *jump = j;
    *record_cnt = (uint32_t)((index_len - 1) / j);      /* :323-325 */
    const uint64_t problem = (index_len - 1) % j;       /* :327 */
    if (problem != 0) return ORACLE_ERR_INVALID_CSV_FORMAT; /* :342-344 */
    return 0;
}

/* src/record_source.rs:70-102 seek_record (u32 wrapping arithmetic as in a
 * release build; out-of-bounds index / inverted slice = Rust panic). */
int oracle_seek_record(const uint64_t *index, size_t index_len, size_t data_len,
                       uint32_t record_cnt, uint64_t jump, uint32_t field_cnt,
                       uint32_t record_idx, uint64_t *start, uint64_t *end, int *found)
{
    *found = 0;
    if ((uint32_t)(record_idx + 1u) >= record_cnt) return 0;            /* :77-81 */
    const uint32_t idx_start = (uint32_t)(record_idx + 1u) * (uint32_t)jump; /* :83 */
    const size_t a = (size_t)idx_start, b = (size_t)idx_start + (size_t)field_cnt;
    if (a >= index_len || b >= index_len) return ORACLE_ERR_PANIC;      /* :94-95 */
    const uint64_t ms = index[a] + 1, me = index[b];
    if (ms > me || me > data_len) return ORACLE_ERR_PANIC;              /* :99 */
    *start = ms;
    *end = me;
    *found = 1;
    return 0;
}

/* src/record_source.rs:104-140 seek_field */
int oracle_seek_field(const uint64_t *index, size_t index_len, size_t data_len,
                      uint32_t record_cnt, uint32_t field_cnt, int crlf,
                      uint32_t record_idx, uint32_t field_idx,
                      uint64_t *start, uint64_t *end, int *found)
{
    *found = 0;
    if ((uint32_t)(record_idx + 1u) >= record_cnt) return 0;            /* :112-116 */
    if (field_idx >= field_cnt) return 0;                               /* :117-119 */
    const uint32_t row_size = crlf ? field_cnt + 1u : field_cnt;        /* :123-126 */
    const uint32_t idx_start = (uint32_t)(record_idx + 1u) * row_size + field_idx; /* :129 */
    const size_t a = (size_t)idx_start;
    if (a + 1 >= index_len) return ORACLE_ERR_PANIC;                    /* :132-133 */
    const uint64_t ms = index[a] + 1, me = index[a + 1];
    if (ms > me || me > data_len) return ORACLE_ERR_PANIC;              /* :137 */
    *start = ms;
    *end = me;
    *found = 1;
    return 0;
}

/* Timed CPU baseline for the batched-lookup config: the scalar seek_field above over nq queries,
 * single thread, no printing.  Returns a checksum (sum of start ^ end over the hits) and the hit count. */
int oracle_seek_fields_timed(const uint64_t *index, size_t index_len, size_t data_len,
                             uint32_t record_cnt, uint32_t field_cnt, int crlf,
                             const uint32_t *rec, const uint32_t *fld, size_t nq,
                             uint64_t *checksum, uint64_t *hits)
{
    uint64_t cs = 0, h = 0;
    for (size_t i = 0; i < nq; ++i) {
        uint64_t s, e;
        int found;
        const int rc = oracle_seek_field(index, index_len, data_len, record_cnt, field_cnt, crlf,
                                         rec[i], fld[i], &s, &e, &found);
        if (rc) return rc;
        if (found) {
            cs += s ^ (e << 1);
            ++h;
        }
    }
    *checksum = cs;
    *hits = h;
    return 0;
}

/* src/tape.rs:385-428 boundaries(task_size: u32, job_count: u8).
 * Returns the number of boundaries written (0 = None). out must hold 255. */
int oracle_boundaries(uint32_t task_size, uint8_t job_count, oracle_boundary *out)
{
    if (task_size == 0 || job_count == 0) return 0;                     /* :387-389 */
    if (task_size < (uint32_t)job_count) {                              /* :390-395 */
        out[0].start = 0;
        out[0].len = task_size;
        return 1;
    }
    const uint32_t job_size = task_size / (uint32_t)job_count;          /* :405 */
    const uint32_t remainder = task_size % (uint32_t)job_count;         /* :406 */
    uint32_t acc_end = 0, share_remainder = 1;
    for (uint8_t i = 0; i < job_count; ++i) {                           /* :412-421 */
        if (share_remainder == 1 && i >= (uint8_t)remainder) share_remainder = 0;
        out[i].start = acc_end;
        out[i].len = job_size + share_remainder;
        acc_end += job_size + share_remainder;
    }
    return (int)job_count;
}

/* src/tape.rs:95-140 Tape::chunks: boundaries * jump, chunk 0 patched to skip
 * the header row.  Returns chunk count, 0 => Err(InvalidState). */
int oracle_chunks(uint32_t record_cnt, uint64_t jump, uint8_t num, oracle_chunk *out)
{
    oracle_boundary b[256];
    const int nb = oracle_boundaries(record_cnt, num, b);
    if (nb == 0) return 0;
    for (int i = 0; i < nb; ++i) {
        out[i].id = (uint8_t)i;
        out[i].start = b[i].start * jump;                               /* :105-107 */
        out[i].end = (b[i].start + b[i].len) * jump;                    /* :108-111 */
        out[i].record_cnt = (uint32_t)b[i].len;                         /* :112 */
    }
    out[0].start = jump;                                                /* :117-123 */
    out[0].record_cnt = out[0].record_cnt - 1;
    return nb;
}

/* src/lib.rs:139-152 blsr identity: x & (x - 1) clears the lowest set bit. */
/* src/reader.rs:26-132 is_ascii ("Non-core", word-at-a-time): scalar loop for short inputs (:54-61),
 * first word read unaligned (:72-77), aligned words up to len - 8 (:94-118), last word read unaligned
 * (:130-134).  usize = u64. */
int oracle_is_ascii(const uint8_t *s, size_t len)
{
    const uint64_t mask = 0x8080808080808080ull;            /* :29-31 */
    const size_t W = 8;
    size_t align_offset = (size_t)((W - ((uintptr_t)s & (W - 1))) & (W - 1));
    if (len < W || len < align_offset) {                    /* :54-61 */
        for (size_t i = 0; i < len; ++i)
            if (s[i] >= 128) return 0;
        return 1;
    }
    const size_t offset_to_aligned = align_offset == 0 ? W : align_offset;   /* :65-69 */
    uint64_t w;
    memcpy(&w, s, W);                                       /* :74 read_unaligned */
    if (w & mask) return 0;
    size_t byte_pos = offset_to_aligned;
    while (byte_pos + W <= len) {                           /* :94 byte_pos <= len - USIZE_SIZE */
        memcpy(&w, s + byte_pos, W);
        if (w & mask) return 0;
        byte_pos += W;
    }
    if (byte_pos == len) return 1;                          /* :122-124 */
    memcpy(&w, s + len - W, W);                             /* :130-132 */
    return (w & mask) == 0;
}

/* ---- definitions BEYOND the reference (SURVEY 8f "next" rows) -------------------------------------
 * The reference has no counterpart for these, so there is nothing to pin them to: they are scalar
 * statements of the definitions in include/csvb200.h, used to check the CUDA kernels.  "parity unpinned". */

/* csvb200_tape_validate: first slot s >= 1 whose separator does not fit k = (s-1) % jump. */
uint64_t oracle_tape_first_bad_slot(const uint8_t *bytes, size_t n, const uint64_t *index,
                                    size_t index_len, uint32_t field_cnt, int crlf)
{
    const uint64_t jump = crlf ? (uint64_t)field_cnt + 1 : (uint64_t)field_cnt;
    for (size_t s = 1; s < index_len; ++s) {
        const uint64_t k = (s - 1) % jump, pos = index[s];
        int ok = pos < n;
        if (ok) {
            const uint8_t b = bytes[pos];
            if (crlf) {
                if (k + 2 < jump) ok = b == ',';
                else if (k + 2 == jump) ok = b == '\r';
                else ok = b == '\n' && index[s - 1] + 1 == pos;
            } else {
                ok = k + 1 < jump ? b == ',' : (b == '\n' || b == '\r');
            }
        }
        if (!ok) return s;
    }
    return UINT64_MAX;
}

/* csvb200_materialize_*: the value of a raw field slice (src/record_source.rs:135-139 returns the raw
 * slice) after optional trim (ASCII space / tab at both ends) and optional RFC-4180 unquoting (outer
 * quotes stripped when both present, "" -> ").  Returns the length; writes to out when non-null. */
size_t oracle_field_value(const uint8_t *raw, size_t len, uint32_t flags, uint8_t *out)
{
    size_t a = 0, b = len, o = 0;
    if (flags & 2u) {
        while (a < b && (raw[a] == ' ' || raw[a] == '\t')) ++a;
        while (b > a && (raw[b - 1] == ' ' || raw[b - 1] == '\t')) --b;
    }
    if ((flags & 1u) && b - a >= 2 && raw[a] == '"' && raw[b - 1] == '"') {
        ++a;
        --b;
        while (a < b) {
            if (raw[a] == '"' && a + 1 < b && raw[a + 1] == '"') ++a;
            if (out) out[o] = raw[a];
            ++o;
            ++a;
        }
        return o;
    }
    if (out) memcpy(out, raw + a, b - a);
    return b - a;
}

/* whole column through the seek_field restatement + oracle_field_value; offsets[nrec + 1]; out may be
 * NULL (sizing pass).  Returns the total length. */
uint64_t oracle_materialize_column(const uint8_t *bytes, size_t n, const uint64_t *index, size_t index_len,
                                   uint32_t record_cnt, uint32_t field_cnt, int crlf, uint32_t field_idx,
                                   uint32_t first_record, uint32_t nrec, uint32_t flags,
                                   uint64_t *offsets, uint8_t *out)
{
    uint64_t acc = 0;
    for (uint32_t i = 0; i < nrec; ++i) {
        uint64_t a = 0, b = 0;
        int found = 0;
        offsets[i] = acc;
        if (oracle_seek_field(index, index_len, n, record_cnt, field_cnt, crlf, first_record + i, field_idx,
                              &a, &b, &found) == 0 && found && a <= b && b <= n)
            acc += oracle_field_value(bytes + a, (size_t)(b - a), flags, out ? out + acc : NULL);
    }
    offsets[nrec] = acc;
    return acc;
}

uint64_t oracle_blsr(uint64_t x) { return x & (x ? x - 1 : 0); }
