/*
 * gen_csv.c -- deterministic synthetic CSV generators for tests and bench.py
 * (bench/test tooling; not part of the product library).
 *
 * Row r is a pure function of (seed, r): its RNG is a splitmix64 stream whose
 * state starts at mix64(seed, r), so any party (CPU thread, another rank) that
 * generates row r produces the same bytes.  Grammars follow SURVEY.md 8(d):
 *
 *   unquoted (cfg2 / cfg5): header "c0,...,c{F-1}\n"; every field is the decimal
 *     of (next() % modulus); F-1 commas + "\n" per row.
 *   quoted (cfg3 / cfg4): header "c0,...,c15\r\n"; 16 fields per row, rows end
 *     "\r\n"; even fields numeric (next() % 10^6); odd fields are quoted, body of
 *     8..40 tokens, each token: letter 70 %, "," 10 %, "\n" 5 %, "\r\n" 5 %,
 *     "\"\"" 10 %.
 *
 * Whole rows are appended until the size reaches `target` bytes, so the output
 * always ends with its row terminator.
 */
#include <stddef.h>
#include <stdint.h>
#include <string.h>

static inline uint64_t splitmix64(uint64_t *s)
{
    uint64_t z = (*s += 0x9E3779B97F4A7C15ull);
    z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
    z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
    return z ^ (z >> 31);
}

static inline uint64_t row_state(uint64_t seed, uint64_t row)
{
    /* hash (seed, row) through two splitmix64 finalisers so that consecutive
     * rows do not land on shifted copies of the same stream */
    uint64_t a = seed ^ 0x2545F4914F6CDD1Dull;
    const uint64_t h = splitmix64(&a);
    uint64_t b = (row + 1) * 0xD1342543DE82EF95ull ^ h;
    return splitmix64(&b);
}

static inline size_t put_u64(uint8_t *dst, uint64_t v)
{
    uint8_t tmp[24];
    size_t n = 0;
    do {
        tmp[n++] = (uint8_t)('0' + v % 10);
        v /= 10;
    } while (v);
    for (size_t i = 0; i < n; ++i) dst[i] = tmp[n - 1 - i];
    return n;
}

static size_t put_header(uint8_t *dst, uint32_t nfields, int crlf)
{
    size_t p = 0;
    for (uint32_t f = 0; f < nfields; ++f) {
        if (f) dst[p++] = ',';
        dst[p++] = 'c';
        p += put_u64(dst + p, f);
    }
    if (crlf) dst[p++] = '\r';
    dst[p++] = '\n';
    return p;
}

/* Upper bound of one generated row (for caller-side buffer slack). */
size_t gen_max_row_bytes(uint32_t nfields, int quoted)
{
    if (quoted) return (size_t)nfields * (2 + 40 * 2 + 1) + 16;
    return (size_t)nfields * 21 + 16;
}

/* Returns bytes written.  cap must be >= target + gen_max_row_bytes() + header. */
size_t gen_unquoted(uint64_t seed, uint32_t nfields, uint64_t modulus, uint64_t first_row,
                    int with_header, size_t target, uint8_t *out, size_t cap, uint64_t *rows_out)
{
    size_t p = 0;
    uint64_t rows = 0;
    const size_t maxrow = gen_max_row_bytes(nfields, 0);
    if (with_header) p += put_header(out, nfields, 0);
    while (p < target && p + maxrow <= cap) {
        uint64_t st = row_state(seed, first_row + rows);
        for (uint32_t f = 0; f < nfields; ++f) {
            if (f) out[p++] = ',';
            p += put_u64(out + p, splitmix64(&st) % modulus);
        }
        out[p++] = '\n';
        ++rows;
    }
    if (rows_out) *rows_out = rows;
    return p;
}

size_t gen_quoted(uint64_t seed, uint64_t first_row, int with_header, size_t target,
                  uint8_t *out, size_t cap, uint64_t *rows_out)
{
    const uint32_t nfields = 16;
    size_t p = 0;
    uint64_t rows = 0;
    const size_t maxrow = gen_max_row_bytes(nfields, 1);
    if (with_header) p += put_header(out, nfields, 1);
    while (p < target && p + maxrow <= cap) {
        uint64_t st = row_state(seed, first_row + rows);
        for (uint32_t f = 0; f < nfields; ++f) {
            if (f) out[p++] = ',';
            if ((f & 1) == 0) {
                p += put_u64(out + p, splitmix64(&st) % 1000000ull);
            } else {
                out[p++] = '"';
                const uint32_t ntok = 8 + (uint32_t)(splitmix64(&st) % 33);
                for (uint32_t t = 0; t < ntok; ++t) {
                    const uint64_t r = splitmix64(&st);
                    const uint32_t pct = (uint32_t)(r % 100);
                    if (pct < 70) out[p++] = (uint8_t)('a' + (r >> 32) % 26);
                    else if (pct < 80) out[p++] = ',';
                    else if (pct < 85) out[p++] = '\n';
                    else if (pct < 90) { out[p++] = '\r'; out[p++] = '\n'; }
                    else { out[p++] = '"'; out[p++] = '"'; }
                }
                out[p++] = '"';
            }
        }
        out[p++] = '\r';
        out[p++] = '\n';
        ++rows;
    }
    if (rows_out) *rows_out = rows;
    return p;
}

/* 10 M lookup queries of cfg5: rec = next % (record_cnt - 1), fld = next % field_cnt */
void gen_queries(uint64_t seed, uint64_t nq, uint32_t record_cnt, uint32_t field_cnt,
                 uint32_t *rec, uint32_t *fld)
{
    uint64_t st = row_state(seed, 0);
    const uint64_t rmod = record_cnt > 1 ? (uint64_t)record_cnt - 1 : 1;
    for (uint64_t i = 0; i < nq; ++i) {
        rec[i] = (uint32_t)(splitmix64(&st) % rmod);
        fld[i] = (uint32_t)(splitmix64(&st) % field_cnt);
    }
}
