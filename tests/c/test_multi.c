/* test_multi.c -- a plain C caller (no Python, no torch) of the multi-GPU C ABI: one byte buffer in, one contiguous
 * index out, checked entry by entry against the oracle (tests link liboracle: test infrastructure).
 *
 *   test_multi <ndev> [device list repeats device 0 when fewer GPUs are visible] [bytes]
 *
 * The input is a deterministic quote-heavy CSV (embedded commas, CRLF, newlines and "" escapes inside quoted fields)
 * generated here, so that the program has no file dependencies. */
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include "csv_oracle.h"
#include "csvb200.h"

static uint64_t sm64(uint64_t* s)
{
    uint64_t z = (*s += 0x9E3779B97F4A7C15ull);
    z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
    z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
    return z ^ (z >> 31);
}

static size_t gen(uint8_t* out, size_t target)
{
    uint64_t s = 20261018;
    size_t n = 0;
    while (n + 1024 < target) {
        for (int f = 0; f < 8; ++f) {
            if (f & 1) {
                out[n++] = '"';
                int len = 8 + (int)(sm64(&s) % 33);
                for (int i = 0; i < len; ++i) {
                    unsigned t = (unsigned)(sm64(&s) % 100);
                    if (t < 70) out[n++] = (uint8_t)('a' + t % 26);
                    else if (t < 80) out[n++] = ',';
                    else if (t < 85) out[n++] = '\n';
                    else if (t < 90) { out[n++] = '\r'; out[n++] = '\n'; }
                    else { out[n++] = '"'; out[n++] = '"'; }
                }
                out[n++] = '"';
            } else {
                n += (size_t)sprintf((char*)out + n, "%u", (unsigned)(sm64(&s) % 1000000));
            }
            out[n++] = f == 7 ? '\r' : ',';
        }
        out[n++] = '\n';
    }
    return n;
}

int main(int argc, char** argv)
{
    int ndev = argc > 1 ? atoi(argv[1]) : 2;
    size_t target = argc > 2 ? (size_t)atoll(argv[2]) : (size_t)64 << 20;
    int devices[16];
    if (ndev < 1 || ndev > 16) return 2;
    /* CSVB200_TEST_DEVICES="0,0,1": explicit list; default 0..ndev-1 */
    const char* env = getenv("CSVB200_TEST_DEVICES");
    for (int k = 0; k < ndev; ++k) devices[k] = k;
    if (env) {
        int k = 0;
        for (const char* p = env; *p && k < ndev; ++k) {
            devices[k] = atoi(p);
            while (*p && *p != ',') ++p;
            if (*p == ',') ++p;
        }
    }
    uint8_t* bytes = NULL;
    if (posix_memalign((void**)&bytes, 64, target + 4096)) return 2;
    const size_t n = gen(bytes, target);

    uint64_t* want = NULL;
    size_t want_len = 0;
    int end_parity = 0;
    if (oracle_read_closed_form(bytes, n, 0, 0, 1, &want, &want_len, &end_parity) != ORACLE_OK) return 2;

    csvb200_multi* m = NULL;
    int rc = csvb200_multi_create(devices, ndev, &m);
    if (rc) {
        fprintf(stderr, "csvb200_multi_create: %s\n", csvb200_status_string(rc));
        return 1;
    }
    uint64_t* got = (uint64_t*)malloc((want_len + 16) * sizeof(uint64_t));
    size_t len = 0;
    int bad = 0;
    for (int rep = 0; rep < 3 && !bad; ++rep) {
        memset(got, 0xee, (want_len + 16) * sizeof(uint64_t));
        rc = csvb200_multi_index_build_to_host(m, bytes, n, NULL, got, want_len + 16, &len);
        if (rc) {
            fprintf(stderr, "csvb200_multi_index_build_to_host: %s: %s\n", csvb200_status_string(rc), csvb200_multi_last_error(m));
            return 1;
        }
        if (len != want_len || memcmp(got, want, want_len * sizeof(uint64_t)) != 0) bad = 1;
    }
    csvb200_multi_stats st;
    csvb200_multi_last_stats(m, &st);
    printf("{\"ndev\": %d, \"bytes\": %zu, \"entries\": %zu, \"equal_oracle\": %s, \"seconds\": %.4f, \"gbs\": %.2f, "
           "\"upload_s\": %.4f, \"download_s\": %.4f, \"carry_mask\": %u, \"redone_mask\": %u}\n",
           ndev, n, len, bad ? "false" : "true", st.seconds, (double)n / st.seconds / 1e9, st.upload_seconds,
           st.download_seconds, st.carry_mask, st.redone_mask);
    /* too small a destination is reported, with the needed size */
    size_t need = 0;
    rc = csvb200_multi_index_build_to_host(m, bytes, n, NULL, got, 10, &need);
    if (rc != CSVB200_ERR_CAPACITY || need != want_len) bad = 1;
    csvb200_multi_destroy(m);
    oracle_free(want);
    free(got);
    free(bytes);
    return bad;
}
