/* csv_oracle.h -- TEST INFRASTRUCTURE ONLY. See csv_oracle.c. */
#ifndef CSV_ORACLE_H
#define CSV_ORACLE_H
#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

enum {
    ORACLE_OK = 0,
    ORACLE_ERR_PANIC = 1,              /* the reference would panic / hit UB here */
    ORACLE_ERR_UNALIGNED = 2,          /* input not 16-byte aligned (align_to head non-empty) */
    ORACLE_ERR_OOM = 3,
    ORACLE_ERR_INVALID_CSV_FORMAT = 4, /* StructureError::InvalidCsvFormat */
};

typedef struct {
    uint32_t field_cnt;
    uint32_t record_offset;
    int crlf;
    size_t header_start;
    size_t header_end;
} oracle_header;

typedef struct {
    uint64_t start;
    uint64_t len;
} oracle_boundary;

typedef struct {
    uint8_t id;
    uint64_t start;
    uint64_t end;
    uint32_t record_cnt;
} oracle_chunk;

int oracle_read_sse(const uint8_t *bytes, size_t n, uint64_t **out, size_t *out_len);
int oracle_read_sse_timed(const uint8_t *bytes, size_t n, size_t *out_len, uint64_t *checksum);
int oracle_read_closed_form(const uint8_t *bytes, size_t n, int start_parity,
                            uint64_t pos_bias, int with_sentinel,
                            uint64_t **out, size_t *out_len, int *end_parity);
void oracle_shard_summary(const uint8_t *bytes, size_t n, uint64_t *parity,
                          uint64_t *c0, uint64_t *s);
void oracle_free(void *p);
void oracle_structure_run(const uint8_t *chunk, size_t at, uint8_t out16[16]);
int oracle_header_new(const uint8_t *bytes, size_t n, oracle_header *h);
int oracle_header_name(const uint8_t *bytes, const oracle_header *h, uint32_t i,
                       size_t *name_start, size_t *name_end);
int oracle_tape_init(size_t index_len, uint32_t field_cnt, int crlf,
                     uint64_t *jump, uint32_t *record_cnt);
int oracle_seek_record(const uint64_t *index, size_t index_len, size_t data_len,
                       uint32_t record_cnt, uint64_t jump, uint32_t field_cnt,
                       uint32_t record_idx, uint64_t *start, uint64_t *end, int *found);
int oracle_seek_field(const uint64_t *index, size_t index_len, size_t data_len,
                      uint32_t record_cnt, uint32_t field_cnt, int crlf,
                      uint32_t record_idx, uint32_t field_idx,
                      uint64_t *start, uint64_t *end, int *found);
int oracle_seek_fields_timed(const uint64_t *index, size_t index_len, size_t data_len,
                             uint32_t record_cnt, uint32_t field_cnt, int crlf,
                             const uint32_t *rec, const uint32_t *fld, size_t nq,
                             uint64_t *checksum, uint64_t *hits);
int oracle_boundaries(uint32_t task_size, uint8_t job_count, oracle_boundary *out);
int oracle_chunks(uint32_t record_cnt, uint64_t jump, uint8_t num, oracle_chunk *out);
int oracle_is_ascii(const uint8_t *s, size_t len);
/* definitions beyond the reference (see csv_oracle.c) */
uint64_t oracle_tape_first_bad_slot(const uint8_t *bytes, size_t n, const uint64_t *index,
                                    size_t index_len, uint32_t field_cnt, int crlf);
size_t oracle_field_value(const uint8_t *raw, size_t len, uint32_t flags, uint8_t *out);
uint64_t oracle_materialize_column(const uint8_t *bytes, size_t n, const uint64_t *index, size_t index_len,
                                   uint32_t record_cnt, uint32_t field_cnt, int crlf, uint32_t field_idx,
                                   uint32_t first_record, uint32_t nrec, uint32_t flags,
                                   uint64_t *offsets, uint8_t *out);
uint64_t oracle_blsr(uint64_t x);

#ifdef __cplusplus
}
#endif
#endif
