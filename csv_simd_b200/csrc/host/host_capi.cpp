// host_capi.cpp -- flat C shim over the C++ host mirror (csv_simd.hpp) so the pytest parity suite can
// drive the SAME code a C++ user of the drop-in would call.  Status: 0 ok, 1 Io, 2 MissingValue,
// 3 InvalidState, 4 InvalidCsvFormat, 5 ReferencePanic, 6 Gpu, 7 other.
#include <cstring>
#include <string>

#include "csv_simd.hpp"

using namespace csv_simd;

namespace {
thread_local std::string g_err;
int to_status(const StructureError& e)
{
    g_err = e.what();
    switch (e.kind) {
    case ErrorKind::Io: return 1;
    case ErrorKind::MissingValue: return 2;
    case ErrorKind::InvalidState: return 3;
    case ErrorKind::InvalidCsvFormat: return 4;
    case ErrorKind::ReferencePanic: return 5;
    case ErrorKind::Gpu: return 6;
    }
    return 7;
}
template <class F>
int guard(F&& f)
{
    try {
        f();
        return 0;
    } catch (const StructureError& e) {
        return to_status(e);
    } catch (const std::out_of_range& e) {
        g_err = e.what();
        return 5;  // Vec / slice bounds check = reference panic
    } catch (const std::exception& e) {
        g_err = e.what();
        return 7;
    }
}
struct CoreBox {
    TapeCore core;
};
}  // namespace

extern "C" {

const char* csvsimd_last_error() { return g_err.c_str(); }

// reader::read_multi on a file: returns the entry count; dst may be NULL to size the destination
int csvsimd_read_multi(const char* filename, const int* devices, int ndev, uint64_t* dst, size_t cap, size_t* len_out)
{
    return guard([&] {
        const Mmap mm = Mmap::map(filename);
        const std::vector<int> devs(devices, devices + ndev);
        const std::vector<uint64_t> idx = reader::read_multi(mm, devs);
        *len_out = idx.size();
        if (dst) {
            if (idx.size() > cap) throw std::out_of_range("destination too small");
            std::memcpy(dst, idx.data(), idx.size() * sizeof(uint64_t));
        }
    });
}

int csvsimd_create(const char* filename, void** tape_out)
{
    return guard([&] { *tape_out = new Tape(create(filename)); });
}
void csvsimd_tape_free(void* t) { delete static_cast<Tape*>(t); }
uint32_t csvsimd_tape_record_cnt(void* t) { return static_cast<Tape*>(t)->record_cnt_value; }
uint64_t csvsimd_tape_jump(void* t) { return static_cast<Tape*>(t)->record_jump; }
uint32_t csvsimd_tape_field_cnt(void* t) { return static_cast<Tape*>(t)->field_cnt(); }
int csvsimd_tape_is_crlf(void* t) { return static_cast<Tape*>(t)->new_line_tag() == NewLine::CRLF; }
uint32_t csvsimd_tape_record_offset(void* t) { return static_cast<Tape*>(t)->header_info.record_offset; }
size_t csvsimd_tape_index_len(void* t) { return static_cast<Tape*>(t)->index().len(); }
int csvsimd_tape_index_copy(void* t, uint64_t* dst, size_t cap)
{
    return guard([&] {
        const auto& h = static_cast<Tape*>(t)->index().host();
        if (h.size() > cap) throw std::out_of_range("destination too small");
        std::memcpy(dst, h.data(), h.size() * sizeof(uint64_t));
    });
}
int csvsimd_tape_header_name(void* t, uint32_t i, char* dst, size_t cap)
{
    return guard([&] {
        const std::string& s = static_cast<Tape*>(t)->header().at(i);
        if (s.size() + 1 > cap) throw std::out_of_range("destination too small");
        std::memcpy(dst, s.c_str(), s.size() + 1);
    });
}
// found: 1 = Some (start/len relative to the file), 0 = None
int csvsimd_tape_seek_record(void* t, uint32_t r, uint64_t* start, uint64_t* len, int* found)
{
    return guard([&] {
        const Tape* tp = static_cast<Tape*>(t);
        const auto v = tp->seek_record(r);
        *found = v.has_value();
        if (v) {
            *start = static_cast<uint64_t>(v->data() - tp->data_bytes().data());
            *len = v->size();
        }
    });
}
int csvsimd_tape_seek_field(void* t, uint32_t r, uint32_t f, uint64_t* start, uint64_t* len, int* found)
{
    return guard([&] {
        const Tape* tp = static_cast<Tape*>(t);
        const auto v = tp->seek_field(r, f);
        *found = v.has_value();
        if (v) {
            *start = static_cast<uint64_t>(v->data() - tp->data_bytes().data());
            *len = v->size();
        }
    });
}
int csvsimd_tape_seek_fields(void* t, const uint32_t* rec, const uint32_t* fld, size_t nq, csvb200_range* out)
{
    return guard([&] {
        const std::vector<uint32_t> r(rec, rec + nq), f(fld, fld + nq);
        const auto v = static_cast<Tape*>(t)->seek_fields(r, f);
        std::memcpy(out, v.data(), nq * sizeof(csvb200_range));
    });
}
// chunks: writes up to cap entries {id, start, end, record_cnt}; returns count through n_out
int csvsimd_tape_chunks(void* t, uint8_t num, uint64_t* out4, size_t cap, size_t* n_out)
{
    return guard([&] {
        const auto ch = static_cast<Tape*>(t)->chunks(num);
        *n_out = ch.size();
        for (size_t i = 0; i < ch.size() && i < cap; ++i) {
            out4[4 * i + 0] = ch[i].id;
            out4[4 * i + 1] = ch[i].start;
            out4[4 * i + 2] = ch[i].end;
            out4[4 * i + 3] = ch[i].record_cnt;
        }
    });
}
int csvsimd_tape_validate(void* t, csvb200_tape_report* out)
{
    return guard([&] { *out = static_cast<Tape*>(t)->validate(); });
}
// column as packed bytes: offsets[nrec + 1]; returns the total through total_out, copies min(total, cap) bytes
int csvsimd_tape_column(void* t, uint32_t field_idx, uint32_t first_record, uint32_t nrec, uint32_t flags, uint64_t* offsets,
                        uint8_t* dst, size_t cap, size_t* total_out)
{
    return guard([&] {
        const Tape::Column c = static_cast<Tape*>(t)->column(field_idx, first_record, nrec, flags);
        std::memcpy(offsets, c.offsets.data(), c.offsets.size() * sizeof(uint64_t));
        *total_out = c.bytes.size();
        std::memcpy(dst, c.bytes.data(), c.bytes.size() < cap ? c.bytes.size() : cap);
    });
}
int csvsimd_tape_utf8(void* t, uint64_t* valid_up_to, int* well_formed)
{
    return guard([&] {
        const auto v = static_cast<Tape*>(t)->utf8_valid_up_to();
        *well_formed = v ? 0 : 1;
        *valid_up_to = v ? *v : 0;
    });
}
// boundaries (no GPU needed): returns count (0 = None)
int csvsimd_boundaries(uint32_t task_size, uint8_t job_count, uint64_t* out2, size_t cap)
{
    const auto b = boundaries(task_size, job_count);
    if (!b) return 0;
    for (size_t i = 0; i < b->size() && i < cap; ++i) {
        out2[2 * i] = (*b)[i].start;
        out2[2 * i + 1] = (*b)[i].len;
    }
    return static_cast<int>(b->size());
}
// Header::new on a caller buffer (no GPU needed)
int csvsimd_header(const uint8_t* data, size_t n, uint32_t* field_cnt, uint32_t* record_offset, int* crlf)
{
    return guard([&] {
        const Mmap m = Mmap::borrow(data, n);
        const Header h = Header::make(m);
        *field_cnt = h.field_cnt;
        *record_offset = h.record_offset;
        *crlf = h.new_line == NewLine::CRLF;
    });
}
// TapeCore before init: seek must fail with InvalidState (record_source.rs:77-79)
int csvsimd_core_seek_before_init(const char* filename)
{
    return guard([&] {
        Mmap memmap = Mmap::map(filename);
        Header header = Header::make(memmap);
        StructureIndex index = reader::read(memmap);
        TapeCore core = TapeCore::create(std::move(memmap), std::move(index), std::move(header));
        (void)core.seek_record(0);
    });
}

}  // extern "C"
