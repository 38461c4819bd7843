// utf8slice.cuh -- bit-sliced UTF-8 well-formedness check of 32 bytes (K7, validate.cu).
//
// Answers what core::str::from_utf8 answers (the reference builds &str with from_utf8_unchecked,
// src/record_source.rs:97-101, and carries a dead SIMD checker, src/avx/utf8check.rs): is the input
// well-formed, and if not, where does the first ill-formed sequence START (Utf8Error::valid_up_to).
//
// Every rule of Unicode table 3-7 is local to a 4-byte window, so once the 32 bytes are bit-planes
// (bitslice.cuh) the rules are a handful of 32-bit boolean operations:
//   C   continuation 10xxxxxx            L2/L3/L4  valid leads C2..DF / E0..EF / F0..F4
//   BAD C0, C1, F5..FF                    X  = positions that MUST be continuations
//                                             (L234 << 1 | L34 << 2 | L4 << 3, plus the bytes owed by a
//                                              sequence that began in the 3 bytes before this group)
//   stray = C & ~X    miss = X & ~C       special = E0/ED/F0/F4 followed by a second byte outside
//                                             A0..BF / 80..9F / 90..BF / 80..8F
// Each thread judges the sequences whose LEAD lies in its own 32 bytes, looking 3 bytes ahead; the
// first `owed` positions belong to a lead in the previous group and are judged there.
// __host__ __device__: tests/test_bitslice_host.py runs the same code on the CPU against the scalar rule.
#pragma once
#include "bitslice.cuh"

namespace csvb200 {

constexpr uint32_t kUtf8None = 0xffffffffu;

CSVB_HD bool utf8_is_cont(uint32_t b) { return (b & 0xC0u) == 0x80u; }
// sequence length a VALID lead byte announces, 0 otherwise (ASCII, continuation, C0, C1, F5..FF)
CSVB_HD uint32_t utf8_lead_len(uint32_t b)
{
    return (b >= 0xC2u && b <= 0xDFu) ? 2u : (b >= 0xE0u && b <= 0xEFu) ? 3u : (b >= 0xF0u && b <= 0xF4u) ? 4u : 0u;
}

// continuation bytes a sequence begun in the 3 bytes BEFORE a group still owes at the group's start.
// b1 = byte just before the group, b2 and b3 the ones before that (0 where the input starts).
CSVB_HD uint32_t utf8_owed(uint32_t b1, uint32_t b2, uint32_t b3)
{
    const uint32_t l1 = utf8_lead_len(b1), l2 = utf8_lead_len(b2), l3 = utf8_lead_len(b3);
    if (l1 >= 2u) return l1 - 1u;
    if (utf8_is_cont(b1)) {
        if (l2 >= 3u) return l2 - 2u;
        if (utf8_is_cont(b2) && l3 == 4u) return 1u;
    }
    return 0u;
}

// w: the group's 32 bytes (zero past the end of the input); owed: utf8_owed of the 3 bytes before;
// n0, n1, n2: the 3 bytes after the group (0x100 = past the end of the input).
// Returns the offset (0..31) at which the first ill-formed sequence led from this group starts, or kUtf8None.
CSVB_HD uint32_t utf8_check32(const uint32_t w[8], uint32_t owed, uint32_t n0, uint32_t n1, uint32_t n2)
{
    uint32_t P[8];
    bitplanes32(w, P);
    const uint32_t hi = P[7] & P[6];
    const uint32_t C = P[7] & ~P[6];
    const uint32_t L2r = hi & ~P[5];                                 // C0..DF
    const uint32_t L3 = hi & P[5] & ~P[4];                           // E0..EF
    const uint32_t L4r = hi & P[5] & P[4] & ~P[3];                   // F0..F7
    const uint32_t low3z = ~(P[2] | P[1] | P[0]);
    const uint32_t badC = L2r & ~(P[4] | P[3] | P[2] | P[1]);        // C0, C1
    const uint32_t L4 = L4r & (~P[2] | (P[2] & ~P[1] & ~P[0]));      // F0..F4
    const uint32_t BAD = badC | (L4r & ~L4) | (hi & P[5] & P[4] & P[3]);   // C0 C1 F5..F7 F8..FF
    const uint32_t L2 = L2r & ~badC;
    const uint32_t L34 = L3 | L4, L234 = L2 | L34;
    const uint32_t carry = (1u << owed) - 1u;                        // owed <= 3
    const uint32_t X = (L234 << 1) | (L34 << 2) | (L4 << 3) | carry;
    const uint32_t stray = C & ~X;
    const uint32_t miss = X & ~C & ~carry;
    // second-byte ranges: E0 needs A0..BF, ED 80..9F, F0 90..BF, F4 80..8F
    const uint32_t E0 = L3 & ~P[3] & low3z, ED = L3 & P[3] & P[2] & ~P[1] & P[0];
    const uint32_t F0 = L4 & low3z, F4 = L4 & P[2];
    const uint32_t sE0 = C & ~P[5], sED = C & P[5], sF0 = C & ~P[5] & ~P[4], sF4 = C & (P[5] | P[4]);
    // the byte after position 31 is n0
    const uint32_t c0 = n0 < 0x100u && utf8_is_cont(n0) ? 1u : 0u;
    const uint32_t c1 = n1 < 0x100u && utf8_is_cont(n1) ? 1u : 0u;
    const uint32_t c2 = n2 < 0x100u && utf8_is_cont(n2) ? 1u : 0u;
    const uint32_t top = 0x80000000u;
    const uint32_t nE0 = (c0 && !(n0 & 0x20u)) ? top : 0u, nED = (c0 && (n0 & 0x20u)) ? top : 0u;
    const uint32_t nF0 = (c0 && !(n0 & 0x30u)) ? top : 0u, nF4 = (c0 && (n0 & 0x30u)) ? top : 0u;
    const uint32_t special = (E0 & ((sE0 >> 1) | nE0)) | (ED & ((sED >> 1) | nED)) | (F0 & ((sF0 >> 1) | nF0)) |
                             (F4 & ((sF4 >> 1) | nF4));
    uint32_t best = kUtf8None;
    const uint32_t here = BAD | special | stray;                     // these start where they are seen
    if (here) best = (uint32_t)ctz32(here);
    if (miss) {
        const uint32_t p = (uint32_t)ctz32(miss);                    // first missing continuation: its lead is the
        const uint32_t q = 31u - (uint32_t)clz32(L234 & ((1u << p) - 1u));   // nearest lead below it
        best = q < best ? q : best;
    }
    // continuations owed past the end of the group: positions 32, 33, 34
    const uint32_t XN = (L234 >> 31) | ((L34 >> 30) & 3u) | ((L4 >> 29) & 7u);
    const uint32_t NC = c0 | (c1 << 1) | (c2 << 2);
    if (XN & ~NC) {
        const uint32_t q = 31u - (uint32_t)clz32(L234);              // the last lead of the group is the one cut short
        best = q < best ? q : best;
    }
    return best;
}

}  // namespace csvb200
