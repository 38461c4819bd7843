// bitslice.cuh -- bit-sliced byte classification for the CSV structural indexer.
//
// Replaces the reference's nibble-LUT classify + movemask bit-pack
// (src/avx/stage1.rs:249-316 classify, :111-187 get_struct_positions, LUTs at
// src/stage1.rs:24-35).  Enumerating the two LUTs over all 256 byte values,
// class(b) = LO[b & 15] & HI[b >> 4] is non-zero for exactly six bytes; the two
// masks the live path uses are
//     quote (search = 16): b == 0x22
//     struct (search = 3): b in {0x2C ',', 0x0D CR, 0x0A LF}
// so the SSE sequence is equivalent to two set-membership tests per byte.
//
// B200 formulation: one thread owns 32 contiguous bytes (8 x u32).  Instead of
// testing bytes one at a time (>= 4 ALU ops per byte with SWAR compares plus a
// movemask emulation), the 32 bytes are transposed into 8 bit-planes
// (plane b, bit i = bit b of byte i) with 16 PRMT + 3 delta-swap rounds
// (48 SHF/LOP3), after which BOTH 32-bit masks fall out of 8 LOP3s in natural
// bit order (bit i <-> byte i), i.e. ~2.25 integer ops per input byte total.
//
// Everything here is __host__ __device__ so tests/ can exercise the exact same
// code on the CPU (tests/test_bitslice_host.py builds a tiny g++ harness).
#pragma once
#include <stdint.h>

#if defined(__CUDACC__)
#define CSVB_HD __host__ __device__ __forceinline__
#else
#define CSVB_HD inline
#endif

namespace csvb200 {

// PRMT (default mode): bytes 0-3 = a, bytes 4-7 = b, one selector nibble per output byte.
CSVB_HD uint32_t prmt(uint32_t a, uint32_t b, uint32_t sel)
{
#if defined(__CUDA_ARCH__)
    return __byte_perm(a, b, sel);
#else
    const uint64_t src = ((uint64_t)b << 32) | a;
    uint32_t r = 0;
    for (int i = 0; i < 4; ++i) {
        const uint32_t s = (sel >> (4 * i)) & 0x7u;
        r |= (uint32_t)((src >> (8 * s)) & 0xFFu) << (8 * i);
    }
    return r;
#endif
}

// count trailing / leading zeros of a non-zero word
CSVB_HD int ctz32(uint32_t v)
{
#if defined(__CUDA_ARCH__)
    return __ffs((int)v) - 1;
#else
    return __builtin_ctz(v);
#endif
}
CSVB_HD int clz32(uint32_t v)
{
#if defined(__CUDA_ARCH__)
    return __clz((int)v);
#else
    return __builtin_clz(v);
#endif
}

struct Masks32 {
    uint32_t quote;  // bit i set <=> byte i == '"'
    uint32_t sep;    // bit i set <=> byte i in {',', CR, LF}
};

// bitselect: (a & m) | (b & ~m) as ONE LOP3 (immLut 0xE4).  Written in PTX because
// nvcc otherwise treats m and ~m as two unrelated immediates and emits two LOP3s.
CSVB_HD uint32_t bitselect(uint32_t m, uint32_t a, uint32_t b)
{
#if defined(__CUDA_ARCH__)
    uint32_t d;
    asm("lop3.b32 %0, %1, %2, %3, 0xE4;" : "=r"(d) : "r"(a), "r"(b), "r"(m));
    return d;
#else
    return (a & m) | (b & ~m);
#endif
}

// (a, b) -> a' = (a & M) | ((b << S) & ~M),  b' = ((a >> S) & M) | (b & ~M)
// One delta-swap step of the 8x8 bit-matrix transpose across a register pair:
// 2 shifts + 2 LOP3.
template <int S, uint32_t M>
CSVB_HD void delta_swap_pair(uint32_t& a, uint32_t& b)
{
    const uint32_t na = bitselect(M, a, b << S);
    const uint32_t nb = bitselect(M, a >> S, b);
    a = na;
    b = nb;
}

// w[0..7]: the 32 bytes in memory order (little-endian words).  x[b] <- bit-plane b in natural order:
// bit i of x[b] = bit b of byte i.  16 PRMT + 3 delta-swap rounds.
CSVB_HD void bitplanes32(const uint32_t w[8], uint32_t x[8])
{
    // 1) byte transpose: x[j] byte k = input byte 8k + j   (two 4x4 byte transposes)
#pragma unroll
    for (int e = 0; e < 2; ++e) {
        const uint32_t A = w[e], B = w[2 + e], C = w[4 + e], D = w[6 + e];
        const uint32_t t0 = prmt(A, B, 0x5140), t1 = prmt(A, B, 0x7362);
        const uint32_t t2 = prmt(C, D, 0x5140), t3 = prmt(C, D, 0x7362);
        x[4 * e + 0] = prmt(t0, t2, 0x5410);
        x[4 * e + 1] = prmt(t0, t2, 0x7632);
        x[4 * e + 2] = prmt(t1, t3, 0x5410);
        x[4 * e + 3] = prmt(t1, t3, 0x7632);
    }
    // 2) 8x8 bit transpose inside every byte lane: afterwards x[b] bit (8k + j)
    //    = bit b of input byte 8k + j, i.e. x[b] is bit-plane b in natural order.
    delta_swap_pair<4, 0x0F0F0F0Fu>(x[0], x[4]);
    delta_swap_pair<4, 0x0F0F0F0Fu>(x[1], x[5]);
    delta_swap_pair<4, 0x0F0F0F0Fu>(x[2], x[6]);
    delta_swap_pair<4, 0x0F0F0F0Fu>(x[3], x[7]);
    delta_swap_pair<2, 0x33333333u>(x[0], x[2]);
    delta_swap_pair<2, 0x33333333u>(x[1], x[3]);
    delta_swap_pair<2, 0x33333333u>(x[4], x[6]);
    delta_swap_pair<2, 0x33333333u>(x[5], x[7]);
    delta_swap_pair<1, 0x55555555u>(x[0], x[1]);
    delta_swap_pair<1, 0x55555555u>(x[2], x[3]);
    delta_swap_pair<1, 0x55555555u>(x[4], x[5]);
    delta_swap_pair<1, 0x55555555u>(x[6], x[7]);
}

// classify32 plus the two by-products the planes give away for one more LOP3 each (CSVB200_BUILD_VALIDATE):
//   nl : CR / LF (separators with bit 5 clear; ',' = 0x2C has it set)
//   hi : bytes >= 0x80 (plane 7) -- is_ascii of the reference (src/reader.rs:26-132) is "no such byte"
struct Masks32x {
    uint32_t quote, sep, nl, hi;
};

CSVB_HD Masks32x classify32x(const uint32_t w[8])
{
    uint32_t x[8];
    bitplanes32(w, x);
    const uint32_t P0 = x[0], P1 = x[1], P2 = x[2], P3 = x[3];
    const uint32_t P4 = x[4], P5 = x[5], P6 = x[6], P7 = x[7];
    const uint32_t c = ~(P7 | P6 | P4);
    const uint32_t u = (P5 ^ P0) & P2;
    const uint32_t v = ~(P5 | P2 | P0);
    const uint32_t m = (u & ~P1) | (v & P1);
    Masks32x r;
    r.sep = c & P3 & m;
    r.quote = c & ~P3 & (P5 & ~P2 & ~P0) & P1;
    r.nl = r.sep & ~P5;
    r.hi = P7;
    return r;
}

CSVB_HD Masks32 classify32(const uint32_t w[8])
{
    uint32_t x[8];
    bitplanes32(w, x);
    const uint32_t P0 = x[0], P1 = x[1], P2 = x[2], P3 = x[3];
    const uint32_t P4 = x[4], P5 = x[5], P6 = x[6], P7 = x[7];
    // 3) boolean membership on the planes (ptxas fuses these into 8 LOP3):
    //    0x0A = 0000 1010, 0x0D = 0000 1101, 0x2C = 0010 1100, 0x22 = 0010 0010
    const uint32_t c = ~(P7 | P6 | P4);          // bits 7,6,4 clear in all four
    const uint32_t u = (P5 ^ P0) & P2;           // CR (P5=0,P0=1) or ',' (P5=1,P0=0), needs P1=0
    const uint32_t v = ~(P5 | P2 | P0);          // LF, needs P1=1
    const uint32_t m = (u & ~P1) | (v & P1);
    Masks32 r;
    r.sep = c & P3 & m;
    r.quote = c & ~P3 & (P5 & ~P2 & ~P0) & P1;
    return r;
}

// Inclusive prefix-XOR of a 32-bit word (bit i = XOR of bits 0..i): the
// per-word part of the reference's clmul(quote_bits, ~0)
// (src/avx/stage1.rs:342-361).
CSVB_HD uint32_t prefix_xor32(uint32_t m)
{
    m ^= m << 1;
    m ^= m << 2;
    m ^= m << 4;
    m ^= m << 8;
    m ^= m << 16;
    return m;
}

// Class byte of the reference LUTs for one input byte (debug / known-answer
// export for src/structure.rs:10-58): newline=1, comma=2, space=4,
// backslash=8, quote=16.
CSVB_HD uint8_t class_byte(uint8_t b)
{
    return b == 0x0A || b == 0x0D ? 1 : b == 0x2C ? 2 : b == 0x20 ? 4 : b == 0x5C ? 8 : b == 0x22 ? 16 : 0;
}

}  // namespace csvb200
