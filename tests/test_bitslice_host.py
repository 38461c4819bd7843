"""CPU: the bit-sliced classifier (csrc/bitslice.cuh) is __host__ __device__; compile it with g++
and check it against direct byte tests for every byte value at every lane position."""
import os
import subprocess
import tempfile

from tests.conftest import ROOT

HARNESS = r"""
#include "bitslice.cuh"
#include <cstdio>
#include <cstdlib>
#include <cstring>
using namespace csvb200;
int main() {
    srand(1);
    long bad = 0;
    for (int it = 0; it < 40000; ++it) {
        uint8_t b[32];
        for (int i = 0; i < 32; ++i) b[i] = (rand() % 10 < 6) ? (uint8_t)rand() : (uint8_t)("\",\r\n ,\\\"\n"[rand() % 9]);
        if (it < 256 * 32) b[it % 32] = (uint8_t)(it / 32);   // every value at every position
        uint32_t w[8];
        memcpy(w, b, 32);
        const Masks32 m = classify32(w);
        uint32_t q = 0, s = 0;
        for (int i = 0; i < 32; ++i) {
            if (b[i] == 0x22) q |= 1u << i;
            if (b[i] == 0x2c || b[i] == 0x0d || b[i] == 0x0a) s |= 1u << i;
        }
        bad += (q != m.quote) + (s != m.sep);
    }
    for (int it = 0; it < 2000; ++it) {
        const uint32_t m = (uint32_t)rand() * 2654435761u;
        uint32_t e = 0, p = 0;
        for (int i = 0; i < 32; ++i) { p ^= (m >> i) & 1; e |= p << i; }
        bad += prefix_xor32(m) != e;
    }
    // LUT classes of the reference (src/stage1.rs:24-35) over all 256 bytes
    const uint8_t LO[16] = {4, 0, 16, 0, 0, 0, 0, 0, 0, 0, 1, 0, 10, 1, 0, 0};
    const uint8_t HI[16] = {1, 0, 22, 0, 0, 8, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0};
    for (int v = 0; v < 256; ++v) bad += class_byte((uint8_t)v) != (LO[v & 15] & HI[v >> 4]);
    printf("bad=%ld\n", bad);
    return bad != 0;
}
"""


def test_bitslice_classifier_on_host():
    with tempfile.TemporaryDirectory() as d:
        src = os.path.join(d, "t.cpp")
        exe = os.path.join(d, "t")
        open(src, "w").write(HARNESS)
        subprocess.check_call(["g++", "-O2", "-I", os.path.join(ROOT, "csv_simd_b200", "csrc"), "-x", "c++", src,
                               "-o", exe])
        out = subprocess.run([exe], capture_output=True, text=True)
        assert out.returncode == 0 and "bad=0" in out.stdout, out.stdout + out.stderr


UTF8_HARNESS = r"""
#include "utf8slice.cuh"
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <vector>
using namespace csvb200;

// scalar statement of Unicode table 3-7 (what core::str::from_utf8 checks): start of the first ill-formed
// sequence, or -1
static long scalar_valid_up_to(const uint8_t* s, long n)
{
    long i = 0;
    while (i < n) {
        const uint8_t b = s[i];
        if (b < 0x80) { ++i; continue; }
        int len; uint8_t lo = 0x80, hi = 0xBF;
        if (b >= 0xC2 && b <= 0xDF) len = 2;
        else if (b >= 0xE0 && b <= 0xEF) { len = 3; if (b == 0xE0) lo = 0xA0; if (b == 0xED) hi = 0x9F; }
        else if (b >= 0xF0 && b <= 0xF4) { len = 4; if (b == 0xF0) lo = 0x90; if (b == 0xF4) hi = 0x8F; }
        else return i;
        if (i + 1 >= n || s[i + 1] < lo || s[i + 1] > hi) return i;
        for (int k = 2; k < len; ++k)
            if (i + k >= n || (s[i + k] & 0xC0) != 0x80) return i;
        i += len;
    }
    return -1;
}

// the kernel's decomposition: 32-byte groups, each judged with utf8_check32 + its halos, minimum of the starts
static long sliced_valid_up_to(const uint8_t* s, long n)
{
    long best = -1;
    for (long i0 = 0; i0 < n; i0 += 32) {
        uint8_t g[32] = {0};
        memcpy(g, s + i0, (size_t)((n - i0) < 32 ? (n - i0) : 32));
        uint32_t w[8];
        memcpy(w, g, 32);
        const uint32_t b1 = i0 >= 1 ? s[i0 - 1] : 0, b2 = i0 >= 2 ? s[i0 - 2] : 0, b3 = i0 >= 3 ? s[i0 - 3] : 0;
        const uint32_t n0 = i0 + 32 < n ? s[i0 + 32] : 0x100, n1 = i0 + 33 < n ? s[i0 + 33] : 0x100,
                       n2 = i0 + 34 < n ? s[i0 + 34] : 0x100;
        const uint32_t r = utf8_check32(w, utf8_owed(b1, b2, b3), n0, n1, n2);
        if (r != kUtf8None && (best < 0 || i0 + r < best)) best = i0 + r;
    }
    return best;
}

int main() {
    srand(7);
    long bad = 0, checked = 0;
    const uint8_t alpha[] = {0x41, 0x2C, 0x0A, 0x7F, 0x80, 0x8F, 0x90, 0x9F, 0xA0, 0xBF, 0xC0, 0xC1, 0xC2, 0xDF, 0xE0, 0xE1,
                             0xEC, 0xED, 0xEE, 0xEF, 0xF0, 0xF1, 0xF3, 0xF4, 0xF5, 0xF8, 0xFF};
    const int na = (int)sizeof(alpha);
    // exhaustive over 1..3 byte strings from the tricky alphabet at every offset around the group boundary
    for (int len = 1; len <= 4; ++len) {
        long combos = 1;
        for (int k = 0; k < len; ++k) combos *= na;
        for (long c = 0; c < combos; ++c) {
            uint8_t seq[4];
            long t = c;
            for (int k = 0; k < len; ++k) { seq[k] = alpha[t % na]; t /= na; }
            for (int pad = 27; pad <= 33; ++pad) {
                for (int tail = 0; tail <= 1; ++tail) {
                    std::vector<uint8_t> s((size_t)pad, 'x');
                    s.insert(s.end(), seq, seq + len);
                    if (tail) s.insert(s.end(), 5, 'y');
                    const long a = scalar_valid_up_to(s.data(), (long)s.size()), b = sliced_valid_up_to(s.data(), (long)s.size());
                    if (a != b && bad++ < 5) {
                        printf("mismatch len=%d pad=%d tail=%d scalar=%ld sliced=%ld:", len, pad, tail, a, b);
                        for (int k = 0; k < len; ++k) printf(" %02x", seq[k]);
                        printf("\n");
                    }
                    ++checked;
                }
            }
        }
    }
    // random mixtures, all lengths
    for (int it = 0; it < 200000; ++it) {
        const int n = 1 + rand() % 150;
        std::vector<uint8_t> s((size_t)n);
        const int mode = rand() % 3;
        for (int i = 0; i < n; ++i)
            s[(size_t)i] = mode == 0 ? (uint8_t)rand() : (rand() % 4 ? 'a' : alpha[rand() % na]);
        const long a = scalar_valid_up_to(s.data(), n), b = sliced_valid_up_to(s.data(), n);
        if (a != b && bad++ < 5) printf("random mismatch n=%d scalar=%ld sliced=%ld\n", n, a, b);
        ++checked;
    }
    // well-formed text of every sequence length stays well-formed at every alignment
    const char* txt = "a\xc3\xa9\xe2\x82\xac\xf0\x9f\x99\x82z\xed\x9f\xbf\xee\x80\x80\xf4\x8f\xbf\xbf\xe0\xa0\x80\xf0\x90\x80\x80";
    for (int pad = 0; pad < 70; ++pad) {
        std::vector<uint8_t> s((size_t)pad, 'x');
        for (int r = 0; r < 6; ++r) s.insert(s.end(), (const uint8_t*)txt, (const uint8_t*)txt + strlen(txt));
        if (scalar_valid_up_to(s.data(), (long)s.size()) != -1 || sliced_valid_up_to(s.data(), (long)s.size()) != -1) ++bad;
        ++checked;
    }
    printf("checked=%ld bad=%ld\n", checked, bad);
    return bad != 0;
}
"""


def test_utf8_slice_on_host():
    """utf8slice.cuh (the K7 kernel's rule engine) on the CPU against the scalar table 3-7 rule: exhaustive
    over short strings from the boundary-value alphabet at every offset around a 32-byte group edge, plus
    random mixtures.  (tests/test_oracle.py pins the same scalar rule to CPython's decoder.)"""
    with tempfile.TemporaryDirectory() as d:
        src = os.path.join(d, "u.cpp")
        with open(src, "w") as f:
            f.write(UTF8_HARNESS)
        exe = os.path.join(d, "u")
        subprocess.check_call(["g++", "-O2", "-std=c++17", "-I", os.path.join(ROOT, "csv_simd_b200", "csrc"), src, "-o", exe])
        out = subprocess.run([exe], capture_output=True, text=True)
        assert out.returncode == 0, out.stdout[-2000:]


UTF8_SO = r"""
#include "utf8slice.cuh"
#include <cstring>
using namespace csvb200;
// the kernel's decomposition on the host: 32-byte groups, utf8_check32 + halos, minimum of the starts (-1 = well-formed)
extern "C" long sliced_valid_up_to(const uint8_t* s, long n)
{
    long best = -1;
    for (long i0 = 0; i0 < n; i0 += 32) {
        uint8_t g[32] = {0};
        memcpy(g, s + i0, (size_t)((n - i0) < 32 ? (n - i0) : 32));
        uint32_t w[8];
        memcpy(w, g, 32);
        const uint32_t b1 = i0 >= 1 ? s[i0 - 1] : 0, b2 = i0 >= 2 ? s[i0 - 2] : 0, b3 = i0 >= 3 ? s[i0 - 3] : 0;
        const uint32_t n0 = i0 + 32 < n ? s[i0 + 32] : 0x100, n1 = i0 + 33 < n ? s[i0 + 33] : 0x100,
                       n2 = i0 + 34 < n ? s[i0 + 34] : 0x100;
        const uint32_t r = utf8_check32(w, utf8_owed(b1, b2, b3), n0, n1, n2);
        if (r != kUtf8None && (best < 0 || i0 + r < best)) best = i0 + r;
    }
    return best;
}
"""


def test_utf8_slice_against_cpython_decoder():
    """The same rule engine, directly against CPython's strict UTF-8 decoder (= core::str::from_utf8's table and
    error position) on random and crafted inputs -- pins K7's logic on the CPU, no GPU involved."""
    import ctypes as C
    import random
    with tempfile.TemporaryDirectory() as d:
        src = os.path.join(d, "u.cpp")
        open(src, "w").write(UTF8_SO)
        so = os.path.join(d, "libu.so")
        subprocess.check_call(["g++", "-O2", "-std=c++17", "-fPIC", "-shared", "-I", os.path.join(ROOT, "csv_simd_b200", "csrc"),
                               src, "-o", so])
        L = C.CDLL(so)
        L.sliced_valid_up_to.argtypes = [C.c_char_p, C.c_long]
        L.sliced_valid_up_to.restype = C.c_long

        def want(b):
            try:
                b.decode("utf-8")
                return -1
            except UnicodeDecodeError as e:
                return e.start

        rnd = random.Random(3)
        tricky = [0x41, 0x80, 0x8F, 0x90, 0x9F, 0xA0, 0xBF, 0xC0, 0xC1, 0xC2, 0xDF, 0xE0, 0xE1, 0xED, 0xEE, 0xEF, 0xF0, 0xF1,
                  0xF4, 0xF5, 0xFF]
        texts = ["naïve café", "日本語のテキスト", "🙂🙃 emoji", "αβγ,δεζ\n", "퟿\U0010ffff\U00010000ࠀ߿"]
        n_bad = 0
        for it in range(20000):
            mode = it % 4
            if mode == 0:
                b = bytes(rnd.choice(tricky) if rnd.random() < 0.3 else 0x61 for _ in range(rnd.randrange(1, 120)))
            elif mode == 1:
                b = bytes(rnd.randrange(256) for _ in range(rnd.randrange(1, 80)))
            elif mode == 2:       # valid text with one byte damaged
                t = bytearray(("x" * rnd.randrange(0, 40) + rnd.choice(texts) * rnd.randrange(1, 4)).encode())
                t[rnd.randrange(len(t))] = rnd.choice(tricky)
                b = bytes(t)
            else:                 # valid text, possibly truncated mid-sequence
                t = ("y" * rnd.randrange(0, 40) + rnd.choice(texts) * rnd.randrange(1, 4)).encode()
                b = t[:rnd.randrange(1, len(t) + 1)]
            got = L.sliced_valid_up_to(b, len(b))
            if got != want(b):
                n_bad += 1
                assert n_bad < 3, (b, got, want(b))
        assert n_bad == 0
