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
