"""Host-side copy helpers of the end-to-end path (csrc/slice_pool.h), compiled with g++ alone: the non-temporal copy and
the sliced parallel copy must be byte-exact for every alignment and length, including the sub-threshold fall-backs."""
import os
import subprocess
import tempfile

SRC = r'''
#include "slice_pool.h"
#include <cstdio>
#include <cstdint>
#include <vector>
using namespace csvb200;
int main()
{
    std::vector<uint8_t> src((9u << 20) + 512), dst(src.size() + 512), ref(dst.size());
    uint32_t x = 12345;
    for (auto& b : src) { x = x * 1664525u + 1013904223u; b = (uint8_t)(x >> 24); }
    const size_t lens[] = {0, 1, 63, 64, 65, 4095, 4096, 4097, 65536 + 17, (4u << 20) - 1, (4u << 20), (9u << 20) + 77};
    SlicePool pool(5);
    int bad = 0;
    for (size_t len : lens)
        for (size_t so = 0; so < 67; so += 11)
            for (size_t dof = 0; dof < 67; dof += 13) {
                if (so + len > src.size() || dof + len + 64 > dst.size()) continue;
                for (int mode = 0; mode < 3; ++mode) {
                    std::fill(dst.begin(), dst.end(), 0xEE);
                    std::fill(ref.begin(), ref.end(), 0xEE);
                    std::memcpy(ref.data() + dof, src.data() + so, len);
                    if (mode == 0) stream_memcpy(dst.data() + dof, src.data() + so, len);
                    if (mode == 1) parallel_memcpy(pool, dst.data() + dof, src.data() + so, len, CopyDir::ToStaging);
                    if (mode == 2) parallel_memcpy(pool, dst.data() + dof, src.data() + so, len, CopyDir::ToCaller);
                    if (dst != ref) { ++bad; std::printf("mismatch len=%zu so=%zu do=%zu mode=%d\n", len, so, dof, mode); }
                }
            }
    std::printf("bad=%d threads=%d\n", bad, default_io_threads());
    return bad != 0;
}
'''


def test_stream_and_parallel_memcpy_are_exact():
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    with tempfile.TemporaryDirectory() as td:
        src = os.path.join(td, "t.cpp")
        exe = os.path.join(td, "t")
        open(src, "w").write(SRC)
        subprocess.run(["g++", "-O2", "-std=c++17", "-pthread", "-I", os.path.join(root, "csv_simd_b200", "csrc"), src, "-o", exe],
                       check=True)
        out = subprocess.run([exe], capture_output=True, text=True, timeout=300)
        assert out.returncode == 0, out.stdout[-2000:] + out.stderr[-2000:]
        assert "bad=0" in out.stdout
