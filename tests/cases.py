"""Deterministic test inputs shared by the oracle (CPU) and the CUDA parity tests."""
from __future__ import annotations

import numpy as np

ALPHABET = np.frombuffer(b'ab1,"\r\n \\\x00\xff', dtype=np.uint8)
TILE = 32768  # kTileBytes of the CUDA kernel (also exercises multiples of the older 16 KiB tile)


def rand_bytes(n: int, seed: int, weights=None) -> bytes:
    rng = np.random.default_rng(seed)
    if weights is None:
        return ALPHABET[rng.integers(0, ALPHABET.size, size=n)].tobytes()
    p = np.asarray(weights, dtype=np.float64)
    return ALPHABET[rng.choice(ALPHABET.size, size=n, p=p / p.sum())].tobytes()


def full_random(n: int, seed: int) -> bytes:
    return np.random.default_rng(seed).integers(0, 256, size=n, dtype=np.uint8).tobytes()


def edge_cases():
    """(name, bytes) pairs: the edge semantics of SURVEY.md 7 'hard parts' 3 and 8(c)."""
    out = []
    for n in (64, 65, 79, 80, 127, 128, 129, 191, 192, 255, 256, 1000):
        out.append((f"rand{n}", rand_bytes(n, 100 + n)))
    out.append(("all_quotes_200", b'"' * 200))
    out.append(("all_commas_200", b"," * 200))
    out.append(("all_lf_130", b"\n" * 130))
    out.append(("all_crlf_128", b"\r\n" * 64))
    out.append(("no_structure_300", b"abcdefghij" * 30))
    out.append(("high_bytes_256", bytes(range(256))))
    out.append(("zeros_100", b"\x00" * 100))
    # quote exactly at byte 63 / 64 (block carry of the reference's inside_str)
    for qpos in (62, 63, 64, 65, 127, 128):
        b = bytearray(b"a,b,c,d," * 32)
        b[qpos] = 0x22
        out.append((f"quote_at_{qpos}", bytes(b)))
        b[qpos + 40] = 0x22
        out.append((f"quote_pair_from_{qpos}", bytes(b)))
    # "" escape straddling a 64-byte block boundary and a 32-byte word boundary
    b = bytearray(b"x" * 200)
    b[10] = 0x22
    b[63] = 0x22
    b[64] = 0x22
    b[70] = 0x2C
    b[100] = 0x22
    b[101] = 0x2C
    out.append(("escape_across_block", bytes(b)))
    b = bytearray(b"y" * 96)
    b[31] = 0x22
    b[32] = 0x22
    b[33] = 0x0A
    out.append(("escape_across_word", bytes(b)))
    # CR | LF split across a block boundary
    b = bytearray(b"f1,f2" * 40)
    b[63] = 0x0D
    b[64] = 0x0A
    out.append(("crlf_across_block", bytes(b)))
    # tile-boundary cases for the CUDA kernel (16 KiB tiles)
    for n in (TILE - 1, TILE, TILE + 1, 2 * TILE - 1, 2 * TILE, 2 * TILE + 17, 5 * TILE + 3, 8 * TILE, 9 * TILE + 129,
              12 * TILE - 1):
        out.append((f"tile_rand_{n}", rand_bytes(n, 7 * n)))
    for qpos in (TILE - 1, TILE, TILE + 1):
        b = bytearray(b"ab,cd\n" * ((3 * TILE) // 6))
        b[qpos] = 0x22
        out.append((f"open_quote_at_{qpos}", bytes(b)))
        b[qpos + TILE] = 0x22
        out.append((f"quote_span_tile_{qpos}", bytes(b)))
    b = bytearray(b"q,r\r\n" * ((2 * TILE) // 5 + 8))
    b[TILE - 1] = 0x22
    b[TILE] = 0x22
    out.append(("escape_across_tile", bytes(b)))
    # worst-case density: one entry per byte over several tiles
    out.append(("all_commas_3tiles", b"," * (3 * TILE + 5)))
    out.append(("all_quotes_2tiles", b'"' * (2 * TILE + 1)))
    out.append(("alt_quote_comma", b'",' * (TILE + 3)))
    return out


def small_cases():
    """Inputs shorter than 64 bytes: the reference panics; the C ABI returns the closed form."""
    return [(f"small{n}", rand_bytes(n, 900 + n)) for n in (0, 1, 2, 15, 16, 17, 31, 32, 33, 63)]
