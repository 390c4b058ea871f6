"""The element order inside a packed 8 x 8 tile (csrc/common.cuh: magi_tile_slot / magi_tile_pos), compiled for the
host with nvcc and checked for the two properties the fast posterior path relies on once a tile sits in shared memory:
it is a bijection, and both MMA fragment shapes are bank-conflict-free reads (32 banks of 4 bytes; a 16-byte load is
served a quarter warp at a time, an 8-byte load a half warp at a time)."""
import ctypes
import os
import shutil
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def lib(tmp_path_factory):
    nvcc = shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"
    if not os.path.exists(nvcc):
        pytest.skip("nvcc not available")
    so = str(tmp_path_factory.mktemp("tile") / "tile_host.so")
    subprocess.run([nvcc, "-O1", "-shared", "-Xcompiler", "-fPIC", "-o", so,
                    os.path.join(ROOT, "tests", "harness", "tile_host.cu")], check=True)
    return ctypes.CDLL(so)


def test_tile_order_is_a_bijection(lib):
    pos = sorted(lib.tile_pos_host(r, c) for r in range(8) for c in range(8))
    assert pos == list(range(64))
    # a pair (r, 2cp), (r, 2cp+1) stays 16 contiguous, 16-byte aligned bytes
    for r in range(8):
        for cp in range(4):
            assert lib.tile_pos_host(r, 2 * cp) == 2 * lib.tile_slot_host(r, cp)
            assert lib.tile_pos_host(r, 2 * cp + 1) == 2 * lib.tile_slot_host(r, cp) + 1


def test_forward_fragment_is_conflict_free(lib):
    """lane 4g+c loads the pair (g, c) with one 16-byte access: every quarter warp covers all 32 banks once."""
    for q in range(4):
        banks = []
        for lane in range(8 * q, 8 * q + 8):
            g, c = lane >> 2, lane & 3
            w0 = 2 * lib.tile_pos_host(g, 2 * c)          # first 4-byte word of the pair
            banks += [(w0 + k) % 32 for k in range(4)]
        assert sorted(banks) == list(range(32))


def test_transposed_fragment_is_conflict_free(lib):
    """lane 4g+c loads elements (2c+h, g), h = 0, 1, with two 8-byte accesses: every half warp covers all 32 banks once
    (the plain row-major tile is a 4-way conflict here)."""
    for h in range(2):
        for half in range(2):
            banks = []
            for lane in range(16 * half, 16 * half + 16):
                g, c = lane >> 2, lane & 3
                w0 = 2 * lib.tile_pos_host(2 * c + h, g)
                banks += [w0 % 32, (w0 + 1) % 32]
            assert sorted(banks) == list(range(32))
    # the row-major order it replaced: 4-way
    banks = [(2 * ((2 * (lane & 3)) * 8 + (lane >> 2))) % 32 for lane in range(16)]
    assert len(set(banks)) == 4
