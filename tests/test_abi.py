"""CPU checks of the C-ABI: the library builds for sm_100a, loads, and exports exactly the entry points
include/concepthash_b200.h declares (no compute calls without a GPU)."""
import ctypes
import os
import re
import subprocess

import pytest
import torch

from concepthash_b200 import _lib, build

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def header_functions():
    text = open(os.path.join(ROOT, "include", "concepthash_b200.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return set(re.findall(r"\b(ch_[a-z0-9_]+)\s*\(", text))


@pytest.fixture(scope="module")
def lib():
    build.build()
    return _lib.load()


def test_header_and_binding_agree(lib):
    declared = header_functions()
    assert declared == set(_lib.SIGNATURES), declared ^ set(_lib.SIGNATURES)
    for name in declared:
        assert hasattr(lib, name), name


def test_exported_symbols_match_header(lib):
    out = subprocess.run(["nm", "-D", "--defined-only", _lib.LIB_PATH], capture_output=True, text=True).stdout
    exported = set(re.findall(r"\b(ch_[a-z0-9_]+)$", out, flags=re.M))
    exported = {s for s in exported if not s.startswith("ch_ws_") and s != "ch_set_error"}
    assert header_functions() <= exported


def test_pure_host_entry_points(lib):
    assert lib.ch_abi_version() == 5
    assert lib.ch_padded_rows(0) == 64 and lib.ch_padded_rows(1) == 128 and lib.ch_padded_rows(64) == 128
    assert [lib.ch_code_words(n) for n in (1, 16, 32, 33, 64, 65, 128, 129, 256, 257, 0)] == \
           [1, 1, 1, 2, 2, 4, 4, 8, 8, 0, 0]


def test_struct_layout_matches_c(lib):
    # sizes as the C compiler lays them out (pointers 8, int64 8, int32 4, natural alignment)
    assert ctypes.sizeof(_lib.HistArgs) == 14 * 8 + 3 * 8 + 8 * 4 + 8 + 4 + 4     # ..., int64 row_base, int32 + padding
    assert ctypes.sizeof(_lib.FinalArgs) == 10 * 8 + 2 * 8 + 5 * 4 + 4 + 8 * 8 + 32 * 8
    assert ctypes.sizeof(_lib.SelectArgs) == 8 * 8 + 4 * 8 + 6 * 4 + 8 + 8      # ..., int32 pair, bad, int64 q_stripe_bytes
    assert ctypes.sizeof(_lib.CandArgs) == 22 * 8 + 4 * 8 + 9 * 4 + 4 + 8 * 8 + 32 * 8


def test_sass_is_blackwell_native():
    out = subprocess.run(["cuobjdump", "-sass", _lib.LIB_PATH], capture_output=True, text=True).stdout
    assert "sm_100a" in out
    assert "UBLKCP" in out          # cp.async.bulk (TMA engine) streaming of the gallery tiles
    assert "POPC" in out
    assert "UTCIMMA" in out         # tcgen05.mma kind::i8 (select pass)
    assert "LDTM" in out            # tcgen05.ld (TMEM -> registers) in its epilogue


@pytest.mark.skipif(torch.cuda.is_available(), reason="CPU-only behaviour")
def test_fails_loudly_without_gpu():
    from concepthash_b200 import hashing
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        hashing.calculate_mAP(torch.randn(4, 8), torch.eye(4), torch.randn(2, 8), torch.eye(4)[:2], -1)
    ws = ctypes.c_void_p()
    assert lib_call_fails(ws)


def lib_call_fails(ws):
    lib = _lib.load()
    rc = lib.ch_workspace_create(0, ctypes.byref(ws))
    return rc != 0 and len(lib.ch_last_error()) > 0


@pytest.mark.parametrize("isa", ["avx2", "plain"])
def test_host_pack_narrower_isa_variants(lib, isa, monkeypatch):
    """The packer picks AVX-512 on the B200 hosts; CH_HOST_PACK_ISA narrows it so that the AVX2 and scalar variants
    (what a host without AVX-512 runs) are held to the same bits and flags: every width around the vector and word
    boundaries, strided rows, +-0.0, NaN, row counts that are no multiple of anything."""
    import numpy as np
    rng = np.random.RandomState(7)
    for nbit in (1, 7, 8, 9, 15, 16, 17, 31, 32, 33, 48, 63, 64, 65, 100, 128, 129, 200, 255, 256):
        n, stride = 613, nbit + 3
        buf = rng.randn(n, stride).astype(np.float32)
        x = buf[:, :nbit]
        x[2, 0] = 0.0
        x[4, nbit - 1] = -0.0
        x[6, nbit // 2] = np.float32(1e-45)
        words = lib.ch_code_words(nbit)
        for nan in (False, True):
            if nan:
                x[n - 1, nbit - 1] = np.nan
            want = np.zeros((n, words * 32), dtype=np.uint8)
            want[:, :nbit] = x > 0
            want = np.packbits(want, axis=1, bitorder="little").view(np.uint32)
            res = {}
            for which in (None, isa):
                if which is None:
                    monkeypatch.delenv("CH_HOST_PACK_ISA", raising=False)
                else:
                    monkeypatch.setenv("CH_HOST_PACK_ISA", which)
                out = np.full((n, words), 0xDEADBEEF, dtype=np.uint32)
                flags = ctypes.c_uint32(0)
                assert lib.ch_host_pack_sign(buf.ctypes.data, n, nbit, stride, out.ctypes.data, ctypes.byref(flags), 3) == 0
                res[which] = (out, flags.value)
            assert np.array_equal(res[isa][0], want), (isa, nbit)
            assert np.array_equal(res[None][0], want), nbit
            assert res[isa][1] == res[None][1] == (3 if nan else 1), (isa, nbit, res[isa][1], res[None][1])


@pytest.mark.parametrize("nbit", [1, 8, 31, 32, 33, 64, 100, 128, 200, 256])
def test_host_pack_sign_bit_exact(lib, nbit):
    """ch_host_pack_sign (the host half of K1 for pageable fp32 codes; runs without a GPU): bit = (x > 0) exactly --
    +-0.0, denormals, infinities, NaN flag, zero flag, strided rows, every thread count."""
    import numpy as np
    rng = np.random.RandomState(nbit)
    n, stride = 1037, nbit + 5
    buf = rng.randn(n, stride).astype(np.float32)
    x = buf[:, :nbit]
    x[3, 0] = 0.0
    x[5, nbit - 1] = -0.0
    x[7, nbit // 2] = np.float32(1e-45)       # denormal > 0
    x[9, 0] = -np.inf
    x[11, nbit - 1] = np.inf
    words = lib.ch_code_words(nbit)
    want = np.zeros((n, words * 32), dtype=np.uint8)
    want[:, :nbit] = x > 0
    want = np.packbits(want, axis=1, bitorder="little").view(np.uint32)
    for threads in (1, 3, 8):
        out = np.full((n, words), 0xDEADBEEF, dtype=np.uint32)
        flags = ctypes.c_uint32(0)
        rc = lib.ch_host_pack_sign(buf.ctypes.data, n, nbit, stride, out.ctypes.data, ctypes.byref(flags), threads)
        assert rc == 0 and np.array_equal(out, want) and flags.value == 1
    x[100, 1 % nbit] = np.nan
    flags = ctypes.c_uint32(0)
    lib.ch_host_pack_sign(buf.ctypes.data, n, nbit, stride, out.ctypes.data, ctypes.byref(flags), 2)
    assert flags.value == 3
    clean = np.abs(rng.randn(64, nbit)).astype(np.float32) + 0.5
    flags = ctypes.c_uint32(0)
    out = np.zeros((64, words), dtype=np.uint32)
    lib.ch_host_pack_sign(clean.ctypes.data, 64, nbit, nbit, out.ctypes.data, ctypes.byref(flags), 4)
    assert flags.value == 0 and int(np.unpackbits(out.view(np.uint8)).sum()) == 64 * nbit
    # big enough for the thread pool to really split the rows (>= 1 MB per thread)
    big = rng.randn(40000, nbit).astype(np.float32) if nbit >= 64 else None
    if big is not None:
        w2 = np.zeros((40000, words * 32), dtype=np.uint8)
        w2[:, :nbit] = big > 0
        w2 = np.packbits(w2, axis=1, bitorder="little").view(np.uint32)
        out = np.zeros((40000, words), dtype=np.uint32)
        lib.ch_host_pack_sign(big.ctypes.data, 40000, nbit, nbit, out.ctypes.data, None, 8)
        assert np.array_equal(out, w2)
