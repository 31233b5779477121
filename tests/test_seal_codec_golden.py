"""The SEAL wire-format codec against committed byte streams that SEAL 4.0 itself wrote (tests/golden/seal_wire_n256.npz, generated
from the reference's vendored libseal by tests/golden/make_wire_golden.py). Needs neither /root/reference nor oracle/_ref: this is the
pin that travels. The stateless codec entry points are host-side byte work in libhhe_b200.so; no device is used."""
import os

import numpy as np
import pytest

import common

pkg = common.package()
seal_io = __import__("importlib").import_module(common.PKG + ".seal_io")
FX = np.load(os.path.join(common.ROOT, "tests", "golden", "seal_wire_n256.npz"))


@pytest.fixture(scope="module")
def ring():
    if not os.path.exists(pkg.LIB_PATH):
        import __graft_entry__ as g
        g.build()
    return seal_io.Ring(int(FX["N"]), int(FX["t"]), FX["q"])


def test_parms_ids(ring):
    assert np.array_equal(ring.parms_id(0), FX["parms_id_data"]) and np.array_equal(ring.parms_id(1), FX["parms_id_key"])


def test_seal_streams_decode_to_the_saved_arrays(ring):
    for name in ("ct_none", "ct_zlib", "ct_zstd"):
        data = FX[name].tobytes()
        got, used = ring.load_ciphertext(data)
        assert used == len(data) and np.array_equal(got, FX["ct"]), name
    got3, _ = ring.load_ciphertext(FX["ct3_zstd"].tobytes())
    assert got3.shape[0] == 3 and np.array_equal(got3, FX["ct3"])
    rk = ring.unpack_keys(FX["rk_zstd"].tobytes())
    assert list(rk) == [0] and np.array_equal(rk[0], FX["rk"])
    gk = ring.unpack_keys(FX["gk_none"].tobytes())
    assert sorted(gk) == sorted((int(e) - 1) // 2 for e in FX["gk_elts"])
    for e, k in zip(FX["gk_elts"], FX["gk"]):
        assert np.array_equal(gk[(int(e) - 1) // 2], k)


def test_uncompressed_save_reproduces_seals_bytes(ring):
    assert ring.save_ciphertext(FX["ct"], seal_io.COMPR_NONE) == FX["ct_none"].tobytes()
    for compr in (seal_io.COMPR_ZLIB, seal_io.COMPR_ZSTD):  # compressed output need not be byte-identical, it must decode to the same
        got, _ = ring.load_ciphertext(ring.save_ciphertext(FX["ct"], compr))
        assert np.array_equal(got, FX["ct"])
