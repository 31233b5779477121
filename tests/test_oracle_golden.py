"""The oracle (oracle/hhe_oracle.c) against the golden vectors generated from the reference (tests/golden/make_golden.py).
Runs without a GPU and without /root/reference."""
import os

import numpy as np
import pytest

import common
from oracle import oracle as O

FX = np.load(os.path.join(common.ROOT, "tests", "golden", "pasta_n512.npz"))


@pytest.fixture(scope="module")
def orc():
    o = O.Oracle(int(FX["N"]), int(FX["t"]), FX["q"])
    for name, kind in (("gk_m1", 0), ("gk_p128", 0), ("gk_col", 0), ("rk", 2)):
        o.load_ksk(kind, int(FX[name + "_elt"]), FX[name])
    yield o
    o.close()


def test_shake128_against_hashlib():
    import hashlib
    for msg in (b"", b"abc", bytes(range(200)), b"x" * 168, b"y" * 167):
        assert O.shake128(msg, 500) == hashlib.shake_128(msg).digest(500)


def test_constants_match_seal(orc):
    psi, psi_t = orc.ntt_roots()
    assert np.array_equal(psi, FX["psi"]) and psi_t == int(FX["psi_t"])
    b = orc.behz()
    assert b["m_sk"] == int(FX["m_sk"]) and b["gamma"] == int(FX["gamma"])
    assert np.array_equal(b["base_B"], FX["base_B"]) and np.array_equal(b["bsk_roots"], FX["bsk_roots"])
    for name, step in (("gk_m1", -1), ("gk_p128", 128), ("gk_col", 0)):
        assert orc.galois_elt(step) == int(FX[name + "_elt"])


def test_survey_appendix_e_kat():
    # plaintext [0..127], reference key, nonce 123456789
    ct = O.pasta_plain(FX["sym_key"], common.T, np.arange(128, dtype=np.uint64))
    assert list(ct[:8]) == [30446, 62406, 62716, 38766, 43125, 6036, 63532, 7424]
    assert list(ct[126:128]) == [63544, 48230]
    ct2 = O.pasta_plain(FX["sym_key"], common.T, np.arange(256, dtype=np.uint64))
    assert list(ct2[128:132]) == [11021, 61753, 2637, 43237]
    m1, _, _ = O.pasta_layer_material(common.T, common.NONCE, 0, 0)
    assert list(m1[0, :6]) == [34686, 37780, 45807, 58845, 36538, 7530]
    assert list(m1[1, :4]) == [8576, 58655, 54978, 65299] and m1[127, 127] == 55028


def test_plain_pasta_golden():
    assert np.array_equal(O.pasta_plain(FX["sym_key"], common.T, FX["kat_plain"]), FX["kat_cipher"])
    assert np.array_equal(O.pasta_plain(FX["sym_key"], common.T, FX["kat_cipher"], decrypt=True), FX["kat_plain"])
    assert np.array_equal(O.pasta_plain(FX["sym_key"], common.T, FX["plain"]), FX["sym_ct"])


@pytest.mark.parametrize("tag,ctr,layer", [("c0l0", 0, 0), ("c0l3", 0, 3), ("c5l1", 5, 1)])
def test_round_material_golden(tag, ctr, layer):
    m1, m2, rc = O.pasta_layer_material(common.T, common.NONCE, ctr, layer)
    assert np.array_equal(m1, FX["mat1_" + tag]) and np.array_equal(m2, FX["mat2_" + tag]) and np.array_equal(rc, FX["rc_" + tag])


def test_primitive_kats(orc):
    a, b, pt = FX["ct_a"], FX["ct_b"], FX["pt"]
    assert np.array_equal(orc.ntt(0, a[0, 0]), FX["kat_ntt_fwd"])
    assert np.array_equal(orc.ntt(3, a[1, 3], inverse=True), FX["kat_ntt_inv"])
    assert np.array_equal(orc.encode(FX["slots_p"]), pt)
    assert np.array_equal(orc.add(a, b), FX["kat_add"])
    assert np.array_equal(orc.negate(a), FX["kat_negate"])
    assert np.array_equal(orc.add_plain(a, pt), FX["kat_add_plain"])
    assert np.array_equal(orc.multiply_plain(a, pt), FX["kat_multiply_plain"])
    assert np.array_equal(orc.rotate_rows(a, -1), FX["kat_rot_m1"])
    assert np.array_equal(orc.rotate_rows(a, 128), FX["kat_rot_p128"])
    assert np.array_equal(orc.rotate_columns(a), FX["kat_rot_col"])
    m3 = orc.multiply(a, b)
    assert np.array_equal(m3, FX["kat_multiply"])
    assert np.array_equal(orc.multiply(a, a), FX["kat_square"])
    assert np.array_equal(orc.relinearize(m3), FX["kat_relin"])
    assert np.array_equal(orc.exponentiate3(a), FX["kat_exp3"])
    assert np.array_equal(orc.mask(a, np.ones(44, dtype=np.uint64)), FX["kat_mask"])


def test_missing_galois_key_is_an_error(orc):
    with pytest.raises(O.OracleError):
        orc.rotate_rows(FX["ct_a"], -2)  # NAF of -2 is the single term -2: SEAL throws "Galois key not present"


def test_transciphering_golden(orc):
    got = orc.pasta_decompose(FX["enc_key"], FX["sym_ct"])
    assert np.array_equal(got, FX["decomposed"])
    # the golden ciphertexts decrypt (by the reference) to the PASTA plaintext
    assert np.array_equal(FX["decomposed_slots"][0][:128], FX["plain"][:128])
    assert np.array_equal(FX["decomposed_slots"][1][:72], FX["plain"][128:])
