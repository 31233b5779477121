"""Generates tests/golden/pasta_n512.npz from the UNMODIFIED reference (oracle/_ref/libhhe_ref.so, i.e. the reference's
own src/pasta + src/util/sealhelper.cpp + vendored libseal-4.0.a). Run in the build container only:

    make -C oracle ref && python tests/golden/make_golden.py

The fixture pins the oracle and the CUDA path on machines where /root/reference does not exist.
Ring: N=512 (the smallest ring PASTA-3's packing accepts: 4*128 <= slots), t=65537, six 50-bit data primes + one
51-bit special prime, SEAL PRNG seed 11.
"""
import os
import re
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
import common  # noqa: E402
from oracle import refshim as R  # noqa: E402

REF_ROOT = "/root/reference"


def reference_symmetric_key():
    """The fixed 256-word key of pastahelper::get_symmetric_key (src/util/pastahelper.cpp:37-297)."""
    src = open(os.path.join(REF_ROOT, "src/util/pastahelper.cpp")).read()
    body = src[src.index("get_symmetric_key"):]
    body = body[body.index("{"): body.index("};") + 1]
    vals = [int(v, 16) for v in re.findall(r"0x[0-9a-fA-F]+", body)]
    assert len(vals) == 256, len(vals)
    return np.array(vals, dtype=np.uint64)


def main():
    N, t = 512, common.T
    q = common.small_params(N, 6)
    ref = R.Ref(N, t, q, seed=11, steps=(0, -1, 128), default_gk=False)
    rng = np.random.default_rng(2026)
    out = dict(N=np.uint64(N), t=np.uint64(t), q=np.array(q, dtype=np.uint64))
    for name, kind, elt in (("gk_m1", 0, ref.galois_elt(-1)), ("gk_p128", 0, ref.galois_elt(128)),
                            ("gk_col", 0, ref.galois_elt(0)), ("rk", 2, 0)):
        out[name] = ref.ksk(kind, elt)
        out[name + "_elt"] = np.uint64(elt)
    roots, root_t = ref.ntt_roots()
    out["psi"], out["psi_t"] = roots, np.uint64(root_t)
    bz = ref.behz()
    out["m_sk"], out["gamma"], out["base_B"], out["bsk_roots"] = np.uint64(bz["m_sk"]), np.uint64(bz["gamma"]), bz["base_B"], bz["bsk_roots"]

    # plain PASTA-3 with the reference's fixed key (SURVEY.md Appendix E KAT)
    key = reference_symmetric_key()
    out["sym_key"] = key
    kat_pt = np.arange(256, dtype=np.uint64)
    out["kat_plain"], out["kat_cipher"] = kat_pt, R.pasta_plain(key, t, kat_pt)
    for tag, ctr, layer in (("c0l0", 0, 0), ("c0l3", 0, 3), ("c5l1", 5, 1)):
        m1, m2, rc = R.pasta_layer_material(t, common.NONCE, ctr, layer)
        out["mat1_" + tag], out["mat2_" + tag], out["rc_" + tag] = m1.astype(np.uint32), m2.astype(np.uint32), rc.astype(np.uint32)

    # transciphering: 200 words -> 2 blocks (the second one short)
    plain = rng.integers(0, t, 200, dtype=np.uint64)
    sym_ct = R.pasta_plain(key, t, plain)
    enc_key = ref.encrypt(common.pack_key(key, N))
    dec = ref.pasta_decompose(enc_key, sym_ct, use_bsgs=False)
    out.update(plain=plain, sym_ct=sym_ct, enc_key=enc_key, decomposed=dec)
    slots = [ref.decrypt(c) for c in dec]
    out["decomposed_slots"] = np.stack([s[0] for s in slots])
    out["decomposed_budget"] = np.array([s[1] for s in slots], dtype=np.int64)
    assert np.array_equal(slots[0][0][:128], plain[:128]) and np.array_equal(slots[1][0][:72], plain[128:])

    # primitive known-answer vectors
    sa, sb = rng.integers(0, t, N, dtype=np.uint64), rng.integers(0, t, N, dtype=np.uint64)
    a, b = ref.encrypt(sa), ref.encrypt(sb)
    sp = rng.integers(0, t, 300, dtype=np.uint64)
    pt = ref.encode(sp)
    m3 = ref.multiply(a, b)
    out.update(slots_a=sa, slots_b=sb, ct_a=a, ct_b=b, slots_p=sp, pt=pt,
               kat_ntt_fwd=ref.ntt(0, a[0, 0]), kat_ntt_inv=ref.ntt(3, a[1, 3], inverse=True),
               kat_add=ref.add(a, b), kat_negate=ref.negate(a), kat_add_plain=ref.add_plain(a, pt),
               kat_multiply_plain=ref.multiply_plain(a, pt), kat_rot_m1=ref.rotate_rows(a, -1),
               kat_rot_p128=ref.rotate_rows(a, 128), kat_rot_col=ref.rotate_columns(a), kat_multiply=m3,
               kat_square=ref.square(a), kat_relin=ref.relinearize(m3), kat_exp3=ref.exponentiate3(a),
               kat_mask=ref.mask(a, np.ones(44, dtype=np.uint64)))
    path = os.path.join(HERE, "pasta_n512.npz")
    np.savez_compressed(path, **out)
    print("wrote", path, os.path.getsize(path) // 1024, "KiB")


if __name__ == "__main__":
    main()
