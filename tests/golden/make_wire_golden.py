"""Generates tests/golden/seal_wire_n256.npz: byte streams written by SEAL 4.0 itself (the reference's vendored libseal through
oracle/_ref) for one ciphertext (compression none / zlib / zstd), a size-3 ciphertext, a RelinKeys and a GaloisKeys object, together
with the arrays they hold. Pins csrc/seal_codec.cpp on machines where neither /root/reference nor oracle/_ref exists.

    make -C oracle ref && python tests/golden/make_wire_golden.py

Ring: N=256, t=65537, two 48-bit data primes + one 49-bit special prime (common.small_params(256, 2, 48)), SEAL PRNG seed 11
(small on purpose: the fixture is committed).
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
import common  # noqa: E402
from oracle import refshim as R  # noqa: E402

N = 256
q = common.small_params(N, 2, 48)
ref = R.Ref(N, common.T, q, seed=11, steps=(0, -1), default_gk=False)
ct = ref.encrypt(np.arange(200, dtype=np.uint64))
ct3 = ref.multiply(ct, ct)
b = lambda x: np.frombuffer(x, dtype=np.uint8)  # noqa: E731
elts = ref.list_galois(0)
np.savez_compressed(
    os.path.join(HERE, "seal_wire_n256.npz"), N=N, t=common.T, q=np.array(q, dtype=np.uint64),
    parms_id_data=ref.parms_id(0), parms_id_key=ref.parms_id(1),
    ct=ct, ct_none=b(ref.ct_save(ct, 0)), ct_zlib=b(ref.ct_save(ct, 1)), ct_zstd=b(ref.ct_save(ct, 2)),
    ct3=ct3, ct3_zstd=b(ref.ct_save(ct3, 2)),
    rk=ref.ksk(2), rk_zstd=b(ref.keys_save(2, 2)),
    gk_elts=np.array(elts, dtype=np.uint64), gk=np.stack([ref.ksk(0, e) for e in elts]), gk_none=b(ref.keys_save(0, 0)))
print("written", os.path.getsize(os.path.join(HERE, "seal_wire_n256.npz")), "bytes")
