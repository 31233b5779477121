"""Drop-in proof on the B200: oracle/_ref/shim_demo runs the reference's call sequence (decomposition -> flatten ->
packed_enc_multiply -> relinearize_inplace -> encrypted_vec_sum) once with the reference's own classes on the CPU
(libseal) and once with the shim classes of host/hhe_seal_shim.h on the GPU, and compares every seal::Ciphertext. The service-level mirrors csp_b200::decompose / evaluate_model (one engine call per CSP
request) are checked against the same reference ciphertexts."""
import json
import os
import subprocess

import pytest

import common

DEMO = os.path.join(common.ROOT, "oracle", "_ref", "shim_demo")
pytestmark = [pytest.mark.gpu, pytest.mark.skipif(not os.path.exists(DEMO), reason="oracle/_ref/shim_demo not built")]


@pytest.mark.parametrize("input_len,sum_len", [("150", "16"), ("140", "140")])  # the second also runs csp_b200::evaluate_model
def test_reference_call_sites_give_identical_ciphertexts(input_len, sum_len):
    out = subprocess.run([DEMO, "16384", input_len, sum_len], capture_output=True, text=True, timeout=900)
    assert out.returncode == 0, out.stdout + out.stderr
    res = json.loads(out.stdout.strip().splitlines()[-1])
    assert res["ciphertexts_identical"] and res["decrypts_to_plaintext"] and res["noise_budget"] > 0
    assert res["evaluator_facade_ok"]  # the reference's own call-site lines with hhe_shim::Evaluator (keys passed as temporaries, recognised by content)
    assert res["serialized_keys_ok"]  # keys re-uploaded from their GaloisKeys::save / RelinKeys::save bytes give the same ciphertext
