"""Regenerates tests/golden/containers.json: the byte containers of one small seeded Groth16 instance as oracle/py/containers.py writes
them (verifying key and proof in full, the proving key by length and SHA-256), so that a change of the restated layouts shows up as a
diff of committed bytes and the GPU box can compare zkpor_vk_decode / zkpor_proof_decode / zkpor_pk_write against them.
    python tests/golden/make_containers.py"""
import hashlib
import json
import os
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(HERE, "..", "..", "oracle", "py"))

import containers as ct  # noqa: E402
import groth16 as g16  # noqa: E402

SEED = dict(n_constraints=12, nb_secret=6, circuit=3, toxic=4, inputs=5, r=7, s=9)


def build():
    cs = g16.synth_r1cs(SEED["n_constraints"], SEED["nb_secret"], SEED["circuit"])
    pk, vk = g16.setup(cs, g16.toxic_from_seed(SEED["toxic"]))
    pub, sec = g16.synth_inputs(cs, SEED["inputs"])
    proof, _ = g16.prove(cs, pk, pub, sec, SEED["r"], SEED["s"])
    sha = lambda b: hashlib.sha256(b).hexdigest()
    return dict(seed=SEED, vk_compressed=ct.vk_bytes(vk).hex(), vk_raw_sha256=sha(ct.vk_bytes(vk, raw=True)),
                proof_compressed=ct.proof_bytes(proof).hex(), proof_raw=ct.proof_bytes(proof, raw=True).hex(),
                pk_compressed_len=len(ct.pk_bytes(pk)), pk_compressed_sha256=sha(ct.pk_bytes(pk)),
                pk_raw_len=len(ct.pk_bytes(pk, raw=True)), pk_raw_sha256=sha(ct.pk_bytes(pk, raw=True)))


if __name__ == "__main__":
    json.dump(build(), open(os.path.join(HERE, "containers.json"), "w"), indent=1)
    print("wrote containers.json")
