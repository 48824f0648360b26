#!/usr/bin/env python3
"""Make tests/golden/ipa_pallas_k15.json: ONE opening proof of kimchi's shape over Pallas (15 rounds over g[0..2^15),
47 commitments of one chunk each, 2 evaluation points) produced by the oracle prover (oracle/ipa.py, a restatement of
poly-commitment `SRS::open`) under an ARBITRARY seeded Poseidon table -- the kimchi constants are unavailable
(DESIGN.md section 0), so this fixture pins self-consistency of the device verifier, not parity with the reference.

usage: python tools/make_ipa_fixture.py   (CPU only; ~1 minute)
"""
import copy
import json
import os
import random
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import cref, ipa, poseidon  # noqa: E402

CURVE, K, N_COMM, SEED = 0, 15, 47, 0x4D494E41


def main():
    cv = ipa.CurveCtx(CURVE)
    table = poseidon.random_table(cv.base, SEED)
    g, h = cref.srs_derive(cv.base_fid, 0, 1 << K, True)
    h_pt = cref.bytes_to_point(h)
    rng = random.Random(SEED)
    polys = [[rng.randrange(cv.scalar) for _ in range(1 << K)] for _ in range(N_COMM)]
    comms = [ipa.commit(cv, g, f) for f in polys]
    elm = [rng.randrange(cv.scalar) for _ in range(2)]
    polyscale, evalscale = rng.randrange(cv.scalar), rng.randrange(cv.scalar)
    sp = ipa.FqSponge(cv, table)
    for _ in range(5):
        sp.absorb_fq(rng.randrange(cv.base))
    sp.challenge()  # Squeezed(1), like kimchi's fq_sponge_before_evaluations
    state, mode, count = sp.export()
    opening, cip = ipa.open_proof(cv, g, h_pt, K, polys, elm, polyscale, evalscale, copy.deepcopy(sp), rng)
    assert ipa.verify_one(cv, g, h_pt, K, comms, elm, polyscale, evalscale, copy.deepcopy(sp), opening, cip)
    hx = lambda x: "%064x" % x
    pt = lambda p: [hx(p[0]), hx(p[1])]
    out = {"curve": CURVE, "rounds": K, "sponge_mode": mode, "sponge_count": count, "table": [hx(x) for x in table],
           "state": [hx(x) for x in state], "cip": hx(cip), "polyscale": hx(polyscale), "evalscale": hx(evalscale),
           "z1": hx(opening["z1"]), "z2": hx(opening["z2"]), "elm": [hx(x) for x in elm], "delta": pt(opening["delta"]),
           "sg": pt(opening["sg"]), "commitments": [pt(p) for p in comms], "lr": [[pt(l), pt(r)] for l, r in opening["lr"]]}
    path = os.path.join(ROOT, "tests", "golden", "ipa_pallas_k15.json")
    json.dump(out, open(path, "w"), indent=0)
    print("wrote", path, os.path.getsize(path), "bytes")


if __name__ == "__main__":
    main()
