"""GPU tier: the batched IPA final check (poly-commitment `SRS::verify`) through the C ABI.

The reference holds no vector for an opening proof on its own and the Fq-sponge needs the (unavailable) kimchi
Poseidon table, so this row is checked for SELF-CONSISTENCY under arbitrary tables: openings made by the oracle
prover (oracle/ipa.py, a restatement of `SRS::open`) must be accepted, every mutation rejected, the per-opening bits
of a mixed batch must be exact, and the oracle's own `verify_one` must agree on every case."""
import copy
import random

import pytest

from oracle import cref, ipa, pasta, poseidon as oposeidon

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("curve,k", [(0, 9), (1, 8)])
def test_ipa_openings_accept_and_mutations_reject(gpu, curve, k):
    cv = ipa.CurveCtx(curve)
    table = oposeidon.random_table(cv.base, 77 + curve)
    g = gpu.srs_points(curve, 0, 1 << k)
    h_pt = cref.bytes_to_point(gpu.srs_points(curve, 0, 1, want_h=True)[1])
    rng = random.Random(31 + curve)
    n, n_comm = 5, 3
    items = []
    for _ in range(n):
        polys = [[rng.randrange(cv.scalar) for _ in range(1 << k)] for _ in range(n_comm)]
        comms = [ipa.commit(cv, g, f) for f in polys]
        elm = [rng.randrange(cv.scalar) for _ in range(2)]
        polyscale, evalscale = rng.randrange(cv.scalar), rng.randrange(cv.scalar)
        sp = ipa.FqSponge(cv, table)
        for _ in range(3):
            sp.absorb_fq(rng.randrange(cv.base))
        sp.challenge()
        state, mode, count = sp.export()
        opening, cip = ipa.open_proof(cv, g, h_pt, k, polys, elm, polyscale, evalscale, copy.deepcopy(sp), rng)
        item = {"state": state, "cip": cip, "polyscale": polyscale, "evalscale": evalscale, "z1": opening["z1"], "z2": opening["z2"],
                "elm": elm, "delta": opening["delta"], "sg": opening["sg"], "commitments": comms, "lr": opening["lr"]}
        assert ipa.verify_one(cv, g, h_pt, k, comms, elm, polyscale, evalscale, copy.deepcopy(sp), opening, cip)
        items.append((item, sp))
    assert (mode, count) == (1, 1)
    tb = oposeidon.table_bytes(table)
    good = [it for it, _ in items]
    assert gpu.ipa_verify(curve, tb, good, mode, count) == [1] * n
    assert gpu.ipa_verify(curve, tb, good[:1], mode, count) == [1]

    def oracle_says(item, sp):
        opening = {"lr": item["lr"], "delta": item["delta"], "z1": item["z1"], "z2": item["z2"], "sg": item["sg"]}
        return int(ipa.verify_one(cv, g, h_pt, k, item["commitments"], item["elm"], item["polyscale"], item["evalscale"], copy.deepcopy(sp), opening, item["cip"]))

    # one mutation per opening of the batch; the last stays good
    other = cref.bytes_to_point(g[64 * 7: 64 * 8])
    muts = [("z1", lambda it: it.update(z1=(it["z1"] + 1) % cv.scalar)),
            ("cip", lambda it: it.update(cip=(it["cip"] + 5) % cv.scalar)),
            ("L_2", lambda it: it.update(lr=[(other, r) if j == 2 else (l, r) for j, (l, r) in enumerate(it["lr"])])),
            ("commitment", lambda it: it.update(commitments=[other] + it["commitments"][1:]))]
    batch, want = [], []
    for (it, sp), (name, mut) in zip(items, muts):
        m = copy.deepcopy(it)
        mut(m)
        batch.append(m)
        want.append(0)
        assert oracle_says(m, sp) == 0, name
    batch.append(good[-1])
    want.append(1)
    assert gpu.ipa_verify(curve, tb, batch, mode, count) == want
    # more single mutations against the first opening: z2, delta, sg, polyscale, evalscale, an evaluation point, the sponge state
    it0, sp0 = items[0]
    for name, mut in [("z2", lambda it: it.update(z2=(it["z2"] + 1) % cv.scalar)), ("delta", lambda it: it.update(delta=other)),
                      ("sg", lambda it: it.update(sg=other)), ("polyscale", lambda it: it.update(polyscale=(it["polyscale"] + 1) % cv.scalar)),
                      ("evalscale", lambda it: it.update(evalscale=(it["evalscale"] + 1) % cv.scalar)),
                      ("elm", lambda it: it.update(elm=[it["elm"][0], (it["elm"][1] + 1) % cv.scalar])),
                      ("state", lambda it: it.update(state=[(it["state"][0] + 1) % cv.base] + it["state"][1:]))]:
        m = copy.deepcopy(it0)
        mut(m)
        assert gpu.ipa_verify(curve, tb, [good[1], m, good[2]], mode, count) == [1, 0, 1], name
    # an off-curve point or a non-canonical scalar rejects that opening only, without an exception
    m = copy.deepcopy(it0)
    m["delta"] = (m["delta"][0], (m["delta"][1] + 1) % cv.base)
    m2 = copy.deepcopy(it0)
    m2["z1"] = cv.scalar  # not canonical
    assert gpu.ipa_verify(curve, tb, [m, good[1], m2], mode, count) == [0, 1, 0]
    # a different (wrong) Poseidon table changes every challenge: nothing verifies
    assert gpu.ipa_verify(curve, oposeidon.table_bytes(oposeidon.random_table(cv.base, 5)), good[:2], mode, count) == [0, 0]


def test_ipa_kimchi_shape_fixture(gpu):
    """One opening of the wrap proof's shape (Pallas, 15 rounds over the full 2^15 SRS, 47 commitments, 2 points) made
    by the oracle prover (tools/make_ipa_fixture.py): accepted alone and in a batch; corrupted copies are singled out."""
    import os
    from conftest import GOLDEN

    curve, table, op, mode, count = gpu.load_ipa_fixture(os.path.join(GOLDEN, "ipa_pallas_k15.json"))
    assert (curve, len(op["lr"]), len(op["commitments"])) == (0, 15, 47)
    assert gpu.ipa_verify(curve, table, [op], mode, count) == [1]
    bad1 = dict(op, z2=(op["z2"] + 1) % pasta.Q)
    bad2 = dict(op, lr=[(r, l) if j == 14 else (l, r) for j, (l, r) in enumerate(op["lr"])])  # L and R swapped in the last round
    batch = [op] * 70
    batch[3], batch[64], batch[69] = bad1, bad2, bad1
    want = [1] * 70
    want[3] = want[64] = want[69] = 0
    assert gpu.ipa_verify(curve, table, batch, mode, count) == want


@pytest.mark.parametrize("prefix", ["a", "aa", "aaa", "as", "ass", "asa"])
def test_ipa_resumes_the_sponge_in_every_mode(gpu, prefix):
    """The Fq-sponge is handed over mid-stream: Absorbed(1), Absorbed(2), Absorbed(1) after a permutation, Squeezed(1),
    Squeezed(2), and absorbing again after a squeeze.  The device transcript must resume each of them like the oracle."""
    curve, k = 0, 8
    cv = ipa.CurveCtx(curve)
    table = oposeidon.random_table(cv.base, 4242)
    g = gpu.srs_points(curve, 0, 1 << k)
    h_pt = cref.bytes_to_point(gpu.srs_points(curve, 0, 1, want_h=True)[1])
    rng = random.Random(len(prefix) * 7 + sum(map(ord, prefix)))
    items = []
    for _ in range(2):
        polys = [[rng.randrange(cv.scalar) for _ in range(1 << k)] for _ in range(2)]
        comms = [ipa.commit(cv, g, f) for f in polys]
        elm = [rng.randrange(cv.scalar) for _ in range(2)]
        polyscale, evalscale = rng.randrange(cv.scalar), rng.randrange(cv.scalar)
        sp = ipa.FqSponge(cv, table)
        for ch in prefix:
            sp.absorb_fq(rng.randrange(cv.base)) if ch == "a" else sp.challenge_fq()
        state, mode, count = sp.export()
        opening, cip = ipa.open_proof(cv, g, h_pt, k, polys, elm, polyscale, evalscale, copy.deepcopy(sp), rng)
        items.append({"state": state, "cip": cip, "polyscale": polyscale, "evalscale": evalscale, "z1": opening["z1"], "z2": opening["z2"],
                      "elm": elm, "delta": opening["delta"], "sg": opening["sg"], "commitments": comms, "lr": opening["lr"]})
    tb = oposeidon.table_bytes(table)
    assert gpu.ipa_verify(curve, tb, items, mode, count) == [1, 1], (mode, count)
    bad = dict(items[1], cip=(items[1]["cip"] + 1) % cv.scalar)
    assert gpu.ipa_verify(curve, tb, [items[0], bad], mode, count) == [1, 0]
