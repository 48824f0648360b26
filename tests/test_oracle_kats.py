"""Pin the CPU oracle against everything the reference committed for this path."""
import hashlib
import json
import os
import random

import pytest

from oracle import cref, pasta, wire
from conftest import GOLDEN, golden

SRS_PINS = json.load(open(os.path.join(GOLDEN, "srs_sha256.json")))


@pytest.fixture(scope="module")
def vesta_srs():
    g, _ = cref.srs_derive(cref.FQ, 0, 65536, False)
    return g


@pytest.fixture(scope="module")
def pallas_srs():
    g, _ = cref.srs_derive(cref.FP, 0, 32768, False)
    return g


def test_fixture_sizes_and_full_consumption(state_proof):
    # byte-exact consumption of every fixture
    assert state_proof["_proof_end"] == 13849 and state_proof["_consumed"] == 48342
    pub = wire.decode_state_pub(golden("mina_state.pub"))
    assert pub["_consumed"] == 1057 and pub["is_state_proof_from_devnet"] is True
    acc = wire.decode_account_proof(golden("mina_account.proof"))
    assert acc["_consumed"] == 1832 and len(acc["merkle_path"]) == 35
    apub = wire.decode_account_pub(golden("mina_account.pub"))
    assert apub["_consumed"] == 3496 and len(apub["encoded_account"]) == 3456


def test_bad_hash_pub_fails_at_bool_decode():
    # AL/operator/mina/lib/src/lib.rs:246-261 -- rejects while deserialising (first byte 0x5d)
    with pytest.raises(wire.DecodeError):
        wire.decode_state_pub(golden("mina_state_bad_hash.pub"))


def test_structural_chain_checks(state_proof):
    # K-F: stored previous_state_hash chain and ledger hashes agree with the public input
    pub = wire.decode_state_pub(golden("mina_state.pub"))
    st = state_proof["candidate_chain_states"]
    for i in range(15):
        assert st[i + 1]["previous_state_hash"] == pub["candidate_chain_state_hashes"][i]
    for i in range(16):
        tgt = st[i]["body"]["blockchain_state"]["ledger_proof_statement"]["target"]["first_pass_ledger"]
        assert tgt == pub["candidate_chain_ledger_hashes"][i]
    assert pub["bridge_tip_state_hash"] == pub["candidate_chain_state_hashes"][14]


def test_points_on_curve(state_proof):
    pr = state_proof["candidate_tip_proof"]
    assert pasta.is_on_curve(pr["wrap_challenge_polynomial_commitment"], pasta.Q)
    for pt in pr["step_challenge_polynomial_commitments"] + pr["proof"]["w_comm"] + pr["proof"]["t_comm"]:
        assert pasta.is_on_curve(pt, pasta.P)
    for l, r in pr["proof"]["lr"]:
        assert pasta.is_on_curve(l, pasta.P) and pasta.is_on_curve(r, pasta.P)


def test_endo_unit_vectors(state_proof):
    # SURVEY Appendix C.2b
    pr = state_proof["candidate_tip_proof"]
    assert (
        pasta.endo_to_field(pr["bulletproof_challenges"][0], pasta.ENDO_FP, pasta.P)
        == 14263189808346682768889437752131109163706449277114519493415332712824597396161
    )
    assert (
        pasta.endo_to_field(pr["wrap_old_bulletproof_challenges"][0][0], pasta.ENDO_FQ, pasta.Q)
        == 18897605359931041753525855026506093960596801015646020115629628159590724249968
    )
    pre = b"".join(x.to_bytes(16, "little") for x in pr["bulletproof_challenges"])
    assert cref.endo_to_field(cref.FP, pre, pasta.ENDO_FP) == cref.ints_to_bytes(
        [pasta.endo_to_field(x, pasta.ENDO_FP, pasta.P) for x in pr["bulletproof_challenges"]]
    )


def test_srs_derivation_matches_committed_files(vesta_srs, pallas_srs):
    # K-D: the digests were produced from srs/vesta.srs and srs/pallas.srs by tools/make_golden.py
    assert hashlib.sha256(vesta_srs).hexdigest() == SRS_PINS["vesta"]["sha256_g_65536"]
    assert hashlib.sha256(pallas_srs).hexdigest() == SRS_PINS["pallas"]["sha256_g_32768"]
    for fid, name in ((cref.FQ, "vesta"), (cref.FP, "pallas")):
        _, h = cref.srs_derive(fid, 0, 0, True)
        assert h.hex() == SRS_PINS[name]["h"]


def test_srs_python_and_c_agree():
    gm = pasta.GroupMap(pasta.Q)
    pts = [pasta.srs_point(i, gm) for i in range(12)]
    c, h = cref.srs_derive(cref.FQ, 0, 12, True)
    assert c == cref.points_to_bytes(pts)
    assert cref.bytes_to_point(h) == pasta.srs_blinding(gm)


@pytest.mark.skipif(not os.path.exists("/root/reference/srs/vesta.srs"), reason="reference tree not mounted")
def test_srs_file_decompression_live():
    g, h = pasta.srs_compressed_bytes("/root/reference/srs/vesta.srs")
    aff = cref.decompress(cref.FQ, b"".join(g[:2048]))
    der, hd = cref.srs_derive(cref.FQ, 0, 2048, True)
    assert aff == der and cref.decompress(cref.FQ, h) == hd
    assert pasta.decompress_point(g[7], pasta.Q) == cref.bytes_to_point(aff[7 * 64 : 8 * 64])


def test_msm_c_vs_python_small():
    random.seed(11)
    pts, _ = cref.srs_derive(cref.FP, 0, 70, False)
    ptl = [cref.bytes_to_point(pts[64 * i : 64 * i + 64]) for i in range(70)]
    sc = [random.randrange(pasta.Q) for _ in range(70)]
    sc[0], sc[1], sc[2] = 0, 1, pasta.Q - 1
    want = pasta.msm_naive(sc, ptl, pasta.P)
    assert pasta.msm_pippenger(sc, ptl, pasta.P) == want
    for threads in (1, 4):
        got, inf = cref.msm(cref.FP, cref.ints_to_bytes(sc), pts, threads)
        assert not inf and cref.bytes_to_point(got) == want


def test_ark_window_rule():
    # ark-ec 0.3: c = 3 if n < 32 else ln(n) + 2, ln(n) = ceil(log2 n) * 69 / 100
    lib = cref.lib()
    assert lib.oracle_ark_window_bits(31) == 3
    assert lib.oracle_ark_window_bits(65537) == 13  # ceil(log2)=17 -> 11 + 2
    assert lib.oracle_ark_window_bits(32768) == 12
    assert lib.oracle_ark_window_bits(1 << 20) == 15
    for n in (31, 32, 41, 65537, 32850, 1 << 20):
        assert lib.oracle_ark_window_bits(n) == pasta.ark_window_bits(n)


def test_accumulator_kat_vesta(state_proof, vesta_srs):
    # K-A: <b_poly_coefficients(chals), vesta.g> == challenge_polynomial_commitment (2^16 points)
    pr = state_proof["candidate_tip_proof"]
    pre = b"".join(x.to_bytes(16, "little") for x in pr["bulletproof_challenges"])
    chals = cref.endo_to_field(cref.FP, pre, pasta.ENDO_FP)
    s = cref.bpoly_coeffs(cref.FP, chals)
    got, inf = cref.msm(cref.FQ, s, vesta_srs, 8)
    assert not inf and cref.bytes_to_point(got) == pr["wrap_challenge_polynomial_commitment"]


@pytest.mark.parametrize("k", [0, 1])
def test_accumulator_kat_pallas(state_proof, pallas_srs, k):
    # K-B / K-C: the two step-side accumulators (2^15 points each)
    pr = state_proof["candidate_tip_proof"]
    pre = b"".join(x.to_bytes(16, "little") for x in pr["wrap_old_bulletproof_challenges"][k])
    chals = cref.endo_to_field(cref.FQ, pre, pasta.ENDO_FQ)
    s = cref.bpoly_coeffs(cref.FQ, chals)
    got, inf = cref.msm(cref.FP, s, pallas_srs, 8)
    assert not inf and cref.bytes_to_point(got) == pr["step_challenge_polynomial_commitments"][k]


def test_bpoly_python_vs_c_and_eval():
    random.seed(5)
    chals = [random.randrange(pasta.P) for _ in range(7)]
    s = pasta.b_poly_coefficients(chals, pasta.P)
    assert cref.bpoly_coeffs(cref.FP, cref.ints_to_bytes(chals)) == cref.ints_to_bytes(s)
    x = random.randrange(pasta.P)
    horner = 0
    for coef in reversed(s):
        horner = (horner * x + coef) % pasta.P
    assert horner == pasta.b_poly(chals, x, pasta.P)
    # ordering unit vector from SURVEY C.2b: s[1] = c[k-1], s[2] = c[k-2], s[3] = c[k-1]*c[k-2]
    assert s[1] == chals[6] and s[2] == chals[5] and s[3] == chals[6] * chals[5] % pasta.P


def test_domain_generator():
    # K-G
    gen = pow(pow(5, (pasta.Q - 1) >> 32, pasta.Q), 1 << 18, pasta.Q)
    assert gen == 0x1E5587687024253BB079B38D9C5371594958E496C605D3BD898B34D068AFBEE7
    assert pow(gen, 1 << 14, pasta.Q) == 1 and pow(gen, 1 << 13, pasta.Q) != 1
