"""GPU tier: accumulator_check from raw proof bytes, the IPA scalar kernels (K2/K4/K5), the Poseidon
kernels (K3) and the verifier boundary -- all through the C ABI, bit-exact against the oracle."""
import random

import pytest

from oracle import cref, pasta, poseidon as oposeidon, wire
from conftest import golden

pytestmark = pytest.mark.gpu

FIELD_MOD = {0: pasta.P, 1: pasta.Q}
ENDO = {0: pasta.ENDO_FP, 1: pasta.ENDO_FQ}

# byte offsets inside mina_state.proof
OFF_WRAP_PRECHAL = 73          # 16 x 16 B, Vesta-side accumulator challenges
OFF_WRAP_SG_X = 374            # accumulator point (x 374..406, y 414..446)
OFF_STEP_PRECHAL = 446         # 2 x 15 x 16 B, Pallas-side
OFF_STEP_SG = 934              # pair k at 934 + 80k (x at +8, y at +48)


# ---- K4: ScalarChallenge::to_field ------------------------------------------------------------------------
def test_endo_to_field_unit_vectors(gpu):
    # SURVEY Appendix C.2b, taken from the reference's own proof
    pre = (0x6240EE90971028BF | (0xB410EEB2C2577FAB << 64)).to_bytes(16, "little")
    assert int.from_bytes(gpu.endo_to_field(0, pre), "little") == 14263189808346682768889437752131109163706449277114519493415332712824597396161
    pre = (0xDBEC2BE1808004C1 | (0x33BD213574BEB614 << 64)).to_bytes(16, "little")
    assert int.from_bytes(gpu.endo_to_field(1, pre), "little") == 18897605359931041753525855026506093960596801015646020115629628159590724249968


@pytest.mark.parametrize("fid", [0, 1])
def test_endo_to_field_matches_oracle(gpu, fid):
    rng = random.Random(40 + fid)
    pre = b"".join(rng.getrandbits(128).to_bytes(16, "little") for _ in range(777))
    pre += (0).to_bytes(16, "little") + ((1 << 128) - 1).to_bytes(16, "little") + (1).to_bytes(16, "little") + (1 << 127).to_bytes(16, "little")
    assert gpu.endo_to_field(fid, pre) == cref.endo_to_field(fid, pre, ENDO[fid])
    assert gpu.endo_to_field(fid, b"") == b""


# ---- K2: b_poly_coefficients ----------------------------------------------------------------------------------
@pytest.mark.parametrize("fid,k", [(0, 16), (1, 15), (0, 8), (1, 11)])
def test_bpoly_coeffs_match_oracle(gpu, fid, k):
    rng = random.Random(50 + k)
    m = FIELD_MOD[fid]
    chals = [[rng.randrange(m) for _ in range(k)] for _ in range(3)]
    chals[1][0], chals[1][k - 1] = 0, 1  # edge values
    got = gpu.bpoly_coeffs(fid, b"".join(cref.ints_to_bytes(c) for c in chals), k)
    want = b"".join(cref.bpoly_coeffs(fid, cref.ints_to_bytes(c)) for c in chals)
    assert got == want
    # ordering facts from SURVEY C.2b: s[1] = c[k-1], s[2] = c[k-2], s[3] = c[k-1] c[k-2], s[2^k - 1] = prod c
    s = [int.from_bytes(got[32 * i : 32 * i + 32], "little") for i in range(1 << k)]
    c = chals[0]
    prod = 1
    for x in c:
        prod = prod * x % m
    assert s[0] == 1 and s[1] == c[k - 1] and s[2] == c[k - 2] and s[3] == c[k - 1] * c[k - 2] % m and s[(1 << k) - 1] == prod


@pytest.mark.parametrize("fid,k", [(0, 16), (1, 15)])
def test_bpoly_combine_is_the_random_linear_combination(gpu, fid, k):
    rng = random.Random(60 + k)
    m = FIELD_MOD[fid]
    n = 5
    chals = [[rng.randrange(m) for _ in range(k)] for _ in range(n)]
    r = [rng.getrandbits(128) for _ in range(n)]
    got = gpu.bpoly_combine(fid, b"".join(cref.ints_to_bytes(c) for c in chals), cref.ints_to_bytes(r), k)
    coeffs = [cref.bpoly_coeffs(fid, cref.ints_to_bytes(c)) for c in chals]
    idx = [0, 1, 255, 256, 257, (1 << k) - 1] + [rng.randrange(1 << k) for _ in range(200)]
    for i in idx:
        want = sum(r[j] * int.from_bytes(coeffs[j][32 * i : 32 * i + 32], "little") for j in range(n)) % m
        assert int.from_bytes(got[32 * i : 32 * i + 32], "little") == want
    # empty batch: the zero vector
    assert gpu.bpoly_combine(fid, b"", b"", k) == b"\0" * (32 << k)


@pytest.mark.parametrize("fid,k", [(0, 16), (1, 15)])
def test_bpoly_eval_matches_the_product_formula(gpu, fid, k):
    rng = random.Random(70 + k)
    m = FIELD_MOD[fid]
    chals = [[rng.randrange(m) for _ in range(k)] for _ in range(4)]
    xs = [[rng.randrange(m) for _ in range(2)] for _ in range(4)]
    xs[0][0] = 0
    got = gpu.bpoly_eval(fid, b"".join(cref.ints_to_bytes(c) for c in chals), b"".join(cref.ints_to_bytes(x) for x in xs), k, 2)
    for p in range(4):
        for t in range(2):
            want = 1
            for i in range(k):
                want = want * (1 + chals[p][i] * pow(xs[p][t], 1 << (k - 1 - i), m)) % m
            assert int.from_bytes(got[32 * (2 * p + t) : 32 * (2 * p + t) + 32], "little") == want
    # consistency with the coefficients: b(x) = sum_i s[i] x^i
    s = cref.bpoly_coeffs(fid, cref.ints_to_bytes(chals[1]))
    x, acc = xs[1][1], 0
    for i in reversed(range(1 << k)):
        acc = (acc * x + int.from_bytes(s[32 * i : 32 * i + 32], "little")) % m
    assert int.from_bytes(got[32 * 3 : 32 * 4], "little") == acc


# ---- a7: accumulator_check from raw proof bytes -------------------------------------------------------------------
def test_accumulator_check_accepts_the_reference_proof(gpu):
    proof = golden("mina_state.proof")
    assert gpu.accumulator_check([proof]) == [(1, 1, 1)]  # K-A, K-B, K-C


def _flip(data: bytes, off: int, bit: int = 0) -> bytes:
    m = bytearray(data)
    m[off] ^= 1 << bit
    return bytes(m)


def test_accumulator_check_rejects_single_bit_mutations(gpu):
    proof = golden("mina_state.proof")
    cases = {
        "wrap prechallenge": (_flip(proof, OFF_WRAP_PRECHAL + 5), (0, 1, 1)),
        "last wrap prechallenge": (_flip(proof, OFF_WRAP_PRECHAL + 255, 7), (0, 1, 1)),
        "wrap sg.x": (_flip(proof, OFF_WRAP_SG_X + 3), (0, 1, 1)),       # off-curve -> reject without an MSM
        "step prechallenge 0": (_flip(proof, OFF_STEP_PRECHAL + 17), (1, 0, 1)),
        "step prechallenge 1": (_flip(proof, OFF_STEP_PRECHAL + 240 + 100), (1, 1, 0)),
        "step sg 1 y": (_flip(proof, OFF_STEP_SG + 80 + 48 + 9), (1, 1, 0)),
    }
    for mode in (gpu.MODE_PER_PROOF, gpu.MODE_RLC):
        got = gpu.accumulator_check([c[0] for c in cases.values()] + [proof], mode)
        assert got == [c[1] for c in cases.values()] + [(1, 1, 1)], mode
    # swapping the two step accumulators' challenges breaks both
    m = bytearray(proof)
    m[OFF_STEP_PRECHAL : OFF_STEP_PRECHAL + 240], m[OFF_STEP_PRECHAL + 240 : OFF_STEP_PRECHAL + 480] = (
        proof[OFF_STEP_PRECHAL + 240 : OFF_STEP_PRECHAL + 480], proof[OFF_STEP_PRECHAL : OFF_STEP_PRECHAL + 240])
    assert gpu.accumulator_check([bytes(m)]) == [(1, 0, 0)]
    # the negated accumulator point is on the curve but wrong
    y = int.from_bytes(proof[414:446], "little")
    m = bytearray(proof)
    m[414:446] = (pasta.Q - y).to_bytes(32, "little")
    assert gpu.accumulator_check([bytes(m)]) == [(0, 1, 1)]
    # undecodable / empty input: all zero, no exception
    assert gpu.accumulator_check([b"", proof[:100], b"\0" * 48342]) == [(0, 0, 0)] * 3


def test_rlc_batch_returns_exactly_the_per_proof_bits(gpu):
    """Batch of 64 with k corrupted members: random-linear-combination + bisection == per-proof path."""
    proof = golden("mina_state.proof")
    rng = random.Random(0x4D494E41)
    batch, want = [], []
    for i in range(64):
        kind = rng.choice(["ok"] * 9 + ["wrap", "step0", "step1", "both"])
        p = proof
        if kind in ("wrap", "both"):
            p = _flip(p, OFF_WRAP_PRECHAL + rng.randrange(256), rng.randrange(8))
        if kind in ("step0", "both"):
            p = _flip(p, OFF_STEP_PRECHAL + rng.randrange(240), rng.randrange(8))
        if kind == "step1":
            p = _flip(p, OFF_STEP_PRECHAL + 240 + rng.randrange(240), rng.randrange(8))
        batch.append(p)
        want.append((int(kind not in ("wrap", "both")), int(kind not in ("step0", "both")), int(kind != "step1")))
    assert any(w != (1, 1, 1) for w in want)
    b = gpu.Batch(batch)
    assert gpu.accumulator_check(b, gpu.MODE_PER_PROOF) == want
    assert gpu.accumulator_check(b, gpu.MODE_RLC) == want
    # all good / all bad / singleton batches
    assert gpu.accumulator_check([proof] * 7, gpu.MODE_RLC) == [(1, 1, 1)] * 7
    bad = _flip(proof, OFF_WRAP_PRECHAL)
    assert gpu.accumulator_check([bad] * 5, gpu.MODE_RLC) == [(0, 1, 1)] * 5
    assert gpu.accumulator_check([bad], gpu.MODE_RLC) == [(0, 1, 1)]
    assert gpu.accumulator_check([], gpu.MODE_RLC) == []


@pytest.mark.parametrize("n", [130, 333])
def test_group_testing_patterns(gpu, n):
    """The locator / splitting logic of the batched check on batches that are not a multiple of the slice size,
    with the corruption patterns that stress it: exactly one bad proof (first, last, middle: resolved by the
    single-error locator), two identical bad proofs in one slice, neighbours across a slice boundary, a whole
    slice bad, every other proof bad, everything bad.  Expected bits are known by construction."""
    proof = golden("mina_state.proof")
    bad_w = _flip(proof, OFF_WRAP_PRECHAL + 9, 3)          # wrap accumulator wrong
    bad_s = _flip(proof, OFF_STEP_PRECHAL + 240 + 31, 1)   # second step accumulator wrong
    patterns = {
        "single first": {0: bad_w}, "single last": {n - 1: bad_w}, "single middle": {n // 2: bad_s},
        "two identical in one slice": {5: bad_w, 6: bad_w}, "across a slice boundary": {63: bad_w, 64: bad_w, 65: bad_s},
        "one per family": {70: bad_w, 71: bad_s},
        "whole slice": {i: bad_w for i in range(64, 128)},
        "every other": {i: (bad_w if i % 4 else bad_s) for i in range(0, n, 2)},
        "all": {i: bad_w for i in range(n)},
    }
    for name, bad in patterns.items():
        batch = [bad.get(i, proof) for i in range(n)]
        want = [(0, 1, 1) if batch[i] is bad_w else (1, 1, 0) if batch[i] is bad_s else (1, 1, 1) for i in range(n)]
        assert gpu.accumulator_check(gpu.Batch(batch), gpu.MODE_RLC) == want, name


# ---- the boundary with a device present -------------------------------------------------------------------------------
def test_state_ffi_runs_every_built_stage_and_still_refuses_a_partial_accept(gpu):
    S = gpu.STAGES
    assert gpu.verify_mina_state(golden("mina_state.proof"), golden("mina_state.pub")) is False
    rep = gpu.last_stages()
    built = S["lengths"] | S["decode_proof"] | S["decode_pub"] | S["pub_structure"] | S["consensus"] | S["accumulator"] | S["step_accumulators"]
    assert rep.passed == built and rep.failed == 0 and rep.unavailable == S["pub_hashes"] | S["kimchi"]


def test_state_batch_stage_reports(gpu):
    S = gpu.STAGES
    proof, pub = golden("mina_state.proof"), golden("mina_state.pub")
    o = wire.decode_state_proof(proof)
    ledger_off = None
    # a ledger hash inside the pub input (bytes 545..1057): flipping it breaks the structural comparison only
    bad_pub = _flip(pub, 545 + 32 * 3 + 1)
    # swap candidate tip and bridge tip heights: consensus must say Bridge
    st15 = o["candidate_chain_states"][15]
    proofs = [proof, _flip(proof, OFF_WRAP_PRECHAL + 1), proof, proof[:20000], proof]
    pubs = [pub, pub, bad_pub, pub, golden("mina_state_bad_hash.pub")]
    for mode in (gpu.MODE_PER_PROOF, gpu.MODE_RLC):
        accept, reps = gpu.verify_state_stages(proofs, pubs, mode)
        assert accept == [0] * 5
        assert reps[0].failed == 0 and reps[0].passed & S["accumulator"] and reps[0].passed & S["step_accumulators"]
        assert reps[1].failed == S["accumulator"] and reps[1].passed & S["step_accumulators"]
        assert reps[2].failed == S["pub_structure"] and reps[2].passed & S["accumulator"]
        assert reps[3].failed == S["decode_proof"] and reps[3].passed == S["lengths"]
        assert reps[4].failed == S["decode_pub"]
    assert gpu.verify_state_batch(proofs, pubs) == [0] * 5
    assert gpu.verify_state_batch([], []) == []


def test_state_consensus_stage_rejects_a_worse_candidate(gpu):
    S = gpu.STAGES
    proof, pub = golden("mina_state.proof"), golden("mina_state.pub")
    o = wire.decode_state_proof(proof)
    # make the candidate tip shorter than the bridge tip (blockchain_length is a u32 inside state 15)
    st = o["candidate_chain_states"][15]
    enc = wire.encode_protocol_state(st)
    st2 = dict(st)
    import copy

    st2 = copy.deepcopy(st)
    st2["body"]["consensus_state"]["blockchain_length"] = o["bridge_tip_state"]["body"]["consensus_state"]["blockchain_length"] - 1
    enc2 = wire.encode_protocol_state(st2)
    assert len(enc) == len(enc2)
    mutated = proof[: st["_start"]] + enc2 + proof[st["_end"] :]
    accept, reps = gpu.verify_state_stages([mutated], [pub])
    assert accept == [0] and reps[0].failed == S["consensus"]


def test_concurrent_ffi_callers_are_coalesced_and_agree(gpu):
    import threading

    S = gpu.STAGES
    proof, pub = golden("mina_state.proof"), golden("mina_state.pub")
    bad = _flip(proof, OFF_WRAP_PRECHAL + 2)
    out = [None] * 24

    def work(i):
        gpu.verify_mina_state(bad if i % 3 == 0 else proof, pub)
        out[i] = gpu.last_stages().failed

    ts = [threading.Thread(target=work, args=(i,)) for i in range(24)]
    [t.start() for t in ts]
    [t.join(300) for t in ts]
    assert out == [S["accumulator"] if i % 3 == 0 else 0 for i in range(24)]


# ---- K3: Poseidon kernels on an arbitrary table (constants unavailable => parity unpinned) --------------------------------
@pytest.mark.parametrize("fid", [0, 1])
def test_poseidon_permutation_kernel_matches_oracle(gpu, fid):
    mod = FIELD_MOD[fid]
    table = oposeidon.random_table(mod, 500 + fid)
    tb = oposeidon.table_bytes(table)
    rng = random.Random(8)
    states = [[rng.randrange(mod) for _ in range(3)] for _ in range(300)] + [[0, 0, 0], [mod - 1, mod - 1, mod - 1]]
    sb = b"".join(cref.ints_to_bytes(s) for s in states)
    assert gpu.poseidon_permute(fid, tb, sb) == cref.poseidon_permute(fid, tb, sb)


def test_merkle_fold_kernel_matches_oracle(gpu):
    table = oposeidon.random_table(pasta.P, 900)
    tb = oposeidon.table_bytes(table)
    rng = random.Random(12)
    acct = wire.decode_account_proof(golden("mina_account.proof"))
    paths = [acct["merkle_path"], [(0, 0), (1, 0)], [], [(rng.randrange(2), rng.randrange(pasta.P)) for _ in range(20)]]
    leaves = [rng.randrange(pasta.P), 0, 5, rng.randrange(pasta.P)]
    want = [oposeidon.merkle_root(table, leaves[i], paths[i], pasta.P) for i in range(4)]
    roots = list(want)
    roots[3] = (roots[3] + 1) % pasta.P  # one wrong root
    ok, folded = gpu.merkle_fold(tb, paths, leaves, roots)
    assert folded == want and ok == [1, 1, 1, 0]


def test_poseidon_is_untrusted_without_a_constants_table(gpu):
    import os

    from conftest import ROOT

    have = os.path.exists(os.path.join(ROOT, "mina_bridge_b200", "data", "poseidon_fp_kimchi.bin"))
    assert gpu.poseidon_trusted() == have


# ---- a9 shape and K5 ------------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("cid", [0, 1])
def test_msm_over_srs_prefix_plus_per_proof_points(gpu, cid):
    """The IPA final-check MSM mixes the resident SRS prefix with ~80 per-proof points."""
    rng = random.Random(90 + cid)
    sm, fm = (pasta.Q, pasta.P) if cid == 0 else (pasta.P, pasta.Q)
    n = 32768 if cid == 0 else 65536
    srs, h = gpu.srs_points(cid, 0, n, True)
    s1 = cref.ints_to_bytes([rng.randrange(sm) for _ in range(n)])
    extra_pts = h + gpu.srs_points(cid, 5, 79)  # any on-curve points will do: h plus 79 more
    s2 = cref.ints_to_bytes([rng.randrange(sm) for _ in range(79)] + [0])
    want, _ = cref.msm(cid, s1 + s2, srs + extra_pts, 8)
    assert gpu.msm_srs_plus(cid, s1, s2, extra_pts) == want
    # either part empty; cancellation to the identity
    w1, _ = cref.msm(cid, s1, srs, 8)
    assert gpu.msm_srs_plus(cid, s1, b"", b"") == w1
    w2, _ = cref.msm(cid, s2, extra_pts, 2)
    assert gpu.msm_srs_plus(cid, b"", s2, extra_pts) == w2
    g0 = srs[:64]
    k = rng.randrange(sm)
    assert gpu.msm_srs_plus(cid, cref.ints_to_bytes([k]), cref.ints_to_bytes([sm - k]), g0) == b"\0" * 64
    with pytest.raises(gpu.MinaB200Error, match="not on the curve"):
        gpu.msm_srs_plus(cid, s1, cref.ints_to_bytes([1]), b"\1" + b"\0" * 63)


@pytest.mark.parametrize("fid", [0, 1])
def test_combined_inner_product(gpu, fid):
    rng = random.Random(95 + fid)
    m = FIELD_MOD[fid]
    nproofs, npolys, npts = 5, 47, 2  # 47 polynomials at zeta, zeta*omega
    evals = [[[rng.randrange(m) for _ in range(npts)] for _ in range(npolys)] for _ in range(nproofs)]
    scales = [(rng.randrange(m), rng.randrange(m)) for _ in range(nproofs)]
    scales[0] = (0, 5)
    scales[1] = (7, 0)
    got = gpu.combined_inner_product(fid, b"".join(cref.ints_to_bytes(row) for p in evals for row in p),
                                     b"".join(cref.ints_to_bytes(s) for s in scales), npolys, npts)
    for p in range(nproofs):
        v, u = scales[p]
        want = sum(pow(v, i, m) * sum(pow(u, j, m) * evals[p][i][j] for j in range(npts)) for i in range(npolys)) % m
        assert int.from_bytes(got[32 * p : 32 * p + 32], "little") == want
