"""CPU tier: the drop-in boundary (negative paths), the C++ wire decoders, the consensus rule, the VK
loader and the host Poseidon sponge -- none of which needs a device.

These mirror the reference's own tests:
  AL/operator/mina/lib/src/lib.rs:246-331              (state: bad pub, empty proof, empty pub, oversize lens)
  AL/operator/mina_account/lib/src/lib.rs:111-187      (account twins)
  AL/operator/mina/lib/src/consensus_state.rs:170-303  (four fork-choice cases, rebuilt by mutating the fixture)
  AL/operator/mina/lib/src/verifier_index.rs:278-286   (VK loads)
  AL/operator/mina_account/lib/src/merkle_verifier.rs:43-58 (Poseidon KAT; skipped without a constants table)
"""
import copy
import json
import os
import random
import subprocess
import threading

import pytest

from oracle import cref, pasta, poseidon as oposeidon, wire
from conftest import GOLDEN, ROOT, golden

S = None  # stage bit table, filled from the package


@pytest.fixture(scope="module")
def mb(native):
    global S
    S = native.STAGES
    return native


# ---- exported boundary --------------------------------------------------------------------------------
def test_drop_in_symbols_are_exported(mb):
    out = subprocess.run(["nm", "-D", "--defined-only", mb.library_path()], capture_output=True, text=True, check=True).stdout
    names = {line.split()[-1] for line in out.splitlines() if line.strip()}
    for sym in ("verify_mina_state_ffi", "verify_account_inclusion_ffi", "verify_mina_state_batch_ffi",
                "verify_account_inclusion_batch_ffi", "mina_verifier_init", "mina_verifier_shutdown"):
        assert sym in names
    # the library names the reference's cgo LDFLAGS link (mina.go:3-8, mina_account.go:3-8)
    libdir = os.path.dirname(mb.library_path())
    for alias in ("libmina_state_verifier_ffi.so", "libmina_account_verifier_ffi.so"):
        assert os.path.realpath(os.path.join(libdir, alias)) == os.path.realpath(mb.library_path())


def test_a_c_program_links_against_the_reference_header_names(mb, tmp_path):
    """What cgo does: include the header, link the .so by its reference name, call the symbol."""
    src = tmp_path / "t.c"
    src.write_text(
        '#include "mina_verifier.h"\n#include "mina_account_verifier.h"\n#include <stdio.h>\n'
        "static unsigned char proof[MINA_STATE_MAX_PROOF_SIZE], pub[MINA_STATE_MAX_PUB_INPUT_SIZE];\n"
        "int main(void){ int a = verify_mina_state_ffi(proof, MINA_STATE_MAX_PROOF_SIZE + 1, pub, 10);\n"
        " int b = verify_account_inclusion_ffi(proof, 10, pub, MINA_ACCOUNT_MAX_PUB_INPUT_SIZE + 1);\n"
        ' printf("%d %d\\n", a, b); return 0; }\n'
    )
    libdir = os.path.dirname(mb.library_path())
    exe = tmp_path / "t"
    subprocess.run(["gcc", "-I", os.path.join(ROOT, "include"), str(src), "-o", str(exe),
                    os.path.join(libdir, "libmina_state_verifier_ffi.so"), "-Wl,-rpath," + libdir], check=True)
    out = subprocess.run([str(exe)], capture_output=True, text=True, check=True).stdout
    assert out.split() == ["0", "0"]


# ---- negative paths of the FFI (reference tests 2-6 and 8-11) ----------------------------------------------
def test_state_bad_hash_pub_rejects_at_bool_decode(mb):
    assert mb.verify_mina_state(golden("mina_state.proof"), golden("mina_state_bad_hash.pub")) is False
    rep = mb.last_stages()
    assert rep.failed == S["decode_pub"] and rep.passed == S["lengths"] | S["decode_proof"]  # Q10: first byte 0x5d


def test_state_empty_proof_rejects(mb):
    n = len(golden("mina_state.proof"))
    assert mb.verify_mina_state(b"\0" * n, golden("mina_state.pub")) is False
    assert mb.last_stages().failed == S["decode_proof"]


def test_state_empty_pub_does_not_verify(mb):
    assert mb.verify_mina_state(golden("mina_state.proof"), b"\0" * len(golden("mina_state.pub"))) is False
    rep = mb.last_stages()
    # an all-zero pub decodes (is_devnet = false, zero hashes); it is the ledger-hash comparison that fails
    assert rep.failed & S["pub_structure"] and rep.passed & S["decode_pub"]


def test_state_oversize_lengths_reject(mb):
    p, q = golden("mina_state.proof"), golden("mina_state.pub")
    assert mb.verify_mina_state(p, q, proof_len=mb.MAX_STATE_PROOF_SIZE + 1) is False
    assert mb.last_stages().failed == S["lengths"] and mb.last_stages().passed == 0
    assert mb.verify_mina_state(p, q, pub_len=mb.MAX_PUB_INPUT_SIZE + 1) is False
    assert mb.last_stages().failed == S["lengths"]
    assert mb.verify_mina_state(p, q, proof_len=2**63) is False  # a usize, not an unsigned int


def test_account_negative_paths(mb):
    p, q = golden("mina_account.proof"), golden("mina_account.pub")
    assert mb.verify_account_inclusion(b"\0" * len(p), q) is False
    # zeroed proof: path length 0, then an all-zero account ... `token_id` BigInt length 0 != 32
    assert mb.last_stages().failed == S["decode_proof"]
    assert mb.verify_account_inclusion(p, b"\0" * len(q)) is False
    assert mb.verify_account_inclusion(p, q, proof_len=mb.MAX_ACCOUNT_PROOF_SIZE + 1) is False
    assert mb.last_stages().failed == S["lengths"]
    assert mb.verify_account_inclusion(p, q, pub_len=mb.MAX_PUB_INPUT_SIZE + 1) is False
    assert mb.last_stages().failed == S["lengths"]


def test_valid_vectors_are_never_accepted_on_a_partial_check(mb):
    """Vectors (1) and (7) are `true` in the reference.  Until every stage exists this build must say false
    and say why: the unbuilt stages are listed as unavailable, everything that ran passed."""
    assert mb.verify_mina_state(golden("mina_state.proof"), golden("mina_state.pub")) is False
    rep = mb.last_stages()
    assert rep.unavailable == S["pub_hashes"] | S["kimchi"]
    assert rep.passed & (S["lengths"] | S["decode_proof"] | S["decode_pub"] | S["pub_structure"] | S["consensus"]) == (
        S["lengths"] | S["decode_proof"] | S["decode_pub"] | S["pub_structure"] | S["consensus"])
    assert rep.failed & ~S["internal_error"] == 0  # internal_error only when there is no device (this tier)
    assert mb.verify_account_inclusion(golden("mina_account.proof"), golden("mina_account.pub")) is False
    rep = mb.last_stages()
    assert rep.failed == 0 and rep.unavailable == S["account_leaf"] | S["merkle"]
    assert rep.passed == S["lengths"] | S["decode_proof"] | S["decode_pub"] | S["account_abi"]


# ---- account ABI re-encoding (mina_account lib.rs:54-66, core/src/sol/account.rs) ---------------------------------
def test_account_abi_encoding_matches_the_fixture(mb):
    """The reference rejects unless abi_encode(Account::try_from(&proof.account)) == pub.encoded_account; its own
    fixture pair therefore pins the encoder: 3 456 bytes, identical."""
    enc = mb.host_account_abi_encode(golden("mina_account.proof"))
    assert enc is not None and len(enc) == 3456
    assert enc == golden("mina_account.pub")[40:]


def _account_fields():
    """byte offsets inside mina_account.proof: the account starts at 1 548"""
    a = 1548
    return {"public_key_x": a + 8, "is_odd": a + 40, "token_id": a + 49, "symbol_len": a + 81, "balance": a + 89,
            "nonce": a + 97, "receipt": a + 109, "delegate_tag": a + 141, "delegate_x": a + 150, "voting_for": a + 191,
            "timing_tag": a + 223, "perm0": a + 227, "zkapp_tag": a + 283}


def test_account_abi_stage_rejects_a_mismatch(mb):
    """one flipped byte in any account field of the proof, or in the encoded account of the public input, fails
    the ABI stage (and nothing else) -- the reference's `expected_encoded_account != encoded_account` branch"""
    p, q = golden("mina_account.proof"), golden("mina_account.pub")
    f = _account_fields()
    for name in ("public_key_x", "token_id", "balance", "nonce", "receipt", "delegate_x", "voting_for"):
        m = bytearray(p)
        m[f[name]] ^= 1
        assert mb.verify_account_inclusion(bytes(m), q) is False
        rep = mb.last_stages()
        assert rep.failed == S["account_abi"], name
    for off in (40 + 32 + 5, 40 + 5 * 32 + 31, len(q) - 1):
        m = bytearray(q)
        m[off] ^= 1
        assert mb.verify_account_inclusion(p, bytes(m)) is False
        assert mb.last_stages().failed == S["account_abi"]


def test_account_abi_encoding_of_variants(mb):
    """Branches the fixture does not take (account.rs:57-69, 73-110), checked against a Python restatement of the
    Solidity ABI rules: no delegate -> (0, isOdd = true); a timed account; a non-empty token symbol (dynamic tail
    grows, zkapp offset moves); permissions other than Signature."""
    p = bytearray(golden("mina_account.proof"))
    f = _account_fields()
    base = mb.host_account_abi_encode(bytes(p))
    word = lambda v: v.to_bytes(32, "big")
    # (a) delegate = None: drop the 41-byte public key
    m = bytes(p[:f["delegate_tag"]]) + b"\0" + bytes(p[f["delegate_tag"] + 42:])
    enc = mb.host_account_abi_encode(m)
    want = bytearray(base)
    want[32 * 8:32 * 9] = b"\0" * 32
    want[32 * 9:32 * 10] = word(1)
    assert enc == bytes(want)
    # (b) timed account
    timed = (1).to_bytes(4, "little") + (1000).to_bytes(8, "little") + (7).to_bytes(4, "little") + (11).to_bytes(8, "little") + \
        (13).to_bytes(4, "little") + (17).to_bytes(8, "little")
    m = bytes(p[:f["timing_tag"]]) + timed + bytes(p[f["timing_tag"] + 4:])
    enc = mb.host_account_abi_encode(m)
    want = bytearray(base)
    for i, v in enumerate((1000, 7, 11, 13, 17)):
        want[32 * (11 + i):32 * (12 + i)] = word(v)
    assert enc == bytes(want)
    # (c) token symbol "MINA": the string tail gains one data word and the zkapp offset moves by 32
    m = bytes(p[:f["symbol_len"]]) + (4).to_bytes(8, "little") + b"MINA" + bytes(p[f["symbol_len"] + 8:])
    enc = mb.host_account_abi_encode(m)
    want = bytes(base[:32 * 30]) + word(0x3e0 + 32) + word(4) + b"MINA" + b"\0" * 28 + bytes(base[32 * 32:])
    assert enc == want
    # ... and one that is not UTF-8 is a conversion failure
    m = bytes(p[:f["symbol_len"]]) + (2).to_bytes(8, "little") + b"\xff\xfe" + bytes(p[f["symbol_len"] + 8:])
    assert mb.host_account_abi_encode(m) is None
    # (d) permissions: edit_state = Impossible (tag 4)
    m = bytearray(p)
    m[f["perm0"]:f["perm0"] + 4] = (4).to_bytes(4, "little")
    enc = mb.host_account_abi_encode(bytes(m))
    want = bytearray(base)
    want[32 * 16:32 * 17] = word(4)
    assert enc == bytes(want)


# ---- wire writers (core/src/aligned.rs:33-49) ---------------------------------------------------
@pytest.mark.parametrize("kind,name", [(0, "mina_state.proof"), (1, "mina_state.pub"), (1, "mina_state_bad_hash.pub"),
                                       (2, "mina_account.proof"), (3, "mina_account.pub")])
def test_writers_reproduce_the_fixtures(mb, kind, name):
    """decode -> encode through the C++ writers gives back the reference's committed bytes exactly"""
    data = golden(name)
    if name == "mina_state_bad_hash.pub":
        assert mb.host_reencode(kind, data) is None  # first byte 0x5d is not a bool
        return
    assert mb.host_reencode(kind, data) == data


def test_writers_round_trip_mutated_proofs(mb):
    """the writers are what synthetic batches are made of: mutate -> encode -> decode must be stable"""
    data = bytearray(golden("mina_state.proof"))
    for off in (80, 380, 5000, 14000, 30000):
        data[off] ^= 0x10
        enc = mb.host_reencode(0, bytes(data))
        assert enc == bytes(data)


def test_concurrent_callers_do_not_deadlock(mb):
    p, q = golden("mina_state.proof"), golden("mina_state_bad_hash.pub")
    ap, aq = golden("mina_account.proof"), golden("mina_account.pub")
    results = []

    def work(i):
        results.append(mb.verify_account_inclusion(ap, aq) if i % 2 else mb.verify_mina_state(p, q))

    ts = [threading.Thread(target=work, args=(i,)) for i in range(16)]
    [t.start() for t in ts]
    [t.join(60) for t in ts]
    assert len(results) == 16 and not any(results)


# ---- wire decoders against the oracle ------------------------------------------------------------------------
def test_cpp_decoders_consume_the_fixtures_byte_exactly(mb):
    d = golden("mina_state.proof")
    o = wire.decode_state_proof(d)
    s = mb.host_decode(0, d)
    assert s.consumed == o["_consumed"] == len(d) == 48342 and s.proof_end == o["_proof_end"] == 13849
    pr = o["candidate_tip_proof"]
    assert s.n_step_comms == 2 and s.n_lr == len(pr["proof"]["lr"]) == 15
    assert bytes(s.wrap_sg) == cref.points_to_bytes([pr["wrap_challenge_polynomial_commitment"]])
    for k in range(2):
        assert bytes(s.step_sg[k]) == cref.points_to_bytes([pr["step_challenge_polynomial_commitments"][k]])
    states = o["candidate_chain_states"] + [o["bridge_tip_state"]]
    for i, st in enumerate(states):
        cs = st["body"]["consensus_state"]
        assert (s.state_begin[i], s.state_end[i]) == (st["_start"], st["_end"])
        assert s.blockchain_length[i] == cs["blockchain_length"] and s.epoch_count[i] == cs["epoch_count"]
        assert s.curr_global_slot[i] == cs["curr_global_slot"]["slot_number"] and s.min_window_density[i] == cs["min_window_density"]
        assert int.from_bytes(bytes(s.previous_state_hash[i]), "little") == st["previous_state_hash"]
        assert int.from_bytes(bytes(s.first_pass_ledger[i]), "little") == st["body"]["blockchain_state"]["ledger_proof_statement"]["target"]["first_pass_ledger"]
    q = golden("mina_state.pub")
    sq, oq = mb.host_decode(1, q), wire.decode_state_pub(q)
    assert sq.consumed == 1057 and sq.is_devnet == 1 and int.from_bytes(bytes(sq.hash0), "little") == oq["bridge_tip_state_hash"]
    for i in range(16):
        assert int.from_bytes(bytes(sq.previous_state_hash[i]), "little") == oq["candidate_chain_state_hashes"][i]
        assert int.from_bytes(bytes(sq.first_pass_ledger[i]), "little") == oq["candidate_chain_ledger_hashes"][i]
    assert mb.host_decode(1, golden("mina_state_bad_hash.pub")) is None
    a = golden("mina_account.proof")
    sa, oa = mb.host_decode(2, a), wire.decode_account_proof(a)
    assert sa.merkle_depth == len(oa["merkle_path"]) == 35 and sa.balance == oa["account"]["balance"] and sa.nonce == oa["account"]["nonce"]
    assert sa.has_zkapp == 0
    b = golden("mina_account.pub")
    sb, ob = mb.host_decode(3, b), wire.decode_account_pub(b)
    assert sb.consumed == 3496 and sb.encoded_account_len == len(ob["encoded_account"]) == 3456
    assert int.from_bytes(bytes(sb.hash0), "little") == ob["ledger_hash"]


def test_cpp_and_oracle_decoders_agree_on_truncated_and_corrupted_input(mb):
    rng = random.Random(5)
    for kind, name, dec in ((0, "mina_state.proof", wire.decode_state_proof), (1, "mina_state.pub", wire.decode_state_pub),
                            (2, "mina_account.proof", wire.decode_account_proof), (3, "mina_account.pub", wire.decode_account_pub)):
        d = golden(name)
        cuts = sorted({0, 1, 7, 8, 9, 40, len(d) - 1, len(d)} | {rng.randrange(len(d)) for _ in range(40)})
        for cut in cuts:
            try:
                dec(d[:cut])
                want = True
            except wire.DecodeError:
                want = False
            assert (mb.host_decode(kind, d[:cut]) is not None) == want, (name, cut)
        for _ in range(60):  # single-byte corruption: both sides accept or both reject
            m = bytearray(d)
            m[rng.randrange(len(m))] ^= 1 << rng.randrange(8)
            try:
                dec(bytes(m))
                want = True
            except wire.DecodeError:
                want = False
            assert (mb.host_decode(kind, bytes(m)) is not None) == want, name


# ---- consensus rule: the reference's four cases (consensus_state.rs:170-303) -----------------------------------
SUB_WINDOWS_PER_WINDOW = 11


@pytest.fixture(scope="module")
def tips():
    o = wire.decode_state_proof(golden("mina_state.proof"))
    return o["bridge_tip_state"], o["candidate_chain_states"][-1]  # (old tip, new tip)


def _force_long_fork(old_tip, new_tip):
    cs = old_tip["body"]["consensus_state"]
    cs["epoch_count"] = new_tip["body"]["consensus_state"]["epoch_count"]
    cs["staking_epoch_data"]["lock_checkpoint"] = (new_tip["body"]["consensus_state"]["staking_epoch_data"]["lock_checkpoint"] - 1) % pasta.P
    cs["sub_window_densities"] = cs["sub_window_densities"][:-1] + [1]


def test_new_mina_state_passes_consensus_checks(mb, tips):
    old, new = tips
    assert mb.host_select_secure_chain(wire.encode_protocol_state(new), wire.encode_protocol_state(old)) == 1


def test_old_mina_state_fails_consensus_checks(mb, tips):
    old, new = tips
    assert mb.host_select_secure_chain(wire.encode_protocol_state(old), wire.encode_protocol_state(new)) == 0


def test_candidate_state_with_smaller_global_slot_than_tip_state(mb, tips):
    new_tip, old_tip = copy.deepcopy(tips[0]), copy.deepcopy(tips[1])  # the reference swaps the roles here
    _force_long_fork(old_tip, new_tip)
    old_tip["body"]["consensus_state"]["curr_global_slot"]["slot_number"] = (
        new_tip["body"]["consensus_state"]["curr_global_slot"]["slot_number"] - (SUB_WINDOWS_PER_WINDOW + 2))
    assert mb.host_select_secure_chain(wire.encode_protocol_state(new_tip), wire.encode_protocol_state(old_tip)) == 0


def test_candidate_state_with_less_sub_windows_densities_than_sub_windows_per_window(mb, tips):
    new_tip, old_tip = copy.deepcopy(tips[0]), copy.deepcopy(tips[1])
    _force_long_fork(old_tip, new_tip)
    old_tip["body"]["consensus_state"]["curr_global_slot"]["slot_number"] = (
        new_tip["body"]["consensus_state"]["curr_global_slot"]["slot_number"] + SUB_WINDOWS_PER_WINDOW)
    new_tip["body"]["consensus_state"]["sub_window_densities"] = new_tip["body"]["consensus_state"]["sub_window_densities"][:-2]
    assert mb.host_select_secure_chain(wire.encode_protocol_state(new_tip), wire.encode_protocol_state(old_tip)) == 0


def test_consensus_quirks(mb, tips):
    old, new = copy.deepcopy(tips[0]), copy.deepcopy(tips[1])
    # different protocol constants: the reference returns Err (-> reject)
    odd = copy.deepcopy(new)
    odd["body"]["constants"]["k"] += 1
    with pytest.raises(mb.MinaB200Error, match="constants differ"):
        mb.host_select_secure_chain(wire.encode_protocol_state(odd), wire.encode_protocol_state(old))
    # Q3: same height -> Blake2b-512 of last_vrf_output decides, compared as hex (= bytewise)
    same = copy.deepcopy(new)
    same["body"]["consensus_state"]["blockchain_length"] = old["body"]["consensus_state"]["blockchain_length"]
    import hashlib

    hv = lambda st: hashlib.blake2b(st["body"]["consensus_state"]["last_vrf_output"]).hexdigest()
    want = 1 if hv(same) > hv(old) else 0
    assert mb.host_select_secure_chain(wire.encode_protocol_state(same), wire.encode_protocol_state(old)) == want
    # exact tie (same state): needs the Poseidon state hash -> undecidable in this build, reported as such
    with pytest.raises(mb.MinaB200Error, match="needs state hash"):
        mb.host_select_secure_chain(wire.encode_protocol_state(old), wire.encode_protocol_state(old))


# ---- verification keys ------------------------------------------------------------------------------------------
@pytest.mark.parametrize("chain", ["devnet", "mainnet"])
def test_vk_loader(mb, chain):
    path = os.path.join(ROOT, "mina_bridge_b200", "data", chain + "_vk.json")
    raw = json.load(open(path))
    v = mb.host_vk_load(path)
    c = raw["commitments"]
    flat = c["sigma_comm"] + c["coefficients_comm"] + [c[k] for k in ("generic_comm", "psm_comm", "complete_add_comm", "mul_comm", "emul_comm", "endomul_scalar_comm")]
    assert v["commitments"] == [(int(x, 16) % pasta.P, int(y, 16) % pasta.P) for x, y in flat]
    for x, y in v["commitments"]:  # all 28 on Pallas (the loader also checks)
        assert (y * y - x * x * x - 5) % pasta.P == 0
    assert v["shifts"] == [int(s, 16) % pasta.Q for s in raw["index"]["shifts"]]
    assert (v["log_size_of_group"], v["max_poly_size"], v["public"], v["prev_challenges"]) == (14, 32768, 40, 2)
    # K-G: the derived domain generator equals the one the file records and has order exactly 2^14
    g = v["group_gen"]
    assert g == int(raw["index"]["domain"]["group_gen"], 16) and pow(g, 1 << 14, pasta.Q) == 1 and pow(g, 1 << 13, pasta.Q) != 1
    n = 1 << 14
    w3 = pow(g, n - 3, pasta.Q)
    w2, w1 = g * w3 % pasta.Q, g * g * w3 % pasta.Q
    assert v["zk_w3"] == w3
    assert v["zkpm"] == [(-w1 * w2 * w3) % pasta.Q, (w1 * w2 + w1 * w3 + w2 * w3) % pasta.Q, (-w1 - w2 - w3) % pasta.Q, 1]
    assert v["endo"] == pasta.OMEGA_Q  # endos::<Vesta>().0 (verifier_index.rs:169)


def test_vk_devnet_and_mainnet_differ_only_where_the_survey_says(mb):
    d = mb.host_vk_load(os.path.join(ROOT, "mina_bridge_b200", "data", "devnet_vk.json"))
    m = mb.host_vk_load(os.path.join(ROOT, "mina_bridge_b200", "data", "mainnet_vk.json"))
    diff = [i for i in range(28) if d["commitments"][i] != m["commitments"][i]]
    assert diff == [7 + 0, 7 + 5]  # coefficients_comm[0] and [5]


def test_vk_loader_rejects_malformed_files(mb, tmp_path):
    good = json.load(open(os.path.join(ROOT, "mina_bridge_b200", "data", "devnet_vk.json")))
    for mutate in (lambda j: j["commitments"].pop("mul_comm"), lambda j: j["commitments"]["sigma_comm"].pop(),
                   lambda j: j["index"].__setitem__("shifts", j["index"]["shifts"][:6]),
                   lambda j: j["commitments"].__setitem__("psm_comm", ["0xZZ", "0x00"])):
        j = copy.deepcopy(good)
        mutate(j)
        p = tmp_path / "bad.json"
        p.write_text(json.dumps(j))
        with pytest.raises(mb.MinaB200Error):
            mb.host_vk_load(str(p))
    (tmp_path / "trunc.json").write_text(json.dumps(good)[:200])
    with pytest.raises(mb.MinaB200Error):
        mb.host_vk_load(str(tmp_path / "trunc.json"))


# ---- Poseidon (table-driven; constants unavailable => parity unpinned) ----------------------------------------------
def test_host_poseidon_matches_the_oracle_on_an_arbitrary_table(mb):
    rng = random.Random(11)
    for fid, mod in ((0, pasta.P), (1, pasta.Q)):
        table = oposeidon.random_table(mod, 100 + fid)
        tb = oposeidon.table_bytes(table)
        states = [[rng.randrange(mod) for _ in range(3)] for _ in range(5)] + [[0, 0, 0], [mod - 1, 1, 0]]
        sb = b"".join(cref.ints_to_bytes(s) for s in states)
        want = b"".join(cref.ints_to_bytes(oposeidon.permute(table, s, mod)) for s in states)
        assert mb.host_poseidon_permute(fid, tb, sb) == want
        assert cref.poseidon_permute(fid, tb, sb) == want  # the C oracle agrees with the Python one


def test_host_sponge_and_hash_with_kimchi_match_the_oracle(mb):
    table = oposeidon.random_table(pasta.P, 7)
    tb = oposeidon.table_bytes(table)
    rng = random.Random(3)
    for prefix, n in (("MinaMklTree000", 2), ("MinaMklTree034", 2), ("MinaProtoState", 2), ("CodaReceiptUC", 0), ("x", 1), ("MinaAccount", 5), ("a" * 20, 7)):
        xs = [rng.randrange(pasta.P) for _ in range(n)]
        assert mb.host_hash_with_kimchi(tb, prefix, xs) == oposeidon.hash_with_kimchi(table, prefix, xs, pasta.P)
    with pytest.raises(mb.MinaB200Error):
        mb.host_hash_with_kimchi(tb, "a" * 21, [1])  # prefixes are at most 20 bytes


def test_poseidon_reference_kat(mb):
    """K-E (merkle_verifier.rs:43-58).  Enforced the moment a constants table is present."""
    path = os.path.join(ROOT, "mina_bridge_b200", "data", "poseidon_fp_kimchi.bin")
    if not os.path.exists(path):
        pytest.skip("Poseidon kimchi constants unavailable (absent from /root/reference and this image): parity unpinned")
    tb = open(path, "rb").read()
    table = [int.from_bytes(tb[32 * i : 32 * i + 32], "little") for i in range(174)]
    assert oposeidon.merkle_root(table, 0, [(0, 0), (1, 0)], pasta.P) == oposeidon.KAT_ROOT
    h0 = mb.host_hash_with_kimchi(tb, "MinaMklTree000", [0, 0])
    assert mb.host_hash_with_kimchi(tb, "MinaMklTree001", [0, h0]) == oposeidon.KAT_ROOT


def test_poseidon_table_tool_refuses_a_table_that_fails_the_kat(mb, tmp_path):
    """tools/make_poseidon_table.py is the one door through which constants enter the library: it must extract 9 + 165
    literals in order, reject a wrong field, and never write a table that fails merkle_verifier.rs:43-58."""
    import sys

    rng = random.Random(11)
    nums = [rng.randrange(10 ** 70, pasta.P) for _ in range(174)]
    src = tmp_path / "fp_kimchi.rs"
    src.write_text("\n".join('Fp::from_str("%d").unwrap(), // 0x%064x' % (x, 7) if i % 2 else "field_from_hex(\"0x%064x\")" % x for i, x in enumerate(nums)))
    tool = os.path.join(ROOT, "tools", "make_poseidon_table.py")
    out = os.path.join(ROOT, "mina_bridge_b200", "data", "poseidon_fp_kimchi.bin")
    existed = os.path.exists(out)
    r = subprocess.run([sys.executable, tool, "fp", str(src)], capture_output=True, text=True)
    assert r.returncode != 0 and "KAT" in (r.stdout + r.stderr)
    assert os.path.exists(out) == existed  # nothing was written
    # literals of the other field's size are refused before the KAT
    src.write_text("\n".join('"%d"' % (pasta.Q - 1 - i) for i in range(174)))
    r = subprocess.run([sys.executable, tool, "fp", str(src)], capture_output=True, text=True)
    assert r.returncode != 0 and "canonical" in (r.stdout + r.stderr)
