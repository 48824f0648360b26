"""CPU oracle (TEST INFRASTRUCTURE ONLY) -- bincode decoders for the two proof formats.

Restates the serde/bincode 1.3 layout of the reference wire types:
  core/src/proof/state_proof.rs:10-41   (MinaStateProof, MinaStatePubInputs)
  core/src/proof/account_proof.rs:9-35  (MerkleNode, MinaAccountProof, MinaAccountPubInputs)
  core/src/sol/serialization.rs:11-86   (SolSerialize: bare 32-byte little-endian field elements)
and of the un-vendored mina-p2p-messages 0.6.4 types they embed.

Only tests/, smoke() and bench.py's cpu_baseline leg may import this.  Parity: pinned by
byte-exact consumption of the five fixtures under tests/golden/ (48342/1057/1832/3496 bytes).
"""
from __future__ import annotations

import struct

MAX_STATE_PROOF = 48 * 1024  # AL/operator/mina/lib/src/lib.rs:38
MAX_ACCOUNT_PROOF = 16 * 1024  # AL/operator/mina_account/lib/src/lib.rs:13
MAX_PUB_INPUT = 6 * 1024  # lib.rs:39 / mina_account lib.rs:14
FRONTIER = 16  # core/src/utils/constants.rs:31 BRIDGE_TRANSITION_FRONTIER_LEN


class DecodeError(ValueError):
    pass


class Cur:
    def __init__(self, data: bytes):
        self.d, self.o = data, 0

    def take(self, n: int) -> bytes:
        if n < 0 or self.o + n > len(self.d):
            raise DecodeError("unexpected end of input")
        b = self.d[self.o : self.o + n]
        self.o += n
        return b

    def u8(self):
        return self.take(1)[0]

    def u32(self):
        return struct.unpack("<I", self.take(4))[0]

    def u64(self):
        return struct.unpack("<Q", self.take(8))[0]

    def i64(self):
        return struct.unpack("<q", self.take(8))[0]

    def boolean(self):
        v = self.u8()
        if v > 1:
            raise DecodeError("invalid bool")
        return bool(v)

    def option_tag(self):
        v = self.u8()
        if v > 1:
            raise DecodeError("invalid option tag")
        return v

    def bigint(self) -> int:
        """mina_p2p_messages::bigint::BigInt = u64 length (must be 32) + 32 bytes little-endian."""
        if self.u64() != 32:
            raise DecodeError("bigint length")
        return int.from_bytes(self.take(32), "little")

    def bytestr(self) -> bytes:
        return self.take(self.u64())

    def limbs2(self) -> int:
        lo = self.i64() & 0xFFFFFFFFFFFFFFFF
        hi = self.i64() & 0xFFFFFFFFFFFFFFFF
        return lo | (hi << 64)

    def point(self):
        return (self.bigint(), self.bigint())


def _signed_amount(c: Cur):
    mag = c.u64()
    sgn = c.u32()
    if sgn > 1:
        raise DecodeError("sign tag")
    return (mag, sgn)  # sgn 0 = Pos, 1 = Neg


def _registers(c: Cur):
    r = {}
    r["first_pass_ledger"] = c.bigint()
    r["second_pass_ledger"] = c.bigint()
    r["pending_coinbase_stack"] = {
        "data": c.bigint(),
        "state_init": c.bigint(),
        "state_curr": c.bigint(),
    }
    ls = {}
    ls["stack_frame"] = c.bigint()
    ls["call_stack"] = c.bigint()
    ls["transaction_commitment"] = c.bigint()
    ls["full_transaction_commitment"] = c.bigint()
    ls["excess"] = _signed_amount(c)
    ls["supply_increase"] = _signed_amount(c)
    ls["ledger"] = c.bigint()
    ls["success"] = c.boolean()
    ls["account_update_index"] = c.u32()
    n = c.u64()
    tbl = []
    for _ in range(n):
        inner = c.u64()
        row = []
        for _ in range(inner):
            row.append(c.u32())
        tbl.append(row)
    ls["failure_status_tbl"] = tbl
    ls["will_succeed"] = c.boolean()
    r["local_state"] = ls
    return r


def _epoch_data(c: Cur):
    return {
        "ledger_hash": c.bigint(),
        "ledger_total_currency": c.u64(),
        "seed": c.bigint(),
        "start_checkpoint": c.bigint(),
        "lock_checkpoint": c.bigint(),
        "epoch_length": c.u32(),
    }


def _pubkey(c: Cur):
    return (c.bigint(), c.boolean())


def protocol_state(c: Cur):
    """MinaStateProtocolStateValueStableV2 (field order)."""
    st = {"_start": c.o}
    st["previous_state_hash"] = c.bigint()
    body = {}
    body["genesis_state_hash"] = c.bigint()
    bs = {}
    bs["staged_ledger_hash"] = {
        "ledger_hash": c.bigint(),
        "aux_hash": c.bytestr(),
        "pending_coinbase_aux": c.bytestr(),
        "pending_coinbase_hash": c.bigint(),
    }
    bs["genesis_ledger_hash"] = c.bigint()
    lps = {}
    lps["source"] = _registers(c)
    lps["target"] = _registers(c)
    lps["connecting_ledger_left"] = c.bigint()
    lps["connecting_ledger_right"] = c.bigint()
    lps["supply_increase"] = _signed_amount(c)
    lps["fee_excess"] = [(c.bigint(), _signed_amount(c)), (c.bigint(), _signed_amount(c))]
    bs["ledger_proof_statement"] = lps
    bs["timestamp"] = c.u64()
    bs["body_reference"] = c.bytestr()
    body["blockchain_state"] = bs
    cs = {}
    cs["blockchain_length"] = c.u32()
    cs["epoch_count"] = c.u32()
    cs["min_window_density"] = c.u32()
    cs["sub_window_densities"] = [c.u32() for _ in range(c.u64())]
    cs["last_vrf_output"] = c.bytestr()
    cs["total_currency"] = c.u64()
    tag = c.u32()
    if tag != 0:
        raise DecodeError("slot tag")
    cs["curr_global_slot"] = {"slot_number": c.u32(), "slots_per_epoch": c.u32()}
    tag = c.u32()
    if tag != 0:
        raise DecodeError("slot tag")
    cs["global_slot_since_genesis"] = c.u32()
    cs["staking_epoch_data"] = _epoch_data(c)
    cs["next_epoch_data"] = _epoch_data(c)
    cs["has_ancestor_in_same_checkpoint_window"] = c.boolean()
    cs["block_stake_winner"] = _pubkey(c)
    cs["block_creator"] = _pubkey(c)
    cs["coinbase_receiver"] = _pubkey(c)
    cs["supercharge_coinbase"] = c.boolean()
    body["consensus_state"] = cs
    body["constants"] = {
        "k": c.u32(),
        "slots_per_epoch": c.u32(),
        "slots_per_sub_window": c.u32(),
        "grace_period_slots": c.u32(),
        "delta": c.u32(),
        "genesis_state_timestamp": c.u64(),
    }
    st["body"] = body
    st["_end"] = c.o
    return st


def _evals_pair(c: Cur):
    a = [c.bigint() for _ in range(c.u64())]
    b = [c.bigint() for _ in range(c.u64())]
    return (a, b)


EVAL_NAMES = (
    ["w%d" % i for i in range(15)]
    + ["coefficients%d" % i for i in range(15)]
    + ["z"]
    + ["s%d" % i for i in range(6)]
    + [
        "generic_selector",
        "poseidon_selector",
        "complete_add_selector",
        "mul_selector",
        "emul_selector",
        "endomul_scalar_selector",
    ]
)


def pickles_proof(c: Cur):
    """MinaBaseProofStableV2 = PicklesProofProofsVerified2ReprStableV2."""
    pr = {}
    plonk = {}
    plonk["alpha"] = c.limbs2()
    plonk["beta"] = c.limbs2()
    plonk["gamma"] = c.limbs2()
    plonk["zeta"] = c.limbs2()
    if c.option_tag():
        plonk["joint_combiner"] = c.limbs2()
    else:
        plonk["joint_combiner"] = None
    plonk["feature_flags"] = [c.boolean() for _ in range(8)]
    pr["plonk"] = plonk
    pr["bulletproof_challenges"] = [c.limbs2() for _ in range(16)]
    pv = c.u32()
    if pv > 2:
        raise DecodeError("proofs_verified tag")
    pr["proofs_verified"] = pv
    pr["domain_log2"] = c.u8()
    pr["sponge_digest_before_evaluations"] = [c.i64() & 0xFFFFFFFFFFFFFFFF for _ in range(4)]
    pr["wrap_challenge_polynomial_commitment"] = c.point()
    pr["wrap_old_bulletproof_challenges"] = [[c.limbs2() for _ in range(15)] for _ in range(2)]
    n = c.u64()
    pr["step_challenge_polynomial_commitments"] = [c.point() for _ in range(n)]
    n = c.u64()
    pr["step_old_bulletproof_challenges"] = [[c.limbs2() for _ in range(16)] for _ in range(n)]
    pe = {}
    pe["public_input"] = (c.bigint(), c.bigint())
    pe["evals"] = {name: _evals_pair(c) for name in EVAL_NAMES}
    for _ in range(19):
        if c.option_tag():
            raise DecodeError("optional evaluation present: unsupported by the blockchain circuit")
    pe["ft_eval1"] = c.bigint()
    pr["prev_evals"] = pe
    wp = {}
    wp["w_comm"] = [c.point() for _ in range(15)]
    wp["z_comm"] = c.point()
    wp["t_comm"] = [c.point() for _ in range(7)]
    wp["evals"] = {name: (c.bigint(), c.bigint()) for name in EVAL_NAMES}
    wp["ft_eval1"] = c.bigint()
    n = c.u64()
    wp["lr"] = [(c.point(), c.point()) for _ in range(n)]
    wp["z_1"] = c.bigint()
    wp["z_2"] = c.bigint()
    wp["delta"] = c.point()
    wp["sg"] = c.point()
    pr["proof"] = wp
    return pr


def decode_state_proof(data: bytes):
    c = Cur(data)
    out = {"candidate_tip_proof": pickles_proof(c)}
    out["_proof_end"] = c.o
    out["candidate_chain_states"] = [protocol_state(c) for _ in range(FRONTIER)]
    out["bridge_tip_state"] = protocol_state(c)
    out["_consumed"] = c.o
    return out


def decode_state_pub(data: bytes):
    c = Cur(data)
    out = {"is_state_proof_from_devnet": c.boolean()}
    out["bridge_tip_state_hash"] = int.from_bytes(c.take(32), "little")
    out["candidate_chain_state_hashes"] = [int.from_bytes(c.take(32), "little") for _ in range(FRONTIER)]
    out["candidate_chain_ledger_hashes"] = [int.from_bytes(c.take(32), "little") for _ in range(FRONTIER)]
    out["_consumed"] = c.o
    return out


def decode_account_proof(data: bytes):
    c = Cur(data)
    n = c.u64()
    path = []
    for _ in range(n):
        tag = c.u32()
        if tag > 1:
            raise DecodeError("merkle node tag")
        if c.u64() != 32:
            raise DecodeError("field length")
        path.append((tag, int.from_bytes(c.take(32), "little")))
    out = {"merkle_path": path, "_account_start": c.o}
    acc = {}
    acc["public_key"] = _pubkey(c)
    acc["token_id"] = c.bigint()
    acc["token_symbol"] = c.bytestr()
    acc["balance"] = c.u64()
    acc["nonce"] = c.u32()
    acc["receipt_chain_hash"] = c.bigint()
    acc["delegate"] = _pubkey(c) if c.option_tag() else None
    acc["voting_for"] = c.bigint()
    ttag = c.u32()
    if ttag == 0:
        acc["timing"] = None
    elif ttag == 1:
        acc["timing"] = {
            "initial_minimum_balance": c.u64(),
            "cliff_time": c.u32(),
            "cliff_amount": c.u64(),
            "vesting_period": c.u32(),
            "vesting_increment": c.u64(),
        }
    else:
        raise DecodeError("timing tag")
    perms = {}
    for name in (
        "edit_state",
        "access",
        "send",
        "receive",
        "set_delegate",
        "set_permissions",
        "set_verification_key",
        "set_zkapp_uri",
        "edit_action_state",
        "set_token_symbol",
        "increment_nonce",
        "set_voting_for",
        "set_timing",
    ):
        tag = c.u32()
        if tag > 4:
            raise DecodeError("auth tag")
        if name == "set_verification_key":
            perms[name] = (tag, c.u32())
        else:
            perms[name] = tag
    acc["permissions"] = perms
    if c.option_tag():
        raise DecodeError("zkapp accounts: not decoded by the oracle yet")
    acc["zkapp"] = None
    out["account"] = acc
    out["_consumed"] = c.o
    return out


def decode_account_pub(data: bytes):
    c = Cur(data)
    out = {"ledger_hash": int.from_bytes(c.take(32), "little")}
    out["encoded_account"] = c.bytestr()
    out["_consumed"] = c.o
    return out


# ---- producer side: re-encode a decoded protocol state (bincode), used to build mutated test vectors the
# way the reference's consensus tests mutate the fixture (consensus_state.rs:197-303) -------------------------
def _w_bigint(x: int) -> bytes:
    return struct.pack("<Q", 32) + int(x).to_bytes(32, "little")


def _w_bytes(b: bytes) -> bytes:
    return struct.pack("<Q", len(b)) + bytes(b)


def _w_signed(a) -> bytes:
    return struct.pack("<QI", a[0], a[1])


def _w_registers(r) -> bytes:
    ls = r["local_state"]
    out = _w_bigint(r["first_pass_ledger"]) + _w_bigint(r["second_pass_ledger"])
    pcs = r["pending_coinbase_stack"]
    out += _w_bigint(pcs["data"]) + _w_bigint(pcs["state_init"]) + _w_bigint(pcs["state_curr"])
    out += _w_bigint(ls["stack_frame"]) + _w_bigint(ls["call_stack"]) + _w_bigint(ls["transaction_commitment"])
    out += _w_bigint(ls["full_transaction_commitment"]) + _w_signed(ls["excess"]) + _w_signed(ls["supply_increase"])
    out += _w_bigint(ls["ledger"]) + struct.pack("<BI", int(ls["success"]), ls["account_update_index"])
    out += struct.pack("<Q", len(ls["failure_status_tbl"]))
    for row in ls["failure_status_tbl"]:
        out += struct.pack("<Q", len(row)) + b"".join(struct.pack("<I", t) for t in row)
    return out + struct.pack("<B", int(ls["will_succeed"]))


def _w_epoch(e) -> bytes:
    return (_w_bigint(e["ledger_hash"]) + struct.pack("<Q", e["ledger_total_currency"]) + _w_bigint(e["seed"])
            + _w_bigint(e["start_checkpoint"]) + _w_bigint(e["lock_checkpoint"]) + struct.pack("<I", e["epoch_length"]))


def _w_pubkey(k) -> bytes:
    return _w_bigint(k[0]) + struct.pack("<B", int(k[1]))


def encode_protocol_state(st) -> bytes:
    body = st["body"]
    bs, cs, k = body["blockchain_state"], body["consensus_state"], body["constants"]
    slh, lps = bs["staged_ledger_hash"], bs["ledger_proof_statement"]
    out = _w_bigint(st["previous_state_hash"]) + _w_bigint(body["genesis_state_hash"])
    out += _w_bigint(slh["ledger_hash"]) + _w_bytes(slh["aux_hash"]) + _w_bytes(slh["pending_coinbase_aux"]) + _w_bigint(slh["pending_coinbase_hash"])
    out += _w_bigint(bs["genesis_ledger_hash"]) + _w_registers(lps["source"]) + _w_registers(lps["target"])
    out += _w_bigint(lps["connecting_ledger_left"]) + _w_bigint(lps["connecting_ledger_right"]) + _w_signed(lps["supply_increase"])
    for tok, amt in lps["fee_excess"]:
        out += _w_bigint(tok) + _w_signed(amt)
    out += struct.pack("<Q", bs["timestamp"]) + _w_bytes(bs["body_reference"])
    out += struct.pack("<III", cs["blockchain_length"], cs["epoch_count"], cs["min_window_density"])
    out += struct.pack("<Q", len(cs["sub_window_densities"])) + b"".join(struct.pack("<I", d) for d in cs["sub_window_densities"])
    out += _w_bytes(cs["last_vrf_output"]) + struct.pack("<Q", cs["total_currency"])
    out += struct.pack("<III", 0, cs["curr_global_slot"]["slot_number"], cs["curr_global_slot"]["slots_per_epoch"])
    out += struct.pack("<II", 0, cs["global_slot_since_genesis"])
    out += _w_epoch(cs["staking_epoch_data"]) + _w_epoch(cs["next_epoch_data"])
    out += struct.pack("<B", int(cs["has_ancestor_in_same_checkpoint_window"]))
    out += _w_pubkey(cs["block_stake_winner"]) + _w_pubkey(cs["block_creator"]) + _w_pubkey(cs["coinbase_receiver"])
    out += struct.pack("<B", int(cs["supercharge_coinbase"]))
    out += struct.pack("<IIIIIQ", k["k"], k["slots_per_epoch"], k["slots_per_sub_window"], k["grace_period_slots"], k["delta"], k["genesis_state_timestamp"])
    return out
