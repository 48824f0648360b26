"""CPU oracle (TEST INFRASTRUCTURE ONLY) -- Poseidon sponge with kimchi's shape, table-driven.

Restates mina-poseidon `ArithmeticSponge<Fp, PlonkSpongeConstantsKimchi>` and mina-p2p-messages
`hash_with_kimchi` (un-vendored; lambdaclass/openmina-proof-systems @ 44e0d3b, lambdaclass/openmina @ 711c99f;
SURVEY B.9), as called from AL/operator/mina_account/lib/src/merkle_verifier.rs:27.

PARITY UNPINNED: the round constants / MDS of the kimchi parameter set are absent from /root/reference and
from this image.  Every function takes the table as an argument (9 MDS entries row-major + 55 x 3 round
constants as Python ints); the only known-answer vector is merkle_verifier.rs:43-58 (`KAT_ROOT` below),
which a candidate table must reproduce before anything downstream can be called pinned.
Only tests/ may import this module.
"""
from __future__ import annotations

import random

WIDTH, RATE, ROUNDS = 3, 2, 55
KAT_ROOT = int.from_bytes(bytes([140, 130, 39, 24, 215, 108, 36, 34, 181, 80, 10, 131, 110, 152, 243, 145, 144, 175, 100,
                                 161, 62, 28, 236, 143, 184, 143, 185, 114, 129, 4, 63, 47]), "little")


def random_table(mod: int, seed: int):
    rng = random.Random(seed)
    return [rng.randrange(mod) for _ in range(9 + 3 * ROUNDS)]


def table_bytes(table) -> bytes:
    return b"".join(int(x).to_bytes(32, "little") for x in table)


def permute(table, st, mod):
    mds, rc = table[:9], table[9:]
    st = list(st)
    for r in range(ROUNDS):
        sb = [pow(x, 7, mod) for x in st]
        st = [(mds[3 * i] * sb[0] + mds[3 * i + 1] * sb[1] + mds[3 * i + 2] * sb[2] + rc[3 * r + i]) % mod for i in range(3)]
    return st


class Sponge:
    def __init__(self, table, mod):
        self.t, self.m = table, mod
        self.state = [0, 0, 0]
        self.absorbing, self.count = True, 0

    def absorb(self, x):
        if self.absorbing:
            if self.count == RATE:
                self.state = permute(self.t, self.state, self.m)
                self.state[0] = (self.state[0] + x) % self.m
                self.count = 1
            else:
                self.state[self.count] = (self.state[self.count] + x) % self.m
                self.count += 1
        else:
            self.state[0] = (self.state[0] + x) % self.m
            self.absorbing, self.count = True, 1

    def squeeze(self):
        if self.absorbing:
            self.state = permute(self.t, self.state, self.m)
            self.absorbing, self.count = False, 1
            return self.state[0]
        if self.count == RATE:
            self.state = permute(self.t, self.state, self.m)
            self.count = 1
            return self.state[0]
        self.count += 1
        return self.state[self.count - 1]


def prefix_to_field(prefix: str) -> int:
    b = prefix.encode()
    assert len(b) <= 20
    return int.from_bytes(b + b"*" * (20 - len(b)) + b"\0" * 12, "little")


def hash_with_kimchi(table, prefix: str, xs, mod):
    s = Sponge(table, mod)
    s.absorb(prefix_to_field(prefix))
    s.squeeze()
    for x in xs:
        s.absorb(x)
    return s.squeeze()


def merkle_root(table, leaf, path, mod):
    """path: [(tag, sibling)], tag 0 = Left(sibling) -> [acc, sibling], 1 = Right(sibling) -> [sibling, acc]."""
    acc = leaf
    for depth, (tag, sib) in enumerate(path):
        acc = hash_with_kimchi(table, "MinaMklTree%03d" % depth, [sib, acc] if tag else [acc, sib], mod)
    return acc
