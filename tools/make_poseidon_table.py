#!/usr/bin/env python3
"""Turn a source file that holds the kimchi Poseidon parameters into the table this library loads, and gate it on the
reference's known-answer test.

The constants are NOT in the reference tree nor in this image: they live in mina-poseidon's
`pasta/fp_kimchi.rs` / `pasta/fq_kimchi.rs` (lambdaclass/openmina-proof-systems @ 44e0d3b) and, identically, in o1js'
`bindings/crypto/constants.ts`.  Given such a file this tool extracts, in order of appearance, the 9 MDS entries
(row-major) and the 55 x 3 round constants -- decimal or 0x-hex literals of >= 60 digits, in quotes or with a trailing
`n` -- writes `mina_bridge_b200/data/poseidon_{fp,fq}_kimchi.bin` (174 x 32 bytes little-endian) and, for Fp, runs the
reference's KAT (AL/operator/mina_account/lib/src/merkle_verifier.rs:43-58) through the library's own host sponge.
A table that fails the KAT is not written.  Once the Fp file exists, `mina_b200_init` loads it, `mina_b200_poseidon_trusted()`
turns 1 and `tests/test_boundary_cpu.py::test_poseidon_reference_kat` stops skipping.

usage: python tools/make_poseidon_table.py fp path/to/fp_kimchi.rs        (or: fq path/to/fq_kimchi.rs)
"""
import os
import re
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

MOD = {"fp": 0x40000000000000000000000000000000224698FC094CF91B992D30ED00000001,
       "fq": 0x40000000000000000000000000000000224698FC0994A8DD8C46EB2100000001}
KAT_ROOT = bytes([140, 130, 39, 24, 215, 108, 36, 34, 181, 80, 10, 131, 110, 152, 243, 145, 144, 175, 100, 161, 62, 28, 236, 143,
                  184, 143, 185, 114, 129, 4, 63, 47])


def literals(text):
    out = []
    for m in re.finditer(r"0x[0-9a-fA-F]{60,64}|\b[0-9]{60,78}\b", text):
        s = m.group(0)
        out.append(int(s, 16) if s.startswith("0x") else int(s))
    return out


def main():
    if len(sys.argv) != 3 or sys.argv[1] not in MOD:
        raise SystemExit(__doc__)
    field, path = sys.argv[1], sys.argv[2]
    nums = literals(open(path).read())
    if len(nums) < 174:
        raise SystemExit("found only %d field-element literals, need 9 + 165" % len(nums))
    nums = nums[:174]
    if any(x >= MOD[field] for x in nums):
        raise SystemExit("a literal is not a canonical %s element: wrong field?" % field)
    table = b"".join(x.to_bytes(32, "little") for x in nums)
    if field == "fp":
        import mina_bridge_b200 as mb

        acc = 0
        for depth in range(2):  # leaf 0, path [Left(0), Right(0)]
            xs = [acc, 0] if depth == 0 else [0, acc]
            acc = mb.host_hash_with_kimchi(table, "MinaMklTree%03d" % depth, xs)
        if acc.to_bytes(32, "little") != KAT_ROOT:
            raise SystemExit("the table does NOT reproduce the reference's Merkle KAT: not written")
        print("KAT passed")
    out = os.path.join(ROOT, "mina_bridge_b200", "data", "poseidon_%s_kimchi.bin" % field)
    open(out, "wb").write(table)
    print("wrote", out)


if __name__ == "__main__":
    main()
