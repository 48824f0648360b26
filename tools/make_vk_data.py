#!/usr/bin/env python3
"""Extract the parts of the reference's blockchain verification keys that its loader consumes.

Source: AL/operator/mina/lib/src/{devnet,mainnet}_vk.json (include_str!'d at
AL/operator/mina/lib/src/verifier_index.rs:20-21).  The reference's serde structs (:28-69) read only
`commitments` and `index.{domain.log_size_of_group, max_poly_size, public, prev_challenges, shifts}`;
this script keeps exactly those keys (plus `index.domain.group_gen`, used as a cross-check of the
derived domain generator, KAT K-G) and drops the ~90 % of each file nothing reads.  Output:
mina_bridge_b200/data/{devnet,mainnet}_vk.json -- verification-key DATA the verifier needs, in the
reference's JSON grammar so the full original files load too.
Run here only (needs /root/reference).
"""
import json
import os

REF = "/root/reference/contract/lib/aligned_layer/operator/mina/lib/src"
OUT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "mina_bridge_b200", "data")

for name in ("devnet", "mainnet"):
    d = json.load(open(os.path.join(REF, name + "_vk.json")))
    ix = d["index"]
    keep = {
        "commitments": d["commitments"],
        "index": {
            "domain": {"log_size_of_group": ix["domain"]["log_size_of_group"], "group_gen": ix["domain"]["group_gen"]},
            "max_poly_size": ix["max_poly_size"],
            "public": ix["public"],
            "prev_challenges": ix["prev_challenges"],
            "shifts": ix["shifts"],
        },
    }
    path = os.path.join(OUT, name + "_vk.json")
    with open(path, "w") as f:
        json.dump(keep, f, indent=1)
        f.write("\n")
    print("wrote", os.path.normpath(path), os.path.getsize(path), "bytes")
