#!/usr/bin/env python3
"""Regenerate tests/golden/srs_sha256.json from the reference's committed SRS files.

Runs only where /root/reference exists (the build container).  It decompresses
srs/{vesta,pallas}.srs with the oracle (ark-serialize 0.3 compressed points),
checks that the result equals the oracle's hash-to-curve derivation (SRS::create), and records
SHA-256 digests of the canonical affine arrays so that every other box can pin its derived SRS.
"""
import hashlib
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import cref, pasta  # noqa: E402

REF = "/root/reference/srs"


def main():
    out = {}
    for name, fid in (("vesta", cref.FQ), ("pallas", cref.FP)):
        g, h = pasta.srs_compressed_bytes(os.path.join(REF, name + ".srs"))
        aff = cref.decompress(fid, b"".join(g))
        haff = cref.decompress(fid, h)
        der, hder = cref.srs_derive(fid, 0, len(g), True)
        assert der == aff and hder == haff, "derivation != committed file for " + name
        out[name] = {
            "depth_in_file": len(g),
            "sha256_g_65536": hashlib.sha256(aff).hexdigest(),
            "sha256_g_32768": hashlib.sha256(aff[: 64 * 32768]).hexdigest(),
            "sha256_g_1024": hashlib.sha256(aff[: 64 * 1024]).hexdigest(),
            "h": haff.hex(),
            "file_sha256": hashlib.sha256(open(os.path.join(REF, name + ".srs"), "rb").read()).hexdigest(),
        }
    path = os.path.join(ROOT, "tests", "golden", "srs_sha256.json")
    json.dump(out, open(path, "w"), indent=1, sort_keys=True)
    print("wrote", path)


if __name__ == "__main__":
    main()
