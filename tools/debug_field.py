import sys, os, random
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import mina_bridge_b200 as mb
from oracle import pasta, cref
mb.init(0)
rng = random.Random(5)
for fid, m in ((1, pasta.Q),):
    Rinv = pow(1 << 256, -1, m)
    cases = [(0, 0), (0, 1), (1, 0), (1, 1), (2, 3), (m - 1, 1), (m - 1, m - 1), (1 << 32, 1), (1 << 64, 1 << 64), (1 << 200, 1 << 50), (1 << 224, 1 << 224), ((1<<32)-1, (1<<32)-1)]
    cases += [(rng.randrange(m), rng.randrange(m)) for _ in range(2000)]
    A = cref.ints_to_bytes([a for a, b in cases]); B = cref.ints_to_bytes([b for a, b in cases])
    M = 1 << 256
    def redc_ref(lo, hi):
        U = lo | (hi << 256)
        mm = (-U * pow(m, -1, M)) % M
        r = (U + mm * m) >> 256
        return r - m if r >= m else r
    for op, name, fn in ((20, "prod_lo", lambda a, b: (a * b) % M), (21, "prod_hi", lambda a, b: (a * b) >> 256), (22, "redc", lambda a, b: redc_ref(a, b % (m >> 1))), (10, "mul_ptx", lambda a, b: a * b * Rinv % m)):
        BB = B if op != 22 else cref.ints_to_bytes([b % (m >> 1) for a, b in cases])
        out = mb.field_op(fid, op, A, BB)
        bad = 0
        for i, (a, b) in enumerate(cases):
            got = int.from_bytes(out[32 * i:32 * i + 32], "little")
            if got != fn(a, b):
                bad += 1
                if bad <= 4:
                    print("field", fid, name, "MISMATCH a=%x b=%x\n   got=%064x\n  want=%064x" % (a, b, got, fn(a, b)))
        print("field", fid, name, "bad:", bad, "of", len(cases))
