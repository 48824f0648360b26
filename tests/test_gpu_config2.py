"""GPU tier: BASELINE config 2 -- a 2^20-point Vesta MSM, bit-exact against the oracle.

Bases: vesta.srs g[0..65536) followed by g[i] = to_group(blake2b512(be32(i))) for i < 2^20 -- the SRS's own
derivation rule (K-D), so the first 2^16 are the committed SRS.  Scalars: SplitMix64 seed 0x4D494E41 (4 draws
per scalar, reduced mod p) plus the edge batches all-zero / all-one / all-(p-1) / the 2^16 K-A scalars tiled."""
import pytest

from oracle import cref, pasta
from conftest import golden
from oracle import wire

pytestmark = pytest.mark.gpu

N = 1 << 20


@pytest.fixture(scope="module")
def bases(gpu):
    pts = gpu.srs_points(1, 0, 65536) + gpu.host_srs_derive(1, 65536, N - 65536)
    assert len(pts) == 64 * N
    # the extension really is the same rule: re-deriving a slice of the committed part reproduces it
    assert gpu.host_srs_derive(1, 65000, 536) == pts[64 * 65000 : 64 * 65536]
    return pts


@pytest.fixture(scope="module")
def batches():
    import bench

    sm = bench.splitmix_scalars(N, 0x4D494E41)
    first = int.from_bytes(sm[:32], "little")
    # SplitMix64 reference values for seed 0x4D494E41: the first scalar is built from the first four outputs
    z, outs = 0x4D494E41, []
    for _ in range(4):
        z = (z + 0x9E3779B97F4A7C15) & (2**64 - 1)
        x = z
        x = ((x ^ (x >> 30)) * 0xBF58476D1CE4E5B9) & (2**64 - 1)
        x = ((x ^ (x >> 27)) * 0x94D049BB133111EB) & (2**64 - 1)
        outs.append(x ^ (x >> 31))
    assert first == (outs[0] | outs[1] << 64 | outs[2] << 128 | outs[3] << 192) % pasta.P
    pr = wire.decode_state_proof(golden("mina_state.proof"))["candidate_tip_proof"]
    pre = b"".join(x.to_bytes(16, "little") for x in pr["bulletproof_challenges"])
    kat = cref.bpoly_coeffs(cref.FP, cref.endo_to_field(cref.FP, pre, pasta.ENDO_FP))
    return {
        "splitmix": sm,
        "zeros": b"\0" * (32 * N),
        "ones": (1).to_bytes(32, "little") * N,
        "minus_one": (pasta.P - 1).to_bytes(32, "little") * N,
        "kat_tiled": kat * 16,
    }


def test_msm20_generic_path_matches_oracle(gpu, bases, batches):
    for name in ("splitmix", "zeros"):
        want, inf = cref.msm(cref.FQ, batches[name], bases, 16)
        got = gpu.msm(1, batches[name], bases)
        assert got == want, name
        assert inf == (name == "zeros")


def test_msm20_resident_table_matches_oracle_on_every_edge_batch(gpu, bases, batches):
    gpu.fixed_base_load(1, bases, 20)
    names = list(batches)
    got = gpu.fixed_base_msm(1, b"".join(batches[n] for n in names), N)
    for name, g in zip(names, got):
        want, _ = cref.msm(cref.FQ, batches[name], bases, 16)
        assert g == want, name
    # size-independent properties at full size: sum of all bases, and MSM(-1) = -MSM(1)
    ones, minus = cref.bytes_to_point(got[names.index("ones")]), cref.bytes_to_point(got[names.index("minus_one")])
    assert minus == (ones[0], pasta.Q - ones[1])
    # adversarial population: every scalar identical -> one bucket per window holds all 2^20 points
    same = (0x0001000100010001000100010001000100010001000100010001000100010001 % pasta.P).to_bytes(32, "little") * N
    want, _ = cref.msm(cref.FQ, same, bases, 16)
    assert gpu.fixed_base_msm(1, same, N) == [want]


def test_fixed_base_load_rejects_bad_points(gpu, bases):
    bad = bytearray(bases[: 64 * 8])
    bad[5] ^= 1
    with pytest.raises(gpu.MinaB200Error, match="not on the curve"):
        gpu.fixed_base_load(1, bytes(bad))
    with pytest.raises(gpu.MinaB200Error, match="not on the curve"):
        gpu.msm(1, b"\1" + b"\0" * 31 + b"\0" * (32 * 7), bytes(bad))
    noncanonical = bytearray(bases[: 64 * 2])
    noncanonical[0:32] = (pasta.Q + 5).to_bytes(32, "little")
    with pytest.raises(gpu.MinaB200Error):
        gpu.fixed_base_load(1, bytes(noncanonical))
    # scalars must be canonical: >= 2^255 is flagged by the engine instead of silently truncated
    gpu.fixed_base_load(1, bases[: 64 * 16], 16)
    with pytest.raises(gpu.MinaB200Error, match="2\\^255"):
        gpu.fixed_base_msm(1, (b"\xff" * 32) * 16, 16)


def test_adversarial_equal_scalars_on_the_srs_engine(gpu):
    """All-equal digits put n*W points into ONE bucket: the overflow kernel must keep it exact (and bounded)."""
    for cid, n, m in ((1, 65536, pasta.P), (0, 32768, pasta.Q)):
        pts = gpu.srs_points(cid, 0, n)
        for s in (1, 0x0001000100010001000100010001000100010001000100010001000100010001 % m, m - 1):
            sc = s.to_bytes(32, "little") * n
            want, _ = cref.msm(cid, sc, pts, 16)
            assert gpu.msm_srs(cid, sc, n) == [want]
