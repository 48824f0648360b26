#!/usr/bin/env python3
"""bench.py -- one JSON line per run (driver contract).

Default workload `state1024` (BASELINE.json: "1024-proof Mina-state batch"): one STEP verifies a fixed
batch of 1024 serialized proof-of-state inputs (the reference's own fixture replicated, 1 % of them
with one flipped bit in an IPA prechallenge, at seeded random positions) through every stage this build
has -- bincode decode, the plain-comparison half of check_pub_inputs, the fork-choice rule,
accumulator_check (Vesta 2^16) and the wrap proof's two previous-challenge accumulators (Pallas 2^15).
The batch is strong-scaled: rank r takes proofs i = r (mod N); the per-proof result bytes are combined
with ONE NCCL all-reduce(MIN) inside the timed region (AL/operator/pkg/operator.go:461-465).

READ `config.absent_stages`: the kimchi transcript and the 17 Poseidon state hashes are NOT built
(Poseidon constants unavailable => parity unpinned), so the full accept bit is
always 0 and the byte that is reduced is "every built stage passed".  The metric is named for what it is.

  value  = proofs/s with the per-proof device inputs (prechallenges + accumulator points) resident in HBM
  e2e    = proofs/s through the C ABI with HOST buffers holding the serialized proofs (decode on host
           threads, pinned staging, H2D, kernels, D2H of the result bytes all inside the timed region)
Other workloads (same JSON schema; none is the driver's default):
  --workload msm20   BASELINE configs[1]: one 2^20-point Vesta MSM over resident bases
  --batch 64         BASELINE configs[2]'s batch size through the default workload
  --workload mixed   BASELINE configs[4]: 512 proof-of-state + 512 proof-of-account inputs
  --workload ipa     row a9 alone: 1024 IPA final checks of the wrap proof's shape (oracle-made fixture)
  --workload merkle  K3 alone: 1024 Merkle paths of 35 levels
  --profile-step rlc wraps ONE extra untimed step in cudaProfilerStart/Stop (ncu --profile-from-start off)
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

BATCH = 1024              # proofs per step, whole job (BASELINE.json configs[3] / north_star target); --batch overrides
N_CORRUPT = 10            # 1 % of the batch gets one flipped prechallenge bit, at seeded random positions; --corrupt overrides
CORRUPT_SEED = 0x4D494E41
ALG_BYTES_PER_POINT = 96  # 64 B affine base + 32 B scalar
OFF_WRAP_PRE, OFF_WRAP_X, OFF_WRAP_Y = 73, 374, 414
OFF_STEP_PRE, OFF_STEP_SG = 446, 934
ABSENT = ["kimchi to_batch + IPA final check (Fiat-Shamir needs Poseidon)", "17 Poseidon protocol-state hashes",
          "full accept bit (always 0 until the two stages above exist)"]
BUILT = ["bincode decode", "check_pub_inputs: ledger-hash + to_fp comparisons", "select_secure_chain",
         "accumulator_check (Vesta MSM 2^16, K-A pinned)", "wrap prev-challenge accumulators (2 x Pallas MSM 2^15, K-B/K-C pinned)"]


def peaks():
    try:
        return json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"], "measured"
    except Exception:
        return 6650.0, "fallback"


class ClockSampler(threading.Thread):
    def __init__(self):
        super().__init__(daemon=True)
        self.rows, self.stop = [], threading.Event()

    def run(self):
        q = "clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap"
        while not self.stop.is_set():
            try:
                out = subprocess.run(["nvidia-smi", "-i", os.environ.get("LOCAL_RANK", "0"), "--query-gpu=" + q,
                                      "--format=csv,noheader,nounits"], capture_output=True, text=True, timeout=5).stdout
                self.rows.append([x.strip() for x in out.strip().split(",")])
            except Exception:
                pass
            self.stop.wait(0.2)

    def summary(self):
        sm = sorted(int(r[0]) for r in self.rows if r and r[0].isdigit())
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = sorted({names[i] for r in self.rows if len(r) >= 6 for i in range(4) if r[2 + i] == "Active"})
        mx = max((int(r[1]) for r in self.rows if len(r) > 1 and r[1].isdigit()), default=None)
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": mx, "reasons": reasons}


def ncu_capture(kernel):
    """DRAM traffic and pipe counters of `kernel` from the committed `ncu --set full` capture of this same command's
    profile step (profiles/r2_rlc_step_kernels.json, made by tools/ncu_raw_summary.py).  None if absent."""
    try:
        d = json.load(open(os.path.join(ROOT, "profiles", "r2_rlc_step_kernels.json")))["kernels"]
    except Exception:
        return None
    hit = [v for k, v in d.items() if k.startswith(kernel + "<") and v["launches"] >= 2] or [v for k, v in d.items() if k.startswith(kernel + "<")]
    return max(hit, key=lambda v: v["ms"]) if hit else None


MODMUL_CEILING = 7.31e10  # dependent-chain Montgomery products per second, all SMs, measured (tools/lab/accum_lab.cu)


def golden(name):
    return open(os.path.join(ROOT, "tests", "golden", name), "rb").read()


def corrupt_positions(n_corrupt=None):
    import random

    return set(random.Random(CORRUPT_SEED).sample(range(BATCH), N_CORRUPT if n_corrupt is None else n_corrupt))


def synth_batch(n_corrupt=None):
    """The fixed 1024-proof batch: (proofs, pubs, expected built-stage bit)."""
    proof, pub = golden("mina_state.proof"), golden("mina_state.pub")
    bad = corrupt_positions(n_corrupt)
    proofs, want = [], []
    for i in range(BATCH):
        if i in bad:
            m = bytearray(proof)
            m[OFF_WRAP_PRE + (i % 256)] ^= 1 << (i % 8)
            proofs.append(bytes(m))
            want.append(0)
        else:
            proofs.append(proof)
            want.append(1)
    return proofs, [pub] * BATCH, want


def device_inputs(proofs, torch, dev):
    """What the host pass extracts from each proof, packed as the device arrays the kernels read."""
    import numpy as np

    pre_w = b"".join(p[OFF_WRAP_PRE:OFF_WRAP_PRE + 256] for p in proofs)
    pts_w = b"".join(p[OFF_WRAP_X:OFF_WRAP_X + 32] + p[OFF_WRAP_Y:OFF_WRAP_Y + 32] for p in proofs)
    pre_s = b"".join(p[OFF_STEP_PRE:OFF_STEP_PRE + 480] for p in proofs)
    pts_s = b"".join(p[OFF_STEP_SG + 80 * k + 8:OFF_STEP_SG + 80 * k + 40] + p[OFF_STEP_SG + 80 * k + 48:OFF_STEP_SG + 80 * k + 80]
                     for p in proofs for k in range(2))
    to_dev = lambda b: torch.from_numpy(np.frombuffer(b, dtype=np.uint8).copy()).to(dev)
    return to_dev(pre_w), to_dev(pts_w), to_dev(pre_s), to_dev(pts_s)


# ---- CPU side: the oracle port doing the same per-proof work the reference does (one MSM per accumulator) ----
def cpu_verify_sample(proofs, pubs, threads):
    from oracle import cref, pasta, wire

    cref.build()
    bases_v = cpu_verify_sample.cache.get("v")
    if bases_v is None:
        bases_v, _ = cref.srs_derive(cref.FQ, 0, 65536, False)
        bases_p, _ = cref.srs_derive(cref.FP, 0, 32768, False)
        cpu_verify_sample.cache.update(v=bases_v, p=bases_p)
    bases_p = cpu_verify_sample.cache["p"]
    t0 = time.perf_counter()
    bits = []
    for pb, qb in zip(proofs, pubs):
        sp, sq = wire.decode_state_proof(pb), wire.decode_state_pub(qb)
        ok = all(sq["candidate_chain_ledger_hashes"][i] ==
                 sp["candidate_chain_states"][i]["body"]["blockchain_state"]["ledger_proof_statement"]["target"]["first_pass_ledger"]
                 for i in range(16))
        pr = sp["candidate_tip_proof"]
        pre = b"".join(x.to_bytes(16, "little") for x in pr["bulletproof_challenges"])
        s = cref.bpoly_coeffs(cref.FP, cref.endo_to_field(cref.FP, pre, pasta.ENDO_FP))
        got, _ = cref.msm(cref.FQ, s, bases_v, threads)
        ok = ok and cref.bytes_to_point(got) == pr["wrap_challenge_polynomial_commitment"]
        for k in range(2):
            pre = b"".join(x.to_bytes(16, "little") for x in pr["wrap_old_bulletproof_challenges"][k])
            s = cref.bpoly_coeffs(cref.FQ, cref.endo_to_field(cref.FQ, pre, pasta.ENDO_FQ))
            got, _ = cref.msm(cref.FP, s, bases_p, threads)
            ok = ok and cref.bytes_to_point(got) == pr["step_challenge_polynomial_commitments"][k]
        bits.append(int(ok))
    return len(proofs) / (time.perf_counter() - t0), bits


cpu_verify_sample.cache = {}
METRIC = "mina_state_proofs_per_sec_built_stages"
UNIT = "proofs/s"


def run_reference(a):
    if int(os.environ.get("RANK", "0")) != 0:
        return
    cores = os.cpu_count() or 1
    proofs, pubs, want = synth_batch()
    sample = 2
    cpu_verify_sample(proofs[:1], pubs[:1], cores)  # builds the library and the SRS outside the timed steps
    for _ in range(a.warmup):
        cpu_verify_sample(proofs[:1], pubs[:1], cores)
    vals = []
    t0 = time.perf_counter()
    for s in range(a.steps):
        lo = (7 + s * sample) % (BATCH - sample)  # slides over the batch, includes corrupted members
        v, bits = cpu_verify_sample(proofs[lo:lo + sample], pubs[lo:lo + sample], cores)
        assert bits == want[lo:lo + sample]
        vals.append(v)
    dt = time.perf_counter() - t0
    v = sum(vals) / len(vals)
    from oracle import cref as _cref

    modmul_ns = _cref.modmul_ns(0, 1_000_000)
    print(json.dumps({
        "impl": "reference", "metric": METRIC, "value": v, "unit": UNIT, "n_gpus": a.gpus, "steps": a.steps, "warmup": a.warmup,
        "ms_per_step": dt / a.steps * 1e3, "higher_is_better": True, "scaling": "strong", "vs_baseline": None,
        "dtype": "u256 modular (4x64 Montgomery)", "data": "synthetic",
        "config": {"workload": "state%d: the same batch, per-proof accumulator MSMs like the reference (verify_block per proof)" % BATCH,
                   "built_stages": BUILT, "absent_stages": ABSENT},
        "cpu_baseline": {"value": v, "unit": UNIT, "cores": cores, "kind": "port", "modmul_ns": modmul_ns,
                         "sample": "%d proofs per step out of the 1024-proof batch; C restatement of arkworks' bucket MSM (oracle/pasta_ref.c, c = ln n + 2, one thread per window, unrolled 4x64 CIOS) + Python bincode decoder; NOT the reference binary (no Rust toolchain); arkworks+asm is usually quoted at 20-25 ns per modmul vs modmul_ns here" % sample},
        "e2e": {"value": v, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }))


# ---- the B200 arm -------------------------------------------------------------------------------------------------
def bench_state(a, torch, dist, mb, rank, world, dev):
    import numpy as np

    proofs, pubs, want = synth_batch()
    from mina_bridge_b200 import shard

    mine = shard.shard_indices(BATCH, rank, world)
    my_proofs, my_pubs = [proofs[i] for i in mine], [pubs[i] for i in mine]
    m = len(mine)
    d_pre_w, d_pts_w, d_pre_s, d_pts_s = device_inputs(my_proofs, torch, dev)
    idx = torch.tensor(mine, dtype=torch.int64, device=dev)
    result = torch.ones(BATCH, dtype=torch.uint8, device=dev)
    pin = torch.empty(m, dtype=torch.uint8).pin_memory()

    def reduce_bits(bits):
        """per-proof result bytes -> the batch vector every rank ends up with (operator.go:461-465)"""
        pin.copy_(torch.frombuffer(bytearray(bits), dtype=torch.uint8))
        return shard.merge_result_bytes(torch, dist, result, idx, pin.to(dev, non_blocking=True), world)

    def step_device(mode, timing=False):
        r = mb.state_accumulators_device(m, d_pre_w.data_ptr(), d_pts_w.data_ptr(), d_pre_s.data_ptr(), d_pts_s.data_ptr(), mode, timing)
        ok3, stats = r if timing else (r, None)
        bits = bytes(ok3[3 * i] & ok3[3 * i + 1] & ok3[3 * i + 2] for i in range(m))
        reduce_bits(bits)
        return stats  # (KernelStats wrap, KernelStats step) when timing

    built = 0
    for k in ("lengths", "decode_proof", "decode_pub", "pub_structure", "consensus", "accumulator", "step_accumulators"):
        built |= mb.STAGES[k]
    hp, hq = mb.Batch(my_proofs), mb.Batch(my_pubs)

    def step_e2e(mode):
        rep = mb.verify_state_stage_masks(hp, hq, mode)  # (m, 3) uint32: passed, failed, unavailable
        bits = ((rep[:, 1] == 0) & ((rep[:, 0] & built) == built)).astype(np.uint8)
        pin.copy_(torch.from_numpy(bits))
        return shard.merge_result_bytes(torch, dist, result, idx, pin.to(dev, non_blocking=True), world)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def timed(fn, steps):
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        t0 = time.perf_counter()
        e0.record()
        extra = [fn() for _ in range(steps)]
        e1.record()
        barrier()
        wall = time.perf_counter() - t0
        # the library drains its own stream before each call returns, so host wall time between the two
        # events IS the device-inclusive time; take the max of both clocks, then the max over ranks
        t = torch.tensor([max(e0.elapsed_time(e1) * 1e-3, wall)], device=dev)
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item()), extra

    W = max(a.warmup, 3)
    expect = torch.tensor(want, dtype=torch.uint8, device=dev)
    for mode in (mb.MODE_RLC, mb.MODE_PER_PROOF):  # correctness gate before any timing
        step_device(mode)
        assert torch.equal(result, expect), "device leg: result bytes differ from the expected bits"
    assert torch.equal(step_e2e(mb.MODE_RLC), expect), "e2e leg: result bytes differ from the expected bits"
    for _ in range(W):
        step_device(mb.MODE_RLC)
        step_e2e(mb.MODE_RLC)
    sampler = ClockSampler()
    sampler.start()
    l0 = mb.launch_count()
    t_dev, kms = timed(lambda: step_device(mb.MODE_RLC, True), a.steps)
    launches = mb.launch_count() - l0
    t_e2e, _ = timed(lambda: step_e2e(mb.MODE_RLC), a.steps)
    # two host threads submitting batches back to back (what the operator's goroutines do): the host pass of one batch
    # (decode, checks, packing) overlaps the device work of the other; only rank-local, no reduction inside the callers
    def two_callers():
        errs = []

        def worker():
            try:
                for _ in range(a.steps):
                    mb.verify_state_stage_masks(hp, hq, mb.MODE_RLC)
            except Exception as e:  # pragma: no cover
                errs.append(e)

        ts = [threading.Thread(target=worker) for _ in range(2)]
        barrier()
        t0 = time.perf_counter()
        [t.start() for t in ts]
        [t.join() for t in ts]
        torch.cuda.synchronize()
        dt = torch.tensor([time.perf_counter() - t0], device=dev)
        if world > 1:
            dist.all_reduce(dt, op=dist.ReduceOp.MAX)
        assert not errs, errs
        return 2 * a.steps * BATCH / float(dt.item())

    e2e_concurrent = two_callers()
    pp_steps = max(2, min(a.steps, 4))
    t_pp, kms_pp = timed(lambda: step_device(mb.MODE_PER_PROOF, True), pp_steps)
    t_pp_e2e, _ = timed(lambda: step_e2e(mb.MODE_PER_PROOF), pp_steps)
    # the same batch with no / one corrupted member (device-resident leg only): how the group testing scales down
    other = {}
    for label, nc in (("all_valid", 0), ("one_corrupted", 1)):
        o_proofs, _, o_want = synth_batch(nc)
        o_dev = device_inputs([o_proofs[i] for i in mine], torch, dev)
        o_expect = torch.tensor(o_want, dtype=torch.uint8, device=dev)

        def o_step():
            ok3 = mb.state_accumulators_device(m, *[t.data_ptr() for t in o_dev], mb.MODE_RLC, False)
            reduce_bits(bytes(ok3[3 * i] & ok3[3 * i + 1] & ok3[3 * i + 2] for i in range(m)))

        o_step()
        assert torch.equal(result, o_expect), "%s batch: result bytes differ from the expected bits" % label
        t_o, _ = timed(o_step, a.steps)
        other[label] = {"value": BATCH * a.steps / t_o, "unit": UNIT, "ms_per_step": t_o / a.steps * 1e3, "corrupted": nc}
    sampler.stop.set()
    sampler.join()
    if a.profile_step:
        # one extra step inside a cudaProfilerStart/Stop range, outside every timed region: what the committed
        # ncu passes capture (`ncu --profile-from-start off ...`, profiles/README in DESIGN.md section 4)
        barrier()
        torch.cuda.profiler.start()
        step_device(mb.MODE_PER_PROOF if a.profile_step == "per_proof" else mb.MODE_RLC)
        torch.cuda.synchronize()
        torch.cuda.profiler.stop()
    if rank != 0:
        return None
    peak, which = peaks()

    def kernel_lines(stats, steps):
        """roofline objects for k_accumulate (96 B per MSM point) and k_bpoly_combine (16 KiB of product
        tables read per proof + one 2^k x 32 B vector written per slice) from the library's own CUDA-event sums"""
        acc_ms = sum(x.accumulate_ms for pair in stats for x in pair) / steps
        comb_ms = sum(x.combine_ms for pair in stats for x in pair) / steps
        pts = sum(x.msm_points for pair in stats for x in pair) / steps
        nmsm = sum(x.msm_count for pair in stats for x in pair) / steps
        comb_bytes = sum(x.combine_proofs * 16384 + x.combine_vectors * (32 << (16 if i == 0 else 15)) for pair in stats for i, x in enumerate(pair)) / steps
        comb_mod = sum(x.combine_proofs * (1 << (16 if i == 0 else 15)) for pair in stats for i, x in enumerate(pair)) / steps
        acc = {"bound": "hbm", "achieved": ALG_BYTES_PER_POINT * pts / (acc_ms * 1e-3) / 1e9 if acc_ms else 0.0, "peak": peak, "unit": "GB/s",
               "traffic": None, "kernel": "k_accumulate", "peak_source": which, "ms_per_step_in_kernel": acc_ms, "msms_per_step": nmsm,
               "algorithmic_bytes_per_step": ALG_BYTES_PER_POINT * pts,
               "note": "96 B x MSM points / summed k_accumulate CUDA-event time; integer-issue-bound: %.3g modmul/s (10 per mixed add, 16 adds per point)" % (160 * pts / (acc_ms * 1e-3) if acc_ms else 0)}
        acc["frac"] = acc["achieved"] / peak
        cap = ncu_capture("k_accumulate")
        if cap:  # per launch, like `achieved`: the capture's DRAM bytes over its launches of this kernel
            acc["traffic"] = (cap["dram_read_bytes"] + cap["dram_write_bytes"]) / cap["launches"]
            acc["traffic_note"] = "ncu --set full of `bench.py --profile-step rlc`: %d launches, %.0f MB read + %.0f MB written; L2 hit %.0f %%; fmaheavy pipe busy %.0f %%" % (
                cap["launches"], cap["dram_read_bytes"] / 1e6, cap["dram_write_bytes"] / 1e6, cap["l2_hit_pct"], cap["fmaheavy_busy_pct"])
        mm = 160 * pts / (acc_ms * 1e-3) if acc_ms else 0.0
        acc["int_pipe"] = {"bound": "int32 multiply pipe (fmaheavy)", "achieved": mm, "peak": MODMUL_CEILING, "unit": "modmul/s", "frac": mm / MODMUL_CEILING,
                           "note": "10 Montgomery products per mixed addition x 16 additions per point; peak = measured dependent-chain ceiling of mul_ptx2"}
        comb = {"bound": "hbm", "achieved": comb_bytes / (comb_ms * 1e-3) / 1e9 if comb_ms else 0.0, "peak": peak, "unit": "GB/s", "traffic": None,
                "kernel": "k_bpoly_combine", "peak_source": which, "ms_per_step_in_kernel": comb_ms, "algorithmic_bytes_per_step": comb_bytes,
                "note": "integer-issue-bound: %.3g modmul/s" % (comb_mod / (comb_ms * 1e-3) if comb_ms else 0)}
        comb["frac"] = comb["achieved"] / peak
        cap = ncu_capture("k_bpoly_combine")
        if cap:
            comb["traffic"] = (cap["dram_read_bytes"] + cap["dram_write_bytes"]) / cap["launches"]
        return acc, comb

    acc_rlc, comb_rlc = kernel_lines(kms, a.steps)
    acc_pp, _ = kernel_lines(kms_pp, pp_steps)
    dominant = acc_rlc if acc_rlc["ms_per_step_in_kernel"] >= comb_rlc["ms_per_step_in_kernel"] else comb_rlc
    cores = os.cpu_count() or 1
    cpu, _ = cpu_verify_sample(proofs[6:8], pubs[6:8], cores)
    cpu1, _ = cpu_verify_sample(proofs[7:8], pubs[7:8], 1)
    from oracle import cref as _cref

    modmul_ns = _cref.modmul_ns(0, 1_000_000)
    return {
        "metric": METRIC, "value": BATCH * a.steps / t_dev, "unit": UNIT, "n_gpus": world, "steps": a.steps, "warmup": W,
        "ms_per_step": t_dev / a.steps * 1e3, "higher_is_better": True, "scaling": "strong", "vs_baseline": None,
        "dtype": "u256 modular (8x32 Montgomery)", "data": "synthetic",
        "config": {"workload": "state%d: serialized proof-of-state inputs, one batch per step (reference fixture x%d, %d at seeded random positions with a flipped prechallenge bit; `other_batches` = none / one corrupted), rank r takes i = r mod N, one NCCL all-reduce(MIN) of the result bytes per step" % (BATCH, BATCH, N_CORRUPT),
                   "mode": "rlc (random linear combination over the shard + bisection; per-proof numbers in `per_proof_mode`)",
                   "built_stages": BUILT, "absent_stages": ABSENT,
                   "l2": "per-step working set: %d x 16 KiB product tables + 64 MiB / 32 MiB fixed-base tables > 126 MB L2 together; inputs differ per proof only in 10 members" % (3 * m)},
        "roofline": dominant,
        "other_kernels": [comb_rlc if dominant is acc_rlc else acc_rlc],
        "other_batches": other,
        "per_proof_mode": {"value": BATCH * pp_steps / t_pp, "e2e": BATCH * pp_steps / t_pp_e2e, "unit": UNIT, "steps": pp_steps,
                           "ms_per_step": t_pp / pp_steps * 1e3, "roofline": acc_pp},
        "cpu_baseline": {"value": cpu, "unit": UNIT, "cores": cores, "kind": "port", "single_thread_value": cpu1, "modmul_ns": modmul_ns,
                         "sample": "2 proofs of the batch (all cores) + 1 proof (one thread), per-proof MSMs; oracle/pasta_ref.c (arkworks window rule c = ln n + 2, 1 thread/window, unrolled 4x64 CIOS) + Python decoder; NOT the reference binary. modmul_ns is the port's dependent-chain latency on this host; arkworks with the x86 asm backend is usually quoted at 20-25 ns, so the real reference is likely 2-3x faster than this port"},
        "e2e": {"value": BATCH * a.steps / t_e2e, "unit": UNIT, "h2d_bytes_per_step": m * (256 + 64 + 480 + 128 + 3 * 32) + m,
                "d2h_bytes_per_step": 2 * 128 + BATCH,
                "two_concurrent_callers": e2e_concurrent,
                "note": "host buffers = %d x 48 342 B serialized proofs + 1 057 B pub inputs read by host threads; only the extracted prechallenges / points / RLC scalars cross PCIe" % m},
        "gpu_launches": int(launches), "clocks": sampler.summary(),
    }


def splitmix_scalars(n, seed=0x4D494E41):
    """SplitMix64, 4 draws per scalar, reduced mod p (vectorised)."""
    import numpy as np

    P = 0x40000000000000000000000000000000224698FC094CF91B992D30ED00000001
    idx = np.arange(1, 4 * n + 1, dtype=np.uint64)
    with np.errstate(over="ignore"):
        z = np.uint64(seed) + idx * np.uint64(0x9E3779B97F4A7C15)
        z = (z ^ (z >> np.uint64(30))) * np.uint64(0xBF58476D1CE4E5B9)
        z = (z ^ (z >> np.uint64(27))) * np.uint64(0x94D049BB133111EB)
        z = z ^ (z >> np.uint64(31))
    limbs = z.reshape(n, 4)
    out = bytearray()
    for row in limbs:
        v = (int(row[0]) | int(row[1]) << 64 | int(row[2]) << 128 | int(row[3]) << 192) % P
        out += v.to_bytes(32, "little")
    return bytes(out)


def bench_msm20(a, torch, dist, mb, rank, world, dev):
    import numpy as np

    n = 1 << 20
    pts = mb.srs_points(mb.CURVE_VESTA, 0, 65536) + mb.host_srs_derive(mb.CURVE_VESTA, 65536, n - 65536)
    mb.fixed_base_load(mb.CURVE_VESTA, pts, 20)
    sc = splitmix_scalars(n, 0x4D494E41 + rank)
    d_sc = torch.from_numpy(np.frombuffer(sc, dtype=np.uint8).copy()).to(dev)
    d_out = torch.zeros(64, dtype=torch.uint8, device=dev)
    stream = torch.cuda.current_stream().cuda_stream

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    W = max(a.warmup, 3)
    for _ in range(W):
        mb.fixed_base_msm_device(mb.CURVE_VESTA, 1, d_sc.data_ptr(), d_out.data_ptr(), stream)
    l0 = mb.launch_count()
    sampler = ClockSampler()
    sampler.start()
    barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    acc_ms = 0.0
    for _ in range(a.steps):
        acc_ms += mb.fixed_base_msm_device(mb.CURVE_VESTA, 1, d_sc.data_ptr(), d_out.data_ptr(), stream, True)
    e1.record()
    barrier()
    launches = mb.launch_count() - l0
    t = torch.tensor([e0.elapsed_time(e1) * 1e-3], device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    for _ in range(2):
        mb.fixed_base_msm(mb.CURVE_VESTA, sc, n)
    barrier()
    t0 = time.perf_counter()
    for _ in range(a.steps):
        mb.fixed_base_msm(mb.CURVE_VESTA, sc, n)
    barrier()
    te = torch.tensor([time.perf_counter() - t0], device=dev)
    if world > 1:
        dist.all_reduce(te, op=dist.ReduceOp.MAX)
    sampler.stop.set()
    sampler.join()
    if rank != 0:
        return None
    from oracle import cref

    peak, which = peaks()
    t0 = time.perf_counter()
    cref.msm(cref.FQ, sc, pts, os.cpu_count() or 1)
    cpu = 1.0 / (time.perf_counter() - t0)
    achieved = ALG_BYTES_PER_POINT * n / (acc_ms / a.steps * 1e-3) / 1e9
    return {
        "metric": "vesta_msm_2e20_per_sec", "value": world * a.steps / float(t.item()), "unit": "MSM/s (n=2^20, Vesta)", "n_gpus": world,
        "steps": a.steps, "warmup": W, "ms_per_step": float(t.item()) / a.steps * 1e3, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "u256 modular (8x32 Montgomery)", "data": "synthetic",
        "config": {"workload": "msm20: BASELINE config 2, one 2^20-point Vesta MSM per step; bases = vesta.srs g[0..65536) extended by the same hash-to-curve rule to 2^20 (K-D), resident as a 13-window fixed-base table (832 MiB); scalars SplitMix64 seed 0x4D494E41",
                   "l2": "table 832 MiB + 32 MiB scalars > 126 MB L2", "window_bits": 20},
        "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak, "traffic": None,
                     "kernel": "k_accumulate", "peak_source": which, "algorithmic_bytes_per_launch": ALG_BYTES_PER_POINT * n},
        "cpu_baseline": {"value": cpu, "unit": "MSM/s (n=2^20, Vesta)", "cores": os.cpu_count(), "kind": "port",
                         "sample": "1 MSM, oracle/pasta_ref.c (arkworks window rule)"},
        "e2e": {"value": world * a.steps / float(te.item()), "unit": "MSM/s (n=2^20, Vesta)", "h2d_bytes_per_step": 32 * n, "d2h_bytes_per_step": 64},
        "gpu_launches": int(launches), "clocks": sampler.summary(),
    }


def bench_ipa(a, torch, dist, mb, rank, world, dev):
    """Row a9 alone: BATCH openings of the wrap proof's shape (Pallas, 15 rounds, 47 commitments, 2 points; the committed
    oracle-made fixture replicated, 1 % with a wrong z2) through mina_b200_ipa_verify.  The API takes host buffers, so the
    number IS end to end; the Poseidon table is arbitrary (constants unavailable), which does not change the work."""
    from mina_bridge_b200 import shard

    curve, table, op, mode, count = mb.load_ipa_fixture(os.path.join(ROOT, "tests", "golden", "ipa_pallas_k15.json"))
    bad = corrupt_positions()
    Q = 0x40000000000000000000000000000000224698FC0994A8DD8C46EB2100000001
    batch = [dict(op, z2=(op["z2"] + 1) % Q) if i in bad else op for i in range(BATCH)]
    want = [0 if i in bad else 1 for i in range(BATCH)]
    mine = shard.shard_indices(BATCH, rank, world)
    my = [batch[i] for i in mine]
    idx = torch.tensor(mine, dtype=torch.int64, device=dev)
    result = torch.ones(BATCH, dtype=torch.uint8, device=dev)
    pin = torch.empty(len(mine), dtype=torch.uint8).pin_memory()

    packed = mb.ipa_pack(my, mode, count)  # the flat host buffers a caller hands to the C ABI

    def step():
        bits = mb.ipa_verify_packed(curve, table, packed)
        pin.copy_(torch.frombuffer(bytearray(bits), dtype=torch.uint8))
        return shard.merge_result_bytes(torch, dist, result, idx, pin.to(dev, non_blocking=True), world)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    assert torch.equal(step(), torch.tensor(want, dtype=torch.uint8, device=dev)), "ipa: result bytes differ from the expected bits"
    W = max(a.warmup, 3)
    for _ in range(W):
        step()
    sampler = ClockSampler()
    sampler.start()
    l0 = mb.launch_count()
    barrier()
    t0 = time.perf_counter()
    for _ in range(a.steps):
        step()
    barrier()
    t = torch.tensor([time.perf_counter() - t0], device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    launches = mb.launch_count() - l0
    sampler.stop.set()
    sampler.join()
    if a.profile_step:  # one extra untimed step inside a cudaProfilerStart/Stop range (ncu --profile-from-start off)
        barrier()
        torch.cuda.profiler.start()
        step()
        torch.cuda.synchronize()
        torch.cuda.profiler.stop()
    if rank != 0:
        return None
    v = BATCH * a.steps / float(t.item())
    npp = 2 * 15 + 47 + 4
    return {"metric": "ipa_final_checks_per_sec", "value": v, "unit": "openings/s", "n_gpus": world, "steps": a.steps, "warmup": W,
            "ms_per_step": float(t.item()) / a.steps * 1e3, "higher_is_better": True, "scaling": "strong", "vs_baseline": None,
            "dtype": "u256 modular (8x32 Montgomery)", "data": "synthetic",
            "config": {"workload": "ipa%d: SRS::verify of %d openings per step (Pallas, 15 rounds, 47 commitments, 2 points; oracle-made fixture x%d, %d with a wrong z2), "
                                   "transcript + scalars + %d scalar multiplications per opening + batched g-side MSM" % (BATCH, BATCH, BATCH, len(bad), npp),
                       "value_is_e2e": True, "poseidon_table": "arbitrary (kimchi constants unavailable): timing is independent of the constants, parity is unpinned"},
            "roofline": None,
            "e2e": {"value": v, "unit": "openings/s", "h2d_bytes_per_step": len(mine) * (96 + 32 * 7 + 64 * (2 + 30 + 47) + 64 * npp), "d2h_bytes_per_step": len(mine) * 33},
            "gpu_launches": int(launches), "clocks": sampler.summary()}


def bench_merkle(a, torch, dist, mb, rank, world, dev):
    """K3 alone: BATCH Merkle paths of the account fixture's shape (35 levels) folded with Poseidon through
    mina_b200_merkle_fold.  The table is arbitrary (kimchi constants unavailable): the work does not depend on it."""
    import random

    P = 0x40000000000000000000000000000000224698FC094CF91B992D30ED00000001
    rng = random.Random(CORRUPT_SEED)
    table = b"".join(rng.randrange(P).to_bytes(32, "little") for _ in range(174))
    depth = 35
    n = BATCH // world
    paths = [[(rng.randrange(2), rng.randrange(P)) for _ in range(depth)] for _ in range(n)]
    leaves = [rng.randrange(P) for _ in range(n)]
    _, roots = mb.merkle_fold(table, paths, leaves, [0] * n)  # the device's own roots: a second fold must reproduce them
    bad = {i for i in corrupt_positions() if i < n}
    want_roots = [(r + 1) % P if i in bad else r for i, r in enumerate(roots)]

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    ok, _ = mb.merkle_fold(table, paths, leaves, want_roots)
    assert ok == [0 if i in bad else 1 for i in range(n)]
    W = max(a.warmup, 3)
    for _ in range(W):
        mb.merkle_fold(table, paths, leaves, want_roots)
    sampler = ClockSampler()
    sampler.start()
    l0 = mb.launch_count()
    barrier()
    t0 = time.perf_counter()
    for _ in range(a.steps):
        mb.merkle_fold(table, paths, leaves, want_roots)
    barrier()
    t = torch.tensor([time.perf_counter() - t0], device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    launches = mb.launch_count() - l0
    sampler.stop.set()
    sampler.join()
    if a.profile_step:
        barrier()
        torch.cuda.profiler.start()
        mb.merkle_fold(table, paths, leaves, want_roots)
        torch.cuda.synchronize()
        torch.cuda.profiler.stop()
    if rank != 0:
        return None
    v = n * world * a.steps / float(t.item())
    return {"metric": "merkle_paths_per_sec", "value": v, "unit": "paths/s (35 levels)", "n_gpus": world, "steps": a.steps, "warmup": W,
            "ms_per_step": float(t.item()) / a.steps * 1e3, "higher_is_better": True, "scaling": "strong", "vs_baseline": None,
            "dtype": "u256 modular (8x32 Montgomery)", "data": "synthetic",
            "config": {"workload": "merkle%d: verify_merkle_proof over %d paths of 35 levels per step (Python list marshalling included in the time)" % (BATCH, BATCH),
                       "value_is_e2e": True, "poseidon_table": "arbitrary (kimchi constants unavailable): parity unpinned, timing unaffected"},
            "roofline": None,
            "e2e": {"value": v, "unit": "paths/s (35 levels)", "h2d_bytes_per_step": n * (35 * 48 + 68), "d2h_bytes_per_step": n * 33},
            "gpu_launches": int(launches), "clocks": sampler.summary()}


def bench_mixed(a, torch, dist, mb, rank, world, dev):
    """BASELINE configs[4]: BATCH/2 proof-of-state + BATCH/2 proof-of-account inputs per step through the two batch entry
    points (host buffers in, result bytes out, one all-reduce(MIN)).  Built stages only: see config.absent_stages."""
    from mina_bridge_b200 import shard

    half = BATCH // 2
    s_proofs, s_pubs, s_want = synth_batch()
    s_proofs, s_pubs, s_want = s_proofs[:half], s_pubs[:half], s_want[:half]
    ap, aq = golden("mina_account.proof"), golden("mina_account.pub")
    bad = {i for i in corrupt_positions() if i >= half}
    a_proofs, a_want = [], []
    for i in range(half, 2 * half):
        if i in bad:  # one flipped balance byte: the ABI comparison must fail
            m = bytearray(ap)
            m[1548 + 89] ^= 1
            a_proofs.append(bytes(m))
            a_want.append(0)
        else:
            a_proofs.append(ap)
            a_want.append(1)
    want = s_want + a_want
    mine = shard.shard_indices(2 * half, rank, world)
    my_s = [i for i in mine if i < half]
    my_a = [i - half for i in mine if i >= half]
    hs, hq = mb.Batch([s_proofs[i] for i in my_s]), mb.Batch([s_pubs[i] for i in my_s])
    ha, haq = mb.Batch([a_proofs[i] for i in my_a]), mb.Batch([aq] * len(my_a))
    S = mb.STAGES
    built_s = 0
    for k in ("lengths", "decode_proof", "decode_pub", "pub_structure", "consensus", "accumulator", "step_accumulators"):
        built_s |= S[k]
    built_a = S["lengths"] | S["decode_proof"] | S["decode_pub"] | S["account_abi"]
    idx = torch.tensor(mine, dtype=torch.int64, device=dev)
    result = torch.ones(2 * half, dtype=torch.uint8, device=dev)
    pin = torch.empty(len(mine), dtype=torch.uint8).pin_memory()

    def step():
        _, rs = mb.verify_state_stages(hs, hq, mb.MODE_RLC)
        _, ra = mb.verify_account_stages(ha, haq)
        bits = [int(r.failed == 0 and (r.passed & built_s) == built_s) for r in rs] + [int(r.failed == 0 and (r.passed & built_a) == built_a) for r in ra]
        order = {g: j for j, g in enumerate(my_s + [half + x for x in my_a])}
        pin.copy_(torch.tensor([bits[order[g]] for g in mine], dtype=torch.uint8))
        return shard.merge_result_bytes(torch, dist, result, idx, pin.to(dev, non_blocking=True), world)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    assert torch.equal(step(), torch.tensor(want, dtype=torch.uint8, device=dev)), "mixed: result bytes differ from the expected bits"
    W = max(a.warmup, 3)
    for _ in range(W):
        step()
    sampler = ClockSampler()
    sampler.start()
    l0 = mb.launch_count()
    barrier()
    t0 = time.perf_counter()
    for _ in range(a.steps):
        step()
    barrier()
    t = torch.tensor([time.perf_counter() - t0], device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    launches = mb.launch_count() - l0
    sampler.stop.set()
    sampler.join()
    if rank != 0:
        return None
    v = 2 * half * a.steps / float(t.item())
    return {"metric": "mina_mixed_proofs_per_sec_built_stages", "value": v, "unit": "proofs/s", "n_gpus": world, "steps": a.steps, "warmup": W,
            "ms_per_step": float(t.item()) / a.steps * 1e3, "higher_is_better": True, "scaling": "strong", "vs_baseline": None,
            "dtype": "u256 modular (8x32 Montgomery)", "data": "synthetic",
            "config": {"workload": "mixed%d: %d proof-of-state + %d proof-of-account inputs per step (reference fixtures replicated, 1 %% corrupted), host buffers through the two batch entry points" % (2 * half, half, half),
                       "value_is_e2e": True, "built_stages": BUILT + ["account: bincode decode", "account: Solidity ABI re-encoding == public input"],
                       "absent_stages": ABSENT + ["account: Account::hash", "account: Merkle fold (kernel exists; needs the leaf hash and a trusted Poseidon table)"]},
            "roofline": None,
            "e2e": {"value": v, "unit": "proofs/s", "h2d_bytes_per_step": len(my_s) * (256 + 64 + 480 + 128 + 3 * 32), "d2h_bytes_per_step": 2 * 128 + 2 * half},
            "gpu_launches": int(launches), "clocks": sampler.summary()}


def main():
    global BATCH, N_CORRUPT
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200")
    ap.add_argument("--workload", default="state1024", choices=["state1024", "msm20", "ipa", "mixed", "merkle"])
    ap.add_argument("--profile-step", default="", choices=["", "rlc", "per_proof"],
                    help="run one extra untimed step inside a cudaProfilerStart/Stop range (for ncu --profile-from-start off)")
    ap.add_argument("--batch", type=int, default=BATCH, help="proofs per step (default: the 1024 of BASELINE.json; 64 = configs[2])")
    ap.add_argument("--corrupt", type=int, default=-1, help="corrupted members (default: 1 %% of the batch)")
    a = ap.parse_args()
    BATCH = a.batch
    N_CORRUPT = a.corrupt if a.corrupt >= 0 else max(1, round(BATCH / 100)) if BATCH != 1024 else N_CORRUPT
    if a.impl == "reference":
        return run_reference(a)

    import torch
    import torch.distributed as dist

    rank, world, local = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device; this library has no CPU fallback")
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    import __graft_entry__ as entry
    import mina_bridge_b200 as mb

    if rank == 0:
        entry.build()  # one builder; the others wait for the finished library / SRS cache
    if world > 1:
        dist.barrier()
    mb.init(local)
    dev = torch.device("cuda", local)
    line = {"state1024": bench_state, "msm20": bench_msm20, "ipa": bench_ipa, "mixed": bench_mixed, "merkle": bench_merkle}[a.workload](a, torch, dist, mb, rank, world, dev)
    if rank == 0:
        print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
