#!/usr/bin/env python3
"""Quick device-side timing of the fixed-base MSM (not the contract bench; see bench.py)."""
import argparse
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402

import mina_bridge_b200 as mb  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--curve", type=int, default=1)
    ap.add_argument("--n", type=int, default=65536)
    ap.add_argument("--nmsm", type=int, nargs="+", default=[1, 16, 64])
    ap.add_argument("--c", type=int, default=16)
    ap.add_argument("--leaf", type=int, default=8)
    ap.add_argument("--iters", type=int, default=5)
    a = ap.parse_args()
    mb.init(0)
    if a.c != 16 or a.leaf != 8:
        mb.msm_configure(a.curve, a.c, True, a.leaf)
    dev = torch.device("cuda:0")
    g = torch.Generator(device=dev)
    g.manual_seed(1)
    for nmsm in a.nmsm:
        sc = torch.randint(0, 2**31 - 1, (nmsm, a.n, 8), dtype=torch.int32, device=dev, generator=g)
        sc[:, :, 7] &= 0x1FFFFFFF  # < 2^253 < p
        out = torch.zeros((nmsm, 16), dtype=torch.int32, device=dev)
        st = torch.cuda.current_stream().cuda_stream
        for _ in range(2):
            mb.msm_srs_device(a.curve, nmsm, a.n, sc.data_ptr(), out.data_ptr(), st)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        acc = 0.0
        for _ in range(a.iters):
            acc += mb.msm_srs_device(a.curve, nmsm, a.n, sc.data_ptr(), out.data_ptr(), st, True)
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / a.iters
        print(
            "curve=%d n=%d nmsm=%d c=%d: %.3f ms/call (%.3f ms/MSM), accumulate kernel %.3f ms, %.1f Mpoints/s"
            % (a.curve, a.n, nmsm, a.c, ms, ms / nmsm, acc / a.iters, nmsm * a.n / ms / 1e3)
        )


if __name__ == "__main__":
    main()
