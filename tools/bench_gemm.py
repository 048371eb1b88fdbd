"""Blendshape GEMM (U1+U2) at the batched sizes of BASELINE.json configs 3 and 5.

    python tools/bench_gemm.py [--T 7680]

Times omfs_flame_blend_gemm (both tcgen05 tf32x3 forms and the CUDA-core kernel) with CUDA events and prints
useful / executed TFLOP/s and the output bandwidth.  `ncu -k regex:flame_blend_tc` on this script
gives the tensor-pipe utilisation quoted in profiles/.
"""
import argparse
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import omfs_b200  # noqa: E402,F401
from omfs_b200 import runtime  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--T", type=int, nargs="+", default=[300, 7680])
    ap.add_argument("--reps", type=int, default=20)
    args = ap.parse_args()
    import torch
    L = runtime.load_library()
    runtime.check(L.omfs_device_check(0))
    dev = torch.device("cuda", 0)
    kpad, npad, n_useful = 136, 15488, 3 * 5143 + 15
    K3 = 3 * kpad
    g = torch.Generator(device="cpu").manual_seed(0)

    def split3(x, order):
        """tf32 hi/lo split laid out along K as the kernels expect ([hi|hi|lo] for A, [hi|lo|hi] for B)."""
        hi = (x.view(torch.int32) & -8192).view(torch.float32)
        r = x - hi
        lo = (r.view(torch.int32) & -8192).view(torch.float32)
        parts = {"A": [hi, hi, lo], "B": [hi, lo, hi]}[order]
        return torch.cat(parts, dim=1).contiguous()

    b_full = torch.randn(npad, kpad, generator=g) * 1e-3
    Bt = split3(b_full, "B").to(dev)
    base = torch.randn(npad, generator=g).to(dev)
    out = []
    for T in args.T:
        a_full = torch.randn(T, kpad, generator=g) * 0.5
        A = split3(a_full, "A").to(dev)
        C = torch.empty(T, npad, device=dev)
        stream = torch.cuda.current_stream()
        ref = (a_full.double().to(dev) @ b_full.double().to(dev).T + base.double()).float()
        names = {0: "tcgen05 tf32x3 (panel re-use)", 2: "tcgen05 tf32x3 (concatenated K)", 1: "cuda-core fp32"}
        for impl in (0, 2, 1):
            def run():
                runtime.check(L.omfs_flame_blend_gemm(T, kpad, npad, A.data_ptr(), Bt.data_ptr(), base.data_ptr(),
                                                      C.data_ptr(), impl, stream.cuda_stream or None))
            for _ in range(3):
                run()
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for _ in range(args.reps):
                run()
            e1.record()
            torch.cuda.synchronize()
            ms = e0.elapsed_time(e1) / args.reps
            useful = 2.0 * T * (100 + 36) * n_useful
            executed = 2.0 * T * K3 * npad
            C.zero_()
            run()
            torch.cuda.synchronize()
            out.append({"T": T, "impl": names[impl], "ms": ms, "max_abs_err_vs_fp64": float((C - ref).abs().max()),
                        "useful_TFLOPs": useful / ms / 1e9, "executed_TFLOPs": executed / ms / 1e9,
                        "output_GBs": 4.0 * T * npad / ms / 1e6})
    for o in out:
        print(json.dumps(o))


if __name__ == "__main__":
    main()
