"""Micro-benchmarks of the opt-in kernel variants at the LLFF 384x256, N=32, B=2 shapes (CUDA events, L2 flushed).

    python scripts/bench_optin.py                       # default kernels
    MINE_B200_HEAD=direct MINE_B200_BN_REDUCE=v2 python scripts/bench_optin.py
"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from mine_b200.ops import conv_engine as E  # noqa: E402

FLUSH = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")


def timeit(fn, iters=10):
    for _ in range(3):
        fn()
    ms = []
    for _ in range(iters):
        FLUSH.zero_()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        fn()
        b.record()
        b.synchronize()
        ms.append(a.elapsed_time(b))
    ms.sort()
    return ms[len(ms) // 2]


def main():
    n = 64
    print("MINE_B200_HEAD=%s MINE_B200_BN_REDUCE=%s" % (E.head_mode(), os.environ.get("MINE_B200_BN_REDUCE", "v1")))
    for lvl, (h, w, c) in enumerate([(256, 384, 16), (128, 192, 32)]):
        apad = torch.randn((n, h + 2, w + 2, c), device="cuda").to(torch.bfloat16)
        wt, bias = torch.randn((4, c, 3, 3), device="cuda") * 0.1, torch.zeros(4, device="cuda")
        t = timeit(lambda: E.HeadConv.apply(apad, wt, bias, False))
        mb = (apad.numel() * 2 + n * h * w * 17) / 1e6
        print("head_%d fwd   %.3f ms  %.0f GB/s" % (lvl, t, mb / t))
    for h, w, c in [(256, 384, 16), (128, 192, 32), (64, 96, 64)]:
        dapad = torch.randn((n, h + 2, w + 2, c), device="cuda").to(torch.bfloat16)
        y = torch.randn((n, h, w, c), device="cuda").to(torch.bfloat16)
        stats = torch.stack([y.float().sum((0, 1, 2)), (y.float() ** 2).sum((0, 1, 2))])
        gamma, beta = torch.ones(c, device="cuda"), torch.zeros(c, device="cuda")
        t = timeit(lambda: E.ext().bn_act_bwd_reduce(dapad, y, stats, gamma, beta, 0, float(n * h * w), 1e-5))
        mb = (dapad.numel() + 2 * y.numel()) * 2 / 1e6
        print("bn_act_bwd_reduce %dx%dx%d  %.3f ms  %.0f GB/s" % (h, w, c, t, mb / t))


if __name__ == "__main__":
    main()
