"""Per-kernel timing of the hot sm_100a kernels at the headline shapes (LLFF 384x256, N=32, B=2 -> 64 planes).
CUDA events, 3 warm-ups, L2 flushed between timed iterations.  Prints achieved bandwidth / FLOP rate and the
fraction of the measured peaks in MEASURED_PEAKS.json.  `--only NAME` runs a single case (for ncu)."""
import argparse, json, os, sys, torch
sys.path.insert(0, '.')
from mine_b200.ops import conv_engine as E
from mine_b200.ops import cuda as K

ap = argparse.ArgumentParser(); ap.add_argument("--only", default=None); ap.add_argument("--iters", type=int, default=10)
ap.add_argument("--precision", default="tf32", choices=["tf32", "bf16"])
ap.add_argument("--out", default=None)
args = ap.parse_args()
E.set_precision(args.precision)
ACT = E.ACT_DTYPE
ES = 4 if args.precision == "tf32" else 2
# compute roofline denominator: measured bf16 cuBLAS burst; TF32 tensor peak is nominally half of it
PEAK_SCALE = 0.5 if args.precision == "tf32" else 1.0
peaks = json.load(open("MEASURED_PEAKS.json")) if os.path.exists("MEASURED_PEAKS.json") else {"hbm_gbs": 6650.0, "bf16_tflops": 1590.0}
dev = torch.device("cuda")
flush = torch.zeros(64 * 1024 * 1024, device=dev)
N = 64


def timeit(fn):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    tot = 0.0
    for _ in range(args.iters):
        flush.add_(1.0)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); fn(); e1.record(); e1.synchronize()
        tot += e0.elapsed_time(e1)
    return tot / args.iters


rows = []


def report(name, ms, bytes_, flops):
    gbs, tf = bytes_ / ms / 1e6, flops / ms / 1e9
    rows.append((name, ms, gbs, gbs / peaks["hbm_gbs"], tf, tf / (peaks["bf16_tflops"] * PEAK_SCALE)))
    print("%-34s %8.3f ms  %8.1f GB/s (%.2f of measured HBM)  %7.1f TFLOP/s (%.3f of measured bf16%s)" %
          (rows[-1] + ("/2 = tf32" if PEAK_SCALE != 1.0 else "",)), flush=True)


def want(name):
    return args.only is None or args.only == name


layers = [  # name, h (low-res for up), w, Ci, Co, up
    ("up_4_1", 8, 12, 256, 256, True), ("same_3_0", 16, 24, 256, 128, False), ("up_3_1", 16, 24, 128, 128, True),
    ("same_2_0", 32, 48, 128, 64, False), ("up_2_1", 32, 48, 64, 64, True), ("same_1_0", 64, 96, 64, 32, False),
    ("up_1_1", 64, 96, 32, 32, True), ("same_0_0", 128, 192, 32, 16, False), ("up_0_1", 128, 192, 16, 16, True)]
for name, h, w, ci, co, up in layers:
    xp = torch.randn(N, h + 2, w + 2, ci, device=dev).to(ACT)
    wt = torch.randn(co, ci, 3, 3, device=dev) * 0.05
    stats = torch.zeros(2, co, device=dev)
    ho, wo = (2 * h, 2 * w) if up else (h, w)
    dy = torch.randn(N, ho, wo, co, device=dev).to(ACT)
    fl = 2.0 * N * h * w * ci * co * (16 if up else 9)
    by_f = xp.numel() * ES + N * ho * wo * co * ES
    if want("fprop_" + name):
        f = (lambda: E.conv_up_raw(xp, wt, stats=stats)) if up else (lambda: E.conv_same_raw(xp, wt, stats=stats))
        report("fprop_" + name, timeit(f), by_f, fl)
    if want("dgrad_" + name):
        f = (lambda: E.dgrad_up_raw(dy, wt)) if up else (lambda: E.dgrad_same_raw(dy, wt))
        report("dgrad_" + name, timeit(f), by_f, fl)
    if want("wgrad_" + name):
        f = (lambda: E.wgrad_up_raw(dy, xp)) if up else (lambda: E.wgrad_same_raw(dy, xp))
        report("wgrad_" + name, timeit(f), by_f, fl)
    del xp, dy
# heads
for s_, (h, w, c) in enumerate([(256, 384, 16), (128, 192, 32), (64, 96, 64), (32, 48, 128)]):
    if want("head_%d" % s_):
        xp = torch.randn(N, h + 2, w + 2, c, device=dev).to(ACT)
        wt, b = torch.randn(4, c, 3, 3, device=dev) * 0.05, torch.zeros(4, device=dev)
        report("head_%d" % s_, timeit(lambda: E.conv_same_raw(xp, wt, chan_bias=b, head=True)), xp.numel() * ES + N * h * w * 17,
               2.0 * N * h * w * c * 4 * 9)
        del xp
# elementwise companions on the largest activation
if want("bn_act_pad_fwd"):
    y = torch.randn(N, 256, 384, 16, device=dev).to(ACT)
    st = torch.stack([y.float().sum((0, 1, 2)), (y.float() ** 2).sum((0, 1, 2))]).contiguous()
    g, b = torch.ones(16, device=dev), torch.zeros(16, device=dev)
    cnt = float(N * 256 * 384)
    report("bn_act_pad_fwd", timeit(lambda: E.ext().bn_act_pad_fwd(y, st, g, b, 0, cnt, 1e-5)), y.numel() * 2 * ES, 0)
    dap = torch.randn(N, 258, 386, 16, device=dev).to(ACT)
    report("bn_act_bwd_reduce", timeit(lambda: E.ext().bn_act_bwd_reduce(dap, y, st, g, b, 0, cnt, 1e-5)), y.numel() * 3 * ES, 0)
    gg, sums = E.ext().bn_act_bwd_reduce(dap, y, st, g, b, 0, cnt, 1e-5)
    report("bn_bwd_apply", timeit(lambda: E.ext().bn_bwd_apply(gg, y, st, g, sums, 32, False, False, cnt, 1e-5)), y.numel() * 3 * ES, 0)
    # default backward pair: sums only (2 tensors read) + apply from the padded gradient (2 read, 1 written)
    report("bn_act_bwd_sums", timeit(lambda: E.ext().bn_act_bwd_sums(dap, y, st, g, b, 0, cnt, 1e-5)), y.numel() * 2 * ES, 0)
    report("bn_bwd_apply_fused", timeit(lambda: E.ext().bn_bwd_apply_fused(dap, y, st, g, b, sums, 32, False, False, cnt, 1e-5, 0)),
           y.numel() * 3 * ES, 0)
# render
if want("render"):
    from mine_b200 import geometry as geo
    for (B, S, H, W) in [(2, 32, 256, 384), (1, 64, 256, 384)]:
        mpi = torch.rand(B, S, H, W, 4, device=dev)
        disp = torch.linspace(1, 0.01, S, device=dev)[None].repeat(B, 1)
        k = geo.fov_intrinsics(H, W).to(dev)[None].repeat(B, 1, 1); kinv = geo.inv3x3(k)
        G = torch.eye(4, device=dev)[None].repeat(B, 1, 1); G[:, 0, 3] = 0.05
        img = torch.rand(B, 3, H, W, device=dev)
        nb = mpi.numel() * 4
        report("render_src_fwd B%dS%d" % (B, S), timeit(lambda: K._ext.render_src_fwd(mpi, disp, kinv, img, False, True, 0, True)), 2 * nb, 0)
        report("render_tgt_fwd B%dS%d" % (B, S), timeit(lambda: K._ext.render_tgt_fwd(mpi, disp, G, kinv, k, False, 0)), nb, 0)
        rgb, depth, mask, wsum = K._ext.render_tgt_fwd(mpi, disp, G, kinv, k, False, 0)
        grgb, gd = torch.rand_like(rgb), torch.rand_like(depth)
        report("render_tgt_bwd B%dS%d" % (B, S), timeit(lambda: K._ext.render_tgt_bwd(mpi, disp, G, kinv, k, rgb, depth, wsum, grgb, gd, False, 0)), 2 * nb, 0)
if want("ssim"):
    a, b = torch.rand(2, 3, 256, 384, device=dev), torch.rand(2, 3, 256, 384, device=dev)
    report("ssim_fwd(+partials) 2x3x256x384", timeit(lambda: K._ext.ssim_fwd(a, b, True)), a.numel() * 4 * 5, 0)
os.makedirs("gpurun_out", exist_ok=True)
json.dump([dict(zip(["kernel", "ms", "GBps", "frac_hbm", "TFLOPs", "frac_bf16"], r)) for r in rows], open(args.out or "gpurun_out/kernel_bench_%s.json" % args.precision, "w"), indent=1)
