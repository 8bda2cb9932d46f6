"""Worker for tests/test_multigpu.py (launched by torch.distributed.run, one rank per GPU)."""
import json
import os
import sys

import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))


def main():
    from mine_b200.parallel import bootstrap
    from mine_b200.parallel.p2p import P2PComm
    ctx = bootstrap.init_distributed()
    dev, rank, world = ctx.device, ctx.rank, ctx.world_size
    res = {}
    comm = P2PComm(dev)
    res["multicast"] = bool(comm._small["mc"])
    # 1. small one-shot SUM, repeated (flag reuse / epoch bugs), odd sizes
    ok = True
    for it, n in enumerate([1, 7, 128, 513, 4097, 2 * 2048 + 1, 8192, 65536, 33, 33, 33, 256, 256]):
        g = torch.Generator(device="cpu").manual_seed(100 * it + rank)
        x = torch.randn(n, generator=g).to(dev)
        ref = x.clone()
        dist.all_reduce(ref)
        out = comm.allreduce_sum_(x.clone())
        ok &= bool(torch.allclose(out, ref, rtol=1e-5, atol=1e-5))
        gathered = [torch.empty_like(out) for _ in range(world)]
        dist.all_gather(gathered, out)
        ok &= all(torch.equal(gathered[0], t) for t in gathered)          # bit-identical on every rank
    res["small_ok"] = ok
    # 2. in-place mean on the symmetric arena, several buckets, both code paths
    numel = 3 * 1024 * 1024 + 64
    arena = comm.alloc_symmetric(numel)
    for use_mm in ([False, True] if res["multicast"] else [False]):
        comm.use_multimem = use_mm
        okb = True
        for it in range(3):
            g = torch.Generator(device="cpu").manual_seed(7 * it + rank)
            arena.copy_(torch.randn(numel, generator=g).to(dev))
            ref = arena.clone()
            dist.all_reduce(ref)
            ref /= world
            torch.cuda.synchronize()
            dist.barrier()
            side = torch.cuda.Stream(dev)
            side.wait_stream(torch.cuda.current_stream())
            for lo, hi in ((0, 1024 * 1024), (1024 * 1024, 3 * 1024 * 1024), (3 * 1024 * 1024, numel)):
                comm.allreduce_mean_(arena[lo:hi], stream=side)
            torch.cuda.current_stream().wait_stream(side)
            torch.cuda.synchronize()
            okb &= bool(torch.allclose(arena, ref, rtol=1e-5, atol=1e-6))
        res["mean_ok_multimem" if use_mm else "mean_ok_p2p"] = okb
    # 3. timing of the gradient-sized all-reduce (152 MB) vs NCCL
    big = 38 * 1024 * 1024
    comm2 = P2PComm(dev)
    a2 = comm2.alloc_symmetric(big)
    nccl_buf = torch.zeros(big, device=dev)

    def timeit(fn, iters=10):
        for _ in range(3):
            fn()
        torch.cuda.synchronize(); dist.barrier(); torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(iters):
            fn()
        e1.record(); torch.cuda.synchronize()
        t = torch.tensor([e0.elapsed_time(e1) / iters], device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())
    comm2.use_multimem = False
    res["ms_152MB_p2p"] = timeit(lambda: comm2.allreduce_mean_(a2))
    if res["multicast"]:
        comm2.use_multimem = True
        res["ms_152MB_multimem"] = timeit(lambda: comm2.allreduce_mean_(a2))
    res["ms_152MB_nccl"] = timeit(lambda: (dist.all_reduce(nccl_buf), nccl_buf.mul_(1.0 / world)))
    for n_small in (256, 4097):
        small = torch.zeros(n_small, device=dev)
        comm2.use_ll = True
        res["us_small%d_ll" % n_small] = 1e3 * timeit(lambda: comm2.allreduce_sum_(small), 50)
        comm2.use_ll = False
        res["us_small%d_oneshot" % n_small] = 1e3 * timeit(lambda: comm2.allreduce_sum_(small), 50)
        comm2.use_ll = True
        res["us_small%d_nccl" % n_small] = 1e3 * timeit(lambda: dist.all_reduce(small), 50)
    # 4. two-rank training step: own comm == NCCL comm (same weights, same data)
    from mine_b200 import config as C
    from mine_b200.data.synthetic import synthetic_batch
    from mine_b200.task import SynthesisTask
    # 256x256: the receptive-field block then normalises over >= 16 values per channel also at 8 ranks
    base = {"data.img_w": 256, "data.img_h": 256, "mpi.num_bins_coarse": 4, "data.visible_point_count": 32,
            "model.imagenet_pretrained": False, "mpi.fix_disparity": True, "data.per_gpu_batch_size": 1,
            "lr.backbone_lr": 0.0, "lr.decoder_lr": 0.0}
    grads = {}
    for kind in ("p2p", "nccl"):
        cfg = C.config_for_dataset("llff", dict(base, **{"engine.comm": kind}))
        cfg.update({"device": dev, "global_rank": rank})
        torch.manual_seed(0)
        task = SynthesisTask(cfg, None)
        items = synthetic_batch(world, 256, 256, 32, seed=0)
        mine = tuple({k: v[rank:rank + 1] for k, v in d.items()} for d in items)
        task.train_step(mine)
        torch.cuda.synchronize()
        grads[kind] = task.arena.grad.clone()
        res["comm_" + kind] = task.comm.name
    cos = torch.nn.functional.cosine_similarity(grads["p2p"], grads["nccl"], dim=0).item()
    res["step_grad_cos"] = cos
    if rank == 0:
        print("RESULT " + json.dumps(res), flush=True)
    dist.barrier()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
