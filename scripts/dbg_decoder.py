import sys, torch
sys.path.insert(0, '.')
from mine_b200.models.decoder import DepthDecoder
from mine_b200.models.encoder import ResnetEncoder
from mine_b200.ops.conv_engine import ConvEngine
torch.manual_seed(0)
enc, dec = ResnetEncoder().cuda(), DepthDecoder().cuda()
b, s, h, w = 2, 4, 256, 256
img = torch.rand(b, 3, h, w, device="cuda")
disp = torch.rand(b, s, device="cuda") * 0.8 + 0.1
eng = ConvEngine(enc, dec, {}, torch.device("cuda"))
def rel(a, b): return (a.float() - b.float()).norm().item() / (b.float().norm().item() + 1e-12)
smooth = "--smooth" in sys.argv
for trial in range(2):
    outs = eng.predict(img, disp)
    gouts = [torch.randn(o.shape, device='cuda', generator=torch.Generator('cuda').manual_seed(k)) for k, o in enumerate(outs)]
    if smooth:
        for g in gouts: g[..., 3] = 0        # drop the non-smooth |sigma| channel
    sum((o * g).sum() for o, g in zip(outs, gouts)).backward()
    got = {k: p.grad.clone() for k, p in list(dec.named_parameters()) + [('enc.' + k, p) for k, p in enc.named_parameters()] if p.grad is not None}
    for p in list(enc.parameters()) + list(dec.parameters()): p.grad = None
    if trial == 0:
        got0 = got
print('run-to-run (engine vs engine) max rel:', max(rel(got[k], got0[k]) for k in got))
feats = enc(img)
ref = dec(feats, disp)
refs = [ref[("disp", k)].permute(0, 1, 3, 4, 2) for k in range(4)]
for k in range(4): print('fwd', k, rel(outs[k], refs[k]))
sum((o * g).sum() for o, g in zip(refs, gouts)).backward()
ref_grads = {k: p.grad.clone() for k, p in dec.named_parameters() if p.grad is not None}
for kname, p in dec.named_parameters():
    if kname in got: print('%-44s %.4f  |ref|=%.3e' % (kname, rel(got[kname], p.grad), p.grad.norm().item()))
encs = [(k, rel(got['enc.' + k], p.grad)) for k, p in enc.named_parameters() if 'enc.' + k in got]
print('encoder params: median rel', sorted(v for _, v in encs)[len(encs)//2], 'max', max(encs, key=lambda t: t[1]))
# the same comparison for the library bf16 path (autocast) vs fp32: how much of the deviation is just bf16?
for p in list(enc.parameters()) + list(dec.parameters()): p.grad = None
with torch.autocast("cuda", dtype=torch.bfloat16):
    feats_b = enc(img.contiguous(memory_format=torch.channels_last))
    out_b = dec(feats_b, disp)
outs_b = [out_b[("disp", k)].float().permute(0, 1, 3, 4, 2) for k in range(4)]
for k in range(4): print('autocast fwd', k, rel(outs_b[k], refs[k]))
ref_g = {k: p.grad for k, p in dec.named_parameters()}
sum((o * g).sum() for o, g in zip(outs_b, gouts)).backward()
print('autocast-bf16 vs fp32 (weights only):')
for kname, p in dec.named_parameters():
    if kname.endswith('conv.weight') or kname.endswith('0.weight'):
        print('   %-40s %.4f' % (kname, rel(p.grad, ref_grads[kname])))
