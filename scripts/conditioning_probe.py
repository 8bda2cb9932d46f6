"""How well-conditioned is the training gradient of the random-init network?  (CPU, fp32, specification path.)
Perturbs the source image by a relative 1e-7 / 1e-6 and compares the full gradient vectors; repeats with damped
residual branches (gain of the last BatchNorm of every ResNet block).  Evidence for the multi-GPU equivalence test
(tests/multigpu_worker.py): python scripts/conditioning_probe.py > profiles/conditioning_probe_r2.txt"""
import sys, torch, copy
sys.path.insert(0,'.')
from mine_b200 import config as C
from mine_b200.data.synthetic import synthetic_batch
from mine_b200.task import SynthesisTask
torch.set_num_threads(16)
H,W,S,B=128,192,4,2
base={"data.img_w":W,"data.img_h":H,"mpi.num_bins_coarse":S,"data.visible_point_count":64,"model.imagenet_pretrained":False,
      "mpi.fix_disparity":True,"data.per_gpu_batch_size":B,"lr.backbone_lr":0.0,"lr.decoder_lr":0.0}
items=synthetic_batch(B,H,W,64,seed=0)
def run(eps):
    cfg=C.config_for_dataset("llff",dict(base)); cfg["device"]=torch.device("cpu")
    torch.manual_seed(0)
    t=SynthesisTask(cfg,None)
    it=copy.deepcopy(items)
    if eps: it[0]["img"]=it[0]["img"]*(1+eps*torch.randn_like(it[0]["img"]))
    out=t.train_step(it)
    return t.arena.grad.clone().double(), float(out["loss"]), t
g0,l0,t=run(0); g1,l1,_=run(1e-7); g2,l2,_=run(1e-6)
cos=lambda a,b: torch.nn.functional.cosine_similarity(a,b,dim=0).item()
print("loss",l0,l1,l2)
print("cos eps=1e-7:",cos(g0,g1),"relerr",float((g0-g1).norm()/g0.norm()))
print("cos eps=1e-6:",cos(g0,g2),"relerr",float((g0-g2).norm()/g0.norm()))
# per-parameter breakdown for eps=1e-6
names=[n for n,_ in list(t.backbone.named_parameters())]+["dec."+n for n,_ in t.decoder.named_parameters()]
rows=[]
for i,(p,o) in enumerate(zip(t.arena.params,t.arena.offsets)):
    a,b=g0[o:o+p.numel()],g2[o:o+p.numel()]
    rows.append((float((a-b).norm()), float(a.norm()), names[i]))
rows.sort(reverse=True)
tot=float((g0-g2).norm())
for r in rows[:12]: print("%.3e  |g|=%.3e  %s"%r)
print("total diff",tot,"total norm",float(g0.norm()))

print("---- damped residual branches (last BN gamma of every encoder block = G) ----")
def run2(eps, G):
    cfg=C.config_for_dataset("llff",dict(base)); cfg["device"]=torch.device("cpu")
    torch.manual_seed(0)
    t=SynthesisTask(cfg,None)
    with torch.no_grad():
        for li in range(1,5):
            for blk in getattr(t.backbone.encoder,"layer%d"%li):
                blk.bn3.weight.fill_(G)
    it=copy.deepcopy(items)
    if eps: it[0]["img"]=it[0]["img"]*(1+eps*torch.randn_like(it[0]["img"]))
    out=t.train_step(it)
    return t.arena.grad.clone().double()
for G in (0.3, 0.1):
    a,b,c=run2(0,G),run2(1e-7,G),run2(1e-5,G)
    print("G",G,"cos 1e-7",cos(a,b),"rel",float((a-b).norm()/a.norm()),"cos 1e-5",cos(a,c),"rel",float((a-c).norm()/a.norm()), "norm",float(a.norm()))
