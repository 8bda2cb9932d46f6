import sys, torch, torch.nn.functional as F
sys.path.insert(0, '.')
sys.path.insert(0, 'tests')
from test_conv_engine_gpu import _bf, _rand, _nhwc, _nchw, _rel
from mine_b200.models.norm import BatchNorm
from mine_b200.ops import conv_engine as E
n, s, h, w, ci, co = 4, 2, 12, 16, 32, 16
a = _bf(_rand((n, ci, h, w), 0))
wt = (_rand((co, ci, 3, 3), 1, 0.1)).requires_grad_(True)
gamma, beta = (_rand((co,), 4).abs() + 0.5).requires_grad_(True), (_rand((co,), 5) * 0.1).requires_grad_(True)
xlo = a.clone().requires_grad_(True)
up = F.interpolate(xlo, scale_factor=2, mode="nearest")
y = F.conv2d(F.pad(up, (1, 1, 1, 1), mode="reflect"), _bf(wt) + (wt - wt.detach()))
act = F.elu(F.batch_norm(y, None, None, gamma, beta, training=True, eps=1e-5))
ref_apad = F.pad(act, (1,1,1,1), mode='reflect')
gout = _bf(_rand(ref_apad.shape, 8))
(ref_apad * gout).sum().backward()
xlo2 = xlo.detach().clone().requires_grad_(True)
xpad = E.pad_nhwc(_nhwc(xlo2).to(torch.bfloat16), "replicate")
xpad.retain_grad()
bn = BatchNorm(co).cuda()
apad = E.PlaneConvBNAct.apply(xpad, wt.detach().clone().requires_grad_(True), None, None, None, gamma.detach().clone().requires_grad_(True), beta.detach().clone().requires_grad_(True), True, s, 0, bn, None)
print('fwd rel', _rel(_nchw(apad), ref_apad))
(apad * _nhwc(gout).to(torch.bfloat16)).sum().backward()
d = (xlo2.grad - xlo.grad).abs()
print('dx rel', _rel(xlo2.grad, xlo.grad), 'argmax', torch.nonzero(d == d.max())[0].tolist(), 'max ref', xlo.grad.abs().max().item())
print('err map (sum over n,c):'); print((d.sum((0,1)) / xlo.grad.abs().sum((0,1)).clamp(min=1e-6)))
# reference grad wrt replicate-padded input
xp_ref = F.pad(xlo.detach(), (1,1,1,1), mode='replicate').requires_grad_(True)
wp = E.pack_up(_bf(wt.detach()))
out = torch.zeros(n, co, 2*h, 2*w, device='cuda')
for py in range(2):
    for px in range(2):
        acc = 0
        for aa in range(2):
            for bb in range(2):
                win = xp_ref[:, :, py+aa:py+aa+h, px+bb:px+bb+w]
                acc = acc + torch.einsum('nihw,oi->nohw', win, wp[py*2+px, aa*2+bb])
        out[:, :, py::2, px::2] = acc
print('phase fwd vs direct', _rel(out, y))
act2 = F.elu(F.batch_norm(out, None, None, gamma.detach(), beta.detach(), training=True, eps=1e-5))
(F.pad(act2,(1,1,1,1),mode='reflect')*gout).sum().backward()
dp = (_nchw(xpad.grad).float() - xp_ref.grad).abs()
print('dxpad rel', _rel(_nchw(xpad.grad), xp_ref.grad)); print((dp.sum((0,1))/xp_ref.grad.abs().sum((0,1)).clamp(min=1e-6)))
