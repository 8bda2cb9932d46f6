"""Model parity vs upstream modules with shared random weights + checkpoint round trips."""

import pytest
import torch

from mine_b200.models import checkpoint as ckpt
from mine_b200.models.decoder import DepthDecoder
from mine_b200.models.encoder import ResnetEncoder
from mine_b200.spec.embedder import get_embedder


def _randomise_bn(m):
    g = torch.Generator().manual_seed(7)
    for mod in m.modules():
        if hasattr(mod, "running_mean") and mod.running_mean is not None:
            mod.running_mean.copy_(torch.randn(mod.running_mean.shape, generator=g) * 0.1)
            mod.running_var.copy_(torch.rand(mod.running_var.shape, generator=g) + 0.5)
            mod.weight.data.copy_(torch.rand(mod.weight.shape, generator=g) + 0.5)
            mod.bias.data.copy_(torch.randn(mod.bias.shape, generator=g) * 0.1)


def test_mangled_keys():
    assert ckpt.mangle(("upconv", 4, 0)) == "(-'-u-p-c-o-n-v-'-,- -4-,- -0-)"
    dec = DepthDecoder()
    sd = ckpt.decoder_to_reference(dec, module_prefix=True)
    assert len(sd) == 102
    assert "module.convs.(-'-d-i-s-p-c-o-n-v-'-,- -3-).conv.bias" in sd
    assert "module.conv_down1.1.num_batches_tracked" in sd
    enc = ResnetEncoder()
    assert len(ckpt.backbone_to_reference(enc)) == 320


@pytest.mark.parametrize("train", [False, True])
def test_decoder_matches_reference(ref, train):
    rd = ref.load("network.monodepth2.depth_decoder")
    emb, edim = get_embedder(10)
    torch.manual_seed(0)
    ref_dec = rd.DepthDecoder(num_ch_enc=[64, 256, 512, 1024, 2048], embedder=emb, embedder_out_dim=edim).double()
    _randomise_bn(ref_dec)
    ours = DepthDecoder().double()
    res = ckpt.load_decoder(ours, ref_dec.state_dict())
    assert not res.missing_keys and not res.unexpected_keys
    ref_dec.train(train), ours.train(train)
    b, s, h, w = 2, 3, 128, 128
    g = torch.Generator().manual_seed(1)
    feats = [torch.randn(b, c, h // d, w // d, generator=g, dtype=torch.float64) * 0.5
             for c, d in zip([64, 256, 512, 1024, 2048], [2, 4, 8, 16, 32])]
    disp = torch.rand(b, s, generator=g, dtype=torch.float64) + 0.05
    with torch.no_grad():
        want = ref_dec([f.clone() for f in feats], disp)
        got = ours(feats, disp)
    for sc in range(4):
        assert got[("disp", sc)].shape == want[("disp", sc)].shape
        assert torch.allclose(got[("disp", sc)], want[("disp", sc)], rtol=1e-8, atol=1e-9), sc
    if train:   # running statistics advanced identically
        a = dict(ref_dec.named_buffers())
        mine = ckpt.decoder_to_reference(ours, module_prefix=False)
        for k, v in a.items():
            assert torch.allclose(mine[k].double(), v.double(), rtol=1e-8, atol=1e-10), k


def test_encoder_matches_torchvision():
    import torchvision
    torch.manual_seed(0)
    tv = torchvision.models.resnet50(weights=None)
    _randomise_bn(tv)
    enc = ResnetEncoder()
    res = ckpt.load_backbone(enc, {"module.encoder." + k: v for k, v in tv.state_dict().items()})
    assert not res.missing_keys and not res.unexpected_keys
    tv.eval(), enc.eval()
    x = torch.rand(2, 3, 64, 96)
    with torch.no_grad():
        xn = (x - enc.img_mean) / enc.img_std
        c1 = tv.relu(tv.bn1(tv.conv1(xn)))
        b1 = tv.layer1(tv.maxpool(c1)); b2 = tv.layer2(b1); b3 = tv.layer3(b2); b4 = tv.layer4(b3)
        got = enc(x)
    for a, b in zip(got, (c1, b1, b2, b3, b4)):
        assert (a - b).abs().max() <= 1e-4 * b.abs().max() + 1e-5


def test_checkpoint_roundtrip_and_reference_loader(ref, tmp_path):
    enc, dec = ResnetEncoder(), DepthDecoder()
    _randomise_bn(enc), _randomise_bn(dec)
    params = [{"params": enc.parameters(), "lr": 1e-3}, {"params": dec.parameters(), "lr": 1e-3}]
    opt = torch.optim.Adam(params, weight_decay=4e-5)
    for p in list(enc.parameters()) + list(dec.parameters()):
        p.grad = torch.randn_like(p) * 1e-3
    opt.step()
    path = str(tmp_path / "checkpoint_latest.pth")
    ckpt.save_checkpoint(path, enc, dec, opt, meta={"global_step": 12, "epoch": 3})
    raw = torch.load(path, weights_only=False)
    assert len(raw["backbone"]) == 320 and len(raw["decoder"]) == 102
    assert len(raw["optimizer"]["state"]) == 161 + 60
    # ours -> ours
    enc2, dec2 = ResnetEncoder(), DepthDecoder()
    opt2 = torch.optim.Adam([{"params": enc2.parameters(), "lr": 1e-3}, {"params": dec2.parameters(), "lr": 1e-3}],
                            weight_decay=4e-5)
    meta = ckpt.restore_model(path, enc2, dec2, opt2)
    assert meta["global_step"] == 12
    for a, b in zip(enc.state_dict().values(), enc2.state_dict().values()):
        assert torch.equal(a, b)
    for a, b in zip(dec.state_dict().values(), dec2.state_dict().values()):
        assert torch.equal(a, b)
    s1, s2 = opt.state_dict()["state"], opt2.state_dict()["state"]
    assert all(torch.equal(s1[i]["exp_avg"], s2[i]["exp_avg"]) for i in s1)
    # ours -> reference loader (unmodified utils.restore_model on reference modules)
    ru = ref.load("utils")
    rd = ref.load("network.monodepth2.depth_decoder")
    import torchvision
    emb, edim = get_embedder(10)

    class RefBackbone(torch.nn.Module):
        def __init__(self):
            super().__init__()
            self.encoder = torchvision.models.resnet50(weights=None)
    rb, rdec = RefBackbone(), rd.DepthDecoder(num_ch_enc=[64, 256, 512, 1024, 2048], embedder=emb, embedder_out_dim=edim)
    ropt = torch.optim.Adam([{"params": rb.parameters(), "lr": 1e-3}, {"params": rdec.parameters(), "lr": 1e-3}],
                            weight_decay=4e-5)
    ru.restore_model(path, rb, rdec, ropt, logger=None)
    mine = ckpt.decoder_to_reference(dec, module_prefix=False)
    for k, v in rdec.state_dict().items():
        assert torch.equal(v, mine[k]), k
    mine_b = ckpt.backbone_to_reference(enc, module_prefix=False)
    for k, v in rb.state_dict().items():
        if not k.startswith("encoder.fc."):
            assert torch.equal(v, mine_b[k]), k
