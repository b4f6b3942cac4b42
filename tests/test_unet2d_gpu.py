"""2-D U-Net of the feature planes on our kernels (vtaco_conv3d_cl on depth-1 volumes, vtaco_maxpool2d_cl,
vtaco_depth_to_space2_cl; reference src/encoder/unet.py:45-239) against the torch.nn modules in fp32 (cuDNN with
TF32 off), against cuDNN's TF32 kernels, and against `unet2d_shipped.npz` — the output of the REFERENCE's UNet
with the shipped kwargs.  Arithmetic: single-pass TF32 with fp32 accumulation = the reference on a GPU; the bars
are the ones of the UNet3D tests."""
import numpy as np
import pytest
import torch

from util import load, randomise, rs_randn

pytestmark = pytest.mark.gpu


def _err(a, b):   # (max deviation / max magnitude, mean deviation / mean magnitude), as in test_unet3d_gpu.py
    a, b = a.double(), b.double()
    return float((a - b).abs().max() / b.abs().max()), float((a - b).abs().mean() / b.abs().mean())


def _net(seed=5, **kw):
    from vtaco_b200.encoder.unet import UNet
    args = dict(in_channels=32, depth=4, merge_mode='concat', start_filts=32)
    args.update(kw)
    net = UNet(32, **args)
    randomise(net, seed)
    return net.cuda().eval()


def _reference_modules(net, x, tf32):
    prev_c, prev_m = torch.backends.cudnn.allow_tf32, torch.backends.cuda.matmul.allow_tf32
    torch.backends.cudnn.allow_tf32 = torch.backends.cuda.matmul.allow_tf32 = tf32
    net.fused = False
    try:
        with torch.no_grad():
            return net(x)
    finally:
        net.fused = True
        torch.backends.cudnn.allow_tf32, torch.backends.cuda.matmul.allow_tf32 = prev_c, prev_m


def test_plane_ops_match_torch():
    """vtaco_maxpool2d_cl == MaxPool2d(2) bit for bit; 1x1 conv + vtaco_depth_to_space2_cl == ConvTranspose2d(2, 2)."""
    from vtaco_b200 import _abi
    L = _abi.lib()
    st = _abi.stream_ptr(torch.device('cuda'))
    x = torch.randn(3, 12, 20, 32, device='cuda')                       # channels-last (N,H,W,C)
    y = torch.empty(3, 6, 10, 32, device='cuda')
    _abi.check(L.vtaco_maxpool2d_cl(_abi.ptr(x), _abi.ptr(y), 3, 12, 20, 32, st), 'maxpool2d')
    ref = torch.nn.functional.max_pool2d(x.permute(0, 3, 1, 2), 2).permute(0, 2, 3, 1)
    assert torch.equal(y, ref)
    t = torch.randn(2, 5, 7, 4 * 16, device='cuda')
    up = torch.empty(2, 10, 14, 16, device='cuda')
    _abi.check(L.vtaco_depth_to_space2_cl(_abi.ptr(t), _abi.ptr(up), 2, 5, 7, 16, st), 'd2s')
    ref = t.view(2, 5, 7, 2, 2, 16).permute(0, 1, 3, 2, 4, 5).reshape(2, 10, 14, 16)
    assert torch.equal(up, ref)


@pytest.mark.parametrize('B,R', [(1, 32), (3, 64), (2, 128)])
def test_unet2d_fused_vs_torch(B, R):
    """the fused path against the torch.nn modules: fp32 (TF32 off) and cuDNN TF32; planes of the shipped
    resolutions (32 hand encoder, 64 / 128 tri-plane configs), batch of planes."""
    net = _net()
    x = torch.from_numpy(rs_randn(7, B, 32, R, R)).cuda()
    x = x * (torch.rand_like(x) < 0.3)
    with torch.no_grad():
        got = net(x)
        assert net._fusable(x)
    assert got.shape == (B, 32, R, R)
    assert got.permute(0, 2, 3, 1).is_contiguous()                       # channels-last storage for the decoder
    ref32 = _reference_modules(net, x, tf32=False)
    reftf = _reference_modules(net, x, tf32=True)
    emax, emean = _err(got, ref32)
    cmax, cmean = _err(reftf, ref32)
    assert emax <= 2e-2 and emean <= 1e-2, (emax, emean)
    assert emean <= max(1.25 * cmean, 1e-4), (emean, cmean)              # not further from fp32 than cuDNN's TF32 kernels


def test_unet2d_fused_vs_reference_golden():
    """unet2d_shipped.npz: output of the REFERENCE's UNet (shipped kwargs, fp32 CPU) on two seeded 32 x 32 planes;
    the parameters are re-drawn from the same numpy stream (tests/util.randomise)."""
    from vtaco_b200.encoder.unet import UNet
    g = load('unet2d_shipped.npz')
    net = UNet(32, in_channels=32, depth=4, merge_mode='concat', start_filts=32)
    randomise(net, int(g['seed_w']))
    net = net.cuda().eval()
    sx = [int(v) for v in g['seed_x']]
    x = (rs_randn(sx[0], 2, 32, 32, 32) * (np.random.RandomState(sx[1]).rand(2, 32, 32, 32) < 0.3)).astype(np.float32)
    xt = torch.from_numpy(x).cuda()
    with torch.no_grad():
        got = net(xt)
    mod = _reference_modules(net, xt, tf32=False)
    ref = torch.from_numpy(g['y'])
    assert _err(mod.cpu(), ref)[0] < 1e-4          # our module definition == the reference's network
    emax, emean = _err(got.cpu(), ref)
    assert emax <= 2e-2 and emean <= 1e-2, (emax, emean)


def test_unet2d_falls_back_to_modules():
    """widths the kernels do not cover (the 8-channel golden network), grad mode and 'add' merging use torch.nn."""
    from vtaco_b200.encoder.unet import UNet
    small = UNet(8, in_channels=8, depth=3, start_filts=8).cuda().eval()
    x = torch.randn(1, 8, 16, 16, device='cuda')
    with torch.no_grad():
        assert not small._fusable(x)
    net = _net()
    xr = torch.randn(1, 32, 32, 32, device='cuda', requires_grad=True)
    assert not net._fusable(xr)
    out = net(xr)
    out.sum().backward()
    assert xr.grad is not None
    with torch.no_grad():
        assert not _net(merge_mode='add')._fusable(x.new_zeros(1, 32, 32, 32))


def test_triplane_encoder_with_fused_unet():
    """LocalPoolPointnet(plane_type xz/xy/yz, unet=True) end to end on the GPU == the same encoder with the torch.nn
    U-Net in fp32, at the TF32 tolerance."""
    from vtaco_b200.encoder import encoder_dict
    enc = encoder_dict['pointnet_local_pool'](dim=3, c_dim=32, hidden_dim=32, plane_type=['xz', 'xy', 'yz'],
                                              plane_resolution=32, unet=True,
                                              unet_kwargs=dict(depth=4, merge_mode='concat', start_filts=32))
    randomise(enc, 11)
    enc = enc.cuda().eval()
    p = torch.from_numpy(np.random.RandomState(12).uniform(-0.5, 0.5, size=(2, 3000, 3)).astype(np.float32)).cuda()
    with torch.no_grad():
        got = enc(p)
    ref = _reference_modules_enc(enc, p)
    for k in ('xz', 'xy', 'yz'):
        emax, emean = _err(got[k], ref[k])
        assert emax <= 2e-2 and emean <= 1e-2, (k, emax, emean)


def _reference_modules_enc(enc, p):
    prev = torch.backends.cudnn.allow_tf32
    torch.backends.cudnn.allow_tf32 = False
    enc.unet.fused = False
    try:
        with torch.no_grad():
            return enc(p)
    finally:
        enc.unet.fused = True
        torch.backends.cudnn.allow_tf32 = prev
