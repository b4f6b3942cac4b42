"""UNet3D on our own kernels (csrc/conv3d.cu): the fused tcgen05 layer
(GroupNorm-apply + [upsample + concat] + Conv3d + ReLU + next-layer statistics), max-pool and the whole
network against torch fp32 (cuDNN with TF32 off) and the reference-generated golden `unets.npz`.

Tolerance.  The kernels compute in single-pass TF32 with fp32 accumulation — the arithmetic the reference
itself uses on a GPU (cuDNN, torch.backends.cudnn.allow_tf32 = True is torch's default).  TF32 keeps 11
significant bits (relative rounding 4.9e-4 per operand), so against an fp32 result ONE layer is accurate to
~1e-3 of the output scale (tests: max <= 1e-2 * max|b|, mean <= 2e-3 * mean|b|) and the 11-layer network
to ~5e-3 (tests: max <= 2e-2 * max|b|, mean <= 1e-2 * mean|b|).  Measured on B200 for the shipped network at
32^3 / 64^3: ours 4.4e-3 / 4.7e-3 mean, cuDNN's TF32 kernels 4.4e-3 / 4.7e-3 mean against the same fp32
result — the test also requires that we are no further from fp32 than cuDNN's TF32 by more than 25 %."""
import numpy as np
import pytest
import torch
import torch.nn.functional as F

from util import load

pytestmark = pytest.mark.gpu


def _err(a, b):
    a, b = a.double(), b.double()
    return float((a - b).abs().max() / b.abs().max()), float((a - b).abs().mean() / b.abs().mean())


def _layer_ref(x, x2, gn_w, gn_b, groups, w, bias, relu):
    """fp32 torch reference of one fused layer on NCDHW tensors."""
    if x2 is not None:
        x = torch.cat((x, F.interpolate(x2, size=x.shape[2:], mode='nearest')), 1)
    if gn_w is not None:
        x = F.group_norm(x, groups, gn_w, gn_b, 1e-5)
    y = F.conv3d(x, w, bias, padding=w.shape[2] // 2)
    return F.relu(y) if relu else y


@pytest.mark.parametrize('case', ['32to32_16', '32to64_ragged', 'concat_96to32', 'k1_bias', 'batch2_128to32'])
def test_fused_conv_layer(case):
    from vtaco_b200 import _abi
    from vtaco_b200.encoder.unet3d import _pack_conv_weight
    import ctypes as C
    rs = np.random.RandomState(hash(case) % 1000)
    cfg = {'32to32_16': (1, 16, 16, 16, 32, 0, 32, 3, True, True),
           '32to64_ragged': (1, 5, 24, 12, 32, 0, 64, 3, True, True),
           'concat_96to32': (1, 8, 16, 16, 32, 64, 32, 3, True, True),
           'k1_bias': (1, 6, 16, 8, 32, 0, 32, 1, False, False),
           'batch2_128to32': (2, 4, 16, 8, 128, 0, 32, 3, True, True)}[case]
    N, D, H, W, C1, C2, Cout, k, gn, relu = cfg
    Cin = C1 + C2
    x = torch.from_numpy(rs.randn(N, C1, D, H, W).astype(np.float32) * 1.5 + 0.3).cuda()
    x2 = torch.from_numpy(rs.randn(N, C2, D // 2, H // 2, W // 2).astype(np.float32)).cuda() if C2 else None
    w = torch.from_numpy((rs.randn(Cout, Cin, k, k, k) / np.sqrt(Cin * k ** 3)).astype(np.float32)).cuda()
    bias = torch.from_numpy(rs.randn(Cout).astype(np.float32)).cuda() if not gn else None
    gw = torch.from_numpy(rs.uniform(0.5, 1.5, Cin).astype(np.float32)).cuda() if gn else None
    gb = torch.from_numpy(rs.randn(Cin).astype(np.float32) * 0.2).cuda() if gn else None
    prev = torch.backends.cudnn.allow_tf32
    torch.backends.cudnn.allow_tf32 = False
    try:
        ref = _layer_ref(x, x2, gw, gb, 8, w, bias, relu)
    finally:
        torch.backends.cudnn.allow_tf32 = prev
    L = _abi.lib()
    st = _abi.stream_ptr(x.device)
    xcl = x.permute(0, 2, 3, 4, 1).contiguous()
    x2cl = x2.permute(0, 2, 3, 4, 1).contiguous() if C2 else None
    a = _abi.Conv3dArgs()
    a.x, a.N, a.D, a.H, a.W, a.C1 = xcl.data_ptr(), N, D, H, W, C1
    if C2:
        a.x2, a.C2, a.D2, a.H2, a.W2 = x2cl.data_ptr(), C2, D // 2, H // 2, W // 2
    wp = _pack_conv_weight(w)
    a.w_packed, a.Cout, a.ksize, a.relu = wp.data_ptr(), Cout, k, int(relu)
    if bias is not None:
        a.bias = bias.data_ptr()
    if gn:
        stats = torch.zeros(N, Cin, 2, dtype=torch.float64, device='cuda')
        for n in range(N):
            _abi.check(L.vtaco_channel_stats_cl(_abi.ptr(xcl[n]), D * H * W, C1, _abi.ptr(stats[n, :C1]), st), 'stats')
            if C2:
                s2 = torch.zeros(C2, 2, dtype=torch.float64, device='cuda')
                _abi.check(L.vtaco_channel_stats_cl(_abi.ptr(x2cl[n]), D * H * W // 8, C2, _abi.ptr(s2), st), 'stats')
                stats[n, C1:] = 8.0 * s2
        # the statistics kernel itself
        full = x if not C2 else torch.cat((x, F.interpolate(x2, size=x.shape[2:], mode='nearest')), 1)
        want = torch.stack([full.double().sum((2, 3, 4)), (full.double() ** 2).sum((2, 3, 4))], -1)
        assert torch.allclose(stats, want, rtol=1e-5, atol=1e-3)
        a.in_stats, a.gamma, a.beta, a.groups, a.eps = stats.data_ptr(), gw.data_ptr(), gb.data_ptr(), 8, 1e-5
    y = torch.empty(N, D, H, W, Cout, device='cuda')
    ys = torch.zeros(N, Cout, 2, dtype=torch.float64, device='cuda')
    a.y, a.out_stats = y.data_ptr(), ys.data_ptr()
    _abi.check(L.vtaco_conv3d_cl(C.byref(a), st), 'conv3d_cl')
    got = y.permute(0, 4, 1, 2, 3)
    emax, emean = _err(got, ref)
    assert emax <= 1e-2 and emean <= 2e-3, (case, emax, emean)
    want = torch.stack([got.double().sum((2, 3, 4)), (got.double() ** 2).sum((2, 3, 4))], -1)
    assert torch.allclose(ys, want, rtol=1e-4, atol=1e-2), (ys - want).abs().max()


def test_maxpool_cl():
    from vtaco_b200 import _abi
    x = torch.randn(1, 64, 8, 12, 16, device='cuda')
    xcl = x.permute(0, 2, 3, 4, 1).contiguous()
    y = torch.empty(1, 4, 6, 8, 64, device='cuda')
    stats = torch.zeros(64, 2, dtype=torch.float64, device='cuda')
    _abi.check(_abi.lib().vtaco_maxpool2_cl(_abi.ptr(xcl), _abi.ptr(y), 1, 8, 12, 16, 64, _abi.ptr(stats),
                                            _abi.stream_ptr(x.device)), 'maxpool2_cl')
    ref = F.max_pool3d(x, 2)
    assert torch.equal(y.permute(0, 4, 1, 2, 3), ref)
    want = torch.stack([ref.double().sum((0, 2, 3, 4)), (ref.double() ** 2).sum((0, 2, 3, 4))], -1)
    assert torch.allclose(stats, want, rtol=1e-5, atol=1e-4)


@pytest.mark.parametrize('R,B', [(32, 2), (64, 1)])
def test_unet3d_fused_vs_fp32_modules(R, B):
    """the whole network (VTacO_YCB kwargs) on our kernels vs the same module on torch fp32 (cuDNN TF32 off);
    cuDNN's TF32 result is measured beside it as the 'reference on a GPU' yardstick."""
    from vtaco_b200.encoder.unet3d import UNet3D
    torch.manual_seed(3)
    net = UNet3D(num_levels=4, f_maps=32, in_channels=32, out_channels=32).cuda().eval()
    with torch.no_grad():
        for m in net.modules():
            if isinstance(m, torch.nn.GroupNorm):
                m.weight.uniform_(0.5, 1.5)
                m.bias.normal_(0, 0.2)
    x = torch.randn(B, 32, R, R, R, device='cuda') * (torch.rand(B, 32, R, R, R, device='cuda') < 0.05)   # sparse like scatter_mean output
    prev = torch.backends.cudnn.allow_tf32
    with torch.no_grad():
        got = net(x)
        assert got.permute(0, 2, 3, 4, 1).is_contiguous()        # channels_last_3d: the decoder's layout, no copy
        net.fused = False
        try:
            torch.backends.cudnn.allow_tf32 = False
            ref = net(x)
            torch.backends.cudnn.allow_tf32 = True
            cud = net(x)
        finally:
            torch.backends.cudnn.allow_tf32 = prev
            net.fused = True
    emax, emean = _err(got, ref)
    cmax, cmean = _err(cud, ref)
    print('fused vs fp32: max %.2e mean %.2e | cuDNN TF32 vs fp32: max %.2e mean %.2e' % (emax, emean, cmax, cmean))
    assert emax <= 2e-2 and emean <= 1e-2, (emax, emean)
    assert emean <= 1.25 * cmean + 1e-4 and emax <= 1.25 * cmax + 1e-4
    fmax, _ = _err(got, cud)
    assert fmax <= 5e-3, fmax                 # and we agree with cuDNN's TF32 result itself to ~1e-3


def test_unet3d_fused_vs_reference_golden():
    """unet3d_shipped.npz: output of the REFERENCE's UNet3D (shipped kwargs, fp32 CPU) on a seeded 16^3 input;
    the parameters are re-drawn from the same numpy stream (tests/util.randomise)."""
    from util import randomise, rs_randn
    from vtaco_b200.encoder.unet3d import UNet3D
    g = load('unet3d_shipped.npz')
    net = UNet3D(in_channels=32, out_channels=32, num_levels=4, f_maps=32)
    randomise(net, int(g['seed_w']))
    net = net.cuda().eval()
    sx = [int(v) for v in g['seed_x']]
    x = (rs_randn(sx[0], 1, 32, 16, 16, 16) * (np.random.RandomState(sx[1]).rand(1, 32, 16, 16, 16) < 0.2)).astype(np.float32)
    with torch.no_grad():
        got = net(torch.from_numpy(x).cuda())
        net.fused = False
        prev = torch.backends.cudnn.allow_tf32
        torch.backends.cudnn.allow_tf32 = False
        try:
            mod = net(torch.from_numpy(x).cuda())
        finally:
            torch.backends.cudnn.allow_tf32 = prev
    ref = torch.from_numpy(g['y'])
    assert _err(mod.cpu(), ref)[0] < 1e-4          # our module definition == the reference's network
    emax, emean = _err(got.cpu(), ref)
    assert emax <= 2e-2 and emean <= 1e-2, (emax, emean)
