"""3-D U-Net applied to the feature grid — state_dict-compatible with reference
src/encoder/unet3d.py:361-491 (encoders.N.basic_module.SingleConv{1,2}.{groupnorm,conv,...},
decoders.N.basic_module..., final_conv).  Library-backed (torch.nn / cuDNN), SURVEY §2 row 8."""
from collections import OrderedDict

import torch
import torch.nn as nn
import torch.nn.functional as F


class GroupNorm(nn.GroupNorm):
    """nn.GroupNorm with the same parameters / state_dict; on CUDA fp32 inference it runs
    vtaco_group_norm (chip-wide reduction) instead of ATen's one-block-per-group kernel."""
    prefer_channels_last = False

    def forward(self, x):
        if (x.is_cuda and x.dtype == torch.float32 and x.dim() >= 3
                and not (torch.is_grad_enabled() and (x.requires_grad or self.weight.requires_grad))):
            from .. import _abi
            N, Cc = x.shape[0], x.shape[1]
            S = x[0, 0].numel()
            ws = torch.empty(2 * N * self.num_groups, dtype=torch.float64, device=x.device)
            perm = (0,) + tuple(range(2, x.dim())) + (1,)
            cpg = Cc // self.num_groups
            # channels-last kernel only on request: measured on B200, cuDNN's fp32 Conv3d is not faster in
            # channels_last_3d at these sizes (UNet3D 64^3 x 32: e2e 5.27 ms contiguous vs 5.48 ms channels-last)
            if (self.prefer_channels_last and x.dim() in (4, 5) and x.permute(*perm).is_contiguous() and Cc % 4 == 0 and cpg % 4 == 0
                    and 256 % (Cc // 4) == 0 and self.num_groups <= 64):
                y = torch.empty_like(x)          # keeps the channels-last strides
                with torch.cuda.device(x.device):
                    st = _abi.lib().vtaco_group_norm_cl(_abi.ptr(x), _abi.ptr(y), _abi.ptr(self.weight),
                                                        _abi.ptr(self.bias), N, Cc, self.num_groups, S, float(self.eps),
                                                        _abi.ptr(ws), _abi.stream_ptr(x.device))
                _abi.check(st, 'group_norm_cl')
                return y
            xc = x.contiguous()
            y = torch.empty_like(xc)
            with torch.cuda.device(x.device):
                st = _abi.lib().vtaco_group_norm(_abi.ptr(xc), _abi.ptr(y), _abi.ptr(self.weight), _abi.ptr(self.bias),
                                                 N, Cc, self.num_groups, S, float(self.eps), _abi.ptr(ws),
                                                 _abi.stream_ptr(x.device))
            _abi.check(st, 'group_norm')
            return y
        return super().forward(x)


def _upsample_concat(skip, x):
    """cat(skip, interpolate(x, size=skip.shape[2:], mode='nearest'), dim=1) (reference unet3d.py
    Decoder.forward); one fused kernel on CUDA fp32 inference, the two ATen ops otherwise."""
    if (skip.is_cuda and x.is_cuda and skip.dtype == torch.float32 and x.dtype == torch.float32 and skip.dim() == 5
            and skip.size(4) % 4 == 0 and skip.is_contiguous() and x.is_contiguous()
            and not (torch.is_grad_enabled() and (skip.requires_grad or x.requires_grad))):
        from .. import _abi
        N, C1, Do, Ho, Wo = skip.shape
        C2, Di, Hi, Wi = x.shape[1:]
        out = torch.empty((N, C1 + C2, Do, Ho, Wo), dtype=torch.float32, device=skip.device)
        L = _abi.lib()
        with torch.cuda.device(skip.device):
            st = L.vtaco_upsample_concat3d(_abi.ptr(skip), _abi.ptr(x), _abi.ptr(out), N, C1, C2, Do, Ho, Wo, Di, Hi, Wi,
                                           _abi.stream_ptr(skip.device))
        _abi.check(st, 'upsample_concat3d')
        return out
    x = F.interpolate(x, size=skip.size()[2:], mode='nearest')
    return torch.cat((skip, x), dim=1)


def _single_conv(cin, cout, order, num_groups, kernel_size=3, padding=1):
    assert 'c' in order, 'Conv layer MUST be present'
    assert order[0] not in 'rle', 'Non-linearity cannot be the first operation in the layer'
    mods = OrderedDict()
    for i, ch in enumerate(order):
        if ch == 'r':
            mods['ReLU'] = nn.ReLU(inplace=True)
        elif ch == 'l':
            mods['LeakyReLU'] = nn.LeakyReLU(negative_slope=0.1, inplace=True)
        elif ch == 'e':
            mods['ELU'] = nn.ELU(inplace=True)
        elif ch == 'c':
            mods['conv'] = nn.Conv3d(cin, cout, kernel_size, padding=padding, bias=not ('g' in order or 'b' in order))
        elif ch == 'g':
            nch = cin if i < order.index('c') else cout
            groups = 1 if nch < num_groups else num_groups
            assert nch % groups == 0
            mods['groupnorm'] = GroupNorm(num_groups=groups, num_channels=nch)
        elif ch == 'b':
            mods['batchnorm'] = nn.BatchNorm3d(cin if i < order.index('c') else cout)
        else:
            raise ValueError("Unsupported layer type '%s'. MUST be one of ['b', 'g', 'r', 'l', 'e', 'c']" % ch)
    return nn.Sequential(mods)


def _double_conv(cin, cout, encoder, order, num_groups):
    if encoder:
        c1 = max(cout // 2, cin)
        shapes = ((cin, c1), (c1, cout))
    else:
        shapes = ((cin, cout), (cout, cout))
    return nn.Sequential(OrderedDict([
        ('SingleConv1', _single_conv(shapes[0][0], shapes[0][1], order, num_groups)),
        ('SingleConv2', _single_conv(shapes[1][0], shapes[1][1], order, num_groups))]))


class _Encoder(nn.Module):
    def __init__(self, cin, cout, apply_pooling, order, num_groups):
        super().__init__()
        self.pooling = nn.MaxPool3d(kernel_size=(2, 2, 2)) if apply_pooling else None
        self.basic_module = _double_conv(cin, cout, True, order, num_groups)

    def forward(self, x):
        if self.pooling is not None:
            x = self.pooling(x)
        return self.basic_module(x)


class _Decoder(nn.Module):
    def __init__(self, cin, cout, order, num_groups):
        super().__init__()
        self.basic_module = _double_conv(cin, cout, False, order, num_groups)

    def forward(self, skip, x):
        return self.basic_module(_upsample_concat(skip, x))


class UNet3D(nn.Module):
    def __init__(self, in_channels, out_channels, final_sigmoid=True, f_maps=64, layer_order='gcr',
                 num_groups=8, num_levels=4, is_segmentation=True, testing=False, **kwargs):
        super().__init__()
        self.testing = testing
        if isinstance(f_maps, int):
            f_maps = [f_maps * 2 ** k for k in range(num_levels)]
        self.encoders = nn.ModuleList([
            _Encoder(in_channels if i == 0 else f_maps[i - 1], f, i > 0, layer_order, num_groups)
            for i, f in enumerate(f_maps)])
        rev = list(reversed(f_maps))
        self.decoders = nn.ModuleList([
            _Decoder(rev[i] + rev[i + 1], rev[i + 1], layer_order, num_groups) for i in range(len(rev) - 1)])
        self.final_conv = nn.Conv3d(f_maps[0], out_channels, 1)
        if is_segmentation:
            self.final_activation = nn.Sigmoid() if final_sigmoid else nn.Softmax(dim=1)
        else:
            self.final_activation = None

    def forward(self, x):
        feats = []
        for enc in self.encoders:
            x = enc(x)
            feats.insert(0, x)
        for dec, skip in zip(self.decoders, feats[1:]):
            x = dec(skip, x)
        x = self.final_conv(x)
        if self.testing and self.final_activation is not None:
            x = self.final_activation(x)
        return x
