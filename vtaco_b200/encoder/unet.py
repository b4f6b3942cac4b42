"""2-D U-Net applied to the feature planes — state_dict-compatible with reference
src/encoder/unet.py:117-239 (down_convs.N.conv{1,2}, up_convs.N.{upconv,conv1,conv2},
conv_final).  Library-backed (torch.nn / cuDNN): SURVEY §2 row 7 keeps it out of the
hand-written kernel list; it is a "next" row in §8f."""
import torch
import torch.nn as nn
import torch.nn.functional as F


class _Down(nn.Module):
    def __init__(self, cin, cout, pooling):
        super().__init__()
        self.conv1 = nn.Conv2d(cin, cout, 3, padding=1)
        self.conv2 = nn.Conv2d(cout, cout, 3, padding=1)
        self.pooling = pooling
        if pooling:
            self.pool = nn.MaxPool2d(2, 2)

    def forward(self, x):
        x = F.relu(self.conv2(F.relu(self.conv1(x))))
        return (self.pool(x) if self.pooling else x), x


class _Up(nn.Module):
    def __init__(self, cin, cout, merge_mode, up_mode):
        super().__init__()
        self.merge_mode = merge_mode
        if up_mode == 'transpose':
            self.upconv = nn.ConvTranspose2d(cin, cout, 2, stride=2)
        else:
            self.upconv = nn.Sequential(nn.Upsample(mode='bilinear', scale_factor=2), nn.Conv2d(cin, cout, 1))
        self.conv1 = nn.Conv2d(2 * cout if merge_mode == 'concat' else cout, cout, 3, padding=1)
        self.conv2 = nn.Conv2d(cout, cout, 3, padding=1)

    def forward(self, skip, x):
        x = self.upconv(x)
        x = torch.cat((x, skip), 1) if self.merge_mode == 'concat' else x + skip
        return F.relu(self.conv2(F.relu(self.conv1(x))))


class UNet(nn.Module):
    def __init__(self, num_classes, in_channels=3, depth=4, start_filts=32, up_mode='transpose',
                 merge_mode='concat', **kwargs):  # unknown keys (e.g. the YAML typo `start_flits`) are swallowed
        super().__init__()
        if up_mode not in ('transpose', 'upsample'):
            raise ValueError('"{}" is not a valid mode for upsampling. Only "transpose" and '
                             '"upsample" are allowed.'.format(up_mode))
        if merge_mode not in ('concat', 'add'):
            raise ValueError('"{}" is not a valid mode for merging up and down paths. '
                             'Only "concat" and "add" are allowed.'.format(up_mode))
        if up_mode == 'upsample' and merge_mode == 'add':
            raise ValueError('up_mode "upsample" is incompatible with merge_mode "add"')
        self.num_classes, self.in_channels, self.start_filts, self.depth = num_classes, in_channels, start_filts, depth
        self.up_mode, self.merge_mode = up_mode, merge_mode
        downs, ups = [], []
        outs = in_channels
        for i in range(depth):
            ins, outs = outs, start_filts * (2 ** i)
            downs.append(_Down(ins, outs, pooling=i < depth - 1))
        for i in range(depth - 1):
            ins, outs = outs, outs // 2
            ups.append(_Up(ins, outs, merge_mode, up_mode))
        self.down_convs = nn.ModuleList(downs)
        self.up_convs = nn.ModuleList(ups)
        self.conv_final = nn.Conv2d(outs, num_classes, 1)
        for m in self.modules():
            if isinstance(m, nn.Conv2d):  # reference weight_init (ConvTranspose2d keeps its default)
                nn.init.xavier_normal_(m.weight)
                nn.init.constant_(m.bias, 0)

    def forward(self, x):
        skips = []
        for d in self.down_convs:
            x, before = d(x)
            skips.append(before)
        for i, u in enumerate(self.up_convs):
            x = u(skips[-(i + 2)], x)
        return self.conv_final(x)
