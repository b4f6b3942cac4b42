"""Time LocalDecoder forward + backward (vtaco_decoder_backward) at the training shape of
SURVEY §8d config 2/3: B=32 x N=2048 queries, grid-64 features (+ optional c_img)."""
import json
import sys
import torch

sys.path.insert(0, '.')
from vtaco_b200.conv_onet.models import decoder_dict  # noqa


def timeit(fn, n=20, warm=3):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(n):
        fn()
    b.record()
    torch.cuda.synchronize()
    return a.elapsed_time(b) / n


def main():
    torch.manual_seed(0)
    out = {}
    for B, N, R, tag in ((32, 2048, 64, 'train_B32_N2048_grid64'), (1, 1 << 19, 64, 'B1_N524288_grid64')):
        dec = decoder_dict['simple_local'](dim=3, c_dim=32, padding=0.1, hidden_size=32).cuda().train()
        with torch.no_grad():
            for n, prm in dec.named_parameters():
                if n.endswith('fc_1.weight'):
                    prm.normal_(0, 0.1)
        feat = torch.randn(B, 32, R, R, R, device='cuda').contiguous(memory_format=torch.channels_last_3d).requires_grad_(True)
        p = torch.rand(B, N, 3, device='cuda') - 0.5
        c_img = torch.randn(B, N, 32, device='cuda', requires_grad=True)
        r = torch.randn(B, N, device='cuda')
        res = {}
        for mode in ('forward', 'forward_img'):
            def fwd():
                if mode == 'forward':
                    return dec(p, {'grid': feat})
                return dec.forward_img(p, {'grid': feat}, c_img)

            def step():
                dec.zero_grad(set_to_none=True)
                feat.grad = None
                c_img.grad = None
                (fwd() * r).sum().backward()

            with torch.no_grad():
                t_f = timeit(fwd)
            t_fb = timeit(step)
            res[mode] = {'fwd_ms': t_f, 'fwd_bwd_ms': t_fb, 'Mqueries_per_s_fwd_bwd': B * N / t_fb / 1e3}
        out[tag] = res
    # PointNet part of the encoder (no UNet): B clouds of 3640 points into a 64^3 grid
    from vtaco_b200.encoder import encoder_dict
    for B, T, tag in ((8, 3640, 'pointnet_B8_T3640_grid64'), (32, 3000, 'pointnet_B32_T3000_tri32')):
        kw = dict(plane_type='grid', grid_resolution=64) if 'grid' in tag else \
            dict(plane_type=['xz', 'xy', 'yz'], plane_resolution=32)
        enc = encoder_dict['pointnet_local_pool'](dim=3, c_dim=32, padding=0.1, hidden_dim=32, **kw).cuda().train()
        with torch.no_grad():
            for n, prm in enc.named_parameters():
                if n.endswith('fc_1.weight'):
                    prm.normal_(0, 0.1)
        cloud = torch.rand(B, T, 3, device='cuda') - 0.5
        with torch.no_grad():
            fea = enc(cloud)
        rs = {k: torch.randn_like(v) for k, v in fea.items()}

        def efwd():
            return enc(cloud)

        def estep():
            enc.zero_grad(set_to_none=True)
            torch.autograd.backward(list(efwd().values()), list(rs.values()))

        with torch.no_grad():
            t_f = timeit(efwd)
        t_fb = timeit(estep)
        out[tag] = {'fwd_ms': t_f, 'fwd_bwd_ms': t_fb, 'Mpoints_per_s_fwd_bwd': B * T / t_fb / 1e3}
    print(json.dumps(out, indent=1))
    with open('gpurun_out/bwd_probe.json', 'w') as f:
        json.dump(out, f, indent=1)


if __name__ == '__main__':
    main()
