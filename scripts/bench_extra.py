"""Secondary measurements (not the driver's headline): ensemble sampling (BASELINE configs[3]), the deterministic
U-Net training step (configs[4]) and -- informational -- the reference algorithm through PyTorch eager on the same
B200 (the oracle restatement with CUDA tensors, i.e. cuDNN / cuBLAS).  Prints one JSON line per measurement."""
import argparse
import json
import os
import sys
import time

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, 'tests', 'golden'))
import synth  # noqa: E402


def timed(fn, steps, warmup):
    for _ in range(warmup):
        fn()
    torch.cuda.synchronize()
    s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    s.record()
    for _ in range(steps):
        fn()
    e.record()
    torch.cuda.synchronize()
    return s.elapsed_time(e) / steps


def ensemble(args):
    from prob_unet_mds_b200 import ProbabilisticUNet
    m = ProbabilisticUNet(3, 3, latent_dim=16).cuda()
    m.load_state_dict(synth.make_weights(synth.load_schema('schema_probunet_L16.json'), seed=0))
    m.eval()
    B, S = args.inputs, args.members
    x, _ = synth.make_inputs(B, 128, 128, seed=1)
    xp = x.pin_memory()
    xd = xp.cuda()
    out_host = torch.empty((B, S, 3, 128, 128), dtype=torch.float32).pin_memory()
    ms = timed(lambda: m.sample_ensemble(xd, S), args.steps, args.warmup)

    def e2e():
        out = m.sample_ensemble(xp.cuda(non_blocking=True), S)
        out_host.copy_(out, non_blocking=True)
        torch.cuda.synchronize()
    ms_e2e = timed(e2e, max(1, args.steps // 2), 1)
    # encode-only and decode-only split
    with torch.no_grad():
        from prob_unet_mds_b200 import ops
        dt = m.compute_dtype
        from prob_unet_mds_b200 import engine
        feat, _ = m.unet.engine().forward(engine.input_nhwc(xd, dt), False, False)
        mu, ls, _ = m.prior.engine(dt).forward(m.prior._input(xd, None, dt), save=False)
        ms_dec = timed(lambda: m.decode_ensemble(feat, mu, ls, S), args.steps, args.warmup)
    print(json.dumps({'metric': 'ensemble_member_samples_per_s', 'value': B * S / (ms / 1e3), 'unit': 'member-samples/s',
                      'ms_per_step': ms, 'decode_only_ms': ms_dec, 'config': {'inputs': B, 'members': S, 'tile': 128},
                      'e2e': {'value': B * S / (ms_e2e / 1e3), 'ms_per_step': ms_e2e,
                              'h2d_bytes_per_step': int(xp.numel() * 4), 'd2h_bytes_per_step': int(out_host.numel() * 4)}}),
          flush=True)


def detunet(args):
    from prob_unet_mds_b200.baseline.deterministic_unet import UNet
    m = UNet(img_resolution=(64, 64), in_channels=3, out_channels=3, label_dim=0, use_diffuse=False).cuda()
    m.load_state_dict(synth.make_weights(synth.load_schema('schema_detunet.json'), seed=3))
    m.train()
    B = args.det_batch
    x, t = synth.make_inputs(B, 256, 256, seed=5)
    x, t = x.cuda(), t.cuda()
    opt = torch.optim.AdamW(m.parameters(), lr=1e-3, fused=True)
    lossf = torch.nn.MSELoss()

    def step():
        opt.zero_grad(set_to_none=True)
        loss = lossf(m(x, class_labels=None), t)
        loss.backward()
        opt.step()
    ms = timed(step, args.steps, args.warmup)
    tfl = 807.66e9 * B / (ms / 1e3) / 1e12
    print(json.dumps({'metric': 'detunet_train_samples_per_s', 'value': B / (ms / 1e3), 'unit': 'samples/s',
                      'ms_per_step': ms, 'step_tflops_algorithmic': tfl,
                      'config': {'batch': B, 'tile': 256, 'model': 'baseline/deterministic_unet.UNet'}}), flush=True)


def eager(args):
    """The reference algorithm (oracle restatement) run by PyTorch eager on the GPU: cuDNN / cuBLAS kernels."""
    from oracle import probunet_oracle as O
    sd = synth.make_weights(synth.load_schema('schema_probunet_L16.json'), seed=0)
    for autocast in (False, True):
        B = args.eager_batch
        leaf = {k: (v.cuda().requires_grad_(True) if 'resample_filter' not in k else v.cuda()) for k, v in sd.items()}
        live = [v for k, v in leaf.items() if v.requires_grad and 'map_layer' not in k]
        opt = torch.optim.AdamW(live, lr=1e-3, fused=True)
        x, t = synth.make_inputs(B, 128, 128, seed=1)
        x, t = x.cuda(), t.cuda()
        torch.backends.cudnn.allow_tf32 = True
        torch.backends.cuda.matmul.allow_tf32 = True

        def step():
            opt.zero_grad(set_to_none=True)
            eps = torch.randn(B, 16, device='cuda')
            with torch.autocast('cuda', dtype=torch.bfloat16, enabled=autocast):
                r = O.elbo(leaf, x, t, eps)
            r['total'].backward()
            opt.step()
        try:
            ms = timed(step, args.steps, args.warmup)
            print(json.dumps({'metric': 'torch_eager_elbo_train_samples_per_s', 'value': B / (ms / 1e3),
                              'ms_per_step': ms, 'config': {'batch': B, 'tile': 128,
                                                            'precision': 'bf16 autocast' if autocast else 'tf32/fp32'}}),
                  flush=True)
        except torch.OutOfMemoryError:
            print(json.dumps({'metric': 'torch_eager_elbo_train_samples_per_s', 'error': 'OOM', 'batch': B}), flush=True)
        del leaf, opt
        torch.cuda.empty_cache()


def prepare(args):
    """Device-side ClimEx sample preparation for the bench batch, beside the reference's per-item CPU path."""
    import time
    from oracle import climex_oracle as CO
    from prob_unet_mds_b200 import data
    B = args.inputs
    g = torch.Generator().manual_seed(3)
    hr = torch.randn(B, 3, 128, 128, generator=g) * 3 + 1
    stats = CO.compute_stats(hr, 'perpixel')
    hr_d = hr.cuda()
    st_d = [s.cuda() for s in stats]
    ms = timed(lambda: data.prepare_batch(hr_d, 'perpixel', st_d), 50, 10)
    out = data.prepare_batch(hr_d, 'perpixel', st_d)
    res = torch.randn_like(hr_d)
    ms_inv = timed(lambda: data.residual_to_hr(res, out['lrinterp'], 'perpixel', st_d), 50, 10)
    t0 = time.perf_counter()
    for i in range(B):                      # what the DataLoader does: one __getitem__ per sample
        CO.prepare_batch(hr[i:i + 1], 'perpixel', stats)
    cpu_s = time.perf_counter() - t0
    nbytes = hr.numel() * 4 * 4 + stats[0].numel() * 8
    print(json.dumps({'metric': 'climex_prepare_samples_per_s', 'value': B / (ms / 1e3), 'ms_per_batch': ms,
                      'gbs': nbytes / ms / 1e6, 'residual_to_hr_ms': ms_inv,
                      'cpu_baseline': {'value': B / cpu_s, 'unit': 'samples/s', 'kind': 'port',
                                       'cores': torch.get_num_threads()},
                      'config': {'batch': B, 'tile': 128, 'standardization': 'perpixel'}}), flush=True)


def crps(args):
    """Empirical CRPS of the bench ensemble: the fused kernel, the reference function run on the same GPU
    by PyTorch (torch.sort over the member dimension), and the reference function on the host cores (bounded sample)."""
    import time
    from oracle import metrics_oracle as MO
    from prob_unet_mds_b200 import metrics
    B, S = args.inputs, args.members
    ens = torch.randn(B, S, 3, 128, 128, device='cuda')
    truth = torch.randn(B, 3, 128, 128, device='cuda')
    ms = timed(lambda: metrics.crps_ensemble(ens, truth), args.steps, args.warmup)

    def torch_gpu():
        pred = ens.transpose(0, 1)
        p = pred.sort(dim=0).values
        diff = p[1:] - p[:-1]
        w = (torch.arange(1, S, device='cuda') * torch.arange(S - 1, 0, -1, device='cuda')).float()
        w = w.reshape(w.shape + (1,) * (diff.dim() - 1))
        return (p - truth).abs().mean(0) - (diff * w).sum(0) / S ** 2
    ms_t = timed(torch_gpu, args.steps, args.warmup)
    nb = 2
    e_cpu, t_cpu = ens[:nb].transpose(0, 1).contiguous().cpu(), truth[:nb].cpu()
    t0 = time.perf_counter()
    MO.crps_empirical(e_cpu, t_cpu)
    cpu_s = time.perf_counter() - t0
    print(json.dumps({'metric': 'crps_member_samples_per_s', 'value': B * S / (ms / 1e3), 'ms_per_call': ms,
                      'gbs_read': ens.numel() * 4 / ms / 1e6, 'torch_same_gpu_ms': ms_t,
                      'cpu_baseline': {'value': nb * S / cpu_s, 'unit': 'member-samples/s', 'kind': 'port',
                                       'cores': torch.get_num_threads(), 'sample': f'{nb} inputs x {S} members'},
                      'config': {'inputs': B, 'members': S, 'tile': 128}}), flush=True)


if __name__ == '__main__':
    ap = argparse.ArgumentParser()
    ap.add_argument('what', nargs='+', choices=['ensemble', 'detunet', 'eager', 'prepare', 'crps'])
    ap.add_argument('--steps', type=int, default=3)
    ap.add_argument('--warmup', type=int, default=2)
    ap.add_argument('--inputs', type=int, default=64)
    ap.add_argument('--members', type=int, default=100)
    ap.add_argument('--det-batch', type=int, default=32)
    ap.add_argument('--eager-batch', type=int, default=16)
    a = ap.parse_args()
    for w in a.what:
        {'ensemble': ensemble, 'detunet': detunet, 'eager': eager, 'prepare': prepare, 'crps': crps}[w](a)
