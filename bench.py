#!/usr/bin/env python
"""Headline benchmark: ELBO training throughput of the Probabilistic U-Net (BASELINE.json configs[1]).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--mode train|ensemble] [--batch B]

One "step" = the reference's training-loop body (train_prob_unet_model.py:89-92):
    optimizer.zero_grad(); loss, recon, kl = model.elbo(x, target); loss.backward(); optimizer.step()
on a synthetic ClimEx-shaped batch (tests/golden/synth.py) of 64 samples per GPU, 3 variables, 128x128 tiles,
latent_dim 16, bf16 tensor-core arithmetic with fp32 accumulation, fp32 master weights, AdamW(lr=1e-3).

Printed JSON (one line, rank 0): metric / value / e2e / roofline / cpu_baseline / clocks / gpu_launches ... as the
driver's contract asks.  `--impl reference` times the reference algorithm's CPU path (oracle/probunet_oracle.py, the
pinned restatement of prob_unet.py/networks.py; /root/reference itself is not present on the GPU box) on a bounded
sample of the same workload with all host threads.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

import torch

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, 'tests', 'golden'))
import synth  # noqa: E402

# algorithmic work per sample at 128x128, L=16 (FlopCounterMode on the unmodified reference)
GFLOP_FWD_BWD_PER_SAMPLE = 1170.36
LATENT = 16
TILE = 128


def measured_peaks():
    p = os.path.join(ROOT, 'MEASURED_PEAKS.json')
    if os.path.exists(p):
        d = json.load(open(p))
        return dict(source='measured', tflops=d['bf16_tflops'], tflops_sustained=d['bf16_tflops_sustained'], hbm=d['hbm_gbs'])
    return dict(source='fallback', tflops=1590.0, tflops_sustained=1400.0, hbm=6650.0)


class ClockSampler(threading.Thread):
    """Samples SM clocks / throttle reasons with nvidia-smi while the timed region runs."""

    def __init__(self, index=0):
        super().__init__(daemon=True)
        self.index = index
        self.rows = []
        self._halt = threading.Event()

    def run(self):
        if self._run_nvml():
            return
        q = ('clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,'
             'clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap')
        while not self._halt.is_set():
            try:
                out = subprocess.run(['nvidia-smi', f'--query-gpu={q}', '--format=csv,noheader,nounits', '-i',
                                      str(self.index)], capture_output=True, text=True, timeout=5).stdout.strip()
                if out:
                    self.rows.append([c.strip() for c in out.split(',')])
            except Exception:  # noqa: BLE001
                pass
            self._halt.wait(0.2)

    def _run_nvml(self):
        """Same columns through NVML (no subprocess, one sample every 20 ms); False if pynvml is unusable."""
        try:
            import pynvml
            pynvml.nvmlInit()
            h = pynvml.nvmlDeviceGetHandleByIndex(self.index)
            mx = pynvml.nvmlDeviceGetMaxClockInfo(h, pynvml.NVML_CLOCK_SM)
        except Exception:  # noqa: BLE001
            return False
        bits = (('hw_slowdown', 0x8), ('hw_thermal_slowdown', 0x40), ('sw_thermal_slowdown', 0x20), ('sw_power_cap', 0x4))
        while not self._halt.is_set():
            try:
                sm = pynvml.nvmlDeviceGetClockInfo(h, pynvml.NVML_CLOCK_SM)
                r = pynvml.nvmlDeviceGetCurrentClocksThrottleReasons(h)
                self.rows.append([str(sm), str(mx)] + ['Active' if r & b else 'Not Active' for _, b in bits])
            except Exception:  # noqa: BLE001
                pass
            self._halt.wait(0.02)
        return True

    def stop(self):
        self._halt.set()
        self.join(timeout=5)
        sm = sorted(int(r[0]) for r in self.rows if r[0].isdigit())
        reasons = set()
        names = ['hw_slowdown', 'hw_thermal_slowdown', 'sw_thermal_slowdown', 'sw_power_cap']
        for r in self.rows:
            for n, v in zip(names, r[2:6]):
                if v.lower().startswith('active'):
                    reasons.add(n)
        return dict(sm_mhz=sm[len(sm) // 2] if sm else None,
                    sm_max_mhz=int(self.rows[0][1]) if self.rows and self.rows[0][1].isdigit() else None,
                    reasons=sorted(reasons), samples=len(self.rows))


def make_batch(B, seed):
    x, t = synth.make_inputs(B, TILE, TILE, seed=seed)
    return x, t


# ----------------------------------------------------------------------------------------------------------------------
# reference arm / cpu baseline: the reference algorithm (oracle restatement) on the host cores
# ----------------------------------------------------------------------------------------------------------------------
def cpu_reference_steps(steps, warmup, sample_batch=2, threads=None):
    from oracle import probunet_oracle as O
    if threads:
        torch.set_num_threads(threads)
    schema = synth.load_schema(f'schema_probunet_L{LATENT}.json')
    sd = synth.make_weights(schema, seed=0)
    leaf = {k: (v.requires_grad_(True) if 'resample_filter' not in k else v) for k, v in sd.items()}
    live = [v for k, v in leaf.items() if v.requires_grad and 'map_layer' not in k]
    opt = torch.optim.AdamW(live, lr=1e-3)
    x, t = make_batch(sample_batch, seed=1)
    g = torch.Generator().manual_seed(7)
    times = []
    for i in range(warmup + steps):
        t0 = time.perf_counter()
        opt.zero_grad()
        masks = None   # no dropout masks: the CPU arm does marginally less work (<1%) than the reference's step
        eps = torch.randn(sample_batch, LATENT, generator=g)
        r = O.elbo(leaf, x, t, eps, dropout_masks=masks)
        r['total'].backward()
        opt.step()
        r['total'].item()
        dt = time.perf_counter() - t0
        if i >= warmup:
            times.append(dt)
    sec = sum(times) / len(times)
    return dict(value=sample_batch / sec, unit='samples/s', cores=torch.get_num_threads(), kind='port',
                sample=f'{len(times)} timed + {warmup} warm-up steps of zero_grad/elbo/backward/AdamW at batch '
                       f'{sample_batch}, 3x{TILE}x{TILE}, L={LATENT}, fp32, oracle/probunet_oracle.py on '
                       f'{torch.get_num_threads()} of {os.cpu_count()} host threads'), sec


def run_reference(args):
    rank = int(os.environ.get('RANK', '0'))
    if rank != 0:
        return
    steps = max(1, min(args.steps, 3))
    warm = 1
    # torchrun exports OMP_NUM_THREADS=1 for every rank; this arm runs on rank 0 alone and is meant to use the host
    cb, sec = cpu_reference_steps(steps, warm, sample_batch=2, threads=os.cpu_count() or 1)
    line = {
        'impl': 'reference', 'metric': 'elbo_train_samples_per_s', 'value': cb['value'], 'unit': 'samples/s',
        'n_gpus': args.gpus, 'steps': steps, 'warmup': warm, 'ms_per_step': sec * 1e3, 'higher_is_better': True,
        'scaling': 'weak', 'vs_baseline': None, 'dtype': 'f32', 'data': 'synthetic',
        'config': workload_config(args.gpus, 64, note='CPU reference arm: each step is a bounded sample (batch 2) of '
                                                       'the same workload'),
        'cpu_baseline': cb,
        'e2e': {'value': cb['value'], 'unit': 'samples/s', 'h2d_bytes_per_step': 0, 'd2h_bytes_per_step': 0},
    }
    print(json.dumps(line), flush=True)


def ncu_traffic():
    """DRAM bytes (read + write) of the dominant kernel's headline launch from the committed `ncu --set full` capture
    (profiles/r1_ncu_heavy_kernels.tsv, first conv_tc_kernel row: 256->256 3x3 at 128x128, batch 64)."""
    path = os.path.join(os.path.dirname(os.path.abspath(__file__)), 'profiles', 'r1_ncu_heavy_kernels.tsv')
    try:
        rows = [l.rstrip('\n').split('\t') for l in open(path)]
        hdr = rows[0]
        for r in rows[1:]:
            if r[0].startswith('conv_tc_kernel<256, 1>'):
                rd, wr = float(r[hdr.index('dram_rd_MB')]), float(r[hdr.index('dram_wr_MB')])
                alg = 2 * 64 * 128 * 128 * 256 * 2 / 1e6
                return {'traffic': (rd + wr) * 1e6,
                        'traffic_note': f'ncu dram read+write of one conv_tc_kernel<256,1> launch (256->256 3x3, 128x128, '
                                        f'batch 64): {rd + wr:.0f} MB vs {alg:.0f} MB algorithmic (input + output once); '
                                        'source profiles/r1_ncu_heavy_kernels.tsv'}
    except (OSError, ValueError, IndexError):
        pass
    return {}


def workload_config(n_gpus, per_gpu_batch, note=None):
    c = {'workload': 'ProbabilisticUNet ELBO training step (zero_grad, elbo fwd, backward, AdamW), 3x128x128 tiles, '
                     f'latent_dim {LATENT}, num_filters [64,128,256,512] (BASELINE.json configs[1])',
         'per_gpu_batch': per_gpu_batch, 'global_batch': per_gpu_batch * n_gpus, 'tile': TILE,
         'parallelism': f'dp{n_gpus}', 'l2_policy': 'working set >> L2 (activations of one step are tens of GB)'}
    if note:
        c['note'] = note
    return c


# ----------------------------------------------------------------------------------------------------------------------
# our arm
# ----------------------------------------------------------------------------------------------------------------------
def run_ours(args):
    import torch.distributed as dist
    from prob_unet_mds_b200 import ProbabilisticUNet
    from prob_unet_mds_b200 import _lib as L
    from prob_unet_mds_b200 import parallel

    world = int(os.environ.get('WORLD_SIZE', '1'))
    rank = int(os.environ.get('RANK', '0'))
    local_rank = int(os.environ.get('LOCAL_RANK', '0'))
    torch.cuda.set_device(local_rank)
    dev = torch.device('cuda', local_rank)
    if world > 1:
        dist.init_process_group('nccl', device_id=dev)
    B = args.batch
    torch.manual_seed(1234 + rank)
    model = ProbabilisticUNet(3, 3, latent_dim=LATENT, num_filters=[64, 128, 256, 512]).to(dev)
    model.load_state_dict(synth.make_weights(synth.load_schema(f'schema_probunet_L{LATENT}.json'), seed=0))
    model.set_precision(args.precision)
    model.train()
    if args.torch_adamw:
        opt = torch.optim.AdamW(model.parameters(), lr=1e-3, fused=True)
    else:
        from prob_unet_mds_b200 import AdamW          # one-launch multi-tensor AdamW, torch.optim.AdamW semantics
        opt = AdamW(model.parameters(), lr=1e-3)
    # gradient all-reduce (SUM) overlapped with backward; attaches itself to the model
    ddp = parallel.DataParallel(model) if world > 1 else None  # noqa: F841

    x_host, t_host = make_batch(B, seed=1 + rank)
    x_pin, t_pin = x_host.pin_memory(), t_host.pin_memory()
    x_dev, t_dev = x_pin.to(dev), t_pin.to(dev)

    def step(x, t):
        opt.zero_grad(set_to_none=True)
        total, recon, kl = model.elbo(x, t)
        total.backward()
        opt.step()
        return total

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    # ---- warm-up
    for _ in range(args.warmup):
        step(x_dev, t_dev)
    barrier()

    # ---- timed region 1: inputs resident in HBM, per-call events for the roofline of the dominant kernel
    lib = L.lib()
    sampler = ClockSampler(local_rank) if rank == 0 else None
    if sampler:
        sampler.start()
    prof = L.start_profiling() if args.profile_calls else None
    lib.pu_launch_count(1)
    s_evt, e_evt = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    s_evt.record()
    for _ in range(args.steps):
        step(x_dev, t_dev)
    e_evt.record()
    barrier()
    launches = int(L._raw_lib().pu_launch_count(0))
    L.stop_profiling()
    ms_resident = s_evt.elapsed_time(e_evt) / args.steps
    clocks = sampler.stop() if sampler else None

    # ---- timed region 2: end to end through the public API with host buffers (H2D of the batch, D2H of the loss)
    barrier()
    s_evt.record()
    for _ in range(args.steps):
        x = x_pin.to(dev, non_blocking=True)
        t = t_pin.to(dev, non_blocking=True)
        loss = step(x, t)
        loss_host = loss.item()
    e_evt.record()
    barrier()
    ms_e2e = s_evt.elapsed_time(e_evt) / args.steps

    # ---- secondary metric: ensemble member-samples/s.  U-Net + prior once per input, then the
    # fused Fcomb decode for 100 latent samples; inputs sharded over ranks (no communication), outputs stay on device.
    ms_ens = None
    if args.ensemble_members > 0:
        model.eval()
        with torch.no_grad():
            for _ in range(2):
                model.sample_ensemble(x_dev, args.ensemble_members)
            barrier()
            s_evt.record()
            for _ in range(args.steps):
                ens = model.sample_ensemble(x_dev, args.ensemble_members)
            e_evt.record()
            barrier()
            ms_ens = s_evt.elapsed_time(e_evt) / args.steps
            del ens
        model.train()

    if world > 1:
        tms = torch.tensor([ms_resident, ms_e2e, ms_ens or 0.0], device=dev, dtype=torch.float64)
        dist.all_reduce(tms, op=dist.ReduceOp.MAX)
        ms_resident, ms_e2e, ms_ens_max = tms.tolist()
        ms_ens = ms_ens_max if ms_ens is not None else None

    if rank == 0:
        peaks = measured_peaks()
        global_batch = B * world
        value = global_batch / (ms_resident / 1e3)
        e2e_value = global_batch / (ms_e2e / 1e3)
        roof = None
        kinds = {}
        if prof is not None:
            kinds = prof.summary()
            tc = [kinds[k] for k in ('conv_tc', 'wgrad_tc') if k in kinds]
            if tc:
                fl = sum(d['flops'] for d in tc)
                ms = sum(d['ms'] for d in tc)
                calls = sum(d['calls'] for d in tc)
                achieved = fl / (ms / 1e3) / 1e12
                roof = {'bound': 'tensor', 'kernel': 'conv_tc_kernel + wgrad_tc_kernel (tcgen05 implicit GEMM)',
                        'achieved': achieved, 'peak': peaks['tflops_sustained'], 'unit': 'TFLOP/s',
                        'frac': achieved / peaks['tflops_sustained'],
                        'peak_source': f"{peaks['source']} bf16_tflops_sustained (kernel timed inside a long step)",
                        'flops_per_launch': fl / calls, 'avg_launch_ms': ms / calls, 'launches_timed': calls,
                        'share_of_step': ms / (ms_resident * args.steps), 'traffic': None}
                roof.update(ncu_traffic())
        step_tflops = GFLOP_FWD_BWD_PER_SAMPLE * global_batch / 1e3 / (ms_resident / 1e3)
        line = {
            'metric': 'elbo_train_samples_per_s', 'value': value, 'unit': 'samples/s', 'n_gpus': world,
            'steps': args.steps, 'warmup': args.warmup, 'ms_per_step': ms_resident, 'higher_is_better': True,
            'scaling': 'weak', 'vs_baseline': None, 'dtype': 'bf16' if args.precision == 'bf16' else 'f32',
            'data': 'synthetic',
            'config': dict(workload_config(world, B), optimizer='torch.optim.AdamW(fused=True)' if args.torch_adamw
                           else 'prob_unet_mds_b200.AdamW (pu_adamw_multi)'),
            'e2e': {'value': e2e_value, 'unit': 'samples/s', 'ms_per_step': ms_e2e,
                    'h2d_bytes_per_step': int(x_pin.numel() * 4 + t_pin.numel() * 4), 'd2h_bytes_per_step': 4},
            'gpu_launches': launches,
            'step_tflops_algorithmic': step_tflops,
            'step_frac_of_bf16_peak': step_tflops / world / peaks['tflops_sustained'],
            'roofline': roof, 'clocks': clocks, 'last_loss': loss_host,
            'kernel_ms_per_step': {k: round(v['ms'] / args.steps, 3) for k, v in sorted(kinds.items(), key=lambda kv: -kv[1]['ms'])},
        }
        if ms_ens is not None:
            line['ensemble'] = {'metric': 'ensemble_member_samples_per_s',
                                'value': global_batch * args.ensemble_members / (ms_ens / 1e3),
                                'unit': 'member-samples/s', 'ms_per_step': ms_ens,
                                'config': {'inputs_per_gpu': B, 'members': args.ensemble_members, 'tile': TILE,
                                           'sharding': 'inputs over ranks, all members local'}}
        if world == 1 and not args.no_cpu_baseline:
            cb, _ = cpu_reference_steps(1, 1, sample_batch=2)
            line['cpu_baseline'] = cb
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--gpus', type=int, default=1)
    ap.add_argument('--steps', type=int, default=5)
    ap.add_argument('--warmup', type=int, default=3)
    ap.add_argument('--impl', default='ours', choices=['ours', 'reference'])
    ap.add_argument('--batch', type=int, default=64, help='per-GPU batch')
    ap.add_argument('--precision', default='bf16', choices=['bf16', 'fp32'])
    ap.add_argument('--no-cpu-baseline', action='store_true')
    ap.add_argument('--torch-adamw', action='store_true',
                    help="use torch.optim.AdamW(fused=True) instead of the package's multi-tensor AdamW")
    ap.add_argument('--ensemble-members', type=int, default=100,
                    help='members per input of the secondary ensemble metric (0 disables it)')
    ap.add_argument('--no-profile-calls', dest='profile_calls', action='store_false')
    args = ap.parse_args()
    if args.impl == 'reference':
        run_reference(args)
        return
    if not torch.cuda.is_available():
        raise SystemExit('bench.py: no CUDA device; the product path has no CPU fallback '
                         '(use --impl reference for the CPU arm)')
    world = int(os.environ.get('WORLD_SIZE', '1'))
    if args.gpus > 1 and world == 1:
        # convenience: re-launch under torchrun
        cmd = [sys.executable, '-m', 'torch.distributed.run', '--nnodes=1', f'--nproc-per-node={args.gpus}',
               '--master-addr', '127.0.0.1', '--master-port', '29511', os.path.abspath(__file__)] + sys.argv[1:]
        raise SystemExit(subprocess.call(cmd))
    run_ours(args)


if __name__ == '__main__':
    main()
