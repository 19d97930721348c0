#!/usr/bin/env python
"""Headline benchmark: ELBO training throughput of the Probabilistic U-Net (BASELINE.json configs[1]).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--batch B]

One "step" = the reference's training-loop body (train_prob_unet_model.py:89-92):
    optimizer.zero_grad(); loss, recon, kl = model.elbo(x, target); loss.backward(); optimizer.step()
on a synthetic ClimEx-shaped batch (tests/golden/synth.py) of 64 samples per GPU, 3 variables, 128x128 tiles,
latent_dim 16, bf16 tensor-core arithmetic with fp32 accumulation, fp32 master weights, AdamW(lr=1e-3), dropout on.

Printed JSON (one line, rank 0).  Contract keys: metric / value / e2e / roofline / cpu_baseline / clocks / gpu_launches.
Extra keys on the same line, one per BASELINE.json config as written:
    config3_global_batch_512   data-parallel step at GLOBAL batch 512 (512/N per GPU in micro-batches of 64 with local
                               gradient accumulation, one overlapped NCCL all-reduce per step)                 configs[2]
    ensemble / config4_members_sharded   100 latent samples per input: U-Net + prior once per input, Fcomb per sample;
                               members sharded over the ranks (feature all-gather) beside the zero-communication
                               variant (inputs sharded), each also end to end with the result copied to the host   configs[3]
    config5_det_unet           baseline/deterministic_unet.UNet training step, batch 32, 256x256 tiles          configs[4]
    dp_check                   (N > 1) identical gradient checksums on all ranks after the all-reduce, and the summed
                               loss / gradients equal a single-process run over the same global batch
    gpu_eager_reference        (N = 1, informational) the unmodified reference run by PyTorch eager on the same GPU
    parity                     measured logits / ELBO / KL errors of the benchmarked (bf16) mode against the fixtures

`--impl reference` times the UNMODIFIED reference (oracle/_ref, staged by oracle/make_ref.py) on the host cores at
configs[0] (batch 8, fp32, dropout on); the oracle port is the fallback when oracle/_ref is absent.
"""
import argparse
import contextlib
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

# algorithmic work per sample (FlopCounterMode on the unmodified reference)
GFLOP_FWD_BWD_PER_SAMPLE = 1170.36          # ProbabilisticUNet, 128x128, L=16
GFLOP_DET_FWD_BWD_PER_SAMPLE = 807.66       # deterministic U-Net, 256x256
GFLOP_ENSEMBLE_PER_INPUT = 386.13 + 1.87    # U-Net + prior once per input ...
GFLOP_ENSEMBLE_PER_MEMBER = 0.1405          # ... + Fcomb per member after the layer-0 hoist
LATENT = 16
TILE = 128
CPU_BATCH = 8                               # BASELINE.json configs[0]


def measured_peaks():
    p = os.path.join(ROOT, 'MEASURED_PEAKS.json')
    if os.path.exists(p):
        d = json.load(open(p))
        return dict(source='measured', tflops=d['bf16_tflops'], tflops_sustained=d['bf16_tflops_sustained'], hbm=d['hbm_gbs'])
    return dict(source='fallback', tflops=1590.0, tflops_sustained=1400.0, hbm=6650.0)


class ClockSampler(threading.Thread):
    """Samples SM clocks / throttle reasons (NVML, else nvidia-smi) while the timed region runs."""

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


def make_batch(B, seed, tile=TILE):
    return synth.make_inputs(B, tile, tile, seed=seed)


def probunet_weights():
    return synth.make_weights(synth.load_schema(f'schema_probunet_L{LATENT}.json'), seed=0)


# ----------------------------------------------------------------------------------------------------------------------
# reference arm / cpu baseline: the reference itself (oracle/_ref) on the host cores; the oracle port as the fallback
# ----------------------------------------------------------------------------------------------------------------------
def _reference_model(device):
    """The UNMODIFIED reference ProbabilisticUNet with the bench weights, or None when oracle/_ref is not staged."""
    from oracle import make_ref
    if not make_ref.available():
        return None
    ref = make_ref.import_reference()
    ref.device = torch.device(device)            # the module-level global the reference's constructor moves itself to
    torch.manual_seed(0)
    model = ref.ProbabilisticUNet(3, 3, latent_dim=LATENT, num_filters=[64, 128, 256, 512]).to(device)
    model.load_state_dict(probunet_weights())
    return model


def cpu_reference_steps(steps, warmup, batch=CPU_BATCH, threads=None):
    """zero_grad / elbo / backward / AdamW.step (train_prob_unet_model.py:89-92) on the host cores, fp32, dropout on."""
    if threads:
        torch.set_num_threads(threads)
    x, t = make_batch(batch, seed=1)
    model = _reference_model('cpu')
    torch.manual_seed(7)
    if model is not None:
        kind = 'reference'
        model.train()
        opt = torch.optim.AdamW(model.parameters(), lr=1e-3)       # main.py:95

        def step():
            opt.zero_grad()
            loss, _, _ = model.elbo(x, t)
            loss.backward()
            opt.step()
            return loss.item()
        what = 'the unmodified reference (oracle/_ref: prob_unet.py + networks.py), model.train() (dropout 0.10 on)'
    else:
        from oracle import probunet_oracle as O
        kind = 'port'
        sd = probunet_weights()
        leaf = {k: (v.requires_grad_(True) if 'resample_filter' not in k else v) for k, v in sd.items()}
        live = [v for k, v in leaf.items() if v.requires_grad and 'map_layer' not in k]
        opt = torch.optim.AdamW(live, lr=1e-3)
        g = torch.Generator().manual_seed(7)

        def step():
            opt.zero_grad()
            r = O.elbo(leaf, x, t, torch.randn(batch, LATENT, generator=g), dropout_masks=None)
            r['total'].backward()
            opt.step()
            return r['total'].item()
        what = 'oracle/probunet_oracle.py (oracle/_ref is not staged), dropout masks off (<1% less work)'
    times = []
    for i in range(warmup + steps):
        t0 = time.perf_counter()
        step()
        if i >= warmup:
            times.append(time.perf_counter() - t0)
    sec = sum(times) / len(times)
    return dict(value=batch / sec, unit='samples/s', cores=torch.get_num_threads(), kind=kind,
                sample=f'{len(times)} timed + {warmup} warm-up steps of zero_grad/elbo/backward/AdamW at batch {batch} '
                       f'(BASELINE.json configs[0]), 3x{TILE}x{TILE}, L={LATENT}, fp32, {what}, on '
                       f'{torch.get_num_threads()} of {os.cpu_count()} host threads'), sec


def run_reference(args):
    rank = int(os.environ.get('RANK', '0'))
    if rank != 0:
        return
    steps = max(1, min(args.steps, 2))
    warm = 1
    # torchrun exports OMP_NUM_THREADS=1 for every rank; this arm runs on rank 0 alone and is meant to use the host
    cb, sec = cpu_reference_steps(steps, warm, batch=args.cpu_batch, threads=os.cpu_count() or 1)
    line = {
        'impl': 'reference', 'metric': 'elbo_train_samples_per_s', 'value': cb['value'], 'unit': 'samples/s',
        'n_gpus': args.gpus, 'steps': steps, 'warmup': warm, 'ms_per_step': sec * 1e3, 'higher_is_better': True,
        'scaling': 'weak', 'vs_baseline': None, 'dtype': 'f32', 'data': 'synthetic',
        'config': workload_config(args.gpus, 64, note=f'CPU reference arm: each step is a bounded sample (batch {args.cpu_batch}, '
                                                       'BASELINE.json configs[0]) of the same workload'),
        'cpu_baseline': cb,
        'e2e': {'value': cb['value'], 'unit': 'samples/s', 'h2d_bytes_per_step': 0, 'd2h_bytes_per_step': 0},
    }
    print(json.dumps(line), flush=True)


def ncu_traffic():
    """DRAM bytes (read + write) of the dominant kernel's headline launch, read from the newest committed
    `ncu --set full` summary under profiles/ (NOT measured in this run: ncu cannot run inside the timed bench)."""
    pdir = os.path.join(ROOT, 'profiles')
    for name in ('r2_ncu_heavy_kernels.tsv', 'r1_ncu_heavy_kernels.tsv'):
        path = os.path.join(pdir, name)
        try:
            rows = [l.rstrip('\n').split('\t') for l in open(path)]
            hdr = rows[0]
            for r in rows[1:]:
                if r[0].startswith('conv_tc_kernel<256, 1'):      # first row: the plain 256->256 3x3 forward launch
                    rd, wr = float(r[hdr.index('dram_rd_MB')]), float(r[hdr.index('dram_wr_MB')])
                    alg = 2 * 64 * 128 * 128 * 256 * 2 / 1e6
                    return {'traffic': (rd + wr) * 1e6,
                            'traffic_source': f'committed ncu capture profiles/{name} (not measured in this run): dram '
                                              f'read+write of one conv_tc_kernel<256,1> forward launch (256->256 3x3, 128x128, '
                                              f'batch 64) = {rd + wr:.0f} MB vs {alg:.0f} MB algorithmic (input + output once)'}
        except (OSError, ValueError, IndexError):
            continue
    return {}


def workload_config(n_gpus, per_gpu_batch, note=None):
    c = {'workload': 'ProbabilisticUNet ELBO training step (zero_grad, elbo fwd, backward, AdamW), 3x128x128 tiles, '
                     f'latent_dim {LATENT}, num_filters [64,128,256,512], dropout 0.10 on (BASELINE.json configs[1])',
         'per_gpu_batch': per_gpu_batch, 'global_batch': per_gpu_batch * n_gpus, 'tile': TILE,
         'parallelism': f'dp{n_gpus}', 'l2_policy': 'working set >> L2 (activations of one step are tens of GB)'}
    if note:
        c['note'] = note
    return c


# ----------------------------------------------------------------------------------------------------------------------
# our arm
# ----------------------------------------------------------------------------------------------------------------------
class Ctx:
    """Process-wide state of one bench run (rank, device, timing helpers)."""

    def __init__(self):
        import torch.distributed as dist
        self.dist = dist
        self.world = int(os.environ.get('WORLD_SIZE', '1'))
        self.rank = int(os.environ.get('RANK', '0'))
        self.local_rank = int(os.environ.get('LOCAL_RANK', '0'))
        torch.cuda.set_device(self.local_rank)
        self.dev = torch.device('cuda', self.local_rank)
        if self.world > 1:
            dist.init_process_group('nccl', device_id=self.dev)

    def barrier(self):
        if self.world > 1:
            self.dist.barrier()
        torch.cuda.synchronize()

    def timed(self, fn, steps, warmup):
        """warmup untimed calls, then `steps` calls bracketed by barrier + synchronize; CUDA events on the launching
        stream; returns ms per call, MAX over ranks."""
        for _ in range(warmup):
            fn()
        s_evt, e_evt = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        self.barrier()
        s_evt.record()
        for _ in range(steps):
            fn()
        e_evt.record()
        self.barrier()
        return self.max_over_ranks(s_evt.elapsed_time(e_evt) / steps)

    def max_over_ranks(self, v):
        if self.world == 1:
            return v
        t = torch.tensor([v], device=self.dev, dtype=torch.float64)
        self.dist.all_reduce(t, op=self.dist.ReduceOp.MAX)
        return t.item()


def bench_train(ctx, args, model, opt, ddp):
    """Timed regions of the headline metric.  Returns a dict of measurements (rank-independent after max-over-ranks)."""
    from prob_unet_mds_b200 import _lib as L
    B = args.batch
    dev = ctx.dev
    x_host, t_host = make_batch(B, seed=1 + ctx.rank)
    x_pin, t_pin = x_host.pin_memory(), t_host.pin_memory()
    x_dev, t_dev = x_pin.to(dev), t_pin.to(dev)
    out = {}

    def step(x, t):
        opt.zero_grad(set_to_none=True)
        total, recon, kl = model.elbo(x, t)
        total.backward()
        opt.step()
        return total

    for _ in range(args.warmup):
        step(x_dev, t_dev)
    ctx.barrier()

    # ---- region 1: inputs resident in HBM; no per-call instrumentation
    sampler = ClockSampler(ctx.local_rank) if ctx.rank == 0 else None
    if sampler:
        sampler.start()
    L._raw_lib().pu_launch_count(1)
    ncu_range = os.environ.get('PU_NCU_RANGE') == '1'      # scripts/gpu_launch_list.sh: `ncu --profile-from-start off`
    if ncu_range:                                          # captures exactly the steady-state timed steps
        torch.cuda.profiler.start()
    out['ms_resident'] = ctx.timed(lambda: step(x_dev, t_dev), args.steps, 0)
    if ncu_range:
        torch.cuda.profiler.stop()
    out['launches'] = int(L._raw_lib().pu_launch_count(0))
    out['clocks'] = sampler.stop() if sampler else None

    # ---- region 2: end to end through the public API with host buffers (H2D of the batch, D2H of the loss)
    holder = {}

    def e2e_step():
        x = x_pin.to(dev, non_blocking=True)
        t = t_pin.to(dev, non_blocking=True)
        holder['loss'] = step(x, t).item()
    out['ms_e2e'] = ctx.timed(e2e_step, args.steps, 0)
    out['last_loss'] = holder['loss']
    out['h2d'] = int(x_pin.numel() * 4 + t_pin.numel() * 4)

    # ---- separate instrumented pass: CUDA events around every C-ABI call -> per-kernel time and the roofline of the
    # dominant kernels (not part of `value`: the events cost ~2 % of the step)
    out['kinds'], out['ms_profiled'] = {}, None
    if args.profile_calls:
        ctx.barrier()
        prof = L.start_profiling()
        psteps = min(args.steps, 3)
        s_evt, e_evt = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        s_evt.record()
        for _ in range(psteps):
            step(x_dev, t_dev)
        e_evt.record()
        torch.cuda.synchronize()
        L.stop_profiling()
        out['kinds'] = prof.summary()
        out['psteps'] = psteps
        out['ms_profiled'] = s_evt.elapsed_time(e_evt) / psteps
    return out, (x_dev, t_dev)


def bench_config3(ctx, args, model, opt, ddp):
    """BASELINE.json configs[2]: data-parallel training at GLOBAL batch 512.  Each rank owns 512/N samples, processed as
    micro-batches of 64 with local gradient accumulation (DataParallel.no_sync) and ONE bucketed NCCL all-reduce,
    overlapped with the last micro-batch's backward, per optimizer step.  At N = 1 this is 8 accumulated micro-batches."""
    G = 512
    if G % ctx.world or (G // ctx.world) % args.batch:
        return None
    micro = G // ctx.world // args.batch
    dev = ctx.dev
    batches = [tuple(u.to(dev) for u in make_batch(args.batch, seed=100 + ctx.rank * 16 + i)) for i in range(micro)]

    def step():
        opt.zero_grad(set_to_none=True)
        for i, (x, t) in enumerate(batches):
            last = i == micro - 1
            cm = contextlib.nullcontext() if (last or ddp is None) else ddp.no_sync()
            with cm:
                total, _, _ = model.elbo(x, t)
                total.backward()
        opt.step()
    steps = max(2, min(args.steps, 3 if micro > 2 else 5))
    ms = ctx.timed(step, steps, 1)
    return {'metric': 'elbo_train_samples_per_s', 'value': G / (ms / 1e3), 'unit': 'samples/s', 'ms_per_step': ms,
            'scaling': 'strong', 'steps': steps,
            'config': {'global_batch': G, 'per_gpu_batch': G // ctx.world, 'micro_batch': args.batch,
                       'micro_batches_per_step': micro, 'parallelism': f'dp{ctx.world}',
                       'exchange': 'one overlapped NCCL all-reduce(SUM) of the accumulated gradients per step'}}


def bench_ensemble(ctx, args, model, x_dev):
    """BASELINE.json configs[3]: 100 latent samples per input through prior + Fcomb only (U-Net once per input)."""
    from prob_unet_mds_b200 import parallel
    S = args.ensemble_members
    B = args.batch
    res = {}
    model.eval()
    steps = max(2, min(args.steps, 5))
    keep = {}
    with torch.no_grad():
        # (1) weak-scaling line, as in round 1: 64 inputs per GPU, all members local, result stays on the device
        ms = ctx.timed(lambda: keep.__setitem__('o', model.sample_ensemble(x_dev, S)), steps, 2)
        res['ensemble'] = {'metric': 'ensemble_member_samples_per_s', 'value': B * ctx.world * S / (ms / 1e3),
                           'unit': 'member-samples/s', 'ms_per_step': ms,
                           'config': {'inputs_per_gpu': B, 'members': S, 'tile': TILE,
                                      'sharding': 'inputs over ranks, all members local, result resident in HBM'}}
        host = torch.empty(keep['o'].shape, dtype=torch.float32).pin_memory()

        def local_e2e():
            host.copy_(model.sample_ensemble(x_dev, S), non_blocking=True)
            torch.cuda.current_stream().synchronize()
        ms = ctx.timed(local_e2e, steps, 1)
        res['ensemble']['e2e'] = {'value': B * ctx.world * S / (ms / 1e3), 'unit': 'member-samples/s', 'ms_per_step': ms,
                                  'd2h_bytes_per_step': host.numel() * 4,
                                  'note': 'result copied to pinned host memory inside the timed region (the reference '
                                          'does output.cpu() per member, train_prob_unet_model.py:181)'}
        keep.clear()
        del host
        # (2) configs[3] as written: ONE set of 64 inputs for the whole job, the 100 members sharded over the ranks
        # (inputs sharded for the encode, features + (mu, log sigma) all-gathered), beside the zero-communication variant
        # (inputs sharded, all members local) on the same 64 inputs
        if ctx.world > 1:
            xg = make_batch(B, seed=999)[0].to(ctx.dev)
            lo, hi = parallel.shard_range(B, ctx.rank, ctx.world)
            ms_m = ctx.timed(lambda: keep.__setitem__('o', parallel.ensemble_sharded(model, xg, S)[0]), steps, 2)
            hm = torch.empty(keep['o'].shape, dtype=torch.float32).pin_memory()

            def members_e2e():
                hm.copy_(parallel.ensemble_sharded(model, xg, S)[0], non_blocking=True)
                torch.cuda.current_stream().synchronize()
            ms_m_e2e = ctx.timed(members_e2e, steps, 1)
            keep.clear()
            xl = xg[lo:hi].contiguous()
            ms_i = ctx.timed(lambda: keep.__setitem__('o', model.sample_ensemble(xl, S)), steps, 2)
            hi_ = torch.empty(keep['o'].shape, dtype=torch.float32).pin_memory()

            def inputs_e2e():
                hi_.copy_(model.sample_ensemble(xl, S), non_blocking=True)
                torch.cuda.current_stream().synchronize()
            ms_i_e2e = ctx.timed(inputs_e2e, steps, 1)
            keep.clear()
            tot = B * S
            res['config4_members_sharded'] = {
                'metric': 'ensemble_member_samples_per_s', 'unit': 'member-samples/s', 'scaling': 'strong',
                'config': {'inputs_total': B, 'members': S, 'tile': TILE, 'n_gpus': ctx.world},
                'members_sharded': {'value': tot / (ms_m / 1e3), 'ms_per_step': ms_m, 'e2e_value': tot / (ms_m_e2e / 1e3),
                                    'e2e_ms_per_step': ms_m_e2e, 'd2h_bytes_per_step_per_rank': hm.numel() * 4,
                                    'exchange': 'NCCL all-gather of features [64,128,128,64] bf16 + (mu, log sigma)'},
                'inputs_sharded': {'value': tot / (ms_i / 1e3), 'ms_per_step': ms_i, 'e2e_value': tot / (ms_i_e2e / 1e3),
                                   'e2e_ms_per_step': ms_i_e2e, 'd2h_bytes_per_step_per_rank': hi_.numel() * 4,
                                   'exchange': 'none'}}
    model.train()
    return res


def bench_det_unet(ctx, args):
    """BASELINE.json configs[4]: baseline/deterministic_unet.UNet (model_channels 64, no attention) training step as
    trainmodel.train_step runs it (trainmodel.py:157-160, baseline/main.py:69): preds = model(x); MSELoss (mean);
    backward; AdamW.  Batch 32 per GPU, 256x256 tiles, bf16."""
    from prob_unet_mds_b200 import AdamW
    from prob_unet_mds_b200.baseline.deterministic_unet import UNet
    Bd, Hd = 32, 256
    dev = ctx.dev
    torch.manual_seed(4321 + ctx.rank)
    m = UNet(img_resolution=(64, 64), in_channels=3, out_channels=3, label_dim=0, use_diffuse=False).to(dev)
    m.load_state_dict(synth.make_weights(synth.load_schema('schema_detunet.json'), seed=3))
    m.compute_dtype = torch.bfloat16
    m.train()
    opt = AdamW(m.parameters(), lr=1e-3)
    x, t = [u.to(dev) for u in make_batch(Bd, seed=50 + ctx.rank, tile=Hd)]
    lossf = torch.nn.MSELoss()

    def step():
        opt.zero_grad(set_to_none=True)
        loss = lossf(m(x, class_labels=None), t)
        loss.backward()
        opt.step()
    steps = max(2, min(args.steps, 5))
    ms = ctx.timed(step, steps, 2)
    kinds = {}
    if args.profile_calls and ctx.rank == 0:
        from prob_unet_mds_b200 import _lib as L
        prof = L.start_profiling()
        for _ in range(2):
            step()
        torch.cuda.synchronize()
        L.stop_profiling()
        kinds = {k: round(v['ms'] / 2, 3) for k, v in sorted(prof.summary().items(), key=lambda kv: -kv[1]['ms'])[:10]}
    peaks = measured_peaks()
    sps = Bd * ctx.world / (ms / 1e3)
    tf = GFLOP_DET_FWD_BWD_PER_SAMPLE * sps / 1e3
    del m, opt, x, t
    torch.cuda.empty_cache()
    return {'metric': 'det_unet_train_samples_per_s', 'value': sps, 'unit': 'samples/s', 'ms_per_step': ms, 'steps': steps,
            'step_tflops_algorithmic': tf, 'step_frac_of_bf16_peak': tf / ctx.world / peaks['tflops_sustained'],
            'kernel_ms_per_step': kinds,
            'config': {'model': 'baseline/deterministic_unet.UNet (22.8 M parameters)', 'per_gpu_batch': Bd, 'tile': Hd,
                       'loss': 'MSELoss(mean)', 'optimizer': 'prob_unet_mds_b200.AdamW', 'dtype': 'bf16',
                       'gflop_per_sample_fwd_bwd': GFLOP_DET_FWD_BWD_PER_SAMPLE}}


def dp_check(ctx):
    """Driver-visible evidence that the data-parallel exchange is correct (the 2-GPU pytest is skipped on 1-GPU boxes):
    every rank runs elbo + backward on its shard (2 samples, 64x64, fp32 mode, dropout off) through DataParallel; after
    the all-reduce all ranks must hold the same gradient checksum, and rank 0 compares the summed loss and the gradients
    with a single-process run over the whole global batch."""
    from prob_unet_mds_b200 import ProbabilisticUNet, parallel
    dist = ctx.dist
    per, H, Lz = 2, 64, LATENT
    Bg = per * ctx.world
    xg, tg = synth.make_inputs(Bg, H, H, seed=21)
    eps = synth.make_eps(Bg, Lz, seed=22)

    def fresh():
        m = ProbabilisticUNet(3, 3, latent_dim=Lz).to(ctx.dev)
        m.load_state_dict(probunet_weights())
        m.set_precision('fp32')
        for b in m.unet.modules():
            if hasattr(b, 'dropout'):
                b.dropout = 0
        m.train()
        return m
    m = fresh()
    dp = parallel.DataParallel(m)
    lo = ctx.rank * per
    checks = []
    for it in range(2):                 # step 0 learns the bucket order (unbucketed reduce), step 1 uses the buckets
        for p in m.parameters():
            p.grad = None
        m.eps_override = eps[lo:lo + per]
        total, recon, kl = m.elbo(xg[lo:lo + per].to(ctx.dev), tg[lo:lo + per].to(ctx.dev))
        total.backward()
        tot_sum, = parallel.allreduce_losses(total)
        cs = torch.stack([p.grad.double().abs().sum() for p in m.parameters() if p.grad is not None]).sum()
        allcs = [torch.zeros_like(cs) for _ in range(ctx.world)]
        dist.all_gather(allcs, cs)
        checks.append(dict(loss_sum=tot_sum.item(), checksums=[c.item() for c in allcs]))
    res = {'ranks': ctx.world, 'per_rank_batch': per, 'tile': H, 'precision': 'fp32',
           'grad_checksum_identical_on_all_ranks': all(len(set(c['checksums'])) == 1 for c in checks),
           'grad_checksum': checks[-1]['checksums'][0], 'loss_sum_over_ranks': checks[-1]['loss_sum']}
    if ctx.rank == 0:
        s = fresh()
        s.eps_override = eps
        total, _, _ = s.elbo(xg.to(ctx.dev), tg.to(ctx.dev))
        total.backward()
        errs = []
        for (k, p), (_, q) in zip(m.named_parameters(), s.named_parameters()):
            if q.grad is None:
                continue
            d = (p.grad.double() - q.grad.double()).norm().item()
            n = q.grad.double().norm().item()
            if n > 0:
                errs.append(d / n)
        errs.sort()
        res.update(single_process_loss=total.item(),
                   loss_rel_err=abs(checks[-1]['loss_sum'] - total.item()) / abs(total.item()),
                   grad_rel_err_median=errs[len(errs) // 2], grad_rel_err_max=errs[-1])
        res['ok'] = bool(res['grad_checksum_identical_on_all_ranks'] and res['loss_rel_err'] < 1e-5
                         and res['grad_rel_err_max'] < 1e-3)
    m._grad_sink_factory = None
    return res


def parity_numbers(ctx):
    """Measured errors of the benchmarked precision against the committed reference fixtures (tests/golden/*.npz, written
    by the unmodified reference): ELBO / KL / posterior-branch logits at 64x64 and 128x128, batch 1, L=16, dropout off."""
    import numpy as np
    from prob_unet_mds_b200 import ProbabilisticUNet
    out = {}
    for tag, H in (('probunet_64_L16_B1', 64), ('probunet_128_L16_B1', 128)):
        fx = np.load(os.path.join(ROOT, 'tests', 'golden', tag + '.npz'))
        m = ProbabilisticUNet(3, 3, latent_dim=LATENT).to(ctx.dev)
        m.load_state_dict(probunet_weights())
        for b in m.unet.modules():
            if hasattr(b, 'dropout'):
                b.dropout = 0
        m.train()
        x, t = synth.make_inputs(1, H, H, seed=1)
        x, t = x.to(ctx.dev), t.to(ctx.dev)
        row = {}
        for prec in ('bf16', 'fp32'):
            m.set_precision(prec)
            with torch.no_grad():
                m.eps_override = torch.from_numpy(fx['eps'])
                total, recon, kl = m.elbo(x, t)
                m.eps_override = torch.from_numpy(fx['post_eps'])
                y = m(x, t, training=True).cpu().double()
            ref = torch.from_numpy(fx['post_output']).double()
            row[prec] = {'elbo_rel_err': abs(total.item() - float(fx['total'])) / abs(float(fx['total'])),
                         'kl_rel_err': abs(kl.item() - float(fx['kl'])) / abs(float(fx['kl'])),
                         'logits_rel_err': ((y - ref).norm() / ref.norm()).item()}
        row['reference_own_bf16_autocast'] = {'elbo_rel_err': float(fx['ac_total_relerr']), 'kl_rel_err': float(fx['ac_kl_relerr']),
                                              'logits_rel_err': float(fx['ac_post_output_relerr'])}
        out[tag] = row
        del m
    return out


def gpu_eager_reference(ctx, steps=3, batch=16):
    """Informational: the UNMODIFIED reference (oracle/_ref) run by PyTorch eager (cuDNN / cuBLAS, TF32 allowed) on the
    same B200, same training step, batch 16 (its T x T fp32 attention matrices do not fit much more)."""
    try:
        model = _reference_model(ctx.dev)
        if model is None:
            return {'unavailable': 'oracle/_ref is not staged'}
        torch.backends.cudnn.allow_tf32 = True
        torch.backends.cuda.matmul.allow_tf32 = True
        model.train()
        opt = torch.optim.AdamW(model.parameters(), lr=1e-3)
        x, t = [u.to(ctx.dev) for u in make_batch(batch, seed=1)]

        def step():
            opt.zero_grad()
            loss, _, _ = model.elbo(x, t)
            loss.backward()
            opt.step()
        ms = ctx.timed(step, steps, 2)
        res = {'value': batch / (ms / 1e3), 'unit': 'samples/s', 'ms_per_step': ms, 'batch': batch,
               'what': 'unmodified reference, PyTorch eager fp32 with TF32 matmul/conv, same GPU (informational)'}
        del model, opt, x, t
        torch.cuda.empty_cache()
        return res
    except Exception as e:  # noqa: BLE001
        torch.cuda.empty_cache()
        return {'unavailable': f'{type(e).__name__}: {str(e)[:200]}'}


def run_ours(args):
    from prob_unet_mds_b200 import ProbabilisticUNet, parallel

    ctx = Ctx()
    world, rank, dev = ctx.world, ctx.rank, ctx.dev
    B = args.batch
    torch.manual_seed(1234 + rank)
    model = ProbabilisticUNet(3, 3, latent_dim=LATENT, num_filters=[64, 128, 256, 512]).to(dev)
    model.load_state_dict(probunet_weights())
    model.set_precision(args.precision)
    model.train()
    if args.torch_adamw:
        opt = torch.optim.AdamW(model.parameters(), lr=1e-3, fused=True)
    else:
        from prob_unet_mds_b200 import AdamW          # one-launch multi-tensor AdamW, torch.optim.AdamW semantics
        opt = AdamW(model.parameters(), lr=1e-3)
    # gradient all-reduce (SUM) overlapped with backward; attaches itself to the model
    ddp = parallel.DataParallel(model) if world > 1 else None

    tr, (x_dev, t_dev) = bench_train(ctx, args, model, opt, ddp)
    extras = {}
    if not args.headline_only:
        c3 = bench_config3(ctx, args, model, opt, ddp)
        if c3:
            extras['config3_global_batch_512'] = c3
        if args.ensemble_members > 0:
            extras.update(bench_ensemble(ctx, args, model, x_dev))
        del x_dev, t_dev
        model._grad_sink_factory = None
        del opt, ddp
        torch.cuda.empty_cache()
        extras['config5_det_unet'] = bench_det_unet(ctx, args)
        if world > 1:
            extras['dp_check'] = dp_check(ctx)
        if rank == 0:
            extras['parity'] = parity_numbers(ctx)
        if world == 1:
            del model
            torch.cuda.empty_cache()
            extras['gpu_eager_reference'] = gpu_eager_reference(ctx)

    if rank == 0:
        peaks = measured_peaks()
        global_batch = B * world
        ms_resident, ms_e2e = tr['ms_resident'], tr['ms_e2e']
        value = global_batch / (ms_resident / 1e3)
        roof = None
        kinds = tr['kinds']
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
                    'share_of_step': ms / (tr['ms_profiled'] * tr['psteps']), 'traffic': None,
                    'timing': 'CUDA events around every C-ABI call in a separate instrumented pass of '
                              f"{tr['psteps']} steps ({tr['ms_profiled']:.1f} ms/step with the events)"}
            roof.update(ncu_traffic())
        step_tflops = GFLOP_FWD_BWD_PER_SAMPLE * global_batch / 1e3 / (ms_resident / 1e3)
        line = {
            'metric': 'elbo_train_samples_per_s', 'value': value, 'unit': 'samples/s', 'n_gpus': world,
            'steps': args.steps, 'warmup': args.warmup, 'ms_per_step': ms_resident, 'higher_is_better': True,
            'scaling': 'weak', 'vs_baseline': None, 'dtype': 'bf16' if args.precision == 'bf16' else 'f32',
            'data': 'synthetic',
            'config': dict(workload_config(world, B), optimizer='torch.optim.AdamW(fused=True)' if args.torch_adamw
                           else 'prob_unet_mds_b200.AdamW (pu_adamw_multi)'),
            'e2e': {'value': global_batch / (ms_e2e / 1e3), 'unit': 'samples/s', 'ms_per_step': ms_e2e,
                    'h2d_bytes_per_step': tr['h2d'], 'd2h_bytes_per_step': 4},
            'gpu_launches': tr['launches'],
            'step_tflops_algorithmic': step_tflops,
            'step_frac_of_bf16_peak': step_tflops / world / peaks['tflops_sustained'],
            'roofline': roof, 'clocks': tr['clocks'], 'last_loss': tr['last_loss'],
            'kernel_ms_per_step': {k: round(v['ms'] / tr.get('psteps', 1), 3)
                                   for k, v in sorted(kinds.items(), key=lambda kv: -kv[1]['ms'])},
        }
        line.update(extras)
        if world == 1 and not args.no_cpu_baseline and not args.headline_only:
            cb, _ = cpu_reference_steps(1, 1, batch=args.cpu_batch)
            line['cpu_baseline'] = cb
        print(json.dumps(line), flush=True)
    if world > 1:
        ctx.dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--gpus', type=int, default=1)
    ap.add_argument('--steps', type=int, default=5)
    ap.add_argument('--warmup', type=int, default=3)
    ap.add_argument('--impl', default='ours', choices=['ours', 'reference'])
    ap.add_argument('--batch', type=int, default=64, help='per-GPU batch')
    ap.add_argument('--precision', default='bf16', choices=['bf16', 'fp32'])
    ap.add_argument('--no-cpu-baseline', action='store_true')
    ap.add_argument('--cpu-batch', type=int, default=CPU_BATCH, help='batch of the CPU reference arm / cpu_baseline leg '
                    '(default: BASELINE.json configs[0], batch 8)')
    ap.add_argument('--headline-only', action='store_true', help='only the ELBO training metric (no extra configs)')
    ap.add_argument('--torch-adamw', action='store_true',
                    help="use torch.optim.AdamW(fused=True) instead of the package's multi-tensor AdamW")
    ap.add_argument('--ensemble-members', type=int, default=100,
                    help='members per input of the ensemble metric (0 disables it)')
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
