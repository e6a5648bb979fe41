#!/usr/bin/env python
"""Benchmark of the STOVE hot path (BASELINE.json metric).

    python bench.py --gpus N --steps K --warmup W          # this framework on N B200s
    python bench.py --impl reference --steps K --warmup W  # the reference's own code on the host CPU cores

Headline metric: training sequences/s of the sequence-ELBO forward+backward on BASELINE config 1
(billiards, 3 balls, 32x32, 8-frame window, batch 256 per GPU; optimizer excluded, gradient all-reduce
included when N > 1; weak scaling).  One JSON line is printed by rank 0.  Extra keys carry every other
BASELINE config as stated (each with `value`, `roofline` and, at N = 1, `cpu_baseline`):
  video_prediction        cfg 2: 8-frame inference + 92-frame rollout, n = 1024 per GPU at 32x32 (+ `variants`:
                          n = 300 at 32x32 and at the reference's 50x50)
  train_ac                cfg 3: action-conditioned world model with appearance + reward head, batch 512 per GPU,
                          loss = -ELBO + 15000 * rampup * BCE(reward) (train.py:452-465)
  train_multiball         cfg 4: 6 and 9 objects at 50x50, greedy matching, batch 256 per GPU
  rollout_long            cfg 5, weak: 1024 sequences x 2000 frames per GPU
  rollout_long_sharded    cfg 5 as stated: 1024 sequences in total, sharded over the N GPUs
  train_strong            (N > 1) config 1 with the GLOBAL batch fixed at 256 (strong scaling)
  train_step_with_optimizer   config 1 + clip_grad_norm + Adam(amsgrad) in the same graph

Timing: CUDA events around exactly K steps after W warm-up steps, barrier + synchronize on both sides, max
over ranks; the K-step region is repeated until at least ~0.5 s has been measured and the MEDIAN region is
reported (a 15 ms region is at the mercy of one scheduling hiccup).  Inputs rotate through a device-resident
pool of batches larger than L2, so no step finds its frames in L2.  `e2e` repeats the measurement through the
public module call with pinned HOST batches (H2D copy of the frames and D2H read of the loss inside the timed
region).  `roofline` comes from a further pass of K steps with per-kernel CUDA events (stove_profile_*);
`cpu_baseline` times the reference's stock code (oracle/_ref, staged by oracle/build_ref.py) -- or the oracle
port when that copy is absent -- on the host cores (rank 0, N = 1 only), on a bounded sample.
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

BATCH, T, O, RES = 256, 8, 3, 32
POOL = 8                      # device-resident batches: 8 x 25.2 MB > 126 MB L2
WORKLOAD = 'STOVE billiards, 3 balls, 32x32 frames, 8-frame window, fwd+bwd ELBO, batch 256 per GPU'
MIN_REGION_S = 0.5            # repeat the K-step region until this much device time has been measured

VARIANT_KW = {
    'plain': {},
    'ac': dict(action_conditioned=True, action_space=9, debug_core_appearance=True),
    'o6': dict(num_obj=6, width=50, height=50, debug_match_objects='greedy', overlap_beta=100.0, max_obj_scale=0.22),
    'o9': dict(num_obj=9, width=50, height=50, debug_match_objects='greedy', overlap_beta=100.0, max_obj_scale=0.22),
    'g50': dict(width=50, height=50),
}


# ----------------------------------------------------------------------------------------
# helpers
# ----------------------------------------------------------------------------------------
class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled every 50 ms while the timed region runs."""

    FIELDS = ('index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,'
              'clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,'
              'clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap')

    def __init__(self, device_index):
        self.rows, self.proc, self.idx = [], None, device_index

    def __enter__(self):
        try:
            self.proc = subprocess.Popen(
                ['nvidia-smi', '-i', str(self.idx), '--query-gpu=' + self.FIELDS,
                 '--format=csv,noheader,nounits', '-lms', '50'],
                stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._read, daemon=True)
            self.thread.start()
        except OSError:
            self.proc = None
        return self

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(',')])

    def __exit__(self, *a):
        if self.proc is not None:
            self.proc.terminate()
            try:
                self.proc.wait(timeout=2)
            except subprocess.TimeoutExpired:
                self.proc.kill()

    def summary(self):
        sm = sorted(float(r[1]) for r in self.rows if len(r) >= 9 and r[1].replace('.', '').isdigit())
        mx = [float(r[2]) for r in self.rows if len(r) >= 9 and r[2].replace('.', '').isdigit()]
        names = ['hw_slowdown', 'hw_thermal_slowdown', 'sw_thermal_slowdown', 'sw_power_cap']
        reasons = sorted({n for r in self.rows if len(r) >= 9 for n, v in zip(names, r[5:9]) if v == 'Active'})
        return {'sm_mhz': sm[len(sm) // 2] if sm else None, 'sm_max_mhz': max(mx) if mx else None,
                'reasons': reasons, 'samples': len(sm)}


def peaks():
    """Roofline denominators: MEASURED_PEAKS.json (driver-written) + the micro-benchmarks of this repo
    (profiles/r02_microbench.json: FP32 FMA and tcgen05 TF32 peaks measured on the box), else stated fallbacks."""
    out = {'hbm_gbs': 6650.0, 'hbm_src': 'fallback of B200_PROFILING.md (MEASURED_PEAKS.json absent)',
           'tf32_tflops': 1590.0 / 2, 'tf32_src': 'fallback: 1.59 PFLOP/s bf16 / 2',
           'fp32_tflops': 148 * 128 * 2 * 1.965e9 / 1e12, 'fp32_src': 'derived: 148 SMs x 128 lanes x 2 x 1.965 GHz'}
    path = os.path.join(ROOT, 'MEASURED_PEAKS.json')
    if os.path.exists(path):
        with open(path) as f:
            p = json.load(f)
        out.update(hbm_gbs=p['hbm_gbs'], hbm_src='MEASURED_PEAKS.json (measured copy bandwidth)')
        if 'bf16_tflops' in p:
            out.update(tf32_tflops=p['bf16_tflops'] / 2, tf32_src='MEASURED_PEAKS.json bf16_tflops / 2 (TF32 runs at half the bf16 rate)')
    path = os.path.join(ROOT, 'profiles', 'r02_microbench.json')
    if os.path.exists(path):
        with open(path) as f:
            m = json.load(f)
        if m.get('fp32_fma_tflops'):
            out.update(fp32_tflops=m['fp32_fma_tflops'], fp32_src='profiles/r02_microbench.json (FFMA micro-benchmark on a B200 of this pool)')
        if m.get('tcgen05_tf32_tflops'):
            out.update(tf32_mma_tflops=m['tcgen05_tf32_tflops'])
    return out


def make_frames(n, seed, num_obj=O, res=RES, gravity=False, frames=T):
    from stove_b200 import synth
    return synth.billiards(n, frames, num_obj, res=res, seed=seed, gravity=gravity,
                           radius=1.2 if num_obj <= 3 else 1.0)['x']


def build_variant(variant, device, seed=0):
    from stove_b200 import Stove, StoveConfig
    torch.manual_seed(seed)
    kw = dict(width=RES, height=RES, num_obj=O, action_conditioned=False, action_space=None, random_seed=7,
              device=device)
    kw.update(VARIANT_KW[variant])
    return Stove(StoveConfig(**kw)).to(device)


def build_model(device):
    return build_variant('plain', device)


def build_ac_model(device):
    m = build_variant('ac', device, seed=2)     # a seed whose random-init rollout stays finite (SURVEY hard part 13)
    with torch.no_grad():                        # tame exp(attention) of the untrained net (stated in config)
        for core in m.dyn.att_net:
            for lin in core:
                lin.weight.mul_(0.5)
    return m


def dist_setup(n_gpus):
    import torch.distributed as dist
    # keep stdout to the one JSON line: NCCL prints its version banner (and any debug output) to stdout unless told otherwise
    if 'STOVE_NCCL_DEBUG' in os.environ:
        os.environ['NCCL_DEBUG'] = os.environ['STOVE_NCCL_DEBUG']
    os.environ.setdefault('NCCL_DEBUG_FILE', '/dev/stderr')
    world = int(os.environ.get('WORLD_SIZE', '1'))
    local = int(os.environ.get('LOCAL_RANK', '0'))
    if world > 1:
        torch.cuda.set_device(local)
        dist.init_process_group('nccl', device_id=torch.device('cuda', local))
    else:
        torch.cuda.set_device(0)
    return world, int(os.environ.get('RANK', '0')), local


def timed(fn, steps, warmup, world, on_start=None, min_seconds=MIN_REGION_S, max_repeats=200):
    """W warm-up calls, then regions of exactly K timed calls of fn(i), each bracketed by barrier + synchronize;
    device time, max over ranks per region; regions repeat until `min_seconds` are covered.
    -> (median region ms, number of regions)."""
    import torch.distributed as dist
    for i in range(warmup):
        fn(i)
    regions, at, total = [], warmup, 0.0
    while True:
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()
        if on_start is not None and not regions:
            on_start()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        for i in range(steps):
            fn(at + i)
        b.record()
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        ms = torch.tensor([a.elapsed_time(b)], device='cuda')
        if world > 1:
            dist.all_reduce(ms, op=dist.ReduceOp.MAX)      # every rank sees the same number -> the same loop count
        regions.append(float(ms))
        total += regions[-1] * 1e-3
        at += steps
        if total >= min_seconds or len(regions) >= max_repeats:
            break
    regions.sort()
    return regions[len(regions) // 2], len(regions)


# ----------------------------------------------------------------------------------------
# CPU baseline: the reference's own code (oracle/_ref) or the oracle port
# ----------------------------------------------------------------------------------------
def _oracle_config(variant):
    from oracle import stove_oracle as so
    return so.default_config(**VARIANT_KW[variant])


def reference_available():
    from oracle import ref_harness as rh
    return rh.available()


def cpu_train(variant, state_dict, x, actions=None, reward_target=None, steps=3, warmup=1, threads=None,
              dtype=torch.float32, reward_factor=15000.0, reward_weight=1.0 / 20000):
    """fwd+bwd of the training loss on the host cores.  -> dict(value seq/s, kind, cores, ms_per_step)."""
    from oracle import ref_harness as rh
    from oracle import stove_oracle as so
    threads = threads or os.cpu_count()
    torch.set_num_threads(threads)
    oc = _oracle_config(variant)
    n = x.shape[0]
    x = x.to(dtype)
    actions = actions.to(dtype) if actions is not None else None
    reward_target = reward_target.to(dtype) if reward_target is not None else None
    times = []
    if rh.available():
        kind = 'reference'
        sd = {k: v.detach().cpu() for k, v in state_dict.items()}
        ref = rh.build_reference(oc, sd, dtype)
        for i in range(warmup + steps):
            t0 = time.perf_counter()
            with rh.default_dtype(dtype), rh.quiet():
                ref.zero_grad()
                elbo, _, rewards = ref(x, 1, actions=actions)
                loss = -elbo
                if reward_target is not None:      # train.py:452-465
                    loss = loss + reward_factor * reward_weight * torch.nn.functional.binary_cross_entropy(
                        rewards.flatten(), reward_target.flatten())
                loss.backward()
            if i >= warmup:
                times.append(time.perf_counter() - t0)
    else:
        kind = 'port'
        structs = so.structures(oc)
        P = {k: v.detach().to(dtype).cpu().clone().requires_grad_(True) for k, v in state_dict.items()
             if 'output_vector' not in k}
        gen = torch.Generator().manual_seed(0)
        Ov, Tv = oc.num_obj, x.shape[1]
        for i in range(warmup + steps):
            noise = [torch.randn(n, Ov, 12, 1, generator=gen).to(dtype) for _ in range(2)] + \
                    [torch.randn(n, Ov, 18, generator=gen).to(dtype) for _ in range(Tv - 2)]
            for p in P.values():
                p.grad = None
            t0 = time.perf_counter()
            elbo, _, rewards = so.stove_forward(oc, P, x, noise, actions=actions, structs=structs)
            loss = -elbo
            if reward_target is not None:
                loss = loss + reward_factor * reward_weight * torch.nn.functional.binary_cross_entropy(
                    rewards.flatten(), reward_target.flatten())
            loss.backward()
            if i >= warmup:
                times.append(time.perf_counter() - t0)
    sec = sum(times) / len(times)
    return {'value': n / sec, 'unit': 'sequences/s', 'cores': threads, 'kind': kind,
            'dtype': str(dtype).replace('torch.', ''), 'ms_per_step': sec * 1e3,
            'sample': '%d fwd+bwd steps of batch %d (after %d warm-up) of %s, torch CPU' % (
                steps, n, warmup, 'the stock reference code (oracle/_ref)' if kind == 'reference' else 'the oracle port')}


def cpu_predict(variant, state_dict, x, num, threads=None, dtype=torch.float32):
    """8-frame inference + `num`-frame rollout under no_grad on the host cores -> frames/s dict."""
    from oracle import ref_harness as rh
    from oracle import stove_oracle as so
    threads = threads or os.cpu_count()
    torch.set_num_threads(threads)
    oc = _oracle_config(variant)
    n, Tv = x.shape[0], x.shape[1]
    x = x.to(dtype)
    if rh.available():
        kind = 'reference'
        ref = rh.build_reference(oc, {k: v.detach().cpu() for k, v in state_dict.items()}, dtype)
        with torch.no_grad(), rh.default_dtype(dtype), rh.quiet():
            ref(x[:8], 0)
            t0 = time.perf_counter()
            _, prop, _ = ref(x, 0)
            ref.rollout(prop['z'][:, -1], num=num)
            dt = time.perf_counter() - t0
    else:
        kind = 'port'
        P = {k: v.detach().to(dtype).cpu() for k, v in state_dict.items()}
        gen = torch.Generator().manual_seed(0)
        noise = [torch.randn(n, oc.num_obj, 12, 1, generator=gen).to(dtype) for _ in range(2)] + \
                [torch.randn(n, oc.num_obj, 18, generator=gen).to(dtype) for _ in range(Tv - 2)]
        with torch.no_grad():
            t0 = time.perf_counter()
            _, prop, _ = so.stove_forward(oc, P, x, noise)
            so.rollout(oc, P, prop['z'][:, -1], num=num)
            dt = time.perf_counter() - t0
    return {'value': n * (Tv + num) / dt, 'unit': 'frames/s', 'cores': threads, 'kind': kind,
            'dtype': str(dtype).replace('torch.', ''),
            'sample': 'one call on %d sequences (%d-frame inference + %d-frame rollout), torch CPU' % (n, Tv, num)}


def cpu_rollout(state_dict, z_last, actions, app, num, threads=None):
    from oracle import ref_harness as rh
    from oracle import stove_oracle as so
    threads = threads or os.cpu_count()
    torch.set_num_threads(threads)
    oc = _oracle_config('ac')
    n = z_last.shape[0]
    if rh.available():
        kind = 'reference'
        ref = rh.build_reference(oc, {k: v.detach().cpu() for k, v in state_dict.items()}, torch.float32)
        with torch.no_grad(), rh.default_dtype(torch.float32), rh.quiet():
            ref.rollout(z_last[:64], num=2, actions=actions[:64], appearance=app[:64])
            t0 = time.perf_counter()
            ref.rollout(z_last, num=num, actions=actions, appearance=app)
            dt = time.perf_counter() - t0
    else:
        kind = 'port'
        P = {k: v.detach().float().cpu() for k, v in state_dict.items()}
        with torch.no_grad():
            so.rollout(oc, P, z_last[:64], 2, actions[:64], app[:64])
            t0 = time.perf_counter()
            so.rollout(oc, P, z_last, num, actions, app)
            dt = time.perf_counter() - t0
    return {'value': n * num / dt, 'unit': 'sequence-frames/s', 'cores': threads, 'kind': kind, 'dtype': 'float32',
            'sample': '%d sequences x %d steps' % (n, num)}


# ----------------------------------------------------------------------------------------
# reference arm
# ----------------------------------------------------------------------------------------
def run_reference(args):
    """The reference's own CPU implementation of the path on the host cores: the stock `Stove.forward` +
    backward from oracle/_ref (staged copy of /root/reference, oracle/build_ref.py), else the oracle port.
    fp32 is the headline (like for like with the fp32 build); the reference's default fp64 (config.py:58) and
    its default 8 threads (config.py:59) are timed beside it.  Rank 0 only."""
    if int(os.environ.get('RANK', '0')) != 0:
        return
    model = build_model('cpu')
    frames = make_frames(BATCH, 100)
    sd = model.state_dict()
    main = cpu_train('plain', sd, frames, steps=args.steps, warmup=args.warmup)
    extra = {}
    try:
        extra['fp64_all_cores'] = cpu_train('plain', sd, frames, steps=max(2, args.steps // 5), warmup=1, dtype=torch.float64)
        extra['fp32_8_threads'] = cpu_train('plain', sd, frames, steps=max(2, args.steps // 5), warmup=1, threads=8)
    except Exception as e:                       # noqa: BLE001 -- the headline number stands on its own
        extra['error'] = repr(e)
    line = {
        'impl': 'reference', 'metric': 'train_seqs_per_sec', 'value': main['value'], 'unit': 'sequences/s',
        'n_gpus': args.gpus, 'steps': args.steps, 'warmup': args.warmup, 'ms_per_step': main['ms_per_step'],
        'higher_is_better': True, 'scaling': 'weak', 'vs_baseline': None, 'dtype': 'f32', 'data': 'synthetic',
        'config': {'workload': WORKLOAD, 'global_batch': BATCH, 'parallelism': 'cpu',
                   'step': 'one full batch-256 fwd+bwd of the ELBO (zero_grad included, optimizer excluded)'},
        'cpu_baseline': main,
        'e2e': {'value': main['value'], 'unit': 'sequences/s', 'h2d_bytes_per_step': 0, 'd2h_bytes_per_step': 0},
        'gpu_launches': 0, 'other_settings': extra,
    }
    emit(line)


# ----------------------------------------------------------------------------------------
# B200 arm
# ----------------------------------------------------------------------------------------
class TrainBench:
    """One training workload: model + DP engine + graphed step over a rotating device pool."""

    def __init__(self, model, batches, world, graph=True, optimizer=False):
        from stove_b200 import dp
        self.model, self.world = model, world
        self.engine = dp.DataParallel(model)
        self.batches = batches                      # list of tuples (x, actions | None, reward_target | None) on the device
        x, a, r = batches[0]
        self.opt = None
        if optimizer:
            from stove_b200.optim import FusedAdam
            self.opt = FusedAdam(model.parameters(), lr=2e-3, amsgrad=True, max_norm=1.0)
        self.graphed = dp.GraphedStep(self.engine, x, a, r, optimizer=self.opt) if graph else None
        self.losses = []

    def step(self, i):
        x, a, r = self.batches[i % len(self.batches)]
        if self.graphed is not None:
            loss = self.graphed(x, a, r)
        elif self.opt is not None:
            loss = self.engine.train_step(self.opt, x, 1, a, r)
        else:
            loss = self.engine.forward_backward(x, 1, a, r)
        self.losses.append(loss)
        return loss

    def measure(self, steps, warmup, on_start=None):
        ms, reps = timed(self.step, steps, warmup, self.world, on_start=on_start)
        n = self.batches[0][0].shape[0]
        finite = bool(torch.isfinite(torch.stack([l.detach().float() for l in self.losses[-3:]])).all().item())
        self.losses = []
        return {'value': self.world * n * steps / (ms * 1e-3), 'unit': 'sequences/s', 'ms_per_step': ms / steps,
                'regions': reps, 'loss_finite': finite, 'batch_per_gpu': n}


def train_roofline(value_per_gpu, flop_per_seq, bytes_per_seq, pk):
    """Whole-step view: useful fp32-equivalent FLOPs and algorithmic HBM bytes per sequence."""
    return {'bound': 'fp32', 'achieved_tflops': value_per_gpu * flop_per_seq / 1e12, 'peak_tflops': pk['fp32_tflops'],
            'frac': value_per_gpu * flop_per_seq / 1e12 / pk['fp32_tflops'], 'peak_source': pk['fp32_src'],
            'hbm': {'achieved_gbs': value_per_gpu * bytes_per_seq / 1e9, 'peak_gbs': pk['hbm_gbs'],
                    'frac': value_per_gpu * bytes_per_seq / 1e9 / pk['hbm_gbs']},
            'note': 'per GPU, whole step: %.1f MFLOP (fwd+bwd, fp32-equivalent; the input GEMM of the recognition LSTM counted '
                    'ONCE per frame -- the reference computes it num_obj times, which is how SURVEY 8d arrives at 203 MFLOP per '
                    'sequence for config 1) and %.0f KB algorithmic bytes per sequence' % (flop_per_seq / 1e6, bytes_per_seq / 1e3)}


def step_flops(num_obj, res, frames=T, ac=False):
    """Useful fwd FLOPs per sequence: encoder + dynamics + scene likelihood; fwd+bwd = 3x."""
    K, H = res * res, 256
    enc = 2.0 * frames * (K * 4 * H + (num_obj - 1) * H * 4 * H + num_obj * (H * 50 + 50 * 8))
    pair, obj = 13472, 8704 + (7 * 32 if ac else 0)
    dyn = 2.0 * (frames - 2) * (num_obj * num_obj * pair + num_obj * obj)
    scene = (frames - 1 + 1) * (num_obj * (70e3 + 14e3) + 18.0 * K * 5 + 108 * 3)
    return enc + dyn + scene


def run_b200(args):
    import torch.distributed as dist
    from stove_b200 import _native as N
    from stove_b200 import dp, synth
    world, rank, local = dist_setup(args.gpus)
    dev = torch.device('cuda', local if world > 1 else 0)
    torch.backends.cudnn.allow_tf32 = False
    torch.backends.cuda.matmul.allow_tf32 = False
    lib = N.lib()
    pk = peaks()
    graph = not args.no_graph

    # ---- config 1: training, device-resident pool (value) ---------------------------------
    model = build_model(dev)
    host_pool = [make_frames(BATCH, 1000 * rank + i) for i in range(POOL)]
    # end-to-end inputs: uint8 RGB frames in pinned host memory (what a video loader holds); the device converts
    # them in bw_transform (x / 255).  The fp32-host-frame figure is reported beside it.
    host_u8 = [(h * 255).round().to(torch.uint8).pin_memory() for h in host_pool]
    host_f32 = [h.pin_memory() for h in host_pool]
    dev_pool = [(h.to(dev), None, None) for h in host_pool]
    main = TrainBench(model, dev_pool, world, graph=graph)
    engine, graphed = main.engine, main.graphed

    def make_e2e(hosts, example):
        g = dp.GraphedStep(engine, example) if graph else None
        prefetch = dp.HostPrefetcher(lambda i: hosts[i % POOL], dev)
        sink = []

        def step_e2e(i):
            # public call with HOST frames: every step copies its batch from pinned host memory (the copy of step
            # i+1 overlaps the compute of step i on a side stream) and reads its loss back (D2H)
            x = prefetch.get(i)
            if g is not None:
                ev = g.load(x)
                loss = g.run()
                prefetch.release(i, ev if ev is not None else g.done)     # frames read in place: free after the replay
            else:
                loss = engine.forward_backward(x, step_counter=1)
            prefetch.prefetch(i + 1)
            sink.append(loss.item())
        return step_e2e, g, sink

    with ClockSampler(dev.index or 0) as clocks:
        res_main = main.measure(args.steps, args.warmup, on_start=lambda: lib.stove_launch_count(1))
        launches = lib.stove_launch_count(1)
        if graphed is not None:       # replays launch the captured kernels without passing the counter
            launches = graphed.native_launches * args.steps
        e2e_fn, e2e_graph, e2e_losses = make_e2e(host_u8, host_u8[0].to(dev))
        ms_e2e, _ = timed(e2e_fn, args.steps, args.warmup, world)
        del e2e_graph
        e2e32_fn, e2e32_graph, _ = make_e2e(host_f32, host_f32[0].to(dev))
        ms_e2e32, _ = timed(e2e32_fn, args.steps, args.warmup, world)
        del e2e32_graph
    clock_summary = clocks.summary()
    value = res_main['value']
    ms_step = res_main['ms_per_step']
    e2e_value = world * BATCH * args.steps / (ms_e2e * 1e-3)

    # ---- per-kernel pass (roofline of the dominant kernel) ---------------------------------------
    def step_eager(i):
        engine.forward_backward(dev_pool[i % POOL][0], step_counter=1)

    # per-kernel times are taken with the side streams / library forks switched off: overlapped kernels
    # would each be charged the time they spend waiting for SMs
    from stove_b200 import ops
    ops.set_fork(False)
    lib.stove_profile_enable(1)
    N.profile_read()
    timed(step_eager, args.steps, 1, world, min_seconds=0.0)      # events cannot be read back from a captured graph
    lib.stove_profile_enable(0)
    ops.set_fork(True)
    recs = N.profile_read()
    per = {}
    for name, t in recs:
        per.setdefault(name, []).append(t)
    n_steps_prof = args.steps + 1
    share = {k: sum(v) / n_steps_prof for k, v in per.items()}           # ms per step per kernel
    top = max(share, key=share.get) if share else None
    roofline = None
    if top is not None:
        roofline = kernel_roofline(top, per[top], n_steps_prof, pk)
        roofline['kernel_ms_per_step'] = {k: round(v, 4) for k, v in sorted(share.items(), key=lambda kv: -kv[1])}
        roofline['native_serial_ms_over_step_ms'] = sum(share.values()) / ms_step
        roofline['whole_step'] = train_roofline(value / world, 3 * step_flops(O, RES), 142e3, pk)
        # the other tensor-core kernels of the recognition LSTM, for the record
        roofline['other_kernels'] = {k: kernel_roofline(k, per[k], n_steps_prof, pk, brief=True)
                                     for k in ('lstm_gemm_cell_fwd', 'tc3_gemm', 'dynloop_fwd', 'dynloop_bwd', 'scene_ll_fwd', 'scene_ll_bwd')
                                     if k in per and k != top}

    extra = {}

    def guarded(name, fn):
        if args.headline_only:                   # quick iterations: config 1 only (the driver never passes this flag)
            return
        try:
            extra[name] = fn()
        except Exception as e:                   # noqa: BLE001 -- an extra workload must not take the headline down
            extra[name] = {'error': repr(e)[:300]}
        torch.cuda.synchronize()

    k_small = max(3, args.steps // 4)

    # ---- config 5: long rollouts (no collective: sequences shard over ranks) -----------------------
    ac = build_ac_model(dev)
    len5 = 2000
    step_fl5 = 2.0 * (98208 + O * (4 * 9 + 7 * 32) + 32 * 32 * 2 + 32 * 16 + 16 * 8 + 8)

    def rollout_inputs(n5, seed):
        gen = torch.Generator().manual_seed(seed)
        z_last = torch.cat([0.2 + 0.3 * torch.rand(n5, O, 2, generator=gen),
                            torch.rand(n5, O, 16, generator=gen) - 0.5], -1)
        app = torch.rand(n5, O, 3, generator=gen)
        act = synth.random_actions(n5, len5, 9, seed + 4)
        return z_last, app, act

    def rollout_bench(n5, label):
        z_last, app, act = rollout_inputs(n5, 7 + rank)
        zl_d, app_d, act_d = z_last.to(dev), app.to(dev), act.to(dev)
        ms5, reps = timed(lambda i: ac.rollout(zl_d, len5, actions=act_d, appearance=app_d), max(2, args.steps // 5), 1, world)
        k5 = max(2, args.steps // 5)
        per_gpu = n5 * len5 * k5 / (ms5 * 1e-3)
        out = {'workload': label, 'metric': 'rollout_frames_per_sec', 'value': world * per_gpu,
               'unit': 'sequence-frames/s', 'ms_per_rollout': ms5 / k5, 'regions': reps, 'sequences_per_gpu': n5,
               'inputs': 'random z_last / appearance (not config-3 inference output), attention weights of the untrained '
                         'net halved so 2000 steps stay finite',
               'roofline': {'kernel': 'team_rollout', 'bound': 'fp32',
                            'achieved_tflops': per_gpu * step_fl5 / 1e12, 'peak_tflops': pk['fp32_tflops'],
                            'frac': per_gpu * step_fl5 / 1e12 / pk['fp32_tflops'], 'peak_source': pk['fp32_src'],
                            'hbm': {'achieved_gbs': per_gpu * (O * 18 * 4 + 40) / 1e9, 'peak_gbs': pk['hbm_gbs'],
                                    'frac': per_gpu * (O * 18 * 4 + 40) / 1e9 / pk['hbm_gbs']},
                            'note': 'per GPU; %d FLOP and %d algorithmic bytes per sequence-step; a latency-bound '
                                    'kernel: 2000 dependent steps' % (int(step_fl5), O * 18 * 4 + 40)}}
        return out, (z_last, act, app)

    def cfg5_weak():
        out, inputs = rollout_bench(1024, 'action-conditioned world-model rollouts, 1024 sequences x 2000 frames per GPU (weak)')
        if world == 1 and rank == 0 and not args.no_cpu_baseline:
            z_last, act, app = inputs
            out['cpu_baseline'] = cpu_rollout(ac.state_dict(), z_last, act[:, :30], app, 30)
        return out
    guarded('rollout_long', cfg5_weak)
    if 1024 % world == 0:
        guarded('rollout_long_sharded', lambda: rollout_bench(
            1024 // world, 'BASELINE config 5 as stated: 1024 sequences x 2000 frames in total, sharded over %d GPU(s)' % world)[0])

    # ---- config 2: video prediction = 8-frame inference + 92-frame rollout ---------------------------
    def predict_bench(n2, res, gravity, variant, label, cpu_n=None):
        mdl = model if variant == 'plain' else build_variant(variant, dev)
        x2h = make_frames(n2, 77 + rank, res=res, gravity=gravity)
        x2 = x2h.to(dev)

        def predict(i):
            with torch.no_grad():
                _, prop, _ = mdl(x2, 0)
                mdl.rollout(prop['z'][:, -1], num=92)

        ms2, reps = timed(predict, max(2, args.steps // 2), 2, world)
        k2 = max(2, args.steps // 2)
        per_gpu = n2 * 100 * k2 / (ms2 * 1e-3)
        fl = (step_flops(O, res) + 92 * 2.0 * 147360) / 100          # per frame
        out = {'workload': label, 'metric': 'rollout_frames_per_sec', 'value': world * per_gpu, 'unit': 'frames/s',
               'ms_per_call': ms2 / k2, 'regions': reps, 'sequences_per_gpu': n2,
               'roofline': {'bound': 'fp32', 'achieved_tflops': per_gpu * fl / 1e12, 'peak_tflops': pk['fp32_tflops'],
                            'frac': per_gpu * fl / 1e12 / pk['fp32_tflops'], 'peak_source': pk['fp32_src'],
                            'note': 'per GPU; %.2f MFLOP useful per frame (inference amortised over 100 frames); '
                                    'launch/latency bound: ~60 launches + one persistent rollout kernel per call' % (fl / 1e6)}}
        if world == 1 and rank == 0 and not args.no_cpu_baseline and cpu_n:
            out['cpu_baseline'] = cpu_predict(variant, mdl.state_dict(), x2h[:cpu_n], 92)
        return out

    def cfg2():
        out = predict_bench(1024, 32, True, 'plain', 'gravity-like frames, 3 balls, 32x32: 8-frame inference + 92-frame rollout, '
                            '1024 sequences per GPU', cpu_n=300)
        out['variants'] = {}
        for tag, (n2, res, variant) in {'n300_32x32': (300, 32, 'plain'), 'n300_50x50': (300, 50, 'g50')}.items():
            try:
                out['variants'][tag] = predict_bench(n2, res, True, variant, '%d sequences per GPU at %dx%d' % (n2, res, res))
            except Exception as e:               # noqa: BLE001
                out['variants'][tag] = {'error': repr(e)[:300]}
        return out
    guarded('video_prediction', cfg2)

    # ---- config 3: action-conditioned training with the reward loss --------------------------------
    def cfg3():
        n3 = 512
        m3 = build_variant('ac', dev, seed=3)
        batches, hosts = [], []
        for i in range(3):                       # 3 x 50 MB of frames > L2
            x = make_frames(n3, 300 + 10 * rank + i)
            a = synth.random_actions(n3, T, 9, 500 + i)
            r = (torch.rand(n3, T - 2, 1, generator=torch.Generator().manual_seed(i)) < 0.1).float()
            hosts.append((x, a, r))
            batches.append((x.to(dev), a.to(dev), r.to(dev)))
        tb = TrainBench(m3, batches, world, graph=graph)
        tb.engine.set_reward_weight(1)           # step 1 of the ramp-up: weight 1 / 20000 (train.py:458-462)
        res = tb.measure(k_small, 3)
        res.update({'workload': 'action-conditioned avoidance-style world model (9 actions, appearance, reward head), 32x32, '
                                'batch 512 per GPU, loss = -ELBO + 15000 * rampup * BCE(reward), fwd+bwd',
                    'metric': 'train_seqs_per_sec',
                    'roofline': train_roofline(res['value'] / world, 3 * step_flops(O, RES, ac=True), 142e3 + 3 * 36 * T, pk)})
        if world == 1 and rank == 0 and not args.no_cpu_baseline:
            x, a, r = hosts[0]
            res['cpu_baseline'] = cpu_train('ac', m3.state_dict(), x[:128], a[:128], r[:128], steps=2, warmup=1)
        return res
    guarded('train_ac', cfg3)

    # ---- config 4: multiball, 6 and 9 objects at 50x50 -----------------------------------------------
    def cfg4():
        out = {'workload': 'multiball billiards at 50x50, greedy matching, overlap_beta 100, max_obj_scale 0.22, batch 256 per '
                           'GPU, fwd+bwd ELBO', 'metric': 'train_seqs_per_sec', 'objects': {}}
        for variant, num_obj in (('o6', 6), ('o9', 9)):
            try:
                m4 = build_variant(variant, dev, seed=4)
                hosts = [make_frames(BATCH, 400 + 10 * rank + i, num_obj=num_obj, res=50) for i in range(3)]   # 3 x 61 MB
                tb = TrainBench(m4, [(h.to(dev), None, None) for h in hosts], world, graph=graph)
                res = tb.measure(k_small, 3)
                res['roofline'] = train_roofline(res['value'] / world, 3 * step_flops(num_obj, 50),
                                                 8 * 3 * 2500 * 4 + 2 * 4 * 3.1e6 / BATCH, pk)
                if world == 1 and rank == 0 and not args.no_cpu_baseline:
                    res['cpu_baseline'] = cpu_train(variant, m4.state_dict(), hosts[0][:64], steps=1, warmup=1)
                out['objects'][str(num_obj)] = res
                del tb
            except Exception as e:               # noqa: BLE001
                out['objects'][str(num_obj)] = {'error': repr(e)[:300]}
        six = out['objects'].get('6', {})
        out.update({k: six[k] for k in ('value', 'unit', 'ms_per_step') if k in six})
        return out
    guarded('train_multiball', cfg4)

    # ---- strong scaling of config 1: the GLOBAL batch stays 256 ------------------------------------------
    if world > 1 and BATCH % world == 0:
        def strong():
            nb = BATCH // world
            ms_model = build_model(dev)
            ms_model.load_state_dict(model.state_dict())
            pool = [(dev_pool[i][0][rank * nb:(rank + 1) * nb].contiguous(), None, None) for i in range(POOL)]
            tb = TrainBench(ms_model, pool, world, graph=graph)
            res = tb.measure(args.steps, args.warmup)
            res.update({'workload': 'config 1 with the global batch fixed at 256 (%d sequences per GPU)' % nb,
                        'metric': 'train_seqs_per_sec', 'scaling': 'strong'})
            return res
        guarded('train_strong', strong)

    # ---- whole training iteration: + global-norm clip + Adam(amsgrad) (train.py:471-473), one graph -------
    if graph:
        def with_opt():
            model_o = build_model(dev)
            model_o.load_state_dict(model.state_dict())
            tb = TrainBench(model_o, dev_pool, world, graph=True, optimizer=True)
            res = tb.measure(args.steps, args.warmup)
            res.update({'workload': WORKLOAD + ' + clip_grad_norm(1) + Adam(amsgrad) step', 'metric': 'train_seqs_per_sec'})
            return res
        guarded('train_step_with_optimizer', with_opt)

    # ---- N > 1: the sharded gradient equals the single-rank gradient of the concatenated batch -------------
    if world > 1:
        guarded('dp_gradient_check', lambda: dp_gradient_check(model, dev, world, rank))

    def finish():
        # graphs that captured NCCL work must go before the communicator; a hard exit after the flush
        # avoids teardown hangs (the measurement is complete at this point)
        sys.stdout.flush()
        if world > 1:
            torch.cuda.synchronize()
            dist.barrier()
            os._exit(0)

    if rank != 0:
        finish()
        return

    cpu = None
    if world == 1 and not args.no_cpu_baseline:
        cpu = cpu_train('plain', model.state_dict(), host_pool[0], steps=6, warmup=1)
        try:
            cpu['fp64'] = cpu_train('plain', model.state_dict(), host_pool[0], steps=2, warmup=1, dtype=torch.float64)
        except Exception as e:                   # noqa: BLE001
            cpu['fp64'] = {'error': repr(e)[:200]}

    frame_bytes = BATCH * T * 3 * RES * RES
    line = {
        'metric': 'train_seqs_per_sec', 'value': value, 'unit': 'sequences/s', 'n_gpus': world,
        'steps': args.steps, 'warmup': args.warmup, 'ms_per_step': ms_step, 'higher_is_better': True,
        'scaling': 'weak', 'vs_baseline': None, 'dtype': 'f32', 'data': 'synthetic',
        'config': {'workload': WORKLOAD, 'global_batch': BATCH * world, 'parallelism': 'dp%d' % world,
                   'l2_policy': 'inputs rotate through a %d-batch device pool (%.0f MB > 126 MB L2)'
                                % (POOL, POOL * frame_bytes * 4 / 1e6),
                   'timing': 'median of %d regions of exactly %d steps each (>= %.1f s measured in total), CUDA events, '
                             'barrier + synchronize on both sides, max over ranks' % (res_main['regions'], args.steps, MIN_REGION_S),
                   'cuda_graph': graphed is not None,
                   'e2e_input': 'uint8 RGB frames in pinned host memory, scaled by 1/255 inside bw_transform on the device; '
                                '`e2e_f32_frames` is the same with fp32 host frames (4x the bytes)',
                   'optimizer': 'excluded (metric is fwd+bwd); gradient all-reduce of the flat bucket included when N>1',
                   'dp_exchange': None if world == 1 else (
                       'two pieces overlapped with the LSTM backward; ' +
                       ('%s on a symmetric-memory bucket (start-up race, ms per 2 MB piece: %s)' % (
                           getattr(engine._overlap, 'symm_name', '?'),
                           {k: round(v, 4) for k, v in getattr(engine._overlap, 'exchange_times_ms', {}).items()})
                        if getattr(engine._overlap, 'symm', None) else 'NCCL all-reduce'))},
        'e2e': {'value': e2e_value, 'unit': 'sequences/s', 'ms_per_step': ms_e2e / args.steps,
                'h2d_bytes_per_step': frame_bytes, 'd2h_bytes_per_step': 4},
        'e2e_f32_frames': {'value': world * BATCH * args.steps / (ms_e2e32 * 1e-3), 'unit': 'sequences/s',
                           'ms_per_step': ms_e2e32 / args.steps, 'h2d_bytes_per_step': frame_bytes * 4, 'd2h_bytes_per_step': 4},
        'gpu_launches': int(launches), 'gpu_launches_per_step': int(launches) // args.steps,
        'clocks': clock_summary, 'roofline': roofline, 'cpu_baseline': cpu,
        'loss_finite': res_main['loss_finite'] and all(v == v for v in e2e_losses[-3:]),
    }
    line.update(extra)
    emit(line)
    finish()


def dp_gradient_check(model, dev, world, rank):
    """Inside the N > 1 run: all-reduced gradients of the rank-local shards == gradient of the concatenated batch
    computed by one rank (the ELBO is a batch mean, stove.py:748).  Noise is replayed so both see the same draws."""
    import torch.distributed as dist
    from stove_b200 import dp
    n_loc = 16
    full = make_frames(n_loc * world, 4242).to(dev)
    gen = torch.Generator().manual_seed(5)
    draws = [torch.randn(n_loc * world, O, 12, 1, generator=gen) for _ in range(2)] + \
            [torch.randn(n_loc * world, O, 18, generator=gen) for _ in range(T - 2)]

    class Replay:
        stacked = False

        def __init__(self, lo, hi):
            self.d, self.at = [d[lo:hi].to(dev) for d in draws], 0      # kept referenced: consumed on a side stream

        def __call__(self, shape, like):
            self.at += 1
            return self.d[self.at - 1]

    eng = dp.DataParallel(model, broadcast=False)
    shard = full[rank * n_loc:(rank + 1) * n_loc]
    flats = []
    for _ in range(3):                 # pass 1: one all-reduce of the whole bucket; passes 2, 3: overlapped pieces
        model._standard_normal = Replay(rank * n_loc, (rank + 1) * n_loc)
        eng.forward_backward(shard, 1)
        flats.append(eng.flat.clone())
    sharded = flats[-1]
    piecewise_vs_whole = float((flats[-1] - flats[0]).abs().max() / flats[0].abs().max())
    model._standard_normal = Replay(0, n_loc * world)
    for p in eng.params:
        p.grad = None
    elbo, _, _ = model(full, 1)
    (-elbo).backward()
    del model._standard_normal
    ref = torch.cat([p.grad.reshape(-1) for p in eng.live])
    err = float((sharded - ref).abs().max() / ref.abs().max())
    t = torch.tensor([err, piecewise_vs_whole], device=dev)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    err, piecewise_vs_whole = float(t[0]), float(t[1])
    return {'max_rel_err': err, 'tolerance': 1e-4, 'ok': err < 1e-4 and piecewise_vs_whole < 1e-5,
            'overlapped_vs_single_allreduce': piecewise_vs_whole, 'overlap_active': eng._overlap is not None,
            'symmetric_memory': bool(getattr(eng._overlap, 'symm', None)),
            'what': 'flat gradient after the (overlapped, piecewise) all-reduce of %d-sequence shards vs. one rank on the '
                    '%d-sequence batch' % (n_loc, n_loc * world)}



def kernel_roofline(kernel, times, n_steps, pk, brief=False):
    launches = len(times) / n_steps
    avg_ms = sum(times) / len(times)
    units = kernel_algorithmic_bytes(kernel, BATCH)
    achieved = units / (avg_ms * 1e-3) / 1e9 if units else None
    out = {'kernel': kernel, 'avg_launch_ms': avg_ms, 'launches_per_step': launches}
    flops = kernel_flops(kernel, BATCH)
    tflops = kernel_tensor_flops(kernel, BATCH)
    if tflops:
        # executed TF32 tensor-core FLOPs = 3 x the fp32-equivalent product (hi*hi + hi*lo + lo*hi)
        ach = tflops / (avg_ms * 1e-3) / 1e12
        out.update({'bound': 'tensor', 'achieved': ach, 'peak': pk['tf32_tflops'], 'unit': 'TFLOP/s', 'frac': ach / pk['tf32_tflops'],
                    'frac_useful': ach / 3 / pk['tf32_tflops'], 'peak_source': pk['tf32_src'],
                    'note': '`frac` counts the executed 3xTF32 FLOPs, `frac_useful` the fp32-equivalent product (1/3 of them)'})
        if pk.get('tf32_mma_tflops'):
            out['frac_of_measured_tcgen05_tf32_peak'] = ach / pk['tf32_mma_tflops']
    elif flops:
        ach = flops / (avg_ms * 1e-3) / 1e12
        out.update({'bound': 'fp32', 'achieved': ach, 'peak': pk['fp32_tflops'], 'unit': 'TFLOP/s', 'frac': ach / pk['fp32_tflops'],
                    'frac_useful': ach / pk['fp32_tflops'], 'peak_source': pk['fp32_src'],
                    'note': 'neither HBM nor tensor bound (SURVEY 8d: ~80 FLOP per algorithmic byte, fp32 SIMT arithmetic): `achieved` = '
                            'the reference operation count of the launch / its duration against the measured FP32-FMA peak; the HBM '
                            'fraction of the same launch is under `hbm`'})
    else:
        out.update({'bound': 'hbm', 'achieved': achieved, 'peak': pk['hbm_gbs'], 'unit': 'GB/s',
                    'frac': (achieved / pk['hbm_gbs']) if achieved else None, 'peak_source': pk['hbm_src']})
    if not brief:
        out.update({'traffic': measured_traffic(kernel), 'algorithmic_bytes_per_launch': units,
                    'hbm': {'achieved_gbs': achieved, 'peak_gbs': pk['hbm_gbs'], 'frac': (achieved / pk['hbm_gbs']) if achieved else None}})
    return out


def kernel_algorithmic_bytes(kernel, batch):
    """Minimum HBM traffic of ONE launch of `kernel` in the config-1 training step (DESIGN.md
    section 4): bytes that must be read/written if every intermediate stayed on chip."""
    frames = batch * (T - 1)              # one likelihood pass over x[:, 1:]
    patches = frames * O
    D_bg, D_obj = RES * RES, 100
    S, Z, W_DYN = T - 2, 18, 22660        # dynamics steps, state width, GNN weight floats (plain config)
    n, H = batch * T, 256
    table = {
        'spn1_fwd_leaf': frames * D_bg * 4 * 2,                   # frame + mask in, 36 floats out (negligible)
        'spn1_bwd_input': frames * D_bg * 4 * 3,                  # frame + mask in, mask gradient out
        'spn1_bwd_leafparam': frames * D_bg * 4 * 2,
        'spn2_fwd': patches * D_obj * 4 * 2 + patches * 4,
        'spn2_bwd_nodes': patches * (240 + 120) * 4,
        'spn2_bwd_input': patches * D_obj * 4 * 4,
        'spn2_bwd_leafparam': patches * D_obj * 4 * 2,
        'spn2_bwd_sumparam': patches * (12 * 30 + 6 * 21) * 4,
        'scene_fwd': frames * (D_bg * 4 * 2 + O * (16 + 2 * D_obj * 4)),
        'scene_bwd': frames * (D_bg * 4 * 2 + O * (16 + 2 * D_obj * 4)),
        # fused scene likelihood (4.2 KB per scored frame at 32x32, O = 3): frame + z in, 1 + 2 O scalars out,
        # the packed SPN tables once (object leaf / sum / root weights, background leaf table)
        'scene_ll_fwd': frames * (D_bg * 4 + O * 16 + (1 + 2 * O) * 4) + 4 * (21600 + 14400 + 600 + 3 * D_bg * 24),
        'scene_ll_bwd': frames * (D_bg * 4 + O * 16 + (1 + 2 * O) * 4 + O * 16) + 4 * (21600 + 14400 + 600 + 3 * D_bg * 24),
        # z_init, sup, sup_std, eps in; z, z_std, z_dyn, z_dyn_std, logq, trans out; weights once
        'dynloop_fwd': 4 * (batch * O * Z + batch * S * O * (12 + Z) + batch * S * (O * (2 * Z + 2 * (Z - 2)) + 2) + W_DYN),
        'dynloop_bwd': 4 * (batch * S * (O * (Z + 12 + Z + Z) + 2) + batch * S * O * 12 + batch * O * Z + W_DYN),
        'dynloop_wgrad': 4 * (batch * S * 8224 + 148 * W_DYN),     # its input IS the per-step record stream
        'bw_transform': batch * T * D_bg * 4 * 4,
        # recognition LSTM, average launch of the O per step: frames (once) + W_ih + W_hh in; h, c and the saved
        # gate activations out; gx and the TF32 operand planes are intermediates
        'lstm_gemm_cell_fwd': 4 * (n * D_bg + 4 * H * (D_bg + H) + O * n * H * 6) // O,
        # backward GEMMs, average of the 2 (O - 1) hidden-state + 2 weight-gradient launches: gate gradients in,
        # weight gradients out (the operand planes are intermediates)
        'tc3_gemm': 4 * (O * n * 4 * H + 4 * H * (D_bg + H)) // (O + 1),
        'enc_head_fwd': 4 * batch * T * O * (256 + 50 + 8),
        'enc_head_bwd_data': 4 * batch * T * O * (8 + 50 + 256),
        'enc_head_bwd_par': 4 * batch * T * O * (256 + 50 + 8),
        'lstm_cell_bwd': 4 * n * H * (4 + 2 + 2 + 4),
        'sup_prepare_fwd': 4 * batch * T * O * (8 + 4 + 6 + 6),
        'sup_prepare_bwd': 4 * batch * T * O * (8 + 4 + 6 + 6 + 6 + 8),
    }
    return table.get(kernel)


def kernel_flops(kernel, batch):
    """Useful fp32 FLOPs of one launch (2 x FMA count of the factorised GNN)."""
    S = T - 2
    step_fwd = 2 * 98208                   # one dynamics step of one sequence, O = 3, cl = 32
    frames = batch * (T - 1)
    table = {'dynloop_fwd': batch * S * step_fwd, 'dynloop_bwd': batch * S * step_fwd,
             'dynloop_wgrad': batch * S * step_fwd,
             'scene_fwd': frames * 345e3, 'scene_bwd': frames * 2 * 345e3,      # 345 kFLOP/frame
             'scene_ll_fwd': frames * 345e3, 'scene_ll_bwd': frames * 2 * 345e3}
    return table.get(kernel)


def kernel_tensor_flops(kernel, batch):
    """EXECUTED tensor-core FLOPs of one (average) launch: 3 x the fp32-equivalent product (3xTF32)."""
    n, H, K = batch * T, 256, RES * RES
    table = {'lstm_gemm_cell_fwd': 3 * 2.0 * n * 4 * H * (K + (O - 1) * H) / O,
             # (O - 1) hidden-state GEMMs n x H x 4H, the W_hh gradient 4H x H x (O - 1) n, the W_ih gradient 4H x K x n
             'tc3_gemm': 3 * 2.0 * n * 4 * H * (2 * (O - 1) * H + K) / (O + 1)}
    return table.get(kernel)


def measured_traffic(kernel):
    """DRAM bytes of one launch from the committed `ncu --set full` capture (profiles/ncu_traffic.json)."""
    path = os.path.join(ROOT, 'profiles', 'ncu_traffic.json')
    if not os.path.exists(path):
        return None
    with open(path) as f:
        t = json.load(f).get(kernel)
    return (t['dram_read_bytes'] + t['dram_write_bytes']) if t else None


_REAL_STDOUT = None


def claim_stdout():
    """Keep stdout to the ONE JSON line: NCCL prints its version banner from C straight to file descriptor 1 (and
    libraries may print warnings); everything written to fd 1 from here on goes to stderr, the result line is
    written to the original descriptor by emit()."""
    global _REAL_STDOUT
    sys.stdout.flush()
    _REAL_STDOUT = os.dup(1)
    os.dup2(2, 1)


def emit(line):
    data = (json.dumps(line) + '\n').encode()
    if _REAL_STDOUT is None:
        sys.stdout.write(data.decode())
        sys.stdout.flush()
    else:
        os.write(_REAL_STDOUT, data)


def main():
    claim_stdout()
    ap = argparse.ArgumentParser()
    ap.add_argument('--gpus', type=int, default=1)
    ap.add_argument('--steps', type=int, default=20)
    ap.add_argument('--warmup', type=int, default=5)
    ap.add_argument('--impl', default='b200', choices=['b200', 'reference'])
    ap.add_argument('--no-cpu-baseline', action='store_true')
    ap.add_argument('--no-graph', action='store_true', help='run the training step eagerly (no CUDA graph)')
    ap.add_argument('--headline-only', action='store_true', help='skip the extra workloads (configs 2-5, strong scaling)')
    ap.add_argument('--no-scene-seq', action='store_true', help='A/B: ZAll -> SceneLL -> ElboAssemble instead of SceneElbo')
    ap.add_argument('--no-dyn-stream', action='store_true', help='A/B: dynamics weights packed on the SPN packing stream')
    ap.add_argument('--option', action='append', default=[], help='A/B: library option name=value (stove_set_option)')
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == 'b200' else args.warmup
    if args.no_scene_seq:
        from stove_b200 import ops
        ops.set_scene_seq(False)
    if args.no_dyn_stream:
        from stove_b200 import ops
        ops.set_dyn_stream(False)
    for kv in args.option:
        from stove_b200 import _native
        k, v = kv.split('=')
        _native.set_option(k, int(v))
    if args.impl == 'reference':
        run_reference(args)
    else:
        run_b200(args)


if __name__ == '__main__':
    main()
