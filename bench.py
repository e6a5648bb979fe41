#!/usr/bin/env python
"""Benchmark of the STOVE hot path (BASELINE.json metric).

    python bench.py --gpus N --steps K --warmup W          # this framework on N B200s
    python bench.py --impl reference --steps K --warmup W  # the reference algorithm on host CPU cores

Headline metric: training sequences/s of the sequence-ELBO forward+backward on BASELINE
config 1 (billiards, 3 balls, 32x32, 8-frame window, batch 256 per GPU; optimizer excluded,
gradient all-reduce included when N > 1; weak scaling).  One JSON line is printed by rank 0.
Extra keys report the rollout workloads of the same metric family (video prediction:
8-frame inference + 92-frame rollout; MCTS-style long rollouts: 1024 x 2000 frames).

Timing: CUDA events around exactly K steps after W warm-up steps, barrier + synchronize on
both sides, max over ranks.  Inputs rotate through a device-resident pool of batches that
is larger than L2 (8 x 25 MB > 126 MB), so no step finds its frames in L2.  `e2e` repeats the
measurement through the public module call with pinned HOST batches (H2D copy of the frames and
D2H read of the loss inside the timed region).  `roofline` comes from a second pass of K steps
with per-kernel CUDA events (stove_profile_*); `cpu_baseline` times the oracle port of the
reference algorithm on the host cores (rank 0, N = 1 only).
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


# ----------------------------------------------------------------------------------------
# helpers
# ----------------------------------------------------------------------------------------
class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled every 200 ms while the timed region runs."""

    FIELDS = ('index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,'
              'clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,'
              'clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap')

    def __init__(self, device_index):
        self.rows, self.proc, self.idx = [], None, device_index

    def __enter__(self):
        try:
            self.proc = subprocess.Popen(
                ['nvidia-smi', '-i', str(self.idx), '--query-gpu=' + self.FIELDS,
                 '--format=csv,noheader,nounits', '-lms', '200'],
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


def measured_peaks():
    path = os.path.join(ROOT, 'MEASURED_PEAKS.json')
    if os.path.exists(path):
        with open(path) as f:
            p = json.load(f)
        return p['hbm_gbs'], 'MEASURED_PEAKS.json (measured copy bandwidth)'
    return 6650.0, 'fallback of B200_PROFILING.md (MEASURED_PEAKS.json absent)'


def make_frames(n, seed):
    from stove_b200 import synth
    return synth.billiards(n, T, O, res=RES, seed=seed)['x']


def build_model(device):
    from stove_b200 import Stove, StoveConfig
    torch.manual_seed(0)
    cfg = StoveConfig(width=RES, height=RES, num_obj=O, action_conditioned=False, action_space=None,
                      random_seed=7, device=device)
    return Stove(cfg).to(device)


def build_ac_model(device):
    from stove_b200 import Stove, StoveConfig
    torch.manual_seed(2)              # a seed whose random-init rollout stays finite (SURVEY hard part 13)
    cfg = StoveConfig(width=RES, height=RES, num_obj=O, action_conditioned=True, action_space=9,
                      debug_core_appearance=True, random_seed=7, device=device)
    m = Stove(cfg).to(device)
    with torch.no_grad():             # tame exp(attention) of the untrained net
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


def timed(fn, steps, warmup, world, on_start=None):
    """W warm-up + exactly K timed calls of fn(i); device time, max over ranks (ms total)."""
    import torch.distributed as dist
    for i in range(warmup):
        fn(i)
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    if on_start is not None:
        on_start()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for i in range(steps):
        fn(warmup + i)
    b.record()
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    ms = torch.tensor([a.elapsed_time(b)], device='cuda')
    if world > 1:
        dist.all_reduce(ms, op=dist.ReduceOp.MAX)
    return float(ms)


# ----------------------------------------------------------------------------------------
# CPU baseline = oracle port of the reference algorithm (oracle/stove_oracle.py)
# ----------------------------------------------------------------------------------------
def cpu_train_baseline(steps, warmup, state_dict, frames, threads=None):
    """fwd+bwd ELBO of the reference algorithm on the host cores, fp32, full batch-256 steps."""
    from oracle import stove_oracle as so
    threads = threads or os.cpu_count()
    torch.set_num_threads(threads)
    oc = so.default_config()
    structs = so.structures(oc)
    P = {k: v.detach().float().cpu().clone().requires_grad_(True) for k, v in state_dict.items()
         if 'output_vector' not in k}
    gen = torch.Generator().manual_seed(0)
    times = []
    for i in range(warmup + steps):
        x = frames[i % len(frames)]
        noise = [torch.randn(BATCH, O, 12, 1, generator=gen) for _ in range(2)] + \
                [torch.randn(BATCH, O, 18, generator=gen) for _ in range(T - 2)]
        for p in P.values():
            p.grad = None
        t0 = time.perf_counter()
        elbo, _, _ = so.stove_forward(oc, P, x, noise, structs=structs)
        (-elbo).backward()
        if i >= warmup:
            times.append(time.perf_counter() - t0)
    return BATCH * len(times) / sum(times), threads, sum(times) / len(times)


def cpu_rollout_baseline(state_dict, z_last, actions, app, num, threads=None):
    from oracle import stove_oracle as so
    torch.set_num_threads(threads or os.cpu_count())
    oc = so.default_config(action_conditioned=True, action_space=9, debug_core_appearance=True)
    P = {k: v.detach().float().cpu() for k, v in state_dict.items()}
    with torch.no_grad():
        so.rollout(oc, P, z_last[:64], 2, actions[:64], app[:64])
        t0 = time.perf_counter()
        so.rollout(oc, P, z_last, num, actions, app)
        dt = time.perf_counter() - t0
    return z_last.shape[0] * num / dt


# ----------------------------------------------------------------------------------------
# arms
# ----------------------------------------------------------------------------------------
def run_reference(args):
    """The reference's own algorithm on the host CPU (oracle port; the reference is pure Python
    and is not on the GPU box).  Rank 0 only."""
    if int(os.environ.get('RANK', '0')) != 0:
        return
    model = build_model('cpu')
    frames = [make_frames(BATCH, 100 + i) for i in range(2)]
    value, threads, sec = cpu_train_baseline(args.steps, args.warmup, model.state_dict(), frames)
    line = {
        'impl': 'reference', 'metric': 'train_seqs_per_sec', 'value': value, 'unit': 'sequences/s',
        'n_gpus': args.gpus, 'steps': args.steps, 'warmup': args.warmup, 'ms_per_step': sec * 1e3,
        'higher_is_better': True, 'scaling': 'weak', 'vs_baseline': None, 'dtype': 'f32', 'data': 'synthetic',
        'config': {'workload': WORKLOAD, 'batch_per_step': BATCH},
        'cpu_baseline': {'value': value, 'unit': 'sequences/s', 'cores': threads, 'kind': 'port',
                         'sample': '%d full batch-256 fwd+bwd steps of the oracle port (torch CPU, fp32)' % args.steps},
        'e2e': {'value': value, 'unit': 'sequences/s', 'h2d_bytes_per_step': 0, 'd2h_bytes_per_step': 0},
        'gpu_launches': 0,
    }
    emit(line)


def run_b200(args):
    import torch.distributed as dist
    from stove_b200 import _native as N
    from stove_b200 import dp
    world, rank, local = dist_setup(args.gpus)
    dev = torch.device('cuda', local if world > 1 else 0)
    torch.backends.cudnn.allow_tf32 = False
    torch.backends.cuda.matmul.allow_tf32 = False
    model = build_model(dev)
    engine = dp.DataParallel(model)
    lib = N.lib()

    # ---- training: device-resident pool (value) -------------------------------------------
    host_pool = [make_frames(BATCH, 1000 * rank + i).pin_memory() for i in range(POOL)]
    dev_pool = [h.to(dev) for h in host_pool]
    losses = []
    graphed = None
    if not args.no_graph:
        graphed = dp.GraphedStep(engine, dev_pool[0])

    def run_step(x):
        return graphed(x) if graphed is not None else engine.forward_backward(x, step_counter=1)

    def step_resident(i):
        losses.append(run_step(dev_pool[i % POOL]))

    prefetch = dp.HostPrefetcher(lambda i: host_pool[i % POOL], dev)

    def step_e2e(i):
        # public call with HOST frames: every step copies its batch from pinned host memory (the copy
        # of step i+1 overlaps the compute of step i on a side stream) and reads its loss back (D2H)
        x = prefetch.get(i)
        if graphed is not None:
            prefetch.release(i, graphed.load(x))
            loss = graphed.run()
        else:
            loss = engine.forward_backward(x, step_counter=1)
        prefetch.prefetch(i + 1)
        losses.append(loss.item())

    with ClockSampler(dev.index or 0) as clocks:
        ms = timed(step_resident, args.steps, args.warmup, world, on_start=lambda: lib.stove_launch_count(1))
        launches = lib.stove_launch_count(1)
        if graphed is not None:       # replays launch the captured kernels without passing the counter
            launches = graphed.native_launches * args.steps
        ms_e2e = timed(step_e2e, args.steps, args.warmup, world)
    clock_summary = clocks.summary()
    value = world * BATCH * args.steps / (ms * 1e-3)
    e2e_value = world * BATCH * args.steps / (ms_e2e * 1e-3)

    # ---- per-kernel pass (roofline) ----------------------------------------------------------
    def step_eager(i):
        engine.forward_backward(dev_pool[i % POOL], step_counter=1)

    # per-kernel times are taken with the side streams / library forks switched off (STOVE_NO_FORK):
    # overlapped kernels would each be charged the time they spend waiting for SMs
    os.environ['STOVE_NO_FORK'] = '1'
    lib.stove_profile_enable(1)
    N.profile_read()
    timed(step_eager, args.steps, 1, world)      # events cannot be read back from a captured graph
    lib.stove_profile_enable(0)
    os.environ.pop('STOVE_NO_FORK', None)
    recs = N.profile_read()
    per = {}
    for name, t in recs:
        per.setdefault(name, []).append(t)
    n_steps_prof = args.steps + 1
    share = {k: sum(v) / n_steps_prof for k, v in per.items()}           # ms per step per kernel
    top = max(share, key=share.get) if share else None
    peak, peak_src = measured_peaks()
    roofline = None
    if top is not None:
        launches_top = len(per[top]) / n_steps_prof
        avg_ms = sum(per[top]) / len(per[top])
        units = kernel_algorithmic_bytes(top, BATCH)
        achieved = units / (avg_ms * 1e-3) / 1e9 if units else None
        flops = kernel_flops(top, BATCH)
        sm_mhz = clock_summary.get('sm_max_mhz') or 1965.0
        fp32_peak = 148 * 128 * 2 * sm_mhz * 1e6 / 1e12
        fp32 = None
        if flops:
            fp32 = {'achieved_tflops': flops / (avg_ms * 1e-3) / 1e12, 'peak_tflops': fp32_peak,
                    'frac': flops / (avg_ms * 1e-3) / 1e12 / fp32_peak,
                    'peak_source': '148 SMs x 128 FP32 lanes x 2 x max SM clock'}
        tensor = None
        tflops = kernel_tensor_flops(top, BATCH)
        if tflops:
            # executed TF32 tensor-core FLOPs (3 x the fp32-equivalent product: hi*hi + hi*lo + lo*hi); the TF32
            # dense peak is half the measured bf16 one
            tpeak, tsrc = measured_tensor_peak()
            tensor = {'achieved_tflops': tflops / (avg_ms * 1e-3) / 1e12, 'peak_tflops': tpeak,
                      'frac': tflops / (avg_ms * 1e-3) / 1e12 / tpeak, 'peak_source': tsrc,
                      'fp32_equivalent_tflops': tflops / 3 / (avg_ms * 1e-3) / 1e12}
        roofline = {'kernel': top, 'bound': 'hbm', 'achieved': achieved, 'peak': peak, 'unit': 'GB/s',
                    'frac': (achieved / peak) if achieved else None, 'traffic': measured_traffic(top),
                    'fp32': fp32, 'tensor': tensor,
                    'avg_launch_ms': avg_ms, 'launches_per_step': launches_top,
                    'algorithmic_bytes_per_launch': units, 'peak_source': peak_src,
                    'kernel_ms_per_step': {k: round(v, 4) for k, v in sorted(share.items(), key=lambda kv: -kv[1])},
                    'native_serial_ms_over_step_ms': sum(share.values()) / (ms / args.steps),
                    'hbm': {'achieved_gbs': achieved, 'peak_gbs': peak, 'frac': (achieved / peak) if achieved else None},
                    'note': 'the SIMT hot-path kernels are FP32-issue/latency bound, not HBM bound (DESIGN.md): for them '
                            'the HBM fraction is reported because the contract asks for it and `fp32` is the bound '
                            'that applies; the recognition-LSTM kernel (tcgen05, 3xTF32) is reported against the '
                            'tensor pipe; `hbm` always carries the algorithmic-bytes view; `traffic` = DRAM bytes '
                            'of one launch from the committed ncu capture (profiles/ncu_traffic.json)'}

    if roofline is not None and roofline.get('tensor'):
        # the dominant kernel is the tcgen05 GEMM + LSTM cell: its bound is the tensor pipe
        t = roofline['tensor']
        roofline.update({'bound': 'tensor', 'achieved': t['achieved_tflops'], 'peak': t['peak_tflops'],
                         'unit': 'TFLOP/s', 'frac': t['frac'], 'peak_source': t['peak_source']})

    # ---- rollouts (no collective: sequences shard over ranks) ----------------------------------
    extra = {}
    ac = build_ac_model(dev)
    from stove_b200 import synth
    n5, len5 = 1024, 2000
    gen = torch.Generator().manual_seed(7 + rank)
    z_last = torch.cat([0.2 + 0.3 * torch.rand(n5, O, 2, generator=gen),
                        torch.rand(n5, O, 16, generator=gen) - 0.5], -1)
    app = torch.rand(n5, O, 3, generator=gen)
    act = synth.random_actions(n5, len5, 9, 11 + rank)
    zl_d, app_d, act_d = z_last.to(dev), app.to(dev), act.to(dev)

    def roll5(i):
        ac.rollout(zl_d, len5, actions=act_d, appearance=app_d)

    ms5 = timed(roll5, max(2, args.steps // 5), 1, world)
    k5 = max(2, args.steps // 5)
    extra['rollout_long'] = {'workload': 'action-conditioned world-model rollouts, 1024 sequences x 2000 frames per GPU',
                             'metric': 'rollout_frames_per_sec', 'value': world * n5 * len5 * k5 / (ms5 * 1e-3),
                             'unit': 'sequence-frames/s', 'ms_per_rollout': ms5 / k5}
    # roofline of the persistent rollout kernel (one launch per call): useful FP32 work of one sequence-step of
    # the factorised, action-conditioned network against the FP32 pipe; the algorithmic HBM traffic is the state
    # written per step (+ the action read)
    step_flops = 2.0 * (98208 + O * (4 * 9 + 7 * 32) + 32 * 32 * 2 + 32 * 16 + 16 * 8 + 8)
    sm_mhz5 = clock_summary.get('sm_max_mhz') or 1965.0
    fp32_peak5 = 148 * 128 * 2 * sm_mhz5 * 1e6 / 1e12
    per_gpu = n5 * len5 * k5 / (ms5 * 1e-3)
    extra['rollout_long']['roofline'] = {
        'kernel': 'team_rollout', 'bound': 'fp32',
        'achieved_tflops': per_gpu * step_flops / 1e12, 'peak_tflops': fp32_peak5,
        'frac': per_gpu * step_flops / 1e12 / fp32_peak5,
        'hbm': {'achieved_gbs': per_gpu * (O * 18 * 4 + 4 + 36) / 1e9, 'peak_gbs': peak,
                'frac': per_gpu * (O * 18 * 4 + 4 + 36) / 1e9 / peak},
        'note': 'per GPU; %d FLOP and %d algorithmic bytes per sequence-step' % (int(step_flops), O * 18 * 4 + 40)}
    # video prediction: 8-frame inference + 92-frame rollout (BASELINE configs[1])
    n2 = 1024
    x2 = make_frames(n2, 77 + rank).to(dev)

    def predict(i):
        with torch.no_grad():
            _, prop, _ = model(x2, 0)
            model.rollout(prop['z'][:, -1], num=92)

    ms2 = timed(predict, max(2, args.steps // 2), 2, world)
    k2 = max(2, args.steps // 2)
    extra['video_prediction'] = {'workload': '8-frame inference + 92-frame rollout, 1024 sequences per GPU, 32x32',
                                 'metric': 'rollout_frames_per_sec', 'value': world * n2 * 100 * k2 / (ms2 * 1e-3),
                                 'unit': 'frames/s', 'ms_per_call': ms2 / k2}

    # whole training iteration: + global-norm clip + Adam(amsgrad) on the flat bucket (train.py:471-473), one graph
    if not args.no_graph:
        from stove_b200.optim import FusedAdam
        model_o = build_model(dev)
        model_o.load_state_dict(model.state_dict())
        engine_o = dp.DataParallel(model_o, broadcast=False)
        opt = FusedAdam(model_o.parameters(), lr=2e-3, amsgrad=True, max_norm=1.0)
        graphed_o = dp.GraphedStep(engine_o, dev_pool[0], optimizer=opt)
        ms_o = timed(lambda i: graphed_o(dev_pool[i % POOL]), args.steps, args.warmup, world)
        extra['train_step_with_optimizer'] = {
            'workload': WORKLOAD + ' + clip_grad_norm(1) + Adam(amsgrad) step',
            'metric': 'train_seqs_per_sec', 'value': world * BATCH * args.steps / (ms_o * 1e-3), 'unit': 'sequences/s',
            'ms_per_step': ms_o / args.steps, 'loss_finite': bool(torch.isfinite(graphed_o.loss).item())}
        del graphed_o

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
        frames = [h.clone() for h in host_pool[:2]]
        v, threads, sec = cpu_train_baseline(3, 1, model.state_dict(), frames)
        cpu = {'value': v, 'unit': 'sequences/s', 'cores': threads, 'kind': 'port',
               'sample': '3 full batch-256 fwd+bwd steps (after 1 warm-up) of the oracle port, torch CPU fp32',
               'ms_per_step': sec * 1e3}
        v5 = cpu_rollout_baseline(ac.state_dict(), z_last, act, app, 30)
        extra['rollout_long']['cpu_baseline'] = {'value': v5, 'unit': 'sequence-frames/s', 'cores': threads,
                                                 'kind': 'port', 'sample': '1024 sequences x 30 steps'}

    line = {
        'metric': 'train_seqs_per_sec', 'value': value, 'unit': 'sequences/s', 'n_gpus': world,
        'steps': args.steps, 'warmup': args.warmup, 'ms_per_step': ms / args.steps, 'higher_is_better': True,
        'scaling': 'weak', 'vs_baseline': None, 'dtype': 'f32', 'data': 'synthetic',
        'config': {'workload': WORKLOAD, 'global_batch': BATCH * world, 'parallelism': 'dp%d' % world,
                   'l2_policy': 'inputs rotate through a %d-batch device pool (%.0f MB > 126 MB L2)'
                                % (POOL, POOL * BATCH * T * 3 * RES * RES * 4 / 1e6),
                   'cuda_graph': graphed is not None,
                   'optimizer': 'excluded (metric is fwd+bwd); flat-bucket NCCL all-reduce included when N>1'},
        'e2e': {'value': e2e_value, 'unit': 'sequences/s', 'ms_per_step': ms_e2e / args.steps,
                'h2d_bytes_per_step': BATCH * T * 3 * RES * RES * 4, 'd2h_bytes_per_step': 4},
        'gpu_launches': int(launches), 'gpu_launches_per_step': int(launches) // args.steps,
        'clocks': clock_summary, 'roofline': roofline, 'cpu_baseline': cpu,
        'loss_finite': bool(all(map(lambda v: v == v, [float(l) for l in losses[-3:]]))),
    }
    line.update(extra)
    emit(line)
    finish()


def kernel_algorithmic_bytes(kernel, batch):
    """Minimum HBM traffic of ONE launch of `kernel` in the config-1 training step (DESIGN.md
    section 4): bytes that must be read/written if every intermediate stayed on chip."""
    frames = batch * (T - 1)              # one likelihood pass over x[:, 1:]
    patches = frames * O
    D_bg, D_obj = RES * RES, 100
    S, Z, W_DYN = T - 2, 18, 22660        # dynamics steps, state width, GNN weight floats (plain config)
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
        # z_init, sup, sup_std, eps in; z, z_std, z_dyn, z_dyn_std, logq, trans out; weights once
        'dynloop_fwd': 4 * (batch * O * Z + batch * S * O * (12 + Z) + batch * S * (O * (2 * Z + 2 * (Z - 2)) + 2) + W_DYN),
        # z, sup, sup_std, eps, g_z, g_logq, g_trans in; g_sup, g_sup_std, g_z_init out; weights once
        # (the per-step activation / gradient records are intermediates: 16.7 KB + 16.2 KB per
        # sequence-step, L2-resident, they show up in `traffic` only)
        'dynloop_bwd': 4 * (batch * S * (O * (Z + 12 + Z + Z) + 2) + batch * S * O * 12 + batch * O * Z + W_DYN),
        'dynloop_wgrad': 4 * (batch * S * 8224 + 148 * W_DYN),     # its input IS the per-step record stream
        'bw_transform': batch * T * D_bg * 4 * 4,
        # recognition LSTM, average launch of the O per step: frames (once) + W_ih + W_hh in; h, c and the saved
        # gate activations out; gx and the TF32 operand splits are intermediates
        'lstm_gemm_cell_fwd': 4 * (batch * T * D_bg + 4 * 256 * (D_bg + 256) + O * batch * T * 256 * 6) // O,
        'enc_head_fwd': 4 * batch * T * O * (256 + 50 + 8),
        'enc_head_bwd_data': 4 * batch * T * O * (8 + 50 + 256),
        'enc_head_bwd_par': 4 * batch * T * O * (256 + 50 + 8),
        'lstm_cell_bwd': 4 * batch * T * 256 * (4 + 2 + 2 + 4),
        'sup_prepare_fwd': 4 * batch * T * O * (8 + 4 + 6 + 6),
        'sup_prepare_bwd': 4 * batch * T * O * (8 + 4 + 6 + 6 + 6 + 8),
    }
    return table.get(kernel)


def kernel_flops(kernel, batch):
    """Useful fp32 FLOPs of one launch (2 x FMA count of the factorised GNN)."""
    S = T - 2
    step_fwd = 2 * 98208                   # one dynamics step of one sequence, O = 3, cl = 32
    # dynloop_bwd reloads the activations kept by the forward pass: input gradients only
    frames = batch * (T - 1)
    table = {'dynloop_fwd': batch * S * step_fwd, 'dynloop_bwd': batch * S * step_fwd,
             'dynloop_wgrad': batch * S * step_fwd,
             'scene_fwd': frames * 345e3, 'scene_bwd': frames * 2 * 345e3}      # 345 kFLOP/frame
    return table.get(kernel)


def kernel_tensor_flops(kernel, batch):
    """Executed tensor-core FLOPs of one (average) launch: the recognition LSTM runs 3xTF32, K-concatenated."""
    n, H, K = batch * T, 256, RES * RES
    table = {'lstm_gemm_cell_fwd': 3 * 2.0 * n * 4 * H * (K + (O - 1) * H) / O}
    return table.get(kernel)


def measured_tensor_peak():
    path = os.path.join(ROOT, 'MEASURED_PEAKS.json')
    if os.path.exists(path):
        with open(path) as f:
            p = json.load(f)
        for key in ('bf16_tflops', 'bf16_dense_tflops', 'tensor_bf16_tflops'):
            if key in p:
                return p[key] / 2.0, 'MEASURED_PEAKS.json %s / 2 (TF32 runs at half the bf16 rate)' % key
    return 2250.0 / 2 / 2, 'fallback: nominal 1125 TFLOP/s dense TF32 / 2 (MEASURED_PEAKS.json has no bf16 figure)'


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
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == 'b200' else args.warmup
    if args.impl == 'reference':
        run_reference(args)
    else:
        run_b200(args)


if __name__ == '__main__':
    main()
