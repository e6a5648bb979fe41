"""Batch-sharded data parallelism for STOVE training (one process per GPU).

The reference is single-process / single-device (model/main.py:140-141); only the batch
shards naturally.  Every rank holds a full replica, takes its slice of
the batch, and the ELBO gradient -- the 110 live tensors, 1.41 M floats -- is exchanged as ONE
flat fp32 bucket in a single NCCL all-reduce per step, before gradient clipping (which needs
the global norm, model/video_prediction/train.py:471-472) and Adam.  Parameters that never
receive a gradient (the two unused dynamics cores, SURVEY hard part 10) stay out of the
bucket and keep `grad = None`, so Adam skips them exactly like the reference.
Rollouts shard over sequences with no collective at all.
"""
import torch
import torch.distributed as dist


def world():
    return dist.get_world_size() if dist.is_available() and dist.is_initialized() else 1


def rank():
    return dist.get_rank() if dist.is_available() and dist.is_initialized() else 0


def shard(t, dim=0):
    """This rank's contiguous slice of a tensor along `dim` (batch-sharding rule).  The batch must divide
    evenly: the gradient exchange averages per-rank MEAN gradients with equal weights, which equals the
    single-process full-batch gradient only for equal shards."""
    w, r = world(), rank()
    n = t.shape[dim]
    if n % w:
        raise ValueError('batch of %d does not divide over %d ranks' % (n, w))
    return t.narrow(dim, (n // w) * r, n // w)


class DataParallel:
    """Minimal DP engine around a `Stove` replica: zero_grad -> forward -> backward ->
    flat-bucket all-reduce (-> clip -> optimizer step)."""

    def __init__(self, model, reward_factor=None, broadcast=True, mse=None):
        """reward_factor / mse default to the model config's debug_reward_factor / debug_mse (train.py:205-208,
        :463).  The ramp-up weight min(1, step / debug_reward_rampup) (train.py:458-462) is a DEVICE scalar,
        `self.reward_weight`, refreshed by `set_reward_weight(step)` so that a captured graph replays it."""
        self.model = model
        self.params = [p for p in model.parameters() if p.requires_grad]
        c = getattr(model, 'c', None)
        self.reward_factor = getattr(c, 'debug_reward_factor', 0.0) if reward_factor is None else reward_factor
        self.mse = bool(getattr(c, 'debug_mse', False)) if mse is None else mse
        self.rampup = getattr(c, 'debug_reward_rampup', False)
        self.reward_weight = None
        self.flat = None
        self.live = None
        if broadcast and world() > 1:
            for p in model.parameters():
                dist.broadcast(p.data, 0)

    # -- one training iteration (optimizer excluded) ---------------------------------------
    def forward_backward(self, x, step_counter=1, actions=None, reward_target=None):
        for p in self.params:
            p.grad = None
        elbo, prop, rewards = self.model(x, step_counter, actions=actions)
        if reward_target is not None and self.reward_factor:
            # train.py:452-465: reward_target = present_rewards[:, skip:] (the caller slices), both flattened
            fn = torch.nn.functional.mse_loss if self.mse else torch.nn.functional.binary_cross_entropy
            reward_loss = fn(rewards.flatten(), reward_target.flatten())
            if self.reward_weight is None:
                self.set_reward_weight(step_counter, device=rewards.device)
            loss = -elbo + (self.reward_factor * self.reward_weight) * reward_loss
            loss.backward()
            loss = loss.detach()
        else:
            # loss = -elbo: seed the backward pass with d loss / d elbo = -1 directly (no negation kernels and
            # gradient fill between the ELBO kernel and its backward); the loss value is formed afterwards
            elbo.backward(self._minus_one(elbo))
            loss = -elbo.detach()
        self.all_reduce_gradients()
        return loss

    def set_reward_weight(self, step_counter, device=None):
        """min(1, step / debug_reward_rampup) (1 when the ramp-up is off) into the device scalar the loss reads."""
        w = 1.0 if self.rampup is False or not self.rampup else min(1.0, step_counter / self.rampup)
        if self.reward_weight is None:
            self.reward_weight = torch.full((), w, device=device, dtype=torch.float32)
        else:
            self.reward_weight.fill_(w)
        return w

    def _minus_one(self, like):
        key = (like.device, like.dtype)
        cache = self.__dict__.setdefault('_neg_one', {})
        if key not in cache:
            cache[key] = torch.full((), -1.0, device=like.device, dtype=like.dtype)
        return cache[key]

    def all_reduce_gradients(self):
        """Average the gradients over ranks through one flat bucket; afterwards every `p.grad`
        is a view into that bucket (`self.flat`)."""
        live = [p for p in self.params if p.grad is not None]
        if live[0].grad.is_cuda:
            from . import ops
            flat = ops.gather_flat([p.grad for p in live])      # one or two launches instead of a 110-way cat
        else:
            flat = torch.cat([p.grad.reshape(-1) for p in live])
        if world() > 1:
            dist.all_reduce(flat)
            flat.mul_(1.0 / world())
        at = 0
        for p in live:
            n = p.numel()
            p.grad = flat[at:at + n].view_as(p)
            at += n
        self.flat, self.live = flat, live
        return flat

    def train_step(self, optimizer, x, step_counter=1, actions=None, reward_target=None):
        """One full training iteration (train.py:438-473): forward, backward, all-reduce, then the fused
        global-norm clip + Adam(amsgrad) on the flat bucket (`stove_b200.optim.FusedAdam`)."""
        loss = self.forward_backward(x, step_counter, actions, reward_target)
        optimizer.step(self.live, self.flat)
        return loss

    def clip_and_step(self, optimizer, max_norm=1.0):
        """Global-norm clipping on the flat bucket (train.py:471-472), then the optimizer."""
        if max_norm is not None:
            total = self.flat.norm()
            self.flat.mul_(torch.clamp(max_norm / (total + 1e-6), max=1.0))
        optimizer.step()


class GraphedStep:
    """One training iteration (zero_grad -> forward -> backward -> flat-bucket all-reduce)
    captured in a CUDA graph and replayed per step.

    The reference's step issues ~27 000 ATen ops and is launch bound on a GPU (SURVEY.md section
    3.1); after fusion the step is still hundreds of short launches, so removing the per-launch
    host cost matters more than any single kernel.  Capture is possible because the forward has
    no host synchronisation (the reference's matcher has one per time step, stove.py:271-273).
    Inputs are copied into static buffers; `loss` and every `p.grad` live at fixed addresses.

    Logging steps: the graph is captured with step_counter = 1, so `prop_dict` is NOT refreshed by replays;
    a trainer that reads `prop_dict` every `print_every` steps runs those steps eagerly
    (`engine.forward_backward(x, step)` / `engine.train_step`), which shares parameters and optimizer state
    with the graph.  With a reward loss call `engine.set_reward_weight(step)` before a replay (device scalar).

    Construct it before running eager backward passes of the same model on the default stream (or drop
    their autograd graphs first): PyTorch binds each parameter's gradient accumulation to the stream of the
    first autograd forward pass that used it, and the legacy default stream cannot take part in a capture
    (`cudaErrorStreamCaptureImplicit`).  The warm-up passes here run on the capture stream for that reason.
    """

    def __init__(self, engine, example_x, example_actions=None, example_reward_target=None, warmup=3,
                 optimizer=None):
        """`optimizer` (a `stove_b200.optim.FusedAdam`): also capture the clip + Adam step, i.e. replay one
        whole training iteration; its step counter and learning rate are device scalars."""
        if optimizer is not None and warmup < 1:
            raise ValueError('GraphedStep with an optimizer needs warmup >= 1 (the optimizer state binds to the '
                             'gradient bucket of an eager pass)')
        self.engine = engine
        self.optimizer = optimizer
        self.x = example_x.clone()
        self.actions = example_actions.clone() if example_actions is not None else None
        self.target = example_reward_target.clone() if example_reward_target is not None else None
        # capture on a HIGH-priority stream: kernel nodes inherit the priority of the stream they were captured
        # on, and everything the library and the model fork off the chain (parameter-gradient kernels,
        # parameter packing) runs on default = lowest-priority streams.  Several of those kernels fill every
        # SM; with equal priorities they delay the next kernel of the chain by their whole duration.
        side = torch.cuda.Stream(priority=-1)
        side.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(side):
            for _ in range(warmup):
                engine.forward_backward(self.x, 1, self.actions, self.target)
        torch.cuda.current_stream().wait_stream(side)
        torch.cuda.synchronize()
        if optimizer is not None and optimizer.live is None:
            optimizer._bind(engine.live, engine.flat)      # state is allocated (and zeroed) outside the graph
        from . import _native
        before = _native.lib().stove_launch_count(0)
        self.graph = torch.cuda.CUDAGraph()
        with torch.cuda.graph(self.graph, stream=side):
            if optimizer is not None:
                self.loss = engine.train_step(optimizer, self.x, 1, self.actions, self.target)
            else:
                self.loss = engine.forward_backward(self.x, 1, self.actions, self.target)
        # kernels of the native library captured in (hence launched by every replay of) the graph
        self.native_launches = _native.lib().stove_launch_count(0) - before

    def load(self, x, actions=None, reward_target=None):
        """Copy the step's inputs into the graph's static buffers; returns an event after which
        the source tensors may be overwritten (used by HostPrefetcher)."""
        self.x.copy_(x, non_blocking=True)
        if actions is not None:
            self.actions.copy_(actions, non_blocking=True)
        if reward_target is not None:
            self.target.copy_(reward_target, non_blocking=True)
        ev = torch.cuda.Event()
        ev.record()
        return ev

    def run(self):
        self.graph.replay()
        return self.loss

    def __call__(self, x, actions=None, reward_target=None):
        self.load(x, actions, reward_target)
        return self.run()


class HostPrefetcher:
    """Double-buffered host->device input pipeline: the pinned-host batch of step i+1 is copied on
    a side stream while step i computes (the reference uploads synchronously inside the step,
    train.py:436).  `get(i)` returns the device batch of step i and starts the copy of step i+1."""

    def __init__(self, fetch, device):
        self.fetch, self.device = fetch, device          # fetch(i) -> pinned host tensor
        self.stream = torch.cuda.Stream(device=device)
        self.bufs, self.events, self.free, self.ready = [None, None], [None, None], [None, None], -1

    def release(self, i, event):
        """`event`: recorded once the consumer of step i no longer reads its buffer."""
        self.free[i % 2] = event

    def _start(self, i):
        host = self.fetch(i)
        k = i % 2
        if self.bufs[k] is None or self.bufs[k].shape != host.shape:
            self.bufs[k] = torch.empty(host.shape, dtype=host.dtype, device=self.device)
        if self.free[k] is not None:
            self.stream.wait_event(self.free[k])          # only the last reader of buffer k, not the whole step
        else:
            self.stream.wait_stream(torch.cuda.current_stream(self.device))
        with torch.cuda.stream(self.stream):
            self.bufs[k].copy_(host, non_blocking=True)
            ev = torch.cuda.Event()
            ev.record(self.stream)
        self.events[k] = ev
        self.ready = i

    def get(self, i):
        if self.ready != i:
            self._start(i)
        torch.cuda.current_stream(self.device).wait_event(self.events[i % 2])
        return self.bufs[i % 2]

    def prefetch(self, i):
        self._start(i)
