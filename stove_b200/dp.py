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

    def __init__(self, model, reward_factor=None, broadcast=True, mse=None, overlap=True, comm='auto'):
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
        self.overlap = overlap          # R > 1 on CUDA: exchange the bucket in pieces under the backward pass
        self.comm = comm                # 'nccl' | 'symm' (NVLS through symmetric memory) | 'auto' (symm when available)
        self._overlap = None
        self.flat = None
        self.live = None
        if broadcast and world() > 1:
            for p in model.parameters():
                dist.broadcast(p.data, 0)

    # -- one training iteration (optimizer excluded) ---------------------------------------
    def forward_backward(self, x, step_counter=1, actions=None, reward_target=None):
        for p in self.params:
            p.grad = None
        if self._overlap is not None:
            self._overlap.begin()
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

    # -- gradient exchange -----------------------------------------------------------------------------
    def all_reduce_gradients(self):
        """Average the gradients over ranks through one flat bucket; afterwards every `p.grad`
        is a view into that bucket (`self.flat`).

        One rank / CPU tensors / the first pass: gather everything, one all-reduce (the 1 / R of the average is
        folded into the gather).  Later CUDA passes with R > 1 exchange the bucket in pieces WHILE the backward
        pass is still running (`_Overlap`): what was already sent is skipped here."""
        live = [p for p in self.params if p.grad is not None]
        R = world()
        ov = self._overlap
        if ov is not None and ov.active:
            flat = ov.finish(live)
        else:
            if R > 1 and live[0].grad.is_cuda and self.overlap and ov is None:
                # the set of parameters that receive gradients is known now: fix the bucket layout of the
                # overlapped exchange (used from the next pass on) and use the same order for this pass
                ov = self._overlap = _Overlap(self, live)
            if ov is not None:
                live = list(ov.order)
            if live[0].grad.is_cuda:
                from . import ops
                flat = ops.gather_flat([p.grad for p in live], scale=1.0 / R)   # one or two launches, not a 110-way cat
            else:
                flat = torch.cat([p.grad.reshape(-1) for p in live])
                if R > 1:
                    flat.mul_(1.0 / R)
            if R > 1:
                dist.all_reduce(flat)
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


_ALIGN = 128      # floats: piece boundaries of the bucket (16-byte vectors x up to 32 ranks)


class _Overlap:
    """Exchange of the flat gradient bucket in TWO pieces, the first one under the tail of the backward pass
    (R > 1, CUDA).

    Bucket order: [W_ih rows | encoder rest | early].  `early` = every parameter outside the recognition network
    (SPN, GNN): their gradients are complete long before the LSTM backward ends; a post-accumulate hook counts
    them in and the last one gathers them into the bucket on the communication stream.  The recognition network
    hands its gradients over from inside its backward node (ops.LstmEncoder.grad_sink): head, W_hh and biases
    first, then W_ih in two row blocks, upper rows first, written by the last reduction of the GEMM straight into
    the bucket.  An all-reduce costs ~20 us however small it is (profiles/
    r02_timeline_dp2_v1_four_pieces.txt: four pieces made the exchange LONGER than the compute it hides
    behind), so there are exactly two: everything from the first W_ih block (the upper rows) to the end of the
    bucket goes on the wire while the second block's GEMM runs; the second block (2 MB) is the only exposed transfer.  Whatever did not arrive through a
    hook or the sink is exchanged at the end (`finish`).  Works eagerly and under CUDA-graph capture (the
    communication stream forks from / joins the producing streams through events)."""

    def __init__(self, engine, live):
        from . import ops
        self.ops = ops
        self.engine = engine
        self.R = world()
        enc = getattr(getattr(engine.model, 'sup', None), 'encoder', None)
        names = {}
        if enc is not None and hasattr(enc, 'rnn'):
            names = {'w_hh': enc.rnn.weight_hh_l0, 'b_ih': enc.rnn.bias_ih_l0, 'b_hh': enc.rnn.bias_hh_l0,
                     'w1': enc.fc1.weight, 'b1': enc.fc1.bias, 'w2': enc.fc2.weight, 'b2': enc.fc2.bias,
                     'w_ih': enc.rnn.weight_ih_l0}
        live_ids = {id(p) for p in live}
        self.by_name = {k: p for k, p in names.items() if id(p) in live_ids}
        enc_ids = {id(p) for p in self.by_name.values()}
        early = [p for p in live if id(p) not in enc_ids]
        rest = [p for k, p in self.by_name.items() if k != 'w_ih']
        last = [self.by_name['w_ih']] if 'w_ih' in self.by_name else []
        # the bucket layout (== engine.live from now on): W_ih first, so that its row blocks start on aligned offsets
        # (its producer writes them in place) and the block computed FIRST -- the upper rows -- touches the rest
        self.order = last + rest + early
        self.offset, at = {}, 0
        for p in self.order:
            self.offset[id(p)] = at
            at += p.numel()
        self.total = at
        self.early, self.n_early = early, sum(p.numel() for p in early)
        dev = live[0].device
        self.comm = torch.cuda.Stream(device=dev, priority=-1)
        self.flat = None
        self.active = False
        self.symm = None                 # (op, group name) of the NVLS all-reduce on a symmetric-memory bucket
        self.total_padded = (self.total + _ALIGN - 1) // _ALIGN * _ALIGN
        if engine.comm in ('auto', 'symm'):
            self._try_symmetric_memory(dev, required=engine.comm == 'symm')
        self.pending, self.queue, self.sent = 0, [], set()
        for p in early:
            p.register_post_accumulate_grad_hook(self._early_hook)

    def _try_symmetric_memory(self, dev, required):
        """All-reduce through the NVSwitch (NVLS): the bucket lives in symmetric memory (one persistent buffer,
        mapped by every rank and as a multicast object), and each piece is reduced by a `multimem.ld_reduce` /
        `multimem.st` kernel -- torch's symm_mem::multimem_all_reduce_ -- instead of NCCL's ring: 25 us instead
        of 62 us for the whole 5.6 MB bucket on 8 GPUs, 15 us instead of 36 us for 1.4 MB
        (profiles/r02_allreduce_bench_N8.json).  Falls back to NCCL when the fabric has no multicast."""
        def agree(ok):
            """every rank takes part, whatever happened locally: the ranks must not diverge on the collectives"""
            t = torch.tensor([1.0 if ok else 0.0], device=dev)
            dist.all_reduce(t, op=dist.ReduceOp.MIN)
            return float(t) == 1.0

        buf = group = None
        why, has_mc = None, False
        try:
            import torch.distributed._symmetric_memory as symm_mem
            group = dist.group.WORLD.group_name
            buf = symm_mem.empty(self.total_padded, dtype=torch.float32, device=dev)
            hdl = symm_mem.rendezvous(buf, group=group)
            has_mc = bool(getattr(hdl, 'multicast_ptr', 0))
        except Exception as e:                      # noqa: BLE001 -- any failure here means: use NCCL
            why = repr(e)
        if not agree(why is None):
            if required:
                raise RuntimeError('stove_b200.dp: symmetric-memory all-reduce is not available: %s' % (why or 'on another rank'))
            return
        # trial on the two kinds of ranges the exchange uses (a prefix and the remainder), against NCCL -- and a
        # start-up race between the candidates on the size of the exposed piece: the multimem kernel wins on 8 GPUs
        # (17 vs 39 us for 2 MB), the two-shot peer-memory kernel on 2 (19 vs 27 vs 29 us); every rank times all of
        # them and the slowest rank's numbers decide, so all ranks pick the same one
        cands = [(name, getattr(torch.ops.symm_mem, name, None)) for name in ('multimem_all_reduce_', 'two_shot_all_reduce_')]
        cands = [(name, f) for name, f in cands if f is not None and (has_mc or not name.startswith('multimem'))]
        if not agree(len(cands) > 0):
            return
        times, good, why = [], True, None
        try:
            cut = (self.total_padded // 2) // _ALIGN * _ALIGN
            pattern = torch.arange(self.total_padded, device=dev, dtype=torch.float32) % 251 * (1 + rank())
            ref = pattern.clone()
            dist.all_reduce(ref)
            piece = min(self.total_padded, 524288)          # the exposed transfer: half of W_ih

            def clock(fn):
                for _ in range(3):
                    fn()
                a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                torch.cuda.synchronize(dev)
                a.record()
                for _ in range(10):
                    fn()
                b.record()
                torch.cuda.synchronize(dev)
                return a.elapsed_time(b) / 10

            scratch = torch.zeros(piece, device=dev)
            times.append(clock(lambda: dist.all_reduce(scratch)))
            for name, f in cands:
                buf.copy_(pattern)
                f(buf[:cut], 'sum', group)
                f(buf[cut:], 'sum', group)
                torch.cuda.synchronize(dev)
                good = good and bool(((buf - ref).abs() <= 1e-3 * ref.abs()).all())
                buf.zero_()
                times.append(clock(lambda: f(buf[:piece], 'sum', group)))
        except Exception as e:                      # noqa: BLE001
            good, why = False, repr(e)
        if not agree(good):
            if required:
                raise RuntimeError('stove_b200.dp: trial all-reduce through symmetric memory failed: %s' % (why or 'mismatch'))
            return
        t = torch.tensor(times + [0.0] * (3 - len(times)), device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        t = t[:1 + len(cands)].tolist()
        best = min(range(len(t)), key=lambda i: t[i])
        self.exchange_times_ms = dict(zip(['nccl'] + [name for name, _ in cands], t))
        if best == 0 and not required:
            return                                   # NCCL is the fastest here
        name, op = cands[max(best, 1) - 1]
        buf.zero_()
        self.symm = (op, group)
        self.symm_name = name
        self.symm_buf = buf

    def _all_reduce(self, lo, hi):
        """on the communication stream"""
        if self.symm is not None:
            op, group = self.symm
            op(self.flat_padded[lo:hi], 'sum', group)
        else:
            dist.all_reduce(self.flat[lo:hi])

    # -- per pass ---------------------------------------------------------------------------------------
    def begin(self):
        dev = self.order[0].device
        if self.symm is not None:
            self.flat_padded = self.symm_buf            # persistent: registered with the peers once
            self.flat = self.symm_buf[:self.total]
            # the previous pass's readers (optimizer, gradient views) are ordered before this pass's writers
            self.comm.wait_stream(torch.cuda.current_stream(dev))
        else:
            self.flat = torch.empty(self.total, device=dev, dtype=torch.float32)
        self.pending, self.queue, self.sent = len(self.early), [], set()
        self.rows_done = {}
        self.have, self.reduced = [], []              # delivered / all-reduced intervals of the bucket [lo, hi)
        self.active = True
        self.ops.LstmEncoder.grad_sink = self

    def _early_hook(self, p):
        if not self.active:
            return
        # the hook runs on the stream that produced this gradient: the exchange must wait for it
        self.comm.wait_stream(torch.cuda.current_stream(p.device))
        self.pending -= 1
        if self.pending == 0:
            self._gather([(q.grad, self.offset[id(q)], q.numel()) for q in self.early], wait=False)
            if self.early:
                lo = min(self.offset[id(q)] for q in self.early)
                self._deliver(lo, lo + self.n_early)
            self.sent.update(id(q) for q in self.early)

    def __call__(self, name, tensor, row_lo=None, row_hi=None):       # ops.LstmEncoder.grad_sink
        p = self.by_name.get(name)
        if p is None or not self.active:
            return
        off = self.offset[id(p)]
        if row_lo is None:
            self.queue.append((tensor, off, p.numel()))
            self.sent.add(id(p))
        else:
            cols = p.shape[1]
            self.queue.append((tensor[row_lo:row_hi], off + row_lo * cols, (row_hi - row_lo) * cols))
            self.rows_done[id(p)] = self.rows_done.get(id(p), 0) + (row_hi - row_lo)
            if self.rows_done[id(p)] >= p.shape[0]:
                self.sent.add(id(p))

    @property
    def scale(self):
        return 1.0 / self.R

    def slot(self, name, row_lo, row_hi):
        """the rows [row_lo, row_hi) of parameter `name` inside the bucket, for a producer that writes its (already
        1 / R scaled) gradient there itself; None if this parameter is not exchanged"""
        p = self.by_name.get(name)
        if p is None or not self.active:
            return None
        off, cols = self.offset[id(p)], p.shape[1]
        return self.flat[off + row_lo * cols: off + row_hi * cols].view(row_hi - row_lo, cols)

    def delivered(self, name, row_lo, row_hi):
        """the producer has written `slot(name, row_lo, row_hi)` on the current stream"""
        p = self.by_name[name]
        off, cols = self.offset[id(p)], p.shape[1]
        self.flush()                                          # what was queued before comes first
        self.comm.wait_stream(torch.cuda.current_stream(self.flat.device))
        self._deliver(off + row_lo * cols, off + row_hi * cols)
        self.rows_done[id(p)] = self.rows_done.get(id(p), 0) + (row_hi - row_lo)
        if self.rows_done[id(p)] >= p.shape[0]:
            self.sent.add(id(p))

    @staticmethod
    def _merge(intervals):
        out = []
        for lo, hi in sorted(intervals):
            if out and lo <= out[-1][1]:
                out[-1][1] = max(out[-1][1], hi)
            else:
                out.append([lo, hi])
        return out

    def _deliver(self, lo, hi):
        self.have = self._merge(self.have + [[lo, hi]])

    def _pending(self):
        """delivered but not yet all-reduced, as maximal intervals"""
        out = []
        for lo, hi in self.have:
            at = lo
            for rlo, rhi in self.reduced:
                if rhi <= at or rlo >= hi:
                    continue
                if rlo > at:
                    out.append([at, rlo])
                at = max(at, rhi)
            if at < hi:
                out.append([at, hi])
        return out

    def flush(self, send=False):
        """gather what the sink queued into the bucket; `send`: also all-reduce what has been delivered and is not on
        the wire yet.  Piece boundaries inside the bucket stay aligned (the multimem kernels need it); the piece that
        reaches the end of the bucket takes the padding with it."""
        if self.queue:
            lo = min(q[1] for q in self.queue)
            hi = max(q[1] + q[2] for q in self.queue)
            assert sum(q[2] for q in self.queue) == hi - lo, 'sink pieces must tile a contiguous bucket range'
            self._gather(self.queue)
            self._deliver(lo, hi)
            self.queue = []
        if not send:
            return
        for lo, hi in self._pending():
            a = (lo + _ALIGN - 1) // _ALIGN * _ALIGN
            b = hi // _ALIGN * _ALIGN
            if hi == self.total:
                b = self.total_padded if self.symm is not None else self.total
            if b > a:
                with torch.cuda.stream(self.comm):
                    self._all_reduce(a, b)
                self.reduced = self._merge(self.reduced + [[a, min(b, self.total)]])

    def _gather(self, pieces, wait=True):
        dev = self.flat.device
        if wait:
            self.comm.wait_stream(torch.cuda.current_stream(dev))
        with torch.cuda.stream(self.comm):
            self.ops.gather_flat([t for t, _, _ in pieces], out=self.flat, offsets=[o for _, o, _ in pieces],
                                 scale=1.0 / self.R)
        for t, _, _ in pieces:
            t.record_stream(self.comm)

    def finish(self, live):
        """after the backward pass: send what is left, join the communication stream -> the flat bucket"""
        self.ops.LstmEncoder.grad_sink = None
        self.active = False
        left = [p for p in self.order if id(p) not in self.sent]
        cur = torch.cuda.current_stream(self.flat.device)
        left = [q for q in left if q.grad is not None]
        if left:
            self._gather([(q.grad, self.offset[id(q)], q.numel()) for q in left])
            for q in left:
                self._deliver(self.offset[id(q)], self.offset[id(q)] + q.numel())
        self.flush(send=True)
        # slivers between aligned pieces (not in the normal flow: the pieces are aligned by construction) and
        # anything that was never delivered go through NCCL, which has no alignment constraint
        self.have = [[0, self.total]]
        for lo, hi in self._pending():
            with torch.cuda.stream(self.comm):
                dist.all_reduce(self.flat[lo:hi])
        self.reduced = [[0, self.total]]
        cur.wait_stream(self.comm)
        # the bucket order is the engine's `live` order from now on
        live[:] = self.order
        return self.flat


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
                 optimizer=None, indirect=True):
        """`optimizer` (a `stove_b200.optim.FusedAdam`): also capture the clip + Adam step, i.e. replay one
        whole training iteration; its step counter and learning rate are device scalars."""
        if optimizer is not None and warmup < 1:
            raise ValueError('GraphedStep with an optimizer needs warmup >= 1 (the optimizer state binds to the '
                             'gradient bucket of an eager pass)')
        self.engine = engine
        self.optimizer = optimizer
        # Input of the captured step.  Where the frames are only read by the first kernel (bw_transform: no
        # appearances) the graph reads them THROUGH a device cell and `load` re-points the cell (8 bytes) instead of
        # copying the batch (25 MB, ~9 us per step at config 1) into a static buffer.
        c = getattr(engine.model, 'c', None)
        self.indirect = bool(indirect and c is not None and getattr(c, 'debug_bw', False)
                             and not getattr(c, 'debug_core_appearance', False)
                             and not getattr(c, 'debug_match_appearance', False)
                             and example_x.is_contiguous() and example_x.data_ptr() % 16 == 0)
        if self.indirect:
            from . import ops
            self.x = ops.IndirectFrames(example_x)
        else:
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
        """Hand the step's inputs to the graph: copies into its static buffers, or -- indirect frames -- re-points
        the input cell at `x`.  Returns an event after which the sources may be overwritten, or None when the
        frames are read in place: then they must stay untouched until `run()` has finished (`self.done`)."""
        if self.indirect:
            self.x.point_at(x)                 # `x` must stay unchanged until the replay has consumed it
        else:
            self.x.copy_(x, non_blocking=True)
        if actions is not None:
            self.actions.copy_(actions, non_blocking=True)
        if reward_target is not None:
            self.target.copy_(reward_target, non_blocking=True)
        if self.indirect:
            return None
        ev = torch.cuda.Event()
        ev.record()
        return ev

    def run(self):
        self.graph.replay()
        self.done = torch.cuda.Event()         # recorded behind the replay: its inputs are free again
        self.done.record()
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
