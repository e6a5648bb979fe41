"""RAT-SPN module with fused sm_100a kernels (drop-in for model/spn/rat_torch.py:21-401).

Module tree and parameter names are the reference's (`vector_list.{layer}.{idx}.{means|
sigma_params|params}`, root sum also reachable as `output_vector`, rat_torch.py:271-331) so
reference checkpoints load unchanged.  `RatSpn.forward` does not walk the node vectors: it
packs the parameters once (ops.PackLeaf / ops.PackSum) and calls one fused kernel family,
chosen from the structure:
  * "D2"  leaves -> products -> sums -> products -> root   (object SPN, probabilistic_models.py:8-22)
  * "D1"  leaves -> products -> root                        (background SPN, :25-39)
Other region graphs raise: there is no generic or CPU fallback on this path.
The per-vector `forward` methods exist for the visualisation helpers of the reference
(`compute_activations(get_sum_child_acts=True)`, supair.py:371-416); they are plain tensor
code and are not used by `RatSpn.forward`.
"""
import ctypes as C
import math

import numpy as np
import torch
import torch.nn as nn

from .. import _native as N
from .. import ops


def truncated_normal_(tensor, mean=0, std=0.1):
    """N(mean, std) truncated at two standard deviations (rat_torch.py:11-18)."""
    with torch.no_grad():
        draws = tensor.new_empty(tuple(tensor.shape) + (4,)).normal_()
        ok = (draws < 2) & (draws > -2)
        pick = ok.max(-1, keepdim=True)[1]
        tensor.copy_(draws.gather(-1, pick).squeeze(-1))
        tensor.mul_(std).add_(mean)


class BasicParamProvider:
    def grab_sum_parameters(self, num_inputs, num_sums):
        return nn.Parameter(torch.empty(num_inputs, num_sums))

    def grab_leaf_parameters(self, scope, number, name=None):
        return nn.Parameter(torch.empty(len(scope), number))


class SpnArgs(object):
    def __init__(self):
        self.linear_sum_weights = False
        self.normalized_sums = True
        self.num_sums = 20
        self.num_gauss = 20
        self.param_provider = BasicParamProvider()
        self.gauss_min_sigma = 0.1
        self.gauss_max_sigma = 1.0
        self.gauss_mean_of_means = 0.0
        self.dist = 'Gauss'
        self.init_fn = truncated_normal_
        self.gauss_min_mean = None
        self.gauss_max_mean = None


class NodeVector(nn.Module):
    def __init__(self, name):
        super().__init__()
        self.name = name

    def __hash__(self):
        return hash(self.name)

    def __eq__(self, other):
        return self.name == other.name

    def init_params(self, init_fn=None):
        pass

    def num_params(self):
        return 0


_HALF_LOG_2PI = 0.5 * math.log(2 * math.pi)


class GaussVector(NodeVector):
    """Leaf vector: `size` diagonal Gaussians over the pixels in `scope`."""

    def __init__(self, region, args, name, num_dims=0):
        super().__init__(name)
        self.args = args
        self.scope = sorted(int(p) for p in region)
        self.local_size = len(self.scope)
        self.size = args.num_gauss
        self.num_dims = num_dims
        self.means = args.param_provider.grab_leaf_parameters(self.scope, self.size)
        if args.gauss_min_sigma < args.gauss_max_sigma:
            self.sigma_params = args.param_provider.grab_leaf_parameters(self.scope, self.size)
        else:
            self.sigma_params = None

    def variance(self):
        a = self.args
        return a.gauss_min_sigma + (a.gauss_max_sigma - a.gauss_min_sigma) * torch.sigmoid(self.sigma_params)

    def forward(self, inputs, marginalized=None):
        var = self.variance()
        x = inputs[:, self.scope].unsqueeze(-1)
        ll = -(x - self.means) ** 2 / (2 * var) - 0.5 * torch.log(var) - _HALF_LOG_2PI
        if marginalized is not None:
            ll = ll * (1 - marginalized.clamp(0.0, 1.0)[:, self.scope].unsqueeze(-1))
        return ll.sum(1)

    def init_params(self, init_fn=None):
        init_fn = init_fn or truncated_normal_
        init_fn(self.means, mean=self.args.gauss_mean_of_means, std=0.1)
        if self.sigma_params is not None:
            init_fn(self.sigma_params, mean=0.0, std=0.1)

    def num_params(self):
        return self.means.numel() + (self.sigma_params.numel() if self.sigma_params is not None else 0)

    def reconstruct(self, idxs, node_num, sample):
        out = np.zeros((self.num_dims,))
        mu = self.means[:, node_num].detach().cpu().numpy()
        if sample:
            mu = np.random.normal(mu, np.sqrt(self.variance()[:, node_num].detach().cpu().numpy()))
        out[self.scope] = mu
        return out


class ProductVector(NodeVector):
    """Cross product of two vectors with disjoint scopes: out[b, j*n0 + i] = in0[b, i] + in1[b, j]."""

    def __init__(self, vector1, vector2, name):
        super().__init__(name)
        self.inputs = [vector1, vector2]
        assert not set(vector1.scope) & set(vector2.scope)
        self.scope = sorted(set(vector1.scope) | set(vector2.scope))
        self.size = vector1.size * vector2.size

    def forward(self, inputs):
        a, b = inputs
        return (a.unsqueeze(1) + b.unsqueeze(2)).reshape(a.shape[0], -1)

    def reconstruct(self, idxs, node_num, sample):
        n0 = self.inputs[0].size
        return (self.inputs[0].reconstruct(idxs, node_num % n0, sample)
                + self.inputs[1].reconstruct(idxs, node_num // n0, sample))


class SumVector(NodeVector):
    """`size` mixtures over the concatenated product inputs; weights = log_softmax(params, 0)."""

    def __init__(self, prod_vectors, num_sums, args, name=''):
        super().__init__(name)
        self.inputs = prod_vectors              # plain list: children are registered elsewhere
        self.size = num_sums
        self.args = args
        self.scope = self.inputs[0].scope
        for v in self.inputs:
            assert set(v.scope) == set(self.scope)
        if args.linear_sum_weights or not args.normalized_sums:
            raise NotImplementedError('stove_b200: only normalised log-space sum weights are supported')
        self.params = args.param_provider.grab_sum_parameters(sum(v.size for v in prod_vectors), num_sums)

    def forward(self, inputs, get_child_acts=False):
        child = torch.cat(inputs, 1).unsqueeze(-1) + torch.log_softmax(self.params, 0)
        sums = torch.logsumexp(child, 1)
        return (sums, child) if get_child_acts else sums

    def reconstruct(self, idxs, node_num, sample):
        k = idxs[self][node_num]
        for v in self.inputs:
            if k < v.size:
                return v.reconstruct(idxs, k, sample)
            k -= v.size

    def num_params(self):
        return self.params.numel()

    def init_params(self, init_fn=None):
        (init_fn or truncated_normal_)(self.params)


class _Tables:
    """int32 structure tensors (kept on the device of the parameters) + the ctypes struct."""

    def __init__(self, kind, host, meta):
        self.kind, self.host, self.meta = kind, host, meta
        self.device = None
        self.dev = {}
        self.cstruct = None

    def to(self, device):
        if self.device == device:
            return self
        self.dev = {k: torch.from_numpy(v).to(device) for k, v in self.host.items()}
        m = self.meta
        if self.kind == 'D2':
            self.cstruct = N.Spn2Struct(m['D'], m['R'], m['G'], m['S'], m['pmax'],
                                        self.dev['region_scope'].data_ptr(), self.dev['region_n0'].data_ptr(),
                                        self.dev['region_n'].data_ptr(), self.dev['pix_slot'].data_ptr())
        else:
            self.cstruct = N.Spn1Struct(m['D'], m['R'], m['G'], self.dev['side'].data_ptr())
        self.device = device
        return self


class PackedSpn:
    """Parameters of one SPN in kernel layout (autograd-connected to the module parameters)."""

    def __init__(self, kind, tables, leaf, wlog=None, wlin=None, rlog=None, rlin=None):
        self.kind, self.tables = kind, tables
        self.leaf, self.wlog, self.wlin, self.rlog, self.rlin = leaf, wlog, wlin, rlog, rlin
        self.leaf_il_f = self.leaf_il_b = None      # background SPN: lane-interleaved copies (fused scene likelihood)
        # stream the packing ran on: its backward runs there too, so the parameter-gradient kernels of the
        # SPN backward can be handed to it and leave the chain of the caller's stream (ops.Spn2 / ops.Spn1)
        self.stream = torch.cuda.current_stream(leaf.device) if leaf.is_cuda else None


class RatSpn(nn.Module):
    def __init__(self, num_classes, region_graph, args=None, name=None):
        super().__init__()
        args = args or SpnArgs()
        self.name = name if name is not None else str(id(self))
        self._region_graph = region_graph
        self.args = args
        self.num_classes = num_classes
        self._region_distributions = {}
        self._region_products = {}
        self.vector_list = nn.ModuleList()
        self.output_vector = None
        self.num_dims = region_graph.get_num_items()
        self._build()
        self.init_params(init_fn=args.init_fn)
        self._tables = self._analyse()

    # -- construction (rat_torch.py:279-331) -------------------------------------------
    def _build(self):
        layers = self._region_graph.make_layers()
        self.rg_layers = layers
        if self.args.dist != 'Gauss':
            raise NotImplementedError('stove_b200: only Gauss leaves are supported')
        leaves = nn.ModuleList()
        for i, region in enumerate(layers[0]):
            vec = GaussVector(region, self.args, '%s_gauss_%d' % (self.name, i), num_dims=self.num_dims)
            leaves.append(vec)
            self._region_distributions[region] = vec
        self.vector_list.append(leaves)
        for depth in range(1, len(layers)):
            level = nn.ModuleList()
            if depth % 2 == 1:
                for i, partition in enumerate(layers[depth]):
                    first, second = partition[0], partition[1]
                    vec = ProductVector(self._region_distributions[first], self._region_distributions[second],
                                        '%s_prod_%d_%d' % (self.name, depth, i))
                    level.append(vec)
                    self._region_products.setdefault(tuple(sorted(first + second)), []).append(vec)
            else:
                width = self.num_classes if depth == len(layers) - 1 else self.args.num_sums
                for i, region in enumerate(layers[depth]):
                    vec = SumVector(self._region_products[region], width, self.args,
                                    name='%s_sum_%d_%d' % (self.name, depth, i))
                    level.append(vec)
                    self._region_distributions[region] = vec
            self.vector_list.append(level)
        self.output_vector = self._region_distributions[self._region_graph.get_root_region()]

    def init_params(self, init_fn):
        for layer in self.vector_list:
            for vec in layer:
                vec.init_params(init_fn)

    # -- structure analysis for the fused kernels ---------------------------------------
    def _analyse(self):
        root = self.output_vector
        depth = len(self.vector_list)
        if not isinstance(root, SumVector) or self.num_classes != 1 or depth not in (3, 5):
            return None
        D = self.num_dims
        G = self.args.num_gauss
        reps = root.inputs
        R = len(reps)
        if depth == 3:
            if not all(isinstance(v, GaussVector) for p in reps for v in p.inputs):
                return None
            side = np.zeros((D, R), dtype=np.int32)
            order, dst = [], []
            for r, prod in enumerate(reps):
                a, b = prod.inputs
                if sorted(a.scope + b.scope) != list(range(D)):
                    return None
                side[b.scope, r] = 1
                for leaf in (a, b):
                    order.append(leaf)
                    dst.extend(p * R + r for p in leaf.scope)
            # scope lists per leaf l = 2 r + side (ascending pixels), for the fused scene-likelihood kernels
            bg_scope = np.zeros((2 * R, D), dtype=np.int32)
            bg_cnt = np.zeros(2 * R, dtype=np.int32)
            for r in range(R):
                for h in (0, 1):
                    px = np.nonzero(side[:, r] == h)[0]
                    bg_scope[2 * r + h, :len(px)] = px
                    bg_cnt[2 * r + h] = len(px)
            # lane-interleaved copies of the packed leaf table for the fused kernels (ops.interleave_leaf): row order
            # of the forward pass = (leaf, position in its scope list), of the backward pass = (repetition, pixel)
            Ls = (int(bg_cnt.max()) + 31) // 32 * 32
            Dp = (D + 31) // 32 * 32
            il_f = np.full(2 * R * Ls, -1, dtype=np.int32)
            il_b = np.full(R * Dp, -1, dtype=np.int32)
            for r in range(R):
                for h in (0, 1):
                    l = 2 * r + h
                    il_f[l * Ls:l * Ls + bg_cnt[l]] = bg_scope[l, :bg_cnt[l]] * R + r
                il_b[r * Dp:r * Dp + D] = np.arange(D) * R + r
            host = {'side': side, 'dst_row': np.asarray(dst, dtype=np.int32), 'bg_scope': bg_scope, 'bg_cnt': bg_cnt,
                    'il_f': il_f, 'il_b': il_b}
            t = _Tables('D1', host, dict(D=D, R=R, G=G))
            t.il_stride_f, t.il_stride_b = Ls, Dp
            t.leaf_order, t.prow_total, t.GP = order, D * R, (G + 3) // 4 * 4
            return t
        S = self.args.num_sums
        regions = []
        for prod in reps:
            for mid in prod.inputs:
                if not isinstance(mid, SumVector) or len(mid.inputs) != 1 or mid.size != S:
                    return None
                l0, l1 = mid.inputs[0].inputs
                if not (isinstance(l0, GaussVector) and isinstance(l1, GaussVector)):
                    return None
                regions.append((mid, l0, l1))
        Q = len(regions)
        pmax = max(len(l0.scope) + len(l1.scope) for _, l0, l1 in regions)
        scope = np.full((Q, pmax), -1, dtype=np.int32)
        n0 = np.zeros(Q, dtype=np.int32)
        nt = np.zeros(Q, dtype=np.int32)
        slot = np.full((D, R), -1, dtype=np.int32)
        order, dst = [], []
        for q, (_, l0, l1) in enumerate(regions):
            px = list(l0.scope) + list(l1.scope)
            scope[q, :len(px)] = px
            n0[q], nt[q] = len(l0.scope), len(px)
            slot[px, q // 2] = q * pmax + np.arange(len(px))
            order += [l0, l1]
            dst.extend(q * pmax + i for i in range(len(px)))
        if (slot < 0).any():
            return None
        host = {'region_scope': scope, 'region_n0': n0, 'region_n': nt, 'pix_slot': slot,
                'dst_row': np.asarray(dst, dtype=np.int32)}
        t = _Tables('D2', host, dict(D=D, R=R, G=G, S=S, pmax=int(pmax)))
        t.leaf_order, t.prow_total, t.GP = order, Q * pmax, (G + 3) // 4 * 4
        t.mid_sums = [m for m, _, _ in regions]
        t.SP = (S + 3) // 4 * 4
        return t

    @property
    def fused_kind(self):
        return self._tables.kind if self._tables is not None else None

    # -- fused path ----------------------------------------------------------------------
    def pack(self):
        """Gather + transform the parameters into kernel layout (differentiable)."""
        t = self._tables
        if t is None:
            raise NotImplementedError(
                'stove_b200.RatSpn: this region-graph structure has no fused kernel (supported: the '
                'object/background SPNs of probabilistic_models.py); there is no generic fallback')
        dev = self.output_vector.params.device
        t.to(dev)
        a = self.args
        if a.gauss_min_sigma >= a.gauss_max_sigma or a.gauss_min_mean is not None or a.gauss_max_mean is not None:
            # rat_torch.py:83-96 of the reference: constant sigma = 1 / sigmoid-bounded means; not on the hot path
            raise NotImplementedError('stove_b200.RatSpn: the fused leaf kernels need gauss_min_sigma < gauss_max_sigma '
                                      'and unbounded means (gauss_min_mean = gauss_max_mean = None)')
        means = torch.cat([v.means for v in t.leaf_order], 0)
        sigma = torch.cat([v.sigma_params for v in t.leaf_order], 0)
        leaf = ops.PackLeaf.apply(means, sigma, t.dev['dst_row'], t.GP, t.prow_total,
                                  float(a.gauss_min_sigma), float(a.gauss_max_sigma))
        rlog, rlin = ops.PackSum.apply(self.output_vector.params.unsqueeze(0), 1)
        if t.kind == 'D1':
            pk = PackedSpn('D1', t, leaf, rlog=rlog, rlin=rlin)
            if (t.GP == 8) and leaf.is_cuda:
                pk.leaf_il_f = ops.interleave_leaf(leaf, t.dev['il_f'])
                pk.leaf_il_b = ops.interleave_leaf(leaf, t.dev['il_b'])
            return pk
        wlog, wlin = ops.PackSum.apply(torch.stack([m.params for m in t.mid_sums], 0), t.SP)
        return PackedSpn('D2', t, leaf, wlog, wlin, rlog, rlin)

    def forward_packed(self, packed, inputs, marginalized=None):
        if inputs.shape[0] == 0:
            return inputs.new_zeros(0, self.num_classes)
        pstream = packed.stream
        if pstream is not None and pstream == torch.cuda.current_stream(inputs.device):
            pstream = None
        if packed.kind == 'D2':
            out = ops.Spn2.apply(inputs, marginalized, packed.leaf, packed.wlog, packed.wlin, packed.rlog,
                                 packed.rlin, packed.tables, pstream)
        else:
            out = ops.Spn1.apply(inputs, marginalized, packed.leaf, packed.rlog, packed.rlin, packed.tables,
                                 pstream)
        return out.unsqueeze(1)

    def forward(self, inputs, marginalized=None):
        """Root log-likelihood (N, 1) of inputs (N, D) [rat_torch.py:354-357]."""
        return self.forward_packed(self.pack(), inputs, marginalized)

    # -- reference helpers kept for visualisation code -----------------------------------
    def compute_activations(self, inputs, marginalized=None, get_sum_child_acts=False):
        acts, child_acts = {}, {}
        for leaf in self.vector_list[0]:
            acts[leaf] = leaf.forward(inputs, marginalized)
        for layer in list(self.vector_list)[1:]:
            for vec in layer:
                ins = [acts[v] for v in vec.inputs]
                if isinstance(vec, SumVector) and get_sum_child_acts:
                    acts[vec], child_acts[vec] = vec.forward(ins, get_child_acts=True)
                else:
                    acts[vec] = vec.forward(ins)
        return (acts, child_acts) if get_sum_child_acts else acts

    def reconstruct(self, idxs, node_num, sample):
        return self.output_vector.reconstruct(idxs, node_num, sample)

    def get_sum_params(self):
        return {v: v.params for layer in self.vector_list for v in layer if isinstance(v, SumVector)}

    def num_params(self):
        return sum(v.num_params() for layer in self.vector_list for v in layer)
