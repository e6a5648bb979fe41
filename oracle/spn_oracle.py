"""ORACLE (test infrastructure only) -- CPU restatement of the reference RAT-SPN.

This file is the *checker*, never the product: only `tests/`, `__graft_entry__.smoke()`
and `bench.py`'s cpu_baseline / `--impl reference` leg may import it.  Nothing under
`stove_b200/` imports `oracle/`.

Parity status: **pinned** against outputs of the reference itself, executed in the build
container (the reference is pure Python and imports here).  The generating script is
`oracle/make_golden.py`; the vectors live in `tests/golden/`.  The reference ships no tests
or golden vectors of its own.

What is restated (reference file:line):
  * random region graph + layering      model/spn/region_graph.py:54-95, 118-155
  * leaf / product / sum vector wiring   model/spn/rat_torch.py:279-331
  * Gauss leaf                           model/spn/rat_torch.py:83-109
  * product (log-space outer sum)        model/spn/rat_torch.py:147-163
  * sum (log_softmax + logsumexp)        model/spn/rat_torch.py:202-222
  * object / background SPN builders     model/spn/probabilistic_models.py:8-39

The evaluation is tensorised (all leaves of a layer in one gather) but follows the
reference's layer-by-layer algorithm, dtype-agnostic (fp64 = gold, fp32 = like-for-like).
"""
import math

import numpy as np
import torch


# --------------------------------------------------------------------------------------
# Region graph (structure only).  The partition-layer order of the reference comes from
# iterating a Python `set` of nested int tuples (region_graph.py:141-142); hashing of int
# tuples is deterministic within one interpreter, so re-creating the same set with the
# same insertion history reproduces the reference's order.  Pinned by
# tests/golden/spn_structure.json.
# --------------------------------------------------------------------------------------
def build_region_layers(num_items, seed, splits):
    """splits: list of (num_parts, num_recursions), applied to the root in order.

    Returns the layered structure `[leaf regions, partitions, regions, ..., [root]]`
    exactly as region_graph.py:118-155 would.
    """
    rng = np.random.RandomState(seed)
    root = tuple(range(num_items))
    regions = {root}
    partitions = set()
    children = {}

    def split(region, parts, depth):
        # region_graph.py:54-95
        if depth < 1 or len(region) == 1:
            return
        perm = [int(v) for v in rng.permutation(list(region))]
        parts_here = min(len(perm), parts)
        base, extra = divmod(len(perm), parts_here)
        subs, at = [], 0
        for k in range(parts_here):
            width = base + (1 if k < extra else 0)
            sub = tuple(sorted(perm[at:at + width]))
            subs.append(sub)
            regions.add(sub)
            at += width
        partition = tuple(sorted(subs))
        if partition not in partitions:
            partitions.add(partition)
            children[region] = children.get(region, []) + [partition]
        if depth > 1:
            for sub in partition:
                split(sub, parts, depth - 1)

    for parts, depth in splits:
        split(root, parts, depth)

    # region_graph.py:118-155
    leaves = sorted(r for r in regions if r not in children)
    layers = [leaves]
    if len(leaves) == 1 and root in leaves:
        return layers
    seen_r, seen_p = set(leaves), set()
    while len(seen_r) != len(regions) or len(seen_p) != len(partitions):
        p_layer = [p for p in partitions
                   if p not in seen_p and all(r in seen_r for r in p)]
        layers.append(p_layer)
        seen_p.update(p_layer)
        r_layer = sorted(r for r in regions
                         if r not in seen_r and all(p in seen_p for p in children[r]))
        layers.append(r_layer)
        seen_r.update(r_layer)
    return layers


class SpnStructure:
    """Wiring of a RAT-SPN as `RatSpn._make_spn_from_region_graph` builds it
    (rat_torch.py:279-331).  `layers[l]` is a list of node-vector descriptors:
      leaf    : {'kind': 'leaf', 'scope': [...], 'size': G}
      product : {'kind': 'prod', 'inputs': [(layer, idx), (layer, idx)], 'size': ..}
      sum     : {'kind': 'sum',  'inputs': [(layer, idx), ...],          'size': ..}
    """

    def __init__(self, num_items, seed, splits, num_gauss, num_sums, num_classes=1):
        rg = build_region_layers(num_items, seed, splits)
        self.num_items = num_items
        self.layers = [[]]
        where = {}          # region -> (layer, idx) of its distribution vector
        prods_of = {}       # region -> list of (layer, idx) of its product vectors
        for i, region in enumerate(rg[0]):
            self.layers[0].append({'kind': 'leaf', 'scope': list(region), 'size': num_gauss})
            where[region] = (0, i)
        for l in range(1, len(rg)):
            self.layers.append([])
            if l % 2 == 1:
                for i, partition in enumerate(rg[l]):
                    a, b = partition[0], partition[1]
                    ia, ib = where[a], where[b]
                    size = self._size(ia) * self._size(ib)
                    self.layers[l].append({'kind': 'prod', 'inputs': [ia, ib], 'size': size})
                    prods_of.setdefault(tuple(sorted(a + b)), []).append((l, i))
            else:
                width = num_classes if l == len(rg) - 1 else num_sums
                for i, region in enumerate(rg[l]):
                    self.layers[l].append({'kind': 'sum', 'inputs': list(prods_of[region]),
                                           'size': width})
                    where[region] = (l, i)
        self.root = where[tuple(range(num_items))]

    def _size(self, ref):
        return self.layers[ref[0]][ref[1]]['size']

    def param_shapes(self):
        """{'vector_list.L.I.name': shape} in module registration order."""
        out = {}
        for l, layer in enumerate(self.layers):
            for i, v in enumerate(layer):
                if v['kind'] == 'leaf':
                    out['vector_list.%d.%d.means' % (l, i)] = (len(v['scope']), v['size'])
                    out['vector_list.%d.%d.sigma_params' % (l, i)] = (len(v['scope']), v['size'])
                elif v['kind'] == 'sum':
                    n_in = sum(self._size(r) for r in v['inputs'])
                    out['vector_list.%d.%d.params' % (l, i)] = (n_in, v['size'])
        return out

    def describe(self):
        """JSON-able summary used for the structure golden."""
        return {
            'leaf_scopes': [v['scope'] for v in self.layers[0]],
            'wiring': [[[list(r) for r in v['inputs']] for v in layer]
                       for layer in self.layers[1:]],
        }


def obj_spn_structure(patch_size, seed, num_gauss=10, num_sums=10):
    """probabilistic_models.py:8-22: six `random_split(2, 2)` of the patch pixels."""
    return SpnStructure(patch_size, seed, [(2, 2)] * 6, num_gauss, num_sums)


def bg_spn_structure(image_size, seed):
    """probabilistic_models.py:25-39: three `random_split(2, 1)`; 6 gaussians, 3 sums."""
    return SpnStructure(image_size, seed, [(2, 1)] * 3, 6, 3)


_HALF_LOG_2PI = 0.5 * math.log(2.0 * math.pi)


def spn_forward(struct, params, x, marg, min_var, max_var, prefix='', return_all=False):
    """Root log-likelihood (N, num_classes) of a RAT-SPN, rat_torch.py:333-357.

    params: dict name -> tensor with the reference's state_dict names
            (`prefix + 'vector_list.L.I.means'` ...).
    """
    acts = {}
    if marg is not None:
        keep = 1.0 - torch.clamp(marg, 0.0, 1.0)                     # rat_torch.py:104-106
    for i, v in enumerate(struct.layers[0]):
        mu = params['%svector_list.0.%d.means' % (prefix, i)]
        sp = params['%svector_list.0.%d.sigma_params' % (prefix, i)]
        var = min_var + (max_var - min_var) * torch.sigmoid(sp)       # rat_torch.py:85-87,98-99
        xi = x[:, v['scope']].unsqueeze(-1)
        logpdf = -(xi - mu) ** 2 / (2.0 * var) - 0.5 * torch.log(var) - _HALF_LOG_2PI
        if marg is not None:
            logpdf = logpdf * keep[:, v['scope']].unsqueeze(-1)
        acts[(0, i)] = logpdf.sum(1)                                  # rat_torch.py:108
    for l in range(1, len(struct.layers)):
        for i, v in enumerate(struct.layers[l]):
            ins = [acts[tuple(r)] for r in v['inputs']]
            if v['kind'] == 'prod':
                # rat_torch.py:155-161: out[b, j*n0 + i] = in0[b, i] + in1[b, j]
                a, b = ins
                acts[(l, i)] = (a.unsqueeze(1) + b.unsqueeze(2)).reshape(a.shape[0], -1)
            else:
                w = torch.log_softmax(params['%svector_list.%d.%d.params' % (prefix, l, i)], 0)
                cat = torch.cat(ins, 1)
                acts[(l, i)] = torch.logsumexp(cat.unsqueeze(-1) + w, 1)  # rat_torch.py:214-219
    if return_all:
        return acts
    return acts[struct.root]


def init_spn_params(struct, generator=None, dtype=torch.float64, prefix=''):
    """Truncated-normal(std 0.1) init in the spirit of rat_torch.py:11-18 (values differ from
    the reference's RNG stream; parity tests load the reference's own state_dict instead)."""
    out = {}
    for name, shape in struct.param_shapes().items():
        t = torch.randn(*shape, generator=generator, dtype=torch.float64).clamp_(-2, 2) * 0.1
        out[prefix + name] = t.to(dtype)
    return out
