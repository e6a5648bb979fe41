"""Random region graphs for RAT-SPNs (drop-in for model/spn/region_graph.py:5-155).

Same public surface (`RegionGraph(items, seed)`, `random_split`, `make_layers`, ...) and --
this is what matters for checkpoints -- the same resulting structure: leaf regions come out
lexicographically sorted, partition layers in the iteration order of a Python `set` of nested
int tuples (region_graph.py:141-142).  Hashing of int tuples is deterministic, so replaying
the same insertions into the same container types reproduces the reference's order; pinned
by tests/golden/spn_structure.json.  Structure only -- no tensor work happens here.
"""
import numpy as np


def _canon(ints):
    return tuple(sorted(int(v) for v in ints))


class RegionGraph(object):
    def __init__(self, items, seed=12345):
        self._items = _canon(items)
        self._regions = {self._items}
        self._partitions = set()
        self._child_partitions = {}
        self._rand_state = np.random.RandomState(seed)
        self._layers = []

    # -- accessors -------------------------------------------------------------------
    def get_root_region(self):
        return self._items

    def get_num_items(self):
        return len(self._items)

    def get_regions(self):
        return self._regions

    def get_child_partitions(self, region):
        return self._child_partitions[region]

    def get_region(self, region):
        region = _canon(region)
        if not set(region) <= set(self._items):
            raise ValueError('Argument region is not a sub-set of _items.')
        self._regions.add(region)
        return region

    def get_leaf_regions(self):
        return [r for r in self._regions if r not in self._child_partitions]

    # -- construction ----------------------------------------------------------------
    def _register(self, region, parts):
        partition = tuple(sorted(parts))
        if partition not in self._partitions:
            self._partitions.add(partition)
            self._child_partitions[region] = self._child_partitions.get(region, []) + [partition]
        return partition

    def random_split(self, num_parts, num_recursions=1, region=None):
        """Shuffle `region` with the private RNG, cut it into `num_parts` nearly equal
        pieces and recurse `num_recursions - 1` levels into every piece."""
        if num_recursions < 1:
            return None
        region = region if region else self._items
        if region not in self._regions:
            raise LookupError('Trying to split non-existing region.')
        if len(region) == 1:
            return None
        shuffled = self._rand_state.permutation(list(region))
        num_parts = min(len(shuffled), num_parts)
        bounds = np.cumsum([0] + [len(shuffled) // num_parts + (1 if k < len(shuffled) % num_parts else 0)
                                  for k in range(num_parts)])
        parts = [_canon(shuffled[lo:hi]) for lo, hi in zip(bounds[:-1], bounds[1:])]
        self._regions.update(parts)
        partition = self._register(region, parts)
        if num_recursions > 1:
            for sub in partition:
                self.random_split(num_parts, num_recursions - 1, sub)
        return partition

    def make_split(self, region, sub_region):
        region, sub_region = set(region), set(sub_region)
        key = _canon(region)
        if key not in self._regions:
            raise LookupError('Trying to split non-existing region.')
        if not sub_region < region or not sub_region:
            raise AssertionError('sub-region is not a proper sub-set.')
        a, b = _canon(sub_region), _canon(region - sub_region)
        self._regions.update((a, b))
        return self._register(key, [a, b])

    # -- layering --------------------------------------------------------------------
    def make_layers(self):
        """[leaf regions, partitions, regions, ..., [root]] (odd layers: partitions)."""
        leaves = sorted(self.get_leaf_regions())
        self._layers = [leaves]
        if leaves == [self._items]:
            return self._layers
        done_r, done_p = set(leaves), set()
        while len(done_r) < len(self._regions) or len(done_p) < len(self._partitions):
            ready_p = [p for p in self._partitions
                       if p not in done_p and all(r in done_r for r in p)]
            done_p.update(ready_p)
            ready_r = sorted(r for r in self._regions
                             if r not in done_r and all(p in done_p for p in self._child_partitions[r]))
            done_r.update(ready_r)
            self._layers += [ready_p, ready_r]
        return self._layers
