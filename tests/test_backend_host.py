"""CPU-only: host logic of the SLOS_B200 backend that needs no device -- the compute tree shared by several input states
(reference perceval/backends/_slos.py:36-86) and the mask plumbing when Perceval hands the backend an exqalibur FSMask."""
import numpy as np

from perceval_b200 import BasicState
from perceval_b200.backends.slos import SLOSB200Backend, _PathNode


def ref_decompose(states):
    """_Path.__init__ + _decompose (_slos.py:39-86) restated on plain lists: returns {mode: subtree}, ends-here list."""
    def build(n, sts, targets):
        ends, t2, s2 = [], [], []
        for t, s in zip(targets, sts):
            if sum(t) == 0:
                ends.append(s)
            else:
                t2.append(t)
                s2.append(s)
        children = {}
        while t2:
            counts = [0] * len(t2[0])
            for one in t2:
                counts = [x + y for x, y in zip(counts, one)]
            max_index = counts.index(max(counts))
            cur_t, cur_s, new_t, new_s = [], [], [], []
            for one_t, one_s in zip(t2, s2):
                if one_t[max_index]:
                    one_t[max_index] -= 1
                    cur_t.append(one_t)
                    cur_s.append(one_s)
                else:
                    new_t.append(one_t)
                    new_s.append(one_s)
            children[max_index] = build(n + 1, cur_s, cur_t)
            t2, s2 = new_t, new_s
        return (n, ends, children)
    return build(0, list(states), [list(s) for s in states])


def same_tree(node, ref):
    n, ends, children = ref
    assert node.depth == n and node.states == ends and list(node.children.keys()) == list(children.keys())
    for mode, child in node.children.items():
        same_tree(child, children[mode])


def test_path_tree_matches_reference_decomposition():
    rng = np.random.default_rng(0)
    cases = [[(1, 1, 1, 0)], [(1, 1, 0, 0), (1, 0, 1, 0), (1, 1, 1, 0)], [(2, 1, 0), (0, 3, 1), (0, 0, 0), (2, 1, 0)]]
    for _ in range(20):
        m = int(rng.integers(2, 7))
        cases.append([tuple(int(x) for x in rng.integers(0, 3, m)) for _ in range(int(rng.integers(1, 6)))])
    for states in cases:
        sts = [BasicState(list(s)) for s in states]
        root = _PathNode(0, list(sts), [list(s) for s in states])
        same_tree(root, ref_decompose(sts))


def test_shared_prefix_layer_count():
    # three noisy inputs sharing two photons: 2 shared layers + 1 + 1 + (0: the two-photon input ends on the shared node)
    sts = [BasicState([1, 1, 1, 0, 0]), BasicState([1, 1, 0, 1, 0]), BasicState([1, 1, 0, 0, 0])]
    root = _PathNode(0, list(sts), [list(s) for s in sts])
    assert root.count_layers() == 4          # separate chains would need 3 + 3 + 2 = 8
    assert root.nmax == 3


class _XqMaskStub:
    """Stands in for exqalibur's FSMask (what Perceval's own _init_mask stores in self._mask): only .match()."""

    def __init__(self, strings):
        self.strings = strings

    def match(self, state, allow_missing=False):
        return any(all(ch in " *" or int(state[i]) == ord(ch) - 0x30 for i, ch in enumerate(s)) for s in self.strings)


class _Circ:
    m = 4
    requires_polarization = False


def test_device_mask_is_built_from_mask_strings_not_from_backend_mask_object():
    """with a real Perceval install self._mask is an xq.FSMask without conds_array()/at_least_bits(); the device
    path must use its own FockMask built from _masks_str / _mask_n / _no_limit_modes."""
    b = SLOSB200Backend()
    b._circuit = _Circ()
    b._input_state = BasicState([1, 0, 1, 0])
    b.set_mask(["**00", "1***"], at_least_modes=[0])
    b._mask = _XqMaskStub(["**00", "1***"])          # what Perceval's base class would have stored
    dm = b._dev_mask
    assert dm is not None and dm.conds_array().shape == (2, 4)
    assert dm.conds_array().tolist() == [[-1, -1, 0, 0], [1, -1, -1, -1]]
    assert dm.at_least_bits() == 1
    assert dm.match([2, 0, 0, 0]) and not dm.match([0, 1, 1, 0])
    b.clear_mask()
    assert b._dev_mask is None and b._masks_str is None
