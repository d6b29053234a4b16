"""Result wire formats of the step after the hot path: the text forms Perceval ships results in,
built from DEVICE tensors so that 1e5 - 1e6 samples never become one Python object each.

Reference: perceval/serialization/_state_serialization.py:68-92 (``serialize_bssamples``: unique states in order of first
appearance + one index per sample), perceval/serialization/serialize.py:175-206 (``:PCVL:BSDistribution:{state=p;...}``,
``:PCVL:BSCount:{state=count;...}``, ``:PCVL:BSSamples:...``, optional zlib + base64 behind ``:PCVL:zip:``),
perceval/utils/conversion.py:52-70,143-155 (samples -> sample count -> probabilities), perceval/utils/format.py:34-85
(``simple_float(v, nsimplify=False)``, the number format of a serialised distribution).

The de-duplication (unique + first-appearance order + counts) runs on the device: states are mapped to their FSArray rank
(C ABI ``fock_rank``) when the rank fits 63 bits, otherwise rows are compared directly; only the U unique states and the
per-sample indices cross PCIe.  String assembly is host work by nature.
"""
from __future__ import annotations

import zlib
from base64 import b64decode, b64encode

import numpy as np
import torch

from . import fsarray

SEP = ":"
PCVL_PREFIX = f"{SEP}PCVL{SEP}"
ZIP_PREFIX = f"{PCVL_PREFIX}zip{SEP}"
BSD_TAG, BSC_TAG, BSS_TAG = "BSDistribution", "BSCount", "BSSamples"


# ------------------------------------------------------------------ numbers and states
def format_probability(v: float, precision: float = 1e-6) -> str:
    """``simple_float(v, nsimplify=False)[1]`` of perceval/utils/format.py:34-85: six significant decimals of the mantissa,
    values below 1e-3 as ``<mantissa>e-<k>``, trailing zeros dropped."""
    v = float(v)
    sign = ""
    if v < 0:
        sign, v = "-", -v
    alpha, mult10 = v, 0
    while alpha and alpha < 1:
        mult10 += 1
        alpha = alpha * 10
    if mult10 <= 3:
        while mult10:
            alpha = alpha / 10
            mult10 -= 1
    alpha = float(np.float64(alpha / precision).round()) * precision
    s = "%.15g" % alpha
    if "e" in s or "E" in s:            # sympy prints 15 significant digits positionally in the range that occurs here
        s = np.format_float_positional(alpha, precision=15, unique=False, trim="-")
    if "." in s:
        s = s.rstrip("0").rstrip(".")
    if not s:
        s = "0"
    s = sign + s
    if mult10:
        s += "e-%d" % mult10
    return s


def state_str(row) -> str:
    """``str(BasicState)`` = ``serialize_state`` (_state_serialization.py:36-37)."""
    return "|" + ",".join(str(int(x)) for x in row) + ">"


def _compress(text: str, compress: bool) -> str:
    if not compress:
        return text
    return ZIP_PREFIX + b64encode(zlib.compress(text.encode("utf-8"))).decode("utf-8")


def decompress(text: str) -> str:
    if text.startswith(ZIP_PREFIX):
        return zlib.decompress(b64decode(text[len(ZIP_PREFIX):])).decode("utf-8")
    return text


# ------------------------------------------------------------------ device side: unique + first-appearance order
def samples_first_appearance(samples: torch.Tensor, engine=None):
    """(unique_states (U, m) uint8, order (count,) int64, counts (U,) int64) of a chronological (count, m) uint8 sample
    tensor: ``unique_states`` in order of first appearance -- the key order of ``Counter(sample_list)``
    (conversion.py:52-60) and of the mapping in serialize_bssamples -- and ``order[i]`` the index of sample i in it."""
    assert samples.dim() == 2 and samples.dtype == torch.uint8
    count, m = samples.shape
    dev = samples.device
    if count == 0:
        return samples.new_empty((0, m)), torch.empty(0, dtype=torch.int64, device=dev), torch.empty(0, dtype=torch.int64, device=dev)
    n = int(samples[0].sum().item())
    keys = None
    if engine is not None and samples.is_cuda and m <= 64 and n <= 32 and fsarray.count(m, n) < (1 << 62):
        same_n = bool((samples.sum(dim=1, dtype=torch.int64) == n).all().item())
        if same_n:
            keys = engine.rank(m, n, samples)
    if keys is not None:
        uniq_keys, inverse, counts = torch.unique(keys, return_inverse=True, return_counts=True)
        nu = uniq_keys.numel()
    else:
        uniq_rows, inverse, counts = torch.unique(samples, dim=0, return_inverse=True, return_counts=True)
        nu = uniq_rows.shape[0]
    first = torch.full((nu,), count, dtype=torch.int64, device=dev)
    first.scatter_reduce_(0, inverse, torch.arange(count, dtype=torch.int64, device=dev), reduce="amin")
    by_first = torch.argsort(first)                       # unique ids in order of first appearance
    label = torch.empty(nu, dtype=torch.int64, device=dev)
    label[by_first] = torch.arange(nu, dtype=torch.int64, device=dev)
    order = label[inverse]
    states = samples[first[by_first]]                     # the first sample of each unique state, in that order
    return states, order, counts[by_first]


def samples_to_sample_count(samples: torch.Tensor, engine=None):
    """conversion.py:52-60 on tensors: (states (U, m), counts (U,)) in Counter order (first appearance)."""
    states, _order, counts = samples_first_appearance(samples, engine)
    return states, counts


def samples_to_probs(samples: torch.Tensor, engine=None):
    """conversion.py:63-70,143-155: sample count normalised to a distribution (states, float64 probabilities)."""
    states, counts = samples_to_sample_count(samples, engine)
    total = counts.sum()
    return states, counts.to(torch.float64) / total.to(torch.float64) if counts.numel() else counts.to(torch.float64)


# ------------------------------------------------------------------ text forms
def serialize_bssamples(samples: torch.Tensor, engine=None) -> str:
    """The string of reference serialize_bssamples (_state_serialization.py:68-78) for the chronological list ``samples``."""
    states, order, _ = samples_first_appearance(samples, engine)
    st = states.cpu().numpy()
    return ";".join(state_str(r) for r in st) + "/" + ";".join(map(str, order.cpu().tolist()))


def deserialize_bssamples(text: str) -> np.ndarray:
    """(count, m) uint8 array back from serialize_bssamples (_state_serialization.py:81-92)."""
    parts = text.split("/")
    assert len(parts) == 2, f"Bad serialized BSSamples: {text[:80]}"
    if not parts[0]:
        return np.zeros((0, 0), dtype=np.uint8)
    table = np.array([[int(x) for x in s.strip()[1:-1].split(",")] if s.strip()[1:-1] else [] for s in parts[0].split(";")], dtype=np.uint8)
    order = np.array([int(x) for x in parts[1].split(";")], dtype=np.int64)
    return table[order]


def serialize_samples(samples: torch.Tensor, engine=None, compress: bool = True) -> str:
    """serialize(BSSamples) of serialize.py:199-206."""
    return _compress(f"{PCVL_PREFIX}{BSS_TAG}{SEP}" + serialize_bssamples(samples, engine), compress)


def serialize_count(states, counts, compress: bool = True) -> str:
    """serialize(BSCount) of serialize.py:187-196 from (states (U, m), counts (U,))."""
    st = states.cpu().numpy() if isinstance(states, torch.Tensor) else np.asarray(states)
    ct = counts.cpu().tolist() if isinstance(counts, torch.Tensor) else list(counts)
    body = ";".join("%s=%s" % (state_str(r), str(int(c))) for r, c in zip(st, ct))
    return _compress(f"{PCVL_PREFIX}{BSC_TAG}{SEP}{{" + body + "}", compress)


def serialize_distribution(states, probs, compress: bool = True) -> str:
    """serialize(BSDistribution) of serialize.py:175-184 from (states (U, m), probabilities (U,))."""
    st = states.cpu().numpy() if isinstance(states, torch.Tensor) else np.asarray(states)
    pr = probs.cpu().tolist() if isinstance(probs, torch.Tensor) else list(probs)
    body = ";".join("%s=%s" % (state_str(r), format_probability(p)) for r, p in zip(st, pr))
    return _compress(f"{PCVL_PREFIX}{BSD_TAG}{SEP}{{" + body + "}", compress)


def serialize_backend_distribution(backend, min_p: float = 1e-16, compress: bool = True) -> str:
    """The BSDistribution text of ``backend.prob_distribution()`` without building it: thresholding and un-ranking on the
    device (SLOSB200Backend.prob_iterator_tensors), only the kept states cross PCIe.  BSDistribution.add drops
    p <= min_p exactly like this (perceval/utils/globals.py:30-34)."""
    st = backend._input_state
    ranks, probs = backend.prob_iterator_tensors(min_p)
    occ = backend._eng().unrank(st.m, st.n, ranks)
    return serialize_distribution(occ, probs, compress)
