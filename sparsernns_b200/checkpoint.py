"""Parameter-tree I/O for the NDNS model (host side).

The reference keeps its state in nested dicts (Flax ``params`` / ``batch_stats``) and, after conversion,
in pickles: ``sc_calibrated_params.pkl`` and ``sc_cal_stats.pkl`` (convert.py:963-971), which
``fxprun.py`` merges with ``load_modeldict`` (fxputils.py:121-134).  This module reads those pickles when
they exist and offers a flat ``.npz`` form of the same trees for the fixtures shipped with this repo.
"""

from __future__ import annotations

import pickle

import numpy as np

SEP = "/"


def flatten(tree: dict, prefix: str = "") -> dict:
    out = {}
    for k, v in tree.items():
        key = f"{prefix}{SEP}{k}" if prefix else str(k)
        if isinstance(v, dict):
            out.update(flatten(v, key))
        else:
            out[key] = np.asarray(v)
    return out


def unflatten(flat: dict) -> dict:
    tree: dict = {}
    for key, v in flat.items():
        node = tree
        parts = key.split(SEP)
        for p in parts[:-1]:
            node = node.setdefault(p, {})
        node[parts[-1]] = v
    return tree


def save_npz(path: str, tree: dict) -> None:
    np.savez_compressed(path, **flatten(tree))


def load_npz(path: str) -> dict:
    with np.load(path, allow_pickle=False) as z:
        return unflatten({k: z[k] for k in z.files})


def _to_numpy(tree):
    if isinstance(tree, dict):
        return {k: _to_numpy(v) for k, v in tree.items()}
    return np.asarray(tree)


def load_reference_pickles(params_fname: str, stats_fname: str):
    """(params, stats) from the reference's conversion output; leaves become NumPy arrays.

    Unpickling arrays written by JAX needs jaxlib on the reading side; that is a property of the files,
    not of this library.
    """
    with open(params_fname, "rb") as f:
        params = _to_numpy(pickle.load(f))
    with open(stats_fname, "rb") as f:
        stats = _to_numpy(pickle.load(f))
    return params, stats
