"""Name-keyed on-disk format for estimator state.

The reference pickles `est.state`, a dict keyed by the `nn.Module` OBJECTS (`scripts/factors.py:122-129`, reloaded at
`scripts/evaluate.py:355-370`): it drags the modules into the file, is identity-keyed (useless in another process
unless the very same pickled modules are reused) and is rejected by `torch.load(weights_only=True)`.  Here a state
file holds only plain tensors and strings: the flat arena and, per state entry, the layer's qualified name in
`model.named_modules()` and the shapes of its views.  Loading checks names and shapes against the receiving
estimator and copies the arena in: resume of an interrupted estimation pass, and reload in another process.
"""
from typing import Dict

import torch

FORMAT = "curvature_b200.factors.v1"


def _layer_names(est) -> Dict[int, str]:
    names = {id(m): n for n, m in est.model.named_modules()}
    names.update({id(k): k for k in ("attn_in", "attn_out")})
    return names


def _key_name(names, key) -> str:
    return key if isinstance(key, str) else names[id(key)]


def save_factors(est, path: str) -> None:
    """Write the running sums of `est` (KFAC, Diagonal or EFB; after at least one `update`) to `path`."""
    if getattr(est, "arena", None) is None or not est.state:
        raise RuntimeError("nothing to save: call 'update' first")
    names = _layer_names(est)
    entries = []
    for key, views in est._views.items():
        vs = views if isinstance(views, (list, tuple)) else [views]
        entries.append({"layer": _key_name(names, key), "shapes": [list(v.shape) for v in vs]})
    torch.save({"format": FORMAT, "estimator": est.__class__.__name__, "entries": entries,
                "arena": est.arena.flat.detach().cpu()}, path)


def load_factors(est, path: str, accumulate: bool = False) -> None:
    """Load a file written by `save_factors` into `est` (same class, same selected layers).  `accumulate=True` adds
    to the current sums instead of replacing them (merging shards estimated by different processes)."""
    blob = torch.load(path, map_location="cpu", weights_only=True)
    if blob.get("format") != FORMAT:
        raise ValueError(f"{path}: not a {FORMAT} file")
    if blob["estimator"] != est.__class__.__name__:
        raise ValueError(f"{path} holds {blob['estimator']} state, not {est.__class__.__name__}")
    est._ensure_arena()
    names = _layer_names(est)
    mine = []
    for key, views in est._views.items():
        vs = views if isinstance(views, (list, tuple)) else [views]
        mine.append({"layer": _key_name(names, key), "shapes": [list(v.shape) for v in vs]})
    if mine != blob["entries"]:
        theirs = [e["layer"] for e in blob["entries"]]
        raise ValueError(f"{path}: layers / shapes differ from this estimator's (file: {theirs[:4]}..., "
                         f"here: {[e['layer'] for e in mine[:4]]}...)")
    arena = blob["arena"].to(est.arena.flat.device)
    if arena.shape != est.arena.flat.shape:
        raise ValueError(f"{path}: arena has {arena.numel()} elements, expected {est.arena.flat.numel()}")
    if accumulate:
        est.arena.flat.add_(arena)
    else:
        est.arena.flat.copy_(arena)
    for key, views in est._views.items():
        if hasattr(est, "diags"):              # EFB: (lambdas, diags) per layer
            est.state[key], est.diags[key] = views
        else:
            est.state[key] = views
