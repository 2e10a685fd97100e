"""The consumer of `invert` / `sample_and_replace`: the Bayesian-network evaluation loop of the reference
(scripts/evaluate.py:86-152, `eval_nn` / `eval_bnn`) with the same call signatures and return values.

What is different underneath:
  * an estimator that has `sample_many` (KFAC) draws ALL posterior samples of all layers up front with one kernel launch
    per chunk (the S draws of a layer are one pair of GEMMs, crv_sample_matrix_normal_multi) instead of two GEMMs per
    layer per sample; installing sample s is then two multi-tensor copies;
  * with `torch.distributed` initialised the samples are sharded over the ranks (`shard_indices`; drawing them needs no
    communication) and ONE all-reduce of the summed predictions at the end gives every rank the ensemble mean.
The metric helpers restate the definitions of curvature/utils.py:79-91, 141-152, 207-247, 250-267 (top-1 accuracy in
percent, NLL with the 1e-12 guard, 10-bin expected calibration error, predictive entropy)."""
from typing import Optional

import numpy as np
import torch


def accuracy(probabilities: np.ndarray, labels: np.ndarray) -> float:
    return 100.0 * float(np.mean(np.argmax(probabilities, axis=1) == labels))


def negative_log_likelihood(probabilities: np.ndarray, labels: np.ndarray) -> float:
    return -float(np.mean(np.log(probabilities[np.arange(probabilities.shape[0]), labels] + 1e-12)))


def expected_calibration_error(probabilities: np.ndarray, labels: np.ndarray, bins: int = 10) -> float:
    conf = probabilities.max(axis=1)
    hit = (np.argmax(probabilities, axis=1) == labels).astype(np.float64)
    edges = np.linspace(0, 1, bins + 1)
    ece = 0.0
    for lo, hi in zip(edges[:-1], edges[1:]):
        mask = (conf > lo) & (conf <= hi)
        if mask.any():
            ece += mask.mean() * abs(conf[mask].mean() - hit[mask].mean())
    return float(ece)


def predictive_entropy(probabilities: np.ndarray, mean: bool = False):
    p = probabilities / probabilities.sum(axis=1, keepdims=True)
    ent = -np.sum(np.where(p > 0, p * np.log(np.where(p > 0, p, 1.0)), 0.0), axis=1)
    return float(ent.mean()) if mean else ent


def eval_nn(model, dataset, device=torch.device('cuda'), verbose=False):
    """Softmax predictions and labels over `dataset` (an iterable of (images, labels) batches); scripts/evaluate.py:86-118."""
    model.eval()
    logits_list, labels_list = [], []
    with torch.no_grad():
        for images, labels in dataset:
            logits_list.append(model(images.to(device, non_blocking=True)))
            labels_list.append(labels)
        predictions = torch.nn.functional.softmax(torch.cat(logits_list), dim=1).cpu().numpy()
        labels = torch.cat(labels_list).cpu().numpy()
    if verbose:
        print(f"Accuracy: {accuracy(predictions, labels):.2f}% | ECE: {100 * expected_calibration_error(predictions, labels):.2f}%")
    return predictions, labels


def eval_bnn(model, dataset, estimator, samples=30, stats=False, device=torch.device('cuda'), verbose=True,
             group=None, chunk: Optional[int] = None):
    """Monte-Carlo ensemble prediction: for every posterior sample replace the parameters and run `eval_nn`, average the
    predictions (scripts/evaluate.py:121-152).  Returns (mean_predictions, labels, stats_list) like the reference.
    `chunk` bounds how many samples are drawn per `sample_many` call (default: all of this rank's, at most 32)."""
    import torch.distributed as dist
    from .parallel import shard_indices
    model.eval()
    world = dist.get_world_size(group) if (dist.is_available() and dist.is_initialized()) else 1
    rank = dist.get_rank(group) if world > 1 else 0
    mine = shard_indices(samples, rank, world) if world > 1 else list(range(samples))
    stats_list = {"acc": [], "ece": [], "nll": [], "ent": []}
    total = None
    labels = None
    many = getattr(estimator, "sample_many", None)
    chunk = max(1, min(len(mine), chunk or 32)) if mine else 1
    done = 0
    with torch.no_grad():
        for c0 in range(0, len(mine), chunk):
            n = min(chunk, len(mine) - c0)
            draws = estimator.sample_many(n) if many is not None else None
            for j in range(n):
                if draws is not None:
                    estimator.replace_with(draws, j)
                else:
                    estimator.sample_and_replace()
                predictions, labels = eval_nn(model, dataset, device)
                total = predictions.astype(np.float64) if total is None else total + predictions
                done += 1
                if stats:
                    running = total / done
                    stats_list["acc"].append(accuracy(running, labels))
                    stats_list["ece"].append(100 * expected_calibration_error(running, labels))
                    stats_list["nll"].append(negative_log_likelihood(predictions, labels))
                    stats_list["ent"].append(predictive_entropy(running, mean=True))
            del draws
        if world > 1:
            if total is None:                 # a rank without samples still takes part in the one collective
                predictions, labels = eval_nn(model, dataset, device)
                total = np.zeros_like(predictions, dtype=np.float64)
            t = torch.from_numpy(total).to(device)
            dist.all_reduce(t, group=group)
            total = t.cpu().numpy()
        mean_predictions = (total / samples).astype(np.float32)
    if verbose and rank == 0:
        print(f"Accuracy: {accuracy(mean_predictions, labels):.2f}% | ECE: {100 * expected_calibration_error(mean_predictions, labels):.2f}%")
    return mean_predictions, labels, stats_list
