"""Result I/O of the test entry point (test_recognizer.py:119-139 and codes/core/evaluation/accuracy.py): the per-video
class scores `Recognizer2D.forward_test` returns are stacked into one (videos, classes) array and pickled (`--out`,
default.pkl), scored with top-k / mean-class accuracy, and several such files can be fused with weights.  Host-side numpy:
nothing here touches the GPU path.
"""
from __future__ import annotations

import pickle

import numpy as np


def stack_results(outputs):
    """list of (1, classes) or (classes,) arrays, one per video -> (videos, classes) (np.vstack in test_recognizer.py:122)."""
    return np.vstack([np.asarray(o) for o in outputs])


def dump_results(outputs, path="default.pkl"):
    """What `mmcv.dump(results, args.out)` writes for a .pkl target: the stacked array, pickled."""
    if not str(path).endswith((".pkl", ".pickle")):
        raise ValueError("The output file must be a pkl file.")          # test_recognizer.py:60-61
    results = stack_results(outputs)
    with open(path, "wb") as f:
        pickle.dump(results, f)
    return results


def load_results(path):
    with open(path, "rb") as f:
        return np.asarray(pickle.load(f))


def softmax(x, dim=1):
    """accuracy.py:4-7."""
    x = np.asarray(x)
    e = np.exp(x - x.max(axis=dim, keepdims=True))
    return e / e.sum(axis=dim, keepdims=True)


def top_k_accuracy(scores, labels, k=(1,)):
    """Fraction of samples whose label (an int, or any of a collection of ints) is among the k best-scored classes
    (accuracy.py:82-101; the k best = the last k of an ascending argsort, as there)."""
    scores = np.asarray(scores)
    order = np.argsort(scores, axis=1)
    out = []
    for kk in k:
        best = order[:, -kk:]
        hits = [bool(np.isin(best[i], [y] if np.isscalar(y) or isinstance(y, (int, np.integer)) else list(y)).any())
                for i, y in enumerate(labels)]
        out.append(float(np.mean(hits)))
    return out


def mean_class_accuracy(scores, labels):
    """Mean over the classes that occur (as a label or as a prediction) of hits / samples of that class, classes without
    samples counting 0 (accuracy.py:50-69)."""
    pred = np.argmax(np.asarray(scores), axis=1).astype(np.int64)
    real = np.asarray(labels, dtype=np.int64)
    classes = np.unique(np.concatenate((pred, real)))
    per_class = []
    for c in classes:
        members = real == c
        per_class.append(float((pred[members] == c).mean()) if members.any() else 0.0)
    return float(np.mean(per_class))


def fuse_scores(score_list, coeff_list):
    """Weighted score fusion of several result sets of the same videos: sum_i coeff_i * scores_i (accuracy.py:104-123)."""
    if len(score_list) != len(coeff_list):
        raise ValueError("one coefficient per result set")
    stacked = [np.asarray(s, dtype=np.float64) for s in score_list]
    if any(s.shape != stacked[0].shape for s in stacked):
        raise ValueError("result sets must cover the same videos and classes")
    fused = np.zeros_like(stacked[0])
    for s, c in zip(stacked, coeff_list):
        fused += float(c) * s
    return fused


def fuse_result_files(paths, coeffs, labels=None, k=(1, 5)):
    """Load several default.pkl-style files, fuse them, and (with labels) score the fusion."""
    fused = fuse_scores([load_results(p) for p in paths], coeffs)
    if labels is None:
        return fused, None
    top = top_k_accuracy(fused, labels, k=k)
    return fused, dict(top_k=dict(zip(k, top)), mean_class_accuracy=mean_class_accuracy(fused, labels))
