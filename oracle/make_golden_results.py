#!/usr/bin/env python
"""Golden vectors for mvfnet_b200/results.py from the UNMODIFIED reference (codes/core/evaluation/accuracy.py, numpy only):
random scores / labels -> top-k, mean-class accuracy, weighted fusion.   python oracle/make_golden_results.py
Writes tests/golden/results_cases.npz (runs in the build container only: /root/reference is not on the GPU box)."""
import importlib.util
import os

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
spec = importlib.util.spec_from_file_location("ref_accuracy", "/root/reference/codes/core/evaluation/accuracy.py")
ref = importlib.util.module_from_spec(spec)
spec.loader.exec_module(ref)


def main():
    rng = np.random.RandomState(7)
    rec = {}
    for name, (n, c) in {"small": (37, 11), "k400": (64, 400)}.items():
        scores = rng.randn(n, c).astype(np.float32).astype(np.float64)     # stored as float32
        labels = rng.randint(0, c, size=n)
        labels[: n // 3] = np.argmax(scores[: n // 3], axis=1)          # some hits
        scores2 = (scores + 0.5 * rng.randn(n, c)).astype(np.float32).astype(np.float64)
        rec[name + "/scores"] = scores.astype(np.float32)
        rec[name + "/scores2"] = scores2.astype(np.float32)
        rec[name + "/labels"] = labels.astype(np.int64)
        rec[name + "/top"] = np.array(ref.top_k_accuracy(list(scores), [int(v) for v in labels], k=(1, 5)))
        rec[name + "/mca"] = np.array(ref.mean_class_accuracy(list(scores), [int(v) for v in labels]))
        fused = np.array(ref.get_weighted_score([list(scores), list(scores2)], [1.0, 0.5]))
        if name == "small":
            rec[name + "/softmax"] = ref.softmax(scores, dim=1)
            rec[name + "/fused"] = fused
        rec[name + "/fused_top"] = np.array(ref.top_k_accuracy(list(fused), [int(v) for v in labels], k=(1, 5)))
    np.savez_compressed(os.path.join(ROOT, "tests", "golden", "results_cases.npz"), **rec)
    print({k: v.shape for k, v in rec.items()})


if __name__ == "__main__":
    main()
