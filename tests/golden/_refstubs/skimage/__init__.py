"""Import-only stand-in for scikit-image (reference evaluation.py imports skimage.metrics)."""
import sys
import types

metrics = types.ModuleType("skimage.metrics")


def structural_similarity(*a, **k):
    raise NotImplementedError("skimage stub")


metrics.structural_similarity = structural_similarity
sys.modules["skimage.metrics"] = metrics
