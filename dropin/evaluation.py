"""Drop-in shim: ``import evaluation`` resolves to the metric functions of anoddpm_b200.evaluation (reference
evaluation.py:26-87); put this directory first on sys.path."""
import os as _os
import sys as _sys

_sys.path.insert(0, _os.path.dirname(_os.path.dirname(_os.path.abspath(__file__))))
from anoddpm_b200.evaluation import *  # noqa: F401,F403,E402
from anoddpm_b200.evaluation import (AUC_score, FPR, IoU, PSNR, ROC_AUC, auc_histogram, detection_counts,  # noqa: F401,E402
                                     detection_metrics, dice_coeff, precision, recall)
