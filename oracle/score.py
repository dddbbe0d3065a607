"""Oracle: confusion matrix + score() measures (TEST INFRASTRUCTURE ONLY).

Follows xview/models/base_model.py:140-151 (confusion-matrix graph) and :306-329
(host-side float64 measures).  PINNED: reproduces the stored measures of sacred run 868
(`Experimental Details.ipynb` cell 12) - tests/test_oracle_golden.py.
"""
import numpy as np


def confusion_matrix(labels, prediction, num_classes):
    """base_model.py:140-151: labels < 0 go to an extra class that is sliced away; rows =
    ground-truth label, cols = prediction (tf.confusion_matrix(labels, predictions))."""
    labels = np.asarray(labels).reshape(-1).astype(np.int64)
    prediction = np.asarray(prediction).reshape(-1).astype(np.int64)
    labels = np.where(labels < 0, num_classes, labels)
    # tf.confusion_matrix with num_classes+1 requires every index < num_classes+1
    keep = (labels <= num_classes) & (prediction <= num_classes) & (prediction >= 0)
    idx = labels[keep] * (num_classes + 1) + prediction[keep]
    cm = np.bincount(idx, minlength=(num_classes + 1) ** 2).reshape(num_classes + 1,
                                                                    num_classes + 1)
    return cm[:num_classes, :num_classes].astype(np.int64)


def score_measures(cm):
    """base_model.py:315-329 verbatim semantics on a float64 matrix."""
    cm = np.asarray(cm, np.float64)
    with np.errstate(divide='ignore', invalid='ignore'):
        m = {}
        m['confusion_matrix'] = cm
        m['recall'] = np.diag(cm) / cm.sum(1)
        m['precision'] = np.diag(cm) / cm.sum(0)
        m['F1'] = 2 * m['precision'] * m['recall'] / (m['precision'] + m['recall'])
        m['mean_F1'] = np.nanmean(m['F1'])
        m['total_accuracy'] = np.diag(cm)[1:].sum() / cm[1:, :].sum()
        m['IoU'] = np.diag(cm) / (cm.sum(1) + cm.sum(0) - np.diag(cm))
        m['mean_IoU'] = np.nanmean(m['IoU'][1:])
    return m
