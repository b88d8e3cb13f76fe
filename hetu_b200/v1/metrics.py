"""v1 evaluation metrics on numpy arrays (ref: hetu/v1/python/hetu/metrics.py: confusion matrices at thresholds, ROC / PR
curves, AUC, accuracy, one-hot precision / recall / F-score)."""
from __future__ import annotations

import numpy as np


def softmax_func(y):
    y = np.asarray(y, dtype=np.float64)
    e = np.exp(y - y.max(axis=-1, keepdims=True))
    return e / e.sum(axis=-1, keepdims=True)


def confusion_matrix_at_thresholds(labels, predictions, thresholds, includes=None):
    """-> dict of arrays (one entry per threshold): tp, fn, tn, fp"""
    labels = np.asarray(labels).reshape(-1).astype(bool)
    pred = np.asarray(predictions, dtype=np.float64).reshape(-1)
    th = np.asarray(thresholds, dtype=np.float64).reshape(-1, 1)
    pos = pred[None, :] > th
    out = {"tp": (pos & labels).sum(1), "fn": (~pos & labels).sum(1), "tn": (~pos & ~labels).sum(1), "fp": (pos & ~labels).sum(1)}
    return {k: v.astype(np.float64) for k, v in out.items() if includes is None or k in includes}


def roc_pr_curve(values, curve="ROC"):
    eps = 1e-6
    if curve == "ROC":
        return values["fp"] / (values["fp"] + values["tn"] + eps), values["tp"] / (values["tp"] + values["fn"] + eps)
    rec = values["tp"] / (values["tp"] + values["fn"] + eps)
    prec = values["tp"] / (values["tp"] + values["fp"] + eps)
    return rec, prec


def auc(labels, predictions, num_thresholds=200, curve="ROC"):
    """trapezoidal area under the ROC (or PR) curve sampled at `num_thresholds` thresholds"""
    eps = 1e-7
    th = [0.0 - eps] + [(i + 1) / (num_thresholds - 1) for i in range(num_thresholds - 2)] + [1.0 + eps]
    x, y = roc_pr_curve(confusion_matrix_at_thresholds(labels, predictions, th), curve)
    return float(np.sum((x[:-1] - x[1:]) * (y[:-1] + y[1:]) / 2.0))


def accuracy(labels, predictions):
    labels, predictions = np.asarray(labels), np.asarray(predictions)
    if labels.ndim > 1 and labels.shape[-1] > 1:
        labels = labels.argmax(-1)
    if predictions.ndim > 1 and predictions.shape[-1] > 1:
        predictions = predictions.argmax(-1)
    else:
        predictions = (predictions.reshape(-1) > 0.5).astype(labels.dtype)
    return float((labels.reshape(-1) == predictions.reshape(-1)).mean())


def confusion_matrix_one_hot(labels, predictions):
    """per class: tp, fp, fn, tn from one-hot (or probability) rows"""
    y, p = np.asarray(labels).argmax(-1), np.asarray(predictions).argmax(-1)
    k = np.asarray(labels).shape[-1]
    tp = np.array([((p == c) & (y == c)).sum() for c in range(k)], dtype=np.float64)
    fp = np.array([((p == c) & (y != c)).sum() for c in range(k)], dtype=np.float64)
    fn = np.array([((p != c) & (y == c)).sum() for c in range(k)], dtype=np.float64)
    tn = len(y) - tp - fp - fn
    return tp, fp, fn, tn


def _avg(num, den, support, average):
    per = num / np.maximum(den, 1e-12)
    if average is None:
        return per
    if average == "micro":
        return float(num.sum() / max(den.sum(), 1e-12))
    if average == "macro":
        return float(per.mean())
    if average == "weighted":
        return float((per * support).sum() / max(support.sum(), 1e-12))
    raise ValueError(f"unknown average {average}")


def precision_score_one_hot(labels, predictions, average=None):
    tp, fp, fn, _ = confusion_matrix_one_hot(labels, predictions)
    return _avg(tp, tp + fp, tp + fn, average)


def recall_score_one_hot(labels, predictions, average=None):
    tp, fp, fn, _ = confusion_matrix_one_hot(labels, predictions)
    return _avg(tp, tp + fn, tp + fn, average)


def f_score_one_hot(labels, predictions, beta=1.0, average=None):
    tp, fp, fn, _ = confusion_matrix_one_hot(labels, predictions)
    b2 = beta * beta
    if average == "micro":
        return float((1 + b2) * tp.sum() / max((1 + b2) * tp.sum() + b2 * fn.sum() + fp.sum(), 1e-12))
    per = (1 + b2) * tp / np.maximum((1 + b2) * tp + b2 * fn + fp, 1e-12)
    if average is None:
        return per
    sup = tp + fn
    return float(per.mean()) if average == "macro" else float((per * sup).sum() / max(sup.sum(), 1e-12))
