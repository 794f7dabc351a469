"""TEST INFRASTRUCTURE ONLY -- NumPy restatement of the reference's estimators (estimators/package.scala:18-137).

Pinning: the reference holds no value-level golden vectors for the estimators (its specs only assert thresholds on trained
models, e.g. optimizers/AdamSpec.scala:61 `accuracy(...) should be >= 0.9f`), so these are "parity unpinned" restatements;
the logistic-regression threshold is re-checked in tests/test_gpu_parity.py::test_adam_logistic_regression_spec."""
import numpy as np

from .layers import model_forward
from .train import tensor_iterator

__all__ = ["accuracy", "mse", "rmse", "mae", "r2_score"]


def _result(model, x, params):
    return model_forward(model, x, params, training=False)[0]  # TrainedModel.buildResult (models/Model.scala:154-163)


def accuracy(model, params, x_rows, y_rows, batch=1000):
    """package.scala:18-43: per batch, yPredicted = result.round; a row counts when all outputs match."""
    pos, total = 0, 0
    for x, y in tensor_iterator(x_rows, y_rows, batch, remaining="partial"):
        pred = np.round(_result(model, x, params))
        assert pred.ndim == 2, "rank-2 model outputs (the estimator sums matches over axis 1 only)"
        matched = (y.reshape(pred.shape) == pred).astype(np.int32).sum(axis=1)
        pos += int((matched == pred.shape[-1]).sum())
        total += x.shape[0]
    return float(np.float32(pos) / np.float32(total))


def _mean_error(model, params, x_rows, y_rows, batch, err):
    """package.scala:67-97: Float sum of per-batch error sums / rows."""
    s, total = np.float32(0), 0
    for x, y in tensor_iterator(x_rows, y_rows, batch, remaining="partial"):
        pred = _result(model, x, params)
        s = np.float32(s + np.float32(err(pred, y.reshape(pred.shape)).sum(dtype=np.float32)))
        total += x.shape[0]
    return float(np.float32(s) / np.float32(total))


def mse(model, params, x_rows, y_rows, batch=1000):  # package.scala:51-57
    return _mean_error(model, params, x_rows, y_rows, batch, lambda p, e: (p - e) ** 2)


def rmse(model, params, x_rows, y_rows, batch=1000):  # package.scala:45-49
    return float(np.float32(np.sqrt(mse(model, params, x_rows, y_rows, batch))))


def mae(model, params, x_rows, y_rows, batch=1000):  # package.scala:59-65
    return _mean_error(model, params, x_rows, y_rows, batch, lambda p, e: np.abs(p - e))


def r2_score(model, params, x_rows, y_rows, batch=1000):
    """package.scala:99-137.  The pairs are (expected, predicted) and the reference centres SST on `map(_._2)`, i.e. on the
    mean of the PREDICTIONS (sic); kept."""
    if tuple(y_rows.shape[1:]) != (1,):
        raise ValueError("requirement failed: labels should have shape (1)")
    exp, pred = [], []
    for x, y in tensor_iterator(x_rows, y_rows, batch, remaining="partial"):
        p = _result(model, x, params)
        exp.append(y[:, 0].astype(np.float32)); pred.append(p.reshape(p.shape[0], -1)[:, 0].astype(np.float32))
    exp, pred = np.concatenate(exp).astype(np.float64), np.concatenate(pred).astype(np.float64)
    mean = pred.mean()
    ssr, sst = ((pred - exp) ** 2).sum(), ((pred - mean) ** 2).sum()
    return float(np.float32(1 - ssr / sst))
