"""Evaluation statistics of the reference driver: per-database and overall Pearson
correlation, RMSE, RMSE after a first/second/third-order mapping and the ITU-T P.1401 RMSE* on
per-file and per-condition level - what ``nisqaModel.evaluate()`` prints after ``predict()`` in the
reference's ``run_evaluate.py``.

Mirrors reference nisqa/NISQA_lib.py:1469-1852 (``is_const``, ``calc_eval_metrics``, ``calc_rmse``,
``calc_rmse_star``, ``calc_mapped``, ``fit_*``, ``calc_mapping``, ``eval_results``): same function
names, arguments, returned keys / DataFrame columns and printed lines, so that ``run_evaluate.py``
reads the same.  This is host arithmetic on one number per file (at most a few 10^5 rows): it runs in
NumPy on the host and is not part of the GPU hot path.  ``tests/test_evaluate.py`` holds it to golden
results produced by the reference's own functions (``oracle/make_eval_golden.py``).
"""
import numpy as np
import pandas as pd
from scipy.optimize import minimize
from scipy.stats import pearsonr

_PER_FILE_NAN_KEYS = ("r_p", "r_s", "rmse", "r_p_map", "r_s_map", "rmse_map")
_PER_CON_NAN_KEYS = _PER_FILE_NAN_KEYS + ("rmse_star_map",)


def is_const(x):
    """lib:1469-1475."""
    x = np.asarray(x)
    m = np.mean(x)
    return bool(np.linalg.norm(x - m) < 1e-13 * np.abs(m) or np.all(x == x[0]))


def calc_rmse(y_true, y_pred, d=0):
    """lib:1499-1508; d = degrees of freedom of the mapping, Eq (7-29) of P.1401."""
    sq = np.square(y_true - y_pred)
    if d == 0:
        return np.sqrt(np.mean(sq))
    n = y_true.shape[0]
    if n - d < 1:
        return np.nan
    return np.sqrt(1 / (n - d) * np.sum(sq))


def calc_rmse_star(mos_sub, mos_obj, ci, d):
    """lib:1510-1524: epsilon-insensitive RMSE, Eq (7-27) / (7-29) of P.1401."""
    n = mos_sub.shape[0]
    error = mos_sub - mos_obj
    if np.isnan(ci).any():
        return np.nan, np.nan, error
    p_error = (abs(error) - ci).clip(min=0)
    if n - d < 1:
        return np.nan, p_error, error
    return np.sqrt(1 / (n - d) * sum(p_error ** 2)), p_error, error


def calc_eval_metrics(y, y_hat, y_hat_map=None, d=None, ci=None):
    """lib:1477-1497."""
    r = dict(r_p=np.nan, rmse=np.nan, rmse_map=np.nan, rmse_star_map=np.nan)
    if not (is_const(y_hat) or any(np.isnan(y))):
        r["r_p"] = pearsonr(y, y_hat)[0]
    r["rmse"] = calc_rmse(y, y_hat)
    if y_hat_map is not None:
        r["rmse_map"] = calc_rmse(y, y_hat_map, d=d)
        if ci is not None:
            r["rmse_star_map"] = calc_rmse_star(y, y_hat_map, ci, d)[0]
    return r


def calc_mapped(x, b):
    """lib:1526-1532: sum_i b[i] * x**i."""
    x = np.asarray(x, dtype=np.float64)
    return np.vander(x, N=np.asarray(b).shape[0], increasing=True) @ b


def _fit_poly(y, y_hat, order):
    A = np.vander(np.asarray(y_hat, dtype=np.float64), N=order + 1, increasing=True)
    return np.linalg.lstsq(A, y, rcond=None)[0]


def fit_first_order(y_con, y_con_hat):
    return _fit_poly(y_con, y_con_hat, 1)


def fit_second_order(y_con, y_con_hat):
    return _fit_poly(y_con, y_con_hat, 2)


def fit_third_order(y_con, y_con_hat):
    """lib:1544-1555: unconstrained cubic; reports when it is not monotonic over the data range."""
    b = _fit_poly(y_con, y_con_hat, 3)
    rr = np.roots(np.polyder(np.poly1d(np.flipud(b))))
    r = rr[np.imag(rr) == 0]
    if not all(np.logical_or(r > max(y_con_hat), r < min(y_con_hat))):
        print("Not monotonic!!!")
    return b


def fit_monotonic_third_order(dfile_db, dcon_db=None, pred=None, target_mos=None, target_ci=None, mapping=None):
    """lib:1557-1640: cubic constrained to a non-negative first derivative on the prediction range
    (SLSQP from the identity mapping); per condition when ``dcon_db`` is given."""
    y = dfile_db[target_mos].to_numpy()
    x = dfile_db[pred].to_numpy()
    src = dfile_db if dcon_db is None else dcon_db
    ci = src[target_ci].to_numpy() if target_ci in src else 0
    y_con = None if dcon_db is None else dcon_db[target_mos].to_numpy()
    grid = np.arange(min(x) - 0.01, max(x) + 0.01, 0.1)
    con_codes = None if dcon_db is None else dfile_db["con"]

    def poly(p, t):
        return p[0] + p[1] * t + p[2] * t ** 2 + p[3] * t ** 3

    def loss(err):
        if mapping == "pError":
            return ((abs(err) - ci).clip(min=0) ** 2).sum()
        if mapping == "error":
            return (err ** 2).sum()
        raise NotImplementedError

    def objective(p):
        x_map = poly(p, x)
        if dcon_db is None:
            return loss(x_map - y)
        return loss(pd.Series(x_map, index=dfile_db.index).groupby(con_codes).mean().to_numpy() - y_con)

    res = minimize(objective, x0=np.array([0.0, 1.0, 0.0, 0.0]), method="SLSQP",
                   constraints=dict(type="ineq", fun=lambda p: p[1] + 2 * p[2] * grid + 3 * p[3] * grid ** 2))
    return res.x


def calc_mapping(dfile_db, mapping=None, dcon_db=None, target_mos=None, target_ci=None, pred=None):
    """lib:1642-1685: coefficients b (increasing powers) and degrees of freedom d_map."""
    if dcon_db is not None:
        y = dcon_db[target_mos].to_numpy()
        y_hat = dfile_db.groupby("con")[pred].mean().to_numpy()
    else:
        y = dfile_db[target_mos].to_numpy()
        y_hat = dfile_db[pred].to_numpy()
    if mapping is None:
        return np.array([0, 1, 0, 0]), 0
    if mapping == "first_order":
        return fit_first_order(y, y_hat), 1
    if mapping == "second_order":
        return fit_second_order(y, y_hat), 3
    if mapping == "third_order_not_monotonic":
        return fit_third_order(y, y_hat), 4
    if mapping == "third_order":
        return fit_monotonic_third_order(dfile_db, dcon_db=dcon_db, pred=pred, target_mos=target_mos,
                                         target_ci=target_ci, mapping="error"), 4
    raise NotImplementedError


def _scatter(db_name, level, x, y, b, target_mos, marker):
    try:
        import matplotlib.pyplot as plt
    except ImportError:           # plots are optional; the numbers do not depend on them
        return
    xx = np.arange(0, 6, 0.01)
    plt.figure(figsize=(3.0, 3.0), dpi=300)
    plt.clf()
    plt.plot(x, y, "o", label="Original data", markersize=marker)
    plt.plot([0, 5], [0, 5], "gray")
    plt.plot(xx, calc_mapped(xx, b), "r", label="Fitted line")
    plt.axis([1, 5, 1, 5])
    plt.gca().set_aspect("equal", adjustable="box")
    plt.grid(True)
    plt.xticks(np.arange(1, 6))
    plt.yticks(np.arange(1, 6))
    plt.title(db_name + " per " + level)
    plt.ylabel(("Subjective " if level == "file" else "Sub ") + target_mos.upper())
    plt.xlabel(("Predicted " if level == "file" else "Pred ") + target_mos.upper())
    plt.show()


def eval_results(df, dcon=None, target_mos="mos", target_ci="mos_ci", pred="mos_pred", mapping=None,
                 do_print=False, do_plot=False):
    """lib:1687-1852.  Returns (per-database DataFrame, dict of overall results) and writes the
    condition-level mapped predictions into ``df['y_hat_map']`` like the reference."""
    rows = []
    df["y_hat_map"] = np.nan
    dcon_db = None
    for db_name in df.db.astype("category").cat.categories:
        sel = df.db == db_name
        df_db = df.loc[sel]
        dcon_db = dcon.loc[dcon.db == db_name] if dcon is not None else None
        has_con = dcon_db is not None and "con" in df_db

        # ---- per file
        y = df_db[target_mos].to_numpy()
        y_hat = df_db[pred].to_numpy()
        b = None
        if np.isnan(y).any():
            r = dict.fromkeys(_PER_FILE_NAN_KEYS, np.nan)
        else:
            b, d = calc_mapping(df_db, mapping=mapping, target_mos=target_mos, target_ci=target_ci, pred=pred)
            r = calc_eval_metrics(y, y_hat, y_hat_map=calc_mapped(y_hat, b), d=d)
            r.pop("rmse_star_map")
        res = {k + "_file": v for k, v in r.items()}

        # ---- per condition
        r_con = dict.fromkeys(_PER_CON_NAN_KEYS, np.nan)
        y_con = y_con_hat = b_con = None
        if has_con:
            y_con = dcon_db[target_mos].to_numpy()
            y_con_hat = df_db.groupby("con")[pred].mean().to_numpy()
            if not np.isnan(y_con).any():
                ci_con = dcon_db[target_ci].to_numpy() if target_ci in dcon_db else None
                b_con, d = calc_mapping(df_db, dcon_db=dcon_db, mapping=mapping, target_mos=target_mos,
                                        target_ci=target_ci, pred=pred)
                mapped = pd.Series(calc_mapped(y_hat, b_con), index=df_db.index)
                df.loc[sel, "y_hat_map"] = mapped
                r_con = calc_eval_metrics(y_con, y_con_hat, y_hat_map=mapped.groupby(df_db["con"]).mean().to_numpy(),
                                          d=d, ci=ci_con)
        res.update({k + "_con": v for k, v in r_con.items()})
        rows.append({"db": db_name, **res})

        if not np.isnan(y).any():
            if do_plot:
                _scatter(db_name, "file", y_hat, y, b, target_mos, 2)
                if has_con and b_con is not None:
                    _scatter(db_name, "con", y_con_hat, y_con, b_con, target_mos, 3)
            if do_print:
                if has_con:
                    print("%-30s r_p_file: %0.2f, rmse_map_file: %0.2f, r_p_con: %0.2f, rmse_map_con: %0.2f, rmse_star_map_con: %0.2f"
                          % (db_name + ":", res["r_p_file"], res["rmse_map_file"], res["r_p_con"], res["rmse_map_con"],
                             res["rmse_star_map_con"]))
                else:
                    print("%-30s r_p_file: %0.2f, rmse_map_file: %0.2f" % (db_name + ":", res["r_p_file"], res["rmse_map_file"]))

    db_results_df = pd.DataFrame(rows)
    overall = {}
    y_all, y_hat_all = df[target_mos].to_numpy(), df[pred].to_numpy()
    tot = calc_eval_metrics(y_all, y_hat_all)
    overall["r_p_all"], overall["rmse_all"] = tot["r_p"], tot["rmse"]
    for key in ("r_p", "rmse", "rmse_map"):
        overall[key + "_mean_file"] = db_results_df[key + "_file"].mean()
    for key in ("r_p", "rmse", "rmse_map", "rmse_star_map"):
        # the reference keys this on the LAST database having a condition table (lib:1829)
        overall[key + "_mean_con"] = db_results_df[key + "_con"].mean() if dcon_db is not None else np.nan
    return db_results_df, overall
