"""Synthetic sparse systems for the tests, tools and bench.py (numpy/scipy on the host; nothing here is on the solve path)."""
import numpy as np
import scipy.sparse as sp


def variable_laplacian(nx, ny, beta, neumann=True):
    """-div(beta grad) on nx x ny cells, face coefficient = harmonic mean; negative diagonal like fv::laplacian.
    All-Neumann (rows sum to zero) or Dirichlet on the left wall."""
    idx = np.arange(nx * ny).reshape(ny, nx)
    b = beta.reshape(ny, nx)
    rows, cols, vals = [], [], []
    diag = np.zeros((ny, nx))
    for (sl_a, sl_b) in (((slice(None), slice(0, nx - 1)), (slice(None), slice(1, nx))),
                         ((slice(0, ny - 1), slice(None)), (slice(1, ny), slice(None)))):
        f = 2.0 * b[sl_a] * b[sl_b] / (b[sl_a] + b[sl_b])
        rows += [idx[sl_a].ravel(), idx[sl_b].ravel()]; cols += [idx[sl_b].ravel(), idx[sl_a].ravel()]
        vals += [f.ravel(), f.ravel()]
        np.add.at(diag, sl_a, -f); np.add.at(diag, sl_b, -f)
    if not neumann:
        diag[:, 0] -= 2.0 * b[:, 0]
    rows.append(idx.ravel()); cols.append(idx.ravel()); vals.append(diag.ravel())
    A = sp.csr_matrix((np.concatenate(vals), (np.concatenate(rows), np.concatenate(cols))), shape=(nx * ny, nx * ny))
    A.sort_indices()
    return A


def beta_field(nx, ny, cx, ratio):
    """1 / rho of a bubble of radius 0.2 centred at (cx, 0.5) in the unit square, density ratio `ratio`"""
    x = (np.arange(nx) + 0.5) / nx
    y = (np.arange(ny) + 0.5) / ny
    X, Y = np.meshgrid(x, y)
    g = 0.5 * (1.0 + np.tanh((0.2 - np.hypot(X - cx, Y - 0.5)) / 0.02))
    return (1.0 / (1.0 + (ratio - 1.0) * g)).ravel()
