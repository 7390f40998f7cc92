"""The north_star parity bar, per vertex: "optimised pose/plane parameters within 1e-5 relative".  Rotation entries are
O(1) => 1e-5 absolute per entry; a translation / landmark position is compared relative to ITS OWN magnitude (floor 1 m),
never to the largest coordinate of the graph."""
import numpy as np

TOL = 1e-5


def pose_errors(P, Po):
    """P, Po: (n, 3, 4) [R|t].  Returns (max rotation-entry error, max per-vertex relative translation error)."""
    P = np.asarray(P).reshape(-1, 3, 4)
    Po = np.asarray(Po).reshape(-1, 3, 4)
    rot = float(np.abs(P[:, :, :3] - Po[:, :, :3]).max()) if P.size else 0.0
    dt = np.linalg.norm(P[:, :, 3] - Po[:, :, 3], axis=1)
    scale = np.maximum(1.0, np.linalg.norm(Po[:, :, 3], axis=1))
    return rot, float((dt / scale).max()) if P.size else 0.0


def point_error(X, Xo):
    X = np.asarray(X).reshape(-1, 3)
    Xo = np.asarray(Xo).reshape(-1, 3)
    if not X.size:
        return 0.0
    return float((np.linalg.norm(X - Xo, axis=1) / np.maximum(1.0, np.linalg.norm(Xo, axis=1))).max())


def assert_parity(P, X, Po, Xo, tol=TOL, what=""):
    rot, tr = pose_errors(P, Po)
    lm = point_error(X, Xo)
    assert rot <= tol, f"{what} rotation entries differ by {rot:.3e} (> {tol})"
    assert tr <= tol, f"{what} translations differ by {tr:.3e} relative per vertex (> {tol})"
    assert lm <= tol, f"{what} landmarks differ by {lm:.3e} relative per vertex (> {tol})"
    return rot, tr, lm
