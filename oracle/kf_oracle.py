"""ORACLE (test infrastructure, not product code): CPU restatement of Torch_KF.predict / update / view of
util_track/kf.py, with plain PyTorch CPU ops in the reference's operation order.

    predict   kf.py:292-336   (F_rep[:,0,5] = D * dt; X = F_rep X; P = F_rep P F_rep^T + Q * dt / dt_default)
    update    kf.py:339-403   (y = z + mu_R - X H^T; S = H P H^T + R; K = P H^T S^-1; X += K y; P = (I - K H) P)

Parity pin: tests/test_oracle_golden.py checks these against tests/golden/kf.npz, produced by the UNMODIFIED reference
class (tests/golden/make_golden.py imports util_track.kf with a matplotlib stub).
Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline legs may import this module.
"""
import torch


def predict(X, P, D, dt, F, Q, dt_default=1 / 30.0):
    """returns (X', P').  dt: python float or float64 tensor [n] (kf.py:308-333)."""
    n = len(X)
    F_rep = F.unsqueeze(0).repeat(n, 1, 1)
    F_rep[:, 0, 5] = D * dt
    Xn = torch.bmm(F_rep, X.unsqueeze(2)).squeeze(2)
    step3 = torch.bmm(torch.bmm(F_rep, P.float()), F_rep.transpose(1, 2))
    step4 = Q.reshape(1, *Q.shape[-2:]).repeat(n, 1, 1)
    if isinstance(dt, torch.Tensor):
        step4 = step4 * dt.unsqueeze(1).unsqueeze(2).repeat(1, Q.shape[-2], Q.shape[-1]) / dt_default
    else:
        step4 = step4 * dt / dt_default
    return Xn, (step3 + step4).float()


def update(X, P, rows, z, H, R, mu_R):
    """returns updated copies of (X, P); rows: list / int64 tensor of object rows, z: [m, M] (kf.py:369-403)."""
    X, P = X.clone(), P.clone()
    X_up, P_up = X[rows, :], P[rows, :, :].float()
    y = z.double() + mu_R.reshape(1, -1) - torch.mm(X_up, H.transpose(0, 1))
    H_rep = H.unsqueeze(0).repeat(len(P_up), 1, 1)
    S = torch.bmm(torch.bmm(H_rep, P_up), H_rep.transpose(1, 2)) + R.reshape(1, *R.shape[-2:]).repeat(len(P_up), 1, 1)
    K = torch.bmm(torch.bmm(P_up, H_rep.transpose(1, 2)), S.inverse())
    X_up = X_up + torch.bmm(K, y.unsqueeze(-1).float()).squeeze(-1)
    I = torch.eye(X.shape[1]).unsqueeze(0).repeat(len(P_up), 1, 1)
    P_up = torch.bmm(I - torch.bmm(K, H_rep), P_up)
    X[rows, :] = X_up
    P[rows, :, :] = P_up
    return X, P
