"""Host side of the Kalman-filter drop-in: `Torch_KF` with the reference's constructor, attributes and
methods (util_track/kf.py:14-428); predict / update run in the CUDA kernels of csrc/kf.cu.

What the trackers touch is kept: `Torch_KF(device, state_err, meas_err, mod_err, INIT, ADD_MEAN_Q, ADD_MEAN_R)`, the
attributes X [n,S] float32, P [n,S,S] float32, D [n], T [n] float64, obj_idxs {id: row}, the model matrices F, H, Q, R,
R2 / R3, mu_Q, mu_R, P0, and add / remove / get_dt / view / predict / update / objs with the same argument meaning
(measurement_idx 1/2/3 selects H/R, H2/R2, H3/R3).  The state lives on the GPU whatever `device` says (there is no CPU
arithmetic path).  The host part is organised differently from the reference: the model matrices are assembled by
`_model_from_scalars` / `_model_from_init`, and the id <-> row bookkeeping is a `_RowTable` (append / drop / order).
"""
import numpy as np
import torch

from . import ops

_STATE, _MEAS = 6, 5                  # (x, y, l, w, h, v) and (x, y, l, w, h) of the reference filter (kf.py:44-45)


def _dev():
    return torch.device("cuda", torch.cuda.current_device())


def _host32(t):
    return torch.as_tensor(t).detach().cpu().float()


def _model_from_scalars(state_err, meas_err, mod_err):
    """the default model of kf.py:56-69: identity dynamics, the first four state entries observed, scaled identities"""
    eye_s, eye_m = torch.eye(_STATE), torch.eye(_MEAS)
    H = torch.zeros(_MEAS, _STATE)
    H[:4, :4] = torch.eye(4)
    return dict(P0=(eye_s * state_err)[None], F=eye_s.clone(), H=H, Q=(eye_s * mod_err)[None], R=(eye_m * meas_err)[None],
                R2=(eye_m * meas_err)[None], mu_Q=torch.zeros(1, _STATE), mu_R=torch.zeros(1, _MEAS))


def _model_from_init(init, add_mean_q, add_mean_r):
    """a fitted model (kf.py:72-103): INIT holds P, F, H, Q, R, mu_Q, mu_R and optionally the second / third measurement
    models (R2, mu_R2, H2, R3, ...) and the priors mu_v, class_size, class_covariance"""
    m = dict(P0=init["P"][None], F=init["F"], H=init["H"], Q=init["Q"][None], R=init["R"][None],
             mu_Q=init["mu_Q"][None], mu_R=init["mu_R"][None])
    for suffix in ("2", "3"):
        if "R" + suffix in init:
            m["R" + suffix] = init["R" + suffix][None].float()
            m["mu_R" + suffix] = init["mu_R" + suffix][None].float()
            m["H" + suffix] = init["H" + suffix].float()
    m.update({k: init[k] for k in ("mu_v", "class_size", "class_covariance") if k in init})
    if not add_mean_q:
        m["mu_Q"] = torch.zeros(1, init["F"].shape[0])
    if not add_mean_r:
        m["mu_R"] = torch.zeros(1, init["H"].shape[0])
    return m


class _RowTable:
    """object id <-> filter row.  Rows are dense 0..n-1 in insertion order; dropping ids closes the gaps and keeps the
    order of the survivors (what kf.py:225-261 does with its keepers / new_id loop)."""

    def __init__(self):
        self.row_of = {}

    def append(self, ids, first_row):
        self.row_of.update({oid: first_row + k for k, oid in enumerate(ids)})

    def drop(self, ids, n_rows):
        """returns the boolean keep mask over the n_rows current rows and renumbers the survivors"""
        gone = {self.row_of.pop(oid) for oid in ids}
        keep = np.ones(n_rows, dtype=bool)
        keep[list(gone)] = False
        new_row = np.cumsum(keep) - 1
        self.row_of = {oid: int(new_row[r]) for oid, r in self.row_of.items()}
        return keep

    def ids_in_row_order(self):
        return [oid for oid, _ in sorted(self.row_of.items(), key=lambda kv: kv[1])]


class Torch_KF(object):
    def __init__(self, device=None, state_err=10000, meas_err=1, mod_err=1, INIT=None, ADD_MEAN_Q=False, ADD_MEAN_R=False):
        self.dt_default = 1 / 30.0
        self.device = _dev()
        self.X = self.P = self.D = self.T = None
        self._rows = _RowTable()
        model = _model_from_scalars(state_err, meas_err, mod_err) if INIT is None else _model_from_init(INIT, ADD_MEAN_Q, ADD_MEAN_R)
        for name, value in model.items():
            setattr(self, name, value)
        self.state_size, self.meas_size = self.F.shape[0], self.H.shape[0]
        for name in ("F", "H", "Q", "R", "P0", "mu_Q", "mu_R"):      # tiny: kept on the host in float32, passed by value
            setattr(self, name, _host32(getattr(self, name)))

    # the reference exposes the id -> row dictionary as an attribute (trackers read it)
    @property
    def obj_idxs(self):
        return self._rows.row_of

    def _n(self):
        return 0 if self.X is None else len(self.X)

    def _on_device(self, x, dtype=None):
        t = torch.from_numpy(x) if isinstance(x, np.ndarray) else torch.as_tensor(x)
        return t.to(self.device) if dtype is None else t.to(self.device, dtype)

    # ------------------------------------------------------------------------------------------------ time bookkeeping
    def get_dt(self, target_time, idxs=None, use_default=True):
        """seconds from each object's last time stamp T to target_time (kf.py:118-155): a float or a tensor gives one value
        per object; a list with `idxs` fills the named rows and leaves dt_default (or 0) elsewhere"""
        if self._n() == 0:
            return None
        if isinstance(target_time, list):
            target = torch.tensor(target_time, dtype=torch.float64, device=self.device)
            if idxs is None:
                return target - self.T
            base = self.dt_default if use_default else 0.0
            dt = torch.full((self._n(),), base, dtype=torch.float32, device=self.device)
            rows = torch.as_tensor(idxs, dtype=torch.int64, device=self.device)
            dt[rows] = (target - self.T[rows]).to(dt.dtype)
            return dt
        if isinstance(target_time, float):
            return target_time - self.T
        return target_time.to(self.device) - self.T

    # ------------------------------------------------------------------------------------------------ add / remove
    def add(self, detections, obj_ids, directions, times, init_speed=False, classes=None):
        """new objects (kf.py:158-222): measurements fill the leading state entries (a full state is taken as is), the
        covariance starts from P0; with `classes` the size entries and their covariance come from the class priors"""
        det = self._on_device(detections).float()
        k = len(det)
        if det.shape[1] == self.meas_size:
            rows = torch.zeros((k, self.state_size), device=self.device)
            rows[:, :self.meas_size] = det
        else:
            rows = det.clone()
        if init_speed:
            rows[:, -1] = torch.as_tensor(self.mu_v).reshape(-1)[0].to(self.device)
        cov = self.P0.to(self.device).repeat(len(obj_ids), 1, 1)
        if classes is not None:
            for i, c in enumerate(classes[:k]):
                rows[i, 2:5] = torch.as_tensor(self.class_size[c]).to(self.device)
                cov[i, 2:5, 2:5] = torch.as_tensor(self.class_covariance[c]).to(self.device)
        direction, stamp = self._on_device(directions), self._on_device(times).double()
        first = self._n()
        if first:
            self.X = torch.cat((self.X, rows)).contiguous()
            self.P = torch.cat((self.P, cov)).contiguous()
            self.D = torch.cat((self.D, direction.to(self.D.dtype)))
            self.T = torch.cat((self.T, stamp)).contiguous()
        else:
            self.X, self.P, self.D, self.T = rows.float().contiguous(), cov.float().contiguous(), direction, stamp.contiguous()
        self._rows.append(obj_ids, first)

    def remove(self, obj_ids):
        """forget objects (kf.py:225-261); the remaining rows keep their order"""
        if self.X is None:
            return
        keep = torch.from_numpy(self._rows.drop(obj_ids, self._n())).to(self.device)
        self.X, self.P = self.X[keep].contiguous(), self.P[keep].contiguous()
        self.D, self.T = self.D[keep], self.T[keep].contiguous()

    # ------------------------------------------------------------------------------------------------ read-out
    def view(self, dt=None, with_direction=False):
        """(ids in row order, states), the states advanced by dt on a copy if dt is given (kf.py:263-289); with_direction
        inserts the direction before the last (speed) column"""
        if self._n() == 0:
            return [], []
        states = self.X
        if dt is not None:
            states = self.X.clone()
            ops.kf_predict_(states, self.P.clone(), self.D, dt, self.F, self.Q[0], self.dt_default, None)
        if with_direction:
            states = torch.cat((states[:, :-1], self.D.float().unsqueeze(1), states[:, -1:]), dim=1)
        return self._rows.ids_in_row_order(), states

    def objs(self, with_direction=False, with_time=False):   # kf.py:421-428
        return self.view(dt=None, with_direction=with_direction)

    # ------------------------------------------------------------------------------------------------ the two kernels
    def predict(self, dt=None):
        """X, P one step ahead and T += dt for every object (kf.py:292-336): dt a scalar, or one value per object"""
        if self._n() == 0:
            return
        ops.kf_predict_(self.X, self.P, self.D, self.dt_default if dt is None else dt, self.F, self.Q[0], self.dt_default, self.T)

    def _measurement_model(self, measurement_idx):
        try:
            suffix = {1: "", 2: "2", 3: "3"}[measurement_idx]
        except KeyError:
            print("This measurement index does not exist in this filter")   # the reference's message (kf.py:358-360)
            raise ValueError from None
        return getattr(self, "H" + suffix), getattr(self, "R" + suffix), getattr(self, "mu_R" + suffix)

    def update(self, detections, obj_ids, measurement_idx=1):
        """measurement update of the named objects (kf.py:339-403)"""
        H, R, mu_R = self._measurement_model(measurement_idx)
        rows = torch.as_tensor([self._rows.row_of[oid] for oid in obj_ids], dtype=torch.int64, device=self.device)
        ops.kf_update_(self.X, self.P, rows, self._on_device(detections).double(), H, R[0], mu_R[0])
