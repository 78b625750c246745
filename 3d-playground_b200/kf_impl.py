"""Host side of the Kalman-filter drop-in: Torch_KF with the reference's constructor, attributes and
methods (util_track/kf.py:14-428), predict / update executed by the CUDA kernels of csrc/kf.cu.

Kept from the reference: `Torch_KF(device, state_err, meas_err, mod_err, INIT, ADD_MEAN_Q, ADD_MEAN_R)`, the attributes
X [n,S] float32, P [n,S,S] float32, D [n], T [n] float64, obj_idxs {id: row}, the model matrices F, H, Q, R, R2/R3,
mu_Q, mu_R, P0, and the methods add / remove / get_dt / view / predict / update / objs with the same argument meaning
(measurement_idx 1/2/3 selects H/R, H2/R2, H3/R3).  The state lives on the GPU whatever `device` says (there is no CPU
arithmetic path); id bookkeeping stays on the host exactly as in the reference.
"""
import numpy as np
import torch

from . import ops


def _dev():
    return torch.device("cuda", torch.cuda.current_device())


class Torch_KF(object):
    def __init__(self, device=None, state_err=10000, meas_err=1, mod_err=1, INIT=None, ADD_MEAN_Q=False, ADD_MEAN_R=False):
        self.meas_size, self.state_size = 5, 6
        self.dt_default = 1 / 30.0
        self.device = _dev()
        self.X = self.D = self.T = self.P = None
        self.obj_idxs = {}
        if INIT is None:   # kf.py:56-69
            self.P0 = torch.eye(self.state_size).unsqueeze(0) * state_err
            self.F = torch.eye(self.state_size).float()
            self.H = torch.zeros(self.meas_size, self.state_size)
            self.H[:4, :4] = torch.eye(4)
            self.Q = torch.eye(self.state_size).unsqueeze(0) * mod_err
            self.R = torch.eye(self.meas_size).unsqueeze(0) * meas_err
            self.R2 = torch.eye(self.meas_size).unsqueeze(0) * meas_err
            self.mu_Q = torch.zeros([1, self.state_size])
            self.mu_R = torch.zeros([1, self.meas_size])
        else:              # kf.py:72-103
            self.P0 = INIT["P"].unsqueeze(0)
            self.F, self.H = INIT["F"], INIT["H"]
            self.Q, self.R = INIT["Q"].unsqueeze(0), INIT["R"].unsqueeze(0)
            self.mu_Q, self.mu_R = INIT["mu_Q"].unsqueeze(0), INIT["mu_R"].unsqueeze(0)
            for k in ("2", "3"):
                if "R" + k in INIT:
                    setattr(self, "R" + k, INIT["R" + k].unsqueeze(0).float())
                    setattr(self, "mu_R" + k, INIT["mu_R" + k].unsqueeze(0).float())
                    setattr(self, "H" + k, INIT["H" + k].float())
            for k in ("mu_v", "class_size", "class_covariance"):
                if k in INIT:
                    setattr(self, k, INIT[k])
            self.state_size, self.meas_size = self.F.shape[0], self.H.shape[0]
            if not ADD_MEAN_Q:
                self.mu_Q = torch.zeros([1, self.state_size])
            if not ADD_MEAN_R:
                self.mu_R = torch.zeros([1, self.meas_size])
        for k in ("F", "H", "Q", "R", "P0", "mu_Q", "mu_R"):   # model matrices are tiny: kept on the host, float32
            setattr(self, k, getattr(self, k).detach().cpu().float())

    # ---- bookkeeping (host logic of the reference, tensors on the GPU)
    def _t(self, x, dtype=None):
        t = torch.from_numpy(x) if isinstance(x, np.ndarray) else torch.as_tensor(x)
        return t.to(self.device) if dtype is None else t.to(self.device, dtype)

    def get_dt(self, target_time, idxs=None, use_default=True):   # kf.py:118-155
        if self.X is None or len(self.X) == 0:
            return None
        if type(target_time) == float:
            return target_time - self.T
        if type(target_time) == list:
            target_time = torch.tensor(target_time, dtype=torch.double, device=self.device)
            if idxs is None:
                return target_time - self.T
            dt = torch.zeros(len(self.X), device=self.device)
            dt = dt + self.dt_default if use_default else dt
            ii = torch.as_tensor(idxs, dtype=torch.int64, device=self.device)
            dt[ii] = (target_time - self.T[ii]).to(dt.dtype)
            return dt
        return target_time.to(self.device) - self.T

    def add(self, detections, obj_ids, directions, times, init_speed=False, classes=None):   # kf.py:158-222
        det = self._t(detections).float()
        newX = torch.zeros((len(det), self.state_size), device=self.device)
        if det.shape[1] == self.meas_size:
            newX[:, :self.meas_size] = det
        else:
            newX = det.clone()
        newD, newT = self._t(directions), self._t(times)
        if init_speed:
            newX[:, -1] = torch.as_tensor(self.mu_v).reshape(-1)[0].to(self.device)
        newP = self.P0.to(self.device).repeat(len(obj_ids), 1, 1)
        if classes is not None:
            for i in range(len(newX)):
                newX[i, 2:5] = torch.as_tensor(self.class_size[classes[i]]).to(self.device)
                newP[i, 2:5, 2:5] = torch.as_tensor(self.class_covariance[classes[i]]).to(self.device)
        if self.X is not None and len(self.X) > 0:
            new_idx = len(self.X)
            self.X = torch.cat((self.X, newX), dim=0).contiguous()
            self.P = torch.cat((self.P, newP), dim=0).contiguous()
            self.D = torch.cat((self.D, newD.to(self.D.dtype)), dim=0)
            self.T = torch.cat((self.T, newT.double()), dim=0).contiguous()
        else:
            new_idx = 0
            self.X, self.P = newX.float().contiguous(), newP.float().contiguous()
            self.D, self.T = newD, newT.double().contiguous()
        for idx, id in enumerate(obj_ids):
            self.obj_idxs[id] = new_idx + idx

    def remove(self, obj_ids):   # kf.py:225-261
        if self.X is None:
            return
        keepers = list(range(len(self.X)))
        for id in obj_ids:
            keepers.remove(self.obj_idxs[id])
            self.obj_idxs[id] = None
        k = torch.as_tensor(keepers, dtype=torch.int64, device=self.device)
        self.X, self.P = self.X[k].contiguous(), self.P[k].contiguous()
        self.D, self.T = self.D[k], self.T[k].contiguous()
        new_id, removals = 0, []
        for id in self.obj_idxs:
            if self.obj_idxs[id] is not None:
                self.obj_idxs[id] = new_id
                new_id += 1
            else:
                removals.append(id)
        for id in removals:
            del self.obj_idxs[id]

    def view(self, dt=None, with_direction=False):   # kf.py:263-289: predict() on a copy, states only
        if self.X is None or len(self.X) == 0:
            return [], []
        states = self.X
        if dt is not None:
            states = self.X.clone()
            ops.kf_predict_(states, self.P.clone(), self.D, dt, self.F, self.Q[0], self.dt_default, None)
        inverted = dict([(self.obj_idxs[key], key) for key in self.obj_idxs.keys()])
        id_list = [inverted[i] for i in range(states.shape[0])]
        if with_direction:
            states = torch.cat((states[:, :-1], self.D.float().unsqueeze(1), states[:, -1:]), dim=1)
        return id_list, states

    def objs(self, with_direction=False, with_time=False):   # kf.py:421-428
        return self.view(dt=None, with_direction=with_direction)

    # ---- the two kernels
    def predict(self, dt=None):   # kf.py:292-336
        if self.X is None or len(self.X) == 0:
            return
        if dt is None:
            dt = self.dt_default
        ops.kf_predict_(self.X, self.P, self.D, dt, self.F, self.Q[0], self.dt_default, self.T)

    def update(self, detections, obj_ids, measurement_idx=1):   # kf.py:339-403
        if measurement_idx == 1:
            mu_R, H, R = self.mu_R, self.H, self.R
        elif measurement_idx == 2:
            mu_R, R, H = self.mu_R2, self.R2, self.H2
        elif measurement_idx == 3:
            mu_R, R, H = self.mu_R3, self.R3, self.H3
        else:
            print("This measurement index does not exist in this filter")
            raise ValueError
        relevant = torch.as_tensor([self.obj_idxs[id] for id in obj_ids], dtype=torch.int64, device=self.device)
        z = self._t(detections).double()
        ops.kf_update_(self.X, self.P, relevant, z, H, R[0], mu_R[0])
