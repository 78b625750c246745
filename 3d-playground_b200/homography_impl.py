"""Host side of the homography drop-in: Homography / Homography_Wrapper with the reference's method surface
(homography.py:156-748, :793-901), every transform executed by the CUDA kernels of csrc/homography.cu.

What is kept from the reference: class and method names, argument meaning (`name` = None | str | list[str], plus a
uint8/int tensor of camera indices as the fast form of the list), output shapes and dtypes (float64 image points,
float32 states), the `correspondence[name] = {"H", "H_inv", "P", ...}` dictionary of numpy float64 matrices (so pickled
reference objects load into these classes), `default_correspondence`, `class_heights` / `guess_heights`.
What is out of scope (calibration-time fitting, cv2 plotting): add_i24_camera, find_vanishing_point, scale_Z, plot_*.
Matrices can be supplied with add_correspondence_matrices().

Tensors on the CPU are accepted for convenience (the trackers keep their state on the host): they are copied to the
GPU, transformed by the kernel and the result is returned on the caller's device.  There is no CPU arithmetic path.
"""
import numpy as np
import torch

from . import ops

_CLASS_HEIGHTS = {"sedan": 4, "midsize": 5, "van": 6, "pickup": 5, "semi": 12, "truck (other)": 12, "truck": 12,
                  "motorcycle": 4, "trailer": 3, "other": 5}
_CLASS_DIMS = {"sedan": [16, 6, 4], "midsize": [18, 6.5, 5], "van": [20, 6, 6.5], "pickup": [20, 6, 5],
               "semi": [55, 9, 12], "truck (other)": [25, 9, 12], "truck": [25, 9, 12], "motorcycle": [7, 3, 4],
               "trailer": [16, 7, 3], "other": [18, 6.5, 5]}
_CLASS_NAMES = ["sedan", "midsize", "van", "pickup", "semi", "truck (other)", "motorcycle", "trailer"]


def _class_dict():
    d = {n: i for i, n in enumerate(_CLASS_NAMES)}
    d["truck"] = 5
    d.update({i: n for i, n in enumerate(_CLASS_NAMES)})
    return d


def _default_device():
    return torch.device("cuda", torch.cuda.current_device())


class _MatrixBank:
    """Device copies of the per-camera matrices of one or two Homography objects: P[ncam,2,3,4], H[ncam,2,3,3]."""

    def __init__(self, hg1, hg2=None):
        self.names = list(hg1.correspondence.keys())
        self.index = {n: i for i, n in enumerate(self.names)}
        hgs = (hg1, hg2 if hg2 is not None else hg1)
        n = max(len(self.names), 1)
        P = np.zeros((n, 2, 3, 4), dtype=np.float64)
        H = np.zeros((n, 2, 3, 3), dtype=np.float64)
        for i, name in enumerate(self.names):
            for j, hg in enumerate(hgs):
                corr = hg.correspondence.get(name, hg1.correspondence[name])
                P[i, j] = np.asarray(corr["P"], dtype=np.float64).reshape(3, 4)
                H[i, j] = np.asarray(corr["H"], dtype=np.float64).reshape(3, 3)
        self.P_host, self.H_host = P, H
        self._dev = {}
        self.signature = _signature(hg1, hg2)

    def on(self, device):
        key = str(device)
        if key not in self._dev:
            self._dev[key] = (torch.from_numpy(self.P_host).to(device), torch.from_numpy(self.H_host).to(device))
        return self._dev[key]


def _signature(hg1, hg2):
    sig = []
    for hg in (hg1, hg2):
        if hg is None:
            continue
        for name, corr in hg.correspondence.items():
            sig.append((name, id(corr.get("P")), id(corr.get("H"))))
    return tuple(sig)


def _bank_for(owner, hg1, hg2=None):
    bank = owner.__dict__.get("_g3d_bank")
    if bank is None or bank.signature != _signature(hg1, hg2):
        bank = _MatrixBank(hg1, hg2)
        owner.__dict__["_g3d_bank"] = bank
    return bank


def _camera_arg(bank, name, default, d, device):
    """name: None | str | int | list[str] | tensor of camera indices -> (cam argument for ops.*)"""
    if name is None:
        name = default
    if isinstance(name, str):
        if name not in bank.index:
            raise KeyError(name)
        return bank.index[name]
    if isinstance(name, (int, np.integer)):
        return int(name)
    if isinstance(name, torch.Tensor):
        return name.to(device=device, dtype=torch.uint8)
    if isinstance(name, (list, tuple, np.ndarray)):
        if len(name) != d:
            raise ValueError(f"{len(name)} camera names for {d} objects")
        idx = np.fromiter((bank.index[n] for n in name), dtype=np.uint8, count=len(name))
        return torch.from_numpy(idx).to(device)
    raise TypeError(f"unsupported camera name type {type(name).__name__}")


def _to_dev(t, device):
    if not isinstance(t, torch.Tensor):
        t = torch.as_tensor(t)
    return t if t.is_cuda else t.to(device)


def _ret(t, like):
    """results go back to the device the caller's tensor lives on"""
    if isinstance(like, torch.Tensor) and not like.is_cuda:
        return t.cpu()
    return t


def _exec_device(*tensors):
    for t in tensors:
        if isinstance(t, torch.Tensor) and t.is_cuda:
            return t.device
    return _default_device()


class Homography:
    """homography.py:156-748 (transform methods).  One space/state formulation, many camera correspondences."""

    def __init__(self, f1=None, f2=None):
        if f1 is not None:
            self.f1, self.f2 = f1, f2
        self.correspondence = {}
        self.class_heights = dict(_CLASS_HEIGHTS)
        self.class_dims = {k: list(v) for k, v in _CLASS_DIMS.items()}
        self.class_dict = _class_dict()
        self.default_correspondence = None

    # ---- correspondences
    def add_correspondence_matrices(self, name, H, P, H_inv=None, **extra):
        """Register a camera from already-fitted matrices (the fitting itself, homography.py:336-377, is out of scope)."""
        H = np.asarray(H, dtype=np.float64).reshape(3, 3)
        P = np.asarray(P, dtype=np.float64).reshape(3, 4)
        entry = {"H": H, "H_inv": np.linalg.inv(H) if H_inv is None else np.asarray(H_inv, dtype=np.float64), "P": P}
        entry.update(extra)
        self.correspondence[name] = entry
        if self.default_correspondence is None:
            self.default_correspondence = name

    def remove_correspondence(self, name):
        try:
            del self.correspondence[name]
            print("Deleted correspondence for {}".format(name))
        except KeyError:
            print("Tried to delete correspondence {}, but this does not exist".format(name))

    def _custom_space(self):
        return "f1" in self.__dict__ or "f2" in self.__dict__

    # ---- state <-> space
    def i24_state_to_space(self, points):
        dev = _exec_device(points)
        return _ret(ops.state_to_space(_to_dev(points, dev)), points)

    def i24_space_to_state(self, points):
        dev = _exec_device(points)
        return _ret(ops.space_to_state(_to_dev(points, dev)), points)

    def state_to_space(self, points):
        return self.f2(points) if "f2" in self.__dict__ else self.i24_state_to_space(points)

    def space_to_state(self, points):
        return self.f1(points) if "f1" in self.__dict__ else self.i24_space_to_state(points)

    # ---- space / state <-> image
    def _bank(self):
        return _bank_for(self, self)

    def space_to_im(self, points, name=None):
        dev = _exec_device(points)
        bank = self._bank()
        P, _ = bank.on(dev)
        cam = _camera_arg(bank, name, self.default_correspondence, points.shape[0], dev)
        return _ret(ops.space_to_im(_to_dev(points, dev), P, cam, wrapper=False), points)

    def im_to_space(self, points, name=None, heights=None):
        if heights is None:
            print("No heights were input")
            return None
        dev = _exec_device(points, heights)
        bank = self._bank()
        _, H = bank.on(dev)
        cam = _camera_arg(bank, name, self.default_correspondence, points.shape[0], dev)
        return _ret(ops.im_to_space(_to_dev(points, dev), _to_dev(heights, dev), H, cam, wrapper=False), points)

    def state_to_im(self, points, name=None, out_dtype=torch.float64):
        if self._custom_space():
            return self.space_to_im(self.state_to_space(points), name=name)
        dev = _exec_device(points)
        bank = self._bank()
        P, _ = bank.on(dev)
        cam = _camera_arg(bank, name, self.default_correspondence, points.shape[0], dev)
        return _ret(ops.state_to_im(_to_dev(points, dev), P, cam, wrapper=False, out_dtype=out_dtype), points)

    def im_to_state(self, points, name=None, heights=None):
        if self._custom_space() or heights is None:
            return self.space_to_state(self.im_to_space(points, heights=heights, name=name))
        dev = _exec_device(points, heights)
        bank = self._bank()
        _, H = bank.on(dev)
        cam = _camera_arg(bank, name, self.default_correspondence, points.shape[0], dev)
        return _ret(ops.im_to_state(_to_dev(points, dev), _to_dev(heights, dev), H, cam, wrapper=False), points)

    def im_to_state_refined(self, points, name=None, heights=None, return_heights=False):
        """Fused form of the trackers' idiom im_to_state -> state_to_im -> height_from_template -> im_to_state
        (MC3D_crop_tracker.py:364-370, :1222-1227; mot_evaluator.py:169-176)."""
        dev = _exec_device(points, heights)
        bank = self._bank()
        P, H = bank.on(dev)
        cam = _camera_arg(bank, name, self.default_correspondence, points.shape[0], dev)
        out = ops.im_to_state_refined(_to_dev(points, dev), _to_dev(heights, dev), H, P, cam, wrapper=False,
                                      return_heights=return_heights)
        if return_heights:
            return _ret(out[0], points), _ret(out[1], points)
        return _ret(out, points)

    # ---- helpers
    def guess_heights(self, classes):
        heights = torch.zeros(len(classes))
        for i in range(len(classes)):
            heights[i] = self.class_heights.get(classes[i], self.class_heights["other"])
        return heights

    def height_from_template(self, template_boxes, template_space_heights, boxes):
        dev = _exec_device(template_boxes, template_space_heights, boxes)
        out = ops.height_from_template(_to_dev(template_boxes, dev), _to_dev(template_space_heights, dev),
                                       _to_dev(boxes, dev))
        return _ret(out, boxes)

    def test_transformation(self, points, classes=None, name=None, im=None, heights=None, verbose=True):
        """im -> state -> im round trip; returns mean top + bottom reprojection error in pixels (homography.py:554-604)."""
        if heights is None:
            if classes is None:
                print("Must either specify heights or classes for boxes")
                return None
            heights = self.guess_heights(classes)
        state_pts = self.im_to_state(points, heights=heights, name=name)
        repro = self.state_to_im(state_pts, name=name)
        error = torch.abs(points.to(repro.dtype).to(repro.device) - repro)
        bottom = torch.sqrt(error[:, :4, 0] ** 2 + error[:, :4, 1] ** 2).mean()
        top = torch.sqrt(error[:, 4:8, 0] ** 2 + error[:, 4:8, 1] ** 2).mean()
        if verbose:
            print("Average distance between reprojected points and original points:")
            print("-----------------------------")
            print("Top: {} pixels".format(top))
            print("Bottom: {} pixels".format(bottom))
        return top + bottom


class Homography_Wrapper:
    """homography.py:793-901: two correspondences per camera; the second is used for objects whose road-plane y
    (of corner 0) exceeds 60 ft.  The selection happens inside the kernels - both matrices are never evaluated for
    every row and then overwritten, as the reference does."""

    def __init__(self, hg1=None, hg2=None):
        if hg1 is None or hg2 is None:
            raise ValueError("Homography_Wrapper needs two initialised Homography objects (loading the reference's "
                             "pickled calibration files is outside this package)")
        self.hg1, self.hg2 = hg1, hg2

    def _bank(self):
        return _bank_for(self, self.hg1, self.hg2)

    def guess_heights(self, classes):
        return self.hg1.guess_heights(classes)

    def state_to_space(self, points):
        return self.hg1.state_to_space(points)

    def space_to_state(self, points):
        return self.hg1.space_to_state(points)

    def height_from_template(self, template_boxes, template_space_heights, boxes):
        return self.hg1.height_from_template(template_boxes, template_space_heights, boxes)

    def _default(self):
        return self.hg1.default_correspondence

    def im_to_space(self, points, name=None, heights=None):
        if heights is None:
            print("No heights were input")
            return None
        dev = _exec_device(points, heights)
        bank = self._bank()
        _, H = bank.on(dev)
        cam = _camera_arg(bank, name, self._default(), points.shape[0], dev)
        return _ret(ops.im_to_space(_to_dev(points, dev), _to_dev(heights, dev), H, cam, wrapper=True), points)

    def space_to_im(self, points, name=None):
        dev = _exec_device(points)
        bank = self._bank()
        P, _ = bank.on(dev)
        cam = _camera_arg(bank, name, self._default(), points.shape[0], dev)
        return _ret(ops.space_to_im(_to_dev(points, dev), P, cam, wrapper=True), points)

    def im_to_state(self, points, name=None, heights=None):
        dev = _exec_device(points, heights)
        bank = self._bank()
        _, H = bank.on(dev)
        cam = _camera_arg(bank, name, self._default(), points.shape[0], dev)
        return _ret(ops.im_to_state(_to_dev(points, dev), _to_dev(heights, dev), H, cam, wrapper=True), points)

    def state_to_im(self, points, name=None, out_dtype=torch.float64):
        dev = _exec_device(points)
        bank = self._bank()
        P, _ = bank.on(dev)
        cam = _camera_arg(bank, name, self._default(), points.shape[0], dev)
        return _ret(ops.state_to_im(_to_dev(points, dev), P, cam, wrapper=True, out_dtype=out_dtype), points)

    def state_to_im_all(self, points, out_dtype=torch.float64):
        """Every state into every camera: [d, ncam, 8, 2] (BASELINE config 4 mode ii)."""
        dev = _exec_device(points)
        P, _ = self._bank().on(dev)
        return _ret(ops.state_to_im(_to_dev(points, dev), P, None, wrapper=True, all_cams=True, out_dtype=out_dtype), points)

    def im_to_state_refined(self, points, name=None, heights=None, return_heights=False):
        dev = _exec_device(points, heights)
        bank = self._bank()
        P, H = bank.on(dev)
        cam = _camera_arg(bank, name, self._default(), points.shape[0], dev)
        out = ops.im_to_state_refined(_to_dev(points, dev), _to_dev(heights, dev), H, P, cam, wrapper=True,
                                      return_heights=return_heights)
        if return_heights:
            return _ret(out[0], points), _ret(out[1], points)
        return _ret(out, points)
