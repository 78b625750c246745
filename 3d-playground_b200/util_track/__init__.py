"""Drop-in for the reference's `util_track` package: kf.Torch_KF on the CUDA kernels."""
