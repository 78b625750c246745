"""Drop-in for the reference's `pytorch_retinanet_detector_directional` tree (3D directional copy)."""
