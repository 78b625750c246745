"""Drop-in for pytorch_retinanet_detector_directional/retinanet: losses, utils, model (post-processing), anchors."""
