"""Drop-in for the reference's top-level `retinanet` package (2D copy): losses, utils, model (post-processing), anchors."""
