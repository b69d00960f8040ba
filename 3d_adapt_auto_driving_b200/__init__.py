"""B200-native PointRCNN inference hot path (drop-in for cxy1997/3D_adapt_auto_driving).

The directory name starts with a digit, so import it with
    importlib.import_module("3d_adapt_auto_driving_b200")
Sub-modules mirror the reference's own module names (pointnet2_cuda, pointnet2_utils,
pointnet2_modules, pytorch_utils, iou3d_utils, roipool3d_utils, rotate_iou, norm ...).
"""
__all__ = ["cabi"]
