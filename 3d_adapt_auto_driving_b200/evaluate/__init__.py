"""KITTI AP evaluation around rotate_iou (mirror of the reference's evaluate/ directory):
kitti_common (label readers, skimage-free) and eval2 (get_official_eval_result and the functions under it)."""
