"""Synthetic KITTI-format dataset trees (no dataset is available offline): the directory layout
eval_rcnn.py reads, <root>/multi_data/<name>/KITTI/{ImageSets/<split>.txt, object/training/{velodyne/
%06d.bin (N,4) f32, calib/%06d.txt, label_2/%06d.txt, image_2/%06d.png}} (tools/generate_multi_data.py:7-17,
lib/datasets/kitti_dataset.py:16-40), filled with seeded lidar-like scenes from synthetic.py."""
import os

import numpy as np

from . import synthetic

CALIB_LINES = [
    "P0: 7.215377e+02 0.0 6.095593e+02 0.0 0.0 7.215377e+02 1.728540e+02 0.0 0.0 0.0 1.0 0.0",
    "P1: 7.215377e+02 0.0 6.095593e+02 -3.875744e+02 0.0 7.215377e+02 1.728540e+02 0.0 0.0 0.0 1.0 0.0",
    "P2: 7.215377e+02 0.0 6.095593e+02 4.485728e+01 0.0 7.215377e+02 1.728540e+02 2.163791e-01 0.0 0.0 1.0 2.745884e-03",
    "P3: 7.215377e+02 0.0 6.095593e+02 -3.395242e+02 0.0 7.215377e+02 1.728540e+02 2.199936e+00 0.0 0.0 1.0 2.729905e-03",
    "R0_rect: 9.999239e-01 9.837760e-03 -7.445048e-03 -9.869795e-03 9.999421e-01 -4.278459e-03 7.402527e-03 4.351614e-03 9.999631e-01",
    "Tr_velo_to_cam: 7.533745e-03 -9.999714e-01 -6.166020e-04 -4.069766e-03 1.480249e-02 7.280733e-04 -9.998902e-01 -7.631618e-02 9.998621e-01 7.523790e-03 1.480755e-02 -2.717806e-01",
    "Tr_imu_to_velo: 9.999976e-01 7.553071e-04 -2.035826e-03 -8.086759e-01 -7.854027e-04 9.998898e-01 -1.482298e-02 3.195559e-01 2.024406e-03 1.482454e-02 9.998881e-01 -7.997231e-01",
]
IMAGE_SIZE = (1242, 375)


def _rect_to_velo(pts_rect):
    """inverse of calibration.Calibration.lidar_to_rect for the calibration above (float64)."""
    vals = {l.split(':')[0]: np.array(l.split(':')[1].split(), np.float64) for l in CALIB_LINES}
    R0 = vals['R0_rect'].reshape(3, 3)
    V2C = vals['Tr_velo_to_cam'].reshape(3, 4)
    ref = pts_rect.astype(np.float64) @ np.linalg.inv(R0).T
    return (ref - V2C[:, 3]) @ np.linalg.inv(V2C[:, :3]).T


def write_scene(train_dir, sample_id, rng, npoints=20000, n_cars=4, n_invisible=0):
    """one scene: lidar-like cloud in rect coordinates -> velodyne .bin, calib, labels for the car clusters, PNG.
    n_invisible: additional returns BEHIND the camera (a real 360-degree sweep has ~120 k points of which ~20 k fall into
    the image): they cost the data path file reading and the lidar -> camera transform and are then filtered out."""
    from PIL import Image
    pts_rect = synthetic.lidar_cloud(rng, npoints, n_cars=n_cars)
    if n_invisible:
        back = synthetic.uniform_cloud(rng, n_invisible)
        back[:, 2] = -back[:, 2] - 1.0                 # z < 0: behind the image plane
        pts_rect = np.concatenate([pts_rect, back], axis=0)
        pts_rect = pts_rect[rng.permutation(len(pts_rect))]
        npoints = len(pts_rect)
    velo = np.concatenate([_rect_to_velo(pts_rect), rng.random_sample((npoints, 1))], axis=1).astype(np.float32)
    velo.tofile(os.path.join(train_dir, 'velodyne', '%06d.bin' % sample_id))
    with open(os.path.join(train_dir, 'calib', '%06d.txt' % sample_id), 'w') as f:
        f.write("\n".join(CALIB_LINES) + "\n")
    labels = []
    for _ in range(2):
        x, z = rng.uniform(-10, 10), rng.uniform(8, 40)
        labels.append("Car 0.00 0 %.2f 600.00 150.00 700.00 220.00 1.50 1.60 3.90 %.2f 1.65 %.2f %.2f"
                      % (rng.uniform(-3, 3), x, z, rng.uniform(-3, 3)))
    with open(os.path.join(train_dir, 'label_2', '%06d.txt' % sample_id), 'w') as f:
        f.write("\n".join(labels) + "\n")
    png = os.path.join(train_dir, 'image_2', '%06d.png' % sample_id)
    if not os.path.exists(png):
        Image.new("L", IMAGE_SIZE).save(png)       # only the size header is read (kitti_dataset.py:50-55)


def make_dataset(root, name="kitti", n_scenes=8, split="val", seed=666, npoints=20000, n_invisible=0, alias_to=None):
    """-> the dataset root eval_rcnn.py derives from its own location: <root>/multi_data/<name>"""
    data_root = os.path.join(root, "multi_data", name)
    train_dir = os.path.join(data_root, "KITTI", "object", "training")
    for sub in ("velodyne", "calib", "label_2", "image_2"):
        os.makedirs(os.path.join(train_dir, sub), exist_ok=True)
    os.makedirs(os.path.join(data_root, "KITTI", "ImageSets"), exist_ok=True)
    rng = np.random.RandomState(seed)
    for i in range(n_scenes):
        write_scene(train_dir, i, rng, npoints=npoints, n_invisible=n_invisible)
    total = n_scenes
    if alias_to and alias_to > n_scenes:
        # a large data set from a pool of distinct scenes: ids n_scenes .. alias_to-1 are hard links to id % n_scenes
        for i in range(n_scenes, alias_to):
            for sub, ext in (("velodyne", ".bin"), ("calib", ".txt"), ("label_2", ".txt"), ("image_2", ".png")):
                dst = os.path.join(train_dir, sub, "%06d%s" % (i, ext))
                if not os.path.exists(dst):
                    os.link(os.path.join(train_dir, sub, "%06d%s" % (i % n_scenes, ext)), dst)
        total = alias_to
    with open(os.path.join(data_root, "KITTI", "ImageSets", split + ".txt"), "w") as f:
        f.write("\n".join("%06d" % i for i in range(total)) + "\n")
    return data_root
