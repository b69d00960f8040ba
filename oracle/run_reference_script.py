"""Run one of the reference's UNMODIFIED scripts (tools/eval_rcnn.py) on a GPU, over the reference's own Python and its own
CUDA kernels (test / baseline infrastructure, the GPU twin of tools/refnet_cpu.py --run):

    python oracle/run_reference_script.py --work DIR eval_rcnn.py <script arguments>      (cwd: any)

DIR/pointrcnn is staged with lib/, pointnet2_lib/, tools/train_utils, tools/cfgs symlinked to the reference tree
(/root/reference in the build container, baseline/_ref/pointrcnn on the GPU box) and tools/{eval_rcnn.py,_init_path.py}
copied (the script derives its data root from its own realpath, eval_rcnn.py:854); DIR/pointrcnn/multi_data must exist
(symlink it to the data set).  Only the three compiled extension modules are supplied (oracle/refnet_gpu.py, backend
"legacy" = the reference's .cu files compiled unchanged), plus easydict / tensorboardX stand-ins, yaml.load's Loader
argument (PyYAML >= 6), and the one accommodation the script needs against its own data set class: eval_rcnn.py:862
passes far_points= to a constructor that calls the parameter npoints_faraway (a TypeError upstream)."""
import os
import runpy
import shutil
import sys
import types

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
sys.path.insert(0, ROOT)

_TENSORBOARDX = "class SummaryWriter(object):\n    def __init__(self, *a, **k): pass\n    def add_scalar(self, *a, **k): pass\n"


def stage(work):
    """-> DIR/pointrcnn/tools (idempotent)"""
    from oracle import refnet_gpu as rg
    ref = rg.ref_root()
    if ref is None:
        raise SystemExit("no reference tree (neither /root/reference nor baseline/_ref/pointrcnn)")
    root = os.path.join(work, "pointrcnn")
    tools = os.path.join(root, "tools")
    os.makedirs(tools, exist_ok=True)
    for name in ("lib", "pointnet2_lib"):
        if not os.path.lexists(os.path.join(root, name)):
            os.symlink(os.path.join(ref, name), os.path.join(root, name))
    for name in ("train_utils", "cfgs"):
        if not os.path.lexists(os.path.join(tools, name)):
            os.symlink(os.path.join(ref, "tools", name), os.path.join(tools, name))
    for name in ("eval_rcnn.py", "_init_path.py"):
        shutil.copyfile(os.path.join(ref, "tools", name), os.path.join(tools, name))
    os.makedirs(os.path.join(tools, "tensorboardX"), exist_ok=True)
    with open(os.path.join(tools, "tensorboardX", "__init__.py"), "w") as f:
        f.write(_TENSORBOARDX)
    return tools


def main():
    if len(sys.argv) < 4 or sys.argv[1] != "--work":
        raise SystemExit(__doc__)
    work, script, argv = sys.argv[2], sys.argv[3], sys.argv[4:]
    import yaml
    from oracle import refnet_gpu as rg
    tools = stage(work)
    os.chdir(tools)
    ed = types.ModuleType("easydict")
    ed.EasyDict = rg._AttrDict
    sys.modules.update(dict(rg._legacy_stubs(), easydict=ed))
    old_load = yaml.load
    yaml.load = lambda f, *a, **k: old_load(f, Loader=yaml.SafeLoader)
    sys.argv = [script] + argv
    sys.path.insert(0, tools)
    import _init_path  # noqa: F401  (the script's own path set-up)
    import lib.datasets.kitti_rcnn_dataset as ds_mod
    ref_init = ds_mod.KittiRCNNDataset.__init__

    def init_accepting_far_points(self, *a, far_points=None, **k):
        if far_points is not None:
            k["npoints_faraway"] = far_points
        ref_init(self, *a, **k)

    ds_mod.KittiRCNNDataset.__init__ = init_accepting_far_points
    runpy.run_path(os.path.join(tools, script), run_name="__main__")


if __name__ == "__main__":
    main()
