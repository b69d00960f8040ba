"""Build recipe for the checker binaries (TEST INFRASTRUCTURE, never imported by the product).

  oracle/libpn2_oracle.so      <- oracle/pn2_oracle.c                      (gcc, always)
  oracle/_ref/libpn2_legacy.so <- the reference's own .cu files, compiled UNCHANGED from
                                  where they lie under /root/reference, plus
                                  oracle/legacy_shim.cu (extern "C" doors, no algorithm).
                                  Only built when /root/reference exists (this container);
                                  the GPU box uses the prebuilt file that travels with the
                                  snapshot (oracle/_ref/ is git-ignored, not gpurun-ignored).
"""
import os
import shutil
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
REF = os.environ.get("PN2_REFERENCE_ROOT", "/root/reference")
ARCH = ["-gencode", "arch=compute_100a,code=sm_100a"]

LEGACY_SOURCES = [
    "pointrcnn/pointnet2_lib/pointnet2/src/sampling_gpu.cu",
    "pointrcnn/pointnet2_lib/pointnet2/src/ball_query_gpu.cu",
    "pointrcnn/pointnet2_lib/pointnet2/src/group_points_gpu.cu",
    "pointrcnn/pointnet2_lib/pointnet2/src/interpolate_gpu.cu",
    "pointrcnn/lib/utils/iou3d/src/iou3d_kernel.cu",
    "pointrcnn/lib/utils/roipool3d/src/roipool3d_kernel.cu",
]


def _newer(dst, srcs):
    if not os.path.exists(dst):
        return False
    t = os.path.getmtime(dst)
    return all(os.path.getmtime(s) <= t for s in srcs)


def build_oracle(verbose=False):
    src = [os.path.join(HERE, "pn2_oracle.c"), os.path.join(HERE, "geom_oracle.c"), os.path.join(HERE, "datapath_oracle.c"),
           os.path.join(HERE, "fps_pruned_model.c")]
    src = [s for s in src if os.path.exists(s)]
    dst = os.path.join(HERE, "libpn2_oracle.so")
    if _newer(dst, src):
        return dst
    cmd = ["gcc", "-O2", "-ffp-contract=off", "-fno-fast-math", "-shared", "-fPIC", "-o", dst] + src + ["-lm"]
    if verbose:
        print(" ".join(cmd))
    subprocess.check_call(cmd)
    return dst


def build_legacy(verbose=False):
    """Returns the path of libpn2_legacy.so, or None when neither the reference tree nor a
    prebuilt library is available."""
    out_dir = os.path.join(HERE, "_ref")
    dst = os.path.join(out_dir, "libpn2_legacy.so")
    if not os.path.isdir(REF):
        return dst if os.path.exists(dst) else None
    os.makedirs(out_dir, exist_ok=True)
    shim = os.path.join(HERE, "legacy_shim.cu")
    srcs = [os.path.join(REF, s) for s in LEGACY_SOURCES]
    if _newer(dst, srcs + [shim]):
        return dst
    nvcc = shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"
    # the reference builds with nvcc -O2 and no arch flags (pointnet2/setup.py:19-20)
    cmd = [nvcc, "-O2", "-shared", "-Xcompiler", "-fPIC", "-lineinfo"] + ARCH + [
        "-I", os.path.join(HERE, "stub_include"),
        "-I", os.path.join(REF, "pointrcnn/pointnet2_lib/pointnet2/src"),
        "-o", dst, shim] + srcs
    if verbose:
        print(" ".join(cmd))
    subprocess.check_call(cmd)
    # the numba rotated-IoU reference is a Python file: it travels as a git-ignored copy
    shutil.copyfile(os.path.join(REF, "evaluate/rotate_iou.py"), os.path.join(out_dir, "rotate_iou.py"))
    return dst


def stage_reference_scripts():
    """git-ignored verbatim copies of reference Python files the GPU box needs (it has no /root/reference):
    evaluate/rotate_iou.py (numba goldens) and pointrcnn/tools/eval_rcnn.py, the unmodified driver that
    tests/test_eval_rcnn_dropin_gpu.py runs on top of the package."""
    out_dir = os.path.join(HERE, "_ref")
    if not os.path.isdir(REF):
        return
    os.makedirs(out_dir, exist_ok=True)
    for rel, name in (("evaluate/rotate_iou.py", "rotate_iou.py"), ("pointrcnn/tools/eval_rcnn.py", "eval_rcnn.py")):
        shutil.copyfile(os.path.join(REF, rel), os.path.join(out_dir, name))


if __name__ == "__main__":
    print(build_oracle(verbose=True))
    print(build_legacy(verbose=True))
