"""End-to-end golden of the REFERENCE pipeline: its unmodified tools/eval_rcnn.py, dataset class, network, post-
processing and result writer run on the CPU of the build container (tools/refnet_cpu.py: CUDA extensions replaced by
the C restatements of their kernels; the only accommodation is the `far_points` keyword eval_rcnn.py:862 passes to a
constructor that calls it `npoints_faraway`) on a synthetic KITTI tree with seeded random-init weights.
    python tools/make_refeval_fixture.py     ->  tests/golden/refeval/00000{0,1,2}.txt     (~40 s)
tests/test_refeval_cpu.py re-runs it live and checks that the CPU port + host mirrors write byte-identical files;
tests/test_refeval_golden_gpu.py checks the sm_100a Detector against the committed files."""
import os
import sys
import tempfile

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "tests"))
sys.path.insert(0, os.path.join(ROOT, "tools"))
GOLD = os.path.join(ROOT, "tests", "golden", "refeval")
N_SCENES, DATA_SEED = 3, 666


def seeded_model(device):
    """default.yaml PointRCNN, seed 0; the RCNN score head is shifted so that boxes survive the 0.3 threshold
    (random-init heads score everything below it)."""
    from conftest import load
    model = load("inference").build_model(seed=0, device=device)
    with torch.no_grad():
        model.rcnn_net.cls_layer[-1].conv.bias.fill_(1.0)
    return model


def make_dataset(root):
    from conftest import load
    return load("synthetic_kitti").make_dataset(root, name="kitti", n_scenes=N_SCENES, split="val", seed=DATA_SEED)


def run_reference(workdir):
    """-> directory with the reference's final_result/data/*.txt"""
    from conftest import load
    import refnet_cpu as rn
    tu = load("train_utils")
    tools = rn.stage_reference_tree(workdir)
    make_dataset(os.path.dirname(tools))
    ckpt = os.path.join(workdir, "ckpt")
    os.makedirs(ckpt)
    tu.save_checkpoint(tu.checkpoint_state(seeded_model("cpu"), None, 1, 1), filename=os.path.join(ckpt, "checkpoint_epoch_1"))
    out = os.path.join(workdir, "out")
    r = rn.run_reference_eval(tools, ["--cfg_file", "cfgs/default.yaml", "--eval_mode", "rcnn", "--ckpt",
                                      os.path.join(ckpt, "checkpoint_epoch_1.pth"), "--batch_size", str(N_SCENES),
                                      "--workers", "0", "--output_dir", out])
    if r.returncode != 0:
        raise RuntimeError("reference eval_rcnn.py failed:\n" + r.stderr[-3000:])
    return os.path.join(out, "eval", "epoch_1", "val", "final_result", "data")


def main():
    with tempfile.TemporaryDirectory() as d:
        final = run_reference(d)
        os.makedirs(GOLD, exist_ok=True)
        for name in sorted(os.listdir(final)):
            text = open(os.path.join(final, name)).read()
            open(os.path.join(GOLD, name), "w").write(text)
            print(name, len(text.splitlines()), "boxes")


if __name__ == "__main__":
    main()
