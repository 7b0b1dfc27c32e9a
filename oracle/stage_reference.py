"""Stage the UNMODIFIED reference sources the caller-level GPU tests need into the git-ignored `baseline/_ref/`.

TEST INFRASTRUCTURE ONLY.  `/root/reference` exists in the build container but not on the GPU box; `baseline/_ref/`
is git-ignored (no reference source ever enters the history) yet travels with the gpurun snapshot, so
`tests/test_gpu_reference_callers.py` can drive the drop-in from the reference's OWN `core.trainer.Trainer.train_batch`
and `run_nerf.render_path` and compare against the reference's own `RayCaster` running eagerly on the B200
("second oracle").  Files are copied byte for byte: `core/` (the package), `run_nerf.py`,
`configs/surreal/`, `configs/h36m/`.

    python oracle/stage_reference.py            # no-op (exit 0) when /root/reference is absent
"""
import os
import shutil
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SRC = os.environ.get("POSEGEN_REFERENCE", "/root/reference")
DST = os.path.join(ROOT, "baseline", "_ref")


def stage(verbose: bool = False) -> bool:
    if not os.path.isdir(os.path.join(SRC, "core")):
        if verbose:
            print(f"stage_reference: {SRC} not present, nothing staged")
        return False
    os.makedirs(DST, exist_ok=True)
    for rel in ("core", os.path.join("configs", "surreal"), os.path.join("configs", "h36m")):
        dst = os.path.join(DST, rel)
        if os.path.isdir(dst):
            shutil.rmtree(dst)
        shutil.copytree(os.path.join(SRC, rel), dst, ignore=shutil.ignore_patterns("__pycache__", "*.pyc"))
    shutil.copy2(os.path.join(SRC, "run_nerf.py"), os.path.join(DST, "run_nerf.py"))
    if verbose:
        print(f"stage_reference: staged core/, run_nerf.py, configs/ -> {DST}")
    return True


if __name__ == "__main__":
    stage(verbose=True)
    sys.exit(0)
