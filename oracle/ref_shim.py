"""Import shim for running the UNMODIFIED reference (/root/reference) on CPU.

TEST INFRASTRUCTURE ONLY.  Used by ``oracle/make_golden.py`` in the build
container (where /root/reference is mounted) to produce the committed golden
fixtures under ``tests/golden/``.  Nothing in ``posegen_b200/`` imports this and
nothing on the GPU box needs it (/root/reference does not exist there).

What it does:
  1. registers empty stub modules for third-party packages the reference imports
     at module scope but never touches on the render path
     (core/utils/skeleton_utils.py:5,12-14 -> plotly / matplotlib / pytorch3d;
     run_nerf.py -> imageio, h5py, deepdish, smplx, pytorch_msssim, configargparse);
  2. when no GPU is present, rewrites the reference's hard-coded ``.to('cuda')``
     device strings (core/encoders.py:17, core/networks/nerf.py:167,192,196-202,
     core/utils/ray_utils.py:161,166,185,186,218, core/raycasters.py:687) to 'cpu'.
No reference file is modified or copied.
"""
import argparse
import os
import sys
import types

import torch

REFERENCE_ROOT = os.environ.get("POSEGEN_REFERENCE", "/root/reference")


class _PermissiveModule(types.ModuleType):
    """Any attribute the reference imports by name resolves to a placeholder class."""

    def __getattr__(self, item):
        if item.startswith("__"):
            raise AttributeError(item)
        return type(item, (), {})


def _stub(name, **attrs):
    mod = _PermissiveModule(name)
    mod.__dict__.update(attrs)
    sys.modules[name] = mod
    parent, _, child = name.rpartition(".")
    if parent and parent in sys.modules:
        setattr(sys.modules[parent], child, mod)
    return mod


class _ConfigArgParser(argparse.ArgumentParser):
    """Tiny stand-in for configargparse.ArgumentParser: ``--config file`` holds
    ``key = value`` lines that are expanded into argv (True -> bare flag)."""

    def __init__(self, *a, **k):
        k.pop("config_file_parser_class", None)
        k.pop("default_config_files", None)
        super().__init__(*a, **k)
        self._config_dest = None

    def add_argument(self, *a, **k):
        if k.pop("is_config_file", False):
            self._config_dest = a[0]
        return super().add_argument(*a, **k)

    def parse_args(self, args=None, namespace=None):
        args = list(sys.argv[1:] if args is None else args)
        expanded = []
        if self._config_dest in args:
            path = args[args.index(self._config_dest) + 1]
            for line in open(path):
                line = line.split("#")[0].strip()
                if not line or "=" not in line:
                    continue
                key, val = [s.strip() for s in line.split("=", 1)]
                if val == "True":
                    expanded.append("--" + key)
                elif val == "False":
                    continue
                else:
                    expanded += ["--" + key] + val.split()
        return super().parse_args(expanded + args, namespace)


def install():
    """Install stubs + device rewrite and put the reference on sys.path."""
    if not os.path.isdir(REFERENCE_ROOT):
        raise FileNotFoundError(
            f"reference not mounted at {REFERENCE_ROOT}; golden fixtures can only be "
            "regenerated in the build container")
    for name in ["plotly", "plotly.graph_objects", "matplotlib", "matplotlib.pyplot",
                 "pytorch3d", "pytorch3d.transforms",
                 "pytorch3d.transforms.rotation_conversions",
                 "imageio", "h5py", "deepdish", "smplx", "smplx.lbs",
                 "pytorch_msssim"]:
        if name not in sys.modules:
            _stub(name)
    sys.modules["smplx"].SMPL = object
    sys.modules["pytorch_msssim"].SSIM = object
    if "configargparse" not in sys.modules:
        _stub("configargparse", ArgumentParser=_ConfigArgParser)

    if not torch.cuda.is_available() and not getattr(torch.Tensor.to, "_pgn_shim", False):
        def _fix(args):
            return ["cpu" if isinstance(x, str) and x.startswith("cuda") else x for x in args]
        _tensor_to = torch.Tensor.to
        _module_to = torch.nn.Module.to

        def tensor_to(self, *a, **k):
            return _tensor_to(self, *_fix(a), **k)

        def module_to(self, *a, **k):
            return _module_to(self, *_fix(a), **k)
        tensor_to._pgn_shim = True
        torch.Tensor.to = tensor_to
        torch.nn.Module.to = module_to

    if REFERENCE_ROOT not in sys.path:
        sys.path.insert(0, REFERENCE_ROOT)


def build_reference_raycaster(tmpdir):
    """create_raycaster (core/raycasters.py:17) for configs/surreal/surreal.txt.
    Returns (render_kwargs_test, args)."""
    install()
    import contextlib
    import io
    import numpy as np
    with contextlib.redirect_stdout(io.StringIO()):
        import run_nerf  # noqa: F401  (config_parser, render_path)
        from core.raycasters import create_raycaster
        from core.utils.skeleton_utils import SMPLSkeleton, smpl_rest_pose, get_per_joint_coords
        os.makedirs(os.path.join(tmpdir, "x"), exist_ok=True)
        args = run_nerf.config_parser().parse_args(
            ["--config", os.path.join(REFERENCE_ROOT, "configs/surreal/surreal.txt"),
             "--basedir", tmpdir, "--expname", "x", "--no_reload"])
        data_attrs = {"skel_type": SMPLSkeleton, "near": 60., "far": 100., "n_views": 1,
                      "joint_coords": get_per_joint_coords(smpl_rest_pose.astype(np.float32))}
        _, render_kwargs_test, _, _, _, _ = create_raycaster(args, data_attrs)
    return render_kwargs_test, args
