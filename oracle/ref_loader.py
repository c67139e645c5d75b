"""ORACLE — test infrastructure, NOT product code.

Imports the UNMODIFIED reference Python from /root/reference on the CPU of THIS container
(recipe of SURVEY.md Appendix B).  /root/reference does not exist on the GPU box, so nothing that
runs there (the `-m gpu` tests, smoke(), bench.py) may call this; it exists to (a) generate the
golden vectors committed under tests/golden/ (oracle/make_golden.py) and (b) pin the portable
restatement oracle/genpose_oracle.py against the real thing (tests/test_oracle_vs_reference.py,
skipped when the reference tree is absent).

What is substituted, and why it does not change reference arithmetic:
  * ipdb / tensorboardX / matplotlib  -> import-only stubs (oracle/shims), never executed.
  * pytorch3d.transforms              -> v0.7.2 published definitions restated (oracle/shims).
  * pointnet2_cuda                    -> oracle/pointnet2_cpu.c, the fmaf-exact C restatement
                                         of the reference's own kernels (the reference has no
                                         CPU path for these ops).
  * torch.cuda.{Float,Int}Tensor      -> CPU aliases (pointnet2_utils.py:26-27,56,173,219
                                         allocate through them).
  * sys.argv                          -> the reference parses argv at import time
                                         (networks/pts_encoder/pointnet2.py:13).
"""
import contextlib
import importlib
import os
import sys

import torch

REFERENCE_ROOT = os.environ.get("GENPOSE_REFERENCE_ROOT", "/root/reference")
_HERE = os.path.dirname(os.path.abspath(__file__))
_SHIMS = os.path.join(_HERE, "shims")

_REF_TOPLEVEL = ("networks", "configs", "utils", "runners", "datasets")


def available() -> bool:
    return os.path.isdir(os.path.join(REFERENCE_ROOT, "networks", "gf_algorithms"))


def _purge_modules():
    for name in list(sys.modules):
        if name.split(".")[0] in _REF_TOPLEVEL or name in ("pointnet2_cuda",):
            del sys.modules[name]
    importlib.invalidate_caches()


@contextlib.contextmanager
def reference_env(argv):
    """Context in which `import networks...` resolves to /root/reference with the shims active."""
    if not available():
        raise RuntimeError(f"reference tree not found at {REFERENCE_ROOT}")
    from oracle import pointnet2_cpu

    saved_argv, saved_path = sys.argv, list(sys.path)
    saved_ft = getattr(torch.cuda, "FloatTensor", None)
    saved_it = getattr(torch.cuda, "IntTensor", None)
    saved_mods = {k: v for k, v in sys.modules.items()
                  if k.split(".")[0] in _REF_TOPLEVEL or k == "pointnet2_cuda"}
    _purge_modules()
    try:
        sys.argv = ["oracle"] + list(argv)
        sys.path[:0] = [_SHIMS, REFERENCE_ROOT]
        sys.modules["pointnet2_cuda"] = pointnet2_cpu
        torch.cuda.FloatTensor = torch.FloatTensor
        torch.cuda.IntTensor = torch.IntTensor
        yield
    finally:
        sys.argv = saved_argv
        sys.path[:] = saved_path
        if saved_ft is not None:
            torch.cuda.FloatTensor = saved_ft
        if saved_it is not None:
            torch.cuda.IntTensor = saved_it
        _purge_modules()
        sys.modules.update(saved_mods)


def default_argv(sampler_mode="pc", sampling_steps=None, posenet_mode="score", extra=()):
    argv = ["--device", "cpu", "--sampler_mode", sampler_mode, "--posenet_mode", posenet_mode]
    if sampling_steps is not None:
        argv += ["--sampling_steps", str(sampling_steps)]
    return argv + list(extra)


def build_agent(state_dict, argv):
    """Construct the reference `PoseNet` agent (networks/posenet_agent.py:46) on CPU inside an
    active reference_env() and load `state_dict` strictly (the load_ckpt contract, :143-173)."""
    from configs.config import get_config
    from networks.posenet_agent import PoseNet

    cfg = get_config()
    agent = PoseNet(cfg)
    agent.net.load_state_dict({k: v.clone() for k, v in state_dict.items()}, strict=True)
    agent.net.eval()
    return agent, cfg
