"""ORACLE — test infrastructure, NOT product code.

Generates tests/golden/*.npz by EXECUTING THE UNMODIFIED REFERENCE (/root/reference, imported by
oracle/ref_loader.py) on the synthetic inputs of genpose_b200/synth.py.  Run in the build
container (the reference tree is not present on the GPU box):

    python -m oracle.make_golden            # rewrites every golden file
    python -m oracle.make_golden --check    # regenerates in memory and compares with the committed files

The reference ships no tests, fixtures or golden vectors and no checkpoint is
available offline, so these vectors — reference code, synthetic weights, injected noise — are what
pins the oracle port (oracle/genpose_oracle.py) and, through it, the CUDA path.

Each file stores the case parameters, checksums of every input (so a drifted generator is caught
before any comparison) and the reference's outputs.  Inputs themselves are re-derived from seeds.
"""
import argparse
import os
import sys

import numpy as np
import torch

_HERE = os.path.dirname(os.path.abspath(__file__))
_ROOT = os.path.dirname(_HERE)
if _ROOT not in sys.path:
    sys.path.insert(0, _ROOT)

from genpose_b200 import synth  # noqa: E402
from oracle import genpose_oracle as O  # noqa: E402
from oracle import ref_loader, ref_runner  # noqa: E402

GOLDEN_DIR = os.path.join(_ROOT, "tests", "golden")

# name -> case description.  kappa < 0 makes the reference's PC update (predictor sign as written,
# samplers.py:147-148) contractive; kappa > 0 does the same for the probability-flow ODE.
CASES = {
    # BASELINE.json configs[0]: 1 object, 1024 pts, K=1, T=10, ScoreNet only
    "config1_pc_B1_K1_T10": dict(sampler="pc", B=1, K=1, T=10, seed=0, kappa=-0.02, energy=False),
    "pc_B3_K4_T50": dict(sampler="pc", B=3, K=4, T=50, seed=1, kappa=-0.3, energy=True),
    "pc_B2_K5_T500": dict(sampler="pc", B=2, K=5, T=500, seed=2, kappa=-0.3, energy=True),
    "ode_B3_K4_T055": dict(sampler="ode", B=3, K=4, T0=0.55, seed=3, kappa=0.3, energy=True),
    "ode_B2_K3_T100": dict(sampler="ode", B=2, K=3, T0=1.0, seed=4, kappa=0.05, energy=False),
    # K = 50 (the candidates per object every BASELINE config and scripts/eval_single.sh:8 use): the cases the tensor-core samplers
    # accept (a 128-row tile must span <= 4 objects), so those kernels are held to reference-generated vectors directly
    "pc_B3_K50_T500": dict(sampler="pc", B=3, K=50, T=500, seed=5, kappa=-0.3, energy=True),
    "ode_B3_K50_T055": dict(sampler="ode", B=3, K=50, T0=0.55, seed=6, kappa=0.3, energy=True),
    # no stabilising linear field at all (kappa = 0): a short chain shows the error of the undamped reference dynamics
    "pc_B3_K50_T12_kappa0": dict(sampler="pc", B=3, K=50, T=12, seed=7, kappa=0.0, energy=False),
}


def case_inputs(case):
    """Everything a test needs to re-create the case's inputs from its seeds."""
    seed = case["seed"]
    sd = synth.make_state_dict(seed, kappa=case["kappa"])
    esd = synth.make_state_dict(seed + 100, kappa=case["kappa"]) if case.get("energy") else None
    clouds = synth.make_clouds(case["B"], seed)
    rows = case["B"] * case["K"]
    if case["sampler"] == "pc":
        x0 = synth.make_prior_noise(rows, seed, sigma=O.SIGMA_MAX)
        step_noise = synth.make_step_noise(case["T"], rows, seed)
    else:
        sig = float(O.sigma_of_t(torch.tensor(case["T0"])))
        x0 = synth.make_prior_noise(rows, seed, sigma=sig)
        step_noise = None
    return sd, esd, clouds, x0, step_noise


def input_checksums(sd, esd, clouds, x0, step_noise):
    cs = {
        "cs_clouds": synth.checksum(clouds),
        "cs_x0": synth.checksum(x0),
        "cs_weights": sum(synth.checksum(v.numpy()) for k, v in sd.items() if v.dtype == torch.float32),
    }
    if step_noise is not None:
        cs["cs_step_noise"] = synth.checksum(step_noise)
    if esd is not None:
        cs["cs_energy_weights"] = sum(synth.checksum(v.numpy()) for k, v in esd.items() if v.dtype == torch.float32)
    return cs


def reference_index_trace(sd, clouds):
    """FPS / ball-query indices and per-level features from the reference's own Python wrappers
    (pointnet2_utils.py) over the C restatement of its kernels."""
    with ref_loader.reference_env(ref_loader.default_argv("pc", 10)):
        from networks.pts_encoder.pointnet2_utils.pointnet2 import pointnet2_utils as pu
        xyz = torch.from_numpy(clouds).contiguous()
        out = {}
        cur = xyz
        for l in range(3):
            fps = pu.furthest_point_sample(cur, O.NPOINTS[l])
            new_xyz = pu.gather_operation(cur.transpose(1, 2).contiguous(), fps).transpose(1, 2).contiguous()
            out[f"fps_idx_l{l}"] = fps.numpy().copy()
            for s in range(2):
                bq = pu.ball_query(O.RADIUS[l][s], O.NSAMPLE[l][s], cur, new_xyz)
                out[f"ball_idx_l{l}_s{s}"] = bq.numpy().copy()
            cur = new_xyz
        return out


def generate(name):
    case = CASES[name]
    sd, esd, clouds, x0, step_noise = case_inputs(case)
    ref = ref_runner.run_reference(
        sd, clouds, case["K"], case["sampler"], num_steps=case.get("T"), T0=case.get("T0"),
        x0=x0, step_noise=step_noise, energy_sd=esd)
    rec = {f"ref_{k}": v for k, v in ref.items()}
    rec.update(reference_index_trace(sd, clouds))
    rec.update({k: np.float64(v) for k, v in input_checksums(sd, esd, clouds, x0, step_noise).items()})
    for k, v in case.items():
        rec[f"case_{k}"] = np.array(v)
    return rec


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--check", action="store_true")
    ap.add_argument("--only", nargs="*")
    args = ap.parse_args()
    if not ref_loader.available():
        raise SystemExit("reference tree not found; goldens can only be generated in the build container")
    os.makedirs(GOLDEN_DIR, exist_ok=True)
    torch.set_num_threads(os.cpu_count() or 1)
    bad = 0
    for name in (args.only or CASES):
        rec = generate(name)
        path = os.path.join(GOLDEN_DIR, name + ".npz")
        if args.check:
            old = np.load(path)
            for k in rec:
                a, b = np.asarray(rec[k]), old[k]
                same = np.array_equal(a, b) if a.dtype.kind in "iuUSb" else np.allclose(a, b, rtol=1e-5, atol=1e-6)
                if not same:
                    bad += 1
                    print(f"[MISMATCH] {name}:{k}")
            print(f"checked {name}")
        else:
            np.savez_compressed(path, **rec)
            print(f"wrote {path} ({os.path.getsize(path) / 1024:.0f} KiB)")
    if bad:
        raise SystemExit(f"{bad} golden entries differ")


if __name__ == "__main__":
    main()
