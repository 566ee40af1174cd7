#!/usr/bin/env python
"""Generates tests/golden/*.npz: outputs of the REFERENCE'S OWN GPU KERNELS (oracle/_ref: the PTX the reference ships,
assembled for sm_100a and driven by oracle/ref_harness.cpp with the launch sequence of src/cuda/cuda_mpm_pipeline.rs)
on small seeded scenes.  Needs a GPU; run on the B200 box:

    gpurun -- 'python scripts/make_golden.py gpurun_out/golden'      # then copy gpurun_out/golden/*.npz to tests/golden/

The reference holds no golden vectors for this path, so these fixtures are what pins the CPU oracle in
the `-m "not gpu"` suite (tests/test_golden_cpu.py) and the CUDA product in the `-m gpu` suite (tests/test_golden_gpu.py).
Per scene: the initial positions, the dt the reference chose at every substep, and the full particle state + active block
set after selected substeps, + the grid (mass, velocity, projection tag) after the last grid update.
"""
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from sparkl_b200 import abi, scenes, scenes2d  # noqa: E402

# name -> (generator, args, jitter amplitude, seed, substeps, recorded substeps)
CASES = {
    "sand3": ("sand3", (10, 5, 8), 0.02, 11, 12, (0, 3, 7, 11)),
    "fluids3": ("fluids3", (12, 8, 8), 0.01, 12, 8, (0, 3, 7)),
    "neo_hookean_block": ("neo_hookean_block", (10, 10, 6), 0.002, 13, 8, (0, 3, 7)),
    "cube_through_sand3": ("cube_through_sand3", (10, 5, 6, 4), 0.02, 14, 8, (0, 3, 7)),
    # the plasticity / failure models no example scene reaches (pre-strained block, scenes.model_block)
    "nacc": ("model_block", ("nacc", 8, 8, 6), 0.01, 15, 8, (0, 3, 7)),
    "rankine": ("model_block", ("rankine", 8, 8, 6), 0.01, 16, 8, (0, 3, 7)),
    "snow": ("model_block", ("snow", 8, 8, 6), 0.01, 17, 8, (0, 3, 7)),
    "failure": ("model_block", ("failure", 8, 8, 6), 0.01, 18, 8, (0, 3, 7)),
    "failure_split": ("model_block", ("failure_split", 8, 8, 6), 0.01, 19, 8, (0, 3, 7)),
}


# dim2 cases (reference: resources/sparkl2d-kernels.ptx through the -DSPK_DIM=2 harness).  All collider-free: the byte image of
# the dim2 GpuCollider was not recovered (include/sparkl_abi.h), so the dim2 harness cannot upload colliders.
CASES_2D = {
    "elasticity2": ("elasticity2", (2, 0.3, False), 0.002, 31, 8, (0, 3, 7)),   # BASELINE.json configs[0] (pre-strained, see build_case)
    "fluids2": ("fluids2", (24, 16, False), 0.002, 32, 8, (0, 3, 7)),
    "corotated2": ("model_block2", ("corotated", 14, 12), 0.01, 33, 8, (0, 3, 7)),
    "neo_hookean2": ("model_block2", ("neo_hookean", 14, 12), 0.01, 34, 8, (0, 3, 7)),
    "sand2": ("model_block2", ("sand", 14, 12), 0.01, 35, 8, (0, 3, 7)),
    "nacc2": ("model_block2", ("nacc", 14, 12), 0.01, 36, 8, (0, 3, 7)),
    "rankine2": ("model_block2", ("rankine", 14, 12), 0.01, 37, 8, (0, 3, 7)),
    "snow2": ("model_block2", ("snow", 14, 12), 0.01, 38, 8, (0, 3, 7)),
    "failure2": ("model_block2", ("failure", 14, 12), 0.01, 39, 8, (0, 3, 7)),
    "failure_split2": ("model_block2", ("failure_split", 14, 12), 0.01, 40, 8, (0, 3, 7)),
}


def case_dim(name):
    return 2 if name in CASES_2D else 3


def build_case(name):
    if name in CASES_2D:
        gen, args, amp, seed, nsub, rec = CASES_2D[name]
        sc = scenes2d.jitter(scenes2d.SCENES[gen](*args), amp, seed=seed)
        sc.colliders = []
        if name == "elasticity2":  # free-falling stars would only exercise advection: pre-strain them so corotated + Rankine act
            rng = np.random.Generator(np.random.PCG64(seed + 100))
            n = len(sc.particles)
            F = np.tile(np.eye(2, dtype=np.float32), (n, 1, 1)) + rng.uniform(-0.05, 0.05, (n, 2, 2)).astype(np.float32) * np.array(
                [[1.0, 0.3], [0.3, 1.0]], np.float32)
            sc.particles.volume["deformation_gradient"] = F.transpose(0, 2, 1).reshape(n, 4)
        if name == "fluids2":  # a free-falling block at J = 1 feels no pressure: start from a seeded volume ratio J = 1 -+ 2 %
            rng = np.random.Generator(np.random.PCG64(seed + 100))
            sc.particles.volume["deformation_gradient"][:, 0] = (1.0 + rng.uniform(-0.02, 0.02, len(sc.particles))).astype(np.float32)
        return sc, nsub, rec
    gen, args, amp, seed, nsub, rec = CASES[name]
    sc = scenes.jitter(scenes.SCENES[gen](*args), amp, seed=seed)
    return sc, nsub, rec


def record_particles(out, s, pc):
    out[f"pos_{s}"] = pc.position["point"].copy()
    out[f"vel_{s}"] = pc.velocity["vector"].copy()
    out[f"F_{s}"] = pc.volume["deformation_gradient"].copy()
    out[f"jp_{s}"] = pc.volume["plastic_deformation_gradient_det"].copy()
    out[f"phase_{s}"] = np.stack([pc.phase["phase"], pc.phase["psi_pos"]], axis=1)
    out[f"data_{s}"] = np.stack([pc.data["nacc_alpha"], pc.data["plastic_hardening"], pc.data["elastic_hardening"], pc.data["log_vol_gain"]], axis=1)
    out[f"failed_{s}"] = pc.status["failed"].copy()


def main():
    from oracle import binding as ob
    args = [a for a in sys.argv[1:] if not a.startswith("--")]
    outdir = args[0] if args else os.path.join(ROOT, "tests", "golden")
    os.makedirs(outdir, exist_ok=True)
    names = list(CASES_2D) if "--dim2" in sys.argv else (list(CASES) + list(CASES_2D) if "--all" in sys.argv else list(CASES))
    for name in names:
        sc, nsub, rec = build_case(name)
        table = CASES_2D if name in CASES_2D else CASES
        if case_dim(name) == 2:
            assert ob.ref2d_available(), "oracle/_ref (dim2) not built or no CUDA device"
            e = ob.RefEngine2D(sc.cell_width)
        else:
            assert ob.ref_available(), "oracle/_ref not built or no CUDA device"
            e = ob.RefEngine(sc.cell_width)
        e.models_upload(sc.models); e.colliders_upload(sc.colliders); e.upload(sc.particles)
        p = sc.params
        e.frame_begin(p)
        remaining = np.float32(p.dt)
        min_dt = float(np.float32(p.dt) / np.float32(p.max_num_substeps))
        out = {"meta": np.array(json.dumps({"case": name, "generator": table[name][0], "args": table[name][1], "jitter": table[name][2], "seed": table[name][3],
                                             "nsub": nsub, "recorded": list(rec), "dim": case_dim(name),
                                             "what": f"reference sparkl{case_dim(name)}d-kernels.ptx on B200 (oracle/_ref)"})),
               "pos_init": sc.particles.position["point"].copy()}
        dts, accs = [], []
        for s in range(nsub):
            e.stage_sort()
            hdr = e.grid_headers(1)
            order = np.argsort(hdr["virtual_id"], kind="stable")
            out[f"vids_{s}"] = hdr["virtual_id"][order].copy()
            out[f"counts_{s}"] = hdr["num_particles"][order].copy()
            e.stage_reset_grid()
            dt = e.stage_estimate_dt(min_dt, float(min(remaining, np.float32(p.max_substep_dt))))
            dts.append(dt); accs.append(e.dt_acc())
            e.stage_g2p2g(dt)
            if s in rec:
                record_particles(out, s, e.read_all())
            e.stage_grid_update(dt, sc.gravity)
            remaining = np.float32(remaining - np.float32(dt))
            if remaining <= 1e-11:
                e.frame_begin(p)
                remaining = np.float32(p.dt)
        hdr = e.grid_headers(0)
        nodes = e.grid_nodes(0)
        order = np.argsort(hdr["virtual_id"], kind="stable")
        out["grid_vids"] = hdr["virtual_id"][order].copy()
        n = nodes[order]
        out["grid_mass"] = n["mass"].copy()
        out["grid_vel"] = n["momentum_velocity"].copy()
        out["grid_tag"] = n["projection_status_tag"].astype(np.uint8)
        out["dts"] = np.array(dts, np.float32)
        out["dt_accs"] = np.array(accs, np.uint64)
        path = os.path.join(outdir, name + ".npz")
        np.savez_compressed(path, **out)
        print(name, sc.num_particles, "particles ->", path, os.path.getsize(path), "bytes")


if __name__ == "__main__":
    main()
