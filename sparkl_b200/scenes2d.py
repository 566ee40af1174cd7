"""Deterministic synthetic restatements of the reference's 2D example scenes (examples2d/*.rs) for the `dim2` build.

Every generator returns a `scenes.Scene` whose particles / colliders use the dim2 layouts of abi2d.  BASELINE.json configs[0]
is `elasticity2`.  The reference places its blocks with the un-vendored `oorandom` crate; the placements here are fixed
lattice slots of the same generator range (ground_shift + 40 h k, k in 1..=5), stated in the code, so runs are reproducible.
"""
from __future__ import annotations

import numpy as np

from . import abi2d
from .abi2d import (BoundaryHandling, CorotatedLinearElasticity, CuboidCollider, DruckerPragerPlasticity, MaximumStressFailure, MonaghanSphEos,
                    NeoHookeanElasticity, ParticleComponents, ParticleModel, RankinePlasticity, SnowPlasticity, SolverParameters,
                    particles_from_positions, rect_particles)
from .scenes import Scene

GRAVITY = (0.0, -9.81)


def _points_in_polygon(pts: np.ndarray, poly: np.ndarray) -> np.ndarray:
    """even-odd rule; stands in for `shape.contains_local_point` of the convex decomposition (examples2d/helper.rs:66)"""
    x, y = pts[:, 0], pts[:, 1]
    inside = np.zeros(len(pts), bool)
    n = len(poly)
    for i in range(n):
        x0, y0 = poly[i]
        x1, y1 = poly[(i + 1) % n]
        crosses = (y0 > y) != (y1 > y)
        with np.errstate(divide="ignore", invalid="ignore"):
            xi = (x1 - x0) * (y - y0) / (y1 - y0) + x0
        inside ^= crosses & (x < xi)
    return inside


def sample_polygon(poly, translation, model_index, particle_rad, density) -> ParticleComponents:
    """examples2d/helper.rs:38-83 `sample_shape` with randomize = false: lattice of spacing 2 r over the shape's AABB, x-major"""
    poly = np.asarray(poly, np.float32)
    r2 = np.float32(particle_rad) * np.float32(2.0)
    lo, hi = poly.min(axis=0), poly.max(axis=0)
    xs, ys = [], []
    x = np.float32(lo[0])
    while x <= hi[0]:
        xs.append(x)
        x = np.float32(x + r2)
    y = np.float32(lo[1])
    while y <= hi[1]:
        ys.append(y)
        y = np.float32(y + r2)
    gx, gy = np.meshgrid(np.array(xs, np.float32), np.array(ys, np.float32), indexing="ij")
    pts = np.stack([gx.ravel(), gy.ravel()], axis=1)
    pts = pts[_points_in_polygon(pts, poly)]
    pts = (pts + np.asarray(translation, np.float32)).astype(np.float32)
    return particles_from_positions(pts, model_index, particle_rad, density)


def star(w):
    return np.array([[w, -w], [w * 0.5, 0.0], [w, w], [0.0, w * 0.5], [-w, w], [-w * 0.5, 0.0], [-w, -w], [0.0, -w * 0.5]], np.float32)


def ground_and_walls(cell_width, friction=0.5, handling=BoundaryHandling.Friction):
    """the three cuboids of examples2d/elasticity2.rs:26-48 / fluids2.rs:17-40"""
    gh, gs = cell_width * 10.0, cell_width * 40.0
    return [CuboidCollider((1000.0, gh), (0.0, gs - gh), friction=friction, grid_boundary_handling=handling),
            CuboidCollider((gh, 1000.0), (gs - gh, 0.0), friction=friction, grid_boundary_handling=handling),
            CuboidCollider((gh, 1000.0), (gs - gh + gs * 8.0, 0.0), friction=friction, grid_boundary_handling=handling)]


def elasticity2(nblocks=5, scale=1.0, with_colliders=True) -> Scene:
    """examples2d/elasticity2.rs:18-100 (BASELINE.json configs[0]): star-shaped corotated + Rankine blocks (E = 2e4, nu = 0.35,
    tensile strength 1e2, softening 5) of half-width 20 h sampled at r = h / 4, density 2, dropped on a cuboid ground between
    two cuboid walls; h = 0.05.  `scale` < 1 shrinks the stars for the parity tests."""
    E, NU = 2.0e4, 0.35
    h = 0.05
    gs = h * 40.0
    models = [ParticleModel.with_plasticity(CorotatedLinearElasticity(E, NU), RankinePlasticity(E, NU, 1.0e2, 5.0))]
    slots = [(1, 1), (3, 2), (5, 1), (2, 4), (4, 4)]  # stand-in for `rng.rand_u32() % 5 + 1` (elasticity2.rs:66-67)
    parts = []
    for k in range(nblocks):
        sx, sy = slots[k % len(slots)]
        parts.append(sample_polygon(star(h * 20.0 * scale), (gs + h * 40.0 * sx, gs + h * 40.0 * sy), 0, h / 4.0, 2.0))
    pc = ParticleComponents.concatenate(parts)
    return Scene("elasticity2", h, pc, models, ground_and_walls(h) if with_colliders else [], gravity=GRAVITY)


def fluids2(ni=300, nj=300, with_colliders=True) -> Scene:
    """examples2d/fluids2.rs:15-62: a 300 x 300 block of Monaghan weakly-compressible fluid (pressure0 1e4, gamma 7), h = 0.1,
    r = h / 4, frictionless ground and walls"""
    h = 0.1
    gs = h * 40.0
    models = [ParticleModel(MonaghanSphEos(1.0e4, 7, 1.01e-3))]
    pc = rect_particles((gs + h * 2.0 + h / 4.0, gs + h * 2.0 + h / 4.0), ni, nj, 0, h / 4.0, 1000.0)
    return Scene("fluids2", h, pc, models, ground_and_walls(h, friction=0.0) if with_colliders else [], gravity=GRAVITY)


def sand2(ni=40, nj=80, with_colliders=True) -> Scene:
    """the snow + sand columns of examples2d/basic2.rs:43-72 (corotated E = 1e5, nu = 0.2 with Snow / Drucker-Prager plasticity,
    h = 1/128, r = h / 4, density 1000) on a flat cuboid ground (`basic2` uses the reference's heightfield ground)"""
    h = 1.0 / 128.0
    models = [ParticleModel.with_plasticity(CorotatedLinearElasticity(1.0e5, 0.2), SnowPlasticity()),
              ParticleModel.with_plasticity(CorotatedLinearElasticity(1.0e5, 0.2), DruckerPragerPlasticity(1.0e5, 0.2))]
    r = h / 4.0
    x0, y0 = h * 40.0, h * 10.0 + 0.4
    snow = rect_particles((x0, y0), ni, nj, 0, r, 1000.0)
    sand = rect_particles((x0, y0 + nj * 2.0 * r + 4 * h), ni, nj, 1, r, 1000.0)
    pc = ParticleComponents.concatenate([snow, sand])
    cols = [CuboidCollider((10.0, 0.2), (0.0, y0 - 2 * h - 0.2))] if with_colliders else []
    return Scene("sand2", h, pc, models, cols, gravity=GRAVITY, params=SolverParameters(max_num_substeps=50))


def basic2_ground(friction=0.5, handling=BoundaryHandling.Friction):
    """the ground of examples2d/basic2.rs:32-41: a 41-sample heightfield, heights[i] = -sin(i pi / 40), scale (2, 1), translated by
    (0.5, 1.5): a trough whose bottom (y = 0.5) is at x = 0.5"""
    n = 40
    i = np.arange(n + 1, dtype=np.float32)
    heights = -np.sin(i * np.float32(np.pi) / np.float32(n)).astype(np.float32)
    return abi2d.HeightFieldCollider(heights, (2.0, 1.0), translation=(0.5, 1.5), friction=friction, grid_boundary_handling=handling)


def basic2(ni=12, nj=24, drop=0.02) -> Scene:
    """examples2d/basic2.rs:43-72 reduced: a snow column and a sand column (corotated E = 1e5, nu = 0.2; Snow / Drucker-Prager
    plasticity, h = 1/128, r = h / 4, density 1000) starting `drop` above the reference's heightfield ground"""
    h = 1.0 / 128.0
    models = [ParticleModel.with_plasticity(CorotatedLinearElasticity(1.0e5, 0.2), SnowPlasticity()),
              ParticleModel.with_plasticity(CorotatedLinearElasticity(1.0e5, 0.2), DruckerPragerPlasticity(1.0e5, 0.2))]
    r = h / 4.0
    x0 = h * 40.0
    ground = basic2_ground()
    y_ground = 1.5 - np.sin((x0 + 0.5) / 2.0 * np.pi)  # height of the trough around x0 (heights are -sin(pi (x + 0.5) / 2) before scaling)
    snow = rect_particles((x0, y_ground + 0.05 + drop), ni, nj, 0, r, 1000.0)
    sand = rect_particles((x0 + ni * 2.0 * r + 6 * h, y_ground + 0.05 + drop), ni, nj, 1, r, 1000.0)
    pc = ParticleComponents.concatenate([snow, sand])
    return Scene("basic2", h, pc, models, [ground], gravity=GRAVITY, params=SolverParameters(max_num_substeps=50))


def l_panel2(scale=1.0) -> Scene:
    """examples2d/l_panel2.rs:18-113: two L-shaped concrete panels (corotated E = 25.85e9, nu = 0.18; the second one with a
    MaximumStressFailure(2.7e6, MAX)), h = 0.005, r = h / 4, density 2500, a cuboid ground, no gravity, dt = 1 / 6000.  The scene's
    pulled-node boundary condition is an `MpmHooks::post_grid_update_hook` on the CPU solver's SpGrid (:124-138) and its Eigenerosion
    damage model is CPU-only: neither exists on the GPU path, the GPU run of this scene is the free panels.
    `scale` < 1 shrinks the panels for the parity tests."""
    E, NU = 25.85e9, 0.18
    h = 0.005
    gs = h * 40.0
    models = [ParticleModel(CorotatedLinearElasticity(E, NU)),
              ParticleModel(CorotatedLinearElasticity(E, NU), None, MaximumStressFailure(2.7e6, 3.4028235e38))]
    panel = np.array([[0.0, 0.0], [0.25, 0.0], [0.25, 0.25], [0.5, 0.25], [0.5, 0.5], [0.0, 0.5]], np.float32) * np.float32(scale)
    parts = [sample_polygon(panel, (gs, gs), 0, h / 4.0, 2500.0), sample_polygon(panel, (gs * 8.0, gs), 1, h / 4.0, 2500.0)]
    pc = ParticleComponents.concatenate(parts)
    cols = [CuboidCollider((1000.0, h * 10.0), (0.0, gs - h * 10.0))]
    params = SolverParameters(boundary_handling=BoundaryHandling.Stick)
    params.dt = 1.0 / 6000.0
    return Scene("l_panel2", h, pc, models, cols, gravity=(0.0, 0.0), params=params)


def model_block2(kind: str, ni=16, nj=16, seed: int = 21) -> Scene:
    """dim2 counterpart of scenes.model_block: a small pre-strained block over a cuboid floor for every constitutive / plastic /
    failure model: corotated | neo_hookean | fluid | sand | nacc | rankine | snow | failure | failure_split"""
    E, NU = 1.0e6, 0.3
    h = 0.2
    r = np.float32(h) / np.float32(4.0)
    if kind == "corotated":
        model = ParticleModel(CorotatedLinearElasticity(E, NU))
    elif kind == "neo_hookean":
        model = ParticleModel(NeoHookeanElasticity(E, NU))
    elif kind == "fluid":
        model = ParticleModel(MonaghanSphEos(1.0e5, 7, 1.01e-3))
    elif kind == "sand":
        model = ParticleModel.with_plasticity(CorotatedLinearElasticity(E, NU), DruckerPragerPlasticity(E, NU))
    elif kind == "nacc":
        model = ParticleModel.with_plasticity(NeoHookeanElasticity(E, NU), abi2d.NaccPlasticity(E, NU, 0.04, True, 3.0, np.radians(40.0)))
    elif kind == "rankine":
        model = ParticleModel.with_plasticity(CorotatedLinearElasticity(E, NU), RankinePlasticity(E, NU, 1.0e2, 5.0))
    elif kind == "snow":
        model = ParticleModel.with_plasticity(CorotatedLinearElasticity(E, NU), SnowPlasticity())
    elif kind in ("failure", "failure_split"):
        c = CorotatedLinearElasticity(E, NU)
        c.split_stress_on_failure = kind == "failure_split"
        model = ParticleModel(c, None, MaximumStressFailure(2.0e4, 3.0e38))
    else:
        raise ValueError(kind)
    pc = rect_particles((0.0, 2.6), ni, nj, 0, r, 1000.0)
    rng = np.random.Generator(np.random.PCG64(seed))
    n = len(pc)
    if kind != "fluid":
        F = np.tile(np.eye(2, dtype=np.float32), (n, 1, 1)) + rng.uniform(-0.08, 0.08, (n, 2, 2)).astype(np.float32) * np.array(
            [[1.0, 0.3], [0.3, 1.0]], np.float32)
        pc.volume["deformation_gradient"] = F.transpose(0, 2, 1).reshape(n, 4)  # column-major like nalgebra
    else:
        pc.volume["deformation_gradient"][:, 0] = (1.0 + rng.uniform(-0.02, 0.02, n)).astype(np.float32)
    ext = 2.0 * float(r)
    cols = [CuboidCollider((ni * ext + 8.0, 1.0), (ni * ext / 2, 1.0))]
    return Scene("model_block2_" + kind, h, pc, [model], cols, gravity=GRAVITY, params=SolverParameters(max_substep_dt=1.0e-3))


def jitter(scene: Scene, amplitude: float, seed: int = 42) -> Scene:
    rng = np.random.Generator(np.random.PCG64(seed))
    p = scene.particles.position["point"]
    p += rng.uniform(-amplitude, amplitude, p.shape).astype(np.float32)
    return scene


SCENES = {"elasticity2": elasticity2, "fluids2": fluids2, "sand2": sand2, "basic2": basic2, "l_panel2": l_panel2, "model_block2": model_block2}
