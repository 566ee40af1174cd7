"""Deterministic synthetic restatements of the reference's example scenes.

Every generator returns a `Scene`: lattice particles built by `abi.cube_particles` (examples3d/helper.rs:6-40,
randomize = false), the model list, cuboid colliders, gravity and `SolverParameters::default()` with dt = 1/60.
Sizes are parameters so the same scene serves the parity tests (thousands of particles) and the benchmark
(millions).  No RNG is used unless `jitter` is requested (numpy PCG64, seed stated by the caller).
"""
from __future__ import annotations

from dataclasses import dataclass, field

import numpy as np

from . import abi
from .abi import (BoundaryHandling, CorotatedLinearElasticity, CuboidCollider, HeightFieldCollider, DruckerPragerPlasticity, MonaghanSphEos,
                  NeoHookeanElasticity, ParticleComponents, ParticleModel, SolverParameters, cube_particles)


@dataclass
class Scene:
    name: str
    cell_width: float
    particles: ParticleComponents
    models: list
    colliders: list
    gravity: tuple = (0.0, -9.81, 0.0)
    params: SolverParameters = field(default_factory=SolverParameters)

    @property
    def num_particles(self):
        return len(self.particles)


def box_colliders(lo, hi, thickness=2.0, friction=0.5, handling=BoundaryHandling.Friction, walls=True):
    """Cuboid floor (top at lo.y) and optionally 4 walls around [lo, hi] — the cuboid stand-in for the reference's
    heightfield ground (examples3d/sand3.rs:23-38), see SURVEY §8d."""
    lo, hi = np.asarray(lo, np.float64), np.asarray(hi, np.float64)
    cx, cz = (lo[0] + hi[0]) / 2, (lo[2] + hi[2]) / 2
    hx, hz = (hi[0] - lo[0]) / 2 + thickness, (hi[2] - lo[2]) / 2 + thickness
    hy = (hi[1] - lo[1]) / 2 + thickness
    cy = (lo[1] + hi[1]) / 2
    out = [CuboidCollider((hx, thickness / 2, hz), (cx, lo[1] - thickness / 2, cz), friction=friction, grid_boundary_handling=handling)]
    if walls:
        out += [
            CuboidCollider((thickness / 2, hy, hz), (lo[0] - thickness / 2, cy, cz), friction=friction, grid_boundary_handling=handling),
            CuboidCollider((thickness / 2, hy, hz), (hi[0] + thickness / 2, cy, cz), friction=friction, grid_boundary_handling=handling),
            CuboidCollider((hx, hy, thickness / 2), (cx, cy, lo[2] - thickness / 2), friction=friction, grid_boundary_handling=handling),
            CuboidCollider((hx, hy, thickness / 2), (cx, cy, hi[2] + thickness / 2), friction=friction, grid_boundary_handling=handling),
        ]
    return out


def sand3_heightfield_ground(friction=0.5, handling=BoundaryHandling.Friction) -> HeightFieldCollider:
    """the ground of examples3d/sand3.rs:23-38 and cube_through_sand3.rs:26-38: a 41 x 41 heightfield, heights[(i, j)] =
    -sin(i pi / 40), scale (40, 10, 40), translated by (0, 10, 0): a trough along x whose bottom (y = 0) is at z = 0"""
    n = 40
    i = np.arange(n + 1, dtype=np.float32)
    col = -np.sin(i * np.float32(np.pi) / np.float32(n)).astype(np.float32)
    heights = np.repeat(col[:, None], n + 1, axis=1)
    return HeightFieldCollider(heights, (40.0, 10.0, 40.0), translation=(0.0, 10.0, 0.0), friction=friction, grid_boundary_handling=handling)


def sand3(ni=100, nj=50, nk=50, x_offset=0.0, with_colliders=True, ground="cuboid", y0=None) -> Scene:
    """examples3d/sand3.rs:40-75: an upper Drucker-Prager sand lattice stacked on a corotated elastic lattice,
    E = 1e7, nu = 0.2, h = 0.2, r = h/4, rho0 = 2700.  Reference size is (100, 50, 50) x 2 = 500k particles;
    the 4M benchmark configuration is (200, 100, 100) x 2."""
    E, NU = 1.0e7, 0.2
    h = 0.2
    r = np.float32(h) / np.float32(4.0)
    models = [ParticleModel.with_plasticity(CorotatedLinearElasticity(E, NU), DruckerPragerPlasticity(E, NU)),
              ParticleModel(CorotatedLinearElasticity(E, NU))]
    y0 = np.float32(h) * np.float32(3.0) + np.float32(2.0) if y0 is None else np.float32(y0)
    upper = cube_particles((x_offset, y0 + r * np.float32(2.0) * np.float32(nj), 0.0), ni, nj, nk, 0, r, 2700.0)
    lower = cube_particles((x_offset, y0, 0.0), ni, nj, nk, 1, r, 2700.0)
    pc = ParticleComponents.concatenate([upper, lower])
    cols = []
    if with_colliders and ground == "heightfield":  # the reference scene's own ground
        cols = [sand3_heightfield_ground()]
    elif with_colliders:
        ext = 2.0 * float(r)
        cols = box_colliders((x_offset - 4.0, 2.0, -4.0), (x_offset + ni * ext + 4.0, 2.0 + 2 * nj * ext + 8.0, nk * ext + 4.0))
    return Scene("sand3", h, pc, models, cols)


def _axis_angle_x(angle):
    """rapier `.rotation(vector![angle, 0, 0])`: UnitQuaternion::new(axis-angle) in nalgebra coordinate order (i, j, k, w)"""
    half = np.float32(angle) * np.float32(0.5)
    return (float(np.sin(half)), 0.0, 0.0, float(np.cos(half)))


def sand_penetration3(ni=100, nj=50, nk=50, origin=None) -> Scene:
    """examples3d/sand_penetration3.rs:11-152: a Drucker-Prager sand lattice (E = 1e7, nu = 0.2, h = 0.2, r = h / 4, rho0 = 2700)
    inside four 41 x 41 heightfields, heights[(i, j)] = sin(10 i pi / 40), scale (40, 1, 40): a floor at y = 10, a ceiling at y = 12
    turned by pi about x, and two walls at z = +-5 turned by -+pi / 2 about x — the reference's scene with ROTATED heightfields.
    `origin` moves the lattice (reference: (0, 15.2, 0)); the reference inserts only the sand batch (:136-137)."""
    E, NU = 1.0e7, 0.2
    h = 0.2
    r = np.float32(h) / np.float32(4.0)
    models = [ParticleModel.with_plasticity(CorotatedLinearElasticity(E, NU), DruckerPragerPlasticity(E, NU)),
              ParticleModel(CorotatedLinearElasticity(E, NU))]
    n = 40
    i = np.arange(n + 1, dtype=np.float32)
    col = np.sin(i * np.float32(np.pi) / np.float32(n) * np.float32(10.0)).astype(np.float32)
    heights = np.repeat(col[:, None], n + 1, axis=1)
    pi = float(np.float32(np.pi))
    cols = [HeightFieldCollider(heights, (40.0, 1.0, 40.0), translation=(0.0, 10.0, 0.0)),
            HeightFieldCollider(heights, (40.0, 1.0, 40.0), translation=(0.0, 12.0, 0.0), rotation=_axis_angle_x(pi)),
            HeightFieldCollider(heights, (40.0, 1.0, 40.0), translation=(0.0, 0.0, 5.0), rotation=_axis_angle_x(-pi / 2.0)),
            HeightFieldCollider(heights, (40.0, 1.0, 40.0), translation=(0.0, 0.0, -5.0), rotation=_axis_angle_x(pi / 2.0))]
    if origin is None:
        origin = (0.0, float(np.float32(2.0) * (np.float32(h) * np.float32(3.0) + np.float32(2.0) + r * np.float32(2.0) * np.float32(50.0))), 0.0)
    pc = cube_particles(origin, ni, nj, nk, 0, r, 2700.0)
    return Scene("sand_penetration3", h, pc, models, cols)


def fluids3(ni=38, nj=20, nk=20, with_colliders=True) -> Scene:
    """examples3d/fluids3.rs:17-38: Monaghan weakly-compressible fluid, radius 0.1, h = 0.8 (64 particles per cell
    at rest), rho0 = 1000.  The reference has no collider; the dam-break variant adds a cuboid floor and walls."""
    h = 0.8
    r = 0.1
    models = [ParticleModel(MonaghanSphEos(1.0e6, 7, 1.01e-3))]
    pc = cube_particles((1.6, 1.6, 1.6), ni, nj, nk, 0, r, 1000.0)
    cols = []
    if with_colliders:
        cols = box_colliders((0.8, 0.8, 0.8), (1.6 + ni * 0.2 * 2.0, 1.6 + nj * 0.2 * 3.0, 1.6 + nk * 0.2 + 0.8))
    return Scene("fluids3", h, pc, models, cols)


def cube_through_sand3(ni=100, nj=50, nk=50, nc=25) -> Scene:
    """examples3d/cube_through_sand3.rs:51-75: a Drucker-Prager sand lattice hit by a block of kinematic particles
    (kinematic_vel = (10, 0, 0)); the "cube" is particles, not an SDF body."""
    E, NU = 1.0e7, 0.2
    h = 0.2
    r = np.float32(h) / np.float32(4.0)
    models = [ParticleModel.with_plasticity(CorotatedLinearElasticity(E, NU), DruckerPragerPlasticity(E, NU)),
              ParticleModel(CorotatedLinearElasticity(E, NU))]
    y0 = np.float32(h) * np.float32(3.0) + np.float32(2.0)
    sand = cube_particles((0.0, y0, 0.0), ni, nj, nk, 0, r, 2700.0)
    ext = 2.0 * float(r)
    block = cube_particles((-nc * ext - 1.0, y0, (nk - nc) * ext / 2), nc, nc, nc, 1, r, 2700.0)
    block.status["kinematic_vel_enabled"] = 1
    block.status["kinematic_vel"] = np.array([10.0, 0.0, 0.0], np.float32)
    pc = ParticleComponents.concatenate([sand, block])
    cols = box_colliders((-nc * ext - 6.0, 2.0, -4.0), (ni * ext + 4.0, 2.0 + nj * ext + 8.0, nk * ext + 4.0))
    return Scene("cube_through_sand3", h, pc, models, cols)


def neo_hookean_block(ni=40, nj=40, nk=20, x_offset=0.0, with_floor=True) -> Scene:
    """The synthetic uniform-density Neo-Hookean block of BASELINE.json configs[4] (per-GPU slab 400x400x200 at
    full size): spacing h/2, h = 0.2, E = 1e6, nu = 0.3, with a small sinusoidal initial displacement (initial
    velocities are discarded by the reference's first substep)."""
    h = 0.2
    r = np.float32(h) / np.float32(4.0)
    models = [ParticleModel(NeoHookeanElasticity(1.0e6, 0.3))]
    pc = cube_particles((x_offset, 2.6, 0.0), ni, nj, nk, 0, r, 1000.0)
    p = pc.position["point"]
    p[:, 1] += (np.float32(0.01) * np.sin(p[:, 0] * np.float32(3.0))).astype(np.float32)
    cols = []
    if with_floor:
        ext = 2.0 * float(r)
        cols = box_colliders((x_offset - 4.0, 2.0, -4.0), (x_offset + ni * ext + 4.0, 2.0 + nj * ext + 8.0, nk * ext + 4.0), walls=False)
    # The reference's Neo-Hookean dt bound is evaluated with cell_width and hardening swapped: it
    # returns ~0.03 s here, far above the stable step (~0.5 h / c = 2.7e-3 s), so the scene caps the substep explicitly.
    return Scene("neo_hookean_block", h, pc, models, cols, params=SolverParameters(max_substep_dt=1.5e-3))


def model_block(kind: str, ni=8, nj=8, nk=6, seed: int = 21) -> Scene:
    """A small pre-strained block dropped on a cuboid floor for the plasticity / failure models the example scenes of
    scenes.py do not reach: `kind` in nacc | rankine | snow | failure | failure_split.  Model parameters follow the
    reference's 2D examples (examples2d/basic2.rs:43-46,91-94, examples2d/elasticity2.rs:68-71, l_panel2.rs:49-52).
    Every particle starts with a seeded random deformation gradient (stretch up to +-8 %, small shear) so the return
    mappings and the failure criterion fire from the first substep on."""
    from .abi import MaximumStressFailure, NaccPlasticity, RankinePlasticity, SnowPlasticity
    E, NU = 1.0e6, 0.3
    h = 0.2
    r = np.float32(h) / np.float32(4.0)
    if kind == "nacc":
        model = ParticleModel.with_plasticity(NeoHookeanElasticity(E, NU), NaccPlasticity(E, NU, 0.04, True, 3.0, np.radians(40.0)))
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
    pc = cube_particles((0.0, 2.6, 0.0), ni, nj, nk, 0, r, 1000.0)
    rng = np.random.Generator(np.random.PCG64(seed))
    n = len(pc)
    F = np.tile(np.eye(3, dtype=np.float32), (n, 1, 1)) + rng.uniform(-0.08, 0.08, (n, 3, 3)).astype(np.float32) * np.array(
        [[1.0, 0.3, 0.3], [0.3, 1.0, 0.3], [0.3, 0.3, 1.0]], np.float32)
    pc.volume["deformation_gradient"] = F.transpose(0, 2, 1).reshape(n, 9)  # column-major like nalgebra
    ext = 2.0 * float(r)
    cols = box_colliders((-4.0, 2.0, -4.0), (ni * ext + 4.0, 2.0 + nj * ext + 8.0, nk * ext + 4.0), walls=False)
    return Scene("model_block_" + kind, h, pc, [model], cols, params=SolverParameters(max_substep_dt=1.0e-3))


def jitter(scene: Scene, amplitude: float, seed: int = 42) -> Scene:
    """Uniform position jitter (numpy PCG64(seed)); NOT the reference's un-vendored oorandom stream."""
    rng = np.random.Generator(np.random.PCG64(seed))
    p = scene.particles.position["point"]
    p += rng.uniform(-amplitude, amplitude, p.shape).astype(np.float32)
    return scene


SCENES = {"sand3": sand3, "fluids3": fluids3, "cube_through_sand3": cube_through_sand3, "sand_penetration3": sand_penetration3, "neo_hookean_block": neo_hookean_block,
          "model_block": model_block}
