"""Canonical comparison forms shared by the parity tests: the reference assigns block ids and
in-block particle order by atomics races, so integer structures are compared on canonical forms only."""
import numpy as np

from sparkl_b200 import abi


def canonical_sort_state(engine):
    """(sorted virtual ids, {vid: sorted particle ids}, dispatch multiset) of the NEXT grid after stage_sort."""
    hdr = engine.grid_headers(1)
    ids = engine.read(abi.ATTR_SORTED_IDS)
    order = np.argsort(hdr["virtual_id"], kind="stable")
    vids = hdr["virtual_id"][order]
    per_block = {}
    for h in hdr:
        s = ids[h["first_particle"]: h["first_particle"] + h["num_particles"]]
        per_block[int(h["virtual_id"])] = np.sort(s)
    per = getattr(getattr(engine, "T", abi), "REF_G2P2G_THREADS", 128)  # G2P2G_THREADS of the reference: 128 (dim3) / 64 (dim2)
    ndisp = int(sum((int(h["num_particles"]) + per - 1) // per for h in hdr))
    disp = engine.dispatch(ndisp)
    multiset = sorted((int(hdr["virtual_id"][d["active_block_id"]]), int(d["first_particle"]) - int(hdr["first_particle"][d["active_block_id"]])) for d in disp)
    return vids, per_block, multiset, hdr, ids


def nodes_by_key(engine, which):
    """dict-free canonical node table: (sorted vids, nodes[nblocks, 64] reordered by vid)."""
    hdr = engine.grid_headers(which)
    nodes = engine.grid_nodes(which)
    order = np.argsort(hdr["virtual_id"], kind="stable")
    return hdr["virtual_id"][order], nodes[order]


def rel_err(a, b, floor=0.0, atol=0.0, few=0.0):
    """max |a-b| minus an absolute allowance, relative to the scale of b.  `few` > 0: the max over all but that fraction (at least one)
    of the entries — for quantities that depend discontinuously on the state (see Tol.few)"""
    a, b = np.asarray(a, np.float64), np.asarray(b, np.float64)
    if not a.size:
        return 0.0
    scale = max(float(np.max(np.abs(b))), floor, 1e-30)
    d = np.abs(a - b)
    if few > 0.0:
        d = np.sort(d.ravel())
        k = max(1, int(np.ceil(few * d.size)))
        worst = float(d[-k - 1]) if d.size > k else 0.0
    else:
        worst = float(np.max(d))
    return max(worst - atol, 0.0) / scale


def rel_err_but_few(a, b, floor=0.0, atol=0.0, few=1e-3):
    """rel_err over all but the `few` (fraction, at least one) largest deviations: for fields that are discontinuous functions of
    the state (psi_pos is a running max of an energy with a positive-part clamp: one particle whose singular value crosses 1 within
    round-off, or flips a Drucker-Prager branch, moves it by its whole size — in the reference's own run-to-run spread too)"""
    a, b = np.asarray(a, np.float64).ravel(), np.asarray(b, np.float64).ravel()
    if not a.size:
        return 0.0
    d = np.sort(np.abs(a - b))
    k = max(1, int(np.ceil(few * d.size)))
    scale = max(float(np.max(np.abs(b))), floor, 1e-30)
    return max(float(d[-k - 1] if d.size > k else 0.0) - atol, 0.0) / scale


class Tol:
    few = 0.0  # set by a caller that compares against the REFERENCE's kernels on a scene with Drucker-Prager plasticity: the return mapping
               # is discontinuous at zero volumetric strain (plasticity_drucker_prager.rs:47-49), so round-off flips single particles between
               # branches — between two runs of the reference itself too (test_reference_run_to_run_spread_and_product_within_it).  A
               # fraction `few` of the entries of the stress-dependent fields is then exempt from the max-norm check.
    """Tolerance classes of the parity tests (all relative to the field's scale over the scene):
      base       well-conditioned state: positions (in units of h), F, Jp, masses, hardening variables
      20 x base  velocity-like: particle velocities, node momenta / velocities (they carry the stress error below)
      50 x base  energy / stress-like: psi_pos, psi momenta, log_vol_gain.  tau = 2 mu (F - R) F^T and
                 psi = mu sum (sigma - 1)^2 subtract nearly equal numbers at small strain, so in f32 they carry an
                 absolute error of ~1e-6 * modulus whatever the implementation (the oracle's SVD runs in double);
                 `modulus` (max Young modulus of the scene) sets that absolute allowance.
    """

    def __init__(self, base, modulus=0.0):
        self.base, self.modulus = base, modulus

    @property
    def energy_atol(self):
        # 4e-6 * (lambda + 2 mu) ~ mu * (3.3e-3)^2: the reference's own f32 SVD returns singular values that are off by up to
        # ~2-3e-3 when two of them coalesce, and psi = mu sum max(sigma - 1, 0)^2 turns that into this much
        # absolute noise on a particle at rest; the reference is also non-deterministic run to run (atomics), so the bound is
        # an envelope (2e-6 was exceeded by 2 % in one of many B200 runs of the kinematic-block scene)
        return 4e-6 * self.modulus


def scene_tol(scene, base):
    mod = 0.0
    for m in scene.models:
        c = m.constitutive_model
        mod = max(mod, float(getattr(c, "lam", 0.0)) + 2 * float(getattr(c, "mu", 0.0)))
    return Tol(base, mod)


def assert_particles_close(pa: abi.ParticleComponents, pb: abi.ParticleComponents, h, tol, what=""):
    if not isinstance(tol, Tol):
        tol = Tol(tol)
    assert len(pa) == len(pb)
    b = tol.base
    checks = {  # name: (error, allowed)
        "pos": (float(np.max(np.abs(pa.position["point"].astype(np.float64) - pb.position["point"]))) / max(h, 1e-30) if len(pa) else 0.0, b),
        "vel": (rel_err(pa.velocity["vector"], pb.velocity["vector"], floor=1e-2, few=tol.few), 20 * b),
        "F": (rel_err(pa.volume["deformation_gradient"], pb.volume["deformation_gradient"], floor=1.0, few=tol.few), b),
        "Jp": (rel_err(pa.volume["plastic_deformation_gradient_det"], pb.volume["plastic_deformation_gradient_det"], floor=1.0, few=tol.few), b),
        "psi_pos": (rel_err_but_few(pa.phase["psi_pos"], pb.phase["psi_pos"], floor=1e-3, atol=tol.energy_atol), 50 * b),
        "phase": (rel_err(pa.phase["phase"], pb.phase["phase"], floor=1.0), b),
        "plastic_hardening": (rel_err(pa.data["plastic_hardening"], pb.data["plastic_hardening"], floor=1.0, few=tol.few), b),
        "elastic_hardening": (rel_err(pa.data["elastic_hardening"], pb.data["elastic_hardening"], floor=1.0), b),
        "log_vol_gain": (rel_err(pa.data["log_vol_gain"], pb.data["log_vol_gain"], floor=5e-2, few=tol.few), 50 * b),
        "nacc_alpha": (rel_err(pa.data["nacc_alpha"], pb.data["nacc_alpha"], floor=1e-2), 50 * b),
    }
    bad = {k: v for k, v in checks.items() if not (v[0] <= v[1])}
    assert not bad, f"{what}: fields out of tolerance (error, allowed): {bad} (all: {checks})"
    for f in ("failed", "is_static", "kinematic_vel_enabled", "model_index"):
        assert np.array_equal(pa.status[f], pb.status[f]), f"{what}: status.{f} differs"
    assert np.array_equal(pa.status["kinematic_vel"], pb.status["kinematic_vel"])
    assert np.array_equal(pa.volume["mass"], pb.volume["mass"]) and np.array_equal(pa.volume["volume0"], pb.volume["volume0"])
    assert np.array_equal(pa.volume["radius0"], pb.volume["radius0"])
    return {k: v[0] for k, v in checks.items()}


# The reference's Snow and Rankine models recompose F = U S' V^T from nalgebra's f32 SVD at every substep.  When two singular
# values coalesce (both clamped to the same bound, or a nearly isotropic stretch) that SVD has a backward error of up to
# ~2e-3 (measured on the B200: scripts/diag_models.py — the f32 one-sided Jacobi of the product and the f64 Jacobi of the
# oracle agree with each other to 1e-7 on the same particles and both differ from the reference by the same 2e-3).  The
# affected particles (a few %) then perturb their neighbours through the grid.  For these two models the comparison with
# the reference is therefore statistical: median / 90 % quantile tight, maximum bounded, positions still tight.
ROBUST_MODELS = ("snow", "rankine")


def assert_particles_close_robust(pa, pb, h, what=""):
    def per_particle(a, b, scale):
        a, b = np.asarray(a, np.float64), np.asarray(b, np.float64)
        d = np.abs(a - b)
        return (d.reshape(len(d), -1).max(axis=1) if d.ndim > 1 else d) / scale
    vscale = max(float(np.abs(pb.velocity["vector"]).max()), 1e-2)
    fields = {
        "pos": per_particle(pa.position["point"], pb.position["point"], h),
        "vel": per_particle(pa.velocity["vector"], pb.velocity["vector"], vscale),
        "F": per_particle(pa.volume["deformation_gradient"], pb.volume["deformation_gradient"], 1.0),
        "Jp": per_particle(pa.volume["plastic_deformation_gradient_det"], pb.volume["plastic_deformation_gradient_det"], 1.0),
        "plastic_hardening": per_particle(pa.data["plastic_hardening"], pb.data["plastic_hardening"], 1.0),
        "elastic_hardening": per_particle(pa.data["elastic_hardening"], pb.data["elastic_hardening"], 1.0),
    }
    limits = {"pos": (5e-6, 2e-5, 1e-4), "vel": (2e-4, 1e-3, 5e-3)}  # (median, q90, max); default below
    out = {}
    for k, d in fields.items():
        med, q90, mx = (float(np.median(d)), float(np.quantile(d, 0.9)), float(d.max())) if len(d) else (0.0, 0.0, 0.0)
        lm, lq, lx = limits.get(k, (2e-5, 2e-4, 2e-2))
        assert med <= lm and q90 <= lq and mx <= lx, f"{what}: {k} median {med:.2e} (<= {lm}), q90 {q90:.2e} (<= {lq}), max {mx:.2e} (<= {lx})"
        out[k] = mx
    for f in ("failed", "is_static", "kinematic_vel_enabled", "model_index"):
        assert np.array_equal(pa.status[f], pb.status[f]), f"{what}: status.{f} differs"
    assert np.array_equal(pa.phase["phase"], pb.phase["phase"]), f"{what}: phase differs"
    return out


def node_positions(vids, h, dim=3):
    """world position of every node of the blocks `vids` (grid_update.rs:34-38): [nblocks, 64, 3] float32 (dim2: [nblocks, 16, 2])"""
    if dim == 2:
        bx = (vids & np.uint64(0xFFFFFFFF)).astype(np.int64) - (1 << 31)
        by = (vids >> np.uint64(32)).astype(np.int64) - (1 << 31)
        cell = np.arange(16)
        out = np.empty((len(vids), 16, 2), np.float32)
        out[..., 0] = (bx[:, None] * 4 + (cell & 3)[None, :]).astype(np.float32) * np.float32(h)
        out[..., 1] = (by[:, None] * 4 + (cell >> 2)[None, :]).astype(np.float32) * np.float32(h)
        return out
    m = (1 << 21) - 1
    org = 1 << 20
    bx = (vids & m).astype(np.int64) - org
    by = ((vids >> 21) & m).astype(np.int64) - org
    bz = ((vids >> 42) & m).astype(np.int64) - org
    cell = np.arange(64)
    out = np.empty((len(vids), 64, 3), np.float32)
    out[..., 0] = (bx[:, None] * 4 + (cell & 3)[None, :]).astype(np.float32) * np.float32(h)
    out[..., 1] = (by[:, None] * 4 + ((cell >> 2) & 3)[None, :]).astype(np.float32) * np.float32(h)
    out[..., 2] = (bz[:, None] * 4 + (cell >> 4)[None, :]).astype(np.float32) * np.float32(h)
    return out


def assert_nodes_close(na, nb, tol, what="", scene=None, vids=None, oracle=None, velocity_field=False):
    """grid nodes re-indexed by (virtual id, cell).  Projection tags must be equal; the collider index may differ only
    where two colliders are equidistant from the node up to round-off (the reference resolves such ties with a strict
    `<` on f32 distances, grid_update.rs:121-124, which no re-implementation can reproduce bit for bit)."""
    if not isinstance(tol, Tol):
        tol = Tol(tol)
    assert na.shape == nb.shape
    b = tol.base
    mscale = float(np.max(np.abs(nb["mass"]))) if nb.size else 0.0
    # a node whose mass is within rounding of the f32 underflow threshold is flushed to zero by one implementation and
    # not by the other (the reference flushes at every `red.add`): its velocity is 0 on one side and finite on the other,
    # and it weighs < 1e-20 of a regular node in every particle update.  Compare momenta / velocities on massive nodes.
    # After grid_update the field holds a velocity = momentum / mass: on light nodes (a few tiny contributions) that
    # ratio amplifies round-off, while what reaches the particles is weighted by the node's mass.  So once the field is a
    # velocity (has_velocity) the comparison is mass-weighted, i.e. a comparison of momenta.
    massive = (nb["mass"] > 1e-20 * mscale) & (na["mass"] > 1e-20 * mscale)
    wgt = np.where(massive, nb["mass"].astype(np.float64) / max(mscale, 1e-30), 0.0) if velocity_field else massive.astype(np.float64)
    mv_a = na["momentum_velocity"].astype(np.float64) * wgt[..., None]
    mv_b = nb["momentum_velocity"].astype(np.float64) * wgt[..., None]
    pv_a = na["psi_momentum_velocity"].astype(np.float64) * wgt
    pv_b = nb["psi_momentum_velocity"].astype(np.float64) * wgt
    checks = {
        "mass": (rel_err(na["mass"], nb["mass"]), b),
        "momentum_velocity": (rel_err(mv_a, mv_b, floor=1e-2, few=tol.few), 50 * b),
        "psi_momentum_velocity": (rel_err(pv_a, pv_b, floor=1e-3, atol=tol.energy_atol * max(mscale, 1.0)), 50 * b),
        "psi_mass": (rel_err(na["psi_mass"], nb["psi_mass"]), b),
        "prev_mass": (rel_err(na["prev_mass"], nb["prev_mass"]), b),
        # the normal is a central difference of the SDF over +-0.1 h (grid_update.rs:74-94): f32 round-off of the projected
        # points (~1e-7 * |x|) is amplified by 1 / (0.2 h), i.e. a few 1e-5 for coordinates of order 10
        "collision_normal": (rel_err(na["collision_normal"], nb["collision_normal"], floor=1.0), max(b, 1e-4)),
    }
    bad = {k: v for k, v in checks.items() if not (v[0] <= v[1])}
    assert not bad, f"{what}: node fields out of tolerance (error, allowed): {bad} (all: {checks})"
    if not np.array_equal(na["projection_status_tag"], nb["projection_status_tag"]):
        diff = na["projection_status_tag"] != nb["projection_status_tag"]
        where = ""
        if vids is not None and scene is not None:
            pos = node_positions(vids, scene.cell_width, na["momentum_velocity"].shape[-1])[diff]
            where = f"; first nodes {pos[:4].tolist()} tags {na['projection_status_tag'][diff][:4].tolist()} vs {nb['projection_status_tag'][diff][:4].tolist()}"
        raise AssertionError(f"{what}: projection tags differ on {int(diff.sum())} of {diff.size} nodes{where}")
    same = na["projection_status_payload"] == nb["projection_status_payload"]
    dir_err = np.abs(na["projection_scaled_dir"].astype(np.float64) - nb["projection_scaled_dir"]).max(axis=-1) if na.size else np.zeros(0)
    assert (dir_err[same] <= 1e-5 * max(1.0, float(np.abs(nb["projection_scaled_dir"]).max()))).all(), f"{what}: projection_scaled_dir differs"
    if not same.all():
        assert scene is not None and oracle is not None and vids is not None, f"{what}: projection payload differs"
        dim = na["momentum_velocity"].shape[-1]
        pos = node_positions(vids, scene.cell_width, dim)[~same]
        ia, ib = na["projection_status_payload"][~same], nb["projection_status_payload"][~same]
        project = oracle.cuboid_project if dim == 3 else (lambda c, q: oracle.collider_project2d(c, q, 2 * scene.cell_width))
        for p, i, j in zip(pos, ia, ib):
            pi, _ = project(scene.colliders[int(i)], p)
            pj, _ = project(scene.colliders[int(j)], p)
            di, dj = np.linalg.norm(pi - p), np.linalg.norm(pj - p)
            assert abs(di - dj) <= 1e-5 * scene.cell_width, f"{what}: collider index differs at a non-tie node {p}: {i}@{di} vs {j}@{dj}"
    return {k: v[0] for k, v in checks.items()}


def boundary_particles(pos, h, eps=2e-5):
    """particles whose round(p / h) is decided by less than `eps` cells: their block membership is ill-defined at the
    f32 position tolerance"""
    q = np.asarray(pos, np.float64) / h
    return (np.abs(np.abs(q - np.floor(q)) - 0.5) < eps).any(axis=1)


def assert_membership_equal(canon_a, canon_b, boundary, what=""):
    """canonical sort states equal; particles on a rounding boundary are exempt (and so are the block sets when any exist)"""
    bad = set(np.nonzero(boundary)[0].tolist())
    if not bad:
        assert np.array_equal(canon_a[0], canon_b[0]), f"{what}: active block sets differ"
        assert canon_a[2] == canon_b[2], f"{what}: dispatch multisets differ"
    keys = set(canon_a[1]) | set(canon_b[1])
    for k in keys:
        a = set(canon_a[1].get(k, np.zeros(0, np.uint32)).tolist()) - bad
        b = set(canon_b[1].get(k, np.zeros(0, np.uint32)).tolist()) - bad
        assert a == b, f"{what}: particle set of block {k:#x} differs: {sorted(a ^ b)[:8]}"
    return len(bad)


def robust_particle_agreement(pa, pb, h, pos_tol_cells=1e-3):
    """fraction of particles whose positions differ by more than pos_tol_cells * h, and the deviation of the means"""
    d = np.abs(pa.position["point"].astype(np.float64) - pb.position["point"]).max(axis=1) / h
    return float((d > pos_tol_cells).mean()), float(np.abs(pa.position["point"].mean(axis=0) - pb.position["point"].mean(axis=0)).max() / h)


def assert_grids_close(eng_a, eng_b, which, tol, what, scene=None, oracle=None, velocity_field=False):
    """compare two engines' grids re-indexed by (virtual id, cell).  Blocks present on one side only are legal when a
    particle sits on a rounding boundary (its stencil shifts by one cell and the extra nodes receive ~zero weight):
    such blocks must carry negligible mass."""
    ka, na = nodes_by_key(eng_a, which)
    kb, nb = nodes_by_key(eng_b, which)
    common, ia, ib = np.intersect1d(ka, kb, return_indices=True)
    mscale = float(np.max(np.abs(nb["mass"]))) if nb.size else 0.0
    for k_only, n_only, side in ((np.setdiff1d(ka, kb), na, ka), (np.setdiff1d(kb, ka), nb, kb)):
        if len(k_only):
            sel = np.isin(side, k_only)
            assert float(np.abs(n_only[sel]["mass"]).max()) <= 1e-6 * max(mscale, 1e-30), f"{what}: a block active on one side only carries mass"
    return assert_nodes_close(na[ia], nb[ib], tol, what, scene, common, oracle, velocity_field)


# ---------------------------------------------------------------------------------------------- golden fixtures
GOLDEN_DIR = __import__("os").path.join(__import__("os").path.dirname(__import__("os").path.abspath(__file__)), "golden")


def golden_cases():
    import os
    return sorted(f[:-4] for f in os.listdir(GOLDEN_DIR) if f.endswith(".npz"))


def golden_particles(fx, s, template: abi.ParticleComponents) -> abi.ParticleComponents:
    """ParticleComponents holding the reference kernels' state after substep `s` (immutable fields from the scene)"""
    pc = type(template)(len(template))  # abi.ParticleComponents or abi2d.ParticleComponents
    for name in ("status", "position", "velocity", "volume", "phase", "data"):
        setattr(pc, name, getattr(template, name).copy())
    pc.position["point"] = fx[f"pos_{s}"]
    pc.velocity["vector"] = fx[f"vel_{s}"]
    pc.volume["deformation_gradient"] = fx[f"F_{s}"]
    pc.volume["plastic_deformation_gradient_det"] = fx[f"jp_{s}"]
    pc.phase["phase"], pc.phase["psi_pos"] = fx[f"phase_{s}"][:, 0], fx[f"phase_{s}"][:, 1]
    d = fx[f"data_{s}"]
    pc.data["nacc_alpha"], pc.data["plastic_hardening"], pc.data["elastic_hardening"], pc.data["log_vol_gain"] = d[:, 0], d[:, 1], d[:, 2], d[:, 3]
    pc.status["failed"] = fx[f"failed_{s}"]
    return pc


def run_against_golden(make_engine, case, base=2e-5, growth=3.0, dt_rtol=1e-5):
    """Drives `make_engine(scene)` in lock-step with the committed outputs of the reference's own GPU kernels
    (tests/golden/<case>.npz, generated by scripts/make_golden.py on a B200): same initial state, the reference's dt at
    every substep; active block set + per-block particle counts exact (up to particles on a rounding boundary), dt
    accumulator within 1e-5 (`dt_rtol`; looser for the dim2 fluid whose single-particle bound is a 0/0 at J = 1, see
    tests/test_dim2_gpu.py), particle state within Tol at the recorded substeps, final grid mass / velocity / tags."""
    import json
    import os
    from scripts.make_golden import build_case
    fx = np.load(os.path.join(GOLDEN_DIR, case + ".npz"))
    meta = json.loads(str(fx["meta"]))
    sc, nsub, rec = build_case(case)
    assert nsub == meta["nsub"] and list(rec) == meta["recorded"]
    if not np.array_equal(sc.particles.position["point"], fx["pos_init"]):  # a different numpy stream: use the recorded state
        sc.particles.position["point"][...] = fx["pos_init"]
    e = make_engine(sc)
    p = sc.params
    e.frame_begin(p)
    remaining = np.float32(p.dt)
    min_dt = float(np.float32(p.dt) / np.float32(p.max_num_substeps))
    worst = {}
    for s in range(nsub):
        tol = scene_tol(sc, base * (1 + growth * s))
        e.stage_sort()
        hdr = e.grid_headers(1)
        order = np.argsort(hdr["virtual_id"], kind="stable")
        boundary = boundary_particles(e.read(abi.ATTR_POSITION)["point"], sc.cell_width)
        if not boundary.any():
            assert np.array_equal(hdr["virtual_id"][order], fx[f"vids_{s}"]), f"{case} substep {s}: active block set differs from the reference"
            assert np.array_equal(hdr["num_particles"][order], fx[f"counts_{s}"]), f"{case} substep {s}: particles per block differ from the reference"
        e.stage_reset_grid()
        e.stage_estimate_dt(min_dt, float(min(remaining, np.float32(p.max_substep_dt))))
        acc, ref_acc = e.dt_acc(), int(fx["dt_accs"][s])
        assert abs(acc - ref_acc) <= max(2, int(dt_rtol * ref_acc)), f"{case} substep {s}: dt accumulator {acc} vs reference {ref_acc}"
        e.set_dt_acc(ref_acc)
        dt = float(fx["dts"][s])
        e.stage_g2p2g(dt)
        if s in rec and case in ROBUST_MODELS:
            errs = assert_particles_close_robust(e.read_all(), golden_particles(fx, s, sc.particles), sc.cell_width, f"{case} substep {s} vs reference fixture")
            for f, v in errs.items():
                worst[f] = max(worst.get(f, 0.0), v)
        elif s in rec:
            errs = assert_particles_close(e.read_all(), golden_particles(fx, s, sc.particles), sc.cell_width, tol, f"{case} substep {s} vs reference fixture")
            for f, v in errs.items():
                worst[f] = max(worst.get(f, 0.0), v)
        e.stage_grid_update(dt, sc.gravity)
        remaining = np.float32(remaining - np.float32(dt))
        if remaining <= 1e-11:
            e.frame_begin(p)
            remaining = np.float32(p.dt)
    if case in ROBUST_MODELS:
        return worst
    # final grid
    tol = scene_tol(sc, base * (1 + growth * nsub))
    vids, nodes = nodes_by_key(e, 0)
    common, ia, ib = np.intersect1d(vids, fx["grid_vids"], return_indices=True)
    assert len(common) >= 0.98 * len(fx["grid_vids"])
    mass, rmass = nodes["mass"][ia], fx["grid_mass"][ib]
    mscale = float(np.abs(rmass).max())
    assert rel_err(mass, rmass) <= tol.base, f"{case}: final grid mass {rel_err(mass, rmass)}"
    w = np.where((rmass > 1e-20 * mscale) & (mass > 1e-20 * mscale), rmass.astype(np.float64) / mscale, 0.0)[..., None]
    verr = rel_err(nodes["momentum_velocity"][ia] * w, fx["grid_vel"][ib] * w, floor=1e-2)
    assert verr <= 50 * tol.base, f"{case}: final grid velocity (mass weighted) {verr}"
    tag, rtag = nodes["projection_status_tag"][ia].astype(np.uint8), fx["grid_tag"][ib]
    if not np.array_equal(tag, rtag):
        # legal only where two colliders are equidistant from the node up to round-off (strict `<` on f32 distances,
        # grid_update.rs:121-124): e.g. a node 0.8 inside one wall and 0.8 outside the next one
        pos = node_positions(common, sc.cell_width, nodes["momentum_velocity"].shape[-1])[tag != rtag]
        for pt in pos:
            d = np.sort([_axis_aligned_cuboid_distance(c, pt) for c in sc.colliders])
            assert len(d) >= 2 and abs(d[0] - d[1]) <= 1e-5 * sc.cell_width, f"{case}: projection tag differs at a non-tie node {pt}"
    return worst


def _axis_aligned_cuboid_distance(col, pt):
    """distance from pt to the surface of an un-rotated cuboid collider (the scenes of scenes.py use identity rotations)"""
    he = np.asarray(col.half_extents, np.float64)
    loc = np.abs(np.asarray(pt, np.float64) - np.asarray(col.translation, np.float64))
    if (loc <= he).all():
        return float((he - loc).min())
    return float(np.linalg.norm(np.maximum(loc - he, 0.0)))
