"""The dim3 DEVICE updater, compiled for the host, against the CPU oracle — on GPU-less CI.

sparkl_b200/csrc/spk_updater.cuh (`DefaultParticleUpdater::update`: advection + velocity clamp, F update, ONE shared one-sided
Jacobi SVD, Drucker-Prager / NACC / Rankine / Snow return mappings, maximum-stress failure, corotated / Neo-Hookean / Monaghan
stress) is the code the fused G2P2G kernel inlines.  tests/host/updater3d_host.cpp compiles it with g++ and this file feeds it and
the oracle's `update_particle` (oracle/mpm_oracle.cpp, pinned against the reference's own kernels) the same seeded particles."""
import ctypes as C
import os
import subprocess

import numpy as np
import pytest

from oracle import binding as ob
from sparkl_b200 import abi, scenes
from tests.host.emu2d import DEVMODEL, _build, devmodel

f32 = np.float32
vp = C.c_void_p


@pytest.fixture(scope="module")
def hu():
    l = _build("updater3d_host")
    assert l.hu_sizeof_devmodel() == DEVMODEL.itemsize
    l.hu_update.argtypes = [vp, vp, vp, vp, vp, vp, vp, C.c_float, C.c_float]
    l.hu_dt_candidate.restype = C.c_float
    l.hu_dt_candidate.argtypes = [vp] + [C.c_float] * 8
    return l


def _p(a):
    return a.ctypes.data_as(vp)


def _oracle_update(o, st, pos, vel, vol, ph, data, iv, G, dt):
    stress = np.zeros(9, np.float32)
    G = G.copy()
    lib = ob.lib()
    lib.orc_update_particle.argtypes = [vp, C.c_float] + [vp] * 9
    ok = lib.orc_update_particle(o._h, dt, _p(st), _p(pos), _p(vel), _p(vol), _p(ph), _p(data), _p(iv), _p(G), _p(stress))
    return ok, G, stress


def _random_F(rng, i, amp):
    F = np.eye(3) + rng.uniform(-amp, amp, (3, 3)) * (np.full((3, 3), 0.4) + 0.6 * np.eye(3))
    if i % 5 == 0:
        q, _ = np.linalg.qr(rng.normal(size=(3, 3)))
        F = q @ F
    if i % 9 == 0:
        F = np.diag(rng.uniform(1 - amp, 1 + amp, 3))
    return F.astype(np.float32)


KINDS = ["sand", "corotated", "neo_hookean", "fluid", "nacc", "rankine", "snow", "failure", "failure_split"]


def _model(kind):
    if kind == "sand":
        return scenes.sand3(2, 2, 2).models[0]
    if kind == "corotated":
        return scenes.sand3(2, 2, 2).models[1]
    if kind == "neo_hookean":
        return scenes.neo_hookean_block(2, 2, 2).models[0]
    if kind == "fluid":
        return scenes.fluids3(2, 2, 2).models[0]
    return scenes.model_block(kind, 2, 2, 2).models[0]


@pytest.mark.parametrize("kind", KINDS)
def test_device_updater_matches_oracle(hu, kind):
    rng = np.random.default_rng(100 + KINDS.index(kind))
    model = _model(kind)
    dm = devmodel(model)
    h, dt = 0.2, 1.0e-3
    o = ob.OracleEngine(h)
    o.models_upload([model])
    lam_mu = float(getattr(model.constitutive_model, "lam", 0.0)) + 2 * float(getattr(model.constitutive_model, "mu", 0.0))
    n_p2g = n_changed = 0
    for i in range(400):
        F = _random_F(rng, i, 0.08)
        st = np.zeros(1, abi.STATUS)
        if i % 13 == 0:
            st["kinematic_vel_enabled"], st["kinematic_vel"] = 1, rng.uniform(-2, 2, 3).astype(np.float32)
        if i % 17 == 0:
            st["is_static"] = 1
        if i % 41 == 0:
            st["failed"] = 1
        pos = np.zeros(1, abi.POSITION); pos["point"] = rng.uniform(-3, 3, 3).astype(np.float32)
        vel = np.zeros(1, abi.VELOCITY)
        vol = np.zeros(1, abi.VOLUME)
        vol["mass"], vol["volume0"], vol["radius0"] = 0.125, 1.25e-4, 0.05
        vol["deformation_gradient"] = F.T.reshape(9)
        vol["plastic_deformation_gradient_det"] = f32(rng.uniform(0.9, 1.1))
        if kind == "fluid":
            vol["deformation_gradient"] = np.array([rng.uniform(0.9, 1.05), 0, 0, 0, 1, 0, 0, 0, 1], np.float32)
        ph = np.zeros(1, abi.PHASE); ph["phase"] = [1.0, 0.0, 0.6][i % 3] if "failure" in kind else 1.0; ph["psi_pos"] = f32(rng.uniform(0, 10))
        data = np.zeros(1, abi.DATA)
        data["nacc_alpha"], data["plastic_hardening"], data["elastic_hardening"] = rng.uniform(-0.05, 0.02), 1.0 + (rng.uniform(0, 0.5) if i % 2 else 0.0), rng.uniform(0.8, 1.3)
        data["log_vol_gain"] = rng.uniform(-0.02, 0.02) if i % 3 == 0 else 0.0
        iv = (rng.uniform(-3, 3, 3) * (100.0 if i % 29 == 0 else 1.0)).astype(np.float32)  # sometimes beyond the h / dt clamp
        G = rng.uniform(-2, 2, 9).astype(np.float32)
        # ---- product (device source on the host)
        flags = np.array([(1 if st["failed"][0] else 0) | (2 if st["is_static"][0] else 0) | (4 if st["kinematic_vel_enabled"][0] else 0)], np.uint32)
        part = np.concatenate([pos["point"][0], vol["deformation_gradient"][0], [vol["plastic_deformation_gradient_det"][0], vol["mass"][0], vol["volume0"][0],
                               ph["phase"][0], ph["psi_pos"][0], data["elastic_hardening"][0], data["plastic_hardening"][0], data["log_vol_gain"][0],
                               data["nacc_alpha"][0]]]).astype(np.float32)
        pvel, pG, pstress = iv.copy(), G.copy(), np.zeros(9, np.float32)
        ok_p = hu.hu_update(_p(dm), _p(flags), _p(part), _p(pvel), _p(pG), _p(pstress), _p(st["kinematic_vel"][0].copy()), h, dt)
        # ---- oracle
        ok_o, oG, ostress = _oracle_update(o, st, pos, vel, vol, ph, data, iv, G, dt)
        assert ok_p == ok_o, (kind, i)
        assert bool(flags[0] & 1) == bool(st["failed"][0])
        assert np.abs(part[:3] - pos["point"][0]).max() <= 1e-6 * max(1.0, np.abs(pos["point"]).max())
        assert np.abs(pvel - vel["vector"][0]).max() <= 1e-6 * max(1.0, np.abs(vel["vector"]).max())
        sv = np.linalg.svd(F.astype(np.float64), compute_uv=False)
        coalescing = kind in ("rankine", "snow") and min(abs(sv[0] - sv[1]), abs(sv[1] - sv[2])) < 3e-3 * sv[0]
        if not coalescing:  # clamping two nearly equal singular values: U, V ill-defined, f32 and f64 recompositions differ (util.ROBUST_MODELS)
            assert np.abs(part[3:12] - vol["deformation_gradient"][0]).max() <= 3e-5, (kind, i)
            assert abs(part[12] - vol["plastic_deformation_gradient_det"][0]) <= 3e-5
            assert abs(part[18] - data["plastic_hardening"][0]) <= 3e-5 * max(1.0, abs(data["plastic_hardening"][0]))
            assert abs(part[17] - data["elastic_hardening"][0]) <= 3e-4 * max(1.0, abs(data["elastic_hardening"][0]))
            assert abs(part[19] - data["log_vol_gain"][0]) <= 5e-5 and abs(part[20] - data["nacc_alpha"][0]) <= 1e-4
            assert part[15] == ph["phase"][0] or "failure" in kind
        if ok_p and not coalescing:
            n_p2g += 1
            assert np.abs(pG - oG).max() == 0.0
            assert np.abs(pstress - ostress).max() <= 1e-4 * np.abs(ostress).max() + 4e-6 * lam_mu + 1e-3, (kind, i, pstress, ostress)
            assert abs(part[16] - ph["psi_pos"][0]) <= 1e-3 * abs(ph["psi_pos"][0]) + 4e-6 * lam_mu
        if np.abs(part[3:12] - F.T.reshape(9)).max() > 1e-3:
            n_changed += 1
    assert n_p2g > 300 and n_changed > 50


def test_device_svd3_reconstructs(hu):
    rng = np.random.default_rng(7)
    for i in range(500):
        F = _random_F(rng, i, 0.3) if i % 4 else rng.normal(size=(3, 3)).astype(np.float32)
        if i % 50 == 7:
            F[:, 2] = F[:, 0] * f32(0.5) - F[:, 1]  # rank two
        u, s, v = np.zeros(9, np.float32), np.zeros(3, np.float32), np.zeros(9, np.float32)
        hu.hu_svd3(_p(np.ascontiguousarray(F.T)), _p(u), _p(s), _p(v))
        U, V = u.reshape(3, 3).T, v.reshape(3, 3).T
        scale = max(1.0, np.abs(F).max())
        assert np.abs(U @ np.diag(s) @ V.T - F).max() <= 5e-6 * scale
        assert np.abs(U @ U.T - np.eye(3)).max() <= 3e-6 and np.abs(V @ V.T - np.eye(3)).max() <= 3e-6 and (s >= 0).all()
        assert np.allclose(np.sort(s), np.sort(np.linalg.svd(F.astype(np.float64), compute_uv=False)), atol=5e-6 * scale)


def test_device_dt_candidate_matches_oracle(hu):
    rng = np.random.default_rng(9)
    for kind in ("corotated", "neo_hookean", "fluid"):
        model = _model(kind)
        dm = devmodel(model)
        h = 0.2
        for i in range(40):
            pc = abi.particles_from_positions(np.zeros((1, 3), np.float32), 0, 0.05, 1000.0)
            v = rng.uniform(-30, 30, 3).astype(np.float32) * (0.0 if i == 0 else 1.0)
            pc.velocity["vector"][0] = v
            pc.volume["deformation_gradient"][0, 0] = f32(rng.uniform(0.9, 1.1))
            pc.data["elastic_hardening"][0] = f32(rng.uniform(0.8, 1.3))
            o = ob.OracleEngine(h)
            o.models_upload([model]); o.upload(pc)
            o.set_dt_acc(2 ** 63)
            o.stage_estimate_dt(0.0, 1.0e6)
            cand = hu.hu_dt_candidate(_p(dm), h, float(pc.volume["mass"][0]), float(pc.volume["volume0"][0]), float(pc.volume["deformation_gradient"][0, 0]),
                                      float(pc.data["elastic_hardening"][0]), float(v[0]), float(v[1]), float(v[2]))
            want = o.dt_acc() / 1e12
            cand = 1.0e6 if np.isnan(cand) else min(cand, 1.0e6)
            assert abs(cand - want) <= 2e-5 * want, (kind, i, cand, want)
