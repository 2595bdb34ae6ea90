"""Shared helpers of the parity tests: seeded synthetic inputs and error measures."""
import numpy as np

from ntrtsim_b200 import _capi as capi


def quat_mul(a, b):
    """Hamilton product, (x, y, z, w) layout, broadcasting over leading axes."""
    ax, ay, az, aw = a[..., 0], a[..., 1], a[..., 2], a[..., 3]
    bx, by, bz, bw = b[..., 0], b[..., 1], b[..., 2], b[..., 3]
    return np.stack([aw * bx + ax * bw + ay * bz - az * by,
                     aw * by + ay * bw + az * bx - ax * bz,
                     aw * bz + az * bw + ax * by - ay * bx,
                     aw * bw - ax * bx - ay * by - az * bz], axis=-1)


def quat_rotate(q, v):
    """Rotate vectors v by unit quaternions q."""
    qv = np.concatenate([v, np.zeros(v.shape[:-1] + (1,))], axis=-1)
    qc = q * np.array([-1.0, -1.0, -1.0, 1.0])
    return quat_mul(quat_mul(q, qv), qc)[..., :3]


def random_unit_quat(rng, n):
    q = rng.normal(size=(n, 4))
    return q / np.linalg.norm(q, axis=1, keepdims=True)


def random_poses(model, n_env, seed, translate=2.0, rotate=True, lift=None, yaw_only=False):
    """Per-env rigid transform of the whole structure: rotation about the structure's
    centroid, then a translation. Returns rod states [n_env, R, 13] with zero velocities."""
    rng = np.random.default_rng(seed)
    base = model.initial_rod_state()
    centroid = base[:, 0:3].mean(axis=0)
    if rotate:
        if yaw_only:
            ang = rng.uniform(0, 2 * np.pi, size=n_env)
            q = np.stack([np.zeros(n_env), np.sin(ang / 2), np.zeros(n_env), np.cos(ang / 2)], axis=1)
        else:
            q = random_unit_quat(rng, n_env)
    else:
        q = np.tile(np.array([0.0, 0.0, 0.0, 1.0]), (n_env, 1))
    t = rng.uniform(-translate, translate, size=(n_env, 3))
    if lift is not None:
        t[:, 1] = rng.uniform(lift[0], lift[1], size=n_env)
    out = np.zeros((n_env, model.n_bodies, capi.ROD_STATE_DOUBLES))
    rel = base[:, 0:3] - centroid
    qb = np.broadcast_to(q[:, None, :], (n_env, model.n_bodies, 4))
    out[:, :, 0:3] = quat_rotate(qb, np.broadcast_to(rel[None], (n_env,) + rel.shape)) + centroid + t[:, None, :]
    out[:, :, 3:7] = quat_mul(qb, np.broadcast_to(base[None, :, 3:7], (n_env, model.n_bodies, 4)))
    return out


def align_quat_sign(a, ref):
    """q and -q are the same rotation: flip a where it is closer to -ref."""
    s = np.sign(np.sum(a[..., 3:7] * ref[..., 3:7], axis=-1, keepdims=True))
    s[s == 0] = 1.0
    out = a.copy()
    out[..., 3:7] *= s
    return out


def rod_errors(dev, ora, pos_scale):
    """(max position error / pos_scale, max quaternion error, max linear velocity error, max angular velocity error)"""
    dev = align_quat_sign(dev, ora)
    return (np.abs(dev[..., 0:3] - ora[..., 0:3]).max() / pos_scale, np.abs(dev[..., 3:7] - ora[..., 3:7]).max(),
            np.abs(dev[..., 7:10] - ora[..., 7:10]).max(), np.abs(dev[..., 10:13] - ora[..., 10:13]).max())


def tension_rel_error(dev_obs, ora_obs, floor):
    """max over cables of |T_dev - T_ora| / max(|T_ora|, floor)"""
    td, to = dev_obs[..., 2], ora_obs[..., 2]
    return (np.abs(td - to) / np.maximum(np.abs(to), floor)).max()
