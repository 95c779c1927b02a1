"""Synthetic cone tracks and pose-landmark graphs for the BASELINE.json configurations.

One generator feeds every consumer (CUDA path, CPU oracle, bench), so both sides always see the
same bytes.  Rules follow SURVEY.md section 8(d): own PRNG (splitmix64 + Box-Muller), observation
fields rounded to float32 and widened (wire types of opendlv-standard-message-set-v0.9.5.odvd:294-303,
widened at slam.cpp:83-84,108), zenith = 0, |azimuth| >= 1e-3 deg (slam.cpp:515 divides by
fabs(angle)), thresholds of the example command line (sameConeThreshold 1.2, coneMappingThreshold 50),
lidar 1.5 m ahead of the pose origin (slam.cpp:514).

Nothing here touches the GPU or the oracle; it is plain numpy.
"""
from __future__ import annotations

import dataclasses
import numpy as np

DEG2RAD = 0.017453292522222          # slam.hpp:134
RAD2DEG = 57.295779513082325         # slam.hpp:135
PI_REF = float(np.float32(3.14159265))  # slam.hpp:136 (float literal widened)
LIDAR_TO_COG = 1.5                   # slam.cpp:514

SAME_CONE_THRESHOLD = 1.2
CONE_MAPPING_THRESHOLD = 50.0
POSE_ID_BASE = 1000                  # slam.hpp:118
INFO_ODOMETRY = 5.0                  # slam.cpp:456
INFO_CONE = 0.01                     # slam.cpp:546

_U64 = np.uint64


class SplitMix64:
    """Counter-based splitmix64; vectorised. uniform() in (0,1), normal() via Box-Muller."""

    def __init__(self, seed: int):
        self.state = _U64(seed & 0xFFFFFFFFFFFFFFFF)

    def next_u64(self, n: int) -> np.ndarray:
        with np.errstate(over="ignore"):
            k = np.arange(1, n + 1, dtype=np.uint64)
            z = self.state + k * _U64(0x9E3779B97F4A7C15)
            self.state = self.state + _U64(n) * _U64(0x9E3779B97F4A7C15)
            z = (z ^ (z >> _U64(30))) * _U64(0xBF58476D1CE4E5B9)
            z = (z ^ (z >> _U64(27))) * _U64(0x94D049BB133111EB)
            z = z ^ (z >> _U64(31))
        return z

    def uniform(self, n: int) -> np.ndarray:
        return ((self.next_u64(n) >> _U64(11)).astype(np.float64) + 0.5) * (1.0 / 9007199254740992.0)

    def normal(self, n: int) -> np.ndarray:
        m = (n + 1) // 2
        u1 = self.uniform(m)
        u2 = self.uniform(m)
        r = np.sqrt(-2.0 * np.log(u1))
        out = np.concatenate([r * np.cos(2 * np.pi * u2), r * np.sin(2 * np.pi * u2)])
        return out[:n]

    def integers(self, n: int, lo: int, hi: int) -> np.ndarray:
        return (lo + (self.next_u64(n) % _U64(hi - lo))).astype(np.int64)


def f32(x):
    """Round to float32 and widen, as the OD4 message fields do."""
    return np.asarray(x, dtype=np.float32).astype(np.float64)


# ------------------------------------------------------------------------------------------------
# reference polar -> Cartesian maths in numpy (generator-side only: used to build *inputs* such as
# edge measurements for directly-constructed graphs; parity of the conversion itself is tested
# against the oracle, not against this)
# ------------------------------------------------------------------------------------------------
def transform_cone_to_cog(az_deg, dist):
    az_deg = np.asarray(az_deg, dtype=np.float64)
    dist = np.asarray(dist, dtype=np.float64)
    sign = az_deg / np.abs(az_deg)
    ang = PI_REF - np.abs(az_deg * DEG2RAD)
    dnew = np.sqrt(LIDAR_TO_COG * LIDAR_TO_COG + dist * dist - 2 * LIDAR_TO_COG * dist * np.cos(ang))
    anew = np.arcsin((np.sin(ang) * dist) / dnew) * RAD2DEG
    return anew * sign, dnew


def spherical_to_cartesian(az_deg, zen_deg, dist):
    a, d = transform_cone_to_cog(az_deg, dist)
    zen = np.asarray(zen_deg, dtype=np.float64)
    x = d * np.cos(zen * DEG2RAD) * np.cos(a * DEG2RAD)
    y = d * np.cos(zen * DEG2RAD) * np.sin(a * DEG2RAD)
    z = d * np.sin(zen * DEG2RAD)
    return x, y, z


def normalize_theta(t):
    t = np.asarray(t, dtype=np.float64)
    return t - 2 * np.pi * np.floor((t + np.pi) / (2 * np.pi))


def se2_between(prev, cur):
    """prev^-1 * cur for arrays (...,3)."""
    prev = np.asarray(prev, dtype=np.float64)
    cur = np.asarray(cur, dtype=np.float64)
    dx = cur[..., 0] - prev[..., 0]
    dy = cur[..., 1] - prev[..., 1]
    c, s = np.cos(prev[..., 2]), np.sin(prev[..., 2])
    return np.stack([c * dx + s * dy, -s * dx + c * dy, normalize_theta(cur[..., 2] - prev[..., 2])], axis=-1)


# ------------------------------------------------------------------------------------------------
# track geometry
# ------------------------------------------------------------------------------------------------
@dataclasses.dataclass
class Track:
    cones_xy: np.ndarray      # (M,2) ground truth
    cones_type: np.ndarray    # (M,) 1 yellow right, 2 blue left
    length: float
    centre: "callable"        # s -> (x, y, heading)


def ellipse_track(n_pairs=150, a=90.0, b=45.0, half_width=1.5) -> Track:
    t = np.linspace(0.0, 2 * np.pi, 200001)
    x, y = a * np.cos(t), b * np.sin(t)
    seg = np.hypot(np.diff(x), np.diff(y))
    s = np.concatenate([[0.0], np.cumsum(seg)])
    length = float(s[-1])

    def centre(sq):
        sq = np.mod(np.asarray(sq, dtype=np.float64), length)
        tt = np.interp(sq, s, t)
        cx, cy = a * np.cos(tt), b * np.sin(tt)
        hx, hy = -a * np.sin(tt), b * np.cos(tt)
        return cx, cy, np.arctan2(hy, hx)

    sc = (np.arange(n_pairs) + 0.5) * (length / n_pairs)
    cx, cy, h = centre(sc)
    nx, ny = -np.sin(h), np.cos(h)     # left normal
    left = np.stack([cx + half_width * nx, cy + half_width * ny], axis=1)
    right = np.stack([cx - half_width * nx, cy - half_width * ny], axis=1)
    cones = np.empty((2 * n_pairs, 2))
    cones[0::2] = left
    cones[1::2] = right
    types = np.empty(2 * n_pairs, dtype=np.int32)
    types[0::2] = 2
    types[1::2] = 1
    return Track(cones, types, length, centre)


def corridor_track(n_pairs, spacing=3.0, half_width=1.5, wiggle=40.0, wavelength=2000.0) -> Track:
    """Gently curving open corridor for the large-graph configuration (C5)."""
    length = n_pairs * spacing

    def centre(sq):
        sq = np.asarray(sq, dtype=np.float64)
        x = sq
        y = wiggle * np.sin(2 * np.pi * sq / wavelength)
        h = np.arctan2(wiggle * 2 * np.pi / wavelength * np.cos(2 * np.pi * sq / wavelength), 1.0)
        return x, y, h

    sc = (np.arange(n_pairs) + 0.5) * spacing
    cx, cy, h = centre(sc)
    nx, ny = -np.sin(h), np.cos(h)
    cones = np.empty((2 * n_pairs, 2))
    cones[0::2] = np.stack([cx + half_width * nx, cy + half_width * ny], axis=1)
    cones[1::2] = np.stack([cx - half_width * nx, cy - half_width * ny], axis=1)
    types = np.empty(2 * n_pairs, dtype=np.int32)
    types[0::2] = 2
    types[1::2] = 1
    return Track(cones, types, length, centre)


# ------------------------------------------------------------------------------------------------
# frames: what Slam::performSLAM receives (4xN frame matrix + pose), slam.cpp:298-338
# ------------------------------------------------------------------------------------------------
@dataclasses.dataclass
class Drive:
    track: Track
    poses_true: np.ndarray        # (P,3)
    poses_noisy: np.ndarray       # (P,3)  what the UKF would have delivered
    # all observations, flat, ordered by pose then by measured range (the column order of a frame)
    obs_pose: np.ndarray          # (E,) pose index
    obs_cone: np.ndarray          # (E,) ground-truth cone id
    obs_az: np.ndarray            # (E,) degrees, float32-rounded
    obs_range: np.ndarray         # (E,) float32-rounded
    obs_type: np.ndarray          # (E,)
    _frames: list = None
    _ids: list = None

    def _split(self):
        P = len(self.poses_true)
        counts = np.bincount(self.obs_pose, minlength=P)
        cuts = np.cumsum(counts)[:-1]
        fr = np.zeros((4, len(self.obs_pose)), order="F")
        fr[0] = self.obs_az; fr[2] = self.obs_range; fr[3] = self.obs_type
        self._frames = [np.asfortranarray(a) for a in np.split(fr, cuts, axis=1)]
        self._ids = np.split(self.obs_cone.astype(np.int64), cuts)

    @property
    def frames(self):
        """P arrays, each (4,N) float64 F-order (az deg, zen deg, range, type): what performSLAM gets."""
        if self._frames is None:
            self._split()
        return self._frames

    @property
    def frame_cone_ids(self):
        if self._ids is None:
            self._split()
        return self._ids


def simulate_drive(track: Track, n_poses: int, s_start=0.0, s_step=None, seed=18, r_min=0.5,
                   r_max=12.0, az_max=100.0, sigma_xy=0.05, sigma_th=0.005, sigma_r=0.03,
                   sigma_az=0.15, closed=True, chunk=65536) -> Drive:
    rng = SplitMix64(seed)
    if s_step is None:
        s_step = track.length / n_poses
    s = s_start + s_step * np.arange(n_poses)
    cx, cy, h = track.centre(s)
    poses_true = np.stack([cx, cy, h], axis=1)
    noise = rng.normal(3 * n_poses).reshape(n_poses, 3)
    poses_noisy = poses_true + noise * np.array([sigma_xy, sigma_xy, sigma_th])
    poses_noisy[:, 2] = normalize_theta(poses_noisy[:, 2])

    M = track.cones_xy.shape[0]
    n_pairs = M // 2
    spacing = track.length / n_pairs
    # candidate cones per pose: pairs whose arclength is within r_max+4 m of the lidar
    half = int(np.ceil((r_max + 4.0) / spacing)) + 1
    offs = np.arange(-half, half + 1)
    lid_x = cx + LIDAR_TO_COG * np.cos(h)
    lid_y = cy + LIDAR_TO_COG * np.sin(h)
    pair0 = np.floor((s % track.length if closed else s) / spacing).astype(np.int64)
    out = {k: [] for k in ("pose", "cone", "az", "rng")}
    for c0 in range(0, n_poses, chunk):
        c1 = min(n_poses, c0 + chunk)
        pr = pair0[c0:c1, None] + offs[None, :]
        if closed:
            pr = np.mod(pr, n_pairs)
            valid = np.ones_like(pr, dtype=bool)
        else:
            valid = (pr >= 0) & (pr < n_pairs)
            pr = np.clip(pr, 0, n_pairs - 1)
        cid = np.concatenate([2 * pr, 2 * pr + 1], axis=1)
        valid = np.concatenate([valid, valid], axis=1)
        dx = track.cones_xy[cid, 0] - lid_x[c0:c1, None]
        dy = track.cones_xy[cid, 1] - lid_y[c0:c1, None]
        ch, sh = np.cos(h[c0:c1, None]), np.sin(h[c0:c1, None])
        vx = ch * dx + sh * dy
        vy = -sh * dx + ch * dy
        rng_true = np.hypot(vx, vy)
        az_true = np.arctan2(vy, vx) * (180.0 / np.pi)
        vis = valid & (rng_true >= r_min) & (rng_true <= r_max) & (np.abs(az_true) <= az_max)
        nz = rng.normal(2 * vis.size).reshape(2, *vis.shape)
        r_obs = f32(rng_true + sigma_r * nz[0])
        az_obs = f32(az_true + sigma_az * nz[1])
        tiny = np.abs(az_obs) < 1e-3
        az_obs = np.where(tiny, f32(np.where(az_obs < 0, -1e-3, 1e-3)), az_obs)
        rows, cols = np.nonzero(vis)                       # row-major: by pose, then candidate column
        order = np.lexsort((r_obs[rows, cols], rows))      # by pose, then measured range (stable)
        rows, cols = rows[order], cols[order]
        out["pose"].append(rows + c0)
        out["cone"].append(cid[rows, cols])
        out["az"].append(az_obs[rows, cols])
        out["rng"].append(r_obs[rows, cols])
    obs_pose = np.concatenate(out["pose"]).astype(np.int64)
    obs_cone = np.concatenate(out["cone"]).astype(np.int64)
    return Drive(track, poses_true, poses_noisy, obs_pose, obs_cone, np.concatenate(out["az"]),
                 np.concatenate(out["rng"]), track.cones_type[obs_cone].astype(np.float64))


def trackdrive(n_laps=1, poses_per_lap=1000, seed=18) -> Drive:
    """C1 (n_laps=1) / C2 (n_laps=10): closed ellipse track a=90 b=45, 150 cone pairs."""
    trk = ellipse_track()
    n = n_laps * poses_per_lap
    return simulate_drive(trk, n, s_step=trk.length / poses_per_lap, seed=seed, closed=True)


# ------------------------------------------------------------------------------------------------
# graphs in SoA form (what the bulk C-ABI loader takes)
# ------------------------------------------------------------------------------------------------
@dataclasses.dataclass
class GraphSoA:
    pose_ids: np.ndarray      # (P,) int32
    pose_est: np.ndarray      # (P,3)
    lm_ids: np.ndarray        # (L,) int32
    lm_est: np.ndarray        # (L,2)
    eo_from: np.ndarray       # (Eo,) vertex ids
    eo_to: np.ndarray
    eo_z: np.ndarray          # (Eo,3)
    eo_info: np.ndarray       # (Eo,9) row-major
    el_pose: np.ndarray       # (El,) vertex ids
    el_lm: np.ndarray
    el_z: np.ndarray          # (El,2)
    el_info: np.ndarray       # (El,4) row-major
    fixed_ids: np.ndarray     # ids with setFixed(true)

    @property
    def sizes(self):
        return dict(P=len(self.pose_ids), L=len(self.lm_ids), Eo=len(self.eo_from), El=len(self.el_pose))


def graph_from_drive(drive: Drive, pose_id_base=None) -> GraphSoA:
    """Pose-landmark graph with ground-truth association (C2/C5; SURVEY 8(d): 'build the graph
    directly with the row-8/9 edge semantics').  Landmark ids = order of first observation
    (the order addConesToMap would create them, slam.cpp:610); initial landmark estimate = first
    observation mapped through the noisy pose (coneToGlobal, slam.cpp:499-510); odometry
    measurement = prev^-1 * cur of the noisy poses (slam.cpp:452-455); gauge = first two poses
    and first two cones fixed (slam.cpp:464-474)."""
    P = len(drive.poses_noisy)
    El = len(drive.obs_pose)
    cone = drive.obs_cone
    pose_of = drive.obs_pose
    zx, zy, _ = spherical_to_cartesian(drive.obs_az, np.zeros(El), drive.obs_range)
    # landmark numbering by first appearance
    uniq, first = np.unique(cone, return_index=True)
    order = np.argsort(first, kind="stable")
    uniq, first = uniq[order], first[order]
    lm_of_cone = -np.ones(drive.track.cones_xy.shape[0], dtype=np.int64)
    lm_of_cone[uniq] = np.arange(len(uniq))
    L = len(uniq)
    if pose_id_base is None:
        pose_id_base = max(POSE_ID_BASE, L)
    pn = drive.poses_noisy
    fp = pose_of[first]
    c, s = np.cos(pn[fp, 2]), np.sin(pn[fp, 2])
    lm_est = np.stack([zx[first] * c - zy[first] * s + pn[fp, 0], zx[first] * s + zy[first] * c + pn[fp, 1]], axis=1)
    eo_z = se2_between(pn[:-1], pn[1:])
    Eo = P - 1
    pose_ids = (pose_id_base + np.arange(P)).astype(np.int32)
    return GraphSoA(
        pose_ids=pose_ids,
        pose_est=np.ascontiguousarray(pn, dtype=np.float64),
        lm_ids=np.arange(L, dtype=np.int32),
        lm_est=np.ascontiguousarray(lm_est),
        eo_from=pose_ids[:-1].copy(), eo_to=pose_ids[1:].copy(),
        eo_z=np.ascontiguousarray(eo_z),
        eo_info=np.tile(np.eye(3).reshape(1, 9) * INFO_ODOMETRY, (Eo, 1)),
        el_pose=pose_ids[pose_of].astype(np.int32),
        el_lm=lm_of_cone[cone].astype(np.int32),
        el_z=np.ascontiguousarray(np.stack([zx, zy], axis=1)),
        el_info=np.tile(np.eye(2).reshape(1, 4) * INFO_CONE, (El, 1)),
        fixed_ids=np.array([pose_ids[0], pose_ids[1], 0, 1], dtype=np.int32),
    )


def c2_graph(n_laps=10, poses_per_lap=1000, seed=18) -> GraphSoA:
    return graph_from_drive(trackdrive(n_laps, poses_per_lap, seed))


def c5_graph(n_poses=1_000_000, n_pairs=100_000, seed=5) -> GraphSoA:
    trk = corridor_track(n_pairs)
    drive = simulate_drive(trk, n_poses, s_start=0.0, s_step=trk.length / n_poses * 0.999, seed=seed, closed=False)
    return graph_from_drive(drive)


def perturb_replicas(g: GraphSoA, n_replicas: int, seed=18, sigma_xy=0.05, sigma_th=0.005,
                     sigma_z=0.03, first=0):
    """C3: Monte-Carlo replicas of one topology.  Replica r perturbs the initial poses and the
    landmark measurements with seed+r.  Returns (pose_est (R,P,3), lm_est (R,L,2), el_z (R,El,2),
    eo_z (R,Eo,3))."""
    P, L, El, Eo = len(g.pose_ids), len(g.lm_ids), len(g.el_pose), len(g.eo_from)
    pe = np.empty((n_replicas, P, 3)); le = np.empty((n_replicas, L, 2))
    ez = np.empty((n_replicas, El, 2)); oz = np.empty((n_replicas, Eo, 3))
    for r in range(n_replicas):
        rng = SplitMix64(seed + first + r)
        pe[r] = g.pose_est + rng.normal(3 * P).reshape(P, 3) * np.array([sigma_xy, sigma_xy, sigma_th])
        pe[r, :, 2] = normalize_theta(pe[r, :, 2])
        le[r] = g.lm_est + rng.normal(2 * L).reshape(L, 2) * sigma_xy
        ez[r] = g.el_z + rng.normal(2 * El).reshape(El, 2) * sigma_z
        oz[r] = se2_between(pe[r, :-1], pe[r, 1:]) if Eo == P - 1 else g.eo_z
    return pe, le, ez, oz


# ------------------------------------------------------------------------------------------------
# C4: large cone field, association only
# ------------------------------------------------------------------------------------------------
@dataclasses.dataclass
class Field:
    map_x: np.ndarray
    map_y: np.ndarray
    map_type: np.ndarray     # int32
    pose: np.ndarray         # (3,)
    frame: np.ndarray        # (4,N) F-order


def cone_field(n_map=1_000_000, n_obs=100_000, seed=4, density=0.1, sigma=0.2, frac_matched=0.9) -> Field:
    rng = SplitMix64(seed)
    side = float(np.sqrt(n_map / density))
    mx = (rng.uniform(n_map) - 0.5) * side
    my = (rng.uniform(n_map) - 0.5) * side
    mt = rng.integers(n_map, 1, 5).astype(np.int32)
    pose = np.array([0.0, 0.0, 0.3])
    n_m = int(round(n_obs * frac_matched))
    # The reference conversion folds anything behind the vehicle to the front (asin in
    # transformConeToCoG, slam.cpp:518), like a forward-looking lidar: observe only cones ahead.
    c, s = np.cos(pose[2]), np.sin(pose[2])
    ahead = np.nonzero(c * (mx - pose[0]) + s * (my - pose[1]) > LIDAR_TO_COG + 2.0)[0]
    pick = ahead[rng.integers(n_m, 0, len(ahead))]
    n_u = n_obs - n_m
    ux = 2.0 + LIDAR_TO_COG + rng.uniform(n_u) * (0.5 * side - 4.0)      # vehicle frame, ahead
    uy = (rng.uniform(n_u) - 0.5) * side
    gx = np.concatenate([mx[pick] + sigma * rng.normal(n_m), pose[0] + c * ux - s * uy])
    gy = np.concatenate([my[pick] + sigma * rng.normal(n_m), pose[1] + s * ux + c * uy])
    ty = np.concatenate([mt[pick], rng.integers(n_obs - n_m, 1, 5).astype(np.int32)]).astype(np.float64)
    sh = rng.integers(n_obs, 0, 1 << 62)
    perm = np.argsort(sh, kind="stable")
    gx, gy, ty = gx[perm], gy[perm], ty[perm]
    # global -> vehicle (CoG) frame -> lidar frame -> (az deg, range)
    dx, dy = gx - pose[0], gy - pose[1]
    c, s = np.cos(pose[2]), np.sin(pose[2])
    vx = c * dx + s * dy - LIDAR_TO_COG
    vy = -s * dx + c * dy
    az = f32(np.arctan2(vy, vx) * (180.0 / np.pi))
    az = np.where(np.abs(az) < 1e-3, f32(1e-3), az)
    fr = np.zeros((4, n_obs), dtype=np.float64, order="F")
    fr[0] = az
    fr[2] = f32(np.hypot(vx, vy))
    fr[3] = ty
    return Field(mx, my, mt, pose, fr)
