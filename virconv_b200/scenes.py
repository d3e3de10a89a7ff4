"""Seeded synthetic KITTI-shaped scenes (LiDAR + depth-completion virtual points).

This is bench/test *input preparation*, the stand-in for the reference's CPU
dataloader (`pcdet/datasets/kitti/kitti_dataset_mm.py`, `dataset.py:191-313`,
`data_processor.py:128-187`): no KITTI data exists offline, so scenes are
generated from a fixed camera model.  The point format is the one PENet writes
(`tools/PENet/vis_utils.py:126-152`): `[x, y, z, intensity, r, g, b, indicator]`
with indicator 2 = LiDAR return, 1 = virtual point.

Nothing here is on the timed hot path; the backbone consumes `voxel_features`
/ `voxel_coords` exactly as `load_data_to_gpu` (`pcdet/models/__init__.py:16-24`)
hands them over (float32 arrays, coordinates included).
"""
from __future__ import annotations

import math
from dataclasses import dataclass, field

import numpy as np

# KITTI P_rect_02 (tools/PENet/dataloaders/calib_cam_to_cam.txt:26) — the P2 the reference projects with.
P2 = np.array([[721.5377, 0.0, 609.5593, 44.85728],
               [0.0, 721.5377, 172.854, 0.2163791],
               [0.0, 0.0, 1.0, 0.002745884]], dtype=np.float32)
# default R0 hard-coded in pcdet/utils/calibration_kitti.py:30-32
R0 = np.array([[0.99992624, 0.00965411, -0.0072371],
               [-0.00968531, 0.99994343, -0.00433077],
               [0.00719491, 0.00440054, 0.99996366]], dtype=np.float32)
# public KITTI 2011_09_26 velodyne->camera extrinsics (any rigid transform works for synthetic scenes)
TR_VELO_TO_CAM = np.array([[7.533745e-03, -9.999714e-01, -6.166020e-04, -4.069766e-03],
                           [1.480249e-02, 7.280733e-04, -9.998902e-01, -7.631618e-02],
                           [9.998621e-01, 7.523790e-03, 1.480755e-02, -2.717806e-01]], dtype=np.float32)

POINT_CLOUD_RANGE = np.array([0.0, -40.0, -3.0, 70.4, 40.0, 1.0], dtype=np.float32)  # kitti_dataset.yaml:7
VOXEL_SIZE = (0.05, 0.05, 0.05)  # VirConv-L.yaml:105
IMG_W, IMG_H = 1216, 352  # PENet crop the virtual points are generated on


class Calib:
    """Attribute-compatible subset of `pcdet.utils.calibration_kitti.Calibration`
    (:60-76): the backbone's voxel->pixel projection only reads P2 / R0 / V2C."""

    def __init__(self, calib: dict | None = None):
        calib = calib or {'P2': P2, 'R0': R0, 'Tr_velo2cam': TR_VELO_TO_CAM}
        self.P2 = np.asarray(calib['P2'], dtype=np.float32).reshape(3, 4)
        self.R0 = np.asarray(calib['R0'], dtype=np.float32).reshape(3, 3)
        self.V2C = np.asarray(calib['Tr_velo2cam'], dtype=np.float32).reshape(3, 4)
        self.cu, self.cv = self.P2[0, 2], self.P2[1, 2]
        self.fu, self.fv = self.P2[0, 0], self.P2[1, 1]
        self.tx = self.P2[0, 3] / (-self.fu)
        self.ty = self.P2[1, 3] / (-self.fv)

    # calibration_kitti.py:164-174
    def img_to_rect(self, u, v, depth):
        x = ((u - self.cu) * depth) / self.fu + self.tx
        y = ((v - self.cv) * depth) / self.fv + self.ty
        return np.stack([x, y, depth], axis=1)

    # calibration_kitti.py:95-108
    def rect_to_lidar(self, pts_rect):
        hom = np.hstack([pts_rect, np.ones((pts_rect.shape[0], 1), dtype=np.float32)])
        r0 = np.eye(4, dtype=np.float32)
        r0[:3, :3] = self.R0
        v2c = np.eye(4, dtype=np.float32)
        v2c[:3, :4] = self.V2C
        return (hom @ np.linalg.inv((r0 @ v2c).T))[:, :3]


def _scene_depth(rng: np.random.Generator):
    """Facade depth per image column: piecewise constant over 48-px runs, U(8, 60) m."""
    n_runs = (IMG_W + 47) // 48 + 2
    runs = rng.uniform(8.0, 60.0, size=n_runs).astype(np.float32)
    return np.repeat(runs, 48)[:IMG_W + 48]


def _depth_at(u, v, facade, calib: Calib):
    """Depth (rect-z) seen at pixel (u, v): min(ground plane at camera-y 1.65 m, facade of that column)."""
    yprime = (v - calib.cv) / calib.fv
    with np.errstate(divide='ignore'):
        z_ground = np.where(yprime > 1e-6, 1.65 / np.maximum(yprime, 1e-6), np.inf)
    col = np.clip(u.astype(np.int64), 0, facade.shape[0] - 1)
    return np.minimum(z_ground, facade[col]).astype(np.float32)


def make_points(scene_id: int, n_lidar: int = 16384, n_virtual: int = 80000, calib: Calib | None = None):
    """One scene's fused cloud `[N, 8]` float32, LiDAR rows first (data_processor.py:152-155)."""
    calib = calib or Calib()
    rng = np.random.default_rng(scene_id)
    facade = _scene_depth(rng)

    # LiDAR: 64 elevation rings x uniform azimuth inside the camera FOV
    rings = 64
    per_ring = max(1, n_lidar // rings)
    elev = np.deg2rad(np.linspace(-24.8, 2.0, rings, dtype=np.float32))
    half_fov = math.atan((IMG_W / 2) / float(calib.fu))
    azim = np.linspace(-half_fov, half_fov, per_ring, dtype=np.float32)
    ee, aa = np.meshgrid(elev, azim, indexing='ij')
    # ray direction in the rect camera frame (x right, y down, z forward)
    dx = np.sin(aa) * np.cos(ee)
    dy = -np.sin(ee)
    dz = np.cos(aa) * np.cos(ee)
    u = calib.fu * dx / dz + calib.cu
    v = calib.fv * dy / dz + calib.cv
    depth = _depth_at(u.ravel(), v.ravel(), facade, calib)
    ok = np.isfinite(depth) & (depth < 80.0)
    pts_rect = calib.img_to_rect(u.ravel()[ok], v.ravel()[ok], depth[ok])
    lidar_xyz = calib.rect_to_lidar(pts_rect.astype(np.float32))
    lidar_xyz = lidar_xyz + rng.normal(0.0, 0.01, size=lidar_xyz.shape)
    lidar = np.zeros((lidar_xyz.shape[0], 8), dtype=np.float32)
    lidar[:, :3] = lidar_xyz
    lidar[:, 3] = rng.uniform(0.0, 1.0, size=lidar.shape[0])
    lidar[:, 7] = 2.0

    clouds = [lidar]
    if n_virtual > 0:
        pix = rng.choice(IMG_W * IMG_H, size=min(n_virtual, IMG_W * IMG_H), replace=False)
        vu = (pix % IMG_W).astype(np.float32)
        vv = (pix // IMG_W).astype(np.float32)
        vdepth = _depth_at(vu, vv, facade, calib)
        vdepth = vdepth * (1.0 + rng.normal(0.0, 0.01, size=vdepth.shape)).astype(np.float32)
        vok = np.isfinite(vdepth) & (vdepth > 0.5) & (vdepth < 80.0)
        vrect = calib.img_to_rect(vu[vok], vv[vok], vdepth[vok])
        vxyz = calib.rect_to_lidar(vrect.astype(np.float32))
        virt = np.zeros((vxyz.shape[0], 8), dtype=np.float32)
        virt[:, :3] = vxyz
        virt[:, 4:7] = rng.uniform(0.0, 85.0, size=(virt.shape[0], 3))
        virt[:, 7] = 1.0
        virt = virt[virt[:, 2] < 1.0]  # tools/PENet/dataloaders/my_loader.py:402
        clouds.append(virt)
    return np.concatenate(clouds).astype(np.float32)


def augment(points: np.ndarray, aug_param):
    """Forward world rotation -> flip -> scaling with explicit params (X_transform.py:49-120)."""
    rot, flip, scale = float(aug_param[0]), float(aug_param[1]), float(aug_param[2])
    c, s = math.cos(rot), math.sin(rot)
    rm = np.array([[c, s, 0.0], [-s, c, 0.0], [0.0, 0.0, 1.0]], dtype=np.float32)  # common_utils.py:45-54
    out = points.copy()
    out[:, :3] = out[:, :3] @ rm
    if flip:
        out[:, 1] = -out[:, 1]
    out[:, :3] *= np.float32(scale)
    return out


def mask_points_by_range(points, rng=POINT_CLOUD_RANGE):
    # common_utils.py:59-63 masks x/y only; the voxelizer drops z outliers itself
    m = (points[:, 0] >= rng[0]) & (points[:, 0] <= rng[3]) & (points[:, 1] >= rng[1]) & (points[:, 1] <= rng[4])
    return points[m]


def voxelize_first_come(points, voxel_size=VOXEL_SIZE, pc_range=POINT_CLOUD_RANGE,
                        max_points=5, max_voxels=40000):
    """Vectorised numpy statement of the first-come voxeliser the dataloader calls
    (`Point2VoxelCPU3d.point_to_voxel`, `data_processor.py:43-59`): voxel id in order of first
    appearance until `max_voxels`, at most `max_points` points per voxel in point order,
    zero padded; coordinates emitted zyx.  Returns (voxels [M,P,C], coords [M,3] i32, num [M] i32)."""
    pts = np.asarray(points, dtype=np.float32)
    vs = np.asarray(voxel_size, dtype=np.float32)
    lo = np.asarray(pc_range[:3], dtype=np.float32)
    grid = np.round((np.asarray(pc_range[3:], dtype=np.float64) - np.asarray(pc_range[:3], dtype=np.float64))
                    / np.asarray(voxel_size, dtype=np.float64)).astype(np.int64)
    c = np.floor((pts[:, :3] - lo) / vs).astype(np.int64)
    ok = np.all((c >= 0) & (c < grid), axis=1)
    idx = np.nonzero(ok)[0]
    c = c[ok]
    lin = (c[:, 2] * grid[1] + c[:, 1]) * grid[0] + c[:, 0]
    uniq, first, inv = np.unique(lin, return_index=True, return_inverse=True)
    order = np.argsort(first, kind='stable')           # voxels in order of first appearance
    rank_of_uniq = np.empty_like(order)
    rank_of_uniq[order] = np.arange(order.size)
    vid = rank_of_uniq[inv]                             # voxel id of every in-range point
    keep = vid < max_voxels
    vid, idx, c = vid[keep], idx[keep], c[keep]
    m = int(min(order.size, max_voxels))
    # slot of each point inside its voxel = number of earlier points in the same voxel
    srt = np.argsort(vid, kind='stable')
    vs_sorted = vid[srt]
    start = np.searchsorted(vs_sorted, np.arange(m))
    slot_sorted = np.arange(vs_sorted.size) - start[vs_sorted]
    slot = np.empty_like(slot_sorted)
    slot[srt] = slot_sorted
    voxels = np.zeros((m, max_points, pts.shape[1]), dtype=np.float32)
    take = slot < max_points
    voxels[vid[take], slot[take]] = pts[idx[take]]
    num = np.minimum(np.bincount(vid, minlength=m), max_points).astype(np.int32)
    coords = np.zeros((m, 3), dtype=np.int32)
    firstpt = srt[start]                                 # first point (in kept order) of each voxel
    coords[:, 0] = c[firstpt, 2]
    coords[:, 1] = c[firstpt, 1]
    coords[:, 2] = c[firstpt, 0]
    return voxels, coords, num


def mean_vfe(voxels, num, model='max'):
    """`MeanVFE.forward` (mean_vfe.py:39-47): per-voxel mean, last channel := max when MODEL == 'max'."""
    s = voxels.sum(axis=1, dtype=np.float32)
    out = s / np.maximum(num.astype(np.float32), 1.0)[:, None]
    if model == 'max':
        out[:, -1] = voxels.max(axis=1)[:, -1]
    return out.astype(np.float32)


@dataclass
class SceneBatch:
    """Host-side batch exactly as it sits in `batch_dict` after collate."""
    voxel_features: np.ndarray          # [M, 8] f32
    voxel_coords: np.ndarray            # [M, 4] f32 (b, z, y, x), batch-contiguous
    batch_size: int
    calib: list = field(default_factory=list)
    aug_param: np.ndarray | None = None  # [B, 3] f32 (rot, flip, scale), training only
    grid_size: tuple = (1408, 1600, 80)  # xyz

    def sparse_shape(self):
        gx, gy, gz = self.grid_size
        return [gz + 1, gy, gx]          # spconv_backbone.py:552


def make_batch(scene_ids, n_lidar=16384, n_virtual=80000, max_voxels=40000, training=False,
               voxel_size=VOXEL_SIZE, pc_range=POINT_CLOUD_RANGE) -> SceneBatch:
    feats, coords, calibs, augs = [], [], [], []
    for b, sid in enumerate(scene_ids):
        calib = Calib()
        pts = make_points(sid, n_lidar, n_virtual, calib)
        if training:
            rng = np.random.default_rng(10_000_019 * (sid + 1))
            aug = np.array([rng.uniform(-0.78539816, 0.78539816), float(rng.integers(0, 2)),
                            rng.uniform(0.95, 1.05)], dtype=np.float32)
            pts = augment(pts, aug)
            augs.append(aug)
        pts = mask_points_by_range(pts, pc_range)
        vox, c, num = voxelize_first_come(pts, voxel_size, pc_range, 5, max_voxels)
        f = mean_vfe(vox, num, 'max')
        cb = np.concatenate([np.full((c.shape[0], 1), b, dtype=np.int32), c], axis=1)
        feats.append(f)
        coords.append(cb.astype(np.float32))
        calibs.append(calib)
    gs = np.round((np.asarray(pc_range[3:], dtype=np.float64) - np.asarray(pc_range[:3], dtype=np.float64))
                  / np.asarray(voxel_size, dtype=np.float64)).astype(np.int64)
    return SceneBatch(np.concatenate(feats), np.concatenate(coords), len(scene_ids), calibs,
                      np.stack(augs) if training else None, tuple(int(g) for g in gs))


@dataclass
class PointsBatch:
    """The same scenes as `make_batch`, one step EARLIER in the reference's pipeline: the collated `batch_dict['points']`
    ([N, 1+8] f32 rows (b, x, y, z, intensity, r, g, b, indicator), samples contiguous in batch order, dataset.py:349-353)
    before `transform_points_to_voxels` — the input of the model-side voxeliser (preprocess.PointsToVoxels,
    graph.GraphedStep(voxelizer=...))."""
    points: np.ndarray
    batch_size: int
    calib: list = field(default_factory=list)
    aug_param: np.ndarray | None = None
    grid_size: tuple = (1408, 1600, 80)


def make_points_batch(scene_ids, n_lidar=16384, n_virtual=80000, training=False, pc_range=POINT_CLOUD_RANGE,
                      voxel_size=VOXEL_SIZE) -> PointsBatch:
    clouds, calibs, augs = [], [], []
    for b, sid in enumerate(scene_ids):
        calib = Calib()
        pts = make_points(sid, n_lidar, n_virtual, calib)
        if training:
            rng = np.random.default_rng(10_000_019 * (sid + 1))
            aug = np.array([rng.uniform(-0.78539816, 0.78539816), float(rng.integers(0, 2)),
                            rng.uniform(0.95, 1.05)], dtype=np.float32)
            pts = augment(pts, aug)
            augs.append(aug)
        pts = mask_points_by_range(pts, pc_range)
        clouds.append(np.concatenate([np.full((pts.shape[0], 1), b, dtype=np.float32), pts.astype(np.float32)], axis=1))
        calibs.append(calib)
    gs = np.round((np.asarray(pc_range[3:], dtype=np.float64) - np.asarray(pc_range[:3], dtype=np.float64))
                  / np.asarray(voxel_size, dtype=np.float64)).astype(np.int64)
    return PointsBatch(np.ascontiguousarray(np.concatenate(clouds)), len(scene_ids), calibs,
                       np.stack(augs) if training else None, tuple(int(g) for g in gs))


# ------------------------------------------------------------------------------------------------------
# VirConv-T / -S: LiDAR stream + virtual ("MM") stream (LATER_FUSION, dataset.py:270-281), optional test-time
# transformed copies (X_TRANS.input_transform, X_transform.py:156-193)
# ------------------------------------------------------------------------------------------------------
TEST_TRANSFORMS = np.array([[0.3, 0.0, 0.98], [0.3, 1.0, 1.02], [0.0, 1.0, 1.0]], dtype=np.float32)  # VirConv-T.yaml X_TRANS


@dataclass
class SceneBatchMM:
    """batch_dict arrays for VirConv8x: keys carry the reference's suffixes ('', '1', '2') per transformed copy."""
    arrays: dict                           # voxel_features{,i}, voxel_coords{,i}, voxel_features_mm{,i}, voxel_coords_mm{,i}
    batch_size: int
    calib: list
    aug_param: np.ndarray | None = None          # [B,3] training
    transform_param: np.ndarray | None = None    # [B,R,3] test
    grid_size: tuple = (1408, 1600, 80)


def _voxel_stream(pts, b, max_voxels, vfe_model, voxel_size, pc_range):
    pts = mask_points_by_range(pts, pc_range)
    vox, c, num = voxelize_first_come(pts, voxel_size, pc_range, 5, max_voxels)
    f = mean_vfe(vox, num, vfe_model)
    cb = np.concatenate([np.full((c.shape[0], 1), b, dtype=np.int32), c], axis=1).astype(np.float32)
    return f, cb


def make_batch_mm(scene_ids, n_lidar=16384, n_virtual=80000, max_voxels=16000, training=True, rot_num=1,
                  voxel_size=VOXEL_SIZE, pc_range=POINT_CLOUD_RANGE) -> SceneBatchMM:
    """training: one copy, random aug_param; test: `rot_num` copies transformed with TEST_TRANSFORMS[:rot_num]."""
    n_copies = 1 if training else rot_num
    parts = {(k, i): [] for k in ('voxel_features', 'voxel_coords', 'voxel_features_mm', 'voxel_coords_mm')
             for i in range(n_copies)}
    calibs, augs = [], []
    for b, sid in enumerate(scene_ids):
        calib = Calib()
        calibs.append(calib)
        pts = make_points(sid, n_lidar, n_virtual, calib)
        if training:
            rng = np.random.default_rng(10_000_019 * (sid + 1))
            aug = np.array([rng.uniform(-0.78539816, 0.78539816), float(rng.integers(0, 2)), rng.uniform(0.95, 1.05)],
                           dtype=np.float32)
            augs.append(aug)
            copies = [augment(pts, aug)]
        else:
            copies = [augment(pts, TEST_TRANSFORMS[i]) for i in range(n_copies)]
        for i, p in enumerate(copies):
            f, c = _voxel_stream(p[p[:, 7] == 2], b, max_voxels, 'max', voxel_size, pc_range)
            fm, cm = _voxel_stream(p[p[:, 7] == 1], b, max_voxels, None, voxel_size, pc_range)   # mm stream: plain mean
            parts[('voxel_features', i)].append(f)
            parts[('voxel_coords', i)].append(c)
            parts[('voxel_features_mm', i)].append(fm)
            parts[('voxel_coords_mm', i)].append(cm)
    arrays = {}
    for (k, i), v in parts.items():
        arrays[k + ('' if i == 0 else str(i))] = np.concatenate(v)
    gs = np.round((np.asarray(pc_range[3:], dtype=np.float64) - np.asarray(pc_range[:3], dtype=np.float64))
                  / np.asarray(voxel_size, dtype=np.float64)).astype(np.int64)
    tp = None if training else np.tile(TEST_TRANSFORMS[None, :n_copies], (len(scene_ids), 1, 1)).astype(np.float32)
    return SceneBatchMM(arrays, len(scene_ids), calibs, np.stack(augs) if training else None, tp,
                        tuple(int(g) for g in gs))
