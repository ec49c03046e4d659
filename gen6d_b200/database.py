"""In-memory object databases for the estimator.

The reference's dataset/database.py readers (LINEMOD, GenMOP, COLMAP projects ...) are disk
formats and out of scope.  The estimator only needs the small interface
below; `SyntheticObjectDatabase` provides it without any file, for tests and bench.py, and
`ReferenceDatabaseAdapter` wraps a reference `BaseDatabase` object for drop-in use.
"""
import cv2
import numpy as np


class ObjectDatabase:
    """Interface the estimator consumes (mirrors BaseDatabase.get_* plus the
    get_object_center / get_diameter / get_object_vert free functions, database.py:30-54,346-397)."""
    database_name = 'object'

    def get_image(self, img_id):
        raise NotImplementedError

    def get_K(self, img_id):
        raise NotImplementedError

    def get_pose(self, img_id):
        raise NotImplementedError

    def get_img_ids(self):
        raise NotImplementedError

    def object_center(self):
        raise NotImplementedError

    def object_diameter(self):
        return 2.0

    def object_vert(self):
        return np.asarray([0, 0, 1], np.float32)


class ReferenceDatabaseAdapter(ObjectDatabase):
    """Wraps a reference-repo database object (needs the reference's `dataset.database` importable)."""

    def __init__(self, ref_db):
        from dataset.database import get_diameter, get_object_center, get_object_vert  # reference package
        self.db = ref_db
        self.database_name = ref_db.database_name
        self._c, self._d, self._v = get_object_center(ref_db), get_diameter(ref_db), get_object_vert(ref_db)

    def get_image(self, i): return self.db.get_image(i)
    def get_K(self, i): return self.db.get_K(i)
    def get_pose(self, i): return self.db.get_pose(i)
    def get_img_ids(self): return self.db.get_img_ids()
    def object_center(self): return self._c
    def object_diameter(self): return self._d
    def object_vert(self): return self._v


def as_object_database(db):
    """`db` itself if it already offers the ObjectDatabase interface, else a ReferenceDatabaseAdapter
    around it (a reference-repo BaseDatabase: CustomDatabase, LINEMOD, GenMOP ...)."""
    if all(hasattr(db, m) for m in ('object_center', 'object_diameter', 'object_vert')):
        return db
    return ReferenceDatabaseAdapter(db)


def _look_at(cam, up):
    z = -cam / np.linalg.norm(cam)
    x = np.cross(z, up)
    x /= np.linalg.norm(x)
    y = np.cross(z, x)
    R = np.stack([x, y, z], 0)
    return np.concatenate([R, (-R @ cam)[:, None]], 1)


class SyntheticObjectDatabase(ObjectDatabase):
    """A procedurally rendered object (a cloud of coloured blobs inside the unit sphere) seen from
    `n_views` cameras on a jittered upper hemisphere.  Fully determined by `seed`."""

    def __init__(self, n_views=80, height=480, width=640, seed=0, n_blobs=260, radius=5.0, name='synthetic/blobs'):
        self.database_name = name
        rng = np.random.RandomState(seed)
        self.h, self.w = height, width
        # object: blobs on a bumpy ellipsoid shell, diameter 2, centre at the origin
        d = rng.randn(n_blobs, 3)
        d /= np.linalg.norm(d, axis=1, keepdims=True)
        self.points = (d * np.array([0.55, 0.8, 0.95]) * (0.85 + 0.15 * rng.rand(n_blobs, 1))).astype(np.float64)
        self.colors = rng.randint(30, 255, size=(n_blobs, 3))
        self.sizes = 0.05 + 0.07 * rng.rand(n_blobs)
        self.object_point_cloud = self.points.astype(np.float32)
        self.center = np.zeros(3, np.float32)
        f = 0.9 * width
        self.K = np.array([[f, 0, width / 2], [0, f, height / 2], [0, 0, 1]], np.float32)
        self.img_ids = [str(i) for i in range(n_views)]
        self.poses, self.Ks, self._imgs = {}, {}, {}
        bg = cv2.resize(rng.randint(60, 200, size=(height // 40, width // 40, 3)).astype(np.uint8), (width, height),
                        interpolation=cv2.INTER_CUBIC)
        self._bg = bg
        for i in self.img_ids:
            v = rng.randn(3)
            v[2] = abs(v[2]) * 0.8 + 0.15
            v /= np.linalg.norm(v)
            cam = v * radius * (0.9 + 0.2 * rng.rand())
            up = np.array([0, 0, 1.0]) + 0.15 * rng.randn(3)
            pose = _look_at(cam, up / np.linalg.norm(up))
            # push the object off-centre by a small camera rotation so detection is not trivial
            ax, ay = 0.08 * rng.randn(2)
            Rx = np.array([[1, 0, 0], [0, np.cos(ax), -np.sin(ax)], [0, np.sin(ax), np.cos(ax)]])
            Ry = np.array([[np.cos(ay), 0, np.sin(ay)], [0, 1, 0], [-np.sin(ay), 0, np.cos(ay)]])
            Rj = Rx @ Ry
            pose = np.concatenate([Rj @ pose[:, :3], Rj @ pose[:, 3:]], 1)
            self.poses[i] = pose.astype(np.float32)
            self.Ks[i] = self.K.copy()

    def render(self, pose, K=None):
        K = self.K if K is None else K
        img = self._bg.copy()
        p = self.points @ pose[:, :3].T.astype(np.float64) + pose[:, 3].astype(np.float64)
        order = np.argsort(-p[:, 2])
        for j in order:
            z = p[j, 2]
            if z < 0.1:
                continue
            u, v = K[0, 0] * p[j, 0] / z + K[0, 2], K[1, 1] * p[j, 1] / z + K[1, 2]
            r = max(1, int(round(K[0, 0] * self.sizes[j] / z)))
            cv2.circle(img, (int(round(u)), int(round(v))), r, tuple(int(c) for c in self.colors[j]), -1,
                       lineType=cv2.LINE_AA)
        return img

    def get_image(self, img_id):
        if img_id not in self._imgs:
            self._imgs[img_id] = self.render(self.poses[img_id], self.Ks[img_id])
        return self._imgs[img_id]

    def get_K(self, img_id):
        return self.Ks[img_id].copy()

    def get_pose(self, img_id):
        return self.poses[img_id].copy()

    def get_img_ids(self):
        return list(self.img_ids)

    def object_center(self):
        return self.center
