"""Result containers of the ray path: ``Ray``, ``RayFan``, ``EigenRays``.

Output contract of the reference (/root/reference/src/pygenray/ray_objects.py:7-72, 75-430,
433-636): attribute names, array shapes, the negative-z storage convention (``z = -y[1]``,
``p = -y[2]``), ray-id strings and the indexing/concatenation quirks are
kept so downstream code written against pygenray keeps working.  Containers are host-side numpy;
they are filled from device arrival tables / trajectories by :mod:`pygenray_b200.launch_rays`.
The reference's plot methods are out of scope and not provided.
"""
import numpy as np
from scipy import io

__all__ = ["Ray", "RayFan", "EigenRays", "DeviceRows"]


class DeviceRows:
    """A float64 ``(M, N)`` array that stays in GPU memory; indexing copies just the requested rows / columns /
    slices to the host as numpy, ``np.asarray(x)`` fetches (and caches) everything.

    ``shoot_rays(..., device_trajectories=True)`` returns a ``RayFan`` whose ``ts`` / ``zs`` / ``ps`` are of this
    type: the 24 bytes per save point per ray -- 241 MB for a 10k-ray fan with 1001 save points, the whole
    end-to-end cost at 8 GPUs on one host -- cross PCIe only when, and as far as, somebody looks at them
    (``find_eigenrays`` reads one column, ``fan[i]`` one row)."""

    def __init__(self, tensor):
        self._t = tensor
        self._host = None

    shape = property(lambda self: tuple(self._t.shape))
    ndim = property(lambda self: self._t.ndim)
    size = property(lambda self: int(self._t.numel()))
    dtype = property(lambda self: np.dtype(np.float64))
    nbytes = property(lambda self: int(self._t.numel()) * 8)
    device_tensor = property(lambda self: self._t)

    def __len__(self):
        return int(self._t.shape[0])

    def __array__(self, dtype=None, copy=None):
        if self._host is None:
            self._host = self._t.cpu().numpy()
        return self._host if dtype is None else self._host.astype(dtype, copy=False)

    def _index(self, key):
        import torch  # noqa: PLC0415

        def conv(k):
            if isinstance(k, np.ndarray):
                return torch.from_numpy(np.ascontiguousarray(k)).to(self._t.device)
            if isinstance(k, (list,)):
                return torch.as_tensor(k, device=self._t.device)
            if isinstance(k, (np.integer,)):
                return int(k)
            return k

        return tuple(conv(k) for k in key) if isinstance(key, tuple) else conv(key)

    def __getitem__(self, key):
        if self._host is not None:
            return self._host[key]
        try:
            out = self._t[self._index(key)]
        except Exception:  # noqa: BLE001  (index forms torch does not take, e.g. negative steps)
            return np.asarray(self)[key]
        return out.cpu().numpy() if out.ndim else float(out.item())

    def min(self):
        return float(self._t.min().item())

    def max(self):
        return float(self._t.max().item())

    def __repr__(self):
        return f"DeviceRows(shape={self.shape}, device={self._t.device})"


class Ray:
    """One ray: ``r`` [m], ``t`` [s], ``z`` [m, negative down], ``p`` [s/m, sign-flipped]."""

    def __init__(self, r, y, n_bottom, n_surface, launch_angle=None, source_depth=None):
        self.r = r
        self.t = y[0, :]
        self.z = -y[1, :]
        self.p = -y[2, :]
        self.n_bottom = n_bottom
        self.n_surface = n_surface
        if launch_angle is not None:
            self.launch_angle = launch_angle
        if source_depth is not None:
            self.source_depth = source_depth


def _ray_id_numbers(ps, thetas):
    """(# sign changes of p on the saved grid) * sign(launch angle)  (ray_objects.py:142-144)."""
    return np.sum(np.diff(np.sign(ps)) != 0, axis=1) * np.sign(thetas)


class RayFan:
    """A fan of rays sharing one save grid: ``thetas (M,)``, ``rs/ts/zs/ps (M,N)``, bounce counts."""

    def __init__(self, Rays):
        self.thetas = np.array([ray.launch_angle for ray in Rays])
        self.rs = np.array([ray.r for ray in Rays])
        self.ts = np.array([ray.t for ray in Rays])
        self.zs = np.array([ray.z for ray in Rays])
        self.ps = np.array([ray.p for ray in Rays])
        self.n_botts = np.array([ray.n_bottom for ray in Rays])
        self.n_surfs = np.array([ray.n_surface for ray in Rays])
        self.source_depths = np.array([ray.source_depth for ray in Rays])
        self.compute_rayids()

    @classmethod
    def from_arrays(cls, thetas, rs, ts, zs, ps, n_botts, n_surfs, source_depths):
        """Build a fan directly from (M,N) arrays in STORED sign convention (no per-ray objects)."""
        fan = cls.__new__(cls)
        fan.thetas = np.asarray(thetas)
        fan.rs, fan.ts, fan.zs, fan.ps = rs, ts, zs, ps
        fan.n_botts = np.asarray(n_botts)
        fan.n_surfs = np.asarray(n_surfs)
        fan.source_depths = np.asarray(source_depths)
        fan.compute_rayids()
        return fan

    def compute_rayids(self):
        if self.rs.ndim != 2 or self.rs.shape[0] == 0:
            self.ray_ids = np.array([], dtype=str)
            return
        numbers = getattr(self, "_ray_id_numbers", None)  # counted on the device when the fan came from there
        if numbers is None:
            numbers = _ray_id_numbers(self.ps, self.thetas)
        self.ray_ids = _ray_id_strings(numbers, (self.n_botts == 0) & (self.n_surfs == 0))

    @property
    def ray_ids(self):
        """Ray identifiers ('-3.0', '4.0b', ...), built lazily for fans assembled from device arrays."""
        if self.__dict__.get("_ray_ids") is None:
            self.compute_rayids()
        return self.__dict__["_ray_ids"]

    @ray_ids.setter
    def ray_ids(self, value):
        self.__dict__["_ray_ids"] = value

    # ---- container protocol -------------------------------------------------------------
    def __len__(self):
        return len(self.thetas)

    def _ray_at(self, i):
        # re-negate: Ray.__init__ flips z and p again (ray_objects.py:388-395)
        return Ray(
            r=self.rs[i],
            y=np.array([self.ts[i], -self.zs[i], -self.ps[i]]),
            n_bottom=self.n_botts[i],
            n_surface=self.n_surfs[i],
            launch_angle=self.thetas[i],
            source_depth=self.source_depths[i],
        )

    def __getitem__(self, key):
        if isinstance(key, int):
            n = len(self.thetas)
            if key < 0:
                key += n
            if key < 0 or key >= n:
                raise IndexError(f"Index {key} is out of bounds for RayFan with {n} rays")
            return self._ray_at(key)
        if isinstance(key, slice):
            idx = np.arange(len(self.thetas))[key]
        else:
            idx = np.asarray(key)
            if idx.dtype == bool:
                idx = np.where(idx)[0]
        if np.isscalar(idx) or idx.ndim == 0:
            idx = [int(idx)]
        elif idx.ndim == 1:
            idx = idx.tolist()
        else:
            raise ValueError("Invalid indexing array shape")
        return RayFan([self._ray_at(i) for i in idx])

    def __add__(self, other):
        if not isinstance(other, RayFan):
            raise TypeError("Can only add RayFan objects together")
        if not np.array_equal(self.rs[0], other.rs[0]):
            raise ValueError("Range arrays (rs) must be equivalent for concatenation")
        rays = []
        for fan in (self, other):
            for i in range(len(fan.thetas)):
                # NOT re-negated: the sum of two fans has z and p flipped (reference quirk, Q3)
                rays.append(
                    Ray(
                        r=fan.rs[i],
                        y=np.array([fan.ts[i], fan.zs[i], fan.ps[i]]),
                        n_bottom=fan.n_botts[i],
                        n_surface=fan.n_surfs[i],
                        launch_angle=fan.thetas[i],
                        source_depth=fan.source_depths[i],
                    )
                )
        return RayFan(rays)

    def save_mat(self, filename):
        io.savemat(
            filename,
            {
                "rayfan": {
                    "thetas": self.thetas,
                    "xs": self.rs,
                    "ts": self.ts,
                    "zs": self.zs,
                    "ps": self.ps,
                    "n_botts": self.n_botts,
                    "n_surfs": self.n_surfs,
                    "source_depths": self.source_depths,
                }
            },
        )


def _ray_id_strings(numbers, free):
    """'-3.0', '4.0b', ...: the number, then 'b' when the ray touched a boundary (ray_objects.py:145-155, 536-541)."""
    numbers = np.ascontiguousarray(numbers, dtype=np.float64)
    if numbers.shape[0] == 0:
        return np.array([], dtype=str)
    # few distinct values: format each once (by bit pattern, so that 0.0 and -0.0 keep their own spelling)
    bits, inverse = np.unique(numbers.view(np.int64), return_inverse=True)
    text = np.array([str(v) for v in bits.view(np.float64)])
    return np.char.add(text[inverse], np.where(free, "", "b"))


def _received_angles_device(x, zs_end, ps_end, environment):
    """EigenRays.received_angles (ray_objects.py:519-531) for all rays of a receiver at once: ``prt_received_angles``
    on the stored end states and the un-flattened sound speed (reference quirk Q10), staged once per environment."""
    import ctypes as C  # noqa: PLC0415

    import torch  # noqa: PLC0415

    from . import _lib  # noqa: PLC0415

    L = _lib.require_gpu()
    n = int(np.shape(zs_end)[0])
    if n == 0:
        return np.zeros(0)
    dev = torch.device("cuda", torch.cuda.current_device())
    grid = unflattened_grid_on_device(environment, dev)
    with torch.cuda.device(dev):
        z_d = torch.from_numpy(np.ascontiguousarray(zs_end, dtype=np.float64)).to(dev)
        p_d = torch.from_numpy(np.ascontiguousarray(ps_end, dtype=np.float64)).to(dev)
        out = torch.empty(n, dtype=torch.float64, device=dev)
        rc = L.prt_received_angles(n, float(x), C.c_void_p(z_d.data_ptr()), C.c_void_p(p_d.data_ptr()), grid["n_r"],
                                   grid["n_z"], C.c_void_p(grid["rin"].data_ptr()), C.c_void_p(grid["zin"].data_ptr()),
                                   C.c_void_p(grid["cin"].data_ptr()), C.c_void_p(out.data_ptr()),
                                   C.c_void_p(torch.cuda.current_stream().cuda_stream))
        _lib.check(rc, "prt_received_angles")
        return out.cpu().numpy()


def unflattened_grid_on_device(environment, dev):
    """``environment.sound_speed`` (values + coordinates, NOT flat-earth transformed) as CUDA tensors, cached on the
    environment object per device: the grid ``ray_angle`` is evaluated on for received angles."""
    import torch  # noqa: PLC0415

    cache = environment.__dict__.setdefault("_prt_unflattened_cache", {})
    ss = environment.sound_speed
    cin = np.ascontiguousarray(np.asarray(ss.values), dtype=np.float64)
    key = (dev.index, cin.shape, float(cin[0, 0]), float(cin[-1, -1]), float(cin.sum()))
    hit = cache.get("grid")
    if hit is None or hit["key"] != key:
        rin = np.ascontiguousarray(np.asarray(ss.coords["range"].values), dtype=np.float64)
        zin = np.ascontiguousarray(np.asarray(ss.coords["depth"].values), dtype=np.float64)
        hit = dict(key=key, n_r=int(rin.shape[0]), n_z=int(zin.shape[0]), cin=torch.from_numpy(cin).to(dev),
                   rin=torch.from_numpy(rin).to(dev), zin=torch.from_numpy(zin).to(dev))
        cache["grid"] = hit
    return hit


class EigenRays:
    """Eigenrays per receiver depth (dicts keyed by receiver index), ray_objects.py:433-548.

    ``eigenray_dict[ridx]`` is a ``RayFan`` or a list of ``Ray`` (as in the reference).  No per-ray host loop: received
    angles come from ``prt_received_angles`` (or, when the fan was assembled by :func:`pygenray_b200.find_eigenrays`,
    from the device-built eigenray table it carries), ray ids from vectorised numpy.  A receiver without eigenrays
    yields empty arrays (the reference raises an AxisError there: ``RayFan([])``, ray_objects.py:135-144).
    """

    def __init__(
        self, receiver_depths, eigenray_dict, environment, num_eigenrays, num_eigenrays_found,
        failed_eray_theta_brackets,
    ):
        self.receiver_depths = receiver_depths
        self.rs, self.ts, self.zs, self.ps = {}, {}, {}, {}
        self.received_angles, self.launch_angles = {}, {}
        self.n_botts, self.n_surfs = {}, {}
        self.ray_id, self.ray_id_int = {}, {}
        self.num_eigenrays = num_eigenrays
        self.num_eigenrays_found = num_eigenrays_found
        self.failed_eray_theta_brackets = failed_eray_theta_brackets

        for ridx in range(len(receiver_depths)):
            fan = eigenray_dict[ridx]
            if not isinstance(fan, RayFan):
                fan = RayFan(fan)
            self.rs[ridx], self.ts[ridx], self.zs[ridx], self.ps[ridx] = fan.rs, fan.ts, fan.zs, fan.ps
            self.n_botts[ridx], self.n_surfs[ridx] = fan.n_botts, fan.n_surfs
            n_rays = fan.rs.shape[0] if fan.rs.ndim == 2 else 0
            table = getattr(fan, "_eigen_table", None)  # (received angles, ray id numbers) built on the device
            if n_rays == 0:
                recv, numbers = np.zeros(0), np.zeros(0)
            elif table is not None:
                recv, numbers = table
            else:
                # received angle from the STORED (negative-z) state and the un-flattened sound speed --
                # numerically odd but it is the reference's output
                recv = _received_angles_device(fan.rs[0, -1], fan.zs[:, -1], fan.ps[:, -1], environment)
                numbers = _ray_id_numbers(fan.ps, fan.thetas)
            free = (np.asarray(fan.n_botts) == 0) & (np.asarray(fan.n_surfs) == 0) if n_rays else np.zeros(0, bool)
            self.received_angles[ridx] = np.asarray(recv, dtype=np.float64)
            self.launch_angles[ridx] = fan.thetas
            self.ray_id[ridx] = _ray_id_strings(numbers, free)
            self.ray_id_int[ridx] = np.asarray(numbers).astype(np.int64)

    def save_mat(self, filename):
        data = {}
        for ridx, rdepth in enumerate(self.receiver_depths):
            data[f"receiver_depth_{ridx}"] = {
                "receiver_depth": rdepth,
                "xs": self.rs[ridx],
                "ts": self.ts[ridx],
                "zs": self.zs[ridx],
                "ps": self.ps[ridx],
                "received_angles": self.received_angles[ridx],
                "launch_angles": self.launch_angles[ridx],
                "ray_id": self.ray_id[ridx],
                "ray_id_int": self.ray_id_int[ridx],
                "n_bottom": self.n_botts[ridx],
                "n_surface": self.n_surfs[ridx],
                "source_depth": np.nan,
                "num_eigenrays": self.num_eigenrays,
                "num_eigenrays_found": self.num_eigenrays_found,
            }
        io.savemat(filename, {"eigenrays": data})
