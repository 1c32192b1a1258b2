"""Result containers of the ray path: ``Ray``, ``RayFan``, ``EigenRays``.

Output contract of the reference (/root/reference/src/pygenray/ray_objects.py:7-72, 75-430,
433-636): attribute names, array shapes, the negative-z storage convention (``z = -y[1]``,
``p = -y[2]``), ray-id strings and the indexing/concatenation quirks are
kept so downstream code written against pygenray keeps working.  Containers are host-side numpy;
they are filled from device arrival tables / trajectories by :mod:`pygenray_b200.launch_rays`.
Plot methods import matplotlib lazily (not installed in this image; cosmetic, out of scope).
"""
import numpy as np
from scipy import io

__all__ = ["Ray", "RayFan", "EigenRays"]


def _plt():
    from matplotlib import pyplot as plt  # noqa: PLC0415 - optional dependency

    return plt


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

    def plot(self, **kwargs):
        plt = _plt()
        plt.plot(self.r, self.z, **kwargs)
        plt.xlabel("time [s]")
        plt.ylabel("depth [m]")
        plt.ylim([self.z.min(), self.z.max()])


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
        free = (self.n_botts == 0) & (self.n_surfs == 0)
        self.ray_ids = np.array([str(n) if f else f"{n}b" for n, f in zip(numbers, free)])

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

    # ---- plots (optional matplotlib) ------------------------------------------------------
    def plot_time_front(self, include_lines=False, range_idx=-1, add_colorbar=True, ray_id=False, **kwargs):
        plt = _plt()
        if include_lines:
            plt.plot(self.ts[:, range_idx], self.zs[:, range_idx], c="#aaaaaa", lw=0.5, zorder=5)
        style = {"c": self.thetas, "cmap": "managua", "s": 2, "lw": 0, "zorder": 6}
        style.update(kwargs)
        if ray_id:
            cats = np.unique(self.ray_ids)
            colors = plt.cm.tab20(np.linspace(0, 1, len(cats)))
            lut = dict(zip(cats, colors))
            style["c"] = [lut[c] for c in self.ray_ids]
            add_colorbar = False
            for i, cat in enumerate(cats):
                plt.scatter([], [], c=[colors[i]], label=cat)
            plt.legend(ncols=3, loc="lower left")
        plt.scatter(x=self.ts[:, range_idx], y=self.zs[:, range_idx], **style)
        plt.ylim([self.zs.min(), self.zs.max()])
        if add_colorbar:
            plt.colorbar(label="launch angle [degrees]")
            plt.xlabel("time [s]")
            plt.ylabel("depth [m]")
            plt.title("Time Front")

    def plot_ray_fan(self, **kwargs):
        plt = _plt()
        alpha = min(1.0, 10.0 / max(len(self.thetas), 1))
        style = {"c": "k", "lw": 1, "alpha": alpha}
        style.update(kwargs)
        plt.plot(self.rs.T, self.zs.T, **style)
        plt.xlabel("range [m]")
        plt.ylabel("depth [m]")
        plt.ylim([self.zs.min(), self.zs.max()])
        plt.title("Ray Fan")

    def plot_depth_v_angle(self, include_line=False, **kwargs):
        plt = _plt()
        if include_line:
            plt.plot(self.thetas, self.zs[:, -1], c="#aaaaaa", lw=0.5, zorder=5)
        plt.scatter(x=self.thetas, y=self.zs[:, -1], **kwargs)


def _ray_angle_host(x, y, cin, rin, zin):
    """Host evaluation of ray_angle (integration_processes.py:306-334) for EigenRays.received_angles.

    Same cell rule and weight expression as the reference's bilinear_interp, in numpy scalars.
    """
    i = int(np.searchsorted(rin, x)) - 1
    j = int(np.searchsorted(zin, y[1])) - 1
    i = max(0, min(i, len(rin) - 2))
    j = max(0, min(j, len(zin) - 2))
    wx = (x - rin[i]) / (rin[i + 1] - rin[i])
    wy = (y[1] - zin[j]) / (zin[j + 1] - zin[j])
    c = (
        (1 - wx) * (1 - wy) * cin[i, j]
        + wx * (1 - wy) * cin[i + 1, j]
        + (1 - wx) * wy * cin[i, j + 1]
        + wx * wy * cin[i + 1, j + 1]
    )
    with np.errstate(invalid="ignore"):
        theta = np.degrees(np.arcsin(y[2] * c))
    return theta, c


class EigenRays:
    """Eigenrays per receiver depth (dicts keyed by receiver index), ray_objects.py:433-548."""

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

        ss = environment.sound_speed
        cin = np.asarray(ss.values)
        rin = np.asarray(ss.coords["range"].values)
        zin = np.asarray(ss.coords["depth"].values)

        for ridx in range(len(receiver_depths)):
            fan = eigenray_dict[ridx]
            if not isinstance(fan, RayFan):
                fan = RayFan(fan)
            self.rs[ridx], self.ts[ridx], self.zs[ridx], self.ps[ridx] = fan.rs, fan.ts, fan.zs, fan.ps
            self.n_botts[ridx], self.n_surfs[ridx] = fan.n_botts, fan.n_surfs
            angles, ids, ids_int = [], [], []
            n_rays = fan.rs.shape[0] if fan.rs.ndim == 2 else 0
            for k in range(n_rays):
                # received angle from the STORED (negative-z) state and the un-flattened sound
                # speed -- numerically odd but it is the reference's output
                y_last = np.stack((fan.ts[k, -1], fan.zs[k, -1], fan.ps[k, -1]))
                theta, _ = _ray_angle_host(fan.rs[k, -1], y_last, cin, rin, zin)
                angles.append(theta)
                number = np.sum(np.diff(np.sign(fan.ps[k, :])) != 0) * np.sign(fan.thetas[k])
                free = fan.n_botts[k] == 0 and fan.n_surfs[k] == 0
                ids.append(f"{number}{'' if free else 'b'}")
                ids_int.append(int(number))
            self.received_angles[ridx] = np.array(angles)
            self.launch_angles[ridx] = fan.thetas
            self.ray_id[ridx] = np.array(ids)
            self.ray_id_int[ridx] = np.array(ids_int)

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

    def plot_angle_time(self, ridxs=None, **kwargs):
        plt = _plt()
        for ridx in ridxs if ridxs is not None else list(self.received_angles.keys()):
            plt.scatter(self.ts[ridx][:, -1], self.received_angles[ridx], **kwargs)
        plt.xlabel("time [s]")
        plt.ylabel("received angle [deg]")
        plt.title("Received Angle vs Time")

    def plot(self, ridxs=(0,), **kwargs):
        plt = _plt()
        if isinstance(ridxs, int):
            ridxs = [ridxs]
        style = {"c": "k"}
        style.update(kwargs)
        for ridx in ridxs:
            plt.plot(self.rs[ridx].T, self.zs[ridx].T, **style)
        plt.xlabel("range [m]")
        plt.ylabel("depth [m]")
        plt.title("Eigen Rays")

    def plot_ducted(self, **kwargs):
        plt = _plt()
        style = {"c": "k"}
        style.update(kwargs)
        for ridx in self.ray_id:
            free = (self.n_botts[ridx] == 0) & (self.n_surfs[ridx] == 0)
            plt.plot(self.rs[ridx][free].T, -self.zs[ridx][free].T, **style)
        plt.xlabel("range [m]")
        plt.ylabel("depth [m]")
        plt.title("Ducted Eigen Rays")
