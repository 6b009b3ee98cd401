"""``xr.Dataset`` stand-in used when xarray is not installed.

``ODESampleSimulator.simulate_samples`` returns an ``xr.Dataset`` with dims (sim, time) and one
variable per observable named "[sid]" (reference petab_factory.py:269-272).  xarray is not part
of this image; ``SimDataset`` offers the subset of that interface the reference's callers use
(``data_vars``, ``["sim"]/["time"]`` coords, ``dset[var].isel(sim=k)``, ``isel(sim=k).to_dataframe()``,
``to_xarray()`` when xarray is present).
"""
from __future__ import annotations

from typing import Iterable

import numpy as np
import pandas as pd


class _Var:
    def __init__(self, values: np.ndarray, dims: tuple[str, ...]):
        self.values = values
        self.dims = dims

    def isel(self, **idx):
        v = self.values
        dims = list(self.dims)
        for d, i in idx.items():
            ax = dims.index(d)
            v = np.take(v, i, axis=ax)
            if np.ndim(i) == 0:
                dims.pop(ax)
        return _Var(v, tuple(dims))

    def __array__(self, dtype=None, copy=None):
        return np.asarray(self.values, dtype=dtype)

    def __len__(self):
        return len(self.values)

    @property
    def shape(self):
        return self.values.shape


class SimDataset:
    """dims (sim, time); data_vars: {"[sid]": float64[sim, time]}."""

    def __init__(self, time: np.ndarray, data: dict[str, np.ndarray], sim: np.ndarray | None = None):
        self.coords = {"time": np.asarray(time, dtype=float)}
        n = next(iter(data.values())).shape[0] if data else 0
        self.coords["sim"] = np.arange(n) if sim is None else np.asarray(sim)
        self.data_vars = dict(data)
        self.dims = {"sim": len(self.coords["sim"]), "time": len(self.coords["time"])}

    def __getitem__(self, key: str) -> _Var:
        if key in self.coords:
            return _Var(self.coords[key], (key,))
        return _Var(self.data_vars[key], ("sim", "time"))

    def __iter__(self) -> Iterable[str]:
        return iter(self.data_vars)

    def isel(self, sim: int) -> "SimDataset":
        return SimDataset(self.coords["time"], {k: v[sim:sim + 1] for k, v in self.data_vars.items()},
                          sim=self.coords["sim"][sim:sim + 1])

    def to_dataframe(self) -> pd.DataFrame:
        n, T = self.dims["sim"], self.dims["time"]
        idx = pd.MultiIndex.from_product([self.coords["sim"], self.coords["time"]], names=["sim", "time"])
        df = pd.DataFrame({k: v.reshape(n * T) for k, v in self.data_vars.items()}, index=idx)
        if n == 1:
            df = df.droplevel("sim")
        return df

    def to_netcdf(self, path) -> None:
        """``xr.Dataset.to_netcdf`` for the call in the reference's workflow (``dset.to_netcdf(xp_path / f"{category}.nc")``,
        petab_factory.py:601-602): a NetCDF-3 classic file with dims (sim, time), the two coordinates and one float64
        variable per observable, written with scipy (no netCDF4 / h5netcdf needed); xarray opens it as the same Dataset."""
        from scipy.io import netcdf_file
        with netcdf_file(str(path), "w") as f:
            f.createDimension("sim", self.dims["sim"])
            f.createDimension("time", self.dims["time"])
            sim = f.createVariable("sim", "i4", ("sim",))
            sim[:] = np.asarray(self.coords["sim"], dtype=np.int32)
            time = f.createVariable("time", "d", ("time",))
            time[:] = self.coords["time"]
            for name, values in self.data_vars.items():
                var = f.createVariable(name, "d", ("sim", "time"))
                var[:] = values

    def to_xarray(self):
        import xarray as xr   # optional
        return xr.Dataset({k: (("sim", "time"), v) for k, v in self.data_vars.items()},
                          coords={"sim": self.coords["sim"], "time": self.coords["time"]})
