"""Golden vectors for the ClimEx sample preparation, produced by the UNMODIFIED reference methods
(`climex_utils.climex2torch.__getitem__`, `compute_stats`, `residual_to_hr`).  Run in the authoring container:

    python tests/golden/make_golden_climex.py

`climex_utils` imports xarray / dask / cartopy / matplotlib at module level; those packages are absent here and are never
touched by the three methods, so they are stubbed in `sys.modules` before the import.  The dataset object is created
without running `__init__` (which reads NetCDF files) and given the attributes `__init__` would have set."""
import importlib.abc
import importlib.machinery
import os
import sys
import types

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))


STUBS = ('xarray', 'dask', 'cartopy', 'matplotlib', 'bottleneck', 'h5netcdf', 'netCDF4', 'seaborn', 'wandb')


class _Anything(types.ModuleType):
    __path__ = []                       # looks like a package: `import stub.sub` goes back through the finder

    def __getattr__(self, name):
        if name.startswith('__'):
            raise AttributeError(name)
        return _Anything(self.__name__ + '.' + name)

    def __call__(self, *a, **k):
        return self


class _StubFinder(importlib.abc.MetaPathFinder, importlib.abc.Loader):
    """Serves an empty stand-in for the plotting / NetCDF packages that are not installed here."""

    def find_spec(self, fullname, path=None, target=None):
        if fullname.split('.')[0] in self.missing:
            return importlib.machinery.ModuleSpec(fullname, self, is_package=True)
        return None

    def create_module(self, spec):
        return _Anything(spec.name)

    def exec_module(self, module):
        pass


def load_reference():
    finder = _StubFinder()
    finder.missing = set()
    for name in STUBS:
        try:
            __import__(name)
        except Exception:  # noqa: BLE001
            finder.missing.add(name)
    sys.meta_path.append(finder)
    sys.path.insert(0, '/root/reference')
    import climex_utils
    return climex_utils


def make_dataset(cu, hr_all, standardization, lowres_scale=4):
    ds = object.__new__(cu.climex2torch)
    ds.hr = hr_all
    ds.lowres_scale = lowres_scale
    ds.standardization = standardization
    ds.epsilon = 1e-10
    ds.lrstats = None
    ds.timestamps = np.arange(hr_all.shape[0])
    return ds


def main():
    cu = load_reference()
    g = torch.Generator().manual_seed(5)
    T, Cc, H, W = 4, 3, 16, 24
    hr_all = torch.randn(T, Cc, H, W, generator=g) * torch.tensor([2.0, 8.0, 9.0]).view(1, 3, 1, 1) + \
        torch.tensor([1.0, -5.0, 4.0]).view(1, 3, 1, 1)
    residual = torch.randn(T, Cc, H, W, generator=g)
    out = {'hr_all': hr_all.numpy(), 'residual': residual.numpy()}
    for mode in ('none', 'perpixel', 'pertimestep', 'minmax'):
        ds = make_dataset(cu, hr_all, mode)
        items = [ds[i] for i in range(T)]
        out[f'{mode}_inputs'] = torch.stack([it['inputs'] for it in items]).numpy()
        out[f'{mode}_targets'] = torch.stack([it['targets'] for it in items]).numpy()
        out[f'{mode}_lr'] = torch.stack([it['lr'] for it in items]).numpy()
        out[f'{mode}_lrinterp'] = torch.stack([it['lrinterp'] for it in items]).numpy()
        if mode != 'none':
            out[f'{mode}_s0'] = ds.lrstats[0].numpy()
            out[f'{mode}_s1'] = ds.lrstats[1].numpy()
        hp = torch.stack([ds.residual_to_hr(residual[i], items[i]['lrinterp'], items[i]['stand_stats']) for i in range(T)])
        out[f'{mode}_hr_pred'] = hp.numpy()
    path = os.path.join(HERE, 'climex_prepare.npz')
    np.savez_compressed(path, **out)
    print('wrote', path, {k: v.shape for k, v in out.items()})


if __name__ == '__main__':
    main()
