"""Array container I/O for the input ``*.snps.hdf5`` and the working ``{name}.database.hdf5``.

The reference reads and writes these with h5py (tetrad/src/write_database.py:132-168,
tetrad/src/resolve_quartets.py:33-35).  h5py/libhdf5 are not part of this image, so:

* a path ending in ``.npz`` (or any file that starts with the zip magic) is a numpy archive
  with the same dataset names (``snps``, ``snpsmap``, ``names`` / ``seqarr``, ``spans``,
  ``tmparr``, ``tmpmap``);
* a real HDF5 file is read/written through h5py when it can be imported, else through the
  bundled subset implementation ``tetrad_b200.hdf5`` (contiguous datasets, fixed-length string
  attributes -- what ipyrad and this package write).
"""
from __future__ import annotations

from pathlib import Path

import numpy as np

_ZIP_MAGIC = b"PK\x03\x04"
_HDF5_MAGIC = b"\x89HDF\r\n\x1a\n"


def _kind(path: Path) -> str:
    path = Path(path)
    if path.exists():
        with open(path, "rb") as fh:
            head = fh.read(8)
        if head.startswith(_ZIP_MAGIC):
            return "npz"
        if head == _HDF5_MAGIC:
            return "hdf5"
    return "npz" if path.suffix == ".npz" else "hdf5"


def _h5():
    try:
        import h5py  # noqa: F401
        return h5py
    except ImportError:
        from . import hdf5
        return hdf5


def read_arrays(path: Path, names) -> dict:
    """Datasets ``names`` of ``path`` as numpy arrays (missing names are skipped)."""
    path = Path(path)
    out = {}
    if _kind(path) == "npz":
        with np.load(path, allow_pickle=False) as io5:
            for key in names:
                if key in io5.files:
                    out[key] = io5[key]
        return out
    with _h5().File(path, "r") as io5:
        for key in names:
            if key in io5:
                out[key] = io5[key][:]
    return out


def read_snps_input(path: Path):
    """(snps u8[n,S], snpsmap int[S,>=2], names list[str]) of an ipyrad-style SNP file
    (reference write_database.py:57-78,132-141)."""
    path = Path(path)
    if _kind(path) == "npz":
        with np.load(path, allow_pickle=False) as io5:
            snps, snpsmap = io5["snps"], io5["snpsmap"]
            names = [n.decode() if isinstance(n, bytes) else str(n) for n in io5["names"]]
        return snps, snpsmap, names
    with _h5().File(path, "r") as io5:
        snps = io5["snps"][:]
        snpsmap = io5["snpsmap"][:]
        names = [n.decode("utf-8") if isinstance(n, bytes) else str(n) for n in io5["snps"].attrs["names"]]
    return snps, snpsmap, names


def write_snps_input(path: Path, snps, snpsmap, names) -> None:
    """Write an input SNP file (used by the synthetic generator and the tests)."""
    path = Path(path)
    if path.suffix == ".npz":
        np.savez(path, snps=snps, snpsmap=snpsmap, names=np.array([n.encode() for n in names]))
        return
    with _h5().File(path, "w") as io5:
        dset = io5.create_dataset("snps", data=np.asarray(snps, dtype=np.uint8))
        dset.attrs["names"] = np.array([n.encode() for n in names])
        io5.create_dataset("snpsmap", data=np.asarray(snpsmap))


def write_arrays(path: Path, arrays: dict, as_npz: bool | None = None) -> None:
    """(Re)write the working database with exactly ``arrays`` (the reference deletes and
    recreates ``tmparr``/``tmpmap`` every replicate, run_inference.py:138-143)."""
    path = Path(path)
    if as_npz is None:
        as_npz = path.suffix == ".npz" or (path.exists() and _kind(path) == "npz")
    if as_npz:
        with open(path, "wb") as fh:
            np.savez(fh, **arrays)
        return
    with _h5().File(path, "w") as io5:
        for key, val in arrays.items():
            io5.create_dataset(key, data=val)
