"""CPU: the bundled HDF5 subset (used when h5py is absent): round trips of what this package writes,
and reading of structures it does not write itself (v2 object headers with link messages, chunked +
shuffle + deflate datasets), assembled here byte by byte from the format specification."""
import struct
import zlib

import numpy as np
import pytest

from tetrad_b200 import database, hdf5, synth


def test_roundtrip_database_arrays(tmp_path):
    rng = np.random.default_rng(0)
    arrays = dict(seqarr=rng.integers(0, 255, (7, 333)).astype(np.uint8), spans=rng.integers(0, 10**9, (50, 2)).astype(np.int64),
                  tmparr=rng.integers(0, 4, (7, 333)).astype(np.uint8), tmpmap=rng.integers(0, 2**32, (333, 2)).astype(np.uint32),
                  f4=rng.random((3, 5)).astype(np.float32), f8=rng.random(9), empty=np.zeros((0, 2), np.uint32))
    path = tmp_path / "t.hdf5"
    with hdf5.File(path, "w", libver="latest") as f:
        for k, v in arrays.items():
            f.create_dataset(k, data=v)
    assert path.read_bytes()[:8] == hdf5.SIGNATURE
    with hdf5.File(path, "r", swmr=True) as f:
        assert sorted(f.keys()) == sorted(arrays)
        for k, v in arrays.items():
            assert k in f and f[k].shape == v.shape and f[k].dtype == v.dtype
            assert (f[k][:] == v).all()
        assert (f["tmparr"][2:4] == arrays["tmparr"][2:4]).all()
        assert "nope" not in f
        with pytest.raises(KeyError):
            f["nope"]


def test_snps_input_with_names_attribute(tmp_path):
    d = synth.generate(9, 30, 120, seed=4)
    path = tmp_path / "x.snps.hdf5"
    database.write_snps_input(path, d.snps, d.snpsmap, d.names)
    snps, snpsmap, names = database.read_snps_input(path)
    assert (snps == d.snps).all() and (snpsmap == d.snpsmap).all() and names == d.names
    with hdf5.File(path, "r") as f:
        assert f["snps"].attrs["names"].dtype.kind == "S" and len(f["snps"].attrs["names"]) == 9


def test_rewrite_and_delete_like_the_reference(tmp_path):
    """resample_tmp_database deletes and recreates tmparr/tmpmap (run_inference.py:138-143)."""
    path = tmp_path / "db.hdf5"
    with hdf5.File(path, "w") as f:
        f.create_dataset("seqarr", data=np.arange(12, dtype=np.uint8).reshape(3, 4))
        f.create_dataset("tmparr", data=np.zeros((3, 4), np.uint8))
    with hdf5.File(path, "r+") as f:
        del f["tmparr"]
        f.create_dataset("tmparr", data=np.ones((3, 6), np.uint8))
    with hdf5.File(path, "r") as f:
        assert f["tmparr"].shape == (3, 6) and (f["seqarr"][:] == np.arange(12).reshape(3, 4)).all()


def test_many_datasets_span_several_symbol_nodes(tmp_path):
    path = tmp_path / "many.hdf5"
    with hdf5.File(path, "w") as f:
        for i in range(20):
            f.create_dataset(f"d{i:02d}", data=np.full((2,), i, np.int64))
    with hdf5.File(path, "r") as f:
        assert len(f.keys()) == 20 and all(int(f[f"d{i:02d}"][0]) == i for i in range(20))


def _v2_header(messages):
    body = b"".join(struct.pack("<BHB", t, len(d), 0) + d for t, d in messages)
    hdr = b"OHDR" + bytes([2, 0x02]) + struct.pack("<I", len(body)) + body      # flags: 4-byte chunk size
    return hdr + b"\0\0\0\0"                                                     # checksum (not verified)


def test_reads_v2_headers_links_and_chunked_deflate(tmp_path):
    """A file in the style libver='latest' + chunks + gzip + shuffle produces: superblock v2, OHDR
    headers, a compact group of link messages, layout v3 chunked with a v1 B-tree, filter pipeline v2."""
    data = (np.arange(7 * 10, dtype=np.uint16).reshape(7, 10) * 37) % 1000
    chunk = (4, 6)
    buf = bytearray(48)                                           # superblock v2 placeholder
    def put(b):
        while len(buf) % 8:
            buf.append(0)
        addr = len(buf)
        buf.extend(b)
        return addr
    chunks = []
    for r0 in range(0, 7, 4):
        for c0 in range(0, 10, 6):
            block = np.zeros(chunk, np.uint16)
            part = data[r0:r0 + 4, c0:c0 + 6]
            block[:part.shape[0], :part.shape[1]] = part
            raw = block.tobytes()
            shuffled = np.frombuffer(raw, np.uint8).reshape(-1, 2).T.tobytes()
            comp = zlib.compress(shuffled)
            chunks.append(((r0, c0), put(comp), len(comp)))
    node = b"TREE" + bytes([1, 0]) + struct.pack("<HQQ", len(chunks), hdf5.UNDEF, hdf5.UNDEF)
    for (r0, c0), addr, size in chunks:
        node += struct.pack("<IIQQQ", size, 0, r0, c0, 0) + struct.pack("<Q", addr)
    node += struct.pack("<IIQQQ", 0, 0, 8, 12, 0)
    btree = put(node)
    space = struct.pack("<BBBB", 2, 2, 0, 1) + struct.pack("<QQ", 7, 10)
    dtype = struct.pack("<BBBBIHH", 0x10, 0, 0, 0, 2, 0, 16)
    layout = struct.pack("<BBB", 3, 2, 3) + struct.pack("<Q", btree) + struct.pack("<III", 4, 6, 2)
    filters = struct.pack("<BB", 2, 2) + struct.pack("<HHHI", 2, 0, 1, 2) + struct.pack("<HHHI", 1, 0, 1, 6)
    names = np.array([b"tip_a", b"tip_b"])
    attr = struct.pack("<BBHHH", 3, 0, 6, 8, 12) + b"\0" + b"names\0" + struct.pack("<BBBBI", 0x13, 1, 0, 0, 5) + \
        struct.pack("<BBBB", 2, 1, 0, 1) + struct.pack("<Q", 2) + names.tobytes()
    dset = put(_v2_header([(0x01, space), (0x03, dtype), (0x08, layout), (0x0B, filters), (0x0C, attr)]))
    link = bytes([1, 0x00, 4]) + b"snps" + struct.pack("<Q", dset)
    root = put(_v2_header([(0x06, link)]))
    buf[0:8] = hdf5.SIGNATURE
    buf[8:12] = bytes([2, 8, 8, 0])
    struct.pack_into("<QQQQ", buf, 12, 0, hdf5.UNDEF, len(buf), root)
    path = tmp_path / "latest.hdf5"
    path.write_bytes(bytes(buf))
    with hdf5.File(path, "r") as f:
        assert f.keys() == ["snps"]
        assert f["snps"].shape == (7, 10) and f["snps"].dtype == np.uint16
        assert (f["snps"][:] == data).all()
        assert f["snps"].attrs["names"].tolist() == [b"tip_a", b"tip_b"]


def test_reads_a_libhdf5_written_file():
    """The one file in this image that libhdf5 itself wrote: scipy's MATLAB v7.3 test file (BSD-licensed test data, copied
    to tests/golden/): 512-byte user block, superblock v0 with base address 512, symbol-table root group (v1 B-tree, local
    heap, SNOD), v1 object header, contiguous little-endian f64 dataset, v1 attribute with a fixed-length string."""
    from pathlib import Path
    path = Path(__file__).parent / "golden" / "libhdf5_matlab73_testdouble.mat"
    with hdf5.File(path, "r") as f:
        assert f.keys() == ["testdouble"]
        d = f["testdouble"]
        assert d.shape == (9, 1) and d.dtype == np.float64
        assert np.array_equal(d[:].ravel(), np.linspace(0, 2 * np.pi, 9))     # scipy's own expectation for this file
        assert bytes(d.attrs["MATLAB_class"]) == b"double"


def test_variable_length_string_names(tmp_path):
    """``names`` stored the way h5py stores a ``list[str]``: variable-length UTF-8 strings in a global heap collection
    (the reference reads them with ``i.decode() if isinstance(i, bytes) else i``, write_database.py:66-69)."""
    names = ["tip_a", "tïp_β", "", "a_rather_longer_tip_name_that_spans_more_than_sixteen_bytes"]
    path = tmp_path / "v.hdf5"
    with hdf5.File(path, "w") as f:
        ds = f.create_dataset("snps", data=np.arange(12, dtype=np.uint8).reshape(4, 3))
        ds.attrs["names"] = np.array(names, dtype=object)
        ds.attrs["version"] = np.int64(3)
        f.create_dataset("snpsmap", data=np.zeros((3, 5), np.uint32)).attrs["columns"] = np.array(["a", "bc"], dtype=object)
    raw = path.read_bytes()
    g = raw.index(b"GCOL")
    assert raw[g + 4] == 1 and struct.unpack_from("<Q", raw, g + 8)[0] == 4096      # one minimum-size collection
    with hdf5.File(path, "r") as f:
        got = f["snps"].attrs["names"]
        assert got.dtype == object and got.tolist() == names
        assert int(np.asarray(f["snps"].attrs["version"]).ravel()[0]) == 3
        assert f["snpsmap"].attrs["columns"].tolist() == ["a", "bc"]
        assert (f["snps"][:] == np.arange(12, dtype=np.uint8).reshape(4, 3)).all()
    snps, snpsmap, got_names = database.read_snps_input(path)            # the input reader the CLI uses
    assert got_names == names and snps.shape == (4, 3) and snpsmap.shape == (3, 5)


def test_unsupported_features_fail_loudly(tmp_path):
    with pytest.raises(ValueError):
        p = tmp_path / "junk.hdf5"
        p.write_bytes(b"not hdf5" * 100)
        hdf5.File(p, "r")
    with pytest.raises(NotImplementedError):
        hdf5._Reader.dtype(struct.pack("<BBBBI", 0x19, 0, 0, 0, 16))     # variable-length
    with pytest.raises(FileNotFoundError):
        hdf5.File(tmp_path / "missing.hdf5", "r")


def test_cli_init_on_hdf5_input(tmp_path):
    from tetrad_b200 import cli
    from tetrad_b200.schema import Project
    d = synth.generate(6, 40, 160, seed=6)
    database.write_snps_input(tmp_path / "in.snps.hdf5", d.snps, d.snpsmap, d.names)
    assert cli.main(["init", str(tmp_path / "in.snps.hdf5"), "-n", "H", "-w", str(tmp_path), "-s", "-r", "1"]) == 0
    proj = Project.load_json(tmp_path / "H.json")
    assert (tmp_path / "H.database.hdf5").read_bytes()[:8] == hdf5.SIGNATURE
    arrays = database.read_arrays(proj.database_file, ["seqarr", "spans", "tmparr", "tmpmap"])
    assert arrays["tmparr"].shape == (6, 160) and arrays["spans"].shape == (40, 2) and proj.samples[0] == "t0"
