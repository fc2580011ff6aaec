"""Host-side logic that needs no GPU: the C-ABI library loads and exports every declared symbol, the
.npz loader/saver matches what the reference reads/writes, the Python mirror keeps the reference's
API contracts, and the product path refuses to run without CUDA (no CPU fallback)."""
import ctypes
import os
import re
import zipfile

import numpy as np
import pytest
import scipy.sparse as sp
import torch

import vsearch_b200 as vs
from tests.util import GOLDEN
from vsearch_b200 import _native, npz_io

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_library_exports_every_declared_symbol():
    hdr = open(os.path.join(ROOT, "include", "vsearch_b200.h")).read()
    body = re.sub(r"/\*.*?\*/", "", hdr, flags=re.S)
    declared = set(re.findall(r"\b(vs_[a-z_0-9]+)\s*\(", body))
    assert declared, "no declarations parsed"
    lib = ctypes.CDLL(_native.LIB_PATH)
    for name in sorted(declared):
        assert hasattr(lib, name), f"{name} declared in include/vsearch_b200.h but not exported"
    assert declared == set(_native.SYMBOLS), declared ^ set(_native.SYMBOLS)
    assert lib.vs_abi_version() == 2


def test_abi_argument_validation_without_gpu():
    """No compute: NULL / bad arguments are rejected before any CUDA call."""
    lib = _native.LIB
    out = ctypes.c_void_p()
    rc = lib.vs_index_create_csr(0, 10, 70000, 0, None, _native.VS_I64, None, _native.VS_I64, None, _native.VS_NONE,
                                 _native.VS_NONE, None, ctypes.byref(out))
    assert rc == _native.VS_ERR_UNSUPPORTED and "uint16" in _native.last_error()  # V must be < 32768
    rc = lib.vs_index_create_csr(0, 10, 100, 0, None, _native.VS_F32, None, _native.VS_I64, None, _native.VS_NONE,
                                 _native.VS_NONE, None, ctypes.byref(out))
    assert rc == _native.VS_ERR_INVALID
    assert lib.vs_index_destroy(None) == 0
    with pytest.raises(ValueError):
        _native.check(rc)


@pytest.mark.parametrize("shift", [0, 100])
def test_loader_matches_reference_loader(shift):
    """sorted-glob order (index10 before index2), column shift, row concatenation (index.py:172-175)."""
    z = np.load(os.path.join(GOLDEN, f"load_shift{shift}.npz"))
    idx = vs.SparseIndex(os.path.join(GOLDEN, "shards", "index*.npz"), None, fp16=False, device="cpu", shift=shift)
    v = idx.vector
    assert tuple(v.shape) == tuple(z["shape"])
    assert np.array_equal(v.crow_indices().numpy(), z["crow"])
    assert np.array_equal(v.col_indices().numpy(), z["col"])
    assert np.array_equal(v.values().numpy(), z["val"])
    assert v.values().dtype == torch.float32
    assert len(idx) == 0 and "SparseIndex" in str(idx) and "torch.sparse_csr" in str(idx)


def test_loader_fp16_default_like_reference():
    idx = vs.SparseIndex(os.path.join(GOLDEN, "shards", "index1.npz"))  # fp16=True is the upstream default
    assert idx.vector.values().dtype == torch.float16


def test_reads_file_saved_by_reference_and_writes_same_layout(tmp_path):
    ref_file = os.path.join(GOLDEN, "saved_by_reference.npz")
    idx = vs.SparseIndex(ref_file, None, fp16=False)
    src = sp.load_npz(os.path.join(GOLDEN, "shards", "index1.npz"))
    assert np.array_equal(idx.vector.values().numpy(), src.data)
    assert np.array_equal(idx.vector.col_indices().numpy(), src.indices)
    out = str(tmp_path / "mine.npz")
    idx.save(out)
    with zipfile.ZipFile(ref_file) as a, zipfile.ZipFile(out) as b:
        assert [i.filename for i in a.infolist()] == [i.filename for i in b.infolist()]
        assert all(i.compress_type == zipfile.ZIP_DEFLATED for i in b.infolist())
    za, zb = np.load(ref_file), np.load(out)
    for k in za.files:
        assert za[k].dtype == zb[k].dtype or k in ("indices", "indptr"), k
        assert np.array_equal(za[k], zb[k]), k
    m = sp.load_npz(out)  # scipy (what the reference's loader calls) reads our file
    assert (m != src).nnz == 0


def test_fp16_npz_roundtrip(tmp_path):
    """fp16 data members (BoT indices built in memory upstream are fp16, retriever.py:232)."""
    p = str(tmp_path / "h.npz")
    npz_io.save_csr_npz(p, np.array([0, 2, 3]), np.array([1, 5, 2]), np.array([1, 1, 1], dtype=np.float16), (2, 9))
    idx = vs.BoTIndex(p, None, fp16=True)
    assert idx.vector.values().dtype == torch.float16 and tuple(idx.vector.shape) == (2, 9)


def test_index_type_contract_and_retriever_dispatch(tmp_path):
    assert [t.value for t in vs.IndexType] == ["dense", "sparse", "bag_of_token"]
    r = vs.Retriever(device="cpu")
    with pytest.raises(ValueError):
        r.load_index("foo.bin")
    with pytest.raises(TypeError):
        r.load_index("foo.npz", index_type=3)
    with pytest.raises(ValueError):
        r.load_index("foo.npz", index_type="nope")
    with pytest.raises(TypeError):
        r.build_index(vectors=torch.zeros(2, 3), index_type=7)
    r.load_index(os.path.join(GOLDEN, "shards", "index0.npz"))
    assert r.index_type == vs.IndexType.SPARSE and isinstance(r.index, vs.SparseIndex)
    r.load_index(os.path.join(GOLDEN, "shards", "index0.npz"), index_type="BAG_OF_TOKEN")
    assert isinstance(r.index, vs.BoTIndex)
    assert torch.equal(r.process_query(np.ones((2, 3), dtype=np.float64)), torch.ones(2, 3))
    t = torch.rand(2, 3)
    assert r.process_query(t) is t
    with pytest.raises(NotImplementedError):
        r.process_query(3.5)


def test_no_cpu_fallback():
    idx = vs.SparseIndex(os.path.join(GOLDEN, "shards", "index0.npz"), None, fp16=False, device="cpu")
    with pytest.raises(RuntimeError, match="no CPU search path"):
        idx.search(torch.zeros(1, 700), 3)
    with pytest.raises(RuntimeError):
        vs.merge_keys(torch.zeros(2, 1, 3, dtype=torch.int64), 3)
    assert "oracle" not in open(os.path.join(ROOT, "vsearch_b200", "index.py")).read().replace("oracle merge", "")


def test_text_store_and_low_memory(tmp_path):
    p = tmp_path / "texts.jsonl"
    texts = ["alpha", "béta ünïcode", "gamma"]
    p.write_text("\n".join(__import__("json").dumps(t) for t in texts) + "\n", encoding="utf-8")
    a = vs.Index(None, str(p))
    b = vs.Index(None, str(p), low_memory=True)
    assert len(a) == 3 and [a.get_sample(i) for i in range(3)] == texts
    assert [b.get_sample(i) for i in range(3)] == texts


def test_row_partition_covers_rows():
    for n, w in [(21015324, 8), (10, 4), (3, 8), (100, 1)]:
        spans = [vs.row_partition(n, w, r) for r in range(w)]
        assert spans[0][0] == 0 and spans[-1][1] == n
        assert all(a[1] == b[0] for a, b in zip(spans[:-1], spans[1:]))
    assert vs.row_partition(21015324, 8, 0) == (0, 2626916)


def test_native_npz_reader_matches_numpy_reader(tmp_path):
    """csrc/npz.cu (zip central directory, .npy headers, streaming inflate with on-the-fly int64 -> int32, row-pointer
    offsets and astype(float16)) against the numpy member reader, on files written by scipy (deflate), by numpy
    (stored) and by our saver (fp16 data), int32 and int64 index members, an empty-row and a one-row shard."""
    rng = np.random.default_rng(0)
    files = []
    for i, (n, idt) in enumerate([(3000, np.int32), (2000, np.int64), (1, np.int32), (2500, np.int32)]):
        m = sp.random(n, 29523, density=0.003, format="csr", random_state=i, dtype=np.float32)
        m.indices, m.indptr = m.indices.astype(idt), m.indptr.astype(idt)
        f = str(tmp_path / f"index{i}.npz")
        sp.save_npz(f, m)
        files.append(f)
    a = npz_io.load_csr_shards(files)
    b = npz_io.load_csr_shards_native(files, threads=3)
    assert a[3] == b[3]
    for x, y in zip(a[:3], b[:3]):
        assert x.dtype == y.dtype and np.array_equal(x, y)
    c = npz_io.load_csr_shards_native(files, fp16=True)
    assert c[2].dtype == np.float16 and np.array_equal(c[2], a[2].astype(np.float16))
    stored = str(tmp_path / "stored.npz")
    np.savez(stored, indices=a[1], indptr=a[0], format=np.array(b"csr"), shape=np.asarray(a[3], dtype=np.int64),
             data=a[2], _is_array=np.array(True))
    d = npz_io.load_csr_shards_native([stored])
    assert np.array_equal(d[0], a[0]) and np.array_equal(d[1], a[1]) and np.array_equal(d[2], a[2])
    half = str(tmp_path / "half.npz")
    npz_io.save_csr_npz(half, a[0], a[1], a[2].astype(np.float16), a[3])
    e = npz_io.load_csr_shards_native([half])
    assert e[2].dtype == np.float16 and np.array_equal(e[1], a[1])
    with pytest.raises(ValueError):
        npz_io.load_csr_shards_native([str(tmp_path / "missing.npz")])
    bad = tmp_path / "bad.npz"
    bad.write_bytes(b"not a zip archive at all" * 10)
    with pytest.raises(ValueError):
        npz_io.load_csr_shards_native([str(bad)])


def test_native_npz_writer_is_scipy_loadable(tmp_path):
    """vs_npz_write (parallel block deflate, csrc/npz.cu): same members, order, dtypes and contents as
    scipy.sparse.save_npz; zipfile's CRC check passes; scipy, numpy and the native reader load it; multi-block members."""
    rng = np.random.default_rng(1)
    n, m = 60_000, 40                                   # data member = 9.6 MB -> three 4 MB deflate blocks
    cols = (np.arange(m) * 29523 // m)[None, :] + rng.integers(0, 29523 // m, (n, m))
    M = sp.csr_array((rng.random(n * m, dtype=np.float32), cols.reshape(-1).astype(np.int32),
                      np.arange(n + 1, dtype=np.int32) * m), shape=(n, 29523))
    ref, mine = str(tmp_path / "ref.npz"), str(tmp_path / "mine.npz")
    sp.save_npz(ref, M)
    npz_io.save_csr_npz_native(mine, M.indptr, M.indices, M.data, M.shape, threads=3)
    with zipfile.ZipFile(ref) as a, zipfile.ZipFile(mine) as b:
        assert [i.filename for i in a.infolist()] == [i.filename for i in b.infolist()]
        assert all(i.compress_type == zipfile.ZIP_DEFLATED for i in b.infolist())
        assert b.testzip() is None
    za, zb = np.load(ref), np.load(mine)
    for k in za.files:
        assert za[k].dtype == zb[k].dtype and np.array_equal(za[k], zb[k]), k
    assert (sp.load_npz(mine) != M).nnz == 0
    back = npz_io.load_csr_shards_native([mine])
    assert np.array_equal(back[1], M.indices) and np.array_equal(back[2], M.data)
    tiny = str(tmp_path / "tiny")                       # suffix appended like numpy; empty matrix; fp16 / int64
    npz_io.save_csr_npz_native(tiny, np.zeros(1, dtype=np.int64), np.zeros(0, dtype=np.int64), np.zeros(0, dtype=np.float16), (0, 7))
    z = np.load(tiny + ".npz")
    assert z["indices"].size == 0 and z["data"].dtype == np.float16 and z["shape"].tolist() == [0, 7]


def test_native_npz_reader_rejects_corrupt_files(tmp_path):
    """truncations, byte flips and overwritten spans of a valid shard: the native reader reports an error (bad zip
    structure, corrupt deflate stream, short member, size / CRC-32 mismatch) -- never a crash, never wrong data."""
    import random

    m = sp.random(1500, 29523, density=0.003, format="csr", random_state=1, dtype=np.float32)
    good = str(tmp_path / "good.npz")
    sp.save_npz(good, m)
    truth = npz_io.load_csr_shards_native([good])
    raw = open(good, "rb").read()
    rnd = random.Random(0)
    bad = str(tmp_path / "bad.npz")
    for t in range(90):
        b = bytearray(raw)
        if t % 3 == 0:
            b = b[:rnd.randrange(1, len(b))]
        elif t % 3 == 1:
            for _ in range(rnd.randrange(1, 20)):
                b[rnd.randrange(len(b))] = rnd.randrange(256)
        else:
            i = rnd.randrange(len(b) - 64)
            b[i:i + 32] = bytes(rnd.randrange(256) for _ in range(32))
        if bytes(b) == raw:
            continue
        with open(bad, "wb") as f:
            f.write(bytes(b))
        try:
            got = npz_io.load_csr_shards_native([bad])
        except Exception:  # noqa: BLE001  (ValueError from the ABI, or numpy/zipfile on the tiny members)
            continue
        assert all(np.array_equal(x, y) for x, y in zip(got[:3], truth[:3])), f"case {t}: corrupt file loaded as different data"


def test_native_npz_writer_zip64_records(tmp_path, monkeypatch):
    """A 21M-row index has members beyond 4 GB, i.e. ZIP64 records in the local headers, the central directory and the
    end-of-directory locator.  VSEARCH_B200_NPZ_FORCE_ZIP64 takes that path with a small matrix: zipfile (CRC check),
    scipy and the native reader must all read the file."""
    monkeypatch.setenv("VSEARCH_B200_NPZ_FORCE_ZIP64", "1")
    m = sp.random(3000, 29523, density=0.003, format="csr", random_state=1, dtype=np.float32)
    f = str(tmp_path / "z64.npz")
    npz_io.save_csr_npz_native(f, m.indptr, m.indices, m.data, m.shape, threads=2)
    with zipfile.ZipFile(f) as z:
        assert all(i.extract_version == 45 for i in z.infolist()) and z.testzip() is None
    assert (sp.load_npz(f) != m).nnz == 0
    r = npz_io.load_csr_shards_native([f])
    assert np.array_equal(r[0], m.indptr) and np.array_equal(r[1], m.indices) and np.array_equal(r[2], m.data)


def test_native_npz_roundtrip_property(tmp_path):
    """hypothesis: random CSR shapes / dtypes / thread counts through vs_npz_write then vs_npz_read, with every
    supported conversion (index narrowing and widening, float16 <-> float32) and partial reads."""
    from hypothesis import given, settings
    from hypothesis import strategies as st_

    counter = [0]

    @settings(max_examples=25, deadline=None)
    @given(n=st_.integers(0, 300), width=st_.integers(1, 40), idt=st_.sampled_from([np.int32, np.int64]),
           fdt=st_.sampled_from([np.float32, np.float16]), threads=st_.integers(1, 4), seed=st_.integers(0, 10**6))
    def run(n, width, idt, fdt, threads, seed):
        rng = np.random.default_rng(seed)
        lens = rng.integers(0, width + 1, n)
        indptr = np.concatenate(([0], np.cumsum(lens))).astype(idt)
        nnz = int(indptr[-1])
        indices = rng.integers(0, 29523, nnz).astype(idt)
        data = (rng.integers(-64, 64, nnz) / 8).astype(fdt)          # exact in fp16
        counter[0] += 1
        f = str(tmp_path / f"p{counter[0]}.npz")
        npz_io.save_csr_npz_native(f, indptr, indices, data, (n, 29523), threads=threads)
        z = np.load(f)
        assert np.array_equal(z["indptr"], indptr) and np.array_equal(z["indices"], indices) and np.array_equal(z["data"], data)
        sh = npz_io._NativeShard(f)
        try:
            for dst in (np.int32, np.int64):
                out = np.empty(nnz, dtype=dst)
                sh.read_into("indices", out)
                assert np.array_equal(out, indices.astype(dst))
            for dst in (np.float32, np.float16):
                out = np.empty(nnz, dtype=dst)
                sh.read_into("data", out)
                assert np.array_equal(out, data.astype(dst))
            out = np.empty(n, dtype=np.int64)                           # row pointers: skip the leading 0, add an offset
            sh.read_into("indptr", out, skip=1, add=1000)
            assert np.array_equal(out, indptr[1:].astype(np.int64) + 1000)
        finally:
            sh.close()

    run()
