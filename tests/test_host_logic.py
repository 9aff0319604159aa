"""
CPU-only checks of the host side: C-ABI library presence / exported symbols,
bit marshalling against the reference's golden vectors, plugin discovery,
configuration round trips and the exception contract that needs no compute.
"""
import ctypes
import json
import os

import numpy as np
import pytest

from smqtk_core.configuration import configuration_test_helper
from smqtk_dataprovider.exceptions import ReadOnlyError
from smqtk_dataprovider.impls.data_element.memory import DataMemoryElement
from smqtk_dataprovider.impls.key_value_store.memory import MemoryKeyValueStore
from smqtk_descriptors.impls.descriptor_element.memory import DescriptorMemoryElement
from smqtk_descriptors.impls.descriptor_set.memory import MemoryDescriptorSet

from smqtk_indexing_b200 import _lib
from smqtk_indexing_b200.interfaces import HashIndex, LshFunctor, NearestNeighborsIndex
from smqtk_indexing_b200.impls.hash_index.linear import LinearHashIndex
from smqtk_indexing_b200.impls.lsh_functor.itq import ItqFunctor
from smqtk_indexing_b200.impls.nn_index.lsh import LSHNearestNeighborIndex
from smqtk_indexing_b200.utils import bits as B


# ------------------------------------------------------------------ C ABI
def test_library_loads_and_exports_every_declared_symbol():
    assert os.path.exists(_lib.LIB_PATH), "run __graft_entry__.build() first"
    raw = ctypes.CDLL(_lib.LIB_PATH)
    declared = _lib.declared_symbols()
    assert len(declared) >= 16
    for name in declared:
        assert hasattr(raw, name), "missing export %s" % name
    assert set(declared) == set(_lib.SIGNATURES), "ctypes table out of sync with include/smqtk_b200.h"
    lib = _lib.load()
    assert lib.sb_version() >= 100
    assert lib.sb_launch_count() == 0          # no compute happened


def test_workspace_query_needs_no_device():
    lib = _lib.load()
    assert lib.sb_hamming_scan_workspace_bytes(10_000_000, 8, 4096, 10) > 0
    assert lib.sb_hamming_scan_workspace_bytes(10, 3, 1, 1) == 0      # unsupported word count


def test_tensor_core_entry_points_shape_rules_need_no_device():
    """`*_supported` / `*_workspace_bytes` of the tcgen05 entry points are host-side arithmetic: the
    shape rules stated in include/smqtk_b200.h, and the dispatcher's work threshold."""
    from smqtk_indexing_b200 import _lib, device
    lib = _lib.load()
    for W in (1, 2, 4, 8):
        assert lib.sb_hamming_scan_tc_supported(10_000_000, W, 4096, 10) == 1
    for U, W, Q, k in [(10_000_000, 16, 4096, 10), (10_000_000, 3, 4096, 10), (10_000_000, 8, 4096, 257),
                       (0, 8, 4096, 10), (10_000_000, 8, 0, 10), (1 << 38, 8, 1, 1)]:
        assert lib.sb_hamming_scan_tc_supported(U, W, Q, k) == 0
        assert lib.sb_hamming_scan_tc_workspace_bytes(U, W, Q, k) == 0
    small = lib.sb_hamming_scan_tc_workspace_bytes(10_000_000, 8, 256, 10)
    big = lib.sb_hamming_scan_tc_workspace_bytes(10_000_000, 8, 4096, 10)
    assert 0 < small < big < (1 << 30)
    # query image (72 KB per 256 queries) + candidate buffers (4096 keys per query) dominate
    assert big >= 16 * 73728 + 4096 * 4096 * 8
    assert lib.sb_hamming_scan_tc_workspace_bytes(10_000_000, 8, 4096, 256) > big
    # dispatcher: large batches over large tables only
    assert device._tc_scan_pays(10_000_000, 8, 4096, 10) and device._tc_scan_pays(1_250_000, 8, 512, 10)
    assert not device._tc_scan_pays(10_000_000, 8, 1, 10) and not device._tc_scan_pays(50_000, 8, 4096, 10)
    assert not device._tc_scan_pays(10_000_000, 16, 4096, 10)
    # training Gram: D % 32, 16-byte rows, b % 4, b <= 256
    assert lib.sb_fit_gram_bits_tc_supported(50_000_000, 256, 256, 256) == 1
    assert lib.sb_fit_gram_bits_tc_supported(1000, 48, 48, 32) == 0          # D % 32
    assert lib.sb_fit_gram_bits_tc_supported(1000, 64, 66, 32) == 0          # row pitch not 16-byte aligned
    assert lib.sb_fit_gram_bits_tc_supported(1000, 64, 64, 30) == 0          # b % 4
    assert lib.sb_fit_gram_bits_tc_supported(1000, 512, 512, 512) == 0       # b > 256
    assert lib.sb_fit_gram_bits_tc_workspace_bytes(50_000_000, 256, 256) >= 256 * 256 * 8


def test_no_silent_cpu_fallback():
    """Without a CUDA device every compute entry point must raise."""
    import torch
    if torch.cuda.is_available():
        pytest.skip("CUDA present")
    f = ItqFunctor(bit_length=1)
    f.mean_vec = np.zeros(2)
    f.rotation = np.ones((2, 1))
    with pytest.raises(RuntimeError, match="CUDA"):
        f.get_hash(np.array([1., 2.]))
    with pytest.raises(RuntimeError, match="CUDA"):
        LinearHashIndex().build_index([[0, 1, 0]])
    with pytest.raises(RuntimeError, match="CUDA"):
        LSHNearestNeighborIndex._get_dist_func("euclidean")(np.zeros(2), np.ones(2))


def test_product_does_not_import_the_oracle():
    import smqtk_indexing_b200
    root = os.path.dirname(smqtk_indexing_b200.__file__)
    for dp, _, files in os.walk(root):
        for fn in files:
            if fn.endswith(".py"):
                src = open(os.path.join(dp, fn)).read()
                assert "np_oracle" not in src and "import oracle" not in src and "from oracle" not in src, fn


# ------------------------------------------------------------------ bit marshalling
def test_bits_against_reference_golden(golden):
    g = golden("bits")
    for b, pv, hx in zip(g["lens"], g["vec_packed"], g["ints_hex"]):
        v = np.unpackbits(pv, count=int(b)).astype(bool)
        i = int(str(hx), 16)
        assert B.bit_vector_to_int_large(v) == i
        assert np.array_equal(B.int_to_bit_vector_large(i, int(b)), v)
        w = B.pack_bits(v, B.words_for_bits(int(b)))
        assert B.words_to_ints(w) == [i]
        assert np.array_equal(B.ints_to_words([i], w.shape[1]), w)
        assert np.array_equal(B.unpack_bits(w, int(b))[0], v)
    for hx, bits, ln, pv in zip(g["back_int_hex"], g["back_bits"], g["back_len"], g["back_vec_packed"]):
        v = B.int_to_bit_vector_large(int(str(hx), 16), int(bits))
        assert len(v) == ln and np.array_equal(v, np.unpackbits(pv, count=int(ln)).astype(bool))


def test_bits_errors_and_widths():
    with pytest.raises(ValueError):
        B.int_to_bit_vector_large(8, 3)
    with pytest.raises(ValueError):
        B.unpack_bits(np.array([[8]], np.uint32), 3)
    with pytest.raises(ValueError):
        B.words_for_bits(0)
    with pytest.raises(ValueError):
        B.words_for_bits(B.MAX_BITS + 1)
    assert [B.words_for_bits(b) for b in (1, 32, 33, 64, 65, 256, 257, 1024)] == [1, 1, 2, 2, 4, 8, 16, 32]
    w = B.pack_bits(np.array([[1, 0, 1]], bool), 1)
    assert w[0, 0] == 5
    assert np.array_equal(B.widen_words(w, 4), [[0, 0, 0, 5]])
    assert np.array_equal(B.unpack_bits(B.widen_words(w, 4), 3), [[True, False, True]])


# ------------------------------------------------------------------ plugins / config
def test_plugins_are_discoverable():
    assert LinearHashIndex in HashIndex.get_impls()
    assert ItqFunctor in LshFunctor.get_impls()
    assert LSHNearestNeighborIndex in NearestNeighborsIndex.get_impls()
    assert LinearHashIndex.is_usable() and ItqFunctor.is_usable() and LSHNearestNeighborIndex.is_usable()


def test_is_usable_needs_library_and_device(monkeypatch):
    """is_usable() = C-ABI library present AND a CUDA device visible."""
    import torch
    monkeypatch.setattr(_lib, "_usable", None)
    monkeypatch.delenv(_lib.ENV_ASSUME_USABLE, raising=False)
    assert LinearHashIndex.is_usable() == torch.cuda.is_available()
    assert ItqFunctor.is_usable() == torch.cuda.is_available()
    assert LSHNearestNeighborIndex.is_usable() == torch.cuda.is_available()
    if not torch.cuda.is_available():
        assert LinearHashIndex not in HashIndex.get_impls()
        with pytest.raises(RuntimeError):                      # Pluggable.__init__ refuses unusable impls
            LinearHashIndex()
    monkeypatch.setenv(_lib.ENV_ASSUME_USABLE, "1")
    assert LinearHashIndex.is_usable()
    monkeypatch.setattr(_lib, "LIB_PATH", _lib.LIB_PATH + ".missing")
    assert not LinearHashIndex.is_usable()                     # no library: never usable


def test_entry_points_discover_plugins_from_a_fresh_interpreter(tmp_path):
    """The package is INSTALLED (pip, offline) into a scratch directory and a fresh interpreter that
    imports nothing but the interfaces must find every implementation through the `smqtk_plugins`
    entry-point group -- the reference's registration mechanism (pyproject.toml:71-82), which is what
    LSHNearestNeighborIndex.from_config relies on (impls/nn_index/lsh.py:88-97)."""
    import subprocess
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    target = str(tmp_path / "site")
    src = str(tmp_path / "src")
    import shutil
    os.makedirs(src)
    shutil.copy(os.path.join(root, "pyproject.toml"), src)
    shutil.copy(os.path.join(root, "README.md"), src)
    shutil.copytree(os.path.join(root, "smqtk_indexing_b200"), os.path.join(src, "smqtk_indexing_b200"),
                    ignore=shutil.ignore_patterns("__pycache__", "obj", "*.o"))
    r = subprocess.run([sys.executable, "-m", "pip", "install", "--no-index", "--no-build-isolation", "--no-deps",
                        "--quiet", "--target", target, src], capture_output=True, text=True)
    assert r.returncode == 0, r.stdout + r.stderr
    code = (
        "import sys, json\n"
        "from smqtk_indexing_b200.interfaces import HashIndex, LshFunctor, NearestNeighborsIndex\n"
        "import smqtk_indexing_b200\n"
        "assert smqtk_indexing_b200.__file__.startswith(%r), smqtk_indexing_b200.__file__\n"
        "pre = [m for m in sys.modules if m.startswith('smqtk_indexing_b200.impls.')]\n"
        "out = {'pre': pre}\n"
        "for i in (HashIndex, LshFunctor, NearestNeighborsIndex):\n"
        "    out[i.__name__] = sorted(c.__module__ + '.' + c.__name__ for c in i.get_impls())\n"
        "print(json.dumps(out))\n" % target)
    env = dict(os.environ, PYTHONPATH=target, SMQTK_B200_ASSUME_USABLE="1")
    r = subprocess.run([sys.executable, "-c", code], capture_output=True, text=True, env=env, cwd=str(tmp_path))
    assert r.returncode == 0, r.stdout + r.stderr
    out = json.loads(r.stdout.strip().splitlines()[-1])
    assert out["pre"] == []                                    # nothing imported the implementations by hand
    p = "smqtk_indexing_b200.impls."
    assert {p + "hash_index.linear.LinearHashIndex",
            p + "hash_index.sklearn_balltree.SkLearnBallTreeHashIndex"} <= set(out["HashIndex"])
    assert {p + "lsh_functor.itq.ItqFunctor", p + "lsh_functor.simple_rp.SimpleRPFunctor"} <= set(out["LshFunctor"])
    assert {p + "nn_index.lsh.LSHNearestNeighborIndex",
            p + "nn_index.flat.FlatL2NearestNeighborsIndex"} <= set(out["NearestNeighborsIndex"])


def test_balltree_named_alias_config_and_reference_cache():
    """`SkLearnBallTreeHashIndex` by name (reference sklearn_balltree.py:112-151): same
    constructor / config keys; reads the reference's .npz cache layout (:179-184)."""
    from io import BytesIO
    from smqtk_indexing_b200.impls.hash_index.sklearn_balltree import SkLearnBallTreeHashIndex
    assert SkLearnBallTreeHashIndex in HashIndex.get_impls()
    c = SkLearnBallTreeHashIndex.get_default_config()
    assert c['leaf_size'] == 40 and c['random_seed'] is None and c['cache_element']['type'] is None
    i = SkLearnBallTreeHashIndex.from_config(c)
    assert i.get_config() == c and i.count() == 0
    i = SkLearnBallTreeHashIndex(DataMemoryElement(), leaf_size=52, random_seed=42)
    for inst in configuration_test_helper(i):
        assert inst.leaf_size == 52 and inst.random_seed == 42 and isinstance(inst.cache_element, DataMemoryElement)
    # a cache written by the reference: np.savez(data_arr=<0/1 matrix>, idx_array_arr, node_data_arr, node_bounds_arr, tail)
    rng = np.random.RandomState(3)
    bits = rng.rand(50, 70) > 0.5
    bits[7] = bits[3]                                          # the tree stores duplicates happily
    buff = BytesIO()
    np.savez(buff, data_arr=bits.astype(np.float64), idx_array_arr=np.arange(50), node_data_arr=np.zeros(3),
             node_bounds_arr=np.zeros((1, 3, 70)), tail=np.array([40, 1, 2, 0, 0, 0, 0, None], dtype=object))
    i = SkLearnBallTreeHashIndex(DataMemoryElement(buff.getvalue()))
    assert i.count() == 49
    assert i.index == {B.bit_vector_to_int_large(v) for v in bits}
    with pytest.raises(ValueError):
        SkLearnBallTreeHashIndex().nn(bits[0], 1)             # empty index (interface check, hash_index.py:108-109)
    ro = DataMemoryElement(readonly=True)
    with pytest.raises(ValueError):
        SkLearnBallTreeHashIndex(ro).save_model()


def test_linear_config():
    c = LinearHashIndex.get_default_config()
    assert len(c) == 1 and c['cache_element']['type'] is None
    i = LinearHashIndex.from_config(c)
    assert i.cache_element is None and i.index == set() and i.count() == 0
    assert i.get_config() == c
    c['cache_element']['type'] = 'smqtk_dataprovider.impls.data_element.memory.DataMemoryElement'
    i = LinearHashIndex.from_config(c)
    assert isinstance(i.cache_element, DataMemoryElement)
    assert i.get_config()['cache_element']['type'] == c['cache_element']['type']
    for inst in configuration_test_helper(LinearHashIndex(DataMemoryElement())):
        assert isinstance(inst.cache_element, DataMemoryElement)


def test_itq_config_roundtrip():
    c = ItqFunctor.get_default_config()
    assert ItqFunctor.from_config(c).get_config() == c
    f = ItqFunctor(DataMemoryElement(), DataMemoryElement(), bit_length=153, itq_iterations=7,
                   normalize=2, random_seed=58)
    for inst in configuration_test_helper(f):
        assert (inst.bit_length, inst.itq_iterations, inst.normalize, inst.random_seed) == (153, 7, 2, 58)
        assert isinstance(inst.mean_vec_cache_elem, DataMemoryElement)
        assert isinstance(inst.rotation_cache_elem, DataMemoryElement)


def test_itq_model_cache_bytes_are_npy(tmp_path):
    from io import BytesIO
    mean, rot = np.array([1., 2., 3.]), np.eye(3)
    mb, rb = BytesIO(), BytesIO()
    np.save(mb, mean)
    np.save(rb, rot)
    f = ItqFunctor(DataMemoryElement(mb.getvalue()), DataMemoryElement(rb.getvalue()), bit_length=3)
    assert f.has_model()
    np.testing.assert_array_equal(f.mean_vec, mean)
    np.testing.assert_array_equal(f.rotation, rot)
    # written caches are byte-identical .npy blobs (reference itq.py:222-237)
    g = ItqFunctor(DataMemoryElement(), DataMemoryElement(), bit_length=3)
    g.mean_vec, g.rotation = mean, rot
    g.save_model()
    assert g.mean_vec_cache_elem.get_bytes() == mb.getvalue()
    assert g.rotation_cache_elem.get_bytes() == rb.getvalue()


def test_itq_errors_without_compute():
    f = ItqFunctor(bit_length=8)
    with pytest.raises(Exception, match="mean vector is none"):
        f.get_hash(np.zeros(4))
    f.mean_vec = np.zeros(4)
    with pytest.raises(Exception, match="rotation matrix is none"):
        f.get_hash(np.zeros(4))
    descr = [DescriptorMemoryElement(i).set_vector([-1. + i, -1. + i]) for i in range(3)]
    g = ItqFunctor(bit_length=8)
    for arg in (descr, iter(descr)):
        with pytest.raises(ValueError, match="fewer features than requested bit encoding"):
            g.fit(arg)
        assert g.mean_vec is None and g.rotation is None
    h = ItqFunctor(bit_length=1)
    h.mean_vec, h.rotation = np.zeros(2), np.ones((2, 1))
    with pytest.raises(RuntimeError, match="already been loaded"):
        h.fit(descr)
    with pytest.raises(ValueError):
        ItqFunctor(normalize="fro")


def test_itq_norm_vector_known_answers():
    f = ItqFunctor(normalize=2)
    np.testing.assert_array_almost_equal(f._norm_vector(np.array([0, 1])), [0, 1])
    np.testing.assert_array_almost_equal(f._norm_vector(np.array([[3., 4.], [0., 0.]])), [[.6, .8], [0, 0]])
    v = np.array([1., 2.])
    assert ItqFunctor()._norm_vector(v) is v


def test_lsh_config():
    i = LSHNearestNeighborIndex(lsh_functor=ItqFunctor(), descriptor_set=MemoryDescriptorSet(),
                                hash2uuids_kvstore=MemoryKeyValueStore(), hash_index=LinearHashIndex(),
                                distance_method='euclidean', read_only=True)
    for inst in configuration_test_helper(i):
        assert isinstance(inst.lsh_functor, ItqFunctor)
        assert isinstance(inst.descriptor_set, MemoryDescriptorSet)
        assert isinstance(inst.hash_index, LinearHashIndex)
        assert isinstance(inst.hash2uuids_kvstore, MemoryKeyValueStore)
        assert inst.distance_method == 'euclidean' and inst.read_only is True

    c = LSHNearestNeighborIndex.get_default_config()
    assert json.loads(json.dumps(c)) == c
    c['lsh_functor']['type'] = 'smqtk_indexing_b200.impls.lsh_functor.itq.ItqFunctor'
    c['descriptor_set']['type'] = 'smqtk_descriptors.impls.descriptor_set.memory.MemoryDescriptorSet'
    c['hash2uuids_kvstore']['type'] = 'smqtk_dataprovider.impls.key_value_store.memory.MemoryKeyValueStore'
    c['hash_index']['type'] = None
    c['hash_index_comment'] = "ignored"
    idx = LSHNearestNeighborIndex.from_config(c)
    assert isinstance(idx.lsh_functor, ItqFunctor) and idx.hash_index is None
    assert json.loads(json.dumps(idx.get_config())) == idx.get_config()


def test_lsh_dist_func_label_and_count():
    with pytest.raises(ValueError, match="Invalid distance method"):
        LSHNearestNeighborIndex._get_dist_func('not-valid-string')
    with pytest.raises(ValueError):
        LSHNearestNeighborIndex(ItqFunctor(), MemoryDescriptorSet(), MemoryKeyValueStore(), distance_method="nope")
    import types
    for m in ("euclidean", "cosine", "hik"):
        assert isinstance(LSHNearestNeighborIndex._get_dist_func(m), types.FunctionType)
    # count() sums the KVS value sets, not the descriptor set (reference lsh.py:271-281)
    kvs = MemoryKeyValueStore()
    idx = LSHNearestNeighborIndex(ItqFunctor(), MemoryDescriptorSet(), kvs)
    assert idx.count() == 0
    kvs.add(0, {0}).add(1, {1, 2}).add(2, frozenset({3, 4, 5}))
    assert idx.count() == 6 and len(idx) == 6


def test_validation_errors_need_no_device():
    idx = LSHNearestNeighborIndex(ItqFunctor(), MemoryDescriptorSet(), MemoryKeyValueStore())
    for call in (idx.build_index, idx.update_index, idx.remove_from_index):
        with pytest.raises(ValueError, match="No DescriptorElement"):
            call([])
    with pytest.raises(ValueError, match="did not have a vector"):
        idx.nn(DescriptorMemoryElement(0))
    with pytest.raises(ValueError, match="No index currently set"):
        idx.nn(DescriptorMemoryElement(0).set_vector([1., 2.]))
    hi = LinearHashIndex()
    for call in (hi.build_index, hi.update_index, hi.remove_from_index):
        with pytest.raises(ValueError, match="No hash vectors"):
            call([])
    with pytest.raises(ValueError, match="No index currently set"):
        hi.nn(np.zeros(3, bool))
    ro = LSHNearestNeighborIndex(ItqFunctor(), MemoryDescriptorSet(), MemoryKeyValueStore(), read_only=True)
    d = [DescriptorMemoryElement(0).set_vector([1., 2.])]
    for call, arg in ((ro.build_index, d), (ro.update_index, d), (ro.remove_from_index, [0])):
        with pytest.raises(ReadOnlyError):
            call(arg)


def test_linear_index_attribute_roundtrip_on_host():
    """`index` get/set works on ints without a device (upload is lazy)."""
    hi = LinearHashIndex()
    hi.index = {1, 2, 3, 4, 2 ** 70 + 5}
    assert hi.count() == 5 and hi.index == {1, 2, 3, 4, 2 ** 70 + 5}
    hi.index = set()
    assert hi.count() == 0


def test_linear_cache_formats():
    from io import BytesIO
    ce = DataMemoryElement()
    hi = LinearHashIndex(ce)
    hi.index = {1, 2, 3, 4}
    hi.save_cache()
    # narrow codes: the reference's 1-D integer .npy (linear.py:131-142)
    assert set(np.load(BytesIO(ce.get_bytes()))) == {1, 2, 3, 4}
    assert LinearHashIndex(ce).index == {1, 2, 3, 4}
    # wide codes: packed table (the reference format cannot represent them)
    ce2 = DataMemoryElement()
    h2 = LinearHashIndex(ce2)
    wide = {2 ** 255 + 7, 2 ** 64, 3}
    h2.index = wide
    h2.save_cache()
    arr = np.load(BytesIO(ce2.get_bytes()))
    assert arr.ndim == 2 and arr.dtype == np.uint32
    assert LinearHashIndex(ce2).index == wide
    # a cache written by the reference itself
    b = BytesIO()
    np.save(b, tuple({5, 9, 12}))
    assert LinearHashIndex(DataMemoryElement(b.getvalue())).index == {5, 9, 12}
    ro = LinearHashIndex(DataMemoryElement(readonly=True))
    ro.index = {1}
    with pytest.raises(ValueError, match="is read-only"):
        ro.save_cache()


def test_code_table_build_matches_oracle():
    """codes.build_table (radix sort / unique / CSR; torch plumbing, runs on CPU tensors too)
    vs the oracle's unique_code_table, including codes with the top bit set (unsigned order)."""
    import torch
    import np_oracle as O
    from smqtk_indexing_b200 import codes as C
    rng = np.random.RandomState(0)
    for W in (1, 2, 4, 8):
        cw = rng.randint(0, 3, size=(4000, W)).astype(np.uint32)
        cw[:, 0] |= rng.randint(0, 2, size=4000).astype(np.uint32) << 31
        t, inv, off, rows = O.unique_code_table(cw)
        tt, ri, co, cr = C.build_table(torch.from_numpy(cw.view(np.int32)))
        assert np.array_equal(tt.numpy().view(np.uint32), t)
        assert np.array_equal(ri.numpy(), inv) and np.array_equal(co.numpy(), off) and np.array_equal(cr.numpy(), rows)
    e = C.build_table(torch.empty((0, 4), dtype=torch.int32))
    assert e[0].shape == (0, 4) and e[2].tolist() == [0]


def test_device_index_ingest_and_snapshot_logic_on_host_tensors(tmp_path):
    """DeviceLshIndex bookkeeping (append into growing buffers, overwrite, tombstones, compaction,
    snapshot round trip) is torch plumbing around codes.build_table: exercised here on CPU tensors
    against the oracle's unique_code_table of the live rows."""
    import torch
    import np_oracle as O
    from smqtk_indexing_b200.engine import DeviceLshIndex
    rng = np.random.RandomState(3)

    def rand_codes(n):
        return torch.from_numpy(rng.randint(0, 40, size=(n, 2)).astype(np.uint32).view(np.int32))

    def check(m, live_rows):
        t, inv, off, rows = O.unique_code_table(m.codes.numpy().view(np.uint32)[live_rows])
        assert np.array_equal(m.table.numpy().view(np.uint32), t)
        assert np.array_equal(m.csr_off.numpy(), off)
        assert np.array_equal(m.csr_rows.numpy(), np.asarray(live_rows)[rows])
        assert np.array_equal(m.row_code.numpy()[live_rows], inv)
        assert m.num_live == len(live_rows)

    m = DeviceLshIndex()
    c0, x0 = rand_codes(500), torch.rand(500, 8)
    m.set_rows(x0, c0)
    check(m, np.arange(500))
    for n in (1, 300, 7):                                   # appends: buffers grow, views stay consistent
        first = m.append_rows(torch.rand(n, 8), rand_codes(n))
        assert first == m.num_rows - n
        m.reindex()
        check(m, np.arange(m.num_rows))
    assert m._x_buf.shape[0] >= m.num_rows and torch.equal(m.x[:500], x0)
    m.overwrite_rows(torch.tensor([3, 4]), torch.ones(2, 8), rand_codes(2))
    m.reindex()
    check(m, np.arange(m.num_rows))
    assert torch.equal(m.x[3], torch.ones(8))
    dead = rng.choice(m.num_rows, 200, replace=False)
    m.remove_rows(torch.from_numpy(dead))
    m.reindex()
    live = np.setdiff1d(np.arange(m.num_rows), dead)
    check(m, live)
    assert (m.row_code.numpy()[dead] == -1).all() and m.num_dead == 200
    # snapshot: same state after a round trip through torch.save (table / CSR re-derived)
    path = str(tmp_path / "index.pt")
    torch.save(m.state_dict(), path)
    m2 = DeviceLshIndex()
    m2.load_state_dict(torch.load(path, weights_only=False), "cpu")
    check(m2, live)
    assert torch.equal(m2.x, m.x) and m2.num_dead == 200
    with pytest.raises(ValueError):
        m2.load_state_dict({"format": "something else"}, "cpu")
    # append after a removal keeps the tombstones; compaction drops them and returns the row map
    m.append_rows(torch.rand(5, 8), rand_codes(5))
    m.reindex()
    live = np.concatenate([live, np.arange(m.num_rows - 5, m.num_rows)])
    check(m, live)
    codes_before = m.codes.numpy().copy()
    remap = m.compact()
    assert m.num_dead == 0 and m.num_rows == len(live)
    assert np.array_equal(remap.numpy()[live], np.arange(len(live))) and (remap.numpy()[dead] == -1).all()
    assert np.array_equal(m.codes.numpy(), codes_before[live])
    check(m, np.arange(len(live)))
    m.remove_rows(torch.arange(m.num_rows))
    m.reindex()
    assert m.table is None and m.num_live == 0


def test_flat_l2_index_config_and_errors_need_no_device():
    from smqtk_descriptors.impls.descriptor_set.memory import MemoryDescriptorSet
    from smqtk_dataprovider.exceptions import ReadOnlyError
    from smqtk_indexing_b200.impls.nn_index.flat import FlatL2NearestNeighborsIndex
    assert FlatL2NearestNeighborsIndex in NearestNeighborsIndex.get_impls()
    c = FlatL2NearestNeighborsIndex.get_default_config()
    assert c["read_only"] is False and "type" in c["descriptor_set"]
    json.dumps(c)
    c["descriptor_set"]["type"] = "smqtk_descriptors.impls.descriptor_set.memory.MemoryDescriptorSet"
    idx = FlatL2NearestNeighborsIndex.from_config(c)
    assert idx.count() == 0
    c2 = idx.get_config()
    json.dumps(c2)
    assert FlatL2NearestNeighborsIndex.from_config(c2).get_config() == c2
    with pytest.raises(ValueError):
        idx.build_index([])
    with pytest.raises(ValueError):                       # empty index
        idx.nn(DescriptorMemoryElement(0).set_vector(np.zeros(4)), 1)
    ro = FlatL2NearestNeighborsIndex(MemoryDescriptorSet(), read_only=True)
    with pytest.raises(ReadOnlyError):
        ro.build_index([DescriptorMemoryElement(0).set_vector(np.zeros(4))])
    with pytest.raises(ReadOnlyError):
        ro.update_index([DescriptorMemoryElement(0).set_vector(np.zeros(4))])
    with pytest.raises(ReadOnlyError):
        ro.remove_from_index([0])


def test_simple_rp_functor_config_and_errors():
    from smqtk_indexing_b200.impls.lsh_functor.simple_rp import SimpleRPFunctor
    assert SimpleRPFunctor in LshFunctor.get_impls()
    f = SimpleRPFunctor(bit_length=16, normalize=2, random_seed=3)
    assert f.get_config() == {"bit_length": 16, "normalize": 2, "random_seed": 3}
    assert SimpleRPFunctor.from_config(f.get_config()).get_config() == f.get_config()
    assert not f.has_model()
    with pytest.raises(RuntimeError):
        f.get_hash(np.zeros(4))


# ------------------------------------------------------------------ FP4 scan: the 8-bit field windows, modelled on the host
def _fp4_window_model(T, tq, d, K=256):
    """Arithmetic of hamming_tc4.cu's epilogue for ONE accumulator column (three queries a, b, c), restated with
    numpy: the accumulator the MMAs produce, `v = bits(acc +rz 1.5 * 2^24)`, and the three window tests.
    tq, d: int arrays [..., 3].  Returns (pass_a, pass_b, pass_c) as the kernel would flag them."""
    beta, gamma = (254 - T) & ~1, T + 1
    t = np.minimum(tq, T)
    u_a = t[..., 0] - d[..., 0] + beta
    w_b = d[..., 1] - t[..., 1] + gamma
    u_c = t[..., 2] - d[..., 2] + beta
    acc = (u_a + 256 * w_b + 65536 * (u_c - 128)).astype(np.int64)      # exact in FP32: |acc| < 2^25 here, see below
    x = acc + 25165824                                                   # acc + 1.5 * 2^24
    # round toward zero to FP32: ulp 2 in [2^24, 2^25), ulp 1 in [2^23, 2^24), ...
    xf = np.where(x >= (1 << 24), (x >> 1) << 1, x).astype(np.float64)
    v = xf.astype(np.float32).view(np.uint32).astype(np.uint64)
    w1, w9 = (v << np.uint64(1)) & np.uint64(0xffffffff), (v << np.uint64(9)) & np.uint64(0xffffffff)
    under = (w1 >> np.uint64(16)) < 0x9700
    a_min, c_min, b_lim = (beta >> 1) << 9, beta << 8, (gamma + 1) << 8
    pa = under | ((w9 & np.uint64(0xffff)) >= a_min)
    pb = under | ((w1 & np.uint64(0xffff)) < b_lim)
    pc = under | ((w9 >> np.uint64(16)) >= c_min)
    return pa, pb, pc


def test_fp4_scan_field_windows_never_lose_a_survivor():
    """Whatever the three queries' distances are -- including distances that push a field out of its window
    (borrow into the b field, carry into the c field, exponent drop of the c field) -- a pair with d <= tq is
    flagged for the re-check; and while every field is inside its window the test is exact."""
    rng = np.random.RandomState(5)
    for T in (0, 1, 7, 64, 100, 127, 128, 200, 251, 252):
        n = 200000
        tq = rng.randint(0, T + 1, size=(n, 3))
        d = rng.randint(0, 257, size=(n, 3))
        # edge values: on / next to the threshold, the extremes of the distance range
        edge = rng.randint(0, 6, size=(n, 3))
        d = np.where(edge == 0, tq, d)
        d = np.where(edge == 1, np.minimum(tq + 1, 256), d)
        d = np.where(edge == 2, rng.choice([0, 255, 256], size=(n, 3)), d)
        tq[:1000] = T
        tq[1000:2000] = 0
        passed = _fp4_window_model(T, tq, d)
        truth = d <= tq
        beta, gamma = (254 - T) & ~1, T + 1
        inside = np.stack([(tq[:, 0] - d[:, 0] + beta >= 0), (d[:, 1] - tq[:, 1] + gamma <= 255),
                           (tq[:, 2] - d[:, 2] + beta >= 0)], axis=1).all(axis=1)
        for h in range(3):
            assert not (truth[:, h] & ~passed[h]).any(), (T, h)          # never a false negative
            assert np.array_equal(passed[h][inside], truth[inside, h]), (T, h)   # exact inside the windows
        assert inside.mean() > 0.2                                       # (the sample covers both regimes)
